// oracle/restate/csc_restate.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A plain, single-threaded CPU restatement of libcsc's compression path for levels 1..4 and of its
// decoder, written from the reference's published behaviour.  It is deliberately the "textbook"
// formulation (real hash tables, a ring window, byte-at-a-time coder buffers) and shares no code
// with csc_b200/csrc -- the product computes the same streams with sort-based chains, liveness
// bitmaps and a linear history, so agreement between the two is meaningful.
//
// Pinning: tests/test_oracle_cpu.py checks this restatement byte-for-byte against the unmodified
// reference (oracle/_ref, built from /root/reference) and against tests/golden/golden.json.
// Level 5 (binary-tree match finder, csc_mf.cpp:369-451) is not restated: its parity is ratio-only.
//
// Each function cites the reference lines it follows (paths relative to /root/reference/src/libcsc).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

const u32 KB = 1024, MB = 1048576, BLK = 8192;
enum { T_NORMAL = 1, T_TXT = 2, T_EXE = 3, T_FAST = 4, T_NOLZ = 5, T_ENTROPY = 7, T_BAD = 8, T_EOF = 9, T_DLT = 0x10, T_SKIP = 0x1E };
const u32 kChannels[5] = {1, 2, 3, 4, 8};

struct Props {   // csc_common.h:19-63 / csc_enc.cpp:16-97
    u32 dict, blocksize, rawblock, hbits, hwidth, good, mode, fdlt, ftxt, fexe;
};
void props_for(Props &p, u32 dict, int level) {
    dict += 10 * KB;
    if (dict < 32 * KB) dict = 32 * KB;
    if (dict > 1024 * MB) dict = 1024 * MB;
    if (level < 1) level = 1;
    if (level > 5) level = 5;
    p.dict = dict; p.blocksize = 64 * KB; p.rawblock = 2 * MB; p.fdlt = p.ftxt = p.fexe = 1;
    u32 hb;
    if (dict < MB) hb = 19; else if (dict <= 4 * MB) hb = 20; else if (dict <= 16 * MB) hb = 21;
    else if (dict <= 64 * MB) hb = 22; else if (dict <= 256 * MB) hb = 23; else hb = 24;
    while ((1u << hb) > dict) hb--;
    switch (level) {
        case 1: p.hwidth = 1; p.mode = 2; p.good = 32; p.hbits = hb + 1; break;
        case 2: p.hwidth = 8; p.mode = 2; p.good = 24; p.hbits = hb - 1; break;
        case 3: p.hwidth = 2; p.mode = 3; p.good = 16; p.hbits = hb + 1; break;
        case 4: p.hwidth = 8; p.mode = 3; p.good = 24; p.hbits = hb - 1; break;
        default: p.hwidth = 0; p.mode = 3; p.good = 48; p.hbits = hb; break;
    }
}

// ------------------------------------------------------------------------------------------
// stream framing + coder (csc_memio.cpp:83-108, csc_coder.{h,cpp})
struct Coder {
    std::vector<u8> *out;
    u32 bsize;
    std::vector<u8> rcbuf, bcbuf;   // never cleared: the byte skipped by flush keeps its old value (csc_coder.cpp:46-47)
    u32 rcn, bcn;
    u64 low, pending;
    u32 range, nbits, acc;
    u8 cache;

    void init(std::vector<u8> *o, u32 bs) {
        out = o; bsize = bs; rcbuf.assign(bs, 0); bcbuf.assign(bs, 0);
        reset();
    }
    void reset() { low = 0; range = 0xFFFFFFFFu; pending = 1; cache = 0; rcn = bcn = 0; nbits = acc = 0; }
    void block(const u8 *b, u32 n, int rc) {   // MemIO::WriteBlock
        u8 h = (u8)(rc << 7);
        if (n == bsize) out->push_back(h | 0x40);
        else { out->push_back(h); out->push_back((n >> 16) & 255); out->push_back((n >> 8) & 255); out->push_back(n & 255); }
        out->insert(out->end(), b, b + n);
    }
    void rc_byte(u8 b) {
        rcbuf[rcn++] = b;
        if (rcn == bsize) { block(rcbuf.data(), bsize, 1); rcn = 0; }
    }
    void bc_byte(u8 b) {
        bcbuf[bcn++] = b;
        if (bcn == bsize) { block(bcbuf.data(), bsize, 0); bcn = 0; }
    }
    void shift() {   // Coder::RC_ShiftLow, csc_coder.cpp:89-112
        if ((u32)low < 0xFF000000u || (low >> 32) != 0) {
            u8 t = cache;
            do { rc_byte((u8)(t + (u8)(low >> 32))); t = 0xFF; } while (--pending);
            cache = (u8)((u32)low >> 24);
        }
        pending++;
        low = (u32)low << 8;
    }
    void bit(u32 &p, u32 v) {   // EncodeBit, csc_coder.h:67-81
        u32 bound = (range >> 12) * p;
        if (v) { range = bound; p += (0xFFF - p) >> 5; }
        else { low += bound; range -= bound; p -= p >> 5; }
        if (range < (1u << 24)) { range <<= 8; shift(); }
    }
    void raw16(u32 v, u32 n) {   // Coder::EncDirect16, csc_coder.cpp:76-87
        acc = (acc << n) | v;
        nbits += n;
        while (nbits >= 8) { bc_byte((acc >> (nbits - 8)) & 255); nbits -= 8; }
    }
    void raw(u32 v, u32 n) { if (n <= 16) raw16(v, n); else { raw16(v >> 16, n - 16); raw16(v & 0xFFFF, 16); } }
    void number(u32 v) {   // Model::EncodeInt, csc_model.cpp:389-414
        u32 slot = 0;
        for (u32 t = v; t; t >>= 1) slot++;
        if (slot) slot--;
        raw(slot, 5);
        if (!slot) raw(v, 1); else raw(v - (1u << slot), slot);
    }
    void flush() {   // Coder::Flush, csc_coder.cpp:40-74
        for (int i = 0; i < 5; i++) shift();
        rcn++;   // skipped, not stored
        bc_byte((acc << (8 - nbits)) & 255);
        bc_byte(0);
        block(rcbuf.data(), rcn, 1);
        block(bcbuf.data(), bcn, 0);
        reset();
    }
};

// ------------------------------------------------------------------------------------------
// model (csc_model.{h,cpp})
struct Model {
    Coder *c;
    std::vector<u32> lit, dlt;
    u32 st[192], rp[192], dist[168], slot[2], e1[8], e2[8], e3[128], dx[464], longlen, rleflag;
    u32 state, ctx, price_tab[512], lenprice[32], lp_count;
    void init(Coder *cc) {
        c = cc;
        lit.assign(65536, 2048); dlt.clear();
        for (u32 &v : st) v = 2048; for (u32 &v : rp) v = 2048; for (u32 &v : dist) v = 2048;
        for (u32 &v : e1) v = 2048; for (u32 &v : e2) v = 2048; for (u32 &v : e3) v = 2048; for (u32 &v : dx) v = 2048;
        slot[0] = slot[1] = longlen = rleflag = 2048;
        state = ctx = lp_count = 0;
        for (int i = 0; i < 512; i++) price_tab[i] = (u32)(128 * log((float)(i * 8 + 4) / 4096) / log(0.5));   // csc_model.cpp:68-70
    }
    void tree(u32 *p, u32 sym, u32 nbits) {   // MSB-first bit tree of nbits bits
        u32 node = 1;
        for (u32 i = nbits; i--;) { u32 b = (sym >> i) & 1; c->bit(p[node], b); node = node * 2 + b; }
    }
    void literal_body(u32 *row, u32 ch) { tree(row, ch, 8); }
    void literal(u32 ch) {   // EncodeLiteral, csc_model.cpp:169-183
        c->bit(st[state * 3], 0);
        state = (state * 4) & 63;
        u32 *row = &lit[ctx * 256];
        ctx = ch;
        literal_body(row, ch);
    }
    void len1(u32 l) {   // encode_matchlen_1, csc_model.cpp:113-145
        if (l < 8) { c->bit(slot[0], 0); tree(e1, l, 3); }
        else if (l < 16) { c->bit(slot[0], 1); c->bit(slot[1], 0); tree(e2, l - 8, 3); }
        else { c->bit(slot[0], 1); c->bit(slot[1], 1); tree(e3, l - 16, 7); }
    }
    void len2(u32 l) {   // encode_matchlen_2, csc_model.cpp:147-159
        if (l >= 143) {
            len1(143);
            for (l -= 143; l >= 143; l -= 143) c->bit(longlen, 0);
            c->bit(longlen, 1);
        }
        len1(l);
    }
    static u32 slot_of(u32 d) {   // largest s with dist_table_[s] <= d (csc_model.cpp:45-55)
        static const u32 tab[32] = {0, 1, 2, 3, 5, 9, 17, 33, 65, 129, 257, 513, 1025, 2049, 4097, 8193, 16385, 32769, 65537, 131073,
                                    262145, 524289, 1048577, 2097153, 4194305, 8388609, 16777217, 33554433, 67108865,
                                    134217729, 268435457, 536870913};
        u32 s = 0;
        while (s + 1 < 32 && tab[s + 1] <= d) s++;
        return s;
    }
    void match(u32 d, u32 l) {   // EncodeMatch, csc_model.cpp:301-366
        c->bit(st[state * 3], 1);
        c->bit(st[state * 3 + 1], 1);
        len2(l);
        u32 base, sb;
        if (l == 0) { base = 0; sb = 3; } else if (l <= 2) { base = 8 + 16 * (l - 1); sb = 4; }
        else if (l <= 5) { base = 40 + 32 * (l - 3); sb = 5; } else { base = 136; sb = 5; }
        u32 s = slot_of(d);
        tree(dist + base, s, sb);
        if (s > 2) {
            u32 eb = s - 2, extra = d - (1u << eb) - 1;
            if (eb > 4) c->raw(extra >> 4, eb - 4);
            u32 low4 = extra & 15, rev = ((low4 & 1) << 3) | ((low4 & 2) << 1) | ((low4 & 4) >> 1) | (low4 >> 3);
            tree(dx + (eb - 1) * 16, rev, 4);
        }
        state = (state * 4 + 1) & 63;
    }
    void rep0len1() {   // EncodeRep0Len1, csc_model.cpp:198-207
        c->bit(st[state * 3], 1); c->bit(st[state * 3 + 1], 0); c->bit(st[state * 3 + 2], 0);
        ctx = 0;
        state = (state * 4 + 2) & 63;
    }
    void rep(u32 idx, u32 l) {   // EncodeRepDistMatch, csc_model.cpp:218-232
        c->bit(st[state * 3], 1); c->bit(st[state * 3 + 1], 0); c->bit(st[state * 3 + 2], 1);
        u32 hi = idx >> 1, lo = idx & 1;
        c->bit(rp[state * 3], hi);
        c->bit(rp[state * 3 + 1 + hi], lo);
        len2(l);
        state = (state * 4 + 3) & 63;
    }
    // prices (csc_model.cpp:161-167,185-196,209-216,273-299,368-387)
    u32 pb(u32 p, u32 v) const { return price_tab[(v ? p : 4096 - p) >> 3]; }
    u32 price_tree(const u32 *p, u32 sym, u32 nbits) const {
        u32 node = 1, r = 0;
        for (u32 i = nbits; i--;) { u32 b = (sym >> i) & 1; r += pb(p[node], b); node = node * 2 + b; }
        return r;
    }
    u32 price_literal(u32 s, u32 cx, u32 ch) const { return pb(st[s * 3], 0) + price_tree(&lit[cx * 256], ch, 8); }
    u32 price_rep0len1(u32 s) const { return pb(st[s * 3], 1) + pb(st[s * 3 + 1], 0) + pb(st[s * 3 + 2], 0); }
    u32 price_rep(u32 s, u32 idx) const {
        u32 hi = idx >> 1, lo = idx & 1;
        return pb(st[s * 3], 1) + pb(st[s * 3 + 1], 0) + pb(st[s * 3 + 2], 1) + pb(rp[s * 3], hi) + pb(rp[s * 3 + 1 + hi], lo);
    }
    u32 price_dist(u32 s, u32 d) const {
        u32 sl = slot_of(d);
        return pb(st[s * 3], 1) + pb(st[s * 3 + 1], 1) + (sl > 2 ? sl + 2 : 2) * 128;
    }
    u32 price_len(u32 l) {
        if (l >= 32) return 768;
        if (lp_count-- == 0) {   // len_price_rebuild every 4096 lookups, csc_model.cpp:234-270
            for (u32 i = 0; i < 32; i++) {
                if (i < 8) lenprice[i] = pb(slot[0], 0) + price_tree(e1, i, 3);
                else if (i < 16) lenprice[i] = pb(slot[0], 1) + pb(slot[1], 0) + price_tree(e2, i - 8, 3);
                else lenprice[i] = pb(slot[0], 1) + pb(slot[1], 1) + price_tree(e3, i - 16, 7);
            }
            lp_count = 4096;
        }
        return lenprice[l];
    }
    // non-LZ payloads (csc_model.cpp:448-513)
    void raw_bytes(const u8 *s, u32 n) { c->number(n); for (u32 i = 0; i < n; i++) c->raw16(s[i], 8); }
    void literals(const u8 *s, u32 n) {
        c->number(n);
        for (u32 i = 0; i < n; i++) { u32 *row = &lit[ctx * 256]; ctx = s[i]; literal_body(row, s[i]); }
    }
    void rle(const u8 *s, u32 n) {
        c->number(n);
        if (dlt.empty()) dlt.assign(65536, 2048);
        u32 cx = 0;
        for (u32 i = 0; i < n;) {
            if (i > 0 && n - i > 3 && s[i - 1] == s[i] && s[i] == s[i + 1] && s[i] == s[i + 2]) {
                u32 j = i + 3;
                while (j < n && s[j] == s[j - 1]) j++;
                if (j - i > 10) { cx = s[j - 1]; c->bit(rleflag, 1); len2(j - i - 11); i = j; continue; }
            }
            c->bit(rleflag, 0);
            literal_body(&dlt[cx * 256], s[i]);
            cx = s[i++];
        }
    }
};

// ------------------------------------------------------------------------------------------
// analyzer (csc_analyzer.cpp) and filters (csc_filters.cpp)
struct Analyzer {
    u32 logt[BLK / 16 + 1];
    Analyzer() {
        for (u32 i = 0; i < BLK / 16; i++) logt[i] = (double)100 * log((double)i * 16 + 8) / log((double)2);
        logt[BLK / 16] = (double)100 * log((double)BLK) / log((double)2);
    }
    u32 cost(const u32 *f, u32 n) const { u32 e = n * logt[n >> 4]; for (int i = 0; i < 256; i++) e -= f[i] * logt[f[i] >> 4]; return e; }
    u32 delta_bpb(const u8 *s, u32 n, u32 ch) const {   // GetDltBpb :166-182
        u32 f[256] = {0};
        u8 prev = 0;
        for (u32 c = 0; c < ch; c++) for (u32 j = c; j < n; j += ch) { f[(u8)(s[j] - prev)]++; prev = s[j]; }
        return cost(f, n) / n;
    }
    u32 classify(const u8 *s, u32 n, u32 *bpb) const {   // Analyze :184-239
        if (n > BLK) n = BLK;
        if (n < 512) return T_SKIP;
        u32 f[256] = {0};
        for (u32 i = 0; i < n; i++) f[s[i]]++;
        u32 e = cost(f, n), distinct = 0, high = 0, letters = 0;
        for (int i = 0; i < 256; i++) { distinct += f[i] != 0; if (i >= 128) high += f[i]; if (i >= 'a' && i <= 'z') letters += f[i]; }
        *bpb = e / n;
        if (high < n / 8 && f[' '] + f['\n'] + f[':'] + f['.'] + f['/'] > n / 16 && f['a'] + f['e'] + f['t'] > n / 16 && e > 300 * n && letters > n / 3) return T_TXT;
        if (f[0x8b] > (n >> 8) && f[0] > (n >> 8) * 2 && f[0xE8] > 6) return T_EXE;
        if (e > (log((double)distinct - 2) / log((double)2) - 0.6) * 100.0 * n && distinct < 16 && distinct >= 6) return T_ENTROPY;
        if (e < 400 * n && distinct < 200) return T_NORMAL;
        // get_channel_idx :122-164
        static const u32 lag[5] = {1, 2, 3, 4, 8};
        u32 same[5] = {0}, dif[5] = {0};
        for (u32 i = 0; i + 16 < n; i++)
            for (int k = 0; k < 5; k++) { int a = s[i], b = s[i + lag[k]]; same[k] += a == b; dif[k] += abs(a - b); }
        u32 best = 0, mins = same[0], maxd = dif[0];
        for (int k = 0; k < 5; k++) { if (same[k] < mins) mins = same[k]; if (dif[k] > maxd) maxd = dif[k]; if (dif[k] < dif[best]) best = k; }
        if ((maxd > dif[best] * 4 || maxd > dif[best] + 40 * n) && same[best] > mins * 3 && same[0] < 0.3 * n) return T_DLT + best;
        if (e > 795 * n) return T_BAD;
        if (e > 780 * n) return T_FAST;
        return T_NORMAL;
    }
};

// e8e9 (csc_filters.cpp:508-609): 8-byte delay line; when the byte 5 back is E8/E9 and >= 4 bytes after
// the previous hit, a rel32 within +-16 MiB becomes a 25-bit absolute address with swizzled bytes.
void e8e9(u8 *s, u32 n, bool inverse) {
    u32 next_ok = 5;   // position (bytes consumed) from which a new opcode may be recognised
    for (u32 i = 5; i < n; i++) {   // i = bytes consumed before the check; opcode sits at i-5, operand at i-4..i-1
        if (i < next_ok || (s[i - 5] & 0xFE) != 0xE8) continue;
        next_ok = i + 4;
        u32 v = (u32)s[i - 4] | ((u32)s[i - 3] << 8) | ((u32)s[i - 2] << 16) | ((u32)s[i - 1] << 24);
        u32 x = v - 0xFF000000u;
        if (x >= 0x02000000u) continue;
        if (!inverse) {
            x = ((x + i) & 0x01FFFFFFu) << 7;
            x = (x >> 24) | (((x >> 16) & 255) << 8) | (((x >> 8) & 255) << 16) | ((x & 255) << 17);
        } else {
            x = (((x >> 24) & 255) << 7) | (((x >> 16) & 255) << 8) | (((x >> 8) & 255) << 16) | (x << 24);
            x = ((x >> 7) - i) & 0x01FFFFFFu;
        }
        v = x + 0xFF000000u;
        s[i - 4] = (u8)v; s[i - 3] = (u8)(v >> 8); s[i - 2] = (u8)(v >> 16); s[i - 1] = (u8)(v >> 24);
    }
}

struct Dict {   // csc_filters.cpp:8-111
    u16 next[300][26];
    u8 sym[300];
    const char *word[256];
    u32 limit;
    Dict() {
        static const char *w[] = {"ac", "ad", "ai", "al", "am", "an", "ar", "as", "at", "ea", "ec", "ed", "ee", "el", "en", "er", "es",
            "et", "id", "ie", "ig", "il", "in", "io", "is", "it", "of", "ol", "on", "oo", "or", "os", "ou", "ow", "ul", "un", "ur", "us",
            "ba", "be", "ca", "ce", "co", "ch", "de", "di", "ge", "gh", "ha", "he", "hi", "ho", "ra", "re", "ri", "ro", "rs", "la", "le",
            "li", "lo", "ld", "ll", "ly", "se", "si", "so", "sh", "ss", "st", "ma", "me", "mi", "ne", "nc", "nd", "ng", "nt", "pa", "pe",
            "ta", "te", "ti", "to", "th", "tr", "wa", "ve", "all", "and", "but", "dow", "for", "had", "hav", "her", "him", "his", "man",
            "mor", "not", "now", "one", "out", "she", "the", "was", "wer", "whi", "whe", "wit", "you", "any", "are", "that", "said",
            "with", "have", "this", "from", "were", "tion"};
        memset(next, 0, sizeof(next)); memset(sym, 0, sizeof(sym)); memset(word, 0, sizeof(word));
        u32 nodes = 1, code = 0x82;
        for (const char *s : w) {
            u32 at = 0;
            for (const char *q = s; *q; q++) { u16 &nx = next[at][*q - 'a']; if (!nx) nx = (u16)nodes++; at = nx; }
            sym[at] = (u8)code; word[code] = s; code++;
        }
        limit = code;
    }
    bool forward(u8 *s, u32 n) const {   // Foward_Dict :256-335
        if (n < 16384) return false;
        std::vector<u8> o;
        o.reserve(n + 16);
        u32 i = 0;
        while (i < n - 5) {
            if (o.size() > n - 16) return false;
            u8 ch = s[i];
            if (ch >= 'a' && ch <= 'z') {
                u32 at = 0, hit = 0, hitlen = 0;
                for (u32 j = 0;; j++) {
                    u32 k = (u32)s[i + j] - 'a';
                    if (k > 25 || !next[at][k]) break;
                    at = next[at][k];
                    if (sym[at]) { hit = sym[at]; hitlen = j + 1; }
                }
                if (hit) { o.push_back((u8)hit); i += hitlen; continue; }
            } else if (ch >= 0x82) o.push_back(254);
            o.push_back(ch);
            i++;
        }
        for (; i < n; i++) { if (s[i] >= 0x82) o.push_back(254); o.push_back(s[i]); }
        if (o.size() > n * 0.82) return false;
        memcpy(s, o.data(), o.size());
        memset(s + o.size(), 0x20, n - o.size());
        return true;
    }
    void inverse(u8 *s, u32 n) const {   // Inverse_Dict :337-369
        std::vector<u8> o(n);
        u32 i = 0, k = 0;
        while (k < n && i < n) {
            u8 ch = s[i];
            if (ch >= 0x82 && ch < limit) { for (const char *q = word[ch]; *q && k < n; q++) o[k++] = (u8)*q; }
            else if (ch == 254 && i + 1 < n && s[i + 1] >= 0x82) o[k++] = s[++i];
            else o[k++] = ch;
            i++;
        }
        memcpy(s, o.data(), n);
    }
};
void delta_fwd(u8 *s, u32 n, u32 ch) {   // Forward_Delta :132-164
    if (n < 512) return;
    std::vector<u8> t(s, s + n);
    u32 k = 0; u8 prev = 0;
    for (u32 c = 0; c < ch; c++) for (u32 j = c; j < n; j += ch) { s[k++] = (u8)(t[j] - prev); prev = t[j]; }
}
void delta_inv(u8 *s, u32 n, u32 ch) {   // Inverse_Delta :371-398
    if (n < 512) return;
    std::vector<u8> t(s, s + n);
    u32 k = 0; u8 prev = 0;
    for (u32 c = 0; c < ch; c++) for (u32 j = c; j < n; j += ch) { prev = (u8)(t[k++] + prev); s[j] = prev; }
}

// ------------------------------------------------------------------------------------------
// match finder (csc_mf.cpp) over a ring window, levels 1..4 (no binary tree)
struct Cand { u32 len, dist; };
struct MF {
    u8 *w; u32 wsize, valid, pos, bits, width, good;
    std::vector<u32> t2, t3, t6;
    void init(u8 *wnd, u32 ws, u32 hb, u32 hw, u32 gl) {
        w = wnd; wsize = ws; valid = ws - BLK - 4; pos = valid; bits = hb; width = hw; good = gl;
        t2.assign(16384, 0); t3.assign(65536, 0); t6.assign((size_t)width << bits, 0);
    }
    static u32 h2(const u8 *p) { u16 v; memcpy(&v, p, 2); return (v * 65521u) & 0x3FFF; }
    static u32 h3(const u8 *p) { return (p[0] << 8) ^ (p[1] << 5) ^ p[2]; }
    u32 h6(const u8 *p) const { u32 a; u16 b; memcpy(&a, p, 4); memcpy(&b, p + 4, 2); return ((a ^ ((u32)b << 13)) * 2654435761u) >> (32 - bits); }
    static u32 bound(u32 l) { static const u32 b[7] = {0, 0, 64, 1024, 16 * KB, 256 * KB, 4 * MB}; return b[l]; }
    u32 common(u32 a, u32 b, u32 lim) const { u32 l = 0; while (l < lim && w[a + l] == w[b + l]) l++; return l; }

    // find_match :243-495 (hash tables only)
    u32 find(Cand *out, const u32 *reps, u32 wp, u32 lim) {
        u32 n = 0, minlen = 1, dist = 0;
        for (u32 i = 0; i < 4; i++) {
            if (reps[i] >= valid) continue;
            u32 cp = wp >= reps[i] ? wp - reps[i] : wp + wsize - reps[i];
            u32 cl = lim < wsize - cp ? lim : wsize - cp;
            if (minlen >= cl || w[cp + minlen] != w[wp + minlen]) continue;
            u32 ml = common(wp, cp, cl);
            if (ml && i == 0) out[n++] = Cand{1, 1};
            if (ml > minlen) { minlen = ml; out[n++] = Cand{ml, 1 + i}; if (ml >= good) { dist = 0xFFFFFFFFu; break; } }
        }
        u32 k2 = h2(w + wp), k3 = h3(w + wp);
        for (int tab = 0; tab < 2; tab++) {
            u32 e = tab ? t3[k3] : t2[k2];
            if (pos - e <= dist) continue;
            dist = pos - e;
            if (dist >= valid) continue;
            bool inside = tab ? wp >= dist : wp > dist;   // the 2-byte table uses '>' (:306)
            u32 cp = inside ? wp - dist : wp + wsize - dist;
            u32 cl = lim < wsize - cp ? lim : wsize - cp;
            if (minlen >= cl || w[cp + minlen] != w[wp + minlen]) continue;
            u32 ml = common(wp, cp, cl);
            if (ml <= minlen) continue;
            minlen = ml;
            if (ml <= 6 && dist >= bound(ml)) continue;
            out[n++] = Cand{ml, 4 + dist};
            if (ml >= good) dist = 0xFFFFFFFFu;
        }
        t2[k2] = pos; t3[k3] = pos;
        if (width) {
            u32 *b = &t6[(size_t)h6(w + wp) * width];
            for (u32 i = 0; i < width; i++) {
                if (pos - b[i] <= dist) continue;
                dist = pos - b[i];
                if (dist >= valid) continue;
                u32 cp = wp >= dist ? wp - dist : wp + wsize - dist;
                u32 cl = lim < wsize - cp ? lim : wsize - cp;
                if (minlen >= cl || w[cp + minlen] != w[wp + minlen]) continue;
                u32 ml = common(wp, cp, cl);
                if (ml <= minlen) continue;
                minlen = ml;
                if (ml <= 6 && dist >= bound(ml)) continue;
                out[n++] = Cand{ml, 4 + dist};
                if (ml >= good) { dist = 0xFFFFFFFFu; break; }
            }
            for (u32 i = width - 1; i > 0; i--) b[i] = b[i - 1];
            b[0] = pos;
        }
        pos++;
        return n;
    }
    void slide(u32 wp, u32 len) {   // SlidePos :134-206
        u32 last = 0;
        for (u32 i = 1; i < len;) {
            u32 p = wp + i;
            t2[h2(w + p)] = pos; t3[h3(w + p)] = pos;
            if (i + 128 < len) { i += 4; pos += 4; continue; }
            if (width) {
                u32 h = h6(w + p), *b = &t6[(size_t)h * width];
                if (h != last) for (u32 j = width - 1; j > 0; j--) b[j] = b[j - 1];
                b[0] = pos; last = h;
            }
            pos++; i++;
        }
    }
    void slide_sampled(u32 wp, u32 len) {   // SlidePosFast :208-241
        for (u32 i = 0; i < len; i++, pos++) {
            if (h2(w + wp + i) % 16 || !width) continue;
            u32 *b = &t6[(size_t)h6(w + wp + i) * width];
            for (u32 j = width - 1; j > 0; j--) b[j] = b[j - 1];
            b[0] = pos;
        }
    }
    bool probe_duplicate(u32 wp, const u8 *s, u32 lim) const {   // TestFind :526-568
        if (h2(s) % 16 || !width) return false;
        u32 dist = pos - t6[(size_t)h6(s) * width];
        if (dist >= valid) return false;
        u32 cp = wp >= dist ? wp - dist : wp + wsize - dist;
        u32 cl = lim < wsize - cp ? lim : wsize - cp, l = 0;
        while (l < cl && s[l] == w[cp + l]) l++;
        return l > 18;
    }
    static bool better(Cand a, Cand b) {   // SecondMatchBetter :570-582
        static const u32 cof[4] = {0, 4, 8, 12};
        return b.len > 1 && (b.len > a.len + 3 || (b.len > a.len && b.dist <= 4) || (b.len + 2 > a.len && b.dist <= 4 && a.dist > 4) ||
                             (b.len >= a.len && (b.dist >> cof[b.len - a.len]) <= a.dist) ||
                             (b.len < a.len && b.len + 2 >= a.len && a.dist > 4 && (a.dist >> cof[a.len - b.len]) > b.dist));
    }
    Cand best(const u32 *reps, u32 wp, u32 lim) {   // FindMatch :497-524
        Cand c[32], r = {1, 0};
        u32 n = find(c, reps, wp, lim);
        for (u32 i = 0; i < n; i++) if (i == 0 || better(r, c[i])) r = c[i];
        return r;
    }
};

// ------------------------------------------------------------------------------------------
// LZ parsers (csc_lz.cpp) and block dispatch (csc_encoder_main.cpp)
struct Encoder {
    Props p;
    Coder coder;
    Model m;
    MF mf;
    Analyzer an;
    Dict dict;
    std::vector<u8> wnd;
    u32 wsize, wcur, reps[4];
    struct Node { u32 dist, state, price, lit, rep[4]; int back, next; };
    std::vector<Node> ap;
    std::vector<Cand> tab;

    void init(const Props &pp, std::vector<u8> *out) {
        p = pp;
        coder.init(out, p.blocksize);
        m.init(&coder);
        wsize = p.dict;
        wnd.assign((size_t)wsize + 8, 0);
        mf.init(wnd.data(), wsize, p.hbits, p.hwidth, p.good);
        wcur = 0;
        for (u32 &r : reps) r = wsize;
        ap.resize(2049); tab.resize(p.good + 2);
    }
    void emit(Cand u) {   // encode_nonlit :127-154
        if (u.dist <= 4) {
            if (u.len == 1 && u.dist == 1) { m.rep0len1(); return; }
            m.rep(u.dist - 1, u.len - 2);
            u32 d = reps[u.dist - 1];
            for (u32 j = u.dist - 1; j > 0; j--) reps[j] = reps[j - 1];
            reps[0] = d;
        } else {
            m.match(u.dist - 5, u.len - 2);
            reps[3] = reps[2]; reps[2] = reps[1]; reps[1] = reps[0]; reps[0] = u.dist - 4;
        }
    }
    void lazy(u32 size) {   // compress_normal :156-199
        Cand a = {1, 0}, b;
        bool have = false;
        for (u32 i = 0; i < size;) {
            if (!have) a = mf.best(reps, wcur, size - i);
            if (a.len == 1 || a.len >= p.good) {
                if (!a.dist) m.literal(wnd[wcur]); else emit(a);
                mf.slide(wcur, a.len);
                i += a.len; wcur += a.len;
                if (a.dist) m.ctx = wnd[wcur - 1];
                have = false;
                continue;
            }
            b = mf.best(reps, wcur + 1, size - i - 1);
            if (MF::better(a, b)) { m.literal(wnd[wcur]); i++; wcur++; a = b; have = true; }
            else { emit(a); mf.slide(wcur + 1, a.len - 1); i += a.len; wcur += a.len; m.ctx = wnd[wcur - 1]; have = false; }
        }
    }
    void priced(u32 state, const u32 *rp, u32 lim) {   // FindMatchWithPrice, csc_mf.cpp:584-625
        Cand c[32];
        u32 n = mf.find(c, rp, wcur, lim);
        tab[0] = n ? c[n - 1] : Cand{1, 0};
        if (tab[0].len >= p.good) return;
        tab[1].dist = 0;
        u32 l = 1;
        for (u32 i = 0; i < n; i++) {
            u32 dp, rd;
            if (c[i].len == 1 && c[i].dist == 1) { tab[1].len = m.price_rep0len1(state); tab[1].dist = 1; continue; }
            if (c[i].dist <= 4) { dp = m.price_rep(state, c[i].dist - 1); rd = 0; }
            else { dp = m.price_dist(state, c[i].dist - 5); rd = c[i].dist - 4; }
            while (l < c[i].len) {
                l++;
                if (l <= 6 && rd >= MF::bound(l)) { tab[l].dist = 0; continue; }
                tab[l].dist = c[i].dist;
                tab[l].len = dp + m.price_len(l - 2);
            }
        }
    }
    void backtrack(int end) {   // ap_backward :335-362
        for (int i = end; i;) { ap[ap[i].back].next = i; i = ap[i].back; }
        for (int i = 0; i != end;) {
            int nx = ap[i].next;
            u32 d = ap[nx].dist;
            if (!d) m.literal(ap[i].lit);
            else {
                if (d <= 4) { if (nx - i == 1 && d == 1) m.rep0len1(); else m.rep(d - 1, nx - i - 2); }
                else m.match(d - 5, nx - i - 2);
                m.ctx = ap[nx - 1].lit;
            }
            i = nx;
        }
        memcpy(reps, ap[end].rep, sizeof(reps));
    }
    void optimal(u32 size) {   // compress_advanced :207-333
        for (u32 i = 0; i < size;) {
            priced(m.state, reps, size - i);
            if (!tab[0].dist) { m.literal(wnd[wcur]); i++; wcur++; continue; }
            u32 cur = 0, end = 1, lim = size - i < 2048 ? size - i : 2048;
            ap[0].price = 0; ap[0].back = 0; ap[0].state = m.state;
            memcpy(ap[0].rep, reps, sizeof(reps));
            for (;;) {
                ap[cur].lit = wnd[wcur];
                if (cur) {
                    Node &n = ap[cur];
                    const Node &f = ap[n.back];
                    memcpy(n.rep, f.rep, sizeof(n.rep));
                    if (!n.dist) n.state = (f.state * 4) & 63;
                    else if (n.dist <= 4) {
                        if (cur - n.back == 1 && n.dist == 1) n.state = (f.state * 4 + 2) & 63;
                        else {
                            n.state = (f.state * 4 + 3) & 63;
                            u32 t = n.rep[n.dist - 1];
                            for (u32 j = n.dist - 1; j > 0; j--) n.rep[j] = n.rep[j - 1];
                            n.rep[0] = t;
                        }
                    } else {
                        n.state = (f.state * 4 + 1) & 63;
                        n.rep[0] = n.dist - 4; n.rep[1] = f.rep[0]; n.rep[2] = f.rep[1]; n.rep[3] = f.rep[2];
                    }
                    if (cur < lim) priced(n.state, n.rep, size - i - cur);
                }
                if (cur == lim) { backtrack(cur); i += cur; break; }
                if (tab[0].len == 1 && cur + 1 == end) {
                    backtrack(cur); m.literal(ap[cur].lit); i += cur + 1; wcur++; break;
                }
                if (cur + 1 >= end) ap[end++].price = 0xFFFFFFFFu;
                if (tab[0].len >= p.good || (tab[0].len > 1 && tab[0].len + cur >= lim)) {
                    backtrack(cur); i += cur;
                    emit(tab[0]);
                    mf.slide(wcur, tab[0].len);
                    i += tab[0].len; wcur += tab[0].len;
                    m.ctx = wnd[wcur - 1];
                    break;
                }
                u32 base = ap[cur].price;
                u32 lp = m.price_literal(ap[cur].state, wcur ? wnd[wcur - 1] : 0, wnd[wcur]);
                if (lp + base < ap[cur + 1].price) { ap[cur + 1].dist = 0; ap[cur + 1].back = cur; ap[cur + 1].price = lp + base; }
                if (tab[1].dist && tab[1].len + base < ap[cur + 1].price) { ap[cur + 1].dist = 1; ap[cur + 1].back = cur; ap[cur + 1].price = tab[1].len + base; }
                u32 ml = tab[0].len;
                while (cur + ml >= end) ap[end++].price = 0xFFFFFFFFu;
                for (u32 l = ml; l > 1; l--)
                    if (tab[l].dist && tab[l].len + base < ap[cur + l].price) { ap[cur + l].dist = tab[l].dist; ap[cur + l].back = cur; ap[cur + l].price = tab[l].len + base; }
                cur++; wcur++;
            }
        }
    }
    void lz(const u8 *s, u32 n, u32 mode) {   // LZ::EncodeNormal :61-100
        for (u32 i = 0; i < n;) {
            u32 part = wsize - wcur < n - i ? wsize - wcur : n - i;
            if (part > BLK) part = BLK;
            memcpy(&wnd[wcur], s + i, part);
            if (mode == 2) lazy(part); else if (mode == 3) optimal(part);
            else { mf.slide_sampled(wcur, part); wcur += part; }
            if (wcur >= wsize) wcur = 0;
            i += part;
        }
        if (mode != 5) m.match(64, 0);
    }
    bool duplicate(const u8 *s, u32 n) const {   // IsDuplicateBlock :102-112
        for (u32 i = 0; i < n; i++) if (mf.probe_duplicate(wcur, s + i, n - i)) return true;
        return false;
    }
    void run(u8 *s, u32 n, u32 type) {   // compress_block, csc_encoder_main.cpp:35-83
        if (!n) return;
        if (type == T_NORMAL || type == T_FAST) { coder.number(T_NORMAL); lz(s, n, p.mode); }
        else if (type == T_EXE) { coder.number(type); e8e9(s, n, false); lz(s, n, p.mode); }
        else if (type == T_TXT) {
            if (dict.forward(s, n)) { coder.number(type); coder.number(n); } else coder.number(T_NORMAL);
            lz(s, n, p.mode);
        } else if (type == T_BAD) { coder.number(type); lz(s, n, 5); m.raw_bytes(s, n); }
        else if (type == T_ENTROPY) { coder.number(type); lz(s, n, 5); m.literals(s, n); }
        else { coder.number(type); lz(s, n, 5); delta_fwd(s, n, kChannels[type - T_DLT]); m.rle(s, n); }
    }
    void segment(u8 *s, u32 n) {   // CSCEncoder::Compress :85-146
        u32 last = T_NORMAL, begin = 0, have = 0, bpb = 0;
        bool filters = p.fdlt + p.ftxt + p.fexe != 0;
        for (u32 i = 0; i < n;) {
            u32 cur = n - i < BLK ? n - i : BLK;
            u32 t = filters ? an.classify(s + i, cur, &bpb) : T_NORMAL;
            if (t == T_SKIP) t = last;
            if (t == T_EXE && !p.fexe) t = T_NORMAL;
            else if (t == T_TXT && !p.ftxt) t = T_NORMAL;
            else if (t >= T_DLT && !p.fdlt) t = T_NORMAL;
            if (t >= T_DLT && an.delta_bpb(s + i, cur, kChannels[t - T_DLT]) >= bpb * 0.95) t = T_NORMAL;
            if (t >= T_NOLZ && duplicate(s + i, cur)) t = T_NORMAL;
            if (t != last || have + cur > p.rawblock) {
                if (have) { run(s + begin, have, last); coder.number(0); }
                begin = i; have = 0;
            }
            last = t; have += cur; i += cur;
        }
        if (have) { run(s + begin, have, last); coder.number(1); coder.flush(); }
    }
};

// ------------------------------------------------------------------------------------------
// decoder (csc_dec.cpp)
struct Decoder {
    const u8 *in; size_t n, at;
    u32 bsize;
    std::vector<std::vector<u8> > pend[2];   // blocks of the other kind read ahead (csc_memio.cpp:55-77)
    size_t pend_at[2];
    std::vector<u8> buf[2];
    u32 cur[2];
    u32 range, code, nbits, acc;
    std::vector<u32> lit, dlt;
    u32 st[192], rp[192], dist[168], slot[2], e1[8], e2[8], e3[128], dx[464], longlen, rleflag, state, ctx;
    std::vector<u8> wnd;
    u32 wsize, wcur, reps[4];
    bool fail;

    bool fetch(int kind) {   // MemIO::ReadBlock :5-81
        if (pend_at[kind] < pend[kind].size()) { buf[kind].swap(pend[kind][pend_at[kind]++]); cur[kind] = 0; return true; }
        for (;;) {
            if (at >= n) return false;
            u8 h = in[at++];
            u32 sz;
            if (h & 0x40) sz = bsize;
            else { if (at + 3 > n) return false; sz = ((u32)in[at] << 16) | ((u32)in[at + 1] << 8) | in[at + 2]; at += 3; }
            if (!sz || sz > bsize || at + sz > n) return false;
            std::vector<u8> b(in + at, in + at + sz);
            at += sz;
            if ((h >> 7) == kind) { buf[kind].swap(b); cur[kind] = 0; return true; }
            pend[1 - kind].push_back(b);
        }
    }
    u8 next(int kind) {
        if (fail) return 0;
        u8 v = buf[kind][cur[kind]++];
        if (cur[kind] >= buf[kind].size() && !fetch(kind)) { buf[kind].assign(8, 0); cur[kind] = 0; eof_hits++; }
        return v;
    }
    u32 eof_hits;
    bool open() {   // start of a flush segment (csc_dec.cpp:336-345, 657-680)
        if (!fetch(1) || !fetch(0)) return false;
        if (buf[1].size() < 5) return false;
        range = 0xFFFFFFFFu; nbits = acc = 0;
        code = ((u32)buf[1][1] << 24) | ((u32)buf[1][2] << 16) | ((u32)buf[1][3] << 8) | buf[1][4];
        cur[1] = 5;
        if (cur[1] >= buf[1].size() && !fetch(1)) { buf[1].assign(8, 0); cur[1] = 0; }
        return true;
    }
    u32 bit(u32 &p) {   // DecodeBit :10-35
        if (range < (1u << 24)) { range <<= 8; code = (code << 8) + next(1); }
        u32 bound = (range >> 12) * p;
        if (code < bound) { range = bound; p += (0xFFF - p) >> 5; return 1; }
        range -= bound; code -= bound; p -= p >> 5;
        return 0;
    }
    u32 raw16(u32 k) {
        while (nbits < k) { acc = (acc << 8) | next(0); nbits += 8; }
        u32 r = (acc >> (nbits - k)) & ((1u << k) - 1);
        nbits -= k;
        return r;
    }
    u32 raw(u32 k) { if (k <= 16) return raw16(k); u32 h = raw16(k - 16) << 16; return h | raw16(16); }
    u32 number() { u32 s = raw(5), v = raw(s ? s : 1); return s ? v + (1u << s) : v; }
    u32 tree(u32 *p, u32 k) { u32 node = 1; for (u32 i = 0; i < k; i++) node = node * 2 + bit(p[node]); return node - (1u << k); }
    u32 len1() { if (!bit(slot[0])) return tree(e1, 3); if (!bit(slot[1])) return 8 + tree(e2, 3); return 16 + tree(e3, 7); }
    u32 len2() { u32 l = len1(); if (l != 143) return l; while (!bit(longlen) && !fail) l += 143; return l + len1(); }

    bool lz(std::vector<u8> &out, u32 limit) {   // lz_decode :476-571
        u32 i = 0;
        size_t start = out.size();
        while (i <= limit) {
            if (!bit(st[state * 3])) {
                u32 ch = tree(&lit[ctx * 256], 8);
                ctx = ch; state = (state * 4) & 63;
                wnd[wcur++] = (u8)ch; out.push_back((u8)ch); i++;
            } else {
                u32 d, l;
                if (bit(st[state * 3 + 1])) {
                    l = len2();
                    u32 base, sb;
                    if (l == 0) { base = 0; sb = 3; } else if (l <= 2) { base = 8 + 16 * (l - 1); sb = 4; }
                    else if (l <= 5) { base = 40 + 32 * (l - 3); sb = 5; } else { base = 136; sb = 5; }
                    u32 s = tree(dist + base, sb);
                    if (s <= 2) d = s;
                    else {
                        u32 eb = s - 2, hi = eb > 4 ? raw(eb - 4) : 0, lo = tree(dx + (eb - 1) * 16, 4);
                        lo = ((lo & 1) << 3) | ((lo & 2) << 1) | ((lo & 4) >> 1) | (lo >> 3);
                        d = (1u << eb) + 1 + (hi << 4) + lo;
                    }
                    state = (state * 4 + 1) & 63;
                    if (l == 0 && d == 64) break;
                    d++; l += 2;
                    reps[3] = reps[2]; reps[2] = reps[1]; reps[1] = reps[0]; reps[0] = d;
                } else if (!bit(st[state * 3 + 2])) {
                    state = (state * 4 + 2) & 63;
                    u32 cp = wcur > reps[0] ? wcur - reps[0] : wcur + wsize - reps[0];
                    if (cp >= wsize) return false;
                    u8 v = wnd[cp];
                    wnd[wcur++] = v; out.push_back(v); i++;
                    ctx = v;
                    if (wcur == wsize) wcur = 0;
                    continue;
                } else {
                    u32 hi = bit(rp[state * 3]), lo = bit(rp[state * 3 + 1 + hi]), idx = hi * 2 + lo;
                    l = len2() + 2;
                    state = (state * 4 + 3) & 63;
                    d = reps[idx];
                    for (u32 j = idx; j > 0; j--) reps[j] = reps[j - 1];
                    reps[0] = d;
                }
                u32 cp = wcur >= d ? wcur - d : wcur + wsize - d;
                if (fail || cp >= wsize || cp + l > wsize || l + i > limit || wcur + l > wsize) return false;
                for (u32 k = 0; k < l; k++) { u8 v = wnd[cp + k]; wnd[wcur + k] = v; out.push_back(v); }
                wcur += l; i += l;
                ctx = wnd[wcur - 1];
            }
            if (wcur == wsize) wcur = 0;
            if (fail) return false;
        }
        (void)start;
        return true;
    }
    void to_window(const u8 *s, u32 n) {   // lz_copy2dict :573-584
        for (u32 i = 0; i < n;) {
            u32 part = wsize - wcur < n - i ? wsize - wcur : n - i;
            if (part > BLK) part = BLK;
            memcpy(&wnd[wcur], s + i, part);
            wcur += part; if (wcur >= wsize) wcur = 0;
            i += part;
        }
    }
    int run(const u8 *stream, size_t len, std::vector<u8> &out) {   // CSCDec_* :586-777
        if (len < 10) return -1;
        wsize = ((u32)stream[0] << 24) | ((u32)stream[1] << 16) | ((u32)stream[2] << 8) | stream[3];
        bsize = ((u32)stream[4] << 16) | ((u32)stream[5] << 8) | stream[6];
        u32 rawblock = ((u32)stream[7] << 16) | ((u32)stream[8] << 8) | stream[9];
        if (wsize < 32 * KB || wsize > 1024 * MB || !bsize) return -1;
        in = stream + 10; n = len - 10; at = 0; pend_at[0] = pend_at[1] = 0; fail = false; eof_hits = 0;
        lit.assign(65536, 2048); dlt.clear();
        for (u32 &v : st) v = 2048; for (u32 &v : rp) v = 2048; for (u32 &v : dist) v = 2048;
        for (u32 &v : e1) v = 2048; for (u32 &v : e2) v = 2048; for (u32 &v : e3) v = 2048; for (u32 &v : dx) v = 2048;
        slot[0] = slot[1] = longlen = rleflag = 2048; state = ctx = 0;
        wnd.assign((size_t)wsize + 8, 0); wcur = 0; memset(reps, 0, sizeof(reps));
        Dict dict;
        if (!open()) return -1;
        for (;;) {
            u32 type = number();
            size_t mark = out.size();
            if (type == T_EOF) break;
            if (type == T_NORMAL || type == T_EXE || type == T_TXT) {
                if (type == T_TXT) number();
                if (!lz(out, rawblock)) return -96;
                if (type == T_EXE) e8e9(out.data() + mark, (u32)(out.size() - mark), true);
                if (type == T_TXT) dict.inverse(out.data() + mark, (u32)(out.size() - mark));
            } else if (type == T_BAD || type == T_ENTROPY || (type >= T_DLT && type < T_DLT + 5)) {
                u32 sz = number();
                if (sz > rawblock) return -1;
                std::vector<u8> b(sz);
                if (type == T_BAD) for (u32 i = 0; i < sz; i++) b[i] = (u8)raw16(8);
                else if (type == T_ENTROPY) for (u32 i = 0; i < sz; i++) { u32 ch = tree(&lit[ctx * 256], 8); ctx = ch; b[i] = (u8)ch; }
                else {   // decode_rle :107-153
                    if (dlt.empty()) dlt.assign(65536, 2048);
                    u32 cx = 0;
                    for (u32 i = 0; i < sz;) {
                        if (!bit(rleflag)) { u32 ch = tree(&dlt[cx * 256], 8); b[i++] = (u8)ch; cx = ch; }
                        else {
                            u32 l = len2() + 11;
                            if (!i) return -1;
                            while (l-- && i < sz) { b[i] = b[i - 1]; i++; }
                            cx = b[i - 1];
                        }
                        if (eof_hits > 64) return -98;
                    }
                    delta_inv(b.data(), sz, kChannels[type - T_DLT]);
                }
                to_window(b.data(), sz);
                out.insert(out.end(), b.begin(), b.end());
            } else return -96;
            if (eof_hits > 64) return -98;
            if (number() == 1 && !open()) return -1;
        }
        return 0;
    }
};

}  // namespace

extern "C" {

// same contract as oracle/ref_shim.cpp ref_encode (levels 1..4)
int rs_encode(const uint8_t *in, size_t n, uint32_t dict_size, int level, uint32_t read_size, int flt_mask, int with_props,
              uint8_t *out, size_t cap, size_t *out_n) {
    if (level < 1 || level > 4) return -91;
    Props p;
    props_for(p, dict_size, level);
    p.fdlt = flt_mask & 1; p.ftxt = (flt_mask >> 1) & 1; p.fexe = (flt_mask >> 2) & 1;
    std::vector<u8> res;
    if (with_props) {
        u8 h[10] = {(u8)(p.dict >> 24), (u8)(p.dict >> 16), (u8)(p.dict >> 8), (u8)p.dict, (u8)(p.blocksize >> 16), (u8)(p.blocksize >> 8),
                    (u8)p.blocksize, (u8)(p.rawblock >> 16), (u8)(p.rawblock >> 8), (u8)p.rawblock};
        res.insert(res.end(), h, h + 10);
    }
    Encoder *e = new Encoder;
    e->init(p, &res);
    if (!read_size || read_size > p.rawblock) read_size = p.rawblock;
    std::vector<u8> seg(read_size);
    for (size_t off = 0; off < n; off += read_size) {
        u32 k = (u32)(n - off < read_size ? n - off : read_size);
        memcpy(seg.data(), in + off, k);   // filters run in place on the caller's buffer copy (csc_enc.cpp:169-176)
        e->segment(seg.data(), k);
    }
    e->coder.number(T_EOF);   // CSCEnc_Encode_Flush
    e->coder.flush();
    delete e;
    *out_n = res.size();
    if (res.size() > cap) return -3;
    memcpy(out, res.data(), res.size());
    return 0;
}

int rs_decode(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_n) {
    std::vector<u8> res;
    Decoder *d = new Decoder;
    int r = d->run(in, n, res);
    delete d;
    *out_n = res.size();
    if (r < 0) return r;
    if (res.size() > cap) return -3;
    memcpy(out, res.data(), res.size());
    return 0;
}

}  // extern "C"
