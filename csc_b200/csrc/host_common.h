// host_common.h -- host-side pieces shared by the C-ABI library and the test
// host simulation: libm-derived tables, the level -> parameter
// mapping of CSCEncProps_Init, and the RC/BC block framing of MemIO.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>

#include "csc_defs.h"
#include "blocks.cuh"
#include "../../include/csc_common.h"

namespace cscb {

// ---- tables ----------------------------------------------------------------------
inline void build_analyzer_tables(AnalyzerTables &T) {
    // same expressions as csc_analyzer.cpp:11-14 (host libm, truncation to u32)
    for (u32 i = 0; i < (kBlock >> 4); i++) T.logtab[i] = (double)100 * log((double)i * 16 + 8) / log((double)2);
    T.logtab[kBlock >> 4] = (double)100 * log((double)(kBlock)) / log((double)2);
    for (u32 d = 0; d < 16; d++) T.ent_thr[d] = 0;
    // (log(diffNum - 2)/log 2 - 0.6) * 100.0, then * size at the use site (csc_analyzer.cpp:223)
    for (u32 d = 6; d < 16; d++) T.ent_thr[d] = (log((double)d - 2) / log((double)2) - 0.6) * 100.0;
}

inline void build_p2bits(u32 *t) {
    // csc_model.cpp:68-70: the float argument selects the float log overload
    for (int i = 0; i < (4096 >> 3); i++) t[i] = (uint32_t)(128 * log((float)(i * 8 + 4) / 4096) / log(0.5));
}

// the 122 n-grams of the English dictionary filter (data of the stream format,
// csc_filters.cpp:8-38); symbol of word k (1-based) is 0x81 + k
static const char *const kDictWords[] = {
    "ac", "ad", "ai", "al", "am", "an", "ar", "as", "at", "ea", "ec", "ed", "ee", "el", "en", "er", "es", "et",
    "id", "ie", "ig", "il", "in", "io", "is", "it", "of", "ol", "on", "oo", "or", "os", "ou", "ow", "ul", "un",
    "ur", "us", "ba", "be", "ca", "ce", "co", "ch", "de", "di", "ge", "gh", "ha", "he", "hi", "ho", "ra", "re",
    "ri", "ro", "rs", "la", "le", "li", "lo", "ld", "ll", "ly", "se", "si", "so", "sh", "ss", "st", "ma", "me",
    "mi", "ne", "nc", "nd", "ng", "nt", "pa", "pe", "ta", "te", "ti", "to", "th", "tr", "wa", "ve", "all", "and",
    "but", "dow", "for", "had", "hav", "her", "him", "his", "man", "mor", "not", "now", "one", "out", "she",
    "the", "was", "wer", "whi", "whe", "wit", "you", "any", "are", "that", "said", "with", "have", "this",
    "from", "were", "tion"};
static const u32 kDictWordCount = sizeof(kDictWords) / sizeof(kDictWords[0]);

inline void build_dict_tables(DictTables &D) {   // MakeWordTree, csc_filters.cpp:87-111
    memset(&D, 0, sizeof(D));
    u32 nodes = 1;
    u32 sym = 0x82;
    for (u32 w = 0; w < kDictWordCount; w++) {
        u32 node = 0;
        const char *s = kDictWords[w];
        for (u32 j = 0; s[j]; j++) {
            u32 idx = (u32)(s[j] - 'a');
            if (!D.next[node][idx]) D.next[node][idx] = (u16)nodes++;
            node = D.next[node][idx];
        }
        D.symbol[node] = (u8)sym;
        strncpy((char *)D.word[sym], s, 7);
        sym++;
    }
    D.max_symbol = sym;
}

// ---- CSCEncProps_Init (csc_enc.cpp:16-97) -------------------------------------------
inline void props_init(CSCProps *p, uint32_t dict_size, int level) {
    const uint32_t KB = 1024, MB = 1048576;
    dict_size += 10 * KB;
    if (dict_size < 32 * KB) dict_size = 32 * KB;
    if (dict_size > 1024 * MB) dict_size = 1024 * MB;
    if (level < 1) level = 1;
    if (level > 5) level = 5;
    p->dict_size = dict_size;
    p->DLTFilter = p->TXTFilter = p->EXEFilter = 1;
    p->csc_blocksize = 64 * KB;
    p->raw_blocksize = 2 * MB;
    uint32_t hb = dict_size < MB ? 19 : dict_size <= 4 * MB ? 20 : dict_size <= 16 * MB ? 21
                  : dict_size <= 64 * MB ? 22 : dict_size <= 256 * MB ? 23 : 24;
    while (((uint32_t)1 << hb) > dict_size) hb--;
    uint32_t bt;
    if (dict_size <= 16 * MB) bt = dict_size;
    else if (dict_size <= 64 * MB) bt = (dict_size - 16 * MB) / 2 + 16 * MB;
    else if (dict_size <= 256 * MB) bt = (dict_size - 64 * MB) / 4 + 40 * MB;
    else bt = (dict_size - 256 * MB) / 8 + 88 * MB;
    struct Lv { uint8_t width, mode, good; int dbits; bool bt; };
    static const Lv lv[6] = {{0, 0, 0, 0, false}, {1, 2, 32, +1, false}, {8, 2, 24, -1, false},
                             {2, 3, 16, +1, false}, {8, 3, 24, -1, false}, {0, 3, 48, 0, true}};
    const Lv &L = lv[level];
    p->bt_size = L.bt ? bt : 0;
    p->good_len = L.good;
    p->hash_bits = (uint8_t)(hb + L.dbits);
    p->bt_hash_bits = (uint8_t)(hb + 1);
    p->hash_width = L.width;
    p->lz_mode = L.mode;
    p->bt_cyc = L.bt ? 32 : 0;   // the reference leaves it uninitialised below -m5 (unused there)
    if (p->bt_size == p->dict_size) p->hash_width = 0;
}

inline uint64_t est_mem_usage(const CSCProps *p) {   // csc_enc.cpp:99-112
    const uint64_t KB = 1024, MB = 1048576;
    uint64_t r = p->dict_size;
    r += p->csc_blocksize * 2;
    if (p->bt_size) r += ((1 << p->bt_hash_bits) + 2 * p->bt_size) * sizeof(uint32_t);
    if (p->hash_width) r += (p->hash_width * (1 << p->hash_bits)) * sizeof(uint32_t);
    r += 80 * KB * sizeof(uint32_t);
    r += 256 * 256 * sizeof(uint32_t) * 2;
    r += 2 * MB;
    return r;
}

inline void write_props(const CSCProps *p, uint8_t *s) {   // csc_enc.cpp:145-158
    s[0] = (p->dict_size >> 24) & 0xff; s[1] = (p->dict_size >> 16) & 0xff;
    s[2] = (p->dict_size >> 8) & 0xff;  s[3] = p->dict_size & 0xff;
    s[4] = (p->csc_blocksize >> 16) & 0xff; s[5] = (p->csc_blocksize >> 8) & 0xff; s[6] = p->csc_blocksize & 0xff;
    s[7] = (p->raw_blocksize >> 16) & 0xff; s[8] = (p->raw_blocksize >> 8) & 0xff; s[9] = p->raw_blocksize & 0xff;
}
inline void read_props(CSCProps *p, const uint8_t *s) {   // csc_dec.cpp:733-738
    p->dict_size = ((uint32_t)s[0] << 24) + (s[1] << 16) + (s[2] << 8) + s[3];
    p->csc_blocksize = ((uint32_t)s[4] << 16) + (s[5] << 8) + s[6];
    p->raw_blocksize = ((uint32_t)s[7] << 16) + (s[8] << 8) + s[9];
}

inline EncCfg cfg_from_props(const CSCProps *p) {   // LZ::Init / MatchFinder::Init
    EncCfg c;
    memset(&c, 0, sizeof(c));
    u32 w = (u32)p->dict_size;
    if (w < 32 * 1024) w = 32 * 1024;
    if (w > 1024u * 1048576u) w = 1024u * 1048576u;
    c.wnd_size = w;
    c.vld_rge = w - kBlock - 4;
    c.good_len = p->good_len;
    c.ht_width = (p->hash_bits && p->hash_width) ? p->hash_width : 0;
    c.ht_bits = c.ht_width ? p->hash_bits : 0;
    c.lz_mode = p->lz_mode;
    c.flt_dlt = p->DLTFilter; c.flt_txt = p->TXTFilter; c.flt_exe = p->EXEFilter;
    c.csc_blocksize = p->csc_blocksize;
    c.bt_bits = (p->bt_hash_bits && p->bt_size) ? p->bt_hash_bits : 0;
    c.bt_size = c.bt_bits ? p->bt_size : 0;
    c.bt_cyc = p->bt_cyc;
    if (c.bt_bits && !c.ht_width) {
        // -m5 (csc_enc.cpp:84-91): the binary-tree search becomes a bounded per-position search over the
        // kBtRowK nearest earlier positions with equal 6-byte hash (ratio-gated, not stream-identical).
        c.ht_width = kBtRowK;
        c.ht_bits = kBtHashBits;
        c.bounded = 1;
        if (c.good_len > 32) c.good_len = 32;   // the optimal parser's 32 virtual lanes
        c.bt_bits = 0; c.bt_size = 0;
    }
    return c;
}

// ---- MemIO framing (csc_memio.cpp:83-108) + the stale flush byte (H3) -------------
// Replays one stream's segments into the byte sequence the reference hands to
// ISeqOutStream::Write.  `ring` emulates the never-cleared rc_buf_ (csc_coder.cpp:46-47).
struct Framer {
    u32 bsize;
    std::vector<u8> ring;
    explicit Framer(u32 blocksize) : bsize(blocksize), ring(blocksize, 0) {}

    template <class Sink> void block(Sink &sink, const u8 *data, u32 size, int rc) {
        u8 hdr[4];
        hdr[0] = (u8)(rc << 7);
        if (size == bsize) { hdr[0] |= 0x40; sink(hdr, 1); }
        else { sink(hdr, 1); hdr[1] = (size >> 16) & 0xff; hdr[2] = (size >> 8) & 0xff; hdr[3] = size & 0xff; sink(hdr + 1, 3); }
        if (size) sink(data, size);
    }
    // rc/bc: the segment's sub-streams as produced by the engine (rc's last byte is the
    // skipped one); events: 1 = RC block filled, 0 = BC block filled, in temporal order
    template <class Sink> void segment(Sink &sink, u8 *rc, u32 rc_len, const u8 *bc, u32 bc_len, const u8 *ev, u32 nev) {
        u32 rpos = 0, bpos = 0;
        for (u32 i = 0; i < nev; i++) {
            if (ev[i]) { memcpy(ring.data(), rc + rpos, bsize); block(sink, rc + rpos, bsize, 1); rpos += bsize; }
            else { block(sink, bc + bpos, bsize, 0); bpos += bsize; }
        }
        u32 rrem = rc_len - rpos;          // >= 1: includes the skipped byte
        if (rrem > 1) memcpy(ring.data(), rc + rpos, rrem - 1);
        rc[rpos + rrem - 1] = ring[rrem - 1];   // value last written at this ring offset
        block(sink, rc + rpos, rrem, 1);
        block(sink, bc + bpos, bc_len - bpos, 0);
    }
    // EncodeInt(SIG_EOF) + Flush from a reset coder (csc_encoder_main.cpp:149-157)
    template <class Sink> void eof(Sink &sink) {
        u8 rc[6] = {0, 0, 0, 0, 0, 0};
        rc[5] = ring[5];
        memset(ring.data(), 0, 5);
        u8 bc[3] = {0x19, 0x00, 0x00};
        block(sink, rc, 6, 1);
        block(sink, bc, 3, 0);
    }
};


// Temporal order of a segment's block completions (1 = RC block, 0 = BC block) from the two time-stamp
// lists the device passes keep (csc_defs.h: SegOut); RC stamps are odd, BC stamps even.
inline void merge_events(const u32 *rc_ev, u32 nr, const u32 *bc_ev, u32 nb, std::vector<u8> &ev) {
    ev.clear();
    u32 i = 0, j = 0;
    while (i < nr || j < nb) {
        if (j >= nb || (i < nr && rc_ev[i] < bc_ev[j])) { ev.push_back(1); i++; }
        else { ev.push_back(0); j++; }
    }
}

// ---- MemIO de-framing (csc_memio.cpp:5-81) ---------------------------------------
// Splits a libcsc stream (after the 10-byte header) into the concatenated RC and BC block
// payloads plus the payload start offset of every block (+ one sentinel).
struct DecInput {
    std::vector<u8> rc, bc;
    std::vector<u32> rc_off, bc_off;
};
inline int deframe(const u8 *s, size_t n, u32 bsize, DecInput &d) {
    size_t i = 0;
    d.rc.clear(); d.bc.clear(); d.rc_off.clear(); d.bc_off.clear();
    d.rc.reserve(n + 16); d.bc.reserve(n / 4 + 64);
    while (i < n) {
        u8 fb = s[i++];
        u32 sz;
        if ((fb >> 6) & 1) sz = bsize;
        else {
            if (i + 3 > n) return -1;
            sz = ((u32)s[i] << 16) + ((u32)s[i + 1] << 8) + s[i + 2];
            i += 3;
        }
        if (!sz || sz > bsize) return -1;   // "must be a abnormal stream" (csc_memio.cpp:43-46)
        if (i + sz > n) return -1;
        if (fb >> 7) { d.rc_off.push_back((u32)d.rc.size()); d.rc.insert(d.rc.end(), s + i, s + i + sz); }
        else { d.bc_off.push_back((u32)d.bc.size()); d.bc.insert(d.bc.end(), s + i, s + i + sz); }
        i += sz;
    }
    d.rc_off.push_back((u32)d.rc.size());
    d.bc_off.push_back((u32)d.bc.size());
    d.rc.resize(d.rc.size() + 16, 0);
    d.bc.resize(d.bc.size() + 16, 0);
    return 0;
}

}  // namespace cscb
