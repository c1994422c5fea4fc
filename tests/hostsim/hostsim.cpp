// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the very same engine sources the CUDA kernels are built from
// (csc_b200/csrc/enc_engine.cuh, dec_engine.cuh, blocks.cuh, plan.h) for the
// host with a single-lane "warp", and feeds them match-finder arrays produced
// by a brute-force CPU builder.  There is no GPU in the development container,
// so this is how the serial logic is iterated against the oracle before a
// gpurun call; the product library never links or calls any of this.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include <stdio.h>

#include "../../csc_b200/csrc/host_common.h"
#include "../../csc_b200/csrc/plan.h"
#include "../../csc_b200/csrc/enc_engine.cuh"
#include "../../csc_b200/csrc/rc_engine.cuh"
#include "../../csc_b200/csrc/dec_engine.cuh"
#include "../../csc_b200/csrc/frame.cuh"

using namespace cscb;

namespace {

struct Sink {
    std::vector<u8> *v;
    void operator()(const u8 *p, u32 n) { v->insert(v->end(), p, p + n); }
};

// Brute-force "last-K prior occurrences per position" (what kernels.cuh: k_mf_hash / k_mf_rows / k_mf_extend compute with a
// sort + segmented scan) through an actual last-occurrence table, over the sub-block grid of plan.h.
void build_mf_cpu(const u8 *lz, u32 n, const std::vector<SubBlock> &subs, u32 wnd, u32 bits, u32 K, std::vector<u32> &P6,
                  std::vector<u16> &L6, std::vector<u32> &Z2, std::vector<u32> &Z6) {
    u32 words = (n + 31) / 32 + 1;
    Z2.assign(words, 0);
    Z6.assign(words, 0);
    P6.assign((size_t)n * (K ? K : 1), kNone);
    L6.assign((size_t)n * (K ? K : 1), 0);
    std::vector<u32> last;
    if (K) last.assign((size_t)1 << bits, kNone);
    for (const SubBlock &sb : subs) {
        u32 sb_end = sb.x + sb.size;
        for (u32 x = sb.x; x < sb_end; x++) {
            u32 avail = sb_end - x;
            if (hash2_at(lz, x, sb_end, wnd) % 16 == 0) Z2[x >> 5] |= 1u << (x & 31);
            if (!K) continue;
            u32 h = hash6_at(lz, x, sb_end, wnd, bits);
            if (h == 0) Z6[x >> 5] |= 1u << (x & 31);
            u32 prev = last[h];
            last[h] = x;
            u32 e = prev;
            for (u32 k = 0; k < K && e != kNone; k++) {
                P6[(size_t)x * K + k] = e;
                u32 l = 0;
                while (l < avail && lz[x + l] == lz[e + l]) l++;
                L6[(size_t)x * K + k] = (u16)l;
                e = P6[(size_t)e * K];
            }
        }
    }
}


// -m5 bounded search rows (csc_defs.h: kBt*): nearest earlier positions by 6-byte hash, 8- and 12-byte context
// hash, merged by distance, duplicates removed -- what k_mf_hash / k_mf_rows / k_mf_merge compute on the device.
// EXPERIMENT hook: CSC_M5_CLASSES="6:5,8:5,12:6" overrides the context classes (len:count), generic byte hash.
void build_mf5_cpu(const u8 *lz, u32 n, const std::vector<SubBlock> &subs, u32 wnd, std::vector<u32> &P6, std::vector<u16> &L6) {
    const u32 bits = kBtHashBits, K = kBtRowK;
    struct Cls { u32 len, cnt; std::vector<u32> last, prev; };
    std::vector<Cls> cls;
    const char *env = getenv("CSC_M5_CLASSES");
    if (env) {
        for (const char *q = env; *q;) { char *e; u32 l = strtoul(q, &e, 10); q = e + 1; u32 c = strtoul(q, &e, 10); q = *e ? e + 1 : e; cls.push_back({l, c, {}, {}}); }
    } else {
        cls.push_back({6, kBtKA, {}, {}}); cls.push_back({kBtCtxB, kBtKB, {}, {}}); cls.push_back({kBtCtxC, kBtKC, {}, {}});
    }
    for (auto &c : cls) { c.last.assign((size_t)1 << bits, kNone); c.prev.assign(n, kNone); }
    P6.assign((size_t)n * K, kNone);
    L6.assign((size_t)n * K, 0);
    for (const SubBlock &sb : subs) {
        u32 sb_end = sb.x + sb.size;
        for (u32 x = sb.x; x < sb_end; x++) {
            u32 avail = sb_end - x;
            std::vector<u32> c;
            for (auto &cl : cls) {
                u32 h;
                if (cl.len == 6) h = hash6_at(lz, x, sb_end, wnd, bits);
                else if (!env) h = hash_ctx_at(lz, x, sb_end, wnd, cl.len, bits);
                else { u32 f = 2166136261u; for (u32 k = 0; k < cl.len; k++) f = (f ^ window_byte(lz, x + k, sb_end, wnd)) * 16777619u; h = (f * 0x27D4EB2Fu) >> (32 - bits); }
                cl.prev[x] = cl.last[h]; cl.last[h] = x;
                u32 q = cl.prev[x];
                for (u32 k = 0; k < cl.cnt && q != kNone; k++) { c.push_back(q); q = cl.prev[q]; }
            }
            std::sort(c.begin(), c.end(), [](u32 u, u32 v) { return u > v; });
            c.erase(std::unique(c.begin(), c.end()), c.end());
            for (u32 k = 0; k < K && k < c.size(); k++) {
                P6[(size_t)x * K + k] = c[k];
                u32 l = 0;
                while (l < avail && lz[x + l] == lz[c[k] + l]) l++;
                L6[(size_t)x * K + k] = (u16)l;
            }
        }
    }
}

// nearest earlier position with equal (padded) HASH3 / HASH2 among the sub-blocks not statically excluded
void build_chain23_cpu(const u8 *lz, u32 n, const std::vector<SubBlock> &subs, u32 wnd, int which, std::vector<u32> &P,
                       std::vector<u16> &L) {
    P.assign(n + 1, kNone);
    L.assign(n + 1, 0);
    std::vector<u32> last(which == 3 ? 65536 : 16384, kNone);
    for (const SubBlock &sb : subs) {
        if (sb.flags & 1) continue;
        u32 sb_end = sb.x + sb.size;
        for (u32 x = sb.x; x < sb_end; x++) {
            u32 avail = sb_end - x;
            u32 h = which == 3 ? hash3_at(lz, x, sb_end, wnd) : hash2_at(lz, x, sb_end, wnd);
            u32 e = last[h];
            last[h] = x;
            if (e == kNone) continue;
            P[x] = e;
            u32 l = 0;
            while (l < avail && lz[x + l] == lz[e + l]) l++;
            L[x] = (u16)l;
        }
    }
}

}  // namespace

static std::vector<u8> g_frame_dev;   // the stream as frame.cuh writes it (checked against Framer in hs_encode)
static int g_replans = 0;   // how often hs_encode had to plan again (tests look at it)
struct StageDump { u32 *binfo, *bbpb; u8 *lz; u32 *P6; u16 *L6; u32 *Z2, *Z6; u8 *dict_ok; u32 *tail_dlt; bool on; };
static StageDump g_dump = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, false};

extern "C" {

// capture the intermediate arrays of the next hs_encode (compared with the GPU's in tests)
void hs_debug_dump(uint32_t *binfo, uint32_t *bbpb, uint8_t *lz, uint32_t *P6, uint16_t *L6, uint32_t *Z2, uint32_t *Z6,
                   uint8_t *dict_ok, uint32_t *tail_dlt) {
    g_dump.binfo = binfo; g_dump.bbpb = bbpb; g_dump.lz = lz; g_dump.P6 = P6; g_dump.L6 = L6; g_dump.Z2 = Z2; g_dump.Z6 = Z6;
    g_dump.dict_ok = dict_ok; g_dump.tail_dlt = tail_dlt; g_dump.on = binfo != nullptr;
}

// Whole stream through the host-simulated product path. Same contract as oracle ref_encode.
// trace: optional (kind,a,b) triples; returns 0 or negative.
int hs_encode(const uint8_t *in, size_t n_, uint32_t dict_size, int level, uint32_t read_size, int flt_mask,
              int with_props, uint8_t *out, size_t cap, size_t *out_n, uint32_t *trace, size_t trace_cap,
              size_t *trace_n) {
    u32 n = (u32)n_;
    CSCProps p;
    props_init(&p, dict_size, level);
    p.DLTFilter = flt_mask & 1; p.TXTFilter = (flt_mask >> 1) & 1; p.EXEFilter = (flt_mask >> 2) & 1;
    EncCfg cfg = cfg_from_props(&p);
    if ((u64)n + cfg.vld_rge >= 0xFFFFFFF0ull) return kErrUnsupported;   // MatchFinder::normalize (csc_mf.cpp:108-114) is not modelled
    std::vector<u8> res;
    Sink sink{&res};
    if (with_props) { u8 hdr[10]; write_props(&p, hdr); sink(hdr, 10); }
    Framer fr(p.csc_blocksize);
    if (n) {
        if (!read_size || read_size > p.raw_blocksize) read_size = p.raw_blocksize;
        std::vector<u32> seg, seg_blk0;
        u32 nblk = 0;
        for (u32 a = 0; a < n; a += read_size) {
            u32 s = n - a < read_size ? n - a : read_size;
            seg.push_back(s);
            seg_blk0.push_back(nblk);
            nblk += (s + kBlock - 1) / kBlock;
        }
        std::vector<u8> flip(nblk, 0);   // assumed outcomes of the duplicate-block test (plan.h); refined on kErrReplan
        std::vector<u8> rc, bc;
        std::vector<u32> rc_ev, bc_ev;
        std::vector<DecEntry> dec;
        std::vector<SegOut> so;
        u32 tn = 0;
        for (;;) {
        std::vector<u8> lz(in, in + n);
        lz.resize(n + 64, 0);
        AnalyzerTables AT;
        build_analyzer_tables(AT);
        DictTables DT;
        build_dict_tables(DT);
        std::vector<u32> binfo(nblk), bbpb(nblk), tail_dlt(seg.size() * 5, 0);
        std::vector<u8> dict_ok(nblk, 0), nolz(nblk, 0);
        std::vector<SubBlock> subs;
        u32 a = 0;
        bool use_filters = (cfg.flt_dlt + cfg.flt_txt + cfg.flt_exe) != 0;
        for (u32 s = 0; s < seg.size(); s++) {
            u32 b = seg_blk0[s];
            for (u32 off = 0; off < seg[s]; off += kBlock, b++) {
                u32 sz = seg[s] - off < kBlock ? seg[s] - off : kBlock;
                if (use_filters) analyze_block(lz.data() + a + off, sz, AT, &binfo[b], &bbpb[b], &tail_dlt[s * 5]);
                else { binfo[b] = DT_NORMAL; bbpb[b] = 0; }
            }
            const u32 nb = (seg[s] + kBlock - 1) / kBlock;
            std::vector<FilterRun> runs(nb + 1);
            std::vector<SubBlock> sg(nb + wraps_inside(a, seg[s], cfg.wnd_size) + 1);
            u32 nsg = 0;
            u32 nr = plan_segment(binfo.data(), bbpb.data(), &tail_dlt[s * 5], flip.data(), seg_blk0[s], a, seg[s], cfg.wnd_size,
                                  cfg.flt_dlt, cfg.flt_txt, cfg.flt_exe, runs.data(), nolz.data(), sg.data(), &nsg);
            subs.insert(subs.end(), sg.begin(), sg.begin() + nsg);
            std::vector<u8> tmp;
            for (u32 r = 0; r < nr; r++) {
                if (runs[r].kind == DT_EXE) e8e9_run(lz.data() + runs[r].off, runs[r].size, false);
                else {
                    tmp.resize(runs[r].size);
                    dict_ok[runs[r].first_blk] = (u8)dict_forward(lz.data() + runs[r].off, runs[r].size, tmp.data(), DT);
                }
            }
            a += seg[s];
        }
        std::vector<u32> P6, Z2, Z6;
        std::vector<u16> L6;
        u32 K = cfg.ht_width;
        build_mf_cpu(lz.data(), n, subs, cfg.wnd_size, cfg.ht_bits, K, P6, L6, Z2, Z6);
        if (cfg.bounded) build_mf5_cpu(lz.data(), n, subs, cfg.wnd_size, P6, L6);
        std::vector<u32> P3, P2;
        std::vector<u16> L3, L2;
        build_chain23_cpu(lz.data(), n, subs, cfg.wnd_size, 3, P3, L3);
        build_chain23_cpu(lz.data(), n, subs, cfg.wnd_size, 2, P2, L2);
        std::vector<u32> D23((n + 31) / 32 + 1, 0);
        // precomputed table candidates (k_mf_tlist on the device): the same scalar statement, per LZ-parsed sub-block
        std::vector<TRec> T;
        std::vector<u32> B6((n + 31) / 32 + 2, 0);
        if (n <= cfg.wnd_size && !cfg.bounded && !getenv("CSC_HS_NO_TREC")) {
            T.assign(n + 1, TRec());
            for (const SubBlock &sb : subs) {
                if (sb.flags & 1) continue;
                for (u32 i = 0; i < sb.size; i++) {
                    const u32 x = sb.x + i;
                    T[x] = Encoder::tlist_of(x, sb.size - i, cfg.vld_rge, cfg.ht_width, K ? K : 1, cfg.good_len, P2[x], L2[x], P3[x], L3[x],
                                             P6.data() + (size_t)x * (K ? K : 1), L6.data() + (size_t)x * (K ? K : 1));
                    if ((K && x && P6[(size_t)x * K] == x - 1) || ((Z6[x >> 5] >> (x & 31)) & 1u)) B6[x >> 5] |= 1u << (x & 31);
                }
            }
        }
        if (g_dump.on) {
            memcpy(g_dump.binfo, binfo.data(), nblk * 4); memcpy(g_dump.bbpb, bbpb.data(), nblk * 4);
            memcpy(g_dump.lz, lz.data(), n); memcpy(g_dump.P6, P6.data(), (size_t)n * K * 4);
            memcpy(g_dump.L6, L6.data(), (size_t)n * K * 2); memcpy(g_dump.Z2, Z2.data(), (size_t)(n + 31) / 32 * 4);
            memcpy(g_dump.Z6, Z6.data(), (size_t)(n + 31) / 32 * 4); memcpy(g_dump.dict_ok, dict_ok.data(), nblk);
            memcpy(g_dump.tail_dlt, tail_dlt.data(), seg.size() * 5 * 4);
        }
        std::vector<u32> D6((n + 31) / 32 + 1, 0), ht2(16384, 0), ht3(65536, 0);
        std::vector<u16> p_lit(131072, 2048);
        std::vector<ApUnit> ap(kApLimit + 2);
        u32 rc_cap = n * 2 + 65536 * 2, bc_cap = n * 2 + 65536 * 2;
        rc.assign(rc_cap, 0); bc.assign(bc_cap, 0);
        rc_ev.assign(rc_cap / 32768 + 64, 0); bc_ev.assign(bc_cap / 32768 + 64, 0);
        dec.assign((size_t)n * 9 + 64 * seg.size() + 1024, 0);
        so.assign(seg.size(), SegOut());
        i32 status = 0;
        u32 replan[2] = {0, 0};
        tn = 0;
        EncStream S;
        memset(&S, 0, sizeof(S));
        std::vector<u8> tmpbuf(n + 64);
        S.tmp = tmpbuf.data();
        S.cfg = cfg; S.lz = lz.data(); S.n = n; S.nseg = (u32)seg.size();
        S.seg_size = seg.data(); S.seg_blk0 = seg_blk0.data(); S.binfo = binfo.data(); S.bbpb = bbpb.data();
        S.tail_dlt = tail_dlt.data(); S.dict_ok = dict_ok.data();
        S.P6 = P6.data(); S.L6 = L6.data(); S.K = K ? K : 1; S.Z2 = Z2.data(); S.Z6 = Z6.data(); S.D6 = D6.data();
        S.P3 = P3.data(); S.L3 = L3.data(); S.P2 = P2.data(); S.L2 = L2.data(); S.D23 = D23.data(); S.blk_nolz = nolz.data();
        S.flip = flip.data(); S.replan = replan; S.T = T.empty() ? nullptr : T.data(); S.B6 = B6.data();
        S.ht2 = ht2.data(); S.ht3 = ht3.data(); S.p_lit = p_lit.data(); S.p_delta = p_lit.data() + 65536; S.ap = ap.data();
        S.rc_out = rc.data(); S.rc_cap = rc_cap; S.bc_out = bc.data(); S.bc_cap = bc_cap;
        S.dec = dec.data(); S.dec_cap = (u32)dec.size();
        S.rc_ev = rc_ev.data(); S.bc_ev = bc_ev.data(); S.ev_cap = (u32)rc_ev.size(); S.segout = so.data(); S.status = &status;
        S.trace = trace; S.trace_cap = (u32)trace_cap; S.trace_n = &tn;
        ModelSmall *M = new ModelSmall;
        static u32 p2bits_tab[512];
        build_p2bits(p2bits_tab);
        M->p2bits = p2bits_tab;
        Encoder *E = new Encoder(S, *M);
        E->run();
        delete E;
        delete M;
        if (status == kErrReplan) { flip[replan[0]] = (u8)replan[1]; g_replans++; continue; }   // plan again with the observed outcome
        if (trace_n) *trace_n = tn;
        if (status < 0) return status;
        RcCoder R;   // the range-coder pass over the decision lists (k_rc on the device)
        R.run(S, nullptr);
        if (status < 0) return status;
        break;
        }
        // the device-side framing statement (frame.cuh, k_frame) on the same sub-streams: must give the bytes Framer gives
        FrameStream FS;
        memset(&FS, 0, sizeof(FS));
        FS.rc = rc.data(); FS.bc = bc.data(); FS.rc_ev = rc_ev.data(); FS.bc_ev = bc_ev.data(); FS.so = so.data();
        FS.nseg = (u32)seg.size(); FS.bsize = p.csc_blocksize; FS.with_props = with_props ? 1 : 0;
        const i32 frame_status = 0;
        FS.status = &frame_status;
        write_props(&p, FS.props);
        g_frame_dev.assign(frame_size(FS), 0);
        frame_write(FS, g_frame_dev.data());
        std::vector<u8> ev;
        for (u32 s = 0; s < seg.size(); s++) {
            merge_events(rc_ev.data() + so[s].rev_off, so[s].rev_len, bc_ev.data() + so[s].bev_off, so[s].bev_len, ev);
            fr.segment(sink, rc.data() + so[s].rc_off, so[s].rc_len, bc.data() + so[s].bc_off, so[s].bc_len, ev.data(), (u32)ev.size());
        }
    }
    fr.eof(sink);
    if (n && (g_frame_dev.size() != res.size() || memcmp(g_frame_dev.data(), res.data(), res.size()))) return -77;   // frame.cuh disagrees with Framer
    *out_n = res.size();
    if (res.size() > cap) return kErrOverflow;
    memcpy(out, res.data(), res.size());
    return 0;
}

int hs_replans(void) { int r = g_replans; g_replans = 0; return r; }

// frame.cuh against Framer on synthetic sub-streams: `nseg` segments with random RC / BC lengths around multiples of the
// block size (so that segments with and without full blocks, long and short remainders follow each other: every
// branch of the stale-byte search), random interleaving of the block completions.  Returns 0 when the bytes agree.
int hs_frame_selftest(uint32_t seed, uint32_t nseg, uint32_t bsize) {
    u32 x = seed * 2654435761u + 12345u;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 17; x ^= x << 5; return x; };
    std::vector<u8> rc, bc;
    std::vector<u32> rc_ev, bc_ev;
    std::vector<SegOut> so(nseg);
    for (u32 s = 0; s < nseg; s++) {
        SegOut &o = so[s];
        memset(&o, 0, sizeof(o));
        const u32 nr = rnd() % 4, nb = rnd() % 3;                    // full blocks
        const u32 rrem = 1 + (rnd() % 3 == 0 ? rnd() % 8 : rnd() % bsize);   // remainder incl. the skipped byte (1 .. bsize)
        const u32 brem = 2 + rnd() % (bsize - 1);
        o.rc_off = (u32)rc.size(); o.rc_len = nr * bsize + rrem;
        o.bc_off = (u32)bc.size(); o.bc_len = nb * bsize + brem;
        for (u32 i = 0; i < o.rc_len; i++) rc.push_back((u8)rnd());
        for (u32 i = 0; i < o.bc_len; i++) bc.push_back((u8)rnd());
        o.rev_off = (u32)rc_ev.size(); o.rev_len = nr;
        o.bev_off = (u32)bc_ev.size(); o.bev_len = nb;
        u32 t = 0;
        for (u32 i = 0; i < nr; i++) { t += 1 + rnd() % 50; rc_ev.push_back(2 * t + 1); }   // RC stamps odd, increasing
        t = 0;
        for (u32 i = 0; i < nb; i++) { t += 1 + rnd() % 50; bc_ev.push_back(2 * t); }       // BC stamps even, increasing
    }
    const i32 status = 0;
    FrameStream FS;
    memset(&FS, 0, sizeof(FS));
    FS.rc = rc.data(); FS.bc = bc.data(); FS.rc_ev = rc_ev.data(); FS.bc_ev = bc_ev.data(); FS.so = so.data();
    FS.nseg = nseg; FS.bsize = bsize; FS.with_props = 1; FS.status = &status;
    for (int i = 0; i < 10; i++) FS.props[i] = (u8)(i + 1);
    std::vector<u8> dev(frame_size(FS), 0xEE);
    frame_write(FS, dev.data());
    std::vector<u8> res(FS.props, FS.props + 10);
    Sink sink{&res};
    Framer fr(bsize);
    std::vector<u8> ev;
    for (u32 s = 0; s < nseg; s++) {
        merge_events(rc_ev.data() + so[s].rev_off, so[s].rev_len, bc_ev.data() + so[s].bev_off, so[s].bev_len, ev);
        fr.segment(sink, rc.data() + so[s].rc_off, so[s].rc_len, bc.data() + so[s].bc_off, so[s].bc_len, ev.data(), (u32)ev.size());
    }
    fr.eof(sink);
    if (dev.size() != res.size()) return 1;
    return memcmp(dev.data(), res.data(), res.size()) ? 2 : 0;
}

// Whole stream (with 10-byte props header) through the host-simulated decoder engine.
int hs_decode(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_n) {
    if (n < 10) return -1;
    CSCProps p;
    memset(&p, 0, sizeof(p));
    read_props(&p, in);
    DecInput di;
    int r = deframe(in + 10, n - 10, p.csc_blocksize, di);
    if (r < 0) return r;
    std::vector<u8> wnd(cap + 64);
    std::vector<u16> p_lit(65536, 2048), p_delta(65536, 2048);
    std::vector<DecRun> runs(cap / 512 + 1024);
    u32 n_runs = 0, out_len = 0;
    i32 status = 0;
    DecStream D;
    memset(&D, 0, sizeof(D));
    D.wnd_size = (u32)p.dict_size; D.raw_blocksize = p.raw_blocksize;
    D.rc = di.rc.data(); D.rc_len = (u32)di.rc.size(); D.bc = di.bc.data(); D.bc_len = (u32)di.bc.size();
    D.rc_blk_off = di.rc_off.data(); D.n_rc_blk = (u32)di.rc_off.size() - 1;
    D.bc_blk_off = di.bc_off.data(); D.n_bc_blk = (u32)di.bc_off.size() - 1;
    D.wnd = wnd.data(); D.out_cap = (u32)cap;
    D.p_lit = p_lit.data(); D.p_delta = p_delta.data();
    D.runs = runs.data(); D.runs_cap = (u32)runs.size(); D.n_runs = &n_runs; D.out_len = &out_len; D.status = &status;
    ModelSmall *M = new ModelSmall;
    Decoder *dec = new Decoder(D, *M);
    dec->run();
    delete dec;
    delete M;
    if (status < 0) return status;
    DictTables DT;
    build_dict_tables(DT);
    // inverse filters per run (csc_dec.cpp:596-651)
    for (u32 i = 0; i < n_runs; i++) {
        const DecRun &rn = runs[i];
        if (rn.type == DT_EXE) { memcpy(out + rn.off, wnd.data() + rn.off, rn.size); e8e9_run(out + rn.off, rn.size, true); }
        else if (rn.type == DT_ENGTXT) dict_inverse(wnd.data() + rn.off, rn.size, out + rn.off, DT);
        else memcpy(out + rn.off, wnd.data() + rn.off, rn.size);
    }
    *out_n = out_len;
    return 0;
}

// class-level helpers (compared with the oracle in tests)
void hs_analyze_blocks(const uint8_t *buf, size_t n, uint32_t *info, uint32_t *bpb) {
    AnalyzerTables AT;
    build_analyzer_tables(AT);
    u32 b = 0;
    u32 tail[5];
    for (size_t off = 0; off < n; off += kBlock, b++) {
        u32 sz = (u32)(n - off < kBlock ? n - off : kBlock);
        analyze_block(buf + off, sz, AT, &info[b], &bpb[b], tail);
    }
}
int hs_filter(int which, int inverse, uint8_t *buf, uint32_t n, uint32_t chn) {
    std::vector<u8> tmp(n + 16);
    int r = 1;
    if (which == 0) e8e9_run(buf, n, inverse != 0);
    else if (which == 1) {
        DictTables DT;
        build_dict_tables(DT);
        if (inverse) { dict_inverse(buf, n, tmp.data(), DT); memcpy(buf, tmp.data(), n); }
        else r = (int)dict_forward(buf, n, tmp.data(), DT);
    } else {
        if (inverse) delta_inverse(buf, n, chn, tmp.data()); else delta_forward(buf, n, chn, tmp.data());
        memcpy(buf, tmp.data(), n);
    }
    return r;
}
void hs_props_init(uint32_t dict_size, int level, uint64_t *o) {
    CSCProps p;
    memset(&p, 0, sizeof(p));
    props_init(&p, dict_size, level);
    o[0] = p.dict_size; o[1] = p.csc_blocksize; o[2] = p.raw_blocksize; o[3] = p.hash_bits; o[4] = p.hash_width;
    o[5] = p.bt_hash_bits; o[6] = p.bt_size; o[7] = p.bt_cyc; o[8] = p.good_len; o[9] = p.lz_mode;
    o[10] = p.DLTFilter; o[11] = p.TXTFilter; o[12] = p.EXEFilter; o[13] = est_mem_usage(&p); o[14] = sizeof(CSCProps);
}

}  // extern "C"
