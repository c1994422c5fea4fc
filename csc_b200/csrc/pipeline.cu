// pipeline.cu -- host side of the B200 libcsc hot path: batch layout in HBM, kernel
// sequencing on one CUDA stream, result framing, and the C-ABI entry points
// (include/csc_enc.h, include/csc_dec.h, include/csc_b200.h).
//
// There is NO CPU fallback: without a usable CUDA device every compute entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <thread>
#include <atomic>

#include "../../include/csc_enc.h"
#include "../../include/csc_dec.h"
#include "../../include/csc_b200.h"
#include "host_common.h"
#include "kernels.cuh"

using namespace cscb;

namespace {

const int kErrCuda = -92;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "csc_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return kErrCuda;                                                                       \
        }                                                                                          \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e != cudaSuccess) { fprintf(stderr, "csc_b200: cudaMalloc(%zu) failed: %s\n", bytes, cudaGetErrorString(e)); p = nullptr; return kErrCuda; }
        cap = want;
        return 0;
    }
    template <class T> T *as() const { return (T *)p; }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 4096;
        if (cudaMallocHost(&p, want) != cudaSuccess) { p = nullptr; return kErrCuda; }
        cap = want;
        return 0;
    }
    template <class T> T *as() const { return (T *)p; }
};

struct Timer {
    cudaEvent_t ev[16];
    int n = 0;
    bool ok = false;
    void init() { if (!ok) { for (auto &e : ev) cudaEventCreate(&e); ok = true; } }
    void mark(cudaStream_t st) { if (n < 16) cudaEventRecord(ev[n++], st); }
    float ms(int a, int b) { float t = 0; if (a < n && b < n) cudaEventElapsedTime(&t, ev[a], ev[b]); return t; }
};

struct Ctx {
    bool init = false;
    int dev = -1;
    bool busy = false;
    cudaStream_t st = nullptr;
    AnalyzerTables *d_AT = nullptr;
    DictTables *d_DT = nullptr;
    u32 *d_p2bits = nullptr;
    DevBuf packdesc, lz, tmp, mfs, P6, L6, Z2, Z6, D6, ht, plit, ap, rc, bc, ev, segout, status, blkdesc,
        segdesc, binfo, bbpb, taildlt, dictok, planscratch, runs, counters, streams, segsize, segblk0, sbits, sgoff, sK,
        dwnd, dout, doffs, druns, dmeta, doutptr, P3, L3, P2, L2, D23, nolz, rstiles, rsH, rsB, trace,
        flip, replan, subdesc, subscratch, rcev, bcev, dec, B6, E6, adl, fdesc, fsz, foff, swnd;
    PinBuf h_segout, h_status, h_rcev, h_bcev, h_rc, h_bc, h_meta, h_adler, h_fsz, h_replan;
    Timer tm;
    CSCB200_Stats stats;
};

// frees every cached device buffer of a context (the buffers only ever grow one by one; a wave whose mix of sizes
// differs from the last one's -- another level, another kind of call -- may need the room they hold)
void ctx_release_device(Ctx &C) {
    for (DevBuf *b = &C.packdesc; b != &C.swnd + 1; b++) { if (b->p) cudaFree(b->p); b->p = nullptr; b->cap = 0; }
}

// Device contexts (stream, cached device / pinned buffers) are pooled per GPU and leased for the duration of one
// API call: concurrent callers (csarc runs up to 8 worker threads, archiver/csarc.cpp:338-352) each get their own
// context, sequential callers reuse one, and nothing is tied to the life of a thread.
const int kMaxCtxPerDevice = 16;
struct Pool {
    std::mutex m;
    std::condition_variable cv;
    std::vector<Ctx *> all;
} g_pool;
thread_local int t_device = -1;                 // CSCB200_SetDevice: GPU of this thread's calls (-1: the current CUDA device)
thread_local std::vector<int> t_devices;        // CSCB200_SetDevices: GPUs a batched call of this thread fans out over
thread_local CSCB200_Stats t_stats;             // statistics of this thread's last call
std::atomic<unsigned> g_rr{0};                  // round-robin of handles over the device list

// debug: copy the intermediate arrays of a single-stream wave back to the host (tests only)
struct DebugDump { u32 *binfo, *bbpb; u8 *lz; u32 *P6; u16 *L6; u32 *Z2, *Z6; u8 *dict_ok; u32 *tail_dlt; bool on; };
thread_local DebugDump g_dump = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, false};

// debug: packet + match-candidate trace of the next single-stream encode (tests only; see Encoder::trace)
struct DebugTrace { u32 *buf; size_t cap; size_t *n_out; bool on; };
thread_local DebugTrace g_trace = {nullptr, 0, nullptr, false};

int device_count() {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) return 0;
    return ndev;
}

int ctx_init(Ctx &C, int dev) {
    if (C.init) return 0;
    CK(cudaSetDevice(dev));
    C.dev = dev;
    CK(cudaStreamCreateWithFlags(&C.st, cudaStreamNonBlocking));
    AnalyzerTables AT;
    build_analyzer_tables(AT);
    DictTables *DT = new DictTables;
    build_dict_tables(*DT);
    u32 p2[512];
    build_p2bits(p2);
    CK(cudaMalloc(&C.d_AT, sizeof(AT)));
    CK(cudaMalloc(&C.d_DT, sizeof(DictTables)));
    CK(cudaMalloc(&C.d_p2bits, sizeof(p2)));
    CK(cudaMemcpy(C.d_AT, &AT, sizeof(AT), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(C.d_DT, DT, sizeof(DictTables), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(C.d_p2bits, p2, sizeof(p2), cudaMemcpyHostToDevice));
    delete DT;
    C.tm.init();
    C.init = true;
    return 0;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// fn(i) for i in [0, n) on up to 16 host threads (framing / de-framing / staging copies are per-stream work)
template <class Fn> void parallel_for(size_t n, size_t grain, Fn fn) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned nt = (unsigned)std::min<size_t>(std::min(hw, 16u), (n + grain - 1) / std::max<size_t>(grain, 1));
    std::atomic<size_t> next{0};
    auto work = [&]() { for (;;) { const size_t i = next.fetch_add(1); if (i >= n) break; fn(i); } };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
}

// bounded memory sink for framed output
struct MemSink {
    u8 *p; size_t cap, off; bool overflow;
    void operator()(const u8 *d, u32 n) {
        if (off + n > cap) { overflow = true; off += n; return; }
        memcpy(p + off, d, n);
        off += n;
    }
};

// one stream to encode
struct EncJob {
    const u8 *in; u32 n;
    bool in_dev = false;       // `in` is a device pointer (kernel-only timing path)
    bool no_payload = false;   // leave RC/BC sub-streams in HBM, only fetch the per-segment records
    CSCProps props;
    std::vector<u32> segs;     // Read() sizes
    std::vector<u8> flip;      // per block: assumed outcome of the duplicate-block test (plan.h); refined on kErrReplan
    int status;
    // results (host, pinned staging): filled by encode_wave
    std::vector<SegOut> so;
    const u8 *rc, *bc;
    const u32 *rc_ev, *bc_ev;   // block-completion time stamps (SegOut::rev_off / bev_off index these)
    u32 adler = 1;              // Adler-32 of the raw input (when the wave is asked for it)
    // framing on the device (batched calls): the finished stream in pinned staging memory
    bool frame_dev = false; bool with_props = false;
    const u8 *framed = nullptr; size_t framed_size = 0;
};

// framing of the batched calls' streams on the device (frame.cuh); 0: host replay (host_common.h: Framer), A/B only
#ifndef CSC_FRAME_ON_DEVICE
#define CSC_FRAME_ON_DEVICE 1
#endif

// ----------------------------------------------------------------------------- Adler-32
// adler32() (archiver/csa_adler32.cpp:63-128) of n byte ranges of one device buffer, results into pinned host memory
// (valid after the next synchronisation of C.st).  init == nullptr: every checksum starts at 1.
int adler_ranges(Ctx &C, const u8 *d_base, const std::vector<u64> &off, const std::vector<u64> &len, const u32 *init, u32 **h_out) {
    const u32 n = (u32)off.size();
    std::vector<AdlerTile> tiles;
    std::vector<u32> first(n + 1);
    for (u32 i = 0; i < n; i++) {
        first[i] = (u32)tiles.size();
        for (u64 o = 0; o < len[i]; o += kAdlerTile) tiles.push_back({off[i] + o, (u32)std::min<u64>(kAdlerTile, len[i] - o), i});
    }
    first[n] = (u32)tiles.size();
    const size_t nt = tiles.size();
    // layout of C.adl: tiles | partA | partB | first | init | out
    const size_t o_pa = align_up(nt * sizeof(AdlerTile) + 16, 256), o_pb = o_pa + align_up(nt * 4 + 16, 256), o_first = o_pb + align_up(nt * 4 + 16, 256);
    const size_t o_init = o_first + align_up((size_t)(n + 1) * 4, 256), o_out = o_init + align_up((size_t)n * 4 + 16, 256);
    if (C.adl.ensure(o_out + (size_t)n * 4 + 256) || C.h_adler.ensure((size_t)n * 4 + 16)) return kErrCuda;
    u8 *b = C.adl.as<u8>();
    cudaStream_t st = C.st;
    if (nt) CK(cudaMemcpyAsync(b, tiles.data(), nt * sizeof(AdlerTile), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b + o_first, first.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st));
    if (init) CK(cudaMemcpyAsync(b + o_init, init, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    if (nt) k_adler32_tiles<<<(u32)nt, 256, 0, st>>>(d_base, (const AdlerTile *)b, (u32 *)(b + o_pa), (u32 *)(b + o_pb));
    k_adler32_combine<<<(n + 127) / 128, 128, 0, st>>>((const AdlerTile *)b, (const u32 *)(b + o_pa), (const u32 *)(b + o_pb), (const u32 *)(b + o_first), n,
                                                      init ? (const u32 *)(b + o_init) : nullptr, (u32 *)(b + o_out));
    CK(cudaMemcpyAsync(C.h_adler.p, b + o_out, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaGetLastError());
    *h_out = C.h_adler.as<u32>();
    return 0;
}

// ----------------------------------------------------------------------------- encode
// Runs the whole device pipeline for a set of streams that fits in HBM at once and leaves,
// per job, the RC/BC sub-streams + block events in pinned host memory.
int encode_wave(Ctx &C, std::vector<EncJob *> &jobs, bool want_adler = false) {
    const u32 ns = (u32)jobs.size();
    if (!ns) return 0;
    cudaStream_t st = C.st;
    // ---- layout
    std::vector<u64> goff(ns);
    std::vector<u32> seg0(ns), sbits(ns), sK(ns), swnd(ns), sblk0(ns), rcoff(ns), bcoff(ns), rccap(ns), evcap(ns), evoff(ns), deccap(ns);
    std::vector<u64> decoff(ns);
    u64 dectot = 0;
    std::vector<BlockDesc> blks;
    std::vector<SegDesc> segs;
    std::vector<u32> segsize, segblk0;
    u64 total = 0, rctot = 0, evtot = 0;
    u32 cpos = 0, K = 0, key_shift = 0, nsub = 0;
    bool bounded = false;   // -m5: bounded multi-context search rows
    for (u32 s = 0; s < ns; s++) {
        EncJob &J = *jobs[s];
        EncCfg cfg = cfg_from_props(&J.props);
        // (a stream longer than its window wraps: plan.h re-grids the sub-blocks; MatchFinder::normalize, csc_mf.cpp:108-114,
        // which only triggers beyond ~4 GiB per stream, is not modelled)
        if ((u64)J.n + cfg.vld_rge >= 0xFFFFFFF0ull || cfg.bt_bits || !cfg.ht_width) { J.status = kErrUnsupported; return kErrUnsupported; }
        if (s == 0) { K = cfg.ht_width; bounded = cfg.bounded != 0; }
        if (cfg.ht_width != K || (cfg.bounded != 0) != bounded) return kErrUnsupported;   // one level per wave
        goff[s] = total;
        seg0[s] = (u32)segs.size();
        sbits[s] = cfg.ht_bits; sK[s] = cfg.ht_width; swnd[s] = cfg.wnd_size; sblk0[s] = (u32)blks.size();
        key_shift = std::max(key_shift, cfg.ht_bits);
        u32 x = 0;
        u32 flt = (J.props.DLTFilter ? 1 : 0) | (J.props.TXTFilter ? 2 : 0) | (J.props.EXEFilter ? 4 : 0);
        for (u32 sz : J.segs) {
            SegDesc sd;
            sd.goff = total + x; sd.x = x; sd.size = sz; sd.stream = s; sd.blk0 = (u32)blks.size(); sd.flt = flt;
            sd.wnd = cfg.wnd_size; sd.sub0 = nsub; sd.nslots = (sz + kBlock - 1) / kBlock + wraps_inside(x, sz, cfg.wnd_size);
            sd.cpos = cpos; sd.seg = (u32)segs.size();
            nsub += sd.nslots;
            segsize.push_back(sz);
            segblk0.push_back(sd.blk0);
            for (u32 off = 0; off < sz; off += kBlock) {
                BlockDesc bd;
                bd.goff = total + x + off; bd.x = x + off; bd.size = std::min(kBlock, sz - off); bd.stream = s;
                bd.seg = (u32)segs.size(); bd.cpos = cpos; bd.flags = 0;
                cpos += bd.size;
                blks.push_back(bd);
            }
            segs.push_back(sd);
            x += sz;
        }
        total = align_up(total + J.n + 64, 256);
        rccap[s] = (u32)align_up((size_t)J.n + J.n / 8 + 66000, 16);
        rcoff[s] = (u32)rctot; bcoff[s] = (u32)rctot;
        rctot += rccap[s];
        evcap[s] = rccap[s] / 65536 + 8;
        evoff[s] = (u32)evtot;
        evtot += evcap[s];
        // decision list: at most 9 binary decisions per input byte (a literal), 8-entry alignment per segment
        const u64 dc = (u64)J.n * 9 + 64 * (u64)J.segs.size() + 1024;
        if (dc >= 0xFFFFFFF0ull) return kErrUnsupported;
        deccap[s] = (u32)dc;
        decoff[s] = dectot;
        dectot += (dc + 7) & ~7ull;
        if (rctot >= 0xFFFF0000ull || total >= 0xFFFF0000ull) return kErrUnsupported;
    }
    const u32 nblk = (u32)blks.size(), nseg = (u32)segs.size(), npos = cpos;
    // radix-sort tiles: never across streams; histogram laid out [stream][digit][tile-of-stream]
    std::vector<RsTile> rst;
    {
        u32 cstart = 0, hb = 0;
        for (u32 s = 0; s < ns; s++) {
            const u32 n_s = jobs[s]->n, nts = (n_s + kRsTile - 1) / kRsTile;
            for (u32 j = 0; j < nts; j++) {
                RsTile t;
                t.start = cstart + j * kRsTile; t.count = std::min(kRsTile, n_s - j * kRsTile); t.hbase = hb + j; t.nts = nts;
                t.stream = s; t.sstart = cstart;
                rst.push_back(t);
            }
            cstart += n_s; hb += 256 * nts;
        }
    }
    const u32 nrst = (u32)rst.size(), rsHn = 256 * nrst, rsNb = (rsHn + kScanChunk - 1) / kScanChunk;
    total += 256;
    C.stats.mf_positions += npos;
    // ---- device memory
    size_t sortq = 0;
    const size_t bmbytes = total / 8 + 64;
    auto alloc_all = [&]() -> int {
    int r = 0;
    r |= C.lz.ensure(total); r |= C.tmp.ensure(total);
    // One scratch area serves three consumers that never overlap in time: the radix sort's key / value ping-pong buffers
    // (match-finder stage), the candidate records T written by k_mf_tlist after the last sort and read by k_encode, and
    // the packed RC / BC sub-streams gathered for the download once k_rc is done.
    sortq = align_up((size_t)npos * 4 + 16, 256);
    const size_t mfs_bytes = std::max<size_t>({4 * sortq, total * sizeof(TRec), (size_t)2 * align_up(rctot + 64, 256)});
    r |= C.mfs.ensure(mfs_bytes);
    r |= C.P6.ensure(total * K * 4); r |= C.L6.ensure(total * K * 2);
    r |= C.Z2.ensure(bmbytes); r |= C.Z6.ensure(bmbytes); r |= C.D6.ensure(bmbytes); r |= C.D23.ensure(bmbytes);
    r |= C.P3.ensure(total * 4); r |= C.P2.ensure(total * 4); r |= C.L3.ensure(total * 2); r |= C.L2.ensure(total * 2);
    r |= C.B6.ensure(bmbytes); r |= C.E6.ensure(bmbytes);
    r |= C.nolz.ensure(nblk + 16); r |= C.flip.ensure(nblk + 16); r |= C.replan.ensure((size_t)ns * 8);
    r |= C.subdesc.ensure((size_t)nsub * sizeof(BlockDesc)); r |= C.subscratch.ensure((size_t)nsub * sizeof(SubBlock));
    r |= C.swnd.ensure((size_t)ns * 4);
    r |= C.ht.ensure((size_t)ns * 81920 * 4);
    r |= C.plit.ensure((size_t)ns * 131072 * 2);
    r |= C.ap.ensure((size_t)ns * (kApLimit + 2) * sizeof(ApUnit));
    r |= C.rc.ensure(rctot); r |= C.bc.ensure(rctot); r |= C.rcev.ensure(evtot * 4 + 16); r |= C.bcev.ensure(evtot * 4 + 16);
    r |= C.dec.ensure(dectot * sizeof(DecEntry) + 16);
    r |= C.segout.ensure((size_t)nseg * sizeof(SegOut)); r |= C.status.ensure((size_t)ns * 4);
    r |= C.blkdesc.ensure((size_t)nblk * sizeof(BlockDesc)); r |= C.segdesc.ensure((size_t)nseg * sizeof(SegDesc));
    r |= C.binfo.ensure((size_t)nblk * 4); r |= C.bbpb.ensure((size_t)nblk * 4);
    r |= C.taildlt.ensure((size_t)nseg * 5 * 4); r |= C.dictok.ensure(nblk);
    r |= C.planscratch.ensure((size_t)nblk * sizeof(FilterRun)); r |= C.runs.ensure((size_t)nblk * sizeof(GRun));
    r |= C.counters.ensure(64); r |= C.streams.ensure((size_t)ns * sizeof(EncStream));
    r |= C.segsize.ensure((size_t)nseg * 4); r |= C.segblk0.ensure((size_t)nseg * 4);
    r |= C.sbits.ensure((size_t)ns * 4); r |= C.sgoff.ensure((size_t)ns * 8); r |= C.sK.ensure((size_t)ns * 4);
    r |= (C.rstiles.ensure((size_t)nrst * sizeof(RsTile)) || C.rsH.ensure((size_t)rsHn * 4 + 64) || C.rsB.ensure((size_t)rsNb * 4 + 64)) ? 1 : 0;
    return r;
    };
    if (alloc_all()) {   // make room: drop what the context cached for earlier waves and try once more
        CK(cudaStreamSynchronize(st));
        ctx_release_device(C);
        if (alloc_all()) return kErrCuda;
    }
    RsKey *const keys0 = (RsKey *)C.mfs.as<u8>(), *const keys1 = (RsKey *)(C.mfs.as<u8>() + sortq);
    u32 *const vals0 = (u32 *)(C.mfs.as<u8>() + 2 * sortq), *const vals1 = (u32 *)(C.mfs.as<u8>() + 3 * sortq);
    u32 launches = 1;
    static bool smem_opt_in = false;
    if (!smem_opt_in) {
        CK(cudaFuncSetAttribute(rs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kRsTile * 8)));
        smem_opt_in = true;
    }
    // sorts (keys0,vals0) by the low `hbits` key bits inside every stream; returns the buffers holding the result
    auto sort_pairs = [&](u32 hbits, RsKey *&kres, u32 *&vres) -> int {
        RsKey *ka = keys0, *kb = keys1;
        u32 *va = vals0, *vb = vals1;
        for (u32 shift = 0; shift < hbits; shift += 8) {
            const u32 bits = std::min<u32>(8, hbits - shift), mask = (1u << bits) - 1;
            rs_hist<<<nrst, kRsThreads, 0, st>>>(ka, C.rstiles.as<RsTile>(), shift, mask, C.rsH.as<u32>());
            rs_scan_a<<<rsNb, kRsThreads, 0, st>>>(C.rsH.as<u32>(), rsHn, C.rsB.as<u32>());
            rs_scan_b<<<1, kRsThreads, 0, st>>>(C.rsB.as<u32>(), rsNb);
            rs_scan_c<<<rsNb, kRsThreads, 0, st>>>(C.rsH.as<u32>(), rsHn, C.rsB.as<u32>());
            rs_scatter<<<nrst, kRsThreads, kRsTile * 8, st>>>(ka, va, kb, vb, C.rstiles.as<RsTile>(), shift, mask, C.rsH.as<u32>());
            std::swap(ka, kb); std::swap(va, vb);
            launches += 5;
        }
        kres = ka; vres = va;
        return 0;
    };

    // ---- stream descriptors
    std::vector<EncStream> es(ns);
    for (u32 s = 0; s < ns; s++) {
        EncJob &J = *jobs[s];
        EncStream &S = es[s];
        memset(&S, 0, sizeof(S));
        S.cfg = cfg_from_props(&J.props);
        S.lz = C.lz.as<u8>() + goff[s]; S.n = J.n; S.tmp = C.tmp.as<u8>() + goff[s]; S.nseg = (u32)J.segs.size();
        S.seg_size = C.segsize.as<u32>() + seg0[s]; S.seg_blk0 = C.segblk0.as<u32>() + seg0[s];
        S.binfo = C.binfo.as<u32>(); S.bbpb = C.bbpb.as<u32>();
        S.tail_dlt = C.taildlt.as<u32>() + (size_t)seg0[s] * 5; S.dict_ok = C.dictok.as<u8>();
        S.P6 = C.P6.as<u32>() + goff[s] * K; S.L6 = C.L6.as<u16>() + goff[s] * K; S.K = K;
        S.Z2 = C.Z2.as<u32>() + goff[s] / 32; S.Z6 = C.Z6.as<u32>() + goff[s] / 32; S.D6 = C.D6.as<u32>() + goff[s] / 32;
        S.P3 = C.P3.as<u32>() + goff[s]; S.L3 = C.L3.as<u16>() + goff[s];
        S.P2 = C.P2.as<u32>() + goff[s]; S.L2 = C.L2.as<u16>() + goff[s];
        S.D23 = C.D23.as<u32>() + goff[s] / 32; S.blk_nolz = C.nolz.as<u8>();
        S.T = (J.n > S.cfg.wnd_size || bounded) ? nullptr : C.mfs.as<TRec>() + goff[s];
        S.B6 = C.B6.as<u32>() + goff[s] / 32; S.E6 = C.E6.as<u32>() + goff[s] / 32;
        S.flip = C.flip.as<u8>(); S.replan = C.replan.as<u32>() + 2 * (size_t)s;
        S.ht2 = C.ht.as<u32>() + (size_t)s * 81920; S.ht3 = S.ht2 + 16384;
        S.p_lit = C.plit.as<u16>() + (size_t)s * 131072; S.p_delta = S.p_lit + 65536;
        S.ap = C.ap.as<ApUnit>() + (size_t)s * (kApLimit + 2);
        S.rc_out = C.rc.as<u8>() + rcoff[s]; S.rc_cap = rccap[s];
        S.bc_out = C.bc.as<u8>() + bcoff[s]; S.bc_cap = rccap[s];
        S.dec = C.dec.as<DecEntry>() + decoff[s]; S.dec_cap = deccap[s];
        S.rc_ev = C.rcev.as<u32>() + evoff[s]; S.bc_ev = C.bcev.as<u32>() + evoff[s]; S.ev_cap = evcap[s];
        S.segout = C.segout.as<SegOut>() + seg0[s];
        S.status = C.status.as<i32>() + s;
    }
    const bool tracing = g_trace.on && ns == 1;
    if (tracing) {
        if (C.trace.ensure(g_trace.cap * 12 + 16)) return kErrCuda;
        es[0].trace = C.trace.as<u32>(); es[0].trace_cap = (u32)g_trace.cap; es[0].trace_n = C.counters.as<u32>() + 2;
    }
    Timer &T = C.tm;
    T.n = 0;
    T.mark(st);                                                                   // 0
    // ---- uploads
    for (u32 s = 0; s < ns; s++)
        if (jobs[s]->n)
            CK(cudaMemcpyAsync(C.lz.as<u8>() + goff[s], jobs[s]->in, jobs[s]->n,
                               jobs[s]->in_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(C.blkdesc.p, blks.data(), (size_t)nblk * sizeof(BlockDesc), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(C.segdesc.p, segs.data(), (size_t)nseg * sizeof(SegDesc), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(C.rstiles.p, rst.data(), (size_t)nrst * sizeof(RsTile), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(C.segsize.p, segsize.data(), (size_t)nseg * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(C.segblk0.p, segblk0.data(), (size_t)nseg * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(C.sbits.p, sbits.data(), (size_t)ns * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(C.sgoff.p, goff.data(), (size_t)ns * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(C.sK.p, sK.data(), (size_t)ns * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(C.swnd.p, swnd.data(), (size_t)ns * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(C.flip.p, 0, nblk, st));
    for (u32 s = 0; s < ns; s++)
        if (!jobs[s]->flip.empty())
            CK(cudaMemcpyAsync(C.flip.as<u8>() + sblk0[s], jobs[s]->flip.data(), jobs[s]->flip.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(C.streams.p, es.data(), (size_t)ns * sizeof(EncStream), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(C.Z2.p, 0, bmbytes, st));
    CK(cudaMemsetAsync(C.Z6.p, 0, bmbytes, st));
    CK(cudaMemsetAsync(C.D6.p, 0, bmbytes, st));
    CK(cudaMemsetAsync(C.D23.p, 0, bmbytes, st));
    CK(cudaMemsetAsync(C.B6.p, 0, bmbytes, st));
    CK(cudaMemsetAsync(C.E6.p, 0, bmbytes, st));
    CK(cudaMemsetAsync(C.P3.p, 0xFF, total * 4, st));
    CK(cudaMemsetAsync(C.P2.p, 0xFF, total * 4, st));
    CK(cudaMemsetAsync(C.ht.p, 0, (size_t)ns * 81920 * 4, st));
    CK(cudaMemsetAsync(C.status.p, 0, (size_t)ns * 4, st));
    CK(cudaMemsetAsync(C.counters.p, 0, 64, st));
    CK(cudaMemsetAsync(C.dictok.p, 0, nblk, st));
    CK(cudaMemsetAsync(C.taildlt.p, 0, (size_t)nseg * 5 * 4, st));
    CK(cudaMemsetAsync(C.segout.p, 0, (size_t)nseg * sizeof(SegOut), st));
    k_fill_u16<<<592, 256, 0, st>>>(C.plit.as<u16>(), (u64)ns * 131072, 2048);
    u32 *h_adl = nullptr;
    if (want_adler) {   // checksum of the raw chunks while they sit in HBM, before the filters rewrite them in place
        std::vector<u64> ao(goff.begin(), goff.end()), al(ns);
        for (u32 s = 0; s < ns; s++) al[s] = jobs[s]->n;
        int ar = adler_ranges(C, C.lz.as<u8>(), ao, al, nullptr, &h_adl);
        if (ar < 0) return ar;
        launches += 2;
    }
    T.mark(st);                                                                   // 1
    // ---- analyzer, static plan, filters
    k_analyze<<<(nblk + kAnWarps - 1) / kAnWarps, kAnWarps * 32, 0, st>>>(C.lz.as<u8>(), C.blkdesc.as<BlockDesc>(), nblk, C.d_AT,
                                                                           C.binfo.as<u32>(), C.bbpb.as<u32>(), C.taildlt.as<u32>());
    T.mark(st);                                                                   // 2
    u32 *d_nruns = C.counters.as<u32>();
    u32 *d_queue = C.counters.as<u32>() + 1;
    k_plan<<<(nseg + 127) / 128, 128, 0, st>>>(C.segdesc.as<SegDesc>(), nseg, C.binfo.as<u32>(), C.bbpb.as<u32>(), C.taildlt.as<u32>(),
                                               C.flip.as<u8>(), C.planscratch.as<FilterRun>(), C.subscratch.as<SubBlock>(),
                                               C.runs.as<GRun>(), d_nruns, C.nolz.as<u8>(), C.subdesc.as<BlockDesc>());
    k_filter_runs<<<std::min<u32>(nblk, 148 * 8), kFltThreads, 0, st>>>(C.lz.as<u8>(), C.tmp.as<u8>(), C.runs.as<GRun>(), d_nruns, C.d_DT, C.dictok.as<u8>());
    T.mark(st);                                                                   // 3
    launches += 3;
    // ---- match finder precomputation: 6-byte rows, then the 3-byte and 2-byte chains
    k_mf_hash<<<nsub, kMfThreads, 0, st>>>(C.lz.as<u8>(), C.subdesc.as<BlockDesc>(), C.sbits.as<u32>(), C.sgoff.as<u64>(), C.swnd.as<u32>(), key_shift,
                                           keys0, vals0, C.Z2.as<u32>(), C.Z6.as<u32>(), 6);
    T.mark(st);                                                                   // 4
    RsKey *ksorted = nullptr;
    u32 *vsorted = nullptr;
    sort_pairs(key_shift, ksorted, vsorted);
    T.mark(st);                                                                   // 5
    k_mf_rows<<<nrst, 256, 0, st>>>(ksorted, vsorted, C.rstiles.as<RsTile>(), C.sgoff.as<u64>(), C.sK.as<u32>(), C.P6.as<u32>(),
                                    bounded ? kBtKA : 0, bounded ? kBtRowK : 0, 0);
    if (bounded) {   // two more candidate lists by 8- and 12-byte context, then one distance-ordered row
        for (u32 which = kBtCtxB; which <= kBtCtxC; which += kBtCtxC - kBtCtxB) {
            k_mf_hash<<<nsub, kMfThreads, 0, st>>>(C.lz.as<u8>(), C.subdesc.as<BlockDesc>(), C.sbits.as<u32>(), C.sgoff.as<u64>(), C.swnd.as<u32>(), key_shift,
                                                   keys0, vals0, C.Z2.as<u32>(), C.Z6.as<u32>(), which);
            sort_pairs(key_shift, ksorted, vsorted);
            k_mf_rows<<<nrst, 256, 0, st>>>(ksorted, vsorted, C.rstiles.as<RsTile>(), C.sgoff.as<u64>(), C.sK.as<u32>(), C.P6.as<u32>(),
                                            which == kBtCtxB ? kBtKB : kBtKC, kBtRowK, which == kBtCtxB ? kBtKA : kBtKA + kBtKB);
        }
        k_mf_merge<<<nsub, 256, 0, st>>>(C.subdesc.as<BlockDesc>(), C.sgoff.as<u64>(), C.P6.as<u32>());
        launches += 5;
    }
    T.mark(st);                                                                   // 6
    if (K == 1) k_mf_extend<1><<<nsub, kMfThreads, 0, st>>>(C.lz.as<u8>(), C.subdesc.as<BlockDesc>(), C.sgoff.as<u64>(), C.P6.as<u32>(), C.L6.as<u16>());
    else if (K == 2) k_mf_extend<2><<<nsub, kMfThreads, 0, st>>>(C.lz.as<u8>(), C.subdesc.as<BlockDesc>(), C.sgoff.as<u64>(), C.P6.as<u32>(), C.L6.as<u16>());
    else if (K == 8) k_mf_extend<8><<<nsub, kMfThreads, 0, st>>>(C.lz.as<u8>(), C.subdesc.as<BlockDesc>(), C.sgoff.as<u64>(), C.P6.as<u32>(), C.L6.as<u16>());
    else if (K == 16) k_mf_extend<16><<<nsub, kMfThreads, 0, st>>>(C.lz.as<u8>(), C.subdesc.as<BlockDesc>(), C.sgoff.as<u64>(), C.P6.as<u32>(), C.L6.as<u16>());
    else return kErrUnsupported;
    T.mark(st);                                                                   // 7
    for (u32 which = 3; which >= 2; which--) {
        const u32 shift = which == 3 ? 16 : 14;
        u32 *P = which == 3 ? C.P3.as<u32>() : C.P2.as<u32>();
        u16 *L = which == 3 ? C.L3.as<u16>() : C.L2.as<u16>();
        k_mf_hash<<<nsub, kMfThreads, 0, st>>>(C.lz.as<u8>(), C.subdesc.as<BlockDesc>(), C.sbits.as<u32>(), C.sgoff.as<u64>(), C.swnd.as<u32>(), shift,
                                               keys0, vals0, C.Z2.as<u32>(), C.Z6.as<u32>(), which);
        sort_pairs(shift, ksorted, vsorted);
        k_mf_rows<<<nrst, 256, 0, st>>>(ksorted, vsorted, C.rstiles.as<RsTile>(), C.sgoff.as<u64>(), C.sK.as<u32>(), P, 1, 0, 0);
        k_mf_extend<1><<<nsub, kMfThreads, 0, st>>>(C.lz.as<u8>(), C.subdesc.as<BlockDesc>(), C.sgoff.as<u64>(), P, L);
    }
    k_mf_tlist<<<nsub, 256, 0, st>>>(C.subdesc.as<BlockDesc>(), C.streams.as<EncStream>());
    T.mark(st);                                                                   // 8
    launches += 3 + 1 + 2 * 3 + 1;
    if (g_dump.on && ns == 1) {
        CK(cudaStreamSynchronize(st));
        u32 n1 = jobs[0]->n;
        CK(cudaMemcpy(g_dump.binfo, C.binfo.p, (size_t)nblk * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(g_dump.bbpb, C.bbpb.p, (size_t)nblk * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(g_dump.lz, C.lz.p, n1, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(g_dump.P6, C.P6.p, (size_t)n1 * K * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(g_dump.L6, C.L6.p, (size_t)n1 * K * 2, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(g_dump.Z2, C.Z2.p, (size_t)(n1 + 31) / 32 * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(g_dump.Z6, C.Z6.p, (size_t)(n1 + 31) / 32 * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(g_dump.dict_ok, C.dictok.p, nblk, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(g_dump.tail_dlt, C.taildlt.p, (size_t)nseg * 5 * 4, cudaMemcpyDeviceToHost));
    }
    // ---- serial engines, one stream per warp
    int nsm = 148;
    struct TraceFetch {   // copies the debug trace back once the encoder has run (scope exit on every path)
        Ctx &C; bool on; cudaStream_t st;
        ~TraceFetch() {
            if (!on) return;
            u32 n = 0;
            cudaStreamSynchronize(st);
            cudaMemcpy(&n, C.counters.as<u32>() + 2, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(g_trace.buf, C.trace.p, std::min<size_t>(n, g_trace.cap) * 12, cudaMemcpyDeviceToHost);
            if (g_trace.n_out) *g_trace.n_out = n;
        }
    } trace_fetch{C, tracing, st};
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, C.dev);
    {
        u32 grid = std::min<u32>((ns + kEncWarps - 1) / kEncWarps, (u32)nsm * 8);
        k_encode<<<grid, kEncWarps * 32, 0, st>>>(C.streams.as<EncStream>(), ns, d_queue, C.d_p2bits);
    }
    T.mark(st);                                                                   // 9
    k_rc<<<(ns + kRcWarps - 1) / kRcWarps, kRcWarps * 32, 0, st>>>(C.streams.as<EncStream>(), ns);
    T.mark(st);                                                                   // 10
    launches += 2;
    CK(cudaGetLastError());
    bool frame_dev = true;
    for (u32 s = 0; s < ns; s++) frame_dev = frame_dev && jobs[s]->frame_dev && !jobs[s]->no_payload;
    if (frame_dev) {
        // ---- framing on the device (frame.cuh): sizes -> dense offsets -> the finished streams in ONE download
        std::vector<FrameStream> fs(ns);
        for (u32 s = 0; s < ns; s++) {
            FrameStream &F = fs[s];
            memset(&F, 0, sizeof(F));
            F.rc = C.rc.as<u8>() + rcoff[s]; F.bc = C.bc.as<u8>() + bcoff[s];
            F.rc_ev = C.rcev.as<u32>() + evoff[s]; F.bc_ev = C.bcev.as<u32>() + evoff[s];
            F.so = C.segout.as<SegOut>() + seg0[s]; F.nseg = (u32)jobs[s]->segs.size();
            F.bsize = jobs[s]->props.csc_blocksize; F.with_props = jobs[s]->with_props ? 1 : 0;
            write_props(&jobs[s]->props, F.props);
            F.status = C.status.as<i32>() + s;
        }
        if (C.fdesc.ensure((size_t)ns * sizeof(FrameStream)) || C.fsz.ensure((size_t)ns * 4 + 16) || C.foff.ensure((size_t)ns * 8 + 16) ||
            C.h_fsz.ensure((size_t)ns * 4 + 16) || C.h_status.ensure((size_t)ns * 4) || C.h_replan.ensure((size_t)ns * 8))
            return kErrCuda;
        CK(cudaMemcpyAsync(C.fdesc.p, fs.data(), (size_t)ns * sizeof(FrameStream), cudaMemcpyHostToDevice, st));
        k_frame_sizes<<<(ns + 127) / 128, 128, 0, st>>>(C.fdesc.as<FrameStream>(), ns, C.fsz.as<u32>());
        CK(cudaMemcpyAsync(C.h_fsz.p, C.fsz.p, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(C.h_status.p, C.status.p, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(C.h_replan.p, C.replan.p, (size_t)ns * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const u32 *hsz = C.h_fsz.as<u32>();
        const i32 *hst = C.h_status.as<i32>();
        if (h_adl) for (u32 s = 0; s < ns; s++) jobs[s]->adler = h_adl[s];
        std::vector<u64> foff(ns);
        u64 ftot = 0;
        for (u32 s = 0; s < ns; s++) { foff[s] = ftot; ftot += align_up(hsz[s], 16); }
        if (ftot + 64 > C.mfs.cap || C.h_rc.ensure(ftot + 64)) return kErrCuda;   // (the scratch area holds 2 x the sub-stream capacity)
        CK(cudaMemcpyAsync(C.foff.p, foff.data(), (size_t)ns * 8, cudaMemcpyHostToDevice, st));
        k_frame<<<(ns + 3) / 4, 128, 0, st>>>(C.fdesc.as<FrameStream>(), ns, C.foff.as<u64>(), C.mfs.as<u8>());
        if (ftot) CK(cudaMemcpyAsync(C.h_rc.p, C.mfs.p, ftot, cudaMemcpyDeviceToHost, st));
        launches += 2;
        T.mark(st);                                                               // 11
        CK(cudaStreamSynchronize(st));
        int ret = 0;
        for (u32 s = 0; s < ns; s++) {
            EncJob &J = *jobs[s];
            J.status = hst[s];
            if (J.status == kErrReplan) {
                const u32 *rp = C.h_replan.as<u32>() + 2 * (size_t)s;
                if (J.flip.empty()) J.flip.assign((s + 1 < ns ? sblk0[s + 1] : nblk) - sblk0[s], 0);
                if (rp[0] >= sblk0[s] && rp[0] - sblk0[s] < J.flip.size()) J.flip[rp[0] - sblk0[s]] = (u8)rp[1];
                else { J.status = kErrUnsupported; ret = kErrUnsupported; }
                C.stats.replans += 1;
                continue;
            }
            if (J.status < 0) { ret = J.status; continue; }
            J.framed = C.h_rc.as<u8>() + foff[s]; J.framed_size = hsz[s];
        }
        CSCB200_Stats &S = C.stats;
        S.ms_h2d += T.ms(0, 1); S.ms_analyze += T.ms(1, 2); S.ms_filters += T.ms(2, 3); S.ms_mf_hash += T.ms(3, 4);
        S.ms_mf_sort += T.ms(4, 5); S.ms_mf_rows += T.ms(5, 6); S.ms_mf_extend += T.ms(6, 7); S.ms_mf_chains += T.ms(7, 8);
        S.ms_encode += T.ms(8, 9); S.ms_rc += T.ms(9, 10); S.ms_d2h += T.ms(10, 11); S.ms_total_gpu += T.ms(0, 11);
        S.kernel_launches += launches; S.streams += ns; S.waves += 1;
        return ret;
    }
    // ---- results
    if (C.h_segout.ensure((size_t)nseg * sizeof(SegOut)) || C.h_status.ensure((size_t)ns * 4) || C.h_rcev.ensure(evtot * 4 + 16) ||
        C.h_bcev.ensure(evtot * 4 + 16) ||
        C.h_replan.ensure((size_t)ns * 8))
        return kErrCuda;
    CK(cudaMemcpyAsync(C.h_segout.p, C.segout.p, (size_t)nseg * sizeof(SegOut), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(C.h_status.p, C.status.p, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(C.h_replan.p, C.replan.p, (size_t)ns * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(C.h_rcev.p, C.rcev.p, evtot * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(C.h_bcev.p, C.bcev.p, evtot * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const SegOut *hso = C.h_segout.as<SegOut>();
    const i32 *hst = C.h_status.as<i32>();
    if (h_adl) for (u32 s = 0; s < ns; s++) jobs[s]->adler = h_adl[s];
    std::vector<u32> rcn(ns, 0), bcn(ns, 0), hrcoff(ns), hbcoff(ns);
    size_t hrc = 0, hbc = 0;
    int ret = 0;
    for (u32 s = 0; s < ns; s++) {
        EncJob &J = *jobs[s];
        J.status = hst[s];
        if (J.status == kErrReplan) {
            // the engine saw a duplicate-block outcome the static plan had not assumed (wrapped stream): record it;
            // the caller submits the stream again (prefix-stable, so every round fixes the earliest wrong guess)
            const u32 *rp = C.h_replan.as<u32>() + 2 * (size_t)s;
            if (J.flip.empty()) J.flip.assign((s + 1 < ns ? sblk0[s + 1] : nblk) - sblk0[s], 0);
            if (rp[0] >= sblk0[s] && rp[0] - sblk0[s] < J.flip.size()) J.flip[rp[0] - sblk0[s]] = (u8)rp[1];
            else { J.status = kErrUnsupported; ret = kErrUnsupported; }
            C.stats.replans += 1;
            continue;
        }
        if (J.status < 0) { ret = J.status; continue; }
        u32 nsg = (u32)J.segs.size();
        J.so.assign(hso + seg0[s], hso + seg0[s] + nsg);
        if (nsg) {
            rcn[s] = J.so[nsg - 1].rc_off + J.so[nsg - 1].rc_len;
            bcn[s] = J.so[nsg - 1].bc_off + J.so[nsg - 1].bc_len;
        }
        hrcoff[s] = (u32)hrc; hbcoff[s] = (u32)hbc;
        hrc += align_up(rcn[s], 16); hbc += align_up(bcn[s], 16);
    }
    if (C.h_rc.ensure(hrc + 16) || C.h_bc.ensure(hbc + 16)) return kErrCuda;
    bool any_payload = false;
    std::vector<PackDesc> pd(ns);
    for (u32 s = 0; s < ns; s++) {
        PackDesc &d = pd[s];
        memset(&d, 0, sizeof(d));
        if (jobs[s]->status < 0) continue;
        jobs[s]->rc_ev = C.h_rcev.as<u32>() + evoff[s];
        jobs[s]->bc_ev = C.h_bcev.as<u32>() + evoff[s];
        if (jobs[s]->no_payload) continue;
        any_payload = true;
        d.rc_src = rcoff[s]; d.rc_dst = hrcoff[s]; d.rc_n = rcn[s];
        d.bc_src = bcoff[s]; d.bc_dst = hbcoff[s]; d.bc_n = bcn[s];
        jobs[s]->rc = C.h_rc.as<u8>() + hrcoff[s];
        jobs[s]->bc = C.h_bc.as<u8>() + hbcoff[s];
    }
    if (any_payload) {   // gather on the device, then two dense downloads
        // the sort buffers are free again: 8 bytes per position each, the sub-streams need < 1.2
        u8 *const pk_rc = C.mfs.as<u8>(), *const pk_bc = C.mfs.as<u8>() + align_up(rctot + 64, 256);   // (hrc, hbc <= rctot + alignment)
        if (hrc + 16 > align_up(rctot + 64, 256) || hbc + 16 > align_up(rctot + 64, 256)) return kErrCuda;
        if (C.packdesc.ensure((size_t)ns * sizeof(PackDesc))) return kErrCuda;
        CK(cudaMemcpyAsync(C.packdesc.p, pd.data(), (size_t)ns * sizeof(PackDesc), cudaMemcpyHostToDevice, st));
        k_pack_results<<<dim3(ns, 4), 256, 0, st>>>(C.rc.as<u8>(), C.bc.as<u8>(), pk_rc, pk_bc, C.packdesc.as<PackDesc>());
        if (hrc) CK(cudaMemcpyAsync(C.h_rc.p, pk_rc, hrc, cudaMemcpyDeviceToHost, st));
        if (hbc) CK(cudaMemcpyAsync(C.h_bc.p, pk_bc, hbc, cudaMemcpyDeviceToHost, st));
        launches += 1;
    }
    T.mark(st);                                                                   // 11
    CK(cudaStreamSynchronize(st));
    CSCB200_Stats &S = C.stats;
    S.ms_h2d += T.ms(0, 1); S.ms_analyze += T.ms(1, 2); S.ms_filters += T.ms(2, 3); S.ms_mf_hash += T.ms(3, 4);
    S.ms_mf_sort += T.ms(4, 5); S.ms_mf_rows += T.ms(5, 6); S.ms_mf_extend += T.ms(6, 7); S.ms_mf_chains += T.ms(7, 8);
    S.ms_encode += T.ms(8, 9); S.ms_rc += T.ms(9, 10); S.ms_d2h += T.ms(10, 11); S.ms_total_gpu += T.ms(0, 11);
    S.kernel_launches += launches; S.streams += ns; S.waves += 1;
    return ret;
}

// frame one encoded job through `sink`
template <class Sink> void frame_segment(EncJob &J, size_t sg, Framer &fr, Sink &sink) {
    const SegOut &o = J.so[sg];
    std::vector<u8> ev;   // RC/BC block completions of the segment in temporal order (H4)
    merge_events(J.rc_ev + o.rev_off, o.rev_len, J.bc_ev + o.bev_off, o.bev_len, ev);
    fr.segment(sink, const_cast<u8 *>(J.rc) + o.rc_off, o.rc_len, J.bc + o.bc_off, o.bc_len, ev.data(), (u32)ev.size());
}
template <class Sink> void frame_job(EncJob &J, Framer &fr, Sink &sink) {
    for (size_t s = 0; s < J.so.size(); s++) frame_segment(J, s, fr, sink);
}

// size of the framed stream without touching the payload (device-resident path)
size_t framed_size(const EncJob &J, bool with_props) {
    const u32 bs = J.props.csc_blocksize;
    size_t t = with_props ? CSC_PROP_SIZE : 0;
    for (const SegOut &o : J.so) {
        const u32 nr = o.rev_len, nb = o.bev_len;
        u32 rrem = o.rc_len - nr * bs, brem = o.bc_len - nb * bs;
        t += (size_t)(nr + nb) * (1 + bs) + (rrem == bs ? 1 : 4) + rrem + (brem == bs ? 1 : 4) + brem;
    }
    return t + 4 + 6 + 4 + 3;   // EOF segment
}

size_t per_stream_bytes() { return 81920 * 4 + 131072 * 2 + (kApLimit + 2) * sizeof(ApUnit) + 140000 * 2 + 4096; }
// lz + tmp, scratch (sort buffers / records), rows, 2-/3-byte chains, six bitmaps, RC + BC output (n + n/8 each), decisions
size_t per_pos_bytes(u32 K) { return 2 + sizeof(TRec) + 6 * (size_t)K + 12 + 1 + 3 + 9 * sizeof(DecEntry); }

// ----------------------------------------------------------------------------- decode
struct DecJob {
    const u8 *in; size_t n;        // stream incl. 10-byte props
    u8 *out; size_t out_cap;
    size_t out_size;
    int status;
    CSCProps props;
    DecInput di;
    bool predeframed = false;      // di was filled block by block by the caller (handle API)
    std::vector<DecRun> runs;      // filled when want_runs
    u32 adler = 1;                 // Adler-32 of the decoded bytes (when the wave is asked for it)
};

int decode_wave(Ctx &C, std::vector<DecJob *> &jobs, bool want_runs, bool want_adler = false) {
    const u32 ns = (u32)jobs.size();
    if (!ns) return 0;
    cudaStream_t st = C.st;
    std::vector<size_t> woff(ns), rcoff(ns), bcoff(ns), offoff(ns), runoff(ns);
    std::vector<u32> runcap(ns);
    size_t wtot = 0, rctot = 0, offtot = 0, runtot = 0;
    auto t0 = std::chrono::steady_clock::now();
    std::atomic<int> bad{0};
    parallel_for(ns, 8, [&](size_t s) {   // de-framing is independent per stream
        DecJob &J = *jobs[s];
        J.status = 0; J.out_size = 0;
        if (!J.predeframed) {
            if (J.n < CSC_PROP_SIZE) { J.status = -1; bad = 1; return; }
            memset(&J.props, 0, sizeof(J.props));
            read_props(&J.props, J.in);
        }
        if (J.props.dict_size > 1024u * 1048576u || J.props.dict_size < 32 * 1024 || !J.props.csc_blocksize) { J.status = -1; bad = 1; return; }
        if (!J.predeframed && deframe(J.in + CSC_PROP_SIZE, J.n - CSC_PROP_SIZE, J.props.csc_blocksize, J.di) < 0) { J.status = -1; bad = 1; return; }
        if (J.di.rc_off.size() < 2 || J.di.bc_off.size() < 2) { J.status = -1; bad = 1; }
    });
    if (bad) return -1;
    for (u32 s = 0; s < ns; s++) {
        DecJob &J = *jobs[s];
        woff[s] = wtot; wtot = align_up(wtot + J.out_cap + 64, 256);
        rcoff[s] = rctot; rctot = align_up(rctot + J.di.rc.size(), 16);
        bcoff[s] = rctot; rctot = align_up(rctot + J.di.bc.size(), 16);
        offoff[s] = offtot; offtot += J.di.rc_off.size() + J.di.bc_off.size();
        runcap[s] = (u32)(J.out_cap / 4096 + 64);
        runoff[s] = runtot; runtot += runcap[s];
    }
    double ms_deframe = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    auto alloc_all = [&]() -> int {
    int r = 0;
    r |= C.dwnd.ensure(wtot + 256); r |= C.dout.ensure(wtot + 256); r |= C.rc.ensure(rctot + 64);
    r |= C.doffs.ensure(offtot * 4 + 16); r |= C.druns.ensure(runtot * sizeof(DecRun));
    r |= C.plit.ensure((size_t)ns * 131072 * 2); r |= C.dmeta.ensure((size_t)ns * 16);
    r |= C.streams.ensure((size_t)ns * sizeof(DecStream)); r |= C.counters.ensure(64); r |= C.doutptr.ensure((size_t)ns * 8);
    return r;
    };
    if (alloc_all()) {   // make room: drop what the context cached for earlier waves and try once more
        CK(cudaStreamSynchronize(st));
        ctx_release_device(C);
        if (alloc_all()) return kErrCuda;
    }
    std::vector<DecStream> ds(ns);
    std::vector<u8 *> outp(ns);
    std::vector<u32> offs(offtot);
    for (u32 s = 0; s < ns; s++) {
        DecJob &J = *jobs[s];
        DecStream &D = ds[s];
        memset(&D, 0, sizeof(D));
        D.wnd_size = (u32)J.props.dict_size; D.raw_blocksize = J.props.raw_blocksize;
        D.rc = C.rc.as<u8>() + rcoff[s]; D.rc_len = (u32)J.di.rc.size() - 16;
        D.bc = C.rc.as<u8>() + bcoff[s]; D.bc_len = (u32)J.di.bc.size() - 16;
        memcpy(&offs[offoff[s]], J.di.rc_off.data(), J.di.rc_off.size() * 4);
        memcpy(&offs[offoff[s] + J.di.rc_off.size()], J.di.bc_off.data(), J.di.bc_off.size() * 4);
        D.rc_blk_off = C.doffs.as<u32>() + offoff[s]; D.n_rc_blk = (u32)J.di.rc_off.size() - 1;
        D.bc_blk_off = D.rc_blk_off + J.di.rc_off.size(); D.n_bc_blk = (u32)J.di.bc_off.size() - 1;
        D.wnd = C.dwnd.as<u8>() + woff[s]; D.out_cap = (u32)J.out_cap;
        D.p_lit = C.plit.as<u16>() + (size_t)s * 131072; D.p_delta = D.p_lit + 65536;
        D.runs = C.druns.as<DecRun>() + runoff[s]; D.runs_cap = runcap[s];
        u32 *meta = C.dmeta.as<u32>() + (size_t)s * 4;
        D.n_runs = meta; D.out_len = meta + 1; D.status = (i32 *)(meta + 2);
        outp[s] = C.dout.as<u8>() + woff[s];
    }
    Timer &T = C.tm;
    T.n = 0;
    T.mark(st);                                                                   // 0
    // the sub-streams of all jobs go through one pinned staging buffer: one upload instead of two per stream
    if (C.h_rc.ensure(rctot + 64)) return kErrCuda;
    parallel_for(ns, 8, [&](size_t s) {
        DecJob &J = *jobs[s];
        memcpy(C.h_rc.as<u8>() + rcoff[s], J.di.rc.data(), J.di.rc.size());
        memcpy(C.h_rc.as<u8>() + bcoff[s], J.di.bc.data(), J.di.bc.size());
    });
    CK(cudaMemcpyAsync(C.rc.p, C.h_rc.p, rctot, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(C.doffs.p, offs.data(), offtot * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(C.streams.p, ds.data(), (size_t)ns * sizeof(DecStream), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(C.doutptr.p, outp.data(), (size_t)ns * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(C.dmeta.p, 0, (size_t)ns * 16, st));
    CK(cudaMemsetAsync(C.counters.p, 0, 64, st));
    k_fill_u16<<<592, 256, 0, st>>>(C.plit.as<u16>(), (u64)ns * 131072, 2048);
    T.mark(st);                                                                   // 1
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, C.dev);
    u32 grid = std::min<u32>((ns + kDecWarps - 1) / kDecWarps, (u32)nsm * 16);   // up to 64 resident warps per SM pull streams from the queue
    k_decode<<<grid, kDecWarps * 32, 0, st>>>(C.streams.as<DecStream>(), ns, C.counters.as<u32>() + 1);
    T.mark(st);                                                                   // 2
    CK(cudaMemcpyAsync(C.dout.p, C.dwnd.p, wtot, cudaMemcpyDeviceToDevice, st));
    k_unfilter<<<ns, kFltThreads, 0, st>>>(C.streams.as<DecStream>(), ns, C.doutptr.as<u8 *>(), C.d_DT);
    T.mark(st);                                                                   // 3
    CK(cudaGetLastError());
    if (C.h_meta.ensure((size_t)ns * 16)) return kErrCuda;
    CK(cudaMemcpyAsync(C.h_meta.p, C.dmeta.p, (size_t)ns * 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const u32 *hm = C.h_meta.as<u32>();
    int ret = 0;
    for (u32 s = 0; s < ns; s++) {
        DecJob &J = *jobs[s];
        J.status = (i32)hm[s * 4 + 2];
        if (J.status < 0) { ret = J.status; if (!want_runs) continue; }   // (the handle API still writes the runs decoded before the error)
        J.out_size = hm[s * 4 + 1];
        if (J.out_size) CK(cudaMemcpyAsync(J.out, C.dout.as<u8>() + woff[s], J.out_size, cudaMemcpyDeviceToHost, st));
        if (want_runs) {
            J.runs.resize(hm[s * 4]);
            if (!J.runs.empty())
                CK(cudaMemcpyAsync(J.runs.data(), C.druns.as<DecRun>() + runoff[s], J.runs.size() * sizeof(DecRun), cudaMemcpyDeviceToHost, st));
        }
    }
    u32 *h_adl = nullptr;
    if (want_adler) {   // checksum of the decoded chunks on the device, next to their download
        std::vector<u64> ao(ns), al(ns);
        for (u32 s = 0; s < ns; s++) { ao[s] = woff[s]; al[s] = jobs[s]->status < 0 ? 0 : jobs[s]->out_size; }
        int ar = adler_ranges(C, C.dout.as<u8>(), ao, al, nullptr, &h_adl);
        if (ar < 0) return ar;
    }
    T.mark(st);                                                                   // 4
    CK(cudaStreamSynchronize(st));
    if (h_adl) for (u32 s = 0; s < ns; s++) jobs[s]->adler = h_adl[s];
    CSCB200_Stats &S = C.stats;
    S.ms_h2d += T.ms(0, 1); S.ms_decode += T.ms(1, 2); S.ms_unfilter += T.ms(2, 3); S.ms_d2h += T.ms(3, 4);
    S.ms_total_gpu += T.ms(0, 4); S.ms_frame += ms_deframe;
    S.kernel_launches += 3; S.streams += ns; S.waves += 1;
    return ret;
}

// ----------------------------------------------------------------------------- contexts
// memory a wave may use on the leased context's device: what is free now plus what this context already holds
size_t mem_budget(const Ctx &C) {
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return (size_t)8 << 30;
    size_t held = 0;
    const DevBuf *b = &C.packdesc;
    const DevBuf *e = &C.swnd + 1;   // the DevBuf members are declared contiguously
    for (; b != e; b++) held += b->cap;
    const size_t avail = fr + held;
    return avail - avail / 8;
}

int pick_device() {   // device of this thread's next call, or a negative error
    const int n = device_count();
    if (n <= 0) {
        fprintf(stderr, "csc_b200: no CUDA device available (this library has no CPU fallback)\n");
        return kErrCuda;
    }
    int dev = t_device;
    if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return kErrCuda;
    return dev < n ? dev : kErrCuda;
}

// devices a batched call of this thread fans out over: CSCB200_SetDevices, else the environment
// (CSC_B200_DEVICES=all | 0,2,5 -- how an unmodified caller such as csarc is pointed at several GPUs), else one
std::vector<int> device_list() {
    std::vector<int> d = t_devices;
    if (d.empty() && t_device < 0) {
        const char *e = getenv("CSC_B200_DEVICES");
        if (e && *e) {
            const int n = device_count();
            if (!strcmp(e, "all")) for (int i = 0; i < n; i++) d.push_back(i);
            else for (const char *q = e; *q;) { char *end; long v = strtol(q, &end, 10); if (end == q) break; if (v >= 0 && v < n) d.push_back((int)v); q = *end ? end + 1 : end; }
        }
    }
    if (d.empty()) { const int one = pick_device(); if (one >= 0) d.push_back(one); }
    return d;
}

struct Lease {
    Ctx *c = nullptr;
    int prev = -1;
    explicit Lease(int dev) {
        if (dev < 0) return;
        cudaGetDevice(&prev);
        std::unique_lock<std::mutex> lk(g_pool.m);
        for (;;) {
            int have = 0;
            for (Ctx *x : g_pool.all) {
                if (x->dev != dev) continue;
                have++;
                if (!x->busy) { c = x; break; }
            }
            if (c) break;
            if (have < kMaxCtxPerDevice) { c = new Ctx; c->dev = dev; g_pool.all.push_back(c); break; }
            g_pool.cv.wait(lk);
        }
        c->busy = true;
        lk.unlock();
        if (ctx_init(*c, dev) < 0) { release(); return; }
        cudaSetDevice(dev);   // every entry point runs on its context's device, whatever the caller selected meanwhile
        memset(&c->stats, 0, sizeof(c->stats));
    }
    void release() {
        if (!c) return;
        { std::lock_guard<std::mutex> lk(g_pool.m); c->busy = false; }
        g_pool.cv.notify_one();
        c = nullptr;
    }
    ~Lease() {
        if (c) t_stats = c->stats;
        release();
        if (prev >= 0) cudaSetDevice(prev);
    }
    Lease(const Lease &) = delete;
    Lease &operator=(const Lease &) = delete;
};

void add_stats(CSCB200_Stats &a, const CSCB200_Stats &b) {   // devices work side by side: times are the slowest device's
    double *x = &a.ms_h2d;
    const double *y = &b.ms_h2d;
    for (int i = 0; i < 15; i++) x[i] = std::max(x[i], y[i]);
    a.bytes_in += b.bytes_in; a.bytes_out += b.bytes_out; a.mf_positions += b.mf_positions;
    a.kernel_launches += b.kernel_launches; a.streams += b.streams; a.waves += b.waves; a.replans += b.replans;
}

// chunks [lo, hi) of n for device k of g: contiguous ranges, chunk c belongs to device c * g / n
void chunk_range(int n, int k, int g, int &lo, int &hi) {
    lo = (int)(((long long)n * k + g - 1) / g);
    hi = (int)(((long long)n * (k + 1) + g - 1) / g);
}

// ----------------------------------------------------------------------------- batched encode on one device
int encode_chunks_on(Ctx &C, int n, const uint8_t *const *in, const size_t *in_size, uint32_t dict_size, int level, int flt_mask,
                     int with_props, uint8_t *const *out, const size_t *out_cap, size_t *out_size, bool device, uint32_t *adler = nullptr) {
    if (n <= 0) return 0;
    std::vector<EncJob> jobs(n);
    for (int i = 0; i < n; i++) {
        EncJob &J = jobs[i];
        if (in_size[i] >= 0xF0000000ull) return kErrUnsupported;
        J.in = in[i]; J.n = (u32)in_size[i]; J.status = 0; J.rc = J.bc = nullptr; J.rc_ev = J.bc_ev = nullptr;
        J.in_dev = device; J.no_payload = device;
        J.frame_dev = !device && CSC_FRAME_ON_DEVICE; J.with_props = with_props != 0;
        uint32_t d = (uint64_t)dict_size > in_size[i] ? (uint32_t)in_size[i] : dict_size;   // csa_worker.cpp:35
        props_init(&J.props, d, level);
        J.props.DLTFilter = flt_mask & 1; J.props.TXTFilter = (flt_mask >> 1) & 1; J.props.EXEFilter = (flt_mask >> 2) & 1;
        for (u32 a = 0; a < J.n; a += J.props.raw_blocksize) J.segs.push_back(std::min<u32>(J.props.raw_blocksize, J.n - a));
        C.stats.bytes_in += J.n;
    }
    // waves: as many streams as fit the memory budget (largest first inside a wave for balance)
    const size_t budget = mem_budget(C);
    const u32 K = cfg_from_props(&jobs[0].props).ht_width;   // (-m5: width of the bounded-search rows)
    int ret = 0;
    size_t at = 0;
    std::vector<EncJob *> again;   // streams whose static plan has to be redone (kErrReplan)
    while (at < jobs.size() || !again.empty()) {
        std::vector<EncJob *> wave;
        if (!again.empty()) wave.swap(again);
        const bool fresh = wave.empty();
        size_t bytes = 0, pos = 0;
        while (fresh && at < jobs.size()) {
            EncJob &J = jobs[at];
            size_t need = per_stream_bytes() + (size_t)J.n * per_pos_bytes(K) + 4096;
            if (!wave.empty() && (bytes + need > budget || pos + J.n > 0x70000000ull)) break;
            bytes += need; pos += J.n;
            if (J.n) wave.push_back(&J);
            at++;
        }
        std::stable_sort(wave.begin(), wave.end(), [](const EncJob *a, const EncJob *b) { return a->n > b->n; });
        int r = encode_wave(C, wave, adler != nullptr);
        if (r < 0) { ret = r; break; }
        if (adler) for (EncJob *J : wave) adler[J - jobs.data()] = J->adler;
        auto t0 = std::chrono::steady_clock::now();
        // host framing (MemIO replay) of the wave's streams: independent per stream, so spread over host threads
        std::atomic<size_t> next{0}, bytes_out{0};
        std::atomic<int> overflow{0};
        auto work = [&]() {
            for (;;) {
                const size_t w = next.fetch_add(1);
                if (w >= wave.size()) break;
                EncJob *J = wave[w];
                size_t i = J - jobs.data();
                if (J->status == kErrReplan) continue;   // goes round again below
                if (device) { out_size[i] = framed_size(*J, with_props != 0); bytes_out += out_size[i]; continue; }
                if (J->framed) {   // framed on the device: the stream is final, copy it out
                    out_size[i] = J->framed_size;
                    if (J->framed_size > out_cap[i]) { overflow = 1; continue; }
                    memcpy(out[i], J->framed, J->framed_size);
                    bytes_out += J->framed_size;
                    continue;
                }
                MemSink sink{out[i], out_cap[i], 0, false};
                if (with_props) { u8 hdr[CSC_PROP_SIZE]; write_props(&J->props, hdr); sink(hdr, CSC_PROP_SIZE); }
                Framer fr(J->props.csc_blocksize);
                frame_job(*J, fr, sink);
                fr.eof(sink);
                out_size[i] = sink.off;
                if (sink.overflow) overflow = 1;
                bytes_out += sink.off;
            }
        };
        {
            const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
            const unsigned nt = (unsigned)std::min<size_t>(std::min(hw, 16u), (wave.size() + 7) / 8);
            std::vector<std::thread> pool;
            for (unsigned t = 1; t < nt; t++) pool.emplace_back(work);
            work();
            for (auto &th : pool) th.join();
        }
        if (overflow) ret = kErrOverflow;
        for (EncJob *J : wave) if (J->status == kErrReplan) again.push_back(J);
        C.stats.bytes_out += bytes_out;
        C.stats.ms_frame += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    if (ret < 0) return ret;
    for (int i = 0; i < n; i++) {
        if (jobs[i].n) continue;   // empty chunk: header + EOF segment only
        if (adler) adler[i] = 1;
        if (device) { out_size[i] = (with_props ? CSC_PROP_SIZE : 0) + 17; continue; }
        MemSink sink{out[i], out_cap[i], 0, false};
        if (with_props) { u8 hdr[CSC_PROP_SIZE]; write_props(&jobs[i].props, hdr); sink(hdr, CSC_PROP_SIZE); }
        Framer fr(jobs[i].props.csc_blocksize);
        fr.eof(sink);
        out_size[i] = sink.off;
        if (sink.overflow) ret = kErrOverflow;
    }
    return ret;
}

int decode_chunks_on(Ctx &C, int n, const uint8_t *const *in, const size_t *in_size, uint8_t *const *out, const size_t *out_cap,
                     size_t *out_size, uint32_t *adler = nullptr) {
    if (n <= 0) return 0;
    std::vector<DecJob> jobs(n);
    for (int i = 0; i < n; i++) {
        if (out_cap[i] >= 0xFFFF0000ull) return kErrUnsupported;
        jobs[i].in = in[i]; jobs[i].n = in_size[i]; jobs[i].out = out[i]; jobs[i].out_cap = out_cap[i];
        C.stats.bytes_in += in_size[i];
    }
    const size_t budget = mem_budget(C);
    int ret = 0;
    size_t at = 0;
    while (at < jobs.size()) {
        std::vector<DecJob *> wave;
        size_t bytes = 0;
        while (at < jobs.size()) {
            size_t need = jobs[at].out_cap * 2 + jobs[at].n + 131072 * 2 + jobs[at].out_cap / 256 + 8192;
            if (!wave.empty() && bytes + need > budget) break;
            bytes += need;
            wave.push_back(&jobs[at]);
            at++;
        }
        int r = decode_wave(C, wave, false, adler != nullptr);
        if (r < 0) { ret = r; break; }
    }
    for (int i = 0; i < n; i++) { out_size[i] = jobs[i].out_size; C.stats.bytes_out += jobs[i].out_size; if (adler) adler[i] = jobs[i].adler; }
    return ret;
}

// Runs fn(context, lo, hi) for the chunk range of every device of this thread's device list, side by side
// (one host thread and one leased context per GPU; no data-path collective, results are already in chunk order).
template <class Fn> int fan_out(int n, bool may_split, Fn fn) {
    std::vector<int> devs = device_list();
    if (devs.empty()) return kErrCuda;
    if (!may_split || n < 2) devs.resize(1);
    if ((int)devs.size() > n) devs.resize(std::max(n, 1));
    const int g = (int)devs.size();
    if (g == 1) {
        Lease L(devs[0]);
        if (!L.c) return kErrCuda;
        return fn(*L.c, 0, n);
    }
    std::vector<int> rets(g, 0);
    std::vector<CSCB200_Stats> st(g);
    std::vector<std::thread> th;
    for (int k = 0; k < g; k++)
        th.emplace_back([&, k]() {
            int lo, hi;
            chunk_range(n, k, g, lo, hi);
            Lease L(devs[k]);
            if (!L.c) { rets[k] = kErrCuda; return; }
            rets[k] = fn(*L.c, lo, hi);
            st[k] = L.c->stats;
        });
    for (auto &t : th) t.join();
    memset(&t_stats, 0, sizeof(t_stats));
    int ret = 0;
    for (int k = 0; k < g; k++) { add_stats(t_stats, st[k]); if (rets[k] < 0 && !ret) ret = rets[k]; }
    return ret;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

// tests only: capture the intermediate arrays of the next single-stream encode
void CSCB200_DebugDump(uint32_t *binfo, uint32_t *bbpb, uint8_t *lz, uint32_t *P6, uint16_t *L6, uint32_t *Z2, uint32_t *Z6,
                       uint8_t *dict_ok, uint32_t *tail_dlt) {
    g_dump.binfo = binfo; g_dump.bbpb = bbpb; g_dump.lz = lz; g_dump.P6 = P6; g_dump.L6 = L6; g_dump.Z2 = Z2; g_dump.Z6 = Z6;
    g_dump.dict_ok = dict_ok; g_dump.tail_dlt = tail_dlt; g_dump.on = binfo != nullptr;
}

// tests only: record the packet / match-candidate trace of the next single-stream encode as (kind, a, b) triples
void CSCB200_DebugTrace(uint32_t *buf, size_t cap_triples, size_t *n_out) {
    g_trace.buf = buf; g_trace.cap = cap_triples; g_trace.n_out = n_out; g_trace.on = buf != nullptr;
}

int CSCB200_Available(void) { return device_count() > 0 ? 1 : 0; }
// the contiguous chunk range [lo, hi) of part k of g (chunk c belongs to part c * g / n): the rule the batched
// calls fan out by, exported so that a one-process-per-GPU launcher (bench.py under torchrun) partitions identically
void CSCB200_ChunkRange(int n_chunks, int k, int g, int *lo, int *hi) { chunk_range(n_chunks, k, g, *lo, *hi); }
int CSCB200_DeviceCount(void) { return device_count(); }

int CSCB200_SetDevice(int device) {
    if (device >= device_count()) return kErrUnsupported;   // (a failed call leaves the selection untouched)
    t_device = device < 0 ? -1 : device;
    t_devices.clear();
    return 0;
}
int CSCB200_SetDevices(const int *devices, int n) {
    const int have = device_count();
    for (int i = 0; i < n; i++) if (devices[i] < 0 || devices[i] >= have) return kErrUnsupported;
    t_devices.assign(devices, devices + std::max(n, 0));
    return 0;
}
void CSCB200_GetStats(CSCB200_Stats *out) { *out = t_stats; }

// release the cached device / pinned memory of every idle context (a long-lived host process calls this between jobs)
void CSCB200_Trim(void) {
    std::lock_guard<std::mutex> lk(g_pool.m);
    int prev = -1;
    cudaGetDevice(&prev);
    for (Ctx *c : g_pool.all) {
        if (c->busy || !c->init) continue;
        cudaSetDevice(c->dev);
        for (DevBuf *b = &c->packdesc; b != &c->swnd + 1; b++) { if (b->p) cudaFree(b->p); b->p = nullptr; b->cap = 0; }
        for (PinBuf *b = &c->h_segout; b != &c->h_replan + 1; b++) { if (b->p) cudaFreeHost(b->p); b->p = nullptr; b->cap = 0; }
    }
    if (prev >= 0) cudaSetDevice(prev);
}

size_t CSCB200_EncodeBound(size_t n) { return n + n / 4 + (n / 65536 + 8) * 8 + 256; }

int CSCB200_EncodeChunks(int n, const uint8_t *const *in, const size_t *in_size, uint32_t dict_size, int level, int flt_mask,
                         int with_props, uint8_t *const *out, const size_t *out_cap, size_t *out_size) {
    return fan_out(n, true, [&](Ctx &C, int lo, int hi) {
        return encode_chunks_on(C, hi - lo, in + lo, in_size + lo, dict_size, level, flt_mask, with_props, out + lo, out_cap + lo, out_size + lo, false);
    });
}

int CSCB200_EncodeChunksDevice(int n, const uint8_t *const *d_in, const size_t *in_size, uint32_t dict_size, int level,
                               int flt_mask, int with_props, size_t *out_size) {
    return fan_out(n, false, [&](Ctx &C, int lo, int hi) {   // the inputs live on one device: no fan-out
        return encode_chunks_on(C, hi - lo, d_in + lo, in_size + lo, dict_size, level, flt_mask, with_props, nullptr, nullptr, out_size + lo, true);
    });
}

int CSCB200_DecodeChunks(int n, const uint8_t *const *in, const size_t *in_size, uint8_t *const *out, const size_t *out_cap,
                         size_t *out_size) {
    return fan_out(n, true, [&](Ctx &C, int lo, int hi) {
        return decode_chunks_on(C, hi - lo, in + lo, in_size + lo, out + lo, out_cap + lo, out_size + lo);
    });
}

// The same two calls with csarc's per-file checksum as a by-product: adler32() (archiver/csa_adler32.cpp:63-128) of every
// raw chunk is taken from its copy in HBM on the way in / of every decoded chunk before it is downloaded, instead of a
// separate pass over the bytes on a host thread (csa_io.h:250, :342).
int CSCB200_EncodeChunksAdler(int n, const uint8_t *const *in, const size_t *in_size, uint32_t dict_size, int level, int flt_mask,
                              int with_props, uint8_t *const *out, const size_t *out_cap, size_t *out_size, uint32_t *adler_in) {
    return fan_out(n, true, [&](Ctx &C, int lo, int hi) {
        return encode_chunks_on(C, hi - lo, in + lo, in_size + lo, dict_size, level, flt_mask, with_props, out + lo, out_cap + lo, out_size + lo, false,
                                adler_in ? adler_in + lo : nullptr);
    });
}
int CSCB200_DecodeChunksAdler(int n, const uint8_t *const *in, const size_t *in_size, uint8_t *const *out, const size_t *out_cap,
                              size_t *out_size, uint32_t *adler_out) {
    return fan_out(n, true, [&](Ctx &C, int lo, int hi) {
        return decode_chunks_on(C, hi - lo, in + lo, in_size + lo, out + lo, out_cap + lo, out_size + lo, adler_out ? adler_out + lo : nullptr);
    });
}
// adler32(init[i], buf[i], size[i]) for n host buffers (init == NULL: 1), in waves that fit the device
int CSCB200_Adler32(int n, const uint8_t *const *buf, const size_t *size, const uint32_t *init, uint32_t *out) {
    if (n <= 0) return 0;
    return fan_out(n, true, [&](Ctx &C, int lo, int hi) {
        const size_t budget = mem_budget(C) / 2;
        int at = lo;
        while (at < hi) {
            std::vector<u64> off, len;
            size_t tot = 0;
            int first = at;
            while (at < hi && (off.empty() || tot + size[at] + 256 <= budget)) { off.push_back(tot); len.push_back(size[at]); tot = align_up(tot + size[at], 256); at++; }
            if (C.lz.ensure(tot + 256)) return (int)kErrCuda;
            for (int i = first; i < at; i++)
                if (size[i]) CK(cudaMemcpyAsync(C.lz.as<u8>() + off[i - first], buf[i], size[i], cudaMemcpyHostToDevice, C.st));
            u32 *h = nullptr;
            int r = adler_ranges(C, C.lz.as<u8>(), off, len, init ? init + first : nullptr, &h);
            if (r < 0) return r;
            CK(cudaStreamSynchronize(C.st));
            for (int i = first; i < at; i++) out[i] = h[i - first];
        }
        return 0;
    });
}

// ------------------------------------------------------------------- libcsc handle API
// A handle owns no device state: every call leases a context.  One handle = one serially dependent stream
// (SURVEY.md H1), so the stream's whole history is kept on the host and a further CSCEnc_Encode call codes the
// stream again from its start and emits only the new segments (csa_worker.cpp / csc.cpp call it once).
struct EncHandle {
    CSCProps props;
    ISeqOutStream *os;
    Framer *fr;
    int dev;
    std::vector<u8> data;        // everything read so far
    std::vector<u32> segs;       // one entry per Read() = per flush segment (csc_enc.cpp:169-176)
    std::vector<u8> flip;        // duplicate-block outcomes learned so far (plan.h)
    size_t segs_done;            // segments already framed and written
    uint64_t outsize;
};

void CSCEncProps_Init(CSCProps *p, uint32_t dict_size, int level) { props_init(p, dict_size, level); }
void CSCEnc_WriteProperties(const CSCProps *props, uint8_t *s, int full) { (void)full; write_props(props, s); }
uint64_t CSCEnc_EstMemUsage(const CSCProps *p) { return est_mem_usage(p); }

CSCEncHandle CSCEnc_Create(const CSCProps *props, ISeqOutStream *outstream, ISzAlloc *alloc) {
    (void)alloc;   // device memory is managed by the library; host scratch uses the C++ heap
    if (!props) return NULL;
    const std::vector<int> devs = device_list();
    if (devs.empty()) return NULL;
    EncHandle *h = new (std::nothrow) EncHandle;
    if (!h) return NULL;
    h->props = *props; h->os = outstream; h->segs_done = 0; h->outsize = 0;
    h->dev = devs[devs.size() > 1 ? g_rr.fetch_add(1) % devs.size() : 0];   // handles of one process spread over the device list
    h->fr = new Framer(props->csc_blocksize);
    return h;
}
void CSCEnc_Destroy(CSCEncHandle p) {
    EncHandle *h = (EncHandle *)p;
    if (!h) return;
    delete h->fr;
    delete h;
}

struct CbSink {
    ISeqOutStream *os; bool fail; uint64_t payload;
    void operator()(const u8 *d, u32 n) {
        if (fail || !n) return;
        if (os->Write(os, d, n) != n) fail = true;
    }
};

int CSCEnc_Encode(CSCEncHandle p, ISeqInStream *is, ICompressProgress *progress) {
    EncHandle *h = (EncHandle *)p;
    if (!h || device_count() <= 0) {
        // Callers such as libcsc/csc.cpp do not check CSCEnc_Create's result, and csarc feeds the stream from
        // a producer thread that blocks until its queue is read: consume the input before failing.
        if (is) {
            std::vector<u8> sink(1 << 20);
            for (;;) { size_t got = sink.size(); if (is->Read(is, sink.data(), &got) < 0 || !got) break; }
        }
        return kErrCuda;
    }
    if (!is) return kErrCuda;
    // pull the input, one Read() of raw_blocksize per flush segment (csc_enc.cpp:169-176)
    std::vector<u8> &data = h->data;
    uint64_t insize = 0;
    const size_t segs_before = h->segs.size();
    for (;;) {
        size_t want = h->props.raw_blocksize;
        size_t old = data.size();
        data.resize(old + want);
        size_t got = want;
        int rr = is->Read(is, data.data() + old, &got);
        if (rr < 0) { data.resize(old); return READ_ERROR; }
        data.resize(old + got);
        if (!got) break;
        h->segs.push_back((u32)got);
        insize += got;
        if (progress) progress->Progress(progress, insize, h->outsize);
        if (data.size() >= 0xF0000000ull) return kErrUnsupported;
    }
    if (h->segs.size() == segs_before) return 0;
    EncJob J;
    J.props = h->props; J.status = 0; J.segs = h->segs; J.flip = h->flip;
    J.in = data.data(); J.n = (u32)data.size();
    Lease L(h->dev);
    if (!L.c) return kErrCuda;
    Ctx &C = *L.c;
    std::vector<EncJob *> wave{&J};
    int r;
    do { r = encode_wave(C, wave); } while (r >= 0 && J.status == kErrReplan);   // wrapped window: plan again (plan.h)
    if (r < 0) return r;
    h->flip = J.flip;
    CbSink sink{h->os, false, 0};
    for (size_t sgi = h->segs_done; sgi < J.so.size(); sgi++) {
        const SegOut &o = J.so[sgi];
        frame_segment(J, sgi, *h->fr, sink);
        h->outsize += o.rc_len + o.bc_len;
    }
    h->segs_done = J.so.size();
    if (progress) progress->Progress(progress, insize, h->outsize);
    return sink.fail ? WRITE_ERROR : 0;
}

int CSCEnc_Encode_Flush(CSCEncHandle p) {
    EncHandle *h = (EncHandle *)p;
    if (!h) return kErrCuda;
    CbSink sink{h->os, false, 0};
    h->fr->eof(sink);
    return sink.fail ? WRITE_ERROR : 0;
}

struct DecHandle {
    CSCProps props;
    ISeqInStream *is;
    int dev;
};

void CSCDec_ReadProperties(CSCProps *props, uint8_t *s) { read_props(props, s); }

CSCDecHandle CSCDec_Create(const CSCProps *props, ISeqInStream *instream, ISzAlloc *alloc) {
    (void)alloc;
    if (!props || props->dict_size > 1024u * 1048576u || props->dict_size < 32 * 1024) return NULL;
    const std::vector<int> devs = device_list();
    if (devs.empty()) return NULL;
    DecHandle *h = new (std::nothrow) DecHandle;
    if (!h) return NULL;
    h->props = *props; h->is = instream;
    h->dev = devs[devs.size() > 1 ? g_rr.fetch_add(1) % devs.size() : 0];
    return h;
}
void CSCDec_Destroy(CSCDecHandle p) { delete (DecHandle *)p; }

// exactly `n` bytes from the caller's stream (the reference asks for header, size and payload separately and
// never for more than the framing announces: csc_memio.cpp:19-57); false on a short read
static bool read_exact(ISeqInStream *is, u8 *dst, size_t n) {
    size_t have = 0;
    while (have < n) {
        size_t got = n - have;
        if (is->Read(is, dst + have, &got) < 0 || !got) return false;
        have += got;
    }
    return true;
}

// Pulls framed blocks up to and including the next candidate for the stream's last segment: a 6-byte RC block
// followed by the 3-byte BC block 19 00 00, which is what EncodeInt(SIG_EOF) + Flush leave behind
// (csc_encoder_main.cpp:149-157).  Returns 1 at such a pair, 0 when the input ended first, -1 on bad framing.
static int pull_until_eof_pair(ISeqInStream *is, u32 bsize, DecInput &d) {
    u32 prev_rc6 = 0;
    for (;;) {
        u8 fb;
        if (!read_exact(is, &fb, 1)) return 0;
        u32 sz = bsize;
        if (!((fb >> 6) & 1)) {
            u8 sb[3];
            if (!read_exact(is, sb, 3)) return -1;
            sz = ((u32)sb[0] << 16) + ((u32)sb[1] << 8) + sb[2];
        }
        if (!sz || sz > bsize) return -1;   // "must be a abnormal stream" (csc_memio.cpp:43-46)
        std::vector<u8> &v = (fb >> 7) ? d.rc : d.bc;
        std::vector<u32> &off = (fb >> 7) ? d.rc_off : d.bc_off;
        const size_t at = v.size();
        off.push_back((u32)at);
        v.resize(at + sz);
        if (!read_exact(is, v.data() + at, sz)) return -1;
        if (fb >> 7) { prev_rc6 = (sz == 6) ? 1 : 0; continue; }
        if (prev_rc6 && sz == 3 && v[at] == 0x19 && v[at + 1] == 0 && v[at + 2] == 0) return 1;
        prev_rc6 = 0;
    }
}

int CSCDec_Decode(CSCDecHandle p, ISeqOutStream *os, ICompressProgress *progress) {
    DecHandle *h = (DecHandle *)p;
    if (!h || !os || !h->is) return kErrCuda;
    if (device_count() <= 0) return kErrCuda;
    // The compressed stream is pulled block by block, never past its final segment (a caller's stream may block
    // or carry other data behind it), then decoded on the device in one go.  The raw size is not part of the
    // stream: the output capacity is a guess that grows until the decoder fits.
    DecInput di;
    size_t cap = 0;
    std::vector<u8> out;
    DecJob J;
    Lease L(h->dev);
    if (!L.c) return kErrCuda;
    Ctx &C = *L.c;
    for (;;) {
        const int pr = pull_until_eof_pair(h->is, h->props.csc_blocksize, di);
        if (pr < 0) return -1;
        if (di.rc_off.empty() || di.bc_off.empty()) return pr ? -1 : READ_ERROR;
        const size_t comp = di.rc.size() + di.bc.size();
        if (!cap) cap = std::min<size_t>(std::max<size_t>(comp * 12, (size_t)32 << 20), 0xFFFF0000ull - 64);
        int r;
        for (;;) {
            out.resize(cap);
            J = DecJob();
            J.props = h->props; J.predeframed = true; J.in = nullptr; J.n = comp;
            J.di.rc = di.rc; J.di.bc = di.bc; J.di.rc_off = di.rc_off; J.di.bc_off = di.bc_off;
            J.di.rc_off.push_back((u32)J.di.rc.size()); J.di.bc_off.push_back((u32)J.di.bc.size());
            J.di.rc.resize(J.di.rc.size() + 16, 0); J.di.bc.resize(J.di.bc.size() + 16, 0);
            J.out = out.data(); J.out_cap = cap;
            std::vector<DecJob *> wave{&J};
            r = decode_wave(C, wave, true);
            if (r == kErrOverflow && cap < 0xFFFF0000ull - 64) { cap = std::min<size_t>(cap * 6, 0xFFFF0000ull - 64); continue; }
            break;
        }
        // ran out of blocks before SIG_EOF although the pair looked like the end (a 6 + 3 byte segment in mid-stream): pull on
        if ((r == kErrRead || r == -1) && pr == 1) continue;
        uint64_t done = 0;
        for (const DecRun &rn : J.runs) {   // one Write per decoded run, as CSCDec_Decode does (csc_dec.cpp:767); also ahead of an error
            if (!rn.size) continue;
            done += rn.size;
            if (progress) progress->Progress(progress, comp, done);
            size_t w = os->Write(os, out.data() + rn.off, rn.size);
            if (w == CSC_WRITE_ABORT) return 0;
            if (w < rn.size) return WRITE_ERROR;
        }
        return r < 0 ? r : 0;
    }
}

}  // extern "C"
