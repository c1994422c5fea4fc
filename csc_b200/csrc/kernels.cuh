// kernels.cuh -- sm_100a kernels of the libcsc hot path.
//
//   k_analyze      one warp per 8 KiB block: histogram -> entropy, type heuristics, channel
//                  statistics, delta-channel entropy          (csc_analyzer.cpp:122-239)
//   k_plan         static EXE / ENGTXT run extraction          (csc_encoder_main.cpp:95-135)
//   k_filter_runs  e8e9 / dictionary forward transforms per run (csc_filters.cpp:256-335,588)
//   k_mf_hash      padded HASH6 keys + HASH2%16 / HASH6==0 bitmaps (csc_mf.cpp:23-42,215,136)
//   k_mf_rows      sorted (stream,hash,pos) -> K nearest earlier same-hash positions per position
//   k_mf_extend    forward match length of every (position, candidate) pair, capped at the
//                  sub-block end                               (csc_mf.cpp:470, 8 B words)
//   k_mf_tlist     thread per position: table candidates of find_match under the all-alive assumption (TRec)
//   k_adler32_*    Adler-32 of many buffers (csarc's checksum): CTA per 64 KiB tile + in-order combine
//   k_frame        one warp per stream: MemIO framing of the finished sub-streams (frame.cuh)
//   k_encode       one stream per warp: enc_engine.cuh (parser + adaptive model -> decision lists)
//   k_rc           one stream per warp, warp-uniform: rc_engine.cuh (range-coder recurrence over the decision lists)
//   k_decode       one stream per warp: dec_engine.cuh
//   k_unfilter     inverse e8e9 / dictionary per decoded run   (csc_filters.cpp:337,600)
//
// Everything here is HBM/latency-bound integer and byte work: no tensor cores on purpose.
#pragma once
#include <cuda_runtime.h>

#include "csc_defs.h"
#include "warp.h"
#include "blocks.cuh"
#include "plan.h"
#include "enc_engine.cuh"
#include "rc_engine.cuh"
#include "frame.cuh"
#include "dec_engine.cuh"
#include "radix_sort.cuh"

namespace cscb {

// Two grids use this descriptor: the analyzer's 8 KiB blocks (built on the host) and the sub-blocks of the
// window copy (built by k_plan from plan.h: identical to the block grid unless the window wraps inside a run).
struct BlockDesc {
    u64 goff;      // byte offset of the block in the batch LZ buffer
    u32 x;         // byte offset of the block in its stream
    u32 size;      // <= 8192 (0: unused slot)
    u32 stream;
    u32 seg;       // batch-global segment index
    u32 cpos;      // compact position index of the block's first byte (sort arrays)
    u32 flags;     // sub-block grid: bit 0 = left out of the 2-/3-byte chains (plan.h)
};
struct SegDesc {
    u64 goff;      // byte offset of the segment in the batch LZ buffer
    u32 x;         // offset in stream
    u32 size;
    u32 stream;
    u32 blk0;      // batch-global index of the segment's first block
    u32 flt;       // bit0 DLT, bit1 TXT, bit2 EXE
    u32 wnd;       // LZ window size of the stream
    u32 sub0;      // first slot of this segment in the sub-block grid
    u32 nslots;    // its slots there: blocks + window ends inside the segment
    u32 cpos;      // compact position index of the segment's first byte
    u32 seg;       // batch-global segment index
};
struct GRun { u64 goff; u32 size; u32 kind; u32 gblk; u32 pad; };

// ------------------------------------------------------------------ analyzer
static const int kAnWarps = 4;
static const int kStageVec = (kBlock + 32) / 16 + 1;

__device__ __forceinline__ void hist_clear(u32 *h, u32 lane) {
#pragma unroll
    for (int i = 0; i < 8; i++) h[lane + 32 * i] = 0;
    __syncwarp();
}
__device__ __forceinline__ u32 hist_entropy(const u32 *h, u32 size, const u32 *logtab, u32 lane) {
    u32 part = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { u32 f = h[lane + 32 * i]; part += f * logtab[f >> 4]; }
    u32 tot = __reduce_add_sync(0xFFFFFFFFu, part);
    return size * logtab[size >> 4] - tot;   // u32 wrap-around arithmetic, order independent
}
// histogram of the channel-delta residual (GetDltBpb): prev of the first byte of channel c is
// the last byte of channel c-1
__device__ __forceinline__ u32 dlt_bpb_warp(const u8 *sb, u32 size, u32 chn, u32 *h, const u32 *logtab, u32 lane) {
    hist_clear(h, lane);
    u32 base = size / chn, rem = size % chn;
    for (u32 j = lane; j < size; j += 32) {
        u32 prev;
        if (j >= chn) prev = sb[j - chn];
        else if (j == 0) prev = 0;
        else { u32 c = j - 1; u32 cnt = base + (c < rem ? 1u : 0u); prev = cnt ? sb[c + (cnt - 1) * chn] : 0; }
        atomicAdd(&h[(sb[j] - prev) & 0xFF], 1u);
    }
    __syncwarp();
    return hist_entropy(h, size, logtab, lane) / size;
}

__global__ void __launch_bounds__(kAnWarps * 32)
k_analyze(const u8 *__restrict__ lz, const BlockDesc *__restrict__ blk, u32 nblk, const AnalyzerTables *__restrict__ Tg,
          u32 *__restrict__ binfo, u32 *__restrict__ bbpb, u32 *__restrict__ tail_dlt) {
    __shared__ uint4 s_data[kAnWarps][kStageVec];
    __shared__ u32 s_hist[kAnWarps][256];
    __shared__ AnalyzerTables T;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const u32 *src = reinterpret_cast<const u32 *>(Tg);
        u32 *dst = reinterpret_cast<u32 *>(&T);
        for (u32 i = threadIdx.x; i < sizeof(AnalyzerTables) / 4; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const u32 b = blockIdx.x * kAnWarps + warp;
    if (b >= nblk) return;
    const BlockDesc bd = blk[b];
    const u32 size = bd.size;
    // stage the block with 16-byte loads from the enclosing aligned span
    const u64 a0 = bd.goff & ~(u64)15;
    const u32 shift = (u32)(bd.goff - a0);
    const u32 nvec = (shift + size + 15) / 16;
    const uint4 *g = reinterpret_cast<const uint4 *>(lz + a0);
    for (u32 i = lane; i < nvec; i += 32) s_data[warp][i] = g[i];
    __syncwarp();
    const u8 *sb = reinterpret_cast<const u8 *>(s_data[warp]) + shift;
    u32 *h = s_hist[warp];

    if (size < 512) {   // DT_SKIP (csc_analyzer.cpp:193): the engine may need GetDltBpb of the inherited type
        for (u32 k = 0; k < DLT_CHANNEL_MAX; k++) {
            u32 v = size ? dlt_bpb_warp(sb, size, dlt_channels(k), h, T.logtab, lane) : 0;
            if (lane == 0) tail_dlt[bd.seg * 5 + k] = v;
        }
        if (lane == 0) { binfo[b] = DT_SKIP; bbpb[b] = 0; }
        return;
    }
    hist_clear(h, lane);
    for (u32 i = lane; i < size; i += 32) atomicAdd(&h[sb[i]], 1u);
    __syncwarp();
    const u32 entropy = hist_entropy(h, size, T.logtab, lane);
    u32 d = 0, hi = 0, al = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u32 idx = lane + 32 * i, f = h[idx];
        d += (f > 0);
        if (idx >= 128) hi += f;
        if (idx >= 'a' && idx <= 'z') al += f;
    }
    d = __reduce_add_sync(0xFFFFFFFFu, d);
    hi = __reduce_add_sync(0xFFFFFFFFu, hi);
    al = __reduce_add_sync(0xFFFFFFFFu, al);
    bool need_chn;
    u32 type = type_from_stats(h, size, entropy, d, hi, al, T, &need_chn);
    if (need_chn) {   // get_channel_idx (csc_analyzer.cpp:122-164)
        u32 same[DLT_CHANNEL_MAX] = {0, 0, 0, 0, 0}, succ[DLT_CHANNEL_MAX] = {0, 0, 0, 0, 0};
        for (u32 i = lane; i + 16 < size; i += 32) {
            u32 v = sb[i];
#pragma unroll
            for (u32 k = 0; k < DLT_CHANNEL_MAX; k++) {
                u32 w = sb[i + dlt_channels(k)];
                same[k] += (v == w);
                succ[k] += absdiff(v, w);
            }
        }
#pragma unroll
        for (u32 k = 0; k < DLT_CHANNEL_MAX; k++) {
            same[k] = __reduce_add_sync(0xFFFFFFFFu, same[k]);
            succ[k] = __reduce_add_sync(0xFFFFFFFFu, succ[k]);
        }
        type = type_after_chn(channel_from_stats(same, succ, size), entropy, size);
    }
    const u32 bpb = entropy / size;
    u32 fail = 0;
    if (type >= DT_DLT) {
        u32 dv = dlt_bpb_warp(sb, size, dlt_channels(type - DT_DLT), h, T.logtab, lane);
        fail = ((double)dv >= (double)bpb * 0.95) ? 1u : 0u;
    }
    if (lane == 0) { binfo[b] = type | (fail << 8); bbpb[b] = bpb; }
}

// -------------------------------------------------------------- plan + filters
__global__ void k_plan(const SegDesc *__restrict__ segs, u32 nseg, const u32 *__restrict__ binfo, const u32 *__restrict__ bbpb,
                       const u32 *__restrict__ tail_dlt, const u8 *__restrict__ flip, FilterRun *scratch, SubBlock *subscratch,
                       GRun *runs, u32 *nruns, u8 *nolz, BlockDesc *subdesc) {
    u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const SegDesc sd = segs[s];
    FilterRun *mine = scratch + sd.blk0;
    SubBlock *subs = subscratch + sd.sub0;
    u32 nsub = 0;
    u32 nr = plan_segment(binfo, bbpb, tail_dlt + (size_t)sd.seg * 5, flip, sd.blk0, sd.x, sd.size, sd.wnd, sd.flt & 1, (sd.flt >> 1) & 1,
                          (sd.flt >> 2) & 1, mine, nolz, subs, &nsub);
    for (u32 i = 0; i < sd.nslots; i++) {   // the sub-block grid the match-finder kernels run over
        BlockDesc bd;
        const u32 rel = i < nsub ? subs[i].x - sd.x : 0;   // relative to the segment
        bd.goff = sd.goff + rel; bd.x = sd.x + rel; bd.size = i < nsub ? subs[i].size : 0; bd.stream = sd.stream; bd.seg = sd.seg;
        bd.cpos = sd.cpos + rel; bd.flags = i < nsub ? subs[i].flags : 0;
        subdesc[sd.sub0 + i] = bd;
    }
    if (!nr) return;
    u32 at = atomicAdd(nruns, nr);
    for (u32 i = 0; i < nr; i++) {
        GRun r;
        r.goff = sd.goff - sd.x + mine[i].off; r.size = mine[i].size; r.kind = mine[i].kind; r.gblk = mine[i].first_blk; r.pad = 0;
        runs[at + i] = r;
    }
}

// ---- parallel filters: one CTA per run, 8 KiB tiles, 32 bytes per thread ---------------
static const int kFltThreads = 256;
static const int kFltSpan = 32;
static const int kFltTile = kFltThreads * kFltSpan;
static const int kFltLook = 64;   // look-back for re-synchronising the greedy tokenisation

__device__ __forceinline__ u32 block_exscan(u32 v, u32 *s_warp, u32 *total) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (u32)d) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    u32 wbase = 0, tot = 0;
    for (u32 w = 0; w < kFltThreads / 32; w++) { u32 x = s_warp[w]; if (w < warp) wbase += x; tot += x; }
    __syncthreads();
    *total = tot;
    return wbase + inc - v;
}

// e8e9 (csc_filters.cpp:508-609).  The greedy opcode selection is a 4-state machine
// (state = masked positions left after an E8/E9 hit); each thread composes the transition
// function of its 32 bytes, a block-wide scan of function compositions gives every thread
// its true entry state, and the operand rewrites are independent of each other.
__device__ __forceinline__ u32 e89_step(u32 st, u32 byte, bool eligible) {
    if (st) return st - 1;
    return (eligible && (byte & 0xFE) == 0xE8) ? 3u : 0u;
}
__device__ void e8e9_run_cta(u8 *run, u32 size, bool inverse, u8 *s, u32 *s_warp) {
    __shared__ u32 s_state;
    if (threadIdx.x == 0) s_state = 0;
    __syncthreads();
    if (size < 6) return;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 t0 = 0; t0 < size; t0 += kFltTile) {
        const u32 tn = min((u32)kFltTile, size - t0);
        for (u32 i = threadIdx.x; i < tn + 8; i += kFltThreads) s[i] = (t0 + i < size) ? run[t0 + i] : (u8)0;
        __syncthreads();
        const u32 a = threadIdx.x * kFltSpan, b = min(a + kFltSpan, tn);
        // transition function of this span: 2 bits per entry state
        u32 f = 0;
        for (u32 st0 = 0; st0 < 4; st0++) {
            u32 st = st0;
            for (u32 i = a; i < b; i++) st = e89_step(st, s[i], t0 + i + 5 < size);
            f |= st << (2 * st0);
        }
        if (a >= tn) f = 0xE4;   // identity
        // inclusive scan of compositions (apply earlier function first)
        u32 inc = f;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (u32)d) {
                u32 c = 0;
                for (u32 k = 0; k < 4; k++) { u32 mid = (o >> (2 * k)) & 3; c |= ((inc >> (2 * mid)) & 3) << (2 * k); }
                inc = c;
            }
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        u32 st = s_state;
        for (u32 w = 0; w < warp; w++) st = (s_warp[w] >> (2 * st)) & 3;
        u32 prevf = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
        if (lane) st = (prevf >> (2 * st)) & 3;
        // st = entry state of this thread; walk the span and rewrite operands
        for (u32 i = a; i < b; i++) {
            u32 ns = e89_step(st, s[i], t0 + i + 5 < size);
            if (!st && ns == 3) e89_apply(run, t0 + i, inverse);
            st = ns;
        }
        __syncthreads();
        if (threadIdx.x == kFltThreads - 1) s_state = st;   // spans past the tile end are identity
        __syncthreads();
    }
}

// dictionary forward (csc_filters.cpp:256-335).  Tokens never cross from a non-letter into a
// letter run, so every thread re-synchronises by walking back to the start of the letter run
// that covers its span start (bounded; pathological runs fall back to the scalar statement).
struct DictSmem { u16 next[300][26]; u8 symbol[300]; };
__device__ __forceinline__ bool is_lower(u32 c) { return c >= 'a' && c <= 'z'; }
__device__ __forceinline__ u32 dict_tok(const u8 *p, const DictSmem &D, u32 *sym) {   // p[0] is a letter
    u32 node = 0, best = 0, bs = 0;
    for (u32 j = 0; j < 8;) {
        u32 idx = (u32)p[j] - 'a';
        if (idx > 25) break;
        u32 nx = D.next[node][idx];
        if (!nx) break;
        node = nx; j++;
        if (D.symbol[node]) { bs = D.symbol[node]; best = j; }
    }
    *sym = bs;
    return bs ? best : 1;
}
__device__ void dict_forward_cta(u8 *run, u32 size, u8 *tmp, const DictTables *Dg, u8 *ok_out, DictSmem &D, u8 *s, u32 *s_warp) {
    __shared__ u32 s_base, s_fallback;
    if (size < 16384) { if (threadIdx.x == 0) *ok_out = 0; return; }
    const u32 main_end = size - 5;   // tokens start below this; the tail is copied byte-wise
    u32 total = 0;
    for (int pass = 0; pass < 2; pass++) {
        if (threadIdx.x == 0) { s_base = 0; s_fallback = 0; }
        __syncthreads();
        for (u32 t0 = 0; t0 < size; t0 += kFltTile) {
            const u32 tn = min((u32)kFltTile, size - t0);
            // s[kFltLook + i] = run[t0 + i]
            for (u32 i = threadIdx.x; i < kFltLook + tn + 8; i += kFltThreads) {
                long long g = (long long)t0 + i - kFltLook;
                s[i] = (g >= 0 && g < (long long)size) ? run[g] : (u8)0;
            }
            __syncthreads();
            const u8 *w = s + kFltLook;   // w[i] = run[t0 + i]
            const u32 a = threadIdx.x * kFltSpan, b = min(a + kFltSpan, tn);
            u32 pos = a, cnt = 0;
            if (a < tn) {
                // first token start >= a
                if (t0 + a > 0 && is_lower(w[(int)a - 1]) && is_lower(w[a])) {
                    int back = (int)a - 1;
                    while (back > -(int)kFltLook && (long long)t0 + back > 0 && is_lower(w[back - 1])) back--;
                    // the letter run may start before the look-back window: give up on the parallel form
                    if (back == -(int)kFltLook && (long long)t0 + back > 0) s_fallback = 1;
                    u32 p = (u32)(back + (int)kFltLook);   // index into s
                    const u32 target = a + kFltLook;
                    while (p < target) {
                        u32 gp = t0 + p - kFltLook;
                        u32 sym, l = (gp < main_end && is_lower(s[p])) ? dict_tok(s + p, D, &sym) : 1;
                        p += l;
                    }
                    pos = p - kFltLook;
                }
                // count / emit the tokens that start in [pos, b)
                for (u32 i = pos; i < b;) {
                    u32 gp = t0 + i, c = w[i];
                    if (gp < main_end && is_lower(c)) { u32 sym; i += dict_tok(w + i, D, &sym); cnt += 1; }
                    else { cnt += (c >= 0x82) ? 2 : 1; i++; }
                }
            }
            u32 tile_total;
            u32 off = block_exscan(cnt, s_warp, &tile_total);
            if (pass == 1 && a < tn) {
                u8 *o = tmp + s_base + off;
                for (u32 i = pos; i < b;) {
                    u32 gp = t0 + i, c = w[i];
                    if (gp < main_end && is_lower(c)) {
                        u32 sym, l = dict_tok(w + i, D, &sym);
                        *o++ = sym ? (u8)sym : (u8)c;
                        i += l;
                    } else {
                        if (c >= 0x82) *o++ = 254;
                        *o++ = (u8)c;
                        i++;
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) s_base += tile_total;
            __syncthreads();
        }
        total = s_base;
        if (s_fallback) {   // pathological letter run: scalar statement (exactness first)
            if (threadIdx.x == 0) *ok_out = (u8)dict_forward(run, size, tmp, *Dg);
            __syncthreads();
            return;
        }
        if (pass == 0 && (double)total > (double)size * 0.82) { if (threadIdx.x == 0) *ok_out = 0; return; }
        __syncthreads();
    }
    for (u32 i = threadIdx.x; i < size; i += kFltThreads) run[i] = i < total ? tmp[i] : (u8)0x20;
    if (threadIdx.x == 0) *ok_out = 1;
}

__global__ void __launch_bounds__(kFltThreads)
k_filter_runs(u8 *lz, u8 *tmp, const GRun *__restrict__ runs, const u32 *__restrict__ nruns, const DictTables *__restrict__ Dg,
              u8 *dict_ok) {
    __shared__ DictSmem D;
    __shared__ u8 s[kFltLook + kFltTile + 16];
    __shared__ u32 s_warp[kFltThreads / 32];
    for (u32 i = threadIdx.x; i < 300 * 26; i += kFltThreads) D.next[i / 26][i % 26] = Dg->next[i / 26][i % 26];
    for (u32 i = threadIdx.x; i < 300; i += kFltThreads) D.symbol[i] = Dg->symbol[i];
    __syncthreads();
    const u32 n = *nruns;
    for (u32 r = blockIdx.x; r < n; r += gridDim.x) {
        const GRun g = runs[r];
        if (g.kind == DT_EXE) e8e9_run_cta(lz + g.goff, g.size, false, s, s_warp);
        else dict_forward_cta(lz + g.goff, g.size, tmp + g.goff, Dg, dict_ok + g.gblk, D, s, s_warp);
        __syncthreads();
    }
}

// ------------------------------------------------------------------ match finder
// One CTA per sub-block of the reference window copy.  The sort key is the hash value alone: the radix
// sort is segmented by stream (radix_sort.cuh), so a stable sort of the hashes groups equal hashes of one
// stream in position order.
static const int kMfThreads = 256;
// Each thread hashes 4 consecutive positions from three staged words (funnel shifts instead of byte loads)
// and writes its keys / positions as 16-byte vectors; the HASH2%16 / HASH6==0 bitmaps are assembled per
// 32 positions with shuffles (one atomic per non-empty word instead of one per set bit).
__global__ void __launch_bounds__(kMfThreads)
k_mf_hash(const u8 *__restrict__ lz, const BlockDesc *__restrict__ blk, const u32 *__restrict__ stream_bits,
          const u64 *__restrict__ stream_goff, const u32 *__restrict__ stream_wnd, u32 key_shift, RsKey *__restrict__ keys,
          u32 *__restrict__ vals, u32 *Z2, u32 *Z6, u32 which) {
    __shared__ __align__(16) u32 sw[kBlock / 4 + 8];
    const BlockDesc bd = blk[blockIdx.x];
    if (!bd.size) return;
    {   // stage the block, zero padding behind its end (H2); blocks start 4-byte aligned in the batch buffer
        // only when the stream offset is, so go through bytes when it is not
        const u8 *src = lz + bd.goff;
        u8 *sb = reinterpret_cast<u8 *>(sw);
        if ((bd.goff & 15) == 0) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            for (u32 i = threadIdx.x; i < kBlock / 16 + 2; i += kMfThreads) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (i * 16 < bd.size) v = s4[i];   // buffers are padded, reads beyond the block stay inside the batch
                reinterpret_cast<uint4 *>(sw)[i] = v;
            }
            __syncthreads();
            // zero the bytes behind the block end inside the last staged vector
            for (u32 i = bd.size + threadIdx.x; i < ((bd.size + 15) & ~15u); i += kMfThreads) sb[i] = 0;
        } else {
            for (u32 i = threadIdx.x; i < kBlock + 32; i += kMfThreads) sb[i] = i < bd.size ? src[i] : (u8)0;
        }
    }
    __syncthreads();
    {   // once the window has wrapped, the bytes behind the sub-block end are the previous lap's, not zeros (plan.h: window_byte)
        const u32 wnd = stream_wnd[bd.stream], sb_end = bd.x + bd.size;
        if (sb_end > wnd) {
            if (threadIdx.x < 16)
                reinterpret_cast<u8 *>(sw)[bd.size + threadIdx.x] = (u8)window_byte(lz + stream_goff[bd.stream], sb_end + threadIdx.x, sb_end, wnd);
            __syncthreads();
        }
    }
    const u32 bits = stream_bits[bd.stream];
    const u64 bitbase = (stream_goff[bd.stream]) + bd.x;   // bitmaps are indexed by batch byte offset
    const bool excluded = which < 6 && (bd.flags & 1);
    const bool vec = (bd.cpos & 3) == 0;
    const u32 lane = threadIdx.x & 31;
    for (u32 i0 = threadIdx.x * 4; i0 < ((bd.size + 127) & ~127u); i0 += kMfThreads * 4) {
        const bool in = i0 < bd.size;
        u32 w0 = 0, w1 = 0, w2 = 0, w3 = 0;
        if (in) { w0 = sw[i0 >> 2]; w1 = sw[(i0 >> 2) + 1]; w2 = sw[(i0 >> 2) + 2]; w3 = sw[(i0 >> 2) + 3]; }
        RsKey k[4];
        u32 z2 = 0, z6 = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const u32 v = __funnelshift_r(w0, w1, 8 * j), u = __funnelshift_r(w1, w2, 8 * j), t = __funnelshift_r(w2, w3, 8 * j);
            const u32 h2 = ((v & 0xFFFF) * 65521u) & 0x3FFF;
            u32 h;
            if (which >= 8) h = hash_ctx_words(v, u, which == 12 ? t : 0, key_shift);          // -m5 context hashes (plan.h)
            else if (which == 6) h = bits ? ((v ^ ((u & 0xFFFF) << 13)) * 2654435761u) >> (32 - bits) : 0;   // HASH6 csc_mf.cpp:35-42
            else if (which == 3) h = ((v & 0xFF) << 8) ^ (((v >> 8) & 0xFF) << 5) ^ ((v >> 16) & 0xFF);   // HASH3 :30-33
            else h = h2;                                                                          // HASH2 :23-28
            // blocks left out of a 2-/3-byte chain carry flag bit 31 and sort into the all-ones bucket, where k_mf_rows skips them
            k[j] = excluded ? (0x80000000u | ((1u << key_shift) - 1)) : h;
            if (in && i0 + j < bd.size) {
                if ((h2 & 15) == 0) z2 |= 1u << j;
                if (bits && h == 0) z6 |= 1u << j;
            }
        }
        if (in) {
            const u32 c = bd.cpos + i0, x = bd.x + i0;
            if (vec && i0 + 4 <= bd.size) {
                *reinterpret_cast<uint4 *>(keys + c) = make_uint4(k[0], k[1], k[2], k[3]);
                *reinterpret_cast<uint4 *>(vals + c) = make_uint4(x, x + 1, x + 2, x + 3);
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) if (i0 + j < bd.size) { keys[c + j] = k[j]; vals[c + j] = x + j; }
            }
        }
        if (which == 6) {   // 8 lanes x 4 positions = one bitmap word (when the block is 32-aligned in the bitmap)
            u32 m2 = z2 << (4 * (lane & 7)), m6 = z6 << (4 * (lane & 7));
#pragma unroll
            for (int sft = 1; sft < 8; sft <<= 1) { m2 |= __shfl_xor_sync(0xFFFFFFFFu, m2, sft); m6 |= __shfl_xor_sync(0xFFFFFFFFu, m6, sft); }
            if ((lane & 7) == 0 && in) {
                const u64 bit = bitbase + (i0 & ~31u);
                const u32 sh = (u32)(bit & 31);
                if (m2) { atomicOr(&Z2[bit >> 5], m2 << sh); if (sh) atomicOr(&Z2[(bit >> 5) + 1], m2 >> (32 - sh)); }
                if (m6) { atomicOr(&Z6[bit >> 5], m6 << sh); if (sh) atomicOr(&Z6[(bit >> 5) + 1], m6 >> (32 - sh)); }
            }
        }
    }
}

// Rows from the sorted array, one CTA per sort tile (a tile never spans two streams).  Element i of a stream's
// sorted range has its same-hash predecessors right in front of it, in position order.
__global__ void __launch_bounds__(256)
k_mf_rows(const RsKey *__restrict__ keys, const u32 *__restrict__ vals, const RsTile *__restrict__ tiles,
          const u64 *__restrict__ stream_goff, const u32 *__restrict__ stream_K, u32 *__restrict__ P6, u32 kforce, u32 row_stride,
          u32 col0) {
    const RsTile t = tiles[blockIdx.x];
    const u32 K = kforce ? kforce : stream_K[t.stream];
    if (!K) return;
    const u64 sg = stream_goff[t.stream];
    for (u32 jj = threadIdx.x; jj < t.count; jj += 256) {
        const u32 i = t.start + jj;
        const RsKey k = keys[i];
        if (k >> 31) continue;   // position of a block left out of this chain (array pre-filled with kNone)
        const u32 before = i - t.sstart;   // elements of this stream in front of i
        // row_stride != 0: write K entries at column col0 of a wider row (-m5 merges several such lists)
        u32 *row = P6 + (sg + vals[i]) * (row_stride ? row_stride : K) + col0;
        if (kforce != 1 && K <= 8) {
            // 6-byte rows (no flagged entries): the K predecessors in the sorted array are read with independent
            // loads, the row is the prefix of them that still has this element's hash
            RsKey kk[8];
            u32 vv[8];
#pragma unroll
            for (u32 c = 0; c < 8; c++) {
                const bool in = c < K && before > c;
                kk[c] = in ? keys[i - 1 - c] : ~k;
                vv[c] = in ? vals[i - 1 - c] : kNone;
            }
            bool ok = true;
#pragma unroll
            for (u32 c = 0; c < 8; c++) { ok = ok && kk[c] == k; vv[c] = ok ? vv[c] : kNone; }
            if (K == 8 && !row_stride) {
                reinterpret_cast<uint4 *>(row)[0] = make_uint4(vv[0], vv[1], vv[2], vv[3]);
                reinterpret_cast<uint4 *>(row)[1] = make_uint4(vv[4], vv[5], vv[6], vv[7]);
            } else {
#pragma unroll
                for (u32 c = 0; c < 8; c++) if (c < K) row[c] = vv[c];
            }
            continue;
        }
        bool live = true;
        u32 j = i;
        for (u32 c = 0; c < K; c++) {
            // previous element of the same hash run; flagged entries only occur in the all-ones bucket
            while (live && j > t.sstart && (keys[j - 1] >> 31) && ((keys[j - 1] ^ k) << 1) == 0) j--;
            live = live && j > t.sstart && keys[j - 1] == k;
            row[c] = live ? vals[j - 1] : kNone;
            if (live) j--;
        }
    }
}

__device__ __forceinline__ u32 ld_u32_unaligned(const u8 *p) {
    const u32 *a = reinterpret_cast<const u32 *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    u32 sh = (u32)(reinterpret_cast<uintptr_t>(p) & 3) * 8;
    u32 lo = a[0];
    if (!sh) return lo;
    return __funnelshift_r(lo, a[1], sh);
}

// Forward match lengths L6[p][k] = LCP(p, P6[p][k]) capped at the sub-block end.
// One CTA per 8 KiB block, each thread walks a strip of 32 consecutive positions.  Inside a
// strip most compares disappear: if (p-1, e-1) was a pair of the previous position with length
// L >= 1 then LCP(p, e) = L - 1 exactly (same cap), which is the common case in repetitive
// data where whole phrases repeat.  Only strip heads and fresh chain entries touch the window,
// 4 bytes per step through aligned loads + funnel shift.
__device__ __forceinline__ u64 ld_u64_unaligned(const u8 *p) {
    const u64 *a = reinterpret_cast<const u64 *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)7);
    u32 sh = (u32)(reinterpret_cast<uintptr_t>(p) & 7) * 8;
    u64 lo = a[0];
    if (!sh) return lo;
    return (lo >> sh) | (a[1] << (64 - sh));
}
__device__ __forceinline__ u32 lcp_words(const u8 *a, const u8 *m, u32 limit) {
    u32 l = 0;
    while (l + 8 <= limit) {
        u64 d = ld_u64_unaligned(a + l) ^ ld_u64_unaligned(m + l);
        if (d) return l + ((__ffsll((long long)d) - 1) >> 3);
        l += 8;
    }
    while (l < limit && a[l] == m[l]) l++;
    return l;
}

// -m5: every 16-entry row holds three lists (nearest by 6-byte hash, 8- and 12-byte context); order it by
// distance (position descending), drop duplicates, absent entries last.  Entries are sorted as e + 1 so that
// kNone becomes 0; two passes of a 16-input bitonic network in registers.
__device__ __forceinline__ void sort16_desc(u32 *v) {
#pragma unroll
    for (int k = 2; k <= 16; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int l = i ^ j;
                if (l > i) {
                    const bool desc = (i & k) == 0;
                    const u32 a = v[i], b = v[l];
                    const bool sw = desc ? a < b : a > b;
                    v[i] = sw ? b : a; v[l] = sw ? a : b;
                }
            }
}
__global__ void __launch_bounds__(256)
k_mf_merge(const BlockDesc *__restrict__ blk, const u64 *__restrict__ stream_goff, u32 *__restrict__ P6) {
    const BlockDesc bd = blk[blockIdx.x];   // (an unused slot has size 0)
    const u64 sg = stream_goff[bd.stream];
    for (u32 i = threadIdx.x; i < bd.size; i += 256) {
        uint4 *row = reinterpret_cast<uint4 *>(P6 + (sg + bd.x + i) * 16);
        u32 v[16];
#pragma unroll
        for (int q = 0; q < 4; q++) { uint4 t = row[q]; v[4 * q] = t.x + 1; v[4 * q + 1] = t.y + 1; v[4 * q + 2] = t.z + 1; v[4 * q + 3] = t.w + 1; }
        sort16_desc(v);
#pragma unroll
        for (int q = 15; q > 0; q--) if (v[q] == v[q - 1]) v[q] = 0;
        sort16_desc(v);
#pragma unroll
        for (int q = 0; q < 4; q++) row[q] = make_uint4(v[4 * q] - 1, v[4 * q + 1] - 1, v[4 * q + 2] - 1, v[4 * q + 3] - 1);
    }
}

template <int K> struct RowIO;
template <> struct RowIO<16> {
    static __device__ __forceinline__ void load(const u32 *p, u32 *e) {
#pragma unroll
        for (int q = 0; q < 4; q++) { uint4 a = reinterpret_cast<const uint4 *>(p)[q]; e[4 * q] = a.x; e[4 * q + 1] = a.y; e[4 * q + 2] = a.z; e[4 * q + 3] = a.w; }
    }
    static __device__ __forceinline__ void store(u16 *p, const u32 *l, const u32 *) {
#pragma unroll
        for (int q = 0; q < 2; q++)
            reinterpret_cast<uint4 *>(p)[q] = make_uint4(l[8 * q] | (l[8 * q + 1] << 16), l[8 * q + 2] | (l[8 * q + 3] << 16),
                                                          l[8 * q + 4] | (l[8 * q + 5] << 16), l[8 * q + 6] | (l[8 * q + 7] << 16));
    }
};
template <> struct RowIO<1> {
    static __device__ __forceinline__ void load(const u32 *p, u32 *e) { e[0] = p[0]; }
    static __device__ __forceinline__ void store(u16 *p, const u32 *l, const u32 *) { p[0] = (u16)l[0]; }
};
template <> struct RowIO<2> {
    static __device__ __forceinline__ void load(const u32 *p, u32 *e) { uint2 v = *reinterpret_cast<const uint2 *>(p); e[0] = v.x; e[1] = v.y; }
    static __device__ __forceinline__ void store(u16 *p, const u32 *l, const u32 *) { *reinterpret_cast<u32 *>(p) = l[0] | (l[1] << 16); }
};
template <> struct RowIO<8> {
    static __device__ __forceinline__ void load(const u32 *p, u32 *e) {
        uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1];
        e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w;
    }
    static __device__ __forceinline__ void store(u16 *p, const u32 *l, const u32 *) {
        *reinterpret_cast<uint4 *>(p) = make_uint4(l[0] | (l[1] << 16), l[2] | (l[3] << 16), l[4] | (l[5] << 16), l[6] | (l[7] << 16));
    }
};

// Forward match lengths L6[p][k] = LCP(p, P6[p][k]) capped at the sub-block end.
//
// One CTA per sub-block, one LANE per position: a warp walks 1024 consecutive positions 32 at a time, so the
// rows arrive as coalesced kilobytes and the lengths leave the same way.  Most compares disappear: if
// (p-1, e-1) sits at the same rank of the previous position's row with length m >= 1, then LCP(p, e) = m - 1
// exactly (same cap), the common case in repetitive data where whole phrases repeat.  Such an entry is
// LINKED to its left neighbour; the others are chain HEADS and touch the window (8-byte words; matches that
// run long are finished by the whole warp).  The lengths then flow from the heads down the lanes as a
// segmented copy-scan of (length + lane), five shuffle steps per rank.  A link whose predecessor ends with
// length 0 (hash collision) breaks its chain: the entry behind it becomes a head and the scan is repeated.
template <int K>
__global__ void __launch_bounds__(kMfThreads)
k_mf_extend(const u8 *__restrict__ lz, const BlockDesc *__restrict__ blk, const u64 *__restrict__ stream_goff,
            const u32 *__restrict__ P6, u16 *__restrict__ L6) {
    const BlockDesc bd = blk[blockIdx.x];
    if (!bd.size) return;
    const u64 sg = stream_goff[bd.stream];
    const u8 *base = lz + sg;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, FULL = 0xFFFFFFFFu;
    const u32 span = 32 * 32;                                  // positions per warp
    const u32 w0 = warp * span;
    if (w0 >= bd.size) return;
    const u32 w1 = min(w0 + span, bd.size);
    u32 le[K], ll[K];                                          // row and lengths of the previous group's last lane
#pragma unroll
    for (int k = 0; k < K; k++) { le[k] = kNone; ll[k] = 0; }
    for (u32 g0 = w0; g0 < w1; g0 += 32) {
        const u32 i = g0 + lane;
        const bool in = i < w1;
        const u32 x = bd.x + i;
        const u32 limit = in ? bd.size - i : 0;
        const u64 at = (sg + x) * K;
        u32 e[K];
#pragma unroll
        for (int k = 0; k < K; k++) e[k] = kNone;
        if (in) RowIO<K>::load(P6 + at, e);
        // links: the same rank of the left neighbour holds e - 1
        u32 v[K];          // heads: length + lane (final); linked entries: the head's value, filled by the scan
        u32 linkm = 0;     // bit k: entry k is linked to lane - 1
        u32 headm = 0;     // bit k: entry k is a head whose v is final
        u32 todo = 0;      // bit k: head that still needs its compare
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 pe = __shfl_up_sync(FULL, e[k], 1);
            if (lane == 0) pe = le[k];
            const bool has = e[k] != kNone && e[k] != 0 && pe == e[k] - 1;
            v[k] = 0;
            if (e[k] == kNone) continue;
            if (lane == 0 && has && ll[k] >= 1) { v[k] = ll[k] - 1; headm |= 1u << k; }   // known from the previous group
            else if (lane != 0 && has) linkm |= 1u << k;
            else todo |= 1u << k;
        }
        for (;;) {
            // ---- compares of the new heads
            u32 longm = 0;
#pragma unroll
            for (int k = 0; k < K; k++) {
                if ((todo >> k) & 1) {
                    const u32 cap = limit < 64 ? limit : 64;
                    const u32 l = lcp_words(base + x, base + e[k], cap);
                    v[k] = l + lane;
                    if (l == 64 && limit > 64) longm |= 1u << k;
                }
            }
            if (__any_sync(FULL, longm != 0)) {
#pragma unroll
                for (int k = 0; k < K; k++) {   // long matches: the warp finishes them 128 bytes per step
                    u32 m = __ballot_sync(FULL, (longm >> k) & 1);
                    while (m) {
                        const u32 src = (u32)__ffs((int)m) - 1;
                        m &= m - 1;
                        const u32 sx = __shfl_sync(FULL, x, src), se = __shfl_sync(FULL, e[k], src), sl = __shfl_sync(FULL, limit, src);
                        const u32 ext = 64 + lcp_bytes(base + sx + 64, base + se + 64, sl - 64);
                        if (lane == src) v[k] = ext + lane;
                    }
                }
            }
            headm |= todo;
            // ---- segmented copy-scan down the lanes, per rank: a linked entry takes the value of its chain's head
            u32 done = headm;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 od = __shfl_up_sync(FULL, done, d);
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const u32 ov = __shfl_up_sync(FULL, v[k], d);
                    if (lane >= (u32)d && ((linkm & ~done) >> k) & 1) v[k] = ov;
                }
                if (lane >= (u32)d) done |= od & linkm;
            }
            // ---- a chain only holds while its lengths stay >= 1: head length - offset must not go below 0.  The entry
            // right behind a length 0 (value - lane == -1) is compared afresh and heads the rest of its chain.
            todo = 0;
#pragma unroll
            for (int k = 0; k < K; k++)
                if (((linkm >> k) & 1) && v[k] - lane == 0xFFFFFFFFu) todo |= 1u << k;
            if (!__any_sync(FULL, todo != 0)) break;
            linkm &= ~todo;
        }
        // lengths = v - lane (never negative once every chain held)
        u32 l[K];
#pragma unroll
        for (int k = 0; k < K; k++) l[k] = e[k] == kNone ? 0 : v[k] - lane;
        if (in) RowIO<K>::store(L6 + at, l, e);
#pragma unroll
        for (int k = 0; k < K; k++) { le[k] = __shfl_sync(FULL, e[k], 31); ll[k] = __shfl_sync(FULL, l[k], 31); }
    }
}

// ------------------------------------------------------------------ precomputed table candidates
// One thread per position: the table half of find_match under the all-alive assumption (enc_engine.cuh:
// tlist_of), 32 bytes per position.  Rows, chains and records are read / written as whole sectors by
// neighbouring threads.  Sub-blocks the static plan types as not LZ-parsed are skipped (their records are never
// read: a duplicate-block flip moves the stream to table mode), and so are wrapped streams.
template <int K>
__device__ __forceinline__ TRec tlist_row(const EncStream &S, u32 x, u32 limit, u32 &first) {
    u32 row[K], len[K];
    u16 lrow[K];
    RowIO<K>::load(S.P6 + (size_t)x * K, row);
    if (K == 1) lrow[0] = S.L6[x];
    else {
        // lengths are packed two per word, K / 2 words per row
        const u32 *lp = reinterpret_cast<const u32 *>(S.L6 + (size_t)x * K);
        if (K == 2) { const u32 w = lp[0]; len[0] = w & 0xFFFF; len[1] = w >> 16; }
        else {
#pragma unroll
            for (int q = 0; q < K / 8; q++) {
                const uint4 w = reinterpret_cast<const uint4 *>(lp)[q];
                len[8 * q] = w.x & 0xFFFF; len[8 * q + 1] = w.x >> 16; len[8 * q + 2] = w.y & 0xFFFF; len[8 * q + 3] = w.y >> 16;
                len[8 * q + 4] = w.z & 0xFFFF; len[8 * q + 5] = w.z >> 16; len[8 * q + 6] = w.w & 0xFFFF; len[8 * q + 7] = w.w >> 16;
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) lrow[k] = (u16)len[k];
    }
    first = row[0];
    return Encoder::tlist_of(x, limit, S.cfg.vld_rge, K, K, S.cfg.good_len, S.P2[x], S.L2[x], S.P3[x], S.L3[x], row, lrow);
}
__global__ void __launch_bounds__(256)
k_mf_tlist(const BlockDesc *__restrict__ blk, const EncStream *__restrict__ streams) {
    const BlockDesc bd = blk[blockIdx.x];
    if (!bd.size || (bd.flags & 1)) return;
    const EncStream &S = streams[bd.stream];
    if (!S.T || S.n > S.cfg.wnd_size) return;
    if ((bd.flags & 2) && S.cfg.lz_mode == 2) return;   // the lazy parser stays on its table path in text runs (enc_engine.cuh: lz_run)
    const u32 K = S.K;
    if (S.cfg.ht_width != K) return;   // (never: the row depth is the table width)
    for (u32 i0 = 0; i0 < bd.size; i0 += 256) {
        const u32 i = i0 + threadIdx.x;
        const u32 x = bd.x + i;
        bool flag = false;   // SlidePos' 6-byte rules (csc_mf.cpp:147-157) can act on this position
        bool has = false;    // the record holds table candidates
        if (i < bd.size) {
            u32 first = kNone;
            TRec t;
            if (K == 1) t = tlist_row<1>(S, x, bd.size - i, first);
            else if (K == 2) t = tlist_row<2>(S, x, bd.size - i, first);
            else if (K == 8) t = tlist_row<8>(S, x, bd.size - i, first);
            else t = tlist_row<16>(S, x, bd.size - i, first);
            uint4 *o = reinterpret_cast<uint4 *>(S.T + x);
            o[0] = make_uint4(t.d[0], t.d[1], t.d[2], t.d[3]);
            o[1] = make_uint4((u32)t.l[0] | ((u32)t.l[1] << 16), (u32)t.l[2] | ((u32)t.l[3] << 16), t.far, t.cnt);
            flag = (first == x - 1 && x) || ((S.Z6[x >> 5] >> (x & 31)) & 1u);
            has = t.cnt != 0;
        }
        const u32 m = __ballot_sync(0xFFFFFFFFu, flag), me = __ballot_sync(0xFFFFFFFFu, has);
        if ((threadIdx.x & 31) == 0) {
            const u32 sh = x & 31;
            if (m) { atomicOr(&S.B6[x >> 5], m << sh); if (sh) atomicOr(&S.B6[(x >> 5) + 1], m >> (32 - sh)); }
            if (me) { atomicOr(&S.E6[x >> 5], me << sh); if (sh) atomicOr(&S.E6[(x >> 5) + 1], me >> (32 - sh)); }
        }
    }
}

// ------------------------------------------------------------------ result packing
// The RC / BC sub-streams of all streams of a wave are gathered into one dense buffer each, so that the
// download is two large copies instead of two small ones per stream.  Offsets and sizes are 16-byte aligned.
struct PackDesc { u32 rc_src, rc_dst, rc_n, bc_src, bc_dst, bc_n, pad0, pad1; };
__global__ void __launch_bounds__(256)
k_pack_results(const u8 *__restrict__ rc, const u8 *__restrict__ bc, u8 *__restrict__ rc_out, u8 *__restrict__ bc_out,
               const PackDesc *__restrict__ desc) {
    const PackDesc d = desc[blockIdx.x];
    const u32 part = blockIdx.y, parts = gridDim.y;
    {
        const uint4 *s = reinterpret_cast<const uint4 *>(rc + d.rc_src);
        uint4 *o = reinterpret_cast<uint4 *>(rc_out + d.rc_dst);
        const u32 nv = (d.rc_n + 15) / 16;
        for (u32 i = part * 256 + threadIdx.x; i < nv; i += parts * 256) o[i] = s[i];
    }
    {
        const uint4 *s = reinterpret_cast<const uint4 *>(bc + d.bc_src);
        uint4 *o = reinterpret_cast<uint4 *>(bc_out + d.bc_dst);
        const u32 nv = (d.bc_n + 15) / 16;
        for (u32 i = part * 256 + threadIdx.x; i < nv; i += parts * 256) o[i] = s[i];
    }
}

// ------------------------------------------------------------------ Adler-32 (csarc's file checksum)
// adler32() of the archiver (csa_adler32.cpp:63-128, zlib's) over many buffers at once.  A buffer is cut into
// 64 KiB tiles; a CTA reduces one tile to (A, B) = (sum of bytes, sum of (len - offset) * byte), both mod 65521,
// with coalesced 16-byte loads; k_adler32_combine then folds the tiles of a buffer in order:
// b += len * a + B_tile, a += A_tile (the running `a` carries the initial value 1).
static const u32 kAdlerTile = 65536, kAdlerBase = 65521;
struct AdlerTile { u64 off; u32 len; u32 buf; };   // byte offset inside the batch buffer, bytes, owning buffer
__global__ void __launch_bounds__(256)
k_adler32_tiles(const u8 *__restrict__ base, const AdlerTile *__restrict__ tiles, u32 *__restrict__ partA, u32 *__restrict__ partB) {
    __shared__ u64 sa[8], sb[8];
    const AdlerTile t = tiles[blockIdx.x];
    const u8 *p = base + t.off;
    u64 a = 0, w = 0;   // sum of bytes, sum of offset * byte
    const u32 mis = (u32)((uintptr_t)p & 15);
    const u32 head = mis ? min(16 - mis, t.len) : 0;
    if (threadIdx.x < head) { const u32 b = p[threadIdx.x]; a += b; w += (u64)threadIdx.x * b; }
    const u32 nv = (t.len - head) / 16;
    const uint4 *v = reinterpret_cast<const uint4 *>(p + head);
    for (u32 i = threadIdx.x; i < nv; i += 256) {
        const uint4 q = v[i];
        const u32 o = head + i * 16;
        const u32 ws[4] = {q.x, q.y, q.z, q.w};
        u32 s = 0, sw = 0;   // bytes of this vector, sum of (index in vector) * byte
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int j = 0; j < 4; j++) { const u32 b = (ws[k] >> (8 * j)) & 0xFF; s += b; sw += (4 * k + j) * b; }
        a += s;
        w += (u64)o * s + sw;
    }
    const u32 tail0 = head + nv * 16;
    if (tail0 + threadIdx.x < t.len) { const u32 b = p[tail0 + threadIdx.x]; a += b; w += (u64)(tail0 + threadIdx.x) * b; }
#pragma unroll
    for (int sft = 16; sft; sft >>= 1) { a += __shfl_xor_sync(0xFFFFFFFFu, a, sft); w += __shfl_xor_sync(0xFFFFFFFFu, w, sft); }
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = w; }
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 A = 0, W = 0;
        for (int k = 0; k < 8; k++) { A += sa[k]; W += sb[k]; }
        // B_tile = sum (len - offset) * byte = len * A - W   (A <= 2^24, W <= 2^40: no overflow)
        partA[blockIdx.x] = (u32)(A % kAdlerBase);
        partB[blockIdx.x] = (u32)(((u64)t.len * A - W) % kAdlerBase);
    }
}
__global__ void k_adler32_combine(const AdlerTile *__restrict__ tiles, const u32 *__restrict__ partA, const u32 *__restrict__ partB,
                                  const u32 *__restrict__ first_tile, u32 nbuf, const u32 *__restrict__ init, u32 *__restrict__ out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbuf) return;
    const u32 v0 = init ? init[i] : 1u;
    u64 a = v0 & 0xFFFF, b = (v0 >> 16) & 0xFFFF;
    for (u32 t = first_tile[i]; t < first_tile[i + 1]; t++) {
        b = (b + (u64)(tiles[t].len % kAdlerBase) * a + partB[t]) % kAdlerBase;
        a = (a + partA[t]) % kAdlerBase;
    }
    out[i] = (u32)a | ((u32)b << 16);
}

// ------------------------------------------------------------------ framing (csc_memio.cpp:83-108) on the device
// frame.cuh: the finished libcsc stream of every chunk, so that the download is final bytes and the host only copies.
__global__ void k_frame_sizes(const FrameStream *__restrict__ fs, u32 ns, u32 *__restrict__ sizes) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ns) sizes[i] = frame_size(fs[i]);
}
__global__ void __launch_bounds__(128)
k_frame(const FrameStream *__restrict__ fs, u32 ns, const u64 *__restrict__ offs, u8 *__restrict__ out) {
    const u32 s = blockIdx.x * 4 + (threadIdx.x >> 5);   // one warp per stream
    if (s >= ns) return;
    frame_write(fs[s], out + offs[s]);
}

// ------------------------------------------------------------------ fill helper
__global__ void k_fill_u16(u16 *p, u64 n, u16 v) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// ------------------------------------------------------------------ serial engines
static const int kEncWarps = 4;
#ifndef CSC_ENC_MIN_CTAS
#define CSC_ENC_MIN_CTAS 5
#endif
__global__ void __launch_bounds__(kEncWarps * 32, CSC_ENC_MIN_CTAS)
k_encode(const EncStream *__restrict__ streams, u32 ns, u32 *queue, const u32 *__restrict__ p2bits) {
    __shared__ ModelSmall sm[kEncWarps];
#if CSC_TMA_STAGE
    __shared__ TStage tsm[kEncWarps];   // candidate records staged by the bulk-copy engine (tstage.cuh)
#endif
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ u32 s_p2bits[512];
    for (u32 i = threadIdx.x; i < 512; i += kEncWarps * 32) s_p2bits[i] = p2bits[i];
    if (lane == 0) sm[warp].p2bits = s_p2bits;
    __syncthreads();
#if CSC_TMA_STAGE
    ts_init(&tsm[warp]);
    TStage *const tsw = &tsm[warp];
#else
    TStage *const tsw = nullptr;
#endif
    __syncwarp();
    for (;;) {
        u32 s = 0;
        if (lane == 0) s = atomicAdd(queue, 1u);
        s = __shfl_sync(0xFFFFFFFFu, s, 0);
        if (s >= ns) break;
        EncStream S = streams[s];
        Encoder E(S, sm[warp], tsw);
        E.run();
        __syncwarp();
    }
}

// The range-coder pass: one stream per warp (rc_engine.cuh), all lanes redundant; the pass is a serial dependency
// chain per stream, so the warps only have to be spread over all SMs.
static const int kRcWarps = 4;
__global__ void __launch_bounds__(kRcWarps * 32)
k_rc(const EncStream *__restrict__ streams, u32 ns) {
    __shared__ uint4 stage[kRcWarps][32];
    const u32 warp = threadIdx.x >> 5;
    const u32 s = blockIdx.x * kRcWarps + warp;
    if (s >= ns) return;
    RcCoder R;
    R.run(streams[s], stage[warp]);
}

static const int kDecWarps = 4;
__global__ void __launch_bounds__(kDecWarps * 32)
k_decode(const DecStream *__restrict__ streams, u32 ns, u32 *queue) {
    __shared__ ModelSmall sm[kDecWarps];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        u32 s = 0;
        if (lane == 0) s = atomicAdd(queue, 1u);
        s = __shfl_sync(0xFFFFFFFFu, s, 0);
        if (s >= ns) break;
        DecStream S = streams[s];
        Decoder D(S, sm[warp]);
        D.run();
        __syncwarp();
    }
}

// dictionary inverse (csc_filters.cpp:337-369), one CTA per run.  A byte starts a token unless it is the second
// byte of an escape pair (254, x >= 0x82); whether a 254 in front of it is itself a token start depends on the
// parity of the run of 254s it ends, so every thread re-synchronises with a short look-back (pathological runs
// fall back to the scalar statement).  Token output lengths -> block scan -> every thread writes its tokens; the
// output is the first `size` bytes of the expansion (the forward transform pads its tail with spaces).
__device__ void dict_inverse_cta(const u8 *src, u32 size, u8 *dst, const DictTables *Dg, u8 *s, u32 *s_warp) {
    __shared__ u32 s_base, s_fallback;
    if (threadIdx.x == 0) { s_base = 0; s_fallback = 0; }
    __syncthreads();
    const u32 max_symbol = Dg->max_symbol;
    for (u32 t0 = 0; t0 < size; t0 += kFltTile) {
        const u32 tn = min((u32)kFltTile, size - t0);
        for (u32 i = threadIdx.x; i < kFltLook + tn + 8; i += kFltThreads) {
            long long g = (long long)t0 + i - kFltLook;
            s[i] = (g >= 0 && g < (long long)size) ? src[g] : (u8)0;
        }
        __syncthreads();
        const u8 *w = s + kFltLook;   // w[i] = src[t0 + i]
        const u32 a = threadIdx.x * kFltSpan, b = min(a + kFltSpan, tn);
        u32 pos = a, cnt = 0;
        if (a < tn) {
            // is w[a] swallowed by an escape?  count the 254s right in front of it
            if (w[a] >= 0x82) {
                int k = 0;
                while (k < (int)kFltLook && (long long)t0 + a - 1 - k >= 0 && w[(int)a - 1 - k] == 254) k++;
                if (k == (int)kFltLook && (long long)t0 + a - 1 - k >= 0) s_fallback = 1;
                if (k & 1) pos = a + 1;
            }
            for (u32 i = pos; i < b;) {
                const u32 c = w[i];
                if (c >= 0x82 && c < max_symbol) { const u8 *wd = Dg->word[c]; u32 l = 0; while (wd[l]) l++; cnt += l; i++; }
                else if (c == 254 && t0 + i + 1 < size && w[i + 1] >= 0x82) { cnt += 1; i += 2; }
                else { cnt += 1; i++; }
            }
        }
        u32 tile_total;
        u32 off = block_exscan(cnt, s_warp, &tile_total);
        if (a < tn && !s_fallback) {
            u32 o = s_base + off;
            for (u32 i = pos; i < b && o < size;) {
                const u32 c = w[i];
                if (c >= 0x82 && c < max_symbol) {
                    const u8 *wd = Dg->word[c];
                    for (u32 j = 0; wd[j] && o < size; j++) dst[o++] = wd[j];
                    i++;
                } else if (c == 254 && t0 + i + 1 < size && w[i + 1] >= 0x82) { dst[o++] = w[i + 1]; i += 2; }
                else { dst[o++] = (u8)c; i++; }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += tile_total;
        __syncthreads();
        if (s_fallback) break;
    }
    if (s_fallback) {   // exactness first
        if (threadIdx.x == 0) dict_inverse(src, size, dst, *Dg);
        __syncthreads();
        return;
    }
    for (u32 o = s_base + threadIdx.x; o < size; o += kFltThreads) dst[o] = 0;   // (an expansion shorter than the run: csc_filters.cpp:367)
}

// inverse filters: one CTA per stream walks the run list the decoder left; `out` already holds a copy of the window
__global__ void __launch_bounds__(kFltThreads)
k_unfilter(const DecStream *__restrict__ streams, u32 ns, u8 *const *__restrict__ outs, const DictTables *__restrict__ D) {
    __shared__ u8 s[kFltLook + kFltTile + 16];
    __shared__ u32 s_warp[kFltThreads / 32];
    const u32 sid = blockIdx.x;
    if (sid >= ns) return;
    const DecStream &S = streams[sid];
    if (*S.status < 0) return;
    const u32 nr = *S.n_runs;
    u8 *out = outs[sid];
    for (u32 r = 0; r < nr; r++) {
        const DecRun rn = S.runs[r];
        if (rn.type == DT_EXE) e8e9_run_cta(out + rn.off, rn.size, true, s, s_warp);
        else if (rn.type == DT_ENGTXT) dict_inverse_cta(S.wnd + rn.off, rn.size, out + rn.off, D, s, s_warp);
        __syncthreads();
    }
}

}  // namespace cscb
