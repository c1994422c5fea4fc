// radix_sort.cuh -- segmented LSD radix sort of (hash key u32, position u32) pairs, hand-written for
// the match-finder precomputation ("parallel hash-key sort").
//
// The pairs arrive grouped by stream and in position order inside a stream.  Every pass is a
// stable counting sort on one 8-bit digit of the HASH part of the key, done independently inside
// each stream's range: tiles never span two streams and the digit histogram is laid out as
// [stream][digit][tile-of-stream], so one exclusive scan over that array yields, for every
// (tile, digit), the destination of its first element inside the stream's own range.  The stream id
// therefore never has to be sorted NOR carried in the key (7 passes instead of 11 for the three hash
// kinds, 8 bytes per element instead of 12), and equal hashes keep their position order -- which is
// what turns the sorted array into hash chains.
//
// Per pass: rs_hist (4 B/elem read) -> 3 scan kernels over the histogram -> rs_scatter (8 B read +
// 8 B write per element; ranks by warp match_any, tile reordered in shared memory so that the global
// writes are coalesced runs).
#pragma once
#include <cuda_runtime.h>

#include "csc_defs.h"

namespace cscb {

static const u32 kRsTile = 4096;     // elements per tile
static const u32 kRsThreads = 256;   // 8 warps, 512 elements each
static const u32 kRsPerThread = kRsTile / kRsThreads;

typedef u32 RsKey;   // hash value; bit 31 flags an element that is left out of its chain (sorts last)
struct RsTile {
    u32 start;    // first element (compact index)
    u32 count;    // <= kRsTile
    u32 hbase;    // histogram index of (this tile, digit 0)
    u32 nts;      // tiles of this tile's stream = histogram stride between digits
    u32 stream;   // the stream all elements of the tile belong to
    u32 sstart;   // first element of that stream
};

__global__ void __launch_bounds__(kRsThreads)
rs_hist(const RsKey *__restrict__ keys, const RsTile *__restrict__ tiles, u32 shift, u32 mask, u32 *__restrict__ H) {
    __shared__ u32 h[256];
    const RsTile t = tiles[blockIdx.x];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (u32 i = threadIdx.x; i < t.count; i += kRsThreads) atomicAdd(&h[(keys[t.start + i] >> shift) & mask], 1u);
    __syncthreads();
    H[t.hbase + threadIdx.x * t.nts] = h[threadIdx.x];
}

// ---- exclusive scan of H (n entries) in three small kernels -------------------------------------
static const u32 kScanChunk = 4096;
__device__ __forceinline__ u32 block_exscan_256(u32 v, u32 *s_w, u32 *total) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (u32)d) inc += t; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    u32 base = 0, tot = 0;
#pragma unroll
    for (u32 w = 0; w < kRsThreads / 32; w++) { u32 x = s_w[w]; if (w < warp) base += x; tot += x; }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}
__global__ void __launch_bounds__(kRsThreads) rs_scan_a(const u32 *__restrict__ H, u32 n, u32 *__restrict__ bsum) {
    __shared__ u32 s_w[kRsThreads / 32];
    const u32 base = blockIdx.x * kScanChunk;
    u32 v = 0;
    for (u32 i = threadIdx.x; i < kScanChunk; i += kRsThreads) if (base + i < n) v += H[base + i];
    u32 tot;
    block_exscan_256(v, s_w, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(kRsThreads) rs_scan_b(u32 *bsum, u32 nb) {
    __shared__ u32 s_w[kRsThreads / 32];
    __shared__ u32 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u32 b0 = 0; b0 < nb; b0 += kRsThreads) {
        u32 i = b0 + threadIdx.x;
        u32 v = i < nb ? bsum[i] : 0, tot;
        u32 ex = block_exscan_256(v, s_w, &tot);
        if (i < nb) bsum[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kRsThreads) rs_scan_c(u32 *H, u32 n, const u32 *__restrict__ bsum) {
    __shared__ u32 s_w[kRsThreads / 32];
    const u32 base = blockIdx.x * kScanChunk + threadIdx.x * (kScanChunk / kRsThreads);
    u32 v[kScanChunk / kRsThreads], sum = 0;
#pragma unroll
    for (u32 j = 0; j < kScanChunk / kRsThreads; j++) { v[j] = base + j < n ? H[base + j] : 0; sum += v[j]; }
    u32 tot;
    u32 run = bsum[blockIdx.x] + block_exscan_256(sum, s_w, &tot);
#pragma unroll
    for (u32 j = 0; j < kScanChunk / kRsThreads; j++) { if (base + j < n) H[base + j] = run; run += v[j]; }
}

// ---- stable scatter of one tile -------------------------------------------------------------------
// Register-lean on purpose (4 CTAs per SM): phase 1 keeps only the 16 ranks of a thread; the keys and
// positions are read again in phase 3, when they still sit in L1/L2.
__global__ void __launch_bounds__(kRsThreads, 4)
rs_scatter(const RsKey *__restrict__ kin, const u32 *__restrict__ vin, RsKey *__restrict__ kout, u32 *__restrict__ vout,
           const RsTile *__restrict__ tiles, u32 shift, u32 mask, const u32 *__restrict__ Hs) {
    extern __shared__ __align__(16) unsigned char dyn[];
    RsKey *skey = reinterpret_cast<RsKey *>(dyn);                  // [kRsTile]
    u32 *sval = reinterpret_cast<u32 *>(dyn + kRsTile * 4);        // [kRsTile]
    __shared__ u16 wcnt[kRsThreads / 32][256];
    __shared__ u32 gofs[256], lbase[256], s_w[kRsThreads / 32];
    const RsTile t = tiles[blockIdx.x];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    gofs[threadIdx.x] = Hs[t.hbase + threadIdx.x * t.nts];
#pragma unroll
    for (u32 w = 0; w < kRsThreads / 32; w++) wcnt[w][threadIdx.x] = 0;
    __syncthreads();
    // phase 1: each warp ranks its 512 consecutive elements, 32 at a time, in order
    u32 rk[kRsPerThread];
    const u32 wbeg = warp * (kRsTile / (kRsThreads / 32));
    const RsKey *kt = kin + t.start;
    const u32 *vt = vin + t.start;
#pragma unroll
    for (u32 r = 0; r < kRsPerThread; r++) {
        const u32 i = wbeg + r * 32 + lane;
        const bool live = i < t.count;
        const u32 dg = live ? ((kt[i] >> shift) & mask) : 0xFFFFFFFFu;
        const u32 peers = __match_any_sync(0xFFFFFFFFu, dg);
        const u32 below = __popc(peers & ((1u << lane) - 1));
        u32 base = 0;
        if (live) base = wcnt[warp][dg];
        __syncwarp();
        if (live && below == 0) wcnt[warp][dg] = (u16)(base + __popc(peers));
        __syncwarp();
        rk[r] = base + below;
    }
    __syncthreads();
    // phase 2: per digit, offsets of the warps inside the tile and of the digit inside the tile
    u32 run = 0;
#pragma unroll
    for (u32 w = 0; w < kRsThreads / 32; w++) { u32 c = wcnt[w][threadIdx.x]; wcnt[w][threadIdx.x] = (u16)run; run += c; }
    u32 tot;
    lbase[threadIdx.x] = block_exscan_256(run, s_w, &tot);
    __syncthreads();
    // phase 3: reorder the tile in shared memory
#pragma unroll
    for (u32 r = 0; r < kRsPerThread; r++) {
        const u32 i = wbeg + r * 32 + lane;
        if (i < t.count) {
            const RsKey kk = kt[i];
            const u32 dg = (kk >> shift) & mask;
            const u32 p = lbase[dg] + wcnt[warp][dg] + rk[r];
            skey[p] = kk;
            sval[p] = vt[i];
        }
    }
    __syncthreads();
    // phase 4: coalesced runs to the digit's destination inside the stream range
    for (u32 i = threadIdx.x; i < t.count; i += kRsThreads) {
        const RsKey kk = skey[i];
        const u32 dg = (kk >> shift) & mask;
        const u32 dst = gofs[dg] + (i - lbase[dg]);
        kout[dst] = kk;
        vout[dst] = sval[i];
    }
}

}  // namespace cscb
