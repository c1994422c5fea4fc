// tstage.cuh -- TMA staging of the precomputed candidate records (TRec) for the serial warp.
//
// The parser of a stream walks its positions front to back and reads the 32-byte TRec of the positions it
// stops at and of the position behind it (paired probe).  The records of a stream are one dense array, so they
// are brought into shared memory by the bulk-copy engine (cp.async.bulk global -> shared, completion on an
// mbarrier) 16 records = 512 bytes at a time, one chunk ahead of the chunk being parsed: the serial warp then
// reads them with shared-memory latency and issues no load instructions for them.  Ring of two chunks per
// warp: record p lives in slot p & 31.  The ring is kept this small on purpose: every kilobyte of shared
// memory per warp is 20 KB of L1 the SM's resident streams lose, and k_encode lives on its L1 hit rate (a
// 128-record ring measured 20 % slower end to end, profiles/r02_ab_tma.log).
#pragma once
#include "csc_defs.h"

namespace cscb {

static const u32 kTsChunk = 16;                       // records per bulk copy (512 bytes)
static const u32 kTsRing = 2 * kTsChunk;              // records per warp
static const u32 kTsBytes = kTsChunk * sizeof(TRec);

struct alignas(128) TStage {
    TRec ring[kTsRing];
    unsigned long long bar[2];   // one mbarrier per half
    u32 lo;                      // chunk index held by half (lo & 1); chunk lo + 1 is in the other half (kNone: nothing staged)
    u32 pend;                    // bit h: a copy into half h has been issued and not yet waited for
    u32 phase;                   // bit h: parity of the next completion of bar[h]
    u32 pad;
};

#if defined(__CUDACC__)
__device__ __forceinline__ u32 ts_saddr(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ts_init(TStage *ts) {
    if ((threadIdx.x & 31u) == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ts_saddr(&ts->bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ts_saddr(&ts->bar[1])));
        ts->lo = kNone; ts->pend = 0; ts->phase = 0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
}
// forget what is staged (a new stream); outstanding copies are drained first
__device__ __forceinline__ void ts_wait(TStage *ts, u32 half);
__device__ __forceinline__ void ts_reset(TStage *ts) {
    ts_wait(ts, 0);
    ts_wait(ts, 1);
    if ((threadIdx.x & 31u) == 0) ts->lo = kNone;
    __syncwarp();
}
// chunk c of the stream's record array -> half c & 1 (one elected lane issues; the warp has finished reading the half)
__device__ __forceinline__ void ts_issue(TStage *ts, const TRec *T, u32 c) {
    __syncwarp();
    if ((threadIdx.x & 31u) == 0) {
        const u32 half = c & 1u;
        const u32 bar = ts_saddr(&ts->bar[half]), dst = ts_saddr(&ts->ring[half * kTsChunk]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of the half are ordered before the async write
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kTsBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(T + (size_t)c * kTsChunk), "r"(kTsBytes), "r"(bar) : "memory");
        ts->pend |= 1u << half;
    }
    __syncwarp();
}
__device__ __forceinline__ void ts_wait(TStage *ts, u32 half) {
    if (!((ts->pend >> half) & 1u)) return;
    const u32 bar = ts_saddr(&ts->bar[half]), parity = (ts->phase >> half) & 1u;
    u32 done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
    __syncwarp();
    if ((threadIdx.x & 31u) == 0) { ts->pend &= ~(1u << half); ts->phase ^= 1u << half; }
    __syncwarp();
}
// make the records of positions x and x + 1 (the paired probe) readable from the ring
__device__ __forceinline__ void ts_ensure(TStage *ts, const TRec *T, u32 x) {
    const u32 c = x / kTsChunk, lo = ts->lo;
    if (c != lo) {
        if (lo != kNone && c == lo + 1) ts_issue(ts, T, c + 1);            // the half of chunk lo is free now
        else { ts_wait(ts, 0); ts_wait(ts, 1); ts_issue(ts, T, c); ts_issue(ts, T, c + 1); }
        if ((threadIdx.x & 31u) == 0) ts->lo = c;
        __syncwarp();
    }
    ts_wait(ts, c & 1u);
    if ((x & (kTsChunk - 1)) == kTsChunk - 1) ts_wait(ts, (c + 1) & 1u);
}
__device__ __forceinline__ const TRec *ts_rec(const TStage *ts, u32 p) { return &ts->ring[p & (kTsRing - 1)]; }
#endif

}  // namespace cscb
