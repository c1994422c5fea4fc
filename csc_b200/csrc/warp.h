// warp.h -- the few warp-level primitives the per-stream engines are written in.
//
// The serial part of libcsc (parser + adaptive coder) runs ONE
// STREAM PER WARP.  The engine code is warp-uniform: all 32 lanes execute the
// same scalar control flow on the same values (so no shuffles are needed to
// broadcast state), and the lanes fan out only inside these primitives:
// cooperative byte compares, lane-strided loops over bitmaps / positions, and
// ballots.  For host-side simulation in tests (tests/hostsim) the same code is
// compiled with a single lane.
#pragma once
#include "csc_defs.h"

namespace cscb {

#if defined(__CUDA_ARCH__)
#define CSC_LANE (threadIdx.x & 31u)
#define CSC_NLANES 32u
CSC_HDF void w_sync() { __syncwarp(); }
CSC_HDF bool w_any(bool p) { return __any_sync(0xFFFFFFFFu, p); }
// A warp-uniform condition routed through a vote: the compiler then knows the branch cannot diverge and
// drops the convergence-barrier bookkeeping (BSSY/BSYNC/BMOV) it would otherwise emit around it.
CSC_HDF bool w_uni(bool p) { return __any_sync(0xFFFFFFFFu, p); }
// same for a warp-uniform value: the result of a warp reduction lives in a uniform register
CSC_HDF u32 w_uni_u32(u32 v) { return __reduce_max_sync(0xFFFFFFFFu, v); }
CSC_HDF u32 w_ballot(bool p) { return __ballot_sync(0xFFFFFFFFu, p); }
CSC_HDF u32 w_min(u32 v) { return __reduce_min_sync(0xFFFFFFFFu, v); }
CSC_HDF u32 w_add(u32 v) { return __reduce_add_sync(0xFFFFFFFFu, v); }
CSC_HDF void g_or(u32 *p, u32 m) { atomicOr(p, m); }
CSC_HDF void g_max(u32 *p, u32 v) { atomicMax(p, v); }
CSC_HDF u32 g_ld(const u32 *p) { return __ldcg(p); }   // L2-coherent read of lane-written data
CSC_HDF u32 ffs32(u32 m) { return (u32)__ffs((int)m); }
CSC_HDF u32 popc32(u32 m) { return (u32)__popc(m); }
CSC_HDF void pf_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
CSC_HDF void pf_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }   // read-only arrays only
CSC_HDF void s_or(u32 *p, u32 m) { atomicOr(p, m); }   // shared-memory word, lanes may collide
CSC_HDF void s_max(u32 *p, u32 v) { atomicMax(p, v); }
#else
#define CSC_LANE 0u
#define CSC_NLANES 1u
CSC_HDF void w_sync() {}
CSC_HDF bool w_any(bool p) { return p; }
CSC_HDF bool w_uni(bool p) { return p; }
CSC_HDF u32 w_uni_u32(u32 v) { return v; }
CSC_HDF u32 w_ballot(bool p) { return p ? 1u : 0u; }
CSC_HDF u32 w_min(u32 v) { return v; }
CSC_HDF u32 w_add(u32 v) { return v; }
CSC_HDF void g_or(u32 *p, u32 m) { *p |= m; }
CSC_HDF void g_max(u32 *p, u32 v) { if (v > *p) *p = v; }
CSC_HDF u32 g_ld(const u32 *p) { return *p; }
CSC_HDF u32 ffs32(u32 m) { return (u32)__builtin_ffs((int)m); }
CSC_HDF u32 popc32(u32 m) { return (u32)__builtin_popcount(m); }
CSC_HDF void pf_l2(const void *) {}
CSC_HDF void pf_l1(const void *) {}
CSC_HDF void s_or(u32 *p, u32 m) { *p |= m; }
CSC_HDF void s_max(u32 *p, u32 v) { if (v > *p) *p = v; }
#endif

#if defined(__CUDA_ARCH__)
CSC_HDF u32 ld32u(const u8 *p) {   // unaligned little-endian u32 through two aligned loads (buffers are padded)
    const u32 *a = reinterpret_cast<const u32 *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    u32 sh = (u32)(reinterpret_cast<uintptr_t>(p) & 3) * 8;
    return __funnelshift_r(a[0], a[1], sh);
}
#else
CSC_HDF u32 ld32u(const u8 *p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
#endif

// ---- lane vectors ------------------------------------------------------------------------------------
// A few parts of the engine are written over 32 "virtual lanes" (the optimal parser's look-ahead window and
// its price table).  On the device element j of a lane vector is a register of lane j and the helpers are
// shuffles / ballots; the host simulation keeps the 32 elements in an array and loops, so the same source is
// validated against the reference on the CPU.
#if defined(__CUDA_ARCH__)
#define CSC_VL 1
#define LV_FOR(j) for (u32 j = CSC_LANE, j##_once = 1; j##_once; j##_once = 0)
#define LV_AT(j) 0
#else
#define CSC_VL 32
#define LV_FOR(j) for (u32 j = 0; j < 32; j++)
#define LV_AT(j) (j)
#endif
struct Lv { u32 v[CSC_VL]; };
CSC_HDF u32 lv_get(const Lv &a, u32 k) {   // element k, to every lane
#if defined(__CUDA_ARCH__)
    return __shfl_sync(0xFFFFFFFFu, a.v[0], k);
#else
    return a.v[k];
#endif
}
CSC_HDF void lv_shift_down(Lv &a, u32 fill) {   // a[j] = a[j + 1], a[31] = fill
#if defined(__CUDA_ARCH__)
    const u32 t = __shfl_down_sync(0xFFFFFFFFu, a.v[0], 1);
    a.v[0] = CSC_LANE == 31 ? fill : t;
#else
    for (u32 j = 0; j < 31; j++) a.v[j] = a.v[j + 1];
    a.v[31] = fill;
#endif
}
CSC_HDF u32 lv_ballot(const Lv &a) {   // bit j = (a[j] != 0)
#if defined(__CUDA_ARCH__)
    return __ballot_sync(0xFFFFFFFFu, a.v[0] != 0);
#else
    u32 m = 0;
    for (u32 j = 0; j < 32; j++) m |= (a.v[j] != 0 ? 1u : 0u) << j;
    return m;
#endif
}
CSC_HDF u32 lv_sum(const Lv &a) {
#if defined(__CUDA_ARCH__)
    return __reduce_add_sync(0xFFFFFFFFu, a.v[0]);
#else
    u32 s = 0;
    for (u32 j = 0; j < 32; j++) s += a.v[j];
    return s;
#endif
}

// Longest common prefix of a[0..limit) and b[0..limit): the byte-compare loop of
// every probe in find_match (csc_mf.cpp:279,317,348,470), done 128 bytes per
// step by the warp (4 bytes per lane).
CSC_HDN u32 lcp_bytes(const u8 *a, const u8 *b, u32 limit) {
#if defined(__CUDA_ARCH__)
    for (u32 k = 0; k < limit; k += 128) {
        u32 i = k + CSC_LANE * 4;
        u32 d = 0xFFFFFFFFu;
        if (i < limit) {
            d = ld32u(a + i) ^ ld32u(b + i);
            u32 rem = limit - i;   // bytes of this word that count
            if (rem < 4) d |= 0xFFFFFFFFu << (rem * 8);
        }
        u32 m = __ballot_sync(0xFFFFFFFFu, d != 0);
        if (m) {
            u32 l = (u32)__ffs((int)m) - 1;
            u32 dd = __shfl_sync(0xFFFFFFFFu, d, l);
            return k + l * 4 + (((u32)__ffs((int)dd) - 1) >> 3);
        }
    }
    return limit;
#else
    u32 l = 0;
    while (l < limit && a[l] == b[l]) l++;
    return l;
#endif
}

// same, but bytes of b at or beyond b_valid read as zero (window area that has
// not been written yet, csc_lz.cpp:50,67)
CSC_HD u32 lcp_bytes_zpad(const u8 *a, const u8 *b, u32 limit, u32 b_valid) {
    for (u32 k = 0; k < limit; k += CSC_NLANES) {
        u32 i = k + CSC_LANE;
        bool mis = true;
        if (i < limit) {
            u8 bv = (i < b_valid) ? b[i] : (u8)0;
            mis = (a[i] != bv);
        }
        u32 m = w_ballot(mis);
        if (m) return k + ffs32(m) - 1;
    }
    return limit;
}

CSC_HDF bool bit_get(const u32 *map, u32 i) { return (g_ld(&map[i >> 5]) >> (i & 31)) & 1u; }
CSC_HDF bool bit_get_ro(const u32 *map, u32 i) { return (map[i >> 5] >> (i & 31)) & 1u; }

// set bits [a, b) of a bitmap (lane-strided over words)
CSC_HDN void bits_set_range(u32 *map, u32 a, u32 b) {
    if (a >= b) return;
    u32 w0 = a >> 5, w1 = (b - 1) >> 5;
    for (u32 w = w0 + CSC_LANE; w <= w1; w += CSC_NLANES) {
        u32 m = 0xFFFFFFFFu;
        if (w == w0) m &= 0xFFFFFFFFu << (a & 31);
        if (w == w1) m &= 0xFFFFFFFFu >> (31 - ((b - 1) & 31));
        g_or(&map[w], m);
    }
    w_sync();
}

// dst[a,b) |= ~src[a,b)
CSC_HDN void bits_or_not_range(u32 *dst, const u32 *src, u32 a, u32 b) {
    if (a >= b) return;
    u32 w0 = a >> 5, w1 = (b - 1) >> 5;
    for (u32 w = w0 + CSC_LANE; w <= w1; w += CSC_NLANES) {
        u32 m = 0xFFFFFFFFu;
        if (w == w0) m &= 0xFFFFFFFFu << (a & 31);
        if (w == w1) m &= 0xFFFFFFFFu >> (31 - ((b - 1) & 31));
        g_or(&dst[w], m & ~src[w]);
    }
    w_sync();
}

}  // namespace cscb
