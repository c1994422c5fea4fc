// rc_engine.cuh -- the range-coder pass: one stream per WARP, warp-uniform.
//
// The reference's EncodeBit (csc_coder.h:67-81) adapts a probability and then advances the coder
// registers (range, 64-bit low with carry / cache bytes, csc_coder.cpp:89-112).  Only the second half is
// serially dependent on every earlier decision of the segment, and it needs nothing but the sequence of
// (probability, bit) pairs.  The parser / model engine (enc_engine.cuh) therefore only APPENDS those pairs
// to the stream's decision list; this pass replays them: about a dozen integer instructions per decision, no
// model memory, no match finder.
//
// The pass is a dependency chain per stream (latency, not throughput), so what counts is the length of that
// chain.  A thread-per-stream layout (round 2, first version) made every lane of a warp pay for the
// renormalisation / byte-output path of any lane: ~140 cycles per decision.  Here all 32 lanes of a warp run
// the same stream redundantly: branches are uniform and mostly not taken, the decision list is fetched 512
// bytes per load instruction by the whole warp and staged in shared memory one block ahead, and the chain is
// the ~25 cycles of the range recurrence itself.
//
// Output per segment: the RC sub-stream (whose last byte is the one the reference skips instead of
// storing, H3) and the time stamps of the 64 KiB block completions (csc_coder.cpp:97-104) for the host's
// replay of the RC/BC block order (H4).
#pragma once
#include "csc_defs.h"

namespace cscb {

struct RcCoder {
    // registers of Coder (csc_coder.h:44-50)
    u64 low;
    u32 range, cachesize, cache;
    // output
    u8 *out;
    u32 cap, n, blk, bsize;
    u32 *ev;
    u32 ev_cap, ev_n;
    i32 err;
    // bytes are gathered into one aligned 4-byte word before they go to memory
    u32 acc;

    // One output byte (cold path: pending 0xFF runs, flush).
    CSC_HDF void put(u32 b, u32 stamp) {
        const u32 sh = (n & 3) * 8;
        acc |= (b & 0xFF) << sh;
        if (sh == 24) {   // the aligned 4-byte word is complete
            if (n < cap) *reinterpret_cast<u32 *>(out + (n & ~3u)) = acc; else err = kErrOverflow;
            acc = 0;
        }
        n++;
        if (++blk == bsize) {   // a 64 KiB RC block completed (csc_coder.cpp:97-104)
            if (ev_n < ev_cap) ev[ev_n] = stamp; else err = kErrOverflow;
            ev_n++;
            blk = 0;
        }
    }
    CSC_HDF void drain() {   // pending bytes of the gather word
        const u32 base = n & ~3u;
        for (u32 k = 0; k < (n & 3); k++) { if (base + k < cap) out[base + k] = (u8)(acc >> (8 * k)); else err = kErrOverflow; }
    }
    CSC_HDF void shift_low(u32 stamp) {   // RC_ShiftLow, csc_coder.cpp:89-112 (cold path: flush)
        const u32 lo32 = (u32)low, carry = (u32)(low >> 32);
        if (lo32 < 0xFF000000u || carry) {
            put(cache + carry, stamp);
            while (cachesize > 1) { put(0xFF + carry, stamp); cachesize--; }   // pending 0xFF bytes
            cache = lo32 >> 24;
            cachesize = 0;
        }
        cachesize++;
        low = (u64)(lo32 << 8);
    }
    CSC_HDF void step(u32 e, u32 stamp) {   // the coder half of EncodeBit, csc_coder.h:67-81
        const u32 bound = (range >> 12) * (e & 0xFFFu);
        if (e & 0x8000u) range = bound;
        else { low += bound; range -= bound; }
        if (range < (1u << 24)) { range <<= 8; shift_low(stamp); }
    }

    // all segments of one stream; the decision lists were written by Encoder (enc_engine.cuh)
    CSC_HD void run(const EncStream &S, void *stage_mem) {   // stage_mem: 512 bytes of shared memory per warp (device)
        (void)stage_mem;
        out = S.rc_out; cap = S.rc_cap; bsize = S.cfg.csc_blocksize;
        ev = S.rc_ev; ev_cap = S.ev_cap; ev_n = 0;
        n = 0; err = 0; acc = 0;
        if (*S.status < 0) return;
        for (u32 sg = 0; sg < S.nseg; sg++) {
            SegOut so = S.segout[sg];
            low = 0; range = 0xFFFFFFFFu; cachesize = 1; cache = 0; blk = 0;   // Coder::Flush leaves the coder reset (:64-73)
            const u32 rc0 = n, ev0 = ev_n;
            const DecEntry *d = S.dec + so.dec_off;
            const u32 cnt = so.dec_len;
            u32 i = 0;
#if defined(__CUDA_ARCH__)
            // dec_off is 8-aligned: the warp fetches 256 decisions (512 bytes) per load, one block ahead of the
            // one being coded, and every lane reads them back 8 at a time from shared memory (broadcast)
            if (cnt >= 8) {
                const u32 lane = threadIdx.x & 31u, full = cnt & ~7u;
                const uint4 *g = reinterpret_cast<const uint4 *>(d);
                uint4 *stage = reinterpret_cast<uint4 *>(stage_mem);
                uint4 nx = make_uint4(0, 0, 0, 0);
                if (lane * 8 < full) nx = g[lane];
                for (u32 base = 0; base < full; base += 256) {
                    __syncwarp();
                    stage[lane] = nx;
                    __syncwarp();
                    if (base + 256 + lane * 8 < full) nx = g[(base + 256) / 8 + lane];
                    const u32 m = full - base < 256 ? (full - base) / 8 : 32;
                    uint4 vn = stage[0];
#pragma unroll 1
                    for (u32 q = 0; q < m; q++) {
                        const uint4 v = vn;
                        vn = stage[(q + 1) & 31];   // the next eight are on their way while these are coded
                        const u32 w[4] = {v.x, v.y, v.z, v.w};
                        const u32 i0 = base + q * 8;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            step(w[k] & 0xFFFFu, 2 * (i0 + 2 * k) + 1);
                            step(w[k] >> 16, 2 * (i0 + 2 * k + 1) + 1);
                        }
                    }
                }
                i = full;
            }
#endif
            for (; i < cnt; i++) step(d[i], 2 * i + 1);
            // Coder::Flush (csc_coder.cpp:40-48): five more bytes, then one byte is skipped, not stored (H3):
            // the host substitutes the stale ring byte
            for (int k = 0; k < 5; k++) shift_low(2 * cnt + 1);
            if ((n & 3) == 3) { if (n < cap) *reinterpret_cast<u32 *>(out + (n & ~3u)) = acc; else err = kErrOverflow; acc = 0; }
            n++; blk++;
            so.rc_off = rc0; so.rc_len = n - rc0;
            so.rev_off = ev0; so.rev_len = ev_n - ev0;
            S.segout[sg] = so;
        }
        drain();
        if (err) *S.status = err;
    }
};

}  // namespace cscb
