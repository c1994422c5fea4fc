// plan.h -- static part of the block typing (csc_encoder_main.cpp:95-126).
//
// The content of the LZ window only depends on which runs get the e8e9 or the
// dictionary filter.  EXE and ENGTXT blocks are never re-typed by the
// duplicate-block test (it only applies to types >= DT_NO_LZ), so those runs
// are known before the serial engine starts and the filters can run as
// parallel kernels in front of the match finder.
#pragma once
#include "csc_defs.h"

namespace cscb {

struct FilterRun { u32 kind; u32 off; u32 size; u32 first_blk; };   // kind: DT_EXE or DT_ENGTXT; off relative to stream

// Walks the blocks of one segment and appends its EXE / ENGTXT runs.
// Returns the number of runs appended (runs must have room for nblk entries).
// nolz[blk] = 1 for blocks the analyzer statically types as not LZ-parsed (BAD / ENTROPY / DLT that pass the
// delta-bpb veto): they are left out of the 2-/3-byte chains.  A duplicate-block flip of such a block
// (csc_encoder_main.cpp:123-126) makes the engine fall back to real tables for the rest of the stream.
CSC_HD u32 plan_segment(const u32 *binfo, u32 blk0, u32 seg_off, u32 seg_size, u32 flt_dlt, u32 flt_txt, u32 flt_exe,
                        FilterRun *runs, u8 *nolz) {
    u32 nruns = 0;
    if (flt_dlt + flt_txt + flt_exe == 0) {
        u32 b = blk0;
        for (u32 i = 0; i < seg_size; i += kBlock, b++) nolz[b] = 0;
        return 0;
    }
    u32 last = DT_NORMAL, run_begin = 0, run_size = 0, run_blk = blk0;
    u32 blk = blk0;
    for (u32 i = 0; i < seg_size; blk++) {
        u32 cur = seg_size - i < kBlock ? seg_size - i : kBlock;
        u32 t = binfo[blk] & 0xFF;
        {
            u32 st = t;
            if (st >= DT_DLT && st != DT_SKIP && (!flt_dlt || ((binfo[blk] >> 8) & 1))) st = DT_NORMAL;
            nolz[blk] = (st != DT_SKIP && st >= DT_NO_LZ) ? 1 : 0;
        }
        if (t == DT_SKIP) t = last;
        if (t == DT_EXE && !flt_exe) t = DT_NORMAL;
        else if (t == DT_ENGTXT && !flt_txt) t = DT_NORMAL;
        if (t != last) {
            if (run_size && (last == DT_EXE || last == DT_ENGTXT)) {
                runs[nruns].kind = last; runs[nruns].off = seg_off + run_begin; runs[nruns].size = run_size;
                runs[nruns].first_blk = run_blk; nruns++;
            }
            run_begin = i; run_size = 0; run_blk = blk;
        }
        last = t;
        run_size += cur;
        i += cur;
    }
    if (run_size && (last == DT_EXE || last == DT_ENGTXT)) {
        runs[nruns].kind = last; runs[nruns].off = seg_off + run_begin; runs[nruns].size = run_size;
        runs[nruns].first_blk = run_blk; nruns++;
    }
    return nruns;
}

// 6-byte / 2-byte hashes with the sub-block padding rule: bytes at or beyond the end of the
// current <= 8 KiB sub-block have not been copied into the window yet and read as 0
// (csc_lz.cpp:50,67).  `avail` = bytes from p to the sub-block end.
CSC_HD u32 hash6_padded(const u8 *p, u32 avail, u32 bits) {   // HASH6, csc_mf.cpp:35-42
    u32 b[6];
#pragma unroll
    for (u32 k = 0; k < 6; k++) b[k] = k < avail ? p[k] : 0;
    u32 v = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
    u32 v2 = b[4] | (b[5] << 8);
    return ((v ^ (v2 << 13)) * 2654435761u) >> (32 - bits);
}
CSC_HD u32 hash3_padded(const u8 *p, u32 avail) {   // HASH3, csc_mf.cpp:30-33
    u32 b0 = p[0], b1 = avail > 1 ? p[1] : 0, b2 = avail > 2 ? p[2] : 0;
    return (b0 << 8) ^ (b1 << 5) ^ b2;
}
CSC_HD u32 hash2_padded(const u8 *p, u32 avail) {   // HASH2, csc_mf.cpp:23-28
    u32 b0 = p[0], b1 = avail > 1 ? p[1] : 0;
    return ((b0 | (b1 << 8)) * 65521u) & 0x3FFF;
}

// -m5 bounded search: hash of an 8- or 12-byte context (three little-endian words, w2 = 0 for 8 bytes), bytes
// beyond the sub-block end read as 0 like the reference hashes above
CSC_HD u32 hash_ctx_words(u32 w0, u32 w1, u32 w2, u32 bits) {
    u32 a = w0 * 0x9E3779B1u, b = w1 * 0x85EBCA77u, c = w2 * 0xC2B2AE3Du;
    u32 h = a ^ ((b << 15) | (b >> 17)) ^ ((c << 7) | (c >> 25));
    return (h * 0x27D4EB2Fu) >> (32 - bits);
}
CSC_HD u32 hash_ctx_padded(const u8 *p, u32 avail, u32 len, u32 bits) {
    u32 w[3] = {0, 0, 0};
    for (u32 k = 0; k < len; k++) w[k >> 2] |= (u32)(k < avail ? p[k] : 0) << (8 * (k & 3));
    return hash_ctx_words(w[0], w[1], w[2], bits);
}

}  // namespace cscb
