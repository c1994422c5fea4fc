// plan.h -- static part of the block typing (csc_encoder_main.cpp:95-126) and the sub-block grid of
// the window copy (csc_lz.cpp:61-100).
//
// The content of the LZ window only depends on which runs get the e8e9 or the
// dictionary filter.  EXE and ENGTXT blocks are never re-typed by the
// duplicate-block test (it only applies to types >= DT_NO_LZ), so those runs
// are known before the serial engine starts and the filters can run as
// parallel kernels in front of the match finder.  The one dynamic input of the typing, the
// duplicate-block test, enters as a per-block assumption that the serial engine verifies.
#pragma once
#include "csc_defs.h"

namespace cscb {

struct FilterRun { u32 kind; u32 off; u32 size; u32 first_blk; };   // kind: DT_EXE or DT_ENGTXT; off relative to stream

// One sub-block of the window copy (csc_lz.cpp:63-67): at most 8 KiB, never across a run end or the
// window end.  These are the units the match-finder kernels hash / extend over.
struct SubBlock { u32 x; u32 size; u32 flags; };   // x relative to the stream; flags bit 0: left out of the 2-/3-byte chains, bit 1: text run

// Type of one block BEFORE the duplicate-block test (csc_encoder_main.cpp:98-121): analyzer type, DT_SKIP
// inheritance, disabled filters, delta-bpb veto.  `last` is the final type of the previous block of the
// segment (DT_NORMAL at the segment start), `bpb` the analyzer's persistent bits-per-byte value (:90,193).
CSC_HD u32 static_type(const u32 *binfo, const u32 *bbpb, const u32 *seg_tail_dlt, u32 blk, u32 last, u32 &bpb,
                       u32 flt_dlt, u32 flt_txt, u32 flt_exe) {
    if (flt_dlt + flt_txt + flt_exe == 0) return DT_NORMAL;
    u32 t = binfo[blk] & 0xFF;
    bool inherited = false;
    if (t != DT_SKIP) bpb = bbpb[blk];
    else { t = last; inherited = true; }
    if (t != DT_NORMAL) {
        if (t == DT_EXE && !flt_exe) t = DT_NORMAL;
        else if (t == DT_ENGTXT && !flt_txt) t = DT_NORMAL;
        else if (t >= DT_DLT && !flt_dlt) t = DT_NORMAL;
    }
    if (t >= DT_DLT) {   // GetDltBpb(...) >= bpb * 0.95 (csc_encoder_main.cpp:118-121)
        bool fail;
        if (inherited) fail = (double)seg_tail_dlt[t - DT_DLT] >= (double)bpb * 0.95;
        else fail = ((binfo[blk] >> 8) & 1) != 0;
        if (fail) t = DT_NORMAL;
    }
    return t;
}

// Static plan of one segment (= one Compress() call): the typing loop of csc_encoder_main.cpp:95-135 with
// the duplicate-block test replaced by an ASSUMPTION per block (flip[blk] != 0: the test will re-type the
// block to DT_NORMAL; null = no block flips).  The serial engine verifies the assumption where it matters
// and asks for a re-plan otherwise (kErrReplan).  Outputs:
//   runs[]   EXE / ENGTXT runs (their filters change the window content, so they run before the match finder)
//   nolz[b]  block b is not LZ-parsed (BAD / ENTROPY / DLT): left out of the 2-/3-byte chains
//   subs[]   the sub-block grid of the window copy: every run is cut into <= 8 KiB pieces from its start,
//            and additionally at the window end (csc_lz.cpp:65-66,91-92), after which the 8 KiB grid
//            restarts.  Without a window wrap the grid is the analyzer's block grid.
// seg_x = offset of the segment in its stream, wnd = LZ window size.  Returns the number of runs; *nsubs
// receives the number of sub-blocks (at most blocks + window ends inside the segment).
CSC_HD u32 plan_segment(const u32 *binfo, const u32 *bbpb, const u32 *seg_tail_dlt, const u8 *flip, u32 blk0, u32 seg_x,
                        u32 seg_size, u32 wnd, u32 flt_dlt, u32 flt_txt, u32 flt_exe, FilterRun *runs, u8 *nolz,
                        SubBlock *subs, u32 *nsubs) {
    u32 nruns = 0, ns = 0;
    u32 last = DT_NORMAL, run_begin = 0, run_size = 0, run_blk = blk0, bpb = 0;
    u32 blk = blk0;
    for (u32 i = 0; i <= seg_size; blk++) {
        const bool at_end = i == seg_size;
        u32 cur = 0, t = DT_NONE;
        if (!at_end) {
            cur = seg_size - i < kBlock ? seg_size - i : kBlock;
            t = static_type(binfo, bbpb, seg_tail_dlt, blk, last, bpb, flt_dlt, flt_txt, flt_exe);
            if (t >= DT_NO_LZ && flip && flip[blk]) t = DT_NORMAL;
            nolz[blk] = t >= DT_NO_LZ ? 1 : 0;
        }
        if (at_end || t != last) {   // the pending run [run_begin, run_begin + run_size) of type `last` is complete
            if (run_size) {
                if (last == DT_EXE || last == DT_ENGTXT) {
                    runs[nruns].kind = last; runs[nruns].off = seg_x + run_begin; runs[nruns].size = run_size;
                    runs[nruns].first_blk = run_blk; nruns++;
                }
                const u32 fl = (last >= DT_NO_LZ ? 1u : 0u) | (last == DT_ENGTXT ? 2u : 0u);
                for (u32 off = 0; off < run_size;) {   // LZ::EncodeNormal's copy loop
                    const u32 x = seg_x + run_begin + off;
                    u32 sb = run_size - off < kBlock ? run_size - off : kBlock;
                    const u32 to_wrap = wnd - x % wnd;
                    if (to_wrap < sb) sb = to_wrap;
                    subs[ns].x = x; subs[ns].size = sb; subs[ns].flags = fl; ns++;
                    off += sb;
                }
            }
            run_begin = i; run_size = 0; run_blk = blk;
        }
        if (at_end) break;
        last = t;
        run_size += cur;
        i += cur;
    }
    *nsubs = ns;
    return nruns;
}

// number of window ends strictly inside (a, a + size): each can add one sub-block to the block grid
CSC_HD u32 wraps_inside(u32 a, u32 size, u32 wnd) {
    if (!size) return 0;
    const u64 e = (u64)a + size;
    return (u32)((e - 1) / wnd - (u64)a / wnd);
}

// Byte y of the window as the reference's hashes see it while the sub-block ending at sb_end is being
// parsed: bytes below sb_end are the stream itself; beyond it the window still holds what
// the previous lap left there (the stream byte one window earlier), zeros on the first lap and in the 8
// pad bytes behind the window end (csc_lz.cpp:21,50).  lz = linear history of the stream.
CSC_HD u32 window_byte(const u8 *lz, u32 y, u32 sb_end, u32 wnd) {
    if (y < sb_end) return lz[y];
    const u64 lap_end = ((u64)(sb_end - 1) / wnd + 1) * wnd;
    if (y >= lap_end || y < wnd) return 0;
    return lz[y - wnd];
}

// 6-byte / 2-byte hashes with the sub-block padding rule: bytes at or beyond the end of the
// current <= 8 KiB sub-block have not been copied into the window yet and read as 0
// (csc_lz.cpp:50,67).  `avail` = bytes from p to the sub-block end.  (First-lap form; the
// *_at variants below take the window_byte rule and are what the kernels' staging reproduces.)
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

CSC_HD u32 hash6_at(const u8 *lz, u32 x, u32 sb_end, u32 wnd, u32 bits) {
    u32 b[6];
#pragma unroll
    for (u32 k = 0; k < 6; k++) b[k] = window_byte(lz, x + k, sb_end, wnd);
    u32 v = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
    u32 v2 = b[4] | (b[5] << 8);
    return ((v ^ (v2 << 13)) * 2654435761u) >> (32 - bits);
}
CSC_HD u32 hash3_at(const u8 *lz, u32 x, u32 sb_end, u32 wnd) {
    return (window_byte(lz, x, sb_end, wnd) << 8) ^ (window_byte(lz, x + 1, sb_end, wnd) << 5) ^ window_byte(lz, x + 2, sb_end, wnd);
}
CSC_HD u32 hash2_at(const u8 *lz, u32 x, u32 sb_end, u32 wnd) {
    return ((window_byte(lz, x, sb_end, wnd) | (window_byte(lz, x + 1, sb_end, wnd) << 8)) * 65521u) & 0x3FFF;
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
CSC_HD u32 hash_ctx_at(const u8 *lz, u32 x, u32 sb_end, u32 wnd, u32 len, u32 bits) {
    u32 w[3] = {0, 0, 0};
    for (u32 k = 0; k < len; k++) w[k >> 2] |= window_byte(lz, x + k, sb_end, wnd) << (8 * (k & 3));
    return hash_ctx_words(w[0], w[1], w[2], bits);
}

}  // namespace cscb
