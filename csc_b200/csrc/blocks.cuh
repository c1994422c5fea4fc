// blocks.cuh -- per-block / per-run scalar formulations of the analyzer and the
// filters.  These are the reference semantics stated once, as plain functions
// usable from host and device; the CUDA kernels in kernels.cu are the parallel
// versions and are checked against these and against the oracle.
//
//   analyze_block   Analyzer::Analyze + get_channel_idx + GetDltBpb
//                   (csc_analyzer.cpp:122-239)
//   e8e9_run        Filters::Forward_E89 / Inverse_E89 (csc_filters.cpp:508-609),
//                   restated over opcode positions instead of the byte shift register
//   dict_forward / dict_inverse   Filters::Foward_Dict / Inverse_Dict (:256-369)
//   delta_inverse   Filters::Inverse_Delta (:371-398)
#pragma once
#include "csc_defs.h"

namespace cscb {

// Tables derived from host libm exactly as the reference derives them (H5):
// built once on the host (host_tables.h) and uploaded.
struct AnalyzerTables {
    u32 logtab[(kBlock >> 4) + 1];   // logTable, csc_analyzer.cpp:9-16
    double ent_thr[16];              // (log2(d-2) - 0.6) * 100.0 for diffNum d (:223)
};

// Dictionary trie (csc_filters.cpp:87-111): node 0 is the root.
struct DictTables {
    u16 next[300][26];
    u8 symbol[300];
    u8 word[256][8];    // word text per symbol (0-terminated)
    u32 max_symbol;     // first byte value that is NOT a word symbol
};

CSC_HD u32 absdiff(u32 a, u32 b) { return a > b ? a - b : b - a; }

// entropy*100 in the reference's fixed point from a 256-bin histogram
CSC_HD u32 entropy_x100(const u32 *freq, u32 size, const u32 *logtab) {
    u32 e = size * logtab[size >> 4];
    for (u32 i = 0; i < 256; i++) e -= freq[i] * logtab[freq[i] >> 4];
    return e;
}

// GetDltBpb (csc_analyzer.cpp:166-182): order-0 cost of the channel-delta residual
CSC_HD u32 dlt_bpb_block(const u8 *src, u32 size, u32 chn, const u32 *logtab) {
    u32 freq[256];
    for (u32 i = 0; i < 256; i++) freq[i] = 0;
    u8 prev = 0;
    for (u32 c = 0; c < chn; c++)
        for (u32 j = c; j < size; j += chn) {
            freq[(u8)(src[j] - prev)]++;
            prev = src[j];
        }
    return entropy_x100(freq, size, logtab) / size;
}

// channel detection from the 5-lag equality / abs-difference statistics (csc_analyzer.cpp:122-164)
CSC_HD i32 channel_from_stats(const u32 *same, const u32 *succ, u32 size) {
    u32 max_same = same[0], min_same = same[0], max_succ = succ[0], min_succ = succ[0], best = 0;
    for (u32 i = 0; i < DLT_CHANNEL_MAX; i++) {
        if (same[i] < min_same) min_same = same[i];
        if (same[i] > max_same) max_same = same[i];
        if (succ[i] > max_succ) max_succ = succ[i];
        if (succ[i] < min_succ) { min_succ = succ[i]; best = i; }
    }
    if (((max_succ > succ[best] * 4) || (max_succ > succ[best] + 40 * size)) && (same[best] > min_same * 3) &&
        ((double)same[0] < 0.3 * (double)size))
        return (i32)best;
    return -1;
}

// type decision from the histogram statistics (csc_analyzer.cpp:190-238); need_chn is set
// when the channel statistics have to be consulted (returns DT_NONE in that case).
// diff = distinct bytes, hi = mass of bytes >= 0x80, alpha = mass of 'a'..'z'.
CSC_HD u32 type_from_stats(const u32 *freq, u32 size, u32 entropy, u32 diff, u32 hi, u32 alpha,
                           const AnalyzerTables &T, bool *need_chn) {
    u32 avg = size >> 8;
    *need_chn = false;
    if (hi < (size >> 3) && (freq[' '] + freq['\n'] + freq[':'] + freq['.'] + freq['/'] > (size >> 4)) &&
        (freq['a'] + freq['e'] + freq['t'] > (size >> 4)) && entropy > 300 * size && alpha > (size / 3))
        return DT_ENGTXT;
    if (freq[0x8b] > avg && freq[0x00] > avg * 2 && freq[0xE8] > 6) return DT_EXE;
    if (diff < 16 && diff >= 6 && (double)entropy > T.ent_thr[diff] * (double)size) return DT_ENTROPY;
    if (entropy < 400 * size && diff < 200) return DT_NORMAL;
    *need_chn = true;
    return DT_NONE;
}
CSC_HD u32 type_from_hist(const u32 *freq, u32 size, u32 entropy, const AnalyzerTables &T, bool *need_chn) {
    u32 diff = 0, hi = 0, alpha = 0;
    for (u32 i = 0; i < 256; i++) diff += (freq[i] > 0);
    for (u32 i = 128; i < 256; i++) hi += freq[i];
    for (u32 i = 'a'; i <= 'z'; i++) alpha += freq[i];
    return type_from_stats(freq, size, entropy, diff, hi, alpha, T, need_chn);
}
CSC_HD u32 type_after_chn(i32 chn_idx, u32 entropy, u32 size) {
    if (chn_idx != -1) return DT_DLT + (u32)chn_idx;
    if (entropy > 795 * size) return DT_BAD;
    if (entropy > 780 * size) return DT_FAST;
    return DT_NORMAL;
}

// Full scalar analysis of one block.  info = type | dlt_fail << 8 where dlt_fail is the
// GetDltBpb(...) >= bpb * 0.95 veto of csc_encoder_main.cpp:118-121 for the block's own type.
// For blocks shorter than 512 bytes (DT_SKIP) tail_dlt[0..4] receive GetDltBpb for all five
// channel counts, because the inherited type is only known to the serial engine.
CSC_HD void analyze_block(const u8 *src, u32 size, const AnalyzerTables &T, u32 *info, u32 *bpb, u32 *tail_dlt) {
    if (size > kBlock) size = kBlock;
    if (size < 512) {
        *info = DT_SKIP;
        *bpb = 0;
        if (tail_dlt && size)
            for (u32 k = 0; k < DLT_CHANNEL_MAX; k++) tail_dlt[k] = dlt_bpb_block(src, size, dlt_channels(k), T.logtab);
        return;
    }
    u32 freq[256];
    for (u32 i = 0; i < 256; i++) freq[i] = 0;
    for (u32 i = 0; i < size; i++) freq[src[i]]++;
    u32 entropy = entropy_x100(freq, size, T.logtab);
    *bpb = entropy / size;
    bool need_chn;
    u32 type = type_from_hist(freq, size, entropy, T, &need_chn);
    if (need_chn) {
        u32 same[DLT_CHANNEL_MAX] = {0, 0, 0, 0, 0}, succ[DLT_CHANNEL_MAX] = {0, 0, 0, 0, 0};
        for (u32 i = 0; i + 16 < size; i++) {
            u32 v = src[i];
            for (u32 k = 0; k < DLT_CHANNEL_MAX; k++) {
                u32 w = src[i + dlt_channels(k)];
                same[k] += (v == w);
                succ[k] += absdiff(v, w);
            }
        }
        type = type_after_chn(channel_from_stats(same, succ, size), entropy, size);
    }
    u32 fail = 0;
    if (type >= DT_DLT) {
        u32 d = dlt_bpb_block(src, size, dlt_channels(type - DT_DLT), T.logtab);
        fail = ((double)d >= (double)(*bpb) * 0.95) ? 1u : 0u;
    }
    *info = type | (fail << 8);
}

// ---- e8e9 ---------------------------------------------------------------------
CSC_HD u32 e89_xswap(u32 x) {   // E89xswap, csc_filters.cpp:527-531
    x <<= 7;
    return (x >> 24) | (((x >> 16) & 0xFF) << 8) | (((x >> 8) & 0xFF) << 16) | ((x & 0xFF) << 17);
}
CSC_HD u32 e89_yswap(u32 x) {   // E89yswap, csc_filters.cpp:533-537
    x = (((x >> 24) & 0xFF) << 7) | (((x >> 16) & 0xFF) << 8) | (((x >> 8) & 0xFF) << 16) | (x << 24);
    return x >> 7;
}
// rewrite the rel32 operand that follows the opcode at `op` (positions relative to run start)
CSC_HD void e89_apply(u8 *run, u32 op, bool inverse) {
    u32 x0 = (u32)run[op + 1] | ((u32)run[op + 2] << 8) | ((u32)run[op + 3] << 16) | ((u32)run[op + 4] << 24);
    u32 x = x0 - 0xFF000000u;
    if (x >= 0x02000000u) return;
    u32 i = op + 5;
    if (!inverse) { x = (x + i) & 0x01FFFFFFu; x = e89_xswap(x); }
    else { x = e89_yswap(x); x = (x - i) & 0x01FFFFFFu; }
    x0 = x + 0xFF000000u;
    run[op + 1] = (u8)x0; run[op + 2] = (u8)(x0 >> 8); run[op + 3] = (u8)(x0 >> 16); run[op + 4] = (u8)(x0 >> 24);
}
// Greedy left-to-right selection: an E8/E9 byte at op (op + 5 < size) is an instruction
// unless one of the three previous positions was one (k = i + 4, csc_filters.cpp:543-545).
CSC_HD void e8e9_run(u8 *run, u32 size, bool inverse) {
    if (size < 6) return;
    u32 next_ok = 0;
    for (u32 op = 0; op + 5 < size; op++) {
        if (op >= next_ok && (run[op] & 0xFE) == 0xE8) {
            next_ok = op + 4;
            e89_apply(run, op, inverse);
        }
    }
}

// ---- dictionary ---------------------------------------------------------------
// token starting at i: returns symbol (0 = none) and its length
CSC_HD u32 dict_longest(const u8 *src, u32 i, const DictTables &D, u32 *len) {
    u32 node = 0, sym = 0, best = 0;
    for (u32 j = 0;;) {
        u32 idx = (u32)src[i + j] - 'a';
        if (idx > 25) break;
        u32 nx = D.next[node][idx];
        if (!nx) break;
        node = nx;
        j++;
        if (D.symbol[node]) { sym = D.symbol[node]; best = j; }
    }
    *len = best;
    return sym;
}
// Foward_Dict (csc_filters.cpp:256-335).  dst must hold `size` bytes.  Returns 1 and leaves
// the transformed run in src, or 0 and leaves src untouched.
CSC_HD u32 dict_forward(u8 *src, u32 size, u8 *dst, const DictTables &D) {
    if (size < 16384) return 0;
    u32 out = 0, i = 0;
    while (i < size - 5) {
        if (out > size - 16) return 0;
        u32 c = src[i];
        if (c >= 'a' && c <= 'z') {
            u32 len, sym = dict_longest(src, i, D, &len);
            if (sym) { dst[out++] = (u8)sym; i += len; }
            else { dst[out++] = (u8)c; i++; }
        } else {
            if (c >= 0x82) dst[out++] = 254;
            dst[out++] = (u8)c;
            i++;
        }
    }
    for (; i < size; i++) {
        if (out + 2 > size) return 0;
        if (src[i] >= 0x82) dst[out++] = 254;
        dst[out++] = src[i];
    }
    if ((double)out > (double)size * 0.82) return 0;
    for (u32 k = 0; k < out; k++) src[k] = dst[k];
    for (u32 k = out; k < size; k++) src[k] = 0x20;
    return 1;
}
// Inverse_Dict (csc_filters.cpp:337-369): src (size bytes) -> dst (size bytes)
CSC_HD void dict_inverse(const u8 *src, u32 size, u8 *dst, const DictTables &D) {
    u32 i = 0, o = 0;
    while (o < size && i < size) {
        u32 c = src[i];
        if (c >= 0x82 && c < D.max_symbol) {
            const u8 *w = D.word[c];
            for (u32 j = 0; w[j] && o < size; j++) dst[o++] = w[j];
        } else if (c == 254 && i + 1 < size && src[i + 1] >= 0x82) {
            i++;
            dst[o++] = src[i];
        } else {
            dst[o++] = (u8)c;
        }
        i++;
    }
    while (o < size) dst[o++] = 0;
}

// ---- delta ----------------------------------------------------------------------
// Inverse_Delta (csc_filters.cpp:371-398): src = planar residuals, dst = interleaved bytes
CSC_HD void delta_inverse(const u8 *src, u32 size, u32 chn, u8 *dst) {
    if (size < 512) { for (u32 i = 0; i < size; i++) dst[i] = src[i]; return; }
    u32 p = 0; u8 prev = 0;
    for (u32 c = 0; c < chn; c++)
        for (u32 j = c; j < size; j += chn) {
            prev = (u8)(src[p++] + prev);
            dst[j] = prev;
        }
}
CSC_HD void delta_forward(const u8 *src, u32 size, u32 chn, u8 *dst) {   // Forward_Delta (:132-164)
    if (size < 512) { for (u32 i = 0; i < size; i++) dst[i] = src[i]; return; }
    u32 p = 0; u8 prev = 0;
    for (u32 c = 0; c < chn; c++)
        for (u32 j = c; j < size; j += chn) {
            dst[p++] = (u8)(src[j] - prev);
            prev = src[j];
        }
}

}  // namespace cscb
