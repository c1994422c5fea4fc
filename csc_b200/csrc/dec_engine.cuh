// dec_engine.cuh -- libcsc stream decoder, one stream per warp.
//
// Follows CSCDecoder (csc_dec.cpp): range decoder (:10-35), raw-bit reader
// (:64-87), packet grammar (:155-307), LZ copy (:476-571) and the per-run
// dispatch (:586-682).  The window is kept as a LINEAR history of the whole
// stream in HBM (180 GB per B200 makes the reference's ring unnecessary);
// ring semantics only survive as the reference's bounds checks.  Inverse
// e8e9 / dictionary transforms run afterwards as separate per-run kernels,
// driven by the run list this engine emits.
#pragma once
#include "csc_defs.h"
#include "warp.h"

namespace cscb {

struct Decoder {
    const DecStream &S;
    ModelSmall &M;
    u32 range, code;
    u32 rc_pos, rc_end, rc_blk;      // read cursor, end of current RC block, its index
    u32 rc_cur, rc_nxt;              // device: the aligned word under the cursor and the one behind it (already loaded)
    u32 rc_len, n_rc_blk;            // copies of the stream descriptor's fields (registers instead of descriptor loads)
    const u8 *rcp;
    u32 bc_pos, bc_end, bc_blk;
    u32 bc_bits, bc_val;
    u32 state, ctx;
    u32 rep[4];
    u32 x;                           // bytes decoded so far (window position, linear)
    i32 err;

    CSC_HD Decoder(const DecStream &s, ModelSmall &m) : S(s), M(m) {}

    // The RC payloads of a stream are one contiguous, 16-byte aligned and padded buffer: the device reads it a
    // word at a time and always has the next word loaded before the cursor gets there, so the byte a
    // renormalisation needs is never an L2 round trip on the decoder's dependency chain.
    CSC_HD void rc_seek(u32 pos) {
        rc_pos = pos;
#if defined(__CUDA_ARCH__)
        const u32 *w = reinterpret_cast<const u32 *>(rcp) + (pos >> 2);
        rc_cur = w[0]; rc_nxt = w[1];
#endif
    }
    CSC_HD u32 rc_byte() {
        if (rc_pos >= rc_len || rc_blk >= n_rc_blk) { if (!err) err = kErrRead; return 0; }
#if defined(__CUDA_ARCH__)
        const u32 b = (rc_cur >> ((rc_pos & 3) * 8)) & 0xFFu;
        rc_pos++;
        if ((rc_pos & 3) == 0) { rc_cur = rc_nxt; rc_nxt = reinterpret_cast<const u32 *>(rcp)[(rc_pos >> 2) + 1]; }
#else
        const u32 b = rcp[rc_pos++];
#endif
        if (rc_pos == rc_end) {   // block exhausted: the reference fetches the next one at once (:15-21)
            rc_blk++;
            rc_end = rc_blk < n_rc_blk ? S.rc_blk_off[rc_blk + 1] : rc_pos;
        }
        return b;
    }
    CSC_HD u32 bc_byte() {
        if (bc_pos >= S.bc_len || bc_blk >= S.n_bc_blk) { if (!err) err = kErrRead; return 0; }
        u32 b = S.bc[bc_pos++];
        if (bc_pos == bc_end) {
            bc_blk++;
            bc_end = bc_blk < S.n_bc_blk ? S.bc_blk_off[bc_blk + 1] : bc_pos;
        }
        return b;
    }
    // start of a flush segment: position both readers on a fresh block (:336-345, :657-680)
    CSC_HD void open_segment(bool first) {
        if (!first) { rc_blk++; bc_blk++; }
        if (rc_blk >= n_rc_blk || bc_blk >= S.n_bc_blk) { err = -1; return; }
        rc_pos = S.rc_blk_off[rc_blk]; rc_end = S.rc_blk_off[rc_blk + 1];
        bc_pos = S.bc_blk_off[bc_blk]; bc_end = S.bc_blk_off[bc_blk + 1];
        range = 0xFFFFFFFFu; bc_bits = 0; bc_val = 0;
        // byte 0 is the coder's initial cache byte; code = bytes 1..4
        if (rc_end - rc_pos < 5) { err = -1; return; }
        code = ((u32)rcp[rc_pos + 1] << 24) | ((u32)rcp[rc_pos + 2] << 16) | ((u32)rcp[rc_pos + 3] << 8) | rcp[rc_pos + 4];
        rc_seek(rc_pos + 5);
        if (rc_pos == rc_end) { rc_blk++; rc_end = rc_blk < n_rc_blk ? S.rc_blk_off[rc_blk + 1] : rc_pos; }
    }
    // after open_segment rc_blk may already have advanced when a block held exactly 5 bytes;
    // segment switches are relative to the block the segment *ended* in, which is what rc_blk tracks.

    CSC_HD u32 dec_bit(u16 &pr) {   // DecodeBit, csc_dec.cpp:10-35
        if (range < (1u << 24)) {
            range <<= 8;
            code = (code << 8) + rc_byte();
        }
        const u32 p = pr;
        const u32 bound = (range >> 12) * p;
        const bool one = code < bound;
        range = one ? bound : range - bound;
        code = one ? code : code - bound;
        pr = (u16)(one ? p + ((0xFFFu - p) >> 5) : p - (p >> 5));
        return one ? 1u : 0u;
    }
    CSC_HD u32 direct16(u32 len) {   // coder_decode_direct, csc_dec.cpp:64-87
        while (bc_bits < len) {
            bc_val = (bc_val << 8) | bc_byte();
            bc_bits += 8;
        }
        u32 r = (bc_val >> (bc_bits - len)) & ((1u << len) - 1);
        bc_bits -= len;
        return r;
    }
    // `size` whole bytes through the raw-bit reader: the payload shifted by the bits pending in front of it.
    // Lane-parallel per BC block; the block bookkeeping is that of bc_byte().
    CSC_HD void raw_bytes(u8 *dst, u32 size) {
        while (size && bc_bits >= 8) { *dst++ = (u8)direct16(8); size--; }   // (never: at most 7 bits stay pending)
        const u32 k = bc_bits;
        u32 pend = bc_val & ((1u << k) - 1);
        while (size) {
            if (bc_pos >= S.bc_len || bc_blk >= S.n_bc_blk || bc_end <= bc_pos) { if (!err) err = kErrRead; break; }
            const u32 take = size < bc_end - bc_pos ? size : bc_end - bc_pos;
            const u8 *src = S.bc + bc_pos;
            for (u32 i = CSC_LANE; i < take; i += CSC_NLANES) {
                const u32 v = src[i], p = i ? src[i - 1] : pend;
                dst[i] = (u8)((p << (8 - k)) | (v >> k));
            }
            pend = src[take - 1] & ((1u << k) - 1);
            dst += take; size -= take; bc_pos += take;
            if (bc_pos == bc_end) {
                bc_blk++;
                bc_end = bc_blk < S.n_bc_blk ? S.bc_blk_off[bc_blk + 1] : bc_pos;
            }
        }
        bc_val = pend;
        w_sync();
    }
    CSC_HD u32 direct(u32 len) {
        if (len <= 16) return direct16(len);
        u32 hi = direct16(len - 16) << 16;
        return hi | direct16(16);
    }
    CSC_HD u32 decode_int() {   // csc_dec.cpp:89-96
        u32 slot = direct(5);
        u32 num = direct(slot == 0 ? 1 : slot);
        if (slot) num += (1u << slot);
        return num;
    }
    CSC_HD u32 tree(u16 *p, u32 top) {   // MSB-first bit tree until node >= top
        u32 i = 1;
        do { i = i + i + dec_bit(p[i]); } while (i < top);
        return i;
    }
    CSC_HD u32 decode_matchlen_1() {   // csc_dec.cpp:187-220
        if (dec_bit(M.p_matchlen_slot[0]) == 0) return tree(M.p_matchlen_extra1, 8) & 7;
        if (dec_bit(M.p_matchlen_slot[1]) == 0) return 8 + (tree(M.p_matchlen_extra2, 8) & 7);
        return 16 + (tree(M.p_matchlen_extra3, 0x80) & 0x7F);
    }
    CSC_HD u32 decode_matchlen_2() {   // csc_dec.cpp:222-232
        u32 len = decode_matchlen_1();
        if (len != 143) return len;
        while (dec_bit(M.p_longlen) == 0) { len += 143; if (err) return 0; }
        return len + decode_matchlen_1();
    }
    CSC_HD void decode_match(u32 &dist, u32 &len) {   // csc_dec.cpp:234-281
        len = decode_matchlen_2();
        u32 base, sbits;
        if (len == 0) { base = 0; sbits = 3; }
        else if (len <= 2) { base = 16 * (len - 1) + 8; sbits = 4; }
        else if (len <= 5) { base = 32 * (len - 3) + 8 + 32; sbits = 5; }
        else { base = 32 * 3 + 8 + 32; sbits = 5; }
        u32 slot = tree(M.p_dist + base, 1u << sbits) & ((1u << sbits) - 1);
        if (slot <= 2) dist = slot;
        else {
            u32 ebits = slot - 2, elen = 0;
            if (ebits > 4) elen = direct(ebits - 4);
            u32 low = tree(M.p_matchdist_extra + (ebits - 1) * 16, 0x10) & 0x0F;
            u32 rev = ((low & 1) << 3) | ((low & 2) << 1) | ((low & 4) >> 1) | ((low & 8) >> 3);
            dist = (1u << ebits) + 1 + (elen << 4) + rev;   // dist_table_[slot] = 2^(slot-2)+1
        }
        state = (state * 4 + 1) & 0x3F;
    }
    // 8 bits under one order-1 row (decode_literal, csc_dec.cpp:155-167).  On the device the 512-byte row is fetched
    // ONCE, 16 bytes per lane, instead of with eight dependent loads along the tree path; the probability of node i
    // then comes from lane i / 8 by shuffle and only that lane stores the adapted value.
    CSC_HD u32 decode_literal_sym(u16 *row) {
#if defined(__CUDA_ARCH__)
        const uint4 r = reinterpret_cast<const uint4 *>(row)[CSC_LANE];
        u32 i = 1;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (range < (1u << 24)) { range <<= 8; code = (code << 8) + rc_byte(); }
            const u32 sub = i & 7, owner = i >> 3;
            const u32 word = sub < 4 ? (sub < 2 ? r.x : r.y) : (sub < 6 ? r.z : r.w);
            u32 p = (__shfl_sync(0xFFFFFFFFu, word, owner) >> ((sub & 1) * 16)) & 0xFFFFu;
            const u32 bound = (range >> 12) * p;
            u32 bit;
            if (code < bound) { range = bound; p += (0xFFFu - p) >> 5; bit = 1; }
            else { range -= bound; code -= bound; p -= p >> 5; bit = 0; }
            if (CSC_LANE == owner) row[i] = (u16)p;
            i = i + i + bit;
        }
        w_sync();
        return i & 0xFF;
#else
        return tree(row, 0x100) & 0xFF;
#endif
    }

    // copy len bytes from x - dist (overlap allowed, csc_dec.cpp:516-518)
    CSC_HD u32 lz_copy(u32 dist, u32 len) {   // returns the last byte copied (the next literal context)
        u8 *d = S.wnd + x;
        const u8 *s = d - dist;
        const u32 last = s[dist >= len ? len - 1 : (len - 1) % dist];   // on its way while the copy runs
        if (dist >= len) {
            for (u32 k = CSC_LANE; k < len; k += CSC_NLANES) d[k] = s[k];
        } else {
            // overlapping copy = periodic extension: every lane reads the already-final period
            for (u32 k = CSC_LANE; k < len; k += CSC_NLANES) d[k] = s[k % dist];
        }
        w_sync();
        x += len;
        return last;
    }

    // lz_decode (csc_dec.cpp:476-571). Returns decoded size or sets err.
    CSC_HD u32 lz_decode(u32 limit) {
        const u32 wsz = S.wnd_size;
        const bool lin = S.out_cap <= wsz;   // the stream never wraps its window: ring offsets are the linear positions
        u32 i = 0;
        while (i <= limit && !err) {
            if (dec_bit(M.p_state[state * 3 + 0]) == 0) {
                if (x >= S.out_cap) { err = kErrOverflow; break; }
                u32 c = decode_literal_sym(S.p_lit + ctx * 256);
                ctx = c;
                state = (state * 4) & 0x3F;
                S.wnd[x++] = (u8)c;
                i++;
                continue;
            }
            u32 dist, len;
            if (dec_bit(M.p_state[state * 3 + 1]) == 1) {
                decode_match(dist, len);
                if (len == 0 && dist == 64) break;   // end of run
                dist++; len += 2;
                rep[3] = rep[2]; rep[2] = rep[1]; rep[1] = rep[0]; rep[0] = dist;
            } else if (dec_bit(M.p_state[state * 3 + 2]) == 0) {
                // 1-byte match with rep0 (csc_dec.cpp:527-534)
                state = (state * 4 + 2) & 0x3F;
                ctx = 0;
                if (rep[0] == 0 || rep[0] > x || x >= S.out_cap) { err = kErrDecode; break; }
                const u32 b = S.wnd[x - rep[0]];
                S.wnd[x] = (u8)b;
                x++; i++;
                ctx = b;
                continue;
            } else {
                u32 idx = 0;
                u32 t = 1;
                t = t + t + dec_bit(M.p_repdist[state * 3 + t - 1]);
                t = t + t + dec_bit(M.p_repdist[state * 3 + t - 1]);
                idx = t & 3;
                len = decode_matchlen_2() + 2;
                state = (state * 4 + 3) & 0x3F;
                dist = rep[idx];
                for (u32 j = idx; j > 0; j--) rep[j] = rep[j - 1];
                rep[0] = dist;
            }
            if (err) break;
            // bounds checks of the reference ring (csc_dec.cpp:508-510,545-547) on the linear history
            if (dist == 0 || dist > x || dist >= wsz || len + i > limit || x + len > S.out_cap) { err = kErrDecode; break; }
            if (!lin && (x % wsz + len > wsz || (x - dist) % wsz + len > wsz)) { err = kErrDecode; break; }   // (lin: implied by x + len <= out_cap <= wsz)
            ctx = lz_copy(dist, len);
            i += len;
        }
        return i;
    }

    CSC_HD void run() {
        for (u32 i = CSC_LANE; i < 64 * 3; i += CSC_NLANES) { M.p_state[i] = 2048; M.p_repdist[i] = 2048; }
        for (u32 i = CSC_LANE; i < 8 + 16 * 2 + 32 * 4; i += CSC_NLANES) M.p_dist[i] = 2048;
        for (u32 i = CSC_LANE; i < 128; i += CSC_NLANES) M.p_matchlen_extra3[i] = 2048;
        for (u32 i = CSC_LANE; i < 29 * 16; i += CSC_NLANES) M.p_matchdist_extra[i] = 2048;
        for (u32 i = CSC_LANE; i < 8; i += CSC_NLANES) { M.p_matchlen_extra1[i] = 2048; M.p_matchlen_extra2[i] = 2048; }
        if (CSC_LANE == 0) { M.p_matchlen_slot[0] = M.p_matchlen_slot[1] = 2048; M.p_longlen = 2048; M.p_rle_flag = 2048; }
        w_sync();
        state = 0; ctx = 0; err = 0; x = 0;
        rep[0] = rep[1] = rep[2] = rep[3] = 0;
        rc_blk = bc_blk = 0;
        rcp = S.rc; rc_len = S.rc_len; n_rc_blk = S.n_rc_blk;
        rc_cur = rc_nxt = 0;
        u32 nruns = 0;
        open_segment(true);
        const u32 maxb = S.raw_blocksize;
        while (!err) {   // CSCDecoder::Decompress, csc_dec.cpp:586-682
            u32 type = decode_int();
            if (err) break;
            u32 x0 = x, size = 0;
            if (type == SIG_EOF) break;
            if (type == DT_NORMAL || type == DT_EXE) {
                size = lz_decode(maxb);
            } else if (type == DT_ENGTXT) {
                (void)decode_int();
                size = lz_decode(maxb);
            } else if (type == DT_BAD) {
                size = decode_int();
                if (size > maxb || x + size > S.out_cap) { err = -1; break; }
                raw_bytes(S.wnd + x, size);
                x += size;
            } else if (type == DT_ENTROPY) {
                size = decode_int();
                if (size > maxb || x + size > S.out_cap) { err = -1; break; }
                for (u32 k = 0; k < size && !err; k++) {
                    u32 c = decode_literal_sym(S.p_lit + ctx * 256);
                    ctx = c;
                    S.wnd[x + k] = (u8)c;
                }
                x += size;
            } else if (type >= DT_DLT && type < DT_DLT + DLT_CHANNEL_MAX) {
                // decode_rle (csc_dec.cpp:107-153) fused with Inverse_Delta (csc_filters.cpp:371-398):
                // the reconstructed bytes are the running sum of the planar residuals
                u32 chn = dlt_channels(type - DT_DLT);
                size = decode_int();
                if (size > maxb || x + size > S.out_cap) { err = -1; break; }
                const bool filt = size >= 512;
                u32 base = size / chn, rem = size % chn;
                u32 c_ch = 0, c_idx = 0, c_cnt = base + (rem ? 1 : 0);   // planar cursor -> (channel, index)
                u32 sctx = 0, lastres = 0, acc = 0;
                for (u32 k = 0; k < size && !err;) {
                    u32 nrep = 1, res;
                    if (dec_bit(M.p_rle_flag) == 0) {
                        res = decode_literal_sym(S.p_delta + sctx * 256);
                    } else {
                        nrep = decode_matchlen_2() + 11;
                        if (k == 0) { err = -1; break; }
                        res = lastres;
                        if (nrep > size - k) nrep = size - k;
                    }
                    sctx = res; lastres = res;
                    for (u32 r = 0; r < nrep; r++, k++) {
                        if (!filt) { S.wnd[x + k] = (u8)res; continue; }
                        acc = (acc + res) & 0xFF;
                        S.wnd[x + c_ch + c_idx * chn] = (u8)acc;
                        if (++c_idx == c_cnt) { c_ch++; c_idx = 0; c_cnt = base + (c_ch < rem ? 1 : 0); }
                    }
                }
                x += size;
            } else {
                err = kErrDecode;
                break;
            }
            if (err) break;
            if (nruns < S.runs_cap) {
                if (CSC_LANE == 0) { S.runs[nruns].type = type; S.runs[nruns].off = x0; S.runs[nruns].size = size; }
            } else { err = kErrOverflow; break; }
            nruns++;
            if (decode_int() == 1) open_segment(false);
        }
        if (CSC_LANE == 0) { *S.status = err; *S.n_runs = nruns; *S.out_len = x; }
    }
};

}  // namespace cscb
