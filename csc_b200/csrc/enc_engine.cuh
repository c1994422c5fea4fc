// enc_engine.cuh -- the serial half of the libcsc encode path, one stream per warp.
//
// What runs here is everything that SURVEY.md H1 proves serially dependent
// inside one stream: dynamic block typing + run merging
// (csc_encoder_main.cpp:85-146), the state-aware filtering of the parallel
// match-finder precomputation so that it reproduces the reference tables
// exactly (csc_mf.cpp:134-495), the lazy parser (csc_lz.cpp:156-199), the
// price-optimal parser (csc_lz.cpp:207-362), the packet grammar
// (csc_model.cpp) and the range / raw-bit coder (csc_coder.{h,cpp}).
//
// It is NOT a translation of the reference's data structures: the 6-byte hash
// table does not exist here.  The parallel kernels (kernels.cuh: k_mf_*) produce, for
// every position p, the K nearest earlier positions with the same 6-byte hash
// (row P6[p]) and their forward match lengths (L6[p]).  The reference bucket
// at any moment equals "the first W entries along that chain that are still
// alive", where a position is NOT alive if it was skipped by SlidePos' long-
// match rule (csc_mf.cpp:145), sampled out by SlidePosFast (csc_mf.cpp:215),
// or overwritten by the run de-duplication rule (csc_mf.cpp:147-157).  The
// engine keeps one bit per position (D6) for that and walks rows.
// The 2-/3-byte single-slot tables are depth-1 chains (P3/L3, P2/L2) with their own liveness bitmap (D23);
// a stream whose statically excluded blocks get LZ-parsed after all falls back to the real tables (tmode).
#pragma once
#include "csc_defs.h"
#include "warp.h"
#include "plan.h"
#include "tstage.cuh"

// TMA staging of the candidate records through shared memory (tstage.cuh); 0: plain loads + L1 prefetch
#ifndef CSC_TMA_STAGE
#define CSC_TMA_STAGE 0
#endif
// A/B switches of the record path (compile-time, for same-box comparisons: profiles/r02_ab_*.log); the values below are the product build
#ifndef CSC_OPT_TREC
#define CSC_OPT_TREC 1
#endif
#ifndef CSC_OPT_LITRUN
#define CSC_OPT_LITRUN 1
#endif
#ifndef CSC_OPT_PAIR
#define CSC_OPT_PAIR 1
#endif
#ifndef CSC_OPT_PAIR_TAB
#define CSC_OPT_PAIR_TAB 0   // measured slower (text -m2 354 -> 406 ms per GiB): the table path lives on a small hot loop
#endif
#ifndef CSC_OPT_LITW
#define CSC_OPT_LITW 32
#endif
#ifndef CSC_OPT_FUSED
#define CSC_OPT_FUSED 1
#endif
#ifndef CSC_OPT_BESTOF_OUT
#define CSC_OPT_BESTOF_OUT 1
#endif

namespace cscb {

CSC_HD u32 ilog2(u32 v) {
#if defined(__CUDA_ARCH__)
    return 31u - (u32)__clz((int)v);
#else
    return 31u - (u32)__builtin_clz(v);
#endif
}

// distance slot: largest s with dist_table_[s] <= d (csc_model.cpp:45-55,331-342);
// the table is {0,1,2, 2^(s-2)+1 ...} so the search has a closed form.
CSC_HD u32 dist_slot(u32 d) { return d < 3 ? d : 2u + ilog2(d - 1); }

CSC_HD u32 rev4(u32 v) { return ((v & 1) << 3) | ((v & 2) << 1) | ((v & 4) >> 1) | ((v & 8) >> 3); }

// read-mostly context of the match-finder hot loop; a local copy lives in registers inside the parser
// parser control flow: routing these branches through votes (w_uni) was measured slower than leaving them to
// the compiler (r01: -m2 text 705 vs 669 ms), unlike the coder loops where it removes the BMOV traffic
#ifdef CSC_PARSER_UNI
#define p_uni(x) w_uni(x)
#else
#define p_uni(x) (x)
#endif
struct MfCtx {
    const u8 *lz;
    const u32 *P6; const u16 *L6; const u32 *P3; const u16 *L3; const u32 *P2; const u16 *L2;
    const u32 *Z6;
    u32 *D6, *D23, *ht2, *ht3, *cl, *cd;
    u32 *dsum; u32 gs;     // liveness summary (ModelSmall::dsum) and its granule shift
    u32 *dhi;              // ModelSmall::dead_hi
    const TRec *T;         // precomputed table candidates, or null (wrapped stream)
    TStage *ts;            // this warp's staging ring for the records (device)
    const u32 *E6;         // bitmap (with T): the record of the position holds table candidates (or overflowed)
    const u32 *B6;         // bitmap (with T): HASH6 == 0 or equal to the previous position's -- the only positions SlidePos' 6-byte rules act on
    u32 vld, W, K, good;
    u32 wnd;               // window size when the stream wraps, else 0 (no ring arithmetic on the hot path)
    bool tmode;
};

struct Encoder {
    const EncStream &S;
    ModelSmall &M;
    // The range-coder registers (csc_coder.h:44-50) do not live here: the engine hands every binary decision
    // (probability + bit) to the range-coder pass of rc_engine.cuh, which runs after it.  Only the raw-bit
    // coder stays, because its output is a pure function of the packet sequence.
    u32 bc_bits, bc_val;
    u32 dec_n, bc_n, bc_blk, bev_n;
    u32 bq_n, bq_bits;
    // model (csc_model.h)
    u32 state, ctx, lp_rebuild;
    u32 gs;   // granule shift of the liveness summary
    // LZ (csc_lz.h:22)
    u32 rep[4];
    i32 err;
    u32 tr_n;
    u32 vld, W, K, good;
    bool tmode, want_tables;   // 2-/3-byte tables: false = chains + liveness bits, true = real tables (after a flip)
    bool wrap;                 // the stream is longer than its window: ring semantics apply (csc_lz.cpp:91-92)
    bool tracing;              // debug trace requested (tests)
    bool use_rec;              // lazy parser: take the record path for the next sub-block
    bool run_text;             // the run being parsed is typed ENGTXT: no records were built for it (k_mf_tlist), table path only
    TStage *ts;                // staging ring of this warp (device; null on the host)
    u32 seg_dec0, seg_bc0, seg_bev0;

    CSC_HD Encoder(const EncStream &s, ModelSmall &m, TStage *t = nullptr) : S(s), M(m), ts(t) {}
    CSC_HDF MfCtx mfctx() const {
        MfCtx h;
        h.lz = S.lz; h.P6 = S.P6; h.L6 = S.L6; h.P3 = S.P3; h.L3 = S.L3; h.P2 = S.P2; h.L2 = S.L2; h.Z6 = S.Z6;
        h.D6 = S.D6; h.D23 = S.D23; h.ht2 = S.ht2; h.ht3 = S.ht3; h.cl = M.cl; h.cd = M.cd;
        h.dsum = M.dsum; h.gs = gs; h.dhi = &M.dead_hi; h.T = (wrap || !CSC_OPT_TREC) ? nullptr : S.T; h.B6 = S.B6; h.E6 = S.E6; h.ts = ts;
        h.vld = vld; h.W = W; h.K = K; h.good = good; h.tmode = tmode;
        h.wnd = wrap ? S.cfg.wnd_size : 0;
        return h;
    }

    // ------------------------------------------------------------------ output
    CSC_HD void put_bc(u32 b, u32 stamp) {
        if (bc_n < S.bc_cap) S.bc_out[bc_n] = (u8)b; else err = kErrOverflow;
        bc_n++;
        if (++bc_blk == S.cfg.csc_blocksize) {   // BCWCheckBound: a 64 KiB BC block completed
            if (bev_n < S.ev_cap) S.bc_ev[bev_n] = stamp; else err = kErrOverflow;
            bev_n++;
            bc_blk = 0;
        }
    }
    CSC_HD void trace(u32 k, u32 a, u32 b) {
        if (tracing) {
            if (tr_n < S.trace_cap) { S.trace[tr_n * 3] = k; S.trace[tr_n * 3 + 1] = a; S.trace[tr_n * 3 + 2] = b; }
            tr_n++;
        }
    }
    // debug: the candidate list of one find_match call as (4, position, n) + n x (5, len, dist code) records,
    // the same records oracle/ref_mftrace.cpp takes from the reference MatchFinder
    CSC_HDN void trace_cands(u32 x, u32 n) {
        trace(4, x, n);
        for (u32 i = 0; i < n; i++) trace(5, M.cl[i], M.cd[i]);
    }

    // ------------------------------------------------------------- binary decisions
    CSC_HD void coder_reset() {
        bc_bits = 0; bc_val = 0; bc_blk = 0;
        bq_n = 0; bq_bits = 0;
    }
    // EncodeBit (csc_coder.h:67-81) is split in two.  The adaptive-probability half of a decision
    // (load p, adapt, store) does not depend on the coder registers, so a packet first QUEUES its
    // decisions (slot id + bit); flush_bits() then lets lane k fetch/adapt/store probability k (one
    // memory round trip for the whole packet) and appends (p, bit) to the stream's decision list.
    // The range / low recurrence itself is the range-coder pass (rc_engine.cuh).
    // A queue never holds the same slot twice (callers flush in between), and it is flushed before
    // any raw-bit write so that RC/BC block events keep their temporal order (H4).
    CSC_HDF void q_push(u32 slot, u32 bit) {
        M.bq[bq_n] = slot;
        bq_bits |= bit << bq_n;
        bq_n++;
    }
    CSC_HDF void enc_bit(u16 &pr, u32 bit) {   // small-model probability (lives in ModelSmall)
        q_push((u32)(&pr - reinterpret_cast<u16 *>(&M)), bit);
    }
    CSC_HDF void enc_bit_g(u32 gidx, u32 bit) { q_push(kSlotGlobal | gidx, bit); }   // p_lit / p_delta entry
    CSC_HDN void flush_bits() {
        const u32 n = w_uni_u32(bq_n);
        if (!n) return;
        const u32 bits = bq_bits;
        if (dec_n + n > S.dec_cap) { err = kErrOverflow; bq_n = 0; bq_bits = 0; return; }
#if defined(__CUDA_ARCH__)
        if (CSC_LANE < n) {
            const u32 slot = M.bq[CSC_LANE];
            u16 *pp = (slot & kSlotGlobal) ? S.p_lit + (slot & ~kSlotGlobal) : reinterpret_cast<u16 *>(&M) + slot;
            const u32 v = *pp, bit = (bits >> CSC_LANE) & 1;
            *pp = (u16)(bit ? v + ((0xFFFu - v) >> 5) : v - (v >> 5));
            S.dec[dec_n + CSC_LANE] = (DecEntry)(v | (bit << 15));
        }
#else
        for (u32 k = 0; k < n; k++) {
            const u32 slot = M.bq[k];
            u16 *pp = (slot & kSlotGlobal) ? S.p_lit + (slot & ~kSlotGlobal) : reinterpret_cast<u16 *>(&M) + slot;
            const u32 pv = *pp, bit = (bits >> k) & 1;
            *pp = (u16)(bit ? pv + ((0xFFFu - pv) >> 5) : pv - (pv >> 5));
            S.dec[dec_n + k] = (DecEntry)(pv | (bit << 15));
        }
#endif
        dec_n += n;
        w_sync();
        bq_n = 0; bq_bits = 0;
    }
    CSC_HDN void direct16(u32 val, u32 len) {   // csc_coder.cpp:76-87
        flush_bits();   // keep RC decisions ahead of later raw bits (block-event order)
        bc_val = (bc_val << len) | val;
        bc_bits += len;
        while (bc_bits >= 8) {
            put_bc((bc_val >> (bc_bits - 8)) & 0xFF, 2 * (dec_n - seg_dec0));   // after dec_n - seg_dec0 decisions of the segment
            bc_bits -= 8;
        }
    }
    // raw bits whose block completions carry an explicit time stamp (the caller appends its decisions afterwards);
    // the decision queue is known to be empty
    CSC_HDF void raw16(u32 val, u32 len, u32 stamp) {
        bc_val = (bc_val << len) | val;
        bc_bits += len;
        while (bc_bits >= 8) { put_bc((bc_val >> (bc_bits - 8)) & 0xFF, stamp); bc_bits -= 8; }
    }
    CSC_HD void direct(u32 val, u32 len) {     // csc_coder.h:83-88
        if (len <= 16) direct16(val, len);
        else { direct16(val >> 16, len - 16); direct16(val & 0xFFFF, 16); }
    }
    CSC_HDN void flush_segment(u32 seg) {   // csc_coder.cpp:40-74 (the RC half is RcCoder::flush)
        flush_bits();
        SegOut so;
        so.dec_off = seg_dec0; so.dec_len = dec_n - seg_dec0;
        // the two pad bytes of the raw-bit coder follow the five flush bytes of the range coder in time
        put_bc((bc_val << (8 - bc_bits)) & 0xFF, 2 * (so.dec_len + 1));
        put_bc(0, 2 * (so.dec_len + 1));
        so.rc_off = 0; so.rc_len = 0; so.rev_off = 0; so.rev_len = 0;
        so.bc_off = seg_bc0; so.bc_len = bc_n - seg_bc0;
        so.bev_off = seg_bev0; so.bev_len = bev_n - seg_bev0;
        if (CSC_LANE == 0) S.segout[seg] = so;
        dec_n = (dec_n + 7u) & ~7u;   // the next segment's decisions start on a 16-byte boundary
        seg_dec0 = dec_n; seg_bc0 = bc_n; seg_bev0 = bev_n;
        coder_reset();
    }

    // ------------------------------------------------------------------- model
    CSC_HDN void model_reset() {   // csc_model.cpp:88-111
        for (u32 i = CSC_LANE; i < 64 * 3; i += CSC_NLANES) { M.p_state[i] = 2048; M.p_repdist[i] = 2048; }
        for (u32 i = CSC_LANE; i < 8 + 16 * 2 + 32 * 4; i += CSC_NLANES) M.p_dist[i] = 2048;
        for (u32 i = CSC_LANE; i < 128; i += CSC_NLANES) M.p_matchlen_extra3[i] = 2048;
        for (u32 i = CSC_LANE; i < 29 * 16; i += CSC_NLANES) M.p_matchdist_extra[i] = 2048;
        for (u32 i = CSC_LANE; i < 8; i += CSC_NLANES) { M.p_matchlen_extra1[i] = 2048; M.p_matchlen_extra2[i] = 2048; }
        if (CSC_LANE == 0) {
            M.p_matchlen_slot[0] = M.p_matchlen_slot[1] = 2048;
            M.p_longlen = 2048; M.p_rle_flag = 2048;
        }
        for (u32 i = CSC_LANE; i < 32; i += CSC_NLANES) M.len_price[i] = 0;
        w_sync();
        state = 0; ctx = 0; lp_rebuild = 0;
    }

    // 8 literal bits under the order-1 row `row` (index of the row start in p_lit/p_delta, csc_model.cpp:176-182)
    CSC_HDN void lit_bits(u32 row, u32 c) {
        u32 cc = c | 0x100;
#pragma unroll
        for (int k = 0; k < 8; k++) enc_bit_g(row + (cc >> (8 - k)), (cc >> (7 - k)) & 1);
    }
    CSC_HDN void encode_literal(u32 c) {   // csc_model.cpp:169-183
        trace(0, c, 0);
#if defined(__CUDA_ARCH__)
        // device fast path: the 9 decisions (flag + 8 tree nodes) are known up front, lane k owns decision k
        const u32 cc = w_uni_u32(c) | 0x100;   // (the reduction also tells the compiler the warp is converged here)
        const u32 lane = CSC_LANE;
        if (dec_n + 9 > S.dec_cap) { err = kErrOverflow; return; }
        if (lane < 9) {
            u16 *pp = lane == 0 ? &M.p_state[state * 3] : S.p_lit + ctx * 256 + (cc >> (9 - lane));
            const u32 v = *pp;
            const u32 bit = lane == 0 ? 0u : (cc >> (8 - lane)) & 1u;
            *pp = (u16)(bit ? v + ((0xFFFu - v) >> 5) : v - (v >> 5));
            S.dec[dec_n + lane] = (DecEntry)(v | (bit << 15));
        }
        dec_n += 9;
        state = (state * 4) & 0x3F;
        ctx = c;
        __syncwarp();
#else
        enc_bit(M.p_state[state * 3 + 0], 0);
        state = (state * 4) & 0x3F;
        u32 row = ctx * 256;
        ctx = c;
        lit_bits(row, c);
        flush_bits();
#endif
    }
    CSC_HDN void encode_matchlen_1(u32 len) {   // csc_model.cpp:113-145
        len = w_uni_u32(len);
        if (len < 16) {
            u16 *p;
            if (len < 8) {
                enc_bit(M.p_matchlen_slot[0], 0);
                p = M.p_matchlen_extra1;
            } else {
                enc_bit(M.p_matchlen_slot[0], 1);
                enc_bit(M.p_matchlen_slot[1], 0);
                len -= 8;
                p = M.p_matchlen_extra2;
            }
            u32 c = len | 8;
            do { enc_bit(p[c >> 3], (c >> 2) & 1); c <<= 1; } while (w_uni(c < 0x40));
        } else {
            enc_bit(M.p_matchlen_slot[0], 1);
            enc_bit(M.p_matchlen_slot[1], 1);
            u32 c = (len - 16) | 0x80;
            do { enc_bit(M.p_matchlen_extra3[c >> 7], (c >> 6) & 1); c <<= 1; } while (w_uni(c < 0x4000));
        }
    }
    CSC_HDN void encode_matchlen_2(u32 len) {   // csc_model.cpp:147-159
        if (len >= 143) {
            encode_matchlen_1(143);
            flush_bits();
            len -= 143;
            while (len >= 143) { len -= 143; enc_bit(M.p_longlen, 0); flush_bits(); }
            enc_bit(M.p_longlen, 1);
            flush_bits();
        }
        encode_matchlen_1(len);
    }
    // ---- non-literal packets, lane-parallel -------------------------------------------------------------
    // Every binary decision of a packet is a pure function of (state, length code, distance): decision k is
    // worked out by lane k (slot + bit), all probabilities are fetched / adapted / appended in one go, the same
    // shape as the literal fast path.  A packet never uses a slot twice.  Length codes >= 143 (matches longer
    // than 144 bytes: the p_longlen escape chain) take the serial statement below.
    CSC_HDF u32 slot_of(const u16 &pr) const { return (u32)(&pr - reinterpret_cast<const u16 *>(&M)); }
    CSC_HDF static u32 len_ndec(u32 len) { return len < 8 ? 4u : len < 16 ? 5u : 9u; }
    // decision j of encode_matchlen_1(len) (csc_model.cpp:113-145), len < 143
    CSC_HDF u32 len_dec(u32 len, u32 j, u32 &bit) const {
        if (len < 8) {
            if (j == 0) { bit = 0; return slot_of(M.p_matchlen_slot[0]); }
            const u32 c = (len | 8) << (j - 1);
            bit = (c >> 2) & 1;
            return slot_of(M.p_matchlen_extra1[c >> 3]);
        }
        if (j == 0) { bit = 1; return slot_of(M.p_matchlen_slot[0]); }
        if (len < 16) {
            if (j == 1) { bit = 0; return slot_of(M.p_matchlen_slot[1]); }
            const u32 c = ((len - 8) | 8) << (j - 2);
            bit = (c >> 2) & 1;
            return slot_of(M.p_matchlen_extra2[c >> 3]);
        }
        if (j == 1) { bit = 1; return slot_of(M.p_matchlen_slot[1]); }
        const u32 c = ((len - 16) | 0x80) << (j - 2);
        bit = (c >> 6) & 1;
        return slot_of(M.p_matchlen_extra3[c >> 7]);
    }
    // fetch / adapt / append decisions [k0, k1) of the current packet; slot_bit(k, bit) names decision k
    template <class F> CSC_HDF void packet_decisions(u32 k0, u32 k1, F slot_bit) {
        if (dec_n + (k1 - k0) > S.dec_cap) { err = kErrOverflow; return; }
#if defined(__CUDA_ARCH__)
        const u32 k = k0 + CSC_LANE;
        if (k < k1) {
#else
        for (u32 k = k0; k < k1; k++) {
#endif
            u32 bit;
            u16 *pp = reinterpret_cast<u16 *>(&M) + slot_bit(k, bit);
            const u32 v = *pp;
            *pp = (u16)(bit ? v + ((0xFFFu - v) >> 5) : v - (v >> 5));
            S.dec[dec_n + (k - k0)] = (DecEntry)(v | (bit << 15));
        }
        dec_n += k1 - k0;
        w_sync();
    }
    CSC_HDN void encode_rep0len1() {   // csc_model.cpp:198-207
        trace(2, 0, 0);
        const u32 st = state;
        packet_decisions(0, 3, [&](u32 k, u32 &bit) { bit = k == 0 ? 1u : 0u; return slot_of(M.p_state[st * 3 + k]); });
        ctx = 0;
        state = (state * 4 + 2) & 0x3F;
    }
    CSC_HDN void encode_rep(u32 idx, u32 lencode) {   // csc_model.cpp:218-232
        trace(3, lencode, idx);
        idx = w_uni_u32(idx);
        lencode = w_uni_u32(lencode);
        const u32 st = state, b1 = (idx >> 1) & 1, b0 = idx & 1;
        if (lencode < 143) {
            const u32 nl = len_ndec(lencode);
            packet_decisions(0, 5 + nl, [&](u32 k, u32 &bit) {
                if (k < 3) { bit = k == 1 ? 0u : 1u; return slot_of(M.p_state[st * 3 + k]); }
                if (k == 3) { bit = b1; return slot_of(M.p_repdist[st * 3 + 0]); }
                if (k == 4) { bit = b0; return slot_of(M.p_repdist[st * 3 + 1 + b1]); }
                return len_dec(lencode, k - 5, bit);
            });
        } else {
            enc_bit(M.p_state[st * 3 + 0], 1);
            enc_bit(M.p_state[st * 3 + 1], 0);
            enc_bit(M.p_state[st * 3 + 2], 1);
            enc_bit(M.p_repdist[st * 3 + 0], b1);
            enc_bit(M.p_repdist[st * 3 + 1 + b1], b0);
            encode_matchlen_2(lencode);
            flush_bits();
        }
        state = (state * 4 + 3) & 0x3F;
    }
    CSC_HDN void encode_match_general(u32 dist, u32 lencode) {   // csc_model.cpp:301-366, decision by decision
        dist = w_uni_u32(dist);
        lencode = w_uni_u32(lencode);
        const u32 st = state;
        u32 base, sbits;
        if (lencode == 0) { base = 0; sbits = 3; }
        else if (lencode <= 2) { base = 16 * (lencode - 1) + 8; sbits = 4; }
        else if (lencode <= 5) { base = 32 * (lencode - 3) + 8 + 32; sbits = 5; }
        else { base = 32 * 3 + 8 + 32; sbits = 5; }
        const u32 slot = dist_slot(dist);
        const u32 cs = slot | (1u << sbits);
        if (lencode < 143) {
            const u32 nl = len_ndec(lencode);
            packet_decisions(0, 2 + nl + sbits, [&](u32 k, u32 &bit) {
                if (k < 2) { bit = 1; return slot_of(M.p_state[st * 3 + k]); }
                if (k < 2 + nl) return len_dec(lencode, k - 2, bit);
                const u32 c = cs << (k - 2 - nl);   // distance-slot tree
                bit = (c >> (sbits - 1)) & 1;
                return slot_of(M.p_dist[base + (c >> sbits)]);
            });
        } else {
            enc_bit(M.p_state[st * 3 + 0], 1);
            enc_bit(M.p_state[st * 3 + 1], 1);
            encode_matchlen_2(lencode);
            u16 *p = M.p_dist + base;
            u32 c = cs;
            do { enc_bit(p[c >> sbits], (c >> (sbits - 1)) & 1); c <<= 1; } while (w_uni(c < (1u << (sbits * 2))));
            flush_bits();
        }
        if (w_uni(slot > 2)) {
            const u32 ebits = slot - 2;
            const u32 elen = dist - (1u << ebits) - 1;
            if (w_uni(ebits > 4)) direct(elen >> 4, ebits - 4);   // raw bits between the slot and the four modelled low bits
            const u32 ce = rev4(elen & 0x0F) | 0x10;
            packet_decisions(0, 4, [&](u32 k, u32 &bit) {
                const u32 c = ce << k;
                bit = (c >> 3) & 1;
                return slot_of(M.p_matchdist_extra[(ebits - 1) * 16 + (c >> 4)]);
            });
        }
        state = (state * 4 + 1) & 0x3F;
    }
    CSC_HDN void encode_match(u32 dist, u32 lencode) {   // csc_model.cpp:301-366
        trace(1, lencode, dist);
#if defined(__CUDA_ARCH__) && CSC_OPT_FUSED
        dist = w_uni_u32(dist);
        lencode = w_uni_u32(lencode);
        if (lencode < 143) {
            const u32 st = state;
            u32 base, sbits;
            if (lencode == 0) { base = 0; sbits = 3; }
            else if (lencode <= 2) { base = 16 * (lencode - 1) + 8; sbits = 4; }
            else if (lencode <= 5) { base = 32 * (lencode - 3) + 8 + 32; sbits = 5; }
            else { base = 32 * 3 + 8 + 32; sbits = 5; }
            const u32 slot = dist_slot(dist);
            const u32 cs = slot | (1u << sbits);
            // the whole packet in one go: lane k owns decision k of state(2) | length(nl) | distance slot(sbits) | the four
            // modelled low bits of the distance; the raw bits in between are stamped with the decision count in front of them
            const u32 nl = len_ndec(lencode), n1 = 2 + nl + sbits, ne = slot > 2 ? 4u : 0u;
            const u32 ebits = slot > 2 ? slot - 2 : 1u, elen = dist - (1u << ebits) - 1;
            if (dec_n + n1 + ne > S.dec_cap) { err = kErrOverflow; return; }
            const u32 k = CSC_LANE;
            u32 bit = 1, sl = slot_of(M.p_state[st * 3]) + k;
            if (k >= n1) {
                const u32 c = (rev4(elen & 0x0F) | 0x10) << (k - n1);
                bit = (c >> 3) & 1;
                sl = slot_of(M.p_matchdist_extra[(ebits - 1) * 16]) + (c >> 4);
            } else if (k >= 2 + nl) {
                const u32 c = cs << (k - 2 - nl);
                bit = (c >> (sbits - 1)) & 1;
                sl = slot_of(M.p_dist[base]) + (c >> sbits);
            } else if (k >= 2) sl = len_dec(lencode, k - 2, bit);
            if (k < n1 + ne) {
                u16 *pp = reinterpret_cast<u16 *>(&M) + sl;
                const u32 v = *pp;
                *pp = (u16)(bit ? v + ((0xFFFu - v) >> 5) : v - (v >> 5));
                S.dec[dec_n + k] = (DecEntry)(v | (bit << 15));
            }
            if (ebits > 4 && ne) {
                const u32 stamp = 2 * (dec_n + n1 - seg_dec0), rv = elen >> 4, rl = ebits - 4;
                if (rl <= 16) raw16(rv, rl, stamp);
                else { raw16(rv >> 16, rl - 16, stamp); raw16(rv & 0xFFFF, 16, stamp); }
            }
            dec_n += n1 + ne;
            w_sync();
            state = (state * 4 + 1) & 0x3F;
            return;
        }
#endif
        encode_match_general(dist, lencode);
    }
    CSC_HDN void encode_int(u32 num) {   // csc_model.cpp:389-414
        u32 slot = num ? ilog2(num) : 0;
        direct(slot, 5);
        if (slot == 0) direct(num, 1);
        else direct(num - (1u << slot), slot);
    }

    // ------------------------------------------------------------------ prices
    CSC_HD u32 bitprice(u32 p, u32 bit) const { return M.p2bits[(bit ? p : 4096u - p) >> 3]; }
    CSC_HD u32 rep0len1_price(u32 fs) const {   // csc_model.cpp:209-216
        return bitprice(M.p_state[fs * 3], 1) + bitprice(M.p_state[fs * 3 + 1], 0) + bitprice(M.p_state[fs * 3 + 2], 0);
    }
    CSC_HD u32 repdist_price(u32 fs, u32 idx) const {   // csc_model.cpp:273-284
        u32 r = bitprice(M.p_state[fs * 3], 1) + bitprice(M.p_state[fs * 3 + 1], 0) + bitprice(M.p_state[fs * 3 + 2], 1);
        u32 b1 = (idx >> 1) & 1, b0 = idx & 1;
        r += bitprice(M.p_repdist[fs * 3 + 0], b1);
        r += bitprice(M.p_repdist[fs * 3 + 1 + b1], b0);
        return r;
    }
    CSC_HD u32 matchdist_price(u32 fs, u32 dist) const {   // csc_model.cpp:368-387
        u32 r = bitprice(M.p_state[fs * 3], 1) + bitprice(M.p_state[fs * 3 + 1], 1);
        u32 s = dist_slot(dist);
        return r + (s > 2 ? s + 2 : 2) * 128;
    }
    CSC_HD u32 len_price_one(u32 len) const {   // body of len_price_rebuild, csc_model.cpp:234-270
        u32 r = 0;
        if (len < 16) {
            const u16 *p;
            if (len < 8) { r += bitprice(M.p_matchlen_slot[0], 0); p = M.p_matchlen_extra1; }
            else { r += bitprice(M.p_matchlen_slot[0], 1) + bitprice(M.p_matchlen_slot[1], 0); len -= 8; p = M.p_matchlen_extra2; }
            u32 c = len | 8;
            do { r += bitprice(p[c >> 3], (c >> 2) & 1); c <<= 1; } while (c < 0x40);
        } else {
            r += bitprice(M.p_matchlen_slot[0], 1) + bitprice(M.p_matchlen_slot[1], 1);
            u32 c = (len - 16) | 0x80;
            do { r += bitprice(M.p_matchlen_extra3[c >> 7], (c >> 6) & 1); c <<= 1; } while (c < 0x4000);
        }
        return r;
    }

    // ----------------------------------------------------- match finder (state)
    // HASH2 / HASH3 (csc_mf.cpp:23-33) of position x from one unaligned word; window bytes at or
    // beyond the sub-block end have not been written yet and read as 0
    CSC_HDF static u32 word_at(const MfCtx &h, u32 x, u32 sb_end) {
        u32 w = ld32u(h.lz + x);
        u32 avail = sb_end - x;
        if (avail < 4) w &= 0xFFFFFFFFu >> ((4 - avail) * 8);
        return w;
    }
    CSC_HDF static u32 h2_of(u32 w) { return ((w & 0xFFFF) * 65521u) & 0x3FFF; }
    CSC_HDF static u32 h3_of(u32 w) { return ((w & 0xFF) << 8) ^ (((w >> 8) & 0xFF) << 5) ^ ((w >> 16) & 0xFF); }
    // Liveness bitmaps live in global memory (one L2 round trip per look-up, behind the row load it depends
    // on).  Dead positions are rare, so a 2048-bit summary in shared memory answers most look-ups: a clear
    // summary bit means no position of that granule was ever marked in D6 or D23.
    CSC_HDF static bool maybe_dead(const MfCtx &h, u32 e) { const u32 g = e >> h.gs; return (h.dsum[g >> 5] >> (g & 31)) & 1u; }
    CSC_HDF static void sum_mark(const MfCtx &h, u32 e) { const u32 g = e >> h.gs; s_or(&h.dsum[g >> 5], 1u << (g & 31)); s_max(h.dhi, e + 1); }
    CSC_HDF static void sum_mark_range(const MfCtx &h, u32 a, u32 b) {   // [a, b), warp-uniform call
        if (a >= b) return;
        if (CSC_LANE == 0) s_max(h.dhi, b);
        const u32 g0 = a >> h.gs, g1 = (b - 1) >> h.gs;
        for (u32 g = g0 + CSC_LANE; g <= g1; g += CSC_NLANES) s_or(&h.dsum[g >> 5], 1u << (g & 31));
        w_sync();
    }
    CSC_HDF static bool dead6(const MfCtx &h, u32 e) { return maybe_dead(h, e) && bit_get(h.D6, e); }
    CSC_HDF static bool dead23(const MfCtx &h, u32 e) { return maybe_dead(h, e) && bit_get(h.D23, e); }
    CSC_HDF static void kill6(const MfCtx &h, u32 e) { sum_mark(h, e); g_or(&h.D6[e >> 5], 1u << (e & 31)); }

    // newest position < bound on the 6-byte chain of x that is still in the table
    // (= slot 0 of the reference bucket at the time the window ended at `bound`)
    CSC_HDN static u32 newest_alive6(const MfCtx &h, u32 x, u32 bound) {
        u32 src = x, k = 0;
        for (;;) {
            u32 e = h.P6[(size_t)src * h.K + k];
            if (e == kNone) return kNone;
            if (++k == h.K) { src = e; k = 0; }
            if (e >= bound) continue;
            if (!dead6(h, e)) return e;
        }
    }

    // warm the lines the next find_match(y) will need while the coder is busy with a packet
    CSC_HDF static void prefetch_core(const MfCtx &h, u32 y, u32 sb_end) {
        if (y >= sb_end) return;
        if ((CSC_OPT_TREC && h.T) && !h.tmode) {   // the records of y .. y + 31 (literal-run check, paired probe): eight 128-byte lines
#if !CSC_TMA_STAGE
            const u32 q = y + 4 * CSC_LANE;
            if (CSC_LANE < 8 && q < sb_end) pf_l1(h.T + q);
#endif
            return;
        }
        if (h.W) { pf_l1(h.P6 + (size_t)y * h.K); pf_l1(h.L6 + (size_t)y * h.K); }
        if (h.tmode) {
            u32 w = word_at(h, y, sb_end);
            pf_l2(h.ht2 + h2_of(w));
            pf_l2(h.ht3 + h3_of(w));
        } else {
            pf_l1(h.P3 + y); pf_l1(h.P2 + y);
        }
    }
    CSC_HDN void prefetch_fm(u32 y, u32 sb_end) { const MfCtx h = mfctx(); prefetch_core(h, y, sb_end); }

    CSC_HDF static u32 mf_bound(u32 len) { return len < 2 ? 0u : 1u << (4 * len - 2); }   // bound[] csc_mf.cpp:245

    CSC_HDF static void push_cand(const MfCtx &h, u32 &cnt, u32 len, u32 dist) { h.cl[cnt] = len; h.cd[cnt] = dist; cnt++; }

    // a table candidate whose full match length ml is already known: same bookkeeping as probe(h, )
    CSC_HDF static bool take(const MfCtx &h, u32 ml, u32 d, u32 &minlen, u32 &cnt) {
        if (ml <= minlen) return false;
        minlen = ml;
        if (ml <= 6 && d >= mf_bound(ml)) return false;
        push_cand(h, cnt, ml, 4 + d);
        return ml >= h.good;
    }
    // first entry at or behind e on a 2-/3-byte chain that the reference really inserted
    CSC_HDN static u32 alive23(const MfCtx &h, const u32 *chain, u32 e) {
        while (e != kNone && dead23(h, e)) e = chain[e];
        return e;
    }

    // compare limit of a candidate at distance d: the sub-block end on our side, the window end on its side
    CSC_HDF static u32 climit_of(const MfCtx &h, u32 x, u32 d, u32 limit) {
        if (!h.wnd) return limit;
        const u32 c = h.wnd - (x - d) % h.wnd;
        return c < limit ? c : limit;
    }

    // one table / rep probe once the candidate is known to be in range: reference semantics of
    // "pre-check byte at minlen, full compare, update minlen, bound filter, record, good_len stop"
    // returns true when every further probe is disabled (dist poisoned)
    CSC_HDF static bool probe(const MfCtx &h, const u8 *cur, u32 d, u32 limit, u32 pre, u32 c1, u32 code, bool bounded, u32 &minlen, u32 &cnt) {
        const u8 *m = cur - d;
        if (minlen == 1 ? (pre != c1) : (m[minlen] != cur[minlen])) return false;
        u32 ml = lcp_bytes(cur, m, limit);
        if (ml <= minlen) return false;
        minlen = ml;
        if (bounded && ml <= 6 && d >= mf_bound(ml)) return false;
        push_cand(h, cnt, ml, code);
        return ml >= h.good;
    }

    // ---- precomputed table candidates (csc_defs.h: TRec) ------------------------------------------------
    // The table half of find_match (csc_mf.cpp:303-485) at position x with every consulted entry alive, no rep
    // match in front (minlen 1) and chain-mode tables: the statement k_mf_tlist runs for every position.
    CSC_HDF static TRec tlist_of(u32 x, u32 limit, u32 vld, u32 W, u32 K, u32 good, u32 e2, u32 l2, u32 e3, u32 l3, const u32 *row,
                                 const u16 *lrow) {
        TRec t;
        for (int i = 0; i < 4; i++) { t.d[i] = 0; t.l[i] = 0; }
        t.far = 0; t.cnt = 0;
        if (limit <= 1) return t;   // find_match returns nothing here whatever the tables hold
        u32 far = 0, minlen = 1, dist = 0, cnt = 0;
        bool stop = false;
        const u32 pos = vld + x;
        for (int which = 2; which <= 3 && !stop; which++) {
            const u32 e = which == 2 ? e2 : e3, ml = which == 2 ? l2 : l3;
            const u32 d = e == kNone ? pos : x - e;
            if (e != kNone && d > far) far = d;
            if (d <= dist) continue;
            dist = d;
            if (d >= vld || (which == 2 && x == d)) continue;   // (`wpos > dist`, csc_mf.cpp:306)
            if (ml <= minlen) continue;
            minlen = ml;
            if (ml <= 6 && d >= mf_bound(ml)) continue;
            if (cnt < 4) { t.d[cnt] = d; t.l[cnt] = (u16)ml; }
            cnt++;
            stop = ml >= good;
        }
        for (u32 k = 0; k < W && k < K; k++) {   // consulted entries count for `far` even behind a stop
            const u32 e = row[k];
            if (e == kNone) break;
            const u32 d = x - e, ml = lrow[k];
            if (d > far) far = d;
            if (stop || d <= dist) continue;
            dist = d;
            if (d >= vld || ml <= minlen) continue;
            minlen = ml;
            if (ml <= 6 && d >= mf_bound(ml)) continue;
            if (cnt < 4) { t.d[cnt] = d; t.l[cnt] = (u16)ml; }
            cnt++;
            stop = ml >= good;
        }
        t.far = far;
        t.cnt = cnt > 4 ? kTOver : cnt;
        return t;
    }
    // the record of x may stand in for the tables: nothing it consulted has been marked dead
    CSC_HDF static bool trec_valid(const MfCtx &h, u32 x, u32 cnt, u32 far) { return cnt != kTOver && x - far >= *h.dhi; }
    // record of position p; on the device it comes out of the staging ring once stage(h, x) has covered x and x + 1
    CSC_HDF static const TRec *trec_at(const MfCtx &h, u32 p) {
#if defined(__CUDA_ARCH__) && CSC_TMA_STAGE
        return ts_rec(h.ts, p);
#else
        return h.T + p;
#endif
    }
    CSC_HDF static void stage(const MfCtx &h, u32 x) {
#if defined(__CUDA_ARCH__) && CSC_TMA_STAGE
        ts_ensure(h.ts, h.T, x);
#else
        (void)h; (void)x;
#endif
    }

#if defined(__CUDACC__)
    // ---- lane-parallel find_match (device) ------------------------------------------------------
    // Lane c owns candidate c of the reference probe order: 0-3 rep distances, 4 the 2-byte table,
    // 5 the 3-byte table, 8..8+K-1 the 6-byte bucket.  The reference's running state turns into two
    // prefix maxima over the lanes: `dist` (a table candidate is entered iff its distance exceeds all
    // earlier table distances) and `minlen` (a candidate is recorded iff its length exceeds all earlier
    // usable lengths); the good_len cut drops every lane behind the first long candidate.
    static __device__ __forceinline__ u32 lane_lcp(const u8 *a, const u8 *m, u32 lim) {
        u32 l = 0;
        while (l + 4 <= lim) {
            u32 dd = ld32u(a + l) ^ ld32u(m + l);
            if (dd) return l + (((u32)__ffs((int)dd) - 1) >> 3);
            l += 4;
        }
        while (l < lim && a[l] == m[l]) l++;
        return l;
    }
    static __device__ __forceinline__ u32 scan_max_excl(u32 v, u32 lane, bool wide) {   // wide: lanes >= 16 in use
#pragma unroll
        for (int sft = 1; sft < 16; sft <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, v, sft); if (lane >= (u32)sft) v = max(v, t); }
        if (wide) { u32 t = __shfl_up_sync(0xFFFFFFFFu, v, 16); if (lane >= 16u) v = max(v, t); }
        u32 e = __shfl_up_sync(0xFFFFFFFFu, v, 1);
        return lane ? e : 0;
    }
    // find_match from the precomputed records, G positions (x, x + 1, ...) at once, eight lanes each: lanes 0-3 of a
    // group probe the rep distances, lanes 4-7 hold the table candidates; the length filter is a prefix maximum over
    // the group.  Candidate list g goes to cl / cd [16 g ...); returns the counts, 8 bits per group.  (Two positions
    // at once is the lazy parser's u1 / u2 pair: nothing changes between the two calls of the reference.)
    template <int G>
    static __device__ __forceinline__ u32 fm_fast(const MfCtx &h, u32 x, u32 limit, u32 r0, u32 r1, u32 r2, u32 r3, u32 tcnt) {
        const u32 lane = CSC_LANE, FULL = 0xFFFFFFFFu;
        const u32 g = lane >> 3, l8 = lane & 7;
        const bool ing = g < (u32)G;
        const u32 px = x + g, lim = ing && limit > g ? limit - g : 0;   // (lim <= 1: find_match returns nothing)
        const u8 *cur = h.lz + px;
        const bool is_rep = l8 < 4;
        u32 d = 0, L = 0;
        bool usable = false, cmp = false;
        if (lim > 1) {
            if (is_rep) {
                d = l8 == 0 ? r0 : l8 == 1 ? r1 : l8 == 2 ? r2 : r3;
                usable = d < h.vld;
                if (usable) {
                    const u8 *m = cur - d;
                    cmp = m[1] == cur[1];
                    if (cmp) L = lane_lcp(cur, m, lim < 64 ? lim : 64);
                }
            } else {
                const TRec *t = trec_at(h, px);
                const u32 tc = G == 1 ? tcnt : (g == 0 ? tcnt & 0xFF : tcnt >> 8);
                if (l8 - 4 < tc) { d = t->d[l8 - 4]; L = t->l[l8 - 4]; usable = true; }
            }
        }
        u32 more = __ballot_sync(FULL, cmp && L == 64 && lim > 64);
        while (more) {   // long rep matches are finished by the whole warp
            const u32 src = (u32)__ffs((int)more) - 1;
            more &= more - 1;
            const u32 dd = __shfl_sync(FULL, d, src), sp = __shfl_sync(FULL, px, src), sl = __shfl_sync(FULL, lim, src);
            const u32 ext = 64 + lcp_bytes(h.lz + sp + 64, h.lz + sp - dd + 64, sl - 64);
            if (lane == src) L = ext;
        }
        u32 v = usable ? L : 0;   // running maximum of the usable lengths in front of each lane of the group (starts at 1)
#pragma unroll
        for (int sft = 1; sft < 8; sft <<= 1) { const u32 t = __shfl_up_sync(FULL, v, sft); if (l8 >= (u32)sft) v = max(v, t); }
        const u32 prev = __shfl_up_sync(FULL, v, 1);
        const u32 run = max(1u, l8 ? prev : 0u);
        bool acc = usable && L > run;
        const u32 gsh = 8 * g;
        const u32 stopm = (__ballot_sync(FULL, acc && L >= h.good) >> gsh) & 0xFFu;
        if (stopm && l8 > (u32)__ffs((int)stopm) - 1) acc = false;
        const u32 has11 = (__ballot_sync(FULL, l8 == 0 && usable && L >= 2) >> gsh) & 1u;
        const u32 accm = (__ballot_sync(FULL, acc) >> gsh) & 0xFFu;
        if (ing) {
            if (has11 && l8 == 0) { h.cl[16 * g] = 1; h.cd[16 * g] = 1; }
            if (acc) { const u32 idx = 16 * g + has11 + (u32)__popc(accm & ((1u << l8) - 1)); h.cl[idx] = L; h.cd[idx] = is_rep ? 1 + l8 : 4 + d; }
        }
        const u32 cnt = ing ? has11 + (u32)__popc(accm) : 0;
        __syncwarp();
        if (G == 1) return __shfl_sync(FULL, cnt, 0);
        return __shfl_sync(FULL, cnt, 0) | (__shfl_sync(FULL, cnt, 8) << 8);
    }
    // the walk behind a row whose first W entries are not all alive (csc_mf.cpp:453-485 keeps going down the bucket)
    static __device__ __noinline__ u32 fm_walk(const MfCtx &h, u32 x, u32 limit, u32 dm, u32 lu, u32 found, u32 src, u32 cnt) {
        const u32 FULL = 0xFFFFFFFFu;
        const u8 *cur = h.lz + x;
        u32 dist = dm, minlen = lu;
#pragma unroll
        for (int sft = 16; sft; sft >>= 1) { dist = max(dist, __shfl_xor_sync(FULL, dist, sft)); minlen = max(minlen, __shfl_xor_sync(FULL, minlen, sft)); }
        minlen = max(minlen, 1u);
        u32 kk = 0;
        const u32 c1 = cur[1];
        while (found < h.W) {
            u32 e = h.P6[(size_t)src * h.K + kk];
            if (e == kNone) break;
            if (++kk == h.K) { src = e; kk = 0; }
            if (dead6(h, e)) continue;
            found++;
            u32 dd = x - e;
            if (dd <= dist) continue;
            dist = dd;
            if (dist >= h.vld) continue;
            const u32 cl = climit_of(h, x, dd, limit);
            if (minlen >= cl) continue;
            if (probe(h, cur, dd, cl, (cur - dd)[1], c1, 4 + dd, true, minlen, cnt)) break;
        }
        return cnt;
    }
    // find_match on the device.  REC: the caller is on the record path -- the precomputed record answers when it is
    // valid and the general statement is an out-of-line call (cold there); !REC: the general statement inline and no
    // record code at all (the table path of the lazy parser, whose hot loop has to stay small).
    template <bool REC>
    static __device__ __forceinline__ u32 fm_lanes(const MfCtx &h, u32 x, u32 limit, u32 r0, u32 r1, u32 r2, u32 r3) {
        if (REC) {
            if (h.T && !h.tmode) {
                stage(h, x);
                const u32 tcnt = trec_at(h, x)->cnt, far = trec_at(h, x)->far;
                if (trec_valid(h, x, tcnt, far)) return fm_fast<1>(h, x, limit, r0, r1, r2, r3, tcnt);
            }
            return fm_tables_out(h, x, limit, r0, r1, r2, r3);
        }
        if (h.tmode) return fm_tables_out(h, x, limit, r0, r1, r2, r3);   // (after a duplicate-block flip: rare)
        return fm_tables<false>(h, x, limit, r0, r1, r2, r3);
    }
    static __device__ __noinline__ u32 fm_tables_out(const MfCtx &h, u32 x, u32 limit, u32 r0, u32 r1, u32 r2, u32 r3) {
        return fm_tables<true>(h, x, limit, r0, r1, r2, r3);
    }
    // the general statement: candidate rows + liveness.  GEN: with the table-mode branches; !GEN: chain mode only (the
    // instance inlined into the lazy parser's table path, kept as small as it gets: its loop has to stay in the instruction cache)
    template <bool GEN>
    static __device__ __forceinline__ u32 fm_tables(const MfCtx &h, u32 x, u32 limit, u32 r0, u32 r1, u32 r2, u32 r3) {
        const u32 lane = CSC_LANE, FULL = 0xFFFFFFFFu;
        const u8 *cur = h.lz + x;
        const u32 pos = h.vld + x;
        const bool is_rep = lane < 4, is_tab = lane == 4 || lane == 5, is_row = lane >= 8 && lane < 8 + h.K;
        u32 d = 0, L = 0;
        bool cand = false, need_cmp = false;
        if (is_rep) {
            d = lane == 0 ? r0 : lane == 1 ? r1 : lane == 2 ? r2 : r3;
            cand = d < h.vld;
            need_cmp = cand;
        } else if (is_tab) {
            const bool three = lane == 5;
            if (!GEN || !h.tmode) {
                const u32 *P = three ? h.P3 : h.P2;
                u32 e = P[x];
                L = three ? h.L3[x] : h.L2[x];
                if (e != kNone && dead23(h, e)) { e = alive23(h, P, e); need_cmp = true; }
                d = (e == kNone) ? pos : x - e;
            } else {
                const u32 w = word_at(h, x, x + limit);
                u32 *T = three ? h.ht3 : h.ht2;
                const u32 hh = three ? h3_of(w) : h2_of(w);
                d = pos - g_ld(&T[hh]);
                T[hh] = pos;
                need_cmp = true;
            }
            cand = true;
        } else if (is_row) {
            const size_t at = (size_t)x * h.K + (lane - 8);
            const u32 e = h.P6[at];
            L = h.L6[at];
            cand = e != kNone && !dead6(h, e);
            d = e != kNone ? x - e : 0;
        }
        const u32 validrow = __ballot_sync(FULL, is_row && d != 0) >> 8;
        const u32 alivem = __ballot_sync(FULL, is_row && cand);
        if (limit <= 1) return 0;
        const u32 wpos = h.wnd ? x % h.wnd : x;   // window offset of x (csc_mf.cpp: wpos)
        const u32 lt = (1u << lane) - 1;
        const bool cons = is_row && cand && (u32)__popc(alivem & lt) < h.W;
        // distance filter: running maximum over the table candidates that are looked at
        const u32 dm = (is_tab || cons) ? d : 0;
        // rows are ordered nearest first, so among the row lanes the running maximum is always the previous
        // looked-at entry (smaller); only the 2-/3-byte candidates in front of them can be larger
        const u32 d2v = __shfl_sync(FULL, d, 4), d3v = __shfl_sync(FULL, d, 5);
        const bool entered = dm > (lane == 4 ? 0u : lane == 5 ? d2v : max(d2v, d3v));
        bool usable;
        if (is_rep) usable = cand;
        else if (lane == 4) usable = entered && d < h.vld && wpos != d && x >= d;   // `wpos > dist` (csc_mf.cpp:306): offset 0 of the lap is lost
        else if (lane == 5) usable = entered && d < h.vld && x >= d;
        else usable = cons && entered && d < h.vld;
        // a compare never runs past the window end on the candidate's side (climit, csc_mf.cpp:268,307,338,460)
        u32 lim = limit;
        if (h.wnd && usable) { lim = min(limit, h.wnd - (x - d) % h.wnd); L = min(L, lim); }
        // lengths that are not precomputed: rep distances, table-mode / re-walked table candidates
        if (need_cmp && usable) {
            const u8 *m = cur - d;
            L = 0;
            if (m[1] == cur[1]) L = lane_lcp(cur, m, lim < 64 ? lim : 64);
        }
        u32 more = __ballot_sync(FULL, need_cmp && usable && L == 64 && lim > 64);
        while (more) {   // long matches are finished by the whole warp
            const u32 src = (u32)__ffs((int)more) - 1;
            more &= more - 1;
            const u32 dd = __shfl_sync(FULL, d, src), ll = __shfl_sync(FULL, lim, src);
            const u32 ext = 64 + lcp_bytes(cur + 64, cur - dd + 64, ll - 64);
            if (lane == src) L = ext;
        }
        // length filter: running maximum of the usable lengths (starts at 1)
        const u32 lu = usable ? L : 0;
        const u32 run = max(1u, scan_max_excl(lu, lane, h.K > 8));
        bool acc = usable && L > run && !(!is_rep && L <= 6 && d >= mf_bound(L));
        const u32 stopm = __ballot_sync(FULL, acc && L >= h.good);
        if (stopm && lane > (u32)__ffs((int)stopm) - 1) acc = false;
        const u32 has11 = __ballot_sync(FULL, lane == 0 && usable && L >= 2) & 1u;   // rep0-len1 rides on a >= 2 rep0 match
        const u32 accm = __ballot_sync(FULL, acc);
        if (has11 && lane == 0) { h.cl[0] = 1; h.cd[0] = 1; }
        if (acc) { const u32 idx = has11 + (u32)__popc(accm & lt); h.cl[idx] = L; h.cd[idx] = is_rep ? 1 + lane : 4 + d; }
        u32 cnt = has11 + (u32)__popc(accm);
        __syncwarp();
        // a dead entry inside a full row leaves the bucket short: keep walking the chain behind the row (rare)
        const u32 full = (1u << h.K) - 1;
        if (h.W && !stopm && validrow == full && (u32)__popc(alivem) < h.W)
            cnt = fm_walk(h, x, limit, dm, lu, (u32)__popc(alivem), __shfl_sync(FULL, x - d, 8 + h.K - 1), cnt);
        return cnt;
    }
#endif

#if defined(__CUDACC__)
    // The general statement for TWO positions at once, x in lanes 0-15 and x + 1 in lanes 16-31 (chain mode, no window
    // wrap, rows of at most 8 entries): the lazy parser's u1 / u2 pair, which the reference computes back to back with
    // nothing in between that the second call could see (find_match(x) only enters x, which the rows of x + 1 already
    // hold as alive).  List g goes to cl / cd [16 g ...); returns the two counts, or kNone when one of the positions
    // needs the walk behind a short bucket (then the caller probes one position at a time).
    static __device__ __forceinline__ u32 fm_tables2(const MfCtx &h, u32 x, u32 limit, u32 r0, u32 r1, u32 r2, u32 r3) {
        const u32 lane = CSC_LANE, FULL = 0xFFFFFFFFu;
        const u32 g = lane >> 4, l = lane & 15, gb = 16 * g;
        const u32 px = x + g, lim = limit > g ? limit - g : 0;
        const u8 *cur = h.lz + px;
        const u32 pos = h.vld + px;
        const bool is_rep = l < 4, is_tab = l == 4 || l == 5, is_row = l >= 8 && l < 8 + h.K;
        u32 d = 0, L = 0;
        bool cand = false, need_cmp = false;
        if (is_rep) {
            d = l == 0 ? r0 : l == 1 ? r1 : l == 2 ? r2 : r3;
            cand = d < h.vld;
            need_cmp = cand;
        } else if (is_tab) {
            const bool three = l == 5;
            const u32 *P = three ? h.P3 : h.P2;
            u32 e = P[px];
            L = three ? h.L3[px] : h.L2[px];
            if (e != kNone && dead23(h, e)) { e = alive23(h, P, e); need_cmp = true; }
            d = (e == kNone) ? pos : px - e;
            cand = true;
        } else if (is_row) {
            const size_t at = (size_t)px * h.K + (l - 8);
            const u32 e = h.P6[at];
            L = h.L6[at];
            cand = e != kNone && !dead6(h, e);
            d = e != kNone ? px - e : 0;
        }
        const u32 validrow = (__ballot_sync(FULL, is_row && d != 0) >> (gb + 8)) & 0xFFu;
        const u32 alivem = (__ballot_sync(FULL, is_row && cand) >> gb) & 0xFFFFu;
        const bool live = lim > 1;   // (limit <= 1: find_match returns nothing)
        const u32 lt = (1u << l) - 1;
        const bool cons = is_row && cand && (u32)__popc(alivem & lt) < h.W;
        const u32 dm = (is_tab || cons) ? d : 0;
        const u32 d2v = __shfl_sync(FULL, d, gb + 4), d3v = __shfl_sync(FULL, d, gb + 5);
        const bool entered = dm > (l == 4 ? 0u : l == 5 ? d2v : max(d2v, d3v));
        bool usable;
        if (is_rep) usable = cand;
        else if (l == 4) usable = entered && d < h.vld && px != d && px >= d;
        else if (l == 5) usable = entered && d < h.vld && px >= d;
        else usable = cons && entered && d < h.vld;
        usable = usable && live;
        if (need_cmp && usable) {
            const u8 *m = cur - d;
            L = 0;
            if (m[1] == cur[1]) L = lane_lcp(cur, m, lim < 64 ? lim : 64);
        }
        u32 more = __ballot_sync(FULL, need_cmp && usable && L == 64 && lim > 64);
        while (more) {   // long matches are finished by the whole warp
            const u32 src = (u32)__ffs((int)more) - 1;
            more &= more - 1;
            const u32 dd = __shfl_sync(FULL, d, src), sp = __shfl_sync(FULL, px, src), sl = __shfl_sync(FULL, lim, src);
            const u32 ext = 64 + lcp_bytes(h.lz + sp + 64, h.lz + sp - dd + 64, sl - 64);
            if (lane == src) L = ext;
        }
        const u32 lu = usable ? L : 0;
        u32 v = lu;   // running maximum of the usable lengths in front of each lane of its group (starts at 1)
#pragma unroll
        for (int sft = 1; sft < 16; sft <<= 1) { const u32 t = __shfl_up_sync(FULL, v, sft); if (l >= (u32)sft) v = max(v, t); }
        const u32 prev = __shfl_up_sync(FULL, v, 1);
        const u32 run = max(1u, l ? prev : 0u);
        bool acc = usable && L > run && !(!is_rep && L <= 6 && d >= mf_bound(L));
        const u32 stopm = (__ballot_sync(FULL, acc && L >= h.good) >> gb) & 0xFFFFu;
        if (stopm && l > (u32)__ffs((int)stopm) - 1) acc = false;
        const u32 has11 = (__ballot_sync(FULL, l == 0 && usable && L >= 2) >> gb) & 1u;
        const u32 accm = (__ballot_sync(FULL, acc) >> gb) & 0xFFFFu;
        const u32 full = (1u << h.K) - 1;
        const bool walk = live && h.W && !stopm && validrow == full && (u32)__popc(alivem) < h.W;
        if (__any_sync(FULL, walk)) return kNone;
        if (has11 && l == 0) { h.cl[gb] = 1; h.cd[gb] = 1; }
        if (acc) { const u32 idx = gb + has11 + (u32)__popc(accm & lt); h.cl[idx] = L; h.cd[idx] = is_rep ? 1 + l : 4 + d; }
        const u32 cnt = has11 + (u32)__popc(accm);
        __syncwarp();
        return __shfl_sync(FULL, cnt, 0) | (__shfl_sync(FULL, cnt, 16) << 8);
    }
#endif

    // find_match (csc_mf.cpp:243-495): candidate list (M.cl / M.cd) in reference probe order.  On the device
    // this is fm_lanes; the scalar statement below is what the host simulation runs against the oracle.
    template <bool REC = true>
    CSC_HDF static u32 fm_core(const MfCtx &h, u32 x, u32 limit, u32 r0, u32 r1, u32 r2, u32 r3) {
#if defined(__CUDA_ARCH__)
        return fm_lanes<REC>(h, x, limit, r0, r1, r2, r3);
#else
        const u8 *cur = h.lz + x;
        const u32 pos = h.vld + x;
        const u32 sb_end = x + limit;
        const u32 wpos = h.wnd ? x % h.wnd : x;
        u32 cnt = 0, minlen = 1, dist = 0;
        u32 e2 = 0, e3 = 0, l2 = 0, l3 = 0, h2 = 0, h3 = 0;   // table entries as reference pos_ values (0 = empty)
        bool known2 = false, known3 = false;                    // match length already known (chain mode)
        if (!h.tmode) {
            u32 x3 = h.P3[x], x2 = h.P2[x];
            l3 = h.L3[x]; l2 = h.L2[x];
            known3 = known2 = true;
            if (x3 != kNone && dead23(h, x3)) { x3 = alive23(h, h.P3, x3); known3 = false; }
            if (x2 != kNone && dead23(h, x2)) { x2 = alive23(h, h.P2, x2); known2 = false; }
            e3 = (x3 == kNone) ? 0 : h.vld + x3;
            e2 = (x2 == kNone) ? 0 : h.vld + x2;
        } else {
            const u32 w0 = word_at(h, x, sb_end);
            h2 = h2_of(w0); h3 = h3_of(w0);
            e2 = g_ld(&h.ht2[h2]); e3 = g_ld(&h.ht3[h3]);
            h.ht2[h2] = pos;
            h.ht3[h3] = pos;
        }
        const u32 c1 = limit > 1 ? cur[1] : 0;
        bool stop = false;
        // rep-distance probes (:266-299); rep0 also yields the 1-byte match when it matches at all
        const u32 rr[4] = {r0, r1, r2, r3};
        for (u32 i = 0; i < 4 && !stop; i++) {
            if (rr[i] >= h.vld) continue;
            const u32 cl = climit_of(h, x, rr[i], limit);
            if (minlen >= cl || (cur - rr[i])[minlen] != cur[minlen]) continue;
            const u32 ml = lcp_bytes(cur, cur - rr[i], cl);
            if (ml && i == 0) push_cand(h, cnt, 1, 1);
            if (ml > minlen) { minlen = ml; push_cand(h, cnt, ml, 1 + i); stop = ml >= h.good; }
        }
        if (stop) dist = 0xFFFFFFFFu;
        if ((CSC_OPT_TREC && h.T) && !h.tmode && trec_valid(h, x, h.T[x].cnt, h.T[x].far)) {   // the precomputed table half (csc_defs.h: TRec)
            const TRec &t = h.T[x];
            if (limit > 1 && !stop)
                for (u32 i = 0; i < t.cnt; i++)
                    if (t.l[i] > minlen) { minlen = t.l[i]; push_cand(h, cnt, t.l[i], 4 + t.d[i]); }
            return cnt;
        }
        // 2-byte and 3-byte single-slot tables (:303-366).  `wpos > dist` (:306): a 2-byte candidate at window
        // offset 0 of the lap is dropped but still poisons dist.
        const u32 d2 = pos - e2, d3 = pos - e3;
        if (d2 > dist) {
            dist = d2;
            if (d2 < h.vld && wpos != d2 && x >= d2) {
                const u32 cl = climit_of(h, x, d2, limit);
                if (minlen < cl) {
                    if (known2) { if (take(h, l2 < cl ? l2 : cl, d2, minlen, cnt)) dist = 0xFFFFFFFFu; }
                    else if (probe(h, cur, d2, cl, (u32)(cur - d2)[1], c1, 4 + d2, true, minlen, cnt)) dist = 0xFFFFFFFFu;
                }
            }
        }
        if (d3 > dist) {
            dist = d3;
            if (d3 < h.vld && x >= d3) {
                const u32 cl = climit_of(h, x, d3, limit);
                if (minlen < cl) {
                    if (known3) { if (take(h, l3 < cl ? l3 : cl, d3, minlen, cnt)) dist = 0xFFFFFFFFu; }
                    else if (probe(h, cur, d3, cl, (u32)(cur - d3)[1], c1, 4 + d3, true, minlen, cnt)) dist = 0xFFFFFFFFu;
                }
            }
        }
        // 6-byte table of width h.W (:453-485): the first h.W live entries of the chain
        if (h.W && dist != 0xFFFFFFFFu) {
            u32 found = 0, src = x, k = 0;
            while (found < h.W) {
                size_t at = (size_t)src * h.K + k;
                u32 e = h.P6[at];
                if (e == kNone) break;
                bool inrow = (src == x);
                if (++k == h.K) { src = e; k = 0; }
                if (dead6(h, e)) continue;
                found++;
                u32 d = x - e;
                if (d <= dist) continue;
                dist = d;
                if (dist >= h.vld) continue;
                const u32 cl = climit_of(h, x, d, limit);
                if (minlen >= cl) continue;
                u32 ml;
                if (inrow) {
                    ml = h.L6[at];
                    if (ml > cl) ml = cl;
                } else {
                    const u8 *m = cur - dist;
                    if (m[minlen] != cur[minlen]) continue;
                    ml = lcp_bytes(cur, m, cl);
                }
                if (ml > minlen) {
                    minlen = ml;
                    if (ml <= 6 && dist >= mf_bound(ml)) continue;
                    push_cand(h, cnt, ml, 4 + dist);
                    if (ml >= h.good) break;
                }
            }
        }
        // x itself is now the newest entry everywhere (alive by default)
        return cnt;
#endif
    }

    CSC_HDF static bool second_better(Cand u1, Cand u2) {   // csc_mf.cpp:570-582 (== fold of :497-524)
        if (u2.len <= 1) return false;
        if (u2.len > u1.len + 3) return true;
        if (u2.len > u1.len && u2.dist <= 4) return true;
        if (u2.len + 2 > u1.len && u2.dist <= 4 && u1.dist > 4) return true;
        if (u2.len >= u1.len && (u2.dist >> (4 * (u2.len - u1.len))) <= u1.dist) return true;
        if (u2.len < u1.len && u2.len + 2 >= u1.len && u1.dist > 4 && (u1.dist >> (4 * (u1.len - u2.len))) > u2.dist) return true;
        return false;
    }
#if CSC_OPT_BESTOF_OUT
    CSC_HDN
#else
    CSC_HDF
#endif
    static Cand best_of(const MfCtx &h, u32 n, u32 at = 0) {   // fold of FindMatch, csc_mf.cpp:497-524
        Cand best; best.len = 1; best.dist = 0;
        for (u32 i = 0; i < n; i++) {
            Cand c; c.len = h.cl[at + i]; c.dist = h.cd[at + i];
            if (i == 0 || second_better(best, c)) best = c;
        }
        return best;
    }
    // out-of-line entry points for the cold callers (optimal parser, typing)
    CSC_HDN u32 find_match(u32 x, u32 limit, u32 r0, u32 r1, u32 r2, u32 r3) { const MfCtx h = mfctx(); return fm_core(h, x, limit, r0, r1, r2, r3); }
    CSC_HDN void slide(u32 s, u32 len, u32 sb_end) { const MfCtx h = mfctx(); slide_core(h, s, len, sb_end); }
    CSC_HDN u32 newest_alive6(u32 x, u32 bound) const { const MfCtx h = mfctx(); return newest_alive6(h, x, bound); }

    // SlidePos (csc_mf.cpp:134-206): positions s+1 .. s+len-1 are passed over by a match.
    CSC_HDN static void slide_general(const MfCtx &h, u32 s, u32 len, u32 sb_end) {
        if (p_uni(len <= 1)) return;
        u32 i0 = 1;
        while (p_uni(i0 + 128 < len)) i0 += 4;
        if (p_uni(h.tmode)) {
#pragma unroll 1
            for (u32 i = 1 + CSC_LANE; i < len; i += CSC_NLANES) {
                if (i < i0 && ((i - 1) & 3)) continue;   // inside the skip zone only every 4th position (:140-145)
                u32 q = s + i;
                u32 w = word_at(h, q, sb_end);
                g_max(&h.ht2[h2_of(w)], h.vld + q);
                g_max(&h.ht3[h3_of(w)], h.vld + q);
            }
        } else if (p_uni(i0 > 1)) {
            // chain mode: everything is "inserted" by default; only the 3 of 4 skipped zone positions are not
            sum_mark_range(h, s + 1, s + i0);
#pragma unroll 1
            for (u32 i = 1 + CSC_LANE; i < i0; i += CSC_NLANES)
                if ((i - 1) & 3) g_or(&h.D23[(s + i) >> 5], 1u << ((s + i) & 31));
        }
        w_sync();
        if (p_uni(h.W != 0)) {
            if (p_uni(i0 > 1)) { sum_mark_range(h, s + 1, s + i0); bits_set_range(h.D6, s + 1, s + i0); }   // never inserted (:145)
            u32 first = s + i0;
            if (p_uni(i0 < len && bit_get_ro(h.Z6, first))) {    // h6 == lasth6 (== 0): slot 0 overwritten (:147-156)
                u32 e = newest_alive6(h, first, first);
                if (e != kNone) kill6(h, e);
            }
#pragma unroll 1
            for (u32 i = i0 + 1 + CSC_LANE; i < len; i += CSC_NLANES) {
                u32 q = s + i;
                if (h.P6[(size_t)q * h.K] == q - 1) kill6(h, q - 1);   // same hash as the previous slid position
            }
            w_sync();
        }
    }

    // the common case on the device (chain mode, no skip zone: matches of at most 129 bytes) only has the two 6-byte rules
    // to look at; everything else is the out-of-line statement above
    template <bool REC = true>
    CSC_HDF static void slide_core(const MfCtx &h, u32 s, u32 len, u32 sb_end) {
        if (p_uni(len <= 1)) return;
#if defined(__CUDA_ARCH__)
        if (h.tmode || len > 129 || !h.W) { slide_general(h, s, len, sb_end); return; }
        const u32 first = s + 1;
        if (REC && (CSC_OPT_TREC && h.T)) {
            // the rules only act on positions whose HASH6 is 0 or repeats the previous position's: one look at the
            // precomputed bitmap of those (k_mf_tlist) settles the common case
            bool any = false;
            const u32 w0 = first >> 5, w1 = (s + len - 1) >> 5;
            for (u32 w = w0 + CSC_LANE; w <= w1; w += CSC_NLANES) {
                u32 m = h.B6[w];
                if (w == w0) m &= 0xFFFFFFFFu << (first & 31);
                if (w == w1) m &= 0xFFFFFFFFu >> (31 - ((s + len - 1) & 31));
                any = any || m != 0;
            }
            if (!w_any(any)) return;
        }
        if (p_uni(bit_get_ro(h.Z6, first))) {    // h6 == lasth6 (== 0): slot 0 overwritten (:147-156)
            u32 e = newest_alive6(h, first, first);
            if (e != kNone) kill6(h, e);
        }
#pragma unroll 1
        for (u32 i = 2 + CSC_LANE; i < len; i += CSC_NLANES) {
            u32 q = s + i;
            if (h.P6[(size_t)q * h.K] == q - 1) kill6(h, q - 1);   // same hash as the previous slid position
        }
        w_sync();
#else
        slide_general(h, s, len, sb_end);
#endif
    }

    // ---------------------------------------------------- packets
    // The parsers hand packets to the model through emit(); a packet is executed on the spot.
    enum : u32 { OP_LIT = 1, OP_MATCH, OP_REP, OP_REP01, OP_INT, OP_BAD, OP_ENTROPY, OP_DELTA, OP_FLUSH, OP_END };
    CSC_HDF void emit(u32 op, u32 a, u32 b, u32 c) {   // `op` is a constant at every call site: this folds to the one call
        if (err) return;
        switch (op) {
            case OP_LIT: encode_literal(a); break;
            case OP_MATCH: encode_match(a, b); if (c < 0x100) ctx = c; break;
            case OP_REP: encode_rep(a, b); ctx = c; break;
            case OP_REP01: encode_rep0len1(); ctx = c; break;
            case OP_INT: encode_int(a); break;
            case OP_BAD: bad_run(a, b); break;
            case OP_ENTROPY: entropy_run(a, b); break;
            case OP_DELTA: delta_run(a, b, c); break;
            case OP_FLUSH:
                flush_segment(a);
                break;
            default: break;
        }
    }

#if defined(__CUDA_ARCH__)
    // ---- literal runs (device) ---------------------------------------------------------------------------
    // Number of leading positions of [x, end), at most 32, where find_match is certain to come back empty: no
    // table candidate in the (valid) precomputed record and no rep distance that matches two bytes.  Lane j
    // looks at position x + j; literals change neither the rep distances nor the liveness of anything, so the
    // answers are independent.  (The 32 records it touches are the ones the next find_match calls read.)
    static __device__ __forceinline__ u32 literal_run(const MfCtx &h, u32 x, u32 end, u32 r0, u32 r1, u32 r2, u32 r3) {
        const u32 p = x + CSC_LANE;
        bool hit = true;
        if (*h.dhi == 0) {
            // nothing of this stream has been marked dead so far, so every record is valid and "has table candidates" is
            // one bit per position: two words of the E6 bitmap instead of 32 records
            const u32 w0 = h.E6[x >> 5], w1 = h.E6[(x >> 5) + 1];
            const u32 tab = __funnelshift_r(w0, w1, x & 31);
            if (p < end) hit = end - p > 1 && ((tab >> CSC_LANE) & 1u);
            if (p < end && !hit && end - p > 1) {
                const u32 w = (u32)h.lz[p] | ((u32)h.lz[p + 1] << 8);
                if (r0 < h.vld) hit |= ((u32)h.lz[p - r0] | ((u32)h.lz[p + 1 - r0] << 8)) == w;
                if (r1 < h.vld) hit |= ((u32)h.lz[p - r1] | ((u32)h.lz[p + 1 - r1] << 8)) == w;
                if (r2 < h.vld) hit |= ((u32)h.lz[p - r2] | ((u32)h.lz[p + 1 - r2] << 8)) == w;
                if (r3 < h.vld) hit |= ((u32)h.lz[p - r3] | ((u32)h.lz[p + 1 - r3] << 8)) == w;
            }
        } else if (p < end && CSC_LANE < CSC_OPT_LITW) {
            if (end - p <= 1) hit = false;
            else {
                const TRec *t = h.T + p;   // (global: the staging ring only covers the probe positions)
                const u32 tc = t->cnt, far = t->far;
                hit = tc != 0 || p - far < *h.dhi;
                if (!hit) {
                    const u32 w = (u32)h.lz[p] | ((u32)h.lz[p + 1] << 8);
                    if (r0 < h.vld) hit |= ((u32)h.lz[p - r0] | ((u32)h.lz[p + 1 - r0] << 8)) == w;
                    if (r1 < h.vld) hit |= ((u32)h.lz[p - r1] | ((u32)h.lz[p + 1 - r1] << 8)) == w;
                    if (r2 < h.vld) hit |= ((u32)h.lz[p - r2] | ((u32)h.lz[p + 1 - r2] << 8)) == w;
                    if (r3 < h.vld) hit |= ((u32)h.lz[p - r3] | ((u32)h.lz[p + 1 - r3] << 8)) == w;
                }
            }
        }
        const u32 m = __ballot_sync(0xFFFFFFFFu, hit);
        return m ? (u32)__ffs((int)m) - 1 : 32u;
    }
    // n literals starting at x, three per step: lane 9 j codes the packet flag of literal j (its state is the
    // literal state chain of the ones in front), lanes 9 j + 1 .. 9 j + 8 its tree levels (csc_model.cpp:169-183)
    __device__ __noinline__ void emit_literals(u32 x, u32 n) {
        const u32 lane = CSC_LANE, j = lane / 9, kk = lane - 9 * j;
        if (dec_n + 9 * n > S.dec_cap) { err = kErrOverflow; return; }
#pragma unroll 1
        for (u32 i = 0; i < n; i += 3) {
            const u32 nb = n - i < 3 ? n - i : 3;
            const bool on = j < nb;
            u32 c = 0, pc = 0;
            if (on) { c = S.lz[x + i + j]; pc = j ? S.lz[x + i + j - 1] : ctx; }
            const u32 stj = j == 0 ? state : j == 1 ? (state * 4) & 0x3F : (state * 16) & 0x3F;
            const u32 cc = c | 0x100;
            u16 *pp = kk == 0 ? &M.p_state[stj * 3] : S.p_lit + pc * 256 + (cc >> (9 - kk));
            batch_decisions(on, pp, kk == 0 ? 0u : (cc >> (8 - kk)) & 1u, 9 * nb);
            state = nb == 1 ? (state * 4) & 0x3F : nb == 2 ? (state * 16) & 0x3F : 0u;
            ctx = S.lz[x + i + nb - 1];
        }
    }
#endif

    // ------------------------------------------------------------- lazy parser
    CSC_HDN void encode_nonlit(Cand u, u32 newctx) {   // csc_lz.cpp:127-154 (+ SetLiteralCtx of the caller)
        if (u.dist <= 4) {
            if (u.len == 1 && u.dist == 1) { emit(OP_REP01, 0, 0, newctx); return; }
            emit(OP_REP, u.dist - 1, u.len - 2, newctx);
            u32 d = rep[u.dist - 1];
            for (u32 j = u.dist - 1; j > 0; j--) rep[j] = rep[j - 1];
            rep[0] = d;
        } else {
            emit(OP_MATCH, u.dist - 5, u.len - 2, newctx);
            rep[3] = rep[2]; rep[2] = rep[1]; rep[1] = rep[0];
            rep[0] = u.dist - 4;
        }
    }
    // compress_normal(size, true), csc_lz.cpp:156-199.  Two instances on the device, same results: REC reads the precomputed
    // records (literal runs in one go, u1 / u2 as a paired probe) and wins where literals dominate; !REC walks the candidate
    // rows, with no record code in its loop, and wins on match-dense data (text).  lz_run picks per sub-block from the
    // literal share of the sub-block before.  Returns the number of literals coded.
    template <bool REC>
    CSC_HDN u32 parse_lazy_t(u32 x0, u32 size) {
        const MfCtx h = mfctx();   // pointers and constants of the hot loop stay in registers
        const u32 end = x0 + size;
        u32 x = x0;
        Cand u1, u2;
        u1.len = 1; u1.dist = 0;
        u2 = u1;
        u32 nlit = 0;
        bool have = false, have2 = false;   // have2: u2 already holds the look-ahead of x + 1 (device: paired probe)
        while (p_uni(x < end)) {
#if defined(__CUDA_ARCH__)
            if (REC && CSC_OPT_LITRUN && !have && (CSC_OPT_TREC && h.T) && !h.tmode && !tracing) {   // a run of literals is found and coded in one go
                stage(h, x);
                const u32 nl = literal_run(h, x, end, rep[0], rep[1], rep[2], rep[3]);
                if (nl) { emit_literals(x, nl); x += nl; nlit += nl; continue; }
            }
#endif
#if defined(__CUDA_ARCH__)
            if (REC && CSC_OPT_PAIR && !have && !tracing && (CSC_OPT_TREC && h.T) && !h.tmode && x + 1 < end) {
                // u1 at x and the lazy look-ahead u2 at x + 1 in one go, when both records stand in for the tables
                const u32 c0 = trec_at(h, x)->cnt, f0 = trec_at(h, x)->far, c1 = trec_at(h, x + 1)->cnt, f1 = trec_at(h, x + 1)->far;
                if (trec_valid(h, x, c0, f0) && trec_valid(h, x + 1, c1, f1)) {
                    const u32 nn = fm_fast<2>(h, x, end - x, rep[0], rep[1], rep[2], rep[3], c0 | (c1 << 8));
                    u1 = best_of(h, nn & 0xFF, 0);
                    u2 = best_of(h, nn >> 8, 16);
                    have = true; have2 = true;
                }
            }
#endif
#if defined(__CUDA_ARCH__)
            if (!REC && CSC_OPT_PAIR_TAB && !have && !tracing && !h.tmode && !h.wnd && h.K <= 8 && x + 1 < end) {
                // table path: the rows of x and x + 1 through the general statement side by side
                const u32 nn = fm_tables2(h, x, end - x, rep[0], rep[1], rep[2], rep[3]);
                if (nn != kNone) {
                    u1 = best_of(h, nn & 0xFF, 0);
                    u2 = best_of(h, nn >> 8, 16);
                    have = true; have2 = true;
                }
            }
#endif
            bool lazy = false;   // u1 was kept after a look-ahead probe at x + 1 (which entered x + 1 into the tables)
            if (p_uni(!have || !(u1.len == 1 || u1.len >= h.good))) {
                // the one probe site: u1 at x, or the lazy look-ahead u2 at x + 1
                Cand c;
                if (have && have2) { c = u2; have2 = false; }
                else {
                    const u32 px = have ? x + 1 : x;
                    const u32 nc = fm_core<REC>(h, px, end - px, rep[0], rep[1], rep[2], rep[3]);
                    if (tracing) trace_cands(px, nc);
                    c = best_of(h, nc);
                    if (p_uni(!have)) { u1 = c; have = true; continue; }
                }
                if (p_uni(second_better(u1, c))) {
                    emit(OP_LIT, h.lz[x], 0, 0);
                    nlit++;
                    x++;
                    u1 = c;
                    continue;
                }
                lazy = true;
            }
            have2 = false;
            if (p_uni(u1.dist == 0)) { emit(OP_LIT, h.lz[x], 0, 0); nlit++; }
            else { prefetch_core(h, x + u1.len, end); encode_nonlit(u1, h.lz[x + u1.len - 1]); }
            slide_core<REC>(h, lazy ? x + 1 : x, lazy ? u1.len - 1 : u1.len, end);
            x += u1.len;
            have = false;
        }
        return nlit;
    }

    // ---------------------------------------------------------- optimal parser
    // Written over 32 virtual lanes (warp.h: Lv).  Lane L of the price table holds the entry for match
    // length L; lane j of the look-ahead window holds the best known (price, dist, back_pos) of position
    // apcur + j.  good_len <= 24 in every optimal-parser level (csc_enc.cpp:70-82), so lengths and the
    // window both fit 32 lanes; the window moves down one lane per position.
    //
    // FindMatchWithPrice (csc_mf.cpp:584-625) after find_match: t0 = longest candidate, t1 = rep0len1,
    // td/tp[L] = candidate and price for length L (the first candidate that reaches L; lengths in the
    // candidate list strictly increase).  GetMatchLenPrice rebuilds its table every 4096 calls
    // (csc_model.cpp:286-299): the calls of one invocation are the priced lengths in ascending order, so
    // the entry of rank r sees the old table iff r < lp_rebuild.
    CSC_HDF void price_table(u32 n, u32 fstate, Cand &t0, Lv &td, Lv &tp, u32 &t1d, u32 &t1p) {
        if (n == 0) { t0.len = 1; t0.dist = 0; }
        else { t0.len = M.cl[n - 1]; t0.dist = M.cd[n - 1]; }
        if (t0.len >= good) return;
        t1d = 0;
        u32 i0 = 0;
        if (n && M.cl[0] == 1 && M.cd[0] == 1) { t1d = 1; t1p = rep0len1_price(fstate); i0 = 1; }
        Lv need;
        LV_FOR(L) {
            u32 dist = 0, price = 0, nd = 0;
            if (L >= 2 && L <= t0.len) {
                u32 c = i0;
                while (M.cl[c] < L) c++;
                const u32 cdist = M.cd[c];
                const u32 rdist = cdist <= 4 ? 0 : cdist - 4;
                if (!(L <= 6 && rdist >= mf_bound(L))) {
                    dist = cdist;
                    nd = 1;
                    price = cdist <= 4 ? repdist_price(fstate, cdist - 1) : matchdist_price(fstate, cdist - 5);
                }
            }
            td.v[LV_AT(L)] = dist; tp.v[LV_AT(L)] = price; need.v[LV_AT(L)] = nd;
        }
        const u32 needm = lv_ballot(need);
        const u32 calls = popc32(needm), k = lp_rebuild;
        if (calls <= k) {
            lp_rebuild = k - calls;
            LV_FOR(L) if (need.v[LV_AT(L)]) tp.v[LV_AT(L)] += M.len_price[L - 2];
        } else {
            LV_FOR(L) if (need.v[LV_AT(L)] && popc32(needm & ((1u << L) - 1)) < k) tp.v[LV_AT(L)] += M.len_price[L - 2];
            w_sync();
            for (u32 i = CSC_LANE; i < 32; i += CSC_NLANES) M.len_price[i] = len_price_one(i);
            w_sync();
            LV_FOR(L) if (need.v[LV_AT(L)] && popc32(needm & ((1u << L) - 1)) >= k) tp.v[LV_AT(L)] += M.len_price[L - 2];
            lp_rebuild = 4096 - (calls - k - 1);
        }
    }
    CSC_HDF u32 literal_price_w(u32 fstate, u32 fctx, u32 c) const {   // csc_model.cpp:185-196, one bit per lane
        const u16 *row = S.p_lit + fctx * 256;
        const u32 cc = c | 0x100;
        Lv a;
        LV_FOR(kk) {
            u32 r = 0;
            if (kk < 8) r = bitprice(row[cc >> (8 - kk)], (cc >> (7 - kk)) & 1);
            else if (kk == 8) r = bitprice(M.p_state[fstate * 3], 0);
            a.v[LV_AT(kk)] = r;
        }
        return lv_sum(a);
    }

    CSC_HDN void ap_backward(u32 endi) {   // csc_lz.cpp:335-362
        ApUnit *ap = S.ap;
        for (u32 i = endi; i;) {
            u32 b = ap[i].back_pos;
            ap[b].next_pos = (u16)i;
            i = b;
        }
        for (u32 i = 0; i != endi;) {
            u32 nx = ap[i].next_pos;
            u32 d = ap[nx].dist;
            if (d == 0) {
                emit(OP_LIT, ap[i].lit, 0, 0);
            } else if (d <= 4) {
                if (nx - i == 1 && d == 1) emit(OP_REP01, 0, 0, ap[nx - 1].lit);
                else emit(OP_REP, d - 1, nx - i - 2, ap[nx - 1].lit);
            } else {
                emit(OP_MATCH, d - 5, nx - i - 2, ap[nx - 1].lit);
            }
            i = nx;
        }
        for (int j = 0; j < 4; j++) rep[j] = ap[endi].rep[j];
    }

    // compress_advanced (csc_lz.cpp:207-333) as one flat loop with a single match-finder site: outside a
    // round a position is either a literal or starts a round; inside, position apcur is finalised (state
    // and rep distances from its best predecessor, :236-262), searched, and relaxes the window (:296-329).
    CSC_HDN void parse_optimal(u32 x0, u32 size) {
        const MfCtx h = mfctx();
        const u32 end = x0 + size;
        ApUnit *const ap = S.ap;
        u32 x = x0;          // wnd_curpos_
        u32 i = 0;           // bytes of the sub-block already emitted
        bool in_round = false;
        u32 apcur = 0, apend = 1, aplimit = 0;
        Lv wp, wd, wb, td, tp;
        u32 t1d = 0, t1p = 0;
        Cand t0;
        t0.len = 1; t0.dist = 0;
        LV_FOR(j) { wp.v[LV_AT(j)] = 0; wd.v[LV_AT(j)] = 0; wb.v[LV_AT(j)] = 0; td.v[LV_AT(j)] = 0; tp.v[LV_AT(j)] = 0; }
        while (p_uni(in_round || i < size)) {
#if defined(__CUDA_ARCH__)
            if (CSC_OPT_LITRUN && !in_round && (CSC_OPT_TREC && h.T) && !h.tmode && !tracing) {   // outside a round a position without candidates is a literal
                stage(h, x);
                const u32 nl = literal_run(h, x, end, rep[0], rep[1], rep[2], rep[3]);
                if (nl) { emit_literals(x, nl); i += nl; x += nl; continue; }
            }
#endif
            const u32 lit = h.lz[x];
            u32 ns = state, r0 = rep[0], r1 = rep[1], r2 = rep[2], r3 = rep[3];
            if (p_uni(in_round)) {   // apcur >= 1
                const u32 d = lv_get(wd, 0), l = lv_get(wb, 0);
                const ApUnit pl = ap[l];
                r0 = pl.rep[0]; r1 = pl.rep[1]; r2 = pl.rep[2]; r3 = pl.rep[3];
                const u32 ls = pl.state;
                if (d == 0) {
                    ns = (ls * 4) & 0x3F;
                } else if (d <= 4) {
                    if (apcur - l == 1 && d == 1) {
                        ns = (ls * 4 + 2) & 0x3F;
                    } else {
                        ns = (ls * 4 + 3) & 0x3F;
                        // (csc_lz.cpp:246-256: the cases fall through, idx 0 leaves the queue untouched)
                        if (d == 4) { const u32 t = r3; r3 = r2; r2 = r1; r1 = r0; r0 = t; }
                        else if (d == 3) { const u32 t = r2; r2 = r1; r1 = r0; r0 = t; }
                        else if (d == 2) { const u32 t = r1; r1 = r0; r0 = t; }
                    }
                } else {
                    ns = (ls * 4 + 1) & 0x3F;
                    r3 = r2; r2 = r1; r1 = r0; r0 = d - 4;
                }
                ApUnit u;
                u.rep[0] = r0; u.rep[1] = r1; u.rep[2] = r2; u.rep[3] = r3;
                u.dist = d; u.back_pos = (u16)l; u.next_pos = 0; u.state = (u8)ns; u.lit = (u8)lit; u.pad = 0; u.pad2 = 0;
                ap[apcur] = u;
                w_sync();
            }
            if (p_uni(!in_round || apcur < aplimit)) {
                if ((CSC_OPT_TREC && h.T) && !h.tmode) {
#if !CSC_TMA_STAGE
                    if ((x & 3) == 0 && x + 32 < end) pf_l1(h.T + x + 32);   // every position is visited: one line ahead
#endif
                }
                else prefetch_core(h, x + 1, end);
                const u32 n = fm_core(h, x, end - x, r0, r1, r2, r3);
                if (tracing) trace_cands(x, n);
                price_table(n, ns, t0, td, tp, t1d, t1p);
            }
            if (p_uni(!in_round)) {
                if (p_uni(t0.dist == 0)) {
                    emit(OP_LIT, lit, 0, 0);
                    i++; x++;
                    continue;
                }
                in_round = true;
                apcur = 0; apend = 1;
                aplimit = (size - i) < kApLimit ? (size - i) : kApLimit;
                LV_FOR(j) wp.v[LV_AT(j)] = j ? 0xFFFFFFFFu : 0u;
                ApUnit u;
                u.rep[0] = r0; u.rep[1] = r1; u.rep[2] = r2; u.rep[3] = r3;
                u.dist = 0; u.back_pos = 0; u.next_pos = 0; u.state = (u8)ns; u.lit = (u8)lit; u.pad = 0; u.pad2 = 0;
                ap[0] = u;
                w_sync();
            } else if (p_uni(apcur == aplimit)) {
                ap_backward(apcur);
                i += apcur;
                in_round = false;
                continue;
            }
            if (p_uni(t0.len == 1 && apcur + 1 == apend)) {
                ap_backward(apcur);
                emit(OP_LIT, lit, 0, 0);
                i += apcur + 1;
                x++;
                in_round = false;
                continue;
            }
            if (apcur + 1 >= apend) apend++;   // (the window already holds 0xFFFFFFFF there)
            if (p_uni(t0.len >= good || (t0.len > 1 && t0.len + apcur >= aplimit))) {
                ap_backward(apcur);
                encode_nonlit(t0, h.lz[x + t0.len - 1]);
                slide_core(h, x, t0.len, end);
                i += apcur + t0.len;
                x += t0.len;
                in_round = false;
                continue;
            }
            const u32 base = lv_get(wp, 0);
            // lit_ctx = wnd_curpos_ ? wnd_[wnd_curpos_ - 1] : 0 (csc_lz.cpp:297): offset 0 of every lap has no context
            const u32 cprice = literal_price_w(ns, (h.wnd ? x % h.wnd : x) ? h.lz[x - 1] : 0, lit);
            const u32 len = t0.len;
            if (apcur + len >= apend) apend = apcur + len + 1;
            LV_FOR(j) {
                u32 p = wp.v[LV_AT(j)], dd = wd.v[LV_AT(j)], bb = wb.v[LV_AT(j)];
                if (j == 1) {
                    if (cprice + base < p) { dd = 0; bb = apcur; p = cprice + base; }
                    if (t1d && t1p + base < p) { dd = 1; bb = apcur; p = t1p + base; }
                } else if (j >= 2 && j <= len) {
                    const u32 tdd = td.v[LV_AT(j)], tpp = tp.v[LV_AT(j)];
                    if (tdd && tpp + base < p) { dd = tdd; bb = apcur; p = tpp + base; }
                }
                wp.v[LV_AT(j)] = p; wd.v[LV_AT(j)] = dd; wb.v[LV_AT(j)] = bb;
            }
            lv_shift_down(wp, 0xFFFFFFFFu);
            lv_shift_down(wd, 0);
            lv_shift_down(wb, 0);
            apcur++;
            x++;
        }
    }

    // ------------------------------------------------------------ run encoders
    CSC_HDN void lz_run(u32 a, u32 size) {   // LZ::EncodeNormal mode 2/3, csc_lz.cpp:61-100
        for (u32 off = 0; off < size;) {
            u32 sb = size - off;
            if (sb > kBlock) sb = kBlock;
            if (wrap) {   // a sub-block never straddles the window end (csc_lz.cpp:65-66)
                const u32 to_wrap = S.cfg.wnd_size - (a + off) % S.cfg.wnd_size;
                if (to_wrap < sb) sb = to_wrap;
            }
            if (S.cfg.lz_mode == 3) parse_optimal(a + off, sb);
            else {
#if defined(__CUDA_ARCH__)
                // record path when the sub-block before was mostly literals (or the run is typed EXE), else the table path
                const bool rec = use_rec && !run_text && S.T && !wrap && !tmode && !tracing;
                const u32 nl = rec ? parse_lazy_t<true>(a + off, sb) : parse_lazy_t<false>(a + off, sb);
                use_rec = 3 * nl > sb;
#else
                parse_lazy_t<true>(a + off, sb);
#endif
            }
            off += sb;
            if (err) return;
        }
        emit(OP_MATCH, 64, 0, 0x100);   // end marker (:95-98), literal context untouched
    }
    CSC_HDN void skip_run(u32 a, u32 size) {   // mode 5: SlidePosFast (csc_mf.cpp:208-241)
        { const MfCtx h = mfctx(); sum_mark_range(h, a, a + size); }
        if (W) bits_or_not_range(S.D6, S.Z2, a, a + size);
        bits_set_range(S.D23, a, a + size);   // never enters the 2-/3-byte tables
    }
    // A statically excluded block is LZ-parsed after all (duplicate-block flip): the chains do not know its
    // positions, so rebuild the real 2-/3-byte tables from the history [0, upto) and use them from here on.
    CSC_HDN void enter_table_mode(u32 upto) {
        if (wrap) { err = kErrUnsupported; return; }   // cannot happen: a wrapped stream re-plans instead (compress_segment)
        u32 a = 0;
        for (u32 sg = 0; sg < S.nseg && a < upto; sg++) {
            u32 sz = S.seg_size[sg];
            for (u32 off = 0; off < sz && a + off < upto; off += kBlock) {
                u32 sb_end = a + off + (sz - off < kBlock ? sz - off : kBlock);
                u32 stop_at = sb_end < upto ? sb_end : upto;
#pragma unroll 1
                for (u32 y = a + off + CSC_LANE; y < stop_at; y += CSC_NLANES) {
                    if (bit_get(S.D23, y)) continue;
                    u32 w = word_at(mfctx(), y, sb_end);
                    g_max(&S.ht2[h2_of(w)], vld + y);
                    g_max(&S.ht3[h3_of(w)], vld + y);
                }
            }
            a += sz;
        }
        w_sync();
        tmode = true;
    }
    CSC_HDN void bad_run(u32 a, u32 size) {   // CompressBad, csc_model.cpp:463-469
        encode_int(size);
        // `size` raw bytes through the bit coder (csc_coder.cpp:76-87) are the run itself shifted by the bits that
        // were pending in front of it: output byte i = pending/previous byte's low k bits, then the top 8 - k bits
        // of byte i.  The block completions inside the run all carry the same time stamp (no decision in between).
        const u32 k = bc_bits, stamp = 2 * (dec_n - seg_dec0);
        const u32 pend = bc_val & ((1u << k) - 1);
        if (bc_n + size > S.bc_cap) { err = kErrOverflow; return; }
        for (u32 i = CSC_LANE; i < size; i += CSC_NLANES) {
            const u32 v = S.lz[a + i], p = i ? S.lz[a + i - 1] : pend;
            S.bc_out[bc_n + i] = (u8)((p << (8 - k)) | (v >> k));
        }
        const u32 tot = bc_blk + size, nev = tot / S.cfg.csc_blocksize;
        if (bev_n + nev > S.ev_cap) err = kErrOverflow;
        else for (u32 e = CSC_LANE; e < nev; e += CSC_NLANES) S.bc_ev[bev_n + e] = stamp;
        bev_n += nev;
        bc_blk = tot % S.cfg.csc_blocksize;
        bc_n += size;
        if (size) bc_val = S.lz[a + size - 1];
        w_sync();
    }
#if defined(__CUDA_ARCH__)
    // A batch of binary decisions, one per lane (`on`), in lane order.  Lanes that name the same probability see
    // it as the earlier lanes of the batch leave it (the adaptation is replayed from their bits), and the last of
    // them stores the result: exactly the serial statement, whatever the collisions.
    __device__ __forceinline__ void batch_decisions(bool on, u16 *pp, u32 bit, u32 n) {
        const u32 lane = CSC_LANE, FULL = 0xFFFFFFFFu;
        const unsigned long long key = on ? (unsigned long long)(uintptr_t)pp : (unsigned long long)lane;
        const u32 peers = __match_any_sync(FULL, key);
        const u32 bitsm = __ballot_sync(FULL, on && bit);
        if (on) {
            u32 v = *pp;
            u32 earlier = peers & ((1u << lane) - 1);
            while (earlier) {
                const u32 l = (u32)__ffs((int)earlier) - 1;
                earlier &= earlier - 1;
                v = ((bitsm >> l) & 1u) ? v + ((0xFFFu - v) >> 5) : v - (v >> 5);
            }
            S.dec[dec_n + lane] = (DecEntry)(v | (bit << 15));
            if ((peers >> lane) == 1u) *pp = (u16)(bit ? v + ((0xFFFu - v) >> 5) : v - (v >> 5));   // last user of the slot
        }
        dec_n += n;
        __syncwarp();
    }
#endif
    CSC_HDN void entropy_run(u32 a, u32 size) {   // CompressLiterals, csc_model.cpp:448-461
        encode_int(size);
        flush_bits();
#if defined(__CUDA_ARCH__)
        // four literals per step: lane 8 j + k codes tree level k of byte j under the row of the byte in front of it
        if (dec_n + 8 * size > S.dec_cap) { err = kErrOverflow; return; }
        const u32 lane = CSC_LANE, j = lane >> 3, kk = lane & 7;
#pragma unroll 1
        for (u32 i = 0; i < size; i += 4) {
            const u32 nb = size - i < 4 ? size - i : 4;
            const bool on = j < nb;
            u32 c = 0, pc = 0;
            if (on) { c = S.lz[a + i + j]; pc = (i + j) ? S.lz[a + i + j - 1] : 0; }
            if (i == 0 && j == 0) pc = ctx;
            const u32 cc = c | 0x100;
            batch_decisions(on, S.p_lit + pc * 256 + (cc >> (8 - kk)), (cc >> (7 - kk)) & 1, 8 * nb);
        }
        if (size) ctx = S.lz[a + size - 1];
#else
        for (u32 i = 0; i < size; i++) {
            u32 c = S.lz[a + i];
            u32 row = ctx * 256;
            ctx = c;
            lit_bits(row, c);
            flush_bits();
        }
#endif
    }
    // byte k of Forward_Delta's output (csc_filters.cpp:132-164) without materialising it
    CSC_HD u32 delta_byte(u32 a, u32 size, u32 chn, u32 k) const {
        if (size < 512) return S.lz[a + k];
        // channel c holds ceil((size - c)/chn) bytes
        u32 base = size / chn, rem = size % chn;
        u32 c, idx;
        u32 big = (base + 1) * rem;           // bytes in the first `rem` channels
        if (k < big) { c = k / (base + 1); idx = k % (base + 1); }
        else { c = rem + (k - big) / base; idx = (k - big) % base; }
        u32 j = c + idx * chn;
        u32 prev;
        if (idx) prev = S.lz[a + j - chn];
        else if (c == 0) prev = 0;
        else {   // last byte of channel c-1
            u32 cnt = base + ((c - 1) < rem ? 1 : 0);
            prev = S.lz[a + (c - 1) + (cnt - 1) * chn];
        }
        return (S.lz[a + j] - prev) & 0xFF;
    }
    // CompressRLE's run rule at residual i (csc_model.cpp:476-497): three equal neighbours and more than ten equal
    // bytes from i on.  Pure function of the residuals, so every lane can evaluate it for its own position.
    CSC_HDF static bool rle_starts(const u8 *f, u32 i, u32 size) {
        if (i == 0 || size - i <= 3) return false;
        const u32 c0 = f[i];
        if (f[i - 1] != c0 || f[i + 1] != c0 || f[i + 2] != c0) return false;
        if (i + 10 >= size) return false;
        for (u32 t = 3; t <= 10; t++) if (f[i + t] != c0) return false;
        return true;
    }
    CSC_HDN void delta_run(u32 a, u32 size, u32 chn) {   // Forward_Delta + CompressRLE, csc_model.cpp:471-513
        // the filtered run is materialised lane-parallel into the stream's scratch, then coded
        u8 *f = S.tmp + a;
#pragma unroll 1
        for (u32 k = CSC_LANE; k < size; k += CSC_NLANES) f[k] = (u8)delta_byte(a, size, chn, k);
        w_sync();
        encode_int(size);
        flush_bits();
        for (u32 i = 0; i < size;) {
#if defined(__CUDA_ARCH__)
            // up to three literals per step: lane 9 j codes the flag of residual i + j, lanes 9 j + 1 .. 9 j + 8 its
            // tree levels under the row of the residual in front of it; a residual that starts a run ends the batch
            const u32 lane = CSC_LANE, j = lane / 9, kk = lane - 9 * j;
            const bool mine = j < 3 && i + j < size;
            const u32 runm = __ballot_sync(0xFFFFFFFFu, mine && kk == 0 && rle_starts(f, i + j, size));   // bits 0, 9, 18
            u32 nb = size - i < 3 ? size - i : 3;
            if (runm & 1u) nb = 0; else if (runm & (1u << 9)) nb = 1; else if ((runm & (1u << 18)) && nb > 2) nb = 2;
            if (nb) {
                if (dec_n + 9 * nb > S.dec_cap) { err = kErrOverflow; return; }
                const bool on = j < nb;
                u32 c = 0, pc = 0;
                if (on) { c = f[i + j]; pc = (i + j) ? f[i + j - 1] : 0; }
                const u32 cc = c | 0x100;
                u16 *pp = kk == 0 ? &M.p_rle_flag : S.p_delta + pc * 256 + (cc >> (9 - kk));
                batch_decisions(on, pp, kk == 0 ? 0u : (cc >> (8 - kk)) & 1u, 9 * nb);
                i += nb;
                continue;
            }
            {
                const u32 c0 = f[i];
                u32 jj = i + 11;   // rle_starts: f[i .. i + 10] are equal
                while (jj < size && f[jj] == c0) jj++;
                enc_bit(M.p_rle_flag, 1);
                encode_matchlen_2(jj - i - 11);
                flush_bits();
                i = jj;
            }
#else
            const u32 c0 = f[i];
            if (rle_starts(f, i, size)) {
                u32 j = i + 11;
                while (j < size && f[j] == c0) j++;
                enc_bit(M.p_rle_flag, 1);
                encode_matchlen_2(j - i - 11);
                flush_bits();
                i = j;
                continue;
            }
            enc_bit(M.p_rle_flag, 0);
            lit_bits(65536 + (i ? f[i - 1] : 0u) * 256, c0);   // p_delta follows p_lit; the context is the residual in front
            flush_bits();
            i++;
#endif
        }
    }

    // IsDuplicateBlock / TestFind (csc_lz.cpp:102-112, csc_mf.cpp:526-568): window ends at run_start
    CSC_HDN bool dup_block(u32 b0, u32 bsize, u32 run_start) {
        if (!W) return false;   // -m5 head table: handled by the bounded search path
        bool hit = false;
        for (u32 base = 0; base + 18 < bsize && !hit; base += CSC_NLANES) {
            u32 off = base + CSC_LANE;
            bool mine = false;
            if (off + 18 < bsize) {
                u32 p = b0 + off;
                if (bit_get_ro(S.Z2, p)) {
                    u32 e = newest_alive6(p, run_start);
                    if (e != kNone && run_start - e < vld) {
                        u32 limit = bsize - off, k = 0;
                        const u8 *a = S.lz + p, *m = S.lz + e;
                        const u32 valid = run_start - e;   // bytes beyond are not written yet: zero on the first lap,
                        const u32 wsz = S.cfg.wnd_size;     // the previous lap's bytes afterwards; compares stop at the window end
                        if (wrap) { const u32 c = wsz - e % wsz; if (c < limit) limit = c; }
                        while (k < limit) {
                            u8 mv = k < valid ? m[k] : (e + k >= wsz ? S.lz[e + k - wsz] : (u8)0);
                            if (a[k] != mv) break;
                            k++;
                            if (k > 18) break;
                        }
                        mine = k > 18;
                    }
                }
            }
            hit = w_any(mine);
        }
        return hit;
    }

    CSC_HDN void compress_run(u32 a, u32 size, u32 type, u32 first_blk) {   // compress_block, csc_encoder_main.cpp:35-83
        if (!size) return;
        if (want_tables && !tmode && type < DT_NO_LZ) enter_table_mode(a);
        run_text = type == DT_ENGTXT;
        if (type == DT_NORMAL || type == DT_FAST) {
            emit(OP_INT, DT_NORMAL, 0, 0);
            lz_run(a, size);
        } else if (type == DT_EXE) {
            emit(OP_INT, DT_EXE, 0, 0);
            use_rec = true;   // machine code: literal-heavy
            lz_run(a, size);
        } else if (type == DT_ENGTXT) {
            if (S.dict_ok[first_blk]) { emit(OP_INT, DT_ENGTXT, 0, 0); emit(OP_INT, size, 0, 0); }
            else emit(OP_INT, DT_NORMAL, 0, 0);
            lz_run(a, size);
        } else if (type == DT_BAD) {
            emit(OP_INT, DT_BAD, 0, 0);
            skip_run(a, size);
            emit(OP_BAD, a, size, 0);
        } else if (type == DT_ENTROPY) {
            emit(OP_INT, DT_ENTROPY, 0, 0);
            skip_run(a, size);
            emit(OP_ENTROPY, a, size, 0);
        } else {
            emit(OP_INT, type, 0, 0);
            skip_run(a, size);
            emit(OP_DELTA, a, size, dlt_channels(type - DT_DLT));
        }
    }

    // one Compress() call (csc_encoder_main.cpp:85-146)
    CSC_HDN void compress_segment(u32 seg, u32 a, u32 size) {
        u32 last_type = DT_NORMAL, last_begin = 0, last_size = 0;
        u32 bpb = 0;
        u32 blk = S.seg_blk0[seg];
        u32 last_blk = blk;
        for (u32 i = 0; i < size; blk++) {
            u32 cur = size - i < kBlock ? size - i : kBlock;
            u32 t = static_type(S.binfo, S.bbpb, S.tail_dlt + seg * 5, blk, last_type, bpb, S.cfg.flt_dlt, S.cfg.flt_txt, S.cfg.flt_exe);
            if (t >= DT_NO_LZ) {
                const bool dup = dup_block(a + i, cur, a + last_begin);
                // The static plan (plan.h) assumed an outcome for this test.  Once the window wraps, run ends
                // move the sub-block grid the match-finder kernels hashed over, so a wrong guess is fatal
                // there: report the block and let the host plan again with the observed outcome.
                if (wrap && dup != (S.flip && S.flip[blk] != 0)) {
                    if (S.replan) { S.replan[0] = blk; S.replan[1] = dup ? 1u : 0u; }
                    err = kErrReplan;
                    return;
                }
                if (dup) t = DT_NORMAL;
            }
            if (t < DT_NO_LZ && S.blk_nolz[blk] && !want_tables) {
                // the pending run (if it continues into this block) must already be parsed with tables
                want_tables = true;
            }
            if (last_type != t) {   // (size limit of :128 cannot trigger inside one Read())
                if (last_size) {
                    compress_run(a + last_begin, last_size, last_type, last_blk);
                    emit(OP_INT, 0, 0, 0);
                }
                last_begin = i;
                last_size = 0;
                last_blk = blk;
            }
            last_type = t;
            last_size += cur;
            i += cur;
            if (err) return;
        }
        if (last_size) {
            compress_run(a + last_begin, last_size, last_type, last_blk);
            emit(OP_INT, 1, 0, 0);
            emit(OP_FLUSH, seg, 0, 0);
        }
    }

    CSC_HD void init_common() {
        vld = S.cfg.vld_rge; W = S.cfg.ht_width; K = S.K; good = S.cfg.good_len;
        err = 0; tr_n = 0;
        tmode = false; want_tables = false; use_rec = false; run_text = false;
        wrap = S.n > S.cfg.wnd_size;
        tracing = S.trace != nullptr;
        gs = 5;
        while (((S.n + 63) >> gs) > 2048) gs++;   // 2048 summary bits cover the stream
        w_sync();
        for (u32 i = CSC_LANE; i < 64; i += CSC_NLANES) M.dsum[i] = 0;
        if (CSC_LANE == 0) M.dead_hi = 0;
#if defined(__CUDA_ARCH__) && CSC_TMA_STAGE
        ts_reset(ts);
#endif
        w_sync();
        dec_n = bc_n = bev_n = 0;
        seg_dec0 = seg_bc0 = seg_bev0 = 0;
        coder_reset();
        state = 0; ctx = 0; lp_rebuild = 0;
        rep[0] = rep[1] = rep[2] = rep[3] = S.cfg.wnd_size;   // csc_lz.cpp:46-49
    }
    CSC_HD void parse_stream() {
        u32 a = 0;
        for (u32 s = 0; s < S.nseg && !err; s++) {
            compress_segment(s, a, S.seg_size[s]);
            a += S.seg_size[s];
        }
    }
    // whole stream in one warp
    CSC_HD void run() {
        init_common();
        model_reset();
        parse_stream();
        if (CSC_LANE == 0) {
            *S.status = err;
            if (S.trace_n) *S.trace_n = tr_n;
        }
    }
};

}  // namespace cscb
