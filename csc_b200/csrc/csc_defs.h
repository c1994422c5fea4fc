// csc_defs.h -- shared host/device definitions for the B200 libcsc hot path.
//
// Vocabulary follows the reference's domain: a *stream* is one CSCEnc handle's
// output (one chunk of csarc's -p split), a *segment* is the data of one
// Read()/Compress() call (csc_enc.cpp:169-176), a *block* is the analyzer's
// 8 KiB unit (csc_typedef.h:10), a *run* is a maximal sequence of equal-typed
// blocks coded as one unit (csc_encoder_main.cpp:128-135).
#pragma once
#include <stdint.h>
#include <stddef.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;

#if defined(__CUDACC__)
// CSC_HD: ordinary function (inlined into its caller unless CSC_OUTLINE_ALL), CSC_HDF: always inline,
// CSC_HDN: always out of line.  The serial engines are large; CSC_HDN marks the boundaries where
// inlining would duplicate the hot loops (code size costs instruction-cache misses).
#ifdef CSC_OUTLINE_ALL
#define CSC_HD __host__ __device__ __noinline__
#else
#define CSC_HD __host__ __device__ inline
#endif
#define CSC_HDF __host__ __device__ __forceinline__
#define CSC_HDN __host__ __device__ __noinline__
#else
#define CSC_HD inline
#define CSC_HDF inline
#define CSC_HDN
#endif

#if !defined(__CUDACC__)
struct uint4 { unsigned int x, y, z, w; };   // host simulation only
#endif

namespace cscb {

// ---- block types (csc_typedef.h:20-41) -------------------------------------
enum : u32 {
    DT_NONE = 0x00, DT_NORMAL = 0x01, DT_ENGTXT = 0x02, DT_EXE = 0x03, DT_FAST = 0x04,
    DT_NO_LZ = 0x05, DT_ENTROPY = 0x07, DT_BAD = 0x08, SIG_EOF = 0x09, DT_DLT = 0x10,
    DLT_CHANNEL_MAX = 5, DT_SKIP = 0x1E
};
static const u32 kBlock = 8192;          // MinBlockSize
static const u32 kNone = 0xFFFFFFFFu;    // "no earlier occurrence" in chain arrays
static const u32 kApLimit = 2048;        // AP_LIMIT (csc_lz.h:43)
// -m5: the binary-tree search is replaced by a bounded per-position search: the nearest earlier positions
// with equal 6-byte hash (kBtKA), equal 8-byte context hash (kBtKB) and equal 12-byte context hash (kBtKC),
// merged by distance into one row of kBtRowK candidates.
static const u32 kBtKA = 5, kBtKB = 5, kBtKC = 6, kBtRowK = 16, kBtCtxB = 8, kBtCtxC = 12, kBtHashBits = 24;
static const u32 kMaxCand = 32;          // MF_CAND_LIMIT (csc_mf.h:33)

CSC_HD u32 dlt_channels(u32 idx) {       // DltIndex[] (csc_typedef.h:35)
    return idx == 0 ? 1u : idx == 1 ? 2u : idx == 2 ? 3u : idx == 3 ? 4u : 8u;
}

// ---- per-stream encoder configuration (subset of CSCProps that matters) ----
struct EncCfg {
    u32 wnd_size;      // LZ window = CSCProps::dict_size (csc_lz.cpp:15-20)
    u32 vld_rge;       // wnd_size - 8192 - 4 (csc_mf.cpp:56)
    u32 good_len;
    u32 ht_width;      // 6-byte table width W (0 for -m5)
    u32 ht_bits;       // 6-byte table hash bits
    u32 lz_mode;       // 2 lazy, 3 price-optimal
    u32 flt_dlt, flt_txt, flt_exe;
    u32 csc_blocksize; // RC/BC framing block, 64 KiB default
    u32 bt_bits, bt_size, bt_cyc;  // -m5 as configured by the reference (zeroed once mapped to the bounded search)
    u32 bounded;       // 1: -m5 bounded multi-context search (ratio-gated, not stream-identical)
};

// ---- optimal-parser unit (role of APUnit, csc_lz.h:33-41) -------------------
struct alignas(16) ApUnit {   // one position of the optimal parser's path (APUnit, csc_lz.h:45-60); 2 x 16 bytes
    u32 rep[4];
    u32 dist;
    u16 back_pos;
    u16 next_pos;
    u8 state;
    u8 lit;
    u16 pad;
    u32 pad2;              // prices of the look-ahead window live in lanes, not here
};


struct Cand { u32 len; u32 dist; };

// ---- table candidates of one position, precomputed (k_mf_tlist) ----------------------------------------------
// What find_match (csc_mf.cpp:303-485) takes from the 2-/3-/6-byte tables at position x does not depend on the
// parse as long as every table entry it looks at is still in the tables ("alive"): the distance filter, the
// length filter started at 1 and the bound[] cut are functions of the candidate rows alone.  The rep-distance
// probes in front of it (:266-299) only raise the starting length R, which drops the entries with len <= R, or
// stop the search altogether.  TRec holds that list (strictly increasing lengths, cut at the first good_len
// match); `far` is the largest distance of any entry consulted, so "all alive" is checked against the stream's
// highest dead position.  cnt == kTOver: more than four entries, take the general path.
struct alignas(16) TRec {
    u32 d[4];
    u16 l[4];
    u32 far;
    u32 cnt;
};
static const u32 kTOver = 0xFF;   // MFUnit: dist code 0 literal, 1..4 rep, 4+d match

// ---- small adaptive-probability tables (csc_model.h:95-123), 12-bit in u16 --
struct ModelSmall {
    u16 p_state[64 * 3];
    u16 p_repdist[64 * 3];
    u16 p_dist[8 + 16 * 2 + 32 * 4];
    u16 p_matchlen_slot[2];
    u16 p_matchlen_extra1[8];
    u16 p_matchlen_extra2[8];
    u16 p_matchlen_extra3[128];
    u16 p_matchdist_extra[29 * 16];
    u16 p_longlen;
    u16 p_rle_flag;
    u32 len_price[32];
    const u32 *p2bits;    // p_2_bits_ (csc_model.cpp:68-70), host-libm derived: 512 entries, one copy per CTA (shared memory is L1 the warps need)
    u32 row_e[8];         // scratch: candidate row of the current position (lane-gathered)
    u32 row_l[8];
    u32 cl[kMaxCand];     // scratch: candidate list of the current find_match (len, dist code)
    u32 cd[kMaxCand];
    u32 dead_hi;          // 1 + highest position ever marked dead in D6/D23 (0: none): entries at or above it are alive
    u32 dsum[64];         // coarse summary of the liveness bitmaps D6/D23: bit g set = some position of granule g may be dead
    u32 bq[32];           // pending range-coder decisions of the current packet: probability slot ids
    u16 bq_p[32];         //   ... and the probabilities fetched for them (lane-parallel)
};
static const u32 kSlotGlobal = 0x80000000u;   // slot id flag: index into p_lit/p_delta (global), else u16 index into ModelSmall

// Per-segment output record, read by host framing.  The parser / model engine (enc_engine.cuh) fills the
// decision and raw-bit fields, the range-coder pass (rc_engine.cuh) the rc_* fields.  Block completions are
// logged with a time stamp so that the host can replay their temporal order (H4): 2 * d for a BC block that
// filled after d decisions of the segment, 2 * i + 1 for an RC block that filled while decision i was coded
// (i = dec_len for the flush bytes).
struct SegOut {
    u32 rc_off, rc_len;     // bytes of the range-coder sub-stream of this segment
    u32 bc_off, bc_len;     // bytes of the raw-bit sub-stream of this segment
    u32 dec_off, dec_len;   // binary decisions of this segment (DecEntry), dec_off 8-aligned
    u32 rev_off, rev_len;   // RC block completions (time stamps)
    u32 bev_off, bev_len;   // BC block completions (time stamps)
};
// one binary decision handed from the model to the range coder: probability (12 bits) | bit << 15
typedef u16 DecEntry;

// per-stream descriptor handed to the encoder engine (all device pointers)
struct EncStream {
    EncCfg cfg;
    const u8 *lz;          // LZ stream: input after the static e8e9/dict filters
    u32 n;                 // stream bytes
    u8 *tmp;               // scratch of n bytes (dictionary filter staging, delta-filtered runs)
    u32 nseg;
    const u32 *seg_size;   // [nseg] Read() sizes
    const u32 *seg_blk0;   // [nseg] first block index of each segment
    const u32 *binfo;      // [nblk] analyzer type | dlt_fail<<8
    const u32 *bbpb;       // [nblk] order-0 bpb*100 (csc_analyzer.cpp:207)
    const u32 *tail_dlt;   // [nseg*5] delta-bpb of a <512 B tail block per channel idx
    const u8 *dict_ok;     // [nblk] dictionary filter accepted the TXT run starting here
    // match-finder precomputation ("last-k prior occurrences per position")
    const u32 *P6;         // [n*K] row p = the K nearest earlier positions with HASH6 equal
    const u16 *L6;         // [n*K] forward match length of (p, P6[p][k]) capped at sub-block end
    u32 K;                 // row depth (= ht_width)
    const u32 *Z2;         // bitmap: HASH2 % 16 == 0 (SlidePosFast / TestFind sampling)
    const u32 *Z6;         // bitmap: HASH6 == 0 (lasth6 start value, csc_mf.cpp:136)
    u32 *D6;               // bitmap: position not (or no longer) in the 6-byte table
    // 2-/3-byte single-slot tables as chains: nearest earlier position with equal HASH3 / HASH2 among
    // the blocks the analyzer typed as LZ-parsed, + forward match length; D23 marks positions that
    // the reference never inserted (SlidePos skip zone, SlidePosFast runs)
    const u32 *P3; const u16 *L3;
    const u32 *P2; const u16 *L2;
    u32 *D23;
    u32 *B6;               // bitmap (with T): HASH6 == 0 or == the previous position's HASH6
    u32 *E6;               // bitmap (with T): the record holds table candidates (cnt != 0)
    TRec *T;               // [n] precomputed table candidates (k_mf_tlist); null: general path only (wrapped streams)
    const u8 *blk_nolz;    // [nblk] block left out of the 2-/3-byte chains (statically typed BAD/ENTROPY/DLT)
    const u8 *flip;        // [nblk] outcome of the duplicate-block test the static plan assumed (plan.h); may be null
    u32 *replan;           // [2] out: block and observed outcome when the assumption failed (status kErrReplan)
    u32 *ht2;              // [16 Ki] 2-byte table, emulated directly (csc_mf.h:17-18)
    u32 *ht3;              // [64 Ki] 3-byte table
    // model in global memory
    u16 *p_lit;            // [256*256]
    u16 *p_delta;          // [256*256]
    ApUnit *ap;            // [kApLimit+1]
    // outputs
    u8 *rc_out; u32 rc_cap;
    u8 *bc_out; u32 bc_cap;
    DecEntry *dec; u32 dec_cap;          // decisions of the whole stream, in coding order
    u32 *rc_ev; u32 *bc_ev; u32 ev_cap;  // block-completion time stamps (each ev_cap entries)
    SegOut *segout;        // [nseg]
    i32 *status;           // 0 ok, <0 error
    u32 *trace; u32 trace_cap; u32 *trace_n;   // optional symbol trace (debug), may be null
};

// ---- decoder -----------------------------------------------------------------
struct DecRun { u32 type; u32 off; u32 size; };

struct DecStream {
    u32 wnd_size;
    u32 raw_blocksize;
    const u8 *rc; u32 rc_len;            // concatenated RC block payloads
    const u8 *bc; u32 bc_len;            // concatenated BC block payloads
    const u32 *rc_blk_off; u32 n_rc_blk; // payload start of every RC block (+ sentinel)
    const u32 *bc_blk_off; u32 n_bc_blk;
    u8 *wnd;  u32 out_cap;               // linear LZ history (filtered domain)
    u16 *p_lit; u16 *p_delta;
    DecRun *runs; u32 runs_cap; u32 *n_runs;
    u32 *out_len;
    i32 *status;
};

static const i32 kErrDecode = -96, kErrWrite = -97, kErrRead = -98, kErrOverflow = -90, kErrUnsupported = -91;
static const i32 kErrReplan = -93;   // internal: the static plan of a wrapped stream has to be redone (never leaves the library)

}  // namespace cscb
