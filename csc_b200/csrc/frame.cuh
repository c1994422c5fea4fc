// frame.cuh -- MemIO framing of one encoded stream (csc_memio.cpp:83-108), usable on the device.
//
// The engines leave, per flush segment, the RC and BC sub-streams plus the time stamps at which their 64 KiB
// blocks filled (csc_defs.h: SegOut).  The libcsc stream is those blocks in temporal order (H4), each behind a
// 1-byte header (bit 7: RC, bit 6: full block) and, when not full, a 3-byte size; the last RC block of a
// segment ends with the byte Coder::Flush skips instead of storing (H3): whatever the never-cleared 64 KiB
// rc_buf_ held at that offset, i.e. the byte the most recent earlier writer left there.  The writers are the
// full RC blocks (all offsets) and the segments' remainders (offsets below their length - 1), so the stale byte
// is found by walking the stream's own segments backwards -- no ring image is kept.
//
// frame_size / frame_write are the same statement host_common.h: Framer makes for the handle API; the batched
// calls run them in k_frame (one warp per stream) so that the download is the finished streams and the host
// only copies them out.  The host simulation runs them with a single lane against the oracle.
#pragma once
#include "csc_defs.h"
#include "warp.h"

namespace cscb {

struct FrameStream {
    const u8 *rc, *bc;            // sub-streams of the stream (SegOut offsets are relative to these)
    const u32 *rc_ev, *bc_ev;     // block-completion time stamps of the stream
    const SegOut *so;
    u32 nseg;
    u32 bsize;                    // csc_blocksize
    u32 with_props;
    u8 props[12];                 // CSC_PROP_SIZE bytes the caller writes in front of the stream
    const i32 *status;            // *status < 0: the stream failed, nothing is framed
};

CSC_HD u32 frame_size(const FrameStream &F) {
    if (*F.status < 0) return 0;
    u32 t = F.with_props ? 10u : 0u;
    for (u32 s = 0; s < F.nseg; s++) {
        const SegOut o = F.so[s];
        const u32 rrem = o.rc_len - o.rev_len * F.bsize, brem = o.bc_len - o.bev_len * F.bsize;
        t += (o.rev_len + o.bev_len) * (1 + F.bsize) + (rrem == F.bsize ? 1u : 4u) + rrem + (brem == F.bsize ? 1u : 4u) + brem;
    }
    return t + 4 + 6 + 4 + 3;   // EOF segment
}

// header of one block at out[o...]; returns the payload position
CSC_HDF u32 frame_header(u8 *out, u32 o, u32 size, u32 bsize, u32 rc) {
    if (CSC_LANE == 0) {
        out[o] = (u8)((rc << 7) | (size == bsize ? 0x40u : 0u));
        if (size != bsize) { out[o + 1] = (u8)(size >> 16); out[o + 2] = (u8)(size >> 8); out[o + 3] = (u8)size; }
    }
    return o + (size == bsize ? 1u : 4u);
}
CSC_HDF void frame_copy(u8 *dst, const u8 *src, u32 n) {
    for (u32 i = CSC_LANE; i < n; i += CSC_NLANES) dst[i] = src[i];
}
// the byte rc_buf_ holds at offset k: when segment `upto` flushes (flushing: its own remainder has not stored that offset
// yet), or after segment `upto` is complete (!flushing: the EOF segment's flush).  0 if nothing was ever written there.
CSC_HD u32 frame_stale(const FrameStream &F, u32 upto, u32 k, bool flushing) {
    for (u32 s = upto + 1; s-- > 0;) {
        const SegOut o = F.so[s];
        const u32 full = o.rev_len * F.bsize, rrem = o.rc_len - full;
        // chronological order inside a segment: its full blocks, then its remainder (rrem - 1 stored bytes)
        if (!(flushing && s == upto) && rrem - 1 > k) return F.rc[o.rc_off + full + k];
        if (o.rev_len) return F.rc[o.rc_off + full - F.bsize + k];
    }
    return 0;
}

CSC_HD void frame_write(const FrameStream &F, u8 *out) {
    if (*F.status < 0) return;
    u32 o = 0;
    if (F.with_props) { for (u32 i = CSC_LANE; i < 10; i += CSC_NLANES) out[i] = F.props[i]; o = 10; }
    for (u32 s = 0; s < F.nseg; s++) {
        const SegOut so = F.so[s];
        const u32 *re = F.rc_ev + so.rev_off, *be = F.bc_ev + so.bev_off;
        u32 i = 0, j = 0, rpos = 0, bpos = 0;
        while (i < so.rev_len || j < so.bev_len) {   // temporal order of the block completions: RC stamps are odd, BC stamps even
            if (j >= so.bev_len || (i < so.rev_len && re[i] < be[j])) {
                o = frame_header(out, o, F.bsize, F.bsize, 1);
                frame_copy(out + o, F.rc + so.rc_off + rpos, F.bsize);
                o += F.bsize; rpos += F.bsize; i++;
            } else {
                o = frame_header(out, o, F.bsize, F.bsize, 0);
                frame_copy(out + o, F.bc + so.bc_off + bpos, F.bsize);
                o += F.bsize; bpos += F.bsize; j++;
            }
        }
        const u32 rrem = so.rc_len - rpos, brem = so.bc_len - bpos;   // rrem >= 1: includes the skipped byte
        const u32 stale = frame_stale(F, s, rrem - 1, true);
        o = frame_header(out, o, rrem, F.bsize, 1);
        frame_copy(out + o, F.rc + so.rc_off + rpos, rrem - 1);
        if (CSC_LANE == 0) out[o + rrem - 1] = (u8)stale;
        o += rrem;
        o = frame_header(out, o, brem, F.bsize, 0);
        frame_copy(out + o, F.bc + so.bc_off + bpos, brem);
        o += brem;
    }
    // EncodeInt(SIG_EOF) + Flush from a reset coder (csc_encoder_main.cpp:149-157): five zero bytes + the stale one, then 0x19 0 0
    const u32 stale5 = F.nseg ? frame_stale(F, F.nseg - 1, 5, false) : 0;
    o = frame_header(out, o, 6, F.bsize, 1);
    if (CSC_LANE == 0) { for (u32 k = 0; k < 5; k++) out[o + k] = 0; out[o + 5] = (u8)stale5; }
    o += 6;
    o = frame_header(out, o, 3, F.bsize, 0);
    if (CSC_LANE == 0) { out[o] = 0x19; out[o + 1] = 0; out[o + 2] = 0; }
}

}  // namespace cscb
