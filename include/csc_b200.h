/* csc_b200.h -- additive batched entry points of the B200 libcsc library.
 *
 * One libcsc handle is one serially dependent stream, so a GPU
 * is only used well when many independent streams are in flight: exactly the
 * units csarc creates with its -p / -t task split (archiver/csarc.cpp:532-543,
 * one CSCEnc/CSCDec handle per task in archiver/csa_worker.cpp:23-90).  These
 * calls take many chunks at once and produce, per chunk, EXACTLY the bytes the
 * per-handle API (csc_enc.h / csc_dec.h) produces -- they are additive, plain
 * C-ABI (pointers and sizes, no CUDA or torch types), and never touch CSCProps'
 * layout.  Host buffers may be pageable or pinned; device selection is by
 * CSCB200_SetDevice / CSCB200_SetDevices (default: current CUDA device).
 */
#ifndef _CSC_B200_H_
#define _CSC_B200_H_
#include "csc_common.h"

EXTERN_C_BEGIN

/* Worst-case size of one encoded chunk (props header included). */
size_t CSCB200_EncodeBound(size_t in_size);

/* Encode n independent chunks.  Chunk i is coded exactly like
 *   CSCEncProps_Init(&p, min(dict_size, in_size[i]), level);      (csa_worker.cpp:35)
 *   [p.DLTFilter/TXTFilter/EXEFilter from flt_mask bits 0/1/2]
 *   CSCEnc_WriteProperties (if with_props) ; CSCEnc_Encode with Read() == p.raw_blocksize ;
 *   CSCEnc_Encode_Flush.
 * out[i] receives out_size[i] <= out_cap[i] bytes.  Returns 0 or a negative code
 * (READ_ERROR/WRITE_ERROR family, -90 output overflow, -91 unsupported, -92 CUDA failure). */
int CSCB200_EncodeChunks(int n, const uint8_t *const *in, const size_t *in_size, uint32_t dict_size, int level,
                         int flt_mask, int with_props, uint8_t *const *out, const size_t *out_cap, size_t *out_size);

/* Decode n independent libcsc streams (each starting with its 10-byte props header), the
 * counterpart of one CSCDec handle per task (csa_worker.cpp:59-90). */
int CSCB200_DecodeChunks(int n, const uint8_t *const *in, const size_t *in_size, uint8_t *const *out,
                         const size_t *out_cap, size_t *out_size);

/* Device-resident variant used for kernel-only timing: the chunk payloads are already in HBM
 * (d_in[i] are device pointers), the coded sub-streams stay in HBM, and only the size of each
 * framed stream is returned.  Same kernels, same work as CSCB200_EncodeChunks. */
int CSCB200_EncodeChunksDevice(int n, const uint8_t *const *d_in, const size_t *in_size, uint32_t dict_size, int level,
                               int flt_mask, int with_props, size_t *out_size);

/* csarc keeps an Adler-32 per file (archiver/csa_adler32.cpp:63-128, taken over the bytes it reads, csa_io.h:250, and
 * over the bytes it writes back, csa_io.h:342).  These variants return it as a by-product of the pass the data makes
 * through HBM anyway: adler_in[i] = adler32(1, in[i], in_size[i]) of the raw chunk, adler_out[i] = adler32(1, ...) of
 * the decoded chunk.  CSCB200_Adler32 is the stand-alone form, adler32(init[i], buf[i], size[i]) (init NULL: 1);
 * checksums of consecutive pieces of a file chain through init exactly like the reference's running value. */
int CSCB200_EncodeChunksAdler(int n, const uint8_t *const *in, const size_t *in_size, uint32_t dict_size, int level,
                              int flt_mask, int with_props, uint8_t *const *out, const size_t *out_cap, size_t *out_size,
                              uint32_t *adler_in);
int CSCB200_DecodeChunksAdler(int n, const uint8_t *const *in, const size_t *in_size, uint8_t *const *out,
                              const size_t *out_cap, size_t *out_size, uint32_t *adler_out);
int CSCB200_Adler32(int n, const uint8_t *const *buf, const size_t *size, const uint32_t *init, uint32_t *out);

/* GPU selection.  Every call runs on a device context leased from a per-GPU pool for the duration of the call, so
 * handles and batched calls may be used from any number of threads (csarc: one handle per worker thread).
 *   CSCB200_SetDevice(d)      this thread's calls run on GPU d (-1: the thread's current CUDA device, the default)
 *   CSCB200_SetDevices(v, n)  this thread's batched calls split their chunks into n contiguous ranges, one per GPU
 *                             of v, coded side by side (chunk c -> v[c * n / chunks]; no collective, results stay
 *                             in chunk order); handles created by this thread are spread round-robin over v
 * Without either call the environment variable CSC_B200_DEVICES ("all" or "0,1,5") plays the role of SetDevices:
 * that is how an unmodified caller is given several GPUs. */
int CSCB200_SetDevice(int device);
int CSCB200_SetDevices(const int *devices, int n);
int CSCB200_DeviceCount(void);
/* Chunks [*lo, *hi) of n_chunks that part k of g owns (chunk c -> part c * g / n_chunks): the partition of the
 * fan-out above, for launchers that run one process per GPU instead (needs no device). */
void CSCB200_ChunkRange(int n_chunks, int k, int g, int *lo, int *hi);
/* Frees the device and pinned memory cached by idle contexts. */
void CSCB200_Trim(void);

/* Timing / launch statistics of the last Encode/Decode call on this thread's context. */
typedef struct {
    double ms_h2d, ms_analyze, ms_filters, ms_mf_hash, ms_mf_sort, ms_mf_rows, ms_mf_extend, ms_encode, ms_d2h,
        ms_frame, ms_decode, ms_unfilter, ms_total_gpu, ms_mf_chains,   /* mf_hash..extend: 6-byte rows; mf_chains: 3-/2-byte chains */
        ms_rc;                                                          /* ms_encode: parser + model; ms_rc: range-coder pass */
    uint64_t bytes_in, bytes_out, mf_positions;
    uint32_t kernel_launches, streams, waves;
    uint32_t replans;   /* streams encoded again because a duplicate-block outcome moved the sub-block grid of a wrapped window */
} CSCB200_Stats;
void CSCB200_GetStats(CSCB200_Stats *out);

/* 1 if a CUDA device is usable, else 0 (the library has no CPU fallback). */
int CSCB200_Available(void);

EXTERN_C_END
#endif
