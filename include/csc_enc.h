/* csc_enc.h -- encoder side of the libcsc C API, served by the B200 library.
 * Each entry point replaces the reference function of the same name
 * (src/libcsc/csc_enc.h:11-30, implementation src/libcsc/csc_enc.cpp). */
#ifndef _CSC_ENC_H_
#define _CSC_ENC_H_
#include "csc_common.h"

EXTERN_C_BEGIN

/* csc_enc.cpp:16-97: level (1..5) and dictionary size -> all parameters. */
#ifdef __cplusplus
void CSCEncProps_Init(CSCProps *p, uint32_t dict_size = 64000000, int level = 2);
#else
void CSCEncProps_Init(CSCProps *p, uint32_t dict_size, int level);
#endif

/* csc_enc.cpp:145-158: 10-byte stream header (written by the CALLER). */
void CSCEnc_WriteProperties(const CSCProps *props, uint8_t *stream, int full);

/* csc_enc.cpp:99-112 */
uint64_t CSCEnc_EstMemUsage(const CSCProps *props);

typedef void *CSCEncHandle;

/* csc_enc.cpp:114-133; alloc may be NULL. Returns NULL on failure. */
CSCEncHandle CSCEnc_Create(const CSCProps *props, ISeqOutStream *outstream, ISzAlloc *alloc);

/* csc_enc.cpp:135-143 */
void CSCEnc_Destroy(CSCEncHandle p);

/* csc_enc.cpp:160-191: pulls raw_blocksize bytes per Read() until *size == 0;
 * every Read() is one flush segment.  0 or a negative error code. */
int CSCEnc_Encode(CSCEncHandle p, ISeqInStream *instream, ICompressProgress *progress);

/* csc_enc.cpp:193-203: EOF marker + final coder flush. */
int CSCEnc_Encode_Flush(CSCEncHandle p);

EXTERN_C_END

#endif
