/* csc_dec.h -- decoder side of the libcsc C API, served by the B200 library
 * (reference: src/libcsc/csc_dec.h:8-21, implementation src/libcsc/csc_dec.cpp). */
#ifndef _CSC_DEC_H_
#define _CSC_DEC_H_
#include "csc_common.h"

EXTERN_C_BEGIN

/* csc_dec.cpp:733-738: parse the 10-byte header the caller read. */
void CSCDec_ReadProperties(CSCProps *props, uint8_t *stream);

typedef void *CSCDecHandle;

/* csc_dec.cpp:692-720; alloc may be NULL. Returns NULL on failure. */
CSCDecHandle CSCDec_Create(const CSCProps *props, ISeqInStream *instream, ISzAlloc *alloc);

/* csc_dec.cpp:722-731 */
void CSCDec_Destroy(CSCDecHandle p);

/* csc_dec.cpp:740-777: pushes the decoded runs to Write() until the EOF marker. */
int CSCDec_Decode(CSCDecHandle p, ISeqOutStream *outstream, ICompressProgress *progress);

EXTERN_C_END

#endif
