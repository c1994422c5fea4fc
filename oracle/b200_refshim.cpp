// oracle/b200_refshim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The batched entry points of include/csc_b200.h that tools/csarc_gpu.cpp uses, implemented over the UNMODIFIED
// reference library (one CSCEnc / CSCDec handle per chunk, archiver/csa_adler32.cpp for the checksums).  Linking the
// archiver tool against this instead of the product lets the CPU-only test suite check its archive format logic
// against the stock csarc; the product never links or calls any of this.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <csc_enc.h>
#include <csc_dec.h>

extern uint32_t adler32(uint32_t adler, const uint8_t *buf, uint64_t len);   // archiver/csa_adler32.cpp

namespace {
struct In { ISeqInStream is; const uint8_t *p; size_t n, at; };
struct Out { ISeqOutStream os; uint8_t *p; size_t cap, at; bool over; };
int rd(void *o, void *b, size_t *sz) { In *m = (In *)o; size_t k = std::min(*sz, m->n - m->at); memcpy(b, m->p + m->at, k); m->at += k; *sz = k; return 0; }
// zero-filling allocator: the one byte per flush segment that Coder::Flush skips instead of storing is then the
// deterministic 0 the product emulates, not whatever malloc left there
void *zalloc(void *, size_t size) { return calloc(1, size ? size : 1); }
void zfree(void *, void *p) { free(p); }
ISzAlloc g_zero = {zalloc, zfree};
size_t wr(void *o, const void *b, size_t sz) { Out *m = (Out *)o; if (m->at + sz > m->cap) { m->over = true; return 0; } memcpy(m->p + m->at, b, sz); m->at += sz; return sz; }
}

extern "C" {
int CSCB200_Available(void) { return 1; }
size_t CSCB200_EncodeBound(size_t n) { return n + n / 4 + (n / 65536 + 8) * 8 + 256; }
int CSCB200_EncodeChunksAdler(int n, const uint8_t *const *in, const size_t *in_size, uint32_t dict_size, int level, int flt_mask,
                              int with_props, uint8_t *const *out, const size_t *out_cap, size_t *out_size, uint32_t *adler_in) {
    for (int i = 0; i < n; i++) {
        CSCProps p;
        CSCEncProps_Init(&p, (uint32_t)std::min<size_t>(dict_size, in_size[i] ? in_size[i] : dict_size), level);
        p.DLTFilter = flt_mask & 1; p.TXTFilter = (flt_mask >> 1) & 1; p.EXEFilter = (flt_mask >> 2) & 1;
        In mi; mi.is.Read = rd; mi.p = in[i]; mi.n = in_size[i]; mi.at = 0;
        Out mo; mo.os.Write = wr; mo.p = out[i]; mo.cap = out_cap[i]; mo.at = 0; mo.over = false;
        if (with_props) { if (out_cap[i] < CSC_PROP_SIZE) return -90; CSCEnc_WriteProperties(&p, out[i], 0); mo.at = CSC_PROP_SIZE; }
        CSCEncHandle h = CSCEnc_Create(&p, (ISeqOutStream *)&mo, &g_zero);
        if (!h) return -1;
        int r = CSCEnc_Encode(h, (ISeqInStream *)&mi, NULL);
        if (r >= 0) r = CSCEnc_Encode_Flush(h);
        CSCEnc_Destroy(h);
        if (r < 0 || mo.over) return r < 0 ? r : -90;
        out_size[i] = mo.at;
        if (adler_in) adler_in[i] = adler32(1, in[i], in_size[i]);
    }
    return 0;
}
int CSCB200_DecodeChunksAdler(int n, const uint8_t *const *in, const size_t *in_size, uint8_t *const *out, const size_t *out_cap,
                              size_t *out_size, uint32_t *adler_out) {
    for (int i = 0; i < n; i++) {
        if (in_size[i] < CSC_PROP_SIZE) return -1;
        CSCProps p;
        CSCDec_ReadProperties(&p, (uint8_t *)in[i]);
        In mi; mi.is.Read = rd; mi.p = in[i]; mi.n = in_size[i]; mi.at = CSC_PROP_SIZE;
        Out mo; mo.os.Write = wr; mo.p = out[i]; mo.cap = out_cap[i]; mo.at = 0; mo.over = false;
        CSCDecHandle h = CSCDec_Create(&p, (ISeqInStream *)&mi, &g_zero);
        if (!h) return -1;
        int r = CSCDec_Decode(h, (ISeqOutStream *)&mo, NULL);
        CSCDec_Destroy(h);
        if (r < 0 || mo.over) return r < 0 ? r : -90;
        out_size[i] = mo.at;
        if (adler_out) adler_out[i] = adler32(1, out[i], mo.at);
    }
    return 0;
}
int CSCB200_Adler32(int n, const uint8_t *const *buf, const size_t *size, const uint32_t *init, uint32_t *out) {
    for (int i = 0; i < n; i++) {
        // (the reference function answers 1 for a NULL buffer and takes a shortcut for one byte; both agree with the definition)
        uint32_t a = init ? init[i] : 1u;
        out[i] = size[i] ? adler32(a, buf[i], size[i]) : a;
    }
    return 0;
}
}
