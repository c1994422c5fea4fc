// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin C-ABI harness around the UNMODIFIED reference libcsc, compiled from the
// sources where they lie under /root/reference/src/libcsc by oracle/Makefile
// into oracle/_ref/libcsc_ref.so.  Nothing here is copied from the reference:
// it only includes the reference headers at build time and drives the public
// API (csc_enc.h:11-30, csc_dec.h:8-21) plus a few class-level entry points
// (Analyzer, Filters) that are ordinary external symbols of libcsc.
//
// Hygiene: the reference is always given a
// zero-filling ISzAlloc so that the one stale byte per flush segment
// (csc_coder.cpp:46-47) is deterministic, and the Read() granularity is an
// explicit argument because each Read() becomes one flush segment
// (csc_enc.cpp:169-176).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <vector>

#include <csc_enc.h>
#include <csc_dec.h>
#include <csc_analyzer.h>
#include <csc_filters.h>
#include <csc_typedef.h>

namespace {

void *zalloc(void *, size_t size) { return calloc(1, size ? size : 1); }
void zfree(void *, void *addr) { free(addr); }
ISzAlloc g_zero_alloc = {zalloc, zfree};

struct MemIn {
    ISeqInStream is;
    const uint8_t *p;
    size_t n, off;
    size_t read_size;  // max bytes handed out per Read (0 = whatever is asked)
};

SRes mem_read(void *h, void *buf, size_t *size) {
    MemIn *m = (MemIn *)h;
    size_t want = *size;
    if (m->read_size && want > m->read_size) want = m->read_size;
    if (want > m->n - m->off) want = m->n - m->off;
    memcpy(buf, m->p + m->off, want);
    m->off += want;
    *size = want;
    return 0;
}

struct MemOut {
    ISeqOutStream os;
    uint8_t *p;
    size_t cap, off;
    int overflow;
};

size_t mem_write(void *h, const void *buf, size_t size) {
    MemOut *m = (MemOut *)h;
    if (m->off + size > m->cap) {
        m->overflow = 1;
        return 0;
    }
    memcpy(m->p + m->off, buf, size);
    m->off += size;
    return size;
}

}  // namespace

extern "C" {

// Fill the 13 CSCProps fields (csc_common.h:19-63) into a flat u64 array so
// the tests can compare them with the product's CSCEncProps_Init.
void ref_props_init(uint32_t dict_size, int level, uint64_t *out15) {
    CSCProps p;
    memset(&p, 0, sizeof(p));
    CSCEncProps_Init(&p, dict_size, level);
    out15[0] = p.dict_size;
    out15[1] = p.csc_blocksize;
    out15[2] = p.raw_blocksize;
    out15[3] = p.hash_bits;
    out15[4] = p.hash_width;
    out15[5] = p.bt_hash_bits;
    out15[6] = p.bt_size;
    out15[7] = (level >= 5) ? p.bt_cyc : 0;  // uninitialised below -m5
    out15[8] = p.good_len;
    out15[9] = p.lz_mode;
    out15[10] = p.DLTFilter;
    out15[11] = p.TXTFilter;
    out15[12] = p.EXEFilter;
    out15[13] = CSCEnc_EstMemUsage(&p);
    out15[14] = sizeof(CSCProps);
}

// One whole stream through the reference encoder, exactly the call sequence
// of csa_worker.cpp:35-50.  flt_mask: bit0 DLT, bit1 TXT, bit2 EXE (7=default).
// with_props != 0 prepends the caller-written 10-byte header.
// read_size = bytes per Read() (0 -> raw_blocksize).  Returns 0 or negative.
int ref_encode(const uint8_t *in, size_t n, uint32_t dict_size, int level, uint32_t read_size,
               int flt_mask, int with_props, uint8_t *out, size_t cap, size_t *out_n) {
    CSCProps p;
    CSCEncProps_Init(&p, dict_size, level);
    p.DLTFilter = (flt_mask & 1) ? 1 : 0;
    p.TXTFilter = (flt_mask & 2) ? 1 : 0;
    p.EXEFilter = (flt_mask & 4) ? 1 : 0;
    MemIn mi;
    mi.is.Read = mem_read;
    mi.p = in; mi.n = n; mi.off = 0; mi.read_size = read_size;
    MemOut mo;
    mo.os.Write = mem_write;
    mo.p = out; mo.cap = cap; mo.off = 0; mo.overflow = 0;
    if (with_props) {
        if (cap < CSC_PROP_SIZE) return -1;
        CSCEnc_WriteProperties(&p, out, 0);
        mo.off = CSC_PROP_SIZE;
    }
    CSCEncHandle h = CSCEnc_Create(&p, (ISeqOutStream *)&mo, &g_zero_alloc);
    if (!h) return -2;
    int ret = CSCEnc_Encode(h, (ISeqInStream *)&mi, NULL);
    int ret2 = CSCEnc_Encode_Flush(h);
    CSCEnc_Destroy(h);
    *out_n = mo.off;
    if (mo.overflow) return -3;
    if (ret < 0) return ret;
    return ret2;
}

// Same, but the input arrives through SEVERAL CSCEnc_Encode calls on one handle (csc_enc.cpp:160-191 keeps
// model and window across calls): part k holds part_sizes[k] bytes.
int ref_encode_parts(const uint8_t *in, const size_t *part_sizes, int nparts, uint32_t dict_size, int level,
                     int with_props, uint8_t *out, size_t cap, size_t *out_n) {
    CSCProps p;
    CSCEncProps_Init(&p, dict_size, level);
    MemOut mo;
    mo.os.Write = mem_write;
    mo.p = out; mo.cap = cap; mo.off = 0; mo.overflow = 0;
    if (with_props) {
        if (cap < CSC_PROP_SIZE) return -1;
        CSCEnc_WriteProperties(&p, out, 0);
        mo.off = CSC_PROP_SIZE;
    }
    CSCEncHandle h = CSCEnc_Create(&p, (ISeqOutStream *)&mo, &g_zero_alloc);
    if (!h) return -2;
    int ret = 0;
    size_t off = 0;
    for (int k = 0; k < nparts && ret >= 0; k++) {
        MemIn mi;
        mi.is.Read = mem_read;
        mi.p = in + off; mi.n = part_sizes[k]; mi.off = 0; mi.read_size = 0;
        ret = CSCEnc_Encode(h, (ISeqInStream *)&mi, NULL);
        off += part_sizes[k];
    }
    int ret2 = CSCEnc_Encode_Flush(h);
    CSCEnc_Destroy(h);
    *out_n = mo.off;
    if (mo.overflow) return -3;
    if (ret < 0) return ret;
    return ret2;
}

// Whole stream (starting with the 10-byte props header) through the reference
// decoder, the call sequence of csa_worker.cpp:75-83.
int ref_decode(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_n) {
    if (n < CSC_PROP_SIZE) return -1;
    CSCProps p;
    memset(&p, 0, sizeof(p));
    CSCDec_ReadProperties(&p, (uint8_t *)in);
    MemIn mi;
    mi.is.Read = mem_read;
    mi.p = in; mi.n = n; mi.off = CSC_PROP_SIZE; mi.read_size = 0;
    MemOut mo;
    mo.os.Write = mem_write;
    mo.p = out; mo.cap = cap; mo.off = 0; mo.overflow = 0;
    CSCDecHandle h = CSCDec_Create(&p, (ISeqInStream *)&mi, &g_zero_alloc);
    if (!h) return -2;
    int ret = CSCDec_Decode(h, (ISeqOutStream *)&mo, NULL);
    CSCDec_Destroy(h);
    *out_n = mo.off;
    if (mo.overflow) return -3;
    return ret;
}

// ---- multi-threaded chunk harness (CPU baseline: one handle per chunk, the
// ---- csarc -p/-t model of csarc.cpp:532-543 + csa_worker.cpp:23-56) --------
struct MtJob {
    const uint8_t *in; size_t n;
    uint8_t *out; size_t cap; size_t out_n;
    int ret;
};
struct MtCtx {
    MtJob *jobs; int njobs; int next;
    uint32_t dict_size; int level; uint32_t read_size; int decode;
    pthread_mutex_t mu;
};
static void *mt_worker(void *arg) {
    MtCtx *c = (MtCtx *)arg;
    for (;;) {
        pthread_mutex_lock(&c->mu);
        int j = c->next++;
        pthread_mutex_unlock(&c->mu);
        if (j >= c->njobs) break;
        MtJob *job = &c->jobs[j];
        if (c->decode) {
            job->ret = ref_decode(job->in, job->n, job->out, job->cap, &job->out_n);
        } else {
            uint32_t d = c->dict_size;
            if ((uint64_t)d > job->n) d = (uint32_t)job->n;  // csa_worker.cpp:35
            job->ret = ref_encode(job->in, job->n, d, c->level, c->read_size, 7, 1, job->out,
                                  job->cap, &job->out_n);
        }
    }
    return NULL;
}
// in_ptrs/out_ptrs: arrays of njobs pointers; in_n/out_cap/out_n arrays of sizes.
int ref_run_chunks(int decode, int njobs, const uint8_t **in_ptrs, const size_t *in_n,
                   uint8_t **out_ptrs, const size_t *out_cap, size_t *out_n, uint32_t dict_size,
                   int level, uint32_t read_size, int nthreads) {
    std::vector<MtJob> jobs(njobs);
    for (int i = 0; i < njobs; i++) {
        jobs[i].in = in_ptrs[i]; jobs[i].n = in_n[i];
        jobs[i].out = out_ptrs[i]; jobs[i].cap = out_cap[i];
        jobs[i].out_n = 0; jobs[i].ret = 0;
    }
    MtCtx c;
    c.jobs = jobs.data(); c.njobs = njobs; c.next = 0;
    c.dict_size = dict_size; c.level = level; c.read_size = read_size; c.decode = decode;
    pthread_mutex_init(&c.mu, NULL);
    if (nthreads < 1) nthreads = 1;
    std::vector<pthread_t> th(nthreads);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, mt_worker, &c);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&c.mu);
    int ret = 0;
    for (int i = 0; i < njobs; i++) {
        out_n[i] = jobs[i].out_n;
        if (jobs[i].ret < 0) ret = jobs[i].ret;
    }
    return ret;
}

// ---- class-level oracles --------------------------------------------------
// Analyzer::Analyze (csc_analyzer.cpp:184-239) on consecutive 8 KiB blocks of
// one buffer; bpb_out[i] is left at 0xFFFFFFFF when Analyze does not write it.
void ref_analyze_blocks(const uint8_t *buf, size_t n, uint32_t *type_out, uint32_t *bpb_out) {
    Analyzer a;
    a.Init();
    size_t nb = 0;
    for (size_t off = 0; off < n; off += MinBlockSize, nb++) {
        uint32_t sz = (uint32_t)((n - off) < MinBlockSize ? (n - off) : MinBlockSize);
        uint32_t bpb = 0xFFFFFFFFu;
        type_out[nb] = a.Analyze((uint8_t *)buf + off, sz, &bpb);
        bpb_out[nb] = bpb;
    }
}

uint32_t ref_dlt_bpb(const uint8_t *buf, uint32_t n, uint32_t chn) {
    Analyzer a;
    a.Init();
    return a.GetDltBpb((uint8_t *)buf, n, chn);
}

// Filters (csc_filters.cpp) in place on buf. which: 0 e8e9, 1 dict, 2 delta.
// Returns Foward_Dict's accept flag for which==1 forward, else 1.
int ref_filter(int which, int inverse, uint8_t *buf, uint32_t n, uint32_t chn) {
    Filters f;
    f.Init(&g_zero_alloc);
    int r = 1;
    if (which == 0) {
        if (inverse) f.Inverse_E89(buf, n); else f.Forward_E89(buf, n);
    } else if (which == 1) {
        if (inverse) f.Inverse_Dict(buf, n); else r = (int)f.Foward_Dict(buf, n);
    } else {
        if (inverse) f.Inverse_Delta(buf, n, chn); else f.Forward_Delta(buf, n, chn);
    }
    f.Destroy();
    return r;
}

}  // extern "C"
