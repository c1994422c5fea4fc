// oracle/ref_mftrace.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Class-level oracle for the match finder: the per-position candidate lists of the UNMODIFIED
// reference MatchFinder along the reference's own parse.  Nothing of the reference is edited or
// copied: oracle/Makefile links the trace library with
//     -Wl,--wrap=_ZN11MatchFinder9FindMatchEPjjj
//     -Wl,--wrap=_ZN11MatchFinder18FindMatchWithPriceEP5ModeljP6MFUnitPjjj
// so the two calls csc_lz.o makes into csc_mf.o (csc_lz.cpp:161,173,266-268) land here first.  Both
// entry points run the private find_match (csc_mf.cpp:243-495) into mfcand_[1..n] (csc_mf.cpp:501,592);
// the wrapper clears that array, forwards the call and reads the list back (this file alone is built
// with -Dclass=struct -Dprivate=public to see mfcand_ / vld_rge_; layout and mangling are unchanged).
//
// Trace records (same triples as ref_trace.cpp):  (4, linear position, n)  then n x (5, len, dist code).
#include <stdint.h>
#include <stddef.h>
#include <csc_mf.h>

void ref_trace_put(uint32_t k, uint32_t a, uint32_t b);   // ref_trace.cpp
bool ref_trace_cands_on();

extern "C" {
MFUnit __real__ZN11MatchFinder9FindMatchEPjjj(MatchFinder *mf, uint32_t *rep, uint32_t wpos, uint32_t limit);
void __real__ZN11MatchFinder18FindMatchWithPriceEP5ModeljP6MFUnitPjjj(MatchFinder *mf, Model *model, uint32_t state, MFUnit *ret,
                                                                      uint32_t *rep, uint32_t wpos, uint32_t limit);
}

static inline void before(MatchFinder *mf) {
    for (uint32_t i = 1; i < 32; i++) { mf->mfcand_[i].len = 0; mf->mfcand_[i].dist = 0; }
}
static inline void after(MatchFinder *mf, uint32_t x) {
    uint32_t n = 0;
    while (n + 1 < 32 && mf->mfcand_[n + 1].len) n++;
    ref_trace_put(4, x, n);
    for (uint32_t i = 1; i <= n; i++) ref_trace_put(5, mf->mfcand_[i].len, mf->mfcand_[i].dist);
}

extern "C" {
MFUnit __wrap__ZN11MatchFinder9FindMatchEPjjj(MatchFinder *mf, uint32_t *rep, uint32_t wpos, uint32_t limit) {
    if (!ref_trace_cands_on()) return __real__ZN11MatchFinder9FindMatchEPjjj(mf, rep, wpos, limit);
    const uint32_t x = mf->pos_ - mf->vld_rge_;   // pos_ starts at vld_rge_ and advances once per byte (csc_mf.cpp:57,493)
    before(mf);
    MFUnit r = __real__ZN11MatchFinder9FindMatchEPjjj(mf, rep, wpos, limit);
    after(mf, x);
    return r;
}
void __wrap__ZN11MatchFinder18FindMatchWithPriceEP5ModeljP6MFUnitPjjj(MatchFinder *mf, Model *model, uint32_t state, MFUnit *ret,
                                                                      uint32_t *rep, uint32_t wpos, uint32_t limit) {
    if (!ref_trace_cands_on()) { __real__ZN11MatchFinder18FindMatchWithPriceEP5ModeljP6MFUnitPjjj(mf, model, state, ret, rep, wpos, limit); return; }
    const uint32_t x = mf->pos_ - mf->vld_rge_;
    before(mf);
    __real__ZN11MatchFinder18FindMatchWithPriceEP5ModeljP6MFUnitPjjj(mf, model, state, ret, rep, wpos, limit);
    after(mf, x);
}
}
