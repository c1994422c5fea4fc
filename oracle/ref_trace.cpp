// oracle/ref_trace.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Symbol tracer for the UNMODIFIED reference encoder.  The reference calls the
// hooks PEncodeLiteral / PEncodeMatch / PEncodeRepMatch / PEncode1BMatch at the
// top of every Model::Encode* (csc_model.cpp:171,200,220,304) when it is
// compiled with -D_HAVE_PROFILER_ (csc_profiler.h:6-18).  oracle/Makefile
// builds a second copy of the reference objects with that define, leaves the
// reference's own csc_profiler.cpp out, and links these definitions instead,
// so a parity failure can be bisected to the first differing packet.
#include <stdint.h>
#include <stddef.h>

static uint32_t *g_trace = 0;   // triples (kind, a, b)
static size_t g_cap = 0, g_n = 0;

static bool g_cands = false;   // also record the match-finder candidate lists (ref_mftrace.cpp)

static inline void put(uint32_t k, uint32_t a, uint32_t b) {
    if (g_trace && g_n < g_cap) {
        g_trace[g_n * 3 + 0] = k;
        g_trace[g_n * 3 + 1] = a;
        g_trace[g_n * 3 + 2] = b;
    }
    g_n++;
}

// kinds follow the packet kinds of csc_model.h:38-42 (0 literal, 1 match,
// 2 rep0len1, 3 rep match)
void PEncodeLiteral(uint32_t c) { put(0, c, 0); }
void PEncodeMatch(uint32_t len, uint32_t dist) { put(1, len, dist); }
void PEncode1BMatch() { put(2, 0, 0); }
void PEncodeRepMatch(uint32_t len, uint32_t idx) { put(3, len, idx); }
void PWriteLog() {}

void ref_trace_put(uint32_t k, uint32_t a, uint32_t b) { put(k, a, b); }
bool ref_trace_cands_on() { return g_trace && g_cands; }

extern "C" {
void ref_trace_begin(uint32_t *buf, size_t cap_triples) { g_trace = buf; g_cap = cap_triples; g_n = 0; g_cands = false; }
// same, plus (4, position, n) / (5, len, dist) records for every MatchFinder::FindMatch / FindMatchWithPrice call
void ref_trace_begin_cands(uint32_t *buf, size_t cap_triples) { g_trace = buf; g_cap = cap_triples; g_n = 0; g_cands = true; }
size_t ref_trace_end() { g_trace = 0; g_cands = false; return g_n; }
}
