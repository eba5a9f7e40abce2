// TEST INFRASTRUCTURE ONLY.  Link-time stand-ins for the few htslib functions that the
// reference's genotyper.cpp -> FastaReader (VCF header code, never executed by the
// oracle) pulls in.  Reaching any of them is a bug in the oracle driver: abort loudly.
#include <cstdio>
#include <cstdlib>
#include "htslib/faidx.h"
static void unreachable(const char* fn) {
  std::fprintf(stderr, "oracle/_ref: htslib stub %s reached -- not part of the hot path\n", fn);
  std::abort();
}
extern "C" {
faidx_t* fai_load(const char*) { unreachable("fai_load"); return 0; }
void fai_destroy(faidx_t*) {}
char* fai_fetch(const faidx_t*, const char*, int*) { unreachable("fai_fetch"); return 0; }
char* faidx_fetch_seq(const faidx_t*, const char*, int, int, int*) { unreachable("faidx_fetch_seq"); return 0; }
int faidx_has_seq(const faidx_t*, const char*) { unreachable("faidx_has_seq"); return 0; }
int faidx_seq_len(const faidx_t*, const char*) { unreachable("faidx_seq_len"); return 0; }
int faidx_nseq(const faidx_t*) { unreachable("faidx_nseq"); return 0; }
const char* faidx_iseq(const faidx_t*, int) { unreachable("faidx_iseq"); return 0; }
}
