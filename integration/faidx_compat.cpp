// integration/faidx_compat.cpp -- the htslib faidx entry points LongTR's FastaReader binds (src/fasta_reader.h:59-102,
// src/fasta_reader.cpp:25-70: fai_load, fai_destroy, fai_fetch, faidx_fetch_seq, faidx_has_seq, faidx_seq_len, faidx_nseq,
// faidx_iseq) implemented on the library's own indexed-FASTA reader (ltr_fasta_*, csrc/host/fasta_reader.cpp).  Linked in place
// of htslib, the reference's FastaReader runs unmodified on it (oracle/build_ref.sh -> oracle/_ref/libltr_ref_fasta.so,
// tests/test_fasta_bed.py).  Like htslib, fai_fetch / faidx_fetch_seq return malloc'ed, NUL-terminated copies the caller frees.
#include <stdlib.h>
#include <string.h>

#include "longtr_b200.h"

extern "C" {

typedef struct faidx_t faidx_t;  // opaque to LongTR; here: an ltr_fasta

faidx_t* fai_load(const char* fn) {
  ltr_fasta* fa = nullptr;
  if (!fn || ltr_fasta_open(fn, &fa) != LTR_OK) return nullptr;
  return reinterpret_cast<faidx_t*>(fa);
}

void fai_destroy(faidx_t* fai) { ltr_fasta_close(reinterpret_cast<ltr_fasta*>(fai)); }

int faidx_nseq(const faidx_t* fai) { return ltr_fasta_n_seqs(reinterpret_cast<const ltr_fasta*>(fai)); }

const char* faidx_iseq(const faidx_t* fai, int i) { return ltr_fasta_seq_name(reinterpret_cast<const ltr_fasta*>(fai), i); }

int faidx_seq_len(const faidx_t* fai, const char* seq) {
  return (int)ltr_fasta_seq_len(reinterpret_cast<const ltr_fasta*>(fai), seq);
}

int faidx_has_seq(const faidx_t* fai, const char* seq) { return faidx_seq_len(fai, seq) >= 0 ? 1 : 0; }

// [p_beg_i, p_end_i], both inclusive and 0-based, clipped to the sequence like htslib does
char* faidx_fetch_seq(const faidx_t* fai, const char* c_name, int p_beg_i, int p_end_i, int* len) {
  const ltr_fasta* fa = reinterpret_cast<const ltr_fasta*>(fai);
  const long long n = ltr_fasta_seq_len(fa, c_name);
  if (n < 0) {
    if (len) *len = -2;
    return nullptr;
  }
  long long b = p_beg_i < 0 ? 0 : p_beg_i, e = (long long)p_end_i + 1;
  if (e > n) e = n;
  if (b > e) b = e;
  char* out = (char*)malloc((size_t)(e - b) + 1);
  if (!out) return nullptr;
  if (ltr_fasta_fetch(fa, c_name, b, e, (uint8_t*)out) != LTR_OK) {
    free(out);
    if (len) *len = -1;
    return nullptr;
  }
  out[e - b] = 0;
  if (len) *len = (int)(e - b);
  return out;
}

// LongTR only passes bare sequence names as the region (src/fasta_reader.h:73): the whole sequence
char* fai_fetch(const faidx_t* fai, const char* reg, int* len) {
  const long long n = ltr_fasta_seq_len(reinterpret_cast<const ltr_fasta*>(fai), reg);
  if (n < 0) {
    if (len) *len = -2;
    return nullptr;
  }
  return faidx_fetch_seq(fai, reg, 0, (int)(n - 1), len);
}

}  // extern "C"
