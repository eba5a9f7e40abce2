// hts_compat.cpp -- the htslib entry points LongTR's alignment input binds (reference src/bam_io.h:62-210, 367-452;
// src/bam_io.cpp:65-189: sam_open, sam_hdr_read, sam_index_load, sam_itr_querys, sam_itr_next, bam_endpos, bam_aux_*, ...)
// implemented on the library's own BGZF / BAM / BAI reader (ltr_bam_*, include/longtr_b200.h): linked in place of htslib,
// the reference's BamCramReader / BamAlignment run unmodified on top of it.
// BAM only (CRAM is answered with "cannot open").  A region iterator fetches its records when it is created; the
// offsets htslib exposes for the reference's iterator-reuse shortcut (bam_io.cpp:151-166) are reported as "no chunk",
// which makes the reference take its plain path.
#include <stdlib.h>
#include <string.h>

#include <string>

#include "htslib/sam.h"
#include "longtr_b200.h"

namespace {
struct CompatFile {
  ltr_bam* bam;
};
struct CompatIter {
  ltr_bam_reads* reads;
  uint32_t next;
};
int aux_type_size(char t) {
  switch (t) {
    case 'A': case 'c': case 'C': return 1;
    case 's': case 'S': return 2;
    case 'i': case 'I': case 'f': return 4;
    case 'd': return 8;
    default: return 0;
  }
}
// first byte behind the aux field whose type byte is at s (NULL: malformed)
const uint8_t* aux_skip(const uint8_t* s, const uint8_t* e) {
  if (s >= e) return nullptr;
  const char t = (char)*s++;
  if (t == 'Z' || t == 'H') {
    while (s < e && *s) ++s;
    return s < e ? s + 1 : nullptr;
  }
  if (t == 'B') {
    if (s + 5 > e) return nullptr;
    const int w = aux_type_size((char)s[0]);
    uint32_t n;
    memcpy(&n, s + 1, 4);
    s += 5 + (size_t)n * (size_t)w;
    return (w && s <= e) ? s : nullptr;
  }
  const int w = aux_type_size(t);
  return (w && s + w <= e) ? s + w : nullptr;
}
uint8_t* aux_begin(const bam1_t* b) { return bam_get_qual(b) + b->core.l_qseq; }
}  // namespace

extern "C" {

bam1_t* bam_init1(void) { return (bam1_t*)calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t* b) {
  if (!b) return;
  free(b->data);
  free(b);
}
bam1_t* bam_copy1(bam1_t* dst, const bam1_t* src) {
  if (dst->m_data < (uint32_t)src->l_data) {
    dst->data = (uint8_t*)realloc(dst->data, (size_t)src->l_data);
    dst->m_data = (uint32_t)src->l_data;
  }
  if (src->l_data) memcpy(dst->data, src->data, (size_t)src->l_data);
  dst->core = src->core;
  dst->l_data = src->l_data;
  dst->id = src->id;
  return dst;
}
int32_t bam_endpos(const bam1_t* b) {
  const uint32_t* cigar = bam_get_cigar(b);
  int32_t len = 0;
  for (uint32_t k = 0; k < b->core.n_cigar; ++k) {
    const uint32_t op = bam_cigar_op(cigar[k]);
    if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) len += (int32_t)bam_cigar_oplen(cigar[k]);
  }
  return b->core.pos + (len ? len : 1);
}
uint8_t* bam_aux_get(const bam1_t* b, const char tag[2]) {
  const uint8_t* s = aux_begin(b);
  const uint8_t* e = b->data + b->l_data;
  while (s && s + 3 <= e) {
    if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return const_cast<uint8_t*>(s + 2);
    s = aux_skip(s + 2, e);
  }
  return NULL;
}
int bam_aux_del(bam1_t* b, uint8_t* s) {
  uint8_t* e = b->data + b->l_data;
  const uint8_t* after = aux_skip(s, e);
  if (!after) return -1;
  memmove(s - 2, after, (size_t)(e - after));
  b->l_data -= (int)(after - (s - 2));
  return 0;
}
int bam_aux_append(bam1_t* b, const char tag[2], char type, int len, const uint8_t* data) {
  const size_t need = (size_t)b->l_data + 3 + (size_t)len;
  if (b->m_data < need) {
    b->data = (uint8_t*)realloc(b->data, need);
    b->m_data = (uint32_t)need;
  }
  uint8_t* p = b->data + b->l_data;
  p[0] = (uint8_t)tag[0];
  p[1] = (uint8_t)tag[1];
  p[2] = (uint8_t)type;
  memcpy(p + 3, data, (size_t)len);
  b->l_data = (int)need;
  return 0;
}
char bam_aux2A(const uint8_t* s) { return (*s == 'A') ? (char)s[1] : 0; }
int64_t bam_aux2i(const uint8_t* s) {
  switch ((char)*s) {
    case 'c': return (int8_t)s[1];
    case 'C': return s[1];
    case 's': { int16_t v; memcpy(&v, s + 1, 2); return v; }
    case 'S': { uint16_t v; memcpy(&v, s + 1, 2); return v; }
    case 'i': { int32_t v; memcpy(&v, s + 1, 4); return v; }
    case 'I': { uint32_t v; memcpy(&v, s + 1, 4); return v; }
    default: return 0;
  }
}
double bam_aux2f(const uint8_t* s) {
  if (*s == 'f') { float v; memcpy(&v, s + 1, 4); return v; }
  if (*s == 'd') { double v; memcpy(&v, s + 1, 8); return v; }
  return (double)bam_aux2i(s);
}
char* bam_aux2Z(const uint8_t* s) { return (*s == 'Z' || *s == 'H') ? (char*)(s + 1) : NULL; }

samFile* sam_open(const char* fn, const char* mode) {
  if (!fn || !mode || mode[0] != 'r') return NULL;
  ltr_bam* bam = NULL;
  if (ltr_bam_open(fn, NULL, &bam) != LTR_OK) return NULL;
  samFile* fp = (samFile*)calloc(1, sizeof(samFile));
  fp->is_bin = 1;
  fp->is_bgzf = 1;
  CompatFile* cf = new CompatFile();
  cf->bam = bam;
  fp->fp = cf;
  return fp;
}
int sam_close(samFile* fp) {
  if (!fp) return 0;
  CompatFile* cf = static_cast<CompatFile*>(fp->fp);
  if (cf) {
    ltr_bam_close(cf->bam);
    delete cf;
  }
  free(fp);
  return 0;
}
sam_hdr_t* sam_hdr_read(samFile* fp) {
  if (!fp || !fp->fp) return NULL;
  const ltr_bam* bam = static_cast<CompatFile*>(fp->fp)->bam;
  sam_hdr_t* h = (sam_hdr_t*)calloc(1, sizeof(sam_hdr_t));
  h->n_targets = ltr_bam_n_refs(bam);
  h->target_len = (uint32_t*)calloc((size_t)h->n_targets + 1, sizeof(uint32_t));
  h->target_name = (char**)calloc((size_t)h->n_targets + 1, sizeof(char*));
  for (int32_t t = 0; t < h->n_targets; ++t) {
    h->target_len[t] = (uint32_t)ltr_bam_ref_len(bam, t);
    h->target_name[t] = strdup(ltr_bam_ref_name(bam, t));
  }
  const char* text = ltr_bam_header_text(bam);
  h->l_text = strlen(text);
  h->text = strdup(text);
  h->sdict = fp->fp;  // the file the header came from: sam_itr_querys only receives the header and the index
  return h;
}
void sam_hdr_destroy(sam_hdr_t* h) {
  if (!h) return;
  for (int32_t t = 0; t < h->n_targets; ++t) free(h->target_name[t]);
  free(h->target_name);
  free(h->target_len);
  free(h->text);
  free(h);
}
void bam_hdr_destroy(sam_hdr_t* h) { sam_hdr_destroy(h); }
// The index lives inside ltr_bam (loaded by ltr_bam_open from "<file>.bai"): the handle only says whether it is there.
hts_idx_t* sam_index_load(samFile* fp, const char*) {
  if (!fp || !fp->fp) return NULL;
  CompatFile* cf = static_cast<CompatFile*>(fp->fp);
  return ltr_bam_has_index(cf->bam) ? reinterpret_cast<hts_idx_t*>(cf) : NULL;
}
void hts_idx_destroy(hts_idx_t*) {}
// region: "chr" or "chr:beg-end" (1-based, inclusive), as the reference builds it (bam_io.cpp:131, 155-158)
hts_itr_t* sam_itr_querys(const hts_idx_t* idx, sam_hdr_t* hdr, const char* region) {
  if (!idx || !hdr || !region) return NULL;
  const ltr_bam* bam = reinterpret_cast<const CompatFile*>(idx)->bam;
  std::string chrom(region);
  int64_t beg = 0, end = 1ll << 29;
  const size_t colon = chrom.rfind(':');
  if (colon != std::string::npos && ltr_bam_ref_id(bam, chrom.c_str()) < 0) {
    const std::string range = chrom.substr(colon + 1);
    chrom.resize(colon);
    const size_t dash = range.find('-');
    beg = atoll(range.substr(0, dash).c_str()) - 1;
    if (dash != std::string::npos) end = atoll(range.substr(dash + 1).c_str());
    if (beg < 0) beg = 0;
  }
  const int32_t tid = ltr_bam_ref_id(bam, chrom.c_str());
  if (tid < 0) return NULL;
  ltr_bam_reads* reads = NULL;
  if (ltr_bam_fetch(bam, tid, beg, end, 1, &reads) != LTR_OK) return NULL;
  hts_itr_t* it = (hts_itr_t*)calloc(1, sizeof(hts_itr_t));
  it->tid = tid;
  it->beg = beg;
  it->end = end;
  it->n_off = 0;
  it->curr_off = 1;  // "a record has been read": the reference stores it as its restart hint and never uses it with n_off == 0
  CompatIter* ci = new CompatIter();
  ci->reads = reads;
  ci->next = 0;
  it->reg_list = ci;
  return it;
}
int sam_itr_next(samFile*, hts_itr_t* itr, bam1_t* r) {
  if (!itr || !itr->reg_list) return -1;
  CompatIter* ci = static_cast<CompatIter*>(itr->reg_list);
  if (ci->next >= ci->reads->n) return -1;
  const uint32_t i = ci->next++;
  const uint8_t* raw = ci->reads->raw + ci->reads->raw_off[i];
  const size_t n = ci->reads->raw_off[i + 1] - ci->reads->raw_off[i];
  int32_t w[8];
  memcpy(w, raw, 32);
  r->core.tid = w[0];
  r->core.pos = w[1];
  r->core.l_qname = (uint8_t)(w[2] & 0xff);
  r->core.qual = (uint8_t)((w[2] >> 8) & 0xff);
  r->core.bin = (uint16_t)((uint32_t)w[2] >> 16);
  r->core.n_cigar = (uint32_t)w[3] & 0xffff;
  r->core.flag = (uint16_t)((uint32_t)w[3] >> 16);
  r->core.l_qseq = w[4];
  r->core.mtid = w[5];
  r->core.mpos = w[6];
  r->core.isize = w[7];
  const size_t l_data = n - 32;
  if (r->m_data < l_data) {
    r->data = (uint8_t*)realloc(r->data, l_data);
    r->m_data = (uint32_t)l_data;
  }
  memcpy(r->data, raw + 32, l_data);
  r->l_data = (int)l_data;
  return (int)n;
}
void hts_itr_destroy(hts_itr_t* it) {
  if (!it) return;
  CompatIter* ci = static_cast<CompatIter*>(it->reg_list);
  if (ci) {
    ltr_bam_reads_free(ci->reads);
    delete ci;
  }
  free(it);
}
int hts_set_fai_filename(htsFile*, const char*) { return -1; }  // CRAM: not supported

}  // extern "C"
