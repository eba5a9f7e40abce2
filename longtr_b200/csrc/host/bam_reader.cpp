// bam_reader.cpp -- BGZF / BAM / BAI reader behind the C ABI (ltr_bam_*): SURVEY.md section 8f, N3.
//
// LongTR reads its alignments through htslib (reference src/bam_io.h:313-420 BamCramReader, src/bam_io.cpp:94-214:
// sam_open / sam_hdr_read / sam_index_load / sam_itr_querys / sam_itr_next); htslib is neither vendored nor installed
// here, so this is a from-scratch reader of the three formats with zlib as the only dependency:
//   BGZF  RFC 1952 members with the BC extra field; a virtual offset is (offset of the block in the file << 16) | offset
//         inside the inflated block (SAM specification 4.1);
//   BAM   header, reference dictionary, alignment records (SAM specification 4.2);
//   BAI   binning index + 16 kb linear index (SAM specification 5.2): a region query inflates only the blocks the index
//         names.
// The file is mapped once; queries are read-only on the mapping and keep their own block cache, so several host threads
// can fetch different regions from one ltr_bam concurrently.
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <new>
#include <atomic>
#include <string>
#include <utility>
#include <vector>

#include "longtr_b200.h"

namespace {

struct Mapped {
  const uint8_t* p = nullptr;
  size_t n = 0;
  uint64_t id = 0;  // unique per opened file (key of the per-thread block cache)
  bool open(const char* path) {
    static std::atomic<uint64_t> next_id(1);
    id = next_id.fetch_add(1);
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) {
      ::close(fd);
      return false;
    }
    n = (size_t)st.st_size;
    if (n) {
      void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) {
        ::close(fd);
        return false;
      }
      p = static_cast<const uint8_t*>(m);
    }
    ::close(fd);
    return true;
  }
  ~Mapped() {
    if (p) munmap(const_cast<uint8_t*>(p), n);
  }
};

inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t le64(const uint8_t* p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }

// One BGZF block: compressed size from the BC subfield, payload inflated with a raw deflate stream.
// Returns the compressed block size (0 = malformed / end of file).
size_t bgzf_inflate(const Mapped& f, uint64_t coff, std::vector<uint8_t>& out) {
  out.clear();
  if (coff + 18 > f.n) return 0;
  const uint8_t* h = f.p + coff;
  if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return 0;
  const uint32_t xlen = le16(h + 10);
  if (coff + 12 + xlen > f.n) return 0;
  uint32_t bsize = 0;
  for (uint32_t q = 0; q + 4 <= xlen;) {
    const uint8_t* x = h + 12 + q;
    const uint32_t slen = le16(x + 2);
    if (x[0] == 'B' && x[1] == 'C' && slen == 2) bsize = (uint32_t)le16(x + 4) + 1;
    q += 4 + slen;
  }
  if (bsize < 12 + xlen + 8 || coff + bsize > f.n) return 0;
  const uint32_t isize = le32(h + bsize - 4);
  out.resize(isize);
  if (isize) {
    // one inflate state per thread, reset per block (inflateInit2 allocates the 32 KB window every time)
    struct State {
      z_stream zs;
      bool ok;
      State() {
        memset(&zs, 0, sizeof(zs));
        ok = inflateInit2(&zs, -15) == Z_OK;
      }
      ~State() {
        if (ok) inflateEnd(&zs);
      }
    };
    static thread_local State st;
    if (!st.ok || inflateReset(&st.zs) != Z_OK) return 0;
    st.zs.next_in = const_cast<Bytef*>(h + 12 + xlen);
    st.zs.avail_in = bsize - 12 - xlen - 8;
    st.zs.next_out = out.data();
    st.zs.avail_out = isize;
    const int rc = inflate(&st.zs, Z_FINISH);
    if (rc != Z_STREAM_END || st.zs.total_out != isize) return 0;
  }
  return bsize;
}

// The last two blocks a thread inflated, whatever query they were for: consecutive regions of a sorted region list start in
// the block the previous one ended in.
struct ThreadBlocks {
  uint64_t id[2] = {0, 0}, coff[2] = {0, 0};
  size_t csize[2] = {0, 0};
  std::vector<uint8_t> blk[2];
  int victim = 0;
};
static ThreadBlocks& thread_blocks() {
  static thread_local ThreadBlocks t;
  return t;
}

// Sequential reader over virtual offsets with a one-block cache (backed by the thread's last two blocks).
struct Cursor {
  const Mapped* f;
  std::vector<uint8_t> blk;
  uint64_t coff = ~0ull;  // block held in blk
  size_t csize = 0;
  uint64_t at_c = 0;  // position: block offset ...
  uint32_t at_u = 0;  // ... and offset inside it
  explicit Cursor(const Mapped* file) : f(file) {}
  bool load(uint64_t c) {
    if (c == coff) return csize != 0;
    coff = c;
    ThreadBlocks& T = thread_blocks();
    for (int k = 0; k < 2; ++k)
      if (T.id[k] == f->id && T.coff[k] == c && T.csize[k] != 0) {
        blk = T.blk[k];
        csize = T.csize[k];
        T.victim = 1 - k;
        return true;
      }
    csize = bgzf_inflate(*f, c, blk);
    if (csize != 0) {
      const int k = T.victim;
      T.id[k] = f->id;
      T.coff[k] = c;
      T.csize[k] = csize;
      T.blk[k] = blk;
      T.victim = 1 - k;
    }
    return csize != 0;
  }
  void seek(uint64_t voff) {
    at_c = voff >> 16;
    at_u = (uint32_t)(voff & 0xffff);
  }
  uint64_t tell() {  // virtual offset of the next byte (normalised to the start of the next block at a block's end)
    if (load(at_c) && at_u >= blk.size() && at_c + csize < f->n) return (at_c + csize) << 16;
    return (at_c << 16) | at_u;
  }
  // Copies n bytes; false at the end of the file or on a malformed block.
  bool read(uint8_t* dst, size_t n) {
    while (n) {
      if (!load(at_c)) return false;
      if (at_u >= blk.size()) {  // next block (empty blocks -- the EOF marker -- are skipped)
        at_c += csize;
        at_u = 0;
        if (at_c >= f->n) return false;
        continue;
      }
      const size_t take = std::min(n, blk.size() - at_u);
      memcpy(dst, blk.data() + at_u, take);
      dst += take;
      at_u += (uint32_t)take;
      n -= take;
    }
    return true;
  }
};

struct Chunk {
  uint64_t beg, end;
};
struct RefIndex {
  std::vector<std::pair<uint32_t, std::vector<Chunk>>> bins;  // sorted by bin number
  std::vector<uint64_t> linear;
  const std::vector<Chunk>* find(uint32_t bin) const {
    auto it = std::lower_bound(bins.begin(), bins.end(), bin,
                               [](const std::pair<uint32_t, std::vector<Chunk>>& a, uint32_t b) { return a.first < b; });
    return (it != bins.end() && it->first == bin) ? &it->second : nullptr;
  }
};

// bins that may hold records overlapping [beg, end) (SAM specification 5.3)
void reg2bins(int64_t beg, int64_t end, std::vector<uint32_t>& out) {
  out.clear();
  if (beg < 0) beg = 0;
  if (end <= beg) return;
  --end;
  if (end >= (1ll << 29)) end = (1ll << 29) - 1;
  out.push_back(0);
  for (int l = 1, t = 0, s = 26; l <= 5; ++l, s -= 3) {
    t += 1 << ((l - 1) * 3);
    const int64_t b = t + (beg >> s), e = t + (end >> s);
    for (int64_t i = b; i <= e; ++i) out.push_back((uint32_t)i);
  }
}

const char kSeqCodes[] = "=ACMGRSVTWYHKDBN";

}  // namespace

struct ltr_bam {
  Mapped file;
  std::string text;
  std::vector<std::string> ref_names;
  std::vector<int64_t> ref_lens;
  uint64_t first_record = 0;  // virtual offset of the first alignment
  bool has_index = false;
  std::vector<RefIndex> index;
  std::string error;
};

namespace {

struct ReadsOwner {
  ltr_bam_reads pub;
  std::vector<int32_t> tid, pos, end, hp, mate_tid, mate_pos;
  std::vector<uint16_t> flag;
  std::vector<uint8_t> mapq;
  std::vector<uint32_t> name_off, seq_off, cigar_off, cigar_ops, raw_off;
  std::vector<char> names;
  std::vector<uint8_t> seq, qual, raw;
  void publish() {
    pub.n = (uint32_t)tid.size();
    pub.tid = tid.data(); pub.pos = pos.data(); pub.end = end.data(); pub.flag = flag.data(); pub.mapq = mapq.data();
    pub.mate_tid = mate_tid.data(); pub.mate_pos = mate_pos.data();
    pub.name_off = name_off.data(); pub.names = names.data();
    pub.seq_off = seq_off.data(); pub.seq = seq.data(); pub.qual = qual.data();
    pub.cigar_off = cigar_off.data(); pub.cigar_ops = cigar_ops.data();
    pub.hp = hp.data();
    pub.raw_off = raw_off.data(); pub.raw = raw.data();
    pub.owner = this;
  }
};

// Integer value of the HP aux tag (0 when absent or not an integer type).
int32_t aux_int(const uint8_t* p, const uint8_t* e, char t0, char t1) {
  while (p + 3 <= e) {
    const char a = (char)p[0], b = (char)p[1], ty = (char)p[2];
    p += 3;
    const bool hit = (a == t0 && b == t1);
    size_t len = 0;
    switch (ty) {
      case 'A': case 'c': case 'C': len = 1; break;
      case 's': case 'S': len = 2; break;
      case 'i': case 'I': case 'f': len = 4; break;
      case 'd': len = 8; break;
      case 'Z': case 'H': {
        const uint8_t* q = p;
        while (q < e && *q) ++q;
        len = (size_t)(q - p) + 1;
        break;
      }
      case 'B': {
        if (p + 5 > e) return 0;
        const char sub = (char)p[0];
        const uint32_t cnt = le32(p + 1);
        const size_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        len = 5 + (size_t)cnt * w;
        break;
      }
      default: return 0;
    }
    if (p + len > e) return 0;
    if (hit) {
      switch (ty) {
        case 'c': return (int8_t)p[0];
        case 'C': return p[0];
        case 's': return (int16_t)le16(p);
        case 'S': return le16(p);
        case 'i': case 'I': return (int32_t)le32(p);
        default: return 0;
      }
    }
    p += len;
  }
  return 0;
}

// Appends one record (rec = the block_size bytes behind the length word).  Returns false when it is malformed.
bool append_record(ReadsOwner& R, const uint8_t* rec, size_t n, bool keep_raw) {
  if (n < 32) return false;
  const int32_t tid = (int32_t)le32(rec), pos = (int32_t)le32(rec + 4);
  const uint32_t l_name = rec[8], mapq = rec[9], n_cig = le16(rec + 12), flag = le16(rec + 14);
  const uint32_t l_seq = le32(rec + 16);
  const size_t need = 32 + (size_t)l_name + 4 * (size_t)n_cig + (l_seq + 1) / 2 + l_seq;
  if (need > n || l_name == 0) return false;
  R.tid.push_back(tid);
  R.pos.push_back(pos);
  R.flag.push_back((uint16_t)flag);
  R.mapq.push_back((uint8_t)mapq);
  R.mate_tid.push_back((int32_t)le32(rec + 20));
  R.mate_pos.push_back((int32_t)le32(rec + 24));
  const uint8_t* q = rec + 32;
  // l_name counts the terminating NUL; a damaged record may not have one, so it is written here (readers use strlen)
  R.names.insert(R.names.end(), (const char*)q, (const char*)q + l_name - 1);
  R.names.push_back('\0');
  R.name_off.push_back((uint32_t)R.names.size());
  q += l_name;
  int64_t ref_len = 0;
  for (uint32_t k = 0; k < n_cig; ++k) {
    const uint32_t v = le32(q + 4 * k);
    R.cigar_ops.push_back(v);
    const uint32_t op = v & 15;
    if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += (int64_t)(v >> 4);  // M D N = X
  }
  R.cigar_off.push_back((uint32_t)R.cigar_ops.size());
  q += 4 * (size_t)n_cig;
  // as htslib's bam_endpos: an unaligned record covers one base (positions of a damaged record saturate instead of wrapping)
  const int64_t end64 = (int64_t)pos + (ref_len ? ref_len : 1);
  R.end.push_back(end64 > INT32_MAX ? INT32_MAX : (int32_t)end64);
  const size_t s0 = R.seq.size();
  R.seq.resize(s0 + l_seq);
  {
    // two bases per packed byte through a 256-entry table of letter pairs (high nibble first)
    static const struct PairTable {
      uint16_t v[256];
      PairTable() {
        for (int b = 0; b < 256; ++b) v[b] = (uint16_t)((uint8_t)kSeqCodes[b >> 4] | ((uint16_t)(uint8_t)kSeqCodes[b & 15] << 8));
      }
    } pairs;
    uint8_t* dst = R.seq.data() + s0;
    const uint32_t full = l_seq >> 1;
    for (uint32_t k = 0; k < full; ++k) memcpy(dst + 2 * (size_t)k, &pairs.v[q[k]], 2);
    if (l_seq & 1) dst[l_seq - 1] = (uint8_t)kSeqCodes[q[full] >> 4];
  }
  q += (l_seq + 1) / 2;
  R.qual.resize(s0 + l_seq);
  {
    // phred + 33, 0xff (missing) -> '!', above 93 -> '~'
    uint8_t* dst = R.qual.data() + s0;
    uint32_t i = 0;
#if defined(__SSE2__)
    const __m128i c33 = _mm_set1_epi8(33), cap = _mm_set1_epi8(93), ff = _mm_set1_epi8((char)0xff);
    for (; i + 16 <= l_seq; i += 16) {
      const __m128i x = _mm_loadu_si128((const __m128i*)(q + i));
      const __m128i missing = _mm_cmpeq_epi8(x, ff);
      const __m128i v = _mm_add_epi8(_mm_min_epu8(x, cap), c33);  // min(x, 93) + 33: 126 for everything above 93
      _mm_storeu_si128((__m128i*)(dst + i), _mm_or_si128(_mm_andnot_si128(missing, v), _mm_and_si128(missing, c33)));
    }
#endif
    for (; i < l_seq; ++i) dst[i] = (uint8_t)(q[i] == 0xff ? '!' : (q[i] > 93 ? 126 : q[i] + 33));
  }
  R.seq_off.push_back((uint32_t)R.seq.size());
  q += l_seq;
  R.hp.push_back(aux_int(q, rec + n, 'H', 'P'));
  if (keep_raw) R.raw.insert(R.raw.end(), rec, rec + n);
  R.raw_off.push_back((uint32_t)R.raw.size());
  return true;
}

bool load_index(ltr_bam* b, const char* path) {
  Mapped ix;
  if (!ix.open(path) || ix.n < 8 || memcmp(ix.p, "BAI\1", 4) != 0) return false;
  const uint8_t* p = ix.p + 4;
  const uint8_t* e = ix.p + ix.n;
  const uint32_t n_ref = le32(p);
  p += 4;
  std::vector<RefIndex> idx(n_ref);
  for (uint32_t r = 0; r < n_ref; ++r) {
    if (p + 4 > e) return false;
    const uint32_t n_bin = le32(p);
    p += 4;
    for (uint32_t k = 0; k < n_bin; ++k) {
      if (p + 8 > e) return false;
      const uint32_t bin = le32(p), n_chunk = le32(p + 4);
      p += 8;
      if (p + 16ull * n_chunk > e) return false;
      std::vector<Chunk> chunks(n_chunk);
      for (uint32_t c = 0; c < n_chunk; ++c) {
        chunks[c].beg = le64(p + 16 * c);
        chunks[c].end = le64(p + 16 * c + 8);
      }
      p += 16ull * n_chunk;
      if (bin != 37450) idx[r].bins.emplace_back(bin, std::move(chunks));  // 37450: metadata pseudo-bin
    }
    std::sort(idx[r].bins.begin(), idx[r].bins.end(),
              [](const std::pair<uint32_t, std::vector<Chunk>>& a, const std::pair<uint32_t, std::vector<Chunk>>& c) { return a.first < c.first; });
    if (p + 4 > e) return false;
    const uint32_t n_intv = le32(p);
    p += 4;
    if (p + 8ull * n_intv > e) return false;
    idx[r].linear.resize(n_intv);
    for (uint32_t k = 0; k < n_intv; ++k) idx[r].linear[k] = le64(p + 8 * k);
    p += 8ull * n_intv;
  }
  b->index.swap(idx);
  b->has_index = true;
  return true;
}

}  // namespace

static int ltr_bam_open_impl(const char* path, const char* index_path, ltr_bam** out) {
  if (!path || !out) return LTR_ERR_INVALID;
  *out = nullptr;
  ltr_bam* b = new ltr_bam();
  if (!b->file.open(path)) {
    delete b;
    return LTR_ERR_INVALID;
  }
  Cursor c(&b->file);
  c.seek(0);
  uint8_t w[8];
  bool ok = c.read(w, 8) && memcmp(w, "BAM\1", 4) == 0;
  if (ok) {
    const uint32_t l_text = le32(w + 4);
    b->text.resize(l_text);
    ok = l_text == 0 || c.read(reinterpret_cast<uint8_t*>(&b->text[0]), l_text);
  }
  uint32_t n_ref = 0;
  if (ok) ok = c.read(w, 4);
  if (ok) n_ref = le32(w);
  for (uint32_t r = 0; ok && r < n_ref; ++r) {
    ok = c.read(w, 4);
    if (!ok) break;
    const uint32_t l_name = le32(w);
    std::string name(l_name, '\0');
    ok = l_name > 0 && l_name < (1u << 20) && c.read(reinterpret_cast<uint8_t*>(&name[0]), l_name) && c.read(w, 4);
    if (!ok) break;
    name.resize(strlen(name.c_str()));
    b->ref_names.push_back(name);
    b->ref_lens.push_back((int64_t)le32(w));
  }
  if (!ok) {
    delete b;
    return LTR_ERR_INVALID;
  }
  b->first_record = c.tell();
  const std::string ip = index_path ? std::string(index_path) : std::string(path) + ".bai";
  if (!load_index(b, ip.c_str()) && index_path) {  // an index that was asked for must load
    delete b;
    return LTR_ERR_INVALID;
  }
  *out = b;
  return LTR_OK;
}

// C ABI boundary: no exception leaves the library (malformed input and exhausted memory become error codes)
extern "C" int ltr_bam_open(const char* path, const char* index_path, ltr_bam** out) {
  try {
    return ltr_bam_open_impl(path, index_path, out);
  } catch (const std::bad_alloc&) {
    return LTR_ERR_OOM;
  } catch (...) {
    return LTR_ERR_INVALID;
  }
}

// reg2bin of the SAM specification (5.3): the smallest bin that contains [beg, end)
static uint32_t reg2bin(int64_t beg, int64_t end) {
  --end;
  if (beg >> 14 == end >> 14) return (uint32_t)(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (uint32_t)(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (uint32_t)(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (uint32_t)(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (uint32_t)(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

// Builds the binning + linear index of a coordinate-sorted file in memory (what `samtools index` writes to a .bai), for
// files that come without one.  Not thread safe against concurrent fetches on the same handle.
static int ltr_bam_build_index_impl(ltr_bam* b) {
  if (!b) return LTR_ERR_INVALID;
  if (b->has_index) return LTR_OK;
  std::vector<RefIndex> idx(b->ref_names.size());
  std::vector<std::vector<std::pair<uint32_t, Chunk>>> raw(b->ref_names.size());
  Cursor c(&b->file);
  c.seek(b->first_record);
  std::vector<uint8_t> rec;
  int32_t last_tid = 0, last_pos = -1;
  for (;;) {
    const uint64_t v0 = c.tell();
    uint8_t w[4];
    if (!c.read(w, 4)) break;
    const uint32_t bs = le32(w);
    if (bs < 32 || bs > (1u << 28)) return LTR_ERR_INVALID;
    rec.resize(bs);
    if (!c.read(rec.data(), bs)) return LTR_ERR_INVALID;
    const uint64_t v1 = c.tell();
    const int32_t tid = (int32_t)le32(rec.data()), pos = (int32_t)le32(rec.data() + 4);
    if (tid < 0) continue;
    if ((size_t)tid >= idx.size() || pos < 0) return LTR_ERR_INVALID;
    if (tid < last_tid || (tid == last_tid && pos < last_pos)) return LTR_ERR_UNSUPPORTED;  // not coordinate sorted
    last_tid = tid;
    last_pos = pos;
    const uint32_t l_name = rec[8], n_cig = le16(rec.data() + 12);
    if (32 + (size_t)l_name + 4 * (size_t)n_cig > bs) return LTR_ERR_INVALID;
    int64_t ref_len = 0;
    for (uint32_t k = 0; k < n_cig; ++k) {
      const uint32_t v = le32(rec.data() + 32 + l_name + 4 * k), op = v & 15;
      if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += v >> 4;
    }
    const int64_t end = (int64_t)pos + (ref_len ? ref_len : 1);
    const uint32_t bin = reg2bin(pos, end);
    std::vector<std::pair<uint32_t, Chunk>>& rv = raw[(size_t)tid];
    if (!rv.empty() && rv.back().first == bin && rv.back().second.end == v0) rv.back().second.end = v1;  // consecutive records
    else rv.push_back(std::make_pair(bin, Chunk{v0, v1}));
    std::vector<uint64_t>& lin = idx[(size_t)tid].linear;
    const size_t w0 = (size_t)(pos >> 14), w1 = (size_t)((end - 1) >> 14);
    if (lin.size() <= w1) lin.resize(w1 + 1, 0);
    for (size_t k = w0; k <= w1; ++k)
      if (lin[k] == 0) lin[k] = v0;
  }
  for (size_t t = 0; t < idx.size(); ++t) {
    std::stable_sort(raw[t].begin(), raw[t].end(),
                     [](const std::pair<uint32_t, Chunk>& a, const std::pair<uint32_t, Chunk>& d) { return a.first < d.first; });
    for (const std::pair<uint32_t, Chunk>& e : raw[t]) {
      if (idx[t].bins.empty() || idx[t].bins.back().first != e.first)
        idx[t].bins.emplace_back(e.first, std::vector<Chunk>());
      idx[t].bins.back().second.push_back(e.second);
    }
    // windows no record starts in inherit the offset of the next one that has one (as htslib fills its linear index)
    std::vector<uint64_t>& lin = idx[t].linear;
    for (size_t k = lin.size(); k-- > 0;)
      if (lin[k] == 0 && k + 1 < lin.size()) lin[k] = lin[k + 1];
  }
  b->index.swap(idx);
  b->has_index = true;
  return LTR_OK;
}

// C ABI boundary: no exception leaves the library (malformed input and exhausted memory become error codes)
extern "C" int ltr_bam_build_index(ltr_bam* b) {
  try {
    return ltr_bam_build_index_impl(b);
  } catch (const std::bad_alloc&) {
    return LTR_ERR_OOM;
  } catch (...) {
    return LTR_ERR_INVALID;
  }
}

extern "C" void ltr_bam_close(ltr_bam* b) { delete b; }
extern "C" int32_t ltr_bam_n_refs(const ltr_bam* b) { return b ? (int32_t)b->ref_names.size() : 0; }
extern "C" const char* ltr_bam_ref_name(const ltr_bam* b, int32_t tid) {
  return (b && tid >= 0 && tid < (int32_t)b->ref_names.size()) ? b->ref_names[(size_t)tid].c_str() : nullptr;
}
extern "C" int64_t ltr_bam_ref_len(const ltr_bam* b, int32_t tid) {
  return (b && tid >= 0 && tid < (int32_t)b->ref_lens.size()) ? b->ref_lens[(size_t)tid] : -1;
}
extern "C" int32_t ltr_bam_ref_id(const ltr_bam* b, const char* name) {
  if (!b || !name) return -1;
  for (size_t i = 0; i < b->ref_names.size(); ++i)
    if (b->ref_names[i] == name) return (int32_t)i;
  return -1;
}
extern "C" const char* ltr_bam_header_text(const ltr_bam* b) { return b ? b->text.c_str() : nullptr; }
extern "C" int ltr_bam_has_index(const ltr_bam* b) { return b && b->has_index; }

// Records that overlap [beg, end) on reference tid (tid < 0: every record of the file, in file order).
static int ltr_bam_fetch_impl(const ltr_bam* b, int32_t tid, int64_t beg, int64_t end, int32_t keep_raw, ltr_bam_reads** out) {
  if (!b || !out) return LTR_ERR_INVALID;
  *out = nullptr;
  if (tid >= (int32_t)b->ref_names.size()) return LTR_ERR_INVALID;
  ReadsOwner* R = new ReadsOwner();
  R->name_off.push_back(0); R->seq_off.push_back(0); R->cigar_off.push_back(0); R->raw_off.push_back(0);
  Cursor c(&b->file);
  std::vector<uint8_t> rec;
  std::vector<Chunk> todo;
  if (tid < 0 || !b->has_index) {
    todo.push_back(Chunk{b->first_record, ~0ull});
  } else {
    const RefIndex& ix = b->index[(size_t)tid];
    std::vector<uint32_t> bins;
    reg2bins(beg, end, bins);
    uint64_t min_off = 0;
    if (!ix.linear.empty()) {
      size_t w = (size_t)(std::max<int64_t>(beg, 0) >> 14);
      if (w >= ix.linear.size()) w = ix.linear.size() - 1;
      min_off = ix.linear[w];
    }
    for (uint32_t bin : bins) {
      const std::vector<Chunk>* ch = ix.find(bin);
      if (!ch) continue;
      for (const Chunk& k : *ch)
        if (k.end > min_off) todo.push_back(k);
    }
    std::sort(todo.begin(), todo.end(), [](const Chunk& a, const Chunk& d) { return a.beg < d.beg; });
    std::vector<Chunk> merged;  // overlapping / adjacent chunks become one walk (a record is visited once)
    for (const Chunk& k : todo) {
      if (!merged.empty() && k.beg <= merged.back().end) merged.back().end = std::max(merged.back().end, k.end);
      else merged.push_back(k);
    }
    todo.swap(merged);
  }
  int rc = LTR_OK;
  bool past = false;
  for (size_t t = 0; t < todo.size() && !past && rc == LTR_OK; ++t) {
    c.seek(todo[t].beg);
    while (true) {
      if (c.tell() >= todo[t].end) break;
      uint8_t w[4];
      if (!c.read(w, 4)) break;  // end of file
      const uint32_t bs = le32(w);
      if (bs < 32 || bs > (1u << 28)) { rc = LTR_ERR_INVALID; break; }
      rec.resize(bs);
      if (!c.read(rec.data(), bs)) { rc = LTR_ERR_INVALID; break; }
      if (tid >= 0) {
        const int32_t rt = (int32_t)le32(rec.data()), rp = (int32_t)le32(rec.data() + 4);
        if (rt != tid) {
          if (!b->has_index && rt >= 0 && rt < tid) continue;  // coordinate-sorted file without index: not there yet
          past = true;  // a chunk of this reference ran into the next one / the scan left the reference
          break;
        }
        if (rp >= end) { past = true; break; }
        // end position needs the CIGAR: decode cheaply here
        const uint32_t l_name = rec[8], n_cig = le16(rec.data() + 12);
        if (32 + (size_t)l_name + 4 * (size_t)n_cig > bs) { rc = LTR_ERR_INVALID; break; }
        int64_t ref_len = 0;
        for (uint32_t k = 0; k < n_cig; ++k) {
          const uint32_t v = le32(rec.data() + 32 + l_name + 4 * k), op = v & 15;
          if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += v >> 4;
        }
        if ((int64_t)rp + (ref_len ? ref_len : 1) <= beg) continue;
      }
      if (!append_record(*R, rec.data(), bs, keep_raw != 0)) { rc = LTR_ERR_INVALID; break; }
    }
  }
  if (rc != LTR_OK) {
    delete R;
    return rc;
  }
  R->publish();
  *out = &R->pub;
  return LTR_OK;
}

// C ABI boundary: no exception leaves the library (malformed input and exhausted memory become error codes)
extern "C" int ltr_bam_fetch(const ltr_bam* b, int32_t tid, int64_t beg, int64_t end, int32_t keep_raw, ltr_bam_reads** out) {
  try {
    return ltr_bam_fetch_impl(b, tid, beg, end, keep_raw, out);
  } catch (const std::bad_alloc&) {
    return LTR_ERR_OOM;
  } catch (...) {
    return LTR_ERR_INVALID;
  }
}

extern "C" void ltr_bam_reads_free(ltr_bam_reads* r) {
  if (!r) return;
  delete static_cast<ReadsOwner*>(r->owner);
}
