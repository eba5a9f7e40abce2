// fasta_reader.cpp -- ltr_fasta_*: indexed FASTA access for the region loop, and ltr_bed_read.
// The reference goes through htslib's faidx (src/fasta_reader.{h,cpp}: FastaReader::init / add_index / get_sequence /
// get_sequence_length; fai_load, fai_fetch) -- un-vendored here like the rest of htslib, so this is a reader of its own:
// the file is mapped, the `.fai` index (NAME LENGTH OFFSET LINEBASES LINEWIDTH per line, the samtools faidx format) is parsed
// or, when it is missing, built in memory by one pass over the file.  As in FastaReader::init a path is either one FASTA file
// or a directory whose `*.fa` files are all loaded; a sequence name that occurs twice is an error (fasta_reader.cpp:34-35).
// bgzip-compressed FASTA (.gz + .gzi) is not supported.  ltr_bed_read: readRegions + orderRegions (src/region.cpp:26-75).
#include <dirent.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "longtr_b200.h"

namespace {

struct FaiEntry {
  std::string name;
  int64_t length, offset;
  int32_t line_bases, line_width;
  int file;
};

struct FastaFile {
  const uint8_t* data = nullptr;
  size_t size = 0;
};

}  // namespace

struct ltr_fasta {
  std::vector<FastaFile> files;
  std::vector<FaiEntry> seqs;
  std::map<std::string, size_t> by_name;
};

namespace {

bool parse_fai(const std::string& path, int file, std::vector<FaiEntry>& out) {
  std::ifstream f(path.c_str());
  if (!f.is_open()) return false;
  std::string line;
  while (std::getline(f, line)) {
    if (line.empty()) continue;
    std::istringstream ss(line);
    FaiEntry e;
    std::string name;
    if (!std::getline(ss, name, '\t')) return false;
    e.name = name;
    if (!(ss >> e.length >> e.offset >> e.line_bases >> e.line_width)) return false;
    if (e.length < 0 || e.offset < 0 || e.line_bases < 1 || e.line_width < e.line_bases) return false;
    e.file = file;
    out.push_back(e);
  }
  return true;
}

// One pass over the file: what `samtools faidx` would write.  false: lines of unequal length inside a sequence.
bool build_fai(const FastaFile& F, int file, std::vector<FaiEntry>& out) {
  size_t p = 0;
  while (p < F.size) {
    if (F.data[p] != '>') {  // stray text before the first header / blank lines
      const void* nl = memchr(F.data + p, '\n', F.size - p);
      if (!nl) break;
      p = (size_t)((const uint8_t*)nl - F.data) + 1;
      continue;
    }
    const void* nl = memchr(F.data + p, '\n', F.size - p);
    const size_t eol = nl ? (size_t)((const uint8_t*)nl - F.data) : F.size;
    size_t e = p + 1;
    while (e < eol && F.data[e] != ' ' && F.data[e] != '\t' && F.data[e] != '\r') ++e;
    FaiEntry E;
    E.name.assign((const char*)F.data + p + 1, e - p - 1);
    E.file = file;
    E.offset = (int64_t)(eol < F.size ? eol + 1 : F.size);
    E.length = 0;
    E.line_bases = E.line_width = 0;
    p = (size_t)E.offset;
    bool short_line_seen = false;
    while (p < F.size && F.data[p] != '>') {
      const void* n2 = memchr(F.data + p, '\n', F.size - p);
      const size_t end = n2 ? (size_t)((const uint8_t*)n2 - F.data) : F.size;
      size_t bases = end - p;
      if (bases && F.data[end - 1] == '\r') --bases;
      const size_t width = (n2 ? end + 1 : end) - p;
      if (bases) {
        if (short_line_seen) return false;  // a shorter line may only be the last one
        if (E.line_bases == 0) {
          E.line_bases = (int32_t)bases;
          E.line_width = (int32_t)width;
        } else if ((int32_t)bases != E.line_bases) {
          if ((int32_t)bases > E.line_bases) return false;
          short_line_seen = true;
        }
        E.length += (int64_t)bases;
      } else if (E.line_bases) {
        short_line_seen = true;  // blank line: nothing but the end may follow
      }
      p = n2 ? end + 1 : F.size;
    }
    if (E.line_bases == 0) E.line_bases = E.line_width = 1;
    out.push_back(E);
  }
  return true;
}

int add_file(ltr_fasta* fa, const std::string& path) {
  const int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) return LTR_ERR_INVALID;
  struct stat st;
  if (fstat(fd, &st) != 0) {
    close(fd);
    return LTR_ERR_INVALID;
  }
  FastaFile F;
  F.size = (size_t)st.st_size;
  if (F.size) {
    void* m = mmap(nullptr, F.size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) {
      close(fd);
      return LTR_ERR_OOM;
    }
    F.data = (const uint8_t*)m;
  }
  close(fd);
  if (F.size >= 2 && F.data[0] == 0x1f && F.data[1] == 0x8b) {  // gzip / bgzip
    munmap((void*)F.data, F.size);
    return LTR_ERR_UNSUPPORTED;
  }
  const int file = (int)fa->files.size();
  fa->files.push_back(F);
  std::vector<FaiEntry> entries;
  bool ok;
  if (access((path + ".fai").c_str(), F_OK) == 0) ok = parse_fai(path + ".fai", file, entries);
  else ok = build_fai(F, file, entries);
  if (!ok) return LTR_ERR_INVALID;
  for (const FaiEntry& e : entries) {
    if (e.length > 0) {  // the last base must lie inside the file
      const int64_t last = e.offset + (e.length - 1) / e.line_bases * e.line_width + (e.length - 1) % e.line_bases;
      if (last >= (int64_t)F.size) return LTR_ERR_INVALID;
    }
    if (fa->by_name.count(e.name)) return LTR_ERR_INVALID;  // "Multiple entries for chromosome ..."
    fa->by_name[e.name] = fa->seqs.size();
    fa->seqs.push_back(e);
  }
  return LTR_OK;
}

bool ends_with(const std::string& s, const char* suffix) {
  const size_t n = strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

}  // namespace

extern "C" int ltr_fasta_open(const char* path, ltr_fasta** out) {
  if (!path || !out) return LTR_ERR_INVALID;
  *out = nullptr;
  struct stat st;
  if (stat(path, &st) != 0) return LTR_ERR_INVALID;
  ltr_fasta* fa = new ltr_fasta();
  int rc = LTR_OK;
  if (S_ISREG(st.st_mode)) {
    rc = add_file(fa, path);
  } else {
    DIR* dir = opendir(path);
    if (!dir) rc = LTR_ERR_INVALID;
    std::vector<std::string> names;
    if (dir) {
      while (struct dirent* d = readdir(dir))
        if (ends_with(d->d_name, ".fa")) names.push_back(d->d_name);
      closedir(dir);
      std::sort(names.begin(), names.end());
      if (names.empty()) rc = LTR_ERR_INVALID;  // "Failed to locate any FASTA files in the provided directory"
    }
    for (size_t i = 0; i < names.size() && rc == LTR_OK; ++i) rc = add_file(fa, std::string(path) + "/" + names[i]);
  }
  if (rc != LTR_OK) {
    ltr_fasta_close(fa);
    return rc;
  }
  *out = fa;
  return LTR_OK;
}

extern "C" void ltr_fasta_close(ltr_fasta* fa) {
  if (!fa) return;
  for (FastaFile& F : fa->files)
    if (F.data) munmap((void*)F.data, F.size);
  delete fa;
}

extern "C" int32_t ltr_fasta_n_seqs(const ltr_fasta* fa) { return fa ? (int32_t)fa->seqs.size() : 0; }

extern "C" const char* ltr_fasta_seq_name(const ltr_fasta* fa, int32_t i) {
  return (fa && i >= 0 && (size_t)i < fa->seqs.size()) ? fa->seqs[(size_t)i].name.c_str() : nullptr;
}

extern "C" int64_t ltr_fasta_seq_len(const ltr_fasta* fa, const char* name) {
  if (!fa || !name) return -1;
  auto it = fa->by_name.find(name);
  return it == fa->by_name.end() ? -1 : fa->seqs[it->second].length;
}

// Bases [start, end) of the named sequence, as they stand in the file (case kept: the region loop upper-cases what it uses).
extern "C" int ltr_fasta_fetch(const ltr_fasta* fa, const char* name, int64_t start, int64_t end, uint8_t* out) {
  if (!fa || !name || (!out && end > start)) return LTR_ERR_INVALID;
  auto it = fa->by_name.find(name);
  if (it == fa->by_name.end()) return LTR_ERR_INVALID;
  const FaiEntry& e = fa->seqs[it->second];
  if (start < 0 || end < start || end > e.length) return LTR_ERR_INVALID;
  const uint8_t* d = fa->files[(size_t)e.file].data;
  int64_t pos = start;
  while (pos < end) {
    const int64_t in_line = pos % e.line_bases;
    const int64_t n = std::min<int64_t>(e.line_bases - in_line, end - pos);
    memcpy(out + (pos - start), d + e.offset + pos / e.line_bases * e.line_width + in_line, (size_t)n);
    pos += n;
  }
  return LTR_OK;
}

// ---- region file ----------------------------------------------------------------------------------------------------------
namespace {
struct BedOwner {
  ltr_bed pub;
  std::vector<std::string> chroms, names, motifs;
  std::vector<const char*> chrom_p, name_p, motif_p;
  std::vector<int32_t> chrom_of;
  std::vector<ltr_region> regions;
};
struct BedLine {
  std::string chrom, name, motif;
  int32_t start, stop;
};
int motif_period(const std::string& motifs) {  // Region::computePeriod (src/region.h:36-43)
  std::set<int> periods;
  std::stringstream ss(motifs);
  std::string item;
  while (std::getline(ss, item, ',')) periods.insert((int)item.size());
  return periods.size() == 1 ? *periods.begin() : -1;
}
}  // namespace

extern "C" int ltr_bed_read(const char* path, uint32_t max_regions, const char* chrom_limit, ltr_bed** out) {
  if (!path || !out) return LTR_ERR_INVALID;
  *out = nullptr;
  std::ifstream f(path);
  if (!f.is_open()) return LTR_ERR_INVALID;  // "Failed to open region file"
  const std::string limit = chrom_limit ? chrom_limit : "";
  std::vector<BedLine> lines;
  std::string line;
  while (std::getline(f, line) && (max_regions == 0 || lines.size() < max_regions)) {
    std::istringstream iss(line);
    BedLine L;
    if (!(iss >> L.chrom >> L.start >> L.stop >> L.motif)) return LTR_ERR_INVALID;  // "Improperly formatted region file"
    if (L.start < 1 || L.stop <= L.start || L.motif.empty()) return LTR_ERR_INVALID;
    for (char ch : L.motif)
      if (!isalpha((unsigned char)ch) && ch != ',') return LTR_ERR_INVALID;
    if (!limit.empty() && L.chrom != limit) continue;
    if (!(iss >> L.name)) L.name.clear();
    L.start -= 1;  // BED lines hold 1-based starts here (region.cpp:52-54)
    lines.push_back(L);
  }
  if (!limit.empty() && lines.empty()) return LTR_ERR_INVALID;
  // orderRegions: Region::operator< -- chromosome, then start, then stop (src/region.h)
  std::stable_sort(lines.begin(), lines.end(), [](const BedLine& a, const BedLine& b) {
    if (a.chrom != b.chrom) return a.chrom < b.chrom;
    if (a.start != b.start) return a.start < b.start;
    return a.stop < b.stop;
  });
  BedOwner* O = new BedOwner();
  for (const BedLine& L : lines) {
    if (O->chroms.empty() || O->chroms.back() != L.chrom) O->chroms.push_back(L.chrom);
    O->chrom_of.push_back((int32_t)O->chroms.size() - 1);
    ltr_region r;
    r.start = L.start;
    r.stop = L.stop;
    r.period = motif_period(L.motif);
    O->regions.push_back(r);
    O->names.push_back(L.name);
    O->motifs.push_back(L.motif);
  }
  for (const std::string& s : O->chroms) O->chrom_p.push_back(s.c_str());
  for (const std::string& s : O->names) O->name_p.push_back(s.c_str());
  for (const std::string& s : O->motifs) O->motif_p.push_back(s.c_str());
  O->pub.n_regions = (uint32_t)O->regions.size();
  O->pub.regions = O->regions.data();
  O->pub.region_chrom = O->chrom_of.data();
  O->pub.names = O->name_p.data();
  O->pub.motifs = O->motif_p.data();
  O->pub.n_chroms = (uint32_t)O->chroms.size();
  O->pub.chroms = O->chrom_p.data();
  O->pub.owner = O;
  *out = &O->pub;
  return LTR_OK;
}

extern "C" void ltr_bed_free(ltr_bed* b) {
  if (b) delete static_cast<BedOwner*>(b->owner);
}
