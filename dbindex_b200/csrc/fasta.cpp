// fasta.cpp -- host-side FASTA ingest of the C ABI (no CUDA calls): the step before the hot path.
//
// The reference streams the FASTA through the external FastaReader of the un-vendored utilities
// jar, up to three passes (DBIndexer.java:560-571), and hands each record to
// ProteinCache.addProtein(defline, sequence) (DBIndexer.java:605, ProteinCache.java:84-95).  At
// TrEMBL scale (3 G residues) a single-threaded reader is what the GPU build would wait for, so
// this one maps the file and parses it with all host threads in two passes (count, then copy),
// straight into the packed layout dbi_add_proteins takes: residues[] + offsets[n + 1], plus the
// deflines for the host-side protein table.
//
// Record grammar (the contract of dbindex_b200/indexer.py::read_fasta, which the tests hold this
// against): a line starting with '>' opens a record, its defline is the rest of that line without
// the line end; the sequence is the concatenation of the following non-empty lines, each stripped
// of leading / trailing white space and upper-cased (ASCII); lines before the first '>' are ignored.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstring>
#include <exception>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/dbindex_gpu.h"

namespace dbi {
void set_error(const char* fmt, ...);
}

struct dbi_fasta {
  const char* data = nullptr;
  size_t size = 0;
  int fd = -1;
  // per record: byte offset of its '>' in the file; records are in file order
  std::vector<uint64_t> rec_begin;
  std::vector<uint64_t> res_off;  // n + 1: exclusive prefix sums of the residue counts
  std::vector<uint64_t> def_off;  // n + 1: exclusive prefix sums of the defline lengths
  int n_threads = 1;
};

namespace {

inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// [b, e) = one line without its '\n'; returns the stripped range
inline void strip(const char*& b, const char*& e) {
  while (b < e && is_space((unsigned char)*b)) ++b;
  while (e > b && is_space((unsigned char)e[-1])) --e;
}

// Walks the lines of one record body [p, end) and calls f(b, e) for every stripped non-empty line.
template <typename F>
inline void for_each_seq_line(const char* p, const char* end, F&& f) {
  while (p < end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* le = nl ? nl : end;
    const char *b = p, *e = le;
    strip(b, e);
    if (b < e) f(b, e);
    p = nl ? nl + 1 : end;
  }
}

// end of the defline that starts at p (p points at '>'): up to '\n', without a trailing '\r'
inline const char* defline_end(const char* p, const char* end, const char** next_line) {
  const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
  const char* le = nl ? nl : end;
  *next_line = nl ? nl + 1 : end;
  while (le > p && (le[-1] == '\r' || le[-1] == '\n')) --le;
  return le;
}

template <typename F>
void parallel_for(int n_threads, size_t n, F&& f) {
  if (n_threads <= 1 || n < 2) {
    f(0, n, 0);
    return;
  }
  // A worker that throws (bad_alloc of a push_back) must not escape its thread, and a failed
  // emplace_back must not unwind past joinable threads: both would be std::terminate instead of a
  // DBI_* code.  The first failure is kept and rethrown on the calling thread after every join.
  std::vector<std::thread> th;
  const size_t T = std::min<size_t>((size_t)n_threads, n);
  std::exception_ptr first;
  std::mutex mu;
  auto guarded = [&](size_t t) {
    try {
      f(n * t / T, n * (t + 1) / T, (int)t);
    } catch (...) {
      std::lock_guard<std::mutex> lk(mu);
      if (!first) first = std::current_exception();
    }
  };
  try {
    th.reserve(T);
    for (size_t t = 0; t < T; ++t) th.emplace_back(guarded, t);
  } catch (...) {
    std::lock_guard<std::mutex> lk(mu);
    if (!first) first = std::current_exception();
  }
  for (auto& x : th) x.join();
  if (first) std::rethrow_exception(first);
}

}  // namespace

extern "C" {

int dbi_fasta_open(const char* path, int n_threads, dbi_fasta** out) {
  if (!path || !out) {
    dbi::set_error("null argument");
    return DBI_EINVAL;
  }
  *out = nullptr;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) {
    dbi::set_error("cannot open %s: %s", path, strerror(errno));
    return DBI_EINVAL;
  }
  struct stat st;
  if (fstat(fd, &st) != 0) {
    dbi::set_error("cannot stat %s: %s", path, strerror(errno));
    close(fd);
    return DBI_EINVAL;
  }
  dbi_fasta* f = nullptr;
  try {
  f = new dbi_fasta();
  f->fd = fd;
  f->size = (size_t)st.st_size;
  f->n_threads = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
  if (f->size) {
    void* m = mmap(nullptr, f->size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) {
      dbi::set_error("cannot map %s: %s", path, strerror(errno));
      close(fd);
      delete f;
      return DBI_ENOMEM;
    }
    madvise(m, f->size, MADV_SEQUENTIAL);
    f->data = (const char*)m;
  }
  const char* d = f->data;
  const size_t n = f->size;
  // pass 1a: record starts = '>' at the beginning of a line, found chunk by chunk
  const int T = f->n_threads;
  std::vector<std::vector<uint64_t>> found((size_t)std::max(T, 1));
  parallel_for(T, n, [&](size_t lo, size_t hi, int t) {
    auto& v = found[(size_t)t];
    const char* p = d + lo;
    const char* e = d + hi;
    while (p < e) {
      const char* q = (const char*)memchr(p, '>', (size_t)(e - p));
      if (!q) break;
      if (q == d || q[-1] == '\n') v.push_back((uint64_t)(q - d));
      p = q + 1;
    }
  });
  for (auto& v : found) f->rec_begin.insert(f->rec_begin.end(), v.begin(), v.end());
  const size_t R = f->rec_begin.size();
  if (R >= (1ull << 31)) {
    dbi::set_error("more than 2^31 FASTA records");
    dbi_fasta_close(f);
    return DBI_ERANGE;
  }
  // pass 1b: residues and defline bytes per record
  f->res_off.assign(R + 1, 0);
  f->def_off.assign(R + 1, 0);
  parallel_for(T, R, [&](size_t lo, size_t hi, int) {
    for (size_t r = lo; r < hi; ++r) {
      const char* p = d + f->rec_begin[r];
      const char* end = d + (r + 1 < R ? f->rec_begin[r + 1] : n);
      const char* body;
      const char* de = defline_end(p, end, &body);
      f->def_off[r + 1] = (uint64_t)(de - (p + 1));
      uint64_t cnt = 0;
      for_each_seq_line(body, end, [&](const char* b, const char* e) { cnt += (uint64_t)(e - b); });
      f->res_off[r + 1] = cnt;
    }
  });
  for (size_t r = 0; r < R; ++r) {
    f->res_off[r + 1] += f->res_off[r];
    f->def_off[r + 1] += f->def_off[r];
  }
  *out = f;
  return DBI_OK;
  } catch (const std::exception& e) {  // bad_alloc, or a thread that could not be started
    dbi::set_error("FASTA ingest failed: %s", e.what());
    if (f) dbi_fasta_close(f); else close(fd);
    return DBI_ENOMEM;
  }
}

int dbi_fasta_counts(const dbi_fasta* f, uint32_t* n_proteins, uint64_t* n_residues, uint64_t* defline_bytes) {
  if (!f) {
    dbi::set_error("null argument");
    return DBI_EINVAL;
  }
  if (n_proteins) *n_proteins = (uint32_t)f->rec_begin.size();
  if (n_residues) *n_residues = f->res_off.back();
  if (defline_bytes) *defline_bytes = f->def_off.back();
  return DBI_OK;
}

int dbi_fasta_read(const dbi_fasta* f, uint8_t* residues, uint64_t* offsets, char* deflines, uint64_t* defline_off) {
  if (!f) {
    dbi::set_error("null argument");
    return DBI_EINVAL;
  }
  const size_t R = f->rec_begin.size();
  if ((f->res_off.back() && !residues) || !offsets) {
    dbi::set_error("null argument");
    return DBI_EINVAL;
  }
  std::memcpy(offsets, f->res_off.data(), (R + 1) * sizeof(uint64_t));
  if (defline_off) std::memcpy(defline_off, f->def_off.data(), (R + 1) * sizeof(uint64_t));
  const char* d = f->data;
  const size_t n = f->size;
  try {
  parallel_for(f->n_threads, R, [&](size_t lo, size_t hi, int) {
    for (size_t r = lo; r < hi; ++r) {
      const char* p = d + f->rec_begin[r];
      const char* end = d + (r + 1 < R ? f->rec_begin[r + 1] : n);
      const char* body;
      const char* de = defline_end(p, end, &body);
      if (deflines) std::memcpy(deflines + f->def_off[r], p + 1, (size_t)(de - (p + 1)));
      uint8_t* o = residues + f->res_off[r];
      for_each_seq_line(body, end, [&](const char* b, const char* e) {
        for (const char* q = b; q < e; ++q) {
          const unsigned char c = (unsigned char)*q;
          *o++ = (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : (uint8_t)c;
        }
      });
    }
  });
  } catch (const std::exception& e) {
    dbi::set_error("FASTA ingest failed: %s", e.what());
    return DBI_ENOMEM;
  }
  return DBI_OK;
}

void dbi_fasta_close(dbi_fasta* f) {
  if (!f) return;
  if (f->data) munmap((void*)f->data, f->size);
  if (f->fd >= 0) close(f->fd);
  delete f;
}

}  // extern "C"
