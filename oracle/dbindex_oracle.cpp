// dbindex_oracle.cpp -- CPU restatement of the dbIndex index-build + precursor
// lookup path.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT. ***  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library, and
// only as the checker or the reported CPU baseline.  The product path
// (libdbindex_gpu.so) never links, loads or calls it.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
// this path (src/test/.gitignore:1-2), it cannot be compiled here (no JDK, and
// its arithmetic core lives in the un-vendored edu.scripps.yates:utilities
// 1.6-SNAPSHOT, pom.xml:86-90), so this restatement is anchored on the
// reference's own call sites only.  The two external pieces (AssignMass table,
// Enzyme.checkCleavage) follow the written-down contract of SURVEY.md 8(c) and
// are INPUTS (dbi_params tables), not constants.  Differential-mod expansion has
// no reference implementation at all; DESIGN.md "Mod expansion SPEC" defines it.
//
// Paths below are relative to src/main/java/edu/scripps/yates/dbindex/.
//
// Build: make -C oracle   (g++ -O2 -ffp-contract=off -fopenmp, no other deps)

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/dbindex_gpu.h"

namespace {

// One 20-byte record of DBIndexStoreSQLiteByte.updateCachedData
// (DBIndexStoreSQLiteByte.java:202-212; Constants.BYTE_PER_SEQUENCE, Constants.java:60).
struct Rec {
  double mass;
  int32_t off;
  int32_t len;
  int32_t prot;
};

// IndexedSeqMerged (IndexedSeqMerged.java:15-27): first occurrence + all protein ids.
struct Merged {
  double mass;
  int32_t off;
  int32_t len;
  std::vector<int32_t> prots;  // proteinIds, duplicates kept, insertion order (Merge:658-663,678-687)
};

// One searchable entry of the finished index: a unique peptide or one of its
// differential-mod variants (SPEC in DESIGN.md).
struct Entry {
  double mass;
  uint32_t base;    // index into Oracle::uniq
  uint32_t modpat;  // byte k = position+1 of k-th modified residue, 0 = none
};

struct Oracle {
  dbi_params p;
  std::vector<std::string> prot;  // ProteinCache.sequences (ProteinCache.java:24-25)
  std::vector<Rec> emitted;       // addSequence() calls in order
  std::vector<Merged> uniq;       // rows ascending, row order inside (flattened table)
  std::vector<int32_t> uniq_key;  // precursor_mass_key of each uniq entry
  std::vector<Entry> entries;     // searchable entries, mass ascending
  bool built = false;
  int err = 0;
};

inline uint64_t dbits(double d) {
  uint64_t u;
  std::memcpy(&u, &d, 8);
  return u;
}

// Enzyme.checkCleavage(protSeq, start, end, nocut) -- external class; contract of
// SURVEY.md 8(c), call site DBIndexer.java:318-319.  C-terminal cutter (the enzyme
// offset is never handed to Enzyme, SearchParams.java:303).
inline bool check_cleavage(const dbi_params& p, const std::string& s, int start, int end) {
  const int L = (int)s.size();
  const bool n_ok = (start == 0) || (p.is_enzyme[(uint8_t)s[start - 1]] && !p.is_nocut[(uint8_t)s[start]]);
  const bool c_ok = (end == L - 1) || (p.is_enzyme[(uint8_t)s[end]] && !p.is_nocut[(uint8_t)s[end + 1]]);
  return p.semi ? (n_ok || c_ok) : (n_ok && c_ok);
}

// PeptideFilterByMaxOccurrencies.isValid (util/PeptideFilterByMaxOccurrencies.java:22-34).
inline bool filter_is_valid(const dbi_params& p, const char* pep, int n) {
  if (p.filter_aa <= 0) return true;  // peptideFilter == null
  int matches = 0;
  for (int i = 0; i < n; ++i)
    if ((uint8_t)pep[i] == (uint8_t)p.filter_aa) {
      ++matches;
      if (matches > p.filter_max) return false;
    }
  return true;
}

// DBIndexStoreSQLiteMult.filterSequence (DBIndexStoreSQLiteMult.java:245-268): true = INCLUDE.
inline bool store_filter_include(const dbi_params& p, double mass, const char* pep, int n) {
  bool any_mandatory = false;
  for (int c = 0; c < 256; ++c) any_mandatory |= p.is_mandatory[c] != 0;
  if (p.has_mandatory && any_mandatory) {  // "!= null && length > 0"
    for (int i = 0; i + 1 < n; ++i)        // the last AA, the cleavage site, is excluded (:255-256)
      if (p.is_mandatory[(uint8_t)pep[i]]) return true;
    return false;  // SKIP
  }
  return !(p.max_mass < mass || p.min_mass > mass);
}

// DBIndexer.cutSeq (DBIndexer.java:237-405); bracketed-formula PTMs (:288-303) excluded (they need
// the un-vendored FormulaCalculator and rewrite the protein string).
void cut_seq(const dbi_params& p, const std::string& seq, int32_t prot_id, std::vector<Rec>& out) {
  const int length = (int)seq.size();
  const int max_mc = p.max_missed;
  for (int start = 0; start < length; ++start) {  // :256
    int end = start;
    double prec = 0;  // :265
    if (p.add_h2o_proton) prec += p.h2o_proton;  // :268-269
    prec += p.cterm;  // :270
    prec += p.nterm;  // :271
    int pep_size = 0;
    int mc = -1;  // :280
    while (prec <= p.max_mass && end < length) {  // :284
      pep_size++;
      const uint8_t c = (uint8_t)seq[end];
      prec = prec + p.residue_mass[c];           // :306-308
      if (!filter_is_valid(p, seq.data() + start, pep_size)) break;  // :310-313
      if (p.is_enzyme[c]) mc++;                  // :314-316
      if (check_cleavage(p, seq, start, end)) {  // :318-320
        if (mc > max_mc) break;                  // :322-324
        if (prec > p.max_mass) break;            // :327-329
        if (pep_size >= p.min_len && prec >= p.min_mass) {  // :331
          if (p.has_mandatory) {  // mandatoryInternalAAs != null, :334-344
            bool found = false;
            for (int i = 0; i < pep_size; ++i)
              if (p.is_mandatory[(uint8_t)seq[start + i]]) found = true;
            if (!found) break;
          }
          // indexStore.filterSequence (:347); SKIP_PROTEIN_START is never returned by this store
          if (store_filter_include(p, prec, seq.data() + start, pep_size))
            out.push_back(Rec{prec, start, pep_size, prot_id});  // addSequence(mass, start, curSeqI, ...) :388
        }
      }
      ++end;  // :394
    }
  }
}

// precursor_mass_key: (int)(precMass * massGroupFactor), truncation toward zero
// (DBIndexStoreSQLiteByte.java:187).
// Java's (int) cast of a double: truncates toward zero and SATURATES (JLS 5.1.3); a plain C++
// cast is undefined out of range.
inline int32_t java_int(double v) {
  if (v != v) return 0;
  if (v >= 2147483647.0) return INT32_MAX;
  if (v <= -2147483648.0) return INT32_MIN;
  return (int32_t)v;
}
inline int32_t mass_key(const dbi_params& p, double m) { return java_int(m * p.mass_group_factor); }

// DBIndexStoreSQLiteByteIndexMerge.getMergedData (Merge:620-719) for one row.
// THashMap iteration order is unspecified in the reference; here groups keep
// first-occurrence order, then the stable Collections.sort by mass
// (IndexedSeqMerged.compareTo, IndexedSeqMerged.java:30-38).
void merge_row(const Oracle& o, const Rec* recs, size_t n, std::vector<Merged>& out) {
  std::unordered_map<std::string, size_t> pos;
  std::vector<Merged> groups;
  for (size_t i = 0; i < n; ++i) {
    const Rec& r = recs[i];
    // proteinCache.getPeptideSequence(proteinId, offset, length)  (Merge:653, ProteinCache.java:112-127)
    std::string pep = o.prot[r.prot].substr(r.off, r.len);
    auto it = pos.find(pep);
    if (it == pos.end()) {
      pos.emplace(std::move(pep), groups.size());
      groups.push_back(Merged{r.mass, r.off, r.len, {r.prot}});  // firstSeq (Merge:684-687)
    } else {
      groups[it->second].prots.push_back(r.prot);  // Merge:678-681
    }
  }
  std::stable_sort(groups.begin(), groups.end(),
                   [](const Merged& a, const Merged& b) { return a.mass < b.mass; });  // Merge:693
  for (auto& g : groups) out.push_back(std::move(g));
}

// Differential-mod expansion SPEC (no reference implementation; parameters from
// io/SearchParamReader.java:631-687, model/DiffModification.java:11-54).
int expand_mods(Oracle& o) {
  const dbi_params& p = o.p;
  double diff[256];
  bool is_diff[256];
  for (int i = 0; i < 256; ++i) { diff[i] = 0; is_diff[i] = false; }
  for (int i = 0; i < p.n_mods; ++i) {  // DiffModification.setDiffModMass: last one wins
    diff[p.mods[i].residue] = p.mods[i].delta;
    is_diff[p.mods[i].residue] = true;
  }
  const int K = p.n_mods > 0 ? p.max_mods_per_peptide : 0;
  std::vector<int> sites;
  for (uint32_t b = 0; b < o.uniq.size(); ++b) {
    const Merged& u = o.uniq[b];
    const std::string& ps = o.prot[u.prots[0]];
    sites.clear();
    if (K > 0)
      for (int i = 0; i < u.len; ++i)
        if (is_diff[(uint8_t)ps[u.off + i]]) {
          // the mod pattern stores positions in 8 bits: a modifiable residue beyond
          // DBI_MAX_MOD_POS is a hard limit of the index format (DBI_ERANGE)
          if (i > DBI_MAX_MOD_POS) return DBI_ERANGE;
          sites.push_back(i);
        }
    const int n = (int)sites.size();
    // k = 0: the unmodified peptide, always in range
    o.entries.push_back(Entry{u.mass, b, 0u});
    int idx[DBI_MAX_MODS_PER_PEP];
    for (int k = 1; k <= K && k <= n; ++k) {
      for (int i = 0; i < k; ++i) idx[i] = i;
      while (true) {
        double m = u.mass;
        uint32_t pat = 0;
        for (int i = 0; i < k; ++i) {
          const int pp = sites[idx[i]];
          m = m + diff[(uint8_t)ps[u.off + pp]];  // left to right
          pat |= (uint32_t)(pp + 1) << (8 * i);
        }
        if (m >= p.min_mass && m <= p.max_mass) o.entries.push_back(Entry{m, b, pat});
        // next k-subset in lexicographic order
        int i = k - 1;
        while (i >= 0 && idx[i] == n - k + i) --i;
        if (i < 0) break;
        ++idx[i];
        for (int j = i + 1; j < k; ++j) idx[j] = idx[j - 1] + 1;
      }
    }
  }
  // the variants enter the store like any other sequence: key rows, mass order inside
  std::stable_sort(o.entries.begin(), o.entries.end(),
                   [](const Entry& a, const Entry& b) { return a.mass < b.mass; });
  return 0;
}

void flatten_rows(Oracle& o, std::vector<Rec>& recs) {
  // rows: stable order by key = TIntObjectHashMap<DynByteBuffer> append order per key
  // (DBIndexStoreSQLiteByte.java:193-212), then `SELECT ... ` row by row (Merge:85-127).
  const size_t n = recs.size();
  std::vector<int32_t> key(n);
  for (size_t i = 0; i < n; ++i) key[i] = mass_key(o.p, recs[i].mass);
  std::vector<uint32_t> order(n);
  for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
  std::vector<Rec> sorted(n);
  for (size_t i = 0; i < n; ++i) sorted[i] = recs[order[i]];
  // row boundaries
  std::vector<size_t> row_start;
  for (size_t i = 0; i < n; ++i)
    if (i == 0 || key[order[i]] != key[order[i - 1]]) row_start.push_back(i);
  row_start.push_back(n);
  const size_t nrows = row_start.size() - 1;
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  std::vector<std::vector<Merged>> part(nt);
  std::vector<std::vector<int32_t>> part_key(nt);
#pragma omp parallel for schedule(static)
  for (int t = 0; t < nt; ++t) {
    const size_t r0 = nrows * t / nt, r1 = nrows * (t + 1) / nt;
    for (size_t r = r0; r < r1; ++r) {
      const size_t before = part[t].size();
      merge_row(o, &sorted[row_start[r]], row_start[r + 1] - row_start[r], part[t]);
      const int32_t k = key[order[row_start[r]]];
      for (size_t j = before; j < part[t].size(); ++j) part_key[t].push_back(k);
    }
  }
  for (int t = 0; t < nt; ++t) {
    for (auto& m : part[t]) o.uniq.push_back(std::move(m));
    o.uniq_key.insert(o.uniq_key.end(), part_key[t].begin(), part_key[t].end());
  }
}

}  // namespace

extern "C" {

void* orc_create(const dbi_params* p) {
  Oracle* o = new Oracle();
  o->p = *p;
  return o;
}

void orc_destroy(void* h) { delete (Oracle*)h; }

// ProteinCache.addProtein, ids in call order, 0-based (ProteinCache.java:84-95).
int orc_add_proteins(void* h, const uint8_t* residues, const uint64_t* offsets, uint32_t n) {
  Oracle* o = (Oracle*)h;
  for (uint32_t i = 0; i < n; ++i)
    o->prot.emplace_back((const char*)residues + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
  return 0;
}

void orc_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n > 0 ? n : 1);
#else
  (void)n;
#endif
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// DBIndexer.run() loop (DBIndexer.java:600-640) + stopAddSeq (:666).
int orc_build(void* h) {
  Oracle* o = (Oracle*)h;
  const int np = (int)o->prot.size();
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  std::vector<std::vector<Rec>> part(nt);
#pragma omp parallel for schedule(static)
  for (int t = 0; t < nt; ++t) {
    const int p0 = (int)((int64_t)np * t / nt), p1 = (int)((int64_t)np * (t + 1) / nt);
    for (int i = p0; i < p1; ++i) cut_seq(o->p, o->prot[i], i, part[t]);
  }
  for (int t = 0; t < nt; ++t) o->emitted.insert(o->emitted.end(), part[t].begin(), part[t].end());
  flatten_rows(*o, o->emitted);
  o->err = expand_mods(*o);
  o->built = true;
  return o->err;
}

// store-level KAT: records injected instead of digested (DBIndexStoreSQLiteMult.main, Mult:497-524)
int orc_build_from_records(void* h, const double* mass, const uint32_t* prot, const uint32_t* off,
                           const uint16_t* len, uint64_t n) {
  Oracle* o = (Oracle*)h;
  for (uint64_t i = 0; i < n; ++i)
    o->emitted.push_back(Rec{mass[i], (int32_t)off[i], (int32_t)len[i], (int32_t)prot[i]});
  flatten_rows(*o, o->emitted);
  o->err = expand_mods(*o);
  o->built = true;
  return o->err;
}

void orc_counts(void* h, uint64_t* n_emitted, uint64_t* n_unique, uint64_t* n_entries, uint64_t* n_prot_ids) {
  Oracle* o = (Oracle*)h;
  *n_emitted = o->emitted.size();
  *n_unique = o->uniq.size();
  *n_entries = o->entries.size();
  uint64_t t = 0;
  for (auto& e : o->entries) t += o->uniq[e.base].prots.size();
  *n_prot_ids = t;
}

void orc_emitted(void* h, double* mass, uint32_t* prot, uint32_t* off, uint16_t* len) {
  Oracle* o = (Oracle*)h;
  for (size_t i = 0; i < o->emitted.size(); ++i) {
    mass[i] = o->emitted[i].mass;
    prot[i] = (uint32_t)o->emitted[i].prot;
    off[i] = (uint32_t)o->emitted[i].off;
    len[i] = (uint16_t)o->emitted[i].len;
  }
}

// Entries [begin, begin+count) in index order, same shape as dbi_fetch.
void orc_entries(void* h, uint64_t begin, uint64_t count, double* mass, uint32_t* first_prot,
                 uint32_t* first_off, uint16_t* len, uint32_t* modpat, uint64_t* prot_list_off,
                 uint32_t* prot_ids) {
  Oracle* o = (Oracle*)h;
  uint64_t w = 0;
  for (uint64_t i = 0; i < count; ++i) {
    const Entry& e = o->entries[begin + i];
    const Merged& u = o->uniq[e.base];
    if (mass) mass[i] = e.mass;
    if (first_prot) first_prot[i] = (uint32_t)u.prots[0];
    if (first_off) first_off[i] = (uint32_t)u.off;
    if (len) len[i] = (uint16_t)u.len;
    if (modpat) modpat[i] = e.modpat;
    if (prot_list_off) prot_list_off[i] = w;
    for (int32_t pid : u.prots) {
      if (prot_ids) prot_ids[w] = (uint32_t)pid;
      ++w;
    }
  }
  if (prot_list_off) prot_list_off[count] = w;
}

// DBIndexStoreSQLiteByteIndexMerge.getSequences(precMass, tol) (Merge:146-217) +
// parseAddPeptideInfo (Merge:386-481), given lo = max(0, m - tol), hi = m + tol:
// rows minKey..maxKey (Merge:170,178), inside a row stop at mass > hi (:415-417),
// skip mass < lo (:419-434).  Entries are row-ordered and mass-sorted inside a row,
// so the qualifying entries form one contiguous run of the flattened table; the
// run is returned as (begin, count) and `contiguous` is cleared if it is not one.
int orc_query(void* h, const double* lo, const double* hi, uint64_t nq, uint64_t* hit_begin,
              uint64_t* hit_count, int* contiguous) {
  Oracle* o = (Oracle*)h;
  const double f = o->p.mass_group_factor;
  const size_t n = o->entries.size();
  std::vector<int32_t> ekey(n);
  for (size_t i = 0; i < n; ++i) ekey[i] = java_int(o->entries[i].mass * f);
  int contig = 1;
#pragma omp parallel for schedule(static) reduction(&& : contig)
  for (int64_t q = 0; q < (int64_t)nq; ++q) {
    int32_t min_key = java_int(lo[q] * f);  // Merge:170
    if (min_key < 0) min_key = 0;             // Merge:173-175
    const int32_t max_key = java_int(hi[q] * f);  // Merge:178
    // SELECT ... WHERE precursor_mass_key BETWEEN minKey AND maxKey  (Merge:185-188)
    size_t r = std::lower_bound(ekey.begin(), ekey.end(), min_key) - ekey.begin();
    uint64_t first = UINT64_MAX, cnt = 0, last = 0;
    while (r < n && ekey[r] <= max_key) {
      const int32_t row = ekey[r];
      for (; r < n && ekey[r] == row; ++r) {  // walk one row's data blob
        const double m = o->entries[r].mass;
        if (m > hi[q]) {  // "since it's sorted" break (Merge:415-417): skip rest of row
          while (r < n && ekey[r] == row) ++r;
          break;
        }
        if (m < lo[q]) continue;  // Merge:419-434
        if (first == UINT64_MAX) first = r;
        else if (r != last + 1) contig = 0;
        last = r;
        ++cnt;
      }
    }
    if (cnt == 0) {
      // position where hits would be: number of entries with mass < lo
      first = std::lower_bound(o->entries.begin(), o->entries.end(), lo[q],
                               [](const Entry& e, double v) { return e.mass < v; }) -
              o->entries.begin();
    }
    hit_begin[q] = first;
    hit_count[q] = cnt;
  }
  if (contiguous) *contiguous = contig;
  return 0;
}

// What getSequences hands back for a batch of ranges: parseAddPeptideInfo (Merge:386-481) for every
// hit of every query -- exact mass, the first occurrence (protein, offset, length), the peptide
// string cut from the protein (ProteinCache.getPeptideSequence, ProteinCache.java:112-127), the
// flanking residues (Util.getResidues, Merge:456-458) and the protein-id list.  Two-pass: call with
// mass == NULL for the sizes.  hit_off[nq + 1], seq_off / prot_list_off [n_hits + 1].
void orc_get_residues(const uint8_t* prot_seq, uint64_t prot_len, uint32_t seq_off, uint32_t seq_len, char* left3,
                      char* right3);
int orc_query_hits(void* h, const double* lo, const double* hi, uint64_t nq, uint64_t* hit_off, double* mass,
                   uint32_t* first_prot, uint32_t* first_off, uint16_t* len, uint32_t* modpat, uint64_t* seq_off,
                   uint8_t* seq, char* flanks /* 6 per hit */, uint64_t* prot_list_off, uint32_t* prot_ids,
                   uint64_t* n_hits, uint64_t* n_seq_bytes, uint64_t* n_prot_ids) {
  Oracle* o = (Oracle*)h;
  std::vector<uint64_t> b(nq), c(nq);
  int contig = 1;
  orc_query(h, lo, hi, nq, b.data(), c.data(), &contig);
  if (!contig) return -1;
  // sizes per query, then their prefix sums (the fill below runs on every host thread)
  std::vector<uint64_t> qs(nq + 1, 0), qp(nq + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < (int64_t)nq; ++q) {
    uint64_t sb = 0, pi = 0;
    for (uint64_t i = 0; i < c[q]; ++i) {
      const Merged& u = o->uniq[o->entries[b[q] + i].base];
      sb += (uint64_t)u.len;
      pi += u.prots.size();
    }
    qs[q + 1] = sb;
    qp[q + 1] = pi;
  }
  hit_off[0] = 0;
  for (uint64_t q = 0; q < nq; ++q) {
    hit_off[q + 1] = hit_off[q] + c[q];
    qs[q + 1] += qs[q];
    qp[q + 1] += qp[q];
  }
  const uint64_t H = hit_off[nq], SB = qs[nq], PI = qp[nq];
  if (mass) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t q = 0; q < (int64_t)nq; ++q) {
      uint64_t hh = hit_off[q], sb = qs[q], pi = qp[q];
      for (uint64_t i = 0; i < c[q]; ++i, ++hh) {
        const Entry& e = o->entries[b[q] + i];
        const Merged& u = o->uniq[e.base];
        mass[hh] = e.mass;
        first_prot[hh] = (uint32_t)u.prots[0];
        first_off[hh] = (uint32_t)u.off;
        len[hh] = (uint16_t)u.len;
        modpat[hh] = e.modpat;
        seq_off[hh] = sb;
        prot_list_off[hh] = pi;
        const std::string& ps = o->prot[u.prots[0]];
        std::memcpy(seq + sb, ps.data() + u.off, (size_t)u.len);  // protSeq.substring(off, off + len)
        if (flanks)
          orc_get_residues((const uint8_t*)ps.data(), ps.size(), (uint32_t)u.off, (uint32_t)u.len, flanks + 6 * hh,
                           flanks + 6 * hh + 3);
        for (size_t k = 0; k < u.prots.size(); ++k) prot_ids[pi + k] = (uint32_t)u.prots[k];
        sb += (uint64_t)u.len;
        pi += u.prots.size();
      }
    }
    seq_off[H] = SB;
    prot_list_off[H] = PI;
  }
  *n_hits = H;
  *n_seq_bytes = SB;
  *n_prot_ids = PI;
  return 0;
}

// IndexUtil.calculateMass(sequence, isH2OPlusProtonAdded) (util/IndexUtil.java:197-208)
double orc_calculate_mass(void* h, const uint8_t* seq, uint64_t len) {
  Oracle* o = (Oracle*)h;
  double mass = 0;
  if (o->p.add_h2o_proton) mass += o->p.h2o_proton;
  mass += o->p.cterm;
  mass += o->p.nterm;
  for (uint64_t i = 0; i < len; ++i) mass += o->p.residue_mass[seq[i]];
  return mass;
}

// IndexUtil.getToleranceInDalton (util/IndexUtil.java:238-240), Constants.ONE_MILLION
double orc_tolerance_in_dalton(double actual_mass, double ppm) {
  return actual_mass * (1 - 1 / (ppm / 1000000.0 + 1));
}

// DBIndexer.getSequencesUsingPPMTolerance (DBIndexer.java:787-844) for one precursor mass: the
// Dalton query, then exact-mass probes at upperBound, upperBound + PRECISION, ... while the PPM
// window of the probe still reaches down to the precursor mass and the probe finds something.
// Writes the entry indices of the result in list order (first query, then new ones per probe;
// `sequences.contains` is identity of the index entry here); returns how many (<= cap kept).
uint64_t orc_query_ppm(void* h, double precursor_mass, double ppm, uint64_t* out_idx, uint64_t cap,
                       uint64_t* n_probes) {
  std::vector<uint64_t> seqs;
  auto get = [&](double m, double tol, std::vector<uint64_t>& dst) {  // indexStore.getSequences(m, tol)
    double lo = m - tol, hi = m + tol;
    if (lo < 0.0) lo = 0.0;  // Mult:324-329
    uint64_t b = 0, c = 0;
    int contig = 1;
    orc_query(h, &lo, &hi, 1, &b, &c, &contig);
    for (uint64_t i = 0; i < c; ++i) dst.push_back(b + i);
  };
  const double tol = orc_tolerance_in_dalton(precursor_mass, ppm);  // :790
  get(precursor_mass, tol, seqs);                                    // :799
  double upper = precursor_mass + tol;                               // :808
  uint64_t probes = 0;
  while (true) {
    const double tol2 = orc_tolerance_in_dalton(upper, ppm);  // :812
    const double lower_of_upper = upper - tol2;               // :813
    if (lower_of_upper < precursor_mass) {                    // :814
      std::vector<uint64_t> s2;
      get(upper, 0.0, s2);                                    // :815
      ++probes;
      if (s2.empty()) break;                                  // :816-817
      for (uint64_t e : s2)
        if (std::find(seqs.begin(), seqs.end(), e) == seqs.end()) seqs.push_back(e);  // :821-826
    } else {
      break;
    }
    const double next = upper + 1e-6;  // Constants.PRECISION, :833
    if (next == upper) break;
    upper = next;
  }
  for (uint64_t i = 0; i < seqs.size() && i < cap; ++i) out_idx[i] = seqs[i];
  if (n_probes) *n_probes = probes;
  return seqs.size();
}

// Interval.massRangeToInterval (Interval.java:27-38) + MergeIntervals.mergeIntervals
// (MergeIntervals.java:16-46).  In/out: n (mass, tol) pairs -> merged [lo, hi] list.
uint64_t orc_merge_intervals(const double* mass, const double* tol, uint64_t n, double* out_lo, double* out_hi) {
  std::vector<std::pair<double, double>> iv(n);
  for (uint64_t i = 0; i < n; ++i) {
    double lo = mass[i] - tol[i];
    if (lo < 0.0) lo = 0.0;
    iv[i] = {lo, mass[i] + tol[i]};
  }
  if (n < 2) {
    for (uint64_t i = 0; i < n; ++i) { out_lo[i] = iv[i].first; out_hi[i] = iv[i].second; }
    return n;
  }
  std::stable_sort(iv.begin(), iv.end(), [](auto& a, auto& b) { return a.first < b.first; });
  uint64_t w = 0;
  double start = iv[0].first, end = iv[0].second;
  for (uint64_t i = 1; i < n; ++i) {
    if (end >= iv[i].first) {
      end = std::max(end, iv[i].second);
    } else {
      out_lo[w] = start; out_hi[w] = end; ++w;
      start = iv[i].first; end = iv[i].second;
    }
  }
  out_lo[w] = start; out_hi[w] = end; ++w;
  return w;
}

// Util.getResidues (Util.java:130-162) including the right-flank off-by-one (Q8):
// left[3], right[3], '-' padded.
void orc_get_residues(const uint8_t* prot_seq, uint64_t prot_len, uint32_t seq_off, uint32_t seq_len,
                      char* left3, char* right3) {
  const int MAXR = 3;  // Constants.MAX_INDEX_RESIDUE_LEN
  const int64_t pl = (int64_t)prot_len;
  const int64_t left_i = seq_off >= (uint32_t)MAXR ? seq_off - MAXR : 0;
  const int64_t left_len = std::min<int64_t>(MAXR, seq_off);
  std::string l((const char*)prot_seq + left_i, (size_t)left_len);
  const int64_t end = (int64_t)seq_off + seq_len;
  const int64_t right_len = std::min<int64_t>(MAXR, pl - end - 1);
  std::string r;
  if (end < pl && right_len > 0) r.assign((const char*)prot_seq + end, (size_t)right_len);
  while ((int)l.size() < MAXR) l.insert(l.begin(), '-');
  while ((int)r.size() < MAXR) r.push_back('-');
  std::memcpy(left3, l.data(), 3);
  std::memcpy(right3, r.data(), 3);
}

}  // extern "C"
