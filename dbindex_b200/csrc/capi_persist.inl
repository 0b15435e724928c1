// capi_persist.inl -- index persistence / resume; included at the end of capi.cu.
//
// The reference keeps its index on disk under <fasta>_<md5(params)> (util/IndexUtil.java:270-324) and
// skips indexing when it finds one (DBIndexer.java:522-531, "Found existing index, skipping indexing").
// Its payload is a SQLite table of per-key blobs (DBIndexStoreSQLiteByte.java:586-587, merged rows
// DBIndexStoreSQLiteByteIndexMerge.java:696-716); ours is the finished device arrays, written once and
// mapped straight back into HBM: the proteins (offsets + residues, the ProteinCache), the unique tables
// (first occurrence + CSR protein lists) and, with differential mods, the entry arrays.  The file name
// is the host's business (dbindex_b200/indexer.py mirrors createFullIndexFileName).

#include <cstdio>

namespace {

struct IdxHeader {
  char magic[8];  // "DBIGPU2\0"
  uint32_t abi_version;
  uint32_t header_bytes;
  uint64_t n_proteins, n_residues, n_emitted, n_unique, n_entries;
  uint32_t has_entries;  // entry arrays follow (differential mods)
  uint32_t sizeof_params;
  dbi_params params;     // as given to dbi_create (device / diagnostic switches are not compared)
};

struct File {
  FILE* f = nullptr;
  ~File() {
    if (f) std::fclose(f);
  }
};

// the parameters that shape an index: everything but the device and the diagnostic switches
bool same_index_params(const dbi_params& a, const dbi_params& b) {
  dbi_params x = a, y = b;
  x.device = y.device = 0;
  x.keep_emitted = y.keep_emitted = 0;
  x.profile = y.profile = 0;
  std::memset(x.reserved, 0, sizeof(x.reserved));
  std::memset(y.reserved, 0, sizeof(y.reserved));
  x._pad_filters = y._pad_filters = 0;
  for (int i = 0; i < DBI_MAX_MODS; ++i) {
    std::memset(x.mods[i]._pad, 0, sizeof(x.mods[i]._pad));
    std::memset(y.mods[i]._pad, 0, sizeof(y.mods[i]._pad));
    if (i >= x.n_mods) std::memset(&x.mods[i], 0, sizeof(x.mods[i]));
    if (i >= y.n_mods) std::memset(&y.mods[i], 0, sizeof(y.mods[i]));
  }
  // compare field by field through the bytes of the normalised copies (padding of the struct itself
  // is zero in both: dbi_default_params memsets, and callers copy whole structs)
  return std::memcmp(&x, &y, sizeof(dbi_params)) == 0;
}

constexpr size_t kIoChunk = 64u << 20;

void write_dev(dbi_handle* h, FILE* f, const void* d, uint64_t bytes, std::vector<uint8_t>& buf) {
  const uint8_t* p = (const uint8_t*)d;
  for (uint64_t o = 0; o < bytes; o += kIoChunk) {
    const size_t n = (size_t)std::min<uint64_t>(kIoChunk, bytes - o);
    DBI_CUDA(cudaMemcpyAsync(buf.data(), p + o, n, cudaMemcpyDeviceToHost, h->stream));
    DBI_CUDA(cudaStreamSynchronize(h->stream));
    if (std::fwrite(buf.data(), 1, n, f) != n) throw std::runtime_error("write failed");
  }
  static const uint8_t zeros[8] = {0};
  if (bytes & 7)
    if (std::fwrite(zeros, 1, 8 - (bytes & 7), f) != 8 - (bytes & 7)) throw std::runtime_error("write failed");
}

void read_dev(dbi_handle* h, FILE* f, void* d, uint64_t bytes, std::vector<uint8_t>& buf) {
  uint8_t* p = (uint8_t*)d;
  for (uint64_t o = 0; o < bytes; o += kIoChunk) {
    const size_t n = (size_t)std::min<uint64_t>(kIoChunk, bytes - o);
    if (std::fread(buf.data(), 1, n, f) != n) throw std::runtime_error("truncated index file");
    DBI_CUDA(cudaMemcpyAsync(p + o, buf.data(), n, cudaMemcpyHostToDevice, h->stream));
    DBI_CUDA(cudaStreamSynchronize(h->stream));
  }
  if (bytes & 7) {
    uint8_t pad[8];
    if (std::fread(pad, 1, 8 - (bytes & 7), f) != 8 - (bytes & 7)) throw std::runtime_error("truncated index file");
  }
}

}  // namespace

extern "C" {

int dbi_save(dbi_handle* h, const char* path) {
  DBI_API_BEGIN(h)
  if (!h->built) {
    set_error("Indexer is not initialized");
    return DBI_ENOTINIT;
  }
  if (!path) {
    set_error("null argument");
    return DBI_EINVAL;
  }
  if (h->mg_world > 1) {
    set_error("dbi_save stores a single-GPU index; a sharded index is rebuilt from its FASTA shards");
    return DBI_EINVAL;
  }
  const std::string tmp = std::string(path) + ".tmp";
  try {
    File fl;
    fl.f = std::fopen(tmp.c_str(), "wb");
    if (!fl.f) {
      set_error("cannot create %s", tmp.c_str());
      return DBI_EINVAL;
    }
    IdxHeader hd;
    std::memset(&hd, 0, sizeof(hd));
    std::memcpy(hd.magic, "DBIGPU2", 8);
    hd.abi_version = DBI_ABI_VERSION;
    hd.header_bytes = (uint32_t)sizeof(hd);
    hd.n_proteins = h->h_off.size() - 1;
    hd.n_residues = h->n_res;
    hd.n_emitted = h->n_emitted;
    hd.n_unique = h->n_unique;
    hd.n_entries = h->n_entries;
    hd.has_entries = h->e_mass.p ? 1u : 0u;
    hd.sizeof_params = (uint32_t)sizeof(dbi_params);
    hd.params = h->p;
    if (std::fwrite(&hd, sizeof(hd), 1, fl.f) != 1) throw std::runtime_error("write failed");
    if (std::fwrite(h->h_off.data(), 8, h->h_off.size(), fl.f) != h->h_off.size()) throw std::runtime_error("write failed");
    std::vector<uint8_t> buf(kIoChunk);
    const uint64_t U = h->n_unique, N = h->n_emitted, V = h->n_entries;
    write_dev(h, fl.f, h->d_raw.p, h->n_res, buf);
    write_dev(h, fl.f, h->u_mass.p, U * 8, buf);
    write_dev(h, fl.f, h->u_gpos.p, U * 4, buf);
    write_dev(h, fl.f, h->u_prot.p, U * 4, buf);
    write_dev(h, fl.f, h->u_len.p, U * 2, buf);
    write_dev(h, fl.f, h->u_plo.p, (U + 1) * 8, buf);
    write_dev(h, fl.f, h->plist.p, N * 4, buf);
    if (hd.has_entries) {
      write_dev(h, fl.f, h->e_mass.p, V * 8, buf);
      write_dev(h, fl.f, h->e_base.p, V * 4, buf);
      write_dev(h, fl.f, h->e_pat.p, V * 4, buf);
    }
    if (std::fflush(fl.f) != 0) throw std::runtime_error("write failed");
  } catch (const std::runtime_error& e) {
    std::remove(tmp.c_str());
    set_error("dbi_save(%s): %s", path, e.what());
    return DBI_EINVAL;
  }
  if (std::rename(tmp.c_str(), path) != 0) {  // the index appears atomically: a reader never sees half a file
    std::remove(tmp.c_str());
    set_error("cannot rename %s to %s", tmp.c_str(), path);
    return DBI_EINVAL;
  }
  return DBI_OK;
  DBI_API_END
}

int dbi_load(dbi_handle* h, const char* path) {
  DBI_API_BEGIN(h)
  if (h->built || h->h_off.size() > 1) {
    set_error("dbi_load needs a fresh handle (no proteins, no index)");
    return DBI_EALREADY;
  }
  if (!path) {
    set_error("null argument");
    return DBI_EINVAL;
  }
  try {
    File fl;
    fl.f = std::fopen(path, "rb");
    if (!fl.f) {
      set_error("cannot open %s", path);
      return DBI_EINVAL;
    }
    IdxHeader hd;
    if (std::fread(&hd, sizeof(hd), 1, fl.f) != 1 || std::memcmp(hd.magic, "DBIGPU2", 8) != 0 ||
        hd.header_bytes != sizeof(hd) || hd.sizeof_params != sizeof(dbi_params) || hd.abi_version != DBI_ABI_VERSION) {
      set_error("%s is not an index of this library version", path);
      return DBI_EINVAL;
    }
    if (!same_index_params(hd.params, h->p)) {
      set_error("%s was built with other search parameters", path);
      return DBI_EINVAL;
    }
    const uint64_t P = hd.n_proteins, R = hd.n_residues, U = hd.n_unique, N = hd.n_emitted, V = hd.n_entries;
    if (R + P + 1 + 4096 >= (1ull << 32) || P >= (1ull << 31) || U > N || N >= (1ull << 32) || V >= (1ull << 32)) {
      set_error("%s: implausible counts", path);
      return DBI_EINVAL;
    }
    h->h_off.assign(P + 1, 0);
    if (std::fread(h->h_off.data(), 8, P + 1, fl.f) != P + 1 || h->h_off[0] != 0 || h->h_off[P] != R)
      throw std::runtime_error("truncated index file");
    std::vector<uint8_t> buf(kIoChunk);
    h->d_raw.alloc(std::max<uint64_t>(R, 16), h->arena);
    read_dev(h, fl.f, h->d_raw.p, R, buf);
    h->n_res = R;
    h->st.n_proteins = P;
    h->st.n_residues = R;
    ensure_uploaded(h);
    const uint32_t zero = 0;
    DBI_CUDA(cudaMemcpyAsync(h->d_err.p, &zero, 4, cudaMemcpyHostToDevice, h->stream));
    pack_residues(h);  // residues + protein starts as the build lays them out (u_gpos refers to this layout)
    h->u_mass.alloc(std::max<uint64_t>(U, 1) * 8, h->arena);
    h->u_gpos.alloc(std::max<uint64_t>(U, 1) * 4, h->arena);
    h->u_prot.alloc(std::max<uint64_t>(U, 1) * 4, h->arena);
    h->u_len.alloc(std::max<uint64_t>(U, 1) * 2, h->arena);
    h->u_plo.alloc((U + 1) * 8, h->arena);
    h->plist.alloc(std::max<uint64_t>(N, 1) * 4, h->arena);
    read_dev(h, fl.f, h->u_mass.p, U * 8, buf);
    read_dev(h, fl.f, h->u_gpos.p, U * 4, buf);
    read_dev(h, fl.f, h->u_prot.p, U * 4, buf);
    read_dev(h, fl.f, h->u_len.p, U * 2, buf);
    read_dev(h, fl.f, h->u_plo.p, (U + 1) * 8, buf);
    read_dev(h, fl.f, h->plist.p, N * 4, buf);
    if (hd.has_entries) {
      h->e_mass.alloc(std::max<uint64_t>(V, 1) * 8, h->arena);
      h->e_base.alloc(std::max<uint64_t>(V, 1) * 4, h->arena);
      h->e_pat.alloc(std::max<uint64_t>(V, 1) * 4, h->arena);
      read_dev(h, fl.f, h->e_mass.p, V * 8, buf);
      read_dev(h, fl.f, h->e_base.p, V * 4, buf);
      read_dev(h, fl.f, h->e_pat.p, V * 4, buf);
    } else if (V != U) {
      throw std::runtime_error("entry count does not match the unique table");
    }
    if (int rc = check_err_bits(read_err(h))) {
      free_index(h);
      return rc;
    }
    h->n_emitted = N;
    h->n_unique = U;
    h->n_entries = V;
    h->st.n_emitted = N;
    h->st.n_unique = U;
    h->st.n_entries = V;
    h->ent_base_off = 0;
    h->built = true;
    DBI_CUDA(cudaStreamSynchronize(h->stream));
    finish_stats(h);
    return DBI_OK;
  } catch (const std::runtime_error& e) {
    free_index(h);
    h->h_off.assign(1, 0);
    h->n_res = 0;
    set_error("dbi_load(%s): %s", path, e.what());
    return DBI_EINVAL;
  }
  DBI_API_END
}

}  // extern "C"
