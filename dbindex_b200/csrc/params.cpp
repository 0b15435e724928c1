// params.cpp -- host-only parameter helpers of the C ABI (no CUDA calls).
//
// The tables below are the written-down contract for the two classes that live in
// the un-vendored edu.scripps.yates:utilities jar: AssignMass (residue
// masses, H2O_PROTON) and Enzyme.  They are defaults only: every value is an INPUT
// of dbi_create(), so a Java host overwrites them from the live classes.
#include <cstring>

#include "../../include/dbindex_gpu.h"

namespace {

struct AA {
  char c;
  double mono;
  double avg;
};

// standard residue masses (monoisotopic / average), Da
const AA kAA[] = {
    {'G', 57.02146372, 57.0513},   {'A', 71.03711378, 71.0779},   {'S', 87.03202840, 87.0773},
    {'P', 97.05276384, 97.1152},   {'V', 99.06841390, 99.1311},   {'T', 101.04767846, 101.1039},
    {'C', 103.00918447, 103.1429}, {'L', 113.08406396, 113.1576}, {'I', 113.08406396, 113.1576},
    {'N', 114.04292744, 114.1026}, {'D', 115.02694302, 115.0874}, {'Q', 128.05857750, 128.1292},
    {'K', 128.09496300, 128.1723}, {'E', 129.04259308, 129.1140}, {'M', 131.04048459, 131.1961},
    {'H', 137.05891186, 137.1393}, {'F', 147.06841390, 147.1739}, {'R', 156.10111102, 156.1857},
    {'Y', 163.06332852, 163.1733}, {'W', 186.07931294, 186.2099}, {'U', 150.95363559, 150.0379},
    {'O', 237.14772686, 237.2982},
};

}  // namespace

extern "C" {

void dbi_default_params(dbi_params* p, int mono) {
  std::memset(p, 0, sizeof(*p));
  p->abi_version = DBI_ABI_VERSION;
  p->device = 0;
  for (const AA& a : kAA) p->residue_mass[(unsigned char)a.c] = mono ? a.mono : a.avg;
  // ambiguity codes: X and J as Leu/Ile, B = mean(N, D), Z = mean(Q, E)
  p->residue_mass['X'] = p->residue_mass['L'];
  p->residue_mass['J'] = p->residue_mass['L'];
  p->residue_mass['B'] = (p->residue_mass['N'] + p->residue_mass['D']) / 2;
  p->residue_mass['Z'] = (p->residue_mass['Q'] + p->residue_mass['E']) / 2;
  // AssignMass.H2O_PROTON = water + proton
  p->h2o_proton = (mono ? 18.0105646837 : 18.01528) + 1.00727646688;
  p->nterm = 0;
  p->cterm = 0;
  p->add_h2o_proton = 1;  // dbindex.properties:22, SearchParamReader.java:706
  dbi_params_set_enzyme(p, "KR", "");  // dbindex.properties:11-12
  p->max_missed = 2;
  p->semi = 0;  // dbindex.properties:26
  p->min_len = DBI_MIN_PEP_LENGTH;
  p->min_mass = 600.0;
  p->max_mass = 6000.0;
  p->mass_group_factor = DBI_MASS_GROUP_FACTOR;
  p->n_mods = 0;
  p->max_mods_per_peptide = 0;
}

void dbi_params_add_static_mod(dbi_params* p, uint8_t residue, double delta) {
  if (delta > 0) p->residue_mass[residue] += delta;  // AssignMassToStaticParam.java:9-14
}

void dbi_params_set_enzyme(dbi_params* p, const char* residues, const char* nocut) {
  std::memset(p->is_enzyme, 0, sizeof(p->is_enzyme));
  std::memset(p->is_nocut, 0, sizeof(p->is_nocut));
  if (residues)
    for (const char* c = residues; *c; ++c) p->is_enzyme[(unsigned char)*c] = 1;
  if (nocut)
    for (const char* c = nocut; *c; ++c) p->is_nocut[(unsigned char)*c] = 1;
}

void dbi_abi_sizes(uint64_t* sizeof_params, uint64_t* sizeof_stats) {
  if (sizeof_params) *sizeof_params = sizeof(dbi_params);
  if (sizeof_stats) *sizeof_stats = sizeof(dbi_stats);
}

int dbi_params_add_diff_mod(dbi_params* p, const char* residues, double delta) {
  if (delta == 0) return DBI_OK;  // "if (massShift != 0)", SearchParamReader.java:646
  for (const char* c = residues; c && *c; ++c) {
    if (p->n_mods >= DBI_MAX_MODS) return DBI_ERANGE;
    p->mods[p->n_mods].residue = (uint8_t)*c;
    p->mods[p->n_mods].delta = delta;
    p->n_mods++;
  }
  return DBI_OK;
}

}  // extern "C"
