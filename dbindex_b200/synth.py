"""Seeded synthetic inputs for the BASELINE.json configs.

There is no network for real proteomes, so every test and benchmark runs on synthetic
FASTA: amino acids i.i.d. from the Swiss-Prot composition, protein lengths log-normal
(median ~420, sigma 0.6, min 50, cap 36000), seed = 20240601 + cfg.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

AA = "ARNDCQEGHILKMFPSTWYV"
# Swiss-Prot amino-acid composition, percent
AA_FREQ = np.array([8.25, 5.53, 4.06, 5.46, 1.38, 3.93, 6.72, 7.07, 2.27, 5.91, 9.65, 5.80, 2.41, 3.86, 4.74,
                    6.65, 5.36, 1.10, 2.92, 6.85])
BASE_SEED = 20240601

CONFIG_PROTEINS = {1: 20_000, 2: 20_000, 3: 200_000, 4: 8_000_000}


def synth_proteome(n_proteins: int, seed: int, median_len: float = 420.0, sigma: float = 0.6,
                   min_len: int = 50, max_len: int = 36000) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (residues uint8[R], offsets uint64[n+1]); protein i = residues[offsets[i]:offsets[i+1]]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.exp(rng.normal(np.log(median_len), sigma, size=n_proteins))
    lens = np.clip(np.rint(lens), min_len, max_len).astype(np.int64)
    offsets = np.zeros(n_proteins + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    codes = np.frombuffer(AA.encode(), dtype=np.uint8)
    cdf = np.cumsum(AA_FREQ / AA_FREQ.sum())
    residues = np.empty(total, dtype=np.uint8)
    chunk = 1 << 26
    for s in range(0, total, chunk):
        e = min(total, s + chunk)
        u = rng.random(e - s)
        residues[s:e] = codes[np.minimum(np.searchsorted(cdf, u, side="right"), len(codes) - 1)]
    return residues, offsets


def config_proteome(cfg: int, n_proteins: int | None = None) -> Tuple[np.ndarray, np.ndarray]:
    """The synthetic FASTA of BASELINE.json configs[cfg-1]; cfg 2 shares cfg 1's FASTA."""
    seed_cfg = 1 if cfg == 2 else (4 if cfg == 5 else cfg)
    n = n_proteins if n_proteins is not None else CONFIG_PROTEINS[seed_cfg]
    return synth_proteome(n, BASE_SEED + seed_cfg)


def deflines(n: int, start: int = 0) -> List[str]:
    return [f"sp|S{i:07d}|SYN_{i}" for i in range(start, start + n)]


def write_fasta(path: str, residues: np.ndarray, offsets: np.ndarray, width: int = 60):
    with open(path, "w") as f:
        for i in range(len(offsets) - 1):
            seq = residues[int(offsets[i]):int(offsets[i + 1])].tobytes().decode()
            f.write(f">sp|S{i:07d}|SYN_{i}\n")
            for k in range(0, len(seq), width):
                f.write(seq[k:k + width] + "\n")


def synth_queries(index_masses: np.ndarray, nq: int, seed: int, ppm: float = 10.0, min_mass: float = 600.0,
                  max_mass: float = 6000.0, da_fraction: float = 0.0, da_tol: float = 3.0):
    """Precursor queries: half are indexed masses jittered by U(-5, 5) ppm, half
    are U(min, max) decoys; tolerance = getToleranceInDalton(m, ppm) or a fixed +-da_tol window for
    the last `da_fraction` of them.  Returns (mass, tol, lo, hi) with lo = max(0, m - tol)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_hit = nq // 2
    if len(index_masses):
        picks = index_masses[rng.integers(0, len(index_masses), size=n_hit)]
    else:
        picks = rng.uniform(min_mass, max_mass, size=n_hit)
    hits = picks * (1.0 + rng.uniform(-5e-6, 5e-6, size=n_hit))
    decoys = rng.uniform(min_mass, max_mass, size=nq - n_hit)
    mass = np.concatenate([hits, decoys])
    rng.shuffle(mass)
    tol = mass * (1 - 1 / (ppm / 1000000.0 + 1))  # IndexUtil.getToleranceInDalton
    n_da = int(nq * da_fraction)
    if n_da:
        tol[nq - n_da:] = da_tol
    lo = np.maximum(mass - tol, 0.0)
    hi = mass + tol
    return mass, tol, lo, hi
