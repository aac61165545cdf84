#!/usr/bin/env python3
"""Tiny workload that touches every kernel once (for compute-sanitizer racecheck, which is slow)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mchap_b200 import DenovoMCMC  # noqa: E402
from mchap_b200.calling import CallingMCMC  # noqa: E402
from mchap_b200.calling import exact  # noqa: E402
from mchap_b200.synth import synth_haplotype_panel, synth_items  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "assemble"):
    b = synth_items(3, ploidy=4, n_pos=6, depth=12, seed=1)
    reads, counts = [b.item(i)[0] for i in range(3)], [b.item(i)[1] for i in range(3)]
    DenovoMCMC(ploidy=4, n_alleles=[2] * 6, steps=25, chains=2, temperatures=(0.3, 1.0), inbreeding=0.1,
               random_seed=1).fit_posterior_batch(reads, counts, burn=5)
    b = synth_items(2, ploidy=4, n_pos=8, depth=120, seed=2)   # more than 32 unique reads
    DenovoMCMC(ploidy=4, n_alleles=[2] * 8, steps=10, chains=1, random_seed=2).fit_batch(
        [b.item(i)[0] for i in range(2)], [b.item(i)[1] for i in range(2)])
    b = synth_items(2, ploidy=8, n_pos=16, depth=100, seed=3)  # resident-slot swap
    DenovoMCMC(ploidy=8, n_alleles=[2] * 16, steps=3, chains=1, temperatures=(0.01, 0.1, 0.5, 1.0),
               random_seed=3).fit_batch([b.item(i)[0] for i in range(2)], [b.item(i)[1] for i in range(2)])
    rng = np.random.default_rng(9)   # 300 distinct reads: the 16-chunk kernel, Rt in global memory
    r = np.stack([rng.random((300, 5)) * 0.5 + 0.5, np.zeros((300, 5))], axis=-1)
    r[..., 1] = (1 - r[..., 0]) / 3
    DenovoMCMC(ploidy=4, n_alleles=[2] * 5, steps=4, chains=1, temperatures=(0.05, 1.0), random_seed=4).fit_batch([r])
if which in ("all", "assemble", "multiallelic"):
    rng = np.random.default_rng(5)
    reads, counts, nalls = [], [], []
    for i in range(4):   # ragged allele counts: the serial multi-allelic base step, gaps, initial states
        N, A = int(rng.integers(2, 7)), 4
        na = rng.integers(2, A + 1, size=N).astype(np.int8)
        U = int(rng.integers(3, 30))
        r = rng.random((U, N, A)) + 0.05
        for j in range(N):
            r[:, j, na[j]:] = 0
        r /= r.sum(axis=-1, keepdims=True)
        r[rng.random((U, N)) < 0.25] = np.nan
        reads.append(r)
        counts.append(rng.integers(1, 5, size=U))
        nalls.append(na)
    DenovoMCMC(ploidy=4, n_alleles=None, inbreeding=0.05, steps=40, chains=2, temperatures=(0.5, 1.0),
               random_seed=1).fit_batch(reads, counts, n_alleles_list=nalls)
if which in ("all", "call"):
    batch, panels, _ = synth_haplotype_panel(3, 8, 6, 4, depth=10, seed=4)
    reads, counts = [batch.item(i)[0] for i in range(3)], [batch.item(i)[1] for i in range(3)]
    for st in ("Gibbs", "Metropolis-Hastings"):
        CallingMCMC(ploidy=4, haplotypes=panels[0], steps=30, chains=2, random_seed=5, step_type=st,
                    prior=(0.1, None)).fit_posterior_batch(reads, counts, burn=5, haplotypes_list=list(panels))
    exact.posterior_mode_batch(reads, 4, list(panels), counts, [(0.1, None)] * 3)
print("workload done")
