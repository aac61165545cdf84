#!/usr/bin/env python3
"""How far is the float32 screened log-likelihood of a bi-allelic flip from the exact one?

CPU study (numpy emulation of the kernel's screening arithmetic, mutation_compound_step tier 1):
for random states of synthetic configs[1] items, every single-position flip is evaluated
  * exactly (float64, the reference's operation order), and
  * as the kernel screens it: float32 shadow rows, rc = sum of the rows, rt = float32 allele ratio,
    rp = fma(q_h, rt - 1, rc), sum of log(rp) * count accumulated in float32
and the absolute difference is compared with the margin the kernel uses (2 + 2e-3 * sum(counts)).
The fast-math __logf adds at most 2^-21.41 absolute (x in [0.5, 2]) or 3 ulp per read on top (CUDA
C programming guide, intrinsic error table); that term is reported separately as an upper bound.

    python profiles/screening_error_study.py [n_items]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mchap_b200.synth import synth_items  # noqa: E402

f32 = np.float32


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def study(n_items=400, ploidy=4, n_pos=8, depth=40, seed=0, flips=3):
    rng = np.random.default_rng(seed)
    batch = synth_items(n_items, ploidy=ploidy, n_pos=n_pos, depth=depth, seed=seed)
    worst = 0.0
    worst_rel_margin = 0.0
    errs = []
    logf_bound_max = 0.0
    n_insane = 0
    for i in range(n_items):
        reads, counts = batch.item(i)
        U = len(reads)
        if U == 0:
            continue
        R = np.where(np.isnan(reads), 1.0, reads)                      # gaps count as 1 (likelihood.py:55-58)
        g = batch.haplotypes[i].copy()
        for _ in range(flips):                                          # a state near the truth
            g[rng.integers(ploidy), rng.integers(n_pos)] ^= 1
        # cached product rows q[h][r] (compute_row) and their float32 shadows
        q = np.ones((ploidy, U))
        for j in range(n_pos):
            q *= R[:, j, :][np.arange(U)[None, :], g[:, j][:, None]]
        q = q / ploidy
        q32 = q.astype(f32)
        rc = np.zeros(U, dtype=f32)
        for h in range(ploidy):
            rc = (rc + q32[h]).astype(f32)
        c32 = counts.astype(f32)
        margin = 2.0 + 2e-3 * counts.sum()
        for h in range(ploidy):
            for j in range(n_pos):
                cur = g[h, j]
                # exact proposal (reference order: product over positions, / ploidy, sum over haplotypes, log)
                g2 = g.copy()
                g2[h, j] = cur ^ 1
                rowp = np.ones(U)
                for jj in range(n_pos):
                    rowp = rowp * R[:, jj, g2[h, jj]]
                rowp = rowp / ploidy
                rp_exact = np.zeros(U)
                for hh in range(ploidy):
                    rp_exact = rp_exact + (rowp if hh == h else q[hh])
                llk_exact = float(np.sum(np.log(rp_exact) * counts))
                # screened
                rt = (R[:, j, cur ^ 1] / R[:, j, cur]).astype(f32)
                rp = fma32(q32[h], (rt - f32(1.0)).astype(f32), rc)
                sane = bool(np.all((rp > f32(1e-4) * rc) & (rp > f32(1e-30)) & (rp < f32(1e30))))
                if not sane:
                    n_insane += 1
                    continue
                lg = np.log(rp.astype(np.float64)).astype(f32)
                acc = f32(0.0)
                for r in range(U):
                    acc = fma32(np.array(lg[r]), np.array(c32[r]), np.array(acc))
                err = abs(float(acc) - llk_exact)
                errs.append(err)
                worst = max(worst, err)
                worst_rel_margin = max(worst_rel_margin, err / margin)
                # upper bound of what __logf can add: 2^-21.41 abs near 1, else 3 ulp of the result
                ulp = np.spacing(np.abs(lg).astype(f32)).astype(np.float64)
                logf_bound = float(np.sum(np.where((rp >= 0.5) & (rp <= 2.0), 2.0 ** -21.41, 3 * ulp) * counts))
                logf_bound_max = max(logf_bound_max, logf_bound)
    errs = np.array(errs)
    return {
        "items": n_items, "proposals": int(len(errs)), "not_sane_skipped": n_insane,
        "max_abs_error": worst, "p99_abs_error": float(np.quantile(errs, 0.99)), "median_abs_error": float(np.median(errs)),
        "max_error_over_margin": worst_rel_margin, "max_logf_intrinsic_bound": logf_bound_max,
        "kernel_margin_at_depth": 2.0 + 2e-3 * depth, "depth": depth, "ploidy": ploidy, "n_pos": n_pos,
    }


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    for depth in (10, 40, 200):
        print(json.dumps(study(n, depth=depth, seed=depth)), flush=True)
