#!/usr/bin/env python3
"""Kernel and end-to-end rates of the exhaustive caller at BASELINE configs[2] (hexaploid, 8 known
haplotypes, 1716 genotypes) and two neighbouring shapes; prints one JSON line per shape."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import mchap_b200  # noqa: E402
from mchap_b200.api import CallBatch  # noqa: E402
from mchap_b200.synth import synth_haplotype_panel  # noqa: E402

dev = mchap_b200.default_device(0)
shapes = [(50000, 6, 8, 8, 40, (0.1, None)), (50000, 6, 8, 8, 40, None), (20000, 4, 32, 8, 40, (0.1, None)),
          (20000, 2, 16, 8, 20, None)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for n, P, H, N, depth, prior in shapes:
    batch, panels, _ = synth_haplotype_panel(n, H, N, P, depth=depth, seed=777)
    reads = [batch.reads[batch.offsets[i]:batch.offsets[i + 1]] for i in range(n)]
    counts = [batch.counts[batch.offsets[i]:batch.offsets[i + 1]] for i in range(n)]
    t0 = time.perf_counter()
    cb = CallBatch(reads, list(panels), P, counts, None if prior is None else [prior] * n)
    t_pack = time.perf_counter() - t0
    G = int(cb.n_genotypes[0])
    dev.call_exact_mode(cb)
    reps, kms = 3, 0.0
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.call_exact_mode(cb)
        kms += dev.last_kernel_ms
    dt = (time.perf_counter() - t0) / reps
    dev.genotype_likelihoods(cb)
    gl_ms = dev.last_kernel_ms
    cbp = CallBatch(reads, list(panels), P, counts, None if prior is None else [prior] * n, device=dev)
    dev.call_exact_mode(cbp)
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.call_exact_mode(cbp)
    dtp = (time.perf_counter() - t0) / reps
    print(json.dumps({
        "shape": "P=%d H=%d N=%d depth=%d prior=%s" % (P, H, N, depth, prior), "items": n, "genotypes": G,
        "mean_unique_reads": float(np.diff(batch.offsets).mean()),
        "mode_kernel_ms": kms / reps, "mode_genotypes_per_s_kernel": n * G / (kms / reps * 1e-3),
        "mode_genotypes_per_s_e2e": n * G / dt, "mode_call_ms": dt * 1e3, "mode_call_ms_pinned": dtp * 1e3, "mode_genotypes_per_s_e2e_pinned": n * G / dtp, "pack_ms": t_pack * 1e3,
        "gl_kernel_ms": gl_ms, "gl_genotypes_per_s_kernel": n * G / (gl_ms * 1e-3)}), flush=True)
