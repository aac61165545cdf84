#!/usr/bin/env python3
"""Throughput of the other BASELINE.json shapes (not bench lines: bench.py measures configs[1]).

    python profiles/shape_sweep.py            (on a B200 box; prints one JSON object per shape)

Each shape runs through the public batch API with host buffers (fit_batch / fit_posterior_batch /
CallingMCMC.fit_batch), so the numbers include packing, H2D and D2H; `kernel` is the device time
the library reports for the same call."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mchap_b200  # noqa: E402
from mchap_b200 import DenovoMCMC  # noqa: E402
from mchap_b200.synth import synth_items  # noqa: E402


def assemble_shape(name, n_items, ploidy, n_pos, depth, temps, steps, chains=2, posterior=False):
    dev = mchap_b200.default_device(0)
    batch = synth_items(n_items, ploidy=ploidy, n_pos=n_pos, depth=depth, seed=11)
    reads = [batch.item(i)[0] for i in range(n_items)]
    counts = [batch.item(i)[1] for i in range(n_items)]
    model = DenovoMCMC(ploidy=ploidy, n_alleles=[2] * n_pos, steps=steps, chains=chains, temperatures=temps,
                       random_seed=42)
    run = (lambda: model.fit_posterior_batch(reads, counts, burn=steps // 3)) if posterior else \
          (lambda: model.fit_batch(reads, counts, raw=True))
    model.fit_batch(reads, counts, raw=True)     # first call: module loading, page-locked buffers
    model.fit_batch(reads, counts, raw=True)     # kernel time of the sampler alone (traces to the host)
    kernel_ms = dev.last_kernel_ms
    run()
    t0 = time.perf_counter()
    run()
    dt = time.perf_counter() - t0
    n_steps = n_items * chains * steps
    print(json.dumps({
        "shape": name, "items": n_items, "ploidy": ploidy, "n_pos": n_pos, "depth": depth, "temperatures": len(temps),
        "steps": steps, "chains": chains, "api": "fit_posterior_batch" if posterior else "fit_batch",
        "mcmc_steps_per_s_api": n_steps / dt, "mcmc_steps_per_s_kernel": n_steps / (kernel_ms * 1e-3),
        "temperature_steps_per_s_kernel": n_steps * len(temps) / (kernel_ms * 1e-3),
        "mean_unique_reads": float(np.mean([len(r) for r in reads])),
    }), flush=True)


def call_shape(name, n_items, ploidy, n_haps, n_pos, depth, steps, step_type, chains=2):
    from mchap_b200.calling import CallingMCMC
    from mchap_b200.synth import synth_haplotype_panel

    dev = mchap_b200.default_device(0)
    batch, panels, _ = synth_haplotype_panel(n_items, n_haps, n_pos, ploidy, depth=depth, seed=5)
    reads = [batch.item(i)[0] for i in range(n_items)]
    counts = [batch.item(i)[1] for i in range(n_items)]
    model = CallingMCMC(ploidy=ploidy, haplotypes=panels[0], steps=steps, chains=chains, random_seed=42,
                        step_type=step_type)
    run = lambda: model.fit_batch(reads, counts, haplotypes_list=list(panels))
    run()
    t0 = time.perf_counter()
    run()
    dt = time.perf_counter() - t0
    n_steps = n_items * chains * steps
    print(json.dumps({
        "shape": name, "items": n_items, "ploidy": ploidy, "n_haplotypes": n_haps, "n_pos": n_pos, "depth": depth,
        "steps": steps, "chains": chains, "step_type": step_type, "api": "CallingMCMC.fit_batch",
        "mcmc_steps_per_s_api": n_steps / dt, "mcmc_steps_per_s_kernel": n_steps / (dev.last_kernel_ms * 1e-3),
    }), flush=True)


if __name__ == "__main__":
    call_shape("configs[4] call: tetraploid, 32 known haplotypes, 8 SNV, depth 40, Gibbs", 20000, 4, 32, 8, 40, 2000, "Gibbs")
    call_shape("call: tetraploid, 32 known haplotypes, Metropolis-Hastings", 20000, 4, 32, 8, 40, 2000, "Metropolis-Hastings")
    if "--call-only" in sys.argv:
        sys.exit(0)
    assemble_shape("configs[1] tetraploid 8 SNV depth 40", 20000, 4, 8, 40, (1.0,), 1500, posterior=True)
    assemble_shape("configs[1] with inbreeding-free diploid 6 SNV depth 20", 20000, 2, 6, 20, (1.0,), 1500, posterior=True)
    assemble_shape("hexaploid 8 SNV depth 40, 2 temperatures", 8000, 6, 8, 40, (0.2, 1.0), 1500, posterior=True)
    assemble_shape("configs[3] octoploid 16 SNV depth 100, 4 temperatures", 1200, 8, 16, 100, (0.01, 0.1, 0.5, 1.0), 300,
                   posterior=True)
