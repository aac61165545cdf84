#!/usr/bin/env python3
"""Long randomised parity run (not collected by pytest: run it on a B200 box).

    python tests/soak_parity.py [--items 1500] [--steps 400]

Many more items, seeds and shapes than the GPU test-suite: every item's de novo trace from the CUDA
path must be byte-identical to the C oracle's (same MT19937 stream), log-likelihoods within 1e-9
relative, and the consumed word / evaluation counters equal.  The oracle runs on all host cores.
Prints one JSON line per shape and exits non-zero on the first mismatch."""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mchap_b200 import DenovoMCMC  # noqa: E402
from mchap_b200.synth import synth_items  # noqa: E402
from oracle import oracle as O  # noqa: E402


def run_shape(name, n_items, steps, ploidy, n_pos, depth, temps=(1.0,), inbreeding=None, error_rate=None, seed=0,
              per_item_seeds=False):
    kw = {} if error_rate is None else {"error_rate": error_rate}
    batch = synth_items(n_items, ploidy=ploidy, n_pos=n_pos, depth=depth, seed=seed, **kw)
    reads = [batch.item(i)[0] for i in range(n_items)]
    counts = [batch.item(i)[1] for i in range(n_items)]
    model = DenovoMCMC(ploidy=ploidy, n_alleles=[2] * n_pos, inbreeding=inbreeding, steps=steps, chains=2,
                       temperatures=temps, random_seed=seed + 1)
    seeds = [int(s) for s in np.random.default_rng(seed).integers(0, 2 ** 31, size=n_items)] if per_item_seeds else None
    t0 = time.perf_counter()
    out, results = model.fit_batch(reads, counts, seeds=seeds, return_results=True, raw=True)
    t_gpu = time.perf_counter() - t0

    def ref(i):
        return O.denovo_fit(
            reads[i], counts[i], ploidy, [2] * n_pos, inbreeding=inbreeding, steps=steps, chains=2, alpha=model.alpha,
            beta=model.beta, n_intervals=model.n_intervals, fix_homozygous=model.fix_homozygous,
            recombination_step_probability=model.recombination_step_probability,
            partial_dosage_step_probability=model.partial_dosage_step_probability,
            dosage_step_probability=model.dosage_step_probability, temperatures=temps,
            random_seed=model.random_seed if seeds is None else seeds[i])

    t0 = time.perf_counter()
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        refs = list(ex.map(ref, range(n_items)))
    t_cpu = time.perf_counter() - t0
    worst = 0.0
    for i, r in enumerate(refs):
        g, l = out[i]
        if not np.array_equal(g, r["genotypes"]):
            bad = np.argwhere(g != r["genotypes"])[0]
            print(json.dumps({"shape": name, "item": i, "mismatch": "genotypes", "first": [int(x) for x in bad]}))
            sys.exit(1)
        if int(results["rng_words"][i]) != int(r["words"]) or int(results["llk_evals"][i]) != int(r["llk_evals"]):
            print(json.dumps({"shape": name, "item": i, "mismatch": "counters"}))
            sys.exit(1)
        ok = np.isfinite(r["llks"])
        rel = np.max(np.abs(l[ok] - r["llks"][ok]) / np.maximum(np.abs(r["llks"][ok]), 1e-300)) if ok.any() else 0.0
        worst = max(worst, float(rel))
        if not np.array_equal(np.isnan(l), np.isnan(r["llks"])) or rel > 1e-9:
            print(json.dumps({"shape": name, "item": i, "mismatch": "llks", "rel": float(rel)}))
            sys.exit(1)
    print(json.dumps({"shape": name, "items": n_items, "mcmc_steps_compared": n_items * 2 * steps,
                      "identical_traces": True, "worst_llk_rel_err": worst, "gpu_s": round(t_gpu, 2),
                      "oracle_s": round(t_cpu, 2)}), flush=True)


def run_call_shape(name, n_items, steps, ploidy, n_haps, n_pos, depth, step_type, prior=None, seed=0):
    """The same for CallingMCMC.fit_batch (mchap call) against the oracle's calling_fit."""
    from mchap_b200.calling import CallingMCMC
    from mchap_b200.synth import synth_haplotype_panel

    batch, panels, _ = synth_haplotype_panel(n_items, n_haps, n_pos, ploidy, depth=depth, seed=seed)
    reads = [batch.item(i)[0] for i in range(n_items)]
    counts = [batch.item(i)[1] for i in range(n_items)]
    model = CallingMCMC(ploidy=ploidy, haplotypes=None, steps=steps, chains=2, random_seed=seed + 1, step_type=step_type,
                        prior=prior)
    t0 = time.perf_counter()
    traces, results = model.fit_batch(reads, counts, haplotypes_list=list(panels), return_results=True)
    t_gpu = time.perf_counter() - t0

    def ref(i):
        return O.calling_fit(reads[i], counts[i], ploidy, panels[i], prior=prior, steps=steps, chains=2,
                             random_seed=seed + 1, step_type=step_type)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        refs = list(ex.map(ref, range(n_items)))
    t_cpu = time.perf_counter() - t0
    worst = 0.0
    for i, r in enumerate(refs):
        if not np.array_equal(traces[i].genotypes, r["genotypes"]) or int(results["rng_words"][i]) != int(r["words"]):
            print(json.dumps({"shape": name, "item": i, "mismatch": "genotypes or words"}))
            sys.exit(1)
        rel = np.max(np.abs(traces[i].llks - r["llks"]) / np.maximum(np.abs(r["llks"]), 1e-300))
        worst = max(worst, float(rel))
        if rel > 1e-9:
            print(json.dumps({"shape": name, "item": i, "mismatch": "llks", "rel": float(rel)}))
            sys.exit(1)
    print(json.dumps({"shape": name, "items": n_items, "mcmc_steps_compared": n_items * 2 * steps,
                      "identical_traces": True, "worst_llk_rel_err": worst, "gpu_s": round(t_gpu, 2),
                      "oracle_s": round(t_cpu, 2)}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=1500)
    ap.add_argument("--steps", type=int, default=400)
    a = ap.parse_args()
    n, s = a.items, a.steps
    run_shape("tetraploid 8 SNV depth 40 (configs[1])", n, s, 4, 8, 40, seed=101)
    run_shape("tetraploid 8 SNV depth 40, inbreeding 0.2, per-item seeds", n, s, 4, 8, 40, inbreeding=0.2, seed=102,
              per_item_seeds=True)
    run_shape("tetraploid 8 SNV depth 8, noisy reads", n, s, 4, 8, 8, error_rate=0.05, seed=103)
    run_shape("diploid 10 SNV depth 15", n, s, 2, 10, 15, seed=104)
    run_shape("hexaploid 6 SNV depth 60, temperatures 0.3/1", n // 3, s, 6, 6, 60, temps=(0.3, 1.0), seed=105)
    run_shape("tetraploid 12 SNV depth 150 (64+ unique reads)", n // 3, s // 2, 4, 12, 150, seed=106)
    run_shape("octoploid 16 SNV depth 100, 4 temperatures (configs[3])", max(n // 30, 8), max(s // 8, 20), 8, 16, 100,
              temps=(0.01, 0.1, 0.5, 1.0), seed=107)
    # round 2: depth 40 / 100 at every SNV (bench.py's workloads), hot mode over many steps, Rt and the
    # product rows in global memory, a heated replica with the Dirichlet-multinomial prior
    run_shape("tetraploid 8 SNV, 53 fragments (bench headline)", n, s, 4, 8, 53, seed=108)
    run_shape("octoploid 16 SNV, 133 fragments, 4 temperatures (bench configs[3])", max(n // 30, 8), max(s // 4, 20), 8, 16,
              133, temps=(0.01, 0.1, 0.5, 1.0), seed=109)
    run_shape("hexaploid 8 SNV, 80 fragments, temperatures 0.05/0.3/1, inbreeding 0.1", n // 6, s // 2, 6, 8, 80,
              temps=(0.05, 0.3, 1.0), inbreeding=0.1, seed=110)
    # mchap call (two-entry memo of the conditional distributions)
    run_call_shape("call: tetraploid, 32 haplotypes, Gibbs (configs[4])", n // 2, s, 4, 32, 8, 53, "Gibbs", seed=201)
    run_call_shape("call: tetraploid, 32 haplotypes, Gibbs, prior (0.1, flat)", n // 2, s, 4, 32, 8, 53, "Gibbs",
                   prior=(0.1, None), seed=202)
    run_call_shape("call: hexaploid, 12 haplotypes, Metropolis-Hastings", n // 2, s, 6, 12, 8, 53, "Metropolis-Hastings",
                   prior=(0.2, None), seed=203)
