#!/usr/bin/env python3
"""Where the assemble kernel spends its cycles, per temperature (profiling build, -DMCHB_PROFILE).

    python profiles/phase_profile.py [shape ...]     (on a B200 box; one JSON object per shape)

Counters come from clock64() deltas and event counts accumulated by lane 0 of every warp
(assemble_kernel.cuh: MCHB_PROF_ADD); the profiling library is a separate .so and never the
one the tests or bench.py load."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mchap_b200 import build as _build  # noqa: E402

os.environ["MCHB_LIB"] = _build.build_profile()

import mchap_b200  # noqa: E402
from mchap_b200 import DenovoMCMC  # noqa: E402
from mchap_b200.synth import synth_items  # noqa: E402

NAMES = ["cyc_mutation", "cyc_structural", "cyc_slot_copy", "mut_accepts", "windows", "tier2a_evals", "tier2b_windows",
         "base_steps", "str_exact_evals", "str_screen_stay", "str_memo_stay", "intervals", "cyc_swapstep", "str_accepts",
         "tier2b_done"]

SHAPES = {
    "cfg3": dict(n_items=600, ploidy=8, n_pos=16, depth=100, temps=(0.01, 0.1, 0.5, 1.0), steps=300),
    "cfg3d": dict(n_items=592, ploidy=8, n_pos=16, depth=133, temps=(0.01, 0.1, 0.5, 1.0), steps=300),   # depth 100 at every SNV
    "cfg1d": dict(n_items=20000, ploidy=4, n_pos=8, depth=53, temps=(1.0,), steps=1500),                 # depth 40 at every SNV
    "cfg1": dict(n_items=20000, ploidy=4, n_pos=8, depth=40, temps=(1.0,), steps=1500),
    "hex2": dict(n_items=4000, ploidy=6, n_pos=8, depth=40, temps=(0.2, 1.0), steps=1500),
}


def run(name, n_items, ploidy, n_pos, depth, temps, steps, chains=2):
    dev = mchap_b200.default_device(0)
    batch = synth_items(n_items, ploidy=ploidy, n_pos=n_pos, depth=depth, seed=11)
    reads = [batch.item(i)[0] for i in range(n_items)]
    counts = [batch.item(i)[1] for i in range(n_items)]
    model = DenovoMCMC(ploidy=ploidy, n_alleles=[2] * n_pos, steps=steps, chains=chains, temperatures=temps,
                       random_seed=42)
    out = (C.c_uint64 * 128)()
    dev._lib.mchb_debug_counters(dev._h, out, 128, 1)
    _, results = model.fit_batch(reads, counts, raw=True, return_results=True)
    kernel_ms = dev.last_kernel_ms
    dev._check(dev._lib.mchb_debug_counters(dev._h, out, 128, 1))
    v = np.array(list(out), dtype=np.float64).reshape(8, 16)
    n_steps = n_items * chains * steps
    rec = {"shape": name, "items": n_items, "kernel_ms": kernel_ms, "mcmc_steps_per_s": n_steps / (kernel_ms * 1e-3),
           "llk_evals_per_step": float(results["llk_evals"].sum()) / n_steps, "per_temperature": []}
    for t in range(len(temps)):
        d = {"temp": temps[t]}
        for k, nm in enumerate(NAMES):
            d[nm + "_per_step"] = v[t, k] / n_steps
        rec["per_temperature"].append(d)
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(SHAPES)):
        run(nm, **SHAPES[nm])
