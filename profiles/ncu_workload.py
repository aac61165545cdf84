#!/usr/bin/env python3
"""One fit_batch call of a named shape, for ncu captures of the assemble kernel.

    ncu --set full --import-source on --clock-control none --kernel-name-base mangled \\
        -k regex:assemble_kernelILi1ELb0 -c 1 -o gpurun_out/x python profiles/ncu_workload.py cfg1 7104 1500
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mchap_b200  # noqa: E402
from mchap_b200 import DenovoMCMC  # noqa: E402
from mchap_b200.synth import synth_items  # noqa: E402

SHAPES = {
    "cfg1d": dict(ploidy=4, n_pos=8, depth=53, temps=(1.0,)),          # depth 40 at every SNV
    "cfg3d": dict(ploidy=8, n_pos=16, depth=133, temps=(0.01, 0.1, 0.5, 1.0)),
    "cfg1": dict(ploidy=4, n_pos=8, depth=40, temps=(1.0,)),
    "cfg3": dict(ploidy=8, n_pos=16, depth=100, temps=(0.01, 0.1, 0.5, 1.0)),
    "hex2": dict(ploidy=6, n_pos=8, depth=40, temps=(0.2, 1.0)),
}
name = sys.argv[1]
n_items = int(sys.argv[2])
steps = int(sys.argv[3])
sh = SHAPES[name]
dev = mchap_b200.default_device(0)
batch = synth_items(n_items, ploidy=sh["ploidy"], n_pos=sh["n_pos"], depth=sh["depth"], seed=11)
reads = [batch.item(i)[0] for i in range(n_items)]
counts = [batch.item(i)[1] for i in range(n_items)]
model = DenovoMCMC(ploidy=sh["ploidy"], n_alleles=[2] * sh["n_pos"], steps=steps, chains=2, temperatures=sh["temps"],
                   random_seed=42)
model.fit_batch(reads, counts, raw=True)
print(name, n_items, steps, "kernel ms", dev.last_kernel_ms)
