#!/usr/bin/env python3
"""Host-side cost of the public batch API (cProfile of one call each) + raw transfer costs.
    python profiles/host_overheads.py     (on a B200 box)"""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mchap_b200  # noqa: E402
from mchap_b200 import CallingMCMC, DenovoMCMC  # noqa: E402
from mchap_b200.synth import fragments_for_depth, synth_haplotype_panel, synth_items  # noqa: E402

dev = mchap_b200.default_device(0)


def prof(label, fn, top=18):
    fn()
    t0 = time.perf_counter()
    fn()
    dt = time.perf_counter() - t0
    pr = cProfile.Profile()
    pr.enable()
    fn()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(top)
    print("=====", label, "wall %.1f ms, kernel %.1f ms" % (dt * 1e3, dev.last_kernel_ms))
    print("\n".join(l for l in s.getvalue().split("\n")[4:] if l.strip()))


n = 20000
frag = fragments_for_depth(40, 8)
b = synth_items(n, ploidy=4, n_pos=8, depth=frag, seed=11)
reads = [b.item(i)[0] for i in range(n)]
counts = [b.item(i)[1] for i in range(n)]
m = DenovoMCMC(ploidy=4, n_alleles=[2] * 8, steps=1500, chains=2, random_seed=42)
prof("DenovoMCMC.fit_posterior_batch", lambda: m.fit_posterior_batch(reads, counts, burn=500))
prof("DenovoMCMC.fit_batch(raw)", lambda: m.fit_batch(reads, counts, raw=True))
prof("DenovoMCMC.fit_batch", lambda: m.fit_batch(reads, counts))

batch, panels, _ = synth_haplotype_panel(n, 32, 8, 4, depth=frag, seed=5)
reads = [batch.item(i)[0] for i in range(n)]
counts = [batch.item(i)[1] for i in range(n)]
cm = CallingMCMC(ploidy=4, haplotypes=panels[0], steps=2000, chains=2, random_seed=42)
prof("CallingMCMC.fit_batch", lambda: cm.fit_batch(reads, counts, haplotypes_list=list(panels)))
prof("CallingMCMC.fit_posterior_batch", lambda: cm.fit_posterior_batch(reads, counts, burn=500, haplotypes_list=list(panels)))

for mb in (64, 512, 2400):
    t0 = time.perf_counter()
    a = dev.pinned_empty(mb << 20, np.int8)
    t1 = time.perf_counter()
    z = np.empty(mb << 20, dtype=np.int8)
    z[::4096] = 1
    t2 = time.perf_counter()
    print("pinned_empty %d MB: %.1f ms; np.empty + touch: %.1f ms" % (mb, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
    del a, z
