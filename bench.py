#!/usr/bin/env python3
"""bench.py — BASELINE.json's metric on its configs, on N GPUs of one box.

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torchrun)
    python bench.py --impl reference ...                    (CPU arm: the reference's numba path under
                                                             multiprocessing, all host cores)

Headline line (config.workload = BASELINE configs[1]): synthetic tetraploid assemble, 10k loci x 8
SNVs x 100 samples, read depth 40 at every SNV (53 fragments per sample, each covering 50-100 % of
the locus), 2 chains x 1500 MCMC steps, flat prior, CLI-default step probabilities, one seed shared
by all items like the CLI (mchap/application/assemble.py:135).  The K timed bench steps are K
batches that together cover the 10 000 loci once.  With N GPUs every rank runs its own 10k-locus
workload (weak scaling, loci are independent: no collective on the data path); after the timed
region the ranks' per-locus results are gathered on the host in locus order
(mchap_b200.sharding.gather_by_locus), as the north star describes.

JSON keys follow the round contract: value (device-resident throughput, CUDA events, max over
ranks), e2e (host buffers through the C ABI, H2D + D2H inside the timed region), roofline (FP64
SIMT pipe: algorithmic flops of SURVEY.md section 8(d) / kernel time, against a DFMA probe
measured live), cpu_baseline (the reference's numba functions under multiprocessing.Pool on the
host cores, bounded sample; the C port of the oracle beside it), clocks.

The other BASELINE configs are measured by every rank in the same run and reported under
"configs" with the same keys (value / e2e / roofline / cpu_baseline, max over ranks):
  configs[2] call_exact          hexaploid, 8 known haplotypes, all 1716 genotypes, 50k pairs
  configs[3] assemble_octoploid  octoploid, 16 SNVs, depth 100, 4 temperatures
  configs[4] call_mcmc           tetraploid, 32 known haplotypes, Gibbs, 500 samples x a slice of loci
plus e2e_posterior / e2e_from_calls (the headline workload as the application consumes it).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PLOIDY, N_POS, DEPTH, SAMPLES, LOCI = 4, 8, 40, 100, 10000
CHAINS, MCMC_STEPS, SEED = 2, 1500, 42
METRIC = "locus x sample MCMC steps/s (assemble)"
UNIT = "MCMC steps/s"
WORKLOAD = ("synthetic tetraploid assemble: 10k loci x 8 SNVs x 100 samples, depth 40, "
            "2 chains x 1500 steps")

# DRAM bytes per (locus, sample) item of assemble_kernel<1,false>, from the committed ncu capture
# (profiles/r02_d1_ncu_raw.csv): dram__bytes_read.sum + dram__bytes_write.sum = 820.3 MB for the 6801 items
# of that launch (the items with <= 32 distinct reads of a 14208-item batch at depth 40)
NCU_DRAM_BYTES_PER_ITEM = 120620.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--loci", type=int, default=LOCI, help="loci covered by the K timed steps (default: the full config)")
    ap.add_argument("--cpu-items", type=int, default=0, help="items of the CPU baseline sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip configs[2], [3], [4]")
    ap.add_argument("--no-posterior", action="store_true")
    ap.add_argument("--exact-items", type=int, default=50000)
    ap.add_argument("--octoploid-steps", type=int, default=0, help="MCMC steps of the configs[3] line (0: 1500)")
    ap.add_argument("--fragments", type=int, default=0,
                    help="fragments per sample of the headline workload (0: depth 40 at every SNV = 53)")
    return ap.parse_args()


def fragments(depth, n_pos):
    from mchap_b200.synth import fragments_for_depth

    return fragments_for_depth(depth, n_pos)


# ----------------------------------------------------------------------------- CPU arms
def host_cores():
    """Cores this process may really use: os.cpu_count() counts the machine's logical CPUs, the
    scheduler affinity and the cgroup CPU quota say how many of them the container gets (round 1's
    CPU arm did not speed up from 16 to 32 threads: the 8-GPU box reports 32 logical CPUs)."""
    info = {"cpu_count": os.cpu_count() or 1}
    try:
        info["affinity"] = len(os.sched_getaffinity(0))
    except Exception:
        info["affinity"] = info["cpu_count"]
    quota = None
    try:
        txt = open("/sys/fs/cgroup/cpu.max").read().split()
        if txt[0] != "max":
            quota = float(txt[0]) / float(txt[1])
    except Exception:
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / per
        except Exception:
            pass
    info["cgroup_quota"] = quota
    used = min(info["cpu_count"], info["affinity"])
    if quota:
        used = max(1, min(used, int(round(quota))))
    info["used"] = used
    return info


def n_cores():
    return host_cores()["used"]



def port_pool(work, n_items, cores):
    """C oracle (oracle/mchap_oracle.c: a port of the reference's numba path) over `cores` threads,
    items split like the reference's --cores scheme (np.array_split)."""
    from concurrent.futures import ThreadPoolExecutor

    parts = [p for p in np.array_split(np.arange(n_items), cores) if len(p)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        done = sum(ex.map(work, parts))
    return done, time.perf_counter() - t0


def cpu_port_assemble(n_items, cores, frag, seed=12345):
    from mchap_b200.synth import synth_items
    from oracle import oracle as o

    o.lib()
    batch = synth_items(n_items, ploidy=PLOIDY, n_pos=N_POS, depth=frag, seed=seed)
    items = [batch.item(i) for i in range(n_items)]

    def work(idx):
        for i in idx:
            r, c = items[i]
            o.denovo_fit(r, c, PLOIDY, [2] * N_POS, steps=MCMC_STEPS, chains=CHAINS, random_seed=SEED)
        return len(idx)

    o.denovo_fit(items[0][0], items[0][1], PLOIDY, [2] * N_POS, steps=10, chains=1, random_seed=SEED)
    done, dt = port_pool(work, n_items, cores)
    return done * CHAINS * MCMC_STEPS / dt, dt


def cpu_numba_assemble(n_items, cores, frag, seed=12345):
    """The reference itself (oracle/_ref, numba) under multiprocessing.Pool(cores)."""
    from mchap_b200.synth import synth_items
    from oracle import ref_numba as rn

    batch = synth_items(n_items, ploidy=PLOIDY, n_pos=N_POS, depth=frag, seed=seed)
    items = [batch.item(i) for i in range(n_items)]
    kw = dict(ploidy=PLOIDY, n_alleles=[2] * N_POS, steps=MCMC_STEPS, chains=CHAINS, random_seed=SEED)
    rate, dt, used = rn.pool_rate("assemble", items, kw, cores)
    return rate * CHAINS * MCMC_STEPS, dt, used


def reference_kind():
    from oracle import ref_numba as rn

    return "reference" if rn.available() else "port"


def run_reference(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the headline path on the box's
    host cores — the unmodified numba package (oracle/_ref) under multiprocessing.Pool like --cores;
    the C port of the oracle under a thread pool only where numba / oracle/_ref is missing."""
    if rank != 0:
        return
    cores = n_cores()
    frag = args.fragments or fragments(DEPTH, N_POS)
    kind = reference_kind()
    # about 3-4 s of work per step on every core (numba: ~2e4 steps/s/core, port: ~6e4)
    n_items = args.cpu_items or (24 if kind == "reference" else 64) * cores
    from mchap_b200.synth import synth_items

    def sample(n, seed):
        batch = synth_items(n, ploidy=PLOIDY, n_pos=N_POS, depth=frag, seed=seed)
        return [batch.item(i) for i in range(n)]

    runner = None
    if kind == "reference":
        from oracle import ref_numba as rn

        kw = dict(ploidy=PLOIDY, n_alleles=[2] * N_POS, steps=MCMC_STEPS, chains=CHAINS, random_seed=SEED)
        runner = rn.Runner("assemble", kw, cores, sample(1, 1)[0])   # one pool for the whole run, JIT warmed
        run = lambda n, s: runner.rate(sample(n, s))
    else:
        run = lambda n, s: cpu_port_assemble(n, cores, frag, s)
    for w in range(args.warmup):
        run(max(cores, n_items // 8), 900 + w)
    t_all = 0.0
    for k in range(args.steps):
        _, dt = run(n_items, 1000 + k)
        t_all += dt
    if runner is not None:
        runner.close()
    value = n_items * CHAINS * MCMC_STEPS * args.steps / t_all
    sample = "%d locus x sample items per step (of the 1e6 of the config), %d %s" % (
        n_items, cores, "processes (multiprocessing.Pool, numba JIT warmed)" if kind == "reference" else "threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "ploidy": PLOIDY, "n_pos": N_POS, "depth": DEPTH, "fragments": frag,
                   "chains": CHAINS, "mcmc_steps": MCMC_STEPS, "items_per_step": n_items},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "host": host_cores(),
                         "note": "DenovoMCMC.fit of mchap v0.11.1 (numba) per item, as application/assemble.py:123-143 "
                                 "calls it" if kind == "reference" else "C port of the numba path (oracle/mchap_oracle.c)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- B200 arm
class Ranks:
    """torch.distributed plumbing of one bench process (one rank per GPU)."""

    def __init__(self, rank, world):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        self.local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(self.local)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.gpu = torch.device("cuda", self.local)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max(self, *values):
        """Max over ranks of each value (device all-reduce)."""
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.gpu)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def sum(self, *values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.gpu)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t]


def llk_flops(n_reads, ploidy, n_pos):
    """one log-likelihood evaluation = U * (P*N + 2P + 3) flop."""
    return np.asarray(n_reads, dtype=np.float64) * (ploidy * np.asarray(n_pos, dtype=np.float64) + 2 * ploidy + 3)


def fp64_roofline(flops, seconds, peak_tf, kernel, extra=None):
    achieved = flops / seconds / 1e12
    r = {"bound": "fp64_simt", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
         "frac": achieved / peak_tf if peak_tf else None, "traffic": None, "kernel": kernel,
         "peak_source": "DFMA probe kernel measured live in this run (MEASURED_PEAKS.json holds no FP64 figure)"}
    if extra:
        r.update(extra)
    return r


def config_call_exact(R, dev, args, peak_tf, cpu):
    """BASELINE configs[2]: genotype likelihoods/s of call-exact (hexaploid, 8 known haplotypes, all
    1716 genotypes, 50k locus x sample pairs; the work of exact.posterior_mode: mode + normaliser +
    allele frequencies / occurrences).  value: kernel time (CUDA events in the library); e2e: the
    host-buffer C-ABI call (page-locked inputs, H2D + D2H inside)."""
    from mchap_b200.api import CallBatch
    from mchap_b200.synth import synth_haplotype_panel

    n, P, H, N = args.exact_items, 6, 8, 8
    frag = fragments(DEPTH, N)
    batch, panels, _ = synth_haplotype_panel(n, H, N, P, depth=frag, seed=777 + R.rank)
    reads = [batch.reads[batch.offsets[i]:batch.offsets[i + 1]] for i in range(n)]
    counts = [batch.counts[batch.offsets[i]:batch.offsets[i + 1]] for i in range(n)]
    cb = CallBatch(reads, list(panels), P, counts, [(0.1, None)] * n, device=dev)
    G = int(cb.n_genotypes[0])
    dev.call_exact_mode(cb)
    dev.call_exact_mode(cb)
    reps, kms, launches = 5, 0.0, 0
    R.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = dev.call_exact_mode(cb)
        kms += dev.last_kernel_ms
        launches += dev.last_kernel_launches
    R.barrier()
    dt = time.perf_counter() - t0
    assert (out["results"]["status"] == 0).all()
    kms_max, dt_max = R.max(kms, dt)
    flops = float(llk_flops(batch.n_reads(), P, N).sum()) * G * reps
    res = {
        "metric": "genotype likelihoods/s (call-exact)", "unit": "genotypes/s",
        "workload": "synthetic hexaploid call-exact: 8 known haplotypes/locus, all 1716 genotypes, %d locus x sample "
                    "pairs per GPU, depth 40 (%d fragments), prior (F = 0.1, flat frequencies)" % (n, frag),
        "value": R.world * n * G * reps / (kms_max * 1e-3), "n_gpus": R.world, "ms_per_pass": kms_max / reps,
        "e2e": {"value": R.world * n * G * reps / dt_max, "unit": "genotypes/s",
                "h2d_bytes_per_step": int(cb.reads.nbytes + cb.counts.nbytes + cb.haps.nbytes + cb.items.nbytes),
                "d2h_bytes_per_step": int(n * (8 * P + 32 + 24) + 2 * 8 * cb.hap_total), "ms_per_step": 1e3 * dt_max / reps},
        "gpu_launches": launches, "mean_unique_reads": float(batch.n_reads().mean()),
        "roofline": fp64_roofline(flops, kms * 1e-3, peak_tf, "exact_kernel<6>", {
            "note": "one log-likelihood per genotype (G * U * (P*N + 2P + 3) flop per item); the kernel evaluates "
                    "each genotype once and parks the log joints for the frequency pass"}),
    }
    if cpu:
        from oracle import oracle as o

        cores = n_cores()
        n_cpu = min(n, 40 * cores)

        def work(idx):
            for i in idx:
                o.posterior_mode(reads[i], P, panels[i], counts[i], (0.1, None), True, True, True)
            return len(idx)

        done, dcpu = port_pool(work, n_cpu, cores)
        res["cpu_baseline"] = {"value": done * G / dcpu, "unit": "genotypes/s", "cores": cores, "kind": "port",
                               "sample": "%d items, %.1f s" % (n_cpu, dcpu),
                               "note": "the oracle's posterior_mode makes 2 enumeration passes per item like the reference"}
    return res


def config_call_mcmc(R, dev, args, peak_tf, cpu):
    """BASELINE configs[4]: tetraploid `mchap call` MCMC over 32 known haplotypes, 500 samples per
    locus, Gibbs steps, 2 chains x 2000 steps (CLI default); a slice of 40 of the 20k loci per pass
    (20 000 locus x sample items, 8e7 MCMC steps)."""
    from mchap_b200.calling import CallingMCMC
    from mchap_b200.synth import synth_haplotype_panel

    loci, samples, P, H, N, steps, chains = 40, 500, 4, 32, 8, 2000, 2
    n = loci * samples
    frag = fragments(DEPTH, N)
    batch, panels, _ = synth_haplotype_panel(n, H, N, P, depth=frag, seed=555 + R.rank)
    reads = [batch.item(i)[0] for i in range(n)]
    counts = [batch.item(i)[1] for i in range(n)]
    model = CallingMCMC(ploidy=P, haplotypes=None, steps=steps, chains=chains, random_seed=SEED, step_type="Gibbs",
                        device=dev)
    cb, haps, prs, ploidy, pmax, init = model._prepare(reads, counts, None, list(panels), None, None, None)
    st = model._step_type()
    out = dev.call_mcmc(cb, steps, chains, st, init, pmax)
    assert (out["results"]["status"] == 0).all()
    reps, kms, launches = 3, 0.0, 0
    R.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        d2h = int(out["alleles"].nbytes + out["llks"].nbytes + out["results"].nbytes)
        del out                         # the page-locked trace block goes back to the Device's pool
        out = dev.call_mcmc(cb, steps, chains, st, init, pmax)
        kms += dev.last_kernel_ms
        launches += dev.last_kernel_launches
    R.barrier()
    dt = time.perf_counter() - t0
    kms_max, dt_max = R.max(kms, dt)
    n_steps = n * chains * steps
    flops = float(llk_flops(batch.n_reads(), P, N).sum()) * P * H * chains * steps * reps
    res = {
        "metric": "locus x sample MCMC steps/s (call)", "unit": UNIT,
        "workload": "synthetic tetraploid mchap call MCMC: 32 known haplotypes/locus, 500 samples, %d of the 20k loci "
                    "per pass per GPU (%d items), depth 40 (%d fragments), Gibbs, 2 chains x 2000 steps" % (loci, n, frag),
        "value": R.world * n_steps * reps / (kms_max * 1e-3), "n_gpus": R.world, "ms_per_pass": kms_max / reps,
        "e2e": {"value": R.world * n_steps * reps / dt_max, "unit": UNIT,
                "h2d_bytes_per_step": int(cb.reads.nbytes + cb.counts.nbytes + cb.haps.nbytes + cb.items.nbytes),
                "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * dt_max / reps},
        "gpu_launches": launches, "mean_unique_reads": float(batch.n_reads().mean()),
        "roofline": fp64_roofline(flops, kms * 1e-3, peak_tf, "call_mcmc_kernel", {
            "note": "P * H log-likelihood evaluations of U * (P*N + 2P + 3) flop per MCMC step (Gibbs); the kernel "
                    "remembers a slot's categorical distribution per genotype, so a chain sitting in a mode only draws"}),
    }
    if cpu:
        from oracle import oracle as o

        cores = n_cores()
        n_cpu = min(n, 2 * cores)

        def work(idx):
            for i in idx:
                o.calling_fit(reads[i], counts[i], P, panels[i], prior=None, steps=steps, chains=chains,
                              random_seed=SEED, step_type="Gibbs")
            return len(idx)

        done, dcpu = port_pool(work, n_cpu, cores)
        res["cpu_baseline"] = {"value": done * chains * steps / dcpu, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "%d items, %.1f s" % (n_cpu, dcpu)}
    return res


def config_octoploid(R, dev, args, peak_tf, cpu):
    """BASELINE configs[3]: octoploid assemble, 16 SNVs, depth 100, parallel tempering with 4
    temperatures (0.01, 0.1, 0.5, 1.0), 2 chains x 1500 steps; loci sharded over the GPUs (every rank
    its own block, weak scaling).  One MCMC step = all four temperatures."""
    from mchap_b200 import _lib as L
    from mchap_b200.api import make_assemble_params, uniform_assemble_items
    from mchap_b200.assemble.mcmc import break_table
    from mchap_b200.synth import synth_items

    P, N, depth, temps, steps, chains = 8, 16, 100, (0.01, 0.1, 0.5, 1.0), 1500, 2
    if args.octoploid_steps:
        steps = args.octoploid_steps
    frag = fragments(depth, N)
    # three full waves of the kernel: a probe call tells how many items (warps) the GPU holds at once at
    # this shape (the real workload, loci sharded over the GPUs, runs hundreds of waves)
    pb = synth_items(64, ploidy=P, n_pos=N, depth=frag, seed=7)
    ptable, plens = break_table(N)
    pparams, pkeep = make_assemble_params(2, 1, 0.999, 0.5, 0.5, 1.0, ptable, plens, list(temps))
    pitems = uniform_assemble_items(pb.offsets, N, pb.max_allele, P, 1, 2, n_temps=len(temps), seed=SEED)
    dev.assemble_call(pitems, pparams, pb.reads.reshape(-1), pb.counts, np.ascontiguousarray(pb.n_alleles.reshape(-1)), None,
                      np.empty(64 * 2 * P * N, dtype=np.int8), np.empty(64 * 2),
                      (pb.reads.size, pb.counts.size, pb.n_alleles.size, 0, 64 * 2 * P * N, 64 * 2), mem=L.MEM_HOST)
    resident = max(dev.last_resident_warps, dev.sm_count)
    n = 3 * resident
    b = synth_items(n, ploidy=P, n_pos=N, depth=frag, seed=31337 + R.rank)
    items = uniform_assemble_items(b.offsets, N, b.max_allele, P, chains, steps, n_temps=len(temps), seed=SEED)
    table, lens = break_table(N)
    params, keep = make_assemble_params(steps, chains, 0.999, 0.5, 0.5, 1.0, table, lens, list(temps))
    g_len, l_len = n * chains * steps * P * N, n * chains * steps
    reads, cnts, nall = dev.pinned_concatenate([b.reads], np.float64), dev.pinned_concatenate([b.counts], np.int64), \
        np.ascontiguousarray(b.n_alleles.reshape(-1))
    out_g, out_l = dev.pinned_empty(g_len, np.int8), dev.pinned_empty(l_len, np.float64)

    def step():
        return dev.assemble_call(items, params, reads, cnts, nall, None, out_g, out_l,
                                 (reads.size, cnts.size, nall.size, 0, g_len, l_len), mem=L.MEM_HOST)

    # warm-up: the same items for 50 steps (allocations, clocks); the timed pass runs the full 1500
    wparams, wkeep = make_assemble_params(50, chains, 0.999, 0.5, 0.5, 1.0, table, lens, list(temps))
    witems = uniform_assemble_items(b.offsets, N, b.max_allele, P, chains, 50, n_temps=len(temps), seed=SEED)
    res0 = dev.assemble_call(witems, wparams, reads, cnts, nall, None, out_g, out_l,
                             (reads.size, cnts.size, nall.size, 0, g_len, l_len), mem=L.MEM_HOST)
    assert (res0["status"] == 0).all(), np.unique(res0["status"])
    reps, kms, launches = 1, 0.0, 0
    R.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        res0 = step()
        kms += dev.last_kernel_ms
        launches += dev.last_kernel_launches
    R.barrier()
    dt = time.perf_counter() - t0
    kms_max, dt_max = R.max(kms, dt)
    n_steps = n * chains * steps
    flops = float((res0["llk_evals"].astype(np.float64) * llk_flops(b.n_reads(), P, res0["n_het"])).sum()) * reps
    res = {
        "metric": METRIC, "unit": UNIT,
        "workload": "synthetic octoploid assemble: 16-SNV loci, depth 100 (%d fragments), parallel tempering (4 temps: "
                    "0.01, 0.1, 0.5, 1.0), 2 chains x %d steps, %d locus x sample items per GPU per pass" % (frag, steps, n),
        "value": R.world * n_steps * reps / (kms_max * 1e-3), "n_gpus": R.world, "ms_per_pass": kms_max / reps,
        "temperature_steps_per_s": R.world * n_steps * len(temps) * reps / (kms_max * 1e-3),
        "e2e": {"value": R.world * n_steps * reps / dt_max, "unit": UNIT,
                "h2d_bytes_per_step": int(reads.nbytes + cnts.nbytes + nall.nbytes + items.nbytes),
                "d2h_bytes_per_step": int(g_len + 8 * l_len + 24 * n), "ms_per_step": 1e3 * dt_max / reps},
        "gpu_launches": launches, "mean_unique_reads": float(b.n_reads().mean()), "resident_items_per_gpu": resident,
        "llk_evals_per_mcmc_step": float(res0["llk_evals"].sum()) / n_steps,
        "roofline": fp64_roofline(flops, kms * 1e-3, peak_tf, "assemble_kernel<CH = ceil(U / 32)>"),
    }
    if cpu:
        from oracle import oracle as o

        cores = n_cores()
        n_cpu, s_cpu = cores, 150
        its = [b.item(i) for i in range(n_cpu)]

        def work(idx):
            for i in idx:
                o.denovo_fit(its[i][0], its[i][1], P, [2] * N, steps=s_cpu, chains=chains, temperatures=temps,
                             random_seed=SEED)
            return len(idx)

        done, dcpu = port_pool(work, n_cpu, cores)
        res["cpu_baseline"] = {"value": done * chains * s_cpu / dcpu, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "%d items x 2 chains x %d steps, %.1f s" % (n_cpu, s_cpu, dcpu)}
    return res


def run_b200(args, rank, world):
    R = Ranks(rank, world)
    torch = R.torch

    import mchap_b200
    from mchap_b200 import _lib as L
    from mchap_b200.api import make_assemble_params, uniform_assemble_items
    from mchap_b200.assemble.mcmc import break_table
    from mchap_b200.synth import synth_items

    dev = mchap_b200.Device(R.local)
    gpu = R.gpu
    frag = args.fragments or fragments(DEPTH, N_POS)

    K, W = args.steps, args.warmup
    loci_per_step = -(-args.loci // K)
    items_per_step = loci_per_step * SAMPLES
    table, lens = break_table(N_POS)
    params, keep = make_assemble_params(MCMC_STEPS, CHAINS, 0.999, 0.5, 0.5, 1.0, table, lens, [1.0])

    # ---- synthetic inputs, one batch per timed step (distinct items; warm-up reuses batch 0..)
    batches = []
    for k in range(K):
        b = synth_items(items_per_step, ploidy=PLOIDY, n_pos=N_POS, depth=frag, seed=100003 * rank + k)
        items = uniform_assemble_items(b.offsets, N_POS, b.max_allele, PLOIDY, CHAINS, MCMC_STEPS, seed=SEED)
        batches.append((b, items))
    g_len = items_per_step * CHAINS * MCMC_STEPS * PLOIDY * N_POS
    l_len = items_per_step * CHAINS * MCMC_STEPS
    d_out_g = torch.empty(g_len, dtype=torch.int8, device=gpu)
    d_out_l = torch.empty(l_len, dtype=torch.float64, device=gpu)
    dev_in = []
    for b, items in batches:
        dev_in.append((torch.from_numpy(b.reads.reshape(-1)).to(gpu), torch.from_numpy(b.counts).to(gpu),
                       torch.from_numpy(b.n_alleles.reshape(-1)).to(gpu)))
    torch.cuda.synchronize()

    def device_step(k):
        b, items = batches[k]
        dr, dc, dn = dev_in[k]
        res = dev.assemble_call(items, params, dr.data_ptr(), dc.data_ptr(), dn.data_ptr(), None,
                                d_out_g.data_ptr(), d_out_l.data_ptr(),
                                (dr.numel(), dc.numel(), dn.numel(), 0, g_len, l_len), mem=L.MEM_DEVICE)
        return res, dev.last_kernel_ms, dev.last_kernel_launches

    # ---- value: inputs resident in HBM, device time by CUDA events on the launching stream.
    # Between timed steps the kernel writes a fresh multi-GB trace (>> 126 MB L2) and reads a
    # different batch, so no step finds its inputs in L2.
    for k in range(W):
        device_step(k % K)
    sampler = ClockSampler(R.local)
    sampler.start()
    R.barrier()
    kernel_ms, launches, flops, evals = [], 0, 0.0, 0
    per_locus = []
    t0 = time.perf_counter()
    for k in range(K):
        res, ms, nl = device_step(k)
        if (res["status"] != 0).any():
            raise RuntimeError("device reported item errors: %s" % np.unique(res["status"]))
        kernel_ms.append(ms)
        launches += nl
        b = batches[k][0]
        flops += float((res["llk_evals"].astype(np.float64) * llk_flops(b.n_reads(), PLOIDY, res["n_het"])).sum())
        evals += int(res["llk_evals"].sum())
        per_locus.extend(res["llk_evals"].reshape(loci_per_step, SAMPLES).sum(axis=1).tolist())
    R.barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_time = sum(kernel_ms) * 1e-3
    dev_time_max, wall_max = R.max(dev_time, wall)
    total_mcmc_steps = world * K * items_per_step * CHAINS * MCMC_STEPS
    value = total_mcmc_steps / dev_time_max

    # ---- host-side gather of per-locus results in global locus order (no collective on the data
    # path; this is the north star's "host-side gather of per-locus results", outside the timed region)
    from mchap_b200.sharding import gather_by_locus, locus_block

    n_loci_global = world * loci_per_step * K
    t0 = time.perf_counter()
    if world > 1:
        assert locus_block(n_loci_global, rank, world) == (rank * loci_per_step * K, (rank + 1) * loci_per_step * K)
    gathered = gather_by_locus(per_locus, n_loci_global)
    gather_ms = 1e3 * (time.perf_counter() - t0)
    assert len(gathered) == n_loci_global

    # ---- e2e: host (pinned) buffers through the C ABI, H2D + D2H inside the timed region
    e2e = None
    h_reads = h_counts = h_nall = None
    if not args.no_e2e:
        b, items = batches[0]
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_reads, h_counts, h_nall = pin(b.reads.reshape(-1)), pin(b.counts), pin(b.n_alleles.reshape(-1))
        h_out_g = torch.empty(g_len, dtype=torch.int8).pin_memory()
        h_out_l = torch.empty(l_len, dtype=torch.float64).pin_memory()

        def host_step():
            return dev.assemble_call(items, params, h_reads.numpy(), h_counts.numpy(), h_nall.numpy(), None,
                                     h_out_g.numpy(), h_out_l.numpy(),
                                     (h_reads.numel(), h_counts.numel(), h_nall.numel(), 0, g_len, l_len),
                                     mem=L.MEM_HOST)

        host_step()
        n_e2e = min(K, 3)
        R.barrier()
        t0 = time.perf_counter()
        e2e_kernel_ms = 0.0
        for _ in range(n_e2e):
            host_step()
            e2e_kernel_ms += dev.last_kernel_ms
        R.barrier()
        dt, = R.max(time.perf_counter() - t0)
        e2e = {
            "value": world * n_e2e * items_per_step * CHAINS * MCMC_STEPS / dt, "unit": UNIT,
            "h2d_bytes_per_step": int(h_reads.numel() * 8 + h_counts.numel() * 8 + h_nall.numel() + items.nbytes),
            "d2h_bytes_per_step": int(g_len + l_len * 8 + items_per_step * 24),
            "steps": n_e2e, "ms_per_step": 1e3 * dt / n_e2e, "kernel_ms_per_step": e2e_kernel_ms / n_e2e,
            "host_chunks": dev.last_host_chunks,
            "note": "one handle, pinned host buffers: H2D, kernels in chunks of consecutive items, each chunk's trace "
                    "D2H overlapping the next chunks' kernels",
        }
        del h_out_g, h_out_l

    # ---- the same workload as the application consumes it (mchap/application/assemble.py:123-170:
    # fit -> burn -> posterior / incongruence): traces stay in HBM, the tally kernel reduces them,
    # only the tallies come back.  Every rank, max over ranks.
    posterior = from_calls = None
    if not args.no_e2e and not args.no_posterior:
        from mchap_b200.api import TALLY_ITEM_DTYPE

        b, items = batches[0]
        burn, tsize = 500, 64   # CLI defaults: --mcmc-steps 1500 --mcmc-burn 500
        titems = np.zeros(len(items), dtype=TALLY_ITEM_DTYPE)
        titems["genotypes_off"] = items["genotypes_off"]
        titems["n_pos"], titems["ploidy"] = items["n_pos"], items["ploidy"]
        titems["chains"], titems["steps"], titems["burn"], titems["max_unique"] = CHAINS, MCMC_STEPS, burn, tsize
        titems["states_off"] = np.arange(len(items), dtype=np.int64) * tsize * PLOIDY * N_POS
        titems["tallies_off"] = np.arange(len(items), dtype=np.int64) * tsize * CHAINS
        o_states = np.zeros(len(items) * tsize * PLOIDY * N_POS, dtype=np.int8)
        o_counts = np.zeros(len(items) * tsize * CHAINS, dtype=np.int32)
        o_first = np.zeros_like(o_counts)

        def tally_step():
            return dev.assemble_tally_call(
                items, titems, params, h_reads.numpy(), h_counts.numpy(), h_nall.numpy(), None,
                (h_reads.numel(), h_counts.numel(), h_nall.numel(), 0, g_len, l_len), o_states, o_counts, o_first)

        tally_step()
        n_post = min(K, 3)
        R.barrier()
        t0 = time.perf_counter()
        for _ in range(n_post):
            _, tres = tally_step()
        R.barrier()
        dt, = R.max(time.perf_counter() - t0)
        posterior = {
            "value": world * n_post * items_per_step * CHAINS * MCMC_STEPS / dt, "unit": UNIT, "steps": n_post,
            "n_gpus": world, "ms_per_step": 1e3 * dt / n_post, "kernel_ms_per_step": dev.last_kernel_ms,
            "burn": burn, "max_unique": tsize, "items_over_max_unique": int((tres["status"] != 0).sum()),
            "mean_unique_genotypes": float(tres["n_het"].mean()),
            "d2h_bytes_per_step": int(o_states.nbytes + 2 * o_counts.nbytes + 2 * len(items) * 24),
            "note": "mchb_assemble_tally_batch on every rank: host inputs, traces kept in HBM, tallies of the burnt "
                    "trace (distinct genotypes, counts and first occurrences per chain) to the host",
        }

        # ---- the whole per-sample device path: raw allele calls in (9 bytes per base), tallies out
        # (mchb_encode_assemble_tally_batch: read encoding + de-duplication, de novo assembly, tallies)
        if getattr(b, "calls", None) is not None:
            from mchap_b200.encoding import ENCODE_ITEM_DTYPE

            n_it = len(items)
            calls = np.ascontiguousarray(b.calls).reshape(-1)
            probs = np.full(calls.size, 1 - 0.0024)          # io/bam.py:281: error rate only, no base qualities
            enc = np.zeros(n_it, dtype=ENCODE_ITEM_DTYPE)
            idx = np.arange(n_it, dtype=np.int64)
            enc["calls_off"] = enc["probs_off"] = idx * frag * N_POS
            enc["nalleles_off"] = idx * N_POS
            enc["reads_off"] = idx * frag * N_POS * 2
            enc["counts_off"] = idx * frag
            enc["n_reads"], enc["n_pos"], enc["max_allele"] = frag, N_POS, 2
            eres = np.zeros(n_it, dtype=tres.dtype)
            ares = np.zeros(n_it, dtype=tres.dtype)
            tres2 = np.zeros(n_it, dtype=tres.dtype)
            ptr = lambda a: a.ctypes.data_as(C.c_void_p)

            def calls_step():
                rc = dev._lib.mchb_encode_assemble_tally_batch(
                    dev._h, C.byref(params), ptr(enc), ptr(items), ptr(titems), n_it, ptr(calls), calls.size, ptr(probs),
                    probs.size, ptr(h_nall.numpy()), h_nall.numel(), 3.0, g_len, l_len, ptr(o_states), o_states.size,
                    ptr(o_counts), ptr(o_first), o_counts.size, ptr(eres), ptr(ares), ptr(tres2))
                dev._check(rc)

            want_states, want_counts = o_states.copy(), o_counts.copy()   # from the e2e_posterior runs above
            calls_step()
            same = bool(np.array_equal(o_states, want_states) and np.array_equal(o_counts, want_counts) and
                        np.array_equal(eres["n_het"], b.n_reads()))
            R.barrier()
            t0 = time.perf_counter()
            for _ in range(n_post):
                calls_step()
            R.barrier()
            dt, = R.max(time.perf_counter() - t0)
            from_calls = {
                "value": world * n_post * items_per_step * CHAINS * MCMC_STEPS / dt, "unit": UNIT, "steps": n_post,
                "n_gpus": world, "ms_per_step": 1e3 * dt / n_post, "kernel_ms_per_step": dev.last_kernel_ms,
                "h2d_bytes_per_step": int(calls.nbytes + probs.nbytes + h_nall.numel() + enc.nbytes + items.nbytes),
                "same_tallies_as_e2e_posterior": same,
                "note": "mchb_encode_assemble_tally_batch: raw fragments (int8 calls + P(correct)) in, read encoding + "
                        "de-duplication, assembly and tallies on the device, tallies out",
            }
            del want_states, want_counts
        del o_states, o_counts, o_first

    # ---- roofline of the dominant kernels (assemble_kernel<1> and <2>): FP64 SIMT pipe
    n_reads_all = np.concatenate([b.n_reads() for b, _ in batches])
    peak_tf = dev.measure_fp64_peak()
    in_bytes = sum(x[0].numel() * 8 + x[1].numel() * 8 for x in dev_in) / K
    out_bytes = g_len + l_len * 8
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    roofline = fp64_roofline(flops, dev_time, peak_tf, "assemble_kernel<1> + assemble_kernel<2> (items with <= 32 / 33-64 distinct reads)", {
        "traffic": NCU_DRAM_BYTES_PER_ITEM * items_per_step, "traffic_per_item": NCU_DRAM_BYTES_PER_ITEM,
        "traffic_source": "profiles/r02_d1_ncu_raw.csv: (dram__bytes_read.sum + dram__bytes_write.sum) of one "
                          "`ncu --set full` launch of assemble_kernel<1,false> / its 6801 items, scaled to the %d items "
                          "of one bench step; algorithmic bytes per item = %.0f" % (
                              items_per_step, (in_bytes + out_bytes) / items_per_step),
        "llk_evals_per_mcmc_step": evals / (K * items_per_step * CHAINS * MCMC_STEPS),
        "items_with_more_than_32_reads": float((n_reads_all > 32).mean()),
        "hbm_achieved_gbs": (in_bytes + out_bytes) / (dev_time / K) / 1e9,
        "hbm_peak_gbs": hbm_peak if hbm_peak else 6650.0,
        "hbm_peak_source": "MEASURED_PEAKS.json" if hbm_peak else "fallback",
    })
    del d_out_g, d_out_l, dev_in
    torch.cuda.empty_cache()

    cpu_ok = rank == 0 and world == 1 and not args.no_cpu_baseline
    configs = None
    if not args.no_configs:
        configs = {}
        for name, fn in (("call_exact", config_call_exact), ("assemble_octoploid", config_octoploid),
                         ("call_mcmc", config_call_mcmc)):
            configs[name] = fn(R, dev, args, peak_tf, cpu_ok)

    cpu = None
    if cpu_ok:
        cores = n_cores()
        kind = reference_kind()
        if kind == "reference":
            n_cpu = args.cpu_items or 120 * cores
            v, dt, used = cpu_numba_assemble(n_cpu, cores, frag)
            cpu = {"value": v, "unit": UNIT, "cores": used, "kind": "reference",
                   "sample": "%d of the %d locus x sample items of the workload, %.1f s; mchap v0.11.1 DenovoMCMC.fit "
                             "(numba, JIT warmed) under multiprocessing.Pool(%d)" % (n_cpu, LOCI * SAMPLES, dt, used)}
        n_port = args.cpu_items or 100 * cores
        vp, dtp = cpu_port_assemble(n_port, cores, frag)
        port = {"value": vp, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d items, %.1f s; C port of the numba path (oracle/mchap_oracle.c), thread pool" % (n_port, dtp)}
        if cpu is None:
            cpu = port
        else:
            cpu["port"] = port
        cpu["host"] = host_cores()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dev_time_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "ploidy": PLOIDY, "n_pos": N_POS, "depth": DEPTH, "fragments": frag,
                       "depth_definition": "mean reads covering an SNV (fragments cover 50-100 % of the locus)",
                       "samples": SAMPLES, "loci_per_step": loci_per_step, "loci_timed": loci_per_step * K,
                       "chains": CHAINS, "mcmc_steps": MCMC_STEPS, "seed": SEED, "prior": "flat",
                       "l2": "each step reads a different batch and writes a multi-GB trace (> L2)",
                       "mean_unique_reads": float(n_reads_all.mean())},
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "configs": configs, "e2e_posterior": posterior, "e2e_from_calls": from_calls,
            "wall_ms_per_step": 1e3 * wall_max / K, "host_gather_ms": gather_ms,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        R.barrier()
        R.dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_b200(args, rank, world)


if __name__ == "__main__":
    main()
