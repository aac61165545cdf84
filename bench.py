#!/usr/bin/env python3
"""bench.py — locus x sample MCMC steps/s of the assemble hot path (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torchrun)
    python bench.py --impl reference ...                    (CPU arm: the oracle port, all host cores)

Workload (config.workload): synthetic tetraploid assemble, 10k loci x 8 SNVs x 100 samples,
depth 40, 2 chains x 1500 MCMC steps, flat prior, CLI-default step probabilities, seed shared by
all items like the CLI (mchap/application/assemble.py:135).  The K timed bench steps are K batches
that together cover the 10 000 loci once (a bench "step" = one pass of the hot path over one batch
of 10000/K loci x 100 samples).  With N GPUs every rank runs its own 10k-locus workload (weak
scaling, no collective on the data path; loci are independent).

JSON keys follow the round contract: value (device-resident throughput, CUDA events, max over
ranks), e2e (host buffers through the C ABI, H2D + D2H inside the timed region), roofline
(FP64 SIMT pipe: algorithmic flops of SURVEY.md section 8(d) / kernel time, against a DFMA
probe measured live), cpu_baseline (C oracle on the host cores, bounded sample), clocks.
Secondary objects on the same line: e2e_posterior (the same workload as the application consumes
it: traces kept in HBM, tallies of the burnt trace returned) and call_exact (genotype
likelihoods/s of configs[2]).
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
# (profiles/README.md): 250.2 MB read + written by a launch of 1964 items (traces dominate)
NCU_DRAM_BYTES_PER_ITEM = 127416.0

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
    ap.add_argument("--no-call-exact", action="store_true")
    ap.add_argument("--no-posterior", action="store_true")
    ap.add_argument("--exact-items", type=int, default=50000)
    return ap.parse_args()


# ----------------------------------------------------------------------------- CPU arm
def cpu_oracle_rate(n_items, cores, seed=12345):
    """C oracle (oracle/mchap_oracle.c: a port of the reference's numba path) over `cores`
    threads, items split like the reference's --cores scheme (np.array_split)."""
    from concurrent.futures import ThreadPoolExecutor

    from mchap_b200.synth import synth_items
    from oracle import oracle as o

    o.lib()
    batch = synth_items(n_items, ploidy=PLOIDY, n_pos=N_POS, depth=DEPTH, seed=seed)
    items = [batch.item(i) for i in range(n_items)]

    def work(idx):
        for i in idx:
            r, c = items[i]
            o.denovo_fit(r, c, PLOIDY, [2] * N_POS, steps=MCMC_STEPS, chains=CHAINS, random_seed=SEED)
        return len(idx)

    o.denovo_fit(items[0][0], items[0][1], PLOIDY, [2] * N_POS, steps=10, chains=1, random_seed=SEED)
    parts = [p for p in np.array_split(np.arange(n_items), cores) if len(p)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        done = sum(ex.map(work, parts))
    dt = time.perf_counter() - t0
    return done * CHAINS * MCMC_STEPS / dt, dt


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_items = args.cpu_items or 64 * cores
    vals = []
    for _ in range(args.warmup):
        cpu_oracle_rate(max(cores, n_items // 4), cores)
    t_all = 0.0
    for k in range(args.steps):
        v, dt = cpu_oracle_rate(n_items, cores, seed=1000 + k)
        vals.append(v)
        t_all += dt
    value = n_items * CHAINS * MCMC_STEPS * args.steps / t_all
    sample = "%d locus x sample items per step (of the 1e6 of the config), %d threads" % (n_items, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "ploidy": PLOIDY, "n_pos": N_POS, "depth": DEPTH,
                   "chains": CHAINS, "mcmc_steps": MCMC_STEPS, "items_per_step": n_items},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
def algorithmic_flops(results, n_reads):
    """one llk evaluation = U * (P*N + 2P + 3) flop with N = positions passed to
    the sampler (n_het); the kernel reports evaluations and n_het per item."""
    n_het = results["n_het"].astype(np.float64)
    w = n_reads.astype(np.float64) * (PLOIDY * n_het + 2 * PLOIDY + 3)
    return float((results["llk_evals"].astype(np.float64) * w).sum())


def call_exact_line(dev, args):
    """Second half of BASELINE.json's metric: genotype likelihoods/s of call-exact on configs[2]
    (hexaploid, 8 known haplotypes, all 1716 genotypes, 50k locus x sample pairs; the work of
    exact.posterior_mode: mode + normaliser + support + allele frequencies).  Host buffers in,
    results out (e2e) and kernel-only device time; CPU: the C oracle on a bounded sample."""
    from concurrent.futures import ThreadPoolExecutor

    from mchap_b200.api import CallBatch
    from mchap_b200.synth import synth_haplotype_panel
    from oracle import oracle as o

    n, P, H, N = args.exact_items, 6, 8, 8
    batch, panels, _ = synth_haplotype_panel(n, H, N, P, depth=DEPTH, seed=777)
    reads = [batch.reads[batch.offsets[i]:batch.offsets[i + 1]] for i in range(n)]
    counts = [batch.counts[batch.offsets[i]:batch.offsets[i + 1]] for i in range(n)]
    cb = CallBatch(reads, list(panels), P, counts, [(0.1, None)] * n)
    G = int(cb.n_genotypes[0])
    dev.call_exact_mode(cb)
    t0 = time.perf_counter()
    reps = 3
    kms = 0.0
    for _ in range(reps):
        dev.call_exact_mode(cb)
        kms += dev.last_kernel_ms
    dt = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    n_cpu = min(n, 40 * cores)

    def work(idx):
        for i in idx:
            o.posterior_mode(reads[i], P, panels[i], counts[i], (0.1, None), True, True, True)
        return len(idx)

    parts = [p for p in np.array_split(np.arange(n_cpu), cores) if len(p)]
    t1 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        done = sum(ex.map(work, parts))
    dcpu = time.perf_counter() - t1
    return {
        "metric": "genotype likelihoods/s (call-exact)", "unit": "genotypes/s",
        "workload": "synthetic hexaploid call-exact: 8 known haplotypes/locus, all 1716 genotypes, %d locus x sample pairs" % n,
        "value": n * G * reps / (kms * 1e-3), "e2e": n * G * reps / dt, "ms_per_pass": kms / reps,
        "cpu_baseline": {"value": done * G / dcpu, "cores": cores, "kind": "port", "sample": "%d items" % n_cpu,
                         "note": "the oracle's posterior_mode makes 2 enumeration passes per item like the reference"},
    }


def run_b200(args, rank, world):
    import torch
    import torch.distributed as dist

    import mchap_b200
    from mchap_b200 import _lib as L
    from mchap_b200.api import make_assemble_params, uniform_assemble_items
    from mchap_b200.assemble.mcmc import break_table
    from mchap_b200.synth import synth_items

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = mchap_b200.Device(local)
    gpu = torch.device("cuda", local)

    K, W = args.steps, args.warmup
    loci_per_step = -(-args.loci // K)
    items_per_step = loci_per_step * SAMPLES
    table, lens = break_table(N_POS)
    params, keep = make_assemble_params(MCMC_STEPS, CHAINS, 0.999, 0.5, 0.5, 1.0, table, lens, [1.0])

    # ---- synthetic inputs, one batch per timed step (distinct items; warm-up reuses batch 0..)
    batches = []
    for k in range(K):
        b = synth_items(items_per_step, ploidy=PLOIDY, n_pos=N_POS, depth=DEPTH, seed=100003 * rank + k)
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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: inputs resident in HBM, device time by CUDA events on the launching stream.
    # Between timed steps the kernel writes a fresh 9.6 GB trace (>> 126 MB L2) and reads a
    # different batch, so no step finds its inputs in L2.
    for k in range(W):
        device_step(k % K)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    kernel_ms, launches, flops, evals = [], 0, 0.0, 0
    t0 = time.perf_counter()
    for k in range(K):
        res, ms, nl = device_step(k)
        if (res["status"] != 0).any():
            raise RuntimeError("device reported item errors: %s" % np.unique(res["status"]))
        kernel_ms.append(ms)
        launches += nl
        flops += algorithmic_flops(res, batches[k][0].n_reads())
        evals += int(res["llk_evals"].sum())
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_time = sum(kernel_ms) * 1e-3
    t = torch.tensor([dev_time, wall], dtype=torch.float64, device=gpu)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_time_max, wall_max = float(t[0]), float(t[1])
    total_mcmc_steps = world * K * items_per_step * CHAINS * MCMC_STEPS
    value = total_mcmc_steps / dev_time_max

    # ---- e2e: host (pinned) buffers through the C ABI, H2D + D2H inside the timed region
    e2e = None
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
        barrier()
        t0 = time.perf_counter()
        e2e_kernel_ms = 0.0
        for _ in range(n_e2e):
            host_step()
            launches_e2e = dev.last_kernel_launches
            e2e_kernel_ms += dev.last_kernel_ms
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=gpu)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {
            "value": world * n_e2e * items_per_step * CHAINS * MCMC_STEPS / float(t[0]), "unit": UNIT,
            "h2d_bytes_per_step": int(h_reads.numel() * 8 + h_counts.numel() * 8 + h_nall.numel() + items.nbytes),
            "d2h_bytes_per_step": int(g_len + l_len * 8 + items_per_step * 24),
            "steps": n_e2e, "ms_per_step": 1e3 * float(t[0]) / n_e2e, "kernel_ms_per_step": e2e_kernel_ms / n_e2e,
            "host_chunks": dev.last_host_chunks, "note": "one handle, pinned host buffers: H2D, kernels in chunks of consecutive items, each chunk's trace D2H overlapping the next chunks' kernels",
        }

    # ---- the same workload as the application consumes it (mchap/application/assemble.py:123-170:
    # fit -> burn -> posterior / incongruence): traces stay in HBM, the tally kernel reduces them,
    # only the tallies come back
    posterior = None
    if rank == 0 and not args.no_e2e and not args.no_posterior:
        from mchap_b200.api import TALLY_ITEM_DTYPE

        b, items = batches[0]
        burn, table = 500, 64   # CLI defaults: --mcmc-steps 1500 --mcmc-burn 500
        titems = np.zeros(len(items), dtype=TALLY_ITEM_DTYPE)
        titems["genotypes_off"] = items["genotypes_off"]
        titems["n_pos"], titems["ploidy"] = items["n_pos"], items["ploidy"]
        titems["chains"], titems["steps"], titems["burn"], titems["max_unique"] = CHAINS, MCMC_STEPS, burn, table
        titems["states_off"] = np.arange(len(items), dtype=np.int64) * table * PLOIDY * N_POS
        titems["tallies_off"] = np.arange(len(items), dtype=np.int64) * table * CHAINS
        o_states = np.zeros(len(items) * table * PLOIDY * N_POS, dtype=np.int8)
        o_counts = np.zeros(len(items) * table * CHAINS, dtype=np.int32)
        o_first = np.zeros_like(o_counts)

        def tally_step():
            return dev.assemble_tally_call(
                items, titems, params, h_reads.numpy(), h_counts.numpy(), h_nall.numpy(), None,
                (h_reads.numel(), h_counts.numel(), h_nall.numel(), 0, g_len, l_len), o_states, o_counts, o_first)

        tally_step()
        t0 = time.perf_counter()
        n_post = min(K, 3)
        for _ in range(n_post):
            _, tres = tally_step()
        dt = time.perf_counter() - t0
        posterior = {
            "value": n_post * items_per_step * CHAINS * MCMC_STEPS / dt, "unit": UNIT, "steps": n_post,
            "ms_per_step": 1e3 * dt / n_post, "kernel_ms_per_step": dev.last_kernel_ms,
            "burn": burn, "max_unique": table, "items_over_max_unique": int((tres["status"] != 0).sum()),
            "mean_unique_genotypes": float(tres["n_het"].mean()),
            "d2h_bytes_per_step": int(o_states.nbytes + 2 * o_counts.nbytes + 2 * len(items) * 24),
            "note": "mchb_assemble_tally_batch on this rank: host inputs, traces kept in HBM, tallies of the burnt "
                    "trace (distinct genotypes, counts and first occurrences per chain) to the host",
        }

    # ---- the whole per-sample device path: raw allele calls in (9 bytes per base), tallies out
    # (mchb_encode_assemble_tally_batch: read encoding + de-duplication, de novo assembly, tallies)
    from_calls = None
    if posterior is not None and getattr(b, "calls", None) is not None:
        from mchap_b200.encoding import ENCODE_ITEM_DTYPE

        n_it = len(items)
        calls = np.ascontiguousarray(b.calls).reshape(-1)
        probs = np.full(calls.size, 1 - 0.0024)          # io/bam.py:281: error rate only, no base qualities
        enc = np.zeros(n_it, dtype=ENCODE_ITEM_DTYPE)
        idx = np.arange(n_it, dtype=np.int64)
        enc["calls_off"] = enc["probs_off"] = idx * DEPTH * N_POS
        enc["nalleles_off"] = idx * N_POS
        enc["reads_off"] = idx * DEPTH * N_POS * 2
        enc["counts_off"] = idx * DEPTH
        enc["n_reads"], enc["n_pos"], enc["max_allele"] = DEPTH, N_POS, 2
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
        t0 = time.perf_counter()
        for _ in range(n_post):
            calls_step()
        dt = time.perf_counter() - t0
        from_calls = {
            "value": n_post * items_per_step * CHAINS * MCMC_STEPS / dt, "unit": UNIT, "steps": n_post,
            "ms_per_step": 1e3 * dt / n_post, "kernel_ms_per_step": dev.last_kernel_ms,
            "h2d_bytes_per_step": int(calls.nbytes + probs.nbytes + h_nall.numel() + enc.nbytes + items.nbytes),
            "same_tallies_as_e2e_posterior": same,
            "note": "mchb_encode_assemble_tally_batch: raw fragments (int8 calls + P(correct)) in, read encoding + "
                    "de-duplication, assembly and tallies on the device, tallies out",
        }

    # ---- roofline of the dominant kernel (assemble_kernel): FP64 SIMT pipe
    main_items_per_launch = int(np.mean([(b.n_reads() <= 32).sum() for b, _ in batches]))
    peak_tf = dev.measure_fp64_peak()
    achieved_tf = flops / dev_time / 1e12
    in_bytes = sum(x[0].numel() * 8 + x[1].numel() * 8 for x in dev_in) / K
    out_bytes = g_len + l_len * 8
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    roofline = {
        "bound": "fp64_simt", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf if peak_tf else None,
        "traffic": NCU_DRAM_BYTES_PER_ITEM * main_items_per_launch, "traffic_per_item": NCU_DRAM_BYTES_PER_ITEM,
        "traffic_source": "profiles/r01_final_ncu_raw.csv: (dram__bytes_read.sum + dram__bytes_write.sum) of one "
                          "`ncu --set full` launch of assemble_kernel<1,false> / its 1964 items, scaled to the "
                          "%d items of this bench's main-class launch; algorithmic bytes per item = %.0f" % (
                              main_items_per_launch, (in_bytes + out_bytes) / items_per_step),
        "peak_source": "DFMA probe kernel measured live in this run (MEASURED_PEAKS.json holds no FP64 figure)",
        "kernel": "assemble_kernel<1>", "llk_evals_per_mcmc_step": evals / (K * items_per_step * CHAINS * MCMC_STEPS),
        "hbm_achieved_gbs": (in_bytes + out_bytes) / (dev_time / K) / 1e9,
        "hbm_peak_gbs": hbm_peak if hbm_peak else 6650.0,
        "hbm_peak_source": "MEASURED_PEAKS.json" if hbm_peak else "fallback",
    }

    call_exact = None
    if rank == 0 and not args.no_call_exact:
        call_exact = call_exact_line(dev, args)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_items or 200 * cores
        v, dt = cpu_oracle_rate(n_cpu, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d of the %d locus x sample items of the workload, %.1f s" % (n_cpu, LOCI * SAMPLES, dt)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dev_time_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "ploidy": PLOIDY, "n_pos": N_POS, "depth": DEPTH, "samples": SAMPLES,
                       "loci_per_step": loci_per_step, "loci_timed": loci_per_step * K, "chains": CHAINS,
                       "mcmc_steps": MCMC_STEPS, "seed": SEED, "prior": "flat",
                       "l2": "each step reads a different batch and writes a 9.6 GB trace (> L2)",
                       "mean_unique_reads": float(np.mean([b.n_reads().mean() for b, _ in batches]))},
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "call_exact": call_exact, "e2e_posterior": posterior, "e2e_from_calls": from_calls,
            "wall_ms_per_step": 1e3 * wall_max / K,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        barrier()  # rank 0 ran the secondary measurements alone: leave together
        dist.destroy_process_group()


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
