#!/usr/bin/env python3
"""Generate tests/golden/reference_golden.npz by running the REAL reference.

Run in the build container only (needs /root/reference and numba):

    python tests/golden/make_golden.py

The reference (PlantandFoodResearch/MCHap v0.11.1) is imported unmodified from
/root/reference with an empty ``pysam`` stub module on the path (only mchap/io
needs pysam).  Every array saved here is an input to, or an output of, the
reference's own numba functions; nothing from the oracle or the CUDA path is
involved.  The GPU box has no /root/reference, so only the .npz travels.
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MCHAP_REFERENCE", "/root/reference")

_stub = tempfile.mkdtemp(prefix="pysam_stub_")
open(os.path.join(_stub, "pysam.py"), "w").close()
os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "numba_cache_mchap"))
sys.path[:0] = [_stub, REF, ROOT]

import numpy as np  # noqa: E402
import numba  # noqa: E402

import mchap  # noqa: E402
from mchap import jitutils as J  # noqa: E402
from mchap.assemble import likelihood as L  # noqa: E402
from mchap.assemble import mutation, structural, tempering  # noqa: E402
from mchap.assemble import prior as aprior  # noqa: E402
from mchap.assemble import mcmc as amcmc  # noqa: E402
from mchap.assemble.mcmc import DenovoMCMC  # noqa: E402
from mchap.assemble.snpcalling import snp_posterior  # noqa: E402
from mchap.calling import exact, prior as cprior, utils as cutils  # noqa: E402
from mchap.calling import mcmc as cmcmc  # noqa: E402
from mchap.calling.classes import CallingMCMC  # noqa: E402
from mchap import combinatorics  # noqa: E402

from mchap_b200.synth import synth_items, synth_haplotype_panel, encode_calls  # noqa: E402

assert mchap.__version__ == "v0.11.1", mchap.__version__

OUT = {}
META = {"reference_version": mchap.__version__, "numba": numba.__version__, "numpy": np.__version__}


@numba.njit
def _seed(s):
    np.random.seed(s)


@numba.njit
def _next_double():
    return np.random.random()


@numba.njit
def _rng_probe(seed):
    np.random.seed(seed)
    out = np.zeros(64, dtype=np.float64)
    k = 0
    for _ in range(8):
        out[k] = np.random.random()
        k += 1
    for n in (1, 2, 3, 5, 8, 31, 32, 33, 100, 1000, 70000):
        out[k] = np.random.randint(n)
        k += 1
    x = np.arange(10)
    np.random.shuffle(x)
    for i in range(10):
        out[k] = x[i]
        k += 1
    y = np.random.permutation(np.arange(7))
    for i in range(7):
        out[k] = y[i]
        k += 1
    out[k] = np.random.choice(np.array([3, 5, 9, 11]))
    k += 1
    out[k] = np.random.rand()
    k += 1
    z = np.empty((6, 2), dtype=np.int8)
    for i in range(6):
        z[i, 0] = i
        z[i, 1] = 10 + i
    np.random.shuffle(z)
    for i in range(6):
        out[k] = z[i, 0]
        k += 1
    out[k] = np.random.random()
    return out


@numba.njit
def _luh(n_alleles):
    return np.log(n_alleles).sum()


def luh_cases():
    arrs = [[2] * 8, [2, 3, 4, 2], [4] * 16, [2], [3, 3, 3, 2, 2, 4, 4, 2, 2, 2, 3]]
    for k, a in enumerate(arrs):
        a = np.array(a, dtype=np.int8)
        OUT["luh%d_in" % k] = a
        OUT["luh%d_out" % k] = np.array([_luh(a)], dtype=np.float64)
    META["luh_cases"] = len(arrs)


def rng_cases():
    for seed in (0, 11, 42, 123456789):
        OUT["rng_probe_%d" % seed] = _rng_probe(seed)
    META["rng_seeds"] = [0, 11, 42, 123456789]


def random_reads(rng, U, N, A, n_alleles=None, nan_frac=0.2, zero_frac=0.0):
    """Random probabilistic reads with gaps; rows not normalised on purpose (like phred-scaled input)."""
    reads = rng.random((U, N, A))
    reads /= reads.sum(axis=-1, keepdims=True)
    if n_alleles is not None:
        for j, a in enumerate(n_alleles):
            reads[:, j, a:] = 0
    gaps = rng.random((U, N)) < nan_frac
    reads[gaps] = np.nan
    return reads


def likelihood_cases():
    rng = np.random.default_rng(1)
    cases = []
    for k in range(24):
        P = int(rng.choice([2, 3, 4, 6, 8]))
        N = int(rng.integers(1, 9))
        A = int(rng.integers(2, 5))
        U = int(rng.integers(1, 12))
        reads = random_reads(rng, U, N, A)
        if k % 5 == 0:
            reads[0] = np.nan
        if k % 7 == 0:
            reads[-1, 0, 0] = 0.0
        g = rng.integers(0, A, size=(P, N)).astype(np.int8)
        counts = rng.integers(1, 6, size=U).astype(np.int64) if k % 2 else None
        idx = rng.integers(0, P, size=P).astype(np.int8)
        a, b = sorted(rng.integers(0, N + 1, size=2))
        llk = L.log_likelihood(reads, g, read_counts=counts)
        llk_sc = L.log_likelihood_structural_change(reads, g, idx, interval=(int(a), int(b)), read_counts=counts)
        g2 = g.copy()
        J.structural_change(g2, idx, interval=(int(a), int(b)))
        OUT["llk%d_reads" % k] = reads
        OUT["llk%d_genotype" % k] = g
        if counts is not None:
            OUT["llk%d_counts" % k] = counts
        OUT["llk%d_idx" % k] = idx
        OUT["llk%d_interval" % k] = np.array([a, b])
        OUT["llk%d_out" % k] = np.array([llk, llk_sc])
        OUT["llk%d_changed" % k] = g2
        cases.append(k)
    META["llk_cases"] = cases


def jitutils_cases():
    OUT["comb_table"] = np.array(J._COMB_CACHE)
    OUT["cwr_table"] = np.array(J._COMB_WITH_REPLACEMENT_CACHE)
    big = [(120, 5), (150, 7), (1029, 6), (200, 3), (64, 32)]
    OUT["comb_big_in"] = np.array(big)
    OUT["comb_big_out"] = np.array([J.comb(n, k) for n, k in big], dtype=np.int64)
    OUT["cwr_big_out"] = np.array([J.comb_with_replacement(n, k) for n, k in big[:4]], dtype=np.int64)
    for P in (1, 2, 3, 4, 6, 8):
        n = 300
        un = np.zeros((n, P), dtype=np.int64)
        rk = np.zeros(n, dtype=np.int64)
        g = np.zeros(P, dtype=np.int64)
        inc = np.zeros((n, P), dtype=np.int64)
        for i in range(n):
            un[i] = J.index_as_genotype_alleles(i, P)
            rk[i] = J.genotype_alleles_as_index(un[i])
            inc[i] = g
            J.increment_genotype(g)
        OUT["unrank_p%d" % P] = un
        OUT["rank_p%d" % P] = rk
        OUT["increment_p%d" % P] = inc
    idxs = np.array([0, 1, 1715, 1716, 52359, 10 ** 6, 10 ** 9, 1624866254968319], dtype=np.int64)
    OUT["unrank_large_idx"] = idxs
    OUT["unrank_large_p6"] = np.array([J.index_as_genotype_alleles(i, 6) for i in idxs])
    rng = np.random.default_rng(2)
    x = rng.normal(size=40) * 20
    x[3] = -np.inf
    OUT["logspace_in"] = x
    OUT["logspace_sum"] = np.array([J.sum_log_probs(x)])
    OUT["logspace_norm"] = J.normalise_log_probs(x)
    OUT["logspace_add"] = np.array(
        [J.add_log_prob(a, b) for a, b in [(-np.inf, -np.inf), (-1.0, -np.inf), (-3.0, -2.5), (0.0, 0.0)]]
    )
    dos = np.array([[4, 0, 0, 0], [2, 2, 0, 0], [1, 1, 1, 1], [3, 1, 0, 0], [2, 1, 1, 0]], dtype=np.int64)
    OUT["lnperm_in"] = dos
    OUT["lnperm_out"] = np.array([J.ln_equivalent_permutations(d) for d in dos])


def dosage_vectors(P):
    out = []
    rng = np.random.default_rng(P)
    for _ in range(12):
        g = rng.integers(0, 3, size=(P, 3)).astype(np.int8)
        d = np.zeros(P, dtype=np.int8)
        J.get_haplotype_dosage(d, g)
        out.append((g, d))
    return out


def prior_cases():
    k = 0
    for P in (2, 4, 6):
        for g, d in dosage_vectors(P):
            for inb in (0.0, 0.1, 0.5):
                for luh in (np.log(8.0), np.log(256.0)):
                    OUT["aprior%d_genotype" % k] = g
                    OUT["aprior%d_dosage" % k] = d
                    OUT["aprior%d_par" % k] = np.array([luh, inb])
                    OUT["aprior%d_out" % k] = np.array([aprior.log_genotype_prior(d, luh, inb)])
                    k += 1
    META["aprior_cases"] = k
    rng = np.random.default_rng(5)
    k = 0
    for P, H in ((2, 3), (4, 4), (6, 8), (4, 32)):
        for _ in range(6):
            g = np.sort(rng.integers(0, H, size=P)).astype(np.int64)
            freqs = rng.random(H)
            freqs /= freqs.sum()
            for inb in (0.0, 0.2):
                for f in (None, freqs):
                    va = int(rng.integers(0, P))
                    OUT["cprior%d_genotype" % k] = g
                    OUT["cprior%d_par" % k] = np.array([H, inb, va, 0 if f is None else 1])
                    if f is not None:
                        OUT["cprior%d_freqs" % k] = f
                    OUT["cprior%d_out" % k] = np.array([
                        cprior.log_genotype_prior(g, H, inbreeding=inb, frequencies=f),
                        cprior.log_genotype_allele_prior(g, va, H, inbreeding=inb, frequencies=f),
                        cprior.log_genotype_allele_flat_prior(g, va),
                    ])
                    OUT["cprior%d_dosage" % k] = cutils.allelic_dosage(g)
                    k += 1
    META["cprior_cases"] = k


def structural_cases():
    rng = np.random.default_rng(7)
    k = 0
    for P in (2, 3, 4, 6, 8):
        for _ in range(8):
            N = int(rng.integers(2, 9))
            pool = rng.integers(0, 2, size=(max(2, P // 2), N)).astype(np.int8)
            g = pool[rng.integers(0, len(pool), size=P)].copy()
            flip = rng.random((P, N)) < 0.15
            g = np.where(flip, 1 - g, g).astype(np.int8)
            a, b = sorted(rng.integers(0, N + 1, size=2))
            interval = (int(a), int(b))
            labels = structural.haplotype_segment_labels(g, interval)
            labels_none = structural.haplotype_segment_labels(g, None)
            ro = structural.recombination_step_options(labels)
            do = structural.dosage_step_options(labels)
            OUT["struct%d_genotype" % k] = g
            OUT["struct%d_interval" % k] = np.array(interval)
            OUT["struct%d_labels" % k] = labels
            OUT["struct%d_labels_none" % k] = labels_none
            OUT["struct%d_recomb" % k] = ro
            OUT["struct%d_dosage" % k] = do
            OUT["struct%d_n" % k] = np.array([
                structural.recombination_step_n_options(labels),
                structural.dosage_step_n_options(labels),
            ])
            OUT["struct%d_recomb_return" % k] = np.array(
                [structural.recombination_step_n_options(o) for o in ro], dtype=np.int64)
            OUT["struct%d_dosage_return" % k] = np.array(
                [structural.dosage_step_n_options(o) for o in do], dtype=np.int64)
            k += 1
    META["struct_cases"] = k
    k = 0
    for seed, breaks, n in ((1, 0, 5), (2, 1, 5), (3, 3, 8), (4, 7, 8), (5, 4, 16), (6, 2, 3), (7, 1, 2)):
        _seed(seed)
        iv = structural.random_breaks(breaks, n)
        OUT["breaks%d_par" % k] = np.array([seed, breaks, n])
        OUT["breaks%d_out" % k] = iv
        OUT["breaks%d_next" % k] = np.array([_next_double()])
        k += 1
    META["breaks_cases"] = k


def small_item(rng, P, N, A, depth, nan=True):
    """One simulated locus x sample in the reference encoding (ragged allele counts allowed)."""
    n_alleles = rng.integers(2, A + 1, size=N).astype(np.int8) if A > 2 else np.full(N, 2, dtype=np.int8)
    haps = np.stack([rng.integers(0, n_alleles[j], size=P) for j in range(N)], axis=1).astype(np.int8)
    pick = rng.integers(0, P, size=depth)
    calls = haps[pick].copy()
    flips = rng.random(calls.shape) < 0.02
    calls = np.where(flips, (calls + 1) % n_alleles[None, :], calls).astype(np.int8)
    if nan:
        for r in range(depth):
            ln = int(rng.integers(max(1, N // 2), N + 1))
            st = int(rng.integers(0, N - ln + 1))
            calls[r, :st] = -1
            calls[r, st + ln:] = -1
    reads = encode_calls(calls, np.broadcast_to(n_alleles, calls.shape), int(n_alleles.max()))
    from mchap import mset
    ur, uc = mset.unique_counts(reads)
    return ur, uc.astype(np.int64), n_alleles, haps


def step_cases():
    rng = np.random.default_rng(11)
    k = 0
    for P, N, A in ((2, 3, 2), (4, 5, 2), (4, 4, 4), (6, 6, 3), (8, 5, 2), (4, 8, 2)):
        for inb in (None, 0.0, 0.3):
            reads, counts, n_alleles, haps = small_item(rng, P, N, A, 14)
            use_counts = k % 2 == 0
            rc = counts if use_counts else None
            g0 = haps.copy()
            g0[rng.integers(0, P), rng.integers(0, N)] = 0
            temp = [1.0, 0.5, 0.1][k % 3]
            luh = np.log(n_alleles.astype(np.float64)).sum()
            llk0 = L.log_likelihood(reads, g0, read_counts=rc)
            seed = 100 + k
            OUT["step%d_reads" % k] = reads
            OUT["step%d_counts" % k] = counts
            OUT["step%d_nalleles" % k] = n_alleles
            OUT["step%d_g0" % k] = g0
            OUT["step%d_par" % k] = np.array([
                P, N, int(n_alleles.max()), -1.0 if inb is None else inb, temp, int(use_counts), seed, luh, llk0])
            # base step
            g = g0.copy()
            _seed(seed)
            h, j = int(rng.integers(0, P)), int(rng.integers(0, N))
            llk, _ = mutation.base_step(g, reads, llk0, h, j, int(n_alleles[j]), luh, inbreeding=inb,
                                        temp=temp, read_counts=rc, cache=None)
            OUT["step%d_base" % k] = g
            OUT["step%d_base_out" % k] = np.array([h, j, llk, _next_double()])
            # mutation compound step
            g = g0.copy()
            _seed(seed)
            llk, _ = mutation.compound_step(g, reads, llk0, n_alleles, luh, inbreeding=inb, temp=temp,
                                            read_counts=rc, cache=None)
            OUT["step%d_mut" % k] = g
            OUT["step%d_mut_out" % k] = np.array([llk, _next_double()])
            # interval steps
            a, b = sorted(rng.integers(0, N + 1, size=2))
            if a == b:
                a, b = 0, max(1, N // 2)
            for st, name in ((0, "recomb"), (1, "dosage")):
                g = g0.copy()
                _seed(seed)
                llk, _ = structural.interval_step(g, reads, llk0, luh, inbreeding=inb, interval=(int(a), int(b)),
                                                  step_type=st, temp=temp, read_counts=rc, cache=None)
                OUT["step%d_%s" % (k, name)] = g
                OUT["step%d_%s_out" % (k, name)] = np.array([a, b, llk, _next_double()])
            # structural compound step over random breaks
            for st, name in ((0, "crecomb"), (1, "cdosage")):
                g = g0.copy()
                _seed(seed)
                nb = min(2, N - 1)
                iv = structural.random_breaks(nb, N)
                llk, _ = structural.compound_step(g, reads, llk0, iv, luh, inbreeding=inb, step_type=st,
                                                  temp=temp, read_counts=rc, cache=None)
                OUT["step%d_%s" % (k, name)] = g
                OUT["step%d_%s_iv" % (k, name)] = iv
                OUT["step%d_%s_out" % (k, name)] = np.array([llk, _next_double()])
            k += 1
    META["step_cases"] = k


def snp_cases():
    rng = np.random.default_rng(13)
    k = 0
    for P, N, A in ((2, 4, 2), (4, 6, 2), (4, 5, 4), (6, 4, 3)):
        for inb in (None, 0.0, 0.2):
            reads, counts, n_alleles, haps = small_item(rng, P, N, A, 20)
            OUT["snp%d_reads" % k] = reads
            OUT["snp%d_counts" % k] = counts
            OUT["snp%d_nalleles" % k] = n_alleles
            OUT["snp%d_par" % k] = np.array([P, -1.0 if inb is None else inb])
            OUT["snp%d_hom" % k] = amcmc._homozygosity_probabilities(
                reads, n_alleles, P, inbreeding=inb, read_counts=counts)
            _, probs = snp_posterior(reads[:, 0, :], int(n_alleles[0]), P, inb, read_counts=counts)
            OUT["snp%d_post0" % k] = probs
            OUT["snp%d_meandist" % k] = amcmc._read_mean_dist(reads)
            k += 1
    # all-gap column and zero reads
    reads = np.full((3, 2, 2), np.nan)
    reads[:, 0] = [0.7, 0.3]
    OUT["snp%d_reads" % k] = reads
    OUT["snp%d_counts" % k] = np.ones(3, dtype=np.int64)
    OUT["snp%d_nalleles" % k] = np.array([2, 2], dtype=np.int8)
    OUT["snp%d_par" % k] = np.array([4, -1.0])
    OUT["snp%d_hom" % k] = amcmc._homozygosity_probabilities(
        reads, np.array([2, 2], dtype=np.int8), 4, inbreeding=None, read_counts=np.ones(3, dtype=np.int64))
    _, probs = snp_posterior(reads[:, 0, :], 2, 4, None, read_counts=np.ones(3, dtype=np.int64))
    OUT["snp%d_post0" % k] = probs
    OUT["snp%d_meandist" % k] = amcmc._read_mean_dist(reads)
    k += 1
    META["snp_cases"] = k


def fit_cases():
    """DenovoMCMC(...).fit on small items: full traces (unsorted, as returned by _mcmc)."""
    rng = np.random.default_rng(17)
    specs = [
        # P, N, A, depth, inbreeding, temps, steps, chains, seed, fix_hom, probs, n_intervals
        (4, 8, 2, 40, None, (1.0,), 80, 2, 11, 0.999, (0.5, 0.5, 1.0), None),
        (4, 8, 2, 40, 0.0, (1.0,), 60, 2, 42, 0.999, (0.5, 0.5, 1.0), None),
        (4, 8, 2, 40, 0.25, (1.0,), 60, 2, 7, 0.999, (0.5, 0.5, 1.0), None),
        (2, 5, 2, 20, None, (1.0,), 60, 2, 3, 0.999, (0.5, 0.5, 1.0), None),
        (6, 6, 2, 30, None, (0.1, 0.5, 1.0), 40, 2, 5, 0.999, (0.5, 0.5, 1.0), None),
        (8, 10, 2, 60, 0.1, (0.01, 0.1, 0.5, 1.0), 25, 2, 9, 0.999, (0.5, 0.5, 1.0), None),
        (4, 6, 4, 30, None, (1.0,), 60, 2, 13, 0.999, (1.0, 1.0, 1.0), None),
        (4, 6, 3, 30, 0.2, (0.3, 1.0), 50, 1, 21, 2.0, (0.5, 0.5, 1.0), None),
        (4, 8, 2, 40, None, (1.0,), 50, 2, 31, 0.999, (0.5, 0.5, 1.0), 3),
        (3, 4, 2, 12, None, (1.0,), 50, 3, 77, 0.9, (0.0, 0.0, 0.0), None),
        (4, 16, 2, 100, None, (0.25, 1.0), 20, 2, 101, 0.999, (0.5, 0.5, 1.0), None),
        (4, 1, 2, 10, None, (1.0,), 30, 2, 5, 2.0, (0.5, 0.5, 1.0), None),
    ]
    for k, (P, N, A, depth, inb, temps, steps, chains, seed, fix, pr, nint) in enumerate(specs):
        reads, counts, n_alleles, haps = small_item(rng, P, N, A, depth)
        model = DenovoMCMC(ploidy=P, n_alleles=list(n_alleles), inbreeding=inb, steps=steps, chains=chains,
                           n_intervals=nint, fix_homozygous=fix, recombination_step_probability=pr[0],
                           partial_dosage_step_probability=pr[1], dosage_step_probability=pr[2],
                           temperatures=temps, random_seed=seed)
        # mirror fit() without the GenotypeMultiTrace sorting: raw _mcmc traces
        np.random.seed(seed)
        J.seed_numba(seed)
        gens, llks = [], []
        for _ in range(chains):
            g, l = model._mcmc(reads, read_counts=counts, initial=None)
            gens.append(g)
            llks.append(l)
        nxt = _next_double()
        trace = model.fit(reads, read_counts=counts)
        OUT["fit%d_reads" % k] = reads
        OUT["fit%d_counts" % k] = counts
        OUT["fit%d_nalleles" % k] = n_alleles
        OUT["fit%d_genotypes" % k] = np.array(gens).astype(np.int8)
        OUT["fit%d_llks" % k] = np.array(llks)
        OUT["fit%d_sorted" % k] = trace.genotypes.astype(np.int8)
        OUT["fit%d_next" % k] = np.array([nxt])
        hom = amcmc._homozygosity_probabilities(reads, n_alleles, P, inbreeding=inb, read_counts=counts)
        OUT["fit%d_nhet" % k] = np.array([int((~np.any(hom >= fix, axis=-1)).sum())])
        META["fit%d" % k] = dict(P=P, N=N, A=A, inbreeding=inb, temperatures=list(temps), steps=steps,
                                 chains=chains, seed=seed, fix_homozygous=fix, probs=list(pr),
                                 n_intervals=nint)
    # edge cases (tests/test_assemble/test_mcmc.py:95-169)
    k = len(specs)
    edge = []
    reads = np.empty((0, 3, 2))
    edge.append((reads, None, [2, 2, 2], 4))
    reads = np.full((4, 3, 2), np.nan)
    edge.append((reads, None, [2, 2, 2], 4))
    r = encode_calls(np.array([[0, 0, 0]] * 30, dtype=np.int8), np.full((30, 3), 2), 2)
    edge.append((r, None, [2, 2, 2], 2))  # everything fixed homozygous
    for reads, counts, na, P in edge:
        model = DenovoMCMC(ploidy=P, n_alleles=na, steps=20, chains=2, random_seed=4)
        trace = model.fit(reads, read_counts=counts)
        OUT["fit%d_reads" % k] = reads
        OUT["fit%d_nalleles" % k] = np.array(na, dtype=np.int8)
        OUT["fit%d_sorted" % k] = trace.genotypes.astype(np.int8)
        OUT["fit%d_llks" % k] = trace.llks
        META["fit%d" % k] = dict(P=P, N=3, A=2, inbreeding=None, temperatures=[1.0], steps=20, chains=2,
                                 seed=4, fix_homozygous=0.999, probs=[0.5, 0.5, 1.0], n_intervals=None,
                                 edge=True)
        k += 1
    META["fit_cases"] = k


def calling_cases():
    k = 0
    for P, H, N, inb, with_freqs, st in (
        (4, 8, 6, None, False, "Gibbs"),
        (4, 8, 6, 0.1, False, "Gibbs"),
        (4, 8, 6, 0.1, True, "Gibbs"),
        (4, 8, 6, 0.0, True, "Gibbs"),
        (2, 5, 4, None, False, "Metropolis-Hastings"),
        (4, 6, 5, 0.2, True, "Metropolis-Hastings"),
        (6, 8, 8, None, False, "Gibbs"),
        (4, 32, 8, None, False, "Gibbs"),
    ):
        batch, panels, truth = synth_haplotype_panel(1, H, N, P, depth=30, seed=50 + k)
        reads, counts = batch.item(0)
        haps = panels[0]
        rng = np.random.default_rng(k)
        freqs = rng.random(H) + 0.2
        freqs /= freqs.sum()
        prior = None if inb is None else (inb, freqs if with_freqs else None)
        seed = 200 + k
        steps, chains = 40, 2
        model = CallingMCMC(ploidy=P, haplotypes=haps, prior=prior, steps=steps, chains=chains,
                            random_seed=seed, step_type=st)
        trace = model.fit(reads, read_counts=counts)
        nxt = _next_double()
        greedy = cmcmc.greedy_caller(haps, P, reads, counts, prior=prior)
        OUT["call%d_reads" % k] = reads
        OUT["call%d_counts" % k] = counts
        OUT["call%d_haplotypes" % k] = haps
        if prior is not None and prior[1] is not None:
            OUT["call%d_freqs" % k] = freqs
        OUT["call%d_genotypes" % k] = trace.genotypes.astype(np.int64)
        OUT["call%d_llks" % k] = trace.llks
        OUT["call%d_greedy" % k] = greedy.astype(np.int64)
        OUT["call%d_next" % k] = np.array([nxt])
        # one sub-step option table of each kind
        g = greedy.astype(np.int64).copy()
        H_ = len(haps)
        for stype, name in ((0, "gibbs"), (1, "mh")):
            llks = np.full(H_, np.nan)
            lpr = np.full(H_, np.nan)
            pr = np.full(H_, np.nan)
            fn = cmcmc.gibbs_options if stype == 0 else cmcmc.mh_options
            fn(g.copy(), 1, haps, reads, counts, llks, lpr, pr, prior=prior, llk_cache=None)
            OUT["call%d_%s" % (k, name)] = np.stack([llks, lpr, pr])
        META["call%d" % k] = dict(P=P, H=H, N=N, inbreeding=inb, with_freqs=with_freqs, step_type=st,
                                  seed=seed, steps=steps, chains=chains)
        k += 1
    META["call_cases"] = k


def exact_cases():
    k = 0
    for P, H, N, inb, with_freqs in (
        (2, 4, 4, None, False),
        (4, 4, 5, None, False),
        (4, 4, 5, 0.1, False),
        (4, 5, 5, 0.3, True),
        (6, 8, 8, None, False),
        (6, 8, 8, 0.1, False),
        (3, 6, 6, 0.0, True),
    ):
        batch, panels, truth = synth_haplotype_panel(1, H, N, P, depth=24, seed=300 + k)
        reads, counts = batch.item(0)
        haps = panels[0]
        rng = np.random.default_rng(k + 40)
        freqs = rng.random(H) + 0.2
        freqs /= freqs.sum()
        prior = None if inb is None else (inb, freqs if with_freqs else None)
        res = exact.posterior_mode(reads, P, haps, read_counts=counts, prior=prior, return_support_prob=True,
                                   return_posterior_frequencies=True, return_posterior_occurrence=True)
        gl = exact.genotype_likelihoods(reads, P, haps, read_counts=counts)
        gp = exact.genotype_posteriors(gl, P, H, prior=prior)
        fr = exact.posterior_allele_frequencies(gp, P, H)
        alt_g, alt_p = exact.alternate_dosage_posteriors(res[0], gp)
        OUT["exact%d_reads" % k] = reads
        OUT["exact%d_counts" % k] = counts
        OUT["exact%d_haplotypes" % k] = haps
        if prior is not None and prior[1] is not None:
            OUT["exact%d_freqs" % k] = freqs
        OUT["exact%d_mode" % k] = np.asarray(res[0], dtype=np.int64)
        OUT["exact%d_scalars" % k] = np.array([res[1], res[2], res[3]])
        OUT["exact%d_mode_freqs" % k] = res[4]
        OUT["exact%d_mode_occur" % k] = res[5]
        OUT["exact%d_gl" % k] = gl
        OUT["exact%d_gp" % k] = gp
        OUT["exact%d_fr" % k] = np.stack(fr)
        OUT["exact%d_alt_g" % k] = alt_g
        OUT["exact%d_alt_p" % k] = alt_p
        OUT["exact%d_ngen" % k] = np.array([combinatorics.count_unique_genotypes(H, P)])
        META["exact%d" % k] = dict(P=P, H=H, N=N, inbreeding=inb, with_freqs=with_freqs)
        k += 1
    META["exact_cases"] = k


def main():
    rng_cases()
    luh_cases()
    likelihood_cases()
    jitutils_cases()
    prior_cases()
    structural_cases()
    step_cases()
    snp_cases()
    fit_cases()
    calling_cases()
    exact_cases()
    OUT["meta_json"] = np.frombuffer(json.dumps(META).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **OUT)
    print("wrote", path, os.path.getsize(path), "bytes,", len(OUT), "arrays")


if __name__ == "__main__":
    main()
