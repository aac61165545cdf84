#!/usr/bin/env python3
"""Generate tests/golden/reference_trace_classes.npz by running the REAL reference.

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_golden_traces.py

Fixtures for the trace post-processing of SURVEY.md section 8(f) N1: the reference's
``GenotypeMultiTrace`` (mchap/assemble/classes.py:247-376) applied to traces sampled by the
reference's own ``DenovoMCMC.fit`` and to synthetic traces with many ties — per-step haplotype
sort, ``burn``, merged and per-chain ``posterior()``, ``mode_genotype_support()`` and
``replicate_incongruence()``.  The reference is imported unmodified from /root/reference with an
empty ``pysam`` stub (only mchap/io needs pysam); nothing from the oracle or the CUDA path is
involved.  The GPU box has no /root/reference, so only the .npz travels.
"""
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

import mchap  # noqa: E402
from mchap.assemble.mcmc import DenovoMCMC, _point_beta_probabilities  # noqa: E402
from mchap.assemble.classes import GenotypeMultiTrace  # noqa: E402
from mchap.calling.classes import GenotypeAllelesMultiTrace  # noqa: E402
from mchap.encoding.integer import as_probabilistic  # noqa: E402
from mchap import mset  # noqa: E402

from mchap_b200.synth import synth_items  # noqa: E402

assert mchap.__version__ == "v0.11.1", mchap.__version__

OUT = {}


def record(name, raw_genotypes, llks, burn):
    """raw_genotypes: int8[C,S,P,N] as a sampler writes them (haplotypes unsorted)."""
    trace = GenotypeMultiTrace(raw_genotypes, llks)
    OUT[name + "_raw"] = np.asarray(raw_genotypes, dtype=np.int8)
    OUT[name + "_llks"] = np.asarray(llks, dtype=np.float64)
    OUT[name + "_sorted"] = trace.genotypes.astype(np.int8)
    OUT[name + "_burn"] = np.int64(burn)
    burnt = trace.burn(burn)
    post = burnt.posterior()
    OUT[name + "_post_genotypes"] = post.genotypes.astype(np.int8)
    OUT[name + "_post_probs"] = post.probabilities.astype(np.float64)
    mode, prob = post.mode()
    OUT[name + "_mode"] = mode.astype(np.int8)
    OUT[name + "_mode_prob"] = np.float64(prob)
    sup = post.mode_genotype_support()
    OUT[name + "_support_genotypes"] = sup.genotypes.astype(np.int8)
    OUT[name + "_support_probs"] = sup.probabilities.astype(np.float64)
    for c, chain in enumerate(burnt.split()):
        cp = chain.posterior()
        OUT[name + "_chain%d_genotypes" % c] = cp.genotypes.astype(np.int8)
        OUT[name + "_chain%d_probs" % c] = cp.probabilities.astype(np.float64)
    OUT[name + "_incongruence"] = np.array(
        [burnt.replicate_incongruence(threshold=t) for t in (0.6, 0.3, 0.05)], dtype=np.int64)


def sampled_cases():
    k = 0
    for ploidy, n_pos, depth, steps, chains, seed in [
        (4, 6, 12, 300, 2, 1), (4, 6, 6, 300, 2, 2), (4, 8, 40, 200, 2, 3), (2, 5, 8, 250, 3, 4),
        (6, 5, 10, 200, 2, 5), (4, 4, 3, 400, 2, 6),
    ]:
        batch = synth_items(2, ploidy=ploidy, n_pos=n_pos, depth=depth, seed=100 + seed)
        for i in range(2):
            reads, counts = batch.item(i)
            model = DenovoMCMC(ploidy=ploidy, n_alleles=[2] * n_pos, steps=steps, chains=chains,
                               fix_homozygous=1.0, random_seed=seed * 7 + i)
            trace = model.fit(reads, read_counts=counts)
            # un-sort the haplotypes again with a seeded permutation per step, so that the fixture
            # exercises the per-step sort like a sampler's raw output does
            rng = np.random.default_rng(seed * 1000 + i)
            raw = trace.genotypes.copy()
            for c in range(raw.shape[0]):
                for s in range(raw.shape[1]):
                    raw[c, s] = raw[c, s][rng.permutation(ploidy)]
            record("sampled%d" % k, raw, trace.llks, burn=steps // 4)
            k += 1
    OUT["n_sampled"] = np.int64(k)


def synthetic_cases():
    """Random walks over few states: many ties in the counts, chains that disagree, wide alleles."""
    k = 0
    for ploidy, n_pos, n_states, steps, chains, amax, seed in [
        (4, 5, 3, 60, 2, 2, 1), (4, 9, 6, 90, 2, 3, 2), (2, 3, 2, 40, 4, 2, 3), (6, 12, 10, 120, 2, 4, 4),
        (3, 20, 5, 80, 2, 2, 5), (4, 5, 40, 64, 2, 2, 6), (8, 7, 4, 50, 2, 2, 7), (4, 1, 3, 30, 2, 3, 8),
    ]:
        rng = np.random.default_rng(seed)
        states = rng.integers(0, amax, size=(n_states, ploidy, n_pos)).astype(np.int8)
        raw = np.zeros((chains, steps, ploidy, n_pos), dtype=np.int8)
        for c in range(chains):
            cur = int(rng.integers(n_states))
            for s in range(steps):
                if rng.random() < 0.3:
                    # chains prefer different states so that replicate_incongruence has work to do
                    cur = int(rng.integers(n_states)) if rng.random() < 0.5 else (c % n_states)
                raw[c, s] = states[cur][rng.permutation(ploidy)]
        llks = rng.normal(size=(chains, steps))
        record("synthetic%d" % k, raw, llks, burn=int(rng.integers(0, steps // 3)))
        k += 1
    OUT["n_synthetic"] = np.int64(k)


def calling_cases():
    """Calling traces int[C,S,P] (alleles sorted per step like calling/mcmc.py:325-326 leaves them):
    GenotypeAllelesMultiTrace.burn / posterior / mode(genotype_support) / split /
    replicate_incongruence / posterior_frequencies / relabel (calling/classes.py:147-297)."""
    k = 0
    for ploidy, n_allele, n_states, steps, chains, seed in [
        (4, 8, 3, 60, 2, 11), (4, 32, 6, 120, 2, 12), (2, 5, 4, 50, 3, 13), (6, 16, 12, 150, 2, 14),
        (4, 200, 5, 80, 2, 15), (4, 6, 30, 90, 2, 16), (8, 10, 4, 40, 2, 17), (1, 4, 3, 30, 2, 18),
    ]:
        rng = np.random.default_rng(seed)
        states = np.sort(rng.integers(0, n_allele, size=(n_states, ploidy)), axis=1)
        gen = np.zeros((chains, steps, ploidy), dtype=np.int64)
        for c in range(chains):
            cur = int(rng.integers(n_states))
            for s in range(steps):
                if rng.random() < 0.3:
                    cur = int(rng.integers(n_states)) if rng.random() < 0.5 else (c % n_states)
                gen[c, s] = states[cur]
        llks = rng.normal(size=(chains, steps))
        burn = int(rng.integers(0, steps // 3))
        name = "calling%d" % k
        trace = GenotypeAllelesMultiTrace(gen, llks, n_allele)
        OUT[name + "_genotypes"] = gen.astype(np.int32)
        OUT[name + "_llks"] = llks
        OUT[name + "_n_allele"] = np.int64(n_allele)
        OUT[name + "_burn"] = np.int64(burn)
        burnt = trace.burn(burn)
        post = burnt.posterior()
        OUT[name + "_post_genotypes"] = post.genotypes.astype(np.int64)
        OUT[name + "_post_probs"] = post.probabilities
        alleles, gp, sp = post.mode(genotype_support=True)
        OUT[name + "_mode"] = np.asarray(alleles, dtype=np.int64)
        OUT[name + "_mode_probs"] = np.array([gp, sp], dtype=np.float64)
        for c, chain in enumerate(burnt.split()):
            cp = chain.posterior()
            OUT[name + "_chain%d_genotypes" % c] = cp.genotypes.astype(np.int64)
            OUT[name + "_chain%d_probs" % c] = cp.probabilities
        OUT[name + "_incongruence"] = np.array(
            [burnt.replicate_incongruence(threshold=t) for t in (0.6, 0.3, 0.05)], dtype=np.int64)
        fr, cn, oc = burnt.posterior_frequencies()
        OUT[name + "_freqs"] = np.stack([fr, cn, oc])
        labels = np.sort(rng.choice(3 * n_allele, size=n_allele, replace=False))
        rel = burnt.relabel(labels)
        OUT[name + "_labels"] = labels.astype(np.int64)
        rp = rel.posterior()
        OUT[name + "_relabel_post_genotypes"] = rp.genotypes.astype(np.int64)
        OUT[name + "_relabel_post_probs"] = rp.probabilities
        OUT[name + "_relabel_freqs"] = np.stack(rel.posterior_frequencies())
        k += 1
    OUT["n_calling"] = np.int64(k)


def host_tables():
    """Host-side tables the device call takes as inputs: the break-point distributions of
    _point_beta_probabilities (assemble/mcmc.py:429-452) for every number of variable positions."""
    k = 0
    for a, b in [(1.0, 3.0), (1.0, 1.0), (2.5, 0.7)]:
        for n in list(range(1, 18)) + [32, 64, 255]:
            OUT["beta%d_par" % k] = np.array([n, a, b], dtype=np.float64)
            OUT["beta%d_out" % k] = _point_beta_probabilities(n, a, b)
            k += 1
    OUT["n_beta"] = np.int64(k)


def encoding_cases():
    """as_probabilistic (encoding/integer/transcode.py:16-77) + mset.unique_counts (mset.py:361-392)
    the way application/baseclass.py:194-209 chains them (probabilities from phred qualities and an
    extra error rate like io/bam.py:281-286)."""
    k = 0
    for n_reads, n_pos, amax, gap_rate, error_rate, n_quals, seed in [
        (40, 8, 2, 0.3, 0.0024, 1, 21), (60, 8, 2, 0.5, 0.0, 3, 22), (25, 5, 4, 0.2, 0.01, 4, 23),
        (0, 6, 2, 0.0, 0.0, 1, 24), (30, 1, 3, 0.1, 0.0024, 2, 25), (200, 12, 2, 0.6, 0.0024, 2, 26),
        (12, 3, 2, 1.0, 0.0, 1, 27), (50, 16, 3, 0.4, 0.001, 40, 28),
    ]:
        rng = np.random.default_rng(seed)
        n_alleles = rng.integers(2, amax + 1, size=n_pos) if amax > 2 else np.full(n_pos, 2)
        haps = np.stack([rng.integers(0, n_alleles) for _ in range(4)])
        calls = haps[rng.integers(0, 4, size=n_reads)] if n_reads else np.zeros((0, n_pos), dtype=int)
        calls = np.array(calls, dtype=np.int8)
        # some calls beyond the position's allele constraint, some gaps
        flip = rng.random(calls.shape) < 0.03
        calls[flip] = rng.integers(0, amax + 1, size=int(flip.sum()))
        calls[rng.random(calls.shape) < gap_rate] = -1
        quals = rng.choice(rng.integers(2, 42, size=n_quals), size=calls.shape)
        probs = np.ones(calls.shape, dtype=float) * (1 - error_rate)
        probs *= 1 - (10 ** (quals / -10))            # io/util.py prob_of_qual
        dists = as_probabilistic(calls, n_alleles, probs)
        uniq, counts = mset.unique_counts(dists)
        name = "encode%d" % k
        OUT[name + "_calls"] = calls
        OUT[name + "_quals"] = quals.astype(np.int64)
        OUT[name + "_error_rate"] = np.float64(error_rate)
        OUT[name + "_probs"] = probs
        OUT[name + "_n_alleles"] = n_alleles.astype(np.int8)
        OUT[name + "_dists"] = dists
        OUT[name + "_unique"] = uniq
        OUT[name + "_counts"] = np.asarray(counts, dtype=np.int64)
        k += 1
    OUT["n_encode"] = np.int64(k)


if __name__ == "__main__":
    sampled_cases()
    synthetic_cases()
    calling_cases()
    host_tables()
    encoding_cases()
    path = os.path.join(HERE, "reference_trace_classes.npz")
    np.savez_compressed(path, **OUT)
    print("wrote %s: %d arrays, %.1f KB" % (path, len(OUT), os.path.getsize(path) / 1024))
