"""Stand-ins for the device entry points of the application layer, backed by the CPU oracle.

TEST INFRASTRUCTURE ONLY.  The application programs (mchap_b200/application/programs.py) call the
device through a handful of batch functions.  On a machine without a GPU the host-side logic around
those calls (readers, read extraction, summaries, VCF formatting) is still checked against the
reference's golden VCFs by swapping those functions for the ones below, which evaluate every item with
the oracle (oracle/mchap_oracle.c).  The `-m gpu` tests run the same scenarios through the real CUDA
path; nothing in the product imports this module.
"""
import numpy as np

from mchap_b200.assemble.classes import GenotypeMultiTrace, TraceTally
from mchap_b200.calling.classes import AllelesTraceTally, GenotypeAllelesMultiTrace
from oracle import oracle as o


def as_probabilistic(calls, n_alleles, probs, error_factor=3):
    """numpy restatement of encoding/integer/transcode.py:16-77 (same assignment order)."""
    calls = np.asarray(calls)
    n_alleles = np.asarray(n_alleles)
    if calls.shape[-1] == 0:
        return np.empty(calls.shape + (0,), dtype=float)
    alleles = np.arange(np.max(n_alleles))
    onehot = calls[..., None] == alleles
    new = ((1 - probs) / error_factor)[..., None] * ~onehot
    hit = probs[..., None] * onehot
    new[onehot] = hit[onehot]
    new[calls < 0] = np.nan
    new[..., n_alleles[..., None] <= alleles] = 0
    return new


def unique_counts(array):
    """mset.py:242-284, 361-392: distinct rows in first-occurrence order (byte-wise) and counts."""
    seen, order, counts = {}, [], []
    for i, row in enumerate(array):
        key = row.tobytes()
        k = seen.get(key)
        if k is None:
            seen[key] = len(order)
            order.append(i)
            counts.append(1)
        else:
            counts[k] += 1
    return array[order], np.array(counts, dtype=np.int64)


def encode_unique_reads_batch(calls_list, probs_list, n_alleles_list, error_factor=3, device=None):
    out = []
    for calls, probs, na in zip(calls_list, probs_list, n_alleles_list):
        calls = np.asarray(calls)
        if calls.shape[0] == 0:
            A = int(np.max(na, initial=0))
            out.append((np.empty((0, calls.shape[1], A)), np.zeros(0, dtype=np.int64)))
            continue
        dists = as_probabilistic(calls, na, np.broadcast_to(np.asarray(probs, dtype=float), calls.shape), error_factor)
        out.append(unique_counts(dists))
    return out


class DenovoMCMC(object):
    def __init__(self, **kw):
        self.kw = kw

    def fit_posterior_from_calls_batch(self, calls_list, probs_list, burn=0, n_alleles_list=None, seeds=None,
                                       max_unique=128, error_factor=3, ploidy_list=None, inbreeding_list=None,
                                       temperatures_list=None, errors="raise"):
        kw = self.kw
        pairs = encode_unique_reads_batch(calls_list, probs_list, n_alleles_list, error_factor)
        out = []
        for i, (reads, counts) in enumerate(pairs):
            res = o.denovo_fit(
                reads, counts, ploidy_list[i], n_alleles_list[i],
                inbreeding=None if inbreeding_list is None else inbreeding_list[i], steps=kw["steps"],
                chains=kw["chains"], alpha=kw["alpha"], beta=kw["beta"], fix_homozygous=kw["fix_homozygous"],
                recombination_step_probability=kw["recombination_step_probability"],
                partial_dosage_step_probability=kw["partial_dosage_step_probability"],
                dosage_step_probability=kw["dosage_step_probability"],
                temperatures=(1.0,) if temperatures_list is None else temperatures_list[i],
                random_seed=kw["random_seed"])
            trace = GenotypeMultiTrace(res["genotypes"], res["llks"])
            out.append(TraceTally.from_trace(trace.burn(burn)))
        return out, np.array([len(c) for _, c in pairs], dtype=np.int64)


class CallingMCMC(object):
    def __init__(self, **kw):
        self.kw = kw

    def fit_posterior_batch(self, reads_list, counts_list=None, burn=0, initial_list=None, haplotypes_list=None,
                            priors=None, seeds=None, max_unique=128, ploidy_list=None, errors="raise"):
        kw = self.kw
        out = []
        for i, reads in enumerate(reads_list):
            res = o.calling_fit(reads, counts_list[i], ploidy_list[i], haplotypes_list[i],
                                prior=None if priors is None else priors[i], steps=kw["steps"], chains=kw["chains"],
                                random_seed=kw["random_seed"])
            trace = GenotypeAllelesMultiTrace(res["genotypes"], res["llks"], len(haplotypes_list[i]))
            out.append(AllelesTraceTally.from_trace(trace.burn(burn)))
        return out


class exact(object):
    """Namespace with the batch functions of mchap_b200.calling.exact."""

    @staticmethod
    def genotype_likelihoods_batch(reads_list, ploidy, haplotypes_list, counts_list=None, device=None):
        ploidy = np.broadcast_to(np.asarray(ploidy), (len(reads_list),))
        return [o.genotype_likelihoods(r, int(p), h, None if counts_list is None else counts_list[i])
                for i, (r, p, h) in enumerate(zip(reads_list, ploidy, haplotypes_list))]

    @staticmethod
    def genotype_posteriors_batch(llks_list, ploidy, n_alleles_list, priors=None, device=None, with_frequencies=False):
        ploidy = np.broadcast_to(np.asarray(ploidy), (len(llks_list),))
        gps = [o.genotype_posteriors(l, int(p), int(h), None if priors is None else priors[i])
               for i, (l, p, h) in enumerate(zip(llks_list, ploidy, n_alleles_list))]
        if not with_frequencies:
            return gps
        return gps, [o.posterior_allele_frequencies(g, int(p), int(h)) for g, p, h in zip(gps, ploidy, n_alleles_list)]

    @staticmethod
    def posterior_mode_batch(reads_list, ploidy, haplotypes_list, counts_list=None, priors=None, device=None):
        ploidy = np.broadcast_to(np.asarray(ploidy), (len(reads_list),))
        out = []
        for i, (r, p, h) in enumerate(zip(reads_list, ploidy, haplotypes_list)):
            out.append(o.posterior_mode(r, int(p), h, None if counts_list is None else counts_list[i],
                                        None if priors is None else priors[i], True, True, True))
        return out

    @staticmethod
    def alternate_dosage_posteriors(genotype_alleles, probabilities):
        from mchap_b200.calling import exact as real

        return real.alternate_dosage_posteriors(genotype_alleles, probabilities)


class Device(object):
    def minimum_error_correction_batch(self, calls_list, genotypes_list, per_read=False):
        mec, called = [], []
        for calls, g in zip(calls_list, genotypes_list):
            calls, g = np.asarray(calls), np.asarray(g)
            diff = (calls[:, None, :] != g[None, :, :]) & (calls[:, None, :] >= 0)
            per = diff.sum(axis=-1).min(axis=-1) if len(calls) else np.zeros(0, dtype=int)
            mec.append(int(per.sum()))
            called.append(int((calls >= 0).sum()))
        return np.array(mec, dtype=np.int64), np.array(called, dtype=np.int64)

    def index_as_genotype_alleles(self, index, ploidy):
        return np.array([o.index_as_genotype_alleles(int(i), int(ploidy)) for i in index], dtype=np.int64)


def install(monkeypatch):
    """Route the application programs through the stand-ins above."""
    from mchap_b200.application import programs

    def model(cls):
        return lambda **kw: cls(**{k: v for k, v in kw.items() if k != "device"})

    monkeypatch.setattr(programs, "DenovoMCMC", model(DenovoMCMC))
    monkeypatch.setattr(programs, "CallingMCMC", model(CallingMCMC))
    monkeypatch.setattr(programs, "exact", exact)
    monkeypatch.setattr(programs, "encode_unique_reads_batch", encode_unique_reads_batch)
    monkeypatch.setattr(programs.program, "_device", lambda self: Device())
