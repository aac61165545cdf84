"""Exhaustive genotype calling on the GPU — same function names, arguments and return values as
the reference's ``mchap/calling/exact.py`` (posterior_mode 156-249, genotype_likelihoods 266-292,
genotype_posteriors 295-329, posterior_allele_frequencies 332-369, alternate_dosage_posteriors
372-407), each with a ``*_batch`` form that sends many (locus, sample) items in one device call.
"""
from itertools import combinations_with_replacement

import numpy as np

from ..api import CALL_ITEM_DTYPE, CallBatch, count_genotypes, default_device

__all__ = [
    "posterior_mode", "posterior_mode_batch", "genotype_likelihoods", "genotype_likelihoods_batch",
    "genotype_posteriors", "posterior_allele_frequencies", "alternate_dosage_posteriors",
]


def posterior_mode_batch(reads_list, ploidy, haplotypes_list, counts_list=None, priors=None, device=None):
    """posterior_mode for many items -> list of (alleles, llk, probability, support probability,
    mean allele frequencies, allele occurrence)."""
    dev = device or default_device()
    batch = CallBatch(reads_list, haplotypes_list, ploidy, counts_list, priors)
    out = dev.call_exact_mode(batch)
    res = []
    for i in range(batch.n):
        it = batch.items[i]
        P, H, o = int(it["ploidy"]), int(it["n_haps"]), int(it["hap_out_off"])
        st = out["stats"][i]
        res.append((out["alleles"][i, :P].copy(), float(st[0]), float(st[1]), float(st[2]),
                    out["freqs"][o:o + H].copy(), out["occur"][o:o + H].copy()))
    return res


def posterior_mode(reads, ploidy, haplotypes, read_counts=None, prior=None, return_support_prob=False,
                   return_posterior_frequencies=False, return_posterior_occurrence=False, device=None):
    """Call posterior mode genotype with statistics from a set of known haplotypes
    (reference signature; the optional results are selected by the same flags)."""
    alleles, llk, prob, support, freqs, occur = posterior_mode_batch(
        [reads], ploidy, [haplotypes], None if read_counts is None else [read_counts],
        None if prior is None else [prior], device)[0]
    result = [alleles, llk, prob]
    if return_support_prob:
        result.append(support)
    if return_posterior_frequencies:
        result.append(freqs)
    if return_posterior_occurrence:
        result.append(occur)
    return tuple(result)


def genotype_likelihoods_batch(reads_list, ploidy, haplotypes_list, counts_list=None, device=None):
    dev = device or default_device()
    batch = CallBatch(reads_list, haplotypes_list, ploidy, counts_list, None)
    gl = dev.genotype_likelihoods(batch)
    return [gl[int(it["gl_off"]): int(it["gl_off"]) + int(g)] for it, g in zip(batch.items, batch.n_genotypes)]


def genotype_likelihoods(reads, ploidy, haplotypes, read_counts=None, device=None):
    """float32 log likelihood of every genotype in VCF order (reference dtype, exact.py:254)."""
    return genotype_likelihoods_batch([reads], ploidy, [haplotypes],
                                      None if read_counts is None else [read_counts], device)[0]


def _posterior_items(n_genotypes, ploidy, n_alleles, prior):
    items = np.zeros(1, dtype=CALL_ITEM_DTYPE)
    it = items[0]
    it["ploidy"], it["n_haps"] = int(ploidy), int(n_alleles)
    it["freqs_off"] = -1
    it["inbreeding"] = np.nan
    freqs = None
    if prior is not None:
        it["inbreeding"] = float(prior[0])
        if prior[1] is not None:
            freqs = np.ascontiguousarray(prior[1], dtype=np.float64)
            it["freqs_off"] = 0
    assert count_genotypes(n_alleles, ploidy) == n_genotypes
    return items, freqs


def genotype_posteriors_batch(llks_list, ploidy, n_alleles_list, priors=None, device=None, with_frequencies=False):
    """genotype_posteriors for many items in one device call (``ploidy`` scalar or per item):
    list of f64[G] arrays; with_frequencies=True also returns the per-item (frequencies, counts,
    occurrence) triples of posterior_allele_frequencies computed by the same kernel."""
    dev = device or default_device()
    n = len(llks_list)
    ploidies = np.broadcast_to(np.asarray(ploidy, dtype=np.int64), (n,))
    ls = [np.ascontiguousarray(l) for l in llks_list]
    is32 = n > 0 and all(l.dtype == np.float32 for l in ls)
    if not is32:
        ls = [np.ascontiguousarray(l, dtype=np.float64) for l in ls]
    G_ = np.array([len(l) for l in ls], dtype=np.int64)
    H_ = np.asarray(n_alleles_list, dtype=np.int64).reshape(n)
    for g, h, p in zip(G_, H_, ploidies):
        assert count_genotypes(int(h), int(p)) == g
    items = np.zeros(n, dtype=CALL_ITEM_DTYPE)
    excl = lambda x: np.concatenate([[0], np.cumsum(x)[:-1]]) if n else np.zeros(0, dtype=np.int64)
    items["ploidy"], items["n_haps"] = ploidies, H_
    items["gl_off"], items["hap_out_off"] = excl(G_), excl(H_)
    items["freqs_off"], items["inbreeding"] = -1, np.nan
    fs, fo = [], 0
    for i in range(n):
        prior = None if priors is None else priors[i]
        if prior is not None:
            items["inbreeding"][i] = float(prior[0])
            if prior[1] is not None:
                fr = np.ascontiguousarray(prior[1], dtype=np.float64)
                assert len(fr) == H_[i]
                items["freqs_off"][i] = fo
                fs.append(fr)
                fo += len(fr)
    freqs = np.concatenate(fs) if fs else None
    llks = np.concatenate(ls) if n else np.zeros(0)
    gp, of, oc, oo = dev.genotype_posteriors(items, llks, freqs, int(G_.sum()), int(H_.sum()),
                                             with_frequencies=with_frequencies)
    goff, hoff = items["gl_off"], items["hap_out_off"]
    out = [gp[int(o): int(o) + int(g)] for o, g in zip(goff, G_)]
    if not with_frequencies:
        return out
    trip = [(of[int(o): int(o) + int(h)], oc[int(o): int(o) + int(h)], oo[int(o): int(o) + int(h)])
            for o, h in zip(hoff, H_)]
    return out, trip


def genotype_posteriors(log_likelihoods, ploidy, n_alleles, prior=None, device=None):
    """Posterior probability of every genotype in VCF order; a float32 input is handled with the
    reference's mixed precision (sum rounded to float32, normalisation in float64)."""
    dev = device or default_device()
    llks = np.ascontiguousarray(log_likelihoods)
    items, freqs = _posterior_items(len(llks), ploidy, n_alleles, prior)
    gp, _, _, _ = dev.genotype_posteriors(items, llks, freqs, len(llks), int(n_alleles))
    return gp


def posterior_allele_frequencies(posteriors, ploidy, n_alleles, device=None):
    """(mean allele frequencies, posterior allele counts, occurrence) from VCF-ordered posteriors.
    Evaluated as posteriors of log(p) under no prior, which reproduces p up to normalisation."""
    dev = device or default_device()
    p = np.ascontiguousarray(posteriors, dtype=np.float64)
    total = p.sum()
    if not total > 0.0:   # all-zero posteriors: the reference's plain sums give zeros (exact.py:355-369)
        z = np.zeros(int(n_alleles), dtype=np.float64)
        return z, z.copy(), z.copy()
    items, _ = _posterior_items(len(p), ploidy, n_alleles, None)
    with np.errstate(divide="ignore"):
        lp = np.log(p)
    _, freqs, counts, occur = dev.genotype_posteriors(items, lp, None, len(p), int(n_alleles), with_frequencies=True)
    return freqs * total, counts * total, occur * total


def alternate_dosage_posteriors(genotype_alleles, probabilities):
    """Posterior of every dosage variant of the genotype's haplotype set, in VCF order
    (host-side index arithmetic over an array the device produced; exact.py:372-407)."""
    from ..jitutils import genotype_alleles_as_index

    genotype_alleles = np.asarray(genotype_alleles)
    ploidy = len(genotype_alleles)
    support = np.unique(genotype_alleles)
    extras = list(combinations_with_replacement(support, ploidy - len(support)))
    variants = np.array([np.sort(np.concatenate([support, np.array(e, dtype=support.dtype)])) for e in extras])
    ranks = np.array([genotype_alleles_as_index(v) for v in variants], dtype=np.int64)
    order = np.argsort(ranks)
    return variants[order].astype(int), np.asarray(probabilities)[ranks[order]]
