"""Multiset rank / unrank and binomials with the reference's names (mchap/jitutils.py:186-318).

The array forms run on the GPU (bit-exact integer kernels); the scalar helpers used for host-side
index arithmetic are exact Python integers with the reference's conventions
(``comb_with_replacement(0, 0) == 0``, jitutils.py:232-233)."""
from math import comb as _comb

import numpy as np

from .api import default_device


def comb(n, k):
    if n < 0 or k < 0:
        raise ValueError("n and k must be non-negative integers")
    return _comb(int(n), int(k))


def comb_with_replacement(n, k):
    if n < 0:
        raise ValueError("n must be a non-negative integer")
    if n == 0 and k == 0:
        return 0
    return _comb(int(n) + int(k) - 1, int(k))


def genotype_alleles_as_index(alleles):
    """VCF-order index of one sorted genotype (host integers; see genotypes_as_indices for arrays)."""
    index = 0
    for i, a in enumerate(alleles):
        if a < 0:
            raise ValueError("Allele numbers must be >= 0.")
        index += comb_with_replacement(int(a), i + 1)
    return index


def genotypes_as_indices(alleles, device=None):
    """GPU: ranks of an int array [n, ploidy] of sorted genotypes."""
    return (device or default_device()).genotype_alleles_as_index(alleles)


def indices_as_genotypes(index, ploidy, device=None):
    """GPU: genotypes int64[n, ploidy] of an index array (negative index -> -1 alleles)."""
    return (device or default_device()).index_as_genotype_alleles(index, ploidy)


def index_as_genotype_alleles(index, ploidy, device=None):
    """One genotype from its VCF-order index; None for a negative index like the reference
    (jitutils.py:300-303)."""
    if index < 0:
        return None
    return indices_as_genotypes(np.array([index], dtype=np.int64), ploidy, device)[0]


def increment_genotype(genotype):
    """Advance a sorted genotype of allele numbers, in place, to its successor in VCF order
    (jitutils.py:113-146): bump the last copy of the lowest run and reset what lies below it.
    Raises ValueError on descending alleles like the reference."""
    ploidy = len(genotype)
    if ploidy == 1:
        genotype[0] += 1
        return
    for i in range(1, ploidy):
        if genotype[i] < genotype[i - 1]:
            raise ValueError("genotype alleles are not in ascending order")
        if genotype[i] > genotype[i - 1]:
            genotype[i - 1] += 1
            genotype[0:i - 1] = 0
            return
    genotype[-1] += 1
    genotype[0:-1] = 0


_LOG10_E = np.log10(np.exp(1))


def natural_log_to_log10(x):
    """x * log10(e) in the dtype of x (jitutils.py:174-177; the VCF GL field).  numba multiplies a
    float32 array by the float64 constant element-wise in float64 and stores float64."""
    return np.asarray(x, dtype=np.float64) * _LOG10_E if np.ndim(x) else x * _LOG10_E


def greedy_choice(probabilities):
    """Index of the largest probability (jitutils.py:95-110)."""
    return int(np.argmax(probabilities))


def set_haplotype_dosage(genotype, dosage):
    """Rewrite ``genotype`` in place so that haplotype h is present ``dosage[h]`` times: surplus
    copies of a haplotype overwrite the rows whose dosage is zero, scanning rows upwards once
    (jitutils.py:425-461; ``dosage`` itself is left untouched)."""
    dosage = np.array(dosage, copy=True)
    ploidy = len(genotype)
    free = 0
    for h in range(ploidy):
        while dosage[h] > 1:
            while free < ploidy and dosage[free] != 0:
                free += 1
            if free >= ploidy:
                return
            genotype[free] = genotype[h]
            dosage[h] -= 1
            dosage[free] += 1
