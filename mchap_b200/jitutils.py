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
