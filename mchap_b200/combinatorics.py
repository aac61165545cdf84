"""Counting helpers with the reference's names (mchap/combinatorics.py:16-127).

Host-side integer arithmetic: these size the genotype space (G = C(H + P - 1, P)) for the device
calls and the VCF GP/GL arrays.  Everything is exact Python integers; the reference goes through
``scipy.special.comb`` (a float path, combinatorics.py:54) whose result is exact for every size the
CLIs can reach (G < 2**53), so the values agree.
"""
from math import comb as _comb
from math import factorial as _factorial

import numpy as np

__all__ = [
    "count_unique_haplotypes",
    "count_unique_genotypes",
    "count_unique_genotype_permutations",
    "count_haplotype_universial_occurance",
    "count_genotype_permutations",
]


def count_unique_haplotypes(u_alleles):
    """Possible haplotypes of a locus = product of the allele counts of its positions
    (combinatorics.py:16-32; numpy product like the reference, so an empty locus gives 1.0)."""
    return np.prod(u_alleles)


def count_unique_genotypes(u_haps, ploidy):
    """Multisets of size ``ploidy`` over ``u_haps`` haplotypes (combinatorics.py:35-54)."""
    u_haps, ploidy = int(u_haps), int(ploidy)
    if u_haps <= 0:
        return 1 if (u_haps == 0 and ploidy == 0) else 0
    return _comb(u_haps + ploidy - 1, ploidy)


def count_unique_genotype_permutations(u_haps, ploidy):
    """Ordered genotypes, equivalent permutations included (combinatorics.py:57-77)."""
    return u_haps ** ploidy


def count_haplotype_universial_occurance(u_haps, ploidy):
    """Occurrences of one haplotype among all unique genotypes (combinatorics.py:80-100)."""
    return _factorial(u_haps + ploidy - 1) // (_factorial(ploidy - 1) * _factorial(u_haps))


def count_genotype_permutations(dosage):
    """Equivalent permutations of a genotype with the given dosage: P! / prod(d_i!)
    (combinatorics.py:103-127)."""
    ploidy = sum(dosage)
    denominator = 1
    for d in dosage:
        denominator *= _factorial(int(d))
    return _factorial(int(ploidy)) // denominator
