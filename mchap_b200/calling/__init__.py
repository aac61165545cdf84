"""Genotype calling over known haplotypes (mirror of the reference's ``mchap.calling`` surface)."""
from . import exact

__all__ = ["exact"]
