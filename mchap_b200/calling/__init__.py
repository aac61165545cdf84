"""Genotype calling over known haplotypes (mirror of the reference's ``mchap.calling`` surface)."""
from . import exact
from .classes import CallingMCMC, GenotypeAllelesMultiTrace, PosteriorGenotypeAllelesDistribution

__all__ = ["exact", "CallingMCMC", "GenotypeAllelesMultiTrace", "PosteriorGenotypeAllelesDistribution"]
