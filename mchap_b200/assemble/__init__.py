"""De novo haplotype assembly (mirror of the reference's ``mchap.assemble`` call surface)."""
from .mcmc import DenovoMCMC
from .classes import GenotypeMultiTrace, PosteriorGenotypeDistribution, GenotypeSupportDistribution
from .likelihood import log_likelihood, log_likelihood_batch

__all__ = [
    "DenovoMCMC",
    "GenotypeMultiTrace",
    "PosteriorGenotypeDistribution",
    "GenotypeSupportDistribution",
    "log_likelihood",
    "log_likelihood_batch",
]
