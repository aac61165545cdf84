"""Read log-likelihood on the GPU (reference: mchap/assemble/likelihood.py:18-70)."""
import numpy as np

from ..api import default_device


def log_likelihood(reads, genotype, read_counts=None, device=None):
    """Log likelihood of observed reads given a genotype — same arguments and meaning as the
    reference's ``log_likelihood(reads, genotype, read_counts=None)``; evaluated by the CUDA
    kernel ``llk_batch_kernel`` (one warp per pair)."""
    dev = device or default_device()
    out = dev.log_likelihood_batch([reads], [genotype], None if read_counts is None else [read_counts])
    return float(out[0])


def log_likelihood_batch(reads_list, genotypes_list, counts_list=None, device=None):
    """Vector form: one log-likelihood per (reads, genotype[, counts]) triple."""
    dev = device or default_device()
    return np.asarray(dev.log_likelihood_batch(reads_list, genotypes_list, counts_list))
