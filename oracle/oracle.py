"""ctypes front-end of the plain-C oracle (``mchap_oracle.c``) — TEST INFRASTRUCTURE ONLY.

The functions keep the names and argument meaning of the reference functions they
restate (file:line in each docstring, relative to the reference repository) so that
parity tests read like the reference's own tests.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmchap_oracle.so")

ERRORS = {
    1: (ValueError, "Encountered log likelihood of nan"),
    2: (ValueError, "breaks must be smaller then n"),
    3: (ValueError, "genotype alleles are not in ascending order"),
    4: (IndexError, "random_choice returned an out of range option"),
    5: (AssertionError, "initial genotype has the wrong shape"),
    6: (RuntimeError, "replay word stream exhausted"),
    7: (ValueError, "step_type must be 0 (recombination) or 1 (dosage)."),
}


def build(force=False):
    """Compile the oracle with gcc (see oracle/Makefile)."""
    src = os.path.join(_HERE, "mchap_oracle.c")
    if (
        not force
        and os.path.exists(_LIB_PATH)
        and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)
    ):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libmchap_oracle.so"])
    return _LIB_PATH


_lib = None

_f64p = C.POINTER(C.c_double)
_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_i8p = C.POINTER(C.c_int8)
_u32p = C.POINTER(C.c_uint32)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_rng_new.restype = C.c_void_p
        L.orc_rng_new.argtypes = [C.c_uint32]
        L.orc_rng_new_replay.restype = C.c_void_p
        L.orc_rng_new_replay.argtypes = [_u32p, C.c_int64]
        L.orc_rng_free.argtypes = [C.c_void_p]
        L.orc_rng_words.restype = C.c_int64
        L.orc_rng_words.argtypes = [C.c_void_p]
        L.orc_rng_next_u32.restype = C.c_uint32
        L.orc_rng_next_u32.argtypes = [C.c_void_p]
        L.orc_rng_next_double.restype = C.c_double
        L.orc_rng_next_double.argtypes = [C.c_void_p]
        L.orc_rng_next_randint.restype = C.c_int64
        L.orc_rng_next_randint.argtypes = [C.c_void_p, C.c_int64]
        L.orc_rng_shuffle_i64.argtypes = [C.c_void_p, _i64p, C.c_int64]
        L.orc_mt19937_words.argtypes = [C.c_uint32, _u32p, C.c_int64]
        L.orc_random_choice.restype = C.c_int64
        L.orc_random_choice.argtypes = [C.c_void_p, _f64p, C.c_int64]
        for name in ("orc_add_log_prob",):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_double, C.c_double]
        L.orc_sum_log_probs.restype = C.c_double
        L.orc_sum_log_probs.argtypes = [_f64p, C.c_int64]
        L.orc_normalise_log_probs.argtypes = [_f64p, C.c_int64, _f64p]
        L.orc_comb.restype = C.c_int64
        L.orc_comb.argtypes = [C.c_int64, C.c_int64]
        L.orc_comb_with_replacement.restype = C.c_int64
        L.orc_comb_with_replacement.argtypes = [C.c_int64, C.c_int64]
        L.orc_increment_genotype.restype = C.c_int
        L.orc_increment_genotype.argtypes = [_i64p, C.c_int]
        L.orc_genotype_alleles_as_index.restype = C.c_int64
        L.orc_genotype_alleles_as_index.argtypes = [_i64p, C.c_int]
        L.orc_index_as_genotype_alleles.restype = C.c_int
        L.orc_index_as_genotype_alleles.argtypes = [C.c_int64, C.c_int, _i64p]
        L.orc_ln_equivalent_permutations.restype = C.c_double
        L.orc_ln_equivalent_permutations.argtypes = [_i64p, C.c_int]
        L.orc_count_haplotype_copies.restype = C.c_int
        L.orc_count_haplotype_copies.argtypes = [_i8p, C.c_int, C.c_int, C.c_int]
        L.orc_get_haplotype_dosage.argtypes = [_i8p, _i8p, C.c_int, C.c_int]
        L.orc_structural_change.argtypes = [_i8p, C.c_int, C.c_int, _i8p, C.c_int, C.c_int]
        L.orc_log_likelihood.restype = C.c_double
        L.orc_log_likelihood.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, _i8p, C.c_int, _i64p]
        L.orc_log_likelihood_structural_change.restype = C.c_double
        L.orc_log_likelihood_structural_change.argtypes = [
            _f64p, C.c_int, C.c_int, C.c_int, _i8p, C.c_int, _i8p, C.c_int, C.c_int, _i64p,
        ]
        L.orc_assemble_log_genotype_prior.restype = C.c_double
        L.orc_assemble_log_genotype_prior.argtypes = [_i8p, C.c_int, C.c_double, C.c_double]
        L.orc_allelic_dosage.argtypes = [_i64p, C.c_int, _i64p]
        L.orc_calling_log_genotype_prior.restype = C.c_double
        L.orc_calling_log_genotype_prior.argtypes = [_i64p, C.c_int, C.c_int64, C.c_double, _f64p]
        L.orc_log_genotype_allele_flat_prior.restype = C.c_double
        L.orc_log_genotype_allele_flat_prior.argtypes = [_i64p, C.c_int, C.c_int]
        L.orc_log_genotype_allele_prior.restype = C.c_double
        L.orc_log_genotype_allele_prior.argtypes = [
            _i64p, C.c_int, C.c_int, C.c_int64, C.c_double, _f64p,
        ]
        L.orc_random_breaks.restype = C.c_int
        L.orc_random_breaks.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _i64p]
        L.orc_haplotype_segment_labels.argtypes = [
            _i8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i8p,
        ]
        for name in ("orc_recombination_step_n_options", "orc_dosage_step_n_options"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [_i8p, C.c_int]
        for name in ("orc_recombination_step_options", "orc_dosage_step_options"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [_i8p, C.c_int, _i8p]
        L.orc_chain_swap_acceptance.restype = C.c_double
        L.orc_chain_swap_acceptance.argtypes = [C.c_double] * 6
        L.orc_mutation_base_step.restype = C.c_double
        L.orc_mutation_base_step.argtypes = [
            C.c_void_p, _i8p, C.c_int, C.c_int, _f64p, C.c_int, C.c_int, _i64p, C.c_double,
            C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int),
        ]
        L.orc_mutation_compound_step.restype = C.c_double
        L.orc_mutation_compound_step.argtypes = [
            C.c_void_p, _i8p, C.c_int, C.c_int, _f64p, C.c_int, C.c_int, _i64p, C.c_double,
            _i8p, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int),
        ]
        L.orc_structural_interval_step.restype = C.c_double
        L.orc_structural_interval_step.argtypes = [
            C.c_void_p, _i8p, C.c_int, C.c_int, _f64p, C.c_int, C.c_int, _i64p, C.c_double,
            C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int),
        ]
        L.orc_structural_compound_step.restype = C.c_double
        L.orc_structural_compound_step.argtypes = [
            C.c_void_p, _i8p, C.c_int, C.c_int, _f64p, C.c_int, C.c_int, _i64p, C.c_double,
            _i64p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int),
        ]
        L.orc_denovo_assembler.restype = C.c_int
        L.orc_denovo_assembler.argtypes = [
            C.c_void_p, _i8p, C.c_int, C.c_int, _f64p, C.c_int, C.c_int, _i64p, _i8p,
            C.c_double, C.c_int, _f64p, C.c_int, C.c_double, C.c_double, C.c_double, _f64p,
            C.c_int, _i8p, _f64p, _i64p,
        ]
        L.orc_set_llk_cache_threshold.argtypes = [C.c_int64]
        L.orc_log_unique_haplotypes.restype = C.c_double
        L.orc_log_unique_haplotypes.argtypes = [_i8p, C.c_int]
        L.orc_homozygosity_probabilities.argtypes = [
            _f64p, C.c_int, C.c_int, C.c_int, _i8p, C.c_int, C.c_double, _i64p, _f64p,
        ]
        L.orc_read_mean_dist.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, _f64p]
        L.orc_denovo_fit.restype = C.c_int
        L.orc_denovo_fit.argtypes = [
            C.c_uint32, _u32p, C.c_int64, _f64p, C.c_int, C.c_int, C.c_int, _i64p, _i8p,
            C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, _f64p, _i32p, C.c_int,
            C.c_double, C.c_double, C.c_double, _f64p, C.c_int, _i8p, C.c_int, _i8p, _f64p,
            _i32p, _i64p, _i64p,
        ]
        L.orc_greedy_caller.argtypes = [
            _f64p, C.c_int, C.c_int, C.c_int, _i64p, _i8p, C.c_int, C.c_int, C.c_double,
            _f64p, _i64p,
        ]
        L.orc_calling_step_options.argtypes = [
            _f64p, C.c_int, C.c_int, C.c_int, _i64p, _i8p, C.c_int, _i64p, C.c_int, C.c_int,
            C.c_double, _f64p, C.c_int, _f64p, _f64p, _f64p,
        ]
        L.orc_calling_fit.restype = C.c_int
        L.orc_calling_fit.argtypes = [
            C.c_uint32, _u32p, C.c_int64, _f64p, C.c_int, C.c_int, C.c_int, _i64p, _i8p,
            C.c_int, C.c_int, C.c_double, _f64p, C.c_int, C.c_int, C.c_int, _i64p, _i64p,
            _f64p, _i64p, _i64p,
        ]
        L.orc_genotype_likelihoods.argtypes = [
            _f64p, C.c_int, C.c_int, C.c_int, _i64p, _i8p, C.c_int, C.c_int, C.c_int64, _f32p,
        ]
        L.orc_genotype_likelihoods_f64.argtypes = [
            _f64p, C.c_int, C.c_int, C.c_int, _i64p, _i8p, C.c_int, C.c_int, C.c_int64, _f64p,
        ]
        L.orc_genotype_posteriors_f32.argtypes = [
            _f32p, C.c_int64, C.c_int, C.c_int64, C.c_double, _f64p, _f64p,
        ]
        L.orc_genotype_posteriors_f64.argtypes = [
            _f64p, C.c_int64, C.c_int, C.c_int64, C.c_double, _f64p, _f64p,
        ]
        L.orc_posterior_allele_frequencies.argtypes = [
            _f64p, C.c_int64, C.c_int, C.c_int64, _f64p, _f64p, _f64p,
        ]
        L.orc_posterior_mode.argtypes = [
            _f64p, C.c_int, C.c_int, C.c_int, _i64p, _i8p, C.c_int, C.c_int, C.c_int64,
            C.c_double, _f64p, _i64p, _f64p, _f64p, _f64p, _f64p, _f64p,
        ]
        _lib = L
    return _lib


# --------------------------------------------------------------------------- helpers


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i8(a):
    return np.ascontiguousarray(a, dtype=np.int8)


def _p(a, t):
    if a is None:
        return None
    return a.ctypes.data_as(t)


def _raise(err):
    if err:
        exc, msg = ERRORS.get(int(err), (RuntimeError, "oracle error %d" % err))
        raise exc(msg)


def _reads3(reads):
    reads = _f64(reads)
    assert reads.ndim == 3
    return reads, reads.shape[0], reads.shape[1], reads.shape[2]


def _counts(read_counts):
    return None if read_counts is None else _i64(read_counts)


def _inb(inbreeding):
    return float("nan") if inbreeding is None else float(inbreeding)


def _prior(prior):
    """prior = None | (inbreeding, frequencies or None) -> (inbreeding or NaN, freqs array or None)"""
    if prior is None:
        return float("nan"), None
    inb, freqs = prior
    return float(inb), (None if freqs is None else _f64(freqs))


class Rng:
    """numba's thread-local MT19937 after ``np.random.seed(seed)`` (jitutils.py:180-183),
    or a replay of a pre-drawn 32-bit word stream."""

    def __init__(self, seed=None, words=None):
        L = lib()
        if words is not None:
            self._words = np.ascontiguousarray(words, dtype=np.uint32)
            self._h = L.orc_rng_new_replay(_p(self._words, _u32p), len(self._words))
        else:
            self._h = L.orc_rng_new(int(seed))

    def __del__(self):
        try:
            lib().orc_rng_free(self._h)
        except Exception:
            pass

    @property
    def words_consumed(self):
        return lib().orc_rng_words(self._h)

    def u32(self):
        return lib().orc_rng_next_u32(self._h)

    def random(self):
        return lib().orc_rng_next_double(self._h)

    def randint(self, n):
        return lib().orc_rng_next_randint(self._h, int(n))

    def shuffle(self, x):
        x = _i64(x)
        lib().orc_rng_shuffle_i64(self._h, _p(x, _i64p), len(x))
        return x


def set_llk_cache_threshold(threshold):
    """DenovoMCMC.llk_cache_threshold (assemble/mcmc.py:39): -1 disables the llk memo."""
    lib().orc_set_llk_cache_threshold(int(threshold))


def mt19937_words(seed, n):
    out = np.empty(int(n), dtype=np.uint32)
    lib().orc_mt19937_words(int(seed), _p(out, _u32p), int(n))
    return out


# --------------------------------------------------------------------------- jitutils


def add_log_prob(x, y):
    """jitutils.py:6-26"""
    return lib().orc_add_log_prob(float(x), float(y))


def sum_log_probs(a):
    """jitutils.py:29-48"""
    a = _f64(a)
    return lib().orc_sum_log_probs(_p(a, _f64p), len(a))


def normalise_log_probs(a):
    """jitutils.py:51-74"""
    a = _f64(a)
    out = np.empty_like(a)
    lib().orc_normalise_log_probs(_p(a, _f64p), len(a), _p(out, _f64p))
    return out


def random_choice(rng, p):
    """jitutils.py:77-92"""
    p = _f64(p)
    return lib().orc_random_choice(rng._h, _p(p, _f64p), len(p))


def comb(n, k):
    """jitutils.py:186-228"""
    r = lib().orc_comb(int(n), int(k))
    if r < 0:
        raise ValueError("n and k must be non-negative integers")
    return r


def comb_with_replacement(n, k):
    """jitutils.py:231-250"""
    r = lib().orc_comb_with_replacement(int(n), int(k))
    if r < 0:
        raise ValueError("n must be a non-negative integer")
    return r


def increment_genotype(genotype):
    """jitutils.py:113-146 (in place on an int64 array)"""
    assert genotype.dtype == np.int64 and genotype.flags.c_contiguous
    _raise(lib().orc_increment_genotype(_p(genotype, _i64p), len(genotype)))


def genotype_alleles_as_index(alleles):
    """jitutils.py:253-276"""
    a = _i64(alleles)
    r = lib().orc_genotype_alleles_as_index(_p(a, _i64p), len(a))
    if r < 0:
        raise ValueError("Allele numbers must be >= 0.")
    return r


def index_as_genotype_alleles(index, ploidy):
    """jitutils.py:279-318 (returns None for index < 0 like the reference)"""
    out = np.empty(int(ploidy), dtype=np.int64)
    none = lib().orc_index_as_genotype_alleles(int(index), int(ploidy), _p(out, _i64p))
    return None if none else out


def ln_equivalent_permutations(dosage):
    """jitutils.py:149-171"""
    d = _i64(dosage)
    return lib().orc_ln_equivalent_permutations(_p(d, _i64p), len(d))


def count_haplotype_copies(genotype, h):
    """jitutils.py:349-374"""
    g = _i8(genotype)
    return lib().orc_count_haplotype_copies(_p(g, _i8p), g.shape[0], g.shape[1], int(h))


def get_haplotype_dosage(dosage, genotype):
    """jitutils.py:377-422 (in place on an int8 dosage array)"""
    g = _i8(genotype)
    d = np.empty(g.shape[0], dtype=np.int8)
    lib().orc_get_haplotype_dosage(_p(d, _i8p), _p(g, _i8p), g.shape[0], g.shape[1])
    dosage[:] = d


def structural_change(genotype, haplotype_indices, interval=None):
    """jitutils.py:501-544 (in place on an int8 genotype)"""
    assert genotype.dtype == np.int8 and genotype.flags.c_contiguous
    P, N = genotype.shape
    idx = _i8(haplotype_indices)
    start, stop = (0, N) if interval is None else (int(interval[0]), int(interval[1]))
    lib().orc_structural_change(_p(genotype, _i8p), P, N, _p(idx, _i8p), start, stop)


# --------------------------------------------------------------------------- likelihood / priors


def log_likelihood(reads, genotype, read_counts=None):
    """assemble/likelihood.py:18-70"""
    reads, U, N, A = _reads3(reads)
    g = _i8(genotype)
    c = _counts(read_counts)
    return lib().orc_log_likelihood(_p(reads, _f64p), U, N, A, _p(g, _i8p), g.shape[0], _p(c, _i64p))


def log_likelihood_structural_change(reads, genotype, haplotype_indices, interval=None, read_counts=None):
    """assemble/likelihood.py:74-148"""
    reads, U, N, A = _reads3(reads)
    g = _i8(genotype)
    idx = _i8(haplotype_indices)
    c = _counts(read_counts)
    start, stop = (0, N) if interval is None else (int(interval[0]), int(interval[1]))
    return lib().orc_log_likelihood_structural_change(
        _p(reads, _f64p), U, N, A, _p(g, _i8p), g.shape[0], _p(idx, _i8p), start, stop, _p(c, _i64p)
    )


def assemble_log_genotype_prior(dosage, log_unique_haplotypes, inbreeding=0):
    """assemble/prior.py:81-112"""
    d = _i8(dosage)
    return lib().orc_assemble_log_genotype_prior(
        _p(d, _i8p), len(d), float(log_unique_haplotypes), float(inbreeding)
    )


def allelic_dosage(genotype_alleles):
    """calling/utils.py:7-35"""
    g = _i64(genotype_alleles)
    out = np.empty_like(g)
    lib().orc_allelic_dosage(_p(g, _i64p), len(g), _p(out, _i64p))
    return out


def calling_log_genotype_prior(genotype, unique_haplotypes, inbreeding=0, frequencies=None):
    """calling/prior.py:116-179"""
    g = _i64(genotype)
    f = None if frequencies is None else _f64(frequencies)
    return lib().orc_calling_log_genotype_prior(
        _p(g, _i64p), len(g), int(unique_haplotypes), float(inbreeding), _p(f, _f64p)
    )


def log_genotype_allele_flat_prior(genotype, variable_allele):
    """calling/prior.py:30-52"""
    g = _i64(genotype)
    return lib().orc_log_genotype_allele_flat_prior(_p(g, _i64p), len(g), int(variable_allele))


def log_genotype_allele_prior(genotype, variable_allele, unique_haplotypes, inbreeding=0, frequencies=None):
    """calling/prior.py:55-113"""
    g = _i64(genotype)
    f = None if frequencies is None else _f64(frequencies)
    return lib().orc_log_genotype_allele_prior(
        _p(g, _i64p), len(g), int(variable_allele), int(unique_haplotypes), float(inbreeding), _p(f, _f64p)
    )


# --------------------------------------------------------------------------- structural pieces


def random_breaks(rng, breaks, n):
    """assemble/structural.py:23-71"""
    out = np.zeros((int(breaks) + 1, 2), dtype=np.int64)
    _raise(lib().orc_random_breaks(rng._h, int(breaks), int(n), _p(out, _i64p)))
    return out


def haplotype_segment_labels(genotype, interval=None):
    """assemble/structural.py:394-430"""
    g = _i8(genotype)
    P, N = g.shape
    labels = np.zeros((P, 2), dtype=np.int8)
    has = 0 if interval is None else 1
    start, stop = (0, N) if interval is None else (int(interval[0]), int(interval[1]))
    lib().orc_haplotype_segment_labels(_p(g, _i8p), P, N, has, start, stop, _p(labels, _i8p))
    return labels


def recombination_step_n_options(labels):
    """assemble/structural.py:75-121"""
    l8 = _i8(labels)
    return lib().orc_recombination_step_n_options(_p(l8, _i8p), l8.shape[0])


def dosage_step_n_options(labels):
    """assemble/structural.py:182-236"""
    l8 = _i8(labels)
    return lib().orc_dosage_step_n_options(_p(l8, _i8p), l8.shape[0])


def recombination_step_options(labels):
    """assemble/structural.py:124-178"""
    l8 = _i8(labels)
    P = l8.shape[0]
    out = np.zeros((P * P + 1, P, 2), dtype=np.int8)
    n = lib().orc_recombination_step_options(_p(l8, _i8p), P, _p(out, _i8p))
    return out[:n]


def dosage_step_options(labels):
    """assemble/structural.py:239-307"""
    l8 = _i8(labels)
    P = l8.shape[0]
    out = np.zeros((P * P + 1, P, 2), dtype=np.int8)
    n = lib().orc_dosage_step_options(_p(l8, _i8p), P, _p(out, _i8p))
    return out[:n]


def chain_swap_acceptance(llk_i, log_prior_i, temp_i, llk_j, log_prior_j, temp_j):
    """assemble/tempering.py:11-58"""
    return lib().orc_chain_swap_acceptance(llk_i, log_prior_i, temp_i, llk_j, log_prior_j, temp_j)


# --------------------------------------------------------------------------- MCMC steps


def mutation_base_step(rng, genotype, reads, llk, h, j, n_alleles, log_unique_haplotypes,
                       inbreeding=None, temp=1.0, read_counts=None):
    """assemble/mutation.py:15-161 (genotype int8, updated in place)"""
    assert genotype.dtype == np.int8 and genotype.flags.c_contiguous
    reads, U, N, A = _reads3(reads)
    c = _counts(read_counts)
    err = C.c_int(0)
    out = lib().orc_mutation_base_step(
        rng._h, _p(genotype, _i8p), genotype.shape[0], N, _p(reads, _f64p), U, A, _p(c, _i64p),
        float(llk), int(h), int(j), int(n_alleles), float(log_unique_haplotypes), _inb(inbreeding),
        float(temp), C.byref(err),
    )
    _raise(err.value)
    return out


def mutation_compound_step(rng, genotype, reads, llk, n_alleles, log_unique_haplotypes,
                           inbreeding=None, temp=1.0, read_counts=None):
    """assemble/mutation.py:165-246"""
    assert genotype.dtype == np.int8 and genotype.flags.c_contiguous
    reads, U, N, A = _reads3(reads)
    c = _counts(read_counts)
    na = _i8(n_alleles)
    err = C.c_int(0)
    out = lib().orc_mutation_compound_step(
        rng._h, _p(genotype, _i8p), genotype.shape[0], N, _p(reads, _f64p), U, A, _p(c, _i64p),
        float(llk), _p(na, _i8p), float(log_unique_haplotypes), _inb(inbreeding), float(temp),
        C.byref(err),
    )
    _raise(err.value)
    return out


def structural_interval_step(rng, genotype, reads, llk, log_unique_haplotypes, inbreeding=None,
                             interval=None, step_type=0, temp=1.0, read_counts=None):
    """assemble/structural.py:434-587"""
    assert genotype.dtype == np.int8 and genotype.flags.c_contiguous
    reads, U, N, A = _reads3(reads)
    c = _counts(read_counts)
    start, stop = (0, N) if interval is None else (int(interval[0]), int(interval[1]))
    err = C.c_int(0)
    out = lib().orc_structural_interval_step(
        rng._h, _p(genotype, _i8p), genotype.shape[0], N, _p(reads, _f64p), U, A, _p(c, _i64p),
        float(llk), start, stop, int(step_type), float(log_unique_haplotypes), _inb(inbreeding),
        float(temp), C.byref(err),
    )
    _raise(err.value)
    return out


def structural_compound_step(rng, genotype, reads, llk, intervals, log_unique_haplotypes,
                             inbreeding=None, step_type=0, temp=1.0, read_counts=None):
    """assemble/structural.py:591-673"""
    assert genotype.dtype == np.int8 and genotype.flags.c_contiguous
    reads, U, N, A = _reads3(reads)
    c = _counts(read_counts)
    iv = _i64(intervals)
    err = C.c_int(0)
    out = lib().orc_structural_compound_step(
        rng._h, _p(genotype, _i8p), genotype.shape[0], N, _p(reads, _f64p), U, A, _p(c, _i64p),
        float(llk), _p(iv, _i64p), len(iv), int(step_type), float(log_unique_haplotypes),
        _inb(inbreeding), float(temp), C.byref(err),
    )
    _raise(err.value)
    return out


def denovo_assembler(rng, genotype, reads, n_alleles, steps, break_dist, inbreeding=None,
                     read_counts=None, recombination_step_probability=0.5,
                     partial_dosage_step_probability=0.5, dosage_step_probability=1.0,
                     temperatures=(1.0,)):
    """assemble/mcmc.py:269-426 -> (genotypes int8[steps,P,N], llks f64[steps], llk_evals)"""
    reads, U, N, A = _reads3(reads)
    g = _i8(genotype)
    P = g.shape[0]
    c = _counts(read_counts)
    na = _i8(n_alleles)
    bd = _f64(break_dist)
    temps = _f64(temperatures)
    og = np.zeros((steps, P, N), dtype=np.int8)
    ol = np.zeros(steps, dtype=np.float64)
    ev = C.c_int64(0)
    err = lib().orc_denovo_assembler(
        rng._h, _p(g, _i8p), P, N, _p(reads, _f64p), U, A, _p(c, _i64p), _p(na, _i8p),
        _inb(inbreeding), int(steps), _p(bd, _f64p), len(bd),
        float(recombination_step_probability), float(partial_dosage_step_probability),
        float(dosage_step_probability), _p(temps, _f64p), len(temps), _p(og, _i8p), _p(ol, _f64p),
        C.byref(ev),
    )
    _raise(err)
    return og, ol, ev.value


def log_unique_haplotypes(n_alleles):
    """assemble/mcmc.py:294 (float32 arithmetic, see mchap_oracle.c)"""
    na = _i8(n_alleles)
    return lib().orc_log_unique_haplotypes(_p(na, _i8p), len(na))


def homozygosity_probabilities(reads, n_alleles, ploidy, inbreeding=None, read_counts=None):
    """assemble/mcmc.py:495-541"""
    reads, U, N, A = _reads3(reads)
    na = _i8(n_alleles)
    c = _counts(read_counts)
    out = np.zeros((N, A), dtype=np.float64)
    lib().orc_homozygosity_probabilities(
        _p(reads, _f64p), U, N, A, _p(na, _i8p), int(ploidy), _inb(inbreeding), _p(c, _i64p),
        _p(out, _f64p),
    )
    return out


def read_mean_dist(reads):
    """assemble/mcmc.py:455-491"""
    reads, U, N, A = _reads3(reads)
    out = np.zeros((N, A), dtype=np.float64)
    lib().orc_read_mean_dist(_p(reads, _f64p), U, N, A, _p(out, _f64p))
    return out


def point_beta_probabilities(n_base, a=1.0, b=3.0):
    """assemble/mcmc.py:429-452 — host-side scipy call, kept verbatim in meaning:
    CDF differences of Beta(a, b) at k/n_base."""
    from scipy import stats

    points = np.arange(1, n_base + 1) / n_base
    probs = stats.beta(a, b).cdf(points)
    probs[1:] = probs[1:] - probs[:-1]
    return probs


def break_table(n_pos, alpha=1.0, beta=3.0, n_intervals=None):
    """Rows n=0..n_pos of break distributions for n_het == n (assemble/mcmc.py:211-217)."""
    stride = max(int(n_pos), int(n_intervals or 0), 1)
    table = np.zeros((n_pos + 1, stride), dtype=np.float64)
    lens = np.zeros(n_pos + 1, dtype=np.int32)
    for n in range(1, n_pos + 1):
        if n_intervals is None:
            row = point_beta_probabilities(n, alpha, beta)
        else:
            row = np.zeros(n_intervals, dtype=np.float64)
            row[-1] = 1
        table[n, : len(row)] = row
        lens[n] = len(row)
    return table, lens, stride


def denovo_fit(reads, read_counts, ploidy, n_alleles, inbreeding=None, steps=1000, chains=2,
               alpha=1.0, beta=3.0, n_intervals=None, fix_homozygous=0.999,
               recombination_step_probability=0.5, partial_dosage_step_probability=0.5,
               dosage_step_probability=1.0, temperatures=(1.0,), random_seed=42, initial=None,
               replay_words=None):
    """assemble/mcmc.py:103-265 DenovoMCMC(...).fit(reads, read_counts, initial)
    -> dict(genotypes int8[C,S,P,N], llks f64[C,S], n_het, words, llk_evals)"""
    reads, U, N, A = _reads3(reads)
    c = _counts(read_counts)
    na = _i8(n_alleles)
    assert len(na) == N
    temps = np.sort(_f64(temperatures))
    assert temps[0] >= 0.0 and temps[-1] == 1.0
    table, lens, stride = break_table(N, alpha, beta, n_intervals)
    og = np.zeros((chains, steps, ploidy, N), dtype=np.int8)
    ol = np.zeros((chains, steps), dtype=np.float64)
    nhet = C.c_int32(0)
    words = C.c_int64(0)
    ev = C.c_int64(0)
    init = None
    init_nhet = 0
    if initial is not None:
        init = _i8(initial)
        init_nhet = init.shape[-1]
    rw = None if replay_words is None else np.ascontiguousarray(replay_words, dtype=np.uint32)
    err = lib().orc_denovo_fit(
        int(random_seed), _p(rw, _u32p), 0 if rw is None else len(rw), _p(reads, _f64p), U, N, A,
        _p(c, _i64p), _p(na, _i8p), int(ploidy), _inb(inbreeding), int(steps), int(chains),
        float(fix_homozygous), _p(table, _f64p), _p(lens, _i32p), stride,
        float(recombination_step_probability), float(partial_dosage_step_probability),
        float(dosage_step_probability), _p(temps, _f64p), len(temps), _p(init, _i8p), init_nhet,
        _p(og, _i8p), _p(ol, _f64p), C.byref(nhet), C.byref(words), C.byref(ev),
    )
    _raise(err)
    return dict(genotypes=og, llks=ol, n_het=nhet.value, words=words.value, llk_evals=ev.value)


# --------------------------------------------------------------------------- calling


def greedy_caller(haplotypes, ploidy, reads, read_counts, prior=None):
    """calling/mcmc.py:393-453"""
    reads, U, N, A = _reads3(reads)
    haps = _i8(haplotypes)
    c = _counts(read_counts)
    inb, freqs = _prior(prior)
    out = np.zeros(ploidy, dtype=np.int64)
    lib().orc_greedy_caller(
        _p(reads, _f64p), U, N, A, _p(c, _i64p), _p(haps, _i8p), len(haps), int(ploidy), inb,
        _p(freqs, _f64p), _p(out, _i64p),
    )
    return out


def calling_step_options(genotype_alleles, variable_allele, haplotypes, reads, read_counts,
                         prior=None, step_type=0):
    """calling/mcmc.py:143-229 (step_type 0) / 15-140 (step_type 1) -> (llks, lpriors, probs)"""
    reads, U, N, A = _reads3(reads)
    haps = _i8(haplotypes)
    c = _counts(read_counts)
    inb, freqs = _prior(prior)
    g = _i64(genotype_alleles).copy()
    H = len(haps)
    llks = np.zeros(H)
    lpriors = np.zeros(H)
    probs = np.zeros(H)
    lib().orc_calling_step_options(
        _p(reads, _f64p), U, N, A, _p(c, _i64p), _p(haps, _i8p), H, _p(g, _i64p), len(g),
        int(variable_allele), inb, _p(freqs, _f64p), int(step_type), _p(llks, _f64p),
        _p(lpriors, _f64p), _p(probs, _f64p),
    )
    return llks, lpriors, probs


def calling_fit(reads, read_counts, ploidy, haplotypes, prior=None, steps=1000, chains=2,
                random_seed=42, step_type="Gibbs", initial=None, replay_words=None):
    """calling/classes.py:49-124 CallingMCMC(...).fit -> dict(genotypes i64[C,S,P], llks f64[C,S], ...)"""
    reads, U, N, A = _reads3(reads)
    haps = _i8(haplotypes)
    c = _counts(read_counts)
    inb, freqs = _prior(prior)
    st = {"Gibbs": 0, "Metropolis-Hastings": 1}[step_type]
    og = np.zeros((chains, steps, ploidy), dtype=np.int64)
    ol = np.zeros((chains, steps), dtype=np.float64)
    init = None if initial is None else _i64(initial)
    words = C.c_int64(0)
    ev = C.c_int64(0)
    rw = None if replay_words is None else np.ascontiguousarray(replay_words, dtype=np.uint32)
    err = lib().orc_calling_fit(
        int(random_seed), _p(rw, _u32p), 0 if rw is None else len(rw), _p(reads, _f64p), U, N, A,
        _p(c, _i64p), _p(haps, _i8p), len(haps), int(ploidy), inb, _p(freqs, _f64p), int(steps),
        int(chains), st, _p(init, _i64p), _p(og, _i64p), _p(ol, _f64p), C.byref(words),
        C.byref(ev),
    )
    _raise(err)
    return dict(genotypes=og, llks=ol, words=words.value, llk_evals=ev.value)


def count_unique_genotypes(u_haps, ploidy):
    """combinatorics.py:35-54 (exact integer form of scipy comb(repetition=True))"""
    from math import comb as _c

    return _c(int(u_haps) + int(ploidy) - 1, int(ploidy))


def genotype_likelihoods(reads, ploidy, haplotypes, read_counts=None, dtype=np.float32):
    """calling/exact.py:266-292 (float32 like the reference; dtype=float64 for the unrounded table)"""
    reads, U, N, A = _reads3(reads)
    haps = _i8(haplotypes)
    c = _counts(read_counts)
    G = count_unique_genotypes(len(haps), ploidy)
    if dtype == np.float32:
        out = np.zeros(G, dtype=np.float32)
        lib().orc_genotype_likelihoods(
            _p(reads, _f64p), U, N, A, _p(c, _i64p), _p(haps, _i8p), len(haps), int(ploidy), G,
            _p(out, _f32p),
        )
    else:
        out = np.zeros(G, dtype=np.float64)
        lib().orc_genotype_likelihoods_f64(
            _p(reads, _f64p), U, N, A, _p(c, _i64p), _p(haps, _i8p), len(haps), int(ploidy), G,
            _p(out, _f64p),
        )
    return out


def genotype_posteriors(log_likelihoods, ploidy, n_alleles, prior=None):
    """calling/exact.py:295-329 (dtype of log_likelihoods selects the f32 / f64 specialisation)"""
    inb, freqs = _prior(prior)
    llks = np.ascontiguousarray(log_likelihoods)
    out = np.zeros(len(llks), dtype=np.float64)
    if llks.dtype == np.float32:
        lib().orc_genotype_posteriors_f32(
            _p(llks, _f32p), len(llks), int(ploidy), int(n_alleles), inb, _p(freqs, _f64p),
            _p(out, _f64p),
        )
    else:
        llks = _f64(llks)
        lib().orc_genotype_posteriors_f64(
            _p(llks, _f64p), len(llks), int(ploidy), int(n_alleles), inb, _p(freqs, _f64p),
            _p(out, _f64p),
        )
    return out


def posterior_allele_frequencies(posteriors, ploidy, n_alleles):
    """calling/exact.py:332-369 -> (frequencies, counts, occurrence)"""
    p = _f64(posteriors)
    freqs = np.zeros(n_alleles)
    counts = np.zeros(n_alleles)
    occur = np.zeros(n_alleles)
    lib().orc_posterior_allele_frequencies(
        _p(p, _f64p), len(p), int(ploidy), int(n_alleles), _p(freqs, _f64p), _p(counts, _f64p),
        _p(occur, _f64p),
    )
    return freqs, counts, occur


def posterior_mode(reads, ploidy, haplotypes, read_counts=None, prior=None,
                   return_support_prob=False, return_posterior_frequencies=False,
                   return_posterior_occurrence=False):
    """calling/exact.py:156-249"""
    reads, U, N, A = _reads3(reads)
    haps = _i8(haplotypes)
    c = _counts(read_counts)
    inb, freqs = _prior(prior)
    H = len(haps)
    G = count_unique_genotypes(H, ploidy)
    mode = np.zeros(ploidy, dtype=np.int64)
    mode_llk = C.c_double(0)
    mode_prob = C.c_double(0)
    support = C.c_double(0)
    of = np.zeros(H)
    oo = np.zeros(H)
    lib().orc_posterior_mode(
        _p(reads, _f64p), U, N, A, _p(c, _i64p), _p(haps, _i8p), H, int(ploidy), G, inb,
        _p(freqs, _f64p), _p(mode, _i64p), C.byref(mode_llk), C.byref(mode_prob),
        C.byref(support), _p(of, _f64p), _p(oo, _f64p),
    )
    result = [mode, mode_llk.value, mode_prob.value]
    if return_support_prob:
        result.append(support.value)
    if return_posterior_frequencies:
        result.append(of)
    if return_posterior_occurrence:
        result.append(oo)
    return tuple(result)
