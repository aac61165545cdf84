"""ctypes binding of libmchap_b200.so (the C ABI declared in include/mchap_b200.h).

There is no CPU fallback: if the shared library is missing or no sm_100 device is present the
calls raise.
"""
import ctypes as C
import os
import threading

from . import build as _build

MCHB_OK = 0
MCHB_ERR_CUDA = 100
MCHB_ERR_ARGUMENT = 101
MCHB_ERR_NO_DEVICE = 102

ITEM_OK = 0
ITEM_NAN_LLK = 1
ITEM_BREAKS = 2
ITEM_CHOICE_RANGE = 4
ITEM_INITIAL_SHAPE = 5
ITEM_RNG_EXHAUSTED = 6
ITEM_UNSUPPORTED = 8
ITEM_TALLY_OVERFLOW = 9

MEM_HOST = 0
MEM_DEVICE = 1
MEM_LAST_TRACE = 2


class Limits(C.Structure):
    _fields_ = [
        ("max_ploidy", C.c_int32),
        ("max_key_bits", C.c_int32),
        ("max_unique_reads", C.c_int32),
        ("max_temperatures", C.c_int32),
        ("max_haplotypes", C.c_int32),
    ]


class ItemResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("n_het", C.c_int32),
        ("rng_words", C.c_int64),
        ("llk_evals", C.c_int64),
    ]


class LlkItem(C.Structure):
    _fields_ = [
        ("reads_off", C.c_int64),
        ("counts_off", C.c_int64),
        ("geno_off", C.c_int64),
        ("n_reads", C.c_int32),
        ("n_pos", C.c_int32),
        ("max_allele", C.c_int32),
        ("ploidy", C.c_int32),
    ]


class AssembleItem(C.Structure):
    _fields_ = [
        ("reads_off", C.c_int64),
        ("counts_off", C.c_int64),
        ("nalleles_off", C.c_int64),
        ("initial_off", C.c_int64),
        ("genotypes_off", C.c_int64),
        ("llks_off", C.c_int64),
        ("n_reads", C.c_int32),
        ("n_pos", C.c_int32),
        ("max_allele", C.c_int32),
        ("ploidy", C.c_int32),
        ("temps_off", C.c_int32),
        ("n_temps", C.c_int32),
        ("initial_nhet", C.c_int32),
        ("seed", C.c_uint32),
        ("inbreeding", C.c_double),
    ]


class CallItem(C.Structure):
    _fields_ = [
        ("reads_off", C.c_int64),
        ("counts_off", C.c_int64),
        ("haps_off", C.c_int64),
        ("freqs_off", C.c_int64),
        ("hap_out_off", C.c_int64),
        ("gl_off", C.c_int64),
        ("n_reads", C.c_int32),
        ("n_pos", C.c_int32),
        ("max_allele", C.c_int32),
        ("ploidy", C.c_int32),
        ("n_haps", C.c_int32),
        ("reserved", C.c_int32),
        ("inbreeding", C.c_double),
    ]


class TallyItem(C.Structure):
    _fields_ = [
        ("genotypes_off", C.c_int64),
        ("states_off", C.c_int64),
        ("tallies_off", C.c_int64),
        ("n_pos", C.c_int32),
        ("ploidy", C.c_int32),
        ("chains", C.c_int32),
        ("steps", C.c_int32),
        ("burn", C.c_int32),
        ("max_unique", C.c_int32),
    ]


class EncodeItem(C.Structure):
    _fields_ = [
        ("calls_off", C.c_int64),
        ("probs_off", C.c_int64),
        ("nalleles_off", C.c_int64),
        ("reads_off", C.c_int64),
        ("counts_off", C.c_int64),
        ("n_reads", C.c_int32),
        ("n_pos", C.c_int32),
        ("max_allele", C.c_int32),
        ("reserved", C.c_int32),
    ]


class MecItem(C.Structure):
    _fields_ = [
        ("calls_off", C.c_int64),
        ("geno_off", C.c_int64),
        ("per_read_off", C.c_int64),
        ("n_reads", C.c_int32),
        ("n_pos", C.c_int32),
        ("ploidy", C.c_int32),
        ("reserved", C.c_int32),
    ]


class CallMcmcParams(C.Structure):
    _fields_ = [
        ("steps", C.c_int32),
        ("chains", C.c_int32),
        ("step_type", C.c_int32),
        ("reserved", C.c_int32),
        ("replay_words", C.c_void_p),
        ("replay_len", C.c_int64),
        ("rng_words_hint", C.c_int64),
    ]


class AssembleParams(C.Structure):
    _fields_ = [
        ("steps", C.c_int32),
        ("chains", C.c_int32),
        ("fix_homozygous", C.c_double),
        ("p_recombination", C.c_double),
        ("p_partial_dosage", C.c_double),
        ("p_dosage", C.c_double),
        ("break_table", C.c_void_p),
        ("break_len", C.c_void_p),
        ("break_rows", C.c_int32),
        ("break_stride", C.c_int32),
        ("temperatures", C.c_void_p),
        ("temperatures_len", C.c_int32),
        ("sort_haplotypes", C.c_int32),
        ("replay_words", C.c_void_p),
        ("replay_len", C.c_int64),
        ("rng_words_hint", C.c_int64),
    ]


# numpy structured dtypes with the same memory layout (vectorised descriptor construction)
def _np_dtype(struct):
    import numpy as np

    return np.dtype(
        {
            "names": [f[0] for f in struct._fields_],
            "formats": [np.dtype(f[1]) for f in struct._fields_],
            "offsets": [getattr(struct, f[0]).offset for f in struct._fields_],
            "itemsize": C.sizeof(struct),
        }
    )


_lib = None
_lock = threading.Lock()

# exported symbols declared in include/mchap_b200.h
SYMBOLS = [
    "mchb_create", "mchb_destroy", "mchb_last_error", "mchb_get_limits", "mchb_last_kernel_ms",
    "mchb_last_kernel_launches", "mchb_last_host_chunks", "mchb_stream", "mchb_sm_count", "mchb_mt19937_words",
    "mchb_genotype_rank", "mchb_genotype_unrank", "mchb_log_likelihood_batch",
    "mchb_assemble_batch", "mchb_measure_fp64_peak", "mchb_call_exact_mode_batch",
    "mchb_genotype_likelihoods_batch", "mchb_genotype_posteriors_batch", "mchb_call_mcmc_batch",
    "mchb_trace_tally_batch", "mchb_assemble_tally_batch", "mchb_call_trace_tally_batch",
    "mchb_call_mcmc_tally_batch", "mchb_encode_reads_batch", "mchb_encode_assemble_tally_batch",
    "mchb_mec_batch", "mchb_host_alloc", "mchb_host_free", "mchb_debug_counters", "mchb_last_resident_warps",
]


def library_path():
    return _build.LIB_PATH


def load():
    """dlopen the CUDA library (building it in-tree if the sources are newer and nvcc exists)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        if os.environ.get("MCHB_LIB"):
            path = os.environ["MCHB_LIB"]   # an explicitly chosen build (profiling counters)
        elif _build.needs_build() and (_build.have_nvcc() or not os.path.exists(path)):
            _build.build()   # raises without nvcc: there is nothing else to run
        L = C.CDLL(path)
        vp = C.c_void_p
        L.mchb_create.restype = C.c_int
        L.mchb_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.mchb_destroy.restype = None
        L.mchb_destroy.argtypes = [vp]
        L.mchb_last_error.restype = C.c_char_p
        L.mchb_last_error.argtypes = [vp]
        L.mchb_get_limits.restype = None
        L.mchb_get_limits.argtypes = [C.POINTER(Limits)]
        L.mchb_last_kernel_ms.restype = C.c_float
        L.mchb_last_kernel_ms.argtypes = [vp]
        L.mchb_last_kernel_launches.restype = C.c_int32
        L.mchb_last_kernel_launches.argtypes = [vp]
        L.mchb_last_host_chunks.restype = C.c_int32
        L.mchb_last_host_chunks.argtypes = [vp]
        L.mchb_stream.restype = vp
        L.mchb_stream.argtypes = [vp]
        L.mchb_sm_count.restype = C.c_int
        L.mchb_sm_count.argtypes = [vp]
        L.mchb_last_resident_warps.restype = C.c_int32
        L.mchb_last_resident_warps.argtypes = [vp]
        L.mchb_measure_fp64_peak.restype = C.c_int
        L.mchb_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
        L.mchb_mt19937_words.restype = C.c_int
        L.mchb_mt19937_words.argtypes = [vp, C.c_int, C.c_uint32, vp, C.c_int64]
        L.mchb_genotype_rank.restype = C.c_int
        L.mchb_genotype_rank.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int32, vp]
        L.mchb_genotype_unrank.restype = C.c_int
        L.mchb_genotype_unrank.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int32, vp]
        L.mchb_log_likelihood_batch.restype = C.c_int
        L.mchb_log_likelihood_batch.argtypes = [
            vp, C.c_int, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp,
        ]
        L.mchb_assemble_batch.restype = C.c_int
        L.mchb_assemble_batch.argtypes = [
            vp, C.c_int, C.POINTER(AssembleParams), vp, C.c_int64, vp, C.c_int64, vp, C.c_int64,
            vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp,
        ]
        L.mchb_call_exact_mode_batch.restype = C.c_int
        L.mchb_call_exact_mode_batch.argtypes = [
            vp, C.c_int, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64,
            vp, C.c_int32, vp, vp, vp, C.c_int64, vp,
        ]
        L.mchb_genotype_likelihoods_batch.restype = C.c_int
        L.mchb_genotype_likelihoods_batch.argtypes = [
            vp, C.c_int, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp,
        ]
        L.mchb_genotype_posteriors_batch.restype = C.c_int
        L.mchb_genotype_posteriors_batch.argtypes = [
            vp, C.c_int, vp, C.c_int64, vp, C.c_int64, vp, C.c_int, C.c_int64, vp, vp, vp, vp, C.c_int64,
        ]
        L.mchb_call_mcmc_batch.restype = C.c_int
        L.mchb_call_mcmc_batch.argtypes = [
            vp, C.c_int, C.POINTER(CallMcmcParams), vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64,
            vp, C.c_int64, vp, C.c_int32, vp, C.c_int64, vp, C.c_int64, vp,
        ]
        L.mchb_trace_tally_batch.restype = C.c_int
        L.mchb_trace_tally_batch.argtypes = [
            vp, C.c_int, C.c_int, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64, vp,
        ]
        L.mchb_assemble_tally_batch.restype = C.c_int
        L.mchb_assemble_tally_batch.argtypes = [
            vp, C.POINTER(AssembleParams), vp, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64,
            vp, C.c_int64, C.c_int64, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64, vp, vp,
        ]
        L.mchb_call_trace_tally_batch.restype = C.c_int
        L.mchb_call_trace_tally_batch.argtypes = L.mchb_trace_tally_batch.argtypes
        L.mchb_call_mcmc_tally_batch.restype = C.c_int
        L.mchb_call_mcmc_tally_batch.argtypes = [
            vp, C.POINTER(CallMcmcParams), vp, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64,
            vp, C.c_int64, vp, C.c_int32, C.c_int64, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64, vp, vp,
        ]
        L.mchb_encode_reads_batch.restype = C.c_int
        L.mchb_encode_reads_batch.argtypes = [
            vp, C.c_int, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, C.c_double, vp, C.c_int64,
            vp, C.c_int64, vp,
        ]
        L.mchb_encode_assemble_tally_batch.restype = C.c_int
        L.mchb_encode_assemble_tally_batch.argtypes = [
            vp, C.POINTER(AssembleParams), vp, vp, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64,
            C.c_double, C.c_int64, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64, vp, vp, vp,
        ]
        L.mchb_mec_batch.restype = C.c_int
        L.mchb_mec_batch.argtypes = [vp, C.c_int, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, vp, vp, C.c_int64]
        L.mchb_host_alloc.restype = C.c_int
        L.mchb_host_alloc.argtypes = [vp, C.c_int64, C.POINTER(vp)]
        L.mchb_host_free.restype = C.c_int
        L.mchb_host_free.argtypes = [vp, vp]
        L.mchb_debug_counters.restype = C.c_int
        L.mchb_debug_counters.argtypes = [vp, vp, C.c_int32, C.c_int32]
        _lib = L
        return _lib
