"""Read encoding + de-duplication on the device.

``encode_unique_reads_batch`` does for a batch of (locus, sample) items what
``mchap/application/baseclass.py:194-209`` does per sample on the host:
``encode_read_distributions`` (io/bam.py:251-289 -> encoding/integer/transcode.py:16-77
``as_probabilistic``) followed by ``mset.unique_counts`` (mset.py:242-284, 361-392).  The result is
the ``(reads, read_counts)`` pair the samplers take.  ``prob_of_qual`` (io/util.py:40-55) stays on the
host: it is a 94-entry table.
"""
import numpy as np

from . import _lib as L
from .api import ITEM_RESULT_DTYPE, _ptr, default_device, raise_item_status

ENCODE_ITEM_DTYPE = L._np_dtype(L.EncodeItem)

__all__ = ["encode_unique_reads_batch", "call_probabilities"]


def call_probabilities(calls, quals=None, error_rate=0.0):
    """P(call correct) per base like io/bam.py:281-286: (1 - error_rate) * (1 - 10 ** (quals / -10))."""
    probs = np.ones(np.shape(calls), dtype=float) * (1 - error_rate)
    if quals is not None:
        assert np.shape(calls) == np.shape(quals)
        probs *= 1 - (10 ** (np.asarray(quals) / -10))
    return probs


def encode_unique_reads_batch(calls_list, probs_list, n_alleles_list, error_factor=3, device=None):
    """calls int[n_reads, n_pos] (< 0 = gap), probs f64[n_reads, n_pos], n_alleles int[n_pos] per item
    -> list of (reads f64[n_unique, n_pos, max_allele], read_counts int64[n_unique])."""
    dev = device or default_device()
    n = len(calls_list)
    items = np.zeros(n, dtype=ENCODE_ITEM_DTYPE)
    cs, ps, ns = [], [], []
    co = no = ro = uo = 0
    shapes = []
    for i in range(n):
        c = np.ascontiguousarray(calls_list[i], dtype=np.int8)
        assert c.ndim == 2
        R, N = c.shape
        p = np.ascontiguousarray(np.broadcast_to(np.asarray(probs_list[i], dtype=np.float64), (R, N)))
        na = np.ascontiguousarray(np.broadcast_to(np.asarray(n_alleles_list[i]), (N,)), dtype=np.int8)
        A = int(na.max()) if N > 0 else 0
        items[i] = (co, co, no, ro, uo, R, N, A, 0)
        cs.append(c.ravel())
        ps.append(p.ravel())
        ns.append(na)
        shapes.append((R, N, A, ro, uo))
        co += R * N
        no += N
        ro += R * N * A
        uo += R
    calls = np.concatenate(cs) if cs else np.zeros(0, dtype=np.int8)
    probs = np.concatenate(ps) if ps else np.zeros(0)
    nall = np.concatenate(ns) if ns else np.zeros(0, dtype=np.int8)
    out_reads = np.zeros(max(ro, 1), dtype=np.float64)
    out_counts = np.zeros(max(uo, 1), dtype=np.int64)
    results = np.zeros(n, dtype=ITEM_RESULT_DTYPE)
    rc = dev._lib.mchb_encode_reads_batch(
        dev._h, L.MEM_HOST, _ptr(items), n, _ptr(calls), calls.size, _ptr(probs), probs.size, _ptr(nall), nall.size,
        float(error_factor), _ptr(out_reads), ro, _ptr(out_counts), uo, _ptr(results))
    dev._check(rc)
    out = []
    for i, (R, N, A, r0, u0) in enumerate(shapes):
        raise_item_status(int(results["status"][i]), i if n > 1 else None)
        u = int(results["n_het"][i])
        out.append((out_reads[r0: r0 + u * N * A].reshape(u, N, A).copy(), out_counts[u0: u0 + u].copy()))
    return out
