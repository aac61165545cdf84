"""Synthetic locus x sample inputs in the reference's own encoding.

Produces, for a batch of (locus, sample) items, exactly what the reference's
applications hand to the hot path: de-duplicated probabilistic reads
``float64[U, N, A]`` plus ``read_counts int64[U]``
(reference: ``mchap/application/baseclass.py:193-209`` ->
``mchap/encoding/integer/transcode.py:16-77`` (as_probabilistic, p = 1 - 0.0024,
error_factor 3) -> ``mchap/mset.py:361-392`` (unique_counts, first-occurrence order)).

Definition of the synthetic workload:
  * ``ploidy`` true haplotypes per item, alleles uniform over ``n_alleles`` per SNV;
  * ``depth`` read fragments per item; each fragment copies one true haplotype chosen
    uniformly, covers one random contiguous window of the ``n_pos`` SNVs (length uniform in
    ceil(n_pos/2)..n_pos, start uniform; positions outside the window are gaps = NaN);
  * every covered base is flipped to another allele with probability ``error_rate``;
  * encoding: called allele 1 - error_rate, every other allele error_rate / 3, alleles
    >= n_alleles[j] -> 0; identical fragments are merged (first occurrence order) with counts.

Everything is vectorised over items so that 10^6 items are generated in seconds.
"""
from dataclasses import dataclass

import numpy as np

PFEIFFER_ERROR = 0.0024  # reference mchap/constant.py:3


@dataclass
class ItemBatch:
    """Ragged batch of locus x sample items (all with the same n_pos / max_allele)."""

    reads: np.ndarray        # float64 [sum(U), n_pos, max_allele]
    counts: np.ndarray       # int64   [sum(U)]
    offsets: np.ndarray      # int64   [n_items + 1]  (in reads)
    n_alleles: np.ndarray    # int8    [n_items, n_pos]
    haplotypes: np.ndarray   # int8    [n_items, ploidy, n_pos] the simulated truth
    ploidy: int
    n_pos: int
    max_allele: int

    @property
    def n_items(self):
        return len(self.offsets) - 1

    def item(self, i):
        s, e = int(self.offsets[i]), int(self.offsets[i + 1])
        return self.reads[s:e], self.counts[s:e]

    def n_reads(self):
        return np.diff(self.offsets)

    def slice(self, start, stop):
        s, e = int(self.offsets[start]), int(self.offsets[stop])
        return ItemBatch(
            self.reads[s:e], self.counts[s:e], self.offsets[start:stop + 1] - s,
            self.n_alleles[start:stop], self.haplotypes[start:stop],
            self.ploidy, self.n_pos, self.max_allele,
        )


def encode_calls(calls, n_alleles, max_allele, error_rate=PFEIFFER_ERROR):
    """int allele calls (-1 = gap) [..., n_pos] -> float64 [..., n_pos, max_allele]
    with the reference's ``as_probabilistic`` values (transcode.py:61-75)."""
    calls = np.asarray(calls)
    n_alleles = np.asarray(n_alleles)
    p = 1 - error_rate
    other = (1 - p) / 3
    alleles = np.arange(max_allele)
    onehot = calls[..., None] == alleles
    out = np.where(onehot, p, other).astype(np.float64)
    out[calls < 0] = np.nan
    out[np.broadcast_to(n_alleles[..., None] <= alleles, out.shape)] = 0
    return out


def fragments_for_depth(depth_at_snv, n_pos, min_window=None):
    """Fragments per item that give a mean read depth of ``depth_at_snv`` per SNV under the window
    model above (window length uniform in ceil(n_pos/2)..n_pos covers (lo + n_pos) / (2 n_pos) of the
    positions on average): 53 fragments for depth 40 at 8 SNVs, 133 for depth 100 at 16 SNVs
    ("each read covers a random contiguous window ... so depth-at-SNV is about
    the stated depth")."""
    lo = (int(n_pos) + 1) // 2 if min_window is None else int(min_window)
    return int(round(float(depth_at_snv) * 2.0 * n_pos / (lo + n_pos)))


def synth_items(n_items, ploidy=4, n_pos=8, depth=40, n_alleles=2, error_rate=PFEIFFER_ERROR,
                seed=0, min_window=None, window=True):
    """Generate ``n_items`` items; see the module docstring for the model."""
    rng = np.random.default_rng(seed)
    haps = rng.integers(0, int(n_alleles), size=(int(n_items), ploidy, int(n_pos)), dtype=np.int8)
    return _synth_from_haplotypes(haps, depth, n_alleles, error_rate, rng, min_window, window)


def synth_haplotype_panel(n_items, n_haplotypes, n_pos, ploidy, depth=40, n_alleles=2,
                          error_rate=PFEIFFER_ERROR, seed=0):
    """Items for ``mchap call`` / ``call-exact``: per item a panel of ``n_haplotypes`` DISTINCT
    known haplotypes int8[H, n_pos] and reads simulated from ``ploidy`` of them.

    Returns (ItemBatch, panels int8[n_items, H, n_pos], truth int64[n_items, ploidy])."""
    rng = np.random.default_rng(seed + 7919)
    A = int(n_alleles)
    H = int(n_haplotypes)
    assert A ** n_pos >= H
    # distinct haplotypes: sample H distinct codes in [0, A^n_pos) per item
    codes = np.empty((n_items, H), dtype=np.int64)
    space = A ** n_pos
    for i in range(n_items):
        codes[i] = rng.choice(space, size=H, replace=False)
    digits = (codes[:, :, None] // (A ** np.arange(n_pos))[None, None, :]) % A
    panels = digits.astype(np.int8)
    truth = np.sort(rng.integers(0, H, size=(n_items, ploidy)), axis=1)
    haps = np.take_along_axis(panels, truth[:, :, None], axis=1)
    batch = _synth_from_haplotypes(haps, depth, A, error_rate, rng)
    return batch, panels, truth


def _synth_from_haplotypes(haps, depth, n_alleles, error_rate, rng, min_window=None, window=True):
    """Simulate, encode and de-duplicate fragments of the given true haplotypes [n, ploidy, N]."""
    n, ploidy, N = haps.shape
    A, R = int(n_alleles), int(depth)
    pick = rng.integers(0, ploidy, size=(n, R))
    calls = np.take_along_axis(haps, pick[:, :, None].astype(np.int64), axis=1).astype(np.int8)
    flips = rng.random((n, R, N)) < error_rate
    if A > 1:
        shift = rng.integers(1, A, size=(n, R, N), dtype=np.int8)
        calls = np.where(flips, (calls + shift) % A, calls).astype(np.int8)
    if window:
        lo = (N + 1) // 2 if min_window is None else int(min_window)
        length = rng.integers(lo, N + 1, size=(n, R))
        start = np.floor(rng.random((n, R)) * (N - length + 1)).astype(np.int64)
        pos = np.arange(N)[None, None, :]
        covered = (pos >= start[:, :, None]) & (pos < (start + length)[:, :, None])
        calls = np.where(covered, calls, -1).astype(np.int8)
    # de-duplicate fragments per item, first-occurrence order (mset.unique_counts)
    base = A + 1
    assert base ** N < 2 ** 62, "window key does not fit an int64"
    weights = base ** np.arange(N, dtype=np.int64)
    key = ((calls.astype(np.int64) + 1) * weights).sum(axis=-1)          # [n, R]
    kmax = base ** N
    flat = (np.arange(n, dtype=np.int64)[:, None] * kmax + key).ravel()
    uniq, first, cnt = np.unique(flat, return_index=True, return_counts=True)
    item_of = uniq // kmax
    order = np.lexsort((first, item_of))
    first, cnt, item_of = first[order], cnt[order], item_of[order]
    u_calls = calls.reshape(n * R, N)[first]
    n_all = np.full((n, N), A, dtype=np.int8)
    reads = encode_calls(u_calls, n_all[item_of], A, error_rate)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(item_of, minlength=n), out=offsets[1:])
    batch = ItemBatch(reads, cnt.astype(np.int64), offsets, n_all, haps.astype(np.int8), ploidy, N, A)
    batch.calls = calls  # int8 [n_items, depth, n_pos]: the raw fragments before encoding / de-duplication
    return batch
