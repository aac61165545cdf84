"""Host-side batch packing (no device needed): descriptors and flat arrays handed to the C ABI."""
import numpy as np

from mchap_b200 import DenovoMCMC
from mchap_b200.api import CallBatch, count_genotypes
from mchap_b200.synth import synth_haplotype_panel, synth_items


def test_assemble_pack_layout():
    rng = np.random.default_rng(0)
    reads, nalls, inits, counts = [], [], [], []
    for i in range(30):
        N = int(rng.integers(0, 7))
        U = int(rng.integers(0, 20))
        A = int(rng.integers(1, 4)) if N else 0
        reads.append(rng.random((U, N, A)))
        nalls.append(np.full(N, max(A, 1), dtype=np.int8))
        inits.append(rng.integers(0, 2, size=(2, 4, N)).astype(np.int8) if i % 4 == 0 else None)
        counts.append(rng.integers(1, 5, size=U) if i % 3 else None)
    model = DenovoMCMC(ploidy=4, n_alleles=None, steps=20, chains=2, random_seed=3, inbreeding=0.2)
    pk = model._pack(reads, counts, inits, nalls, list(range(100, 130)))
    items = pk["items"]
    ro = co = no = go = lo = io = 0
    for i in range(30):
        U, N, A = reads[i].shape
        it = items[i]
        assert (it["reads_off"], it["counts_off"], it["nalleles_off"]) == (ro, co, no)
        assert (it["genotypes_off"], it["llks_off"]) == (go, lo)
        assert (it["n_reads"], it["n_pos"], it["max_allele"], it["ploidy"]) == (U, N, max(A, 1), 4)
        assert it["seed"] == 100 + i and it["inbreeding"] == 0.2 and it["n_temps"] == 1
        np.testing.assert_array_equal(pk["reads"][ro: ro + reads[i].size], reads[i].ravel())
        want_c = np.ones(U, dtype=np.int64) if counts[i] is None else counts[i]
        np.testing.assert_array_equal(pk["counts"][co: co + U], want_c)
        np.testing.assert_array_equal(pk["n_alleles"][no: no + N], nalls[i])
        if inits[i] is None:
            assert it["initial_off"] == -1
        else:
            assert it["initial_off"] == io and it["initial_nhet"] == N
            np.testing.assert_array_equal(pk["initial"][io: io + inits[i].size], inits[i].ravel())
            io += inits[i].size
        ro += reads[i].size
        co += U
        no += N
        go += 2 * 20 * 4 * N
        lo += 2 * 20
    assert (pk["genotypes_len"], pk["llks_len"]) == (go, lo)
    assert pk["lens"] == (ro, co, no, io)
    assert pk["nmax"] == max(max(r.shape[1] for r in reads), 1)
    empty = model._pack([], None, None, [], None)
    assert len(empty["items"]) == 0 and empty["genotypes_len"] == 0 and empty["counts"] is None


def test_assemble_pack_defaults():
    b = synth_items(12, ploidy=4, n_pos=8, depth=20, seed=2)
    reads = [b.item(i)[0] for i in range(12)]
    model = DenovoMCMC(ploidy=4, n_alleles=[2] * 8, steps=10, chains=3, random_seed=11)
    pk = model._pack(reads, None, None, None, None)
    assert pk["counts"] is None and (pk["items"]["counts_off"] == 0).all()
    assert (pk["items"]["seed"] == 11).all() and np.isnan(pk["items"]["inbreeding"]).all()
    assert pk["n_alleles"].tolist() == [2] * 96


def test_call_batch_layout():
    rng = np.random.default_rng(1)
    batch, panels, _ = synth_haplotype_panel(20, 10, 5, 4, depth=12, seed=3)
    reads = [batch.item(i)[0] for i in range(20)]
    counts = [batch.item(i)[1] if i % 2 else None for i in range(20)]
    haps = [p[: int(rng.integers(1, 10))] for p in panels]
    ploidy = rng.integers(2, 7, size=20)
    priors = [(0.1, rng.dirichlet(np.ones(len(haps[i])))) if i % 3 == 0 else ((0.3, None) if i % 3 == 1 else None)
              for i in range(20)]
    cb = CallBatch(reads, haps, ploidy, counts, priors)
    ro = co = ho = fo = oo = go = 0
    for i in range(20):
        U, N, A = reads[i].shape
        H = len(haps[i])
        it = cb.items[i]
        assert (it["reads_off"], it["counts_off"], it["haps_off"], it["hap_out_off"], it["gl_off"]) == (ro, co, ho, oo, go)
        assert (it["n_reads"], it["n_pos"], it["max_allele"], it["ploidy"], it["n_haps"]) == (U, N, A, ploidy[i], H)
        G = count_genotypes(H, int(ploidy[i]))
        assert cb.n_genotypes[i] == G
        if priors[i] is None:
            assert np.isnan(it["inbreeding"]) and it["freqs_off"] == -1
        elif priors[i][1] is None:
            assert it["inbreeding"] == priors[i][0] and it["freqs_off"] == -1
        else:
            assert it["freqs_off"] == fo
            np.testing.assert_array_equal(cb.freqs[fo: fo + H], priors[i][1])
            fo += H
        np.testing.assert_array_equal(cb.haps[ho: ho + H * N], haps[i].ravel())
        want_c = np.ones(U, dtype=np.int64) if counts[i] is None else counts[i]
        np.testing.assert_array_equal(cb.counts[co: co + U], want_c)
        ro += reads[i].size
        co += U
        ho += H * N
        oo += H
        go += G
    assert (cb.hap_total, cb.gl_total, cb.pmax) == (oo, go, int(ploidy.max()))



def test_many_distinct_seeds_are_split_into_sub_batches(monkeypatch):
    """every distinct seed costs a pre-generated stream on the device; batches with more
    distinct per-item seeds than MAX_DISTINCT_SEEDS run as consecutive sub-batches, results joined."""
    from mchap_b200.assemble import mcmc

    calls = []

    class Fake(object):
        def fit(self, reads_list, seeds, tag):
            calls.append((len(reads_list), list(seeds)))
            return [("r", s) for s in seeds], np.asarray(seeds)

    monkeypatch.setattr(mcmc, "MAX_DISTINCT_SEEDS", 4)
    seeds = np.arange(10)
    out = mcmc.split_by_seeds(Fake(), "fit", 10, seeds, dict(reads_list=list(range(10)), seeds=seeds), dict(tag=1))
    assert [c[0] for c in calls] == [4, 4, 2]
    assert [o[1] for o in out[0]] == list(range(10)) and out[1].tolist() == list(range(10))
    assert mcmc.split_by_seeds(Fake(), "fit", 10, np.zeros(10, dtype=int), {}, {}) is None
    assert mcmc.split_by_seeds(Fake(), "fit", 10, None, {}, {}) is None


def test_fragments_for_depth_matches_the_window_model():
    """53 fragments give depth 40 at every SNV of an 8-SNV locus, 133 give depth 100 at 16 SNVs
    (windows cover 50-100 % of the locus)."""
    from mchap_b200.synth import fragments_for_depth, synth_items

    assert fragments_for_depth(40, 8) == 53 and fragments_for_depth(100, 16) == 133
    b = synth_items(400, ploidy=4, n_pos=8, depth=53, seed=3)
    depth_at_snv = (b.calls >= 0).mean() * 53
    assert 38.5 < depth_at_snv < 41.5


def test_page_locked_size_classes():
    """Pooled page-locked blocks: a size class is at most 12.5 % above the request and classes are
    coarse enough for blocks to be reused by the next batch of about the same size."""
    from mchap_b200.api import Device

    for n in (1, 70000, 10 ** 6, 123456789, 2_400_000_000):
        c = Device._pin_class(n)
        assert c >= n and c >= (1 << 16) and (c <= n * 1.125 + 4096 or n < (1 << 16))
    assert Device._pin_class(2_400_000_000) == Device._pin_class(2_390_000_000)
