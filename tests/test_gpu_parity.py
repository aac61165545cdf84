"""GPU parity tests: the CUDA path (through the C ABI) against the C oracle on the same seeded
inputs, and against the committed reference fixtures.

Bars (BASELINE.json north_star): integers / indices / trajectories bit-exact; fp64
log-likelihoods within 1e-9 relative.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LLK_RTOL = 1e-9


@pytest.fixture(scope="module")
def dev():
    import mchap_b200

    return mchap_b200.default_device(0)


def close(a, b, rtol=LLK_RTOL):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=0, equal_nan=True)


def test_native_library_is_loaded(dev):
    import mchap_b200._lib as L

    with open("/proc/self/maps") as f:
        assert "libmchap_b200.so" in f.read()
    assert dev.sm_count > 0
    assert L.library_path().endswith("libmchap_b200.so")


def test_mt19937_stream(dev, oracle):
    for seed, n in ((0, 1), (11, 623), (42, 624), (42, 5000), (123456789, 70001)):
        np.testing.assert_array_equal(dev.mt19937_words(seed, n), oracle.mt19937_words(seed, n))


@pytest.mark.parametrize("P", [1, 2, 3, 4, 6, 8])
def test_rank_unrank_bit_exact(dev, golden, P):
    un = golden["unrank_p%d" % P]
    idx = np.arange(len(un), dtype=np.int64)
    np.testing.assert_array_equal(dev.index_as_genotype_alleles(idx, P), un)
    np.testing.assert_array_equal(dev.genotype_alleles_as_index(un), idx)


def test_rank_unrank_large_and_negative(dev, golden):
    idx = golden["unrank_large_idx"]
    got = dev.index_as_genotype_alleles(idx, 6)
    np.testing.assert_array_equal(got, golden["unrank_large_p6"])
    np.testing.assert_array_equal(dev.genotype_alleles_as_index(got), idx)
    assert (dev.index_as_genotype_alleles(np.array([-1]), 4) == -1).all()
    # a round trip at full enumeration size: all 52360 tetraploid genotypes of 32 haplotypes
    allidx = np.arange(52360, dtype=np.int64)
    g = dev.index_as_genotype_alleles(allidx, 4)
    assert (np.diff(g, axis=1) >= 0).all() and g.max() == 31
    np.testing.assert_array_equal(dev.genotype_alleles_as_index(g), allidx)


def test_log_likelihood_golden(dev, golden):
    ks = golden.meta["llk_cases"]
    reads = [golden["llk%d_reads" % k] for k in ks]
    genos = [golden["llk%d_genotype" % k] for k in ks]
    counts = [golden.get("llk%d_counts" % k) for k in ks]
    got = dev.log_likelihood_batch(reads, genos, counts)
    want = np.array([golden["llk%d_out" % k][0] for k in ks])
    close(got, want)
    # structural-change llk == llk of the changed genotype
    changed = [golden["llk%d_changed" % k] for k in ks]
    got2 = dev.log_likelihood_batch(reads, changed, counts)
    close(got2, np.array([golden["llk%d_out" % k][1] for k in ks]))


def test_log_likelihood_vs_oracle_fuzz(dev, oracle):
    rng = np.random.default_rng(3)
    reads, genos, counts, want = [], [], [], []
    for _ in range(200):
        P, N, A, U = int(rng.integers(1, 9)), int(rng.integers(1, 17)), int(rng.integers(2, 5)), int(rng.integers(1, 90))
        r = rng.random((U, N, A))
        r[rng.random((U, N)) < 0.3] = np.nan
        g = rng.integers(0, A, size=(P, N)).astype(np.int8)
        c = rng.integers(1, 9, size=U)
        reads.append(r)
        genos.append(g)
        counts.append(c)
        want.append(oracle.log_likelihood(r, g, c))
    close(dev.log_likelihood_batch(reads, genos, counts), np.array(want), rtol=1e-12)


def _fit_model(meta, na):
    from mchap_b200 import DenovoMCMC

    return DenovoMCMC(
        ploidy=meta["P"], n_alleles=list(na), inbreeding=meta["inbreeding"], steps=meta["steps"],
        chains=meta["chains"], n_intervals=meta["n_intervals"], fix_homozygous=meta["fix_homozygous"],
        recombination_step_probability=meta["probs"][0], partial_dosage_step_probability=meta["probs"][1],
        dosage_step_probability=meta["probs"][2], temperatures=tuple(meta["temperatures"]),
        random_seed=meta["seed"])


def test_denovo_fit_golden_trajectories(dev, golden, oracle):
    """Reference fixtures: step-for-step identical genotype traces, llks to 1e-9, and the same
    number of MT19937 words consumed as numba."""
    for k in range(golden.meta["fit_cases"]):
        meta = golden.meta["fit%d" % k]
        reads = golden["fit%d_reads" % k]
        counts = golden.get("fit%d_counts" % k)
        na = golden["fit%d_nalleles" % k]
        model = _fit_model(meta, na)
        (res,), results = model.fit_batch([reads], [counts], return_results=True, raw=True)
        g, l = res
        if meta.get("edge"):
            from mchap_b200.assemble.classes import sort_haplotypes

            np.testing.assert_array_equal(sort_haplotypes(g), golden["fit%d_sorted" % k], err_msg="case %d" % k)
            close(l, golden["fit%d_llks" % k])
            continue
        np.testing.assert_array_equal(g, golden["fit%d_genotypes" % k], err_msg="case %d" % k)
        close(l, golden["fit%d_llks" % k])
        assert results["n_het"][0] == golden["fit%d_nhet" % k][0]
        # RNG position: the next double numba would draw after the fit
        r = oracle.Rng(meta["seed"])
        for _ in range(int(results["rng_words"][0])):
            r.u32()
        assert r.random() == golden["fit%d_next" % k][0]
        # the public fit() returns the sorted trace of the reference
        trace = model.fit(reads, read_counts=counts)
        np.testing.assert_array_equal(trace.genotypes, golden["fit%d_sorted" % k])


def _oracle_fit(oracle, model, reads, counts, na, seed=None, replay=None):
    return oracle.denovo_fit(
        reads, counts, model.ploidy, na, inbreeding=model.inbreeding, steps=model.steps, chains=model.chains,
        alpha=model.alpha, beta=model.beta, n_intervals=model.n_intervals, fix_homozygous=model.fix_homozygous,
        recombination_step_probability=model.recombination_step_probability,
        partial_dosage_step_probability=model.partial_dosage_step_probability,
        dosage_step_probability=model.dosage_step_probability, temperatures=model.temperatures,
        random_seed=model.random_seed if seed is None else seed, replay_words=replay)


@pytest.mark.parametrize("ploidy,n_pos,depth,temps,inbreeding", [
    (4, 8, 40, (1.0,), None),          # BASELINE configs[1] shape
    (4, 8, 40, (1.0,), 0.1),
    (2, 6, 20, (1.0,), None),
    (6, 8, 40, (0.2, 1.0), None),
    (8, 16, 100, (0.01, 0.1, 0.5, 1.0), None),   # BASELINE configs[3] shape (one resident state slot, swapped)
    (8, 14, 90, (0.05, 0.4, 1.0), 0.15),         # the same path with the Dirichlet-multinomial prior
    (4, 12, 60, (0.5, 1.0), 0.3),
    (4, 16, 135, (1.0,), None),                  # 97..128 unique reads: the four-chunk kernel
])
def test_assemble_batch_vs_oracle(dev, oracle, ploidy, n_pos, depth, temps, inbreeding):
    from mchap_b200 import DenovoMCMC
    from mchap_b200.synth import synth_items

    n_items = 24
    steps = 120 if n_pos <= 8 else 40
    batch = synth_items(n_items, ploidy=ploidy, n_pos=n_pos, depth=depth, seed=ploidy * 100 + n_pos)
    model = DenovoMCMC(ploidy=ploidy, n_alleles=[2] * n_pos, inbreeding=inbreeding, steps=steps, chains=2,
                       temperatures=temps, random_seed=11)
    reads = [batch.item(i)[0] for i in range(n_items)]
    counts = [batch.item(i)[1] for i in range(n_items)]
    out, results = model.fit_batch(reads, counts, return_results=True, raw=True)
    total_evals = 0
    for i in range(n_items):
        ref = _oracle_fit(oracle, model, reads[i], counts[i], [2] * n_pos)
        g, l = out[i]
        np.testing.assert_array_equal(g, ref["genotypes"], err_msg="item %d" % i)
        close(l, ref["llks"])
        assert results["rng_words"][i] == ref["words"]
        assert results["n_het"][i] == ref["n_het"]
        assert results["llk_evals"][i] == ref["llk_evals"]
        total_evals += ref["llk_evals"]
    assert total_evals > 0


def test_assemble_per_item_seeds_and_ragged_alleles(dev, oracle):
    from mchap_b200 import DenovoMCMC

    rng = np.random.default_rng(5)
    reads, counts, nalls, seeds = [], [], [], []
    for i in range(12):
        N = int(rng.integers(2, 9))
        A = 4
        na = rng.integers(2, A + 1, size=N).astype(np.int8)
        U = int(rng.integers(3, 45))
        r = rng.random((U, N, A)) + 0.05
        for j in range(N):
            r[:, j, na[j]:] = 0
        r /= r.sum(axis=-1, keepdims=True)
        r[rng.random((U, N)) < 0.25] = np.nan
        reads.append(r)
        counts.append(rng.integers(1, 5, size=U))
        nalls.append(na)
        seeds.append(int(rng.integers(0, 2 ** 31)))
    model = DenovoMCMC(ploidy=4, n_alleles=None, inbreeding=0.05, steps=60, chains=2, random_seed=1)
    out, results = model.fit_batch(reads, counts, n_alleles_list=nalls, seeds=seeds, return_results=True, raw=True)
    for i in range(12):
        ref = _oracle_fit(oracle, model, reads[i], counts[i], nalls[i], seed=seeds[i])
        np.testing.assert_array_equal(out[i][0], ref["genotypes"], err_msg="item %d" % i)
        close(out[i][1], ref["llks"])
        assert results["rng_words"][i] == ref["words"]


def test_assemble_chunked_host_pipeline(dev, oracle, monkeypatch):
    """Host buffers, batch cut into chunks whose trace copies overlap the next chunks' kernels:
    same bytes as the single-launch path and as the oracle (classes CH=1 and CH=2 mixed)."""
    from mchap_b200 import DenovoMCMC
    from mchap_b200.synth import synth_items

    shallow = synth_items(40, ploidy=4, n_pos=8, depth=30, seed=3)
    deep = synth_items(13, ploidy=4, n_pos=8, depth=90, seed=4)
    reads = [shallow.item(i)[0] for i in range(40)] + [deep.item(i)[0] for i in range(13)]
    counts = [shallow.item(i)[1] for i in range(40)] + [deep.item(i)[1] for i in range(13)]
    order = np.random.default_rng(0).permutation(len(reads))
    reads = [reads[i] for i in order]
    counts = [counts[i] for i in order]
    model = DenovoMCMC(ploidy=4, n_alleles=[2] * 8, steps=80, chains=2, random_seed=5)
    monkeypatch.delenv("MCHB_HOST_CHUNKS", raising=False)
    whole, res_w = model.fit_batch(reads, counts, return_results=True, raw=True)
    assert dev.last_host_chunks == 1
    for n_chunks in (2, 5, 16):
        monkeypatch.setenv("MCHB_HOST_CHUNKS", str(n_chunks))
        piped, res_p = model.fit_batch(reads, counts, return_results=True, raw=True)
        assert dev.last_host_chunks > 1
        for i in range(len(reads)):
            np.testing.assert_array_equal(piped[i][0], whole[i][0], err_msg="item %d" % i)
            np.testing.assert_array_equal(piped[i][1], whole[i][1])
        np.testing.assert_array_equal(res_p["rng_words"], res_w["rng_words"])
    monkeypatch.delenv("MCHB_HOST_CHUNKS")
    for i in (0, 17, 52):
        ref = _oracle_fit(oracle, model, reads[i], counts[i], [2] * 8)
        np.testing.assert_array_equal(whole[i][0], ref["genotypes"])
        close(whole[i][1], ref["llks"])


def test_replay_harness(dev, oracle):
    """Both implementations driven by the same pre-drawn word stream (not an MT19937 stream)."""
    from mchap_b200 import DenovoMCMC
    from mchap_b200.synth import synth_items

    batch = synth_items(4, seed=99)
    words = np.random.default_rng(7).integers(0, 2 ** 32, size=200000, dtype=np.uint64).astype(np.uint32)
    model = DenovoMCMC(ploidy=4, n_alleles=[2] * 8, steps=100, chains=2, random_seed=0)
    reads = [batch.item(i)[0] for i in range(4)]
    counts = [batch.item(i)[1] for i in range(4)]
    out = model.fit_batch(reads, counts, raw=True, replay_words=words)
    for i in range(4):
        ref = _oracle_fit(oracle, model, reads[i], counts[i], [2] * 8, replay=words)
        np.testing.assert_array_equal(out[i][0], ref["genotypes"])
        close(out[i][1], ref["llks"])


def test_edge_cases(dev):
    from mchap_b200 import DenovoMCMC

    # zero reads (tests/test_assemble/test_mcmc.py:95-111): a random walk, must run and be seeded
    model = DenovoMCMC(ploidy=4, n_alleles=[2, 2, 2], steps=50, chains=2, random_seed=3)
    t1 = model.fit(np.empty((0, 3, 2)))
    t2 = model.fit(np.empty((0, 3, 2)))
    np.testing.assert_array_equal(t1.genotypes, t2.genotypes)
    assert t1.genotypes.shape == (2, 50, 4, 3)
    # zero SNPs (114-149)
    t = DenovoMCMC(ploidy=4, n_alleles=[], steps=20, chains=2, random_seed=3).fit(np.empty((5, 0, 0)))
    assert t.genotypes.shape == (2, 20, 4, 0) and np.isnan(t.llks).all()
    # unsupported shape fails loudly instead of falling back
    with pytest.raises(NotImplementedError):
        DenovoMCMC(ploidy=2, n_alleles=[2] * 70, steps=5, chains=1, random_seed=1, fix_homozygous=2.0).fit(
            np.full((3, 70, 2), 0.5))


# ----------------------------------------------------------------------------- call-exact (K4)
def _prior(golden, k, meta, name):
    if meta["inbreeding"] is None:
        return None
    return (meta["inbreeding"], golden["%s%d_freqs" % (name, k)] if meta["with_freqs"] else None)


def test_call_exact_golden(dev, golden):
    """posterior_mode / genotype_likelihoods / genotype_posteriors against reference fixtures:
    mode alleles and VCF-order indices bit-exact, float64 statistics to 1e-9, float32 GL exact or
    1 ulp (CUDA vs glibc log)."""
    from mchap_b200.calling import exact

    for k in range(golden.meta["exact_cases"]):
        meta = golden.meta["exact%d" % k]
        reads = golden["exact%d_reads" % k]
        counts = golden["exact%d_counts" % k]
        haps = golden["exact%d_haplotypes" % k]
        prior = _prior(golden, k, meta, "exact")
        P, H = meta["P"], meta["H"]
        mode, llk, prob, support, freqs, occur = exact.posterior_mode(
            reads, P, haps, read_counts=counts, prior=prior, return_support_prob=True,
            return_posterior_frequencies=True, return_posterior_occurrence=True)
        np.testing.assert_array_equal(mode, golden["exact%d_mode" % k], err_msg="case %d" % k)
        close([llk, prob, support], golden["exact%d_scalars" % k])
        close(freqs, golden["exact%d_mode_freqs" % k], rtol=1e-9)
        close(occur, golden["exact%d_mode_occur" % k], rtol=1e-9)
        gl = exact.genotype_likelihoods(reads, P, haps, read_counts=counts)
        assert gl.dtype == np.float32 and len(gl) == golden["exact%d_ngen" % k][0]
        np.testing.assert_allclose(gl, golden["exact%d_gl" % k], rtol=2e-7, atol=0)
        # posteriors from the REFERENCE's float32 GL array: isolates genotype_posteriors
        gp = exact.genotype_posteriors(golden["exact%d_gl" % k], P, H, prior=prior)
        np.testing.assert_allclose(gp, golden["exact%d_gp" % k], rtol=1e-9, atol=1e-300)
        assert int(np.argmax(gp)) == int(np.argmax(golden["exact%d_gp" % k]))
        fr = exact.posterior_allele_frequencies(golden["exact%d_gp" % k], P, H)
        np.testing.assert_allclose(np.stack(fr), golden["exact%d_fr" % k], rtol=1e-9, atol=1e-300)
        alt_g, alt_p = exact.alternate_dosage_posteriors(golden["exact%d_mode" % k], golden["exact%d_gp" % k])
        np.testing.assert_array_equal(alt_g, golden["exact%d_alt_g" % k])
        np.testing.assert_array_equal(alt_p, golden["exact%d_alt_p" % k])


@pytest.mark.parametrize("ploidy,n_haps,n_pos,prior_kind", [
    (6, 8, 8, None),          # BASELINE configs[2] shape: 1716 genotypes
    (6, 8, 8, "flat"),
    (4, 32, 8, "freqs"),      # configs[4] panel: 52360 genotypes
    (2, 5, 3, "freqs0"),
    (8, 4, 6, "flat"),
])
def test_call_exact_batch_vs_oracle(dev, oracle, ploidy, n_haps, n_pos, prior_kind):
    from mchap_b200.calling import exact
    from mchap_b200.synth import synth_haplotype_panel

    n_items = 6 if n_haps == 32 else 16
    batch, panels, truth = synth_haplotype_panel(n_items, n_haps, n_pos, ploidy, depth=40, seed=ploidy + n_haps)
    rng = np.random.default_rng(1)
    reads = [batch.item(i)[0] for i in range(n_items)]
    counts = [batch.item(i)[1] for i in range(n_items)]
    haps = [panels[i] for i in range(n_items)]
    priors = None
    if prior_kind is not None:
        priors = []
        for i in range(n_items):
            f = rng.random(n_haps) + 0.1
            f /= f.sum()
            priors.append({"flat": (0.1, None), "freqs": (0.2, f), "freqs0": (0.0, f)}[prior_kind])
    res = exact.posterior_mode_batch(reads, ploidy, haps, counts, priors)
    gls = exact.genotype_likelihoods_batch(reads, ploidy, haps, counts)
    for i in range(n_items):
        pr = None if priors is None else priors[i]
        want = oracle.posterior_mode(reads[i], ploidy, haps[i], counts[i], pr, True, True, True)
        got = res[i]
        np.testing.assert_array_equal(got[0], want[0], err_msg="item %d" % i)
        close([got[1], got[2], got[3]], [want[1], want[2], want[3]])
        close(got[4], want[4], rtol=1e-9)
        close(got[5], want[5], rtol=1e-9)
        gl64 = oracle.genotype_likelihoods(reads[i], ploidy, haps[i], counts[i], dtype=np.float64)
        np.testing.assert_allclose(gls[i], gl64.astype(np.float32), rtol=2e-7, atol=0)
        # size-independent properties at full enumeration size
        assert abs(got[4].sum() - 1.0) < 1e-9
        assert got[2] <= got[3] + 1e-12 <= 1.0 + 1e-9


# ----------------------------------------------------------------------------- call MCMC (K5)
def test_calling_mcmc_golden(dev, golden):
    """CallingMCMC.fit against reference fixtures: identical genotype traces, greedy start, llks."""
    from mchap_b200 import CallingMCMC

    for k in range(golden.meta["call_cases"]):
        meta = golden.meta["call%d" % k]
        reads = golden["call%d_reads" % k]
        counts = golden["call%d_counts" % k]
        haps = golden["call%d_haplotypes" % k]
        prior = _prior(golden, k, meta, "call")
        model = CallingMCMC(ploidy=meta["P"], haplotypes=haps, prior=prior, steps=meta["steps"],
                            chains=meta["chains"], random_seed=meta["seed"], step_type=meta["step_type"])
        trace = model.fit(reads, read_counts=counts)
        np.testing.assert_array_equal(trace.genotypes, golden["call%d_genotypes" % k], err_msg="case %d" % k)
        close(trace.llks, golden["call%d_llks" % k])
        assert trace.n_allele == len(haps)


def _exact_case(n_items, n_haps, n_pos, ploidies, depth, seed, counts_scale=1):
    from mchap_b200.synth import synth_haplotype_panel

    pmax = int(np.max(ploidies))
    batch, panels, _ = synth_haplotype_panel(n_items, n_haps, n_pos, pmax, depth=depth, seed=seed)
    reads = [batch.item(i)[0] for i in range(n_items)]
    counts = [batch.item(i)[1] * counts_scale for i in range(n_items)]
    return reads, counts, [panels[i] for i in range(n_items)]


def test_call_exact_mixed_ploidy_counts_and_reproducibility(dev, oracle):
    """Odd / mixed ploidies run the generic kernel; read counts above the product-form limit take
    log(rp) * count; zero reads; two runs of the same batch are bit-identical (fixed-order reductions)."""
    from mchap_b200.calling import exact

    ploidies = np.array([3, 5, 2, 7, 4, 6, 1, 3, 9, 4])
    reads, counts, haps = _exact_case(len(ploidies), 6, 5, ploidies, 30, 91)
    counts[1] = counts[1] * 40          # counts far above MCHB_EXACT_POW_MAX
    counts[4] = counts[4] * 16 + 1      # straddles the limit
    reads[7] = reads[7][:0]             # no reads at all: posterior = prior
    counts[7] = counts[7][:0]
    rng = np.random.default_rng(3)
    priors = []
    for i in range(len(ploidies)):
        f = rng.random(6) + 0.05
        priors.append([None, (0.15, None), (0.3, f / f.sum()), (0.0, f / f.sum())][i % 4])
    first = exact.posterior_mode_batch(reads, ploidies, haps, counts, priors)
    again = exact.posterior_mode_batch(reads, ploidies, haps, counts, priors)
    for i, P in enumerate(ploidies):
        want = oracle.posterior_mode(reads[i], int(P), haps[i], counts[i], priors[i], True, True, True)
        got = first[i]
        np.testing.assert_array_equal(got[0], want[0], err_msg="item %d" % i)
        close([got[1], got[2], got[3]], [want[1], want[2], want[3]])
        close(got[4], want[4])
        close(got[5], want[5])
        for x, y in zip(got, again[i]):
            np.testing.assert_array_equal(np.asarray(x), np.asarray(y))   # bit-identical reruns
    gls = exact.genotype_likelihoods_batch(reads, ploidies, haps, counts)
    for i, P in enumerate(ploidies):
        gl64 = oracle.genotype_likelihoods(reads[i], int(P), haps[i], counts[i], dtype=np.float64)
        np.testing.assert_allclose(gls[i], gl64.astype(np.float32), rtol=2e-7, atol=0)


def test_call_exact_impossible_reads(dev, oracle):
    """Per-read probabilities that are exactly zero: a read no allele can produce makes every genotype
    -inf (normaliser -inf, NaN posteriors, mode index 0 like the reference); a read only the all-zero
    haplotype can produce makes the genotypes without that haplotype -inf.  These leave the
    product form and take the reference's log(rp) * count."""
    from mchap_b200.calling import exact

    reads, counts, haps = _exact_case(4, 4, 4, np.array([4, 4, 4, 4]), 12, 5)
    for i, r in enumerate(reads):
        r[0, :, :] = 0.0
        if i % 2:
            r[0, :, 0] = 1.0
            haps[i][0, :] = 0
    res = exact.posterior_mode_batch(reads, 4, haps, counts, None)
    gls = exact.genotype_likelihoods_batch(reads, 4, haps, counts)
    for i in range(4):
        want = oracle.posterior_mode(reads[i], 4, haps[i], counts[i], None, True, True, True)
        gl64 = oracle.genotype_likelihoods(reads[i], 4, haps[i], counts[i], dtype=np.float64)
        np.testing.assert_array_equal(res[i][0], want[0])
        np.testing.assert_allclose(gls[i], gl64.astype(np.float32), rtol=2e-7, atol=0, equal_nan=True)
        close([res[i][1], res[i][2], res[i][3]], [want[1], want[2], want[3]])


def test_call_exact_low_memory_path(dev, oracle, monkeypatch):
    """No room for parked log joints (one large item must not size a scratch of hundreds of
    GB): the second pass evaluates the log joints again, like the reference's
    _posterior_allele_frequencies, and gives the same answers."""
    from mchap_b200.calling import exact

    reads, counts, haps = _exact_case(6, 8, 8, np.array([6] * 6), 40, 17)
    priors = [(0.1, None)] * 6
    parked = exact.posterior_mode_batch(reads, 6, haps, counts, priors)
    monkeypatch.setenv("MCHB_EXACT_SCRATCH_BYTES", "1024")
    twice = exact.posterior_mode_batch(reads, 6, haps, counts, priors)
    for i in range(6):
        want = oracle.posterior_mode(reads[i], 6, haps[i], counts[i], priors[i], True, True, True)
        np.testing.assert_array_equal(twice[i][0], want[0])
        close([twice[i][1], twice[i][2], twice[i][3]], [want[1], want[2], want[3]])
        close(twice[i][4], want[4])
        close(twice[i][5], want[5])
        np.testing.assert_array_equal(twice[i][0], parked[i][0])
        close(twice[i][4], parked[i][4], rtol=1e-12)


def test_genotype_posteriors_batch_vs_oracle(dev, oracle):
    from mchap_b200.calling import exact

    ploidies = np.array([4, 6, 2, 4, 3])
    reads, counts, haps = _exact_case(5, 7, 5, ploidies, 25, 23)
    rng = np.random.default_rng(8)
    f = rng.random(7) + 0.1
    priors = [None, (0.1, None), (0.25, f / f.sum()), (0.0, f / f.sum()), (0.4, None)]
    gls = exact.genotype_likelihoods_batch(reads, ploidies, haps, counts)
    gps, trips = exact.genotype_posteriors_batch(gls, ploidies, [7] * 5, priors, with_frequencies=True)
    again, _ = exact.genotype_posteriors_batch(gls, ploidies, [7] * 5, priors, with_frequencies=True)
    for i, P in enumerate(ploidies):
        want = oracle.genotype_posteriors(gls[i], int(P), 7, priors[i])
        np.testing.assert_allclose(gps[i], want, rtol=1e-9, atol=1e-300)
        wf = oracle.posterior_allele_frequencies(want, int(P), 7)
        for got, w in zip(trips[i], wf):
            np.testing.assert_allclose(got, w, rtol=1e-9, atol=1e-300)
        np.testing.assert_array_equal(gps[i], again[i])
        one = exact.genotype_posteriors(gls[i], int(P), 7, priors[i])
        np.testing.assert_array_equal(one, gps[i])


def test_minimum_error_correction_batch(dev):
    rng = np.random.default_rng(12)
    calls, genos = [], []
    for i in range(40):
        R, N, P = int(rng.integers(0, 90)), int(rng.integers(1, 12)), int(rng.integers(1, 9))
        c = rng.integers(-1, 3, size=(R, N)).astype(np.int8)
        g = rng.integers(-1, 3, size=(P, N)).astype(np.int8)
        calls.append(c)
        genos.append(g)
    mec, called, rows = dev.minimum_error_correction_batch(calls, genos, per_read=True)
    for i, (c, g) in enumerate(zip(calls, genos)):
        # encoding/integer/stats.py:18-39 restated with numpy
        diff = (c[:, None, :] != g[None, :, :]) & (c[:, None, :] >= 0)
        want = diff.sum(axis=-1).min(axis=-1)
        np.testing.assert_array_equal(rows[i], want)
        assert mec[i] == want.sum() and called[i] == (c >= 0).sum()


@pytest.mark.parametrize("ploidy,n_haps,step_type,prior_kind", [
    (4, 32, "Gibbs", None),           # BASELINE configs[4] shape
    (4, 32, "Gibbs", "freqs"),
    (6, 8, "Gibbs", "flat"),
    (4, 8, "Metropolis-Hastings", "freqs"),
    (2, 40, "Gibbs", None),           # more than one round of 32 candidate alleles
])
def test_calling_mcmc_batch_vs_oracle(dev, oracle, ploidy, n_haps, step_type, prior_kind):
    from mchap_b200 import CallingMCMC
    from mchap_b200.synth import synth_haplotype_panel

    n_items, steps = 10, 80
    batch, panels, truth = synth_haplotype_panel(n_items, n_haps, 8, ploidy, depth=40, seed=3 * ploidy + n_haps)
    rng = np.random.default_rng(2)
    reads = [batch.item(i)[0] for i in range(n_items)]
    counts = [batch.item(i)[1] for i in range(n_items)]
    haps = [panels[i] for i in range(n_items)]
    priors = None
    if prior_kind is not None:
        priors = []
        for i in range(n_items):
            f = rng.random(n_haps) + 0.1
            f /= f.sum()
            priors.append({"flat": (0.15, None), "freqs": (0.2, f)}[prior_kind])
    model = CallingMCMC(ploidy=ploidy, haplotypes=None, steps=steps, chains=2, random_seed=5, step_type=step_type)
    traces, results = model.fit_batch(reads, counts, haplotypes_list=haps, priors=priors, return_results=True)
    for i in range(n_items):
        ref = oracle.calling_fit(reads[i], counts[i], ploidy, haps[i], prior=None if priors is None else priors[i],
                                 steps=steps, chains=2, random_seed=5, step_type=step_type)
        np.testing.assert_array_equal(traces[i].genotypes, ref["genotypes"], err_msg="item %d" % i)
        close(traces[i].llks, ref["llks"])
        assert results["rng_words"][i] == ref["words"]


# ----------------------------------------------------------------------------- screening stress
@pytest.mark.parametrize("ploidy,n_pos,depth,error_rate,inbreeding,temps,seed", [
    (4, 8, 5, 0.0024, None, (1.0,), 1),       # very low depth: many plausible proposals
    (4, 8, 12, 0.05, None, (1.0,), 2),        # noisy reads
    (4, 8, 40, 0.0024, None, (1.0,), 3),      # headline shape
    (4, 8, 400, 0.0024, None, (1.0,), 4),     # deep: large read counts, huge |llk|
    (2, 12, 20, 0.01, 0.2, (1.0,), 5),
    (6, 6, 30, 0.02, None, (0.3, 1.0), 6),    # heated chain accepts a lot
    (4, 10, 25, 0.0024, 0.05, (0.1, 0.5, 1.0), 7),
])
def test_assemble_screening_stress_vs_oracle(dev, oracle, ploidy, n_pos, depth, error_rate, inbreeding, temps, seed):
    """The float32 screening passes must never change a decision: long seeded runs over many
    items (duplicated haplotypes, low / high depth, noisy reads) stay step-for-step identical."""
    from mchap_b200 import DenovoMCMC
    from mchap_b200.synth import _synth_from_haplotypes

    rng = np.random.default_rng(seed)
    n_items, steps = 40, 150
    # true haplotypes drawn from a small pool so that duplicated haplotypes (dosage > 1) are common
    pool = rng.integers(0, 2, size=(n_items, max(2, ploidy // 2 + 1), n_pos), dtype=np.int8)
    pick = rng.integers(0, pool.shape[1], size=(n_items, ploidy))
    haps = np.take_along_axis(pool, pick[:, :, None], axis=1)
    batch = _synth_from_haplotypes(haps, depth, 2, error_rate, rng)
    model = DenovoMCMC(ploidy=ploidy, n_alleles=[2] * n_pos, inbreeding=inbreeding, steps=steps, chains=2,
                       temperatures=temps, random_seed=seed)
    reads = [batch.item(i)[0] for i in range(n_items)]
    counts = [batch.item(i)[1] for i in range(n_items)]
    out, results = model.fit_batch(reads, counts, return_results=True, raw=True)
    for i in range(n_items):
        ref = _oracle_fit(oracle, model, reads[i], counts[i], [2] * n_pos)
        np.testing.assert_array_equal(out[i][0], ref["genotypes"], err_msg="item %d" % i)
        close(out[i][1], ref["llks"])
        assert results["rng_words"][i] == ref["words"]
        assert results["llk_evals"][i] == ref["llk_evals"]


def _adversarial_reads(kind, rng, n_items, ploidy, n_pos, depth):
    """Bi-allelic items (the float32-screened fast path) whose read probabilities stress the
    screening arithmetic; returns (reads, counts) lists in the reference's encoding."""
    reads, counts = [], []
    for _ in range(n_items):
        haps = rng.integers(0, 2, size=(ploidy, n_pos))
        haps[rng.integers(0, ploidy)] = haps[0]                       # a duplicated haplotype
        src = rng.integers(0, ploidy, size=depth)
        calls = haps[src].copy()
        flip = rng.random(calls.shape) < 0.02
        calls[flip] ^= 1
        start = rng.integers(0, max(n_pos // 2, 1), size=depth)
        width = rng.integers(max(n_pos // 2, 1), n_pos + 1, size=depth)
        cols = np.arange(n_pos)[None, :]
        gap = (cols < start[:, None]) | (cols >= (start + width)[:, None])
        if kind == "phred":
            # --use-base-phred-scores with --base-error-rate 0: p = 1 - 10^(-q/10), q = 2 .. 41
            q = rng.integers(2, 42, size=calls.shape)
            p = 1.0 - 10.0 ** (q / -10.0)
            other = (1.0 - p) / 3.0
        elif kind == "wide":
            # probabilities anywhere in 1e-9 .. 1 on both alleles: ratios R_new / R_old from 1e-9 to 1e9
            p = 10.0 ** rng.uniform(-9, 0, size=calls.shape)
            other = 10.0 ** rng.uniform(-9, 0, size=calls.shape)
        elif kind == "zeros":
            # error-free calls: the other allele has probability exactly 0 (ratio 0 or inf); some
            # reads are left uninformative so that the chain is not pinned to -inf everywhere
            p = np.ones(calls.shape)
            other = np.zeros(calls.shape)
            soft = rng.random(calls.shape) < 0.6
            p[soft], other[soft] = 0.9976, 0.0008
        elif kind == "ratio_one":
            # nearly uninformative reads: ratios within 1e-6 of one, screened differences ~ rounding
            p = 0.5 + rng.uniform(-1e-6, 1e-6, size=calls.shape)
            other = 0.5 + rng.uniform(-1e-6, 1e-6, size=calls.shape)
        else:
            p = np.full(calls.shape, 0.9976)
            other = np.full(calls.shape, 0.0008)
        r = np.empty(calls.shape + (2,))
        r[..., 0] = np.where(calls == 0, p, other)
        r[..., 1] = np.where(calls == 1, p, other)
        r[gap] = np.nan
        if kind == "counts":
            # few distinct reads with counts up to 10^4
            r = r[:12]
            c = rng.integers(1, 10001, size=len(r))
        else:
            c = np.ones(len(r), dtype=np.int64)
        reads.append(r)
        counts.append(c.astype(np.int64))
    return reads, counts


@pytest.mark.parametrize("kind,ploidy,n_pos,depth,temps,inbreeding", [
    ("phred", 4, 8, 30, (1.0,), None),
    ("phred", 4, 8, 60, (0.3, 1.0), 0.1),        # more than 32 distinct reads: two-chunk kernel
    ("wide", 4, 8, 24, (1.0,), None),
    ("wide", 2, 12, 20, (0.05, 1.0), None),
    ("zeros", 4, 6, 20, (1.0,), None),
    ("zeros", 4, 6, 20, (0.01, 0.2, 1.0), 0.2),
    ("ratio_one", 4, 8, 30, (1.0,), None),
    ("counts", 4, 8, 30, (1.0,), None),
    ("counts", 6, 6, 30, (0.01, 0.1, 0.5, 1.0), None),
    ("plain", 8, 4, 30, (0.01, 1.0), None),       # coldest temperature allowed by the CLIs' docs
    # three and four read chunks: Rt and the product rows in global memory, one direction of the ratio
    # table (the other is its reciprocal), the serial exact tier only
    ("phred", 4, 8, 90, (1.0,), None),
    ("phred", 6, 10, 120, (0.1, 1.0), 0.1),
    ("wide", 4, 8, 100, (0.2, 1.0), None),
    ("zeros", 4, 6, 110, (0.05, 1.0), None),
    ("ratio_one", 8, 6, 70, (1.0,), None),
])
def test_assemble_screening_adversarial_vs_oracle(dev, oracle, kind, ploidy, n_pos, depth, temps, inbreeding):
    """the float32 screening (bi-allelic fast path) must never change a decision on
    inputs outside the CLI defaults: per-base phred probabilities, probabilities over nine decades,
    zero-probability alleles, ratios next to one, counts up to 10^4, temperatures down to 0.01."""
    from mchap_b200 import DenovoMCMC

    seed = sum(map(ord, kind)) + ploidy + n_pos
    rng = np.random.default_rng(seed)
    n_items, steps = 24, 120
    reads, counts = _adversarial_reads(kind, rng, n_items, ploidy, n_pos, depth)
    model = DenovoMCMC(ploidy=ploidy, n_alleles=[2] * n_pos, inbreeding=inbreeding, steps=steps, chains=2,
                       temperatures=temps, random_seed=seed)
    out, results = model.fit_batch(reads, counts, return_results=True, raw=True, errors="return")
    screened = 0
    for i in range(n_items):
        try:
            ref = _oracle_fit(oracle, model, reads[i], counts[i], [2] * n_pos)
        except Exception as e:  # the reference's own failure modes (NaN log-likelihood ...) must match too
            assert isinstance(out[i], type(e)), "item %d: oracle raised %r, device gave %r" % (i, e, out[i])
            continue
        assert not isinstance(out[i], BaseException), "item %d: %r" % (i, out[i])
        np.testing.assert_array_equal(out[i][0], ref["genotypes"], err_msg="item %d" % i)
        close(out[i][1], ref["llks"])
        assert results["rng_words"][i] == ref["words"]
        assert results["llk_evals"][i] == ref["llk_evals"]
        screened += 1
    assert screened >= n_items // 2


@pytest.mark.parametrize("step_type", ["Gibbs", "Metropolis-Hastings"])
def test_calling_mcmc_replay_harness(dev, oracle, step_type):
    """North-star replay bar for the calling sampler: device and oracle driven by the same arbitrary
    pre-drawn 32-bit word stream (not an MT19937 stream) give identical traces and consume the same
    number of words."""
    from mchap_b200.calling import CallingMCMC
    from mchap_b200.synth import synth_haplotype_panel

    n_items, P, H = 6, 4, 12
    batch, panels, _ = synth_haplotype_panel(n_items, H, 8, P, depth=30, seed=41)
    reads = [batch.item(i)[0] for i in range(n_items)]
    counts = [batch.item(i)[1] for i in range(n_items)]
    words = np.random.default_rng(2024).integers(0, 2 ** 32, size=200000, dtype=np.uint64).astype(np.uint32)
    model = CallingMCMC(ploidy=P, haplotypes=panels[0], steps=150, chains=2, random_seed=1, step_type=step_type,
                        prior=(0.1, None))
    traces, results = model.fit_batch(reads, counts, haplotypes_list=list(panels), return_results=True,
                                      replay_words=words)
    for i in range(n_items):
        ref = oracle.calling_fit(reads[i], counts[i], P, panels[i], prior=(0.1, None), steps=150, chains=2,
                                 random_seed=1, step_type=step_type, replay_words=words)
        np.testing.assert_array_equal(traces[i].genotypes, ref["genotypes"], err_msg="item %d" % i)
        close(traces[i].llks, ref["llks"])
        assert results["rng_words"][i] == ref["words"]
        assert results["rng_words"][i] < len(words)


@pytest.mark.parametrize("ploidy,n_pos,n_alleles,depth,temps", [
    (4, 8, 2, 40, (1.0,)),
    (6, 5, 4, 30, (0.2, 1.0)),     # two bits per allele, a heated replica
    (8, 10, 3, 70, (1.0,)),        # three read chunks (Rt in global memory)
    (2, 1, 2, 10, (1.0,)),
])
def test_fit_batch_sorts_haplotypes_on_the_device(dev, ploidy, n_pos, n_alleles, depth, temps):
    """fit_batch returns GenotypeMultiTrace objects whose steps were sorted by the kernel while it
    recorded them (mchb_assemble_params.sort_haplotypes): identical to the host-side lexsort of the
    raw trace (reference: assemble/classes.py:265-278), fixed positions included."""
    from mchap_b200 import DenovoMCMC
    from mchap_b200.assemble.classes import sort_haplotypes
    from mchap_b200.synth import synth_items

    n_items = 12
    batch = synth_items(n_items, ploidy=ploidy, n_pos=n_pos, depth=depth, n_alleles=n_alleles, seed=ploidy * n_pos)
    reads = [batch.item(i)[0] for i in range(n_items)]
    counts = [batch.item(i)[1] for i in range(n_items)]
    model = DenovoMCMC(ploidy=ploidy, n_alleles=[n_alleles] * n_pos, steps=200, chains=2, temperatures=temps,
                       random_seed=9, fix_homozygous=0.9)
    raw = model.fit_batch(reads, counts, raw=True)
    traces = model.fit_batch(reads, counts)
    changed = 0
    for i in range(n_items):
        want = sort_haplotypes(raw[i][0])
        np.testing.assert_array_equal(traces[i].genotypes, want, err_msg="item %d" % i)
        np.testing.assert_array_equal(traces[i].llks, raw[i][1])
        changed += int((want != raw[i][0]).any())
    assert changed > 0 or ploidy * n_pos <= 2


@pytest.mark.parametrize("depth,ploidy,n_pos,temps,inbreeding", [
    (300, 4, 8, (1.0,), None),        # 16 read chunks per lane
    (700, 4, 6, (0.5, 1.0), 0.1),     # 32 read chunks, a heated replica, Dirichlet-multinomial prior
    (1024, 2, 5, (1.0,), None),       # the limit
])
def test_assemble_hundreds_of_distinct_reads_vs_oracle(dev, oracle, depth, ploidy, n_pos, temps, inbreeding):
    """reads encoded from base qualities are nearly all distinct, so items with far more
    than 256 distinct reads are common in deep targeted data.  The 16- and 32-chunk kernels take up to
    1024 distinct reads; trajectories stay bit-identical to the oracle."""
    from mchap_b200 import DenovoMCMC

    rng = np.random.default_rng(depth)
    n_items, steps = 3, 60
    reads, counts = _adversarial_reads("phred", rng, n_items, ploidy, n_pos, depth)
    assert all(len(r) == depth for r in reads)
    model = DenovoMCMC(ploidy=ploidy, n_alleles=[2] * n_pos, inbreeding=inbreeding, steps=steps, chains=2,
                       temperatures=temps, random_seed=depth)
    out, results = model.fit_batch(reads, counts, return_results=True, raw=True)
    for i in range(n_items):
        ref = _oracle_fit(oracle, model, reads[i], counts[i], [2] * n_pos)
        np.testing.assert_array_equal(out[i][0], ref["genotypes"], err_msg="item %d" % i)
        close(out[i][1], ref["llks"])
        assert results["rng_words"][i] == ref["words"]
        assert results["llk_evals"][i] == ref["llk_evals"]
    # one read more than the limit: that item alone is refused
    big = reads + [np.concatenate([reads[0], reads[0][:1]] * 2)[:1025] if depth == 1024 else reads[0]]
    if depth == 1024:
        res = model.fit_batch(big, [None] * 4, errors="return", raw=True)
        assert isinstance(res[3], NotImplementedError) and not isinstance(res[0], BaseException)


@pytest.mark.parametrize("temps,inbreeding,depth", [
    ((0.01, 1.0), None, 30),
    ((0.02, 0.3, 1.0), 0.1, 30),
    ((0.01, 1.0), None, 90),      # three read chunks: tables in global memory
])
def test_assemble_heated_replicas_with_mixed_allele_counts(dev, oracle, temps, inbreeding, depth):
    """Hot mode (a replica that accepts most proposals decides its sub-steps one by one, screened for
    certain acceptance and rejection, the exact log-likelihood evaluated lazily) on items that mix
    bi-allelic positions with three- and four-allelic ones (those take the serial exact step in the
    middle of the hot loop), plus the interval-arithmetic draws of the structural steps."""
    from mchap_b200 import DenovoMCMC

    rng = np.random.default_rng(int(depth + 100 * temps[0] * 100))
    reads, counts, nalls = [], [], []
    for i in range(10):
        N = int(rng.integers(3, 8))
        A = 4
        na = rng.integers(2, A + 1, size=N).astype(np.int8)
        na[rng.integers(0, N)] = 2                      # at least one bi-allelic position
        haps = np.stack([rng.integers(0, na[j], size=4) for j in range(N)], axis=1)
        src = rng.integers(0, 4, size=depth)
        calls = haps[src]
        r = np.full((depth, N, A), 0.0008)
        np.put_along_axis(r, calls[:, :, None], 0.9976, axis=2)
        r *= 1 + rng.random(r.shape) * 1e-3                # all reads distinct
        for j in range(N):
            r[:, j, na[j]:] = 0
        r[rng.random((depth, N)) < 0.2] = np.nan
        reads.append(r)
        counts.append(np.ones(depth, dtype=np.int64))
        nalls.append(na)
    model = DenovoMCMC(ploidy=4, n_alleles=None, inbreeding=inbreeding, steps=80, chains=2, temperatures=temps,
                       random_seed=7)
    out, results = model.fit_batch(reads, counts, n_alleles_list=nalls, return_results=True, raw=True)
    for i in range(10):
        ref = _oracle_fit(oracle, model, reads[i], counts[i], nalls[i])
        np.testing.assert_array_equal(out[i][0], ref["genotypes"], err_msg="item %d" % i)
        close(out[i][1], ref["llks"])
        assert results["rng_words"][i] == ref["words"]
        assert results["llk_evals"][i] == ref["llk_evals"]
