"""Pin the C oracle to the real reference: every fixture in tests/golden/reference_golden.npz
was produced by the reference's own numba functions (tests/golden/make_golden.py).

Integers / trajectories: bit-exact.  fp64 values: exact on the same libm, asserted to 1e-12
relative so that a different glibc on another box cannot make the pin flaky.
"""
import numpy as np
import pytest

RTOL = 1e-12


def close(a, b, rtol=RTOL):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=0, equal_nan=True)


def test_rng_stream_matches_numba(golden, oracle):
    # numba randomimpl.py semantics: random, randint (rejection, n==1 no draw), shuffle (1-D and
    # rows), permutation, choice
    for seed in golden.meta["rng_seeds"]:
        ref = golden["rng_probe_%d" % seed]
        r = oracle.Rng(seed)
        out = []
        for _ in range(8):
            out.append(r.random())
        for n in (1, 2, 3, 5, 8, 31, 32, 33, 100, 1000, 70000):
            out.append(r.randint(n))
        out.extend(r.shuffle(np.arange(10)))
        out.extend(r.shuffle(np.arange(7)))
        out.append([3, 5, 9, 11][r.randint(4)])
        out.append(r.random())
        rows = r.shuffle(np.arange(6))
        out.extend(rows)
        out.append(r.random())
        np.testing.assert_array_equal(np.array(out, dtype=np.float64), ref[: len(out)])


def test_mt19937_words_equal_numpy_legacy(oracle):
    for seed in (0, 11, 42):
        w = oracle.mt19937_words(seed, 2000)
        ref = np.random.RandomState(seed).randint(0, 2 ** 32, 2000, dtype=np.uint64).astype(np.uint32)
        np.testing.assert_array_equal(w, ref)


def test_log_unique_haplotypes_is_float32(golden, oracle):
    # assemble/mcmc.py:294: numba evaluates np.log(int8 array).sum() in float32
    for k in range(golden.meta["luh_cases"]):
        assert oracle.log_unique_haplotypes(golden["luh%d_in" % k]) == golden["luh%d_out" % k][0]


def test_log_likelihood(golden, oracle):
    for k in golden.meta["llk_cases"]:
        reads = golden["llk%d_reads" % k]
        g = golden["llk%d_genotype" % k]
        counts = golden.get("llk%d_counts" % k)
        idx = golden["llk%d_idx" % k]
        a, b = golden["llk%d_interval" % k]
        llk, llk_sc = golden["llk%d_out" % k]
        assert oracle.log_likelihood(reads, g, counts) == llk or np.isclose(
            oracle.log_likelihood(reads, g, counts), llk, rtol=RTOL, atol=0)
        got = oracle.log_likelihood_structural_change(reads, g, idx, (a, b), counts)
        close(got, llk_sc)
        g2 = g.copy()
        oracle.structural_change(g2, idx, (a, b))
        np.testing.assert_array_equal(g2, golden["llk%d_changed" % k])
        # L2 == L1 on the changed genotype, exactly (tests/test_assemble/test_likelihood.py:176-191)
        assert got == oracle.log_likelihood(reads, g2, counts) or (np.isnan(got))


def test_comb_tables(golden, oracle):
    ct = golden["comb_table"]
    cw = golden["cwr_table"]
    for n in range(ct.shape[0]):
        for k in range(ct.shape[1]):
            assert oracle.comb(n, k) == ct[n, k]
            assert oracle.comb_with_replacement(n, k) == cw[n, k]
    big = golden["comb_big_in"]
    for (n, k), v in zip(big, golden["comb_big_out"]):
        assert oracle.comb(n, k) == v
    for (n, k), v in zip(big[:4], golden["cwr_big_out"]):
        assert oracle.comb_with_replacement(n, k) == v
    assert oracle.comb_with_replacement(0, 0) == 0  # jitutils.py:232-233


@pytest.mark.parametrize("P", [1, 2, 3, 4, 6, 8])
def test_rank_unrank_increment(golden, oracle, P):
    un = golden["unrank_p%d" % P]
    rk = golden["rank_p%d" % P]
    inc = golden["increment_p%d" % P]
    g = np.zeros(P, dtype=np.int64)
    for i in range(len(un)):
        np.testing.assert_array_equal(oracle.index_as_genotype_alleles(i, P), un[i])
        assert oracle.genotype_alleles_as_index(un[i]) == rk[i] == i
        np.testing.assert_array_equal(g, inc[i])
        oracle.increment_genotype(g)
    assert oracle.index_as_genotype_alleles(-1, P) is None


def test_unrank_large(golden, oracle):
    for i, ref in zip(golden["unrank_large_idx"], golden["unrank_large_p6"]):
        np.testing.assert_array_equal(oracle.index_as_genotype_alleles(int(i), 6), ref)
        assert oracle.genotype_alleles_as_index(ref) == i


def test_known_answer_tables(oracle):
    # reference tests/test_jitutils.py:220-265 / 268-311 and tests/test_combinatorics.py:18-46
    table = {
        (0, 0): 0, (0, 1): 1, (1, 1): 2, (0, 2): 3, (1, 2): 4, (2, 2): 5,
        (0, 0, 0): 0, (0, 0, 1): 1, (0, 1, 1): 2, (1, 1, 1): 3, (0, 0, 2): 4, (0, 1, 2): 5,
        (0, 0, 0, 0): 0, (0, 0, 0, 1): 1, (0, 0, 1, 1): 2, (0, 1, 1, 1): 3, (1, 1, 1, 1): 4,
        (0, 0, 0, 2): 5, (0, 0, 1, 2): 6,
    }
    for alleles, idx in table.items():
        assert oracle.genotype_alleles_as_index(np.array(alleles)) == idx
        np.testing.assert_array_equal(oracle.index_as_genotype_alleles(idx, len(alleles)), alleles)
    assert oracle.count_unique_genotypes(1024, 6) == 1624866254968320
    assert oracle.count_unique_genotypes(4, 4) == 35
    assert oracle.count_unique_genotypes(8, 6) == 1716
    assert oracle.count_unique_genotypes(32, 4) == 52360


def test_logspace_helpers(golden, oracle):
    x = golden["logspace_in"]
    close(oracle.sum_log_probs(x), golden["logspace_sum"][0])
    close(oracle.normalise_log_probs(x), golden["logspace_norm"])
    pairs = [(-np.inf, -np.inf), (-1.0, -np.inf), (-3.0, -2.5), (0.0, 0.0)]
    close([oracle.add_log_prob(a, b) for a, b in pairs], golden["logspace_add"])
    for d, ref in zip(golden["lnperm_in"], golden["lnperm_out"]):
        close(oracle.ln_equivalent_permutations(d), ref)


def test_assemble_prior(golden, oracle):
    for k in range(golden.meta["aprior_cases"]):
        g = golden["aprior%d_genotype" % k]
        d = np.zeros(len(g), dtype=np.int8)
        oracle.get_haplotype_dosage(d, g)
        np.testing.assert_array_equal(d, golden["aprior%d_dosage" % k])
        luh, inb = golden["aprior%d_par" % k]
        close(oracle.assemble_log_genotype_prior(d, luh, inb), golden["aprior%d_out" % k][0])


def test_assemble_prior_known_answers(oracle):
    # reference tests/test_assemble/test_prior.py:8-76 (tensorflow-probability / scipy values)
    from math import log

    def prior(dosage, unique_haplotypes, inbreeding):
        return oracle.assemble_log_genotype_prior(
            np.array(dosage, dtype=np.int8), log(unique_haplotypes), inbreeding)

    np.testing.assert_almost_equal(prior([1, 1, 1, 1], 16, 0.0), log(24 / 16 ** 4), 10)
    np.testing.assert_almost_equal(prior([4, 0, 0, 0], 16, 0.0), log(1 / 16 ** 4), 10)
    np.testing.assert_almost_equal(prior([2, 2, 0, 0], 16, 0.0), log(6 / 16 ** 4), 10)


def test_calling_prior(golden, oracle):
    for k in range(golden.meta["cprior_cases"]):
        g = golden["cprior%d_genotype" % k]
        H, inb, va, has_f = golden["cprior%d_par" % k]
        f = golden["cprior%d_freqs" % k] if has_f else None
        ref = golden["cprior%d_out" % k]
        close(oracle.calling_log_genotype_prior(g, int(H), inb, f), ref[0])
        close(oracle.log_genotype_allele_prior(g, int(va), int(H), inb, f), ref[1])
        close(oracle.log_genotype_allele_flat_prior(g, int(va)), ref[2])
        np.testing.assert_array_equal(oracle.allelic_dosage(g), golden["cprior%d_dosage" % k])


def test_structural_tables(golden, oracle):
    for k in range(golden.meta["struct_cases"]):
        g = golden["struct%d_genotype" % k]
        a, b = golden["struct%d_interval" % k]
        labels = oracle.haplotype_segment_labels(g, (a, b))
        np.testing.assert_array_equal(labels, golden["struct%d_labels" % k])
        np.testing.assert_array_equal(
            oracle.haplotype_segment_labels(g, None), golden["struct%d_labels_none" % k])
        ro = oracle.recombination_step_options(labels)
        do = oracle.dosage_step_options(labels)
        np.testing.assert_array_equal(ro, golden["struct%d_recomb" % k].reshape(ro.shape))
        np.testing.assert_array_equal(do, golden["struct%d_dosage" % k].reshape(do.shape))
        n = golden["struct%d_n" % k]
        assert oracle.recombination_step_n_options(labels) == n[0] == len(ro)
        assert oracle.dosage_step_n_options(labels) == n[1] == len(do)
        np.testing.assert_array_equal(
            [oracle.recombination_step_n_options(o) for o in ro], golden["struct%d_recomb_return" % k])
        np.testing.assert_array_equal(
            [oracle.dosage_step_n_options(o) for o in do], golden["struct%d_dosage_return" % k])


def test_structural_known_answers(oracle):
    # reference tests/test_assemble/test_structural.py:177-237 (labels), 240-348 (options)
    g = np.array([[0, 1, 0, 1], [0, 1, 1, 1], [0, 1, 1, 1], [0, 1, 0, 0]], dtype=np.int8)
    np.testing.assert_array_equal(
        oracle.haplotype_segment_labels(g, (0, 3)), [[0, 0], [1, 0], [1, 0], [0, 3]])
    np.testing.assert_array_equal(
        oracle.haplotype_segment_labels(g, None), [[0, 0], [1, 0], [1, 0], [3, 0]])
    labels = np.array([[0, 0], [0, 1], [2, 1], [3, 0]], dtype=np.int8)
    np.testing.assert_array_equal(
        oracle.recombination_step_options(labels),
        [[[2, 0], [0, 1], [0, 1], [3, 0]], [[0, 0], [3, 1], [2, 1], [0, 0]], [[0, 0], [0, 1], [3, 1], [2, 0]]])
    labels = np.array([[0, 0], [0, 1], [2, 0], [2, 0]], dtype=np.int8)
    np.testing.assert_array_equal(
        oracle.dosage_step_options(labels),
        [[[2, 0], [0, 1], [2, 0], [2, 0]], [[0, 0], [2, 1], [2, 0], [2, 0]], [[0, 0], [0, 1], [0, 0], [2, 0]]])


def test_random_breaks(golden, oracle):
    for k in range(golden.meta["breaks_cases"]):
        seed, breaks, n = golden["breaks%d_par" % k]
        r = oracle.Rng(int(seed))
        iv = oracle.random_breaks(r, int(breaks), int(n))
        np.testing.assert_array_equal(iv, golden["breaks%d_out" % k])
        assert r.random() == golden["breaks%d_next" % k][0]
    with pytest.raises(ValueError):
        oracle.random_breaks(oracle.Rng(1), 5, 5)


def _step_inputs(golden, k):
    reads = golden["step%d_reads" % k]
    counts = golden["step%d_counts" % k]
    na = golden["step%d_nalleles" % k]
    g0 = golden["step%d_g0" % k]
    P, N, A, inb, temp, use_counts, seed, luh, llk0 = golden["step%d_par" % k]
    inb = None if inb < 0 else float(inb)
    rc = counts if use_counts else None
    return reads, rc, na, g0, inb, float(temp), int(seed), float(luh), float(llk0)


def test_mcmc_steps(golden, oracle):
    for k in range(golden.meta["step_cases"]):
        reads, rc, na, g0, inb, temp, seed, luh, llk0 = _step_inputs(golden, k)
        close(oracle.log_likelihood(reads, g0, rc), llk0)
        # mutation.base_step
        h, j, llk, nxt = golden["step%d_base_out" % k]
        g = g0.copy()
        r = oracle.Rng(seed)
        got = oracle.mutation_base_step(r, g, reads, llk0, int(h), int(j), int(na[int(j)]), luh, inb, temp, rc)
        np.testing.assert_array_equal(g, golden["step%d_base" % k])
        close(got, llk)
        assert r.random() == nxt
        # mutation.compound_step
        llk, nxt = golden["step%d_mut_out" % k]
        g = g0.copy()
        r = oracle.Rng(seed)
        got = oracle.mutation_compound_step(r, g, reads, llk0, na, luh, inb, temp, rc)
        np.testing.assert_array_equal(g, golden["step%d_mut" % k])
        close(got, llk)
        assert r.random() == nxt
        # structural.interval_step
        for st, name in ((0, "recomb"), (1, "dosage")):
            a, b, llk, nxt = golden["step%d_%s_out" % (k, name)]
            g = g0.copy()
            r = oracle.Rng(seed)
            got = oracle.structural_interval_step(r, g, reads, llk0, luh, inb, (int(a), int(b)), st, temp, rc)
            np.testing.assert_array_equal(g, golden["step%d_%s" % (k, name)])
            close(got, llk)
            assert r.random() == nxt
        # structural.compound_step after random_breaks
        for st, name in ((0, "crecomb"), (1, "cdosage")):
            llk, nxt = golden["step%d_%s_out" % (k, name)]
            g = g0.copy()
            r = oracle.Rng(seed)
            N = g.shape[1]
            iv = oracle.random_breaks(r, min(2, N - 1), N)
            np.testing.assert_array_equal(iv, golden["step%d_%s_iv" % (k, name)])
            got = oracle.structural_compound_step(r, g, reads, llk0, iv, luh, inb, st, temp, rc)
            np.testing.assert_array_equal(g, golden["step%d_%s" % (k, name)])
            close(got, llk)
            assert r.random() == nxt


def test_snp_posterior_and_read_mean(golden, oracle):
    for k in range(golden.meta["snp_cases"]):
        reads = golden["snp%d_reads" % k]
        counts = golden["snp%d_counts" % k]
        na = golden["snp%d_nalleles" % k]
        P, inb = golden["snp%d_par" % k]
        inb = None if inb < 0 else float(inb)
        hom = oracle.homozygosity_probabilities(reads, na, int(P), inb, counts)
        close(hom, golden["snp%d_hom" % k])
        close(oracle.read_mean_dist(reads), golden["snp%d_meandist" % k])


def sort_trace(genotypes):
    """Canonical haplotype order per step, as GenotypeMultiTrace.__post_init__ does
    (assemble/classes.py:265-278 -> encoding/integer/sequence.py:78-110: lexicographic rows)."""
    out = genotypes.copy()
    C, S = out.shape[:2]
    for c in range(C):
        for i in range(S):
            g = out[c, i]
            out[c, i] = g[np.lexsort(g.T[::-1])]
    return out


def _fit_kwargs(meta):
    return dict(
        ploidy=meta["P"], inbreeding=meta["inbreeding"], steps=meta["steps"], chains=meta["chains"],
        n_intervals=meta["n_intervals"], fix_homozygous=meta["fix_homozygous"],
        recombination_step_probability=meta["probs"][0],
        partial_dosage_step_probability=meta["probs"][1],
        dosage_step_probability=meta["probs"][2], temperatures=meta["temperatures"],
        random_seed=meta["seed"],
    )


def test_denovo_fit_trajectories(golden, oracle):
    """DenovoMCMC.fit: step-for-step identical traces, llks, and RNG position afterwards."""
    for k in range(golden.meta["fit_cases"]):
        meta = golden.meta["fit%d" % k]
        reads = golden["fit%d_reads" % k]
        counts = golden.get("fit%d_counts" % k)
        na = golden["fit%d_nalleles" % k]
        res = oracle.denovo_fit(reads, counts, n_alleles=na, **_fit_kwargs(meta))
        if meta.get("edge"):
            got = sort_trace(res["genotypes"])
            np.testing.assert_array_equal(got, golden["fit%d_sorted" % k])
            close(res["llks"], golden["fit%d_llks" % k])
            continue
        np.testing.assert_array_equal(res["genotypes"], golden["fit%d_genotypes" % k], err_msg="case %d" % k)
        close(res["llks"], golden["fit%d_llks" % k])
        assert res["n_het"] == golden["fit%d_nhet" % k][0]
        # the reference's numba RNG sits exactly res["words"] words into the stream
        r = oracle.Rng(meta["seed"])
        for _ in range(res["words"]):
            r.u32()
        assert r.random() == golden["fit%d_next" % k][0]
        # replay harness: same trajectory from a pre-drawn word stream
        words = oracle.mt19937_words(meta["seed"], res["words"])
        res2 = oracle.denovo_fit(reads, counts, n_alleles=na, replay_words=words, **_fit_kwargs(meta))
        np.testing.assert_array_equal(res2["genotypes"], res["genotypes"])
        assert res2["words"] == res["words"]


def _call_prior(golden, k, meta, name):
    if meta["inbreeding"] is None:
        return None
    return (meta["inbreeding"], golden["%s%d_freqs" % (name, k)] if meta["with_freqs"] else None)


def test_calling_mcmc(golden, oracle):
    for k in range(golden.meta["call_cases"]):
        meta = golden.meta["call%d" % k]
        reads = golden["call%d_reads" % k]
        counts = golden["call%d_counts" % k]
        haps = golden["call%d_haplotypes" % k]
        prior = _call_prior(golden, k, meta, "call")
        greedy = oracle.greedy_caller(haps, meta["P"], reads, counts, prior)
        np.testing.assert_array_equal(greedy, golden["call%d_greedy" % k])
        res = oracle.calling_fit(reads, counts, meta["P"], haps, prior=prior, steps=meta["steps"],
                                 chains=meta["chains"], random_seed=meta["seed"], step_type=meta["step_type"])
        np.testing.assert_array_equal(res["genotypes"], golden["call%d_genotypes" % k], err_msg="case %d" % k)
        close(res["llks"], golden["call%d_llks" % k])
        r = oracle.Rng(meta["seed"])
        for _ in range(res["words"]):
            r.u32()
        assert r.random() == golden["call%d_next" % k][0]
        for st, name in ((0, "gibbs"), (1, "mh")):
            ref = golden["call%d_%s" % (k, name)]
            llks, lpr, pr = oracle.calling_step_options(greedy, 1, haps, reads, counts, prior, st)
            close(llks, ref[0])
            close(lpr, ref[1])
            close(pr, ref[2], rtol=1e-10)


def test_call_exact(golden, oracle):
    for k in range(golden.meta["exact_cases"]):
        meta = golden.meta["exact%d" % k]
        reads = golden["exact%d_reads" % k]
        counts = golden["exact%d_counts" % k]
        haps = golden["exact%d_haplotypes" % k]
        prior = _call_prior(golden, k, meta, "exact")
        P, H = meta["P"], meta["H"]
        assert oracle.count_unique_genotypes(H, P) == golden["exact%d_ngen" % k][0]
        mode, llk, prob, support, freqs, occur = oracle.posterior_mode(
            reads, P, haps, counts, prior, True, True, True)
        np.testing.assert_array_equal(mode, golden["exact%d_mode" % k])
        close([llk, prob, support], golden["exact%d_scalars" % k], rtol=1e-11)
        close(freqs, golden["exact%d_mode_freqs" % k], rtol=1e-11)
        close(occur, golden["exact%d_mode_occur" % k], rtol=1e-11)
        gl = oracle.genotype_likelihoods(reads, P, haps, counts)
        assert gl.dtype == np.float32
        np.testing.assert_array_equal(gl, golden["exact%d_gl" % k])
        gp = oracle.genotype_posteriors(gl, P, H, prior)
        close(gp, golden["exact%d_gp" % k], rtol=1e-10)
        fr = oracle.posterior_allele_frequencies(golden["exact%d_gp" % k], P, H)
        close(np.stack(fr), golden["exact%d_fr" % k], rtol=1e-12)
