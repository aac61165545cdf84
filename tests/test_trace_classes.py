"""Host-side trace containers against fixtures written by the reference's own
GenotypeMultiTrace (tests/golden/make_golden_traces.py; mchap/assemble/classes.py:247-376)."""
import os

import numpy as np
import pytest

from mchap_b200.assemble.classes import GenotypeMultiTrace, TraceTally

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def fixtures():
    return np.load(os.path.join(HERE, "golden", "reference_trace_classes.npz"))


def case_names(fx):
    return ["sampled%d" % i for i in range(int(fx["n_sampled"]))] + \
           ["synthetic%d" % i for i in range(int(fx["n_synthetic"]))]


def check_case(fx, name, trace):
    """trace: GenotypeMultiTrace-like built from the raw trace of case `name`."""
    np.testing.assert_array_equal(trace.genotypes, fx[name + "_sorted"])
    burnt = trace.burn(int(fx[name + "_burn"]))
    post = burnt.posterior()
    np.testing.assert_array_equal(post.genotypes, fx[name + "_post_genotypes"])
    np.testing.assert_array_equal(post.probabilities, fx[name + "_post_probs"])
    mode, prob = post.mode()
    np.testing.assert_array_equal(mode, fx[name + "_mode"])
    assert prob == float(fx[name + "_mode_prob"])
    sup = post.mode_genotype_support()
    np.testing.assert_array_equal(sup.genotypes, fx[name + "_support_genotypes"])
    np.testing.assert_array_equal(sup.probabilities, fx[name + "_support_probs"])
    for c, chain in enumerate(burnt.split()):
        cp = chain.posterior()
        np.testing.assert_array_equal(cp.genotypes, fx[name + "_chain%d_genotypes" % c])
        np.testing.assert_array_equal(cp.probabilities, fx[name + "_chain%d_probs" % c])
    got = [burnt.replicate_incongruence(threshold=t) for t in (0.6, 0.3, 0.05)]
    assert got == list(fx[name + "_incongruence"])


def test_host_trace_classes_match_reference(fixtures):
    names = case_names(fixtures)
    assert len(names) >= 20
    for name in names:
        trace = GenotypeMultiTrace(fixtures[name + "_raw"], fixtures[name + "_llks"])
        np.testing.assert_array_equal(trace.llks, fixtures[name + "_llks"])
        check_case(fixtures, name, trace)


def check_tally(fx, name, tally):
    """tally: TraceTally of the burnt trace of case `name`."""
    post = tally.posterior()
    np.testing.assert_array_equal(post.genotypes, fx[name + "_post_genotypes"])
    np.testing.assert_array_equal(post.probabilities, fx[name + "_post_probs"])
    sup = post.mode_genotype_support()
    np.testing.assert_array_equal(sup.genotypes, fx[name + "_support_genotypes"])
    np.testing.assert_array_equal(sup.probabilities, fx[name + "_support_probs"])
    for c, chain in enumerate(tally.split()):
        cp = chain.posterior()
        np.testing.assert_array_equal(cp.genotypes, fx[name + "_chain%d_genotypes" % c])
        np.testing.assert_array_equal(cp.probabilities, fx[name + "_chain%d_probs" % c])
    got = [tally.replicate_incongruence(threshold=t) for t in (0.6, 0.3, 0.05)]
    assert got == list(fx[name + "_incongruence"])


def test_host_tally_matches_reference(fixtures):
    for name in case_names(fixtures):
        trace = GenotypeMultiTrace(fixtures[name + "_raw"], fixtures[name + "_llks"])
        check_tally(fixtures, name, TraceTally.from_trace(trace.burn(int(fixtures[name + "_burn"]))))


@pytest.mark.gpu
def test_device_tally_matches_reference(fixtures):
    """mchb_trace_tally_batch on the raw (unsorted) fixture traces, all cases in one call."""
    import mchap_b200
    from mchap_b200.api import TALLY_ITEM_DTYPE

    dev = mchap_b200.default_device(0)
    names = case_names(fixtures)
    for max_unique in (512, 7):
        items = np.zeros(len(names), dtype=TALLY_ITEM_DTYPE)
        raws, go, so, to = [], 0, 0, 0
        for k, name in enumerate(names):
            raw = fixtures[name + "_raw"]
            C, S, P, N = raw.shape
            items[k] = (go, so, to, N, P, C, S, int(fixtures[name + "_burn"]), max_unique)
            raws.append(raw.ravel())
            go += raw.size
            so += max_unique * P * N
            to += max_unique * C
        geno = np.concatenate(raws)
        states = np.zeros(so, dtype=np.int8)
        counts = np.zeros(to, dtype=np.int32)
        first = np.zeros(to, dtype=np.int32)
        res = dev.trace_tally_call(items, geno, geno.size, states, counts, first)
        n_over = 0
        for k, name in enumerate(names):
            raw = fixtures[name + "_raw"]
            C, S, P, N = raw.shape
            want = TraceTally.from_trace(GenotypeMultiTrace(raw, fixtures[name + "_llks"]).burn(int(fixtures[name + "_burn"])))
            if len(want.states) > max_unique:
                assert res["status"][k] == 9  # MCHB_ITEM_TALLY_OVERFLOW
                n_over += 1
                continue
            assert res["status"][k] == 0
            u = int(res["n_het"][k])
            got = TraceTally(
                states[items["states_off"][k]:][: u * P * N].reshape(u, P, N),
                counts[items["tallies_off"][k]:][: u * C].reshape(u, C).astype(np.int64),
                first[items["tallies_off"][k]:][: u * C].reshape(u, C).astype(np.int64))
            np.testing.assert_array_equal(got.states, want.states)
            np.testing.assert_array_equal(got.counts, want.counts)
            np.testing.assert_array_equal(got.first, want.first)
            check_tally(fixtures, name, got)
        assert (n_over > 0) == (max_unique == 7)


@pytest.mark.gpu
def test_fit_posterior_batch_matches_fit_then_host_summaries():
    """Traces kept on the device + device tallies == traces brought to the host + host classes."""
    from mchap_b200 import DenovoMCMC
    from mchap_b200.synth import synth_items

    for ploidy, n_pos, depth, max_unique in [(4, 8, 40, 64), (4, 6, 4, 3), (6, 5, 6, 16), (2, 3, 10, 64)]:
        n_items = 20
        batch = synth_items(n_items, ploidy=ploidy, n_pos=n_pos, depth=depth, seed=7 * ploidy + n_pos)
        reads = [batch.item(i)[0] for i in range(n_items)]
        counts = [batch.item(i)[1] for i in range(n_items)]
        model = DenovoMCMC(ploidy=ploidy, n_alleles=[2] * n_pos, steps=150, chains=2, random_seed=3)
        traces = model.fit_batch(reads, counts)
        tallies = model.fit_posterior_batch(reads, counts, burn=50, max_unique=max_unique)
        for i in range(n_items):
            want = TraceTally.from_trace(traces[i].burn(50))
            np.testing.assert_array_equal(tallies[i].states, want.states, err_msg="item %d" % i)
            np.testing.assert_array_equal(tallies[i].counts, want.counts)
            np.testing.assert_array_equal(tallies[i].first, want.first)
            a, b = tallies[i].posterior(), traces[i].burn(50).posterior()
            np.testing.assert_array_equal(a.genotypes, b.genotypes)
            np.testing.assert_array_equal(a.probabilities, b.probabilities)
            assert tallies[i].replicate_incongruence(0.6) == traces[i].burn(50).replicate_incongruence(0.6)


# ----------------------------------------------------------------------------- calling traces
def calling_names(fx):
    return ["calling%d" % i for i in range(int(fx["n_calling"]))]


def check_calling(fx, name, burnt):
    """burnt: GenotypeAllelesMultiTrace or AllelesTraceTally of the burnt trace of case `name`."""
    post = burnt.posterior()
    np.testing.assert_array_equal(post.genotypes, fx[name + "_post_genotypes"])
    np.testing.assert_array_equal(post.probabilities, fx[name + "_post_probs"])
    alleles, gp, sp = post.mode(genotype_support=True)
    np.testing.assert_array_equal(alleles, fx[name + "_mode"])
    assert [gp, sp] == list(fx[name + "_mode_probs"])
    for c, chain in enumerate(burnt.split()):
        cp = chain.posterior()
        np.testing.assert_array_equal(cp.genotypes, fx[name + "_chain%d_genotypes" % c])
        np.testing.assert_array_equal(cp.probabilities, fx[name + "_chain%d_probs" % c])
    got = [burnt.replicate_incongruence(threshold=t) for t in (0.6, 0.3, 0.05)]
    assert got == list(fx[name + "_incongruence"])
    np.testing.assert_array_equal(np.stack(burnt.posterior_frequencies()), fx[name + "_freqs"])
    rel = burnt.relabel(fx[name + "_labels"])
    rp = rel.posterior()
    np.testing.assert_array_equal(rp.genotypes, fx[name + "_relabel_post_genotypes"])
    np.testing.assert_array_equal(rp.probabilities, fx[name + "_relabel_post_probs"])
    np.testing.assert_array_equal(np.stack(rel.posterior_frequencies()), fx[name + "_relabel_freqs"])


def test_host_calling_classes_match_reference(fixtures):
    from mchap_b200.calling.classes import AllelesTraceTally, GenotypeAllelesMultiTrace

    names = calling_names(fixtures)
    assert len(names) >= 8
    for name in names:
        trace = GenotypeAllelesMultiTrace(fixtures[name + "_genotypes"].astype(np.int64), fixtures[name + "_llks"],
                                          int(fixtures[name + "_n_allele"]))
        burnt = trace.burn(int(fixtures[name + "_burn"]))
        check_calling(fixtures, name, burnt)
        check_calling(fixtures, name, AllelesTraceTally.from_trace(burnt))


@pytest.mark.gpu
def test_device_calling_tally_matches_reference(fixtures):
    import mchap_b200
    from mchap_b200.api import TALLY_ITEM_DTYPE
    from mchap_b200.calling.classes import AllelesTraceTally, GenotypeAllelesMultiTrace

    dev = mchap_b200.default_device(0)
    names = calling_names(fixtures)
    max_unique = 64
    items = np.zeros(len(names), dtype=TALLY_ITEM_DTYPE)
    gens, go, so, to = [], 0, 0, 0
    for k, name in enumerate(names):
        g = fixtures[name + "_genotypes"]
        C, S, P = g.shape
        items[k] = (go, so, to, 1, P, C, S, int(fixtures[name + "_burn"]), max_unique)
        gens.append(g.ravel())
        go += g.size
        so += max_unique * P
        to += max_unique * C
    alleles = np.concatenate(gens).astype(np.int32)
    states = np.zeros(so, dtype=np.int32)
    counts = np.zeros(to, dtype=np.int32)
    first = np.zeros(to, dtype=np.int32)
    res = dev.call_trace_tally_call(items, alleles, alleles.size, states, counts, first)
    for k, name in enumerate(names):
        g = fixtures[name + "_genotypes"]
        C, S, P = g.shape
        assert res["status"][k] == 0
        u = int(res["n_het"][k])
        got = AllelesTraceTally(
            states[items["states_off"][k]:][: u * P].reshape(u, P).astype(np.int64),
            counts[items["tallies_off"][k]:][: u * C].reshape(u, C).astype(np.int64),
            first[items["tallies_off"][k]:][: u * C].reshape(u, C).astype(np.int64), int(fixtures[name + "_n_allele"]))
        want = AllelesTraceTally.from_trace(
            GenotypeAllelesMultiTrace(g.astype(np.int64), fixtures[name + "_llks"],
                                      int(fixtures[name + "_n_allele"])).burn(int(fixtures[name + "_burn"])))
        np.testing.assert_array_equal(got.states, want.states)
        np.testing.assert_array_equal(got.counts, want.counts)
        np.testing.assert_array_equal(got.first, want.first)
        check_calling(fixtures, name, got)


@pytest.mark.gpu
@pytest.mark.parametrize("step_type", ["Gibbs", "Metropolis-Hastings"])
def test_calling_fit_posterior_batch_matches_fit_then_host_summaries(step_type):
    from mchap_b200.calling import CallingMCMC
    from mchap_b200.calling.classes import AllelesTraceTally
    from mchap_b200.synth import synth_haplotype_panel

    for ploidy, n_haps, depth, max_unique in [(4, 12, 30, 64), (4, 6, 3, 2), (2, 5, 8, 64)]:
        n_items = 16
        batch, panels, _ = synth_haplotype_panel(n_items, n_haps, 6, ploidy, depth=depth, seed=ploidy + n_haps)
        reads = [batch.item(i)[0] for i in range(n_items)]
        counts = [batch.item(i)[1] for i in range(n_items)]
        model = CallingMCMC(ploidy=ploidy, haplotypes=panels[0], steps=200, chains=2, random_seed=9,
                            step_type=step_type, prior=(0.1, None))
        traces = model.fit_batch(reads, counts, haplotypes_list=list(panels))
        tallies = model.fit_posterior_batch(reads, counts, burn=50, haplotypes_list=list(panels), max_unique=max_unique)
        for i in range(n_items):
            burnt = traces[i].burn(50)
            want = AllelesTraceTally.from_trace(burnt)
            np.testing.assert_array_equal(tallies[i].states, want.states, err_msg="item %d" % i)
            np.testing.assert_array_equal(tallies[i].counts, want.counts)
            np.testing.assert_array_equal(tallies[i].first, want.first)
            a, b = tallies[i].posterior(), burnt.posterior()
            np.testing.assert_array_equal(a.genotypes, b.genotypes)
            np.testing.assert_array_equal(a.probabilities, b.probabilities)
            assert tallies[i].replicate_incongruence(0.6) == burnt.replicate_incongruence(0.6)
            np.testing.assert_array_equal(np.stack(tallies[i].posterior_frequencies()),
                                          np.stack(burnt.posterior_frequencies()))


def test_break_point_tables_match_reference(fixtures):
    """The break-point distributions handed to the device (host-side scipy, like the reference's
    _point_beta_probabilities, assemble/mcmc.py:429-452) and the rows of break_table built from them."""
    from mchap_b200.assemble.mcmc import break_table, point_beta_probabilities

    n_cases = int(fixtures["n_beta"])
    assert n_cases >= 60
    for k in range(n_cases):
        n, a, b = fixtures["beta%d_par" % k]
        want = fixtures["beta%d_out" % k]
        np.testing.assert_array_equal(point_beta_probabilities(int(n), a, b), want)
        if int(n) <= 32:
            table, lens = break_table(int(n), a, b)
            assert lens[int(n)] == len(want)
            np.testing.assert_array_equal(table[int(n), : len(want)], want)
    table, lens = break_table(6, n_intervals=3)   # fixed number of intervals (mcmc.py:214-217)
    assert list(lens[1:]) == [3] * 6
    np.testing.assert_array_equal(table[4, :3], [0.0, 0.0, 1.0])


# ----------------------------------------------------------------------------- read encoding (N2)
def encode_names(fx):
    return ["encode%d" % i for i in range(int(fx["n_encode"]))]


def test_call_probabilities_match_reference(fixtures):
    from mchap_b200.encoding import call_probabilities

    for name in encode_names(fixtures):
        got = call_probabilities(fixtures[name + "_calls"], fixtures[name + "_quals"], float(fixtures[name + "_error_rate"]))
        np.testing.assert_array_equal(got, fixtures[name + "_probs"])


@pytest.mark.gpu
def test_device_read_encoding_matches_reference(fixtures):
    """mchb_encode_reads_batch: as_probabilistic + unique_counts, byte for byte (NaN gaps included),
    unique reads in the reference's first-occurrence order; all cases in one call."""
    from mchap_b200.encoding import encode_unique_reads_batch

    names = encode_names(fixtures)
    out = encode_unique_reads_batch(
        [fixtures[n + "_calls"] for n in names], [fixtures[n + "_probs"] for n in names],
        [fixtures[n + "_n_alleles"] for n in names])
    n_dup = 0
    for name, (reads, counts) in zip(names, out):
        want_r, want_c = fixtures[name + "_unique"], fixtures[name + "_counts"]
        assert reads.shape == want_r.shape, name
        assert reads.tobytes() == np.ascontiguousarray(want_r).tobytes(), name
        np.testing.assert_array_equal(counts, want_c)
        n_dup += int((want_c > 1).sum())
    assert n_dup > 0
    # a different error factor and scalar probabilities
    calls = fixtures["encode2_calls"]
    (reads, counts), = encode_unique_reads_batch([calls], [0.9], [fixtures["encode2_n_alleles"]], error_factor=2)
    assert np.nanmax(reads) == 0.9 and np.isclose(np.nanmin(reads[reads > 0]), 0.05)
    assert counts.sum() == len(calls)


@pytest.mark.gpu
def test_chained_device_path_matches_the_separate_calls(fixtures):
    """calls + probabilities -> encoded unique reads -> de novo assembly -> tallies in ONE call, with
    the intermediates kept on the device == the three calls with host arrays in between."""
    from mchap_b200 import DenovoMCMC
    from mchap_b200.encoding import encode_unique_reads_batch

    names = [n for n in encode_names(fixtures) if fixtures[n + "_calls"].shape[0] > 0 and
             fixtures[n + "_unique"].shape[0] <= 256]
    assert len(names) >= 5
    calls = [fixtures[n + "_calls"] for n in names]
    probs = [fixtures[n + "_probs"] for n in names]
    nalls = [fixtures[n + "_n_alleles"] for n in names]
    model = DenovoMCMC(ploidy=4, n_alleles=None, steps=120, chains=2, random_seed=17)
    pairs = encode_unique_reads_batch(calls, probs, nalls)
    want = model.fit_posterior_batch([r for r, _ in pairs], [c for _, c in pairs], burn=40, n_alleles_list=nalls)
    got, n_unique = model.fit_posterior_from_calls_batch(calls, probs, burn=40, n_alleles_list=nalls)
    for i, name in enumerate(names):
        assert n_unique[i] == len(pairs[i][0]) == len(fixtures[name + "_unique"])
        np.testing.assert_array_equal(got[i].states, want[i].states, err_msg=name)
        np.testing.assert_array_equal(got[i].counts, want[i].counts)
        np.testing.assert_array_equal(got[i].first, want[i].first)
