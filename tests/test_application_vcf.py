"""BASELINE configs[0]: the three CLIs on the reference's bundled test data must print the reference's
golden VCFs line for line (##commandline / ##source / ##fileDate exempt, like the reference's own
tests: test_application_assemble.py:254-437, test_application_call.py:16-200,
test_application_call_exact.py:16-216; `--mcmc-steps 500 --mcmc-burn 100 --mcmc-seed 11`).

Every scenario runs twice: on the CPU with the device entry points replaced by the oracle
(tests/oracle_engine.py) — this pins the host side: readers, read extraction, summaries, VCF text —
and, marked `gpu`, through the real CUDA library.
"""
import io
import os

import numpy as np
import pytest

from mchap_b200.application import build_program, hostio, vcfout
from mchap_b200.application.loci import Locus, LocusPrior
from mchap_b200.application.reads import extract_read_variants, qual_of_prob

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg1")


def path(name):
    return os.path.join(DATA, name)


SHALLOW = ["simple.sample1.bam", "simple.sample2.bam", "simple.sample3.bam"]
DEEP = ["simple.sample1.deep.bam", "simple.sample2.deep.bam", "simple.sample3.deep.bam"]
MIXED = ["simple.sample1.bam", "simple.sample2.deep.bam", "simple.sample3.bam"]
POOLS = ["--ploidy", path("simple.pools-ploidy"), "--sample-pool", path("simple.pools")]
DIRMUL = ["--use-dirmul-prior", "0.0"]

ASSEMBLE = [
    (SHALLOW, [], "simple.output.assemble.flatprior.vcf"),
    (SHALLOW, DIRMUL, "simple.output.assemble.vcf"),
    (DEEP, DIRMUL, "simple.output.deep.assemble.vcf"),
    (MIXED, DIRMUL + ["--report", "SNVDP"], "simple.output.mixed_depth.assemble.vcf"),
    (MIXED, DIRMUL + ["--report", "AFP"], "simple.output.mixed_depth.assemble.frequencies.vcf"),
    (MIXED, DIRMUL + ["--report", "ACP"], "simple.output.mixed_depth.assemble.counts.vcf"),
    (MIXED, DIRMUL + ["--report", "AOP", "AOPSUM"], "simple.output.mixed_depth.assemble.occurrence.vcf"),
    (MIXED, DIRMUL + ["--sample-pool", "POOL", "--report", "AFP"],
     "simple.output.mixed_depth.assemble.pool.frequencies.vcf"),
    (SHALLOW, DIRMUL + ["--haplotype-posterior-threshold", "1.0", "--base-error-rate", "0.0",
                        "--use-base-phred-scores"], "simple.output.nullallele.assemble.vcf"),
    (DEEP, DIRMUL + POOLS, "simple.output.deep.assemble.pools.vcf"),
]

SKIPRARE = ["--use-dirmul-prior", "0.0", "AFP", "--filter-input-haplotypes", "AFP>=0.1"]
PRIOR = ["--use-dirmul-prior", "0.0", "AFP", "--report", "AFPRIOR", "AFP"]
PHRED_GL = ["--report", "GL", "--base-error-rate", "0.0", "--use-base-phred-scores"]
A_VCF, M_VCF, F_VCF = "simple.output.assemble.vcf", "simple.output.mixed_depth.assemble.vcf", "mock.input.frequencies.vcf"

CALL = [
    (A_VCF, SHALLOW, [], "simple.output.call.vcf"),
    (M_VCF, MIXED, ["--report", "SNVDP"], "simple.output.mixed_depth.call.vcf"),
    (M_VCF, MIXED, ["--report", "AFP"], "simple.output.mixed_depth.call.frequencies.vcf"),
    (M_VCF, MIXED, ["--report", "ACP"], "simple.output.mixed_depth.call.counts.vcf"),
    (M_VCF, MIXED, ["--report", "AOP", "AOPSUM"], "simple.output.mixed_depth.call.occurrence.vcf"),
    (F_VCF, MIXED, SKIPRARE + ["--report", "AFPRIOR", "AFP"], "simple.output.mixed_depth.call.frequencies.skiprare.vcf"),
    (F_VCF, MIXED, PRIOR, "simple.output.mixed_depth.call.frequencies.prior.vcf"),
    (M_VCF, MIXED, PHRED_GL, "simple.output.mixed_depth.call.likelihoods.vcf"),
    (M_VCF, MIXED, ["--report", "GP"], "simple.output.mixed_depth.call.posteriors.vcf"),
    (A_VCF, DEEP, POOLS, "simple.output.deep.call.pools.vcf"),
]

CALL_EXACT = [
    (A_VCF, SHALLOW, [], "simple.output.call-exact.vcf"),
    (M_VCF, MIXED, ["--report", "SNVDP"], "simple.output.mixed_depth.call-exact.vcf"),
    (M_VCF, MIXED, ["--report", "AFP"], "simple.output.mixed_depth.call-exact.frequencies.vcf"),
    (M_VCF, MIXED, ["--report", "ACP"], "simple.output.mixed_depth.call-exact.counts.vcf"),
    (M_VCF, MIXED, ["--report", "AOP", "AOPSUM"], "simple.output.mixed_depth.call-exact.occurrence.vcf"),
    (F_VCF, MIXED, SKIPRARE + ["--report", "AFPRIOR", "AFP"],
     "simple.output.mixed_depth.call-exact.frequencies.skiprare.vcf"),
    (F_VCF, MIXED, SKIPRARE + ["--report", "AFP", "GP"],
     "simple.output.mixed_depth.call-exact.frequencies.posteriors.skiprare.vcf"),
    (F_VCF, MIXED, PRIOR, "simple.output.mixed_depth.call-exact.frequencies.prior.vcf"),
    (M_VCF, MIXED, PHRED_GL, "simple.output.mixed_depth.call-exact.likelihoods.vcf"),
    (M_VCF, MIXED, ["--report", "GP"], "simple.output.mixed_depth.call-exact.posteriors.vcf"),
    (A_VCF, DEEP, POOLS, "simple.output.deep.call-exact.pools.vcf"),
]

MCMC = ["--mcmc-steps", "500", "--mcmc-burn", "100", "--mcmc-seed", "11"]


def assemble_command(bams, extra, cores="1"):
    return (["mchap", "assemble", "--bam"] + [path(b) for b in bams] + [
        "--ploidy", "4", "--targets", path("simple.bed.gz"), "--variants", path("simple.vcf.gz"),
        "--reference", path("simple.fasta")] + MCMC + ["--mcmc-llk-cache-threshold", "10", "--cores", cores] + extra)


def call_command(tool, vcf, bams, extra):
    cmd = ["mchap", tool, "--bam"] + [path(b) for b in bams] + ["--ploidy", "4", "--haplotypes", path(vcf)]
    if tool == "call":
        cmd += MCMC
    return cmd + ["--cores", "1"] + extra


def run_and_compare(command, expected_name, block_loci=None):
    prog = build_program(command)
    if block_loci:
        prog.block_loci = block_loci
    out = io.StringIO()
    prog.run_stdout(out)
    actual = out.getvalue().splitlines(keepends=True)
    with open(path(expected_name)) as f:
        expected = f.readlines()
    assert len(actual) == len(expected)
    for act, exp in zip(actual, expected):
        if act.startswith("##commandline"):
            assert exp.startswith("##commandline")
        elif act.startswith("##source=mchap"):
            assert exp.startswith("##source=mchap")
        elif act.startswith("##fileDate"):
            assert exp.startswith("##fileDate") and act > exp
        else:
            assert act == exp


@pytest.fixture
def oracle_engine(monkeypatch):
    from tests import oracle_engine as engine

    engine.install(monkeypatch)


# ------------------------------------------------------------------ host logic, CPU (oracle engine)
@pytest.mark.parametrize("bams,extra,expected", ASSEMBLE, ids=[e[2] for e in ASSEMBLE])
def test_assemble_golden_vcf_host_side(oracle_engine, bams, extra, expected):
    run_and_compare(assemble_command(bams, extra), expected)


@pytest.mark.parametrize("vcf,bams,extra,expected", CALL, ids=[e[3] for e in CALL])
def test_call_golden_vcf_host_side(oracle_engine, vcf, bams, extra, expected):
    run_and_compare(call_command("call", vcf, bams, extra), expected)


@pytest.mark.parametrize("vcf,bams,extra,expected", CALL_EXACT, ids=[e[3] for e in CALL_EXACT])
def test_call_exact_golden_vcf_host_side(oracle_engine, vcf, bams, extra, expected):
    run_and_compare(call_command("call-exact", vcf, bams, extra), expected)


def test_blocks_of_one_locus_give_the_same_vcf(oracle_engine):
    run_and_compare(assemble_command(MIXED, DIRMUL + ["--report", "AFP"]),
                    "simple.output.mixed_depth.assemble.frequencies.vcf", block_loci=1)


def test_region_argument(oracle_engine):
    # test_application_assemble.py:440-550: one --region gives the matching record of the golden file
    cmd = ["mchap", "assemble", "--bam"] + [path(b) for b in MIXED] + [
        "--ploidy", "4", "--region", "CHR1:5-25", "--region-id", "CHR1_05_25", "--variants", path("simple.vcf.gz"),
        "--reference", path("simple.fasta")] + MCMC + DIRMUL + ["--report", "SNVDP"]
    out = io.StringIO()
    build_program(cmd).run_stdout(out)
    records = [l for l in out.getvalue().splitlines() if not l.startswith("#")]
    with open(path("simple.output.mixed_depth.assemble.vcf")) as f:
        want = [l.rstrip("\n") for l in f if l.startswith("CHR1\t6\t")]
    assert records == want


# ------------------------------------------------------------------ readers
@pytest.mark.parametrize("stem", ["simple.sample1", "simple.sample2.deep"])
def test_sam_and_bam_readers_agree(stem):
    sam, bam = hostio.AlignmentFile(path(stem + ".sam")), hostio.AlignmentFile(path(stem + ".bam"))
    assert [rg["ID"] for rg in sam.header["RG"]] == [rg["ID"] for rg in bam.header["RG"]]
    assert len(sam.records) == len(bam.records) > 0
    for a, b in zip(sam.records, bam.records):
        assert (a.qname, a.flag, a.contig, a.start, a.mapping_quality, a.cigar, a.seq, a.qual) == (
            b.qname, b.flag, b.contig, b.start, b.mapping_quality, b.cigar, b.seq, b.qual)
        assert a.tags["RG"] == b.tags["RG"] and a.tags.get("MD") == b.tags.get("MD")
        assert a.reference_bases() == b.reference_bases()


def test_locus_from_bed_vcf_fasta():
    from mchap_b200.application.loci import read_bed4

    loci = [l.set_sequence(path("simple.fasta")).set_variants(path("simple.vcf.gz")) for l in read_bed4(path("simple.bed"))]
    assert [l.name for l in loci] == ["CHR1_05_25", "CHR1_30_50", "CHR2_10_30", "CHR3_20_40"]
    assert loci[0].positions == [6, 15, 22] and loci[0].alleles == [("A", "C"), ("A", "G"), ("A", "C", "T")]
    assert loci[2].positions == [14, 19] and loci[2].alleles[1] == ("A", "C", "G", "T")  # two records merged
    assert loci[1].variants == () and loci[0].sequence == "A" * 20
    assert loci[0].format_haplotypes(np.array([[0, 1, 2], [1, -1, 0]])) == [
        "AAAAAAAAAAGAAAAAATAA", "ACAAAAAAAA-AAAAAAAAA"]
    with pytest.raises(ValueError, match="does not match reference sequence"):
        bad = Locus("CHR1", 5, 25, "x", "C" * 20, None)
        bad.set_variants(path("simple.vcf"))


def test_known_haplotype_record_with_filter_and_frequencies():
    records = list(hostio.VariantFile(path("mock.input.frequencies.vcf")).fetch())
    locus = LocusPrior.from_variant_record(records[0], frequency_tag="AFP", allele_filter="AFP>=0.1")
    assert locus.mask_reference_allele and locus.frequencies[0] == 0
    np.testing.assert_allclose(locus.frequencies.sum(), 1.0)
    haps = locus.encode_haplotypes()
    assert haps.shape == (1 + len(locus.alts), len(locus.variants)) and (haps[0] == 0).all()
    flat = LocusPrior.from_variant_record(records[0])
    np.testing.assert_allclose(flat.frequencies, 1 / len(flat.frequencies))
    # INFO floats come through float32 like htslib's parsed representation
    afp = records[0].info["AFP"]
    assert all(float(np.float32(v)) == v for v in afp)


def test_read_extraction_pairs_and_filters():
    from mchap_b200.application.loci import read_bed4

    locus = next(read_bed4(path("simple.bed"))).set_sequence(path("simple.fasta")).set_variants(path("simple.vcf"))
    f = hostio.AlignmentFile(path("simple.sample1.bam"))
    chars, quals = extract_read_variants(locus, f, "SAMPLE1")
    assert chars.shape == quals.shape == (20, 3) and chars.dtype == np.dtype("U1") and quals.dtype == np.int16
    assert set(np.unique(chars)) <= set("ACGTN-")
    none, _ = extract_read_variants(locus, f, "SAMPLE2")
    assert none.shape == (0, 3)
    strict, _ = extract_read_variants(locus, f, "SAMPLE1", min_quality=61)
    assert strict.shape == (0, 3)


# ------------------------------------------------------------------ VCF text rules
def test_vcf_value_rules():
    v = vcfout.vcf_value
    assert v(np.array([1.0, 2.5, np.nan, 0.7854])) == "1,2.5,.,0.785"
    assert v(np.array([0.0])) == "0" and v(np.array([])) == "." and v(np.array([3, 4])) == "3,4"
    assert v(0.99951) == "1" and v(0.5) == "0.5" and v(np.nan) == "." and v(None) == "." and v("") == "."
    assert v(np.float64(12.0)) == "12" and v(["a", "b"]) == "a,b" and v([]) == "." and v(7) == "7"
    assert vcfout.info_text(["AN", "REFMASKED", "AC"], {"AN": 12, "REFMASKED": False, "AC": np.array([3, 2])}) == "AN=12;AC=3,2"
    assert vcfout.info_text(["REFMASKED"], {"REFMASKED": True}) == "REFMASKED"


def test_qual_of_prob():
    assert qual_of_prob(1.0) == 60 and qual_of_prob(0.9) == 10 and qual_of_prob(0.785) == 7 and qual_of_prob(0.0) == 0
    np.testing.assert_array_equal(qual_of_prob(np.array([0.5, 0.999999999])), [3, 60])


# ------------------------------------------------------------------ the real device path
@pytest.mark.gpu
@pytest.mark.parametrize("bams,extra,expected", ASSEMBLE, ids=[e[2] for e in ASSEMBLE])
def test_assemble_golden_vcf(bams, extra, expected):
    run_and_compare(assemble_command(bams, extra), expected)


@pytest.mark.gpu
@pytest.mark.parametrize("vcf,bams,extra,expected", CALL, ids=[e[3] for e in CALL])
def test_call_golden_vcf(vcf, bams, extra, expected):
    run_and_compare(call_command("call", vcf, bams, extra), expected)


@pytest.mark.gpu
@pytest.mark.parametrize("vcf,bams,extra,expected", CALL_EXACT, ids=[e[3] for e in CALL_EXACT])
def test_call_exact_golden_vcf(vcf, bams, extra, expected):
    run_and_compare(call_command("call-exact", vcf, bams, extra), expected)


@pytest.mark.gpu
def test_sam_inputs_give_the_same_vcf_as_bam():
    sams = [b.replace(".bam", ".sam") for b in SHALLOW]
    run_and_compare(assemble_command(sams, DIRMUL), "simple.output.assemble.vcf")
    run_and_compare(assemble_command(MIXED, DIRMUL + ["--report", "AFP"]),
                    "simple.output.mixed_depth.assemble.frequencies.vcf", block_loci=1)
