#!/usr/bin/env python3
"""Copy the reference's bundled application test DATA (BASELINE configs[0]) into tests/golden/cfg1/.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_cfg1.py

These are data files of the reference's test-suite (mchap/tests/test_io/data): the alignments,
targets, variants and reference sequence the three CLIs are run on, and the VCFs the reference's own
tests (test_application_assemble.py:254-437, test_application_call.py:16-200,
test_application_call_exact.py:16-216) expect on stdout.  No reference source code is copied.  The GPU
box has no /root/reference, so the files travel as fixtures.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.environ.get("MCHAP_REFERENCE", "/root/reference"), "mchap", "tests", "test_io", "data")
DST = os.path.join(HERE, "cfg1")

INPUTS = [
    "simple.bed", "simple.bed.gz", "simple.vcf", "simple.vcf.gz", "simple.fasta", "simple.fasta.fai",
    "simple.pools", "simple.pools-ploidy", "mock.input.frequencies.vcf",
] + ["simple.sample%d%s.%s" % (i, deep, ext) for i in (1, 2, 3) for deep in ("", ".deep") for ext in ("sam", "bam")]

EXPECTED = [
    "simple.output.assemble.flatprior.vcf", "simple.output.assemble.vcf", "simple.output.deep.assemble.vcf",
    "simple.output.mixed_depth.assemble.vcf", "simple.output.mixed_depth.assemble.frequencies.vcf",
    "simple.output.mixed_depth.assemble.counts.vcf", "simple.output.mixed_depth.assemble.occurrence.vcf",
    "simple.output.mixed_depth.assemble.pool.frequencies.vcf", "simple.output.nullallele.assemble.vcf",
    "simple.output.deep.assemble.pools.vcf",
    "simple.output.call.vcf", "simple.output.mixed_depth.call.vcf", "simple.output.mixed_depth.call.frequencies.vcf",
    "simple.output.mixed_depth.call.counts.vcf", "simple.output.mixed_depth.call.occurrence.vcf",
    "simple.output.mixed_depth.call.frequencies.skiprare.vcf", "simple.output.mixed_depth.call.frequencies.prior.vcf",
    "simple.output.mixed_depth.call.likelihoods.vcf", "simple.output.mixed_depth.call.posteriors.vcf",
    "simple.output.deep.call.pools.vcf",
    "simple.output.call-exact.vcf", "simple.output.mixed_depth.call-exact.vcf",
    "simple.output.mixed_depth.call-exact.frequencies.vcf", "simple.output.mixed_depth.call-exact.counts.vcf",
    "simple.output.mixed_depth.call-exact.occurrence.vcf",
    "simple.output.mixed_depth.call-exact.frequencies.skiprare.vcf",
    "simple.output.mixed_depth.call-exact.frequencies.posteriors.skiprare.vcf",
    "simple.output.mixed_depth.call-exact.frequencies.prior.vcf",
    "simple.output.mixed_depth.call-exact.likelihoods.vcf", "simple.output.mixed_depth.call-exact.posteriors.vcf",
    "simple.output.deep.call-exact.pools.vcf",
]

if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for name in INPUTS + EXPECTED:
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
        os.chmod(os.path.join(DST, name), 0o644)
    print("copied %d files to %s" % (len(INPUTS + EXPECTED), DST))
