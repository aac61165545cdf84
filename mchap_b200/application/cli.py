"""Command line of the three hot-path programs with the reference's option names:

    python -m mchap_b200 assemble   --bam ... --targets BED --variants VCF --reference FASTA ...
    python -m mchap_b200 call       --bam ... --haplotypes VCF ...
    python -m mchap_b200 call-exact --bam ... --haplotypes VCF ...

Options, defaults and their interpretation follow mchap/application/arguments.py (file-or-value
arguments for ``--ploidy`` / ``--use-dirmul-prior`` / ``--mcmc-temperatures``, sample pools,
``--report`` fields).  ``--cores`` is accepted and ignored: the loci of a run are batched onto the GPU
instead of being split over worker processes.  ``mchap call-pedigree``, ``find-snvs`` and ``atomize``
are outside this package.
"""
import argparse
import os
import sys

from . import hostio, vcfout
from .programs import PFEIFFER_ERROR, assemble_program, call_exact_program, call_program

__all__ = ["main", "build_program"]


def _add_common(p, dirmul_nargs):
    p.add_argument("--bam", type=str, nargs="+", default=[])
    p.add_argument("--ploidy", type=str, nargs=1, default=["2"])
    p.add_argument("--sample-pool", type=str, nargs=1, default=[None])
    p.add_argument("--use-dirmul-prior", type=str, nargs=dirmul_nargs, default=[None] * dirmul_nargs)
    p.add_argument("--reference", type=str, nargs=1, default=[None])
    p.add_argument("--base-error-rate", type=float, nargs=1, default=[PFEIFFER_ERROR])
    p.add_argument("--use-base-phred-scores", dest="ignore_base_phred_scores", action="store_false", default=True)
    p.add_argument("--mapping-quality", type=int, nargs=1, default=[20])
    p.add_argument("--keep-duplicate-reads", dest="skip_duplicates", action="store_false", default=True)
    p.add_argument("--keep-qcfail-reads", dest="skip_qcfail", action="store_false", default=True)
    p.add_argument("--keep-supplementary-reads", dest="skip_supplementary", action="store_false", default=True)
    p.add_argument("--read-group-field", type=str, nargs=1, default=["SM"])
    p.add_argument("--report", type=str, nargs="*", default=[])
    p.add_argument("--cores", type=int, nargs=1, default=[1])
    p.add_argument("--device", type=int, nargs=1, default=[0], help="CUDA device ordinal")


def _add_mcmc(p):
    p.add_argument("--mcmc-chains", type=int, nargs=1, default=[2])
    p.add_argument("--mcmc-steps", type=int, nargs=1, default=[2000])
    p.add_argument("--mcmc-burn", type=int, nargs=1, default=[1000])
    p.add_argument("--mcmc-seed", type=int, nargs=1, default=[42])
    p.add_argument("--mcmc-chain-incongruence-threshold", type=float, nargs=1, default=[0.60])


def _parser(tool):
    p = argparse.ArgumentParser("mchap_b200 " + tool)
    if tool == "assemble":
        _add_common(p, 1)
        _add_mcmc(p)
        p.add_argument("--region", type=str, nargs=1, default=[None])
        p.add_argument("--region-id", type=str, nargs=1, default=[None])
        p.add_argument("--targets", type=str, nargs=1, default=[None])
        p.add_argument("--variants", type=str, nargs=1, default=[None])
        p.add_argument("--mcmc-fix-homozygous", type=float, nargs=1, default=[0.999])
        p.add_argument("--mcmc-llk-cache-threshold", type=int, nargs=1, default=[100])
        p.add_argument("--mcmc-recombination-step-probability", type=float, nargs=1, default=[0.5])
        p.add_argument("--mcmc-dosage-step-probability", type=float, nargs=1, default=[1.0])
        p.add_argument("--mcmc-partial-dosage-step-probability", type=float, nargs=1, default=[0.5])
        p.add_argument("--mcmc-temperatures", type=str, nargs="*", default=["1.0"])
        p.add_argument("--haplotype-posterior-threshold", type=float, nargs=1, default=[0.20])
    else:
        _add_common(p, 2)
        p.add_argument("--haplotypes", type=str, nargs=1, default=[None])
        p.add_argument("--filter-input-haplotypes", type=str, nargs=1, default=[None])
        if tool == "call":
            _add_mcmc(p)
    return p


def _is_number(text, integer):
    return text.isdigit() if integer else text.replace(".", "", 1).isdigit()


def sample_values(argument, samples, kind):
    """A constant for every sample, or a two-column file sample<TAB>value (arguments.py:929-960)."""
    if _is_number(argument, kind is int):
        return {s: kind(argument) for s in samples}
    table = {}
    with open(argument) as f:
        for line in f:
            sample, value = line.strip().split("\t")
            table[sample] = kind(value)
    for s in samples:
        if s not in table:
            raise ValueError("Sample '{}' not found in file '{}'".format(s, argument))
    return table


def sample_temperatures(argument, samples):
    """Temperature ladder per sample: values on the command line (for everyone) or a file of
    sample<TAB>t1<TAB>t2...; ladders are sorted and end in 1.0 (arguments.py:1043-1083)."""
    def ladder(values):
        temps = sorted(float(v) for v in values)
        assert temps[0] > 0.0
        assert temps[-1] <= 1.0
        if temps[-1] != 1.0:
            temps.append(1.0)
        return temps

    if len(argument) > 1 or _is_number(argument[0], False):
        temps = ladder(argument)
        return {s: temps for s in samples}
    table = {s: [1.0] for s in samples}
    with open(argument[0]) as f:
        for line in f:
            values = line.strip().split("\t")
            table[values[0]] = ladder(values[1:])
    assert len(samples) == len(table)
    return table


def _alignment_samples(paths, id):
    found = {}
    for path in paths:
        header = hostio.AlignmentFile(path).header
        for sample in dict.fromkeys(rg[id] for rg in header.get("RG", [])):
            if sample in found:
                raise IOError('Duplicate sample with id = "{}" in file "{}"'.format(sample, path))
            found[sample] = path
    return found


def sample_alignments(bam_argument, pool_argument, id):
    """(samples, {sample: [(read-group sample, path)]}) from ``--bam`` (alignment files, or a text
    file listing paths or sample<TAB>path pairs) and ``--sample-pool`` (arguments.py:838-926)."""
    listing = None
    if len(bam_argument) == 1:
        try:
            hostio.AlignmentFile(bam_argument[0])
        except (ValueError, UnicodeDecodeError, IndexError):
            with open(bam_argument[0]) as f:
                listing = [line.strip().split("\t") for line in f if line.strip()]
    if listing is None:
        sample_bams = _alignment_samples(bam_argument, id)
        samples = list(sample_bams)
    else:
        width = len(listing[0])
        if any(len(row) != width for row in listing):
            raise ValueError("Inconsistent number of fields")
        if width == 1:
            sample_bams = _alignment_samples([row[0] for row in listing], id)
            samples = list(sample_bams)
        elif width == 2:
            samples = [row[0] for row in listing]
            sample_bams = dict(listing)
        else:
            raise ValueError("Too many fields")
    if pool_argument is None:
        return samples, {k: [(k, v)] for k, v in sample_bams.items()}
    if not os.path.isfile(pool_argument):
        return [pool_argument], {pool_argument: [(k, v) for k, v in sample_bams.items()]}
    pools, members, assigned = [], {}, set()
    with open(pool_argument) as f:
        for line in f:
            if not line.strip():
                continue
            sample, pool = line.strip().split("\t")
            assigned.add(sample)
            if pool not in members:
                pools.append(pool)
                members[pool] = []
            members[pool].append((sample, sample_bams[sample]))
    missing = set(samples) - assigned
    if missing:
        raise ValueError(f"The following samples have not been assigned to a pool: {missing}")
    unknown = assigned - set(samples)
    if unknown:
        raise ValueError(f"The following names in the sample-pool file do not match a known sample : {unknown}")
    return pools, members


def build_program(command):
    """['mchap', 'assemble' | 'call' | 'call-exact', options...] -> program instance."""
    tool = command[1]
    if tool not in ("assemble", "call", "call-exact"):
        raise SystemExit("unknown tool '%s' (expected assemble, call or call-exact)" % tool)
    a = _parser(tool).parse_args(command[2:])
    if a.ignore_base_phred_scores and a.base_error_rate[0] == 0.0:
        raise ValueError("Cannot ignore base phred scores if --base-error-rate is 0")
    samples, sample_bams = sample_alignments(a.bam, a.sample_pool[0], a.read_group_field[0])
    info_ids, format_ids = vcfout.report_fields(a.report)
    dirmul = a.use_dirmul_prior[0]
    kwargs = dict(
        samples=samples, sample_bams=sample_bams,
        sample_ploidy=sample_values(a.ploidy[0], samples, int),
        sample_inbreeding=None if dirmul is None else sample_values(dirmul, samples, float),
        ref=a.reference[0], read_group_field=a.read_group_field[0], base_error_rate=a.base_error_rate[0],
        ignore_base_phred_scores=a.ignore_base_phred_scores, mapping_quality=a.mapping_quality[0],
        skip_duplicates=a.skip_duplicates, skip_qcfail=a.skip_qcfail, skip_supplementary=a.skip_supplementary,
        info_fields=info_ids, format_fields=format_ids, n_cores=a.cores[0], cli_command=command,
        device_ordinal=a.device[0])
    if tool == "assemble":
        return assemble_program(
            vcf=a.variants[0], bed=a.targets[0], region=a.region[0], region_id=a.region_id[0],
            mcmc_chains=a.mcmc_chains[0], mcmc_steps=a.mcmc_steps[0], mcmc_burn=a.mcmc_burn[0],
            random_seed=a.mcmc_seed[0], mcmc_incongruence_threshold=a.mcmc_chain_incongruence_threshold[0],
            mcmc_fix_homozygous=a.mcmc_fix_homozygous[0], mcmc_llk_cache_threshold=a.mcmc_llk_cache_threshold[0],
            mcmc_recombination_step_probability=a.mcmc_recombination_step_probability[0],
            mcmc_partial_dosage_step_probability=a.mcmc_partial_dosage_step_probability[0],
            mcmc_dosage_step_probability=a.mcmc_dosage_step_probability[0],
            sample_mcmc_temperatures=sample_temperatures(a.mcmc_temperatures, samples),
            haplotype_posterior_threshold=a.haplotype_posterior_threshold[0], **kwargs)
    kwargs.update(vcf=a.haplotypes[0], prior_frequencies_tag=a.use_dirmul_prior[1],
                  filter_input_haplotypes=a.filter_input_haplotypes[0])
    if tool == "call-exact":
        return call_exact_program(random_seed=None, **kwargs)
    return call_program(
        mcmc_chains=a.mcmc_chains[0], mcmc_steps=a.mcmc_steps[0], mcmc_burn=a.mcmc_burn[0],
        random_seed=a.mcmc_seed[0], mcmc_incongruence_threshold=a.mcmc_chain_incongruence_threshold[0], **kwargs)


def main(argv=None):
    command = list(sys.argv if argv is None else argv)
    if len(command) < 3:
        print(__doc__)
        return 1
    prog = build_program(command)
    prog.run_stdout()
    return 0
