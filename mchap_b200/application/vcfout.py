"""VCF text output of the application layer: header lines, INFO / FORMAT field catalogue and the
value formatting rules that make a record line.

The output has to be byte-identical to the reference's (golden VCFs under
mchap/tests/test_io/data/simple.output*.vcf), so the catalogue mirrors mchap/io/vcf/infofields.py and
formatfields.py (ids, Number, Type, description text, default order) and ``vcf_value`` follows
mchap/io/vcf/util.py:4-42 (arrays rounded and joined through numpy's own float printing, trailing
``.0`` trimmed, NaN and empty values written as '.').
"""
from collections import namedtuple
from datetime import date

import numpy as np

__all__ = ["INFO", "FORMAT", "FILTERS", "report_fields", "header_lines", "vcf_value", "record_line"]

Field = namedtuple("Field", "kind id number type descr")


def _line(f):
    return '##%s=<ID=%s,Number=%s,Type=%s,Description="%s">' % (f.kind, f.id, f.number, f.type, f.descr)


def _catalogue(kind, rows):
    return {r[0]: Field(kind, *r) for r in rows}


INFO = _catalogue("INFO", [
    ("AN", 1, "Integer", "Total number of alleles in called genotypes"),
    ("UAN", 1, "Integer", "Total number of unique alleles in called genotypes"),
    ("AC", "A", "Integer", "Allele count in genotypes, for each ALT allele, in the same order as listed"),
    ("REFMASKED", 0, "Flag", "Reference allele is masked"),
    ("NS", 1, "Integer", "Number of samples with data"),
    ("MCI", 1, "Integer", "Number of samples with incongruent Markov chain replicates"),
    ("DP", 1, "Integer", "Combined depth across samples"),
    ("RCOUNT", 1, "Integer", "Total number of observed reads across all samples"),
    ("END", 1, "Integer", "End position on CHROM"),
    ("NVAR", 1, "Integer", "Number of input variants within assembly locus"),
    ("SNVPOS", ".", "Integer", "Relative (1-based) positions of SNVs within haplotypes"),
    ("AFPRIOR", "R", "Float", "Prior allele frequencies"),
    ("ACP", "R", "Float", "Posterior allele counts"),
    ("AFP", "R", "Float", "Posterior mean allele frequencies"),
    ("AOP", "R", "Float", "Posterior probability of allele occurring across all samples"),
    ("AOPSUM", "R", "Float", "Posterior estimate of the number of samples containing an allele"),
    ("SNVDP", ".", "Integer", "Read depth at each SNV position"),
])
INFO_DEFAULT = ["AN", "UAN", "AC", "REFMASKED", "NS", "MCI", "DP", "RCOUNT", "END", "NVAR", "SNVPOS"]
INFO_OPTIONAL = ["AFPRIOR", "ACP", "AFP", "AOP", "AOPSUM", "SNVDP"]

FORMAT = _catalogue("FORMAT", [
    ("GT", 1, "String", "Genotype"),
    ("GQ", 1, "Integer", "Genotype quality"),
    ("SQ", 1, "Integer", "Genotype support quality"),
    ("DP", 1, "Integer", "Read depth"),
    ("RCOUNT", 1, "Integer", "Total count of read pairs within haplotype interval"),
    ("RCALLS", 1, "Integer", "Total count of read base calls matching a known variant"),
    ("MEC", 1, "Integer", "Minimum error correction"),
    ("MECP", 1, "Float", "Minimum error correction proportion"),
    ("GPM", 1, "Float", "Genotype posterior mode probability"),
    ("SPM", 1, "Float", "Genotype support posterior mode probability"),
    ("MCI", 1, "Integer",
     "Replicate Markov-chain incongruence, 0 = none, 1 = incongruence, 2 = putative CNV"),
    ("ACP", "R", "Float", "Posterior allele counts"),
    ("AFP", "R", "Float", "Posterior mean allele frequencies"),
    ("AOP", "R", "Float", "Posterior probability of allele occurring"),
    ("GP", "G", "Float", "Genotype posterior probabilities"),
    ("GL", "G", "Float", "Genotype likelihoods"),
    ("SNVDP", ".", "Integer", "Read depth at each SNV position"),
])
FORMAT_DEFAULT = ["GT", "GQ", "SQ", "DP", "RCOUNT", "RCALLS", "MEC", "MECP", "GPM", "SPM", "MCI"]
FORMAT_OPTIONAL = ["ACP", "AFP", "AOP", "GP", "GL", "SNVDP"]

FILTERS = [
    ("PASS", "All filters passed"),
    ("NOA", "No observed alleles at locus"),
    ("AF0", "All alleles have prior allele frequency of zero"),
]


def report_fields(report):
    """(info ids, format ids) for a ``--report`` list: the defaults plus every requested optional
    field, as plain ids or qualified 'INFO/ID' / 'FORMAT/ID' (arguments.py:1086-1103)."""
    asked = set(report or ())
    info = INFO_DEFAULT + [i for i in INFO_OPTIONAL if i in asked or "INFO/" + i in asked]
    fmt = FORMAT_DEFAULT + [i for i in FORMAT_OPTIONAL if i in asked or "FORMAT/" + i in asked]
    return info, fmt


def header_lines(samples, contigs, info_ids, format_ids, command, random_seed, version):
    """Header of the output VCF in the reference's line order (application/baseclass.py:86-111)."""
    today = date.today()
    if not isinstance(command, str):
        command = '"%s"' % " ".join(command)
    lines = [
        "##fileformat=VCFv4.3",
        "##fileDate=%04d%02d%02d" % (today.year, today.month, today.day),
        "##source=mchap v%s" % version,
        "##phasing=None",
        "##commandline=%s" % command,
        "##randomseed=%s" % random_seed,
    ]
    lines += ["##contig=<ID=%s,length=%s>" % (name, "." if length is None else length) for name, length in contigs]
    lines += ['##FILTER=<ID=%s,Description="%s">' % f for f in FILTERS]
    lines += [_line(INFO[i]) for i in info_ids]
    lines += [_line(FORMAT[i]) for i in format_ids]
    lines.append("#" + "\t".join(
        ["CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + list(samples)))
    return lines


def vcf_value(obj, precision=3):
    """One value as VCF text."""
    if isinstance(obj, np.ndarray) and obj.ndim > 0:
        if len(obj) == 0:
            return "."
        if np.issubdtype(obj.dtype, np.floating):
            text = ",".join(obj.round(precision).astype("U16")).replace("nan", ".").replace(".0,", ",")
            return text[:-2] if text.endswith(".0") else text
        if np.issubdtype(obj.dtype, np.integer):
            return ",".join(obj.astype("U16"))
    if isinstance(obj, str):
        return obj if obj else "."
    if obj is None:
        return "."
    if hasattr(obj, "__iter__"):
        return ",".join(vcf_value(o) for o in obj) if len(obj) else "."
    if isinstance(obj, float):  # includes np.float64
        if np.isnan(obj):
            return "."
        r = np.round(obj, precision)
        return str(int(r)) if int(r) == r else str(r)
    return str(obj)


def info_text(ids, values, precision=3):
    parts = []
    for i in ids:
        v = values[i]
        if isinstance(v, bool):
            if v:
                parts.append(i)   # a flag is written only when set
        else:
            parts.append("%s=%s" % (i, vcf_value(v, precision)))
    return ";".join(parts)


def samples_text(ids, per_sample, precision=3):
    """FORMAT column + one column per sample; per_sample[id] = list of values in sample order."""
    columns = []
    for i in ids:
        vals = per_sample[i]
        if i == "GT":
            vals = ["/".join(str(a) if a >= 0 else "." for a in g) for g in vals]
        columns.append(vals)
    n = len(columns[0])
    assert all(len(c) == n for c in columns)
    cells = [":".join(vcf_value(c[s], precision) for c in columns) for s in range(n)]
    return ":".join(ids) + "\t" + "\t".join(cells)


def record_line(chrom, pos, id, ref, alt, qual, filter, info, samples, precision=3):
    fields = [chrom, pos, id, ref, alt, qual, filter, info, samples]
    return "\t".join(vcf_value(f, precision) for f in fields)
