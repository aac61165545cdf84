"""Read extraction for one (locus, sample): base calls at the locus' SNV positions per read pair.

Host-side step that precedes the device pipeline (mchap/io/bam.py:53-214 ``extract_read_variants``
and the statistics of mchap/application/baseclass.py:176-203).  The product is what the device
encoder ``mchb_encode_reads_batch`` takes: int8 allele calls [n_reads, n_snvs] (-1 = gap / not an
allele) and the phred scores beside them.
"""
import numpy as np

from . import hostio

__all__ = ["extract_read_variants", "prob_of_qual", "qual_of_prob", "read_depth"]


def prob_of_qual(qual):
    """Phred score -> probability that the call is correct (mchap/io/util.py:40-55)."""
    return 1 - (10 ** (np.asarray(qual) / -10))


def qual_of_prob(prob, precision=6):
    """Probability of a correct call -> phred score, capped at 10 * precision and floored to
    ``precision`` decimals before the conversion (mchap/io/util.py:58-88)."""
    cap = 1 - 0.1 ** precision
    scale = 10 ** precision
    if np.shape(prob) == ():
        p = cap if prob > cap else prob
    else:
        p = np.array([cap if x > cap else x for x in prob])
    p = np.floor(p * scale) / scale
    return np.round(-10 * np.log10(1 - p)).astype(int)


def extract_read_variants(locus, alignment_file, sample, id="SM", min_quality=20, skip_duplicates=True,
                          skip_qcfail=True, skip_supplementary=True):
    """(chars U1[n_reads, n_snvs], quals int16[n_reads, n_snvs]) of the read pairs of ``sample``
    overlapping the locus, in order of first appearance in the file.  Mates share a row: agreeing
    calls add their qualities, disagreeing calls become 'N'."""
    assert id in ("ID", "SM")
    n_pos = len(locus.positions)
    column = {pos: j for j, pos in enumerate(locus.positions)}
    group_sample = {rg["ID"]: rg[id] for rg in alignment_file.header.get("RG", [])}
    rows = {}
    for read in alignment_file.fetch(locus.contig, locus.start, locus.stop):
        if read.is_unmapped or read.mapping_quality < min_quality:
            continue
        if (read.is_duplicate and skip_duplicates) or (read.is_qcfail and skip_qcfail) or (
                read.is_supplementary and skip_supplementary):
            continue
        if group_sample[read.tags["RG"]] != sample:
            continue
        pair = rows.get(read.qname)
        if pair is None:
            pair = rows[read.qname] = (["-"] * n_pos, [0] * n_pos)
        chars, quals = pair
        ref_bases = None
        for read_pos, ref_pos in read.aligned_pairs():
            j = column.get(ref_pos)
            if j is None:
                continue
            if ref_bases is None:
                ref_bases = read.reference_bases() or {}
            ref_char = ref_bases.get(ref_pos)
            if ref_char is not None and locus.alleles[j][0].upper() != ref_char.upper():
                where = "'%s:%d'" % (locus.contig, ref_pos + 1)
                if locus.name:
                    where += " in target '%s'" % locus.name
                raise ValueError(
                    "Reference allele of variant '%s' does not match alignment reference allele '%s' at position %s in '%s'"
                    % (locus.alleles[j][0], ref_char, where, alignment_file.filename))
            char = read.seq[read_pos]
            qual = ord(read.qual[read_pos]) - 33
            if chars[j] == "-":
                chars[j], quals[j] = char, qual
            elif chars[j] == char:
                quals[j] += qual
            else:
                chars[j] = "N"
    if not rows:
        return np.empty((0, n_pos), dtype="U1"), np.empty((0, n_pos), dtype=np.int16)
    chars = np.array([c for c, _ in rows.values()], dtype="U1").reshape(len(rows), n_pos)
    quals = np.array([q for _, q in rows.values()], dtype=np.int16).reshape(len(rows), n_pos)
    return chars, quals


def read_depth(chars):
    """Reads with a non-gap character per SNV position (mchap/encoding/character/sequence.py:24-43)."""
    return np.sum(chars != "-", axis=0)


def open_alignment(path, cache):
    """One parsed AlignmentFile per path for the lifetime of a program run."""
    f = cache.get(path)
    if f is None:
        f = cache[path] = hostio.AlignmentFile(path)
    return f
