"""Target loci of the application layer: a genomic interval, its reference sequence and the SNVs
inside it (``Locus``, de novo assembly), or a known multi-nucleotide variant with its alternate
haplotypes and prior allele frequencies (``LocusPrior``, ``mchap call`` / ``call-exact``).

Same attribute and method names as the reference's mchap/io/loci.py:29-325 so that the programs
read alike; built on the in-memory readers of ``hostio`` instead of pysam.
"""
import re
from dataclasses import dataclass, replace

import numpy as np

from . import hostio

__all__ = ["SNP", "Locus", "LocusPrior", "read_bed4", "parse_allele_filter", "apply_allele_filter"]


@dataclass(frozen=True, order=True)
class SNP:
    contig: str
    start: int
    stop: int
    name: str
    alleles: tuple


def _location(contig, pos0, name):
    where = "'%s:%d'" % (contig, pos0 + 1)
    return "%s in target '%s'" % (where, name) if name else where


def _merged(x, y):
    """Two records of one position become one multi-allelic SNP (mchap/io/loci.py:375-388)."""
    same = (x.contig, x.name, x.start, x.stop, x.alleles[0]) == (y.contig, y.name, y.start, y.stop, y.alleles[0])
    if not same:
        raise ValueError('Cannot merge SNPs "{}: {}:{}" and "{}: {}:{}"'.format(
            x.name, x.contig, x.start, y.name, y.contig, y.start))
    extra = tuple(a for a in y.alleles if a not in x.alleles)
    return replace(x, alleles=x.alleles + extra)


def _encode_chars(chars, alleles):
    """Characters -> allele indices per position, -1 for anything that is not a listed allele
    (mchap/encoding/character/transcode.py:4-50)."""
    chars = np.asarray(chars)
    out = np.full(chars.shape, -1, dtype=np.int8)
    if chars.size == 0:
        return out
    for j, tup in enumerate(alleles):
        col = chars[..., j]
        for code, symbol in enumerate(tup):
            out[..., j][col == symbol] = code
    return out


@dataclass(frozen=True, order=True)
class Locus:
    contig: str
    start: int
    stop: int
    name: str
    sequence: str
    variants: tuple

    @property
    def positions(self):
        return [v.start for v in self.variants]

    @property
    def alleles(self):
        return [v.alleles for v in self.variants]

    def count_alleles(self):
        return [len(v.alleles) for v in self.variants]

    def _check_reference_alleles(self):
        for v in self.variants:
            have = self.sequence[v.start - self.start]
            if have != v.alleles[0]:
                raise ValueError(
                    "Reference allele of variant '%s' does not match reference sequence '%s' at %s"
                    % (v.alleles[0], have, _location(self.contig, v.start, self.name)))

    def set_sequence(self, fasta):
        """fasta: path or an open hostio.FastaFile."""
        f = fasta if isinstance(fasta, hostio.FastaFile) else hostio.FastaFile(fasta)
        new = replace(self, sequence=f.fetch(self.contig, self.start, self.stop).upper())
        if new.variants:
            new._check_reference_alleles()
        return new

    def set_variants(self, vcf):
        """Bi- or multi-allelic SNVs of the interval from a VCF (path or open hostio.VariantFile);
        records that are not single-base substitutions are ignored, repeated positions are merged."""
        f = vcf if isinstance(vcf, hostio.VariantFile) else hostio.VariantFile(vcf)
        found = {}
        for rec in f.fetch(self.contig, self.start, self.stop):
            alleles = (rec.ref,) + (rec.alts or ())
            if rec.stop - rec.start != 1 or any(len(a) != 1 for a in alleles):
                continue
            snp = SNP(rec.contig, rec.start, rec.stop, rec.id if rec.id else ".", alleles)
            found[snp.start] = _merged(found[snp.start], snp) if snp.start in found else snp
        new = replace(self, variants=tuple(found.values()))
        if new.sequence:
            new._check_reference_alleles()
        return new

    def format_haplotypes(self, array, gap="-"):
        """Integer haplotypes -> full-length sequences: the reference sequence with the SNV
        positions replaced by the haplotype's alleles."""
        array = np.asarray(array)
        offsets = [p - self.start for p in self.positions]
        out = []
        for hap in array.reshape(-1, array.shape[-1]):
            chars = list(self.sequence)
            for off, tup, a in zip(offsets, self.alleles, hap):
                chars[off] = tup[a] if a >= 0 else gap
            out.append("".join(chars))
        return out

    def encode_read_chars(self, chars):
        return _encode_chars(chars, self.alleles)

    @classmethod
    def from_region_string(cls, string, name=None):
        contig, interval = string.strip().split(":")
        start, stop = interval.strip().split("-")
        return cls(contig, int(start), int(stop), name, None, None)


def read_bed4(bed):
    for contig, start, stop, name in hostio.read_bed4_lines(bed):
        yield Locus(contig, start, stop, name, None, None)


# ---- allele filter of `--filter-input-haplotypes` (mchap/io/filter_alleles.py) -------------
_OPERATORS = {
    "=": np.equal, "==": np.equal, ">": np.greater, ">=": np.greater_equal,
    "<": np.less, "<=": np.less_equal, "!=": np.not_equal,
}
_FILTER_RE = re.compile(r"^(\w+)(==|!=|>=|<=|<>|=|>|<)(\d*[.,]?\d*)$")


def parse_allele_filter(string):
    """'<INFO field><operator><number>' -> (field, numpy comparison, value)."""
    m = _FILTER_RE.match(string)
    if not m:
        raise ValueError("Invalid allele filter '%s'" % string)
    field, op, value = m.groups()
    if op not in _OPERATORS:
        raise ValueError("Invalid operator in allele filter '%s'" % op)
    try:
        number = int(value)
    except ValueError:
        try:
            number = float(value)
        except ValueError:
            raise ValueError("Non-numerical value in allele filter '%s'" % value)
    return field, _OPERATORS[op], number


def apply_allele_filter(record, field, func, value):
    """Boolean keep-mask over (ref, alts...) from an INFO field of Number=R or Number=A."""
    meta = record.header.info.get(field)
    if meta is None:
        raise ValueError("Allele filter field not found in header '%s'" % field)
    if meta.number not in ("R", "A"):
        raise ValueError("Allele filter of field of invalid length '%s'" % meta.number)
    n_alts = len(record.alts) if record.alts else 0
    keep = np.ones(1 + n_alts, dtype=bool)
    obs = record.info.get(field)
    if obs is None:
        return keep
    if meta.number == "R":
        assert len(obs) == 1 + n_alts
        return func(obs, value)
    assert len(obs) == n_alts
    keep[1:] = func(obs, value)
    return keep


@dataclass(frozen=True, order=True)
class LocusPrior(Locus):
    alts: tuple = ()
    frequencies: np.ndarray = None
    mask_reference_allele: bool = False

    def encode_haplotypes(self):
        """Known haplotypes int[n_alleles, n_snvs] (row 0 = reference)."""
        strings = (self.sequence,) + tuple(self.alts)
        offsets = np.array(self.positions, dtype=int) - self.start
        if len(offsets) == 0:
            return np.zeros((len(strings), 0), dtype=int)
        chars = np.array([list(s) for s in strings])[:, offsets]
        return _encode_chars(chars, self.alleles)

    @classmethod
    def from_variant_record(cls, record, use_snvpos=False, frequency_tag=None, allele_filter=None,
                            masked_reference_flag="REFMASKED"):
        """A known MNP record -> locus with SNV positions found by comparing the sequences
        (mchap/io/loci.py:205-324): optional allele filter (a filtered-out reference allele is
        masked rather than dropped), prior frequencies from an INFO tag or flat, renormalised."""
        alts = tuple(record.alts) if record.alts else ()
        assert all(len(a) == len(record.ref) for a in alts)
        masked = masked_reference_flag in record.info
        keep = None
        if allele_filter is not None:
            keep = np.array(apply_allele_filter(record, *parse_allele_filter(allele_filter)), dtype=bool)
            if not keep[0]:
                masked = True
                keep[0] = True
        n_alleles = len(alts) + 1
        if frequency_tag:
            freqs = record.info.get(frequency_tag, ())
            if len(freqs) != n_alleles:
                raise ValueError("Field '%s' does not match number of alleles 'n_alleles'." % frequency_tag)
            freqs = np.array(freqs)
        else:
            freqs = np.ones(n_alleles) / n_alleles
        if masked:
            freqs[0] = 0
        sequences = (record.ref,) + alts
        if keep is not None:
            sequences = tuple(s for s, k in zip(sequences, keep) if k)
            freqs = freqs[keep]
        total = freqs.sum()
        if total > 0:
            freqs /= total
        else:
            freqs[:] = np.nan
        chars = np.array([list(s) for s in sequences])
        if use_snvpos:
            snvpos = record.info["SNVPOS"]
            offsets = np.array(() if snvpos == (None,) else snvpos, int) - 1
        else:
            offsets = np.where((chars != chars[0:1]).any(axis=0))[0]
        snps = []
        for off in offsets:
            column = chars[:, off]
            _, first = np.unique(column, return_index=True)
            first.sort()
            pos = int(off) + record.start
            snps.append(SNP(record.chrom, pos, pos + 1, ".", tuple(str(c) for c in column[first])))
        return cls(
            contig=record.chrom, start=record.start, stop=record.stop, name=record.id if record.id else ".",
            sequence=record.ref, variants=tuple(snps), alts=sequences[1:], frequencies=freqs,
            mask_reference_allele=masked)
