"""Host-side file readers for the application layer: SAM / BAM alignments, FASTA, BED4, VCF.

SURVEY.md section 8(f) N4.  The reference reads these formats through pysam
(mchap/io/bam.py:53-214, mchap/io/loci.py:87-137 and 346-372, application/call_baseclass.py:13-21).
pysam is a C extension that is not part of this package's dependencies, so the few things the
pipeline needs from it are read directly with the standard library: text SAM and BGZF-compressed
BAM (``gzip`` reads multi-member files), FASTA by line, BED4, and VCF / bgzipped VCF with the typed
INFO access of ``pysam.VariantRecord.info``.  Files are read whole and indexed in memory (sorted
start positions per contig): the access pattern is one region query per target locus.

Nothing here touches the GPU.
"""
import bisect
import gzip
import struct
from dataclasses import dataclass, field

import numpy as np

__all__ = ["Alignment", "AlignmentFile", "FastaFile", "VariantFile", "VariantRecord", "read_bed4_lines"]

_GZIP_MAGIC = b"\x1f\x8b\x08"


def _open_maybe_gzip(path):
    with open(path, "rb") as f:
        magic = f.read(3)
    if magic == _GZIP_MAGIC:
        return gzip.open(path, "rb")
    return open(path, "rb")


# ------------------------------------------------------------------------------ alignments
FLAG_UNMAPPED = 0x4
FLAG_QCFAIL = 0x200
FLAG_DUPLICATE = 0x400
FLAG_SUPPLEMENTARY = 0x800

_CIGAR_OPS = "MIDNSHP=X"
_CONSUMES_QUERY = {"M", "I", "S", "=", "X"}
_CONSUMES_REF = {"M", "D", "N", "=", "X"}
_ALIGNED = {"M", "=", "X"}


@dataclass
class Alignment(object):
    """The fields of one alignment record that mchap/io/bam.py:126-195 uses."""

    qname: str
    flag: int
    contig: str
    start: int            # 0-based leftmost reference position
    mapping_quality: int
    cigar: list           # [(op char, length)]
    seq: str
    qual: str             # phred+33 characters, like pysam's AlignedSegment.qual
    tags: dict = field(default_factory=dict)

    @property
    def is_unmapped(self):
        return bool(self.flag & FLAG_UNMAPPED)

    @property
    def is_duplicate(self):
        return bool(self.flag & FLAG_DUPLICATE)

    @property
    def is_qcfail(self):
        return bool(self.flag & FLAG_QCFAIL)

    @property
    def is_supplementary(self):
        return bool(self.flag & FLAG_SUPPLEMENTARY)

    @property
    def reference_end(self):
        n = sum(length for op, length in self.cigar if op in _CONSUMES_REF)
        return self.start + n

    def aligned_pairs(self):
        """(read position, reference position) of every aligned (M/=/X) base: what
        ``get_aligned_pairs(matches_only=True)`` yields."""
        qpos, rpos = 0, self.start
        for op, length in self.cigar:
            if op in _ALIGNED:
                for k in range(length):
                    yield qpos + k, rpos + k
                qpos += length
                rpos += length
            else:
                if op in _CONSUMES_QUERY:
                    qpos += length
                if op in _CONSUMES_REF:
                    rpos += length

    def reference_bases(self):
        """{reference position: reference base} over the aligned bases, rebuilt from the MD tag
        (what ``with_seq=True`` adds); None without an MD tag."""
        md = self.tags.get("MD")
        if md is None:
            return None
        # expand MD into one entry per reference base of the M/D span: None = same as the read
        ref_ops = []
        i, n = 0, len(md)
        while i < n:
            c = md[i]
            if c.isdigit():
                j = i
                while j < n and md[j].isdigit():
                    j += 1
                ref_ops.extend([None] * int(md[i:j]))
                i = j
            elif c == "^":
                j = i + 1
                while j < n and md[j].isalpha():
                    j += 1
                ref_ops.extend(("del", b) for b in md[i + 1:j])
                i = j
            else:
                ref_ops.append(c)
                i += 1
        out = {}
        k = 0  # cursor in ref_ops (covers M/=/X and D operations)
        qpos, rpos = 0, self.start
        for op, length in self.cigar:
            if op in _ALIGNED:
                for t in range(length):
                    e = ref_ops[k] if k < len(ref_ops) else None
                    k += 1
                    out[rpos + t] = self.seq[qpos + t] if e is None else e
                qpos += length
                rpos += length
            elif op == "D":
                k += length
                rpos += length
            else:
                if op in _CONSUMES_QUERY:
                    qpos += length
                if op in _CONSUMES_REF:
                    rpos += length
        return out


def _parse_cigar_string(text):
    if text == "*":
        return []
    out, num = [], 0
    for c in text:
        if c.isdigit():
            num = num * 10 + ord(c) - 48
        else:
            out.append((c, num))
            num = 0
    return out


def _parse_sam_header_line(line, header):
    parts = line.rstrip("\n").split("\t")
    kind = parts[0][1:]
    if kind == "CO":
        return
    entry = {}
    for p in parts[1:]:
        if len(p) >= 3 and p[2] == ":":
            entry[p[:2]] = p[3:]
    header.setdefault(kind, []).append(entry)


def _read_sam(handle):
    header, records = {}, []
    for raw in handle:
        line = raw.decode() if isinstance(raw, bytes) else raw
        if not line.strip():
            continue
        if line.startswith("@"):
            _parse_sam_header_line(line, header)
            continue
        f = line.rstrip("\n").split("\t")
        tags = {}
        for t in f[11:]:
            key, typ, val = t[:2], t[3], t[5:]
            if typ == "i":
                val = int(val)
            elif typ == "f":
                val = float(val)
            tags[key] = val
        records.append(Alignment(
            qname=f[0], flag=int(f[1]), contig=f[2], start=int(f[3]) - 1, mapping_quality=int(f[4]),
            cigar=_parse_cigar_string(f[5]), seq=f[9], qual=f[10], tags=tags))
    return header, records


_BAM_SEQ = "=ACMGRSVTWYHKDBN"
_TAG_FMT = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}


def _parse_bam_tags(buf, pos, end):
    tags = {}
    while pos < end:
        key = buf[pos:pos + 2].decode()
        typ = chr(buf[pos + 2])
        pos += 3
        if typ == "A":
            tags[key] = chr(buf[pos])
            pos += 1
        elif typ in _TAG_FMT:
            fmt = _TAG_FMT[typ]
            tags[key] = struct.unpack_from(fmt, buf, pos)[0]
            pos += struct.calcsize(fmt)
        elif typ in "ZH":
            z = buf.index(b"\x00", pos)
            tags[key] = buf[pos:z].decode()
            pos = z + 1
        elif typ == "B":
            sub = chr(buf[pos])
            count = struct.unpack_from("<i", buf, pos + 1)[0]
            fmt = _TAG_FMT[sub]
            size = struct.calcsize(fmt)
            tags[key] = [struct.unpack_from(fmt, buf, pos + 5 + k * size)[0] for k in range(count)]
            pos += 5 + count * size
        else:
            raise ValueError("unknown BAM tag type %r" % typ)
    return tags


def _read_bam(handle):
    data = handle.read()
    if data[:4] != b"BAM\x01":
        raise ValueError("not a BAM file")
    l_text = struct.unpack_from("<i", data, 4)[0]
    text = data[8:8 + l_text].split(b"\x00")[0].decode()
    header = {}
    for line in text.split("\n"):
        if line.startswith("@"):
            _parse_sam_header_line(line, header)
    pos = 8 + l_text
    n_ref = struct.unpack_from("<i", data, pos)[0]
    pos += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", data, pos)[0]
        name = data[pos + 4:pos + 4 + l_name - 1].decode()
        pos += 4 + l_name + 4
        refs.append(name)
    records = []
    n = len(data)
    while pos + 4 <= n:
        block = struct.unpack_from("<i", data, pos)[0]
        p = pos + 4
        end = p + block
        ref_id, start, l_name, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", data, p)
        p += 32
        qname = data[p:p + l_name - 1].decode()
        p += l_name
        cigar = []
        for k in range(n_cig):
            v = struct.unpack_from("<I", data, p + 4 * k)[0]
            cigar.append((_CIGAR_OPS[v & 0xF], v >> 4))
        p += 4 * n_cig
        packed = data[p:p + (l_seq + 1) // 2]
        seq = "".join(_BAM_SEQ[b >> 4] + _BAM_SEQ[b & 0xF] for b in packed)[:l_seq]
        p += (l_seq + 1) // 2
        q = data[p:p + l_seq]
        qual = "*" if (l_seq and q[0] == 0xFF) else "".join(chr(b + 33) for b in q)
        p += l_seq
        tags = _parse_bam_tags(data, p, end)
        records.append(Alignment(
            qname=qname, flag=flag, contig=refs[ref_id] if ref_id >= 0 else "*", start=start,
            mapping_quality=mapq, cigar=cigar, seq=seq, qual=qual, tags=tags))
        pos = end
    return header, records


class AlignmentFile(object):
    """A whole SAM or BAM file in memory with region queries in file order
    (``header["RG"]`` and ``fetch`` like the pysam object used in mchap/io/bam.py:108-125)."""

    def __init__(self, path):
        self.filename = str(path)
        with open(path, "rb") as f:
            magic = f.read(3)
        if magic == _GZIP_MAGIC:
            with gzip.open(path, "rb") as f:
                self.header, self.records = _read_bam(f)
        elif magic[:1] == b"@" or str(path).endswith(".sam"):
            with open(path, "r") as f:
                self.header, self.records = _read_sam(f)
        else:
            raise ValueError("'%s' is neither a SAM nor a BAM file (CRAM is not supported by the built-in reader)" % path)
        self._by_contig = {}
        for i, r in enumerate(self.records):
            if not r.is_unmapped or r.contig != "*":
                self._by_contig.setdefault(r.contig, []).append(i)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def fetch(self, contig, start, stop):
        """Records overlapping the half-open interval, in file order."""
        for i in self._by_contig.get(contig, ()):
            r = self.records[i]
            end = max(r.reference_end, r.start + 1)
            if r.start < stop and end > start:
                yield r


# ------------------------------------------------------------------------------ reference
class FastaFile(object):
    def __init__(self, path):
        self.filename = str(path)
        self.references, self._seqs = [], {}
        name, chunks = None, []
        with _open_maybe_gzip(path) as f:
            for raw in f:
                line = raw.decode().rstrip("\r\n")
                if line.startswith(">"):
                    if name is not None:
                        self._seqs[name] = "".join(chunks)
                    name = line[1:].split()[0]
                    self.references.append(name)
                    chunks = []
                elif name is not None:
                    chunks.append(line.strip())
        if name is not None:
            self._seqs[name] = "".join(chunks)
        self.lengths = [len(self._seqs[r]) for r in self.references]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def fetch(self, contig, start, stop):
        return self._seqs[contig][start:stop]


# ------------------------------------------------------------------------------ targets
def read_bed4_lines(path):
    """(contig, start, stop, name-or-None) per line of a BED3/BED4 file, plain or gzipped
    (mchap/io/loci.py:327-372)."""
    with _open_maybe_gzip(path) as f:
        for raw in f:
            if raw.startswith(b"#") or not raw.strip():
                continue
            parts = raw.decode().split()
            yield parts[0], int(parts[1]), int(parts[2]), (parts[3] if len(parts) > 3 else None)


# ------------------------------------------------------------------------------ variants
@dataclass
class VariantRecord(object):
    """One VCF data line with the attribute names of pysam.VariantRecord that
    mchap/io/loci.py:100-137 and 205-320 read."""

    contig: str
    start: int          # 0-based
    id: str             # None for "."
    ref: str
    alts: tuple         # None for "."
    info: dict
    header: object

    @property
    def chrom(self):
        return self.contig

    @property
    def stop(self):
        return self.start + len(self.ref)


@dataclass
class _InfoMeta(object):
    number: object
    type: str


class _Header(object):
    def __init__(self):
        self.info = {}
        self.contigs = []   # [(name, length or None)]
        self.samples = []


def _split_meta(body):
    """Key=value pairs of a ##KEY=<...> header body (values may be quoted and contain commas)."""
    out, key, val, in_q, cur = {}, None, [], False, []
    for c in body:
        if in_q:
            if c == '"':
                in_q = False
            else:
                cur.append(c)
        elif c == '"':
            in_q = True
        elif c == "=" and key is None:
            key = "".join(cur)
            cur = []
        elif c == ",":
            out[key] = "".join(cur)
            key, cur = None, []
        else:
            cur.append(c)
    if key is not None:
        out[key] = "".join(cur)
    return out


def _typed_info(raw, meta):
    """Value of an INFO key the way pysam exposes it: flags -> True, Number=1 -> scalar, otherwise a
    tuple; floats pass through float32 like htslib's parsed representation; '.' -> None."""
    if meta is None:
        conv = str
        number = "."
    else:
        number = meta.number
        if meta.type == "Flag":
            return True
        if meta.type == "Integer":
            conv = int
        elif meta.type == "Float":
            conv = lambda s: float(np.float32(s))  # noqa: E731
        else:
            conv = str
    if raw is None:
        return True
    vals = tuple(None if v == "." else conv(v) for v in raw.split(","))
    if number in (1, "1"):
        return vals[0]
    return vals


class VariantFile(object):
    """A whole VCF (plain or bgzipped) in memory with region queries in file order."""

    def __init__(self, path):
        self.filename = str(path)
        self.header = _Header()
        self.records = []
        with _open_maybe_gzip(path) as f:
            for raw in f:
                line = raw.decode().rstrip("\r\n")
                if not line:
                    continue
                if line.startswith("##"):
                    self._meta(line)
                elif line.startswith("#"):
                    self.header.samples = line.split("\t")[9:]
                else:
                    self.records.append(self._record(line))
        self._by_contig = {}
        for i, r in enumerate(self.records):
            self._by_contig.setdefault(r.contig, []).append(i)
        self._starts = {c: [self.records[i].start for i in idx] for c, idx in self._by_contig.items()}
        self._sorted = {c: all(a <= b for a, b in zip(s, s[1:])) for c, s in self._starts.items()}
        self._maxlen = {c: max(len(self.records[i].ref) for i in idx) for c, idx in self._by_contig.items()}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def _meta(self, line):
        if line.startswith("##INFO=<") and line.endswith(">"):
            d = _split_meta(line[8:-1])
            num = d.get("Number", ".")
            self.header.info[d["ID"]] = _InfoMeta(int(num) if num.isdigit() else num, d.get("Type", "String"))
        elif line.startswith("##contig=<") and line.endswith(">"):
            d = _split_meta(line[10:-1])
            length = d.get("length")
            self.header.contigs.append((d["ID"], int(length) if length and length.isdigit() else None))

    def _record(self, line):
        f = line.split("\t")
        info = {}
        if len(f) > 7 and f[7] != ".":
            for item in f[7].split(";"):
                if not item:
                    continue
                key, _, raw = item.partition("=")
                info[key] = _typed_info(raw if _ else None, self.header.info.get(key))
        return VariantRecord(
            contig=f[0], start=int(f[1]) - 1, id=None if f[2] == "." else f[2], ref=f[3],
            alts=None if f[4] == "." else tuple(f[4].split(",")), info=info, header=self.header)

    def fetch(self, contig=None, start=None, stop=None):
        """All records (no arguments) or those overlapping [start, stop) on a contig, in file order."""
        if contig is None:
            yield from self.records
            return
        idx = self._by_contig.get(contig, [])
        if not idx:
            return
        lo, hi = 0, len(idx)
        if self._sorted[contig] and start is not None:
            starts = self._starts[contig]
            lo = bisect.bisect_left(starts, start - self._maxlen[contig] + 1)
            hi = bisect.bisect_left(starts, stop)
        for i in idx[lo:hi]:
            r = self.records[i]
            if start is None or (r.start < stop and r.stop > start):
                yield r
