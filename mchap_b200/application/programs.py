"""``mchap assemble`` / ``mchap call`` / ``mchap call-exact`` as batched device pipelines.

SURVEY.md section 8(f) N3.  The reference walks loci one by one and, inside a locus, samples one by
one, calling a sampler per (locus, sample) (mchap/application/baseclass.py:304-325,
assemble.py:95-252, call.py:49-200, call_exact.py:52-199).  Here a *block* of loci is processed at a
time:

    host   read extraction + allele calls for every (locus, sample) of the block
    device one batched call per stage over all (locus, sample) items of the block
           (read encoding + de-duplication -> sampler / exhaustive caller -> trace tallies,
           minimum error correction, genotype likelihoods)
    host   per-locus summaries and VCF lines, in the original locus order

The VCF text is byte-identical to the reference's (tests/test_application_vcf.py runs the scenarios of
the reference's own golden files).  Field names of the program dataclasses follow the reference so
that ``program(**arguments)`` reads the same.
"""
from dataclasses import dataclass, field

import numpy as np

from .. import __version__ as _PACKAGE_VERSION
from .. import combinatorics
from ..api import default_device
from ..assemble.classes import TraceTally
from ..assemble.mcmc import DenovoMCMC
from ..calling import exact
from ..calling.classes import AllelesTraceTally, CallingMCMC
from ..encoding import call_probabilities, encode_unique_reads_batch
from ..jitutils import genotype_alleles_as_index, natural_log_to_log10
from . import hostio, vcfout
from .loci import Locus, LocusPrior, read_bed4
from .reads import extract_read_variants, open_alignment, qual_of_prob, read_depth

PFEIFFER_ERROR = 0.0024  # default base error rate (reference mchap/constant.py:3)

LOCUS_ERROR = "Exception encountered at locus: '{name}', '{contig}:{start}-{stop}'."
SAMPLE_ERROR = "Exception encountered when assembling sample '{sample}'."


class LocusAssemblyError(Exception):
    pass


class SampleAssemblyError(Exception):
    pass


@dataclass
class LocusData(object):
    """Everything that becomes one VCF record (the reference's LocusAssemblyData)."""

    locus: object
    read_calls: dict = field(default_factory=dict)    # sample -> int8[n_reads, n_snvs]
    read_probs: dict = field(default_factory=dict)    # sample -> f64[n_reads, n_snvs] P(call correct)
    read_dists: dict = field(default_factory=dict)    # sample -> (reads f64[U, N, A], counts) when needed
    columns: dict = field(default_factory=lambda: {"FILTER": []})
    info: dict = field(default_factory=dict)
    sample: dict = field(default_factory=dict)        # FORMAT id -> {sample: value}

    def put(self, fmt_id, sample, value):
        self.sample.setdefault(fmt_id, {})[sample] = value


def _wrap_sample_errors(samples, results):
    """First per-item exception of a locus -> SampleAssemblyError naming the sample."""
    for s, r in zip(samples, results):
        if isinstance(r, BaseException):
            raise SampleAssemblyError(SAMPLE_ERROR.format(sample=s)) from r


@dataclass
class program(object):
    vcf: str
    ref: str
    samples: list
    sample_bams: dict
    sample_ploidy: dict
    sample_inbreeding: dict
    read_group_field: str = "SM"
    base_error_rate: float = PFEIFFER_ERROR
    ignore_base_phred_scores: bool = True
    mapping_quality: int = 20
    skip_duplicates: bool = True
    skip_qcfail: bool = True
    skip_supplementary: bool = True
    info_fields: list = None       # INFO ids in output order
    format_fields: list = None     # FORMAT ids in output order
    n_cores: int = 1               # accepted for drop-in compatibility: the device replaces the process pool
    precision: int = 3
    random_seed: int = 42
    cli_command: str = None
    device: object = None          # mchap_b200.Device; created on first use from device_ordinal
    device_ordinal: int = 0
    block_loci: int = 4096         # loci per device batch

    # ------------------------------------------------------------------ to be specialised
    def loci(self):
        raise NotImplementedError()

    def call_block(self, block):
        raise NotImplementedError()

    def header_contigs(self):
        return list(hostio.VariantFile(self.vcf).header.contigs)

    # ------------------------------------------------------------------ shared machinery
    def _device(self):
        if self.device is None:
            self.device = default_device(self.device_ordinal)
        return self.device

    def require_AFP(self):
        wanted = {"ACP", "AFP", "AOP", "AOPSUM"}
        return bool(wanted & set(self.info_fields)) or bool({"ACP", "AFP", "AOP"} & set(self.format_fields))

    def header(self):
        return vcfout.header_lines(self.samples, self.header_contigs(), self.info_fields, self.format_fields,
                                   self.cli_command, self.random_seed, _PACKAGE_VERSION)

    def encode_sample_reads(self, data, cache):
        """Host part of baseclass.py:134-215: pooled read extraction, depth statistics, integer allele
        calls and the per-base probability of a correct call.  The probabilistic encoding and the
        de-duplication happen on the device."""
        locus = data.locus
        n_pos = len(locus.variants)
        for sample in self.samples:
            try:
                chars, quals = [], []
                for name, path in self.sample_bams[sample]:
                    c, q = extract_read_variants(
                        locus, open_alignment(path, cache), name, id=self.read_group_field,
                        min_quality=self.mapping_quality, skip_duplicates=self.skip_duplicates,
                        skip_qcfail=self.skip_qcfail, skip_supplementary=self.skip_supplementary)
                    chars.append(c)
                    quals.append(q)
                if chars:
                    chars, quals = np.concatenate(chars), np.concatenate(quals)
                else:
                    chars, quals = np.empty((0, n_pos), dtype="U1"), np.empty((0, n_pos), dtype=np.int16)
                data.put("RCOUNT", sample, chars.shape[0])
                depth = read_depth(chars)
                if len(depth) == 0:
                    depth = np.array(np.nan)
                data.put("DP", sample, np.round(np.mean(depth)))
                data.put("SNVDP", sample, np.round(depth))
                calls = locus.encode_read_chars(chars)
                data.read_calls[sample] = calls
                data.read_probs[sample] = call_probabilities(
                    calls, None if self.ignore_base_phred_scores else quals, self.base_error_rate)
                data.put("RCALLS", sample, np.sum(calls >= 0))
            except Exception as e:
                raise SampleAssemblyError(SAMPLE_ERROR.format(sample=sample)) from e
        return data

    def encode_block_reads(self, block, items):
        """Device: distinct encoded reads + counts of the given (data, sample) items."""
        if not items:
            return
        n_alleles = [d.locus.count_alleles() for d, _ in items]
        pairs = encode_unique_reads_batch(
            [d.read_calls[s] for d, s in items], [d.read_probs[s] for d, s in items], n_alleles, 3, self._device())
        for (d, s), pair in zip(items, pairs):
            d.read_dists[s] = pair

    def mec_block(self, items, genotypes):
        """Device: MEC / MECP of the called genotype of every (data, sample) item."""
        live = [k for k, (d, _) in enumerate(items) if len(d.locus.variants) > 0]
        mec = np.zeros(len(items), dtype=np.int64)
        called = np.zeros(len(items), dtype=np.int64)
        if live:
            m, c = self._device().minimum_error_correction_batch(
                [items[k][0].read_calls[items[k][1]] for k in live], [genotypes[k] for k in live])
            mec[live], called[live] = m, c
        for k, (d, s) in enumerate(items):
            d.put("MEC", s, mec[k])
            d.put("MECP", s, mec[k] / called[k] if called[k] > 0 else np.nan)

    def likelihoods_block(self, items, haplotypes):
        """Device: GL arrays (log10) of every (data, sample) item against its locus' haplotypes."""
        need = [(d, s) for d, s in items if s not in d.read_dists]
        self.encode_block_reads(None, need)
        gls = exact.genotype_likelihoods_batch(
            [d.read_dists[s][0] for d, s in items], [self.sample_ploidy[s] for _, s in items], haplotypes,
            [d.read_dists[s][1] for d, s in items], self._device())
        for (d, s), gl in zip(items, gls):
            d.put("GL", s, natural_log_to_log10(gl))

    def mock_invalid_locus(self, data):
        for sample in self.samples:
            ploidy = self.sample_ploidy[sample]
            data.put("GT", sample, np.full(ploidy, -1, int))
            for f in ("GQ", "GPM", "SPM", "SQ", "MCI", "MEC", "MECP"):
                data.put(f, sample, np.nan)
            for f in ("ACP", "AFP", "AOP", "GP", "GL"):
                data.put(f, sample, np.array([np.nan]))

    def summarise(self, data):
        """Record-level columns and INFO values from the per-sample results (baseclass.py:220-302)."""
        locus, col, info, smp = data.locus, data.columns, data.info, data.sample
        col["CHROM"], col["POS"], col["ID"], col["QUAL"] = locus.contig, locus.start + 1, locus.name, np.nan
        info["END"] = locus.stop
        info["NVAR"] = len(locus.variants)
        info["SNVPOS"] = np.subtract(locus.positions, locus.start) + 1
        if len(col["FILTER"]) == 0:
            col["FILTER"] = "PASS"
        n_allele = len(col["ALT"]) + 1
        counts = np.zeros(n_allele, int)
        for gt in smp["GT"].values():
            for a in gt:
                if a >= 0:
                    counts[a] += 1
        info["AC"] = counts[1:]
        info["AN"] = np.sum(counts)
        info["UAN"] = np.sum(counts > 0)
        info["NS"] = sum(np.any(a >= 0) for a in smp["GT"].values())
        info["MCI"] = sum(m > 0 for m in smp["MCI"].values())
        info["DP"] = np.nan if len(locus.variants) == 0 else np.nansum(list(smp["DP"].values()))
        info["RCOUNT"] = np.nansum(list(smp["RCOUNT"].values()))
        blank = np.full(n_allele, np.nan)
        if "ACP" in self.info_fields:
            v = sum(smp["ACP"].values())
            info["ACP"] = blank if np.isnan(v).all() else v
        if "AFP" in self.info_fields:
            v = sum(smp["ACP"].values()) / sum(self.sample_ploidy.values())
            info["AFP"] = blank if np.isnan(v).all() else v
        if "AOPSUM" in self.info_fields:
            v = sum(smp["AOP"].values())
            info["AOPSUM"] = blank if np.isnan(v).all() else v
        if "AOP" in self.info_fields:
            absent = np.ones(n_allele, float)
            for occur in smp["AOP"].values():
                absent = absent * (1 - occur)
            info["AOP"] = 1 - absent
        if "SNVDP" in self.info_fields:
            info["SNVDP"] = sum(smp["SNVDP"].values())
        return data

    def format_record(self, data):
        info = vcfout.info_text(self.info_fields, data.info, self.precision)
        per_sample = {f: [data.sample.get(f, {}).get(s) for s in self.samples] for f in self.format_fields}
        cells = vcfout.samples_text(self.format_fields, per_sample, self.precision)
        c = data.columns
        return vcfout.record_line(c["CHROM"], c["POS"], c["ID"], c["REF"], c["ALT"], c["QUAL"], c["FILTER"], info,
                                  cells, self.precision)

    def records(self):
        """VCF record lines in locus order, one device batch per block of loci."""
        cache = {}
        block = []

        def flush():
            try:
                self.call_block(block)
            except LocusAssemblyError:
                raise
            except Exception as e:
                where = getattr(e, "locus", None) or block[0].locus
                msg = LOCUS_ERROR.format(name=where.name, contig=where.contig, start=where.start, stop=where.stop)
                raise LocusAssemblyError(msg) from e
            lines = [self.format_record(self.summarise(d)) for d in block]
            block.clear()
            return lines

        for locus in self.loci():
            data = LocusData(locus)
            try:
                self.encode_sample_reads(data, cache)
            except Exception as e:
                msg = LOCUS_ERROR.format(name=locus.name, contig=locus.contig, start=locus.start, stop=locus.stop)
                raise LocusAssemblyError(msg) from e
            block.append(data)
            if len(block) >= self.block_loci:
                yield from flush()
        if block:
            yield from flush()

    def run_stdout(self, out=None):
        import sys

        out = out or sys.stdout
        for line in self.header():
            out.write(line + "\n")
        for line in self.records():
            out.write(line + "\n")


# ======================================================================================= assemble
def call_posterior_haplotypes(posteriors, threshold=0.01):
    """Haplotypes reported as VCF alleles: every haplotype whose posterior probability of occurring
    reaches the threshold in at least one sample, ordered by descending summed expected copies with
    the reference allele (all zeros) first (mchap/assemble/haplotype_calling.py:4-64).
    Returns (haplotypes int8[n_alleles, n_snvs], reference allele observed)."""
    seen, weight = {}, {}
    for post in posteriors:
        haps, copies, occur = post.allele_frequencies(dosage=True)
        keep = occur >= threshold
        for h, w in zip(haps[keep], copies[keep]):
            key = h.tobytes()
            if key not in seen:
                seen[key], weight[key] = h, 0
            weight[key] += w
    ref_key = None
    for key, h in seen.items():
        if np.all(h == 0):
            ref_key = key
    if ref_key is not None:
        seen.pop(ref_key)
        weight.pop(ref_key)
    n_base = posteriors[0].genotypes.shape[-1]
    haplotypes = np.full((len(seen) + 1, n_base), -1, np.int8)
    values = np.full(len(seen) + 1, -1, float)
    for i, (key, h) in enumerate(seen.items()):
        haplotypes[i], values[i] = h, weight[key]
    haplotypes[-1][:] = 0
    values[-1] = values.max() + 1
    order = np.flip(np.argsort(values))
    return haplotypes[order], ref_key is not None


def _alleles_of(genotype, labels):
    """VCF allele numbers of a genotype's haplotypes, sorted, unlabelled (-1) last."""
    a = np.sort([labels.get(h.tobytes(), -1) for h in genotype])
    return np.append(a[a >= 0], a[a < 0])


@dataclass
class assemble_program(program):
    bed: str = ""
    region: str = None
    region_id: str = None
    haplotype_posterior_threshold: float = 0.2
    mcmc_chains: int = 1
    mcmc_steps: int = 2000
    mcmc_burn: int = 1000
    mcmc_alpha: float = 1.0
    mcmc_beta: float = 3.0
    mcmc_fix_homozygous: float = 0.999
    mcmc_recombination_step_probability: float = 0.5
    mcmc_partial_dosage_step_probability: float = 0.5
    mcmc_dosage_step_probability: float = 1.0
    mcmc_incongruence_threshold: float = 0.60
    mcmc_llk_cache_threshold: int = 100
    sample_mcmc_temperatures: dict = None

    def loci(self):
        if self.bed is None and self.region is None:
            raise ValueError("No region or targets bedfile is specified.")
        fasta, variants = hostio.FastaFile(self.ref), hostio.VariantFile(self.vcf)
        targets = read_bed4(self.bed) if self.bed is not None else [Locus.from_region_string(self.region, self.region_id)]
        for t in targets:
            yield t.set_sequence(fasta).set_variants(variants)

    def header_contigs(self):
        f = hostio.FastaFile(self.ref)
        return list(zip(f.references, f.lengths))

    def _model(self):
        return DenovoMCMC(
            ploidy=2, n_alleles=None, inbreeding=None, steps=self.mcmc_steps, chains=self.mcmc_chains,
            alpha=self.mcmc_alpha, beta=self.mcmc_beta, fix_homozygous=self.mcmc_fix_homozygous,
            recombination_step_probability=self.mcmc_recombination_step_probability,
            partial_dosage_step_probability=self.mcmc_partial_dosage_step_probability,
            dosage_step_probability=self.mcmc_dosage_step_probability, random_seed=self.random_seed,
            llk_cache_threshold=self.mcmc_llk_cache_threshold, device=self._device())

    def call_block(self, block):
        items = [(d, s) for d in block for s in self.samples]
        kept = max(self.mcmc_steps - self.mcmc_burn, 0)
        # ---- device: encode + assemble + tally for every item with at least one SNV
        live = [k for k, (d, _) in enumerate(items) if len(d.locus.variants) > 0]
        tallies = [None] * len(items)
        if live:
            res, _ = self._model().fit_posterior_from_calls_batch(
                [items[k][0].read_calls[items[k][1]] for k in live],
                [items[k][0].read_probs[items[k][1]] for k in live],
                burn=self.mcmc_burn,
                n_alleles_list=[items[k][0].locus.count_alleles() for k in live],
                ploidy_list=[self.sample_ploidy[items[k][1]] for k in live],
                inbreeding_list=None if self.sample_inbreeding is None else [
                    self.sample_inbreeding[items[k][1]] for k in live],
                temperatures_list=None if self.sample_mcmc_temperatures is None else [
                    self.sample_mcmc_temperatures[items[k][1]] for k in live],
                errors="return")
            for k, r in zip(live, res):
                tallies[k] = r
        for k, (d, s) in enumerate(items):
            if tallies[k] is None:
                # no SNV in the interval: the only genotype is the reference haplotype (mcmc.py:188-199)
                P = self.sample_ploidy[s]
                tallies[k] = TraceTally(np.zeros((1, P, 0), dtype=np.int8),
                                        np.full((1, self.mcmc_chains), kept, dtype=np.int64),
                                        np.zeros((1, self.mcmc_chains), dtype=np.int64))
        ns = len(self.samples)
        for b, d in enumerate(block):
            try:
                _wrap_sample_errors(self.samples, tallies[b * ns:(b + 1) * ns])
            except SampleAssemblyError as e:
                e.locus = d.locus
                raise
        # ---- host: per-sample summaries of the posterior
        posteriors, modes = [], []
        for (d, s), tally in zip(items, tallies):
            posterior = tally.posterior()
            support = posterior.mode_genotype_support()
            support_prob = support.probabilities.sum()
            genotype, genotype_prob = support.mode_genotype()
            d.put("SPM", s, support_prob)
            d.put("SQ", s, qual_of_prob(support_prob))
            d.put("GQ", s, qual_of_prob(genotype_prob))
            d.put("GPM", s, genotype_prob)
            d.put("MCI", s, tally.replicate_incongruence(threshold=self.mcmc_incongruence_threshold))
            posteriors.append(posterior)
            modes.append(genotype)
        self.mec_block(items, modes)
        # ---- host: alleles of every locus from the posteriors of its samples
        gl_items, gl_haps = [], []
        for b, d in enumerate(block):
            posts = posteriors[b * ns:(b + 1) * ns]
            haplotypes, ref_called = call_posterior_haplotypes(posts, threshold=self.haplotype_posterior_threshold)
            labels = {h.tobytes(): i for i, h in enumerate(haplotypes)}
            d.info["REFMASKED"] = not ref_called
            if not ref_called:
                labels.pop(haplotypes[0].tobytes())
                if len(haplotypes) == 1:
                    d.columns["FILTER"].append("NOA")
            d.columns["REF"] = d.locus.sequence
            d.columns["ALT"] = d.locus.format_haplotypes(haplotypes[1:]) if len(haplotypes) > 1 else []
            for j, s in enumerate(self.samples):
                post = posts[j]
                d.put("GT", s, _alleles_of(modes[b * ns + j], labels))
                if self.require_AFP():
                    freqs, occur = np.zeros(len(haplotypes)), np.zeros(len(haplotypes))
                    haps, f, o = post.allele_frequencies()
                    index = {h.tobytes(): i for i, h in enumerate(haps)}
                    idx = np.array([index.get(h.tobytes(), -1) for h in haplotypes], dtype=int)
                    freqs[idx >= 0] = f[idx[idx >= 0]]
                    occur[idx >= 0] = o[idx[idx >= 0]]
                    d.put("AFP", s, freqs)
                    d.put("AOP", s, occur)
                    d.put("ACP", s, freqs * self.sample_ploidy[s])
                if "GP" in self.format_fields:
                    ploidy = post.genotypes.shape[1]
                    probs = np.zeros(combinatorics.count_unique_genotypes(len(labels), ploidy), float)
                    for haps, p in zip(post.genotypes, post.probabilities):
                        alleles = np.sort([labels.get(h.tobytes(), -1) for h in haps])
                        if alleles[0] >= 0:
                            probs[genotype_alleles_as_index(alleles)] = p
                    d.put("GP", s, probs)
                if "GL" in self.format_fields:
                    gl_items.append((d, s))
                    gl_haps.append(haplotypes)
        if gl_items:
            zero = [(d, s) for d, s in gl_items if len(d.locus.variants) == 0]
            for d, s in zero:
                # a locus without SNVs has the reference genotype only: likelihood 1
                d.put("GL", s, natural_log_to_log10(np.zeros(1, dtype=np.float32)))
            live_gl = [(k, ds) for k, ds in enumerate(gl_items) if len(ds[0].locus.variants) > 0]
            if live_gl:
                self.likelihoods_block([ds for _, ds in live_gl], [gl_haps[k] for k, _ in live_gl])


# ======================================================================================= call / call-exact
@dataclass
class known_haplotypes_program(program):
    prior_frequencies_tag: str = None
    filter_input_haplotypes: str = None

    def loci(self):
        for record in hostio.VariantFile(self.vcf).fetch():
            yield LocusPrior.from_variant_record(record, frequency_tag=self.prior_frequencies_tag,
                                                 allele_filter=self.filter_input_haplotypes)

    def _locus_columns(self, d):
        locus = d.locus
        d.columns["REF"], d.columns["ALT"] = locus.sequence, locus.alts
        d.info["REFMASKED"] = locus.mask_reference_allele
        d.info["AFPRIOR"] = locus.frequencies

    def _prior(self, sample, frequencies):
        if self.sample_inbreeding is None:
            return None
        return (self.sample_inbreeding[sample], frequencies)

    def _store_call(self, d, s, alleles, genotype_prob, support_prob, incongruence):
        d.put("GT", s, alleles)
        d.put("GQ", s, qual_of_prob(genotype_prob))
        d.put("GPM", s, genotype_prob)
        d.put("SPM", s, support_prob)
        d.put("SQ", s, qual_of_prob(support_prob))
        d.put("MCI", s, incongruence)


@dataclass
class call_program(known_haplotypes_program):
    mcmc_chains: int = 1
    mcmc_steps: int = 2000
    mcmc_burn: int = 1000
    mcmc_incongruence_threshold: float = 0.60

    def call_block(self, block):
        kept = max(self.mcmc_steps - self.mcmc_burn, 0)
        work = []   # (data, all haplotypes, sampler haplotypes, sampler frequencies, relabelling or None)
        for d in block:
            locus = d.locus
            haplotypes = locus.encode_haplotypes()
            freqs = locus.frequencies
            self._locus_columns(d)
            mask = np.zeros(len(haplotypes), bool)
            mask[0] = locus.mask_reference_allele
            mask |= freqs == 0
            if np.any(mask):
                sub_haps, sub_freqs, relabel = haplotypes[~mask], freqs[~mask], np.where(~mask)[0]
            else:
                sub_haps, sub_freqs, relabel = haplotypes, freqs, None
            if len(sub_haps) == 0:
                d.columns["FILTER"].append("NOA")
                self.mock_invalid_locus(d)
            elif freqs is not None and np.any(np.isnan(freqs)):
                d.columns["FILTER"].append("AF0")
                self.mock_invalid_locus(d)
            else:
                work.append((d, haplotypes, sub_haps, sub_freqs, relabel))
        items = [(w, s) for w in work for s in self.samples]
        live = [k for k, (w, _) in enumerate(items) if len(w[0].locus.variants) > 0]
        self.encode_block_reads(None, [(items[k][0][0], items[k][1]) for k in live])
        tallies = [None] * len(items)
        if live:
            model = CallingMCMC(ploidy=2, haplotypes=None, prior=None, steps=self.mcmc_steps, chains=self.mcmc_chains,
                                random_seed=self.random_seed, device=self._device())
            res = model.fit_posterior_batch(
                [items[k][0][0].read_dists[items[k][1]][0] for k in live],
                [items[k][0][0].read_dists[items[k][1]][1] for k in live],
                burn=self.mcmc_burn,
                haplotypes_list=[items[k][0][2] for k in live],
                priors=None if self.sample_inbreeding is None else [
                    self._prior(items[k][1], items[k][0][3]) for k in live],
                ploidy_list=[self.sample_ploidy[items[k][1]] for k in live], errors="return")
            for k, r in zip(live, res):
                tallies[k] = r
        for k, (w, s) in enumerate(items):
            if tallies[k] is None:
                # no SNV: reference allele only (calling/classes.py:75-82)
                assert len(w[2]) == 1
                P = self.sample_ploidy[s]
                tallies[k] = AllelesTraceTally(np.zeros((1, P), dtype=np.int32),
                                               np.full((1, self.mcmc_chains), kept, dtype=np.int64),
                                               np.zeros((1, self.mcmc_chains), dtype=np.int64), len(w[2]))
        ns = len(self.samples)
        for b, w in enumerate(work):
            try:
                _wrap_sample_errors(self.samples, tallies[b * ns:(b + 1) * ns])
            except SampleAssemblyError as e:
                e.locus = w[0].locus
                raise
        genotypes = []
        for (w, s), tally in zip(items, tallies):
            d, haplotypes, _, _, relabel = w
            if relabel is not None:
                tally = tally.relabel(relabel)
            incongruence = tally.replicate_incongruence(threshold=self.mcmc_incongruence_threshold)
            posterior = tally.posterior()
            alleles, genotype_prob, support_prob = posterior.mode(genotype_support=True)
            self._store_call(d, s, alleles, genotype_prob, support_prob, incongruence)
            genotypes.append(haplotypes[alleles])
            if self.require_AFP():
                frequencies, counts, occurrence = tally.posterior_frequencies()
                d.put("ACP", s, counts)
                d.put("AFP", s, frequencies)
                d.put("AOP", s, occurrence)
            if "GP" in self.format_fields:
                d.put("GP", s, posterior.as_array(len(haplotypes)))
        self.mec_block([(w[0], s) for w, s in items], genotypes)
        if "GL" in self.format_fields:
            self._likelihoods(items)

    def _likelihoods(self, items):
        zero = [(w, s) for w, s in items if len(w[0].locus.variants) == 0]
        for w, s in zero:
            w[0].put("GL", s, natural_log_to_log10(np.zeros(1, dtype=np.float32)))
        live = [(w, s) for w, s in items if len(w[0].locus.variants) > 0]
        if live:
            self.likelihoods_block([(w[0], s) for w, s in live], [w[1] for w, _ in live])


@dataclass
class call_exact_program(known_haplotypes_program):
    random_seed: int = None

    def call_block(self, block):
        work = []
        for d in block:
            locus = d.locus
            haplotypes = locus.encode_haplotypes()
            freqs = locus.frequencies
            self._locus_columns(d)
            if locus.mask_reference_allele:
                assert (freqs[0] == 0) or np.isnan(freqs[0])
            if locus.mask_reference_allele and len(haplotypes) == 1:
                d.columns["FILTER"].append("NOA")
                self.mock_invalid_locus(d)
            elif np.any(np.isnan(freqs)):
                d.columns["FILTER"].append("AF0")
                self.mock_invalid_locus(d)
            else:
                work.append((d, haplotypes, freqs))
        items = [(w, s) for w in work for s in self.samples]
        if not items:
            return
        self.encode_block_reads(None, [(w[0], s) for w, s in items])
        reads = [w[0].read_dists[s][0] for w, s in items]
        counts = [w[0].read_dists[s][1] for w, s in items]
        ploidy = [self.sample_ploidy[s] for _, s in items]
        haps = [w[1] for w, _ in items]
        priors = None if self.sample_inbreeding is None else [self._prior(s, w[2]) for w, s in items]
        dev = self._device()
        full = ("GL" in self.format_fields) or ("GP" in self.format_fields)
        genotypes = []
        if full:
            # every likelihood and posterior is reported: keep the arrays (call_exact.py:126-160)
            llks = exact.genotype_likelihoods_batch(reads, ploidy, haps, counts, dev)
            gps, trips = exact.genotype_posteriors_batch(llks, ploidy, [len(h) for h in haps], priors, dev,
                                                         with_frequencies=True)
            for (w, s), P, llk, probs, (fr, ct, oc) in zip(items, ploidy, llks, gps, trips):
                d = w[0]
                idx = int(np.argmax(probs))
                from ..jitutils import index_as_genotype_alleles

                alleles = index_as_genotype_alleles(idx, P, dev)
                _, support = exact.alternate_dosage_posteriors(alleles, probs)
                self._store_call(d, s, alleles, probs[idx], support.sum(), np.nan)
                if self.require_AFP():
                    d.put("ACP", s, ct)
                    d.put("AFP", s, fr)
                    d.put("AOP", s, oc)
                if "GL" in self.format_fields:
                    d.put("GL", s, natural_log_to_log10(llk))
                if "GP" in self.format_fields:
                    d.put("GP", s, probs)
                genotypes.append(w[1][alleles])
        else:
            res = exact.posterior_mode_batch(reads, ploidy, haps, counts, priors, dev)
            for (w, s), P, (alleles, _, genotype_prob, support_prob, fr, oc) in zip(items, ploidy, res):
                d = w[0]
                self._store_call(d, s, alleles, genotype_prob, support_prob, np.nan)
                d.put("ACP", s, fr * P)
                d.put("AFP", s, fr)
                d.put("AOP", s, oc)
                genotypes.append(w[1][alleles])
        self.mec_block([(w[0], s) for w, s in items], genotypes)
