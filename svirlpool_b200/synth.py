"""Seeded synthetic candidate regions with ONT-like reads (BASELINE.json configs 2-4).

Everything is drawn from numpy.random.Generator(PCG64(seed)); region r of config 2 uses seed
10_000 + r, so any subset of regions can be regenerated independently (bench.py, tests, the CPU
baseline and every rank of a multi-GPU run see the same reads for the same region id).

Error model "ONT-like 10 %": per base substitution 3.5 %, insertion 3.0 %, deletion 3.5 %, indel
lengths geometric(p=0.7), indel rates x3 inside homopolymer runs of length >= 4.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


@dataclass
class Region:
    region_id: int
    names: list[str]
    reads: list[np.ndarray]            # base codes 0..3, ORIGINAL read orientation
    strands: list[str]                 # '+' / '-' relative to the region's reference
    offsets: list[int]                 # start of the read on its haplotype (bases trimmed at the left end)
    haplotypes: list[int]
    sv_size: int = 0                   # signed: >0 insertion on hap1, <0 deletion on hap1
    meta: dict = field(default_factory=dict)

    def read_strings(self) -> dict[str, str]:
        return {n: codes_to_str(r) for n, r in zip(self.names, self.reads)}


def codes_to_str(codes: np.ndarray) -> str:
    return BASES[codes].tobytes().decode()


def revcomp_codes(codes: np.ndarray) -> np.ndarray:
    return _COMP[codes[::-1]]


def _homopolymer_mask(codes: np.ndarray, min_run: int = 4) -> np.ndarray:
    """True for bases inside a run of >= min_run identical bases."""
    n = len(codes)
    if n == 0:
        return np.zeros(0, dtype=bool)
    change = np.flatnonzero(np.diff(codes)) + 1
    starts = np.concatenate(([0], change))
    lengths = np.diff(np.concatenate((starts, [n])))
    return np.repeat(lengths >= min_run, lengths)


def noisy_copy(rng: np.random.Generator, codes: np.ndarray, sub: float = 0.035, ins: float = 0.03, dele: float = 0.035,
               geo_p: float = 0.7, hp_factor: float = 3.0) -> np.ndarray:
    """One noisy read of a template (vectorised)."""
    n = len(codes)
    hp = _homopolymer_mask(codes)
    scale = np.where(hp, hp_factor, 1.0)
    u = rng.random(n)
    # deletions: events start with prob dele*scale, remove geometric(p) bases
    del_start = u < dele * scale
    del_len = np.where(del_start, rng.geometric(geo_p, n), 0)
    removed = np.zeros(n + 1, dtype=np.int64)
    idx = np.flatnonzero(del_start)
    np.add.at(removed, idx, 1)
    np.add.at(removed, np.minimum(idx + del_len[idx], n), -1)
    keep = np.cumsum(removed[:n]) == 0
    # substitutions on kept bases
    subst = (rng.random(n) < sub) & keep
    out = codes.copy()
    out[subst] = (out[subst] + rng.integers(1, 4, int(subst.sum())).astype(np.uint8)) & 3
    # insertions after a base
    ins_ev = (rng.random(n) < ins * scale) & keep
    ins_len = np.where(ins_ev, rng.geometric(geo_p, n), 0)
    reps = keep.astype(np.int64) + ins_len
    pos = np.repeat(np.arange(n), reps)
    read = out[pos]
    # the first copy of each kept base is the base itself, the following copies are random insertions
    first = np.ones(len(read), dtype=bool)
    if len(read) > 1:
        first[1:] = pos[1:] != pos[:-1]
    # positions that were not kept contribute only inserted copies (none: ins_ev requires keep)
    n_ins = int((~first).sum())
    if n_ins:
        read[~first] = rng.integers(0, 4, n_ins).astype(np.uint8)
    return read


def region_config2(region_id: int, n_reads: int = 30, mean_len: int = 8000, sd_len: int = 800) -> Region:
    """BASELINE config 2: 30 ONT-like reads of ~8 kb over one candidate region; two haplotypes, the
    second carries one INS or DEL (size log-uniform 50..2000) in the middle 60 %."""
    rng = np.random.Generator(np.random.PCG64(10_000 + region_id))
    length = int(np.clip(rng.normal(mean_len, sd_len), mean_len // 2, mean_len + mean_len // 2))
    hap0 = rng.integers(0, 4, length).astype(np.uint8)
    size = int(round(np.exp(rng.uniform(np.log(50), np.log(2000)))))
    pos = int(rng.uniform(0.2, 0.8) * length)
    if rng.random() < 0.5:
        hap1 = np.concatenate((hap0[:pos], rng.integers(0, 4, size).astype(np.uint8), hap0[pos:]))
        sv = size
    else:
        size = min(size, length - pos - 100)
        hap1 = np.concatenate((hap0[:pos], hap0[pos + size:]))
        sv = -size
    haps = (hap0, hap1)
    names, reads, strands, offsets, hap_of = [], [], [], [], []
    for k in range(n_reads):
        hp = int(rng.random() < 0.5)
        left, right = int(rng.integers(0, 151)), int(rng.integers(0, 151))
        tmpl = haps[hp][left:len(haps[hp]) - right]
        read = noisy_copy(rng, tmpl)
        strand = "+" if rng.random() < 0.5 else "-"
        if strand == "-":
            read = revcomp_codes(read)
        names.append(f"r{region_id}_{k:02d}")
        reads.append(read)
        strands.append(strand)
        offsets.append(left)
        hap_of.append(hp)
    return Region(region_id, names, reads, strands, offsets, hap_of, sv_size=sv, meta={"sv_pos": pos, "length": length})


def region_config2_shape(region_id: int, mean_len: int = 8000, sd_len: int = 800) -> tuple[int, int]:
    """(template length, signed SV size) of region_config2(region_id) without generating its reads —
    the same first draws of the same generator; used by the scheduler to cost regions cheaply."""
    rng = np.random.Generator(np.random.PCG64(10_000 + region_id))
    length = int(np.clip(rng.normal(mean_len, sd_len), mean_len // 2, mean_len + mean_len // 2))
    rng.integers(0, 4, length)
    size = int(round(np.exp(rng.uniform(np.log(50), np.log(2000)))))
    pos = int(rng.uniform(0.2, 0.8) * length)
    if rng.random() < 0.5:
        return length, size
    return length, -min(size, length - pos - 100)


def region_pairs(region: Region, slack: int = 300, granule: int = 128) -> list[tuple[int, int, int, int, int]]:
    """AVA pair specs (q, t, band_lo, band_w, flags) of one region with local read indices: every
    unordered pair once, query = lexicographically smaller name, strand from the known read
    orientations, band = offset_band() from the known start offsets and the length difference."""
    from .ava import ava_pair_specs
    lens = [len(r) for r in region.reads]
    specs, _ = ava_pair_specs(region.names, lens, dict(zip(region.names, region.strands)),
                              None, slack, granule, False)
    return specs


# ---------------------------------------------------------------------------------------------
# BASELINE config 3: tandem-repeat / MUC1-like regions, long reads, wide bands
# ---------------------------------------------------------------------------------------------

def _log_uniform(rng: np.random.Generator, lo: float, hi: float) -> float:
    return float(np.exp(rng.uniform(np.log(lo), np.log(hi))))


def _tandem_array(rng: np.random.Generator, motif: np.ndarray, length: int, variant_rate: float = 0.05) -> np.ndarray:
    """`length` bases of motif copies, every base replaced by a random other base with probability variant_rate."""
    copies = length // len(motif) + 1
    arr = np.tile(motif, copies)[:length].copy()
    var = rng.random(length) < variant_rate
    arr[var] = (arr[var] + rng.integers(1, 4, int(var.sum())).astype(np.uint8)) & 3
    return arr


def region_config3(region_id: int, n_reads: int = 20, scale: float = 1.0) -> Region:
    """BASELINE config 3: a VNTR with motif length U{20..60} between two 2 kb flanks; haplotype 0 is
    10-30 kb long, haplotype 1 carries 2-20 kb more repeat copies (5 % motif-variant noise everywhere); 20 spanning
    reads, i.e. read lengths 10-50 kb. `scale` < 1 shrinks every length (tests)."""
    rng = np.random.Generator(np.random.PCG64(20_000 + region_id))
    motif = rng.integers(0, 4, int(rng.integers(20, 61))).astype(np.uint8)
    flank = max(64, int(2000 * scale))
    len0 = int(_log_uniform(rng, 10_000, 30_000) * scale)
    delta = int(_log_uniform(rng, 2_000, 20_000) * scale)
    left = rng.integers(0, 4, flank).astype(np.uint8)
    right = rng.integers(0, 4, flank).astype(np.uint8)
    arr0 = _tandem_array(rng, motif, max(len(motif), len0 - 2 * flank))
    extra = _tandem_array(rng, motif, delta)
    cut = int(rng.integers(0, len(arr0) // len(motif) + 1)) * len(motif)       # the extra copies enter on a motif boundary
    cut = min(cut, len(arr0))
    haps = (np.concatenate((left, arr0, right)), np.concatenate((left, arr0[:cut], extra, arr0[cut:], right)))
    names, reads, strands, offsets, hap_of = [], [], [], [], []
    trim_max = max(2, int(150 * scale))
    for k in range(n_reads):
        hp = int(rng.random() < 0.5)
        lo, hi = int(rng.integers(0, trim_max + 1)), int(rng.integers(0, trim_max + 1))
        read = noisy_copy(rng, haps[hp][lo:len(haps[hp]) - hi])
        strand = "+" if rng.random() < 0.5 else "-"
        names.append(f"t{region_id}_{k:02d}")
        reads.append(revcomp_codes(read) if strand == "-" else read)
        strands.append(strand)
        offsets.append(lo)
        hap_of.append(hp)
    return Region(region_id, names, reads, strands, offsets, hap_of, sv_size=delta,
                  meta={"motif_len": len(motif), "hap_lengths": (len(haps[0]), len(haps[1])), "flank": flank})


def region_pairs_config3(region: Region, slack: int = 256, granule: int = 32) -> list[tuple[int, int, int, int, int]]:
    """AVA pair specs of a config-3 region: band = |dlen| + 2*slack (SURVEY: |dlen| + 512) around the co-terminal line."""
    return region_pairs(region, slack, granule)


# ---------------------------------------------------------------------------------------------
# BASELINE configs 4 / 5: HG002-shape candidate set incl. consensus -> reference window; trio with skewed lengths
# ---------------------------------------------------------------------------------------------

@dataclass
class Locus:
    """One candidate region of configs 4/5: its reads, and per haplotype a padded consensus to be aligned back to
    the reference window it came from."""

    region: Region
    chrom: str
    ref_window: np.ndarray            # reference bases of the window (codes)
    window_start: int                 # chromosome coordinate of ref_window[0]
    core_interval: tuple[int, int]    # candidate region on the chromosome
    consensus: list[dict]             # per haplotype: {"id", "padded" (codes), "core" (lo, hi) on padded, "strand"}


def _core_length(rng: np.random.Generator, skew: bool) -> int:
    if skew:                                                     # config 5: Pareto(1.3) tail, max 100 kb
        return int(np.clip(1500.0 * (1.0 + rng.pareto(1.3)), 600, 100_000))
    return int(np.clip(rng.lognormal(np.log(3000.0), 0.8), 600, 50_000))


def locus_config4(region_id: int, seed_base: int = 30_000, skew: bool = False, scale: float = 1.0, max_reads: int = 60) -> Locus:
    """BASELINE config 4: core length log-normal(median 3 kb, sigma 0.8) in [600, 50 000]; coverage
    Poisson(30) in [4, 60]; one or two haplotypes (the second carries an INS or DEL of 50..min(2000, core/3) bp);
    reads are noisy copies of the core +- 300 bp. Per haplotype a padded consensus = pad + core + pad,
    pad = min(2*core, 20 kb), cut from the SAMPLE haplotype (0.5 % diverged from the reference), to be aligned to the
    reference window = locus +- (pad + 10 %). Config 5 = same with seed_base 40/50/60 000 and skew=True."""
    rng = np.random.Generator(np.random.PCG64(seed_base + region_id))
    core = max(200, int(_core_length(rng, skew) * scale))
    pad = min(2 * core, int(20_000 * scale))
    margin = pad + max(100, pad // 10)
    n_reads = int(np.clip(rng.poisson(30), 4, max_reads))
    two_haps = rng.random() < 0.5
    total = core + 2 * margin
    ref = rng.integers(0, 4, total).astype(np.uint8)
    sample = ref.copy()
    div = rng.random(total) < 0.005
    sample[div] = (sample[div] + rng.integers(1, 4, int(div.sum())).astype(np.uint8)) & 3
    c0, c1 = margin, margin + core                       # core on ref / sample hap0 (window coordinates)
    hap0 = sample
    sv = 0
    if two_haps:
        size = int(round(_log_uniform(rng, 50, max(51, min(2000, core // 3)))))
        pos = c0 + int(rng.uniform(0.2, 0.8) * core)
        if rng.random() < 0.5:
            hap1 = np.concatenate((sample[:pos], rng.integers(0, 4, size).astype(np.uint8), sample[pos:]))
            sv = size
        else:
            size = min(size, c1 - pos - 50)
            hap1 = np.concatenate((sample[:pos], sample[pos + size:]))
            sv = -size
    else:
        hap1 = hap0
    haps = (hap0, hap1)
    core_end = (c1, c1 + sv)                              # end of the core on each haplotype
    names, reads, strands, offsets, hap_of = [], [], [], [], []
    flank = min(300, margin)
    for k in range(n_reads):
        hp = int(two_haps and rng.random() < 0.5)
        lo = c0 - int(rng.integers(0, flank + 1))
        hi = core_end[hp] + int(rng.integers(0, flank + 1))
        read = noisy_copy(rng, haps[hp][lo:hi])
        strand = "+" if rng.random() < 0.5 else "-"
        names.append(f"h{region_id}_{k:02d}")
        reads.append(revcomp_codes(read) if strand == "-" else read)
        strands.append(strand)
        offsets.append(lo - (c0 - flank))
        hap_of.append(hp)
    cons = []
    for hp in range(2 if two_haps else 1):
        lo, hi = c0 - pad, core_end[hp] + pad
        padded = haps[hp][lo:hi]
        interval = (pad, pad + (core_end[hp] - c0))
        strand = "+" if rng.random() < 0.5 else "-"
        if strand == "-":
            padded = revcomp_codes(padded)
            interval = (len(padded) - interval[1], len(padded) - interval[0])
        cons.append({"id": f"{region_id}.{hp}", "padded": padded, "core": interval, "strand": strand})
    start = 1_000_000 + 200_000 * (region_id % 1000)
    region = Region(region_id, names, reads, strands, offsets, hap_of, sv_size=sv, meta={"core": core, "pad": pad, "two_haps": two_haps})
    return Locus(region, chrom=f"chr{1 + region_id % 22}", ref_window=ref, window_start=start,
                 core_interval=(start + c0, start + c1), consensus=cons)


def locus_shape(region_id: int, seed_base: int = 30_000, skew: bool = False, scale: float = 1.0, max_reads: int = 60) -> tuple[int, int, int, bool]:
    """(core, pad, n_reads, two_haps) of locus_config4(region_id, ...) without generating its sequences — the same first draws
    of the same generator; what a scheduler knows before it aligns (candidate-region span and coverage)."""
    rng = np.random.Generator(np.random.PCG64(seed_base + region_id))
    core = max(200, int(_core_length(rng, skew) * scale))
    pad = min(2 * core, int(20_000 * scale))
    n_reads = int(np.clip(rng.poisson(30), 4, max_reads))
    two_haps = bool(rng.random() < 0.5)
    return core, pad, n_reads, two_haps


def locus_cost_estimate(region_id: int, seed_base: int = 30_000, skew: bool = False, slack: int = 300, band_w: int = 4096) -> int:
    """Estimated DP cells of a locus from its shape alone: n(n-1)/2 read pairs of ~core + 300 bases in bands of ~2*slack + 200
    diagonals (+ the planted SV between haplotypes), plus both strands of every padded consensus against its window."""
    core, pad, n, two = locus_shape(region_id, seed_base, skew)
    read = core + 300
    pairs = n * (n - 1) // 2
    band = 2 * slack + 200
    if two:                                                       # half the pairs cross haplotypes: + the mean of the log-uniform SV size
        top = max(51, min(2000, core // 3))
        band += int(0.5 * (top - 50) / np.log(top / 50.0))
    cons = (2 if two else 1) * 2 * (core + 2 * pad) * min(band_w, core + 2 * pad)
    return pairs * read * min(band, 2 * read) + cons


def locus_cost(locus: Locus, slack: int = 300, granule: int = 128) -> int:
    """Estimated DP cells of a locus: its all-vs-all plus its consensus -> window pairs (both strands)."""
    from .scheduler import region_cost
    cost = region_cost([len(r) for r in locus.region.reads], slack, granule)
    for c in locus.consensus:
        cost += 2 * len(c["padded"]) * min(4096, len(locus.ref_window))
    return cost
