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


def region_pairs(region: Region, slack: int = 300, granule: int = 512) -> list[tuple[int, int, int, int, int]]:
    """AVA pair specs (q, t, band_lo, band_w, flags) of one region with local read indices: every
    unordered pair once, query = lexicographically smaller name, strand from the known read
    orientations, band = offset_band() from the known start offsets and the length difference."""
    from .ava import ava_pair_specs
    lens = [len(r) for r in region.reads]
    specs, _ = ava_pair_specs(region.names, lens, dict(zip(region.names, region.strands)),
                              None, slack, granule, False)
    return specs
