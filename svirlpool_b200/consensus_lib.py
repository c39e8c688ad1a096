"""Signals -> densities -> pairwise similarity: host mirror of the compute half of
svirlpool.localassembly.consensus_lib (/root/reference/src/svirlpool/localassembly/consensus_lib.py).

Same names, arguments and return types as the reference for everything on the hot path:
    size_damping                               :24-37
    parse_sv_signals_from_ava_alignments       :40-90
    importance_densities_from_ava_signals      :93-180
    sv_size_densities_from_ava_signals         :183-261   (figure output is out of scope)
    DirectedSignals                            :269-276
    parse_directed_signals_from_ava            :279-345
    _directed_distance / _has_interior_bnd     :353-413
    pairwise_similarity_matrix                 :416-592
    detect_outlier_reads (+ the two helpers)   :1004-1135
The two parse_* functions accept, besides a SAM/BAM path as in the reference, an in-memory
iterable of alignment records (datatypes.AlignedRecord) — the S2 seam of SURVEY.md §8b — so the
GPU aligner's output never has to touch a file. The matplotlib/gravis plotting functions of the
reference module are not part of the path and are not provided.
"""
from __future__ import annotations

import logging
import math
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Iterable, Tuple

import numpy as np

from . import alignments_to_rafs
from .datatypes import SVsignal

log = logging.getLogger(__name__)


def size_damping(size: float, s0: float = 30.0, alpha: float = 1.5) -> float:
    """f(s) = 1 - exp(-(s/s0)**alpha), in [0, 1): small indels are damped towards 0."""
    return 1.0 - np.exp(-((size / s0) ** alpha))


# ---------------------------------------------------------------------------------------------
# alignment input: path (SAM text or BAM) or records
# ---------------------------------------------------------------------------------------------

def _records(ava_alignments) -> Iterable:
    if isinstance(ava_alignments, (str, Path)):
        from .sam import read_alignments
        return read_alignments(ava_alignments)
    return ava_alignments


def _mapped(ava_alignments):
    for aln in _records(ava_alignments):
        if aln.is_unmapped or aln.reference_name is None or aln.cigartuples is None:
            continue
        yield aln


def _signals_of(aln, min_signal_size: int, min_bnd_size: int) -> list[SVsignal]:
    ref_start, ref_end = aln.reference_start, aln.reference_end
    read_start, read_end = aln.query_alignment_start, aln.query_alignment_end
    if ref_start is None or ref_end is None or read_start is None or read_end is None:
        raise ValueError(
            "parse_directed_signals_from_ava::Alignment missing required positions: "
            f"ref_start={ref_start}, ref_end={ref_end}, read_start={read_start}, read_end={read_end}")
    return alignments_to_rafs.parse_SVsignals_from_alignment(
        alignment=aln, ref_start=ref_start, ref_end=ref_end, read_start=read_start, read_end=read_end,
        min_signal_size=min_signal_size, min_bnd_size=min_bnd_size)


def parse_sv_signals_from_ava_alignments(ava_alignments, min_signal_size: int, min_bnd_size: int
                                         ) -> Dict[str, list[SVsignal]]:
    """{reference read name: all SVsignals seen on it over every query aligned to it}."""
    pooled: Dict[str, list[SVsignal]] = {}
    for aln in _mapped(ava_alignments):
        sig = _signals_of(aln, min_signal_size, min_bnd_size)
        if sig:
            pooled.setdefault(aln.reference_name, []).extend(sig)
    return pooled


def _plateau(lo: int, hi: int, left_edge: float, right_edge: float, falloff: float) -> np.ndarray:
    """Product of a rising sigmoid at left_edge and a falling one at right_edge on [lo, hi)."""
    x = np.arange(lo, hi, dtype=np.float64)
    rise = 1.0 / (1.0 + np.exp(-falloff * (x - left_edge)))
    fall = 1.0 / (1.0 + np.exp(falloff * (x - right_edge)))
    return rise * fall


def importance_densities_from_ava_signals(
    ava_signals: Dict[str, list[SVsignal]],
    size_densities: Dict[int, Dict[str, np.ndarray]] | None = None,
    window_size: int = 100,
    falloff: float = 0.5,
    damping_s0: float = 30.0,
    damping_alpha: float = 1.5,
    density_power: float = 1.5,
) -> Dict[str, np.ndarray]:
    """Position-space importance tracks, one per reference read; track length is the largest
    ref_end among that read's pooled signals. Each INS/DEL adds a plateau of half-width
    window_size/2 around its reference centre, scaled by size_damping(|size|) and, when
    size_densities is given, by (normalised size density at |size|) ** density_power."""
    tracks: Dict[str, np.ndarray] = {}
    half = window_size / 2.0
    reach = 6.0 / falloff
    for name, signals in ava_signals.items():
        if not signals:
            continue
        length = max(s.ref_end for s in signals)
        if length == 0:
            continue
        scale_tab: Dict[int, np.ndarray] = {}
        if size_densities is not None:
            for sv_type in (0, 1):
                arr = size_densities[sv_type].get(name)
                if arr is not None and arr.max() > 0:
                    scale_tab[sv_type] = (arr / arr.max()) ** density_power
        track = np.zeros(length, dtype=np.float64)
        for s in signals:
            if s.sv_type in (3, 4):
                continue
            centre = (s.ref_start + s.ref_end) / 2.0
            lo = int(max(0, centre - half - reach))
            hi = int(min(length, centre + half + reach + 1))
            if lo >= hi:
                continue
            weight = size_damping(abs(s.size), s0=damping_s0, alpha=damping_alpha)
            tab = scale_tab.get(s.sv_type) if s.sv_type in (0, 1) else None
            if tab is not None and abs(s.size) < len(tab):
                weight = weight * float(tab[abs(s.size)])
            track[lo:hi] += weight * _plateau(lo, hi, centre - half, centre + half, falloff)
        tracks[name] = track
    return tracks


def sv_size_densities_from_ava_signals(
    ava_signals: Dict[str, list[SVsignal]],
    falloff: float = 0.5,
    figures_dir: Path | None = None,
) -> Dict[int, Dict[str, np.ndarray]]:
    """{sv_type (0 INS, 1 DEL): {reference read: density indexed by SV size}}; every signal adds a
    plateau spanning 0.8..1.2 x its size. figures_dir is accepted for signature compatibility; the
    reference's PNG output is not part of the hot path and is not produced."""
    out: Dict[int, Dict[str, np.ndarray]] = {0: {}, 1: {}}
    reach = 6.0 / falloff
    for name, signals in ava_signals.items():
        for sv_type in (0, 1):
            sizes = sorted(abs(s.size) for s in signals if s.sv_type == sv_type)
            if not sizes or sizes[-1] == 0:
                continue
            top = sizes[-1]
            dens = np.zeros(top + 1, dtype=np.float64)
            for size in sizes:
                centre = float(size)
                half = 0.2 * centre
                lo = int(max(0, centre - half - reach))
                hi = int(min(top + 1, centre + half + reach + 1))
                if lo < hi:
                    dens[lo:hi] += _plateau(lo, hi, centre - half, centre + half, falloff)
            out[sv_type][name] = dens
    if figures_dir is not None:
        log.info("sv_size_densities_from_ava_signals: figures_dir ignored (plotting is out of scope)")
    return out


# ---------------------------------------------------------------------------------------------
# directed signals
# ---------------------------------------------------------------------------------------------

@dataclass(frozen=True)
class DirectedSignals:
    """SV signals of one alignment, query -> reference. ref_length is the ALIGNED reference span
    (pysam reference_length, consensus_lib.py:323-325), not the target read's length."""

    query_name: str
    ref_name: str
    signals: list[SVsignal]
    ref_length: int


def parse_directed_signals_from_ava(ava_alignments, min_signal_size: int, min_bnd_size: int
                                    ) -> list[DirectedSignals]:
    """One DirectedSignals per mapped alignment, in input order."""
    out: list[DirectedSignals] = []
    for aln in _mapped(ava_alignments):
        if aln.query_name is None:
            raise ValueError("parse_directed_signals_from_ava::Alignment query name is missing")
        out.append(DirectedSignals(
            query_name=aln.query_name,
            ref_name=aln.reference_name,
            signals=_signals_of(aln, min_signal_size, min_bnd_size),
            ref_length=aln.reference_length if aln.reference_length else 0,
        ))
    return out


# ---------------------------------------------------------------------------------------------
# pairwise similarity
# ---------------------------------------------------------------------------------------------

def _density_weight(track: np.ndarray | None, pos: int, densities_weight: float) -> float:
    if track is None or not 0 <= pos < len(track):
        return 1.0
    return 1.0 + densities_weight * (float(track[pos]) - 1.0)


def _directed_distance(signals: list[SVsignal], ref_density: np.ndarray | None, query_density: np.ndarray | None,
                       damping_s0: float, damping_alpha: float, densities_weight: float = 1.0) -> float:
    """sum over INS/DEL signals of w_ref * w_query * size_damping(|size|) * |size|, with
    w = 1 + densities_weight * (density[centre] - 1) and integer centres (start+end)//2;
    BNDs are skipped."""
    total = 0.0
    for s in signals:
        if s.sv_type in (3, 4):
            continue
        size = abs(s.size)
        w_ref = _density_weight(ref_density, (s.ref_start + s.ref_end) // 2, densities_weight)
        w_query = _density_weight(query_density, (s.read_start + s.read_end) // 2, densities_weight)
        total += w_ref * w_query * size_damping(size, s0=damping_s0, alpha=damping_alpha) * size
    return total


def _has_interior_bnd(signals: list[SVsignal], ref_length: int, bnd_margin: int = 100) -> bool:
    """True if some BND sits strictly more than bnd_margin from both ends of [0, ref_length)."""
    return any(s.sv_type in (3, 4) and bnd_margin < s.ref_start < ref_length - bnd_margin for s in signals)


def pairwise_similarity_matrix(
    directed_signals: list[DirectedSignals],
    densities: Dict[str, np.ndarray],
    read_lengths: Dict[str, int],
    all_read_names: list[str] | None = None,
    length_penalty_lambda: float | None = None,
    damping_s0: float = 30.0,
    damping_alpha: float = 1.5,
    bnd_margin: int = 100,
    densities_weight: float = 1.0,
) -> Tuple[np.ndarray, list[str]]:
    """Symmetric similarity matrix over reads.

    D(A,B) = max(mean d(A->B), mean d(B->A)) + lambda * (1 - min(len)/max(len));
    sim = exp(-D^2 / (2 sigma^2)), sigma = median of the finite non-zero signal-only distances
    (1.0 if there are none), lambda defaults to sigma. Pairs without any alignment, or with an
    interior BND in either direction, get similarity 0; the diagonal is 1.
    """
    if all_read_names is not None:
        read_names = list(all_read_names)
    else:
        seen = set(read_lengths)
        for ds in directed_signals:
            seen.update((ds.query_name, ds.ref_name))
        read_names = sorted(seen)
    n = len(read_names)
    index = {name: k for k, name in enumerate(read_names)}

    d_sum = np.zeros((n, n), dtype=np.float64)
    d_count = np.zeros((n, n), dtype=np.int64)
    interior_bnd = np.zeros((n, n), dtype=bool)
    for ds in directed_signals:
        qi, ri = index.get(ds.query_name), index.get(ds.ref_name)
        if qi is None or ri is None:
            continue
        d_sum[qi, ri] += _directed_distance(ds.signals, densities.get(ds.ref_name), densities.get(ds.query_name),
                                            damping_s0=damping_s0, damping_alpha=damping_alpha,
                                            densities_weight=densities_weight)
        d_count[qi, ri] += 1
        if _has_interior_bnd(ds.signals, ds.ref_length, bnd_margin=bnd_margin):
            interior_bnd[qi, ri] = True

    return _similarity_from_directed(d_sum, d_count, interior_bnd, read_names, read_lengths, length_penalty_lambda), read_names


def _similarity_from_directed(d_sum: np.ndarray, d_count: np.ndarray, interior_bnd: np.ndarray, read_names: list[str],
                              read_lengths: Dict[str, int], length_penalty_lambda: float | None) -> np.ndarray:
    """Directed distance sums / counts / interior-BND flags -> similarity matrix (second half of
    pairwise_similarity_matrix, consensus_lib.py:497-592)."""
    n = len(read_names)
    aligned = d_count > 0
    directed = np.where(aligned, d_sum / np.maximum(d_count, 1), 0.0)
    linked = aligned | aligned.T
    off_diag = ~np.eye(n, dtype=bool)
    signal_dist = np.where(linked & off_diag, np.maximum(directed, directed.T), np.inf)
    np.fill_diagonal(signal_dist, 0.0)

    calib = signal_dist[np.isfinite(signal_dist) & (signal_dist > 0)]
    sigma = float(np.median(calib)) if calib.size else 1.0
    if sigma == 0:
        sigma = 1.0
    if length_penalty_lambda is None:
        length_penalty_lambda = sigma

    lengths = np.array([float(read_lengths.get(r, 0)) for r in read_names], dtype=np.float64)
    lo = np.minimum.outer(lengths, lengths)
    hi = np.maximum.outer(lengths, lengths)
    with np.errstate(divide="ignore", invalid="ignore"):
        size_sim = np.where(hi > 0, lo / np.where(hi > 0, hi, 1.0), 0.0)
    distance = signal_dist + np.where(off_diag, length_penalty_lambda * (1.0 - size_sim), 0.0)

    with np.errstate(over="ignore"):
        similarity = np.where(np.isfinite(distance), np.exp(-(distance ** 2) / (2.0 * sigma ** 2)), 0.0)
    np.fill_diagonal(similarity, 1.0)
    similarity[(interior_bnd | interior_bnd.T) & off_diag] = 0.0
    return similarity


def similarity_from_summary(read_names: list[str], read_lengths: Dict[str, int], q_index: np.ndarray, t_index: np.ndarray,
                            results: np.ndarray, min_score: int = 80, min_bnd_size: int = 100, bnd_margin: int = 100,
                            length_penalty_lambda: float | None = None) -> Tuple[np.ndarray, np.ndarray]:
    """pairwise_similarity_matrix() and consensus.pairwise_alignment_lengths() for unit density weights (the reference's
    production setting --densities-weight 0.0, consensus.py:3565-3572), computed from svp_result rows without CIGARs:

      * the directed distance of an alignment = sum over indel runs >= min_signal_size of size_damping(size) * size, which the
        traceback kernel accumulates per pair (svp_result.sv_dist; damping s0 = 30, alpha = 1.5 = the defaults here);
      * its breakends: a clip >= min_bnd_size before / after the aligned part of the oriented query gives a BND at the
        alignment's start / end on the target, and `_has_interior_bnd` tests it against the ALIGNED span
        (`bnd_margin < pos < span - bnd_margin`, the reference's convention) — only a leading clip can satisfy it.

    q_index / t_index: read index (into read_names) of every result row's query / target; several rows per read pair
    (both orientations) are reduced to the best-scoring one, forward first on ties, like ava.ava_records(). Returns
    (similarity, alignment-length matrix). Bit-identical to the record-based functions (tests/test_consensus_lib.py)."""
    n = len(read_names)
    res = results
    ok = ((res["status"] & 1) == 0) & (res["score"] >= min_score) & (res["cigar_len"] > 0)
    d_sum = np.zeros((n, n), dtype=np.float64)
    d_count = np.zeros((n, n), dtype=np.int64)
    interior = np.zeros((n, n), dtype=bool)
    aln_len = np.zeros((n, n), dtype=int)
    rows = np.flatnonzero(ok)
    if len(rows):
        # best row per (q, t): highest score, earliest row on ties
        key = q_index[rows].astype(np.int64) * n + t_index[rows].astype(np.int64)
        order = np.lexsort((rows, -res["score"][rows].astype(np.int64), key))
        first = np.ones(len(order), dtype=bool)
        first[1:] = key[order][1:] != key[order][:-1]
        best = rows[order[first]]
        q, t = q_index[best], t_index[best]
        r = res[best]
        d_sum[q, t] = r["sv_dist"]
        d_count[q, t] = 1
        lead = r["q_beg"].astype(np.int64)
        span = (r["t_end"] - r["t_beg"]).astype(np.int64)
        t_beg = r["t_beg"].astype(np.int64)
        interior[q, t] = (lead >= min_bnd_size) & (t_beg > bnd_margin) & (t_beg < span - bnd_margin)
        qlen = (r["q_end"] - r["q_beg"]).astype(int)
        aln_len[q, t] = qlen
        aln_len[t, q] = qlen
    sim = _similarity_from_directed(d_sum, d_count, interior, list(read_names), read_lengths, length_penalty_lambda)
    return sim, aln_len


# ---------------------------------------------------------------------------------------------
# outlier detection
# ---------------------------------------------------------------------------------------------

def _outlier_reads_by_length(read_names: list[str], read_lengths: Dict[str, int], n_clusters: int,
                             length_factor: float = 1.2) -> set[str]:
    """Reads in length-clusters too small to be an allele. Clusters: stable sort by length, split
    where curr/prev > length_factor (or prev == 0 < curr); a cluster is too small below
    max(2, int(sqrt(n / max(n_clusters, 1))))."""
    n = len(read_names)
    if n == 0:
        return set()
    ordered = sorted(read_names, key=lambda r: read_lengths.get(r, 0))
    groups: list[list[str]] = [[ordered[0]]]
    for prev, curr in zip(ordered, ordered[1:]):
        a, b = read_lengths.get(prev, 0), read_lengths.get(curr, 0)
        if (a > 0 and b / a > length_factor) or (a == 0 and b > 0):
            groups.append([])
        groups[-1].append(curr)
    min_size = max(2, int(np.sqrt(n / max(n_clusters, 1))))
    return {r for g in groups if len(g) < min_size for r in g}


def _outlier_reads_by_similarity(similarity: np.ndarray, read_names: list[str], n_clusters: int,
                                 min_connections: int | None = None,
                                 connectivity_threshold: float = 0.1) -> set[str]:
    """Reads with fewer than min_connections neighbours above connectivity_threshold. When
    min_connections is None it adapts: max(1, min(int(0.5*ceil(n/k)), int(0.5*median(conn))))."""
    n = len(read_names)
    if n == 0:
        return set()
    above = np.asarray(similarity)[:n, :n] > connectivity_threshold
    if min_connections is None:
        cluster_based = max(1, int(0.5 * int(math.ceil(n / max(n_clusters, 1)))))
        conn_incl_self = above.sum(axis=1).astype(np.int64) - 1
        median_based = max(1, int(0.5 * float(np.median(conn_incl_self))))
        min_connections = max(1, min(cluster_based, median_based))
    neighbours = (above & ~np.eye(n, dtype=bool)).sum(axis=1)
    return {read_names[k] for k in range(n) if neighbours[k] < min_connections}


def detect_outlier_reads(similarity: np.ndarray, read_names: list[str], read_lengths: Dict[str, int],
                         n_clusters: int, length_factor: float = 1.2, min_connections: int | None = None,
                         connectivity_threshold: float = 0.1) -> set[str]:
    """Union of the length outliers and the connectivity outliers."""
    if len(read_names) == 0:
        return set()
    return (_outlier_reads_by_length(read_names, read_lengths, n_clusters, length_factor)
            | _outlier_reads_by_similarity(similarity, read_names, n_clusters, min_connections, connectivity_threshold))
