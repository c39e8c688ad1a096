"""The pieces of svirlpool.localassembly.consensus that consume alignments on the hot path, with the
aligner calls replaced by the CUDA engine (/root/reference/src/svirlpool/localassembly/consensus.py):

    parse_ReadAlignmentSignals_from_alignment   :2440-2464
    pairwise_alignment_lengths                  :1276-1305
    ava_similarity                              :1362-1467   (AVA alignment + consensus_lib pipeline of
                                                              consensus_while_clustering, steps 1-4)
    align_reads_to_consensus                    :851-857 / :2089-2096 (map-ont of reads vs consensus)
    final_consensus                             :821-916
    get_read_alignment_intervals_in_region / _in_cr, get_max_extents_of_read_alignments_on_cr, trim_reads   :382-586

Clustering, consensus assembly (lamassemble/racon) and padding stay with the reference's host code: they
are outside the alignment path.
"""
from __future__ import annotations

import numpy as np

from .crs_to_batches import _load_crIDs_from_batch_tsv  # noqa: F401  (lives in consensus.py in the reference: consensus.py:3134-3160)
from . import alignments_to_rafs, ava, consensus_lib
from .datatypes import AlignedRecord, ReadAlignmentSignals


def parse_ReadAlignmentSignals_from_alignment(samplename: str, alignment, min_signal_size: int, min_bnd_size: int
                                              ) -> ReadAlignmentSignals:
    """Signals of one alignment with read coordinates from get_start_end() (leading clip on forward, trailing
    clip on reverse), sorted by reference position."""
    ref_start, ref_end, read_start, read_end = alignments_to_rafs.get_start_end(alignment)
    signals = alignments_to_rafs.parse_SVsignals_from_alignment(
        alignment=alignment, ref_start=ref_start, ref_end=ref_end, read_start=read_start, read_end=read_end,
        min_signal_size=min_signal_size, min_bnd_size=min_bnd_size)
    return ReadAlignmentSignals(samplename=samplename, read_name=str(alignment.query_name),
                                reference_name=str(alignment.reference_name),
                                alignment_forward=bool(not alignment.is_reverse), SV_signals=sorted(signals))


def pairwise_alignment_lengths(ava_alignments: list[AlignedRecord], read_names: list[str]) -> np.ndarray:
    """Symmetric (n, n) matrix of query_alignment_length per read pair (0 without alignment); when a pair
    has several alignments the last one in input order wins."""
    index = {name: k for k, name in enumerate(read_names)}
    out = np.zeros((len(read_names), len(read_names)), dtype=int)
    for aln in ava_alignments:
        if aln.is_unmapped or aln.query_name is None or aln.reference_name is None:
            continue
        a, b = index.get(aln.query_name), index.get(aln.reference_name)
        if a is None or b is None:
            continue
        out[a, b] = out[b, a] = aln.query_alignment_length or 0
    return out


def _rank_in_cluster(matrix: np.ndarray, matrix_names: list[str], cluster_names: list[str], reduce: str) -> list[str]:
    """Cluster reads ordered by their row statistic inside the cluster's sub-matrix, largest first: the mean over the
    cluster (similarity, diagonal included; consensus.py:583-609) or the sum over the OTHER reads (alignment lengths,
    consensus.py:612-637). Ties fall as numpy's argsort of the negated values leaves them, as in the reference."""
    if len(cluster_names) == 1:
        return list(cluster_names)
    where = {name: k for k, name in enumerate(matrix_names)}
    idx = np.array([where[n] for n in cluster_names])
    sub = np.array(matrix[np.ix_(idx, idx)])
    if reduce == "mean":
        stat = sub.mean(axis=1)
    else:
        np.fill_diagonal(sub, 0)
        stat = sub.sum(axis=1)
    return [cluster_names[k] for k in np.argsort(-stat)]


def find_representative_read(similarity_matrix: np.ndarray, pairwise_alignment_lengths_matrix: np.ndarray,
                             sim_read_names: list[str], cluster_read_names: list[str]) -> str:
    """The read a cluster is polished against in racon mode (consensus.py:640-660): every read of the cluster is ranked by
    its mean similarity to the cluster and by the summed lengths of its alignments to the other reads; the read with the
    smallest rank sum wins (first one on ties). Same signature as the reference."""
    by_sim = _rank_in_cluster(similarity_matrix, sim_read_names, cluster_read_names, "mean")
    by_len = _rank_in_cluster(pairwise_alignment_lengths_matrix, sim_read_names, cluster_read_names, "sum")
    ranksums = np.array([by_sim.index(r) + by_len.index(r) for r in cluster_read_names], dtype=float)
    return cluster_read_names[int(np.argmin(ranksums))]


def ava_similarity(engine, reads: dict[str, str], strands: dict[str, str] | None = None, densities_weight: float = 0.0,
                   min_signal_size: int = 12, min_bnd_size: int = 100, slack: int = 300, granule: int = 512
                   ) -> tuple[np.ndarray, list[str], list[AlignedRecord]]:
    """All-vs-all alignment of one container's reads on the GPU followed by the reference's similarity
    pipeline (signals -> size densities -> importance densities -> directed signals -> similarity matrix),
    i.e. steps 1-4 of consensus_while_clustering. Returns (similarity, read names, alignment records).
    densities_weight defaults to 0.0 like the reference's CLI (consensus.py:3565-3572)."""
    records = ava.ava_align(engine, reads, strands=strands, slack=slack, granule=granule)
    names = list(reads)
    lengths = {n: len(s) for n, s in reads.items()}
    ava_signals = consensus_lib.parse_sv_signals_from_ava_alignments(records, min_signal_size, min_bnd_size)
    size_dens = consensus_lib.sv_size_densities_from_ava_signals(ava_signals)
    dens = consensus_lib.importance_densities_from_ava_signals(ava_signals=ava_signals, size_densities=size_dens)
    directed = consensus_lib.parse_directed_signals_from_ava(records, min_signal_size, min_bnd_size)
    sim, order = consensus_lib.pairwise_similarity_matrix(directed_signals=directed, densities=dens, read_lengths=lengths,
                                                          all_read_names=names, densities_weight=densities_weight)
    return sim, order, records


def ava_similarity_summary(engine, reads: dict[str, str], strands: dict[str, str] | None = None, min_bnd_size: int = 100,
                           slack: int = 300, granule: int = 512) -> tuple[np.ndarray, list[str], np.ndarray]:
    """ava_similarity() for the production setting (unit density weights) without CIGARs: the per-pair distance is reduced
    on the GPU (svp_result.sv_dist), only result rows cross PCIe, and the matrices are assembled with numpy — bit-identical
    to ava_similarity(..., densities_weight=0.0) and pairwise_alignment_lengths(). Returns (similarity, read names,
    alignment-length matrix)."""
    names, qi, ti, res = ava.ava_summary(engine, reads, strands=strands, slack=slack, granule=granule)
    sim, aln_len = consensus_lib.similarity_from_summary(names, {n: len(s) for n, s in reads.items()}, qi, ti, res,
                                                         min_bnd_size=min_bnd_size)
    return sim, names, aln_len


def ava_similarity_summary_many(engine, containers: list[dict[str, str]], strands: list[dict[str, str] | None] | None = None,
                                min_bnd_size: int = 100, slack: int = 300, granule: int = 512
                                ) -> list[tuple[np.ndarray, list[str], np.ndarray]]:
    """ava_similarity_summary() for a whole batch of containers in ONE engine call (what a consensus job over a crs_to_batches
    batch does on the GPU): per container (similarity matrix, read names, alignment-length matrix)."""
    out = []
    for reads, (names, qi, ti, res) in zip(containers, ava.ava_summary_many(engine, containers, strands, slack=slack, granule=granule)):
        sim, aln_len = consensus_lib.similarity_from_summary(names, {n: len(s) for n, s in reads.items()}, qi, ti, res,
                                                             min_bnd_size=min_bnd_size)
        out.append((sim, names, aln_len))
    return out


def align_reads_to_consensus(engine, samplename: str, reads: dict[str, str], consensus_name: str, consensus_seq: str,
                             min_signal_size: int = 8, min_bnd_size: int = 100
                             ) -> tuple[list[AlignedRecord], list[ReadAlignmentSignals]]:
    """Reads of a cluster vs their consensus (`-x map-ont --secondary=no`, consensus.py:851-857) and the
    cut_read_alignment_signals derived from the records (min indel 8 / BND 100, :1616-1617)."""
    records = ava.map_queries(engine, reads, consensus_name, consensus_seq)
    signals = [parse_ReadAlignmentSignals_from_alignment(samplename, r, min_signal_size, min_bnd_size) for r in records]
    return records, signals


def final_consensus(engine, samplename: str, min_indel_size: int, min_bnd_size: int, reads: dict[str, str], consensus_sequence: str,
                    ID: str, crIDs: list[int], original_regions: list[tuple[str, int, int]]):
    """The reference's final_consensus (consensus.py:821-916) with the `minimap2 -x map-ont -Y --sam-hit-only
    --secondary=no` subprocess replaced by the CUDA engine: the cluster's cut reads are aligned to the finished
    consensus; their SV signals (the consensus' "distortions") and reference intervals go into the Consensus record.
    Returns None when no read aligns, as the reference does. In-memory inputs instead of FASTA paths."""
    from . import consensus_class
    if not consensus_sequence:
        raise ValueError("consensus sequence is empty. Cannot proceed.")
    if not reads:
        raise ValueError("reads are empty. Cannot proceed.")
    records, signals = align_reads_to_consensus(engine, samplename, reads, ID, consensus_sequence,
                                                min_signal_size=min_indel_size, min_bnd_size=min_bnd_size)
    if not records:
        return None
    intervals = [(r.reference_start, r.reference_end, r.query_name, r.is_forward) for r in records]
    return consensus_class.Consensus(ID=ID, crIDs=list(crIDs), original_regions=list(original_regions), consensus_sequence=consensus_sequence,
                                     cut_read_alignment_signals=signals, intervals_cutread_alignments=intervals)


def add_cutread_alignments_to_consensus_inplace(alns: list[AlignedRecord], consensus, min_indel_size: int, min_bnd_size: int,
                                                samplename: str) -> None:
    """Append the SV signals and reference intervals of further read alignments to a Consensus record
    (consensus.py:919-943; same signature)."""
    consensus.cut_read_alignment_signals.extend(
        parse_ReadAlignmentSignals_from_alignment(samplename, aln, min_indel_size, min_bnd_size) for aln in alns)
    consensus.intervals_cutread_alignments.extend((aln.reference_start, aln.reference_end, aln.query_name, aln.is_forward) for aln in alns)


def add_unaligned_reads_to_consensuses_inplace(engine, samplename: str, consensus_objects: dict, pool: dict[str, str],
                                               min_indel_size: int, min_bnd_size: int) -> dict[str, str]:
    """Re-home the reads no consensus used (consensus.py:2050-2131): every read of `pool` is aligned to EVERY consensus
    sequence (the reference: one `minimap2 -x map-ont --secondary=no --sam-hit-only -U 25,75 -H` run of the reads against a
    FASTA of all consensuses) and given to the consensus of its primary, i.e. best-scoring, alignment; that consensus receives
    the read's alignment signals and intervals. One engine call per round for all reads x consensuses, both strands
    (ava.map_jobs). Reads without a hit stay unassigned. Returns {read name: consensus ID} (the reference returns None and
    works through the side effect only; the side effect is the same). `pool`: read name -> sequence."""
    if len(consensus_objects) == 0:
        raise AssertionError("consensus_sequences must contain at least one consensus object")
    if len(pool) == 0:
        return {}
    records, redistribution = ava.map_queries_to_best_target(
        engine, pool, {cid: cons.consensus_sequence for cid, cons in consensus_objects.items()})
    for cid, cons in consensus_objects.items():
        assigned = [rec for rec in records if rec.reference_name == cid]
        if assigned:
            add_cutread_alignments_to_consensus_inplace(assigned, cons, min_indel_size, min_bnd_size, samplename)
    return redistribution


# ---------------------------------------------------------------------------------------------
# read cutting: the reads of a container are trimmed to its candidate regions before the all-vs-all
# (consensus.py:382-586); the cut reads are what BASELINE config 2's "30 reads x 8 kb" are
# ---------------------------------------------------------------------------------------------

def get_read_alignment_intervals_in_region(region_start: int, regions_end: int, alignments: list, buffer_clipped_length: int
                                           ) -> dict[str, list[tuple[int, int, str, int, int]]]:
    """readname -> [(read_start, read_end, ref_chr, ref_start, ref_end)] of every alignment inside the region
    (consensus.py:382-418): the read interval covering the region (reaching up to buffer_clipped_length into a clip when
    the region extends beyond the alignment) and the reference interval that read interval maps back to."""
    from . import util
    out: dict[str, list[tuple[int, int, str, int, int]]] = {}
    for aln in alignments:
        read_start, read_end = util.get_interval_on_read_in_region(a=aln, start=region_start, end=regions_end,
                                                                   buffer_clipped_length=buffer_clipped_length)
        ref_start, ref_end = util.get_interval_on_ref_in_region(a=aln, start=read_start, end=read_end)
        if read_start is not None and read_end is not None and read_end > read_start:
            out.setdefault(aln.query_name, []).append((read_start, read_end, aln.reference_name, ref_start, ref_end))
    return out


def get_read_alignment_intervals_in_cr(crs: list, buffer_clipped_length: int, dict_alignments: dict[int, list]
                                       ) -> dict[str, list[tuple[int, int, str, int, int]]]:
    """Intervals over all candidate regions of a container (consensus.py:422-465). `crs` need .crID, .referenceStart,
    .referenceEnd and .sv_signals; the clip buffer is at least the largest insertion signal of the container."""
    extents = {cr.crID: (cr.referenceStart, cr.referenceEnd) for cr in crs}
    max_ins = max([sv.size for cr in crs for sv in cr.sv_signals if sv.sv_type == 0], default=0)
    used = max(buffer_clipped_length, max_ins)
    out: dict[str, list[tuple[int, int, str, int, int]]] = {}
    for crID, alns in dict_alignments.items():
        for name, ivs in get_read_alignment_intervals_in_region(extents[crID][0], extents[crID][1], alns, used).items():
            out.setdefault(name, []).extend(ivs)
    return out


def get_max_extents_of_read_alignments_on_cr(dict_all_intervals: dict[str, list[tuple[int, int, str, int, int]]]
                                             ) -> dict[str, tuple[int, int, str, int, str, int]]:
    """readname -> (min read_start, max read_end, chr of min ref_start, min ref_start, chr of max ref_end, max ref_end)
    (consensus.py:468-496)."""
    out = {}
    for name, ivs in dict_all_intervals.items():
        lo = min((ref_start, ref_chr) for _s, _e, ref_chr, ref_start, _re in ivs)
        hi = max((ref_end, ref_chr) for _s, _e, ref_chr, _rs, ref_end in ivs)
        out[name] = (min(s for s, *_ in ivs), max(e for _s, e, *_ in ivs), lo[1], lo[0], hi[1], hi[0])
    return out


def trim_reads(dict_alignments: dict[int, list], intervals: dict[str, tuple[int, int, str, int, str, int]], read_records: dict[str, str]
               ) -> dict[str, tuple[str, str]]:
    """Cut every read to its interval (consensus.py:499-586). read_records: readname -> full sequence in ORIGINAL read
    orientation (read_cache.ReadSequenceCache). Returns readname -> (cut sequence, description) with the reference's
    description string (crID, cut positions on the read and on the reference)."""
    cut: dict[str, tuple[str, str]] = {}
    for crID, alns in dict_alignments.items():
        for aln in alns:
            name = aln.query_name
            if name not in intervals or name in cut or name not in read_records:
                continue
            start, end, ref_start_chr, ref_start, ref_end_chr, ref_end = intervals[name]
            first = (ref_start_chr, ref_start) if ref_start < ref_end else (ref_end_chr, ref_end)
            last = (ref_start_chr, ref_start) if ref_start >= ref_end else (ref_end_chr, ref_end)
            desc = (f"crID={crID},start={start},end={end},ref_start_chr={first[0]},ref_start={first[1]},"
                    f"ref_end_chr={last[0]},ref_end={last[1]}")
            cut[name] = (read_records[name][start:end], desc)
    return cut
