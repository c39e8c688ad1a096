"""Output records of the consensus stage and the core-interval trace-back of consensus_align — host mirror of
svirlpool.localassembly.consensus_class (/root/reference/src/svirlpool/localassembly/consensus_class.py):

    ConsensusPadding, ConsensusDistortion, Consensus, CrsContainerResult     :64-146   (FORMAT FROZEN: one JSON line per
                                                                                        container, consensus.py:3295-3300)
    get_consensus_core_alignment_interval_on_reference                       :279-303

The reference (de)serialises with cattrs; here `unstructure()` / `structure()` produce and accept exactly the
same plain dict / JSON shapes (tuples become lists in JSON).
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from pathlib import Path
from typing import Iterator

from . import util
from .datatypes import ReadAlignmentSignals, SequenceObject, SVsignal


@dataclass
class ConsensusPadding:
    sequence: str
    readname_left: str
    readname_right: str
    padding_size_left: int       # core interval start
    padding_size_right: int      # len(sequence) - core interval end
    consensus_interval_on_sequence_with_padding: tuple[int, int]

    def unstructure(self) -> dict:
        return {"sequence": self.sequence, "readname_left": self.readname_left, "readname_right": self.readname_right,
                "padding_size_left": self.padding_size_left, "padding_size_right": self.padding_size_right,
                "consensus_interval_on_sequence_with_padding": list(self.consensus_interval_on_sequence_with_padding)}

    @classmethod
    def structure(cls, d: dict) -> "ConsensusPadding":
        return cls(sequence=d["sequence"], readname_left=d["readname_left"], readname_right=d["readname_right"],
                   padding_size_left=int(d["padding_size_left"]), padding_size_right=int(d["padding_size_right"]),
                   consensus_interval_on_sequence_with_padding=tuple(int(v) for v in d["consensus_interval_on_sequence_with_padding"]))


@dataclass
class ConsensusDistortion:
    position: int
    size: int
    type: int        # 0 insertion, 1 deletion
    readname: str
    forward: bool


def _structure_ras(d: dict) -> ReadAlignmentSignals:
    return ReadAlignmentSignals(samplename=d["samplename"], read_name=d["read_name"], reference_name=d["reference_name"],
                                alignment_forward=bool(d["alignment_forward"]),
                                SV_signals=[SVsignal(**{k: int(s[k]) for k in ("ref_start", "ref_end", "read_start", "read_end", "size", "sv_type")})
                                            for s in d.get("SV_signals", [])])


@dataclass
class Consensus:
    ID: str
    crIDs: list
    original_regions: list                       # (chr, start, end)
    consensus_sequence: str
    consensus_padding: ConsensusPadding | None = None
    intervals_cutread_alignments: list = field(default_factory=list)      # (start, end, readname, forward)
    cut_read_alignment_signals: list = field(default_factory=list)        # ReadAlignmentSignals
    clustering_meta_data: dict = field(default_factory=dict)

    def unstructure(self) -> dict:
        return {"ID": self.ID, "crIDs": list(self.crIDs), "original_regions": [list(r) for r in self.original_regions],
                "consensus_sequence": self.consensus_sequence,
                "consensus_padding": self.consensus_padding.unstructure() if self.consensus_padding is not None else None,
                "intervals_cutread_alignments": [list(i) for i in self.intervals_cutread_alignments],
                "cut_read_alignment_signals": [r.unstructure() for r in self.cut_read_alignment_signals],
                "clustering_meta_data": dict(self.clustering_meta_data)}

    @classmethod
    def structure(cls, d: dict) -> "Consensus":
        pad = d.get("consensus_padding")
        return cls(ID=str(d["ID"]), crIDs=[int(c) for c in d["crIDs"]],
                   original_regions=[(str(r[0]), int(r[1]), int(r[2])) for r in d["original_regions"]],
                   consensus_sequence=d["consensus_sequence"],
                   consensus_padding=ConsensusPadding.structure(pad) if pad is not None else None,
                   intervals_cutread_alignments=[(int(i[0]), int(i[1]), str(i[2]), bool(i[3])) for i in d.get("intervals_cutread_alignments", [])],
                   cut_read_alignment_signals=[_structure_ras(r) for r in d.get("cut_read_alignment_signals", [])],
                   clustering_meta_data=dict(d.get("clustering_meta_data", {})))

    def get_used_readnames(self) -> set:
        return {readname for _start, _end, readname, _forward in self.intervals_cutread_alignments}

    def get_consensus_distortions(self) -> list:
        return [ConsensusDistortion(position=s.ref_start, size=s.size, type=s.sv_type, readname=ras.read_name, forward=ras.alignment_forward)
                for ras in self.cut_read_alignment_signals for s in ras.SV_signals if s.sv_type in (0, 1)]

    def get_max_cutread_coverage(self) -> int:
        events = sorted([(s, 1) for s, _e, _n, _f in self.intervals_cutread_alignments] +
                        [(e, -1) for _s, e, _n, _f in self.intervals_cutread_alignments])
        depth = best = 0
        for _pos, ev in events:
            depth += ev
            best = max(best, depth)
        return best


@dataclass
class CrsContainerResult:
    consensus_dicts: dict                        # consensus ID -> Consensus
    unused_reads: dict                           # read ID -> [SequenceObject]

    def unstructure(self) -> dict:
        return {"consensus_dicts": {k: c.unstructure() for k, c in self.consensus_dicts.items()},
                "unused_reads": {k: [s.unstructure() for s in v] for k, v in self.unused_reads.items()}}

    @classmethod
    def structure(cls, d: dict) -> "CrsContainerResult":
        return cls(consensus_dicts={str(k): Consensus.structure(c) for k, c in d["consensus_dicts"].items()},
                   unused_reads={int(k): [SequenceObject(**s) for s in v] for k, v in d.get("unused_reads", {}).items()})


def write_crs_container_results(path: Path | str, results) -> None:
    """One JSON line per container (consensus.py:3295-3300)."""
    with open(path, "w") as f:
        for r in results:
            print(json.dumps(r.unstructure()), file=f)


def parse_crs_container_results(path: Path | str) -> Iterator[CrsContainerResult]:
    """consensus_align.py:58-66."""
    with open(path) as f:
        for line in f:
            if line.strip():
                yield CrsContainerResult.structure(json.loads(line))


def get_consensus_core_alignment_interval_on_reference(consensus: Consensus, alignment) -> tuple[str, int, int]:
    """Reference interval the UPPER-case core of a padded consensus aligns to, traced through the CIGAR of one of its
    alignments (consensus_class.py:279-303)."""
    if consensus.consensus_padding is None:
        raise ValueError("Consensus padding is None, cannot determine core interval.")
    core_start, core_end = consensus.consensus_padding.consensus_interval_on_sequence_with_padding
    a, b = util.get_interval_on_ref_in_region(a=alignment, start=core_start, end=core_end)
    return (str(alignment.reference_name), min(a, b), max(a, b))
