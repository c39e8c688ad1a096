"""Host-side data types of the hot path.

Mirrors the parts of svirlpool.util.datatypes the path touches:
  SVsignal            /root/reference/src/svirlpool/util/datatypes.py:25-72
  SV type codes       datatypes.py:10-22  (0 INS, 1 DEL, 3 BNDL, 4 BNDR)
plus AlignedRecord, a pysam.AlignedSegment stand-in carrying exactly the attributes the
reference's consumers read (seam S1): query_name, reference_name,
reference_start/end, cigartuples, is_reverse, is_unmapped, query_alignment_start/end/length,
reference_length, is_secondary, is_supplementary, flag.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from functools import total_ordering

INS, DEL, BNDL, BNDR = 0, 1, 3, 4
SV_TYPE_DICT = {0: "INS", 1: "DEL", 2: "DEL", 3: "BND", 4: "BND", 5: "INV", 6: "DUP"}

# SAM/BAM CIGAR op codes
CIGAR_M, CIGAR_I, CIGAR_D, CIGAR_N, CIGAR_S, CIGAR_H, CIGAR_P, CIGAR_EQ, CIGAR_X = range(9)
CIGAR_CHARS = "MIDNSHP=X"


@total_ordering
@dataclass(eq=False)
class SVsignal:
    """One SV signal read off a CIGAR (datatypes.py:25-72): ordered by (ref_start, ref_end);
    equality and hash over all six fields."""

    ref_start: int
    ref_end: int
    read_start: int
    read_end: int
    size: int
    sv_type: int

    def _key(self):
        return (self.ref_start, self.ref_end, self.read_start, self.read_end, self.size, self.sv_type)

    def __eq__(self, other):
        return isinstance(other, SVsignal) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __lt__(self, other):
        return (self.ref_start, self.ref_end) < (other.ref_start, other.ref_end)

    def unstructure(self) -> dict:
        return {"ref_start": self.ref_start, "ref_end": self.ref_end, "read_start": self.read_start,
                "read_end": self.read_end, "size": self.size, "sv_type": self.sv_type}

    def _get_description(self, chr: str) -> str:
        return (f"region={chr}:{self.ref_start}-{self.ref_end}, type={SV_TYPE_DICT.get(self.sv_type, 'UNK')}, "
                f"size={self.size}, read_span={self.read_start}-{self.read_end}")


@dataclass
class ReadAlignmentSignals:
    """SV signals of one read alignment (datatypes.py:195-203); SV_signals sorted by reference position."""

    samplename: str
    read_name: str
    reference_name: str
    alignment_forward: bool
    SV_signals: list = field(default_factory=list)

    def unstructure(self) -> dict:
        return {"samplename": self.samplename, "read_name": self.read_name, "reference_name": self.reference_name,
                "alignment_forward": self.alignment_forward, "SV_signals": [s.unstructure() for s in self.SV_signals]}


@dataclass
class AlignedRecord:
    """Duck-typed pysam.AlignedSegment: what `svp_align_batch` results become on the host."""

    query_name: str
    reference_name: str | None
    reference_start: int
    cigartuples: list[tuple[int, int]] | None
    flag: int = 0
    mapping_quality: int = 60
    query_sequence: str | None = None
    tags: dict = field(default_factory=dict)

    # --- flag views -------------------------------------------------------------------------
    @property
    def is_reverse(self) -> bool:
        return bool(self.flag & 16)

    @property
    def is_forward(self) -> bool:
        return not self.is_reverse

    @property
    def is_unmapped(self) -> bool:
        return bool(self.flag & 4)

    @property
    def is_secondary(self) -> bool:
        return bool(self.flag & 256)

    @property
    def is_supplementary(self) -> bool:
        return bool(self.flag & 2048)

    # --- coordinates (pysam semantics) ---------------------------------------------------------
    def _sum(self, ops) -> int:
        return sum(n for op, n in (self.cigartuples or ()) if op in ops)

    @property
    def reference_length(self) -> int:
        return self._sum((CIGAR_M, CIGAR_D, CIGAR_N, CIGAR_EQ, CIGAR_X))

    @property
    def reference_end(self) -> int | None:
        if self.cigartuples is None:
            return None
        return self.reference_start + self.reference_length

    @property
    def query_alignment_start(self) -> int:
        """Offset of the first aligned base in the stored SEQ: soft clips count, hard clips do not."""
        ct = self.cigartuples or ()
        start = 0
        for op, n in ct:
            if op == CIGAR_H:
                continue
            if op == CIGAR_S:
                start += n
                continue
            break
        return start

    @property
    def query_alignment_length(self) -> int:
        return self._sum((CIGAR_M, CIGAR_I, CIGAR_EQ, CIGAR_X))

    @property
    def query_alignment_end(self) -> int:
        return self.query_alignment_start + self.query_alignment_length

    @property
    def query_length(self) -> int:
        return self._sum((CIGAR_M, CIGAR_I, CIGAR_S, CIGAR_EQ, CIGAR_X))

    @property
    def cigarstring(self) -> str | None:
        if self.cigartuples is None:
            return None
        return "".join(f"{n}{CIGAR_CHARS[op]}" for op, n in self.cigartuples)


def cigar_from_string(s: str) -> list[tuple[int, int]]:
    out, num = [], 0
    for ch in s:
        if ch.isdigit():
            num = num * 10 + ord(ch) - 48
        else:
            out.append((CIGAR_CHARS.index(ch), num))
            num = 0
    return out


@dataclass(eq=False)
class MergedSVSignal(SVsignal):
    """SVsignal + chromosome, repeat ids and alt/ref sequences (datatypes.py:409-435); the unit
    consensus_align_lib.parse_sv_signals_from_consensus returns."""

    chr: str = ""
    repeatIDs: list = field(default_factory=list)
    original_alt_sequences: list = field(default_factory=list)
    original_ref_sequences: list = field(default_factory=list)

    def __hash__(self):
        return hash(super().__hash__() + hash(self.chr))

    def unstructure(self) -> dict:
        d = super().unstructure()
        d.update(chr=self.chr, repeatIDs=list(self.repeatIDs), original_alt_sequences=list(self.original_alt_sequences),
                 original_ref_sequences=list(self.original_ref_sequences))
        return d

    def get_alt_sequence(self) -> str:
        s = "".join(self.original_alt_sequences)
        return s[: self.size] if self.sv_type == INS else s

    def get_ref_sequence(self) -> str:
        s = "".join(self.original_ref_sequences)
        return s[: self.size] if self.sv_type == DEL else s


@dataclass
class SequenceObject:
    """datatypes.py:323-341 (a read kept unused by a container)."""

    name: str
    id: str
    sequence: str
    description: str | None = None
    qualities: list | None = None

    def unstructure(self) -> dict:
        return {"name": self.name, "id": self.id, "sequence": self.sequence, "description": self.description,
                "qualities": self.qualities}


@dataclass
class Alignment:
    """Serialisable alignment record (datatypes.py:344-402): the SAM fields as the dict pysam's
    AlignedSegment.to_dict() produces (all values strings, 1-based ref_pos), plus coordinates and annotations.
    One JSON object per line in the partitioned alignment files of consensus_align (consensus_align.py:661-735)."""

    readname: str
    reference_name: str
    reference_ID: int
    reference_start: int
    reference_end: int
    samdict: dict
    headerdict: dict
    annotations: dict | None = None
    samplename: str | None = None

    @classmethod
    def from_record(cls, rec: "AlignedRecord", samplename: str | None = None, annotations: dict | None = None,
                    reference_ID: int = 0, headerdict: dict | None = None) -> "Alignment":
        tags = [f"{k}:{'i' if isinstance(v, int) else 'Z'}:{v}" for k, v in rec.tags.items()]
        samdict = {"name": rec.query_name, "flag": str(rec.flag), "ref_name": rec.reference_name, "ref_pos": str(rec.reference_start + 1),
                   "map_quality": str(rec.mapping_quality), "cigar": rec.cigarstring, "next_ref_name": "*", "next_ref_pos": "0",
                   "length": "0", "seq": rec.query_sequence if rec.query_sequence is not None else "*", "qual": "*", "tags": tags}
        return cls(readname=rec.query_name, reference_name=rec.reference_name, reference_ID=reference_ID,
                   reference_start=rec.reference_start, reference_end=rec.reference_end, samdict=samdict,
                   headerdict=headerdict or {}, annotations=annotations, samplename=samplename)

    def to_record(self) -> "AlignedRecord":
        sd = self.samdict
        tags = {}
        for t in sd.get("tags", []):
            k, ty, v = t.split(":", 2)
            tags[k] = int(v) if ty == "i" else v
        seq = sd.get("seq")
        return AlignedRecord(query_name=sd["name"], reference_name=sd["ref_name"], reference_start=int(sd["ref_pos"]) - 1,
                             cigartuples=cigar_from_string(sd["cigar"]), flag=int(sd["flag"]),
                             mapping_quality=int(sd.get("map_quality", 60)), query_sequence=None if seq in (None, "*") else seq, tags=tags)

    to_pysam = to_record        # the reference's name for the same conversion (datatypes.py:376-378)

    def is_reverse(self) -> bool:
        return int(self.samdict["flag"]) & 0x10 != 0

    def unstructure(self) -> dict:
        return {"readname": self.readname, "reference_name": self.reference_name, "reference_ID": self.reference_ID,
                "reference_start": self.reference_start, "reference_end": self.reference_end, "samdict": self.samdict,
                "headerdict": self.headerdict, "annotations": self.annotations, "samplename": self.samplename}

    @classmethod
    def structure(cls, d: dict) -> "Alignment":
        return cls(**{k: d.get(k) for k in ("readname", "reference_name", "reference_ID", "reference_start", "reference_end",
                                            "samdict", "headerdict", "annotations", "samplename")})

    def __eq__(self, value) -> bool:
        return (self.samdict["ref_pos"] == value.samdict["ref_pos"] and self.samdict["ref_name"] == value.samdict["ref_name"]
                and self.samdict["flag"] == value.samdict["flag"] and self.readname == value.readname
                and self.samdict["cigar"] == value.samdict["cigar"])

    def __hash__(self) -> int:
        return hash((self.readname, self.reference_name, self.samplename, self.samdict["cigar"]))
