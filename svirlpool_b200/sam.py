"""SAM text / BAM reader and writer without htslib.

The file seam of the path is util.align_reads_with_minimap(): FASTA in, coordinate-sorted BAM out
(/root/reference/src/svirlpool/util/util.py:162-238); its consumers open the result with
pysam.AlignmentFile (consensus_lib.py:55, :295; consensus.py:1399). pysam/htslib are not available
here, so this module writes and reads the two formats directly: SAM as text, BAM as BGZF-framed
records (BGZF is a series of gzip members; the SAM/BAM specification v1 §4). Records are
datatypes.AlignedRecord.
"""
from __future__ import annotations

import gzip
import struct
import zlib
from pathlib import Path
from typing import Iterable, Iterator

from .datatypes import CIGAR_CHARS, AlignedRecord, cigar_from_string

_SEQ_CODE = "=ACMGRSVTWYHKDBN"
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def read_alignments(path) -> Iterator[AlignedRecord]:
    """Yield every record of a SAM (text) or BAM file."""
    path = Path(path)
    with open(path, "rb") as f:
        magic = f.read(2)
    if magic == b"\x1f\x8b":
        yield from _read_bam(path)
    else:
        yield from _read_sam(path)


def _read_sam(path: Path) -> Iterator[AlignedRecord]:
    with open(path) as f:
        for line in f:
            if not line or line[0] == "@":
                continue
            c = line.rstrip("\n").split("\t")
            if len(c) < 11:
                continue
            flag = int(c[1])
            tags = {}
            for tag in c[11:]:
                k, ty, v = tag.split(":", 2)
                tags[k] = int(v) if ty == "i" else (float(v) if ty == "f" else v)
            yield AlignedRecord(
                query_name=c[0], reference_name=None if c[2] == "*" else c[2], reference_start=int(c[3]) - 1,
                cigartuples=None if c[5] == "*" else cigar_from_string(c[5]), flag=flag,
                mapping_quality=int(c[4]), query_sequence=None if c[9] == "*" else c[9], tags=tags)


def _read_bam(path: Path) -> Iterator[AlignedRecord]:
    with gzip.open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"BAM\x01":
        raise ValueError(f"{path}: not a BAM file")
    l_text, = struct.unpack_from("<i", data, 4)
    pos = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, pos)
    pos += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", data, pos)
        refs.append(data[pos + 4:pos + 4 + l_name - 1].decode())
        pos += 4 + l_name + 4
    while pos + 4 <= len(data):
        block, = struct.unpack_from("<i", data, pos)
        rec = memoryview(data)[pos + 4:pos + 4 + block]
        pos += 4 + block
        ref_id, p0, l_name, mapq, _bin, n_cig, flag, l_seq, _nr, _np, _tl = struct.unpack_from("<iiBBHHHiiii", rec, 0)
        o = 32
        name = bytes(rec[o:o + l_name - 1]).decode()
        o += l_name
        cig = [(w & 15, w >> 4) for w in struct.unpack_from(f"<{n_cig}I", rec, o)]
        o += 4 * n_cig
        packed = bytes(rec[o:o + (l_seq + 1) // 2])
        seq = "".join(_SEQ_CODE[b >> 4] + _SEQ_CODE[b & 15] for b in packed)[:l_seq]
        yield AlignedRecord(query_name=name, reference_name=refs[ref_id] if ref_id >= 0 else None,
                            reference_start=p0, cigartuples=cig or None, flag=flag, mapping_quality=mapq,
                            query_sequence=seq or None)


def _sam_line(r: AlignedRecord) -> str:
    cigar = r.cigarstring or "*"
    seq = r.query_sequence or "*"
    tags = "".join(f"\t{k}:{'i' if isinstance(v, int) else ('f' if isinstance(v, float) else 'Z')}:{v}"
                   for k, v in r.tags.items())
    return (f"{r.query_name}\t{r.flag}\t{r.reference_name or '*'}\t{r.reference_start + 1 if not r.is_unmapped else 0}"
            f"\t{r.mapping_quality}\t{cigar}\t*\t0\t0\t{seq}\t*{tags}\n")


def _header_text(references: dict[str, int], sort_order: str, command: str | None) -> str:
    head = f"@HD\tVN:1.6\tSO:{sort_order}\n" + "".join(f"@SQ\tSN:{n}\tLN:{ln}\n" for n, ln in references.items())
    head += "@PG\tID:svirlpool_b200\tPN:svirlpool_b200" + (f"\tCL:{command}" if command else "") + "\n"
    return head


def sort_records(records: Iterable[AlignedRecord], references: dict[str, int]) -> list[AlignedRecord]:
    """Coordinate order as `samtools sort` produces it: (reference index, position), unmapped last."""
    order = {n: k for k, n in enumerate(references)}
    return sorted(records, key=lambda r: (order.get(r.reference_name, len(order)) if not r.is_unmapped else len(order) + 1,
                                          r.reference_start))


def write_sam(path, records: Iterable[AlignedRecord], references: dict[str, int], sort_order: str = "coordinate",
              command: str | None = None) -> None:
    with open(path, "w") as f:
        f.write(_header_text(references, sort_order, command))
        for r in records:
            f.write(_sam_line(r))


def _reg2bin(beg: int, end: int) -> int:
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def _bam_record(r: AlignedRecord, ref_index: dict[str, int]) -> bytes:
    name = r.query_name.encode() + b"\0"
    cig = r.cigartuples or []
    seq = r.query_sequence or ""
    codes = [_SEQ_CODE.index(ch) if ch in _SEQ_CODE else 15 for ch in seq.upper()]
    if len(codes) & 1:
        codes.append(0)
    packed = bytes((codes[k] << 4) | codes[k + 1] for k in range(0, len(codes), 2))
    ref_id = ref_index.get(r.reference_name, -1) if not r.is_unmapped else -1
    end = r.reference_end if cig else r.reference_start + 1
    body = struct.pack("<iiBBHHHiiii", ref_id, r.reference_start if ref_id >= 0 else -1, len(name), r.mapping_quality,
                       _reg2bin(max(r.reference_start, 0), max(end or 1, 1)), len(cig), r.flag, len(seq), -1, -1, 0)
    body += name + struct.pack(f"<{len(cig)}I", *((n << 4) | op for op, n in cig)) + packed + b"\xff" * len(seq)
    for k, v in r.tags.items():
        if isinstance(v, int):
            body += k.encode() + b"i" + struct.pack("<i", v)
        else:
            body += k.encode() + b"Z" + str(v).encode() + b"\0"
    return struct.pack("<i", len(body)) + body


def _bgzf_blocks(payload: bytes) -> Iterator[bytes]:
    for off in range(0, len(payload), 0xFF00):
        chunk = payload[off:off + 0xFF00]
        comp = zlib.compressobj(6, zlib.DEFLATED, -15)
        deflated = comp.compress(chunk) + comp.flush()
        bsize = len(deflated) + 25
        yield (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + deflated
               + struct.pack("<II", zlib.crc32(chunk), len(chunk)))


def write_bam(path, records: Iterable[AlignedRecord], references: dict[str, int], sort_order: str = "coordinate",
              command: str | None = None) -> None:
    """Write a BAM file (BGZF + EOF marker). Records must already be in the stated order."""
    text = _header_text(references, sort_order, command).encode()
    payload = bytearray(b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(references)))
    for n, ln in references.items():
        nm = n.encode() + b"\0"
        payload += struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln)
    ref_index = {n: k for k, n in enumerate(references)}
    for r in records:
        payload += _bam_record(r, ref_index)
    with open(path, "wb") as f:
        for blk in _bgzf_blocks(bytes(payload)):
            f.write(blk)
        f.write(_BGZF_EOF)


__all__ = ["read_alignments", "write_sam", "write_bam", "sort_records", "CIGAR_CHARS"]
