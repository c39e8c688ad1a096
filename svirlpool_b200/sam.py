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


def _parse_tags(buf: memoryview) -> dict:
    """Optional fields of a BAM record: integers, floats, strings (Z) — the types this path writes and reads (NM, AS, SA);
    arrays (B) and hex strings are skipped."""
    tags, o, n = {}, 0, len(buf)
    sizes = {"c": ("<b", 1), "C": ("<B", 1), "s": ("<h", 2), "S": ("<H", 2), "i": ("<i", 4), "I": ("<I", 4), "f": ("<f", 4)}
    while o + 3 <= n:
        key, ty = bytes(buf[o:o + 2]).decode(), chr(buf[o + 2])
        o += 3
        if ty in sizes:
            fmt, w = sizes[ty]
            tags[key] = struct.unpack_from(fmt, buf, o)[0]
            o += w
        elif ty == "A":
            tags[key] = chr(buf[o]); o += 1
        elif ty in "ZH":
            end = o
            while end < n and buf[end] != 0:
                end += 1
            tags[key] = bytes(buf[o:end]).decode(); o = end + 1
        elif ty == "B":
            sub, cnt = chr(buf[o]), struct.unpack_from("<i", buf, o + 1)[0]
            o += 5 + cnt * sizes[sub][1]
        else:
            break
    return tags


def _parse_bam_record(rec: memoryview, refs: list[str]) -> AlignedRecord:
    ref_id, p0, l_name, mapq, _bin, n_cig, flag, l_seq, _nr, _np, _tl = struct.unpack_from("<iiBBHHHiiii", rec, 0)
    o = 32
    name = bytes(rec[o:o + l_name - 1]).decode()
    o += l_name
    cig = [(w & 15, w >> 4) for w in struct.unpack_from(f"<{n_cig}I", rec, o)]
    o += 4 * n_cig
    packed = bytes(rec[o:o + (l_seq + 1) // 2])
    o += (l_seq + 1) // 2 + l_seq                   # + qualities
    seq = "".join(_SEQ_CODE[b >> 4] + _SEQ_CODE[b & 15] for b in packed)[:l_seq]
    return AlignedRecord(query_name=name, reference_name=refs[ref_id] if ref_id >= 0 else None, reference_start=p0,
                         cigartuples=cig or None, flag=flag, mapping_quality=mapq, query_sequence=seq or None,
                         tags=_parse_tags(rec[o:]))


def _parse_bam_header(data: bytes) -> tuple[str, list[str], list[int], int]:
    """(header text, reference names, reference lengths, offset of the first record) of uncompressed BAM bytes."""
    if data[:4] != b"BAM\x01":
        raise ValueError("not a BAM file")
    l_text, = struct.unpack_from("<i", data, 4)
    text = data[8:8 + l_text].decode(errors="replace")
    pos = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, pos)
    pos += 4
    names, lens = [], []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", data, pos)
        names.append(data[pos + 4:pos + 4 + l_name - 1].decode())
        lens.append(struct.unpack_from("<i", data, pos + 4 + l_name)[0])
        pos += 4 + l_name + 4
    return text, names, lens, pos


def _read_bam(path: Path) -> Iterator[AlignedRecord]:
    with gzip.open(path, "rb") as f:
        data = f.read()
    _text, refs, _lens, pos = _parse_bam_header(data)
    view = memoryview(data)
    while pos + 4 <= len(data):
        block, = struct.unpack_from("<i", data, pos)
        yield _parse_bam_record(view[pos + 4:pos + 4 + block], refs)
        pos += 4 + block


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


_BGZF_CHUNK = 0xFF00


def _bgzf_blocks(payload: bytes) -> Iterator[bytes]:
    for off in range(0, len(payload), _BGZF_CHUNK):
        chunk = payload[off:off + _BGZF_CHUNK]
        comp = zlib.compressobj(6, zlib.DEFLATED, -15)
        deflated = comp.compress(chunk) + comp.flush()
        bsize = len(deflated) + 25
        yield (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + deflated
               + struct.pack("<II", zlib.crc32(chunk), len(chunk)))


def _write_bai(path, n_ref: int, entries: list[tuple[int, int, int, int, int]]) -> None:
    """BAI index (SAM spec §5.2) from (ref_id, beg, end, virtual offset of the record, virtual offset behind it) of the
    mapped records in file order: per reference the bins with their chunks (adjacent records of a bin share a chunk) and
    the 16 kb linear index."""
    per_ref: list[dict] = [{"bins": {}, "lin": []} for _ in range(n_ref)]
    for ref_id, beg, end, v0, v1 in entries:
        if ref_id < 0:
            continue
        r = per_ref[ref_id]
        chunks = r["bins"].setdefault(_reg2bin(beg, max(end, beg + 1)), [])
        if chunks and chunks[-1][1] == v0:
            chunks[-1][1] = v1
        else:
            chunks.append([v0, v1])
        lin = r["lin"]
        for w in range(beg >> 14, ((max(end, beg + 1) - 1) >> 14) + 1):
            if w >= len(lin):
                lin.extend([0] * (w + 1 - len(lin)))
            if lin[w] == 0 or v0 < lin[w]:
                lin[w] = v0
    out = bytearray(b"BAI\x01" + struct.pack("<i", n_ref))
    for r in per_ref:
        out += struct.pack("<i", len(r["bins"]))
        for b, chunks in sorted(r["bins"].items()):
            out += struct.pack("<Ii", b, len(chunks))
            for v0, v1 in chunks:
                out += struct.pack("<QQ", v0, v1)
        lin, last = r["lin"], 0
        for w in range(len(lin)):                     # empty windows inherit the previous offset
            if lin[w] == 0:
                lin[w] = last
            last = lin[w]
        out += struct.pack("<i", len(lin)) + struct.pack(f"<{len(lin)}Q", *lin)
    with open(path, "wb") as f:
        f.write(bytes(out))


def write_bam(path, records: Iterable[AlignedRecord], references: dict[str, int], sort_order: str = "coordinate",
              command: str | None = None, index: bool = False) -> None:
    """Write a BAM file (BGZF + EOF marker). Records must already be in the stated order. index=True also writes
    `<path>.bai` (what `samtools index` produces at the reference's seam, util/util.py:232-238)."""
    text = _header_text(references, sort_order, command).encode()
    payload = bytearray(b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(references)))
    for n, ln in references.items():
        nm = n.encode() + b"\0"
        payload += struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln)
    ref_index = {n: k for k, n in enumerate(references)}
    spans = []                                       # (ref_id, beg, end, uncompressed offset, uncompressed end)
    for r in records:
        rec = _bam_record(r, ref_index)
        if not r.is_unmapped and r.reference_name in ref_index:
            spans.append((ref_index[r.reference_name], r.reference_start, r.reference_end or r.reference_start + 1,
                          len(payload), len(payload) + len(rec)))
        payload += rec
    block_coff, coff = [], 0
    with open(path, "wb") as f:
        for blk in _bgzf_blocks(bytes(payload)):
            block_coff.append(coff)
            f.write(blk)
            coff += len(blk)
        block_coff.append(coff)                      # the EOF block: where an offset at the very end points
        f.write(_BGZF_EOF)
    if index:
        def voff(u: int) -> int:
            return (block_coff[u // _BGZF_CHUNK] << 16) | (u % _BGZF_CHUNK)
        _write_bai(str(path) + ".bai", len(references), [(rid, b, e, voff(u0), voff(u1)) for rid, b, e, u0, u1 in spans])


# ---------------------------------------------------------------------------------------------
# indexed access (pysam.AlignmentFile.fetch)
# ---------------------------------------------------------------------------------------------

def _reg2bins(beg: int, end: int) -> list[int]:
    end -= 1
    bins = [0]
    for shift, base in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        bins.extend(range(base + (beg >> shift), base + (end >> shift) + 1))
    return bins


class BamFile:
    """Random access to a coordinate-sorted, BAI-indexed BAM without htslib: `fetch(chrom, start, end)` yields the
    records overlapping [start, end) like pysam.AlignmentFile.fetch. BGZF blocks are inflated on demand and cached."""

    def __init__(self, path):
        self.path = Path(path)
        self._f = open(self.path, "rb")
        self._blocks: dict[int, tuple[bytes, int]] = {}           # compressed offset -> (data, compressed size)
        head = bytearray()
        coff = 0
        while True:                                              # the header may span several blocks
            data, size = self._block(coff)
            head += data
            coff += size
            try:
                self.header_text, self.references, self.lengths, first = _parse_bam_header(bytes(head))
                break
            except (struct.error, IndexError):
                if not data:
                    raise ValueError(f"{self.path}: truncated BAM header")
        self._first_voff = self._voff_after(0, first)
        self._index = self._load_bai(Path(str(self.path) + ".bai"))

    def close(self) -> None:
        self._f.close()
        self._blocks.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- BGZF -----------------------------------------------------------------------------------
    def _block(self, coff: int) -> tuple[bytes, int]:
        hit = self._blocks.get(coff)
        if hit is not None:
            return hit
        self._f.seek(coff)
        hdr = self._f.read(18)
        if len(hdr) < 18:
            return b"", 0
        bsize = struct.unpack_from("<H", hdr, 16)[0] + 1
        body = self._f.read(bsize - 18)
        data = zlib.decompress(body[:-8], -15)
        if len(self._blocks) > 256:
            self._blocks.clear()
        self._blocks[coff] = (data, bsize)
        return data, bsize

    def _voff_after(self, coff: int, skip: int) -> int:
        """Virtual offset `skip` uncompressed bytes behind the start of the block at coff."""
        while True:
            data, size = self._block(coff)
            if skip < len(data) or size == 0:
                return (coff << 16) | skip
            skip -= len(data)
            coff += size

    def _read(self, voff: int, n: int) -> tuple[bytes, int]:
        """n uncompressed bytes from a virtual offset (crossing block boundaries) and the virtual offset behind them."""
        coff, uoff = voff >> 16, voff & 0xFFFF
        out = bytearray()
        while n > 0:
            data, size = self._block(coff)
            if size == 0:
                return bytes(out), (coff << 16) | uoff
            take = data[uoff:uoff + n]
            out += take
            n -= len(take)
            uoff += len(take)
            if uoff >= len(data):
                coff, uoff = coff + size, 0
        return bytes(out), (coff << 16) | uoff

    def _records_from(self, voff: int) -> Iterator[tuple[int, AlignedRecord, int, int]]:
        """(virtual offset, record, ref_id, pos) from a virtual offset to the end of the file."""
        while True:
            head, nxt = self._read(voff, 4)
            if len(head) < 4:
                return
            body, nxt = self._read(nxt, struct.unpack("<i", head)[0])
            ref_id, p0 = struct.unpack_from("<ii", body, 0)
            yield voff, _parse_bam_record(memoryview(body), self.references), ref_id, p0
            voff = nxt

    # -- index ----------------------------------------------------------------------------------
    @staticmethod
    def _load_bai(path: Path):
        if not path.exists():
            return None
        d = path.read_bytes()
        if d[:4] != b"BAI\x01":
            raise ValueError(f"{path}: not a BAI index")
        n_ref, = struct.unpack_from("<i", d, 4)
        o, refs = 8, []
        for _ in range(n_ref):
            n_bin, = struct.unpack_from("<i", d, o)
            o += 4
            bins = {}
            for _b in range(n_bin):
                b, n_chunk = struct.unpack_from("<Ii", d, o)
                o += 8
                bins[b] = [struct.unpack_from("<QQ", d, o + 16 * c) for c in range(n_chunk)]
                o += 16 * n_chunk
            n_intv, = struct.unpack_from("<i", d, o)
            lin = list(struct.unpack_from(f"<{n_intv}Q", d, o + 4))
            o += 4 + 8 * n_intv
            refs.append((bins, lin))
        return refs

    def __iter__(self) -> Iterator[AlignedRecord]:
        for _v, rec, _r, _p in self._records_from(self._first_voff):
            yield rec

    def fetch(self, chrom: str, start: int, end: int) -> Iterator[AlignedRecord]:
        if chrom not in self.references:
            raise ValueError(f"invalid contig `{chrom}`")
        rid = self.references.index(chrom)
        start, end = max(int(start), 0), max(int(end), int(start) + 1)
        if self._index is None:                                 # no index: scan (small files, tests)
            for _v, rec, r, _p in self._records_from(self._first_voff):
                if r == rid and rec.reference_start < end and (rec.reference_end or rec.reference_start + 1) > start:
                    yield rec
            return
        bins, lin = self._index[rid]
        min_off = lin[min(start >> 14, len(lin) - 1)] if lin else 0
        chunks = sorted(c for b in _reg2bins(start, end) if b in bins for c in bins[b] if c[1] > min_off)
        seen = -1
        for v0, v1 in chunks:
            for voff, rec, r, p0 in self._records_from(v0):
                if voff >= v1:
                    break
                if voff <= seen:
                    continue
                seen = voff
                if r != rid or p0 >= end:
                    break
                if (rec.reference_end or p0 + 1) > start:
                    yield rec


__all__ = ["read_alignments", "write_sam", "write_bam", "sort_records", "BamFile", "CIGAR_CHARS"]
