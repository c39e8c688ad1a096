"""Group candidate-region containers into coordinate-adjacent batches.

API-compatible host mirror of svirlpool.candidateregions.crs_to_batches
(/root/reference/src/svirlpool/candidateregions/crs_to_batches.py:36-179) and of the TSV reader
svirlpool.localassembly.consensus._load_crIDs_from_batch_tsv (consensus.py:3134-3160):
same function names, arguments, batch dicts, and the frozen TSV layout
    batch_id <TAB> chr <TAB> start <TAB> end <TAB> n_containers <TAB> crIDs(comma separated)
Batches are equal in container COUNT; balancing by DP work happens below this API in
svirlpool_b200.scheduler.
"""
from __future__ import annotations

import argparse
import json
import logging
import sqlite3
from pathlib import Path

log = logging.getLogger(__name__)

TSV_HEADER = ("batch_id", "chr", "start", "end", "n_containers", "crIDs")
Extent = tuple[list[str], int, int]


def _load_container_extents(path_db: Path) -> dict[int, Extent]:
    """{representative crID: (sorted chromosomes, min referenceStart, max referenceEnd)} from the
    `containers(crID, data JSON)` table; containers without CRs are skipped."""
    path_db = Path(path_db)
    if not path_db.exists():
        raise FileNotFoundError(f"Containers database {path_db} does not exist.")
    out: dict[int, Extent] = {}
    conn = sqlite3.connect(f"file:{path_db}?mode=ro", uri=True)
    try:
        for cr_id, data in conn.execute("SELECT crID, data FROM containers"):
            crs = json.loads(data).get("crs", [])
            if crs:
                out[int(cr_id)] = (sorted({str(c["chr"]) for c in crs}),
                                   min(int(c["referenceStart"]) for c in crs),
                                   max(int(c["referenceEnd"]) for c in crs))
    finally:
        conn.close()
    return out


def build_batches(extents: dict[int, Extent], batch_size: int, max_batch_span: int) -> list[dict]:
    """Wide containers (several chromosomes, or span > max_batch_span) become singleton batches,
    emitted first in input order. The rest are sorted by (chr, start, end, crID) and packed
    greedily; a batch closes when the chromosome changes, it already holds batch_size containers,
    or adding the next container would stretch its span beyond max_batch_span. Batch ids are
    sequential from 0."""
    if batch_size <= 0:
        raise ValueError("batch_size must be > 0")
    if max_batch_span <= 0:
        raise ValueError("max_batch_span must be > 0")
    batches: list[dict] = []

    def emit(chrom: str, start: int, end: int, cr_ids: list[int]) -> None:
        batches.append({"batch_id": len(batches), "chr": chrom, "start": start, "end": end, "crIDs": cr_ids})

    narrow: list[tuple[str, int, int, int]] = []
    for cr_id, (chrs, start, end) in extents.items():
        if len(chrs) > 1 or end - start > max_batch_span:
            emit(",".join(chrs), start, end, [cr_id])
        else:
            narrow.append((chrs[0], start, end, cr_id))
    narrow.sort()

    members: list[int] = []
    cur_chr, cur_start, cur_end = None, 0, 0
    for chrom, start, end, cr_id in narrow:
        if members:
            grown = (min(cur_start, start), max(cur_end, end))
            if chrom != cur_chr or len(members) >= batch_size or grown[1] - grown[0] > max_batch_span:
                emit(cur_chr, cur_start, cur_end, members)
                members = []
            else:
                cur_start, cur_end = grown
        if not members:
            cur_chr, cur_start, cur_end = chrom, start, end
        members.append(cr_id)
    if members:
        emit(cur_chr, cur_start, cur_end, members)
    return batches


def write_outputs(batches: list[dict], tsv_path: Path, ids_path: Path) -> None:
    with open(tsv_path, "w") as f:
        f.write("\t".join(TSV_HEADER) + "\n")
        for b in batches:
            ids = ",".join(str(c) for c in b["crIDs"])
            f.write(f"{b['batch_id']}\t{b['chr']}\t{b['start']}\t{b['end']}\t{len(b['crIDs'])}\t{ids}\n")
    with open(ids_path, "w") as f:
        f.writelines(f"{b['batch_id']}\n" for b in batches)


def _load_crIDs_from_batch_tsv(batch_tsv: Path, batch_id: int) -> list[int]:
    """crIDs of one batch row; FileNotFoundError / ValueError exactly where the reference raises them."""
    batch_tsv = Path(batch_tsv)
    if not batch_tsv.exists():
        raise FileNotFoundError(f"Batch TSV {batch_tsv} does not exist.")
    with open(batch_tsv) as f:
        header = f.readline().rstrip("\n").split("\t")
        if "batch_id" not in header or "crIDs" not in header:
            raise ValueError(f"Batch TSV {batch_tsv} is missing expected columns 'batch_id' / 'crIDs'.")
        id_col, crs_col = header.index("batch_id"), header.index("crIDs")
        for line in f:
            fields = line.rstrip("\n").split("\t")
            if not fields or not fields[0]:
                continue
            if int(fields[id_col]) == batch_id:
                csv = fields[crs_col].strip()
                return [int(x) for x in csv.split(",")] if csv else []
    raise ValueError(f"Batch id {batch_id} not found in {batch_tsv}.")


def crs_containers_to_batches(path_containers: Path, tsv_out: Path, ids_out: Path, batch_size: int = 100,
                              max_batch_span: int = 20_000_000) -> list[dict]:
    extents = _load_container_extents(path_containers)
    batches = build_batches(extents=extents, batch_size=batch_size, max_batch_span=max_batch_span)
    write_outputs(batches=batches, tsv_path=tsv_out, ids_path=ids_out)
    log.info("Built %d batches from %d containers (batch_size=%d, max_batch_span=%d).",
             len(batches), len(extents), batch_size, max_batch_span)
    return batches


def get_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Group candidate-region containers into batches of adjacent containers.")
    p.add_argument("-i", "--input", type=Path, required=True, help="crs containers SQLite database")
    p.add_argument("-t", "--tsv", type=Path, required=True, help="output TSV, one row per batch")
    p.add_argument("--ids", type=Path, required=True, help="output text file, one batch id per line")
    p.add_argument("--batch-size", type=int, default=100)
    p.add_argument("--max-batch-span", type=int, default=20_000_000)
    p.add_argument("--log-level", type=str, choices=["DEBUG", "INFO", "WARNING", "ERROR", "CRITICAL"], default="INFO")
    return p


def run(args, **kwargs) -> None:
    logging.basicConfig(level=getattr(logging, getattr(args, "log_level", "INFO")), force=True,
                        format="%(asctime)s - %(name)s - %(levelname)s - %(message)s")
    crs_containers_to_batches(args.input, args.tsv, args.ids, args.batch_size, args.max_batch_span)


def main() -> None:
    run(get_parser().parse_args())


if __name__ == "__main__":
    main()
