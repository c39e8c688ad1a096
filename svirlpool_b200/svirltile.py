"""Output writers on the path's far side: consensus.fasta and the per-sample svirltile.db — host mirror of
svirlpool.localassembly.svirltile (/root/reference/src/svirlpool/localassembly/svirltile.py:20-222),
svirlpool.util.combine_databases (util/combine_databases.py:13-70) and the FASTA layout of
consensus_align.write_consensus_files_for_parallel_processing / write_sequence_fastas_to_file
(consensus_align.py:595-658, 1663-1678). Table layouts and value encodings are the reference's (API-compatible
readers: svirltile.get_consensus_sequences :154-197, get_metadata :200-219):

    consensus_sequences(consensusID VARCHAR(50) PRIMARY KEY, sequence BLOB = pickle.dumps(str), description TEXT)
    metadata(key VARCHAR(50) PRIMARY KEY, value TEXT)
    + every table of the svpatterns / coverage / copy-number databases, copied with CREATE TABLE x AS SELECT *.
"""
from __future__ import annotations

import logging
import pickle
import sqlite3
from pathlib import Path
from typing import Iterable, Iterator

from .consensus_class import Consensus
from .util import read_fasta_records

log = logging.getLogger(__name__)


def write_consensus_fasta(path: Path | str, consensuses: Iterable[Consensus], line_width: int = 60) -> int:
    """consensus.fasta: one record per consensus, id = consensus ID, sequence = the PADDED sequence (lower-case flanks,
    UPPER-case core), wrapped like Bio.SeqIO's FASTA writer (60 columns). Returns the number of records."""
    n = 0
    with open(path, "w") as f:
        for c in consensuses:
            if c.consensus_padding is None:
                raise ValueError(f"Consensus {c.ID} has no consensus_padding. This should not happen. Please check the input data.")
            seq = c.consensus_padding.sequence
            f.write(f">{c.ID}\n")
            for k in range(0, len(seq), line_width):
                f.write(seq[k:k + line_width] + "\n")
            n += 1
    return n


def combine_databases(db_paths: list[Path], output_db_path: Path, overwrite: bool = False) -> None:
    """Copy every table of every input database into the output database (first one wins on a name clash)."""
    output_db_path = Path(output_db_path)
    if overwrite and output_db_path.exists():
        output_db_path.unlink()
    conn = sqlite3.connect(output_db_path)
    try:
        cur = conn.cursor()
        for i, db in enumerate(db_paths):
            cur.execute(f"ATTACH DATABASE ? AS db_{i}", (str(db),))
            cur.execute(f"SELECT name FROM db_{i}.sqlite_master WHERE type='table'")
            for (table,) in cur.fetchall():
                cur.execute("SELECT name FROM sqlite_master WHERE type='table' AND name=?", (table,))
                if cur.fetchone():
                    log.info("Table %s already exists in the output database. Skipping.", table)
                    continue
                cur.execute(f'CREATE TABLE "{table}" AS SELECT * FROM db_{i}."{table}"')
            conn.commit()
            cur.execute(f"DETACH DATABASE db_{i}")
        conn.commit()
    finally:
        conn.close()


def _add_consensus_sequences_to_db(db_path: Path, fasta_path: Path) -> int:
    rows = [(rid, pickle.dumps(seq), desc or "") for rid, desc, seq in read_fasta_records(fasta_path)]
    with sqlite3.connect(str(db_path)) as conn:
        conn.execute("CREATE TABLE IF NOT EXISTS consensus_sequences (consensusID VARCHAR(50) PRIMARY KEY, sequence BLOB NOT NULL, description TEXT)")
        conn.executemany("INSERT OR REPLACE INTO consensus_sequences (consensusID, sequence, description) VALUES (?, ?, ?)", rows)
        conn.execute("CREATE INDEX IF NOT EXISTS idx_consensus_sequences_id ON consensus_sequences (consensusID)")
        conn.commit()
    return len(rows)


def _add_metadata_to_db(db_path: Path, samplename: str) -> None:
    with sqlite3.connect(str(db_path)) as conn:
        conn.execute("CREATE TABLE IF NOT EXISTS metadata (key VARCHAR(50) PRIMARY KEY, value TEXT NOT NULL)")
        conn.execute("INSERT OR REPLACE INTO metadata (key, value) VALUES (?, ?)", ("samplename", str(samplename)))
        conn.commit()


def create_sample_database(svpatterns_db_path, coverage_db_path, consensus_fasta_path, samplename: str, output_db_path,
                           copynumber_db_path=None) -> None:
    """svirltile.db of one sample: the svpatterns / coverage (/ copy-number) tables + consensus sequences + metadata."""
    dbs = [Path(svpatterns_db_path), Path(coverage_db_path)]
    if copynumber_db_path is not None:
        if Path(copynumber_db_path).exists():
            dbs.append(Path(copynumber_db_path))
        else:
            log.warning("Copy number database not found: %s. Continuing without it.", copynumber_db_path)
    combine_databases(dbs, Path(output_db_path), overwrite=True)
    _add_consensus_sequences_to_db(Path(output_db_path), Path(consensus_fasta_path))
    _add_metadata_to_db(Path(output_db_path), samplename)


def get_consensus_sequences(db_path: Path, consensus_ids: list[str] | None = None) -> Iterator[tuple[str, dict]]:
    with sqlite3.connect(str(db_path)) as conn:
        c = conn.cursor()
        try:
            if consensus_ids:
                marks = ",".join("?" * len(consensus_ids))
                c.execute(f"SELECT consensusID, sequence, description FROM consensus_sequences WHERE consensusID IN ({marks})", consensus_ids)
            else:
                c.execute("SELECT consensusID, sequence, description FROM consensus_sequences")
            for cid, blob, desc in c.fetchall():
                yield cid, {"id": cid, "sequence": blob, "description": desc}
        finally:
            c.close()


def get_metadata(db_path: Path) -> dict[str, str]:
    with sqlite3.connect(str(db_path)) as conn:
        return {k: v for k, v in conn.execute("SELECT key, value FROM metadata")}


def decompress_sequence(compressed_sequence: bytes) -> str:
    return pickle.loads(compressed_sequence)
