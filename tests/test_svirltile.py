"""Output writers: consensus.fasta and svirltile.db are readable with the reference's own
reader conventions (svirltile.py:154-222) and carry the reference's table layouts."""
import pickle
import sqlite3

from svirlpool_b200 import svirltile, synth, workloads


def _make_db(path, table, rows, ddl):
    with sqlite3.connect(path) as conn:
        conn.execute(ddl)
        conn.executemany(f"INSERT INTO {table} VALUES ({','.join('?' * len(rows[0]))})", rows)
        conn.commit()


def test_sample_database_roundtrip(tmp_path):
    loc = synth.locus_config4(0, scale=0.1, max_reads=4)
    cons = workloads.consensus_objects(loc)
    fasta = tmp_path / "consensus.fasta"
    assert svirltile.write_consensus_fasta(fasta, cons.values()) == len(cons)
    text = fasta.read_text().splitlines()
    assert text[0] == f">{next(iter(cons))}" and max(len(l) for l in text) <= 60
    # upstream databases with the reference's table layouts (SVpatterns.py:1796-1811, rafs_to_coverage.py:97-100,
    # copynumber_tracks.py:45-49)
    _make_db(tmp_path / "svp.db", "svPatterns", [("0-INS-0.0", 0, "0.0", pickle.dumps({"x": 1}))],
             "CREATE TABLE svPatterns (svPatternID VARCHAR(100) PRIMARY KEY, crID INTEGER, consensusID VARCHAR(50), svPattern BLOB)")
    _make_db(tmp_path / "cov.db", "coverage", [(0, pickle.dumps([1, 2, 3]))], "CREATE TABLE coverage (id INTEGER PRIMARY KEY, data BLOB)")
    _make_db(tmp_path / "cn.db", "copynumber_tracks", [("chr1", pickle.dumps("tree"))],
             "CREATE TABLE copynumber_tracks (chromosome VARCHAR(50) PRIMARY KEY, intervaltree BLOB)")
    out = tmp_path / "svirltile.db"
    svirltile.create_sample_database(tmp_path / "svp.db", tmp_path / "cov.db", fasta, "HG002", out, copynumber_db_path=tmp_path / "cn.db")
    with sqlite3.connect(out) as conn:
        tables = {r[0] for r in conn.execute("SELECT name FROM sqlite_master WHERE type='table'")}
        assert tables == {"svPatterns", "coverage", "copynumber_tracks", "consensus_sequences", "metadata"}
        assert conn.execute("SELECT crID, consensusID FROM svPatterns").fetchall() == [(0, "0.0")]
    got = dict(svirltile.get_consensus_sequences(out))
    assert set(got) == set(cons)
    for cid, c in cons.items():
        seq = svirltile.decompress_sequence(got[cid]["sequence"])
        assert seq == c.consensus_padding.sequence                             # lower-case flanks, UPPER core preserved
        lo, hi = c.consensus_padding.consensus_interval_on_sequence_with_padding
        assert seq[lo:hi].isupper() and (lo == 0 or seq[:lo].islower())
    assert list(svirltile.get_consensus_sequences(out, [next(iter(cons))]))[0][0] == next(iter(cons))
    assert svirltile.get_metadata(out) == {"samplename": "HG002"}
    # a missing copy-number database is skipped, an existing output is overwritten
    svirltile.create_sample_database(tmp_path / "svp.db", tmp_path / "cov.db", fasta, "S2", out, copynumber_db_path=tmp_path / "none.db")
    assert svirltile.get_metadata(out) == {"samplename": "S2"}


def test_consensus_sequences_rows_byte_for_byte(tmp_path):
    """The rows of consensus_sequences against what the reference's writer stores (localassembly/svirltile.py:78-110:
    `SeqIO.parse(fasta, "fasta")`, then (record.id, pickle.dumps(str(record.seq)), record.description or "")), restated by hand
    for a FASTA with Biopython-relevant details: a header with a description (record.description is the WHOLE title line, id
    included), a header with trailing blanks, wrapped sequence lines, CRLF line ends and a blank line inside a record."""
    fasta = tmp_path / "consensus.fasta"
    fasta.write_bytes(b">12.1 crID=12 reads=7 \nacgtACGTNN\nACGTacgt\n>13.0\r\nGGGG\r\n\r\ncccc\r\n>14.0\tsample=HG002\nTTTT\n")
    expected = [
        ("12.1", pickle.dumps("acgtACGTNNACGTacgt"), "12.1 crID=12 reads=7"),
        ("13.0", pickle.dumps("GGGGcccc"), "13.0"),
        ("14.0", pickle.dumps("TTTT"), "14.0\tsample=HG002"),
    ]
    db = tmp_path / "t.db"
    sqlite3.connect(db).close()
    assert svirltile._add_consensus_sequences_to_db(db, fasta) == 3
    with sqlite3.connect(db) as conn:
        rows = conn.execute("SELECT consensusID, sequence, description FROM consensus_sequences ORDER BY consensusID").fetchall()
        ddl = " ".join(conn.execute("SELECT sql FROM sqlite_master WHERE name='consensus_sequences'").fetchone()[0].split())
        idx = conn.execute("SELECT name FROM sqlite_master WHERE type='index' AND tbl_name='consensus_sequences' AND name NOT LIKE 'sqlite_%'").fetchall()
    assert [(a, bytes(b), c) for a, b, c in rows] == expected
    assert ddl == "CREATE TABLE consensus_sequences ( consensusID VARCHAR(50) PRIMARY KEY, sequence BLOB NOT NULL, description TEXT )".replace("( ", "(").replace(" )", ")") \
        or ddl == "CREATE TABLE consensus_sequences (consensusID VARCHAR(50) PRIMARY KEY, sequence BLOB NOT NULL, description TEXT)"
    assert idx == [("idx_consensus_sequences_id",)]
    # metadata table: the reference's layout and its single row
    svirltile._add_metadata_to_db(db, "HG002")
    with sqlite3.connect(db) as conn:
        assert conn.execute("SELECT key, value FROM metadata").fetchall() == [("samplename", "HG002")]
        mddl = " ".join(conn.execute("SELECT sql FROM sqlite_master WHERE name='metadata'").fetchone()[0].split())
    assert mddl == "CREATE TABLE metadata (key VARCHAR(50) PRIMARY KEY, value TEXT NOT NULL)"
    # and the reference's reader convention on those rows (svirltile.py:154-197: pickle.loads of the BLOB)
    got = {cid: svirltile.decompress_sequence(d["sequence"]) for cid, d in svirltile.get_consensus_sequences(db)}
    assert got == {"12.1": "acgtACGTNNACGTacgt", "13.0": "GGGGcccc", "14.0": "TTTT"}
