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
