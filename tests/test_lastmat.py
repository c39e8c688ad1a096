"""last-train .mat parser (svirlpool_b200.lastmat; the reference passes such a file to lamassemble, consensus.py:764-813)."""
from pathlib import Path

import numpy as np
import pytest

from svirlpool_b200 import lastmat


def test_parse_stand_in(tmp_path):
    p = tmp_path / "stand_in.mat"
    p.write_text(lastmat.stand_in_mat_text())
    s = lastmat.parse_last_mat(p)
    m = s["matrix"][0].reshape(4, 4)                     # [query][target]
    assert m[0].tolist() == [6, -18, -9, -19] and m[3, 3] == 6 and m[1, 3] == -8
    assert (int(s["del_open1"][0]), int(s["del_ext1"][0]), int(s["ins_open1"][0]), int(s["ins_ext1"][0])) == (21, 9, 25, 7)
    assert int(s["del_open2"][0]) == -1 and int(s["ins_open2"][0]) == -1 and int(s["min_signal_size"][0]) == 12


def test_rows_are_reference_letters_and_later_blocks_win(tmp_path):
    """An asymmetric matrix: file rows = reference (target) letters, columns = query letters; last-train logs earlier rounds first."""
    p = tmp_path / "train.mat"
    p.write_text("\n".join([
        "# lastTrain round 1", "#last -a 30", "#last -b 2", "       A      C      G      T", "A 1 -1 -1 -1", "C -1 1 -1 -1", "G -1 -1 1 -1", "T -1 -1 -1 1", "",
        "# lastTrain round 2", "#last -a 12", "#last -A 14", "#last -b 3", "#last -B 2", "#last -S 1",
        "# score matrix (query letters = columns, reference letters = rows):",
        "       A      C      G      T      N",
        "A      5    -11    -12    -13     -1",
        "C    -21      6    -22    -23     -1",
        "G    -31    -32      7    -33     -1",
        "T    -41    -42    -43      8     -1",
        "N     -1     -1     -1     -1     -1", ""]))
    s = lastmat.parse_last_mat(p, min_signal_size=8)
    m = s["matrix"][0].reshape(4, 4)
    assert m[1, 0] == -11 and m[0, 1] == -21 and m[3, 2] == -33        # s[query C][target A] = row A, column C
    assert np.diag(m).tolist() == [5, 6, 7, 8]
    assert (int(s["del_open1"][0]), int(s["ins_open1"][0]), int(s["del_ext1"][0]), int(s["ins_ext1"][0])) == (12, 14, 3, 2)
    assert int(s["min_signal_size"][0]) == 8


def test_lfs_pointer_is_reported(tmp_path):
    p = tmp_path / "promethion.mat"
    p.write_text("version https://git-lfs.github.com/spec/v1\noid sha256:c97aae418f798dc0933bf480aa834f279f7fec269f63a2830526cc6c2f1d17ae\nsize 18429\n")
    with pytest.raises(lastmat.LastMatError, match="Git-LFS pointer"):
        lastmat.parse_last_mat(p)
    ref = Path("/root/reference/data/lamassemble-mats/promethion.mat")        # the reference checkout here holds exactly such a pointer
    if ref.exists():
        with pytest.raises(lastmat.LastMatError, match="Git-LFS pointer"):
            lastmat.parse_last_mat(ref)


def test_incomplete_files(tmp_path):
    p = tmp_path / "x.mat"
    p.write_text("#last -a 10\n#last -b 2\n# no matrix here\n")
    with pytest.raises(lastmat.LastMatError, match="no score matrix"):
        lastmat.parse_last_mat(p)
    p.write_text("   A C G T\nA 1 -1 -1 -1\nC -1 1 -1 -1\nG -1 -1 1 -1\nT -1 -1 -1 1\n")
    with pytest.raises(lastmat.LastMatError, match="gap costs"):
        lastmat.parse_last_mat(p)


def test_parsed_scoring_runs_on_the_planner_and_the_oracle(tmp_path):
    """The parsed profile is accepted by the library (general scoring kernels) and aligns on the oracle."""
    from oracle_engine import OracleEngine
    from svirlpool_b200 import aligner, ava
    p = tmp_path / "stand_in.mat"
    p.write_text(lastmat.stand_in_mat_text())
    s = lastmat.parse_last_mat(p)
    rng = np.random.default_rng(3)
    a = "".join("ACGT"[k] for k in rng.integers(0, 4, 600))
    b = a[:300] + a[320:]
    batch = ava.PairBatch.build(["a", "b"], [a, b], [(1, 0, *ava.full_band(len(b), len(a)), 0)])
    plan = aligner.plan_pairs(s, batch.packed.nbytes, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
    assert int(plan[0]["flags"]) & 8                     # SVP_PLAN_GENERAL
    res, cig = OracleEngine(scoring=s).align_batch(batch.packed, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
    assert int(res[0]["n_del_bp"]) == 20 and int(res[0]["score"]) > 2500
