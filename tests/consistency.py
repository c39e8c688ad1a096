"""Size-independent self-checks of alignment results (used where the oracle's full-band matrix would not fit in
memory): a result must be internally consistent with its CIGAR and the sequences — the CIGAR consumes exactly the
reported intervals, the match/mismatch/indel counts are those of the CIGAR walked over the bases, and the score is
the score of that path under the scoring profile (the property `AS == score of the reported CIGAR` that pins the
spec on all 20 frozen minimap2 records)."""
from __future__ import annotations

import numpy as np

from svirlpool_b200._abi import PAIR_REVCOMP, cigartuples

_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


def unpack(packed: np.ndarray, off: int, length: int) -> np.ndarray:
    raw = packed[off:off + (length + 3) // 4]
    codes = np.stack([(raw >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)
    return codes[:length]


def gap_cost(o1: int, e1: int, o2: int, e2: int, k: int) -> int:
    c = o1 + k * e1
    return min(c, o2 + k * e2) if o2 >= 0 else c


def check_result(packed, seq_off, seq_len, pair, res, cig, scoring) -> None:
    """Raises AssertionError if the result row is not a valid alignment with the reported score and counts."""
    if int(res["status"]) & 1:
        return
    q = unpack(packed, int(seq_off[pair["q"]]), int(seq_len[pair["q"]]))
    t = unpack(packed, int(seq_off[pair["t"]]), int(seq_len[pair["t"]]))
    if int(pair["flags"]) & PAIR_REVCOMP:
        q = _COMP[q[::-1]]
    sc = scoring[0]
    mat = np.asarray(sc["matrix"]).reshape(4, 4)
    i, j = int(res["q_beg"]), int(res["t_beg"])
    lo, hi = int(pair["band_lo"]), int(pair["band_lo"]) + int(pair["band_w"]) - 1
    score = n_match = n_mis = ins_bp = del_bp = ins_ops = del_ops = 0
    ops = cigartuples(cig, res)
    assert ops and ops[0][0] == 0 and ops[-1][0] == 0, "a local alignment starts and ends with a match run"
    for op, n in ops:
        assert n > 0
        if op == 0:
            assert lo <= j - i <= hi, "path leaves the band"
            qs, ts = q[i:i + n], t[j:j + n]
            score += int(mat[qs, ts].sum())
            same = int((qs == ts).sum())
            n_match += same; n_mis += n - same
            i += n; j += n
        elif op == 1:
            score -= gap_cost(int(sc["ins_open1"]), int(sc["ins_ext1"]), int(sc["ins_open2"]), int(sc["ins_ext2"]), n)
            ins_bp += n; ins_ops += 1; i += n
        elif op == 2:
            score -= gap_cost(int(sc["del_open1"]), int(sc["del_ext1"]), int(sc["del_open2"]), int(sc["del_ext2"]), n)
            del_bp += n; del_ops += 1; j += n
        else:
            raise AssertionError(f"unexpected CIGAR op {op}")
        assert lo <= j - i <= hi, "path leaves the band"
    assert (i, j) == (int(res["q_end"]), int(res["t_end"])), "CIGAR does not consume the reported intervals"
    assert score == int(res["score"]), f"score of the CIGAR {score} != reported {int(res['score'])}"
    assert (n_match, n_mis, ins_ops, ins_bp, del_ops, del_bp) == tuple(int(res[f]) for f in
                                                                       ("n_match", "n_mismatch", "n_ins_ops", "n_ins_bp", "n_del_ops", "n_del_bp"))
