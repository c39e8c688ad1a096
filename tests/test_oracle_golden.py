"""The CPU oracle (ALIGN_SPEC v1) against the reference's frozen minimap2 alignments
(tests/golden/minimap2_fixtures.json.gz, lifted from /root/reference/tests/data by make_golden.py).

Two levels, reported separately:
  1. score semantics on all 20 records: AS == score of the recorded CIGAR under the map-ont scoring;
  2. re-alignment of the 14 records whose inputs exist: strand, reference start and CIGAR identical
     (clip kind S/H is minimap2's primary/supplementary bookkeeping and is normalised away).
"""
import re

import pytest

from svirlpool_b200 import ava
from svirlpool_b200.datatypes import cigar_from_string


def cigar_score(cigar: str, nm: int) -> int:
    ops = cigar_from_string(cigar)
    n_m = sum(n for op, n in ops if op == 0)
    gaps = [n for op, n in ops if op in (1, 2)]
    mism = nm - sum(gaps)
    return 2 * (n_m - mism) - 4 * mism - sum(min(4 + 2 * g, 24 + g) for g in gaps)


def test_as_formula_all_20_records(golden):
    recs = [r for c in golden["cases"] for r in c["records"]] + golden["score_only"]
    assert len(recs) == 20
    for r in recs:
        assert cigar_score(r["cigar"], r["NM"]) == r["AS"], r["read"]


def norm(cigar: str) -> str:
    return cigar.replace("H", "S")


@pytest.mark.parametrize("case_idx", range(7))
def test_realign_fixture_case(golden, oracle, case_idx):
    case = golden["cases"][case_idx]
    recs = ava.map_queries(oracle, case["reads"], "ref", case["reference"])
    got = sorted((r.query_name, r.reference_start, r.is_reverse, norm(r.cigarstring), r.tags["AS"], r.tags["NM"]) for r in recs)
    want = sorted((r["read"], r["ref_start"], bool(r["flag"] & 16), norm(r["cigar"]), r["AS"], r["NM"]) for r in case["records"])
    assert got == want, case["name"]


def test_left_aligned_gap_tiebreak(golden, oracle):
    """simulated.with_insertion pins the tie-break: 20 A's inserted after an A are reported as
    499M20I501M (gap shifted left), tests/data/consensus_class/simulated.with_insertion.json.gz."""
    case = next(c for c in golden["cases"] if c["name"] == "simulated.with_insertion")
    recs = ava.map_queries(oracle, case["reads"], "ref", case["reference"])
    assert [r.cigarstring for r in recs] == ["499M20I501M"]
    assert re.fullmatch(r"\d+M20I\d+M", recs[0].cigarstring)
