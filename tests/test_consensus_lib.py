"""Host logic downstream of the aligner: the assertions the reference pins in
tests/test_consensus_lib_pairwise_distance.py (:66-449) and tests/test_consensus_lib_importance_densities.py,
restated against svirlpool_b200.consensus_lib, plus a scalar loop restatement of
pairwise_similarity_matrix (consensus_lib.py:477-592) used as a second opinion for the vectorised code."""
import math

import numpy as np
import pytest

from svirlpool_b200.consensus_lib import (DirectedSignals, _directed_distance, _has_interior_bnd, detect_outlier_reads,
                                           importance_densities_from_ava_signals, pairwise_similarity_matrix, size_damping,
                                           sv_size_densities_from_ava_signals)
from svirlpool_b200.datatypes import SVsignal


def ins(ref_start, ref_end, size, read_start=0, read_end=None):
    return SVsignal(ref_start, ref_end, read_start, read_start + size if read_end is None else read_end, size, 0)


def dele(ref_start, ref_end, size, read_start=0):
    return SVsignal(ref_start, ref_end, read_start, read_start, size, 1)


def bnd(ref_start, size, sv_type=3):
    return SVsignal(ref_start, ref_start, 0, 0, size, sv_type)


# ---- size_damping ----------------------------------------------------------------------------
def test_size_damping_shape():
    assert size_damping(10.0) < 0.20 and size_damping(200.0) > 0.99 and size_damping(0.0) == 0.0
    vals = [size_damping(float(s)) for s in range(1, 500)]
    assert all(a <= b for a, b in zip(vals, vals[1:]))
    assert size_damping(30.0) == pytest.approx(1 - math.exp(-1.0))


# ---- densities -------------------------------------------------------------------------------
def test_importance_density_ignores_bnd_and_damps_small():
    only_bnd = importance_densities_from_ava_signals({"a": [bnd(500, 300, 3), bnd(2000, 500, 4)]})
    assert only_bnd["a"].max() == 0.0 and len(only_bnd["a"]) == 2000
    tr = importance_densities_from_ava_signals({"a": [ins(490, 510, 15), ins(2490, 2510, 200)]})["a"]
    assert tr[2490:2520].max() > tr[490:520].max() > 0
    assert importance_densities_from_ava_signals({}) == {}


def test_importance_density_unit_scale_and_size_scaling():
    two = importance_densities_from_ava_signals({"a": [ins(450, 550, 500), ins(2450, 2550, 500)]})["a"]
    np.testing.assert_allclose(two[480:520].max(), two[2480:2520].max(), rtol=1e-6)
    sig = {"a": [ins(450, 550, 100), ins(650, 750, 100), ins(850, 950, 100), ins(2450, 2550, 500)]}
    sd = sv_size_densities_from_ava_signals(sig)
    assert set(sd) == {0, 1} and len(sd[0]["a"]) == 501 and sd[1] == {}
    plain = importance_densities_from_ava_signals(sig)["a"]
    scaled = importance_densities_from_ava_signals(sig, size_densities=sd)["a"]
    assert scaled[2480:2520].max() < plain[2480:2520].max()          # rare size is reduced
    np.testing.assert_allclose(scaled[480:520].max(), plain[480:520].max(), rtol=1e-6)   # common size keeps scale 1


def test_importance_density_bump_support_and_value():
    # one 100 bp INS centred at 1000: plateau 950..1050, sigmoid edges (falloff 0.5), damping(100)
    tr = importance_densities_from_ava_signals({"a": [ins(1000, 1000, 100), ins(3000, 3000, 100)]})["a"]
    assert tr[:938].max() == 0.0 and tr[1063:2900].max() == 0.0
    centre = 1 / (1 + math.exp(-0.5 * 50)) * 1 / (1 + math.exp(0.5 * -50))
    assert tr[1000] == pytest.approx(size_damping(100) * centre, rel=1e-12)


# ---- BND / directed distance -----------------------------------------------------------------
@pytest.mark.parametrize("pos,expected", [(50, False), (4950, False), (2500, True), (100, False), (101, True), (4899, True), (4900, False)])
def test_has_interior_bnd_margins(pos, expected):
    assert _has_interior_bnd([bnd(pos, 300)], ref_length=5000, bnd_margin=100) is expected


def test_has_interior_bnd_ignores_indels():
    assert not _has_interior_bnd([], 5000) and not _has_interior_bnd([ins(2500, 2600, 100)], 5000)


def test_directed_distance_weights():
    assert _directed_distance([bnd(500, 1000)], None, None, 30.0, 1.5) == 0.0
    s = [ins(500, 600, 100, read_start=50, read_end=150)]
    base = _directed_distance(s, None, None, 30.0, 1.5)
    assert base == pytest.approx(size_damping(100) * 100)
    four = _directed_distance(s, np.full(1000, 2.0), np.full(200, 2.0), 30.0, 1.5)
    np.testing.assert_allclose(four, 4.0 * base, rtol=1e-6)
    assert _directed_distance(s, np.full(1000, 2.0), np.full(200, 2.0), 30.0, 1.5, densities_weight=0.0) == pytest.approx(base)
    assert _directed_distance(s, np.full(10, 2.0), np.full(10, 2.0), 30.0, 1.5) == pytest.approx(base)   # out of range -> weight 1
    assert _directed_distance([ins(500, 510, 12)], None, None, 30.0, 1.5) < _directed_distance([ins(500, 600, 200)], None, None, 30.0, 1.5)


# ---- similarity matrix -----------------------------------------------------------------------
def three_reads():
    s = ins(1000, 1100, 50, read_start=1000, read_end=1050)
    ds = [DirectedSignals("A", "B", [s], 5000), DirectedSignals("B", "A", [s], 5000),
          DirectedSignals("A", "C", [bnd(2500, 1000)], 5000), DirectedSignals("C", "A", [s], 5000)]
    return ds, {"A": 5000, "B": 5000, "C": 5000}


def test_similarity_matrix_structure():
    ds, rl = three_reads()
    sim, names = pairwise_similarity_matrix(ds, {}, rl, all_read_names=["A", "B", "C"])
    assert names == ["A", "B", "C"] and sim.dtype == np.float64
    assert np.all(np.diag(sim) == 1.0)
    np.testing.assert_array_equal(sim, sim.T)
    assert sim[0, 2] == 0.0          # interior BND in A->C
    assert sim[1, 2] == 0.0          # no alignment between B and C
    assert sim[0, 1] > 0.3


def test_similarity_length_penalty_and_edge_bnd():
    s = ins(1000, 1100, 100, read_start=1000, read_end=1100)
    eq, _ = pairwise_similarity_matrix([DirectedSignals("X", "Y", [s], 5000), DirectedSignals("Y", "X", [s], 5000)], {},
                                       {"X": 5000, "Y": 5000}, all_read_names=["X", "Y"])
    un, _ = pairwise_similarity_matrix([DirectedSignals("X", "Y", [s], 5000), DirectedSignals("Y", "X", [s], 10000)], {},
                                       {"X": 10000, "Y": 5000}, all_read_names=["X", "Y"])
    assert eq[0, 1] > un[0, 1]
    edge, _ = pairwise_similarity_matrix([DirectedSignals("A", "B", [bnd(50, 300), s], 5000), DirectedSignals("B", "A", [s], 5000)], {},
                                         {"A": 5000, "B": 5000}, all_read_names=["A", "B"])
    assert edge[0, 1] > 0.0
    empty, _ = pairwise_similarity_matrix([], {}, {"A": 5000, "B": 5000}, all_read_names=["A", "B"])
    assert empty[0, 1] == 0.0 and empty[0, 0] == 1.0


def scalar_similarity(directed, densities, read_lengths, names, lam=None, dw=1.0):
    """Loop restatement of consensus_lib.py:477-592 (test-only second opinion)."""
    n = len(names)
    idx = {r: k for k, r in enumerate(names)}
    dsum = [[0.0] * n for _ in range(n)]
    cnt = [[0] * n for _ in range(n)]
    bndm = [[False] * n for _ in range(n)]
    for ds in directed:
        q, r = idx.get(ds.query_name), idx.get(ds.ref_name)
        if q is None or r is None:
            continue
        dsum[q][r] += _directed_distance(ds.signals, densities.get(ds.ref_name), densities.get(ds.query_name), 30.0, 1.5, dw)
        cnt[q][r] += 1
        bndm[q][r] = bndm[q][r] or _has_interior_bnd(ds.signals, ds.ref_length, 100)
    dist = [[math.inf] * n for _ in range(n)]
    for a in range(n):
        dist[a][a] = 0.0
        for b in range(a + 1, n):
            if cnt[a][b] or cnt[b][a]:
                dab = dsum[a][b] / cnt[a][b] if cnt[a][b] else 0.0
                dba = dsum[b][a] / cnt[b][a] if cnt[b][a] else 0.0
                dist[a][b] = dist[b][a] = max(dab, dba)
    fin = [x for row in dist for x in row if math.isfinite(x) and x > 0]
    sigma = float(np.median(fin)) if fin else 1.0
    sigma = sigma or 1.0
    lam = sigma if lam is None else lam
    sim = np.zeros((n, n))
    for a in range(n):
        sim[a, a] = 1.0
        for b in range(a + 1, n):
            if math.isfinite(dist[a][b]):
                la, lb = float(read_lengths.get(names[a], 0)), float(read_lengths.get(names[b], 0))
                hi = max(la, lb)
                d = dist[a][b] + lam * (1.0 - (min(la, lb) / hi if hi > 0 else 0.0))
                sim[a, b] = sim[b, a] = math.exp(-(d ** 2) / (2.0 * sigma ** 2))
            if bndm[a][b] or bndm[b][a]:
                sim[a, b] = sim[b, a] = 0.0
    return sim


@pytest.mark.parametrize("seed", range(6))
def test_vectorised_matrix_matches_scalar_restatement(seed):
    rng = np.random.default_rng(seed)
    names = [f"r{k}" for k in range(int(rng.integers(2, 12)))]
    lengths = {r: int(rng.integers(0, 9000)) for r in names}
    directed = []
    for _ in range(int(rng.integers(0, 60))):
        q, r = rng.choice(names, 2, replace=True)
        sigs = []
        for _ in range(int(rng.integers(0, 4))):
            t = int(rng.choice([0, 1, 3, 4]))
            pos, size = int(rng.integers(0, 8000)), int(rng.integers(12, 900))
            sigs.append(SVsignal(pos, pos + (size if t == 1 else 0), pos, pos + (size if t == 0 else 0), size, t))
        directed.append(DirectedSignals(str(q), str(r), sigs, int(rng.integers(1000, 9000))))
    dens = {r: rng.random(int(rng.integers(1, 9000))) * 2 for r in names[::2]}
    for dw in (0.0, 1.0):
        got, order = pairwise_similarity_matrix(directed, dens, lengths, all_read_names=names, densities_weight=dw)
        want = scalar_similarity(directed, dens, lengths, names, dw=dw)
        assert order == names
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)


def test_matrix_read_names_default_is_sorted_union():
    sim, names = pairwise_similarity_matrix([DirectedSignals("z", "b", [], 10)], {}, {"m": 5})
    assert names == ["b", "m", "z"] and sim.shape == (3, 3)


# ---- outliers --------------------------------------------------------------------------------
def uniform(n, value=0.5):
    sim = np.full((n, n), value)
    np.fill_diagonal(sim, 1.0)
    return sim


def test_outliers_none_when_uniform_and_empty_input():
    names = [f"r{k}" for k in range(10)]
    assert detect_outlier_reads(uniform(10), names, dict.fromkeys(names, 5000), n_clusters=2) == set()
    assert detect_outlier_reads(np.zeros((0, 0)), [], {}, n_clusters=1) == set()


def test_outliers_by_length_or_connectivity():
    names = ["r0", "r1", "r2", "r3", "r4"]
    rl = {"r0": 5000, "r1": 5000, "r2": 5000, "r3": 5000, "r4": 50000}
    assert "r4" in detect_outlier_reads(uniform(5), names, rl, n_clusters=2)
    names10 = [f"r{k}" for k in range(10)]
    sim = uniform(10)
    sim[9, :] = sim[:, 9] = 0.0
    sim[9, 9] = 1.0
    assert detect_outlier_reads(sim, names10, dict.fromkeys(names10, 5000), n_clusters=2) == {"r9"}


def test_coherent_length_groups_are_kept():
    a, b = [f"a{k}" for k in range(18)], [f"b{k}" for k in range(10)]
    names = ["solo"] + a + b
    rl = {"solo": 162, **{r: 1200 + 2 * k for k, r in enumerate(a)}, **{r: 1660 + 2 * k for k, r in enumerate(b)}}
    assert detect_outlier_reads(uniform(len(names)), names, rl, n_clusters=1) == {"solo"}


def test_similarity_from_summary_equals_record_pipeline(oracle):
    """The CIGAR-free path (svp_result rows -> similarity / alignment-length matrices) is bit-identical to the reference's
    record pipeline with unit density weights, for known strands and for both-orientation (unknown strand) batches."""
    from svirlpool_b200 import ava, consensus, consensus_lib, synth
    for rid, strands_known in ((2, True), (5, False), (9, True)):
        region = synth.region_config2(rid, n_reads=9, mean_len=1500, sd_len=80)
        reads = region.read_strings()
        strands = dict(zip(region.names, region.strands)) if strands_known else None
        lengths = {n: len(s) for n, s in reads.items()}
        recs = ava.ava_align(oracle, reads, strands=strands)
        directed = consensus_lib.parse_directed_signals_from_ava(recs, 12, 100)
        sim_a, names_a = consensus_lib.pairwise_similarity_matrix(directed, {}, lengths, all_read_names=list(reads), densities_weight=0.0)
        len_a = consensus.pairwise_alignment_lengths(recs, list(reads))
        names, qi, ti, res = ava.ava_summary(oracle, reads, strands=strands)
        sim_b, len_b = consensus_lib.similarity_from_summary(names, lengths, qi, ti, res)
        assert names == names_a
        np.testing.assert_array_equal(sim_a, sim_b)
        np.testing.assert_array_equal(len_a, len_b)
    # a clipped alignment whose start lies inside the aligned span's margins marks the pair as unrelated (interior BND)
    from svirlpool_b200._abi import RESULT_DTYPE
    res = np.zeros(2, dtype=RESULT_DTYPE)
    res["score"], res["cigar_len"] = 500, 3
    res["q_beg"], res["q_end"], res["t_beg"], res["t_end"] = [150, 0], [900, 800], [120, 0], [900, 800]
    sim, _ = consensus_lib.similarity_from_summary(["a", "b", "c"], {"a": 1000, "b": 1000, "c": 1000}, np.array([0, 0]), np.array([1, 2]), res)
    assert sim[0, 1] == 0.0 and sim[0, 2] > 0.0


def test_parse_sv_signals_from_ava_alignments_pools_by_reference_read():
    from svirlpool_b200 import consensus_lib
    """a3 (consensus_lib.py:40-90): signals are collected under the REFERENCE read of every alignment, over all queries aligned
    to it; unmapped records and records without CIGAR are skipped; alignments without signals leave no key."""
    from svirlpool_b200.datatypes import AlignedRecord
    recs = [
        AlignedRecord("a", "b", 100, [(0, 500), (2, 40), (0, 300), (1, 25), (0, 200)]),                  # DEL 40 @600, INS 25 @940 on b
        AlignedRecord("c", "b", 0, [(4, 150), (0, 400), (1, 10), (0, 100)]),                               # BNDL 150 on b; INS 10 is below 12
        AlignedRecord("a", "c", 10, [(0, 1000)]),                                                           # no signal: no key for c
        AlignedRecord("d", "a", 5, [(0, 300), (2, 12), (0, 300), (5, 120)], flag=16),                      # reverse strand: DEL 12 + BNDR 120 on a
        AlignedRecord("u", None, 0, None, flag=4),
    ]
    out = consensus_lib.parse_sv_signals_from_ava_alignments(recs, 12, 100)
    assert set(out) == {"a", "b"}
    b = sorted(out["b"])
    assert [(s.ref_start, s.ref_end, s.size, s.sv_type) for s in b] == [(0, 0, 150, 3), (600, 640, 40, 1), (940, 940, 25, 0)]
    a = out["a"]
    assert [(s.ref_start, s.ref_end, s.size, s.sv_type) for s in a] == [(617, 617, 120, 4), (305, 317, 12, 1)]
    # the INS on b: read coordinates in original read orientation (forward: leading clip 0 + 500 + 300 consumed)
    ins = next(s for s in b if s.sv_type == 0)
    assert (ins.read_start, ins.read_end) == (800, 825)
    # directed view of the same records: one entry per mapped alignment, ref_length = aligned reference span
    ds = consensus_lib.parse_directed_signals_from_ava(recs, 12, 100)
    assert [(d.query_name, d.ref_name, d.ref_length) for d in ds] == [("a", "b", 1040), ("c", "b", 500), ("a", "c", 1000), ("d", "a", 612)]
