"""The expectations the reference's own pure tests pin on the host logic of this path, one test here per test there:

    tests/test_consensus_lib_pairwise_distance.py   :72-449   (31 tests: size_damping, densities, interior BNDs, directed distance,
                                                                similarity matrix, outlier detection)
    tests/test_consensus_lib_importance_densities.py :32-153  (5 tests)
    tests/test_crs_to_batches.py                     :14-107  (6 tests)

Inputs and asserted values are the reference's; the functions under test are this repo's host mirrors
(svirlpool_b200.consensus_lib, crs_to_batches, consensus._load_crIDs_from_batch_tsv)."""
import numpy as np
import pytest

from svirlpool_b200 import consensus, crs_to_batches
from svirlpool_b200.consensus_lib import (DirectedSignals, _directed_distance, _has_interior_bnd, detect_outlier_reads,
                                           importance_densities_from_ava_signals, pairwise_similarity_matrix, size_damping,
                                           sv_size_densities_from_ava_signals)
from svirlpool_b200.datatypes import SVsignal


def INS(ref_start, ref_end, size, read_start=0, read_end=None):
    return SVsignal(ref_start, ref_end, read_start, read_start + size if read_end is None else read_end, size, 0)


def BND(ref_start, size, sv_type=3):
    return SVsignal(ref_start, ref_start, 0, 0, size, sv_type)


# ---- size_damping (pairwise_distance.py:72-85) -----------------------------------------------------------------------
def test_small_signal_is_strongly_damped():
    assert size_damping(10.0, s0=30.0, alpha=1.5) < 0.20


def test_large_signal_is_nearly_undamped():
    assert size_damping(200.0, s0=30.0, alpha=1.5) > 0.99


def test_transition_is_monotonic():
    v = [size_damping(float(s), s0=30.0, alpha=1.5) for s in range(1, 500)]
    assert all(a <= b for a, b in zip(v, v[1:]))


def test_zero_size_returns_zero():
    assert size_damping(0.0) == 0.0


# ---- importance densities, BNDs excluded and damping (:92-135) ----------------------------------------------------
def test_bnd_signals_are_excluded_from_density():
    track = importance_densities_from_ava_signals({"readA": [BND(500, 300, 3), BND(2000, 500, 4)]}).get("readA")
    assert track is None or track.max() == 0.0


def test_small_signal_damped_versus_large():
    track = importance_densities_from_ava_signals({"readA": [INS(490, 510, 15), INS(2490, 2510, 200)]}, size_densities=None)["readA"]
    assert track[2490:2520].max() > track[490:520].max()


def test_backward_compat_no_size_densities():
    assert importance_densities_from_ava_signals({"readA": [INS(450, 550, 200)]}, size_densities=None)["readA"].max() > 0


# ---- _has_interior_bnd (:138-161) -------------------------------------------------------------------------------------
def test_bnd_at_start_is_not_interior():
    assert not _has_interior_bnd([BND(50, 300, 3)], ref_length=5000, bnd_margin=100)


def test_bnd_at_end_is_not_interior():
    assert not _has_interior_bnd([BND(4950, 300, 4)], ref_length=5000, bnd_margin=100)


def test_bnd_in_middle_is_interior():
    assert _has_interior_bnd([BND(2500, 300, 3)], ref_length=5000, bnd_margin=100)


def test_bnd_exactly_at_margin_is_not_interior():
    assert not _has_interior_bnd([BND(100, 300, 3)], ref_length=5000, bnd_margin=100)


def test_empty_signals():
    assert not _has_interior_bnd([], ref_length=5000)


def test_non_bnd_signals_are_ignored():
    assert not _has_interior_bnd([INS(2500, 2600, 100)], ref_length=5000)


# ---- _directed_distance (:168-199) ------------------------------------------------------------------------------------
def test_bnd_signals_are_excluded():
    assert _directed_distance([BND(500, 1000, 3)], None, None, 30.0, 1.5) == 0.0


def test_basic_ins_contributes():
    assert _directed_distance([INS(500, 600, 100, 50, 150)], None, None, 30.0, 1.5) > 0


def test_importance_weighting_amplifies():
    sig = [INS(500, 600, 100, 50, 150)]
    weighted = _directed_distance(sig, np.full(1000, 2.0), np.full(200, 2.0), 30.0, 1.5)
    np.testing.assert_allclose(weighted, _directed_distance(sig, None, None, 30.0, 1.5) * 4.0, rtol=1e-6)


def test_small_signal_contributes_less():
    assert _directed_distance([INS(500, 510, 12)], None, None, 30.0, 1.5) < _directed_distance([INS(500, 600, 200)], None, None, 30.0, 1.5)


# ---- pairwise_similarity_matrix (:205-372) ------------------------------------------------------------------------------
@pytest.fixture()
def three_reads():
    """A and B similar (small INS both ways), A -> C carries an interior BND, C -> A a small INS, nothing between B and C."""
    small = INS(1000, 1100, 50, 1000, 1050)
    ds = [DirectedSignals("A", "B", [small], 5000), DirectedSignals("B", "A", [small], 5000),
          DirectedSignals("A", "C", [BND(2500, 1000, 3)], 5000), DirectedSignals("C", "A", [small], 5000)]
    return pairwise_similarity_matrix(ds, {}, {"A": 5000, "B": 5000, "C": 5000}, all_read_names=["A", "B", "C"])


def test_diagonal_is_one(three_reads):
    sim, names = three_reads
    assert all(sim[i, i] == 1.0 for i in range(len(names)))


def test_symmetry(three_reads):
    np.testing.assert_array_equal(three_reads[0], three_reads[0].T)


def test_interior_bnd_zeroes_similarity(three_reads):
    sim, names = three_reads
    assert sim[names.index("A"), names.index("C")] == 0.0


def test_no_alignment_gives_zero_similarity(three_reads):
    sim, names = three_reads
    assert sim[names.index("B"), names.index("C")] == 0.0


def test_similar_reads_have_high_similarity(three_reads):
    sim, names = three_reads
    assert sim[names.index("A"), names.index("B")] > 0.3


def test_length_difference_reduces_similarity():
    sig = INS(1000, 1100, 100, 1000, 1100)
    eq, _ = pairwise_similarity_matrix([DirectedSignals("X", "Y", [sig], 5000), DirectedSignals("Y", "X", [sig], 5000)], {},
                                       {"X": 5000, "Y": 5000}, all_read_names=["X", "Y"])
    uneq, _ = pairwise_similarity_matrix([DirectedSignals("X", "Y", [sig], 5000), DirectedSignals("Y", "X", [sig], 10000)], {},
                                         {"X": 10000, "Y": 5000}, all_read_names=["X", "Y"])
    assert eq[0, 1] > uneq[0, 1]


def test_bnd_at_read_edge_does_not_zero_similarity():
    ins = INS(1000, 1100, 100, 1000, 1100)
    ds = [DirectedSignals("A", "B", [BND(50, 300, 3), ins], 5000), DirectedSignals("B", "A", [ins], 5000)]
    sim, _ = pairwise_similarity_matrix(ds, {}, {"A": 5000, "B": 5000}, all_read_names=["A", "B"])
    assert sim[0, 1] > 0.0


def test_empty_directed_signals():
    sim, _ = pairwise_similarity_matrix([], {}, {"A": 5000, "B": 5000}, all_read_names=["A", "B"])
    assert sim[0, 1] == 0.0 and sim[0, 0] == 1.0


# ---- detect_outlier_reads (:378-449) ------------------------------------------------------------------------------------
def half_connected(n):
    sim = np.full((n, n), 0.5)
    np.fill_diagonal(sim, 1.0)
    return sim


def disconnect(sim, k):
    sim[k, :] = 0.0
    sim[:, k] = 0.0
    sim[k, k] = 1.0
    return sim


def test_uniform_reads_no_outliers():
    names = [f"r{i}" for i in range(10)]
    assert detect_outlier_reads(half_connected(10), names, dict.fromkeys(names, 5000), n_clusters=2) == set()


def test_length_only_outlier_is_flagged():
    names = ["r0", "r1", "r2", "r3", "r4"]
    rl = {"r0": 5000, "r1": 5000, "r2": 5000, "r3": 5000, "r4": 50000}
    assert "r4" in detect_outlier_reads(half_connected(5), names, rl, n_clusters=2)


def test_both_criteria_flags_outlier():
    names = ["r0", "r1", "r2", "r3", "r4"]
    rl = {"r0": 5000, "r1": 5000, "r2": 5000, "r3": 5000, "r4": 50000}
    assert "r4" in detect_outlier_reads(disconnect(half_connected(5), 4), names, rl, n_clusters=2)


def test_empty_input():
    assert detect_outlier_reads(np.zeros((0, 0)), [], {}, n_clusters=1) == set()


def test_coherent_length_groups_not_flagged():
    a, b = [f"a{i}" for i in range(18)], [f"b{i}" for i in range(10)]
    names = ["outlier"] + a + b
    rl = {"outlier": 162, **{r: 1200 + 2 * i for i, r in enumerate(a)}, **{r: 1660 + 2 * i for i, r in enumerate(b)}}
    out = detect_outlier_reads(half_connected(len(names)), names, rl, n_clusters=1)
    assert "outlier" in out and not (out & set(a + b))


def test_low_connectivity_outlier_median_based():
    names = [f"r{i}" for i in range(10)]
    assert "r9" in detect_outlier_reads(disconnect(half_connected(10), 9), names, dict.fromkeys(names, 5000), n_clusters=2)


# ---- importance densities with size scaling (importance_densities.py:32-153) --------------------------------------------
def ins0(ref_start, ref_end, size):
    return SVsignal(ref_start, ref_end, 0, size, size, 0)


COMMON_AND_RARE = {"readA": [ins0(450, 550, 100), ins0(650, 750, 100), ins0(850, 950, 100), ins0(2450, 2550, 500)]}


def test_importance_densities_without_size_scaling_uses_unit_scale():
    track = importance_densities_from_ava_signals({"readA": [ins0(450, 550, 500), ins0(2450, 2550, 500)]}, size_densities=None)["readA"]
    np.testing.assert_allclose(track[480:520].max(), track[2480:2520].max(), rtol=1e-6)


def test_importance_densities_size_scaling_reduces_rare_size_contribution():
    sd = sv_size_densities_from_ava_signals(COMMON_AND_RARE)
    unscaled = importance_densities_from_ava_signals(COMMON_AND_RARE, size_densities=None)["readA"]
    scaled = importance_densities_from_ava_signals(COMMON_AND_RARE, size_densities=sd)["readA"]
    assert scaled[2480:2520].max() < unscaled[2480:2520].max()


def test_importance_densities_size_scaling_preserves_common_size_contribution():
    sd = sv_size_densities_from_ava_signals(COMMON_AND_RARE)
    unscaled = importance_densities_from_ava_signals(COMMON_AND_RARE, size_densities=None)["readA"]
    scaled = importance_densities_from_ava_signals(COMMON_AND_RARE, size_densities=sd)["readA"]
    np.testing.assert_allclose(unscaled[480:520].max(), scaled[480:520].max(), rtol=1e-6)


def test_importance_densities_returns_empty_for_empty_input():
    assert importance_densities_from_ava_signals({}, size_densities=None) == {}


def test_importance_densities_bnd_signals_always_use_unit_scale():
    sig = {"readA": [ins0(200, 300, 100), SVsignal(1000, 1000, 0, 0, 300, 3)]}
    sd = sv_size_densities_from_ava_signals(sig)
    unscaled = importance_densities_from_ava_signals(sig, size_densities=None)["readA"]
    scaled = importance_densities_from_ava_signals(sig, size_densities=sd)["readA"]
    np.testing.assert_allclose(unscaled[980:1020].max(), scaled[980:1020].max(), rtol=1e-6)


# ---- crs_to_batches (test_crs_to_batches.py:14-107) -------------------------------------------------------------------
def test_build_batches_groups_adjacent_containers():
    ext = {1: (["chr1"], 100, 200), 2: (["chr1"], 300, 400), 3: (["chr1"], 500, 600), 4: (["chr2"], 100, 200)}
    batches = crs_to_batches.build_batches(extents=ext, batch_size=100, max_batch_span=10_000)
    assert len(batches) == 2
    assert [b for b in batches if b["chr"] == "chr1"][0]["crIDs"] == [1, 2, 3]
    assert [b for b in batches if b["chr"] == "chr2"][0]["crIDs"] == [4]


def test_build_batches_respects_batch_size_cap():
    ext = {i: (["chr1"], i * 100, i * 100 + 50) for i in range(1, 6)}
    assert sorted(len(b["crIDs"]) for b in crs_to_batches.build_batches(extents=ext, batch_size=2, max_batch_span=10_000)) == [1, 2, 2]


def test_build_batches_respects_max_span():
    ext = {1: (["chr1"], 100, 200), 2: (["chr1"], 5_000, 5_100), 3: (["chr1"], 9_000, 9_100)}
    batches = crs_to_batches.build_batches(extents=ext, batch_size=100, max_batch_span=5_000)
    assert sorted(tuple(b["crIDs"]) for b in batches) == [(1, 2), (3,)]


def test_wide_container_becomes_singleton_batch():
    ext = {1: (["chr1"], 100, 200), 2: (["chr1", "chr2"], 100, 200), 3: (["chr3"], 0, 100_000_000), 4: (["chr1"], 300, 400)}
    batches = crs_to_batches.build_batches(extents=ext, batch_size=100, max_batch_span=1_000_000)
    singles = {tuple(b["crIDs"]) for b in batches if len(b["crIDs"]) == 1}
    assert (2,) in singles and (3,) in singles
    local = [b for b in batches if b["chr"] == "chr1" and len(b["crIDs"]) > 1]
    assert len(local) == 1 and local[0]["crIDs"] == [1, 4]


def test_build_batches_assigns_unique_sequential_ids():
    ext = {i: (["chr1"], i * 100, i * 100 + 50) for i in range(1, 8)}
    ids = [b["batch_id"] for b in crs_to_batches.build_batches(extents=ext, batch_size=3, max_batch_span=10_000)]
    assert ids == sorted(set(ids)) and ids[0] == 0


def test_write_outputs_roundtrips_through_consensus_loader(tmp_path):
    ext = {1: (["chr1"], 100, 200), 2: (["chr1"], 300, 400), 7: (["chr2"], 100, 200)}
    batches = crs_to_batches.build_batches(extents=ext, batch_size=100, max_batch_span=10_000)
    tsv, ids_file = tmp_path / "batches.tsv", tmp_path / "ids.txt"
    crs_to_batches.write_outputs(batches=batches, tsv_path=tsv, ids_path=ids_file)
    for b in batches:
        assert consensus._load_crIDs_from_batch_tsv(tsv, b["batch_id"]) == b["crIDs"]
    listed = [int(line) for line in ids_file.read_text().split()]
    assert sorted(listed) == sorted(b["batch_id"] for b in batches)
