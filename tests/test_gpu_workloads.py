"""BASELINE configs 3-5 on the CUDA path: bit-exact against the oracle at sizes the oracle's full-band matrix fits,
and — at the configs' full sizes (reads up to 50 kb, bands beyond 16k diagonals, 50 kb consensuses) — the
size-independent consistency properties of tests/consistency.py on every pair."""
import numpy as np
import pytest

from consistency import check_result
from svirlpool_b200 import consensus_align, synth, workloads
from svirlpool_b200._abi import cigartuples

pytestmark = pytest.mark.gpu

INT_FIELDS = ("score", "status", "q_beg", "q_end", "t_beg", "t_end", "n_match", "n_mismatch", "n_ins_ops", "n_ins_bp",
              "n_del_ops", "n_del_bp", "cigar_len", "cells")


@pytest.fixture(scope="module")
def gpu():
    """The engine of this module plans every call as part of a LARGE call (all kernel widths, as in bench.py's steps): by default a
    call of up to 1024 pairs keeps to the KP = 4 / 8 widths (test_config2_one_container_default_planning)."""
    import os
    from svirlpool_b200.aligner import Aligner
    old = os.environ.get("SVP_SMALL_CALL_PAIRS")
    os.environ["SVP_SMALL_CALL_PAIRS"] = "0"
    try:
        a = Aligner(device=0, preset="map-ont", tb_bytes=24 << 30)
    finally:
        if old is None:
            del os.environ["SVP_SMALL_CALL_PAIRS"]
        else:
            os.environ["SVP_SMALL_CALL_PAIRS"] = old
    yield a
    a.close()


def align(engine, b):
    return engine.align_batch(b["packed"], b["off"], b["lens"], b["pairs"], b["cigar_words"])


def assert_parity(b, gpu, oracle):
    gres, gcig = align(gpu, b)
    ores, ocig = align(oracle, b)
    for f in INT_FIELDS:
        bad = np.nonzero(gres[f] != ores[f])[0]
        assert len(bad) == 0, f"{f}: {len(bad)} mismatches, first at pair {bad[0]}: gpu={gres[bad[0]]} oracle={ores[bad[0]]} pair={b['pairs'][bad[0]]}"
    for k in range(len(gres)):
        assert cigartuples(gcig, gres[k]) == cigartuples(ocig, ores[k]), f"CIGAR differs at pair {k}"
    np.testing.assert_allclose(gres["sv_dist"], ores["sv_dist"], rtol=1e-6, atol=0)
    return gres, gcig


def assert_consistent(b, gres, gcig, scoring, allow_status=0):
    assert not (gres["status"] & ~np.uint32(1 | allow_status)).any(), f"unexpected status bits: {np.unique(gres['status'])}"
    for k in range(len(gres)):
        check_result(b["packed"], b["off"], b["lens"], b["pairs"][k], gres[k], gcig, scoring)


@pytest.fixture(scope="module")
def oracle_mt():
    """The oracle on every host thread (full-size workloads: hundreds of 8 kb pairs)."""
    import os
    from oracle_engine import OracleEngine
    return OracleEngine("map-ont", threads=os.cpu_count() or 1)


def test_config2_full_size_parity(gpu, oracle_mt):
    """The headline workload at its real shape: regions 0 and 1 of config 2 (30 reads x 8 kb, 435 pairs each, bands
    |dlen| + 600 at granule 128 — the regions bench.py's CPU arm aligns), every pair bit-exact against the oracle."""
    b = workloads.batch_config2([0, 1])
    assert len(b["pairs"]) == 870 and b["lens"].mean() > 7000
    gres, gcig = assert_parity(b, gpu, oracle_mt)
    assert (gres["status"] == 0).all() and int(gres["cells"].sum()) > 5e9
    assert len({int(w) for w in b["pairs"]["band_w"]}) > 4             # several kernel geometries in one batch


def test_config2_one_container_default_planning(oracle_mt):
    """One container per call, the way the reference's seam is driven, on a handle with the default planning (small calls keep to the
    KP = 4 / 8 widths): region 2 of config 2, every pair bit-exact against the oracle."""
    from svirlpool_b200.aligner import Aligner
    b = workloads.batch_config2([2])
    with Aligner(device=0, preset="map-ont", tb_bytes=8 << 30) as a:
        assert_parity(b, a, oracle_mt)


def test_config5_reduced_parity(gpu, oracle_mt):
    """Trio-shaped loci (Pareto length tail, seeds 40 000+) at reduced size: all-vs-all + consensus -> window pairs."""
    b = workloads.batch_config4([0, 1, 2, 3, 5], seed_base=40_000, skew=True, scale=0.3, max_reads=8, band_w=1024)
    gres, gcig = assert_parity(b, gpu, oracle_mt)
    assert_consistent(b, gres, gcig, gpu.scoring)


def test_config5_full_size_locus_parity(gpu, oracle_mt):
    """One full-size locus of each trio set (seeds 40 000 / 50 000 / 60 000 + 1): every read pair and both strands of every
    padded consensus against its reference window, bit-exact against the oracle."""
    for seed_base in (40_000, 50_000, 60_000):
        b = workloads.batch_config4([1], seed_base=seed_base, skew=True, max_reads=14)
        gres, gcig = assert_parity(b, gpu, oracle_mt)
        assert (gres["status"] & ~np.uint32(1) == 0).all() and gres["score"].max() > 5000


def test_config3_reduced_parity(gpu, oracle):
    b = workloads.batch_config3([0, 1, 2], n_reads=6, scale=0.3)
    gres, gcig = assert_parity(b, gpu, oracle)
    assert_consistent(b, gres, gcig, gpu.scoring)


def test_config3_full_size_parity_long_reads(gpu, oracle):
    """Region 16: haplotypes 21.7 / 24.8 kb (32-bit lanes), band 3.6k: the oracle still fits."""
    b = workloads.batch_config3([16], n_reads=5)
    assert b["lens"].min() > 16_000
    gres, gcig = assert_parity(b, gpu, oracle)
    assert gres["score"].max() > 16_000


def test_config3_full_size_wide_bands_consistency(gpu):
    """Region 6: haplotypes 15.6 / 33.5 kb, |dlen| = 17.9 kb -> bands beyond 16k diagonals on 32-bit lanes = clusters of 4
    CTAs; 8 reads = 28 pairs incl. same-haplotype narrow ones. No oracle at this size: consistency properties."""
    b = workloads.batch_config3([6], n_reads=8)
    assert b["pairs"]["band_w"].max() > 16_384
    gres, gcig = align(gpu, b)
    assert_consistent(b, gres, gcig, gpu.scoring)
    assert (gres["status"] == 0).all() and gres["score"].min() > 5_000


def test_config4_reduced_parity(gpu, oracle):
    b = workloads.batch_config4([0, 1, 2, 3], scale=0.3, max_reads=8, band_w=1024)
    gres, gcig = assert_parity(b, gpu, oracle)
    assert_consistent(b, gres, gcig, gpu.scoring)


def test_config4_full_size_locus(gpu, oracle):
    """Locus 1: core 12.6 kb, padded consensus 52.6 kb vs a 56.6 kb reference window (band 4096, both strands), plus
    the all-vs-all of 10 of its reads: parity for the consensus pairs, consistency for everything."""
    b = workloads.batch_config4([1], max_reads=10)
    gres, gcig = align(gpu, b)
    assert_consistent(b, gres, gcig, gpu.scoring)
    cons_pairs = np.nonzero(b["lens"][b["pairs"]["q"]] > 50_000)[0]
    assert len(cons_pairs) == 2
    sub = dict(b)
    sub["pairs"] = b["pairs"][cons_pairs].copy()
    assert_parity(sub, gpu, oracle)
    best = gres[cons_pairs][np.argmax(gres[cons_pairs]["score"])]
    assert best["score"] > 90_000 and best["q_end"] - best["q_beg"] > 52_000


def test_config5_skewed_locus_through_the_chain(gpu):
    """A long-tail trio locus through align -> annotate -> core interval: the core maps back to the candidate region."""
    lid = next(k for k in range(400) if synth.locus_config4(k, seed_base=40_000, skew=True).region.meta["core"] > 20_000)
    loc = synth.locus_config4(lid, seed_base=40_000, skew=True, max_reads=4)
    cons = workloads.consensus_objects(loc)
    ref = synth.codes_to_str(loc.ref_window)
    fetch = lambda c, s, e: ref[s - loc.window_start:e - loc.window_start]
    recs = consensus_align.align_consensuses_to_reference(gpu, cons, fetch, {loc.chrom: loc.window_start + len(ref)}, min_margin=0)
    ann = consensus_align.annotate_alignments(recs, "S")
    cores = consensus_align.core_intervals_on_reference(cons, ann)
    for cid in cons:
        assert any(abs(a - loc.core_interval[0]) <= 10 for _i, _c, a, _b, _lo, _hi in cores[cid])


def test_config2_full_step_determinism_and_consistency(gpu):
    """One bench-sized step of config 2 (24 regions = 10 440 pairs here, several waves): two runs give identical result
    rows and CIGARs (a checksum of checksums), the CIGAR-free summary call returns the same rows, and a seeded sample of
    the pairs passes the CIGAR/score/count consistency properties."""
    import zlib
    b = workloads.batch_config2(list(range(200, 224)))
    r1, c1 = align(gpu, b)
    r2, c2 = align(gpu, b)
    fields = [f for f in r1.dtype.names]
    assert all(np.array_equal(r1[f], r2[f]) for f in fields)
    used = int(r1["cigar_off"][-1] + r1["cigar_len"][-1])
    assert zlib.crc32(c1[:used].tobytes()) == zlib.crc32(c2[:used].tobytes())
    rs = gpu.align_batch_summary(b["packed"], b["off"], b["lens"], b["pairs"], b["cigar_words"])
    assert all(np.array_equal(r1[f], rs[f]) for f in fields if f != "cigar_off")
    assert (r1["status"] == 0).all() and int(r1["cells"].sum()) > 50e9
    rng = np.random.default_rng(0)
    for k in rng.choice(len(r1), 400, replace=False):
        check_result(b["packed"], b["off"], b["lens"], b["pairs"][k], r1[k], c1, gpu.scoring)
    # same-haplotype pairs carry no large indel, cross-haplotype pairs carry the planted one (sv_dist separates them)
    assert (r1["sv_dist"] == 0).sum() > 0.3 * len(r1) and (r1["sv_dist"] > 40).sum() > 0.3 * len(r1)
