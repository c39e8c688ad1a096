"""The C-ABI library loads and exports every symbol include/svp_align.h declares (no compute calls:
there is no GPU here), struct layouts match the header, and the product refuses to run without CUDA."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from svirlpool_b200 import _abi, aligner

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "svp_align.h").read_text()


def declared_functions():
    return sorted(set(re.findall(r"\b(svp_[a-z_0-9]+)\s*\(", HEADER)))


def test_header_functions_match_export_list():
    assert declared_functions() == sorted(_abi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = aligner.load_library()
    for name in declared_functions():
        assert getattr(lib, name) is not None
    assert lib.svp_abi_version() == _abi.ABI_VERSION == int(re.search(r"#define SVP_ABI_VERSION (\d+)", HEADER).group(1))


def test_struct_sizes_and_presets():
    lib = aligner.load_library()
    assert (_abi.SCORING_DTYPE.itemsize, _abi.PAIR_DTYPE.itemsize, _abi.RESULT_DTYPE.itemsize, _abi.STATS_DTYPE.itemsize) == (112, 32, 72, 56)
    s = aligner.scoring_preset("map-ont", lib)[0]
    assert s["matrix"].reshape(4, 4).diagonal().tolist() == [2, 2, 2, 2] and s["matrix"][1] == -4
    assert (s["del_open1"], s["del_ext1"], s["del_open2"], s["del_ext2"], s["zdrop"], s["min_signal_size"]) == (4, 2, 24, 1, 400, 12)
    assert aligner.scoring_preset("last", lib)[0]["del_open2"] == -1
    with pytest.raises(ValueError):
        aligner.scoring_preset("nope", lib)


def test_oracle_and_cuda_presets_agree(oracle):
    lib = aligner.load_library()
    for name in ("map-ont", "ava-ont", "last"):
        from oracle_engine import OracleEngine
        assert OracleEngine(name).scoring.tobytes() == aligner.scoring_preset(name, lib).tobytes()


def test_no_cpu_fallback_without_device():
    lib = aligner.load_library()
    if lib.svp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(aligner.AlignerError) as e:
        aligner.Aligner(device=0)
    assert e.value.returncode == _abi.SVP_ERR_NODEVICE and "no CPU fallback" in str(e.value)


def test_pack_2bit_layout():
    packed, off, lens = _abi.pack_2bit(["ACGTAC", "TTTT", "n" * 5, "ACGNNRAC"])
    # the third and fourth sequences hold letters outside ACGT: a 1-bit-per-base mask follows their codes, flagged in seq_off
    assert off.tolist() == [0, 16, 32 | _abi.SEQ_NMASK, 64 | _abi.SEQ_NMASK] and lens.tolist() == [6, 4, 5, 8]
    assert packed[0] == 0b11100100 and packed[1] == 0b0100 and packed[16] == 0xFF and packed[32] == 0 and packed[48] == 0b11111
    assert packed[64] == 0b00100100 and packed[65] == 0b01000000 and packed[80] == 0b00111000 and len(packed) == 96 + 16
    codes = [np.array([0, 1, 2, 3, 0, 1], dtype=np.uint8), np.array([3, 3, 3, 3], dtype=np.uint8)]
    p2, o2, l2 = _abi.pack_codes(codes)
    assert p2[0] == packed[0] and p2[16] == 0xFF and o2.tolist() == [0, 16] and l2.tolist() == [6, 4]


# ---------------------------------------------------------------------------------------------
# host-side builders of the library (svp_pack_2bit, svp_ava_pairs): no device needed
# ---------------------------------------------------------------------------------------------

def _random_reads(rng, n, lo, hi):
    alphabet = np.array(list("ACGTacgtN"))
    return ["".join(alphabet[rng.integers(0, len(alphabet), int(rng.integers(lo, hi)))]) for _ in range(n)]


def test_native_pack_2bit_equals_numpy_packer():
    rng = np.random.default_rng(5)
    seqs = _random_reads(rng, 40, 1, 700) + ["A", "ACGT" * 16, "T" * 63, "G" * 64, "C" * 65]
    p1, o1, l1 = _abi.pack_2bit(seqs)
    p2, o2, l2 = aligner.pack_2bit_native(seqs)
    assert np.array_equal(o1, o2) and np.array_equal(l1, l2) and p1.tobytes() == p2.tobytes()
    p3, o3, l3 = aligner.pack_2bit_native([])
    assert len(o3) == 0 and len(p3) == 16


@pytest.mark.parametrize("strands_known", [False, True])
@pytest.mark.parametrize("with_offsets", [False, True])
@pytest.mark.parametrize("full", [False, True])
def test_native_ava_pairs_equal_python_builder(strands_known, with_offsets, full):
    """svp_ava_pairs + svp_pack_2bit (PairBatch.build_ava) produce exactly the buffers of ava_pair_specs() +
    PairBatch.build(): same pairs, order, bands, flags and CIGAR slots."""
    from svirlpool_b200 import ava
    rng = np.random.default_rng(11 + 2 * strands_known + with_offsets)
    seqs = _random_reads(rng, 13, 200, 3000)
    names = [f"read_{int(k)}" for k in rng.permutation(len(seqs))]           # name order != table order
    strands = {n: "+-"[int(rng.integers(2))] for n in names} if strands_known else None
    offsets = {n: int(rng.integers(0, 400)) for n in names} if with_offsets else None
    specs, meta = ava.ava_pair_specs(names, [len(s) for s in seqs], strands, offsets, 300, 64, full)
    ref = ava.PairBatch.build(names, seqs, specs)
    got, gmeta = ava.PairBatch.build_ava(names, seqs, strands, offsets, 300, 64, full)
    assert gmeta == [(q, t, bool(rc)) for q, t, rc in meta]
    assert got.pairs.tobytes() == ref.pairs.tobytes() and got.cigar_words == ref.cigar_words
    assert got.packed.tobytes() == ref.packed.tobytes() and np.array_equal(got.seq_off, ref.seq_off)


def test_native_ava_pairs_argument_errors():
    lens = np.array([100, 200, 300], dtype=np.uint32)
    with pytest.raises(aligner.AlignerError):
        aligner.ava_pairs_native(lens, None, None, None, 0, 3, 300, 48, False)        # granule must be a multiple of 32
    with pytest.raises(aligner.AlignerError):
        aligner.ava_pairs_native(lens, np.array([0, 1, 7], dtype=np.uint32), None, None, 0, 3, 300, 32, False)   # order index out of range
    pairs, words = aligner.ava_pairs_native(lens, None, np.array([1, -1, 1], dtype=np.int8), None, 0, 3, 300, 32, False, cigar_start=10)
    assert len(pairs) == 3 and pairs["cigar_off"][0] == 10 and words == 10 + int(pairs["cigar_cap"].sum())
    assert pairs["flags"].tolist() == [1, 0, 1]


def test_native_ava_pairs_several_containers_in_one_table():
    """A batch of containers the way INTEGRATION.md describes it: one sequence table over all reads, svp_ava_pairs per
    container with its seq_base and the running CIGAR word count == the per-container Python builds, shifted."""
    from svirlpool_b200 import ava
    rng = np.random.default_rng(21)
    containers = [_random_reads(rng, n, 300, 1500) for n in (5, 2, 9)]
    all_seqs = [s for c in containers for s in c]
    packed, off, lens = aligner.pack_2bit_native(all_seqs)
    base, words, got = 0, 0, []
    for reads in containers:
        pairs, words = aligner.ava_pairs_native(lens, None, np.ones(len(reads), dtype=np.int8), None, base, len(reads), 300, 32, False, cigar_start=words)
        got.append(pairs)
        base += len(reads)
    base, slot = 0, 0
    for reads, pairs in zip(containers, got):
        names = [f"r{k:03d}" for k in range(len(reads))]                     # table order == name order
        specs, _ = ava.ava_pair_specs(names, [len(s) for s in reads], {n: "+" for n in names}, None, 300, 32, False)
        ref = ava.PairBatch.build(names, reads, specs).pairs
        assert np.array_equal(pairs["q"], ref["q"] + base) and np.array_equal(pairs["t"], ref["t"] + base)
        for f in ("band_lo", "band_w", "flags", "cigar_cap"):
            assert np.array_equal(pairs[f], ref[f]), f
        assert np.array_equal(pairs["cigar_off"], ref["cigar_off"] + slot)
        base += len(reads)
        slot += int(ref["cigar_cap"].sum())
    assert words == slot


# ---------------------------------------------------------------------------------------------
# the planner, host only (svp_plan_pairs): which kernel family / geometry a pair gets
# ---------------------------------------------------------------------------------------------

def _plan(lens, specs, preset="map-ont", scoring=None):
    from svirlpool_b200 import ava
    lens = np.asarray(lens, dtype=np.uint32)
    nbytes = ((lens.astype(np.uint64) + 3) // 4 + 15) // 16 * 16
    off = np.concatenate(([0], np.cumsum(nbytes)[:-1])).astype(np.uint64)
    pairs = np.zeros(len(specs), dtype=_abi.PAIR_DTYPE)
    pos = 0
    for k, (q, t, lo, w) in enumerate(specs):
        cap = ava.cigar_cap_for(int(lens[min(q, len(lens) - 1)]), int(lens[min(t, len(lens) - 1)]))
        pairs[k] = (q, t, lo, w, 0, pos, cap, 0)
        pos += cap
    sc = scoring if scoring is not None else aligner.scoring_preset(preset)
    return aligner.plan_pairs(sc, int(nbytes.sum()) + 16, off, lens, pairs, pos), pairs, lens


def test_planner_kernel_families(monkeypatch):
    from svirlpool_b200 import scheduler
    from svirlpool_b200._abi import PLAN_GENERAL, PLAN_INVALID, PLAN_REBASE, PLAN_S32, PLAN_WINDOWS
    for var in ("SVP_NO_REBASE", "SVP_RING_MIN", "SVP_FORCE_RING", "SVP_FORCE_GENERAL", "SVP_LATENCY_PAIRS", "SVP_KP_MASK"):
        monkeypatch.delenv(var, raising=False)
    monkeypatch.setenv("SVP_SMALL_CALL_PAIRS", "0")          # plan these few pairs as part of a large call
    lens = [8000, 8100, 1000, 1100, 20000, 21000, 130000, 120000]
    specs = [(0, 1, -462, 1024),        # 8 kb reads, band 1024: one KP = 8 warp, windows (16 KB of images for one warp)
             (2, 3, -206, 512),         # 1 kb reads, band 512: one KP = 4 warp, whole images
             (0, 1, -1000, 2048),       # two KP = 8 warps
             (0, 1, -700, 1536),        # three KP = 4 warps beat two KP = 8 warps (KP = 5 / 6 / 7 serve single-warp CTAs only; KP = 12 is off by default)
             (4, 5, -500, 2048),        # 20 kb reads: scores leave the s16 range -> rebased 16-bit lanes
             (0, 1, -7999, 20000),      # band beyond one CTA (16 384 diagonals): a cluster of 2 CTAs
             (6, 7, -768, 1536),        # 130 kb x 120 kb: whole images cannot fit -> windows, rebased
             (0, 1, -462, 1000),        # band not a multiple of 32
             (0, 1, -7999, 131072 + 64),  # wider than a cluster of 8 CTAs covers
             (0, 9, -462, 1024),        # sequence index out of range
             (0, 1, -270, 640),         # bands of 640 / 768 / 896 diagonals: one KP = 5 / 6 / 7 warp instead of a padded KP = 8 one
             (0, 1, -334, 768),
             (0, 1, -398, 896),
             (0, 1, -286, 672)]         # 672 = 28 threads x 24 diagonals: a KP = 6 warp with four idle lanes
    plan, pairs, L = _plan(lens, specs)
    fam = [(int(p["flags"]), int(p["kp"]), int(p["warps"]), int(p["cluster"])) for p in plan]
    assert fam[0] == (PLAN_WINDOWS, 8, 1, 1)
    assert fam[1] == (0, 4, 1, 1)
    assert fam[2] == (PLAN_WINDOWS, 8, 2, 1)
    assert fam[3] == (PLAN_WINDOWS, 4, 3, 1)
    assert fam[4] == (PLAN_REBASE | PLAN_WINDOWS, 8, 2, 1)
    assert fam[5][1:] == (8, 16, 2) and not fam[5][0] & (PLAN_S32 | PLAN_REBASE)
    assert fam[6] == (PLAN_REBASE | PLAN_WINDOWS, 4, 3, 1) and plan[6]["smem_bytes"] < 16384
    assert [f[0] for f in fam[7:10]] == [PLAN_INVALID] * 3
    assert [f[1:] for f in fam[10:]] == [(5, 1, 1), (6, 1, 1), (7, 1, 1), (6, 1, 1)]
    for k in (0, 1, 2, 3, 4, 5, 6, 10, 11, 12, 13):                  # cells = the closed form the scheduler uses; 1 traceback byte per cell (+ virtual cells)
        q, t, lo, w = specs[k]
        cells = scheduler.band_cells(int(L[q]), int(L[t]), lo, w)
        assert int(plan[k]["cells"]) == cells and cells <= int(plan[k]["tb_bytes"])
        if k != 5:                      # (the band of pair 5 is mostly outside the matrix: its rows are mostly virtual cells)
            assert int(plan[k]["tb_bytes"]) <= 1.35 * cells + 64 * w        # incl. the padded row slots of KP = 5 (24 B for 20 cells) and 7
    # the same long pair on 32-bit lanes when the rebase is switched off, and under general scoring
    monkeypatch.setenv("SVP_NO_REBASE", "1")
    p2, _, _ = _plan(lens, [specs[4], specs[6]])
    assert [int(x["flags"]) for x in p2] == [PLAN_S32, PLAN_S32 | PLAN_WINDOWS] and [int(x["kp"]) for x in p2] == [4, 4]
    # a small call (one or two containers) keeps to the KP = 4 / 8 widths: fewer, fuller launches and shorter steps per pair
    monkeypatch.setenv("SVP_SMALL_CALL_PAIRS", "1024")
    p3, _, _ = _plan(lens, specs[10:13])
    assert [(int(x["kp"]), int(x["warps"])) for x in p3] == [(8, 1), (8, 1), (8, 1)]
    monkeypatch.delenv("SVP_NO_REBASE")
    p3, _, _ = _plan(lens, [specs[1], specs[0]], preset="last")          # largest matrix entry 6: 8 kb reads already leave the s16 range
    assert [int(x["flags"]) for x in p3] == [PLAN_GENERAL, PLAN_GENERAL | PLAN_S32] and [int(x["warps"]) for x in p3] == [1, 2]
    # every pair on windows when asked to
    monkeypatch.setenv("SVP_RING_MIN", "0")
    p4, _, _ = _plan(lens, [specs[1]])
    assert int(p4[0]["flags"]) == PLAN_WINDOWS


def test_planner_rejects_unsupported_scoring():
    s = aligner.scoring_preset("map-ont").copy()
    s["matrix"][0][0] = 60                       # 2 * (60 + 6) > 127: one-byte traceback bound
    lens = np.array([100, 100], dtype=np.uint32)
    with pytest.raises(aligner.AlignerError) as e:
        aligner.plan_pairs(s, 64, np.array([0, 32], dtype=np.uint64), lens, np.zeros(0, dtype=_abi.PAIR_DTYPE), 1)
    assert e.value.returncode == _abi.SVP_ERR_UNSUPPORTED
