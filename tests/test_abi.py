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
    packed, off, lens = _abi.pack_2bit(["ACGTAC", "TTTT", "n" * 5])
    assert off.tolist() == [0, 16, 32] and lens.tolist() == [6, 4, 5]
    assert packed[0] == 0b11100100 and packed[1] == 0b0100 and packed[16] == 0xFF and packed[32] == 0
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
