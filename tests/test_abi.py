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
    return sorted(set(re.findall(r"\b(svp_[a-z_]+)\s*\(", HEADER)))


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
