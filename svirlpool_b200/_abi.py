"""numpy/ctypes views of the structs in include/svp_align.h (one definition, used by the CUDA
binding and — in tests only — by the oracle binding)."""
from __future__ import annotations

import ctypes as C

import numpy as np

ABI_VERSION = 1

SVP_OK, SVP_ERR_ARG, SVP_ERR_CUDA, SVP_ERR_NOMEM, SVP_ERR_UNSUPPORTED, SVP_ERR_NODEVICE = 0, -1, -2, -3, -4, -5
ST_OK, ST_NOHIT, ST_CIGAR_OVF, ST_SCORE_OVF, ST_BAD_PAIR, ST_TOO_LONG = 0, 1, 2, 4, 8, 16
PAIR_REVCOMP = 1

SCORING_DTYPE = np.dtype([
    ("matrix", "<i4", (16,)),
    ("del_open1", "<i4"), ("del_ext1", "<i4"), ("ins_open1", "<i4"), ("ins_ext1", "<i4"),
    ("del_open2", "<i4"), ("del_ext2", "<i4"), ("ins_open2", "<i4"), ("ins_ext2", "<i4"),
    ("min_signal_size", "<i4"), ("zdrop", "<i4"), ("zdrop_ext", "<i4"), ("reserved", "<i4"),
])
PAIR_DTYPE = np.dtype([
    ("q", "<u4"), ("t", "<u4"), ("band_lo", "<i4"), ("band_w", "<u4"),
    ("flags", "<u4"), ("cigar_off", "<u4"), ("cigar_cap", "<u4"), ("reserved", "<u4"),
])
RESULT_DTYPE = np.dtype([
    ("score", "<i4"), ("status", "<u4"), ("q_beg", "<i4"), ("q_end", "<i4"), ("t_beg", "<i4"), ("t_end", "<i4"),
    ("n_match", "<u4"), ("n_mismatch", "<u4"), ("n_ins_ops", "<u4"), ("n_ins_bp", "<u4"),
    ("n_del_ops", "<u4"), ("n_del_bp", "<u4"), ("cigar_off", "<u4"), ("cigar_len", "<u4"),
    ("sv_dist", "<f8"), ("cells", "<u8"),
])
STATS_DTYPE = np.dtype([
    ("fwd_ms", "<f8"), ("tb_ms", "<f8"), ("total_ms", "<f8"), ("cells", "<u8"), ("tb_bytes", "<u8"),
    ("fwd_launches", "<u4"), ("tb_launches", "<u4"), ("waves", "<u4"), ("reserved", "<u4"),
])
assert SCORING_DTYPE.itemsize == 112 and PAIR_DTYPE.itemsize == 32 and RESULT_DTYPE.itemsize == 72
assert STATS_DTYPE.itemsize == 56

PLAN_DTYPE = np.dtype([("flags", "<u4"), ("kp", "<u4"), ("warps", "<u4"), ("cluster", "<u4"), ("smem_bytes", "<u4"), ("reserved", "<u4"),
                       ("cells", "<u8"), ("tb_bytes", "<u8")])
PLAN_S32, PLAN_REBASE, PLAN_WINDOWS, PLAN_GENERAL, PLAN_INVALID, PLAN_TOO_LONG = 1, 2, 4, 8, 16, 32

# every symbol include/svp_align.h declares (checked by tests/test_abi.py against the built library)
EXPORTS = ("svp_abi_version", "svp_device_count", "svp_scoring_preset", "svp_create", "svp_destroy", "svp_pinned_alloc", "svp_pinned_free",
           "svp_last_error", "svp_pack_bytes", "svp_pack_2bit", "svp_ava_pairs", "svp_plan_pairs", "svp_align_batch", "svp_align_batch_summary", "svp_align_batch_device", "svp_submit_batch_device", "svp_sync", "svp_get_stats", "svp_get_stats_total", "svp_timer_mark", "svp_timer_elapsed_ms")


def ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


SEQ_NMASK = 1          # flag in seq_off: the sequence carries a 1-bit-per-base mask of its non-ACGT letters behind its 2-bit codes


def _code_bytes(lens: np.ndarray) -> np.ndarray:
    return ((lens.astype(np.uint64) + 3) // 4 + 15) // 16 * 16


def _layout(lens: np.ndarray, masked: np.ndarray) -> tuple[np.ndarray, int]:
    """Byte offset of every sequence (16-byte aligned; a masked sequence is followed by ceil(len/8) mask bytes, padded to 16)
    and the buffer size (16 readable bytes behind the last sequence)."""
    nbytes = _code_bytes(lens) + np.where(masked, ((lens.astype(np.uint64) + 7) // 8 + 15) // 16 * 16, 0).astype(np.uint64)
    off = np.zeros(len(lens), dtype=np.uint64)
    if len(lens) > 1:
        off[1:] = np.cumsum(nbytes[:-1])
    return off, int(nbytes.sum()) + 16


def _pack_into(packed: np.ndarray, start: int, codes: np.ndarray, mask: np.ndarray | None) -> None:
    pad = (-len(codes)) % 4
    c = np.concatenate((codes, np.zeros(pad, dtype=np.uint8))) if pad else codes
    quad = (c & 3).reshape(-1, 4)
    packed[start:start + len(quad)] = quad[:, 0] | (quad[:, 1] << 2) | (quad[:, 2] << 4) | (quad[:, 3] << 6)
    if mask is not None:
        bits = np.packbits(mask.astype(np.uint8), bitorder="little")
        m0 = start + int(_code_bytes(np.array([len(codes)], dtype=np.uint32))[0])
        packed[m0:m0 + len(bits)] = bits


_LUT = np.full(256, 4, dtype=np.uint8)
for _ch, _code in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("a", 0), ("c", 1), ("g", 2), ("t", 3)):
    _LUT[ord(_ch)] = _code


def pack_2bit(seqs: list[str | bytes]) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """2-bit pack sequences (A=0 C=1 G=2 T=3, base k in bits 2*(k%4) of byte k//4), each sequence
    starting on a 16-byte boundary. Returns (packed uint8, seq_off uint64, seq_len uint32).
    A sequence with letters outside ACGT (N, IUPAC) carries a 1-bit-per-base mask of those positions behind its codes and has
    SEQ_NMASK set in its seq_off entry (include/svp_align.h: SVP_SEQ_NMASK); such a base never matches anything."""
    codes = [_LUT[np.frombuffer(s.encode() if isinstance(s, str) else bytes(s), dtype=np.uint8)] for s in seqs]
    return pack_codes(codes)


def pack_codes(seqs: list[np.ndarray]) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """pack_2bit() for sequences already given as base-code arrays (uint8, values 0..3; 4 = a letter outside ACGT)."""
    n = len(seqs)
    lens = np.fromiter((len(s) for s in seqs), dtype=np.uint32, count=n)
    masked = np.fromiter((bool(len(s)) and bool((s > 3).any()) for s in seqs), dtype=bool, count=n)
    off, total = _layout(lens, masked)
    packed = np.zeros(total, dtype=np.uint8)
    for k, codes in enumerate(seqs):
        _pack_into(packed, int(off[k]), np.asarray(codes, dtype=np.uint8), (codes > 3) if masked[k] else None)
    return packed, off | masked.astype(np.uint64) * np.uint64(SEQ_NMASK), lens


def cigartuples(cigar_buf: np.ndarray, result) -> list[tuple[int, int]]:
    """(op, len) list of one svp_result's CIGAR."""
    o, n = int(result["cigar_off"]), int(result["cigar_len"])
    w = cigar_buf[o:o + n]
    return list(zip((w & 15).tolist(), (w >> 4).tolist()))
