"""ctypes binding of libsvp_align.so (include/svp_align.h) — the only compute engine of this package.

There is no CPU fallback: a missing library, a missing GPU or an unsupported scoring profile raises
(AlignerError / SVP_ERR_* codes), it never silently degrades. The error mapping mirrors what the
reference's seam raises (util.align_reads_with_minimap, util/util.py:162-238): alignment failures
surface as a subprocess.CalledProcessError subclass, which the reference's callers already catch
(consensus.py:1387-1396).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from . import _abi
from ._abi import PAIR_DTYPE, RESULT_DTYPE, SCORING_DTYPE, STATS_DTYPE, SVP_OK, ptr

LIB_PATH = Path(os.environ.get("SVP_LIB", Path(__file__).resolve().parent / "libsvp_align.so"))       # (SVP_LIB: A/B runs of two builds)


class AlignerError(subprocess.CalledProcessError):
    """Raised for every non-zero return code of the C-ABI."""

    def __init__(self, code: int, message: str):
        super().__init__(code, "libsvp_align", output=message)
        self.message = message

    def __str__(self) -> str:
        return f"libsvp_align error {self.returncode}: {self.message}"


def load_library(path: Path | None = None) -> C.CDLL:
    path = Path(path) if path else LIB_PATH
    if not path.exists():
        raise FileNotFoundError(
            f"{path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C svirlpool_b200/csrc). svirlpool_b200 has no CPU fallback.")
    lib = C.CDLL(str(path))
    lib.svp_abi_version.restype = C.c_int
    lib.svp_device_count.restype = C.c_int
    lib.svp_scoring_preset.restype = C.c_int
    lib.svp_create.restype = C.c_int
    lib.svp_create.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.svp_destroy.restype = None
    lib.svp_destroy.argtypes = [C.c_void_p]
    lib.svp_last_error.restype = C.c_char_p
    lib.svp_last_error.argtypes = [C.c_void_p]
    lib.svp_pinned_alloc.restype = C.c_void_p
    lib.svp_pinned_alloc.argtypes = [C.c_size_t]
    lib.svp_pinned_free.restype = None
    lib.svp_pinned_free.argtypes = [C.c_void_p]
    lib.svp_get_stats.restype = C.c_int
    lib.svp_get_stats.argtypes = [C.c_void_p, C.c_void_p]
    lib.svp_timer_mark.restype = C.c_int
    lib.svp_timer_mark.argtypes = [C.c_void_p, C.c_int]
    lib.svp_timer_elapsed_ms.restype = C.c_double
    lib.svp_timer_elapsed_ms.argtypes = [C.c_void_p]
    batch_args = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                  C.c_void_p, C.c_void_p, C.c_size_t]
    lib.svp_align_batch.restype = C.c_int
    lib.svp_align_batch.argtypes = batch_args
    lib.svp_align_batch_device.restype = C.c_int
    lib.svp_align_batch_device.argtypes = batch_args
    lib.svp_submit_batch_device.restype = C.c_int
    lib.svp_submit_batch_device.argtypes = batch_args
    lib.svp_sync.restype = C.c_int
    lib.svp_sync.argtypes = [C.c_void_p]
    lib.svp_get_stats_total.restype = C.c_int
    lib.svp_get_stats_total.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.svp_align_batch_summary.restype = C.c_int
    lib.svp_align_batch_summary.argtypes = batch_args[:9] + [C.c_size_t]
    lib.svp_pack_bytes.restype = C.c_size_t
    lib.svp_pack_bytes.argtypes = [C.c_void_p, C.c_uint32]
    lib.svp_pack_2bit.restype = C.c_int
    lib.svp_pack_2bit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.svp_ava_pairs.restype = C.c_int
    lib.svp_ava_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, C.c_int,
                                  C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.svp_plan_pairs.restype = C.c_int
    lib.svp_plan_pairs.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_size_t, C.c_void_p]
    if lib.svp_abi_version() != _abi.ABI_VERSION:
        raise RuntimeError(f"ABI mismatch: library {lib.svp_abi_version()}, python {_abi.ABI_VERSION}")
    return lib


_LIB: C.CDLL | None = None


def shared_library() -> C.CDLL:
    """The process-wide handle of libsvp_align.so (host-side builders need no Aligner)."""
    global _LIB
    if _LIB is None:
        _LIB = load_library()
    return _LIB


def pack_2bit_native(seqs: list[str | bytes]) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """svp_pack_2bit: same layout and return value as _abi.pack_2bit (packed uint8, seq_off uint64, seq_len uint32)."""
    lib = shared_library()
    raw = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    n = len(raw)
    lens = np.fromiter((len(b) for b in raw), dtype=np.uint32, count=n)
    ptrs = (C.c_char_p * max(n, 1))(*raw)
    total = lib.svp_pack_bytes(ptr(lens), n)
    packed = np.empty(total, dtype=np.uint8)
    off = np.zeros(n, dtype=np.uint64)
    rc = lib.svp_pack_2bit(ptrs, ptr(lens), n, ptr(packed), total, ptr(off))
    if rc != SVP_OK:
        raise AlignerError(rc, "svp_pack_2bit failed")
    # svp_pack_bytes is the worst case (a mask behind every sequence); trim to what was used + the 16 readable bytes
    used = 16
    if n:
        last = int(lens[-1])
        used += (int(off[-1]) & ~15) + ((last + 3) // 4 + 15) // 16 * 16 + (((last + 7) // 8 + 15) // 16 * 16 if int(off[-1]) & _abi.SEQ_NMASK else 0)
    return packed[:used], off, lens


def ava_pairs_native(seq_len: np.ndarray, name_order: np.ndarray | None, strand: np.ndarray | None, offsets: np.ndarray | None,
                     seq_base: int, n_reads: int, slack: int, granule: int, full: bool, cigar_start: int = 0) -> tuple[np.ndarray, int]:
    """svp_ava_pairs: the pair descriptors of one container's all-vs-all alignment and the next free CIGAR word."""
    lib = shared_library()
    # both orientations of a pair are emitted when either strand is unknown (NULL array, or a 0 entry)
    known = strand is not None and bool(np.all(np.asarray(strand) != 0))
    cap = n_reads * (n_reads - 1) // 2 if known else n_reads * (n_reads - 1)
    pairs = np.zeros(max(cap, 1), dtype=PAIR_DTYPE)
    n_pairs = np.zeros(1, dtype=np.uint32)
    words = np.array([cigar_start], dtype=np.uint64)
    order = None if name_order is None else np.ascontiguousarray(name_order, dtype=np.uint32)
    st = None if strand is None else np.ascontiguousarray(strand, dtype=np.int8)
    offs = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int32)
    rc = lib.svp_ava_pairs(ptr(np.ascontiguousarray(seq_len, dtype=np.uint32)), None if order is None else ptr(order),
                           None if st is None else ptr(st), None if offs is None else ptr(offs), int(seq_base), int(n_reads),
                           int(slack), int(granule), int(bool(full)), ptr(pairs), len(pairs), ptr(n_pairs), ptr(words))
    if rc != SVP_OK:
        raise AlignerError(rc, "svp_ava_pairs failed (invalid argument, or the CIGAR slots of the batch exceed 2^32 words)")
    return pairs[:int(n_pairs[0])], int(words[0])


def plan_pairs(scoring: np.ndarray, packed_bytes: int, seq_off: np.ndarray, seq_len: np.ndarray, pairs: np.ndarray,
               cigar_words: int) -> np.ndarray:
    """svp_plan_pairs: per pair the kernel family / geometry / memory the library would use (host only, no GPU)."""
    lib = shared_library()
    out = np.zeros(max(len(pairs), 1), dtype=_abi.PLAN_DTYPE)
    rc = lib.svp_plan_pairs(ptr(scoring), int(packed_bytes), ptr(seq_off), ptr(seq_len), len(seq_len), ptr(pairs), len(pairs),
                            int(cigar_words), ptr(out))
    if rc != SVP_OK:
        raise AlignerError(rc, (lib.svp_last_error(None) or b"").decode())
    return out[:len(pairs)]


def scoring_preset(name: str, lib: C.CDLL | None = None) -> np.ndarray:
    lib = lib or load_library()
    s = np.zeros(1, dtype=SCORING_DTYPE)
    if lib.svp_scoring_preset(name.encode(), ptr(s)) != SVP_OK:
        raise ValueError(f"unknown scoring preset {name!r}")
    return s


class PinnedBuffer:
    """A page-locked host buffer (svp_pinned_alloc) viewed as a numpy array; grows on demand."""

    def __init__(self, lib: C.CDLL, dtype):
        self.lib, self.dtype, self.ptr, self.cap = lib, np.dtype(dtype), None, 0

    def view(self, n: int) -> np.ndarray:
        n = max(int(n), 1)
        if n > self.cap:
            self.release()
            cap = n + n // 4
            p = self.lib.svp_pinned_alloc(cap * self.dtype.itemsize)
            if not p:
                raise MemoryError("svp_pinned_alloc failed")
            self.ptr, self.cap = p, cap
        buf = (C.c_uint8 * (self.cap * self.dtype.itemsize)).from_address(self.ptr)
        return np.frombuffer(buf, dtype=self.dtype, count=n)

    def release(self) -> None:
        if self.ptr:
            self.lib.svp_pinned_free(self.ptr)
            self.ptr, self.cap = None, 0


_SHARED: dict = {}


def shared_aligner(preset: str = "map-ont", device: int | None = None) -> "Aligner":
    """The process-wide Aligner for (device, preset), created on first use with the default bounded arena and closed at exit —
    what the file seam (util.align_reads_with_minimap) uses when the caller passes no engine, so that repeated calls do not
    pay a handle (and an arena allocation) each. device None: the current CUDA device."""
    import atexit
    key = (device, preset)
    eng = _SHARED.get(key)
    if eng is None or eng._h is None:
        # one container per call: a few GB of traceback rows at most (435 pairs x 8.5 MB); 16 GB leave the device to the other jobs
        gb = int(os.environ.get("SVP_SHARED_ARENA_GB", "16"))
        try:
            eng = Aligner(device=-1 if device is None else device, preset=preset, tb_bytes=gb << 30)
        except AlignerError:                       # less than that is free: the library's default shrinks until it fits
            eng = Aligner(device=-1 if device is None else device, preset=preset)
        if not _SHARED:
            atexit.register(close_shared_aligners)
        _SHARED[key] = eng
    return eng


def close_shared_aligners() -> None:
    for eng in list(_SHARED.values()):
        eng.close()
    _SHARED.clear()


def estimate_seconds(scoring: np.ndarray, batch, gcups: float = 800.0, overhead: float = 0.02) -> float:
    """Device time a batch will take, from the planner alone (svp_plan_pairs: spec cells per pair; no GPU needed), at a conservative
    rate — the budget check behind the `timeout` of the file seam."""
    plan = plan_pairs(scoring, batch.packed.nbytes, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
    return overhead + float(plan["cells"].sum()) / (gcups * 1e9)


class Aligner:
    """One handle = one GPU + one scoring profile. Not thread-safe; one call at a time."""

    def __init__(self, device: int = 0, preset: str = "map-ont", scoring: np.ndarray | None = None,
                 tb_bytes: int = 0, lib_path: Path | None = None):
        self._h = None
        self.lib = load_library(lib_path)
        self.scoring = scoring if scoring is not None else scoring_preset(preset, self.lib)
        self.device = device
        handle = C.c_void_p()
        rc = self.lib.svp_create(int(device), ptr(self.scoring), int(tb_bytes), C.byref(handle))
        if rc != SVP_OK:
            raise AlignerError(rc, (self.lib.svp_last_error(None) or b"").decode())
        self._h = handle
        self._res_buf = PinnedBuffer(self.lib, RESULT_DTYPE)
        self._cig_buf = PinnedBuffer(self.lib, np.uint32)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._res_buf.release()
            self._cig_buf.release()
            self.lib.svp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != SVP_OK:
            raise AlignerError(rc, (self.lib.svp_last_error(self._h) or b"").decode())

    def align_batch(self, packed: np.ndarray, seq_off: np.ndarray, seq_len: np.ndarray, pairs: np.ndarray,
                    cigar_words: int) -> tuple[np.ndarray, np.ndarray]:
        """Host buffers in, host buffers out (svp_align_batch): H2D, kernels, D2H inside the call.
        Returns (results, cigar words); the CIGARs are dense in pair order (results["cigar_off"]).
        With copy=False the arrays are views of the handle's pinned buffers, valid until the next call."""
        return self._align_batch(packed, seq_off, seq_len, pairs, cigar_words, copy=True)

    def align_batch_nocopy(self, packed, seq_off, seq_len, pairs, cigar_words):
        return self._align_batch(packed, seq_off, seq_len, pairs, cigar_words, copy=False)

    def _align_batch(self, packed, seq_off, seq_len, pairs, cigar_words, copy: bool):
        assert pairs.dtype == PAIR_DTYPE and packed.dtype == np.uint8
        res = self._res_buf.view(len(pairs))[:len(pairs)]
        cig = self._cig_buf.view(cigar_words)
        self._check(self.lib.svp_align_batch(self._h, ptr(packed), packed.nbytes, ptr(seq_off), ptr(seq_len), len(seq_len),
                                             ptr(pairs), len(pairs), ptr(res), ptr(cig), len(cig)))
        used = int(res["cigar_off"][-1]) + int(res["cigar_len"][-1]) if len(res) else 0
        if copy:
            return res.copy(), cig[:max(used, 1)].copy()
        return res, cig[:max(used, 1)]

    def align_batch_summary(self, packed: np.ndarray, seq_off: np.ndarray, seq_len: np.ndarray, pairs: np.ndarray,
                            cigar_words: int, copy: bool = True) -> np.ndarray:
        """Result rows only (svp_align_batch_summary): no CIGAR leaves the device. Enough for the similarity matrix with unit
        density weights and the alignment-length matrix (consensus_lib.similarity_from_summary)."""
        assert pairs.dtype == PAIR_DTYPE and packed.dtype == np.uint8
        res = self._res_buf.view(len(pairs))[:len(pairs)]
        self._check(self.lib.svp_align_batch_summary(self._h, ptr(packed), packed.nbytes, ptr(seq_off), ptr(seq_len), len(seq_len),
                                                     ptr(pairs), len(pairs), ptr(res), int(cigar_words)))
        return res.copy() if copy else res

    def align_batch_device(self, d_packed_ptr: int, packed_bytes: int, seq_off: np.ndarray, seq_len: np.ndarray,
                           pairs: np.ndarray, d_results_ptr: int, d_cigar_ptr: int, cigar_words: int) -> None:
        """Sequences/results/CIGARs resident on the device (raw pointers, e.g. torch data_ptr())."""
        self._check(self.lib.svp_align_batch_device(self._h, C.c_void_p(d_packed_ptr), packed_bytes, ptr(seq_off), ptr(seq_len),
                                                    len(seq_len), ptr(pairs), len(pairs), C.c_void_p(d_results_ptr),
                                                    C.c_void_p(d_cigar_ptr), int(cigar_words)))

    def submit_batch_device(self, d_packed_ptr: int, packed_bytes: int, seq_off: np.ndarray, seq_len: np.ndarray,
                            pairs: np.ndarray, d_results_ptr: int, d_cigar_ptr: int, cigar_words: int) -> None:
        """Pipelined align_batch_device (svp_submit_batch_device): returns after the enqueue; results are complete after sync().
        The device buffers must stay untouched until then; consecutive submits overlap on the device."""
        self._check(self.lib.svp_submit_batch_device(self._h, C.c_void_p(d_packed_ptr), packed_bytes, ptr(seq_off), ptr(seq_len),
                                                     len(seq_len), ptr(pairs), len(pairs), C.c_void_p(d_results_ptr),
                                                     C.c_void_p(d_cigar_ptr), int(cigar_words)))

    def sync(self) -> None:
        self._check(self.lib.svp_sync(self._h))

    def stats_total(self, reset: bool = False) -> dict:
        """Sums over all completed batches since the handle was created or last reset."""
        s = np.zeros(1, dtype=STATS_DTYPE)
        self._check(self.lib.svp_get_stats_total(self._h, ptr(s), int(reset)))
        return {k: s[0][k].item() for k in STATS_DTYPE.names}

    def timer_mark(self, which: int) -> None:
        """Record device stopwatch mark 0 / 1 on the handle's main stream (svp_timer_mark)."""
        self._check(self.lib.svp_timer_mark(self._h, int(which)))

    def timer_elapsed_ms(self) -> float:
        ms = self.lib.svp_timer_elapsed_ms(self._h)
        if ms < 0:
            raise AlignerError(int(ms), "svp_timer_elapsed_ms failed")
        return ms

    def stats(self) -> dict:
        s = np.zeros(1, dtype=STATS_DTYPE)
        self._check(self.lib.svp_get_stats(self._h, ptr(s)))
        return {k: s[0][k].item() for k in STATS_DTYPE.names}
