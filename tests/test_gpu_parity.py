"""GPU parity tests proper: libsvp_align.so (through the C-ABI binding) against the CPU oracle on the
same seeded inputs — bit-exact scores, coordinates, counts and CIGARs; sv_dist within 1e-6 relative.
Run on the B200 box: python -m pytest tests -m gpu."""
import numpy as np
import pytest

from svirlpool_b200 import ava
from svirlpool_b200._abi import PAIR_REVCOMP, cigartuples

pytestmark = pytest.mark.gpu

INT_FIELDS = ("score", "status", "q_beg", "q_end", "t_beg", "t_end", "n_match", "n_mismatch", "n_ins_ops", "n_ins_bp",
              "n_del_ops", "n_del_bp", "cigar_len", "cells")


@pytest.fixture(scope="module")
def gpu():
    from svirlpool_b200.aligner import Aligner
    a = Aligner(device=0, preset="map-ont", tb_bytes=4 << 30)
    yield a
    a.close()


def mutate(rng, seq, sub=0.035, ins=0.03, dele=0.035):
    out = []
    for ch in seq:
        r = rng.random()
        if r < dele:
            continue
        if r < dele + sub:
            out.append("ACGT"[(("ACGT".index(ch)) + 1 + rng.integers(3)) % 4])
        else:
            out.append(ch)
        while rng.random() < ins:
            out.append("ACGT"[rng.integers(4)])
    return "".join(out)


def randseq(rng, n):
    return "".join("ACGT"[k] for k in rng.integers(0, 4, n))


def revcomp(s):
    return s.translate(str.maketrans("ACGT", "TGCA"))[::-1]


def assert_same(batch, gres, gcig, ores, ocig):
    for f in INT_FIELDS:
        bad = np.nonzero(gres[f] != ores[f])[0]
        assert len(bad) == 0, (f"{f}: {len(bad)} mismatches, first at pair {bad[0]}: gpu={gres[f][bad[0]]} "
                               f"oracle={ores[f][bad[0]]} pair={batch.pairs[bad[0]]} gpu_row={gres[bad[0]]} oracle_row={ores[bad[0]]}")
    for k in range(len(gres)):
        assert cigartuples(gcig, gres[k]) == cigartuples(ocig, ores[k]), f"CIGAR differs at pair {k}"
    np.testing.assert_allclose(gres["sv_dist"], ores["sv_dist"], rtol=1e-6, atol=0)


def run_both(gpu, oracle, names, seqs, specs):
    batch = ava.PairBatch.build(names, seqs, specs)
    gres, gcig = gpu.align_batch(batch.packed, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
    ores, ocig = oracle.align_batch(batch.packed, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
    assert_same(batch, gres, gcig, ores, ocig)
    return gres, gcig


def test_identical_and_simple_edits(gpu, oracle):
    rng = np.random.default_rng(1)
    ref = randseq(rng, 1000)
    seqs = [ref, ref[100:800], ref[:480] + ref[500:], ref[:500] + "A" * 20 + ref[500:], revcomp(ref[100:800])]
    names = [f"s{k}" for k in range(len(seqs))]
    specs = []
    for q in range(1, len(seqs)):
        lo, w = ava.full_band(len(seqs[q]), len(seqs[0]))
        specs.append((q, 0, lo, w, 0))
        specs.append((q, 0, lo, w, PAIR_REVCOMP))
    gres, gcig = run_both(gpu, oracle, names, seqs, specs)
    assert gres[0]["score"] == 1400


@pytest.mark.parametrize("seed,length,band", [(2, 300, 64), (3, 1200, 256), (4, 2500, 512), (5, 3000, 1024), (6, 5000, 544)])
def test_noisy_pairs_banded(gpu, oracle, seed, length, band):
    rng = np.random.default_rng(seed)
    seqs, specs = [], []
    for p in range(12):
        tmpl = randseq(rng, length + int(rng.integers(-50, 50)))
        a, b = mutate(rng, tmpl), mutate(rng, tmpl)
        if p % 3 == 0:      # one haplotype carries an insertion / deletion
            pos = len(b) // 2
            sv = int(rng.integers(20, max(21, band // 3)))
            b = b[:pos] + (randseq(rng, sv) if p % 2 else "") + b[pos + (0 if p % 2 else sv):]
        rc = bool(rng.integers(2))
        seqs += [revcomp(a) if rc else a, b]
        lo, w = ava.offset_band(len(a), len(b), slack=max(8, band // 2 - 40), granule=32)
        specs.append((2 * p, 2 * p + 1, lo, w, PAIR_REVCOMP if rc else 0))
    names = [f"r{k}" for k in range(len(seqs))]
    run_both(gpu, oracle, names, seqs, specs)


def test_golden_fixtures_on_gpu(gpu, golden):
    """The 14 re-alignable frozen minimap2 records, through the CUDA path."""
    for case in golden["cases"]:
        recs = ava.map_queries(gpu, case["reads"], "ref", case["reference"])
        got = sorted((r.query_name, r.reference_start, r.is_reverse, r.cigarstring.replace("H", "S"), r.tags["AS"], r.tags["NM"]) for r in recs)
        want = sorted((r["read"], r["ref_start"], bool(r["flag"] & 16), r["cigar"].replace("H", "S"), r["AS"], r["NM"]) for r in case["records"])
        assert got == want, case["name"]


def test_edge_cases(gpu, oracle):
    rng = np.random.default_rng(9)
    a = randseq(rng, 700)
    unrelated = randseq(rng, 650)
    tiny = a[10:40]
    seqs = [a, unrelated, tiny, "ACGT" * 50, "A" * 300]
    names = ["a", "u", "tiny", "rep", "homo"]
    specs = []
    for q, t in [(1, 0), (2, 0), (0, 2), (3, 3), (4, 4), (3, 4), (0, 0)]:
        lo, w = ava.full_band(len(seqs[q]), len(seqs[t]))
        specs.append((q, t, lo, w, 0))
    # narrow off-centre bands, bands partly outside the matrix
    specs += [(0, 0, -16, 32, 0), (0, 0, 100, 64, 0), (0, 0, -731, 64, 0), (0, 0, 668, 64, 0), (2, 0, 0, 32, 0)]
    gres, _ = run_both(gpu, oracle, names, seqs, specs)
    assert gres[6]["score"] == 1400 and gres[6]["cigar_len"] == 1


def test_invalid_pairs_are_flagged(gpu):
    rng = np.random.default_rng(11)
    seqs = [randseq(rng, 100), randseq(rng, 100)]
    batch = ava.PairBatch.build(["a", "b"], seqs, [(0, 1, -99, 224, 0), (0, 1, -99, 224, 0), (0, 1, 5000, 32, 0)])
    batch.pairs[1]["band_w"] = 40          # not a multiple of 32
    res, _ = gpu.align_batch(batch.packed, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
    assert res[0]["status"] in (0, 1)
    assert res[1]["status"] == 8 and res[2]["status"] == 8


def test_small_arena(oracle):
    """A 16 MiB arena holds a handful of pairs at a time; results must not depend on it."""
    from svirlpool_b200.aligner import Aligner
    rng = np.random.default_rng(21)
    seqs, specs = [], []
    for p in range(40):
        tmpl = randseq(rng, 3000)
        seqs += [mutate(rng, tmpl), mutate(rng, tmpl)]
        lo, w = ava.offset_band(len(seqs[-2]), len(seqs[-1]), slack=200, granule=512)
        specs.append((2 * p, 2 * p + 1, lo, w, 0))
    with Aligner(device=0, tb_bytes=16 << 20) as small:
        run_both(small, oracle, [f"r{k}" for k in range(len(seqs))], seqs, specs)
        assert small.stats()["fwd_launches"] >= 1


# ---------------------------------------------------------------------------------------------
# general scoring (full 4x4 matrix, separate insertion / deletion costs) and 32-bit lanes
# ---------------------------------------------------------------------------------------------

def noisy_batch(rng, n_pairs, length, band, granule=32, sv_every=3):
    seqs, specs = [], []
    for p in range(n_pairs):
        tmpl = randseq(rng, length + int(rng.integers(-50, 50)))
        a, b = mutate(rng, tmpl), mutate(rng, tmpl)
        if p % sv_every == 0:
            pos = len(b) // 2
            sv = int(rng.integers(20, max(21, band // 3)))
            b = b[:pos] + (randseq(rng, sv) if p % 2 else "") + b[pos + (0 if p % 2 else sv):]
        rc = bool(rng.integers(2))
        seqs += [revcomp(a) if rc else a, b]
        lo, w = ava.offset_band(len(a), len(b), slack=max(8, band // 2 - 40), granule=granule)
        specs.append((2 * p, 2 * p + 1, lo, w, PAIR_REVCOMP if rc else 0))
    return [f"r{k}" for k in range(len(seqs))], seqs, specs


def custom_scoring(matrix, d1, i1, d2=(-1, 0), i2=(-1, 0), zdrop=400, zdrop_ext=2):
    from svirlpool_b200._abi import SCORING_DTYPE
    s = np.zeros(1, dtype=SCORING_DTYPE)
    s["matrix"][0] = np.asarray(matrix, dtype=np.int32).reshape(16)
    s["del_open1"], s["del_ext1"] = d1
    s["ins_open1"], s["ins_ext1"] = i1
    s["del_open2"], s["del_ext2"] = d2
    s["ins_open2"], s["ins_ext2"] = i2
    s["min_signal_size"], s["zdrop"], s["zdrop_ext"] = 12, zdrop, zdrop_ext
    return s


LAST_LIKE = [[6, -18, -9, -19], [-18, 5, -20, -8], [-9, -20, 5, -18], [-19, -8, -18, 6]]
SCORINGS = {
    "last_preset": None,
    "last_like_asym": custom_scoring(LAST_LIKE, (21, 9), (25, 7)),
    "matrix_two_piece_asym": custom_scoring(LAST_LIKE, (8, 4), (10, 3), d2=(30, 2), i2=(36, 1), zdrop_ext=1),
    "two_piece_del_only": custom_scoring([[3, -5, -4, -6], [-5, 3, -6, -4], [-4, -6, 3, -5], [-6, -4, -5, 3]], (5, 3), (6, 2), d2=(26, 1)),
}


@pytest.mark.parametrize("name", list(SCORINGS))
@pytest.mark.parametrize("seed,length,band", [(31, 400, 64), (32, 2000, 512), (33, 3500, 1056)])
def test_general_scoring(name, seed, length, band):
    """4x4 matrices and asymmetric gap costs: svp_fwdg_kernel vs the oracle."""
    from oracle_engine import OracleEngine
    from svirlpool_b200.aligner import Aligner
    scoring = SCORINGS[name]
    kw = {"preset": "last"} if scoring is None else {"scoring": scoring}
    rng = np.random.default_rng(seed)
    names, seqs, specs = noisy_batch(rng, 10, length, band)
    with Aligner(device=0, tb_bytes=2 << 30, **kw) as g:
        run_both(g, OracleEngine(**kw), names, seqs, specs)


def test_general_scoring_edge_cases():
    from oracle_engine import OracleEngine
    from svirlpool_b200.aligner import Aligner
    rng = np.random.default_rng(41)
    a = randseq(rng, 700)
    seqs = [a, randseq(rng, 650), a[10:40], "ACGT" * 50, "A" * 300, "C" * 200]
    names = ["a", "u", "tiny", "rep", "homoA", "homoC"]
    specs = []
    for q, t in [(1, 0), (2, 0), (0, 2), (3, 3), (4, 4), (3, 4), (0, 0), (5, 5), (4, 5)]:
        lo, w = ava.full_band(len(seqs[q]), len(seqs[t]))
        specs.append((q, t, lo, w, 0))
        specs.append((q, t, lo, w, PAIR_REVCOMP))
    specs += [(0, 0, -16, 32, 0), (0, 0, 100, 64, 0), (0, 0, -731, 64, 0), (0, 0, 668, 64, 0), (2, 0, 0, 32, 0)]
    for scoring in (SCORINGS["last_like_asym"], SCORINGS["matrix_two_piece_asym"]):
        with Aligner(device=0, tb_bytes=1 << 30, scoring=scoring) as g:
            run_both(g, OracleEngine(scoring=scoring), names, seqs, specs)


def test_forced_general_equals_fast_path(gpu, oracle, monkeypatch):
    """map-ont scoring pushed through the general kernels (SVP_FORCE_GENERAL) gives the fast path's results."""
    from svirlpool_b200.aligner import Aligner
    rng = np.random.default_rng(43)
    names, seqs, specs = noisy_batch(rng, 12, 2500, 544)
    batch = ava.PairBatch.build(names, seqs, specs)
    fres, fcig = gpu.align_batch(batch.packed, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
    monkeypatch.setenv("SVP_FORCE_GENERAL", "1")
    with Aligner(device=0, preset="map-ont", tb_bytes=1 << 30) as g:
        gres, gcig = g.align_batch(batch.packed, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
    assert_same(batch, gres, gcig, fres, fcig)
    ores, ocig = oracle.align_batch(batch.packed, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
    assert_same(batch, gres, gcig, ores, ocig)


@pytest.mark.parametrize("no_rebase", ["0", "1"])
@pytest.mark.parametrize("seed,length,band", [(51, 17000, 512), (52, 21000, 1056), (53, 30000, 2048), (56, 60000, 544)])
def test_long_reads_beyond_s16(oracle, monkeypatch, no_rebase, seed, length, band):
    """Reads beyond the s16 score range (match * min(m, n) > 32000): the rebased s16x2 fast path (per-warp score offsets,
    default) and the 32-bit lane kernel (SVP_NO_REBASE=1) both reproduce the oracle; scores up to ~70 000."""
    from svirlpool_b200.aligner import Aligner
    monkeypatch.setenv("SVP_NO_REBASE", no_rebase)
    rng = np.random.default_rng(seed)
    names, seqs, specs = noisy_batch(rng, 4 if length < 40000 else 2, length, band)
    with Aligner(device=0, preset="map-ont", tb_bytes=4 << 30) as g:
        gres, _ = run_both(g, oracle, names, seqs, specs)
    assert gres["score"].max() > 12000 and not (gres["status"] & 4).any()


@pytest.mark.parametrize("ring_min", ["0", "1000000000"])
def test_sliding_windows_and_whole_images_fast_path(oracle, monkeypatch, ring_min):
    """The s16x2 fast path with sliding sequence windows for every pair (SVP_RING_MIN=0) and with whole shared-memory
    images for every pair that fits (huge SVP_RING_MIN): single-warp, multi-warp, rebased (long reads) and cluster CTAs,
    forward and reverse-complemented queries, all against the oracle."""
    from svirlpool_b200.aligner import Aligner
    monkeypatch.setenv("SVP_RING_MIN", ring_min)
    rng = np.random.default_rng(58)
    n1, s1, p1 = noisy_batch(rng, 6, 1500, 256)
    n2, s2, p2 = noisy_batch(rng, 4, 6000, 1056)
    n3, s3, p3 = noisy_batch(rng, 3, 9000, 2080, sv_every=1)
    n4, s4, p4 = noisy_batch(rng, 2, 19000, 544)
    seqs, specs = [], []
    for sq, sp in ((s1, p1), (s2, p2), (s3, p3), (s4, p4)):
        off = len(seqs)
        seqs += sq
        specs += [(q + off, t + off, lo, w, fl) for q, t, lo, w, fl in sp]
    t1 = randseq(rng, 20000)
    off = len(seqs)
    seqs += [t1, mutate(rng, t1[12000:17500]), revcomp(mutate(rng, t1[300:6200]))]
    for q, rc in ((off + 1, 0), (off + 2, 1)):
        lo, w = ava.full_band(len(seqs[q]), len(t1))                   # 26k diagonals: a cluster of 2 CTAs
        specs.append((q, off, lo, w, PAIR_REVCOMP if rc else 0))
    with Aligner(device=0, preset="map-ont", tb_bytes=4 << 30) as g:
        gres, _ = run_both(g, oracle, [f"r{k}" for k in range(len(seqs))], seqs, specs)
    assert not gres["status"].any()


def test_rebased_kernel_low_error_reads_and_restarts(gpu, oracle):
    """Rebased kernel: accurate 40 kb reads (scores ~75 000, offsets far beyond 32768), a pair whose alignment only starts
    after 25 kb of unrelated sequence (cells at H = 0 next to nothing, then a fresh start while other pairs' warps sit at
    high offsets), and a pair with a 1.2 kb deletion that moves the path across warps."""
    rng = np.random.default_rng(57)
    t = randseq(rng, 40000)
    a = mutate(rng, t, sub=0.01, ins=0.005, dele=0.005)
    b = randseq(rng, 25000) + mutate(rng, t[25000:], sub=0.02, ins=0.01, dele=0.01)
    c = mutate(rng, t[:20000] + t[21200:], sub=0.02, ins=0.01, dele=0.01)
    seqs, names = [t, a, b, c, revcomp(a)], ["t", "a", "b", "c", "a_rc"]
    specs = []
    for q, rc in ((1, 0), (2, 0), (3, 0), (4, 1)):
        lo, w = ava.offset_band(len(seqs[q]), len(t), slack=700, granule=32)
        specs.append((q, 0, lo, w, PAIR_REVCOMP if rc else 0))
    gres, _ = run_both(gpu, oracle, names, seqs, specs)
    assert gres["score"][0] > 65536 and gres["score"][3] == gres["score"][0] and gres["n_del_bp"][2] >= 1200 and not gres["status"].any()


def test_long_reads_mixed_with_short(gpu, oracle):
    """One batch holding s16 and s32 pairs (different kernel families in one wave)."""
    rng = np.random.default_rng(54)
    n1, s1, p1 = noisy_batch(rng, 6, 3000, 512)
    n2, s2, p2 = noisy_batch(rng, 2, 18000, 512)
    off = len(s1)
    specs = p1 + [(q + off, t + off, lo, w, fl) for q, t, lo, w, fl in p2]
    run_both(gpu, oracle, [f"r{k}" for k in range(off + len(s2))], s1 + s2, specs)


def test_long_reads_general_scoring():
    from oracle_engine import OracleEngine
    from svirlpool_b200.aligner import Aligner
    scoring = SCORINGS["last_like_asym"]
    rng = np.random.default_rng(55)
    names, seqs, specs = noisy_batch(rng, 3, 9000, 512)     # 6 * 9000 > 32000: 32-bit lanes + general scoring
    with Aligner(device=0, tb_bytes=2 << 30, scoring=scoring) as g:
        gres, _ = run_both(g, OracleEngine(scoring=scoring), names, seqs, specs)
    assert gres["score"].max() > 4000


def test_alignment_on_high_band_columns(gpu, oracle):
    """Alignments that run through band columns >= 8192 (KP = 8) / >= 4096 (KP = 4 geometry): the traceback's
    diagonal -> byte-column mapping must not wrap (regression: a 32-bit multiply-shift did)."""
    from oracle_engine import OracleEngine
    from svirlpool_b200.aligner import Aligner
    rng = np.random.default_rng(61)
    t1 = randseq(rng, 12000)
    q1 = mutate(rng, t1[8000:11000])                  # diagonal ~ +8000, band column ~ 11000 of 15008
    t2 = randseq(rng, 6000)
    q2 = mutate(rng, t2[4900:5900])                   # diagonal ~ +4900, band column ~ 5900 of 7008
    seqs, names = [t1, q1, t2, q2], ["t1", "q1", "t2", "q2"]
    specs = []
    for q, t in ((1, 0), (3, 2)):
        lo, w = ava.full_band(len(seqs[q]), len(seqs[t]))
        specs += [(q, t, lo, w, 0)]
    gres, _ = run_both(gpu, oracle, names, seqs, specs)
    assert gres["score"].min() > 1000
    scoring = SCORINGS["last_like_asym"]
    with Aligner(device=0, tb_bytes=2 << 30, scoring=scoring) as g:
        gres, _ = run_both(g, OracleEngine(scoring=scoring), names, seqs, specs)
    assert gres["score"].min() > 1000 and not gres["status"].any()


# ---------------------------------------------------------------------------------------------
# thread-block clusters: bands wider than one CTA
# ---------------------------------------------------------------------------------------------

def test_wide_band_cluster_fast_path(gpu, oracle):
    """Full-matrix bands of 26k and 43k diagonals on the s16x2 fast path: clusters of 2 and 4 CTAs."""
    rng = np.random.default_rng(71)
    t1 = randseq(rng, 20000)
    t2 = randseq(rng, 40000)
    seqs = [t1, mutate(rng, t1[12000:17500]), revcomp(mutate(rng, t1[300:6200])), t2, mutate(rng, t2[33000:36000]), mutate(rng, t2[100:3100])]
    names = ["t1", "a", "b", "t2", "c", "d"]
    specs = []
    for q, t, rc in ((1, 0, 0), (2, 0, 1), (4, 3, 0), (5, 3, 0)):
        lo, w = ava.full_band(len(seqs[q]), len(seqs[t]))
        specs.append((q, t, lo, w, PAIR_REVCOMP if rc else 0))
    gres, _ = run_both(gpu, oracle, names, seqs, specs)
    assert gres["score"].min() > 3000 and not gres["status"].any()


def test_wide_band_cluster_long_reads(gpu, oracle):
    """32-bit lanes + cluster: 18 kb reads whose haplotypes differ by a 9 kb insertion (band |dlen| + 512 > 8192)."""
    rng = np.random.default_rng(72)
    tmpl = randseq(rng, 18000)
    ins = randseq(rng, 9000)
    a = mutate(rng, tmpl, sub=0.01, ins=0.005, dele=0.005)        # accurate reads: bridging the 9 kb gap (cost 9024) pays off
    b = mutate(rng, tmpl[:9000] + ins + tmpl[9000:], sub=0.01, ins=0.005, dele=0.005)
    seqs, names = [a, b, revcomp(a)], ["a", "b", "a_rc"]
    lo, w = ava.offset_band(len(a), len(b), slack=256, granule=32)
    assert w > 8192
    specs = [(0, 1, lo, w, 0), (2, 1, lo, w, PAIR_REVCOMP), (1, 0, -(lo + w - 1), w, 0)]
    gres, gcig = run_both(gpu, oracle, names, seqs, specs)
    assert not gres["status"].any() and gres["n_del_bp"][0] >= 9000 and gres["n_ins_bp"][2] >= 9000


def test_wide_band_cluster_rebased_long_reads(gpu, oracle):
    """Rebased s16x2 kernel on a cluster: 22 kb reads whose haplotypes differ by a 16.6 kb insertion (band > 16384
    diagonals = 2 CTAs at KP = 8); the score offsets of the edge warps travel through DSMEM at every rebase."""
    rng = np.random.default_rng(74)
    tmpl = randseq(rng, 22000)
    ins = randseq(rng, 16600)
    a = mutate(rng, tmpl, sub=0.01, ins=0.005, dele=0.005)
    b = mutate(rng, tmpl[:11000] + ins + tmpl[11000:], sub=0.01, ins=0.005, dele=0.005)
    seqs, names = [a, b], ["a", "b"]
    lo, w = ava.offset_band(len(a), len(b), slack=256, granule=32)
    assert w > 16384
    specs = [(0, 1, lo, w, 0), (1, 0, -(lo + w - 1), w, 0)]
    gres, _ = run_both(gpu, oracle, names, seqs, specs)
    assert not gres["status"].any() and gres["n_del_bp"][0] >= 16000 and gres["score"].min() > 10000


@pytest.mark.parametrize("variant", ["fast_cluster", "rebased_cluster", "s32_cluster", "general_cluster", "fast_one_cta"])
def test_alignment_pressed_against_the_upper_band_edge(oracle, monkeypatch, variant):
    """The target carries a 400-base insertion the band does not allow (upper edge 100 diagonals above the start), and the
    band's thread count is not a multiple of 32, so the last in-band thread sits mid-warp with further (out-of-band) warps /
    CTAs behind it: it must see the "outside the band" constants, not the cells those threads compute.
    (Regression: in cluster launches it read the mailbox of the next, out-of-band warp.)"""
    from oracle_engine import OracleEngine
    from svirlpool_b200.aligner import Aligner
    rng = np.random.default_rng(75)
    long_reads = variant in ("rebased_cluster", "s32_cluster")
    length = 17000 if long_reads else 3000
    base = randseq(rng, length)
    q = mutate(rng, base, sub=0.01, ins=0.005, dele=0.005)
    t = base[:length // 2] + randseq(rng, 400) + base[length // 2:]
    w = {"fast_cluster": 16384 + 544, "rebased_cluster": 16384 + 544, "s32_cluster": 8192 + 288, "general_cluster": 8192 + 288,
         "fast_one_cta": 1024 + 544}[variant]
    hi = 100
    specs = [(0, 1, hi - w + 1, w, 0)]
    if variant == "s32_cluster":
        monkeypatch.setenv("SVP_NO_REBASE", "1")
    kw = {"scoring": SCORINGS["matrix_two_piece_asym"]} if variant == "general_cluster" else {"preset": "map-ont"}
    with Aligner(device=0, tb_bytes=4 << 30, **kw) as g:
        gres, _ = run_both(g, OracleEngine(**kw), ["q", "t"], [q, t], specs)
    assert gres["score"][0] > 500 and gres["t_end"][0] < length // 2 + 700


def test_wide_band_cluster_general_scoring():
    from oracle_engine import OracleEngine
    from svirlpool_b200.aligner import Aligner
    scoring = SCORINGS["matrix_two_piece_asym"]
    rng = np.random.default_rng(73)
    t = randseq(rng, 14000)
    seqs, names = [t, mutate(rng, t[9000:12500]), mutate(rng, t[200:3000])], ["t", "a", "b"]
    specs = []
    for q in (1, 2):
        lo, w = ava.full_band(len(seqs[q]), len(t))
        specs.append((q, 0, lo, w, 0))
    with Aligner(device=0, tb_bytes=4 << 30, scoring=scoring) as g:
        gres, _ = run_both(g, OracleEngine(scoring=scoring), names, seqs, specs)
    assert not gres["status"].any()


# ---------------------------------------------------------------------------------------------
# edge cases of the batch interface
# ---------------------------------------------------------------------------------------------

def test_empty_batch_and_single_base(gpu, oracle):
    from svirlpool_b200._abi import PAIR_DTYPE
    batch = ava.PairBatch.build(["a", "b"], ["ACGT" * 10, "ACGT" * 10], [])
    res, cig = gpu.align_batch(batch.packed, batch.seq_off, batch.seq_len, np.zeros(0, dtype=PAIR_DTYPE), 1)
    assert len(res) == 0
    seqs = ["A", "A", "C", "ACGTACGTAC"]
    specs = [(0, 1, -31, 64, 0), (0, 2, -31, 64, 0), (0, 3, -31, 64, 0), (3, 0, -31, 64, PAIR_REVCOMP)]
    gres, _ = run_both(gpu, oracle, ["a", "a2", "c", "d"], seqs, specs)
    assert list(gres["score"]) == [2, 0, 2, 2] and list(gres["status"]) == [0, 1, 0, 0]


def test_cigar_overflow_is_flagged_and_retried(gpu, oracle):
    rng = np.random.default_rng(81)
    tmpl = randseq(rng, 1500)
    seqs = [mutate(rng, tmpl), mutate(rng, tmpl)]
    lo, w = ava.offset_band(len(seqs[0]), len(seqs[1]), slack=100, granule=32)
    tight = ava.PairBatch.build(["a", "b"], seqs, [(0, 1, lo, w, 0)], cap_fn=lambda m, n: 8)
    res, _ = gpu.align_batch(tight.packed, tight.seq_off, tight.seq_len, tight.pairs, tight.cigar_words)
    ores, _ = oracle.align_batch(tight.packed, tight.seq_off, tight.seq_len, tight.pairs, tight.cigar_words)
    assert res[0]["status"] == 2 == ores[0]["status"] and res[0]["cigar_len"] == 0
    for f in ("score", "q_beg", "q_end", "t_beg", "t_end", "n_match", "n_mismatch", "n_ins_bp", "n_del_bp"):
        assert res[0][f] == ores[0][f]                      # counts and coordinates stay valid
    full, cig = ava.run_batch(gpu, tight)                   # the driver retries with worst-case slots
    assert full[0]["status"] == 0 and full[0]["cigar_len"] > 100
    assert sum(n for op, n in cigartuples(cig, full[0]) if op in (0, 1)) == full[0]["q_end"] - full[0]["q_beg"]


def test_band_beyond_cluster_limit_is_flagged_per_pair(gpu):
    """A band wider than a cluster of 8 CTAs covers -> SVP_ST_BAD_PAIR; the other pairs of the batch are aligned normally."""
    rng = np.random.default_rng(82)
    short = randseq(rng, 500)
    long_a = "".join(np.array(list("ACGT"))[rng.integers(0, 4, 140_000)])
    seqs, names = [short, mutate(rng, short), long_a], ["s", "s2", "la"]
    specs = [(0, 1, -499, 1024, 0), (0, 2, -499, 131_072 + 64, 0)]
    batch = ava.PairBatch.build(names, seqs, specs, cap_fn=lambda m, n: 4096)
    res, _ = gpu.align_batch(batch.packed, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
    assert res[0]["status"] == 0 and res[0]["score"] > 400 and res[1]["status"] == 8


@pytest.mark.parametrize("no_rebase", ["0", "1"])
def test_sequences_beyond_shared_memory_sliding_window(oracle, monkeypatch, no_rebase):
    """130 kb vs 120 kb (the size of the reference's INV.15 consensus record, 185 kb query): the images do not fit shared
    memory, the pair runs on a sliding-window variant — of the rebased s16x2 kernel (scores > 200 000 on 16-bit lanes), or
    of the 32-bit lane kernel (SVP_NO_REBASE=1). Bit-exact against the oracle."""
    from svirlpool_b200.aligner import Aligner
    monkeypatch.setenv("SVP_NO_REBASE", no_rebase)
    rng = np.random.default_rng(83)
    codes = rng.integers(0, 4, 130_000)
    a = "".join(np.array(list("ACGT"))[codes])
    edit = codes[:120_000].copy()
    pos = rng.choice(120_000, 2400, replace=False)
    edit[pos] = (edit[pos] + 1 + rng.integers(0, 3, len(pos))) & 3                   # 2 % substitutions
    b = "".join(np.array(list("ACGT"))[edit])
    b = b[:40_000] + b[40_300:90_000] + randseq(rng, 500) + b[90_000:]                 # a 300 bp deletion and a 500 bp insertion
    seqs, names = [a, b, revcomp(b)], ["a", "b", "b_rc"]
    specs = [(1, 0, -768, 1536, 0), (2, 0, -768, 1536, PAIR_REVCOMP), (0, 1, -768, 1536, 0)]
    with Aligner(device=0, preset="map-ont", tb_bytes=4 << 30) as g:
        gres, _ = run_both(g, oracle, names, seqs, specs)
    assert not gres["status"].any() and gres["score"].min() > 200_000 and (gres["n_ins_bp"] + gres["n_del_bp"]).min() >= 800


@pytest.mark.parametrize("scoring_name", [None, "matrix_two_piece_asym"])
def test_forced_sliding_window_equals_whole_images(oracle, monkeypatch, scoring_name):
    """SVP_FORCE_RING pushes ordinary pairs (single-warp, multi-warp and cluster CTAs, both scoring paths) through the
    sliding-window kernel: same results as the oracle."""
    from oracle_engine import OracleEngine
    from svirlpool_b200.aligner import Aligner
    monkeypatch.setenv("SVP_FORCE_RING", "1")
    rng = np.random.default_rng(84)
    names, seqs, specs = noisy_batch(rng, 6, 3000, 512)
    n2, s2, p2 = noisy_batch(rng, 3, 1500, 96)
    off = len(seqs)
    specs = specs + [(q + off, t + off, lo, w, fl) for q, t, lo, w, fl in p2]
    seqs = seqs + s2
    t = randseq(rng, 12_000)
    seqs += [t, mutate(rng, t[7000:10_000])]
    lo, w = ava.full_band(len(seqs[-1]), len(t))                                     # 15k diagonals: a cluster of 2 CTAs
    specs.append((len(seqs) - 1, len(seqs) - 2, lo, w, 0))
    names = [f"r{k}" for k in range(len(seqs))]
    kw = {"preset": "map-ont"} if scoring_name is None else {"scoring": SCORINGS[scoring_name]}
    with Aligner(device=0, tb_bytes=2 << 30, **kw) as g:
        gres, _ = run_both(g, OracleEngine(**kw) if scoring_name else oracle, names, seqs, specs)
    assert not gres["status"].any()


def test_latency_plan_small_calls(oracle, monkeypatch):
    """Calls of at most SVP_LATENCY_PAIRS pairs (off by default: measured slower) are planned for short steps — KP = 4 warps
    wherever the band allows — instead of fewest issue slots. Same results."""
    from svirlpool_b200.aligner import Aligner
    monkeypatch.setenv("SVP_LATENCY_PAIRS", "1024")
    rng = np.random.default_rng(93)
    names, seqs, specs = noisy_batch(rng, 8, 3000, 1024)             # KP = 8 single warp on the throughput plan, 2 x KP = 4 here
    n2, s2, p2 = noisy_batch(rng, 3, 5000, 2080)
    n3, s3, p3 = noisy_batch(rng, 2, 17000, 1056)                    # rebased, multi-warp
    for sq, sp in ((s2, p2), (s3, p3)):
        off = len(seqs)
        seqs += sq
        specs += [(q + off, t + off, lo, w, fl) for q, t, lo, w, fl in sp]
    with Aligner(device=0, preset="map-ont", tb_bytes=2 << 30) as g:
        run_both(g, oracle, [f"r{k}" for k in range(len(seqs))], seqs, specs)


@pytest.mark.parametrize("arena_mb,streams,mode", [(24, "1", "queue"), (24, "8", "queue"), (160, "8", "queue"), (24, "8", "arena"), (12, "8", "queue"),
                                                   (24, "4", "waves"), (12, "4", "waves"), (160, "4", "waves")])
def test_arena_pressure_and_stream_counts(oracle, monkeypatch, arena_mb, streams, mode):
    """A small arena under the three schedulers (SVP_MODE). queue / arena: the slices of short pairs come from the
    per-SM pools, the long pairs' from the global pool, and most CTAs have to wait for a slice — with a queue they walk queued pairs
    themselves while they wait (pair_begin), so the call ends even where the arena holds fewer pairs than the GPU has CTA slots.
    waves: many small waves, whole-arena waves where a pair exceeds half the arena. Results do not depend on who waited or walked
    (a lane of the walker kernel, a waiting CTA's warp, the pair's own CTA, a bulk kernel) or on the number of launch streams."""
    from svirlpool_b200.aligner import Aligner
    monkeypatch.setenv("SVP_FWD_STREAMS", streams)
    monkeypatch.setenv("SVP_MODE", mode)
    monkeypatch.setenv("SVP_QUEUE_MIN_PAIRS", "0")                    # queue even this small call
    rng = np.random.default_rng(92)
    names, seqs, specs = noisy_batch(rng, 48, 2200, 512)
    n2, s2, p2 = noisy_batch(rng, 3, 14000, 544)                      # long pairs: slices beyond a per-SM region
    off = len(seqs)
    seqs += s2
    specs += [(q + off, t + off, lo, w, fl) for q, t, lo, w, fl in p2]
    with Aligner(device=0, preset="map-ont", tb_bytes=arena_mb << 20) as g:
        for _ in range(2):
            run_both(g, oracle, [f"r{k}" for k in range(len(seqs))], seqs, specs)
        assert g.stats()["waves"] >= 1                                 # launch groups / waves


@pytest.mark.parametrize("mode", ["queue", "arena", "waves"])
def test_schedulers_on_a_mixed_batch(oracle, monkeypatch, mode):
    """The three schedulers on one batch that mixes many short pairs (queued: one walk per lane of the walker kernel), multi-warp
    bands and a few long pairs (walked inside their own CTA), with the default arena."""
    from svirlpool_b200.aligner import Aligner
    monkeypatch.setenv("SVP_MODE", mode)
    monkeypatch.setenv("SVP_QUEUE_MIN_PAIRS", "64")
    rng = np.random.default_rng(93)
    names, seqs, specs = noisy_batch(rng, 150, 1500, 384)
    for n_pairs, length, band in ((40, 2500, 640), (12, 3000, 1536), (3, 13000, 768)):
        _n, s2, p2 = noisy_batch(rng, n_pairs, length, band)
        off = len(seqs)
        seqs += s2
        specs += [(q + off, t + off, lo, w, fl) for q, t, lo, w, fl in p2]
    with Aligner(device=0, preset="map-ont") as g:
        run_both(g, oracle, [f"r{k}" for k in range(len(seqs))], seqs, specs)
        run_both(g, oracle, [f"r{k}" for k in range(len(seqs))], seqs, specs)


def test_pair_beyond_the_arena_is_an_error():
    """A pair whose traceback rows exceed the whole arena cannot run: the call fails with SVP_ERR_NOMEM (AlignerError), it is not
    silently dropped."""
    from svirlpool_b200 import aligner as al
    rng = np.random.default_rng(94)
    names, seqs, specs = noisy_batch(rng, 1, 9000, 1024)
    b = ava.PairBatch.build(names, seqs, specs)
    with al.Aligner(device=0, preset="map-ont", tb_bytes=4 << 20) as g:
        with pytest.raises(al.AlignerError) as e:
            g.align_batch(b.packed, b.seq_off, b.seq_len, b.pairs, b.cigar_words)
        assert e.value.returncode == -3


def test_walk_cases_zdrop_long_gaps_reverse_strands(gpu, oracle):
    """The walk on a mixed batch incl. a Z-trimmed inversion, long gaps and reverse strands."""
    rng = np.random.default_rng(91)
    names, seqs, specs = noisy_batch(rng, 10, 2500, 544)
    ref = randseq(rng, 6000)
    inv = ref[:2000] + revcomp(ref[2000:4000]) + ref[4000:]
    off = len(seqs)
    seqs += [ref, inv, mutate(rng, ref[:1500] + ref[2100:], sub=0.01, ins=0.005, dele=0.005)]
    for q in (off + 1, off + 2):
        lo, w = ava.full_band(len(seqs[q]), len(ref))
        specs += [(q, off, lo, w, 0), (q, off, lo, w, PAIR_REVCOMP)]
    names = [f"r{k}" for k in range(len(seqs))]
    gres, _ = run_both(gpu, oracle, names, seqs, specs)
    assert gres["n_del_bp"].max() >= 600


def test_two_handles_share_a_device(oracle):
    """Two Aligner handles with the default (bounded) arena coexist on one GPU and give the same rows — the reference runs
    several consensus jobs per node (main.smk:649-707)."""
    from svirlpool_b200.aligner import Aligner
    rng = np.random.default_rng(95)
    names, seqs, specs = noisy_batch(rng, 12, 2500, 544)
    with Aligner(device=0, preset="map-ont") as g1, Aligner(device=0, preset="map-ont") as g2:
        r1, _ = run_both(g1, oracle, names, seqs, specs)
        r2, _ = run_both(g2, oracle, names, seqs, specs)
        assert np.array_equal(r1["score"], r2["score"])


def test_pipelined_submits_equal_blocking_calls(oracle):
    """svp_submit_batch_device x 4 back to back (two batches in flight, the per-call buffers alternate across calls; a small
    arena makes their CTAs compete for slices) followed by one svp_sync gives the rows and CIGARs of the blocking host-buffer
    calls, which equal the oracle."""
    import torch
    from svirlpool_b200._abi import RESULT_DTYPE
    from svirlpool_b200.aligner import Aligner
    rng = np.random.default_rng(101)
    batches = []
    for n_pairs, length, band in ((10, 2500, 512), (3, 3000, 1056), (14, 1800, 256), (1, 2000, 512)):
        names, seqs, specs = noisy_batch(rng, n_pairs, length, band)
        batches.append(ava.PairBatch.build(names, seqs, specs))
    dev = torch.device("cuda", 0)
    with Aligner(device=0, preset="map-ont", tb_bytes=12 << 20) as g:
        expected = []
        for b in batches:
            res, cig = g.align_batch(b.packed, b.seq_off, b.seq_len, b.pairs, b.cigar_words)
            ores, ocig = oracle.align_batch(b.packed, b.seq_off, b.seq_len, b.pairs, b.cigar_words)
            assert_same(b, res, cig, ores, ocig)
            expected.append((res, cig))
        assert g.stats()["waves"] >= 1
        bufs = []
        g.stats_total(reset=True)
        for b in batches:
            d_packed = torch.from_numpy(b.packed).to(dev)
            d_res = torch.zeros(len(b.pairs) * RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
            d_cig = torch.zeros(b.cigar_words, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            g.submit_batch_device(d_packed.data_ptr(), b.packed.nbytes, b.seq_off, b.seq_len, b.pairs, d_res.data_ptr(), d_cig.data_ptr(),
                                  b.cigar_words)
            bufs.append((d_packed, d_res, d_cig))
        g.sync()
        total = g.stats_total()
        assert total["waves"] >= len(batches) and total["cells"] == sum(int(e[0]["cells"].sum()) for e in expected)
        for b, (res, cig), (_p, d_res, d_cig) in zip(batches, expected, bufs):
            got = d_res.cpu().numpy().view(RESULT_DTYPE)
            gcig = d_cig.cpu().numpy().view(np.uint32)
            for f in INT_FIELDS:
                if f != "cigar_len":
                    assert np.array_equal(got[f], res[f]), f
            for k in range(len(got)):
                assert cigartuples(gcig, got[k]) == cigartuples(cig, res[k])


# ---------------------------------------------------------------------------------------------
# band widths between 512 and 1024 diagonals per warp (KP = 5, 6, 7) and their multi-warp CTAs
# ---------------------------------------------------------------------------------------------

def fixed_band_batch(rng, n_pairs, length, band_w, err=(0.035, 0.03, 0.035)):
    """Noisy pairs aligned in a band of exactly band_w diagonals centred on the end-to-end diagonal; every third pair carries an
    insertion / deletion of up to a third of the band."""
    seqs, specs = [], []
    for p in range(n_pairs):
        tmpl = randseq(rng, length + int(rng.integers(-50, 50)))
        a, b = mutate(rng, tmpl, *err), mutate(rng, tmpl, *err)
        if p % 3 == 0:
            pos = len(b) // 2
            sv = int(rng.integers(20, band_w // 3))
            b = b[:pos] + (randseq(rng, sv) if p % 2 else "") + b[pos + (0 if p % 2 else sv):]
        rc = bool(rng.integers(2))
        seqs += [revcomp(a) if rc else a, b]
        mid = (len(b) - len(a)) // 2
        specs.append((2 * p, 2 * p + 1, mid - band_w // 2, band_w, PAIR_REVCOMP if rc else 0))
    return [f"r{k}" for k in range(len(seqs))], seqs, specs


@pytest.mark.parametrize("ring_min", ["0", "1000000000"])
@pytest.mark.parametrize("band_w,kp,warps", [(640, 5, 1), (768, 6, 1), (896, 7, 1), (672, 6, 1), (1280, 5, 2), (1536, 6, 2), (1792, 7, 2),
                                              (2304, 6, 3), (3200, 5, 5), (5376, 7, 6), (10752, 7, 12), (13440, 7, 15),
                                              (1152, 12, 1), (1536, 12, 1), (1056, 12, 1)])
def test_band_widths_kp_5_6_7(oracle, monkeypatch, band_w, kp, warps, ring_min):
    """Every KP = 5 / 6 / 7 geometry (single warp, CTAs of 2..16 warps whose warps are linked by mbarriers instead of a CTA
    barrier; whole images and sliding windows) and the single-warp KP = 12 geometry (bands of up to 1536 diagonals) against the
    oracle, and the planner picks the geometry the test names. (By default KP = 5 / 6 / 7 serve single-warp CTAs only: the
    multi-warp cases switch them on, SVP_KP_MULTI, and KP = 12 off.)"""
    from svirlpool_b200 import aligner as al
    monkeypatch.setenv("SVP_RING_MIN", ring_min)
    monkeypatch.setenv("SVP_MODE", "waves")
    monkeypatch.setenv("SVP_SMALL_CALL_PAIRS", "0")                   # (small calls keep to KP = 4 / 8)
    if warps > 1:
        monkeypatch.setenv("SVP_KP_MULTI", "1")
    if kp == 12:
        monkeypatch.setenv("SVP_KP_MASK", str(0x11F))                 # off by default (measured slower than three KP = 4 warps)
    rng = np.random.default_rng(band_w)
    names, seqs, specs = fixed_band_batch(rng, 6, 2600 if band_w < 4000 else 4200, band_w)
    batch = ava.PairBatch.build(names, seqs, specs)
    plan = al.plan_pairs(al.scoring_preset("map-ont"), batch.packed.nbytes, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
    assert {(int(p["kp"]), int(p["warps"])) for p in plan} == {(kp, warps)}
    with al.Aligner(device=0, preset="map-ont", tb_bytes=4 << 30) as g:
        gres, _ = run_both(g, oracle, names, seqs, specs)
    assert not gres["status"].any() and gres["score"].min() > 800


@pytest.mark.parametrize("band_w", [640, 768, 896, 1536, 2688])
def test_band_widths_kp_5_6_7_long_reads_rebased(gpu, oracle, band_w):
    """The same widths on the rebased kernel (18 kb reads: scores beyond the s16 range, per-warp score offsets)."""
    rng = np.random.default_rng(1000 + band_w)
    names, seqs, specs = fixed_band_batch(rng, 3, 18000, band_w, err=(0.02, 0.015, 0.015))
    gres, _ = run_both(gpu, oracle, names, seqs, specs)
    assert not gres["status"].any() and gres["score"].max() > 20000


# ---------------------------------------------------------------------------------------------
# letters outside ACGT (N runs, IUPAC codes): the 1-bit-per-base mask behind a sequence's codes
# ---------------------------------------------------------------------------------------------

def with_n_runs(rng, seq, n_runs=3, max_len=40):
    s = list(seq)
    for _ in range(n_runs):
        pos = int(rng.integers(0, max(1, len(s) - max_len)))
        for k in range(pos, min(len(s), pos + int(rng.integers(1, max_len)))):
            s[k] = "NRYK"[int(rng.integers(4))]
    return "".join(s)


@pytest.mark.parametrize("variant", ["fast_images", "fast_windows", "rebased_long", "s32_long", "general", "general_long_s32"])
def test_reads_with_letters_outside_acgt(oracle, monkeypatch, variant):
    """N runs (and IUPAC letters) in queries and targets, forward and reverse-complemented, on every kernel family: an N never
    matches — a poly-A query against an N run scores mismatches (it used to pack as A and match) — and the rows equal the
    oracle's, which scores such a cell with the lowest matrix entry."""
    from oracle_engine import OracleEngine
    from svirlpool_b200.aligner import Aligner
    rng = np.random.default_rng(["fast_images", "fast_windows", "rebased_long", "s32_long", "general", "general_long_s32"].index(variant) + 300)
    long_reads = variant in ("rebased_long", "s32_long", "general_long_s32")
    if variant == "fast_windows":
        monkeypatch.setenv("SVP_RING_MIN", "0")
    if variant == "fast_images":
        monkeypatch.setenv("SVP_RING_MIN", "1000000000")
    if variant == "s32_long":
        monkeypatch.setenv("SVP_NO_REBASE", "1")
    kw = {"scoring": SCORINGS["last_like_asym"]} if variant.startswith("general") else {"preset": "map-ont"}
    length = 18000 if variant in ("rebased_long", "s32_long") else (9000 if variant == "general_long_s32" else 2600)
    names, seqs, specs = [], [], []
    for p, band in enumerate((544, 768, 1536, 2304) if not long_reads else (544, 1056)):
        tmpl = randseq(rng, length)
        a, b = mutate(rng, tmpl, 0.02, 0.015, 0.015), mutate(rng, tmpl, 0.02, 0.015, 0.015)
        a, b = with_n_runs(rng, a), (with_n_runs(rng, b) if p % 2 == 0 else b)
        rc = p % 2 == 1
        seqs += [revcomp(a) if rc else a, b]
        mid = (len(b) - len(a)) // 2
        specs.append((2 * p, 2 * p + 1, mid - band // 2, band, PAIR_REVCOMP if rc else 0))
    k = len(seqs)
    seqs += ["A" * 300, "N" * 300, "ACGT" * 80, "ACGT" * 30 + "N" * 80 + "ACGT" * 30]       # poly-A vs an N run: no hit; a masked stretch inside a repeat
    lo, w = ava.full_band(300, 300)
    specs += [(k, k + 1, lo, w, 0), (k + 1, k + 1, lo, w, 0)]
    lo, w = ava.full_band(320, 320)
    specs += [(k + 2, k + 3, lo, w, 0), (k + 3, k + 2, lo, w, PAIR_REVCOMP)]
    names = [f"s{i}" for i in range(len(seqs))]
    with Aligner(device=0, tb_bytes=4 << 30, **kw) as g:
        gres, _ = run_both(g, OracleEngine(**kw), names, seqs, specs)
    n_pairs = len(specs) - 4
    assert (gres["status"][:n_pairs] == 0).all() and gres["score"][:n_pairs].min() > 1000
    assert gres["status"][n_pairs] == 1 and gres["status"][n_pairs + 1] == 1                # nothing aligns to an N run, not even itself
    assert gres["n_match"][n_pairs + 2] <= 240 and gres["n_match"][n_pairs + 3] <= 240       # the 80 masked bases never count as matches
