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


def test_multiple_waves_small_arena(oracle):
    """A 16 MiB arena forces several waves; results must not depend on the wave split."""
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
        assert small.stats()["waves"] > 1
