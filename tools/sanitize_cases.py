"""Small workload that touches every kernel family once — for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py
Single-warp and multi-warp fast path (KP = 4 .. 8, mbarrier links between warps), general scoring, a 2-CTA cluster, the
sliding-window kernels (forced), the in-kernel walk and the device-side arena (a small arena makes CTAs wait for slices).
Results are compared with the CPU oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from oracle_engine import OracleEngine  # noqa: E402
from svirlpool_b200 import ava  # noqa: E402
from svirlpool_b200._abi import PAIR_REVCOMP, cigartuples  # noqa: E402
from svirlpool_b200.aligner import Aligner  # noqa: E402


def randseq(rng, n):
    return "".join("ACGT"[k] for k in rng.integers(0, 4, n))


def mutate(rng, s, rate=0.08):
    out = []
    for ch in s:
        r = rng.random()
        if r < rate / 3:
            continue
        out.append("ACGT"[rng.integers(4)] if r < 2 * rate / 3 else ch)
        if rng.random() < rate / 3:
            out.append("ACGT"[rng.integers(4)])
    return "".join(out)


def main():
    rng = np.random.default_rng(7)
    a = randseq(rng, 900)
    seqs = [a, mutate(rng, a), mutate(rng, a[:400] + a[460:]), randseq(rng, 17000)]
    seqs.append(mutate(rng, seqs[3][9000:9300]))
    names = [f"s{k}" for k in range(len(seqs))]
    specs = [(1, 0, -63, 128, 0), (2, 0, -600, 1184, 0), (2, 0, -600, 1184, PAIR_REVCOMP), (1, 2, -1100, 2208, 0),
             (1, 0, -320, 640, 0), (2, 0, -400, 768, 0), (1, 2, -440, 896, 0), (1, 0, -640, 1280, 0), (2, 1, -770, 1536, PAIR_REVCOMP)]
    lo, w = ava.full_band(len(seqs[4]), len(seqs[3]))
    specs.append((4, 3, lo, w, 0))                              # 17k diagonals: cluster of 2 CTAs
    batch = ava.PairBatch.build(names, seqs, specs)
    ok = True
    for label, env, kw in (("fast", {}, {"preset": "map-ont"}), ("general", {}, {"preset": "last"}),
                           ("ring", {"SVP_FORCE_RING": "1"}, {"preset": "map-ont"}), ("small-arena", {}, {"preset": "map-ont", "tb_bytes": 600 << 20})):
        if label == "ring" and os.environ.get("SVP_FORCE_RING") != "1":
            continue                                            # the switch is read once per process: run this case in its own process
        if label != "ring" and os.environ.get("SVP_FORCE_RING") == "1":
            continue
        os.environ.update(env)
        with Aligner(device=0, **{"tb_bytes": 4 << 30, **kw}) as g:
            gres, gcig = g.align_batch(batch.packed, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
        ores, ocig = OracleEngine(**{k: v for k, v in kw.items() if k != "tb_bytes"}).align_batch(batch.packed, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
        same = all(np.array_equal(gres[f], ores[f]) for f in ("score", "status", "q_beg", "q_end", "t_beg", "t_end", "cigar_len"))
        same = same and all(cigartuples(gcig, gres[k]) == cigartuples(ocig, ores[k]) for k in range(len(gres)))
        print(label, "ok" if same else "MISMATCH", gres["score"].tolist())
        ok = ok and same
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
