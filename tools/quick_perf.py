"""Scratch timing of the CUDA path on config-2 regions (not the bench; see bench.py)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from svirlpool_b200 import synth
from svirlpool_b200._abi import PAIR_DTYPE, pack_codes
from svirlpool_b200.aligner import Aligner
from svirlpool_b200.ava import cigar_cap_for

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gran = int(sys.argv[2]) if len(sys.argv) > 2 else 512
t0 = time.time()
seqs, specs = [], []
for r in range(R):
    reg = synth.region_config2(r)
    base = len(seqs)
    seqs += reg.reads
    for (q, t, lo, w, fl) in synth.region_pairs(reg, granule=gran):
        specs.append((base + q, base + t, lo, w, fl))
packed, off, lens = pack_codes(seqs)
pairs = np.zeros(len(specs), dtype=PAIR_DTYPE)
pos = 0
for k, (q, t, lo, w, fl) in enumerate(specs):
    cap = cigar_cap_for(int(lens[q]), int(lens[t]))
    pairs[k] = (q, t, lo, w, fl, pos, cap, 0); pos += cap
print(f"gen {time.time()-t0:.1f}s pairs={len(pairs)} mean band={pairs['band_w'].mean():.0f}", flush=True)
al = Aligner(0, tb_bytes=int(sys.argv[3]) << 30 if len(sys.argv) > 3 else 0)
for rep in range(3):
    t0 = time.time()
    res, cig = al.align_batch_nocopy(packed, off, lens, pairs, pos)
    dt = time.time() - t0
    st = al.stats()
    print(f"rep{rep}: wall {dt*1e3:.1f} ms  fwd {st['fwd_ms']:.1f} ms tb {st['tb_ms']:.1f} ms total {st['total_ms']:.1f} ms waves {st['waves']} "
          f"cells {st['cells']/1e9:.2f} G  fwd GCUPS {st['cells']/st['fwd_ms']/1e6:.1f}  e2e GCUPS {st['cells']/dt/1e9:.1f} tb_bytes {st['tb_bytes']/1e9:.2f} GB", flush=True)
print("status counts", np.unique(res["status"], return_counts=True), "mean score", res["score"].mean(), "mean sv_dist", res["sv_dist"].mean())
bw = np.unique(pairs["band_w"], return_counts=True); print("band_w hist", dict(zip(bw[0].tolist(), bw[1].tolist())))
