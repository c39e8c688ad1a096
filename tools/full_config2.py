"""The whole BASELINE config 2 set once: 10 000 regions x 30 reads x 8 kb (4.35 M pairs), in calls of `--chunk` regions through the
host-buffer API, with the host's share reported beside the device's: synthetic generation (not part of the system), batch
building (2-bit packing + pair lists: svp_pack_2bit / svp_ava_pairs), the engine call (H2D, kernels, D2H), and the per-container
similarity matrices from the summary rows (consensus_lib.similarity_from_summary). Also times the record path (ava.ava_records:
AlignedRecord objects with CIGARs) on a sample of containers.

    python tools/full_config2.py [--regions 10000] [--chunk 96] [--records-sample 8]
"""
import argparse
import json
import sys
import time
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from svirlpool_b200 import ava, consensus, consensus_lib, synth  # noqa: E402


def gen(rid):
    reg = synth.region_config2(rid)
    return reg.read_strings(), dict(zip(reg.names, reg.strands))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=10_000)
    ap.add_argument("--chunk", type=int, default=96)
    ap.add_argument("--records-sample", type=int, default=8)
    ap.add_argument("--workers", type=int, default=12)
    args = ap.parse_args()
    from svirlpool_b200.aligner import Aligner
    al = Aligner(device=0, preset="map-ont")
    t = {"generate": 0.0, "build": 0.0, "engine": 0.0, "similarity": 0.0}
    cells = pairs = 0
    t_all = time.perf_counter()
    with ProcessPoolExecutor(args.workers) as pool:
        nxt = pool.map(gen, range(0, min(args.chunk, args.regions)))
        for first in range(0, args.regions, args.chunk):
            t0 = time.perf_counter()
            cur = list(nxt)
            t["generate"] += time.perf_counter() - t0
            if first + args.chunk < args.regions:         # the next chunk's synthetic reads are generated while this one is aligned
                nxt = pool.map(gen, range(first + args.chunk, min(first + 2 * args.chunk, args.regions)))
            containers, strands = [c[0] for c in cur], [c[1] for c in cur]
            t0 = time.perf_counter()
            batch, spans, bounds = ava.build_ava_many(containers, strands, None, 300, 128, False)
            t1 = time.perf_counter()
            res = ava.run_summary(al, batch)
            t2 = time.perf_counter()
            for c, (base, n) in enumerate(spans):
                lo, hi = bounds[c], bounds[c + 1]
                p = batch.pairs[lo:hi]
                names = list(containers[c])
                consensus_lib.similarity_from_summary(names, {k: len(v) for k, v in containers[c].items()}, p["q"].astype(np.int64) - base,
                                                      p["t"].astype(np.int64) - base, res[lo:hi])
            t3 = time.perf_counter()
            t["build"] += t1 - t0; t["engine"] += t2 - t1; t["similarity"] += t3 - t2
            cells += int(res["cells"].sum()); pairs += len(res)
    wall = time.perf_counter() - t_all
    # record path on a sample: rows + CIGARs -> AlignedRecord objects -> directed signals -> matrix
    rec = {"containers": 0, "engine_s": 0.0, "records_s": 0.0, "signals_matrix_s": 0.0}
    for rid in range(args.records_sample):
        reads, st = gen(rid)
        names = list(reads)
        t0 = time.perf_counter()
        batch, meta = ava.PairBatch.build_ava(names, [reads[n] for n in names], st, None, 300, 128, False)
        r, cig = ava.run_batch(al, batch)
        t1 = time.perf_counter()
        records = ava.ava_records(names, [len(reads[n]) for n in names], meta, r, cig)
        t2 = time.perf_counter()
        directed = consensus_lib.parse_directed_signals_from_ava(records, 12, 100)
        consensus_lib.pairwise_similarity_matrix(directed, {}, {n: len(reads[n]) for n in names}, all_read_names=names, densities_weight=0.0)
        t3 = time.perf_counter()
        rec["containers"] += 1; rec["engine_s"] += t1 - t0; rec["records_s"] += t2 - t1; rec["signals_matrix_s"] += t3 - t2
    al.close()
    sys_s = t["build"] + t["engine"] + t["similarity"]
    out = {"workload": f"config 2, regions 0..{args.regions - 1}, {args.chunk} regions per call, svp_align_batch_summary", "regions": args.regions, "pairs": pairs,
           "gcells": cells / 1e9, "seconds": {k: round(v, 3) for k, v in t.items()}, "wall_s": round(wall, 2),
           "regions_per_s_system": args.regions / sys_s, "regions_per_s_engine_only": args.regions / t["engine"],
           "gcups_system": cells / sys_s / 1e9, "host_ms_per_container": {"build": 1e3 * t["build"] / args.regions, "similarity_from_summary": 1e3 * t["similarity"] / args.regions},
           "note": "system = batch building + engine call (H2D, kernels, D2H) + similarity matrices; synthetic read generation is excluded (waited for only when the generator pool is slower than the GPU)",
           "record_path_ms_per_container": {k[:-2]: round(1e3 * v / max(rec["containers"], 1), 2) for k, v in rec.items() if k.endswith("_s")}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
