"""The N>1 path on CPU: two gloo ranks shard regions with the LPT scheduler, align them with the oracle
engine and gather on rank 0; the result must equal the single-process run."""
import os
import pickle

import numpy as np
import subprocess
import sys
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent("""
    import os, pickle, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
    import numpy as np
    import torch.distributed as dist
    from oracle_engine import OracleEngine
    from svirlpool_b200 import consensus, scheduler, sharding, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("gloo")
    ids = [11, 12, 13, 14, 15]
    shapes = [synth.region_config2_shape(r, mean_len=700, sd_len=200) for r in ids]
    costs = [ln * (abs(sv) + 600) for ln, sv in shapes]
    eng = OracleEngine("map-ont")

    def work(rid):
        reg = synth.region_config2(rid, n_reads=5, mean_len=700, sd_len=200)
        sim, names, recs = consensus.ava_similarity(eng, reg.read_strings(), strands=dict(zip(reg.names, reg.strands)))
        return (rid, names, np.round(sim, 12).tolist(), [(r.query_name, r.reference_name, r.reference_start, r.cigarstring) for r in recs])

    def work_many(rids):
        regs = [synth.region_config2(rid, n_reads=5, mean_len=700, sd_len=200) for rid in rids]
        res = consensus.ava_similarity_summary_many(eng, [r.read_strings() for r in regs], [dict(zip(r.names, r.strands)) for r in regs],
                                                    granule=32)
        return [(rid, names, np.round(sim, 12).tolist(), aln_len.tolist()) for rid, (sim, names, aln_len) in zip(rids, res)]

    out = sharding.run_sharded(ids, costs, work, dist if world > 1 else None)
    out_many = sharding.run_sharded(ids, costs, None, dist if world > 1 else None, process_many=work_many)
    if out is not None:
        with open(sys.argv[1], "wb") as f:
            pickle.dump({{"out": out, "out_many": out_many, "bins": scheduler.lpt_assign(costs, max(world, 1))}}, f)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
""")


def run(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT)))
    out = tmp_path / f"out{world}.pkl"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    if world == 1:
        cmd = [sys.executable, str(script), str(out)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", str(script), str(out)]
    subprocess.run(cmd, check=True, env=env, timeout=300, capture_output=True)
    return pickle.loads(out.read_bytes())


def test_two_ranks_equal_one_rank(tmp_path):
    one, two = run(tmp_path, 1), run(tmp_path, 2)
    assert one["out"] == two["out"]
    # one engine call per rank (process_many): same matrices whatever the sharding, and the similarity equals the record pipeline's
    assert one["out_many"] == two["out_many"] and [r[0] for r in two["out_many"]] == [11, 12, 13, 14, 15]
    for a, b in zip(one["out"], one["out_many"]):
        assert a[1] == b[1] and np.allclose(np.array(a[2]), np.array(b[2]), rtol=1e-9, atol=1e-12)
    assert [r[0] for r in two["out"]] == [11, 12, 13, 14, 15]
    assert sorted(k for b in two["bins"] for k in b) == [0, 1, 2, 3, 4] and all(two["bins"])
