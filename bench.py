#!/usr/bin/env python
"""bench.py — throughput of the alignment hot path on BASELINE.json config 2
("synthetic 10k candidate regions x 30 ONT-like reads x 8 kb, all-vs-all pairwise on 1 B200").

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (oracle on host cores)

A step = one pass of the hot path (forward DP + traceback + per-pair counts/sv_dist) over one batch
of `--regions` regions per GPU (435 read pairs each), drawn from the config's region set (region r
has seed 10000+r). `value` is GCUPS (1e9 spec DP cells per second, whole job) with the packed reads
already resident in HBM; `e2e` is the same metric through the public host-buffer API
(svirlpool_b200.aligner.Aligner.align_batch -> svp_align_batch) with pinned host inputs, H2D and D2H
inside the timed region. One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_REGIONS_CONFIG = 10_000
METRIC = "gcups_ava_config2"
SLACK, GRANULE = 300, 128


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def int_pipe_peak() -> dict:
    """Measured cell-body rate of the two-piece s16x2 recurrence (tools/intpipe_bench.cu)."""
    best = None
    for f in sorted((ROOT / "profiles").glob("intpipe_peaks_r*.json")):
        d = json.loads(f.read_text())
        for r in d["results"]:
            if r["op"].startswith("cell_two_piece"):
                best = {"gcups": r["tera_lane_ops_per_s"] * 1e3, "source": f"profiles/{f.name}"}
    return best or {"gcups": 3056.0, "source": "profiles/intpipe_peaks_r01.json (hard-coded copy)"}


def alu_pipe_peak(sm_mhz: float | None) -> dict:
    """SURVEY.md §8(d) integer roofline: 14 int-ops per two-piece cell, 2 cells per s16x2 instruction, on the MEASURED
    ALU-pipe issue rate (tools/pipe_mix_bench.cu -> profiles/pipe_mix_r01.json: warp-instructions/clk/SMSP of
    VIADDMNMX.S16x2, 4 SMSPs x 148 SMs) at the SM clock sampled during the timed region."""
    ipc, src = 0.5, "nominal 0.5 warp-instr/clk/SMSP"
    f = ROOT / "profiles" / "pipe_mix_r01.json"
    if f.exists():
        d = json.loads(f.read_text())
        for r in d["results"]:
            if r["mix"] == "VIADDMNMX.S16x2":
                ipc, src = float(r["ipc_per_smsp"]), "profiles/pipe_mix_r01.json"
    mhz = sm_mhz or 1965.0
    lane_ops = ipc * 32 * 4 * 148 * mhz * 1e6          # 16-bit-pair lane-ops per second on the ALU pipe
    return {"gcups": lane_ops * 2 / 14 / 1e9, "ipc_per_smsp": ipc, "sm_mhz": mhz, "int_ops_per_cell": 14, "source": src}


def traffic_ratio() -> dict | None:
    """DRAM bytes / algorithmic bytes of the forward kernels from the latest ncu --set full capture in profiles/."""
    best = None
    for f in sorted((ROOT / "profiles").glob("ncu_fwd_r*.json")):
        d = json.loads(f.read_text())
        if "traffic_over_algorithmic" in d:
            best = {"ratio": float(d["traffic_over_algorithmic"]), "source": f"profiles/{f.name}"}
    return best


class ClockSampler:
    """nvidia-smi clocks/throttle reasons of one GPU, sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows: list[list[str]] = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


WORKLOADS = {
    "config2": "config2: 10k regions x 30 ONT-like reads x 8 kb, all-vs-all (435 pairs/region), map-ont scoring",
    "config3": "config3: tandem-repeat regions, 20 spanning reads of 10-50 kb, haplotypes 2-20 kb apart, all-vs-all, band |dlen|+512 (rebased s16x2 lanes, clusters for bands beyond 16k diagonals)",
    "config4": "config4: HG002-shape loci (core log-normal 3 kb, coverage Poisson 30), all-vs-all + padded consensus -> reference window on both strands",
}


def build_batch(region_ids: list[int], workload: str = "config2"):
    """Packed reads + pair descriptors of a list of regions of the workload (svirlpool_b200.workloads)."""
    from svirlpool_b200 import workloads
    if workload == "config3":
        return workloads.batch_config3(region_ids)
    if workload == "config4":
        return workloads.batch_config4(region_ids)
    return workloads.batch_config2(region_ids, SLACK, GRANULE)


def regions_for_step(step: int, world: int, per_gpu: int, workload: str = "config2") -> list[list[int]]:
    """Region ids of one step, assigned to ranks by the size-balanced scheduler (LPT on estimated cells)."""
    from svirlpool_b200 import scheduler, synth
    first = (step * world * per_gpu) % N_REGIONS_CONFIG
    ids = [(first + k) % N_REGIONS_CONFIG for k in range(world * per_gpu)]
    if world == 1:
        return [ids]
    costs = []
    for rid in ids:
        if workload == "config3":
            h0, h1 = synth.region_config3(rid, n_reads=1).meta["hap_lengths"]
            costs.append(h1 * (h1 - h0 + 512))
            continue
        if workload == "config4":
            costs.append(synth.locus_cost(synth.locus_config4(rid)))
            continue
        length, sv = synth.region_config2_shape(rid)
        costs.append(length * (abs(sv) // 2 + 2 * SLACK + GRANULE))
    bins = scheduler.lpt_assign(costs, world)
    return [[ids[k] for k in b] for b in bins]


def cpu_arm(sample_regions: int, threads: int, keep_rows: bool = False) -> dict:
    """The CPU oracle (port of the spec; the reference's own path is an external minimap2 binary that is
    not available) on a bounded sample of the same workload: regions 0 .. sample_regions-1 of config 2."""
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_engine import OracleEngine
    b = build_batch(list(range(sample_regions)))
    eng = OracleEngine("map-ont", threads=threads)
    t0 = time.perf_counter()
    res, cig = eng.align_batch(b["packed"], b["off"], b["lens"], b["pairs"], b["cigar_words"])
    dt = time.perf_counter() - t0
    cells = int(res["cells"].sum())
    out = {"value": cells / dt / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
           "sample": f"{sample_regions} region(s) of config 2 ({len(res)} pairs, {cells / 1e9:.2f} Gcells) in {dt:.2f} s on {threads} thread(s)",
           "regions_per_s": sample_regions / dt, "_seconds": dt, "_cells": cells}
    if keep_rows:
        out["_rows"] = (b, res, cig)
    return out


PARITY_FIELDS = ("score", "status", "q_beg", "q_end", "t_beg", "t_end", "n_match", "n_mismatch", "n_ins_ops", "n_ins_bp",
                 "n_del_ops", "n_del_bp", "cigar_len", "cells")


def parity_check(al, rows) -> dict:
    """GPU rows of the CPU arm's sample against the oracle rows the CPU arm just produced (outside every timed region):
    integer fields and CIGARs bit for bit, sv_dist within 1e-6 relative."""
    from svirlpool_b200._abi import cigartuples
    b, ores, ocig = rows
    gres, gcig = al.align_batch(b["packed"], b["off"], b["lens"], b["pairs"], b["cigar_words"])
    bad = np.zeros(len(ores), dtype=bool)
    for f in PARITY_FIELDS:
        bad |= gres[f] != ores[f]
    bad |= ~np.isclose(gres["sv_dist"], ores["sv_dist"], rtol=1e-6, atol=0)
    for k in range(len(ores)):
        if not bad[k] and cigartuples(gcig, gres[k]) != cigartuples(ocig, ores[k]):
            bad[k] = True
    return {"pairs": int(len(ores)), "mismatches": int(bad.sum()), "regions": f"config 2 regions 0..{b['n_regions'] - 1} at full size",
            "fields": list(PARITY_FIELDS) + ["cigar", "sv_dist (rtol 1e-6)"], "against": "oracle/svp_oracle.c (the rows of the cpu_baseline leg)"}


def cells_at_granule(region_ids: list[int], granule: int) -> int:
    """Spec cells of the same regions with the bands rounded up to `granule` instead of the bench's GRANULE (host only)."""
    from svirlpool_b200 import scheduler, synth
    total = 0
    for rid in region_ids:
        reg = synth.region_config2(rid)
        lens = [len(r) for r in reg.reads]
        for (q, t, lo, w, _fl) in synth.region_pairs(reg, SLACK, granule):
            total += scheduler.band_cells(lens[q], lens[t], lo, w)
    return total


# ---------------------------------------------------------------------------------------------
# strong-scaling sub-records: the north-star workloads (configs 4 and 5) as FIXED pools sharded over the ranks
# ---------------------------------------------------------------------------------------------
SCALING_SETS = {
    "config4": {"sets": [(30_000, False)], "what": "HG002-shape loci (seeds 30000+r): all-vs-all + padded consensus -> reference window, both strands"},
    "config5": {"sets": [(40_000, True), (50_000, True), (60_000, True)],
                "what": "trio: three sets (seeds 40000/50000/60000+r) with a Pareto(1.3) core-length tail, one joint pool"},
}
SCALING_CHUNK = 256          # loci per engine call


def scaling_items(name: str, pool: int) -> list[tuple[int, int, bool]]:
    sets = SCALING_SETS[name]["sets"]
    per = pool // len(sets)
    return [(sb, lid, skew) for sb, skew in sets for lid in range(per)]


def _gen_chunk(items):
    from svirlpool_b200 import workloads
    return workloads.batch_loci(items)


def prepare_scaling(pool: int, rank: int, world: int) -> dict:
    """Host preparation (untimed, before CUDA is touched so that a fork pool is safe): LPT assignment of every pool by estimated
    cells (synth.locus_cost_estimate: shape only, what a scheduler knows before aligning) and the packed batches of THIS rank's
    loci, in chunks of SCALING_CHUNK loci per engine call."""
    import multiprocessing as mp
    from svirlpool_b200 import scheduler, synth
    out = {}
    jobs = []
    for name in SCALING_SETS:
        items = scaling_items(name, pool)
        costs = [synth.locus_cost_estimate(lid, sb, skew) for sb, lid, skew in items]
        mine = scheduler.lpt_assign(costs, world)[rank]
        chunks = [[items[k] for k in mine[c:c + SCALING_CHUNK]] for c in range(0, len(mine), SCALING_CHUNK)]
        out[name] = {"items": items, "costs": costs, "mine": mine, "first": len(jobs), "n_chunks": len(chunks)}
        jobs += chunks
    workers = max(1, min(16, (os.cpu_count() or 1) // max(world, 1)))
    if workers > 1 and len(jobs) > 1:
        with mp.get_context("fork").Pool(workers) as pl:
            batches = pl.map(_gen_chunk, jobs, chunksize=1)
    else:
        batches = [_gen_chunk(j) for j in jobs]
    for name, d in out.items():
        d["batches"] = batches[d["first"]:d["first"] + d["n_chunks"]]
    return out


def run_scaling(al, name: str, prep: dict, rank: int, world: int, dist, torch, dev) -> dict | None:
    """One pass over the fixed pool: every rank aligns its loci through the host-buffer API (H2D, kernels, D2H of rows and
    CIGARs inside the timed region), results are gathered on rank 0 through sharding.run_sharded (host-side gather, no
    collective on the data path). Time = max over ranks of the device stopwatch."""
    from svirlpool_b200 import scheduler, sharding
    d = prep[name]
    pins = []
    for b in d["batches"]:
        pin = torch.empty(b["packed"].nbytes, dtype=torch.uint8, pin_memory=True)
        pin.numpy()[:] = b["packed"]
        pins.append(pin.numpy())
    if d["batches"]:                                   # warm-up: this rank's first chunk, untimed
        b = d["batches"][0]
        al.align_batch_nocopy(pins[0], b["off"], b["lens"], b["pairs"], b["cigar_words"])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    al.stats_total(reset=True)
    state = {}

    def process_many(mine_items):
        rows = []
        t0 = time.perf_counter()
        al.timer_mark(0)
        for b, pin in zip(d["batches"], pins):
            res, _cig = al.align_batch_nocopy(pin, b["off"], b["lens"], b["pairs"], b["cigar_words"])
            pb = b["pair_bounds"]
            cells = np.add.reduceat(res["cells"].astype(np.int64), pb[:-1])
            score = np.add.reduceat(res["score"].astype(np.int64), pb[:-1])
            bad = np.add.reduceat((res["status"] & ~np.uint32(1) != 0).astype(np.int64), pb[:-1])
            rows += [{"pairs": int(pb[k + 1] - pb[k]), "cells": int(cells[k]), "score_sum": int(score[k]), "bad_status": int(bad[k])}
                     for k in range(len(pb) - 1)]
        al.timer_mark(1)
        state["ms"] = al.timer_elapsed_ms() if d["batches"] else 0.0
        state["wall_ms"] = (time.perf_counter() - t0) * 1e3
        return rows

    results = sharding.run_sharded(d["items"], d["costs"], process_many=process_many, dist=dist if world > 1 else None)
    agg = al.stats_total()
    vec = torch.tensor([state["ms"], agg["fwd_ms"], float(agg["cells"]), state["wall_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.zeros_like(vec) for _ in range(world)]
        dist.all_gather(allv, vec)
        allv = torch.stack(allv).cpu().numpy()
    else:
        allv = vec.cpu().numpy()[None, :]
    if rank != 0:
        return None
    ms = allv[:, 0]
    t = float(ms.max()) / 1e3
    cells = sum(r["cells"] for r in results)
    bins = scheduler.lpt_assign(d["costs"], world)
    actual = [r["cells"] for r in results]
    fwd_gcups = float(allv[:, 2].sum()) / (float(allv[:, 1].max()) / 1e3) / 1e9 if allv[:, 1].max() > 0 else 0.0
    ip = int_pipe_peak()
    return {"workload": SCALING_SETS[name]["what"], "scaling": "strong", "loci": len(results), "pairs": sum(r["pairs"] for r in results),
            "bad_status_pairs": sum(r["bad_status"] for r in results),
            "seconds": t, "loci_per_s": len(results) / t, "gcups": cells / t / 1e9,
            "int_pipe_frac": fwd_gcups / (ip["gcups"] * world), "fwd_gcups": fwd_gcups,
            "lpt_imbalance_estimated": scheduler.imbalance(d["costs"], bins), "lpt_imbalance_actual_cells": scheduler.imbalance(actual, bins),
            "rank_ms": [round(float(x), 2) for x in ms], "rank_busy_frac": [round(float(x / ms.max()), 4) for x in ms],
            "rank_wall_ms": [round(float(x), 2) for x in allv[:, 3]],
            "e2e": "host buffers -> svp_align_batch -> host buffers, H2D and D2H (rows + CIGARs) inside the timed region",
            "engine_calls_per_rank": d["n_chunks"], "loci_per_call": SCALING_CHUNK,
            "gather": "sharding.run_sharded (dist.gather_object of per-locus rows to rank 0)"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = args.threads or os.cpu_count() or 1
    total = args.warmup + args.steps
    for _ in range(args.warmup):
        cpu_arm(args.ref_regions, threads)
    secs, cells = 0.0, 0
    last = None
    for _ in range(args.steps):
        last = cpu_arm(args.ref_regions, threads)
        secs += last["_seconds"]; cells += last["_cells"]
    value = cells / secs / 1e9
    base = {k: v for k, v in last.items() if not k.startswith("_")}
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "regions_per_s": args.ref_regions * args.steps / secs,
            "config": {"workload": "config2: 10k regions x 30 ONT-like reads x 8 kb, all-vs-all (435 pairs/region), map-ont scoring",
                       "regions_per_step": args.ref_regions, "band": f"|dlen|+2*{SLACK} rounded up to {GRANULE}", "steps_total": total,
                       "note": "CPU restatement of the alignment spec (oracle/svp_oracle.c), NOT minimap2: the reference's aligner is an external binary absent from this image"},
            "cpu_baseline": base, "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS), help="BASELINE.json config (default: the one the metric is quoted on)")
    ap.add_argument("--regions", type=int, default=None, help="regions per GPU per step (default 80; 8 for config3, 48 for config4)")
    ap.add_argument("--distinct-batches", type=int, default=3, help="distinct synthetic batches cycled over the steps")
    ap.add_argument("--ref-regions", type=int, default=2, help="regions per step of the CPU arm (bounded sample)")
    ap.add_argument("--cpu-sample-regions", type=int, default=2)
    ap.add_argument("--threads", type=int, default=0, help="host threads of the CPU arm (--impl reference); default: all")
    ap.add_argument("--scaling-pool", type=int, default=None,
                    help="loci of the strong-scaling sub-records (config 4: one set; config 5: the three trio sets together); 0 = skip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", dest="pipeline", action="store_false",
                    help="HBM-resident steps as blocking calls instead of pipelined submits (svp_submit_batch_device)")
    args = ap.parse_args()
    if args.regions is None:
        args.regions = {"config2": 80, "config3": 8, "config4": 48}[args.workload]
    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.scaling_pool is None:
        args.scaling_pool = 3072 if args.workload == "config2" else 0
    scaling_prep = prepare_scaling(args.scaling_pool, rank, world) if args.scaling_pool > 0 else None     # fork pool: before CUDA / NCCL exist

    import torch
    import torch.distributed as dist
    from svirlpool_b200._abi import RESULT_DTYPE
    from svirlpool_b200.aligner import Aligner

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version / debug lines on stdout when it creates the communicator; stdout must carry exactly one
        # JSON line, so stdout points at stderr until the communicator exists (first collective)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device=torch.device("cuda", local))
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    dev = torch.device("cuda", local)

    # ---- synthetic input: a few distinct batches, pre-packed, pinned on the host and resident in HBM ----
    n_distinct = max(1, min(args.distinct_batches, args.warmup + args.steps))
    batches = []
    for s in range(n_distinct):
        mine = regions_for_step(s, world, args.regions, args.workload)[rank]
        b = build_batch(mine, args.workload)
        b["d_packed"] = torch.from_numpy(b["packed"]).to(dev)
        pin = torch.empty(b["packed"].nbytes, dtype=torch.uint8, pin_memory=True)
        pin.numpy()[:] = b["packed"]
        b["h_packed_pinned"] = pin.numpy()
        batches.append(b)
    max_pairs = max(len(b["pairs"]) for b in batches)
    max_cig = max(b["cigar_words"] for b in batches)
    d_results = torch.empty(max_pairs * RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_cigar = torch.empty(max_cig, dtype=torch.int32, device=dev)
    # traceback arena of the bench: the free HBM minus 26 GB for the batch buffers (CIGAR slots of the largest strong-scaling chunk).
    # Wave scheduling: the larger the waves, the fewer wave boundaries — a step of 80 regions is 2 whole-arena waves from 148 GB on
    # (the library's own default is a bounded 64 GB); SVP_ARENA_GB overrides
    free_b, _total_b = torch.cuda.mem_get_info(local)
    arena_bytes = 0 if os.environ.get("SVP_ARENA_GB") else max(int(free_b * 0.5), int(free_b) - (26 << 30)) & ~255
    # every call of this handle is planned like a large call (all kernel widths), so that the parity check below — 870 pairs, a small
    # call by itself — runs on the same kernel variants as the timed steps
    os.environ.setdefault("SVP_SMALL_CALL_PAIRS", "0")
    al = None
    for attempt in range(4):                               # an explicit arena size does not shrink by itself: step down if the allocation fails
        try:
            al = Aligner(device=local, preset="map-ont", tb_bytes=arena_bytes)
            break
        except Exception:
            if not arena_bytes or attempt == 3:
                raise
            arena_bytes = int(arena_bytes * 0.85) & ~255

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # result / CIGAR buffers alternate between consecutive steps: the steps are SUBMITTED (svp_submit_batch_device) and
    # overlap on the device — step k+1 is planned and its forward pass starts under step k's last traceback; every step is
    # complete when the timed region ends (svp_sync before the closing mark). --no-pipeline: blocking calls, one at a time.
    d_results2 = torch.empty_like(d_results) if args.pipeline else d_results
    d_cigar2 = torch.empty_like(d_cigar) if args.pipeline else d_cigar

    def step_device(k: int) -> None:
        b = batches[k % n_distinct]
        res, cig = (d_results, d_cigar) if k % 2 == 0 else (d_results2, d_cigar2)
        fn = al.submit_batch_device if args.pipeline else al.align_batch_device
        fn(b["d_packed"].data_ptr(), b["packed"].nbytes, b["off"], b["lens"], b["pairs"], res.data_ptr(), cig.data_ptr(), b["cigar_words"])

    # ---- HBM-resident timing ------------------------------------------------------------------------
    for k in range(args.warmup):
        step_device(k)
    al.sync()
    sync_all()
    al.stats_total(reset=True)
    sampler = ClockSampler(local)
    t0 = time.perf_counter()
    al.timer_mark(0)                                   # device stopwatch: CUDA events on the stream the kernels are launched on
    regions_done, seq_bytes = 0, 0
    for k in range(args.steps):
        step_device(args.warmup + k)
        regions_done += batches[(args.warmup + k) % n_distinct]["n_regions"]
        seq_bytes += batches[(args.warmup + k) % n_distinct]["seq_bytes"]
    al.timer_mark(1)                                   # recorded behind every submitted step's last traceback
    al.sync()
    sync_all()
    agg = {k: v for k, v in al.stats_total().items() if k != "reserved"}
    wall_elapsed = time.perf_counter() - t0
    elapsed = al.timer_elapsed_ms() / 1e3                # device time of the K steps, host gaps between calls included
    clocks = sampler.stop()

    # ---- end to end through the host-buffer API -----------------------------------------------------
    def step_host(k: int):
        b = batches[k % n_distinct]
        return al.align_batch_nocopy(b["h_packed_pinned"], b["off"], b["lens"], b["pairs"], b["cigar_words"])

    step_host(0)
    sync_all()
    t1 = time.perf_counter()
    al.timer_mark(0)
    e2e_cells, h2d, d2h = 0, 0, 0
    for k in range(args.steps):
        b = batches[(args.warmup + k) % n_distinct]
        res, cig = step_host(args.warmup + k)
        e2e_cells += int(res["cells"].sum())
        h2d += b["packed"].nbytes + len(b["pairs"]) * 88 + len(b["pairs"]) * 4
        d2h += res.nbytes + cig.nbytes
    al.timer_mark(1)
    sync_all()
    e2e_wall = time.perf_counter() - t1
    e2e_elapsed = max(al.timer_elapsed_ms() / 1e3, 1e-9)  # first H2D -> last D2H on the device, host planning in between included

    # ---- the same through the CIGAR-free entry point (svp_align_batch_summary): what the similarity matrix needs --------
    b0 = batches[0]
    al.align_batch_summary(b0["h_packed_pinned"], b0["off"], b0["lens"], b0["pairs"], b0["cigar_words"], copy=False)
    sync_all()
    al.timer_mark(0)
    sum_cells, sum_d2h = 0, 0
    for k in range(args.steps):
        b = batches[(args.warmup + k) % n_distinct]
        res = al.align_batch_summary(b["h_packed_pinned"], b["off"], b["lens"], b["pairs"], b["cigar_words"], copy=False)
        sum_cells += int(res["cells"].sum())
        sum_d2h += res.nbytes
    al.timer_mark(1)
    sync_all()
    sum_elapsed = max(al.timer_elapsed_ms() / 1e3, 1e-9)

    # ---- strong scaling of the north-star workloads on fixed pools (configs 4 and 5) ------------------------------------
    scaling = {}
    if scaling_prep is not None:
        for name in SCALING_SETS:
            rec = run_scaling(al, name, scaling_prep, rank, world, dist, torch, dev)
            if rec is not None:
                scaling[name] = rec

    # ---- reduce over ranks: max time, summed work ---------------------------------------------------
    vec = torch.tensor([elapsed, e2e_elapsed, float(agg["cells"]), float(regions_done), float(e2e_cells), agg["fwd_ms"], agg["tb_ms"],
                        float(agg["fwd_launches"] + agg["tb_launches"]), float(agg["tb_bytes"] + seq_bytes), float(agg["fwd_launches"]),
                        sum_elapsed, float(sum_cells)],
                       dtype=torch.float64, device=dev)
    if world > 1:
        mx = vec.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vec.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed, e2e_elapsed, sum_elapsed = mx[0].item(), mx[1].item(), mx[10].item()
        sum_cells = sm[11].item()
        cells, regions_all, e2e_cells_all, launches = sm[2].item(), sm[3].item(), sm[4].item(), sm[7].item()
    else:
        cells, regions_all, e2e_cells_all, launches = vec[2].item(), vec[3].item(), vec[4].item(), vec[7].item()

    if rank == 0:
        pk, ip = peaks(), int_pipe_peak()
        fwd_s = agg["fwd_ms"] / 1e3
        algo_bytes = agg["tb_bytes"] + seq_bytes            # 1 B per spec-band cell row slot + packed reads
        achieved_gbs = algo_bytes / fwd_s / 1e9 if fwd_s > 0 else 0.0
        fwd_gcups = agg["cells"] / fwd_s / 1e9 if fwd_s > 0 else 0.0
        tr = traffic_ratio()
        ap_ = alu_pipe_peak(clocks.get("sm_mhz"))
        n_fwd = max(int(agg["fwd_launches"]), 1)
        line = {
            "metric": METRIC if args.workload == "config2" else f"gcups_{args.workload}", "value": cells / elapsed / 1e9, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s16x2",          # packed 16-bit lanes; reads beyond the s16 score range on per-warp score offsets
            "data": "synthetic", "regions_per_s": regions_all / elapsed,
            "config": {"workload": WORKLOADS[args.workload],
                       "regions_per_step_per_gpu": args.regions, "distinct_batches": n_distinct,
                       "region_seeds": {"config2": "10000+r", "config3": "20000+r", "config4": "30000+r"}[args.workload],
                       "band": f"|dlen|+2*{SLACK} rounded up to {GRANULE} diagonals" if args.workload != "config3" else "|dlen|+512", "sharding": "LPT by estimated cells, no collective",
                       "l2": "inputs+traceback rows per step >> 126 MB L2 (every pair writes 6-12 MB of rows)",
                       "arena_gb": round((arena_bytes or 0) / 2**30, 1) or os.environ.get("SVP_ARENA_GB"),
                       "timing": "CUDA events on the library's launch stream (svp_timer_mark), max over ranks; wall clock reported beside it",
                       "pipeline": ("steps submitted back to back (svp_submit_batch_device), step k+1's kernels start as step k's CTAs retire; "
                                    "all K steps complete inside the timed region") if args.pipeline else "blocking calls"},
            "wall_ms_per_step": wall_elapsed / args.steps * 1e3,
            "e2e": {"value": e2e_cells_all / e2e_elapsed / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps, "regions_per_s": regions_all / e2e_elapsed,
                    "wall_ms_per_step": e2e_wall / args.steps * 1e3},
            "e2e_summary": {"value": sum_cells / sum_elapsed / 1e9, "unit": "GCUPS", "d2h_bytes_per_step": sum_d2h // args.steps,
                            "note": "svp_align_batch_summary: result rows only (score, intervals, counts, sv_dist), no CIGAR D2H"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved_gbs / pk["hbm_gbs"],
                         "traffic": (tr["ratio"] * algo_bytes / n_fwd) if tr else None,
                         "traffic_note": (f"DRAM bytes per forward launch = algorithmic bytes per launch x {tr['ratio']:.4f} "
                                          f"(dram__bytes_read+write / algorithmic, ncu --set full, {tr['source']})") if tr else None,
                         "algorithmic_bytes_per_launch": algo_bytes / n_fwd,
                         "kernel": "svp_fwd_kernel (one launch per kernel variant and wave; the launches of a wave run concurrently)", "peak_source": pk["source"],
                         "launches": int(agg["fwd_launches"]), "avg_launch_ms": agg["fwd_ms"] / n_fwd,
                         "algorithmic_bytes_per_cell": 1.0,
                         "int_pipe": {"achieved_gcups": fwd_gcups, "peak_gcups": ip["gcups"], "frac": fwd_gcups / ip["gcups"],
                                      "peak_source": ip["source"],
                                      "alu_pipe_14op": {"peak_gcups": ap_["gcups"], "frac": fwd_gcups / ap_["gcups"], **{k: ap_[k] for k in ("ipc_per_smsp", "sm_mhz", "int_ops_per_cell", "source")}},
                                      "note": "the kernel is integer-ALU-bound, not HBM-bound (DESIGN.md §5); frac uses the stricter of the two denominators"}},
            "kernel_ms": {"device_total": agg["total_ms"], "fwd": agg["fwd_ms"], "walk": agg["tb_ms"], "waves_or_groups": agg["waves"],
                          "scheduler": os.environ.get("SVP_MODE", "waves"),
                          "note": "waves: fwd = union of the waves' forward intervals, walk = summed spans of the walk kernels (svp_tbg / svp_tbw / svp_tb); "
                                  "arena / queue: fwd = the call's kernel time, walk = its share by the CTAs' own clocks"},
        }
        if scaling:
            line["strong_scaling"] = scaling
        if args.workload == "config2":
            # the same step's cells with bands rounded up to 32 instead of GRANULE diagonals: the un-padded spec figure
            ids0 = regions_for_step(0, world, args.regions, args.workload)[0]
            c_g, c_32 = cells_at_granule(ids0[:8], GRANULE), cells_at_granule(ids0[:8], 32)
            line["cells_granule32"] = {"ratio_to_counted_cells": c_32 / c_g, "gcups": line["value"] * c_32 / c_g,
                                       "note": f"bands |dlen|+2*{SLACK} rounded up to 32 instead of {GRANULE} diagonals (8 regions of step 0); regions_per_s is granule-free"}
        if not args.no_cpu_baseline and world == 1 and args.workload == "config2":
            ncpu = os.cpu_count() or 1
            cb = cpu_arm(args.cpu_sample_regions, ncpu, keep_rows=True)
            rows = cb.pop("_rows")
            line["cpu_baseline"] = {k: v for k, v in cb.items() if not k.startswith("_")}
            cb1 = cpu_arm(1, 1)                         # --threads 1: the reference's per-container setting (consensus.py:3270)
            line["cpu_baseline"]["threads_1"] = {k: v for k, v in cb1.items() if not k.startswith("_")}
            line["parity_checked"] = parity_check(al, rows)
        print(json.dumps(line), flush=True)
        if line.get("parity_checked", {}).get("mismatches", 0):
            print("bench.py: GPU rows differ from the oracle rows of the CPU arm", file=sys.stderr, flush=True)
            al.close()
            os._exit(1)
    al.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
