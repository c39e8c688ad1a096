"""Synthetic BASELINE workloads (configs 2-5): generators are deterministic per region id, pair specs are valid, the
consensus_align chain recovers the planted SV of a config-4 locus (CPU oracle engine), and the LPT scheduler balances
the skewed config-5 cost distribution."""
import numpy as np

from svirlpool_b200 import consensus_align, scheduler, synth, workloads


def test_generators_are_deterministic_and_shaped():
    a, b = synth.region_config3(4, n_reads=6), synth.region_config3(4, n_reads=6)
    assert all(np.array_equal(x, y) for x, y in zip(a.reads, b.reads)) and a.strands == b.strands
    h0, h1 = a.meta["hap_lengths"]
    assert 10_000 <= h0 <= 30_000 and 2_000 <= h1 - h0 <= 20_000 and 20 <= a.meta["motif_len"] <= 60
    assert all(abs(len(r) - (h1 if hp else h0)) < 0.06 * h1 for r, hp in zip(a.reads, a.haplotypes))
    la, lb = synth.locus_config4(7), synth.locus_config4(7)
    assert np.array_equal(la.ref_window, lb.ref_window) and [c["id"] for c in la.consensus] == [c["id"] for c in lb.consensus]
    assert 4 <= len(la.region.reads) <= 60 and 600 <= la.region.meta["core"] <= 50_000
    for c in la.consensus:
        lo, hi = c["core"]
        assert lo == min(2 * la.region.meta["core"], 20_000) or c["strand"] == "-"
        assert 0 < lo < hi < len(c["padded"])
    cores = [synth.locus_config4(k, seed_base=40_000, skew=True).region.meta["core"] for k in range(300)]
    assert max(cores) > 20 * np.median(cores) and max(cores) <= 100_000                     # Pareto tail


def test_batches_are_valid():
    for b in (workloads.batch_config2([1]), workloads.batch_config3([0], n_reads=5, scale=0.3),
              workloads.batch_config4([0, 2], scale=0.5, max_reads=6)):
        p = b["pairs"]
        assert len(p) and (p["band_w"] % 32 == 0).all() and (p["q"] != p["t"]).all()
        assert p["cigar_off"][-1] + p["cigar_cap"][-1] == b["cigar_words"]
        m, n = b["lens"][p["q"]].astype(np.int64), b["lens"][p["t"]].astype(np.int64)
        assert (p["band_lo"] <= n - 1).all() and (p["band_lo"].astype(np.int64) + p["band_w"] - 1 >= -(m - 1)).all()
    b3 = workloads.batch_config3([6], n_reads=20)
    assert b3["pairs"]["band_w"].max() > 16_384 and len(b3["pairs"]) == 190               # 17.9 kb length difference: cluster territory


def test_config4_chain_recovers_planted_sv(oracle):
    loc = next(synth.locus_config4(k, scale=0.25, max_reads=6) for k in range(50)
               if synth.locus_config4(k, scale=0.25, max_reads=6).region.meta["two_haps"]
               and abs(synth.locus_config4(k, scale=0.25, max_reads=6).region.sv_size) >= 60)
    cons = workloads.consensus_objects(loc)
    ref = synth.codes_to_str(loc.ref_window)
    fetch = lambda c, s, e: ref[s - loc.window_start:e - loc.window_start]
    recs = consensus_align.align_consensuses_to_reference(oracle, cons, fetch, {loc.chrom: loc.window_start + len(ref)}, band_w=1024,
                                                          min_margin=0)
    assert {r.query_name for r in recs} == set(cons)
    ann = consensus_align.annotate_alignments(recs, "S")
    cores = consensus_align.core_intervals_on_reference(cons, ann)
    sigs = consensus_align.consensus_sv_signals("S", cons, ann)
    for cid, c in cons.items():
        (idx, chrom, a, b, lo, hi), = cores[cid]
        assert chrom == loc.chrom and abs(a - loc.core_interval[0]) <= 5 and abs(b - loc.core_interval[1]) <= 5
        flat = [s for per in sigs[cid] for s in per]
        if cid.endswith(".1"):
            assert len(flat) == 1 and flat[0].size == abs(loc.region.sv_size) and flat[0].sv_type == (0 if loc.region.sv_size > 0 else 1)
        else:
            assert flat == []


def test_lpt_balances_skewed_trio():
    costs = []
    for base in (40_000, 50_000, 60_000):
        for k in range(120):
            core = synth._core_length(np.random.Generator(np.random.PCG64(base + k)), True)
            costs.append(30 * 29 // 2 * core * 1024)
    bins = scheduler.lpt_assign(costs, 8)
    naive = [list(range(k, len(costs), 8)) for k in range(8)]
    assert scheduler.imbalance(costs, bins) <= scheduler.imbalance(costs, naive)
    assert scheduler.imbalance(costs, bins) < 1.25 or max(costs) > sum(costs) / 8            # one giant region can dominate
