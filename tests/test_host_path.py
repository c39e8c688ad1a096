"""CIGAR consumers, batching, SAM/BAM seam, scheduler, synthetic data — everything on the host side
of the path that needs no GPU. Expected values restate the reference's own tests where they exist."""
import numpy as np
import pytest

from svirlpool_b200 import alignments_to_rafs, ava, consensus, consensus_lib, crs_to_batches, sam, scheduler, synth, util
from svirlpool_b200.datatypes import AlignedRecord, SVsignal, cigar_from_string


def rec(cigar, ref_start, flag=0, name="read-0", ref="ref", seq=None):
    return AlignedRecord(query_name=name, reference_name=ref, reference_start=ref_start, cigartuples=cigar_from_string(cigar),
                         flag=flag, query_sequence=seq)


# ---- dummy inversion fixture: tests/test_alignments_to_rafs.py:46-155 ---------------------------
INVERSION = [("2000M4000H", 0, 2048), ("2000H2000M2000H", 2000, 2064), ("4000S2000M", 4000, 0)]


def test_get_start_end_inversion_records():
    got = [alignments_to_rafs.get_start_end(rec(c, p, f)) for c, p, f in INVERSION]
    assert got == [(0, 2000, 0, 2000), (2000, 4000, 2000, 4000), (4000, 6000, 4000, 6000)]


def test_read_alignment_signals_inversion_records():
    got = [consensus.parse_ReadAlignmentSignals_from_alignment("test", rec(c, p, f), 10, 1000) for c, p, f in INVERSION]
    assert [g.alignment_forward for g in got] == [True, False, True]
    assert got[0].SV_signals == [SVsignal(2000, 2000, 2000, 2000, 4000, 4)]
    assert got[1].SV_signals == [SVsignal(2000, 2000, 4000, 4000, 2000, 3), SVsignal(4000, 4000, 2000, 2000, 2000, 4)]
    assert got[2].SV_signals == [SVsignal(4000, 4000, 4000, 4000, 4000, 3)]
    assert got[0].read_name == "read-0" and got[0].reference_name == "ref" and got[0].samplename == "test"


def test_inversion_records_from_golden_through_oracle(golden, oracle):
    """End to end on the CPU: oracle alignment of the fixture read -> same signals as the frozen records."""
    case = next(c for c in golden["cases"] if c["name"] == "dummy_inversion")
    recs = ava.map_queries(oracle, case["reads"], "ref", case["reference"])
    sig = sorted((s.ref_start, s.read_start, s.size, s.sv_type) for r in recs
                 for s in consensus.parse_ReadAlignmentSignals_from_alignment("t", r, 10, 1000).SV_signals)
    assert sig == [(2000, 2000, 4000, 4), (2000, 4000, 2000, 3), (4000, 2000, 2000, 4), (4000, 4000, 4000, 3)]


# ---- get_starts_ends / indels --------------------------------------------------------------------
def test_starts_ends_forward_and_reverse():
    t, x = np.array([4, 0, 1, 0, 2, 0, 5]), np.array([10, 50, 20, 30, 15, 40, 7])
    rs, re_, fs, fe = util.get_starts_ends(t, x, reference_start=100, is_reverse=False)
    assert rs.tolist() == [10, 10, 60, 80, 110, 110, 150] and re_.tolist() == [10, 60, 80, 110, 110, 150, 150]
    assert fs.tolist() == [100, 100, 150, 150, 180, 195, 235] and fe.tolist() == [100, 150, 150, 180, 195, 235, 235]
    rrs, rre, ffs, ffe = util.get_starts_ends(t, x, reference_start=100, is_reverse=True)
    total = 10 + 50 + 20 + 30 + 40 + 7
    assert rrs.tolist() == (total - re_).tolist() and rre.tolist() == (total - rs).tolist()
    assert ffs.tolist() == fs.tolist() and ffe.tolist() == fe.tolist()


def test_signals_order_thresholds_and_strand():
    r = rec("150S100M20I100M5D100M30D50M12I10M120H", 1000)
    sig = alignments_to_rafs.parse_SVsignals_from_alignment(r, r.reference_start, r.reference_end, r.query_alignment_start,
                                                             r.query_alignment_end, min_signal_size=12, min_bnd_size=100)
    assert [s.sv_type for s in sig] == [3, 4, 1, 0, 0]          # BNDL, BNDR, DELs, INSs
    assert sig[0] == SVsignal(1000, 1000, 150, 150, 150, 3)
    assert sig[1] == SVsignal(1395, 1395, 542, 542, 120, 4)
    assert sig[2] == SVsignal(1305, 1335, 470, 470, 30, 1)
    assert sig[3] == SVsignal(1100, 1100, 250, 270, 20, 0) and sig[4] == SVsignal(1385, 1385, 520, 532, 12, 0)
    rev = rec("150S100M20I100M", 1000, flag=16)
    s2 = alignments_to_rafs.parse_SVsignals_from_alignment(rev, 1000, 1200, rev.query_alignment_start, rev.query_alignment_end, 12, 100)
    assert s2[0].sv_type == 3 and s2[0].read_start == rev.query_alignment_end       # swapped on reverse
    assert s2[1] == SVsignal(1100, 1100, 100, 120, 20, 0)                            # mirrored into read orientation


def test_directed_signals_use_aligned_span_as_ref_length():
    recs = [rec("100S900M", 50, name="q", ref="t"), AlignedRecord("u", None, 0, None, flag=4)]
    ds = consensus_lib.parse_directed_signals_from_ava(recs, 12, 100)
    assert len(ds) == 1 and ds[0].ref_length == 900 and ds[0].query_name == "q" and ds[0].ref_name == "t"
    pooled = consensus_lib.parse_sv_signals_from_ava_alignments(recs, 12, 100)
    assert list(pooled) == ["t"] and pooled["t"][0].sv_type == 3


# ---- crs_to_batches: tests/test_crs_to_batches.py ---------------------------------------------------
def test_build_batches_rules(tmp_path):
    b = crs_to_batches.build_batches({1: (["chr1"], 100, 200), 2: (["chr1"], 300, 400), 3: (["chr1"], 500, 600), 4: (["chr2"], 100, 200)}, 100, 10_000)
    assert [x["crIDs"] for x in b] == [[1, 2, 3], [4]] and [x["chr"] for x in b] == ["chr1", "chr2"]
    five = {k: (["chr1"], k * 100, k * 100 + 50) for k in range(1, 6)}
    assert sorted(len(x["crIDs"]) for x in crs_to_batches.build_batches(five, 2, 10_000)) == [1, 2, 2]
    span = crs_to_batches.build_batches({1: (["chr1"], 100, 200), 2: (["chr1"], 5000, 5100), 3: (["chr1"], 9000, 9100)}, 100, 5000)
    assert sorted(tuple(x["crIDs"]) for x in span) == [(1, 2), (3,)]
    wide = crs_to_batches.build_batches({1: (["chr1"], 100, 200), 2: (["chr1", "chr2"], 100, 200), 3: (["chr3"], 0, 100_000_000),
                                         4: (["chr1"], 300, 400)}, 100, 1_000_000)
    assert [x["crIDs"] for x in wide] == [[2], [3], [1, 4]] and wide[0]["chr"] == "chr1,chr2"
    assert [x["batch_id"] for x in wide] == [0, 1, 2]
    with pytest.raises(ValueError):
        crs_to_batches.build_batches({}, 0, 10)
    tsv, ids = tmp_path / "b.tsv", tmp_path / "ids.txt"
    crs_to_batches.write_outputs(wide, tsv, ids)
    assert tsv.read_text().splitlines()[0] == "batch_id\tchr\tstart\tend\tn_containers\tcrIDs"
    for x in wide:
        assert crs_to_batches._load_crIDs_from_batch_tsv(tsv, x["batch_id"]) == x["crIDs"]
    assert ids.read_text().split() == ["0", "1", "2"]
    with pytest.raises(ValueError):
        crs_to_batches._load_crIDs_from_batch_tsv(tsv, 99)
    with pytest.raises(FileNotFoundError):
        crs_to_batches._load_crIDs_from_batch_tsv(tmp_path / "missing.tsv", 0)


def test_container_db_roundtrip(tmp_path):
    import json
    import sqlite3
    db = tmp_path / "containers.db"
    conn = sqlite3.connect(db)
    conn.execute("CREATE TABLE containers (crID INTEGER PRIMARY KEY, data TEXT)")
    conn.execute("INSERT INTO containers VALUES (5, ?)", (json.dumps({"crs": [{"chr": "1", "referenceStart": 10, "referenceEnd": 90},
                                                                            {"chr": "1", "referenceStart": 50, "referenceEnd": 400}]}),))
    conn.execute("INSERT INTO containers VALUES (6, ?)", (json.dumps({"crs": []}),))
    conn.commit(); conn.close()
    out = crs_to_batches.crs_containers_to_batches(db, tmp_path / "t.tsv", tmp_path / "i.txt")
    assert out == [{"batch_id": 0, "chr": "1", "start": 10, "end": 400, "crIDs": [5]}]


# ---- SAM / BAM seam ---------------------------------------------------------------------------------
def test_sam_and_bam_roundtrip(tmp_path):
    recs = [rec("10S90M5D100M", 7, name="a", ref="t1", seq="ACGT" * 50), rec("200M", 0, flag=16, name="b", ref="t0", seq="T" * 200)]
    recs[0].tags = {"NM": 5, "AS": 321}
    refs = {"t0": 500, "t1": 900}
    ordered = sam.sort_records(recs, refs)
    assert [r.query_name for r in ordered] == ["b", "a"]
    for writer, name in ((sam.write_sam, "x.sam"), (sam.write_bam, "x.bam")):
        path = tmp_path / name
        writer(path, ordered, refs)
        back = list(sam.read_alignments(path))
        assert [(r.query_name, r.reference_name, r.reference_start, r.cigarstring, r.flag, r.query_sequence) for r in back] == \
               [(r.query_name, r.reference_name, r.reference_start, r.cigarstring, r.flag, r.query_sequence) for r in ordered]
    parsed = consensus_lib.parse_directed_signals_from_ava(tmp_path / "x.bam", 5, 10)
    assert [(d.query_name, d.ref_name, d.ref_length) for d in parsed] == [("b", "t0", 200), ("a", "t1", 195)]
    assert parsed[1].signals[0].sv_type == 3 and parsed[1].signals[1] == SVsignal(97, 102, 100, 100, 5, 1)


# ---- scheduler / synth ----------------------------------------------------------------------------------
def test_band_cells_closed_form_matches_enumeration():
    rng = np.random.default_rng(0)
    for _ in range(200):
        m, n = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        lo, w = int(rng.integers(-45, 45)), int(rng.integers(1, 4)) * 32
        brute = sum(1 for i in range(1, m + 1) for j in range(1, n + 1) if lo <= j - i < lo + w)
        assert scheduler.band_cells(m, n, lo, w) == brute


def test_lpt_assignment_is_balanced_and_deterministic():
    rng = np.random.default_rng(3)
    costs = (rng.pareto(1.3, 400) * 1000 + 10).astype(np.int64)
    bins = scheduler.lpt_assign(costs, 8)
    assert sorted(k for b in bins for k in b) == list(range(400))
    assert bins == scheduler.lpt_assign(costs, 8)
    naive = [list(range(k, 400, 8)) for k in range(8)]
    assert scheduler.imbalance(costs, bins) <= scheduler.imbalance(costs, naive)
    assert scheduler.imbalance(costs, bins) < 1.0 + max(costs) / (costs.sum() / 8)


def test_synthetic_regions_are_reproducible_and_oriented():
    a, b = synth.region_config2(7, n_reads=6, mean_len=1500, sd_len=100), synth.region_config2(7, n_reads=6, mean_len=1500, sd_len=100)
    assert all(np.array_equal(x, y) for x, y in zip(a.reads, b.reads)) and a.strands == b.strands and a.sv_size == b.sv_size
    assert (a.meta["length"], a.sv_size) == synth.region_config2_shape(7, mean_len=1500, sd_len=100)
    specs = synth.region_pairs(a, granule=512)
    assert len(specs) == 15 and all(w % 512 == 0 for *_, w, _ in specs)
    order = sorted(range(6), key=lambda k: a.names[k])
    assert all(a.names[q] < a.names[t] for q, t, *_ in specs) and specs[0][:2] == (order[0], order[1])
    assert all((fl & 1) == (a.strands[q] != a.strands[t]) for q, t, _, _, fl in specs)
    err = synth.noisy_copy(np.random.default_rng(1), np.random.default_rng(2).integers(0, 4, 20000).astype(np.uint8))
    assert 0.97 < len(err) / 20000 < 1.03


def test_ava_on_synthetic_region_with_oracle(oracle):
    """Two haplotypes differing by one SV separate in the similarity matrix (CPU path: oracle engine)."""
    reg = synth.region_config2(3, n_reads=8, mean_len=1800, sd_len=50)
    reads = reg.read_strings()
    sim, names, records = consensus.ava_similarity(oracle, reads, strands=dict(zip(reg.names, reg.strands)))
    assert len(records) == 28 and names == reg.names
    hap = np.array(reg.haplotypes)
    same = sim[np.equal.outer(hap, hap) & ~np.eye(8, dtype=bool)]
    diff = sim[~np.equal.outer(hap, hap)]
    if abs(reg.sv_size) >= 50 and len(diff) and len(same):
        assert same.mean() > diff.mean()
    lens = consensus.pairwise_alignment_lengths(records, names)
    assert lens.shape == (8, 8) and (lens == lens.T).all() and lens.max() > 1500


def test_file_seam_align_reads_with_minimap(tmp_path, golden, oracle):
    """S1 seam: FASTA in -> BAM out, consumed by the path-taking consensus_lib functions (oracle engine on CPU)."""
    case = next(c for c in golden["cases"] if c["name"] == "simulated.with_deletion_reverse")
    ref_fa, reads_fa, bam = tmp_path / "ref.fa", tmp_path / "reads.fa", tmp_path / "out.bam"
    ref_fa.write_text(">ref\n" + case["reference"] + "\n")
    reads_fa.write_text("".join(f">{n} desc\n{s}\n" for n, s in case["reads"].items()))
    util.align_reads_with_minimap(reference=ref_fa, reads=reads_fa, bamout=bam, tech="map-ont", threads=1, engine=oracle)
    back = list(sam.read_alignments(bam))
    assert [(r.reference_start, r.is_reverse, r.cigarstring) for r in back] == [(0, True, "480M20D500M")]
    assert back[0].query_sequence == util.reverse_complement(next(iter(case["reads"].values())))
    ds = consensus_lib.parse_directed_signals_from_ava(bam, 12, 100)
    assert ds[0].signals == [SVsignal(480, 500, 500, 500, 20, 1)]


def test_final_consensus_record(oracle):
    """final_consensus (consensus.py:821-916): reads of one haplotype aligned to their consensus; the indel carried by a
    read of the OTHER haplotype shows up as a distortion of the consensus."""
    from svirlpool_b200 import consensus, synth
    from svirlpool_b200.util import reverse_complement
    rid = next(r for r in range(200) if 50 <= abs(synth.region_config2_shape(r, 1800, 60)[1]) <= 120)   # an SV the DP bridges
    region = synth.region_config2(rid, n_reads=8, mean_len=1800, sd_len=60)
    reads = region.read_strings()
    hap0 = [n for n, h in zip(region.names, region.haplotypes) if h == 0]
    other = [n for n, h in zip(region.names, region.haplotypes) if h == 1][:1]
    strand = dict(zip(region.names, region.strands))
    cseq = reads[hap0[0]] if strand[hap0[0]] == "+" else reverse_complement(reads[hap0[0]])
    use = {n: reads[n] for n in hap0 + other}
    c = consensus.final_consensus(oracle, "S", 8, 100, use, cseq, "11.0", [11], [("chr1", 1000, 2800)])
    assert c is not None and c.get_used_readnames() == set(use)
    assert {f for _s, _e, n, f in c.intervals_cutread_alignments if n == hap0[0]} == {strand[hap0[0]] == "+"}
    big = [d for d in c.get_consensus_distortions() if d.readname == other[0] and d.size >= 40]
    assert len(big) == 1 and abs(big[0].size - abs(region.sv_size)) <= 10 and big[0].type == (0 if region.sv_size > 0 else 1)
    assert c.get_max_cutread_coverage() >= len(hap0)
    assert consensus.final_consensus(oracle, "S", 8, 100, {"x": "ACGT" * 30}, cseq, "11.0", [11], []) is None


def test_find_representative_read_rank_sum_medoid():
    """consensus.find_representative_read (reference: consensus.py:583-660): ranks by mean similarity inside the cluster and
    by summed alignment lengths to the other cluster reads, smallest rank sum wins; reads outside the cluster do not count."""
    names = ["a", "b", "c", "d", "x"]
    sim = np.array([[1.0, 0.9, 0.8, 0.2, 0.0],
                    [0.9, 1.0, 0.9, 0.3, 0.0],
                    [0.8, 0.9, 1.0, 0.1, 0.0],
                    [0.2, 0.3, 0.1, 1.0, 0.9],
                    [0.0, 0.0, 0.0, 0.9, 1.0]])
    lens = np.array([[0, 7000, 6500, 100, 0],
                     [7000, 0, 7800, 200, 0],
                     [6500, 7800, 0, 150, 0],
                     [100, 200, 150, 0, 9000],
                     [0, 0, 0, 9000, 0]])
    before = lens.copy()
    assert consensus.find_representative_read(sim, lens, names, ["a", "b", "c"]) == "b"
    assert consensus.find_representative_read(sim, lens, names, ["c", "a", "b"]) == "b"        # cluster order does not matter here
    assert consensus.find_representative_read(sim, lens, names, ["d", "x"]) in ("d", "x")
    assert consensus.find_representative_read(sim, lens, names, ["x"]) == "x"
    assert np.array_equal(lens, before)                                                         # inputs are not modified
    # similarity and alignment length disagree (c: ranks 0 + 1, b: 2 + 0, a: 1 + 2): the smallest rank sum wins
    sim2 = sim.copy(); sim2[0, 1] = sim2[1, 0] = 0.5; sim2[0, 2] = sim2[2, 0] = 0.95
    assert consensus.find_representative_read(sim2, lens, names, ["a", "b", "c"]) == "c"


def test_ava_align_many_equals_per_container_calls(oracle):
    """ava.ava_align_many: several containers through ONE engine call (one sequence table, svp_ava_pairs per container) return,
    per container, the records of ava.ava_align — incl. a container with a single read and one with unknown strands."""
    rng = np.random.default_rng(31)
    containers, strands = [], []
    for c, n_reads in enumerate((4, 1, 3, 5)):
        tmpl = "".join("ACGT"[k] for k in rng.integers(0, 4, 900 + 100 * c))
        reads, st = {}, {}
        for r in range(n_reads):
            seq = list(tmpl[int(rng.integers(0, 40)):len(tmpl) - int(rng.integers(0, 40))])
            for pos in rng.choice(len(seq), 30, replace=False):
                seq[pos] = "ACGT"[(("ACGT".index(seq[pos])) + 1 + int(rng.integers(3))) % 4]
            s = "".join(seq)
            rev = bool(rng.integers(2))
            name = f"c{c}_read{int(rng.integers(1000)):03d}_{r}"
            reads[name] = s.translate(str.maketrans("ACGT", "TGCA"))[::-1] if rev else s
            st[name] = "-" if rev else "+"
        containers.append(reads)
        strands.append(None if c == 2 else st)
    many = ava.ava_align_many(oracle, containers, strands)
    assert len(many) == 4 and many[1] == []
    for reads, st, got in zip(containers, strands, many):
        one = ava.ava_align(oracle, reads, strands=st)
        key = lambda r: (r.query_name, r.reference_name, r.reference_start, r.flag, r.cigartuples, r.tags)
        assert [key(r) for r in got] == [key(r) for r in one]
    assert sum(len(x) for x in many) >= 6 + 3 + 10


def test_add_unaligned_reads_to_consensuses_inplace(oracle):
    """Unused reads are re-homed to the consensus they align best to (reference: consensus.py:2050-2131): reads of two
    haplotypes that differ by a 60-base deletion go to their own consensus and carry no distortion there; an unrelated read stays
    unassigned; signals and intervals are appended to what the consensus already holds."""
    from svirlpool_b200 import consensus_class
    rng = np.random.default_rng(41)
    hap0 = "".join("ACGT"[k] for k in rng.integers(0, 4, 1500))
    hap1 = hap0[:700] + hap0[760:]

    def noisy(seq, n_sub=12):
        s = list(seq)
        for pos in rng.choice(len(s), n_sub, replace=False):
            s[pos] = "ACGT"[(("ACGT".index(s[pos])) + 1 + int(rng.integers(3))) % 4]
        return "".join(s)

    cons = {cid: consensus_class.Consensus(ID=cid, crIDs=[3], original_regions=[("chr1", 1000, 2500)], consensus_sequence=seq)
            for cid, seq in (("3.0", hap0), ("3.1", hap1))}
    cons["3.0"].intervals_cutread_alignments.append((0, 1500, "old_read", True))
    pool = {"r_h0_a": noisy(hap0[20:1480]), "r_h0_b": util.reverse_complement(noisy(hap0[:1450])), "r_h1_a": noisy(hap1[10:]),
            "junk": "".join("ACGT"[k] for k in rng.integers(0, 4, 800))}
    out = consensus.add_unaligned_reads_to_consensuses_inplace(oracle, "S", cons, pool, min_indel_size=8, min_bnd_size=100)
    assert out == {"r_h0_a": "3.0", "r_h0_b": "3.0", "r_h1_a": "3.1"}
    assert cons["3.0"].get_used_readnames() == {"old_read", "r_h0_a", "r_h0_b"} and cons["3.1"].get_used_readnames() == {"r_h1_a"}
    assert {i[2]: i[3] for i in cons["3.0"].intervals_cutread_alignments}["r_h0_b"] is False        # reverse-strand read
    assert len(cons["3.0"].cut_read_alignment_signals) == 2 and len(cons["3.1"].cut_read_alignment_signals) == 1
    assert cons["3.0"].get_consensus_distortions() == [] and cons["3.1"].get_consensus_distortions() == []
    # against the wrong haplotype only, the same read shows the 60-base event
    only0 = {"3.0": consensus_class.Consensus(ID="3.0", crIDs=[3], original_regions=[("chr1", 1000, 2500)], consensus_sequence=hap0)}
    consensus.add_unaligned_reads_to_consensuses_inplace(oracle, "S", only0, {"r_h1_a": pool["r_h1_a"]}, 8, 100)
    d = only0["3.0"].get_consensus_distortions()
    assert len(d) == 1 and d[0].size == 60 and d[0].type == 1 and abs(d[0].position - 700) <= 5
    with pytest.raises(AssertionError):
        consensus.add_unaligned_reads_to_consensuses_inplace(oracle, "S", {}, pool, 8, 100)
    assert consensus.add_unaligned_reads_to_consensuses_inplace(oracle, "S", cons, {}, 8, 100) == {}


def test_similarity_of_many_containers_in_one_call(oracle):
    """consensus.ava_similarity_summary_many == consensus.ava_similarity_summary per container (matrices bit-identical), and ==
    the record pipeline ava_similarity(densities_weight=0)."""
    regions = [synth.region_config2(r, n_reads=n, mean_len=1200, sd_len=60) for r, n in ((0, 6), (1, 1), (2, 5))]
    containers = [reg.read_strings() for reg in regions]
    strands = [dict(zip(reg.names, reg.strands)) for reg in regions]
    many = consensus.ava_similarity_summary_many(oracle, containers, strands, granule=64)
    assert len(many) == 3 and many[1][0].shape == (1, 1)
    for reads, st, (sim, names, aln_len) in zip(containers, strands, many):
        sim1, names1, len1 = consensus.ava_similarity_summary(oracle, reads, strands=st, granule=64)
        assert names == names1 and np.array_equal(sim, sim1) and np.array_equal(aln_len, len1)
        if len(reads) > 1:
            sim2, names2, _recs = consensus.ava_similarity(oracle, reads, strands=st, densities_weight=0.0, granule=64)
            assert names2 == names and np.allclose(sim, sim2, rtol=1e-12, atol=0)


def test_file_seam_multi_reference_and_soft_supplementary(tmp_path, oracle):
    """S1 seam against a FASTA of several sequences (reads vs all consensuses, consensus.py:2089-2096): `--secondary=no` keeps a
    read's records on the sequence of its best primary alignment only; `-Y` makes supplementary records soft-clipped with the
    full read sequence (consensus.py:851-857), without it they are hard-clipped."""
    rng = np.random.default_rng(51)
    a = "".join("ACGT"[k] for k in rng.integers(0, 4, 1600))
    b = "".join("ACGT"[k] for k in rng.integers(0, 4, 1500))
    inv = a[:500] + util.reverse_complement(a[500:1100]) + a[1100:]            # read with an inverted middle: 3 records on a
    reads = {"on_a": a[100:1500], "on_b_rc": util.reverse_complement(b[50:1400]), "inv_a": inv}
    ref_fa, reads_fa = tmp_path / "cons.fa", tmp_path / "reads.fa"
    ref_fa.write_text(f">cA\n{a}\n>cB\n{b}\n")
    reads_fa.write_text("".join(f">{n}\n{s}\n" for n, s in reads.items()))
    out = {}
    for tag, args in (("hard", "--secondary=no --sam-hit-only -U 25,75 -H"), ("soft", "-Y --sam-hit-only --secondary=no -U 10,25 -H")):
        bam = tmp_path / f"{tag}.bam"
        util.align_reads_with_minimap(reference=ref_fa, reads=reads_fa, bamout=bam, tech="map-ont", threads=1, aln_args=args, engine=oracle)
        out[tag] = list(sam.read_alignments(bam))
    for recs in out.values():
        assert sorted((r.query_name, r.reference_name) for r in recs if not r.is_supplementary) == \
               [("inv_a", "cA"), ("on_a", "cA"), ("on_b_rc", "cB")]
        assert {r.reference_name for r in recs if r.query_name == "inv_a"} == {"cA"} and sum(r.query_name == "inv_a" for r in recs) == 3
        assert next(r for r in recs if r.query_name == "on_b_rc").is_reverse
    hard = [r for r in out["hard"] if r.is_supplementary]
    soft = [r for r in out["soft"] if r.is_supplementary]
    assert len(hard) == len(soft) == 2
    assert all(5 in (r.cigartuples[0][0], r.cigartuples[-1][0]) and len(r.query_sequence) == r.query_alignment_length for r in hard)
    assert all(4 in (r.cigartuples[0][0], r.cigartuples[-1][0]) and len(r.query_sequence) == len(inv) for r in soft)
    assert [(r.reference_start, r.reference_end, r.is_reverse) for r in hard] == [(r.reference_start, r.reference_end, r.is_reverse) for r in soft]


def test_file_seam_timeout_contract(tmp_path, oracle):
    """util.align_reads_with_minimap honours `timeout` like the reference (util/util.py:205-226: TimeoutError, which
    consensus.py:1387-1415 catches to retry with half the reads): a container whose estimated work exceeds the budget is refused
    before anything runs and leaves no BAM; a generous budget runs normally."""
    rng = np.random.default_rng(61)
    tmpl = "".join("ACGT"[k] for k in rng.integers(0, 4, 3000))
    reads = {f"r{k:02d}": tmpl[int(rng.integers(0, 50)):len(tmpl) - int(rng.integers(0, 50))] for k in range(12)}
    fa, bam = tmp_path / "reads.fa", tmp_path / "ava.bam"
    fa.write_text("".join(f">{n}\n{s}\n" for n, s in reads.items()))
    with pytest.raises(TimeoutError):
        util.align_reads_with_minimap(reference=fa, reads=fa, bamout=bam, tech="ava-ont", threads=1, timeout=1e-4, engine=oracle)
    assert not bam.exists()
    with pytest.raises(TimeoutError):
        util.align_reads_with_minimap(reference=fa, reads=fa, bamout=bam, tech="map-ont", threads=1, timeout=1e-4, engine=oracle)
    util.align_reads_with_minimap(reference=fa, reads=fa, bamout=bam, tech="ava-ont", threads=1, timeout=120.0, engine=oracle)
    assert bam.exists() and len(list(sam.read_alignments(bam))) >= 60
    # the estimate grows with the work: 4x the reads = ~16x the pairs
    from svirlpool_b200 import aligner, ava
    names = list(reads)
    small = ava.PairBatch.build_ava(names[:3], [reads[n] for n in names[:3]], None, None, 300, 32, False)[0]
    big = ava.PairBatch.build_ava(names, [reads[n] for n in names], None, None, 300, 32, False)[0]
    sc = aligner.scoring_preset("map-ont")
    assert aligner.estimate_seconds(sc, big, overhead=0) > 10 * aligner.estimate_seconds(sc, small, overhead=0)


def test_ava_pairs_native_unknown_strand_entries():
    """svp_ava_pairs emits both orientations of a pair when either strand is 0 (unknown): the binding sizes its buffer for that
    (strands [1, 0, -1] used to fail with SVP_ERR_NOMEM)."""
    from svirlpool_b200.aligner import ava_pairs_native
    lens = np.array([900, 1000, 1100], dtype=np.uint32)
    pairs, words = ava_pairs_native(lens, None, np.array([1, 0, -1], dtype=np.int8), None, 0, 3, 300, 32, False)
    got = sorted((int(p["q"]), int(p["t"]), int(p["flags"]) & 1) for p in pairs)
    assert got == [(0, 1, 0), (0, 1, 1), (0, 2, 1), (1, 2, 0), (1, 2, 1)] and words > 0
