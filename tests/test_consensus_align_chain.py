"""consensus_align consumer chain: host mirrors checked against the expectations of the
reference's own tests (tests/test_consensus_class.py:218-482) on the reference's frozen records
(tests/golden/*.json.gz), plus an end-to-end run of the chain on the CPU oracle engine."""
import gzip
import json
from pathlib import Path

import numpy as np
import pytest

from svirlpool_b200 import consensus_align, consensus_align_lib, consensus_class, util
from svirlpool_b200.datatypes import AlignedRecord, Alignment, cigar_from_string
from svirlpool_b200.util import Direction, get_interval_on_ref_in_region, get_read_position_on_ref

GOLDEN = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def ca_golden():
    with gzip.open(GOLDEN / "consensus_align_fixtures.json.gz", "rt") as f:
        return json.load(f)


def rec_from_fixture(r, name="ref"):
    return AlignedRecord(query_name=r["read"], reference_name=name, reference_start=r["ref_start"],
                         cigartuples=cigar_from_string(r["cigar"]), flag=r["flag"])


def rec_from_slim(a):
    return AlignedRecord(query_name=a["query_name"], reference_name=a["reference_name"], reference_start=a["reference_start"],
                         cigartuples=cigar_from_string(a["cigar"]), flag=a["flag"])


def case(golden, name):
    return next(c for c in golden["cases"] if c["name"] == name)


def test_get_interval_on_ref_in_region(golden):
    """tests/test_consensus_class.py:218-240 on the dummy-inversion records (2000M4000H / 2000H2000M2000H rev / 4000S2000M)."""
    recs = {r["cigar"]: rec_from_fixture(r) for r in case(golden, "dummy_inversion")["records"]}
    assert get_interval_on_ref_in_region(recs["2000M4000H"], 1000, 3000) == (1000, 2000)
    assert get_interval_on_ref_in_region(recs["2000H2000M2000H"], 1000, 3000) == (3000, 4000)
    assert get_interval_on_ref_in_region(recs["4000S2000M"], 3000, 5000) == (4000, 5000)


def test_get_read_position_on_ref_simulated(golden):
    """tests/test_consensus_class.py:265-400, value by value."""
    fwd = rec_from_fixture(case(golden, "simulated.simple_forward")["records"][0])
    assert get_read_position_on_ref(fwd, 100, Direction.RIGHT) == 200
    assert get_read_position_on_ref(fwd, 0, Direction.RIGHT) == 100
    assert get_read_position_on_ref(fwd, 766, Direction.LEFT) == 800
    rc = rec_from_fixture(case(golden, "simulated.simple_reverse")["records"][0])
    assert get_read_position_on_ref(rc, 100, Direction.LEFT) == 700
    assert get_read_position_on_ref(rc, 0, Direction.RIGHT) == 800
    assert get_read_position_on_ref(rc, 700, Direction.LEFT) == 100
    dele = rec_from_fixture(case(golden, "simulated.with_deletion")["records"][0])
    for d in (Direction.LEFT, Direction.RIGHT, Direction.NONE):
        assert get_read_position_on_ref(dele, 479, d) == 479
        assert get_read_position_on_ref(dele, 480, d) == {Direction.LEFT: 480, Direction.RIGHT: 500, Direction.NONE: 500}[d]
        assert get_read_position_on_ref(dele, 481, d) == 501
    dele_rc = rec_from_fixture(case(golden, "simulated.with_deletion_reverse")["records"][0])
    for d in (Direction.LEFT, Direction.RIGHT, Direction.NONE):
        assert get_read_position_on_ref(dele_rc, 500, d) == {Direction.LEFT: 500, Direction.RIGHT: 480, Direction.NONE: 480}[d]
        assert get_read_position_on_ref(dele_rc, 499, d) == 501
    ins = rec_from_fixture(case(golden, "simulated.with_insertion")["records"][0])
    for d in (Direction.LEFT, Direction.RIGHT, Direction.NONE):
        for pos in (501, 500, 499):
            assert get_read_position_on_ref(ins, pos, d) == 499


def test_get_ref_position_on_read_roundtrip(golden):
    """Inverse mapping (util.py:1365-1498) on the same records: a reference position inside a match block maps to a
    read position that maps back to it; positions outside the alignment snap to the clip boundaries."""
    for nm in ("simulated.simple_forward", "simulated.simple_reverse", "simulated.with_deletion", "simulated.with_deletion_reverse",
               "simulated.with_insertion"):
        rec = rec_from_fixture(case(golden, nm)["records"][0])
        t, x = zip(*rec.cigartuples)
        _rs, _re, ref_starts, ref_ends = util.get_starts_ends(np.array(t), np.array(x), rec.reference_start, rec.is_reverse)
        for pos in range(rec.reference_start + 1, rec.reference_end - 1, 37):
            in_match = any(t[b] == 0 and ref_starts[b] <= pos < ref_ends[b] for b in range(len(t)))
            rp = util.get_ref_position_on_read(rec, pos, Direction.RIGHT)
            if in_match:
                side = Direction.LEFT if rec.is_reverse else Direction.RIGHT
                assert get_read_position_on_ref(rec, rp, side) == pos, (nm, pos, rp)
    inv = {r["cigar"]: rec_from_fixture(r) for r in case(golden, "dummy_inversion")["records"]}
    mid = inv["2000H2000M2000H"]                      # reverse strand, ref 2000..4000, read 2000..4000
    assert util.get_interval_on_read_in_region(mid, 2000, 4000) == (2000, 4000)
    assert util.get_interval_on_read_in_region(mid, 0, 6000) == (2000, 4000)
    assert util.get_interval_on_read_in_region(mid, 0, 6000, buffer_clipped_length=500) == (1500, 4500)
    first = inv["2000M4000H"]
    assert util.get_interval_on_read_in_region(first, 0, 2000) == (0, 2000)
    assert util.get_interval_on_read_in_region(first, 500, 1500) == (500, 1500)


def test_core_interval_on_reference_inv15(ca_golden):
    """tests/test_consensus_class.py:405-432."""
    cons = consensus_class.Consensus.structure(ca_golden["consensus"]["INV.15"])
    alns = [rec_from_slim(a) for a in ca_golden["alignments"]["INV.15"] if a["query_name"] == "15.0"]
    got = [consensus_class.get_consensus_core_alignment_interval_on_reference(cons, a) for a in alns]
    assert got == [("6", 169094647, 169095161), ("6", 169093632, 169094647), ("6", 169093075, 169093632)]


def test_core_interval_on_reference_inv11_with_hop(ca_golden):
    """tests/test_consensus_class.py:438-455."""
    cons = consensus_class.Consensus.structure(ca_golden["consensus"]["INV.11"])
    alns = [rec_from_slim(a) for a in ca_golden["alignments"]["INV.11"]]
    got = [consensus_class.get_consensus_core_alignment_interval_on_reference(cons, a) for a in alns]
    assert got == [("3_82201219_82205219", 429, 729), ("3_82201219_82205219", 2323, 2396), ("3_82201219_82205219", 2787, 3087)]


def test_parse_sv_signals_from_consensus_inv15(ca_golden):
    """The reference's saved inputs of parse_sv_signals_from_consensus for consensus 15.0; the output for the reverse-strand
    alignment that ends at the inversion breakpoint (reference 168996671-169093632) is frozen in INV.15.mergedSVs.json.gz."""
    outs = {}
    for e in ca_golden["parse_sv_signals"]:
        rec = rec_from_slim(e["alignment"])
        sig = consensus_align_lib.parse_sv_signals_from_consensus(
            samplename=e["samplename"], reference_name=e["reference_name"], consensus_sequence=e["consensus_sequence"],
            pysam_alignment=rec, interval_core=tuple(e["interval_core"]), trf_intervals=[tuple(t) for t in e["trf_intervals"]],
            min_signal_size=e["min_signal_size"], min_bnd_size=e["min_bnd_size"])
        outs[rec.reference_start] = [s.unstructure() for s in sig]
        assert all(e["interval_core"][0] <= s.read_start and s.read_end <= e["interval_core"][1] for s in sig)
    assert outs[168996671] == ca_golden["mergedSVs_INV15"]
    # the inverted segment itself carries both breakends, the far flank the one facing it
    types = {k: [s["sv_type"] for s in v] for k, v in outs.items()}
    assert types[169093632] == [3, 4] and types[169094647] == [3]


def test_assign_repeat_ids_and_alt_sequences():
    from svirlpool_b200.datatypes import MergedSVSignal
    sigs = [MergedSVSignal(100, 100, 50, 70, 20, 0, chr="1"), MergedSVSignal(200, 260, 80, 80, 60, 1, chr="1")]
    consensus_align_lib.assign_repeat_ids(sigs, [(90, 110, 7), (250, 300, 8), (260, 270, 9), (0, 100, 5)])
    assert sorted(sigs[0].repeatIDs) == [7] and sorted(sigs[1].repeatIDs) == [8]      # intervaltree semantics: begin < other.end and end > other.begin
    seq = "ACGT" * 30
    assert consensus_align_lib.alt_sequence_for_MergedSVSignal(sigs[0], seq, False, (40, 160)) == seq[10:30]
    assert consensus_align_lib.alt_sequence_for_MergedSVSignal(sigs[0], seq, True, (40, 160)) == util.reverse_complement(seq[10:30])
    assert consensus_align_lib.alt_sequence_for_MergedSVSignal(sigs[1], seq, False, (40, 160)) == ""


def test_crs_container_result_roundtrip(tmp_path, ca_golden):
    cons = consensus_class.Consensus.structure(ca_golden["consensus"]["INV.15"])
    assert cons.unstructure() == ca_golden["consensus"]["INV.15"]                      # field-for-field the reference's JSON
    from svirlpool_b200.datatypes import SequenceObject
    res = consensus_class.CrsContainerResult(consensus_dicts={cons.ID: cons}, unused_reads={15: [SequenceObject("r", "r", "ACGT")]})
    p = tmp_path / "consensus.batch_0.jsonl"
    consensus_class.write_crs_container_results(p, [res, res])
    back = list(consensus_class.parse_crs_container_results(p))
    assert len(back) == 2 and back[0].unstructure() == res.unstructure()
    assert cons.get_max_cutread_coverage() >= 1 and len(cons.get_used_readnames()) == len(cons.intervals_cutread_alignments)


def _consensus_case(rng, core_len=900, pad=1500, sv=None, strand="+"):
    """A padded consensus (lower-case flanks, UPPER core) cut from a synthetic chromosome, optionally carrying an SV."""
    chrom = "".join("ACGT"[k] for k in rng.integers(0, 4, 12000))
    r_start, r_end = 5000, 5000 + core_len
    left, core, right = chrom[r_start - pad:r_start], chrom[r_start:r_end], chrom[r_end:r_end + pad]
    if sv == "ins":
        core = core[:400] + "".join("ACGT"[k] for k in rng.integers(0, 4, 150)) + core[400:]
    elif sv == "del":
        core = core[:300] + core[500:]
    elif sv == "inv":        # long enough that aligning through it costs more than the Z-drop (a 400 bp inversion is aligned through)
        core = core[:600] + util.reverse_complement(core[600:1800]) + core[1800:]
    seq = left.lower() + core.upper() + right.lower()
    if strand == "-":
        seq = util.reverse_complement(seq)
        interval = (len(right), len(right) + len(core))
    else:
        interval = (len(left), len(left) + len(core))
    padding = consensus_class.ConsensusPadding(sequence=seq, readname_left="l", readname_right="r", padding_size_left=interval[0],
                                               padding_size_right=len(seq) - interval[1], consensus_interval_on_sequence_with_padding=interval)
    core_seq = seq[interval[0]:interval[1]]
    cons = consensus_class.Consensus(ID="7.0", crIDs=[7], original_regions=[("chrS", r_start, r_end)], consensus_sequence=core_seq,
                                     consensus_padding=padding)
    return chrom, cons


@pytest.mark.parametrize("sv,strand", [(None, "+"), ("ins", "+"), ("del", "-"), ("inv", "+")])
def test_chain_end_to_end_on_oracle(oracle, sv, strand):
    """align -> annotate -> core intervals -> SV signals, engine = CPU oracle (the CUDA engine runs the same chain in
    tests/test_gpu_chain.py)."""
    rng = np.random.default_rng(5)
    chrom, cons = _consensus_case(rng, core_len=2400 if sv == "inv" else 900, sv=sv, strand=strand)
    recs = consensus_align.align_consensuses_to_reference(oracle, {cons.ID: cons}, lambda c, s, e: chrom[s:e], {"chrS": len(chrom)},
                                                          band_w=1024)
    check_chain(recs, cons, sv, strand)


def check_chain(recs, cons, sv, strand):
    assert recs and all(r.reference_name == "chrS" for r in recs)
    assert sum(1 for r in recs if not r.is_supplementary) == 1
    ann = consensus_align.annotate_alignments(recs, "S")
    for cid, qs, qe, aln in ann:
        assert cid == "7.0" and 0 <= qs < qe <= len(cons.consensus_padding.sequence)
        assert aln.annotations["qmid"] == (qs + qe) // 2
        assert Alignment.structure(json.loads(json.dumps(aln.unstructure()))).to_record().cigartuples == aln.to_record().cigartuples
    cores = consensus_align.core_intervals_on_reference({cons.ID: cons}, ann)
    sigs = consensus_align.consensus_sv_signals("S", {cons.ID: cons}, ann, min_signal_size=12, min_bnd_size=50)
    flat = [s for per_aln in sigs["7.0"] for s in per_aln]
    if sv is None:
        assert cores["7.0"] == [(0, "chrS", 5000, 5900, *cons.consensus_padding.consensus_interval_on_sequence_with_padding)]
        assert flat == [] and (recs[0].is_reverse == (strand == "-"))
    elif sv == "ins":
        assert [(s.sv_type, s.size) for s in flat] == [(0, 150)] and abs(flat[0].ref_start - 5400) <= 3
        assert len(flat[0].get_alt_sequence()) == 150
    elif sv == "del":
        assert [(s.sv_type, s.size) for s in flat] == [(1, 200)] and abs(flat[0].ref_start - 5300) <= 3
        assert recs[0].is_reverse
    elif sv == "inv":
        assert len(recs) == 3 and sorted(r.is_reverse for r in recs) == [False, False, True]
        assert sorted(s.sv_type for s in flat if s.sv_type >= 3) == [3, 3, 4, 4] or len([s for s in flat if s.sv_type >= 3]) >= 2
        starts = sorted(c[2] for c in cores["7.0"])
        assert abs(starts[0] - 5000) <= 2 and len(cores["7.0"]) == 3


def test_write_partitioned_alignments_reference_signature(oracle, tmp_path):
    """consensus_align.write_partitioned_alignments (reference: consensus_align.py:661-735): records -> partition TSVs by the
    owner of each consensusID, from an in-memory record list and from a BAM path alike; unknown consensusIDs raise."""
    from svirlpool_b200 import sam
    rng = np.random.default_rng(6)
    chrom, cons = _consensus_case(rng, core_len=2400, sv="inv", strand="+")
    recs = consensus_align.align_consensuses_to_reference(oracle, {cons.ID: cons}, lambda c, s, e: chrom[s:e], {"chrS": len(chrom)},
                                                          band_w=1024)
    assert len(recs) == 3
    paths = {0: tmp_path / "p0.tsv", 1: tmp_path / "p1.tsv"}
    consensus_align.write_partitioned_alignments(recs, paths, {0: {"other.0"}, 1: {cons.ID}}, "S")
    assert paths[0].read_text() == ""
    lines = paths[1].read_text().splitlines()
    assert len(lines) == 3
    expected = consensus_align.annotate_alignments(recs, "S")
    for line, (cid, qs, qe, aln) in zip(lines, expected):
        f = line.split("\t")
        assert (f[0], int(f[1]), int(f[2])) == (cid, qs, qe)
        back = Alignment.structure(json.loads(f[3]))
        assert back.annotations == {"qstart": qs, "qend": qe, "qmid": (qs + qe) // 2}
        assert back.to_record().cigartuples == aln.to_record().cigartuples
    # the same through a coordinate-sorted BAM on disk (the file the reference's minimap2 call leaves behind)
    bam = tmp_path / "consensus.bam"
    sam.write_bam(bam, sam.sort_records(recs, {"chrS": len(chrom)}), {"chrS": len(chrom)}, index=True)
    paths2 = {0: tmp_path / "q0.tsv"}
    consensus_align.write_partitioned_alignments(bam, paths2, {0: {cons.ID}}, "S")
    assert sorted(l.split("\t")[:3] for l in paths2[0].read_text().splitlines()) == sorted(l.split("\t")[:3] for l in lines)
    with pytest.raises(ValueError):
        consensus_align.write_partitioned_alignments(recs, {0: tmp_path / "r0.tsv"}, {0: {"nobody"}}, "S")


@pytest.mark.parametrize("kind,size", [("del", 6000), ("ins", 3000), ("del", 2500)])
def test_sv_larger_than_half_the_band_is_one_record(oracle, kind, size):
    """A consensus carrying a deletion / insertion larger than half the first round's band (band_w 4096): the unaligned flank is
    found over the whole window and the two collinear hits are joined into ONE record whose CIGAR holds the event — what
    minimap2's chained alignment reports — instead of a clipped primary (the 6 kb deletion used to come back as `7001M6999S`)."""
    from svirlpool_b200 import ava
    rng = np.random.default_rng(17)
    chrom = "".join("ACGT"[k] for k in rng.integers(0, 4, 30000))
    if kind == "del":
        query = chrom[3000:10000] + chrom[10000 + size:17000 + size]
    else:
        query = chrom[3000:10000] + "".join("ACGT"[k] for k in rng.integers(0, 4, size)) + chrom[10000:17000]
    window, w0 = chrom[1000:26000], 1000
    job = ava.MapJob(name="c", seq=query, target=0, diag_fwd=3000 - w0, diag_rev=None, band_w=4096)
    recs = ava.map_jobs(oracle, [job], ["w"], [window])
    assert len(recs) == 1 and not recs[0].is_supplementary and not recs[0].is_reverse
    ops = [(op, n) for op, n in recs[0].cigartuples if n >= 1000 and op in (1, 2)]
    assert ops == [(2 if kind == "del" else 1, size)]
    assert recs[0].reference_start == 3000 - w0 and recs[0].query_alignment_length == len(query)
    # the consumer chain sees a DEL / INS signal, not breakends
    from svirlpool_b200 import alignments_to_rafs
    sig = alignments_to_rafs.parse_SVsignals_from_alignment(recs[0], recs[0].reference_start, recs[0].reference_end, 0, len(query), 12, 50)
    assert [(s.sv_type, s.size) for s in sig] == [(1 if kind == "del" else 0, size)]


def test_failure_bits_raise(oracle):
    """A pair the kernels cannot run (band wider than a cluster covers) raises AlignerError from the host drivers instead of
    turning into 'no hit' (aligner.py: never silently degrades)."""
    from svirlpool_b200 import ava
    from svirlpool_b200._abi import RESULT_DTYPE
    from svirlpool_b200.aligner import AlignerError
    res = np.zeros(3, dtype=RESULT_DTYPE)
    res["status"] = [0, 8, 1]
    with pytest.raises(AlignerError) as e:
        ava.check_status(res, None, "test")
    assert "BAD_PAIR 1" in str(e.value)
    res["status"] = [0, 1, 2]
    ava.check_status(res)                              # NOHIT / CIGAR_OVF are not failures
    # 70 kb query vs 70 kb target over the "full matrix": the band is capped at what a cluster covers, the pair runs
    lo, w = ava._job_band(70000, 70000, None, 0, 32, cap=ava.max_band_w(oracle))
    assert w == 131008 and -(70000 - 1) <= lo and lo + w <= -(70000 - 1) + 140000
