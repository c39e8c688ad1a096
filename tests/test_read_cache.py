"""Read ingest: the reference's tests/test_read_cache.py, restated on the htslib-free BAM writer /
indexed reader, plus BAI round trips and the hand-over of a region's reads as 2-bit packed arrays."""
from dataclasses import dataclass

import numpy as np
import pytest

from svirlpool_b200 import sam
from svirlpool_b200.datatypes import AlignedRecord
from svirlpool_b200.read_cache import ReadSequenceCache

REFS = {"chr1": 1_000_000, "chr2": 1_000_000}


def _make_bam(tmp_path, reads):
    recs = [AlignedRecord(query_name=r["name"], reference_name=r["chr"], reference_start=int(r["start"]),
                          cigartuples=r.get("cigar", [(0, len(r["seq"]))]), flag=int(r.get("flag", 0)), query_sequence=r["seq"],
                          tags=r.get("tags", {})) for r in reads]
    path = tmp_path / "test.bam"
    sam.write_bam(path, sam.sort_records(recs, REFS), REFS, index=True)
    return path


@dataclass
class _CR:
    crID: int
    chr: str
    referenceStart: int
    referenceEnd: int


def test_cache_dedup_across_overlapping_crs(tmp_path):
    bam = _make_bam(tmp_path, [{"name": "r1", "chr": "chr1", "start": 100, "seq": "A" * 200},
                               {"name": "r2", "chr": "chr1", "start": 150, "seq": "C" * 200},
                               {"name": "r3", "chr": "chr1", "start": 800, "seq": "G" * 100}])
    with ReadSequenceCache(bam) as cache:
        alns1, seqs1 = cache.fetch_for_cr(_CR(1, "chr1", 100, 300))
        alns2, seqs2 = cache.fetch_for_cr(_CR(2, "chr1", 200, 400))
        assert {a.query_name for a in alns1} == {"r1", "r2"} == {a.query_name for a in alns2}
        assert set(seqs1) == {"r1", "r2"} == set(seqs2)
        assert cache.cache_misses_reads == 2 and cache.cache_hits_reads == 2 and cache.fetch_calls == 2


def test_cache_sliding_window_eviction(tmp_path):
    bam = _make_bam(tmp_path, [{"name": "r1", "chr": "chr1", "start": 100, "seq": "A" * 200},
                               {"name": "r2", "chr": "chr1", "start": 150, "seq": "C" * 400},
                               {"name": "r3", "chr": "chr1", "start": 800, "seq": "G" * 100}])
    with ReadSequenceCache(bam) as cache:
        cache.fetch_for_cr(_CR(1, "chr1", 100, 300))
        assert cache.current_size_reads() == 2 and cache.current_size_bp() == 600
        cache.advance(window_start=400)
        assert cache.current_size_reads() == 1 and "r2" in cache.cached_readnames() and "r1" not in cache.cached_readnames()
        cache.advance(window_start=600)
        assert cache.current_size_reads() == 0


def test_cache_flushes_on_chromosome_switch(tmp_path):
    bam = _make_bam(tmp_path, [{"name": "r1", "chr": "chr1", "start": 100, "seq": "A" * 200},
                               {"name": "r2", "chr": "chr2", "start": 100, "seq": "T" * 200}])
    with ReadSequenceCache(bam) as cache:
        cache.fetch_for_cr(_CR(1, "chr1", 100, 300))
        assert cache.current_size_reads() == 1
        cache.fetch_for_cr(_CR(2, "chr2", 100, 300))
        assert "r1" not in cache.cached_readnames() and "r2" in cache.cached_readnames() and cache.current_size_reads() == 1


def test_cache_reverse_complement_preserves_original_orientation(tmp_path):
    seq_on_ref = "ACGTACGTACGT"
    bam = _make_bam(tmp_path, [{"name": "rfwd", "chr": "chr1", "start": 100, "seq": seq_on_ref, "flag": 0},
                               {"name": "rrev", "chr": "chr1", "start": 200, "seq": seq_on_ref, "flag": 16}])
    with ReadSequenceCache(bam) as cache:
        _alns, seqs = cache.fetch_for_cr(_CR(1, "chr1", 50, 300))
        assert seqs["rfwd"] == seq_on_ref
        assert seqs["rrev"] == seq_on_ref[::-1].translate(str.maketrans("ACGT", "TGCA"))


def test_cache_skips_secondary_alignments(tmp_path):
    bam = _make_bam(tmp_path, [{"name": "r1", "chr": "chr1", "start": 100, "seq": "A" * 200},
                               {"name": "r1", "chr": "chr1", "start": 150, "seq": "A" * 200, "flag": 256}])
    with ReadSequenceCache(bam) as cache:
        alns, _seqs = cache.fetch_for_cr(_CR(1, "chr1", 50, 400))
        assert len(alns) == 1 and int(alns[0].flag) == 0


def test_hard_clipped_record_resolved_through_sa_tag(tmp_path):
    full = "ACGTTGCA" * 50                                    # 400 bp read: primary on chr2 (soft clip), supplementary on chr1 (hard clip)
    bam = _make_bam(tmp_path, [
        {"name": "split", "chr": "chr2", "start": 5000, "seq": full, "cigar": [(0, 250), (4, 150)], "tags": {"SA": "chr1,1001,+,250H150M,60,0;"}},
        {"name": "split", "chr": "chr1", "start": 1000, "seq": full[250:], "flag": 2048, "cigar": [(5, 250), (0, 150)],
         "tags": {"SA": "chr2,5001,+,250M150S,60,0;"}}])
    with ReadSequenceCache(bam) as cache:
        alns, seqs = cache.fetch_for_cr(_CR(1, "chr1", 900, 1200))
        assert len(alns) == 1 and alns[0].is_supplementary and seqs["split"] == full


def test_bai_fetch_matches_linear_scan(tmp_path):
    rng = np.random.default_rng(3)
    reads = []
    for k in range(600):                                      # several BGZF blocks, bins on all levels, two references
        ln = int(rng.integers(50, 3000)) if k % 50 else int(rng.integers(20_000, 90_000))
        reads.append({"name": f"r{k}", "chr": "chr1" if k % 3 else "chr2", "start": int(rng.integers(0, 900_000)),
                      "seq": "".join("ACGT"[c] for c in rng.integers(0, 4, min(ln, 300))), "cigar": [(0, min(ln, 300)), (2, max(ln - 300, 1)), (0, 1)] if ln > 300 else None})
    for r in reads:
        if r["cigar"] is None:
            del r["cigar"]
        else:
            r["seq"] += "A"
    bam = _make_bam(tmp_path, reads)
    everything = list(sam.read_alignments(bam))
    assert len(everything) == 600
    with sam.BamFile(bam) as f:
        assert [r.query_name for r in f] == [r.query_name for r in everything]
        for chrom, start, end in [("chr1", 0, 1000), ("chr1", 16_000, 17_000), ("chr1", 100_000, 400_000), ("chr2", 899_000, 1_000_000),
                                  ("chr2", 0, 1_000_000), ("chr1", 500_000, 500_001)]:
            want = [(r.query_name, r.reference_start) for r in everything
                    if r.reference_name == chrom and r.reference_start < end and r.reference_end > start]
            got = [(r.query_name, r.reference_start) for r in f.fetch(chrom, start, end)]
            assert got == want, (chrom, start, end)
        with pytest.raises(ValueError):
            list(f.fetch("chrX", 0, 10))
    (tmp_path / "test.bam.bai").unlink()                       # without an index fetch() falls back to a scan
    with sam.BamFile(bam) as f:
        assert len(list(f.fetch("chr2", 0, 1_000_000))) == sum(1 for r in everything if r.reference_name == "chr2")


def test_packed_hand_over_feeds_the_aligner(tmp_path, oracle):
    """Reads of a candidate region -> 2-bit arrays + strands -> all-vs-all on the engine, no FASTA in between."""
    from svirlpool_b200 import ava, synth
    region = synth.region_config2(1, n_reads=6, mean_len=900, sd_len=40)
    reads = region.read_strings()
    recs = []
    for name, strand, off in zip(region.names, region.strands, region.offsets):
        seq = reads[name]
        on_ref = seq if strand == "+" else seq[::-1].translate(str.maketrans("ACGT", "TGCA"))
        recs.append({"name": name, "chr": "chr1", "start": 10_000 + off, "seq": on_ref, "flag": 0 if strand == "+" else 16})
    bam = _make_bam(tmp_path, recs)
    with ReadSequenceCache(bam) as cache:
        names, packed, off, lens, strands = cache.packed_for_cr(_CR(7, "chr1", 10_000, 11_000))
    assert names == sorted(reads) and strands == dict(zip(region.names, region.strands))
    assert [int(x) for x in lens] == [len(reads[n]) for n in names]
    specs, meta = ava.ava_pair_specs(names, lens, strands, None, 200, 32, False)
    batch = ava.PairBatch.build(names, [reads[n] for n in names], specs)
    assert np.array_equal(batch.packed, packed) and np.array_equal(batch.seq_off, off)
    res, cig = oracle.align_batch(packed, off, lens, batch.pairs, batch.cigar_words)
    assert (res["score"] > 150).all() and not res["status"].any()


def test_read_cutting_from_bam_to_all_vs_all(tmp_path, oracle):
    """BAM -> ReadSequenceCache -> intervals in the candidate region -> cut reads -> all-vs-all: the ingest chain of
    consensus.py:348-586 without pysam. Reads overhang the candidate region on both sides; after cutting they span it."""
    from svirlpool_b200 import ava, consensus, synth
    from svirlpool_b200.datatypes import SVsignal
    rng = np.random.default_rng(5)
    chrom = "".join("ACGT"[c] for c in rng.integers(0, 4, 6000))
    recs = []
    for k in range(6):
        lo, hi = int(rng.integers(200, 900)), int(rng.integers(4200, 5400))
        seq = chrom[lo:hi]
        if k % 2:                                           # a 40 bp deletion inside the region on half of the reads
            seq = chrom[lo:2500] + chrom[2540:hi]
            cigar = [(0, 2500 - lo), (2, 40), (0, hi - 2540)]
        else:
            cigar = [(0, hi - lo)]
        if k == 4:                                          # one read is soft-clipped inside the region
            cigar = [(4, 300)] + [(0, cigar[0][1] - 300)] + cigar[1:]
            lo += 300
        recs.append({"name": f"read{k}", "chr": "chr1", "start": lo, "seq": seq, "cigar": cigar, "flag": 16 if k == 3 else 0})
    bam = _make_bam(tmp_path, recs)

    @dataclass
    class CR:
        crID: int
        chr: str
        referenceStart: int
        referenceEnd: int
        sv_signals: list

    cr = CR(3, "chr1", 2000, 3200, [SVsignal(2500, 2540, 0, 0, 40, 1)])
    with ReadSequenceCache(bam) as cache:
        alns, seqs = cache.fetch_for_cr(cr)
    assert len(alns) == 6
    intervals = consensus.get_read_alignment_intervals_in_cr([cr], buffer_clipped_length=100, dict_alignments={3: alns})
    extents = consensus.get_max_extents_of_read_alignments_on_cr(intervals)
    cut = consensus.trim_reads({3: alns}, extents, seqs)
    assert set(cut) == {f"read{k}" for k in range(6)}
    for k in range(6):
        seq, desc = cut[f"read{k}"]
        want = 1200 - (40 if k % 2 else 0)
        assert abs(len(seq) - want) <= 1 and desc.startswith("crID=3,start=") and "ref_start_chr=chr1" in desc
        rs, re_, _c0, ref_lo, _c1, ref_hi = extents[f"read{k}"]
        assert (ref_lo, ref_hi) == (2000, 3200) and re_ - rs == len(seq)
    # the reverse-strand read was cut in ORIGINAL read orientation: its cut equals the reverse complement of the region
    region_rc = chrom[2000:3200][::-1].translate(str.maketrans("ACGT", "TGCA"))
    assert cut["read3"][0] == (chrom[2000:2500] + chrom[2540:3200])[::-1].translate(str.maketrans("ACGT", "TGCA")) and len(region_rc) == 1200
    reads = {n: s for n, (s, _d) in cut.items()}
    strands = {f"read{k}": "-" if k == 3 else "+" for k in range(6)}
    records = ava.ava_align(oracle, reads, strands=strands, slack=100)
    assert len(records) == 15
    big = [n for r in records for op, n in r.cigartuples if op in (1, 2) and n >= 30]
    assert len(big) == 9 and set(big) == {40}               # 3 x 3 cross-haplotype pairs see the 40 bp indel
