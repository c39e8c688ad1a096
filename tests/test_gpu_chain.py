"""The host chains (consensus_align consumer chain, AVA similarity, reads -> consensus) driven by the CUDA engine:
identical records to the same chains driven by the CPU oracle."""
import numpy as np
import pytest

from svirlpool_b200 import consensus, consensus_align, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from svirlpool_b200.aligner import Aligner
    a = Aligner(device=0, preset="map-ont", tb_bytes=4 << 30)
    yield a
    a.close()


def rec_key(r):
    return (r.query_name, r.reference_name, r.reference_start, r.flag, r.cigarstring, r.tags["AS"], r.tags["NM"])


@pytest.mark.parametrize("sv,strand", [(None, "+"), ("ins", "+"), ("del", "-"), ("inv", "+"), ("inv", "-")])
def test_consensus_align_chain_gpu(gpu, oracle, sv, strand):
    from test_consensus_align_chain import _consensus_case, check_chain
    rng = np.random.default_rng(5)
    chrom, cons = _consensus_case(rng, core_len=2400 if sv == "inv" else 900, sv=sv, strand=strand)
    fetch = lambda c, s, e: chrom[s:e]
    g = consensus_align.align_consensuses_to_reference(gpu, {cons.ID: cons}, fetch, {"chrS": len(chrom)}, band_w=1024)
    o = consensus_align.align_consensuses_to_reference(oracle, {cons.ID: cons}, fetch, {"chrS": len(chrom)}, band_w=1024)
    assert [rec_key(r) for r in g] == [rec_key(r) for r in o]
    if strand == "+":
        check_chain(g, cons, sv, strand)


def test_ava_similarity_and_read_to_consensus_gpu(gpu, oracle):
    region = synth.region_config2(3, n_reads=10, mean_len=2500, sd_len=150)
    reads = region.read_strings()
    strands = dict(zip(region.names, region.strands))
    sim_g, names_g, recs_g = consensus.ava_similarity(gpu, reads, strands=strands)
    sim_o, names_o, recs_o = consensus.ava_similarity(oracle, reads, strands=strands)
    assert names_g == names_o and [rec_key(r) for r in recs_g] == [rec_key(r) for r in recs_o]
    np.testing.assert_allclose(sim_g, sim_o, rtol=1e-12, atol=0)
    # reads of haplotype 0 vs one of them as a stand-in consensus
    hap0 = {n: reads[n] for n, h in zip(region.names, region.haplotypes) if h == 0}
    cname = next(iter(hap0))
    cseq = reads[cname] if strands[cname] == "+" else consensus_align.util.reverse_complement(reads[cname])
    rg, sg = consensus.align_reads_to_consensus(gpu, "S", hap0, "c0", cseq)
    ro, so = consensus.align_reads_to_consensus(oracle, "S", hap0, "c0", cseq)
    assert [rec_key(r) for r in rg] == [rec_key(r) for r in ro]
    assert [s.unstructure() for s in sg] == [s.unstructure() for s in so]


def test_summary_path_equals_record_path_gpu(gpu):
    """CIGAR-free similarity (svp_align_batch_summary + numpy) == record pipeline, both on the CUDA engine."""
    for rid in (4, 8):
        region = synth.region_config2(rid, n_reads=12, mean_len=3000, sd_len=200)
        reads = region.read_strings()
        strands = dict(zip(region.names, region.strands))
        sim_a, names_a, recs = consensus.ava_similarity(gpu, reads, strands=strands, densities_weight=0.0)
        sim_b, names_b, len_b = consensus.ava_similarity_summary(gpu, reads, strands=strands)
        assert names_a == names_b
        np.testing.assert_array_equal(sim_a, sim_b)
        np.testing.assert_array_equal(consensus.pairwise_alignment_lengths(recs, names_a), len_b)


# ---------------------------------------------------------------------------------------------
# the host-path cases of tests/test_host_path.py (run there on the CPU oracle engine) on the CUDA engine: same inputs,
# same expectations — A4 re-homing, the best-target mapping, both S1 file-seam cases, final_consensus, many containers per call
# ---------------------------------------------------------------------------------------------

def test_rehoming_of_unused_reads_gpu(gpu):
    import test_host_path as hp
    hp.test_add_unaligned_reads_to_consensuses_inplace(gpu)


def test_file_seam_gpu(gpu, tmp_path, golden):
    import test_host_path as hp
    hp.test_file_seam_align_reads_with_minimap(tmp_path, golden, gpu)


def test_file_seam_multi_reference_best_target_gpu(gpu, tmp_path):
    import test_host_path as hp
    hp.test_file_seam_multi_reference_and_soft_supplementary(tmp_path, gpu)


def test_final_consensus_and_many_containers_gpu(gpu):
    import test_host_path as hp
    hp.test_final_consensus_record(gpu)
    hp.test_ava_align_many_equals_per_container_calls(gpu)
    hp.test_similarity_of_many_containers_in_one_call(gpu)


def test_best_target_mapping_gpu_equals_oracle(gpu, oracle):
    """ava.map_queries_to_best_target: identical records and identical target choice on both engines."""
    from svirlpool_b200 import ava, util
    rng = np.random.default_rng(77)
    t0 = "".join("ACGT"[k] for k in rng.integers(0, 4, 2200))
    t1 = t0[:900] + t0[1100:]                                        # the same locus minus 200 bases
    t2 = "".join("ACGT"[k] for k in rng.integers(0, 4, 1800))
    targets = {"c0": t0, "c1": t1, "c2": t2}
    queries = {"q_full": t0[50:2100], "q_del": t1[30:1900], "q_rc": util.reverse_complement(t2[100:1700]),
               "q_none": "".join("ACGT"[k] for k in rng.integers(0, 4, 700))}
    rg, cg = ava.map_queries_to_best_target(gpu, queries, targets)
    ro, co = ava.map_queries_to_best_target(oracle, queries, targets)
    assert cg == co == {"q_full": "c0", "q_del": "c1", "q_rc": "c2"}
    assert [rec_key(r) for r in rg] == [rec_key(r) for r in ro]


def test_file_seam_default_engine_gpu(tmp_path, golden):
    """S1 without an explicit engine: the process-wide Aligner (bounded arena) is created on first use and reused."""
    from svirlpool_b200 import aligner, sam, util
    case = next(c for c in golden["cases"] if c["name"] == "simulated.with_insertion")
    ref_fa, reads_fa = tmp_path / "ref.fa", tmp_path / "reads.fa"
    ref_fa.write_text(">ref\n" + case["reference"] + "\n")
    reads_fa.write_text("".join(f">{n}\n{s}\n" for n, s in case["reads"].items()))
    for k in range(2):
        bam = tmp_path / f"out{k}.bam"
        util.align_reads_with_minimap(reference=ref_fa, reads=reads_fa, bamout=bam, tech="map-ont", threads=1)
        assert [r.cigarstring for r in sam.read_alignments(bam)] == ["499M20I501M"]
    assert len(aligner._SHARED) == 1
