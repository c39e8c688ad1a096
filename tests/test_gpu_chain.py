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
