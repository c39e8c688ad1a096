#!/usr/bin/env python
"""Builds tests/golden/minimap2_fixtures.json.gz from the reference's own frozen minimap2 outputs.

Run in the build container only (needs /root/reference; the GPU box never reads it):
    python tests/golden/make_golden.py

Sources (all under /root/reference/tests/data, produced by minimap2 2.30-r1287, `-x map-ont`):
    consensus_class/simulated.*.json.gz                     generator: tests/test_consensus_class.py:57-137
    signalprocessing/alignments_to_rafs.dummy_inversion.json.gz   generator: tests/test_alignments_to_rafs.py:18-41
    consensus/test_trimming_{alignments.json.gz,ref.fasta,reads.fasta}   generator: tests/test_consensus.py:29-82
    consensus_class/INV.{11,15}.alignments.json.gz          real data; reference sequence not available
Only DATA is lifted (sequences, SAM fields); no reference source code is copied.
"""
import gzip
import json
import re
from pathlib import Path

import numpy as np

REF = Path("/root/reference/tests/data")
OUT = Path(__file__).parent / "minimap2_fixtures.json.gz"
COMP = str.maketrans("ACGTacgt", "TGCAtgca")


def revcomp(s: str) -> str:
    return s.translate(COMP)[::-1]


def legacy_sequence(size: int, seed: int) -> str:
    # the reference's test helper util.generate_sequence (util/util.py:758-760) draws each base with
    # the legacy global RandomState; restated here to regenerate the inversion fixture's reference.
    np.random.seed(seed)
    letters = list("ACGT")
    return "".join(np.random.choice(letters, p=[0.3, 0.2, 0.2, 0.3]) for _ in range(size))


def tag(sd, key):
    for t in sd["tags"]:
        if t.startswith(key + ":"):
            return t.split(":", 2)[2]
    return None


def record(a):
    sd = a["samdict"]
    return {
        "read": a["readname"],
        "flag": int(sd["flag"]),
        "ref_start": int(a["reference_start"]),
        "ref_end": int(a["reference_end"]),
        "cigar": sd["cigar"],
        "NM": int(tag(sd, "NM")),
        "AS": int(tag(sd, "AS")),
        "sam_seq_len": len(sd["seq"]),
        "minimap2": a["headerdict"]["PG"][0]["VN"],
        "cmdline_opts": re.sub(r"\s/\S+", "", a["headerdict"]["PG"][0]["CL"]),
    }


def load(p):
    with gzip.open(REF / p, "rt") as f:
        return json.load(f)


def read_fasta(p):
    seqs, name = {}, None
    for line in open(REF / p):
        line = line.strip()
        if line.startswith(">"):
            name = line[1:].split()[0]
            seqs[name] = ""
        elif name is not None:
            seqs[name] += line
    return seqs


def main():
    cases = []
    # 1) simulated.* — reference stored in the fixture; read = SAM SEQ (reverse-complemented back if flag 16)
    for nm in ["simple_forward", "simple_reverse", "with_deletion", "with_deletion_reverse", "with_insertion"]:
        d = load(f"consensus_class/simulated.{nm}.json.gz")
        ref = "".join(d["reference"])
        recs, reads = [], {}
        for a in d["alignments"]:
            r = record(a)
            seq = a["samdict"]["seq"]
            reads[r["read"]] = revcomp(seq) if r["flag"] & 16 else seq
            recs.append(r)
        cases.append({"name": f"simulated.{nm}", "reference": ref, "reads": reads, "records": recs,
                      "source": f"tests/data/consensus_class/simulated.{nm}.json.gz"})
    # 2) dummy inversion — reference regenerated (seed 1 + two 8-mers), read = SEQ of the primary record
    d = load("signalprocessing/alignments_to_rafs.dummy_inversion.json.gz")
    ref = list(legacy_sequence(6000, 1))
    ref[1996:2004] = "ACGTACGT"
    ref[3996:4004] = "TGCATGCA"
    ref = "".join(ref)
    read = ref[:2000] + revcomp(ref[2000:4000]) + ref[4000:]
    prim = [a for a in d["alignments"] if int(a["samdict"]["flag"]) & 0x900 == 0][0]
    assert prim["samdict"]["seq"] == read, "regenerated inversion read differs from the fixture SEQ"
    cases.append({"name": "dummy_inversion", "reference": ref, "reads": {"read-0": read},
                  "records": [record(a) for a in d["alignments"]],
                  "source": "tests/data/signalprocessing/alignments_to_rafs.dummy_inversion.json.gz"})
    # 3) read trimming — FASTA files are committed in the reference
    d = load("consensus/test_trimming_alignments.json.gz")
    ref = read_fasta("consensus/test_trimming_ref.fasta")
    reads = read_fasta("consensus/test_trimming_reads.fasta")
    assert len(ref) == 1
    cases.append({"name": "test_trimming", "reference": next(iter(ref.values())), "reads": reads,
                  "records": [record(a) for a in d["alignments"]],
                  "source": "tests/data/consensus/test_trimming_alignments.json.gz"})
    # 4) real INV records — CIGAR/NM/AS only (score-formula check)
    score_only = []
    for nm in ["INV.11", "INV.15"]:
        d = load(f"consensus_class/{nm}.alignments.json.gz")
        for a in d["alignments"]:
            r = record(a)
            r["source"] = f"tests/data/consensus_class/{nm}.alignments.json.gz"
            score_only.append(r)
    n_realign = sum(len(c["records"]) for c in cases)
    out = {"provenance": __doc__, "cases": cases, "score_only": score_only,
           "n_realignable_records": n_realign, "n_records": n_realign + len(score_only)}
    with gzip.open(OUT, "wt") as f:
        json.dump(out, f)
    print(f"wrote {OUT}: {n_realign} re-alignable + {len(score_only)} score-only records")


if __name__ == "__main__":
    main()
