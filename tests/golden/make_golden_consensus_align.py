#!/usr/bin/env python
"""Builds tests/golden/consensus_align_fixtures.json.gz: the reference's frozen inputs/outputs of the
consensus_align consumer chain (CIGAR -> core interval on the reference -> SV signals of a consensus).

Run in the build container only (needs /root/reference; the GPU box never reads it):
    python tests/golden/make_golden_consensus_align.py

Sources (all under /root/reference/tests/data):
    consensus_class/INV.{11,15}.consensus.json.gz, INV.{11,15}.alignments.json.gz   (tests/test_consensus_class.py:405-455)
    consensus_class/INV.15.mergedSVs.json.gz
    consensus_align/parse_sv_signals_from_consensus_INV15.*.json.gz                  (inputs saved by consensus_align.py)
Only DATA is lifted (CIGARs, coordinates, intervals, the 2 kb core consensus sequence); the 185 kb SAM SEQ fields
are dropped — the consumers under test do not read them. No reference source code is copied.
"""
import gzip
import json
from pathlib import Path

REF = Path("/root/reference/tests/data")
OUT = Path(__file__).parent / "consensus_align_fixtures.json.gz"


def load(p):
    with gzip.open(REF / p, "rt") as f:
        return json.load(f)


def slim_alignment(a: dict) -> dict:
    sd = a["samdict"]
    return {"query_name": sd["name"], "flag": int(sd["flag"]), "reference_name": sd["ref_name"],
            "reference_start": int(sd["ref_pos"]) - 1, "reference_end": int(a["reference_end"]), "cigar": sd["cigar"],
            "tags": [t for t in sd["tags"] if t.split(":")[0] in ("NM", "AS", "SA")]}


def main():
    out = {"provenance": __doc__, "consensus": {}, "alignments": {}, "parse_sv_signals": []}
    for nm in ("INV.11", "INV.15"):
        c = load(f"consensus_class/{nm}.consensus.json.gz")
        c["consensus_padding"]["sequence"] = ""
        out["consensus"][nm] = c
        out["alignments"][nm] = [slim_alignment(a) for a in load(f"consensus_class/{nm}.alignments.json.gz")["alignments"]]
    for f in sorted((REF / "consensus_align").glob("parse_sv_signals_from_consensus_INV15.*.json.gz")):
        d = load(f"consensus_align/{f.name}")
        aln = d["alignment"]
        entry = {"source": f"tests/data/consensus_align/{f.name}", "samplename": d["samplename"], "reference_name": d["reference_name"],
                 "consensus_sequence": d["consensus_sequence"], "interval_core": d["interval_core"], "trf_intervals": d["trf_intervals"],
                 "min_signal_size": d["min_signal_size"], "min_bnd_size": d["min_bnd_size"],
                 "alignment": slim_alignment(aln) if "samdict" in aln else aln}
        out["parse_sv_signals"].append(entry)
    out["mergedSVs_INV15"] = load("consensus_class/INV.15.mergedSVs.json.gz")
    with gzip.open(OUT, "wt") as f:
        json.dump(out, f)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
