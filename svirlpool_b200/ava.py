"""Host drivers of the alignment hot path: build pair batches, choose bands, turn svp_result rows
into alignment records.

What the reference does with `minimap2 | samtools sort` subprocesses and SAM parsing
    AVA, one container            consensus.py:1362-1400   (-x ava-ont --sam-hit-only -U 25,35 -H)
    reads -> consensus            consensus.py:851-857     (-x map-ont -Y --sam-hit-only --secondary=no)
    consensus -> reference        consensus_align.py:1977-1987
is done here with one `engine.align_batch()` call per batch of pairs. `engine` is
svirlpool_b200.aligner.Aligner (CUDA, the only product engine); tests inject the CPU oracle through
the same interface to check parity.

Alignment semantics: DESIGN.md §3 (banded local affine alignment, two-piece gaps). What is NOT
reproduced from minimap2: seeding/chaining heuristics, MAPQ, the choice of primary vs
supplementary among equal-scoring hits.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ._abi import (PAIR_DTYPE, PAIR_REVCOMP, ST_BAD_PAIR, ST_CIGAR_OVF, ST_NOHIT, ST_SCORE_OVF, ST_TOO_LONG, SVP_ERR_UNSUPPORTED, cigartuples,
                   pack_2bit)
from .datatypes import CIGAR_H, CIGAR_S, AlignedRecord

MIN_DP_SCORE = 80          # minimap2 -s default: hits below this peak DP score are not reported
ST_FAILED = ST_BAD_PAIR | ST_TOO_LONG | ST_SCORE_OVF


def check_status(res: np.ndarray, pairs: np.ndarray | None = None, where: str = "align_batch") -> None:
    """Per-pair failure bits never degrade into 'no hit': a pair the kernels could not run (descriptor rejected, band wider than a
    cluster covers, sequences beyond the descriptors, score range exceeded) raises AlignerError — a subprocess.CalledProcessError,
    what the reference's callers catch around their aligner calls (consensus.py:1387-1396)."""
    bad = np.nonzero(res["status"] & ST_FAILED)[0]
    if len(bad) == 0:
        return
    from .aligner import AlignerError
    k = int(bad[0])
    detail = f"; first: pair {k} status {int(res['status'][k])}"
    if pairs is not None:
        detail += f" (q {int(pairs['q'][k])}, t {int(pairs['t'][k])}, band_lo {int(pairs['band_lo'][k])}, band_w {int(pairs['band_w'][k])})"
    raise AlignerError(SVP_ERR_UNSUPPORTED, f"{where}: {len(bad)} of {len(res)} pairs could not be aligned "
                       f"(BAD_PAIR {int((res['status'] & ST_BAD_PAIR != 0).sum())}, TOO_LONG {int((res['status'] & ST_TOO_LONG != 0).sum())}, "
                       f"SCORE_OVF {int((res['status'] & ST_SCORE_OVF != 0).sum())}){detail}")


def max_band_w(engine) -> int:
    """Widest band the kernels cover for the engine's scoring: a cluster of 8 CTAs x 16 warps — 131 072 diagonals on the s16x2 fast
    path (simple scoring), 65 536 on the general / 32-bit lane kernels (include/svp_align.h: SVP_ST_BAD_PAIR)."""
    sc = getattr(engine, "scoring", None)
    simple = True
    if sc is not None:
        m = np.asarray(sc["matrix"][0]).reshape(4, 4)
        simple = bool((np.diag(m) == m[0, 0]).all() and (m[~np.eye(4, dtype=bool)] == m[0, 1]).all()
                      and int(sc["del_open1"][0]) == int(sc["ins_open1"][0]) and int(sc["del_ext1"][0]) == int(sc["ins_ext1"][0])
                      and int(sc["del_open2"][0]) == int(sc["ins_open2"][0]) and int(sc["del_ext2"][0]) == int(sc["ins_ext2"][0]))
    return (131072 if simple else 65536) - 64


def round_up(x: int, g: int) -> int:
    return (x + g - 1) // g * g


def full_band(m: int, n: int, granule: int = 32) -> tuple[int, int]:
    """Band covering the whole m x n matrix: diagonals -(m-1) .. n-1."""
    lo = -(m - 1)
    return lo, round_up(m + n - 1, granule)


def offset_band(m: int, n: int, d_start: int = 0, d_end: int | None = None, slack: int = 300,
                granule: int = 32) -> tuple[int, int]:
    """Band around the line from diagonal d_start (where the two sequences begin to overlap) to
    d_end (default n - m: both ends co-terminal), widened by slack on both sides and rounded up to
    the granule; never wider than the full matrix."""
    if d_end is None:
        d_end = n - m
    lo = min(d_start, d_end) - slack
    hi = max(d_start, d_end) + slack
    w = round_up(hi - lo + 1, granule)
    flo, fw = full_band(m, n, granule)
    if w >= fw:
        return flo, fw
    return lo, w


def cigar_cap_for(m: int, n: int) -> int:
    """CIGAR slot size (words) for a pair: generous for 10-15 % error reads; an overflow is
    reported per pair (SVP_ST_CIGAR_OVF) and retried by the caller with the worst case m + n."""
    return min(m + n, (m + n) // 3 + 64)


@dataclass
class PairBatch:
    """Sequences + pair descriptors in the layout of svp_align_batch()."""

    names: list[str]
    seqs: list[str]
    packed: np.ndarray
    seq_off: np.ndarray
    seq_len: np.ndarray
    pairs: np.ndarray
    cigar_words: int

    @staticmethod
    def build(names: list[str], seqs: list[str], pair_specs: list[tuple[int, int, int, int, int]],
              cap_fn=cigar_cap_for) -> "PairBatch":
        """pair_specs: (q_index, t_index, band_lo, band_w, flags)."""
        packed, off, lens = pack_2bit(seqs)
        pairs = np.zeros(len(pair_specs), dtype=PAIR_DTYPE)
        pos = 0
        if pair_specs:
            arr = np.asarray(pair_specs, dtype=np.int64).reshape(len(pair_specs), 5)
            m, n = lens[arr[:, 0]].astype(np.int64), lens[arr[:, 1]].astype(np.int64)
            if cap_fn is cigar_cap_for:
                caps = np.minimum(m + n, (m + n) // 3 + 64)
            else:
                caps = np.fromiter((cap_fn(int(a), int(b)) for a, b in zip(m, n)), dtype=np.int64, count=len(m))
            pos = int(caps.sum())
            if pos >= 1 << 32:
                raise ValueError("CIGAR slots of one batch exceed 2^32 words (svp_pair.cigar_off is 32-bit): split the batch")
            pairs["q"], pairs["t"], pairs["band_lo"], pairs["band_w"], pairs["flags"] = arr[:, 0], arr[:, 1], arr[:, 2], arr[:, 3], arr[:, 4]
            pairs["cigar_off"], pairs["cigar_cap"] = np.cumsum(caps) - caps, caps
        return PairBatch(names, seqs, packed, off, lens, pairs, max(pos, 1))


    @staticmethod
    def build_ava(names: list[str], seqs: list[str], strands: dict[str, str] | None, offsets: dict[str, int] | None,
                  slack: int, granule: int, full: bool) -> tuple["PairBatch", list[tuple[int, int, bool]]]:
        """The all-vs-all batch of one container through the library's host-side builders (svp_pack_2bit, svp_ava_pairs):
        the same buffers and pair order as ava_pair_specs() + build(), without the per-pair Python work. Returns the batch
        and the (q, t, revcomp) list of its pairs."""
        from .aligner import ava_pairs_native, pack_2bit_native
        packed, off, lens = pack_2bit_native(seqs)
        order = np.array(sorted(range(len(names)), key=lambda k: names[k]), dtype=np.uint32)
        strand = None if strands is None else np.array([1 if strands[n] == "+" else -1 for n in names], dtype=np.int8)
        offs = None if offsets is None else np.array([offsets[n] for n in names], dtype=np.int32)
        pairs, words = ava_pairs_native(lens, order, strand, offs, 0, len(names), slack, granule, full)
        meta = list(zip(pairs["q"].tolist(), pairs["t"].tolist(), (pairs["flags"] & PAIR_REVCOMP).astype(bool).tolist()))
        return PairBatch(names, seqs, packed, off, lens, pairs, max(words, 1)), meta


def run_batch(engine, batch: PairBatch) -> tuple[np.ndarray, np.ndarray]:
    """Align a batch; pairs whose CIGAR slot overflowed are re-run once with worst-case slots."""
    res, cig = engine.align_batch(batch.packed, batch.seq_off, batch.seq_len, batch.pairs, batch.cigar_words)
    check_status(res, batch.pairs, "run_batch")
    ovf = np.nonzero(res["status"] & ST_CIGAR_OVF)[0]
    if len(ovf):
        specs = [(int(p["q"]), int(p["t"]), int(p["band_lo"]), int(p["band_w"]), int(p["flags"])) for p in batch.pairs[ovf]]
        retry = PairBatch.build(batch.names, batch.seqs, specs, cap_fn=lambda m, n: m + n)
        res2, cig2 = engine.align_batch(retry.packed, retry.seq_off, retry.seq_len, retry.pairs, retry.cigar_words)
        base = len(cig)
        cig = np.concatenate((cig, cig2))
        res2 = res2.copy()
        res2["cigar_off"] += base
        res[ovf] = res2
    return res, cig


def run_summary(engine, batch: PairBatch) -> np.ndarray:
    """Result rows only (Aligner.align_batch_summary: no CIGAR leaves the device). Pairs whose CIGAR slot overflowed —
    their sv_dist is read off the CIGAR on the device — are re-run once with worst-case slots."""
    fn = getattr(engine, "align_batch_summary", None)
    run = (lambda b: fn(b.packed, b.seq_off, b.seq_len, b.pairs, b.cigar_words)) if fn else \
          (lambda b: engine.align_batch(b.packed, b.seq_off, b.seq_len, b.pairs, b.cigar_words)[0])
    res = run(batch)
    check_status(res, batch.pairs, "run_summary")
    ovf = np.nonzero(res["status"] & ST_CIGAR_OVF)[0]
    if len(ovf):
        specs = [(int(p["q"]), int(p["t"]), int(p["band_lo"]), int(p["band_w"]), int(p["flags"])) for p in batch.pairs[ovf]]
        retry = PairBatch.build(batch.names, batch.seqs, specs, cap_fn=lambda m, n: m + n)
        res[ovf] = run(retry)
    return res


def ava_summary(engine, reads: dict[str, str], strands: dict[str, str] | None = None, offsets: dict[str, int] | None = None,
                slack: int = 300, granule: int = 32, full: bool = False) -> tuple[list[str], np.ndarray, np.ndarray, np.ndarray]:
    """All-vs-all of one container without CIGARs: (read names, query index, target index, result rows), the input of
    consensus_lib.similarity_from_summary()."""
    names = list(reads)
    seqs = [reads[k] for k in names]
    if len(names) < 2:
        from ._abi import RESULT_DTYPE
        return names, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=RESULT_DTYPE)
    batch, meta = PairBatch.build_ava(names, seqs, strands, offsets, slack, granule, full)
    res = run_summary(engine, batch)
    return names, np.array([m[0] for m in meta], dtype=np.int64), np.array([m[1] for m in meta], dtype=np.int64), res


def record_from_result(res, cig: np.ndarray, query_name: str, target_name: str, query_len: int, revcomp: bool,
                       supplementary: bool = False, query_sequence: str | None = None,
                       query_offset: int = 0, query_total: int | None = None) -> AlignedRecord | None:
    """One svp_result -> AlignedRecord (None when the pair has no hit). CIGAR = clip + ops + clip in
    the orientation of the target, SAM style: soft clips (S), or hard clips (H) for supplementary
    records. query_offset/query_total place a sub-sequence alignment back on the full read
    (ORIGINAL orientation)."""
    if int(res["status"]) & ST_NOHIT or int(res["cigar_len"]) == 0:
        return None
    total = query_len if query_total is None else query_total
    qb, qe = int(res["q_beg"]), int(res["q_end"])           # oriented sub-sequence coordinates
    if revcomp:
        lead = (total - query_offset - query_len) + qb      # what precedes the hit in the rc'd full read
    else:
        lead = query_offset + qb
    trail = total - lead - (qe - qb)
    clip = CIGAR_H if supplementary else CIGAR_S
    ct = cigartuples(cig, res)
    if lead:
        ct.insert(0, (clip, lead))
    if trail:
        ct.append((clip, trail))
    flag = (16 if revcomp else 0) | (2048 if supplementary else 0)
    nm = int(res["n_mismatch"]) + int(res["n_ins_bp"]) + int(res["n_del_bp"])
    return AlignedRecord(query_name=query_name, reference_name=target_name, reference_start=int(res["t_beg"]),
                         cigartuples=ct, flag=flag, mapping_quality=60, query_sequence=query_sequence,
                         tags={"NM": nm, "AS": int(res["score"])})


# ---------------------------------------------------------------------------------------------
# all-vs-all inside one container
# ---------------------------------------------------------------------------------------------

def ava_pair_specs(names: list[str], lens, strands: dict[str, str] | None, offsets: dict[str, int] | None,
                   slack: int, granule: int, full: bool) -> tuple[list[tuple[int, int, int, int, int]], list[tuple[int, int, bool]]]:
    """Pair list of one container: every unordered pair once, query = the lexicographically smaller
    name (minimap2 -X keeps query < target). Unknown strands -> both orientations are aligned."""
    order = sorted(range(len(names)), key=lambda k: names[k])
    specs, meta = [], []
    for a in range(len(order)):
        for b in range(a + 1, len(order)):
            q, t = order[a], order[b]
            m, n = int(lens[q]), int(lens[t])
            if strands is None:
                orientations = (False, True)
            else:
                orientations = (strands[names[q]] != strands[names[t]],)
            for rc in orientations:
                if full:
                    lo, w = full_band(m, n, granule)
                else:
                    d0 = 0
                    if offsets is not None and not rc:
                        d0 = offsets[names[q]] - offsets[names[t]]
                    lo, w = offset_band(m, n, d_start=d0, d_end=(n - m) if offsets is None or rc else None,
                                        slack=slack, granule=granule)
                specs.append((q, t, lo, w, PAIR_REVCOMP if rc else 0))
                meta.append((q, t, rc))
    return specs, meta


def ava_align(engine, reads: dict[str, str], strands: dict[str, str] | None = None,
              offsets: dict[str, int] | None = None, slack: int = 300, granule: int = 32, full: bool = False,
              min_score: int = MIN_DP_SCORE) -> list[AlignedRecord]:
    """All-vs-all alignment of the reads of one container -> alignment records (the in-memory
    equivalent of the AVA BAM of consensus.py:1378-1400), sorted by (target, position)."""
    names = list(reads)
    seqs = [reads[k] for k in names]
    lens = [len(s) for s in seqs]
    if len(names) < 2:
        return []
    batch, meta = PairBatch.build_ava(names, seqs, strands, offsets, slack, granule, full)
    res, cig = run_batch(engine, batch)
    return ava_records(names, lens, meta, res, cig, min_score)


def build_ava_many(containers: list[dict[str, str]], strands: list[dict[str, str] | None] | None, offsets: list[dict[str, int] | None] | None,
                   slack: int, granule: int, full: bool) -> tuple[PairBatch | None, list[tuple[int, int]], list[int]]:
    """One batch over MANY containers: one sequence table over all reads (svp_pack_2bit), one svp_ava_pairs call per container
    with its seq_base and the running CIGAR word count. Returns (batch or None when no container has two reads,
    [(first sequence, read count)] per container, pair index bounds per container)."""
    from .aligner import ava_pairs_native, pack_2bit_native
    n_c = len(containers)
    strands = strands if strands is not None else [None] * n_c
    offsets = offsets if offsets is not None else [None] * n_c
    all_names, all_seqs, spans = [], [], []
    for reads in containers:
        spans.append((len(all_seqs), len(reads)))
        all_names += list(reads)
        all_seqs += [reads[k] for k in reads]
    packed, off, lens = pack_2bit_native(all_seqs)
    chunks, bounds, words = [], [0], 0
    for (base, n), st, of in zip(spans, strands, offsets):
        names = all_names[base:base + n]
        if n < 2:
            bounds.append(bounds[-1])
            continue
        order = np.array(sorted(range(n), key=lambda k: names[k]), dtype=np.uint32)
        strand = None if st is None else np.array([1 if st[x] == "+" else -1 for x in names], dtype=np.int8)
        offs = None if of is None else np.array([of[x] for x in names], dtype=np.int32)
        pairs, words = ava_pairs_native(lens, order, strand, offs, base, n, slack, granule, full, cigar_start=words)
        chunks.append(pairs)
        bounds.append(bounds[-1] + len(pairs))
    if not chunks:
        return None, spans, bounds
    return PairBatch(all_names, all_seqs, packed, off, lens, np.concatenate(chunks), max(words, 1)), spans, bounds


def ava_align_many(engine, containers: list[dict[str, str]], strands: list[dict[str, str] | None] | None = None,
                   offsets: list[dict[str, int] | None] | None = None, slack: int = 300, granule: int = 32, full: bool = False,
                   min_score: int = MIN_DP_SCORE) -> list[list[AlignedRecord]]:
    """All-vs-all alignment of MANY containers in one engine call — the shape the GPU wants (a crs_to_batches batch holds ~100
    containers, candidateregions/crs_to_batches.py:61-146, and the reference aligns them one minimap2 process at a time,
    consensus.py:3270-3348). Returns per container exactly what ava_align() returns for it."""
    batch, spans, bounds = build_ava_many(containers, strands, offsets, slack, granule, full)
    if batch is None:
        return [[] for _ in containers]
    res, cig = run_batch(engine, batch)
    out = []
    for c, (base, n) in enumerate(spans):
        lo, hi = bounds[c], bounds[c + 1]
        p = batch.pairs[lo:hi]
        meta = list(zip((p["q"] - base).tolist(), (p["t"] - base).tolist(), (p["flags"] & PAIR_REVCOMP).astype(bool).tolist()))
        out.append(ava_records(batch.names[base:base + n], batch.seq_len[base:base + n].tolist(), meta, res[lo:hi], cig, min_score)
                   if hi > lo else [])
    return out


def ava_summary_many(engine, containers: list[dict[str, str]], strands: list[dict[str, str] | None] | None = None,
                     offsets: list[dict[str, int] | None] | None = None, slack: int = 300, granule: int = 32, full: bool = False
                     ) -> list[tuple[list[str], np.ndarray, np.ndarray, np.ndarray]]:
    """ava_summary() for MANY containers in one CIGAR-free engine call: per container (read names, query index, target index,
    result rows), indices local to the container."""
    from ._abi import RESULT_DTYPE
    batch, spans, bounds = build_ava_many(containers, strands, offsets, slack, granule, full)
    res = run_summary(engine, batch) if batch is not None else np.zeros(0, dtype=RESULT_DTYPE)
    out = []
    for c, (base, n) in enumerate(spans):
        lo, hi = bounds[c], bounds[c + 1]
        names = list(containers[c])
        if hi == lo:
            out.append((names, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=RESULT_DTYPE)))
            continue
        p = batch.pairs[lo:hi]
        out.append((names, p["q"].astype(np.int64) - base, p["t"].astype(np.int64) - base, res[lo:hi]))
    return out


def ava_records(names, lens, meta, res, cig, min_score: int = MIN_DP_SCORE) -> list[AlignedRecord]:
    best: dict[tuple[int, int], int] = {}
    for k, (q, t, rc) in enumerate(meta):
        if int(res[k]["status"]) & ST_NOHIT or int(res[k]["score"]) < min_score:
            continue
        cur = best.get((q, t))
        if cur is None or int(res[k]["score"]) > int(res[cur]["score"]):   # equal scores: forward (listed first) wins
            best[(q, t)] = k
    out = []
    for (q, t), k in best.items():
        rec = record_from_result(res[k], cig, names[q], names[t], int(lens[q]), meta[k][2])
        if rec is not None:
            out.append(rec)
    order = {n: k for k, n in enumerate(sorted(names))}
    out.sort(key=lambda r: (order[r.reference_name], r.reference_start, r.query_name))
    return out


# ---------------------------------------------------------------------------------------------
# one query vs one target with supplementary hits (reads -> consensus, consensus -> window)
# ---------------------------------------------------------------------------------------------

@dataclass
class MapJob:
    """One query to be mapped to one target by map_jobs(). diag_fwd / diag_rev: expected diagonal (target position
    minus position on the ORIENTED full query) of the forward / reverse-complement hit, or None when nothing is
    known (full matrix). band_w: diagonals to cover around the expected one."""

    name: str
    seq: str
    target: int = 0
    diag_fwd: int | None = None
    diag_rev: int | None = None
    band_w: int = 0


def _job_band(sub_len: int, n: int, diag: int | None, band_w: int, granule: int, cap: int = 131008) -> tuple[int, int]:
    """Band of one job: band_w diagonals around the expected one, or the whole matrix (nothing known / band_w <= 0) — never more
    than the kernels cover (`cap`): a matrix wider than that keeps its middle `cap` diagonals (around the expected diagonal when
    there is one)."""
    flo, fw = full_band(sub_len, n, granule)
    w = fw if diag is None or band_w <= 0 else min(round_up(band_w, granule), fw)
    if w > cap:
        w = cap // granule * granule
    if w >= fw:
        return flo, fw
    centre = diag if diag is not None else flo + fw // 2
    lo = min(max(centre - w // 2, flo), flo + fw - w)
    return lo, w


def map_jobs(engine, jobs: list[MapJob], target_names: list[str], targets: list[str], min_score: int = MIN_DP_SCORE,
             min_flank: int = 40, max_rounds: int = 8, granule: int = 32) -> list[AlignedRecord]:
    """Map every job's query to its target like `minimap2 -x map-ont --secondary=no`: the best local hit over both
    strands, then — iteratively — the best hit of every unaligned query flank of at least min_flank bases, kept when
    its score reaches min_score. The highest-scoring hit of a query is its primary record (soft clips), the others
    are supplementary (hard clips). All jobs of a round share one svp_align_batch() call. Records come back sorted
    by (target index, target position, query name)."""
    # task: (job index, sub-sequence start, sub-sequence end) in ORIGINAL read coordinates
    tasks = [(k, 0, len(j.seq)) for k, j in enumerate(jobs) if len(j.seq) > 0]
    hits: dict[int, list] = {k: [] for k in range(len(jobs))}
    cap = max_band_w(engine)
    for rnd in range(max_rounds):
        if not tasks:
            break
        names, seqs, specs = list(target_names), list(targets), []
        for k, a, b in tasks:
            job = jobs[k]
            full = len(job.seq)
            names.append(f"{job.name}:{a}-{b}")
            seqs.append(job.seq[a:b])
            n = len(targets[job.target])
            # An unaligned flank is searched over the WHOLE target, not in the band of the first round: what stopped the first hit
            # may be an indel larger than half that band (a 6 kb deletion in a 4 kb band), and the flank then lies on a diagonal
            # the band never reaches — minimap2, seeding against the genome, finds it. The target is a window, so the cost is bounded.
            bw = job.band_w if rnd == 0 else 0
            lo_f, w_f = _job_band(b - a, n, None if job.diag_fwd is None else job.diag_fwd + a, bw, granule, cap)
            lo_r, w_r = _job_band(b - a, n, None if job.diag_rev is None else job.diag_rev + (full - b), bw, granule, cap)
            specs.append((len(seqs) - 1, job.target, lo_f, w_f, 0))
            specs.append((len(seqs) - 1, job.target, lo_r, w_r, PAIR_REVCOMP))
        batch = PairBatch.build(names, seqs, specs)
        res, cig = run_batch(engine, batch)
        nxt = []
        for i, (k, a, b) in enumerate(tasks):
            fwd, rev = res[2 * i], res[2 * i + 1]          # (run_batch raised on failure bits: what is left is NOHIT / CIGAR_OVF after its retry)
            fs = -1 if int(fwd["status"]) & ~ST_CIGAR_OVF else int(fwd["score"])
            rs = -1 if int(rev["status"]) & ~ST_CIGAR_OVF else int(rev["score"])
            pick, rc = (fwd, False) if fs >= rs else (rev, True)
            if max(fs, rs) < min_score:
                continue
            hits[k].append((int(pick["score"]), rc, a, b, pick.copy(), cig))
            qb, qe, sub = int(pick["q_beg"]), int(pick["q_end"]), b - a
            if rc:                                    # oriented -> original coordinates inside the sub-sequence
                qb, qe = sub - qe, sub - qb
            if qb >= min_flank:
                nxt.append((k, a, a + qb))
            if sub - qe >= min_flank:
                nxt.append((k, a + qe, b))
        tasks = nxt
    for _ in range(4):                                # collinear hits on either side of a large indel become one record
        if not _join_collinear(engine, jobs, hits, target_names, targets, granule, cap):
            break
    out = []
    for k, hs in hits.items():
        if not hs:
            continue
        job = jobs[k]
        primary = max(range(len(hs)), key=lambda i: (hs[i][0], -i))
        for i, (score, rc, a, b, r, cig) in enumerate(hs):
            rec = record_from_result(r, cig, job.name, target_names[job.target], b - a, rc, supplementary=(i != primary),
                                     query_offset=a, query_total=len(job.seq))
            if rec is not None:
                out.append((job.target, rec))
    out.sort(key=lambda tr: (tr[0], tr[1].reference_start, tr[1].query_name))
    return [r for _t, r in out]


def _oriented(full: int, hit) -> tuple[int, int, int, int]:
    """(query begin, query end) on the ORIENTED full query and (target begin, target end) of a map_jobs hit."""
    _score, rc, a, b, row, _cig = hit
    off = (full - b) if rc else a
    return off + int(row["q_beg"]), off + int(row["q_end"]), int(row["t_beg"]), int(row["t_end"])


def _join_collinear(engine, jobs: list[MapJob], hits: dict[int, list], target_names: list[str], targets: list[str], granule: int, cap: int,
                    slack: int = 300, tol: int = 50) -> bool:
    """Two hits of a query that follow each other on the same strand in the same order on query and target — the two sides of a
    deletion / insertion larger than the first round's band — are re-aligned as ONE piece in a band that spans both diagonals, and
    replaced by that alignment when it covers both and scores higher than either: the event then is a D / I run of the CIGAR
    (long-gap cost 24 + len for minimap2's scoring), as in minimap2's chained alignment, not a pair of breakends. One join per
    query and call; returns whether anything was joined."""
    cands = []
    for k, hs in hits.items():
        full = len(jobs[k].seq)
        order = sorted(range(len(hs)), key=lambda i: _oriented(full, hs[i])[0])
        for x, y in zip(order, order[1:]):
            if hs[x][1] != hs[y][1]:
                continue
            q1b, q1e, _t1b, t1e = _oriented(full, hs[x])
            q2b, q2e, t2b, _t2e = _oriented(full, hs[y])
            if q2b < q1e - tol or t2b < t1e - tol:
                continue
            d1, d2 = t1e - q1e, t2b - q2b
            w = round_up(abs(d2 - d1) + 2 * slack, granule)
            if w > cap:
                continue
            cands.append((k, x, y, q1b, q2e, min(d1, d2) - slack, w))
            break
    if not cands:
        return False
    names, seqs, specs, pieces = list(target_names), list(targets), [], []
    for k, x, y, qb, qe, lo, w in cands:
        job, rc = jobs[k], hits[k][x][1]
        full = len(job.seq)
        a, b = (full - qe, full - qb) if rc else (qb, qe)              # the piece in ORIGINAL read coordinates
        names.append(f"{job.name}:{a}-{b}")
        seqs.append(job.seq[a:b])
        flo, fw = full_band(b - a, len(targets[job.target]), granule)
        lo_p = lo + qb                                                  # diagonals relative to the piece
        if w >= fw:
            lo_p, w = flo, fw
        specs.append((len(seqs) - 1, job.target, lo_p, w, PAIR_REVCOMP if rc else 0))
        pieces.append((a, b))
    res, cig = run_batch(engine, PairBatch.build(names, seqs, specs))
    joined = False
    for (k, x, y, *_rest), (a, b), row in zip(cands, pieces, res):
        if int(row["status"]) & ~ST_CIGAR_OVF:
            continue
        hs = hits[k]
        if int(row["q_beg"]) > tol or int(row["q_end"]) < (b - a) - tol or int(row["score"]) <= max(hs[x][0], hs[y][0]):
            continue
        merged = (int(row["score"]), hs[x][1], a, b, row.copy(), cig)
        hits[k] = [h for i, h in enumerate(hs) if i not in (x, y)] + [merged]
        joined = True
    return joined


def map_queries_to_best_target(engine, queries: dict[str, str], targets: dict[str, str], min_score: int = MIN_DP_SCORE,
                               min_flank: int = 40, max_rounds: int = 8, granule: int = 32
                               ) -> tuple[list[AlignedRecord], dict[str, str]]:
    """Every query vs EVERY target, keeping — like `minimap2 --secondary=no` against a multi-sequence reference
    (consensus.py:2089-2096: reads vs a FASTA of all consensuses) — only the records on the target of the query's best-scoring
    primary alignment (first target on ties); the primaries on the other targets would be secondary alignments there.
    One engine call per round for all queries x targets. Returns (records sorted by (target, position), {query: target})."""
    names = list(targets)
    jobs = [MapJob(name=q, seq=seq, target=k) for q, seq in queries.items() for k in range(len(names))]
    records = map_jobs(engine, jobs, names, [targets[n] for n in names], min_score, min_flank, max_rounds, granule)
    best: dict[str, tuple[int, int]] = {}
    for rec in records:
        if rec.is_unmapped or rec.is_supplementary:
            continue
        score, k = int(rec.tags.get("AS", 0)), names.index(rec.reference_name)
        cur = best.get(rec.query_name)
        if cur is None or score > cur[0] or (score == cur[0] and k < cur[1]):
            best[rec.query_name] = (score, k)
    chosen = {q: names[k] for q, (_s, k) in best.items()}
    return [rec for rec in records if chosen.get(rec.query_name) == rec.reference_name], chosen


def map_queries(engine, queries: dict[str, str], target_name: str, target: str, min_score: int = MIN_DP_SCORE,
                min_flank: int = 40, max_rounds: int = 8, granule: int = 32) -> list[AlignedRecord]:
    """Every query vs ONE target over the full matrix (reads -> consensus, consensus.py:851-857); see map_jobs()."""
    jobs = [MapJob(name=name, seq=seq) for name, seq in queries.items()]
    return map_jobs(engine, jobs, [target_name], [target], min_score, min_flank, max_rounds, granule)
