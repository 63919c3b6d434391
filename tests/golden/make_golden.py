#!/usr/bin/env python
"""Regenerates tests/golden/golden.json with the independent Python model
(tests/golden/fse_model.py): digests of the FreqTable images, the seq / qual
streams and the side buffers of

  * the reference's four fixtures (tests/data = /root/reference/test/data), whole
    file as one chunk, sample = whole file -- what `fqcomp28 c -S 128 -R 256` does
    on them (src/prepare.cpp:42-47) -- i.e. SURVEY.md Appendix C, and
  * a synthetic Illumina fixture (synth.illumina(0, 9000), ~3 MB) whose sequence
    contexts reach table logs 10-11 and whose quality contexts reach log 11: the
    logs every context has in the 1 GB bench and which the libzstd pin
    (tests/test_fse_vs_libzstd.py, logs 5-9) does not cover; tables from the leading
    2 MB window, chunks of 1 MB (several chunks, so chunk boundaries are pinned too).

Every stream is also decoded by the model and compared with the input, so the
model pins itself by round trip.  Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import fse_model as M  # noqa: E402

FIXTURES = ["SRR065390_1_first5", "SRR065390_sub_1", "SRR065390_sub_2", "without_ns"]


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()[:16]


def log_hist(logs):
    h = {}
    for t in logs:
        h[str(t)] = h.get(str(t), 0) + 1
    return h


def split_chunks(fastq: bytes, R: int):
    """chunk boundary rule of FastqReader::readNextChunk, src/fastq_io.cpp:23-65"""
    ends, pos, line = [], 0, 0
    while True:
        nl = fastq.find(b"\n", pos)
        if nl < 0:
            break
        pos = nl + 1
        line += 1
        if line % 4 == 0:
            ends.append(pos)
    offs, s = [0], 0
    import bisect

    while True:
        w = min(s + R, len(fastq))
        k = bisect.bisect_right(ends, w)
        e = ends[k - 1] if k else 0
        if e <= s:
            break
        offs.append(e)
        s = e
        if w >= len(fastq):
            break
    return offs


def entry(recs, fts, ftq):
    enc = M.encode_chunk(recs, fts, ftq)
    ds, dq = M.decode_chunk(enc, fts, ftq, len(recs))
    assert ds == [r[1] for r in recs] and dq == [r[2] for r in recs], "model round trip failed"
    return {"n_records": len(recs), "seq_len": len(enc["seq"]), "seq": sha(enc["seq"]), "qual_len": len(enc["qual"]),
            "qual": sha(enc["qual"]), "readlens": sha(enc["readlens"]), "n_count_len": len(enc["n_count"]),
            "n_count": sha(enc["n_count"]), "n_pos_len": len(enc["n_pos"]), "n_pos": sha(enc["n_pos"])}


def main():
    out = {"fixtures": {}, "synthetic": {}}
    for name in FIXTURES:
        t0 = time.time()
        data = open(os.path.join(ROOT, "tests", "data", name + ".fastq"), "rb").read()
        recs = M.parse(data)
        fts, ftq = M.freq_tables(recs)
        e = entry(recs, fts, ftq)
        e.update({"ft_seq": sha(M.ft_image(fts)), "ft_qual": sha(M.ft_image(ftq)), "seq_logs": log_hist(fts[1]), "qual_logs": log_hist(ftq[1])})
        out["fixtures"][name] = e
        print(name, e, f"{time.time() - t0:.1f}s", file=sys.stderr)
    import synth

    t0 = time.time()
    data = synth.illumina(0, 9000, seed=30).numpy().tobytes()
    R, S = 1 << 20, 2 << 20
    sample = data[: split_chunks(data, S)[1]]
    fts, ftq = M.freq_tables(M.parse(sample))
    offs = split_chunks(data, R)
    syn = {"generator": "synth.illumina(0, 9000, seed=30)", "bytes": len(data), "sha": sha(data), "reading_size": R, "sample_size": S,
           "ft_seq": sha(M.ft_image(fts)), "ft_qual": sha(M.ft_image(ftq)), "seq_logs": log_hist(fts[1]),
           "qual_logs": log_hist(ftq[1]), "chunk_offsets": offs, "chunks": []}
    for k in range(len(offs) - 1):
        syn["chunks"].append(entry(M.parse(data[offs[k] : offs[k + 1]]), fts, ftq))
    out["synthetic"]["illumina_3mb"] = syn
    print("synthetic", {k: v for k, v in syn.items() if k != "chunks"}, f"{time.time() - t0:.1f}s", file=sys.stderr)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
