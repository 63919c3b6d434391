#!/usr/bin/env python3
"""Wall-clock of the drop-in CLI (`fqcomp28 c|d`, fqcomp28_b200/host/fqcomp28_cli.cpp) on a
synthetic multi-GB FASTQ file, on the GPU box: file -> archive -> file, compared byte for byte.

    python profiles/cli_timing.py [--size-mb 4096] [--gpus 1] > gpurun_out/cli_timing.json

The file is written under $GRAFT_REPO_ROOT/scratch_cli (not copied back).  This times the whole
tool -- file reads, pageable host buffers, header tokenisation, archive writes -- not the codec.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synth  # noqa: E402

CLI = os.path.join(ROOT, "fqcomp28_b200", "fqcomp28")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size-mb", type=int, default=4096)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--reading-mb", default="1,16,256")
    a = ap.parse_args()
    d = os.path.join(os.environ.get("GRAFT_REPO_ROOT", ROOT), "scratch_cli")
    os.makedirs(d, exist_ok=True)
    src, arc, out = (os.path.join(d, n) for n in ("in.fastq", "a.fq28", "out.fastq"))
    piece = 512 << 20
    n = 0
    with open(src, "wb") as f:
        for i in range((a.size_mb << 20) // piece):
            t = synth.illumina_bytes(piece, seed=100, device="cuda", first_record=i * 2_000_000)[0]
            b = t.cpu().numpy().tobytes()
            f.write(b)
            n += len(b)
    os.sync()                      # the input's dirty pages must not be written back under the timed runs
    with open(src, "rb") as f:     # ... and the file is read from the page cache in every run alike
        while f.read(256 << 20):
            pass
    res = {"file_bytes": n, "gpus": a.gpus, "runs": []}
    # fixed cost of the process (CUDA context, library load, table build): a 3 MB file
    tiny = os.path.join(d, "tiny.fastq")
    with open(tiny, "wb") as f:
        f.write(synth.illumina_bytes(3 << 20, seed=5)[0].numpy().tobytes())
    t0 = time.time()
    subprocess.run([CLI, "c", "--i1", tiny, "-o", arc, "-R", "1", "-S", "1"], capture_output=True)
    res["tiny_file_compress_wall_s"] = time.time() - t0
    t0 = time.time()
    subprocess.run([CLI, "d", "-i", arc, "--o1", out], capture_output=True)
    res["tiny_file_decompress_wall_s"] = time.time() - t0
    for R in a.reading_mb.split(","):
        r = {"reading_mb": int(R)}
        t0 = time.time()
        p = subprocess.run([CLI, "c", "--i1", src, "-o", arc, "-R", R, "-S", "128", "--gpus", str(a.gpus)], capture_output=True, text=True)
        r["compress_wall_s"] = time.time() - t0
        r["compress_rc"] = p.returncode
        r["compress_stderr_tail"] = p.stderr.strip().splitlines()[-3:]
        r["archive_bytes"] = os.path.getsize(arc) if os.path.exists(arc) else 0
        t0 = time.time()
        p = subprocess.run([CLI, "d", "-i", arc, "--o1", out, "--gpus", str(a.gpus)], capture_output=True, text=True)
        r["decompress_wall_s"] = time.time() - t0
        r["decompress_rc"] = p.returncode
        r["decompress_stderr_tail"] = p.stderr.strip().splitlines()[-2:]
        r["identical"] = subprocess.run(["cmp", "-s", src, out]).returncode == 0
        r["compress_MBps"] = n / 1e6 / r["compress_wall_s"]
        r["decompress_MBps"] = n / 1e6 / r["decompress_wall_s"]
        res["runs"].append(r)
    t0 = time.time()
    with open(src, "rb") as f:
        while f.read(256 << 20):
            pass
    res["plain_file_read_s"] = time.time() - t0
    print(json.dumps(res))


if __name__ == "__main__":
    main()
