#!/usr/bin/env python3
"""Source-level view of the decoder hot loops from an `ncu --set full --import-source on`
capture (gpurun_out/prof_<tag>_dec.ncu-rep): per source line, warp-stall samples and
executed warp instructions, plus instructions and cycles per decoded symbol.

    python profiles/dec_source.py r02      -> profiles/r02_dec_source.md
    python profiles/dec_source.py r02 REP SUFFIX N_CHUNKS N_RECORDS "TITLE"
                                           -> profiles/r02_dec_source_SUFFIX.md from another capture
"""
import csv
import io
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(root, "gpurun_out", f"prof_{tag}_dec.ncu-rep")
suffix, title = "", "256 MB slab"
if len(sys.argv) > 5:
    rep, suffix = sys.argv[2], "_" + sys.argv[3]
    n_chunks, nsym = int(sys.argv[4]), int(sys.argv[5]) * 150
    title = sys.argv[6] if len(sys.argv) > 6 else sys.argv[3]
else:
    plain = [l for l in open(os.path.join(root, "gpurun_out", "plain.log")) if l.startswith("{")][-1]
    run = json.loads(plain)
    n_chunks = run["stats"]["n_chunks"]
    # Illumina 150 bp synthetic: every record carries 150 bases and 150 qualities
    nsym = run["stats"]["n_records"] * 150

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
kern = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0]
    if name not in kern:
        kern[name] = d

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
# the export is a sequence of blocks: "File Path", "Function Name", header row, then lines
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "Function Name":
        cur = {"fn": r[1].split("(")[0].replace("fq28::", ""), "hdr": None, "lines": []}
        blocks.append(cur)
    elif r[0] == "Line No" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and r[0] not in ("File Path",):
        cur["lines"].append(r)


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


with open(os.path.join(root, "profiles", f"{tag}_dec_source{suffix}.md"), "w") as f:
    f.write(f"# {tag}: decoder hot loops, ncu source view ({title}, {n_chunks} streams per kind, -R 1)\n\n")
    f.write("Same command as the launch list; one launch per kernel; `-lineinfo` maps SASS to `fq28_dec2.cuh`.\n"
            "`samples` = warp-stall samples on the line's instructions, `warp inst` = warp-level instructions executed.\n\n")
    seen = set()
    for b in blocks:
        if b["fn"] in seen or not b["fn"].startswith("k_dec2"):
            continue
        seen.add(b["fn"])
        h = b["hdr"]
        i_line, i_src, i_addr = 0, 1, 2
        i_samp = h.index("# Samples")
        i_inst = h.index("Instructions Executed")
        srcl = [r for r in b["lines"] if r[i_line] != ""]
        sass, addrs = [], set()
        for r in b["lines"]:   # a SASS row is listed under every source line it is attributed to (inlining): once each
            if r[i_line] == "" and r[i_addr] not in addrs:
                addrs.add(r[i_addr])
                sass.append(r)
        tot_s = sum(num(r[i_samp]) for r in sass) or 1.0
        tot_i = sum(num(r[i_inst]) for r in sass)
        k = kern.get(b["fn"]) or kern.get("fq28::" + b["fn"]) or {}
        dur = num(k.get("gpu__time_duration.sum", "0"))
        cyc = num(k.get("sm__cycles_elapsed.max", "0"))
        grid, blk = num(k.get("launch__grid_size", "0")), num(k.get("launch__block_size", "0"))
        warps = grid * blk / 32
        f.write(f"## `{b['fn']}`\n\n")
        f.write(f"duration {dur:.3f} ms ({cyc:.3e} cycles), grid {int(grid)} x {int(blk)} threads = {int(warps)} warps, "
                f"{tot_i:.3e} warp instructions in the kernel")
        if n_chunks and nsym and warps:
            per_stream = nsym / n_chunks
            lanes = n_chunks / warps
            f.write(f"; {per_stream:.0f} symbols per stream, {lanes:.2f} streams per warp -> "
                    f"**{tot_i / warps / per_stream:.1f} warp instructions and {cyc / per_stream:.0f} cycles per symbol step**")
        f.write(".\n\n| line | samples | share | warp inst | source |\n|---:|---:|---:|---:|---|\n")
        for r in sorted(srcl, key=lambda r: -num(r[i_samp]))[:28]:
            f.write(f"| {r[i_line]} | {int(num(r[i_samp]))} | {num(r[i_samp]) / tot_s * 100:.1f}% | {num(r[i_inst]):.3g} | `{r[i_src].strip()[:110]}` |\n")
        f.write("\nTop SASS instructions by samples:\n\n| samples | share | warp inst | SASS |\n|---:|---:|---:|---|\n")
        for r in sorted(sass, key=lambda r: -num(r[i_samp]))[:24]:
            f.write(f"| {int(num(r[i_samp]))} | {num(r[i_samp]) / tot_s * 100:.1f}% | {num(r[i_inst]):.3g} | `{r[3].strip()}` |\n")
        f.write("\n")
print("wrote", f"{tag}_dec_source{suffix}.md")
