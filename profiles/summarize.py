#!/usr/bin/env python
"""Turns gpurun_out/launches_rNN.csv (ncu --metrics gpu__time_duration.sum) and
gpurun_out/prof_rNN.ncu-rep (ncu --set full) into the tracked summaries under
profiles/.  Usage: python profiles/summarize.py r01"""
import collections
import csv
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go = os.path.join(root, "gpurun_out")
out = os.path.join(root, "profiles")

# ---- launch list
rows = [r for r in csv.reader(l for l in open(os.path.join(go, f"launches_{tag}.csv")) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag}: kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n")
    f.write("Command: `python bench.py --size-mb 256 --sample-mb 64 --steps 1 --warmup 1 --no-cpu-baseline --no-sweep --no-parity` "
            "(256 MB slab, -R 1; every launch of the run incl. setup, warm-up, e2e and copy legs).\n"
            "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
    f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"| `{n}` | {c} | {v / 1e6:.3f} | {100 * v / tot:.1f}% |\n")
    f.write(f"\ntotal {tot / 1e6:.2f} ms over {len(rows) - 1} launches\n")

# ---- full-set metrics
rep = os.path.join(go, f"prof_{tag}.ncu-rep")
raw_csv = os.path.join(go, f"prof_{tag}_raw.csv")   # exported on the GPU box when the report is too big to bring back
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
else:
    raw = open(raw_csv).read()
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
want = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("launch__shared_mem_per_block_static", "static smem"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem pipe %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
]
units = rr[1]
seen = set()
with open(os.path.join(out, f"{tag}_ncu_summary.md"), "w") as f:
    f.write(f"# {tag}: ncu --set full, one launch per kernel (same command as {tag}_launches.md)\n\n")
    f.write("dram read+write per launch = `roofline.traffic`; stall columns are average warps stalled per issue-active cycle.\n\n")
    for r in rr[2:]:
        name = r[h.index("Kernel Name")].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
        for k, label in want:
            if k in h:
                f.write(f"| {label} (`{k}`) | {r[h.index(k)]} | {units[h.index(k)]} |\n")
        f.write("\n")
print("wrote", f"{tag}_launches.md", f"{tag}_ncu_summary.md")

# ---- per-symbol DRAM traffic of the stage kernels -> profiles/traffic_<tag>.json (bench.py `roofline.traffic`)
import json

stage_of = [
    ("k_chain<Kind<unsigned int", "chain_qual"), ("k_chain_dom<Kind<unsigned int", "chain_qual"), ("k_chain<Kind<unsigned short", "chain_seq"),
    ("k_decode_seq", "decode_seq"), ("k_decode_qual", "decode_qual"), ("k_dec2_seq", "decode_seq"), ("k_dec2_qual", "decode_qual"),
    ("k_hist", "hist"), ("k_tile_part8", "part_seq"),
    ("k_tile_rank_compact<Kind<unsigned int", "part_qual"), ("k_tile_hist<Kind<unsigned int", "part_qual"),
    ("k_pack_write<256>", "pack_seq"), ("k_pack_count<256>", "pack_seq"), ("k_pack_write<8192>", "pack_qual"), ("k_pack_count<8192>", "pack_qual"),
    ("k_extract", "extract"), ("k_gather_headers", "extract"), ("k_npos", "extract"),
    ("k_count_nl", "parse"), ("k_fill_nl", "parse"), ("k_records", "parse"), ("k_chunk_walk", "parse"),
]
nsym = None
try:
    line = [l for l in open(os.path.join(go, "plain.log")) if l.startswith("{")][-1]
    nsym = json.loads(line)["stats"]["n_records"] * 150
except Exception:
    pass
traffic = {}
done = set()
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
for r in rr[2:]:
    name = r[h.index("Kernel Name")]
    short = name.split("(")[0]
    for pat, stage in stage_of:
        if pat in name and (stage, short) not in done:  # one launch of every kernel of the stage
            done.add((stage, short))
            rd = float(r[h.index("dram__bytes_read.sum")].replace(",", "")) * scale.get(units[h.index("dram__bytes_read.sum")], 1.0)
            wr = float(r[h.index("dram__bytes_write.sum")].replace(",", "")) * scale.get(units[h.index("dram__bytes_write.sum")], 1.0)
            t = traffic.setdefault(stage, {"dram_bytes_per_launch": 0.0, "nsym_of_capture": nsym, "kernels": []})
            t["dram_bytes_per_launch"] += rd + wr
            t["kernels"].append(short)
for t in traffic.values():
    t["dram_bytes_per_symbol"] = t["dram_bytes_per_launch"] / nsym if nsym else None
json.dump(traffic, open(os.path.join(out, f"traffic_{tag}.json"), "w"), indent=1)
print("wrote", f"traffic_{tag}.json", traffic.get("chain_qual"))
