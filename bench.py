#!/usr/bin/env python
"""bench.py -- FASTQ compress / decompress throughput of the fq28 hot path.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm)

Workload (BASELINE.json configs[1]): synthetic Illumina 150 bp single-end
FASTQ, 1 GB per GPU (weak scaling), the reference's single mode -- static
tables from the leading `-S 128` MB -- at reading size `-R 1` MB.  A step is
one pass of the hot path over the slab:
  compress   = analyzeDataset on the sample (histogram [+ NCCL allreduce of the
               525 312 u32 counters when N > 1] + normalise + CTables/DTables)
               + record splitting + encodeChunk of every chunk
  decompress = decodeChunk of every chunk (layout + tANS decode + N insertion)
`value` has the slab resident in HBM; `e2e` goes through the host-buffer C ABI
(pinned host memory, H2D/D2H inside the timed region).  Headline = compress;
the decompress numbers ride in the "decompress" object (or use --mode).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "FASTQ compress/decompress MB/s at 1/2/4/8 B200, bit-exact vs ref CPU threads"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="compress", choices=["compress", "decompress"])
    ap.add_argument("--size-mb", type=int, default=1024, help="FASTQ MB per GPU")
    ap.add_argument("--reading-mb", type=int, default=1, help="-R, chunk size in MB")
    ap.add_argument("--sample-mb", type=int, default=128, help="-S, sample size in MB")
    ap.add_argument("--profile", default="novaseq", choices=["novaseq", "hiseq", "ont"],
                    help="quality model of the synthetic reads; ont = BASELINE config 4 (1-50 kb reads)")
    ap.add_argument("--cpu-mb", type=int, default=2048, help="bounded sample (MB of the slab) for the CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--threads", type=int, default=0, help="CPU threads (0 = all)")
    return ap.parse_args()


def load_traffic():
    """Per-symbol DRAM traffic of the stage kernels from the latest committed
    ncu --set full capture (profiles/traffic_rNN.json, written by
    profiles/summarize.py)."""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_r*.json")))
    if not files:
        return {}
    try:
        return json.load(open(files[-1]))
    except Exception:
        return {}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = mx
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def make_data(args, rank: int, device: str):
    """Rank r's slab: records [r*M, (r+1)*M) of the virtual global file."""
    import synth

    if args.profile == "ont":  # variable-length long reads, ~24.5 kB per record on average
        import torch

        m = max(1, (args.size_mb << 20) // 24500)
        parts = [synth.ont(rank * m + a, min(1024, m - a), seed=32, device=device) for a in range(0, m, 1024)]
        return torch.cat(parts), m
    per = 150 * 2 + 52
    m = (args.size_mb << 20) // per
    t = synth.illumina(rank * m, m, seed=30, profile=args.profile, device=device)
    return t, m


# --------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch

    from oracle import oracle as O

    dev = "cuda" if torch.cuda.is_available() else "cpu"
    saved = args.size_mb
    args.size_mb = min(args.size_mb, args.cpu_mb)
    t, n_rec = make_data(args, 0, dev)
    args.size_mb = saved
    d = t.cpu().numpy()
    del t
    threads = args.threads or os.cpu_count() or 1
    R, S = args.reading_mb << 20, args.sample_mb << 20
    tc, td = [], []
    res = None
    for i in range(args.warmup + args.steps):
        res = O.bench(d, S, R, threads, True)
        if res.err or not res.roundtrip_ok:
            print(json.dumps({"impl": "reference", "unavailable": f"oracle failed err={res.err}"}))
            return 0
        if i >= args.warmup:
            tc.append(res.t_analyze_s + res.t_compress_s)
            td.append(res.t_decompress_s)
    mb = d.size / 1e6
    t_c, t_d = sum(tc) / len(tc), sum(td) / len(td)
    vc, vd = mb / t_c, mb / t_d
    head, t_head = (vc, t_c) if args.mode == "compress" else (vd, t_d)
    sample = f"{d.size} bytes of the workload ({res.n_records} records, {res.n_chunks} chunks) per step, whole slab"
    line = {
        "impl": "reference", "metric": METRIC, "value": head, "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_head * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, d.size),
        "mode": args.mode,
        "cpu_baseline": {"value": head, "unit": "MB/s", "cores": threads, "kind": "port", "sample": sample,
                         "compress_MBps": vc, "decompress_MBps": vd,
                         "note": "reference cannot be compiled offline (un-vendored zstd fork, libbsc, CLI11); "
                                 "oracle/ port with the reference threading model (one chunk per worker thread)"},
        "e2e": {"value": head, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, n_bytes):
    return {
        "workload": ("synthetic variable-length long reads (1-50 kb, ONT-like qualities)" if args.profile == "ont"
                     else f"synthetic Illumina 150bp single-end FASTQ ({args.profile} qualities)") + f", {args.size_mb} MB per GPU, "
                    f"static tables from the leading -S {args.sample_mb} MB (the reference's only mode), -R {args.reading_mb} MB chunks",
        "fastq_bytes_per_gpu": int(n_bytes), "reading_size_mb": args.reading_mb, "sample_size_mb": args.sample_mb,
        "cache": "inputs (>= 1 GB per GPU) are larger than the 126 MB L2; no explicit flush",
    }


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import fqcomp28_b200 as P
    from fqcomp28_b200 import multigpu as MG

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: fqcomp28_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    R, S = args.reading_mb << 20, args.sample_mb << 20

    d_fastq, n_rec = make_data(args, rank, str(dev))
    n_bytes = d_fastq.numel()
    # sample shard of this rank: the sample is the head of the virtual global
    # file = head of rank 0's slab; shard r = its records [K r/N, K (r+1)/N)
    if world > 1:
        import synth

        k = torch.zeros(1, dtype=torch.int64, device=dev)
        if rank == 0:
            k[0] = int((d_fastq[: min(S, n_bytes)] == 10).sum().item()) // 4
        dist.broadcast(k, 0)
        K = int(k.item())
        a, b = MG.shard_range(K, rank, world)
        d_sample = (synth.ont(a, b - a, seed=32, device=str(dev)) if args.profile == "ont"
                    else synth.illumina(a, b - a, seed=30, profile=args.profile, device=str(dev)))
        sample_window = d_sample.numel()
    else:
        d_sample = d_fastq
        sample_window = min(S, n_bytes)

    stream = torch.cuda.current_stream()
    h = P.Handle(local, stream=stream.cuda_stream)
    cs = torch.zeros(256 * 4, dtype=torch.int32, device=dev)
    cq = torch.zeros(8192 * 64, dtype=torch.int32, device=dev)
    max_chunks = 2 * (n_bytes // R) + 8
    infos = (P.ChunkInfo * max_chunks)()
    fs = np.zeros(P.capi.FT_SEQ_BYTES, np.uint8)
    fq = np.zeros(P.capi.FT_QUAL_BYTES, np.uint8)
    state = {}

    def analyze_dev():
        cs.zero_()
        cq.zero_()
        h.hist_dev(d_sample.data_ptr(), sample_window, cs.data_ptr(), cq.data_ptr())
        if world > 1:
            MG.allreduce_counts(cs, cq)  # the path's only collective (C1): 525 312 u32 counters over NCCL
        f = h.build_tables_dev(cs.data_ptr(), cq.data_ptr())
        state["ft"] = f

    def compress_step_dev():
        t = {}
        analyze_dev()
        t.update({k: v for k, v in h.timings().items() if v})
        _, summ = h.compress_dev(d_fastq.data_ptr(), n_bytes, R, eof=True, max_chunks=max_chunks, infos=infos)
        for k, v in h.timings().items():
            t[k] = t.get(k, 0.0) + v
        state["summ"], state["t_c"] = summ, t

    # ---- device-resident decode setup (after one compress)
    def setup_decode():
        compress_step_dev()
        summ = state["summ"]
        state["enc_summ"] = summ
        setup_e2e()  # pinned host arenas, filled by fq28_compress_fetch
        h.compress_fetch(state["hp"]["arenas"])
        # private device copies: the handle's arenas are overwritten by the next compress
        keep = {k: torch.from_numpy(v.view(np.uint8)).to(dev) for k, v in state["hp"]["arenas"].items()}
        nr = int(summ.n_records)
        d = P.DecArenas()
        d.seq, d.seq_bytes = keep["seq"].data_ptr(), int(summ.seq_bytes)
        d.qual, d.qual_bytes = keep["qual"].data_ptr(), int(summ.qual_bytes)
        d.readlens, d.n_count = keep["readlens"].data_ptr(), keep["n_count"].data_ptr()
        d.n_pos, d.n_pos_entries = keep["n_pos"].data_ptr(), int(summ.n_pos_entries)
        d.hdr_lens = keep["hdr_lens"].data_ptr()
        d.headers, d.headers_bytes = keep["headers"].data_ptr(), int(summ.hdr_bytes)
        d.n_records = nr
        state["dec"] = (d, keep, int(summ.n_chunks))
        state["dec_infos"] = (P.ChunkInfo * int(summ.n_chunks))(*[infos[i] for i in range(int(summ.n_chunks))])
        state["d_out"] = torch.empty(n_bytes + 64, dtype=torch.uint8, device=dev)

    def decompress_step_dev():
        d, _, nch = state["dec"]
        wrote = h.decompress_dev(d, state["dec_infos"], nch, state["d_out"].data_ptr(), n_bytes)
        state["t_d"] = {k: v for k, v in h.timings().items() if v}
        state["wrote"] = wrote

    # ---- host-buffer (e2e) setup
    def pinned(n, dtype=np.uint8):
        t = torch.empty(n, dtype=torch.uint8).pin_memory()
        return t, t.numpy()

    def setup_e2e():
        summ = state["enc_summ"]
        hp = {}
        hp["fastq_t"], hp["fastq"] = pinned(n_bytes)
        hp["fastq_t"].copy_(d_fastq.cpu())
        if world > 1:
            hp["sample_t"], hp["sample"] = pinned(sample_window)
            hp["sample_t"].copy_(d_sample.cpu())
        nr = int(summ.n_records)
        caps = {"seq": int(summ.seq_bytes) + 4096, "qual": int(summ.qual_bytes) + 4096, "readlens": nr * 2 + 64,
                "n_count": nr * 2 + 64, "n_pos": int(summ.n_pos_entries) * 2 + 64, "hdr_lens": nr * 2 + 64,
                "headers": int(summ.hdr_bytes) + 64}
        ar = {}
        for k, nb in caps.items():
            t, a = pinned(nb)
            hp[k + "_t"] = t
            ar[k] = a if k in ("seq", "qual", "headers") else a.view(np.uint16)
        hp["arenas"] = ar
        hp["out_t"], hp["out"] = pinned(n_bytes + 64)
        state["hp"] = hp

    def compress_step_e2e():
        hp = state["hp"]
        if world > 1:
            cs.zero_()
            cq.zero_()
            hs = np.zeros((256, 4), np.uint32)
            # H2D of the sample shard + histogram through the host-buffer ABI would
            # need a host-side reduction; keep the collective on device instead:
            d_tmp = torch.empty(sample_window + 64, dtype=torch.uint8, device=dev)
            d_tmp[:sample_window].copy_(hp["sample_t"], non_blocking=True)
            h.hist_dev(d_tmp.data_ptr(), sample_window, cs.data_ptr(), cq.data_ptr())
            MG.allreduce_counts(cs, cq)
            h.build_tables_dev(cs.data_ptr(), cq.data_ptr())
            _, summ, _ = h.compress(hp["fastq"], R, eof=True, arenas=hp["arenas"])
        else:
            _, summ, _ = h.compress(hp["fastq"], R, eof=True, arenas=hp["arenas"], sample_bytes=S, ft_out=(fs, fq))
        state["e2e_summ"] = summ

    def decompress_step_e2e():
        hp = state["hp"]
        summ = state["enc_summ"]
        out = h.decompress(hp["arenas"], state["dec_infos"], int(summ.n_chunks), hp["arenas"]["headers"][: int(summ.hdr_bytes)],
                           int(summ.n_records), out=hp["out"], n_pos_entries=int(summ.n_pos_entries))
        state["e2e_wrote"] = out.size

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = h.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, (w1 - w0) * 1e3 / steps, (h.launches - l0) // steps

    setup_decode()
    decompress_step_dev()
    roundtrip_ok = bool(state["wrote"] == n_bytes and torch.equal(state["d_out"][:n_bytes], d_fastq))

    sampler = ClockSampler(local) if rank == 0 else None
    ms_c, wall_c, launches_c = timed(compress_step_dev, args.steps, args.warmup)
    t_c = dict(state["t_c"])
    ms_d, wall_d, launches_d = timed(decompress_step_dev, args.steps, args.warmup)
    t_d = dict(state["t_d"])
    clocks = sampler.stop() if sampler else {}
    ms_ce, _, _ = timed(compress_step_e2e, args.steps, args.warmup)
    ms_de, _, _ = timed(decompress_step_e2e, args.steps, args.warmup)
    # e2e correctness: the host round trip restores the slab
    e2e_ok = bool(state["e2e_wrote"] == n_bytes and np.array_equal(state["hp"]["out"][:n_bytes], state["hp"]["fastq"]))

    summ = state["enc_summ"]
    tot_bytes = torch.tensor([n_bytes], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot_bytes)
    total_mb = float(tot_bytes.item()) / 1e6
    peak, peak_src = load_peaks()
    traffic_db = load_traffic()
    nsym = int(summ.n_symbols)
    side = 2 * int(summ.n_records) * 2 + 2 * int(summ.n_pos_entries) + int(summ.hdr_bytes)
    algo_pipeline = n_bytes + int(summ.seq_bytes) + int(summ.qual_bytes) + side  # SURVEY 8(d) `B`

    def roofline(stage_ms: dict, kind: str):
        # dominant kernel = longest stage; algorithmic bytes per launch stated in DESIGN.md section 5
        algo = {
            "chain_seq": nsym * 3, "chain_qual": nsym * 3,          # 1 B symbol read + 2 B field written
            "decode_seq": nsym + int(summ.seq_bytes), "decode_qual": nsym + int(summ.qual_bytes),  # stream read + 1 B/sym written
            "part_seq": nsym * (2 + 1 + 4), "part_qual": nsym * (4 + 4 + 1 + 4),
            "pack_seq": nsym * (4 + 2) * 2, "pack_qual": nsym * (4 + 2) * 2, "extract": 2 * nsym + nsym * 6, "parse": 2 * n_bytes,
            "layout": int(summ.hdr_bytes) * 2 + 5 * int(summ.n_records), "hist": 2 * min(S, n_bytes), "tables": 8448 * 2048 * 10,
            "ninsert": 4 * int(summ.n_records),
        }
        name = max(stage_ms, key=lambda k: stage_ms[k])
        ach = algo.get(name, 0) / (stage_ms[name] * 1e-3) / 1e9
        kern_ms = sum(stage_ms.values())
        return {
            "bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": (traffic_db.get(name, {}).get("dram_bytes_per_symbol") or 0) * nsym or None,
            "traffic_source": "ncu dram__bytes_read+write per symbol of the committed capture x symbols of this run" if name in traffic_db else None,
            "peak_source": peak_src, "kernel_ms": stage_ms[name], "kernel_share_of_step": stage_ms[name] / kern_ms,
            "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
            "pipeline": {"algorithmic_bytes": algo_pipeline, "kernels_ms": kern_ms,
                         "achieved": algo_pipeline / (kern_ms * 1e-3) / 1e9, "frac": algo_pipeline / (kern_ms * 1e-3) / 1e9 / peak},
        }

    out_bytes_c = int(summ.seq_bytes) + int(summ.qual_bytes) + 3 * 2 * int(summ.n_records) + 2 * int(summ.n_pos_entries) + int(summ.hdr_bytes)
    comp = {
        "value": total_mb / (ms_c * 1e-3), "ms_per_step": ms_c, "wall_ms_per_step": wall_c,
        "e2e": {"value": total_mb / (ms_ce * 1e-3), "unit": "MB/s", "ms_per_step": ms_ce,
                "h2d_bytes_per_step": n_bytes + (sample_window if world > 1 else 0), "d2h_bytes_per_step": out_bytes_c},
        "gpu_launches": int(launches_c), "roofline": roofline(t_c, "c"),
    }
    deco = {
        "value": total_mb / (ms_d * 1e-3), "ms_per_step": ms_d, "wall_ms_per_step": wall_d,
        "e2e": {"value": total_mb / (ms_de * 1e-3), "unit": "MB/s", "ms_per_step": ms_de,
                "h2d_bytes_per_step": out_bytes_c, "d2h_bytes_per_step": n_bytes},
        "gpu_launches": int(launches_d), "roofline": roofline(t_d, "d"),
    }

    def gpu_fnv(O):
        """FNV-1a over (seq stream, qual stream) of every chunk in order, on the
        bytes the e2e leg fetched to the host -- same walk as fq28o_bench."""
        ar, hh = state["hp"]["arenas"], 1469598103934665603
        for kk in range(int(summ.n_chunks)):
            ci = state["dec_infos"][kk]
            hh = O.fnv1a(ar["seq"][ci.seq_off : ci.seq_off + ci.seq_len], hh)
            hh = O.fnv1a(ar["qual"][ci.qual_off : ci.qual_off + ci.qual_len], hh)
        return hh

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import oracle as O

            threads = args.threads or os.cpu_count() or 1
            nb = min(n_bytes, args.cpu_mb << 20)
            host = state["hp"]["fastq"][:nb]
            # cut at a record boundary so the sample is a valid FASTQ prefix
            offs = O.split_chunks(host, R)
            host = host[: int(offs[-1])]
            res = O.bench(host, S, R, threads, True)
            vc = host.size / 1e6 / (res.t_analyze_s + res.t_compress_s)
            vd = host.size / 1e6 / res.t_decompress_s
            # parity of this run: FNV over all chunk streams of the same prefix
            cpu_baseline = {
                "value": vc if args.mode == "compress" else vd, "unit": "MB/s", "cores": threads, "kind": "port",
                "sample": f"leading {host.size} bytes of the rank-0 slab ({res.n_chunks} chunks), one pass",
                "compress_MBps": vc, "decompress_MBps": vd, "roundtrip_ok": bool(res.roundtrip_ok),
                "streams_match_gpu": bool(host.size == n_bytes
                                          and res.seq_bytes == sum(int(state["dec_infos"][kk].seq_len) for kk in range(int(summ.n_chunks)))
                                          and res.qual_bytes == sum(int(state["dec_infos"][kk].qual_len) for kk in range(int(summ.n_chunks)))
                                          and res.checksum == gpu_fnv(O)),
            }
        except Exception as e:  # the oracle is a checker, never a dependency of the product arm
            cpu_baseline = {"error": repr(e)}

    if rank == 0:
        head, other = (comp, deco) if args.mode == "compress" else (deco, comp)
        line = {
            "metric": METRIC, "value": head["value"], "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(args, n_bytes), "mode": args.mode,
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
            "clocks": clocks, "cpu_baseline": cpu_baseline,
            "compress": {k: v for k, v in comp.items()}, "decompress": {k: v for k, v in deco.items()},
            "parity": {"roundtrip_device": roundtrip_ok, "roundtrip_e2e": e2e_ok},
            "stats": {"n_chunks": int(summ.n_chunks), "n_records": int(summ.n_records), "seq_bytes": int(summ.seq_bytes),
                      "qual_bytes": int(summ.qual_bytes), "ratio_seq_qual": n_bytes / max(1, int(summ.seq_bytes) + int(summ.qual_bytes))},
        }
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL's
    # version banner) go to stderr, the line itself is written to the saved fd
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(saved, "w")
    try:
        if args.impl == "reference":
            return run_reference(args)
        return run_ours(args)
    finally:
        sys.stdout.flush()


if __name__ == "__main__":
    sys.exit(main())
