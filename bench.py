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
    ap.add_argument("--no-sweep", action="store_true", help="skip the reading-size / chunks-resident sweep")
    ap.add_argument("--sweep-reading-mb", default="16,256", help="extra -R values measured on the same slab (few steps each)")
    ap.add_argument("--resident-mb", type=int, default=3840, help="slab size (MB) of the many-chunks decompress point; 0 = skip")
    ap.add_argument("--resident-slabs", type=int, default=3, help="slabs decoded concurrently for the largest chunks-resident point")
    ap.add_argument("--no-parity", action="store_true", help="skip the per-rank oracle check")
    ap.add_argument("--parity-chunks", type=int, default=48, help="chunks per rank checked stream by stream against the oracle")
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


def make_data(args, rank: int, device: str, extra_bytes: int = 0):
    """Rank r's slab: records [r*M, (r+1)*M) of the virtual global file, plus (extra_bytes > 0)
    enough of the following records to cover extra_bytes of lookahead."""
    import synth
    import torch

    if args.profile == "ont":  # variable-length long reads, ~24.5 kB per record on average
        m = max(1, (args.size_mb << 20) // 24500)
        extra = (extra_bytes // 2000 + 8) if extra_bytes else 0   # a record is at least 1000 bases
        n = m + extra
        parts = [synth.ont(rank * m + a, min(1024, n - a), seed=32, device=device) for a in range(0, n, 1024)]
        return torch.cat(parts), m
    per = 150 * 2 + 52
    m = (args.size_mb << 20) // per
    extra = (extra_bytes // 300 + 64) if extra_bytes else 0       # a record is at least ~340 bytes
    n = m + extra
    step = (1 << 30) // per                                        # generate in <= 1 GB pieces (temporaries are 8 B per byte)
    parts = [synth.illumina(rank * m + a, min(step, n - a), seed=30, profile=args.profile, device=device) for a in range(0, n, step)]
    return (parts[0] if len(parts) == 1 else torch.cat(parts)), m


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
    S = args.sample_mb << 20
    last = rank == world - 1
    baton = MG.Baton(rank, world)
    stream = torch.cuda.current_stream()
    h = P.Handle(local, stream=stream.cuda_stream)
    peak, peak_src = load_peaks()
    traffic_db = load_traffic()

    def allmax(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(t.item())

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
        ms = allmax(e0.elapsed_time(e1))
        if world > 1:
            dist.barrier()
        return ms / steps, (w1 - w0) * 1e3 / steps, (h.launches - l0) // steps

    def pinned(n):
        t = torch.empty(n, dtype=torch.uint8).pin_memory()
        return t, t.numpy()

    # ------------------------------------------------------------------ the file
    # One virtual file of `world` x size_mb: rank r owns records [r*M, (r+1)*M).  With several ranks
    # the chunks are those of ONE boundary walk over the whole file: a rank also holds
    # the next R - 1 bytes, so that the chunk that starts in its range and ends in the next one is
    # its own, and it starts encoding where the previous rank's last chunk ended (Baton).
    class Workload:
        def __init__(self, R, size_mb=None):
            self.R = R
            a = argparse.Namespace(**vars(args))
            if size_mb:
                a.size_mb = size_mb
            t, self.M = make_data(a, rank, str(dev), extra_bytes=(R if (world > 1 and not last) else 0))
            if world > 1:
                nl = torch.nonzero(t == 10).flatten()
                own = int(nl[4 * self.M - 1].item()) + 1
                del nl
                self.slab = t[: MG.slab_end(own, R, t.numel(), last)] if not last else t[:own]
            else:
                own = t.numel()
                self.slab = t
            self.own_bytes = own
            sizes = [own]
            if world > 1:
                g = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
                dist.all_gather(g, torch.tensor([own], dtype=torch.int64, device=dev))
                sizes = [int(x.item()) for x in g]
            self.base = sum(sizes[:rank])          # global offset of this rank's first record
            self.file_bytes = sum(sizes)
            self.n = self.slab.numel()
            self.ptr = self.slab.data_ptr()
            self.max_chunks = 2 * (self.n // R) + 8
            self.infos = (P.ChunkInfo * self.max_chunks)()

    wl = Workload(args.reading_mb << 20)

    # sample shard of this rank: the sample is the head of the file = head of rank 0's slab;
    # shard r = its records [K r/N, K (r+1)/N)
    if world > 1:
        import synth

        k = torch.zeros(1, dtype=torch.int64, device=dev)
        if rank == 0:
            k[0] = int((wl.slab[: min(S, wl.own_bytes)] == 10).sum().item()) // 4
        dist.broadcast(k, 0)
        K = int(k.item())
        a, b = MG.shard_range(K, rank, world)
        d_sample = (synth.ont(a, b - a, seed=32, device=str(dev)) if args.profile == "ont"
                    else synth.illumina(a, b - a, seed=30, profile=args.profile, device=str(dev)))
        sample_window = d_sample.numel()
    else:
        d_sample = wl.slab
        sample_window = min(S, wl.n)
    cs = torch.zeros(256 * 4, dtype=torch.int32, device=dev)
    cq = torch.zeros(8192 * 64, dtype=torch.int32, device=dev)
    tables = {}

    def analyze_dev(t=None):
        cs.zero_()
        cq.zero_()
        h.hist_dev(d_sample.data_ptr(), sample_window, cs.data_ptr(), cq.data_ptr())
        if t is not None:  # stage timers are reset by every ABI call: read them call by call
            for kk, v in h.timings().items():
                t[kk] = t.get(kk, 0.0) + v
        if world > 1:
            MG.allreduce_counts(cs, cq)  # the path's only collective (C1): 525 312 u32 counters over NCCL
        tables["ft"] = h.build_tables_dev(cs.data_ptr(), cq.data_ptr())
        if t is not None:
            for kk, v in h.timings().items():
                t[kk] = t.get(kk, 0.0) + v

    # ------------------------------------------------------------------ one (workload, reading size)
    class Run:
        def __init__(self, w):
            self.w = w
            self.st = {}

        # ---- compress, slab resident
        def compress_dev(self):
            w, t = self.w, {}
            analyze_dev(t)
            baton.next_step()
            if world > 1:
                h.preparse_dev(w.ptr, w.n)                      # does not wait for anybody
                for kk, v in h.timings().items():
                    t[kk] = t.get(kk, 0.0) + v
                t0 = time.perf_counter()
                cut = baton.recv()                               # where the previous rank's last chunk ends
                t1 = time.perf_counter()
                consumed, _ = h.plan_cut_dev(w.ptr, w.n, w.R, last, cut - w.base)
                t2 = time.perf_counter()
                baton.send(w.base + consumed)
                self.st["cut_local"] = cut - w.base
                self.st["hop_ms"] = {"wait_for_cut": (t1 - t0) * 1e3, "walk_call": (t2 - t1) * 1e3}
            _, summ = h.compress_dev(w.ptr, w.n, w.R, eof=last, max_chunks=w.max_chunks, infos=w.infos)
            for kk, v in h.timings().items():
                t[kk] = t.get(kk, 0.0) + v
            self.st["summ"], self.st["t_c"] = summ, t

        # chunks of this rank: with a cut inside the slab, chunk 0 is the dropped head
        def kept(self):
            summ = self.st["summ"]
            k0 = 1 if self.st.get("cut_local", 0) > 0 else 0
            return k0, int(summ.n_chunks)

        def setup(self):
            """after one compress: pinned host arenas (= what the e2e legs use), private device
            copies for the device-resident decode, chunk infos rebased to this rank's chunks"""
            w = self.w
            self.compress_dev()
            summ = self.st["summ"]
            nr = int(summ.n_records)
            hp = {}
            hp["fastq_t"], hp["fastq"] = pinned(w.n)
            hp["fastq_t"].copy_(w.slab.cpu())
            if world > 1:
                hp["sample_t"], hp["sample"] = pinned(sample_window)
                hp["sample_t"].copy_(d_sample.cpu())
            caps = {"seq": int(summ.seq_bytes) + 4096, "qual": int(summ.qual_bytes) + 4096, "readlens": nr * 2 + 64,
                    "n_count": nr * 2 + 64, "n_pos": int(summ.n_pos_entries) * 2 + 64, "hdr_lens": nr * 2 + 64,
                    "headers": int(summ.hdr_bytes) + 64}
            ar = {}
            for kk, nb in caps.items():
                t, a = pinned(nb)
                hp[kk + "_t"] = t
                ar[kk] = a if kk in ("seq", "qual", "headers") else a.view(np.uint16)
            hp["arenas"] = ar
            self.hp = hp
            h.compress_fetch(ar)
            k0, k1 = self.kept()
            b = w.infos[k0]
            r0, p0, h0 = int(b.rec_off), int(b.n_pos_off), int(b.hdr_off)
            self.abs_infos = [w.infos[i] for i in range(k0, k1)]
            dec = (P.ChunkInfo * (k1 - k0))()
            out_bytes = 0
            for i in range(k0, k1):
                ci = P.ChunkInfo.from_buffer_copy(bytes(w.infos[i]))
                ci.rec_off -= r0
                ci.n_pos_off -= p0
                ci.hdr_off -= h0
                dec[i - k0] = ci
                out_bytes += int(ci.total)
            self.dec_infos, self.n_dec, self.out_bytes = dec, k1 - k0, out_bytes
            self.first_byte = int(b.fastq_off)
            # host views of this rank's chunks (e2e decompress)
            self.host_ar = {"seq": ar["seq"], "qual": ar["qual"], "readlens": ar["readlens"][r0:], "n_count": ar["n_count"][r0:],
                            "n_pos": ar["n_pos"][p0:], "hdr_lens": ar["hdr_lens"][r0:], "headers": ar["headers"][h0 : int(summ.hdr_bytes)]}
            self.n_rec_dec, self.n_pos_dec = nr - r0, int(summ.n_pos_entries) - p0
            keep = {kk: torch.from_numpy(v.view(np.uint8)).to(dev) for kk, v in ar.items()}
            d = P.DecArenas()
            d.seq, d.seq_bytes = keep["seq"].data_ptr(), int(summ.seq_bytes)
            d.qual, d.qual_bytes = keep["qual"].data_ptr(), int(summ.qual_bytes)
            d.readlens, d.n_count = keep["readlens"].data_ptr() + 2 * r0, keep["n_count"].data_ptr() + 2 * r0
            d.n_pos, d.n_pos_entries = keep["n_pos"].data_ptr() + 2 * p0, self.n_pos_dec
            d.hdr_lens = keep["hdr_lens"].data_ptr() + 2 * r0
            d.headers, d.headers_bytes = keep["headers"].data_ptr() + h0, int(summ.hdr_bytes) - h0
            d.n_records = self.n_rec_dec
            self.dec, self.keep = d, keep
            self.d_out = torch.empty(out_bytes + 64, dtype=torch.uint8, device=dev)
            hp["out_t"], hp["out"] = pinned(out_bytes + 64)
            self.enc_summ = summ
            self.out_bytes_c = int(summ.seq_bytes) + int(summ.qual_bytes) + 3 * 2 * nr + 2 * int(summ.n_pos_entries) + int(summ.hdr_bytes)

        def decompress_dev(self):
            self.st["wrote"] = h.decompress_dev(self.dec, self.dec_infos, self.n_dec, self.d_out.data_ptr(), self.out_bytes)
            self.st["t_d"] = {kk: v for kk, v in h.timings().items() if v}

        def roundtrip_dev(self):
            want = self.w.slab[self.first_byte : self.first_byte + self.out_bytes]
            return bool(self.st["wrote"] == self.out_bytes and torch.equal(self.d_out[: self.out_bytes], want))

        # ---- the same through the host-buffer C ABI (pinned host memory, copies inside the call)
        def compress_e2e(self):
            w, hp = self.w, self.hp
            baton.next_step()
            if world > 1:
                cs.zero_()
                cq.zero_()
                # the sample shard goes H2D, the histogram is reduced on the device (no host-side sum)
                d_tmp = torch.empty(sample_window + 64, dtype=torch.uint8, device=dev)
                d_tmp[:sample_window].copy_(hp["sample_t"], non_blocking=True)
                h.hist_dev(d_tmp.data_ptr(), sample_window, cs.data_ptr(), cq.data_ptr())
                MG.allreduce_counts(cs, cq)
                h.build_tables_dev(cs.data_ptr(), cq.data_ptr())
                h.preparse(hp["fastq"])                         # H2D + record table, before the cut is known
                cut = baton.recv()
                consumed, _ = h.plan_cut(hp["fastq"], w.R, last, cut - w.base)
                baton.send(w.base + consumed)
                _, summ, _ = h.compress(hp["fastq"], w.R, eof=last, arenas=hp["arenas"])
            else:
                fs = np.zeros(P.capi.FT_SEQ_BYTES, np.uint8)
                fq = np.zeros(P.capi.FT_QUAL_BYTES, np.uint8)
                _, summ, _ = h.compress(hp["fastq"], w.R, eof=True, arenas=hp["arenas"], sample_bytes=S, ft_out=(fs, fq))
            self.st["e2e_summ"] = summ

        def decompress_e2e(self):
            out = h.decompress(self.host_ar, self.dec_infos, self.n_dec, self.host_ar["headers"], self.n_rec_dec,
                               out=self.hp["out"], n_pos_entries=self.n_pos_dec)
            self.st["e2e_wrote"] = out.size

        def roundtrip_e2e(self):
            want = self.hp["fastq"][self.first_byte : self.first_byte + self.out_bytes]
            return bool(np.array_equal(self.hp["out"][: self.out_bytes], want))

        # ---- a caller that keeps two slabs in flight: two handles, one host thread each
        def streaming_e2e(self, calls=6):
            """Steady state of the same host-buffer calls when the caller overlaps them (what a file
            compressor does, and what `fqcomp28 c|d --gpus` does per device with its worker threads):
            two handles, one host thread each, `calls` calls in all; the copies of one call run under
            the kernels of the other.  Wall clock around all calls, every call with its H2D / D2H."""
            import threading

            w, hp = self.w, self.hp
            h2 = P.Handle(local)
            h2.load_tables(*tables["ft"])
            keep, ar2 = [], {}
            for kk, v in hp["arenas"].items():
                t, a = pinned(v.view(np.uint8).size)
                keep.append(t)
                ar2[kk] = a if kk in ("seq", "qual", "headers") else a.view(np.uint16)
            out2_t, out2 = pinned(self.out_bytes + 64)
            half = max(1, calls // 2)

            def pair(fa, fb):
                ta, tb = threading.Thread(target=fa), threading.Thread(target=fb)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ta.start()
                tb.start()
                ta.join()
                tb.join()
                torch.cuda.synchronize()
                return (time.perf_counter() - t0) * 1e3

            def c_loop(hh, ar, n):
                fs = np.zeros(P.capi.FT_SEQ_BYTES, np.uint8)
                fq = np.zeros(P.capi.FT_QUAL_BYTES, np.uint8)
                for _ in range(n):
                    hh.compress(hp["fastq"], w.R, eof=True, arenas=ar, sample_bytes=S, ft_out=(fs, fq))

            def d_loop(hh, out, n):
                for _ in range(n):
                    hh.decompress(self.host_ar, self.dec_infos, self.n_dec, self.host_ar["headers"], self.n_rec_dec,
                                  out=out, n_pos_entries=self.n_pos_dec)

            pair(lambda: c_loop(h, hp["arenas"], 1), lambda: c_loop(h2, ar2, 1))
            ms_c2 = pair(lambda: c_loop(h, hp["arenas"], half), lambda: c_loop(h2, ar2, half))
            nb = int(self.enc_summ.seq_bytes)
            same_c = bool(np.array_equal(ar2["seq"][:nb], hp["arenas"]["seq"][:nb]))
            pair(lambda: d_loop(h, hp["out"], 1), lambda: d_loop(h2, out2, 1))
            ms_d2 = pair(lambda: d_loop(h, hp["out"], half), lambda: d_loop(h2, out2, half))
            same_d = bool(np.array_equal(out2[: self.out_bytes], hp["out"][: self.out_bytes]))
            h2.close()
            mb = 2 * half * w.own_bytes / 1e6
            return {"handles": 2, "calls": 2 * half, "compress_MBps": mb / (ms_c2 * 1e-3), "decompress_MBps": mb / (ms_d2 * 1e-3),
                    "compress_ms_per_call": ms_c2 / (2 * half), "decompress_ms_per_call": ms_d2 / (2 * half),
                    "outputs_identical": same_c and same_d, "timer": "host wall clock around all calls",
                    "note": "same calls, buffers and bytes as e2e; the caller keeps two calls in flight"}

        # ---- host link ceiling: the same bytes, the same pinned buffers, copies only
        def copy_legs(self, steps):
            w, hp = self.w, self.hp
            d_in = torch.empty(w.n + 64, dtype=torch.uint8, device=dev)
            d_res = torch.empty(self.out_bytes_c + 64, dtype=torch.uint8, device=dev)
            h_res_t, _ = pinned(self.out_bytes_c + 64)
            side = torch.cuda.Stream()

            def h2d_in():
                d_in[: w.n].copy_(hp["fastq_t"], non_blocking=True)

            def d2h_res():
                h_res_t[: self.out_bytes_c].copy_(d_res[: self.out_bytes_c], non_blocking=True)

            def h2d_res():
                d_res[: self.out_bytes_c].copy_(h_res_t[: self.out_bytes_c], non_blocking=True)

            def d2h_out():
                hp["out_t"][: self.out_bytes].copy_(self.d_out[: self.out_bytes], non_blocking=True)

            def both_c():   # compress direction: input down, result up, at the same time
                side.wait_stream(stream)
                h2d_in()
                with torch.cuda.stream(side):
                    d2h_res()
                stream.wait_stream(side)

            def both_d():
                side.wait_stream(stream)
                h2d_res()
                with torch.cuda.stream(side):
                    d2h_out()
                stream.wait_stream(side)

            r = {}
            for name, fn in (("h2d_fastq", h2d_in), ("d2h_result", d2h_res), ("h2d_result", h2d_res), ("d2h_fastq", d2h_out),
                             ("compress_both_ways", both_c), ("decompress_both_ways", both_d)):
                r[name + "_ms"] = timed(fn, steps, 1)[0]
            return r

    def roofline(run, stage_ms: dict):
        summ, w = run.enc_summ, run.w
        nsym = int(summ.n_symbols)
        side = 2 * int(summ.n_records) * 2 + 2 * int(summ.n_pos_entries) + int(summ.hdr_bytes)
        algo_pipeline = w.n + int(summ.seq_bytes) + int(summ.qual_bytes) + side  # SURVEY 8(d) `B`
        # dominant kernel = longest stage; algorithmic bytes per launch stated in DESIGN.md section 5
        algo = {
            "chain_seq": nsym * 3, "chain_qual": nsym * 3,          # 1 B symbol read + 2 B field written
            "decode_seq": nsym + int(summ.seq_bytes), "decode_qual": nsym + int(summ.qual_bytes),  # stream read + 1 B/sym written
            "part_seq": nsym * (2 + 1 + 4), "part_qual": nsym * (4 + 4 + 1 + 4),
            "pack_seq": nsym * (4 + 2) * 2, "pack_qual": nsym * (4 + 2) * 2, "extract": 2 * nsym + nsym * 6, "parse": 2 * w.n,
            "layout": int(summ.hdr_bytes) * 2 + 5 * int(summ.n_records), "hist": 2 * min(S, w.n), "tables": 8448 * 2048 * 10,
            "ninsert": 4 * int(summ.n_records),
        }
        name = max(stage_ms, key=lambda kk: stage_ms[kk])
        ach = algo.get(name, 0) / (stage_ms[name] * 1e-3) / 1e9
        kern_ms = sum(stage_ms.values())
        return {
            "bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": (traffic_db.get(name, {}).get("dram_bytes_per_symbol") or 0) * nsym or None,
            "traffic_source": "ncu dram__bytes_read+write per symbol of the committed capture x symbols of this run" if name in traffic_db else None,
            "peak_source": peak_src, "kernel_ms": stage_ms[name], "kernel_share_of_step": stage_ms[name] / kern_ms,
            "stage_ms": {kk: round(v, 4) for kk, v in stage_ms.items()},
            "pipeline": {"algorithmic_bytes": algo_pipeline, "kernels_ms": kern_ms,
                         "achieved": algo_pipeline / (kern_ms * 1e-3) / 1e9, "frac": algo_pipeline / (kern_ms * 1e-3) / 1e9 / peak},
        }

    # ------------------------------------------------------------------ headline measurement
    run = Run(wl)
    run.setup()
    run.decompress_dev()
    roundtrip_ok = run.roundtrip_dev()
    total_mb = allsum(float(wl.own_bytes)) / 1e6

    sampler = ClockSampler(local) if rank == 0 else None
    ms_c, wall_c, launches_c = timed(run.compress_dev, args.steps, args.warmup)
    t_c = dict(run.st["t_c"])
    ms_d, wall_d, launches_d = timed(run.decompress_dev, args.steps, args.warmup)
    t_d = dict(run.st["t_d"])
    clocks = sampler.stop() if sampler else {}
    ms_ce, _, _ = timed(run.compress_e2e, args.steps, args.warmup)
    ms_de, _, _ = timed(run.decompress_e2e, args.steps, args.warmup)
    e2e_ok = bool(run.st["e2e_wrote"] == run.out_bytes) and run.roundtrip_e2e()
    copies = run.copy_legs(max(2, args.steps))
    streaming = None
    if world == 1 and not args.no_sweep:
        try:
            streaming = run.streaming_e2e()
        except Exception as e:  # (an extra data point: never the reason a bench line is lost)
            streaming = {"error": repr(e)}
    summ = run.enc_summ

    comp = {
        "value": total_mb / (ms_c * 1e-3), "ms_per_step": ms_c, "wall_ms_per_step": wall_c,
        "e2e": {"value": total_mb / (ms_ce * 1e-3), "unit": "MB/s", "ms_per_step": ms_ce,
                "h2d_bytes_per_step": wl.n + (sample_window if world > 1 else 0), "d2h_bytes_per_step": run.out_bytes_c,
                "copy_only_ms": copies["compress_both_ways_ms"], "fraction_of_copy_ceiling": copies["compress_both_ways_ms"] / ms_ce},
        "gpu_launches": int(launches_c), "roofline": roofline(run, t_c),
    }
    if world > 1:   # the cut chain (last timed step): per rank, how long it waited for its cut and how long its walk took
        hops = [None] * world
        dist.all_gather_object(hops, run.st.get("hop_ms"))
        comp["cut_chain_ms"] = hops
    deco = {
        "value": total_mb / (ms_d * 1e-3), "ms_per_step": ms_d, "wall_ms_per_step": wall_d,
        "e2e": {"value": total_mb / (ms_de * 1e-3), "unit": "MB/s", "ms_per_step": ms_de,
                "h2d_bytes_per_step": run.out_bytes_c, "d2h_bytes_per_step": run.out_bytes,
                "copy_only_ms": copies["decompress_both_ways_ms"], "fraction_of_copy_ceiling": copies["decompress_both_ways_ms"] / ms_de},
        "gpu_launches": int(launches_d), "roofline": roofline(run, t_d),
    }

    # ------------------------------------------------------------------ parity of THIS rank against the oracle
    def parity_rank():
        """chunk boundaries, a spread of chunk streams and side buffers of this rank's slab, and (rank 0)
        the all-reduced FreqTable images, against the CPU oracle.  Returns (ok, detail)."""
        from oracle import oracle as O

        host, ar = run.hp["fastq"], run.hp["arenas"]
        fs, fq = tables["ft"]
        detail = {}
        ok = True
        if rank == 0:  # tables of the sharded sample + all-reduce == tables of the whole sample in one pass
            win = host[: min(S, wl.own_bytes)]
            recs, used = O.parse_records(win)
            ofs, ofq = O.make_ft(*O.hist(win[:used], recs))
            detail["tables_match_single_pass"] = bool(np.array_equal(ofs, fs) and np.array_equal(ofq, fq))
            ok &= detail["tables_match_single_pass"]
        first = run.first_byte
        offs = O.split_chunks(host[first:], wl.R)             # the oracle's walk from this rank's first chunk
        mine = [int(ci.fastq_off) for ci in run.abs_infos] + [int(run.abs_infos[-1].fastq_off) + int(run.abs_infos[-1].total)]
        ncmp = len(mine) if last else min(len(mine), len(offs) - 1)
        detail["boundaries_match"] = bool([int(o) + first for o in offs[:ncmp]] == mine[:ncmp])
        ok &= detail["boundaries_match"]
        cod = O.Codec(fs, fq)
        nk = len(run.abs_infos)
        pick = sorted(set(list(range(min(nk, args.parity_chunks // 2))) + list(range(max(0, nk - args.parity_chunks // 2), nk))))
        bad = 0
        for kk in pick:
            ci = run.abs_infos[kk]
            sub = host[int(ci.fastq_off) : int(ci.fastq_off) + int(ci.total)]
            recs, used = O.parse_records(sub)
            enc = cod.encode_chunk(sub, recs)
            r0, n = int(ci.rec_off), int(ci.n_records)
            same = (used == sub.size and n == len(recs)
                    and np.array_equal(ar["seq"][ci.seq_off : ci.seq_off + ci.seq_len], enc["seq"])
                    and np.array_equal(ar["qual"][ci.qual_off : ci.qual_off + ci.qual_len], enc["qual"])
                    and np.array_equal(ar["readlens"][r0 : r0 + n], enc["readlens"])
                    and np.array_equal(ar["n_count"][r0 : r0 + n], enc["n_count"])
                    and np.array_equal(ar["n_pos"][ci.n_pos_off : ci.n_pos_off + ci.n_pos_len], enc["n_pos"]))
            bad += not same
        detail["chunks_checked"], detail["chunks_differ"] = len(pick), bad
        ok &= bad == 0
        return ok, detail

    parity = {"roundtrip_device": roundtrip_ok, "roundtrip_e2e": e2e_ok}
    if not args.no_parity:
        try:
            ok, detail = parity_rank()
        except Exception as e:
            ok, detail = False, {"error": repr(e)}
        parity["rank0_vs_oracle"] = detail if rank == 0 else None
        parity["this_rank_vs_oracle"] = bool(ok)
        parity["all_ranks_vs_oracle"] = bool(allsum(0.0 if ok else 1.0) == 0.0)
        parity["all_ranks_roundtrip"] = bool(allsum(0.0 if (roundtrip_ok and e2e_ok) else 1.0) == 0.0)
        if world > 1:  # the global chunk chain is gap-free: each rank starts where the previous one stopped
            ends = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
            a0 = wl.base + run.first_byte
            dist.all_gather(ends, torch.tensor([a0, a0 + run.out_bytes], dtype=torch.int64, device=dev))
            e = [(int(x[0]), int(x[1])) for x in ends]
            parity["chunk_chain_contiguous"] = bool(e[0][0] == 0 and all(e[i][1] == e[i + 1][0] for i in range(world - 1))
                                                    and e[-1][1] == wl.file_bytes)

    def gpu_fnv(O):
        """FNV-1a over (seq stream, qual stream) of every chunk in order, on the
        bytes the e2e leg fetched to the host -- same walk as fq28o_bench."""
        ar, hh = run.hp["arenas"], 1469598103934665603
        for ci in run.abs_infos:
            hh = O.fnv1a(ar["seq"][ci.seq_off : ci.seq_off + ci.seq_len], hh)
            hh = O.fnv1a(ar["qual"][ci.qual_off : ci.qual_off + ci.qual_len], hh)
        return hh

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import oracle as O

            threads = args.threads or os.cpu_count() or 1
            nb = min(wl.n, args.cpu_mb << 20)
            host = run.hp["fastq"][:nb]
            # cut at a record boundary so the sample is a valid FASTQ prefix
            offs = O.split_chunks(host, wl.R)
            host = host[: int(offs[-1])]
            res = O.bench(host, S, wl.R, threads, True)
            vc = host.size / 1e6 / (res.t_analyze_s + res.t_compress_s)
            vd = host.size / 1e6 / res.t_decompress_s
            cpu_baseline = {
                "value": vc if args.mode == "compress" else vd, "unit": "MB/s", "cores": threads, "kind": "port",
                "sample": f"leading {host.size} bytes of the rank-0 slab ({res.n_chunks} chunks), one pass",
                "compress_MBps": vc, "decompress_MBps": vd, "roundtrip_ok": bool(res.roundtrip_ok),
                "streams_match_gpu": bool(host.size == wl.n
                                          and res.seq_bytes == sum(int(ci.seq_len) for ci in run.abs_infos)
                                          and res.qual_bytes == sum(int(ci.qual_len) for ci in run.abs_infos)
                                          and res.checksum == gpu_fnv(O)),
            }
        except Exception as e:  # the oracle is a checker, never a dependency of the product arm
            cpu_baseline = {"error": repr(e)}

    # ------------------------------------------------------------------ sweep: reading sizes, chunks resident
    sweep = None
    if not args.no_sweep:
        sweep = {"note": "same slab; tables from the same -S sample; few steps per point (the -R 256 decode takes seconds)",
                 "reading_size_mb": {}, "chunks_resident": {}}
        sweep["reading_size_mb"][str(args.reading_mb)] = {
            "n_chunks_per_gpu": run.n_dec, "compress_MBps": comp["value"], "compress_e2e_MBps": comp["e2e"]["value"],
            "decompress_MBps": deco["value"], "decompress_e2e_MBps": deco["e2e"]["value"], "roundtrip": roundtrip_ok and e2e_ok}
        sweep["chunks_resident"][str(int(allsum(run.n_dec) / world))] = {
            "slabs_in_flight": 1, "fastq_MB_per_gpu": wl.own_bytes / 1e6, "decompress_MBps": deco["value"], "ms": ms_d}
        # (several ranks: the resident slabs hold R - 1 bytes of lookahead for the headline R only)
        for rmb in [int(x) for x in args.sweep_reading_mb.split(",") if x and int(x) != args.reading_mb and world == 1]:
            try:
                w2 = Workload.__new__(Workload)
                w2.__dict__.update(wl.__dict__)
                w2.R = rmb << 20
                w2.max_chunks = 2 * (w2.n // w2.R) + 8
                w2.infos = (P.ChunkInfo * w2.max_chunks)()
                r2 = Run(w2)
                r2.setup()
                heavy = rmb >= 64
                c_ms = timed(r2.compress_dev, 1 if heavy else 2, 1)[0]
                d_ms = timed(r2.decompress_dev, 1, 0 if heavy else 1)[0]
                ok2 = r2.roundtrip_dev()
                ce_ms = timed(r2.compress_e2e, 1 if heavy else 2, 1)[0]
                de_ms = timed(r2.decompress_e2e, 1, 0 if heavy else 1)[0]
                ok2 = ok2 and bool(r2.st["e2e_wrote"] == r2.out_bytes) and r2.roundtrip_e2e()
                mb2 = w2.own_bytes / 1e6
                sweep["reading_size_mb"][str(rmb)] = {
                    "n_chunks_per_gpu": r2.n_dec, "compress_MBps": mb2 / (c_ms * 1e-3), "compress_e2e_MBps": mb2 / (ce_ms * 1e-3),
                    "decompress_MBps": mb2 / (d_ms * 1e-3), "decompress_e2e_MBps": mb2 / (de_ms * 1e-3), "roundtrip": ok2,
                    "stage_ms_compress": {kk: round(v, 3) for kk, v in r2.st["t_c"].items() if v},
                    "stage_ms_decompress": {kk: round(v, 3) for kk, v in r2.st["t_d"].items() if v}}
                del r2
            except Exception as e:
                sweep["reading_size_mb"][str(rmb)] = {"error": repr(e)}
        # BASELINE config 5's shape: decompress only, thousands of chunks resident per GPU.  A slab is
        # < 4 GiB (FQ28_MAX_SLAB), so more chunks = several slabs in flight on separate handles.
        if args.resident_mb and world == 1:
            try:
                import threading

                w3 = Workload(args.reading_mb << 20, size_mb=args.resident_mb)
                r3 = Run(w3)
                r3.setup()
                r3.decompress_dev()
                ok3 = r3.roundtrip_dev()
                d_ms = timed(r3.decompress_dev, 2, 1)[0]
                sweep["chunks_resident"][str(r3.n_dec)] = {"slabs_in_flight": 1, "fastq_MB_per_gpu": w3.own_bytes / 1e6,
                                                           "decompress_MBps": w3.own_bytes / 1e6 / (d_ms * 1e-3), "ms": d_ms, "roundtrip": ok3}
                nh = max(1, args.resident_slabs)
                if nh > 1:
                    hs = [P.Handle(local) for _ in range(nh - 1)]
                    outs = [torch.empty(r3.out_bytes + 64, dtype=torch.uint8, device=dev) for _ in hs]
                    for hh in hs:
                        hh.load_tables(*tables["ft"])

                    def all_at_once():
                        ths = [threading.Thread(target=lambda hh=hh, o=o: hh.decompress_dev(r3.dec, r3.dec_infos, r3.n_dec, o.data_ptr(), r3.out_bytes))
                               for hh, o in zip(hs, outs)]
                        for t in ths:
                            t.start()
                        r3.decompress_dev()
                        for t in ths:
                            t.join()

                    all_at_once()
                    ok3 = all(bool(torch.equal(o[: r3.out_bytes], r3.d_out[: r3.out_bytes])) for o in outs)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    all_at_once()
                    torch.cuda.synchronize()
                    d_ms = (time.perf_counter() - t0) * 1e3
                    sweep["chunks_resident"][str(nh * r3.n_dec)] = {
                        "slabs_in_flight": nh, "fastq_MB_per_gpu": nh * w3.own_bytes / 1e6, "decompress_MBps": nh * w3.own_bytes / 1e6 / (d_ms * 1e-3),
                        "ms": d_ms, "roundtrip": ok3, "timer": "host wall clock around the concurrent calls (one handle and host thread per slab)",
                        "note": "the slabs are copies of one slab (same streams decoded into separate outputs)"}
                    for hh in hs:
                        hh.close()
                del r3, w3
            except Exception as e:
                sweep["chunks_resident"]["error"] = repr(e)

    if rank == 0:
        head, other = (comp, deco) if args.mode == "compress" else (deco, comp)
        cfg = workload_config(args, wl.own_bytes)
        if world > 1:
            cfg["multi_gpu"] = ("one virtual file of %d bytes; chunk boundaries from one sequential walk (each rank starts where the "
                                "previous rank's last chunk ends; that offset is the only thing exchanged, via the process group's "
                                "store); sample sharded by record range + one NCCL all-reduce of the 525 312 counters" % wl.file_bytes)
        line = {
            "metric": METRIC, "value": head["value"], "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": cfg, "mode": args.mode,
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
            "clocks": clocks, "cpu_baseline": cpu_baseline,
            "compress": {kk: v for kk, v in comp.items()}, "decompress": {kk: v for kk, v in deco.items()},
            "host_link": copies, "e2e_two_calls_in_flight": streaming, "parity": parity, "sweep": sweep,
            "stats": {"n_chunks": run.n_dec, "n_records": run.n_rec_dec, "seq_bytes": int(summ.seq_bytes),
                      "qual_bytes": int(summ.qual_bytes), "ratio_seq_qual": wl.n / max(1, int(summ.seq_bytes) + int(summ.qual_bytes))},
        }
        print(json.dumps(line), flush=True)
    h.close()
    baton.close()
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
