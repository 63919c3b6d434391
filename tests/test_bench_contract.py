"""The bench lines committed under profiles/ carry every key the measurement contract asks for
(DESIGN.md section 6), and their numbers are self-consistent.  CPU only:
it reads the recorded JSON lines, it does not run the bench."""
import json
import os

import pytest

from conftest import ROOT

PROFILES = os.path.join(ROOT, "profiles")


def _line(name):
    with open(os.path.join(PROFILES, name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["bench_r02.json", "bench_r02_hiseq.json", "bench_r02_ont.json"])
def test_single_gpu_line(name):
    j = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "cpu_baseline"):
        assert k in j, k
    assert j["n_gpus"] == 1 and j["unit"] == "MB/s" and j["dtype"] == "u8" and j["higher_is_better"] is True
    assert j["warmup"] >= 3 and "workload" in j["config"] and j["vs_baseline"] is None
    # value = bytes / device time of the timed steps
    mb = j["config"]["fastq_bytes_per_gpu"] / 1e6
    assert abs(j["value"] - mb / (j["ms_per_step"] * 1e-3)) / j["value"] < 1e-6
    e = j["e2e"]
    assert e["h2d_bytes_per_step"] >= j["config"]["fastq_bytes_per_gpu"] and e["d2h_bytes_per_step"] > 0
    assert e["value"] < j["value"], "the host-buffer leg cannot beat the device-resident one"
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    c = j["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["streams_match_gpu"] is True
    assert j["gpu_launches"] > 0
    assert not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    p = j["parity"]
    assert p["roundtrip_device"] and p["roundtrip_e2e"] and p["all_ranks_vs_oracle"] and p["all_ranks_roundtrip"]
    assert p["rank0_vs_oracle"]["chunks_differ"] == 0 and p["rank0_vs_oracle"]["boundaries_match"]


def test_reference_arm_line():
    j = _line("bench_r02_reference.json")
    assert j["impl"] == "reference" and j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert abs(j["e2e"]["value"] - j["value"]) < 1e-6
    ours = _line("bench_r02.json")
    assert j["metric"] == ours["metric"] and j["unit"] == ours["unit"]


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_gpu_lines(n):
    j = _line(f"bench_r02_n{n}.json")
    assert j["n_gpus"] == n and j["scaling"] == "weak"
    p = j["parity"]
    assert p["all_ranks_vs_oracle"] and p["all_ranks_roundtrip"] and p["chunk_chain_contiguous"]
    one = _line("bench_r02.json")
    assert j["value"] > 0.8 * n * one["value"], "weak scaling of the device-resident compress"
    assert j["decompress"]["value"] > 0.9 * n * one["decompress"]["value"]
    chain = j["compress"]["cut_chain_ms"]
    assert len(chain) == n and chain[0]["wait_for_cut"] < 0.1
    assert all(chain[r]["wait_for_cut"] <= chain[r + 1]["wait_for_cut"] + 0.3 for r in range(n - 1)), "rank r waits for r walks"
