"""Golden digests from the independent Python model (tests/golden/fse_model.py,
written from SURVEY.md Appendix A / section 8(a), not from the C oracle).

tests/golden/golden.json is what tests/golden/make_golden.py writes.  Here:
  * the committed digests equal SURVEY.md Appendix C on the reference's fixtures;
  * the model regenerates them (so the file cannot drift from the script);
  * the C oracle reproduces every digest, including a ~3 MB synthetic fixture whose
    contexts reach table logs 10-11 -- the logs of the 1 GB bench, outside the
    range the libzstd pin (tests/test_fse_vs_libzstd.py) covers;
  * `-m gpu`: the CUDA path reproduces the same digests through the C ABI.
"""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_fixture

GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)


def sha(a) -> str:
    b = a if isinstance(a, (bytes, bytearray)) else np.ascontiguousarray(a).tobytes()
    return hashlib.sha256(b).hexdigest()[:16]


@pytest.fixture(scope="module")
def G():
    return json.load(open(os.path.join(GOLDEN, "golden.json")))


def log_hist(logs):
    h = {}
    for t in logs:
        h[str(int(t))] = h.get(str(int(t)), 0) + 1
    return h


def test_committed_goldens_are_appendix_c(G):
    from test_oracle_golden import GOLD, LOGS, SIDE

    for name, g in GOLD.items():
        e = G["fixtures"][name]
        assert (e["n_records"], e["seq_len"], e["seq"], e["ft_seq"], e["qual_len"], e["qual"], e["ft_qual"]) == g
        assert e["seq_logs"] == {str(k): v for k, v in LOGS[name][0].items()}
        assert e["qual_logs"] == {str(k): v for k, v in LOGS[name][1].items()}
        s = SIDE[name]
        assert (e["n_count_len"], e["n_count"], e["n_pos_len"]) == (s[0], s[1], s[2])
        if s[3]:
            assert e["n_pos"] == s[3]
        if s[4]:
            assert e["readlens"] == s[4]


@pytest.mark.parametrize("name", ["SRR065390_1_first5", "without_ns"])
def test_model_regenerates_the_goldens(G, name):
    import fse_model as M

    recs = M.parse(load_fixture(name).tobytes())
    fts, ftq = M.freq_tables(recs)
    enc = M.encode_chunk(recs, fts, ftq)
    e = G["fixtures"][name]
    assert (sha(enc["seq"]), sha(enc["qual"]), sha(M.ft_image(fts)), sha(M.ft_image(ftq))) == (e["seq"], e["qual"], e["ft_seq"], e["ft_qual"])
    assert (sha(enc["n_count"]), sha(enc["n_pos"]), sha(enc["readlens"])) == (e["n_count"], e["n_pos"], e["readlens"])
    ds, dq = M.decode_chunk(enc, fts, ftq, len(recs))
    assert ds == [r[1] for r in recs] and dq == [r[2] for r in recs]


def synthetic():
    import synth

    return synth.illumina(0, 9000, seed=30).numpy()


def check_entry(e, enc, n_records):
    assert n_records == e["n_records"]
    assert (enc["seq"].size, sha(enc["seq"])) == (e["seq_len"], e["seq"])
    assert (enc["qual"].size, sha(enc["qual"])) == (e["qual_len"], e["qual"])
    assert sha(np.asarray(enc["readlens"], dtype="<u2")) == e["readlens"]
    assert sha(np.asarray(enc["n_count"], dtype="<u2")) == e["n_count"]
    assert (2 * np.asarray(enc["n_pos"]).size, sha(np.asarray(enc["n_pos"], dtype="<u2"))) == (e["n_pos_len"], e["n_pos"])


def test_oracle_reproduces_the_log_10_11_fixture(G, oracle):
    O = oracle
    s = G["synthetic"]["illumina_3mb"]
    d = synthetic()
    assert (d.size, sha(d)) == (s["bytes"], s["sha"]), "synthetic generator changed: regenerate tests/golden/golden.json"
    R, S = s["reading_size"], s["sample_size"]
    sample = d[: int(O.split_chunks(d, S)[1])]
    recs, _ = O.parse_records(sample)
    fs, fq = O.make_ft(*O.hist(sample, recs))
    assert (sha(fs), sha(fq)) == (s["ft_seq"], s["ft_qual"])
    assert log_hist(O.ft_logs(fs)) == s["seq_logs"] and log_hist(O.ft_logs(fq)) == s["qual_logs"]
    assert {"10", "11"} <= set(s["seq_logs"]) and "11" in s["qual_logs"]
    offs = O.split_chunks(d, R)
    assert [int(o) for o in offs] == s["chunk_offsets"]
    cod = O.Codec(fs, fq)
    for k, e in enumerate(s["chunks"]):
        sub = d[int(offs[k]) : int(offs[k + 1])]
        r, _ = O.parse_records(sub)
        check_entry(e, cod.encode_chunk(sub, r), len(r))


@pytest.mark.gpu
def test_gpu_reproduces_the_goldens(G, oracle):
    """the CUDA path against the model's digests directly (not via the oracle)"""
    import fqcomp28_b200 as P

    h = P.Handle(0)
    for name, e in G["fixtures"].items():
        d = load_fixture(name)
        fs, fq = h.build_tables(*h.hist(d))
        assert (sha(fs), sha(fq)) == (e["ft_seq"], e["ft_qual"]), name
        infos, summ, ar = h.compress(d, 256 << 20, eof=True)
        assert int(summ.n_chunks) == 1
        ci = infos[0]
        enc = {"seq": ar["seq"][ci.seq_off : ci.seq_off + ci.seq_len], "qual": ar["qual"][ci.qual_off : ci.qual_off + ci.qual_len],
               "readlens": ar["readlens"][: ci.n_records], "n_count": ar["n_count"][: ci.n_records], "n_pos": ar["n_pos"][: ci.n_pos_len]}
        check_entry(e, enc, int(ci.n_records))
    s = G["synthetic"]["illumina_3mb"]
    d = synthetic()
    fs = np.zeros(P.capi.FT_SEQ_BYTES, np.uint8)
    fq = np.zeros(P.capi.FT_QUAL_BYTES, np.uint8)
    infos, summ, ar = h.compress(d, s["reading_size"], eof=True, sample_bytes=s["sample_size"], ft_out=(fs, fq))
    assert (sha(fs), sha(fq)) == (s["ft_seq"], s["ft_qual"])
    assert int(summ.n_chunks) == len(s["chunks"])
    for k, e in enumerate(s["chunks"]):
        ci = infos[k]
        assert int(ci.fastq_off) == s["chunk_offsets"][k]
        enc = {"seq": ar["seq"][ci.seq_off : ci.seq_off + ci.seq_len], "qual": ar["qual"][ci.qual_off : ci.qual_off + ci.qual_len],
               "readlens": ar["readlens"][ci.rec_off : ci.rec_off + ci.n_records],
               "n_count": ar["n_count"][ci.rec_off : ci.rec_off + ci.n_records],
               "n_pos": ar["n_pos"][ci.n_pos_off : ci.n_pos_off + ci.n_pos_len]}
        check_entry(e, enc, int(ci.n_records))
    h.close()
