"""Deterministic synthetic FASTQ generators (bench / test infrastructure).

Counter-based: every random draw is a 32-bit hash of (seed, stream, record
index, position), computed with integer tensor ops, so a slice of records
[first, first+n) is bit-identical whether it is generated on the CPU or on a
GPU, in one call or in several.  Not part of the
product path.

Profiles
  illumina : 150 bp single-end, header `@SIM28.<i> FC28:<lane>:<tile>:<x>:<y> length=<L>`,
             reads = substrings (random strand) of a GC-41% pseudo-genome with
             quality-dependent substitutions, 4-bin NovaSeq qualities {# , : F}
             as a first-order Markov chain with P(F->F)=0.97 and a 3' decay,
             0.1% of reads carry a run of 1-20 N with quality '#'.
  hiseq    : same reads, 41-level qualities (Q2..Q41) with a smooth random walk.
  ont      : variable length log-normal(ln 8000, 0.9) clipped to [1000, 50000],
             block-correlated qualities mean Q18 sd 7 clipped to [1, 50].
"""
from __future__ import annotations

import math

import torch

M32 = 0xFFFFFFFF
GENOME_LEN = 1 << 24


def _mix(x: torch.Tensor) -> torch.Tensor:
    """lowbias32 on int64 tensors holding uint32 values."""
    x = x & M32
    x = ((x ^ (x >> 16)) * 0x7FEB352D) & M32
    x = ((x ^ (x >> 15)) * 0x846CA68B) & M32
    return x ^ (x >> 16)


def _rnd(seed: int, stream: int, rec: torch.Tensor, pos=0) -> torch.Tensor:
    """uint32 hash of (seed, stream, rec, pos) as int64 tensor (broadcasting)."""
    a = _mix(rec * 0x9E3779B1 + (seed * 0x85EBCA6B + stream * 0xC2B2AE35))
    return _mix(a ^ _mix((pos + 0x27D4EB2F) * 0x165667B1 if isinstance(pos, int) else (pos + 0x27D4EB2F) * 0x165667B1))


def _u01(h: torch.Tensor) -> torch.Tensor:
    return h.to(torch.float64) / 4294967296.0


_genome_cache: dict = {}


def _genome(seed: int, device) -> torch.Tensor:
    key = (seed, str(device))
    if key not in _genome_cache:
        idx = torch.arange(GENOME_LEN, dtype=torch.int64, device=device)
        u = _rnd(seed, 101, idx) % 100
        # A 29.5 C 20.5 G 20.5 T 29.5 (41% GC), codes A0 C1 G2 T3
        g = (u >= 30).to(torch.int64) + (u >= 50).to(torch.int64) + (u >= 71).to(torch.int64)
        _genome_cache[key] = g.to(torch.uint8)
    return _genome_cache[key]


def _decimal(vals: torch.Tensor, width: int) -> torch.Tensor:
    """[n, width] uint8 of decimal digits, left-aligned, 0-padded on the right
    (0 = absent byte, removed by the final compaction)."""
    n = vals.numel()
    digs = torch.zeros((n, width), dtype=torch.int64, device=vals.device)
    v = vals.clone()
    nd = torch.ones_like(vals)
    t = vals.clone()
    for _ in range(width - 1):
        t = t // 10
        nd = nd + (t > 0).to(torch.int64)
    for k in range(width):
        # digit k (from the left) of a number with nd digits
        p = nd - 1 - k
        valid = p >= 0
        d = torch.where(valid, (v // (10 ** p.clamp(min=0))) % 10, torch.zeros_like(v))
        digs[:, k] = torch.where(valid, d + 48, torch.zeros_like(d))
    return digs.to(torch.uint8)


def _assemble(parts: list[torch.Tensor]) -> torch.Tensor:
    rows = torch.cat(parts, dim=1)
    flat = rows.reshape(-1)
    return flat[flat != 0]


def _lit(s: str, n: int, device) -> torch.Tensor:
    return torch.tensor(list(s.encode()), dtype=torch.uint8, device=device).unsqueeze(0).expand(n, -1)


_ACGT = (65, 67, 71, 84)


def _reads(seed: int, rec: torch.Tensor, L: int, qual: torch.Tensor, device) -> torch.Tensor:
    """[n, L] uint8 bases: genome substrings, random strand, substitution errors
    with p = 10^(-Q/10), N runs on 0.1% of reads."""
    n = rec.numel()
    g = _genome(seed, device)
    start = _rnd(seed, 1, rec) % (GENOME_LEN - L)
    pos = torch.arange(L, dtype=torch.int64, device=device).unsqueeze(0)
    rev = (_rnd(seed, 2, rec) & 1).unsqueeze(1)
    gi = start.unsqueeze(1) + torch.where(rev.bool(), L - 1 - pos, pos)
    b = g[gi].to(torch.int64)
    b = torch.where(rev.bool(), 3 - b, b)
    # substitution errors
    perr = torch.pow(10.0, -(qual.to(torch.float64)) / 10.0)
    u = _u01(_rnd(seed, 3, rec.unsqueeze(1), pos))
    sub = 1 + (_rnd(seed, 4, rec.unsqueeze(1), pos) % 3)
    b = torch.where(u < perr, (b + sub) & 3, b)
    lut = torch.tensor(_ACGT, dtype=torch.uint8, device=device)
    return lut[b]


def _n_runs(seed: int, rec: torch.Tensor, L: int, device):
    """mask [n, L] of positions turned into N (0.1% of reads, run 1..20)."""
    has = (_rnd(seed, 5, rec) % 1000) == 0
    run = 1 + (_rnd(seed, 6, rec) % 20)
    st = _rnd(seed, 7, rec) % L
    pos = torch.arange(L, dtype=torch.int64, device=device).unsqueeze(0)
    return has.unsqueeze(1) & (pos >= st.unsqueeze(1)) & (pos < (st + run).unsqueeze(1))


def _qual_novaseq(seed: int, rec: torch.Tensor, L: int, device) -> torch.Tensor:
    """[n, L] int64 Phred values from {2, 11, 25, 37}: first-order Markov chain."""
    n = rec.numel()
    levels = torch.tensor([2, 11, 25, 37], dtype=torch.int64, device=device)
    state = torch.full((n,), 3, dtype=torch.int64, device=device)
    # first position: mostly F
    u0 = _rnd(seed, 8, rec) % 1000
    state = torch.where(u0 < 30, torch.zeros_like(state), state)
    state = torch.where((u0 >= 30) & (u0 < 60), torch.full_like(state, 2), state)
    out = torch.empty((n, L), dtype=torch.int64, device=device)
    for i in range(L):
        out[:, i] = levels[state]
        u = _rnd(seed, 9, rec, i) % 100000
        decay = int(1500 * i / L)  # 3' decay: leaving F gets likelier
        stayF = 97000 - decay
        nxt = state.clone()
        # from F(3)
        f = state == 3
        nxt = torch.where(f & (u >= stayF), torch.full_like(state, 2), nxt)
        nxt = torch.where(f & (u >= stayF + (100000 - stayF) * 6 // 10), torch.full_like(state, 1), nxt)
        nxt = torch.where(f & (u >= stayF + (100000 - stayF) * 9 // 10), torch.zeros_like(state), nxt)
        # from ':'(2): 55% back to F, 30% stay, 10% ',', 5% '#'
        c = state == 2
        nxt = torch.where(c, torch.full_like(state, 3), nxt)
        nxt = torch.where(c & (u >= 55000), torch.full_like(state, 2), nxt)
        nxt = torch.where(c & (u >= 85000), torch.full_like(state, 1), nxt)
        nxt = torch.where(c & (u >= 95000), torch.zeros_like(state), nxt)
        # from ','(1): 40% F, 25% ':', 25% stay, 10% '#'
        c = state == 1
        nxt = torch.where(c, torch.full_like(state, 3), nxt)
        nxt = torch.where(c & (u >= 40000), torch.full_like(state, 2), nxt)
        nxt = torch.where(c & (u >= 65000), torch.full_like(state, 1), nxt)
        nxt = torch.where(c & (u >= 90000), torch.zeros_like(state), nxt)
        # from '#'(0): 50% stay, 20% ',', 15% ':', 15% F
        c = state == 0
        nxt = torch.where(c, torch.zeros_like(state), nxt)
        nxt = torch.where(c & (u >= 50000), torch.full_like(state, 1), nxt)
        nxt = torch.where(c & (u >= 70000), torch.full_like(state, 2), nxt)
        nxt = torch.where(c & (u >= 85000), torch.full_like(state, 3), nxt)
        state = nxt
    return out


def _qual_hiseq(seed: int, rec: torch.Tensor, L: int, device) -> torch.Tensor:
    """[n, L] Phred 2..41: bounded random walk starting high, drifting down."""
    n = rec.numel()
    q = 34 + (_rnd(seed, 10, rec) % 8)
    out = torch.empty((n, L), dtype=torch.int64, device=device)
    for i in range(L):
        out[:, i] = q
        u = _rnd(seed, 11, rec, i) % 100
        step = torch.zeros_like(q)
        step = torch.where(u < 22, torch.full_like(q, -1), step)
        step = torch.where(u < 6, torch.full_like(q, -4), step)
        step = torch.where(u >= 82, torch.ones_like(q), step)
        step = torch.where(u >= 97, torch.full_like(q, 3), step)
        q = (q + step).clamp(2, 41)
    return out


def illumina(first_record: int, n_records: int, seed: int = 30, read_len: int = 150,
             profile: str = "novaseq", device="cpu") -> torch.Tensor:
    """uint8 tensor with records [first_record, first_record + n_records)."""
    dev = torch.device(device)
    pieces = []
    step = 1 << 18  # bound temporary memory
    for lo in range(0, n_records, step):
        n = min(step, n_records - lo)
        rec = torch.arange(first_record + lo, first_record + lo + n, dtype=torch.int64, device=dev)
        q = _qual_novaseq(seed, rec, read_len, dev) if profile == "novaseq" else _qual_hiseq(seed, rec, read_len, dev)
        nmask = _n_runs(seed, rec, read_len, dev)
        q = torch.where(nmask, torch.full_like(q, 2), q)
        bases = _reads(seed, rec, read_len, q, dev)
        bases = torch.where(nmask, torch.full_like(bases, 78), bases)
        qch = (q + 33).to(torch.uint8)
        lane = 1 + (_rnd(seed, 12, rec) % 8)
        tile = 1101 + (_rnd(seed, 13, rec) % 1578)
        x = 1000 + (_rnd(seed, 14, rec) % 31001)
        y = 1000 + (_rnd(seed, 15, rec) % 31001)
        parts = [
            _lit("@SIM28.", n, dev), _decimal(rec + 1, 12), _lit(" FC28:", n, dev), _decimal(lane, 1),
            _lit(":", n, dev), _decimal(tile, 4), _lit(":", n, dev), _decimal(x, 5), _lit(":", n, dev),
            _decimal(y, 5), _lit(f" length={read_len}\n", n, dev), bases, _lit("\n+\n", n, dev), qch, _lit("\n", n, dev),
        ]
        pieces.append(_assemble(parts))
    return torch.cat(pieces) if pieces else torch.zeros(0, dtype=torch.uint8, device=dev)


def illumina_bytes(n_bytes: int, seed: int = 30, profile: str = "novaseq", device="cpu",
                   first_record: int = 0, read_len: int = 150):
    """Generates whole records until at least n_bytes are produced, then keeps
    the longest record-aligned prefix <= n_bytes.  -> (uint8 tensor, n_records)."""
    per = read_len * 2 + 52
    n_rec = n_bytes // per + 8
    data = illumina(first_record, n_rec, seed, read_len, profile, device)
    while data.numel() < n_bytes:
        more = illumina(first_record + n_rec, max(8, n_rec // 16), seed, read_len, profile, device)
        n_rec += max(8, n_rec // 16)
        data = torch.cat([data, more])
    # cut at the last record end <= n_bytes
    nl = torch.nonzero(data[:n_bytes] == 10).flatten()
    n_lines = (nl.numel() // 4) * 4
    end = int(nl[n_lines - 1].item()) + 1 if n_lines else 0
    return data[:end].contiguous(), n_lines // 4


def ont(first_record: int, n_records: int, seed: int = 32, device="cpu",
        min_len: int = 1000, max_len: int = 50000) -> torch.Tensor:
    """Variable-length long reads with block-correlated qualities."""
    dev = torch.device(device)
    rec = torch.arange(first_record, first_record + n_records, dtype=torch.int64, device=dev)
    # log-normal(ln 8000, 0.9) via Box-Muller on two hashes
    u1 = (_u01(_rnd(seed, 20, rec)) * 0.999998 + 1e-6)
    u2 = _u01(_rnd(seed, 21, rec))
    z = torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2 * math.pi * u2)
    L = torch.exp(math.log(8000.0) + 0.9 * z).to(torch.int64).clamp(min_len, max_len)
    tot = int(L.sum().item())
    rid = torch.repeat_interleave(torch.arange(n_records, device=dev), L)
    starts = torch.cumsum(L, 0) - L
    pos = torch.arange(tot, dtype=torch.int64, device=dev) - starts[rid]
    grec = rec[rid]
    g = _genome(seed, dev)
    gstart = _rnd(seed, 22, rec) % (GENOME_LEN - max_len)
    b = g[gstart[rid] + pos].to(torch.int64)
    # qualities: per-64-base block mean around Q18 (sd 7) + small per-base noise
    blk = pos >> 6
    zb = (_u01(_rnd(seed, 23, grec, blk)) + _u01(_rnd(seed, 24, grec, blk)) + _u01(_rnd(seed, 25, grec, blk)) - 1.5) * 2.0
    q = (18.0 + 7.0 * zb + (_u01(_rnd(seed, 26, grec, pos)) - 0.5) * 6.0).round().to(torch.int64).clamp(1, 50)
    perr = torch.pow(10.0, -q.to(torch.float64) / 10.0)
    sub = 1 + (_rnd(seed, 27, grec, pos) % 3)
    b = torch.where(_u01(_rnd(seed, 28, grec, pos)) < perr, (b + sub) & 3, b)
    lut = torch.tensor(_ACGT, dtype=torch.uint8, device=dev)
    bases = lut[b]
    qch = (q + 33).to(torch.uint8)
    # assemble: header lines are short, build them row-wise, then interleave
    hdr_rows = torch.cat([
        _lit("@SIM28ONT.", n_records, dev), _decimal(rec + 1, 12), _lit(" ch=", n_records, dev),
        _decimal(1 + (_rnd(seed, 29, rec) % 512), 3), _lit(" length=", n_records, dev), _decimal(L, 5),
    ], dim=1)
    hdr_len = (hdr_rows != 0).sum(1)
    rec_bytes = hdr_len + 2 * L + 5
    rec_start = torch.cumsum(rec_bytes, 0) - rec_bytes
    out = torch.empty(int(rec_bytes.sum().item()), dtype=torch.uint8, device=dev)
    hflat = hdr_rows.reshape(-1)
    keep = hflat != 0
    hrid = torch.arange(n_records, device=dev).unsqueeze(1).expand_as(hdr_rows).reshape(-1)[keep]
    hcum = torch.cumsum(hdr_len, 0) - hdr_len
    hpos = torch.arange(int(hdr_len.sum().item()), device=dev) - hcum[hrid]
    out[rec_start[hrid] + hpos] = hflat[keep]
    s0 = rec_start + hdr_len
    out[s0] = 10
    out[s0[rid] + 1 + pos] = bases
    out[s0 + 1 + L] = 10
    out[s0 + 2 + L] = 43
    out[s0 + 3 + L] = 10
    out[s0[rid] + 4 + L[rid] + pos] = qch
    out[s0 + 4 + 2 * L] = 10
    return out


def random_fastq(n_records: int, seed: int = 1, min_len: int = 3, max_len: int = 300, n_rate: float = 0.02,
                 qual_levels: int = 40):
    """Small adversarial FASTQ for parity tests (numpy, CPU): random lengths from
    min_len, N anywhere, long N runs, every quality up to qual_levels-1."""
    import numpy as np

    rng = np.random.default_rng(seed)
    out = bytearray()
    for i in range(n_records):
        L = int(rng.integers(min_len, max_len + 1))
        seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), L)
        mask = rng.random(L) < n_rate
        if rng.random() < 0.1 and L > 8:
            a = int(rng.integers(0, L - 4))
            mask[a : a + int(rng.integers(1, L - a))] = True
        seq = np.where(mask, ord("N"), seq).astype(np.uint8)
        q = np.clip(np.cumsum(rng.integers(-3, 4, L)) + 20, 0, qual_levels - 1).astype(np.uint8) + 33
        out += b"@r%d some:fields:%d/x\n" % (i, int(rng.integers(0, 99999)))
        out += seq.tobytes() + b"\n+\n" + q.tobytes() + b"\n"
    return np.frombuffer(bytes(out), dtype=np.uint8)
