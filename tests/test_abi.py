"""CPU-side checks of the C-ABI boundary: the library loads, exports every
symbol include/fq28.h declares, struct layouts match the header, and the
product never reaches into oracle/."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "fq28.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fq28_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import fqcomp28_b200

    L = fqcomp28_b200.load()
    declared = header_symbols()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(L, name), f"libfq28.so does not export {name}"
    assert sorted(fqcomp28_b200.SYMBOLS) == declared


def test_struct_layouts_match_header(tmp_path):
    """sizeof/offsetof as the C compiler sees them == the ctypes mirrors."""
    import fqcomp28_b200 as P

    prog = tmp_path / "layout.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "fq28.h"\n'
        "int main(void){\n"
        ' printf("%zu %zu %zu %zu\\n", sizeof(fq28_chunk_info), sizeof(fq28_enc_arenas), sizeof(fq28_enc_summary), sizeof(fq28_dec_arenas));\n'
        ' printf("%zu %zu %zu %zu\\n", offsetof(fq28_chunk_info, seq_len), offsetof(fq28_chunk_info, n_pos_off), offsetof(fq28_dec_arenas, hdr_lens), offsetof(fq28_dec_arenas, n_records));\n'
        " return 0; }\n"
    )
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    got = [int(x) for x in out]
    want = [
        C.sizeof(P.ChunkInfo), C.sizeof(P.EncArenas), C.sizeof(P.EncSummary), C.sizeof(P.DecArenas),
        P.ChunkInfo.seq_len.offset, P.ChunkInfo.n_pos_off.offset, P.DecArenas.hdr_lens.offset, P.DecArenas.n_records.offset,
    ]
    assert got == want


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device fq28_create must fail."""
    import torch

    import fqcomp28_b200 as P

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(P.Fq28Error):
        P.Handle(0)


def test_bounds_match_reference_formulas():
    """Workspace::compressBoundSequence/Quality, src/workspace.h:21-35."""
    import fqcomp28_b200 as P

    L = P.load()
    assert L.fq28_bound_seq(10) == 1024 * 256
    assert L.fq28_bound_seq(100000) == 100000 // 4 + 1024
    assert L.fq28_bound_qual(10) == 1024 * 8192
    assert L.fq28_bound_qual(100_000_000) == 100_000_000 * 7 // 8 + 1024


def test_product_does_not_touch_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "fqcomp28_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")) or fn == "Makefile":
                txt = open(os.path.join(dp, fn)).read()
                hits = [l for l in txt.splitlines() if re.search(r"(import|include|from|-l|dlopen).*oracle", l)]
                assert not hits, f"{fn} references oracle/: {hits}"
    out = subprocess.check_output(["ldd", os.path.join(pkg, "libfq28.so")]).decode()
    assert "oracle" not in out
