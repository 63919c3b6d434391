"""Builds and runs tests/cpp/test_facade.cpp: the reference's own unit tests
(test/fse_sequence_test.cpp, test/fse_quality_test.cpp, test/workspace_test.cpp,
test/archive_test.cpp meta part) transliterated onto the C++ facade
fqcomp28_b200/host/fqcomp28_gpu.hpp, which keeps the reference's class names on
top of the C ABI."""
import os
import subprocess

import pytest

from conftest import DATA, ROOT


def build_facade_test(tmp_path):
    exe = str(tmp_path / "test_facade")
    lib_dir = os.path.join(ROOT, "fqcomp28_b200")
    subprocess.check_call(
        ["g++", "-std=c++20", "-O1", "-Wall", "-Wextra", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_facade.cpp"),
         "-L", lib_dir, "-lfq28", f"-Wl,-rpath,{lib_dir}"]
    )
    return exe


def test_facade_compiles(tmp_path):
    """CPU: the facade header is valid C++20 and links against libfq28.so."""
    build_facade_test(tmp_path)


@pytest.mark.gpu
def test_reference_unit_tests_on_facade(tmp_path):
    exe = build_facade_test(tmp_path)
    res = subprocess.run([exe, DATA], capture_output=True, text=True, timeout=300)
    print(res.stdout, res.stderr)
    assert res.returncode == 0, res.stdout + res.stderr
    assert " 0 failed" in res.stdout
