import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA = os.path.join(ROOT, "tests", "data")
FIXTURES = ["SRR065390_1_first5", "SRR065390_sub_1", "SRR065390_sub_2", "without_ns"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_fixture(name: str) -> np.ndarray:
    """The reference's own test inputs (test/data/*.fastq), copied verbatim as
    parity fixtures (public SRA reads, not source code)."""
    return np.fromfile(os.path.join(DATA, name + ".fastq"), dtype=np.uint8)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.lib()
    return O
