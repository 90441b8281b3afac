import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle_py import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def gpu():
    """The product library.  No fallback: a missing .so or device is an error, not a skip."""
    from quilt_b200 import api

    lib = api.GpuLib()
    if lib.device_count() < 1:
        raise RuntimeError("no CUDA device visible: -m gpu tests must run on the GPU box")
    return lib


@pytest.fixture(scope="session")
def small_world():
    """K_full = 600 haplotypes, 3200 common SNPs (T = 100), 9600 SNPs overall (T_all = 300)."""
    from quilt_b200 import synth

    return synth.make_world(20260117, K_full=600, nSNPs=3200, region_bp=300_000, all_snps_factor=3)


@pytest.fixture(scope="session")
def small_reads(small_world):
    from quilt_b200 import synth

    return synth.make_sample_reads(small_world, 7, coverage=1.0, region_bp=300_000)
