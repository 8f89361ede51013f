import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The plain-C restatement, built on demand (gcc only; travels prebuilt to the GPU box)."""
    import ora
    if not os.path.exists(ora.ORACLE_SO):
        ora.build_oracles()
    return ora.OracleLib()


@pytest.fixture(scope="session")
def ref_lib():
    """The unmodified reference sources compiled here; skipped where neither /root/reference nor
    a prebuilt oracle/_ref exists."""
    import ora
    if not os.path.exists(ora.REF_SO):
        if os.path.isdir(ora.REFERENCE_ROOT):
            ora.build_oracles()
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return ora.RefLib()


@pytest.fixture(scope="session")
def gpu_lib():
    from crux_b200 import gpu
    L = gpu.lib()
    if L.pg_device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    return gpu
