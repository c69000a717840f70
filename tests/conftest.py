import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# the product library and the oracle are build artefacts (git-ignored): make sure they exist.  A host
# without nvcc (or make) can still run the suite against prebuilt artefacts that travelled with the tree.
import __graft_entry__  # noqa: E402

try:
    __graft_entry__.build()
except Exception as e:  # noqa: BLE001
    _built = [os.path.join(ROOT, "pointcloud_stitching_b200", "libpcs_b200.so"),
              os.path.join(ROOT, "oracle", "_build", "libpcs_oracle.so")]
    if not all(os.path.exists(p) for p in _built):
        raise
    print("conftest: build() failed (%r); using the prebuilt libraries" % (e,), file=sys.stderr)


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a CPU-only host skips the gpu-marked tests instead of failing them."""
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run with gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def restatement():
    import oracle
    return oracle.restatement()


@pytest.fixture(scope="session")
def ref_camera():
    import oracle
    r = oracle.ref_camera()
    if r is None:
        pytest.skip("oracle/_ref not built (needs /root/reference): golden fixtures cover this")
    return r


@pytest.fixture(scope="session")
def ref_client():
    import oracle
    r = oracle.ref_client()
    if r is None:
        pytest.skip("oracle/_ref not built")
    return r


@pytest.fixture(scope="session")
def ref_optimized():
    import oracle
    r = oracle.ref_optimized()
    if r is None:
        pytest.skip("oracle/_ref not built")
    return r


def roundtrip_inputs():
    """The inputs of tests/golden/roundtrip_all_int16.npz (make_golden.py)."""
    allv = np.zeros((65536, 5), np.int16)
    allv[:, 0] = np.arange(-32768, 32768)
    allv[:, 1] = allv[::-1, 0]
    allv[:, 2] = np.roll(allv[:, 0], 12345)
    allv[:, 3] = np.arange(65536).astype(np.uint16).view(np.int16)
    allv[:, 4] = (np.arange(65536) % 251).astype(np.int16)
    return allv
