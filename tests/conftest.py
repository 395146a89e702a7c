import os
import shutil
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: full BASELINE size, tens of seconds (still part of the default run)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    g.build()
    return g


def _load_oracle(path, fresh=False):
    import tinyrenderder_b200 as trb
    if fresh:  # the reference keeps its counters in process-wide statics: load a private copy
        d = tempfile.mkdtemp(prefix="trb_oracle_")
        p2 = os.path.join(d, os.path.basename(path))
        shutil.copy(path, p2)
        path = p2
    return trb.Api(path, "orc")


@pytest.fixture(scope="session")
def port_api(built):
    return _load_oracle(os.path.join(ROOT, "oracle", "libtrb_port.so"))


@pytest.fixture(scope="session")
def ref_api(built):
    p = os.path.join(ROOT, "oracle", "_ref", "libtrb_ref.so")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/libtrb_ref.so not built (needs /root/reference)")
    return _load_oracle(p)


@pytest.fixture()
def fresh_ref_api(built):
    p = os.path.join(ROOT, "oracle", "_ref", "libtrb_ref.so")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/libtrb_ref.so not built (needs /root/reference)")
    return _load_oracle(p, fresh=True)


@pytest.fixture(scope="session")
def cuda_api(built):
    import tinyrenderder_b200 as trb
    return trb.load_cuda()  # raises when the CUDA library is missing: no silent fallback
