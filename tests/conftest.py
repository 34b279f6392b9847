import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# Every test that takes `orc` runs twice: against oracle/liboracle.so (the restatement, always there) and against
# oracle/_ref/libmodref.so (the UNMODIFIED reference objects behind the same harness API) wherever that was
# built - the GPU box receives the prebuilt file with the snapshot, so the parity tests compare the CUDA path
# with the reference itself, not only with its restatement.
@pytest.fixture(scope="session", params=["port", "reference"])
def orc(request):
    import harness
    if request.param == "port":
        return harness.oracle()
    r = harness.reference()
    if r is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    return r


@pytest.fixture(scope="session")
def ref():
    import harness
    r = harness.reference()
    if r is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    return r
