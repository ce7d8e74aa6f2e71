import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built_artifacts():
    """The in-tree library and the two command lines normally travel with the tree; build whatever is missing
    (nvcc cross-compiles without a GPU), so a bare checkout can run the suite too."""
    from dbslmm_b200 import build as b
    if not (os.path.exists(b.LIB) and os.path.exists(b.CLI) and os.path.exists(b.VALID_CLI)):
        b.build_lib()
        b.build_cli()
    from oracle import oracle
    oracle.build()
    # the unmodified reference over oracle/shim (git-ignored): rebuilt here when the reference sources are present
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")
    if not os.path.exists(ref_so) and os.path.isdir("/root/reference/scr"):
        import subprocess
        subprocess.run(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")], check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    yield


@pytest.fixture(scope="session")
def engine():
    from dbslmm_b200 import _abi
    eng = _abi.Engine(0)          # raises if the CUDA library or a B200 is missing -- no fallback
    yield eng
    eng.close()
