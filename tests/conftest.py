import gzip
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """The in-tree build (library + tools).  Built on demand when missing."""
    lib = os.path.join(ROOT, "rabbitvar_b200", "librvgpu.so")
    tools = [os.path.join(ROOT, "build", t) for t in ("synthgen", "rv_dump", "rabbitvar_b200")]
    if not os.path.exists(lib) or not all(os.path.exists(t) for t in tools):
        import __graft_entry__ as g
        g.build()
    return ROOT


@pytest.fixture(scope="session")
def ref_tools():
    """Paths of the compiled reference (oracle/_ref) or None when it was not built/shipped."""
    d = os.path.join(ROOT, "oracle", "_ref")
    paths = {"ref_dump": os.path.join(d, "ref_dump"), "RabbitVar": os.path.join(d, "RabbitVar")}
    return paths if all(os.path.exists(p) for p in paths.values()) else None


def golden_path(name, kind):
    return os.path.join(ROOT, "tests", "golden", f"{name}.{kind}.gz")


def unpack_golden(name, kind, dest):
    with gzip.open(golden_path(name, kind), "rb") as f, open(dest, "wb") as o:
        o.write(f.read())
    return dest


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)
