"""Drop-in proof for the C++ class API: tests/cpp/dropin_tests.cpp is ONE source that only uses
the reference's public API (CVoxelyze, CVX_Material, CVX_Voxel, CVX_Link, CVX_External ...).

* compiled against the unmodified reference (CPU, only where /root/reference exists at build time)
  it must pass with the reference's own gtest expectations -> the test source is a faithful
  restatement of test/tVoxelyze.h;
* compiled against voxelyze_b200/facade (+ the CUDA library) the very same binary logic must pass
  on the B200 -> existing callers drop in."""
import os
import subprocess

import pytest

from voxelyze_b200 import build


@pytest.fixture(scope="session")
def binaries(built):
    build.build_facade()
    return build.build_cpp_tests()


def _run(exe, *args):
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    return r.stdout


def test_reference_passes_the_dropin_source(binaries):
    if "ref" not in binaries:
        pytest.skip("reference sources not available at build time")
    out = _run(binaries["ref"])
    assert "0 failures" in out and "21 tests" in out


def test_facade_host_only_classes(binaries):
    """Material / external classes work with no device at all (stand-alone objects, tVX_*.h)."""
    out = _run(binaries["b200"], "--host-only")
    assert "0 failures" in out


@pytest.mark.gpu
def test_facade_passes_the_dropin_source_on_gpu(binaries):
    out = _run(binaries["b200"])
    assert "0 failures" in out and "21 tests" in out
