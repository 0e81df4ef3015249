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
    assert "0 failures" in out and "27 tests" in out


def test_facade_host_only_classes(binaries):
    """Material / external classes work with no device at all (stand-alone objects, tVX_*.h)."""
    out = _run(binaries["b200"], "--host-only")
    assert "0 failures" in out


@pytest.mark.gpu
def test_facade_passes_the_dropin_source_on_gpu(binaries):
    out = _run(binaries["b200"])
    assert "0 failures" in out and "31 tests" in out


def test_vxl_json_files_are_interchangeable(binaries, tmp_path):
    """*.vxl.json written by the reference loads into the same model in the facade and vice versa (no device
    needed: loading only builds the host-side model).  Digest = every material, voxel and external field."""
    if "ref" not in binaries:
        pytest.skip("reference sources not available at build time")
    digests = {}
    for writer in ("ref", "b200"):
        path = str(tmp_path / f"{writer}.vxl.json")
        _run(binaries[writer], "--json-save", path)
        for reader in ("ref", "b200"):
            digests[writer, reader] = _run(binaries[reader], "--json-digest", path)
    first = digests["ref", "ref"]
    assert "materials 2 voxels 20" in first and "fixed 111111" in first and "'soft'" in first
    for key, d in digests.items():
        assert d == first, key


def test_facade_loads_the_unterminated_files_the_reference_writes(binaries, tmp_path):
    """Without externals the reference never closes the root object (src/Voxelyze.cpp:211,237); such files must load."""
    path = tmp_path / "open.vxl.json"
    path.write_text('{\n "voxelSize": 0.005,\n "materials": [ {"youngsModulus": 1000000.0, "density": 1000.0}, {"youngsModulus": 1000000, "density": 2000.0} ],\n'
                    ' "voxels": [0,0,0,0, 1,0,0,1, 2,0,0,0]\n')
    out = _run(binaries["b200"], "--json-digest", str(path))
    assert "voxelSize 0.0050000000000000001 materials 2 voxels 3" in out
    assert "mat 0 \'\' rgba -1 -1 -1 -1 linear 1 E 1000000 fail -1 rho 1000" in out
    # an integer literal is not a "double" for the reference's reader (IsDouble): that material has no valid model
    # and keeps the cleared defaults E = 1, rho = 1 (src/VX_Material.cpp:54-73,125-139)
    assert "mat 1 \'\' rgba -1 -1 -1 -1 linear 1 E 1 fail -1 rho 1 " in out
    assert "vox 1 at 1 0 0 mat 1" in out
    if "ref" in binaries:                  # the same file, closed: the unmodified reference reads it to the same model
        closed = tmp_path / "closed.vxl.json"
        closed.write_text(path.read_text() + "}\n")
        assert _run(binaries["ref"], "--json-digest", str(closed)) == _run(binaries["b200"], "--json-digest", str(closed)) == out


@pytest.mark.gpu
def test_mid_run_edits_keep_the_state_of_untouched_links(binaries):
    """A yielded beam, one voxel's material swapped mid-run (only ITS links restart, src/Voxelyze.cpp:485-498), the load
    removed, collisions enabled mid-run: the facade (device re-layouts carrying voxel and link state) ends where the
    unmodified reference ends.  The residual bend is plastic memory of the untouched links."""
    if "ref" not in binaries:
        pytest.skip("reference sources not available at build time")
    ref = _run(binaries["ref"], "--edit-scenario").splitlines()
    got = _run(binaries["b200"], "--edit-scenario").splitlines()
    assert ref[0].split()[:5] == got[0].split()[:5] and int(ref[0].split()[4]) > 0          # same number of yielded links
    assert ref[1] == got[1]
    import numpy as np
    a = np.array([[float(x) for x in l.split()[2:]] for l in ref[2:]])
    b = np.array([[float(x) for x in l.split()[2:]] for l in got[2:]])
    nominal = np.array([[i, j, k] for k in range(2) for j in range(2) for i in range(8)]) * 0.001
    scale = np.abs(a[:32] - nominal).max()          # rows 32.. : a few voxels after setVoxelSize(0.0015) and 200 more steps
    assert scale > 1e-5                                                                    # a residual (plastic) deflection remains
    assert np.abs(a - b).max() <= 1e-7 * scale, np.abs(a - b).max() / scale


def test_json_loader_survives_malformed_files(binaries, tmp_path):
    """Truncated / wrong-typed / out-of-range input never crashes the loader; what is valid is kept
    (like CVoxelyze::loadJSON, which reports success whenever the file could be opened, src/Voxelyze.cpp:61-76)."""
    samples = ['', '{', '{"voxelSize": 0.01', '{"voxelSize": 0.01, "materials": [', '[1,2,3]', '{"voxelSize": "x", "materials": 5}',
               '{"voxelSize": 0.01, "materials": [{"youngsModulus": 1e6}], "voxels": [0,0,0']
    for k, text in enumerate(samples):
        p = tmp_path / f"bad{k}.json"
        p.write_text(text)
        out = _run(binaries["b200"], "--json-digest", str(p))
        assert "voxels 0" in out, (text, out)
    p = tmp_path / "partly.json"
    p.write_text('{"voxelSize": 0.01, "materials": [{"strainData": [0.0, 0.1, 0.2], "stressData": [0.0, 1e5, 1.5e5]}], '
                 '"voxels": [0,0,0,0, 1,0,0,0, 2,0,0,7], '
                 '"externals": [{"voxelIndices": [0, 99], "fixed": [true,true,true,true,true,true]}, {"fixed": [true]}]}')
    out = _run(binaries["b200"], "--json-digest", str(p))
    assert "materials 1 voxels 2" in out and "linear 0" in out and "vox 0 at 0 0 0 mat 0 fixed 111111" in out


# ---- the reference's own gtest files, unmodified (test/VoxelyzeUnitTests.cpp:2-8) -----------------
STALE = {"CVoxelyze.poissonsSmall", "CVoxelyze.deformableMaterialPossions"}     # golden values the reference itself misses


def _gtest_results(exe, *args):
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=1200, cwd=os.path.dirname(exe))     # the tests dump traces into the cwd
    res = {}
    for line in r.stdout.splitlines():
        if line.startswith("[") and "]" in line:
            res[line.split("]", 1)[1].strip()] = "OK" in line.split("]", 1)[0]
    assert res, r.stdout[-2000:] + r.stderr[-2000:]
    return res, r.stdout


def test_reference_passes_49_of_its_51_own_gtests(binaries):
    """Pins the shim: the unmodified reference against its own test headers gives what SURVEY.md section 4 measured."""
    if "gtests_ref" not in binaries:
        pytest.skip("reference sources not available at build time")
    res, _ = _gtest_results(binaries["gtests_ref"])
    assert len(res) == 51 and {k for k, ok in res.items() if not ok} == STALE


def test_facade_passes_the_references_host_only_gtests(binaries):
    """tVX_Material.h, tVX_MaterialLink.h, tVX_Voxel.h against the facade: stand-alone objects, no device needed."""
    if "gtests_b200" not in binaries:
        pytest.skip("reference gtests were not built (needs /root/reference at build time)")
    for suite, n in (("CVX_Material", 23), ("CVX_Voxel", 2)):
        res, out = _gtest_results(binaries["gtests_b200"], suite)
        assert len(res) == n and all(res.values()), out[-3000:]


@pytest.mark.gpu
def test_facade_passes_the_references_own_gtests_on_gpu(binaries):
    """All 51 tests of the reference's own test headers, compiled unmodified against the facade, on the B200: the same 49
    pass and the same two stale golden values fail as on the unmodified reference."""
    if "gtests_b200" not in binaries:
        pytest.skip("reference gtests were not built (needs /root/reference at build time)")
    res, out = _gtest_results(binaries["gtests_b200"])
    assert len(res) == 51, out[-3000:]
    assert {k for k, ok in res.items() if not ok} == STALE, out[-6000:]


@pytest.mark.gpu
def test_mesh_obj_file_equals_the_references(binaries, tmp_path):
    """CVX_MeshRender: the OBJ file of a stepped model written through the facade (device mesh kernels) equals
    the one the unmodified reference writes -- same vertex numbering, same faces, same 6-digit coordinates."""
    if "ref" not in binaries:
        pytest.skip("reference sources not available at build time")
    files = {}
    for which in ("ref", "b200"):
        path = str(tmp_path / f"{which}.obj")
        _run(binaries[which], "--mesh-obj", path)
        files[which] = open(path).read().splitlines()
    assert len(files["ref"]) > 50 and files["ref"][0].startswith("# OBJ")
    assert files["ref"] == files["b200"]


@pytest.mark.gpu
def test_poissons_ratio_toggled_through_the_material_handle_mid_run(binaries):
    """setPoissonsRatio() after the first steps (and back to zero later) through the facade follows the reference."""
    if "ref" not in binaries:
        pytest.skip("reference sources not available at build time")
    out = {w: _run(binaries[w], "--poisson-scenario").splitlines() for w in ("ref", "b200")}
    assert out["ref"][0].startswith("ok 1") and out["b200"][0].split()[:2] == ["ok", "1"]
    ref = [[float(x) for x in l.split()[2:]] for l in out["ref"][1:]]
    got = [[float(x) for x in l.split()[2:]] for l in out["b200"][1:]]
    assert len(ref) == len(got) == 24
    scale = max(abs(c) for r in ref for c in r)
    assert max(abs(a - b) for r, g in zip(ref, got) for a, b in zip(r, g)) <= 1e-6 * scale


@pytest.mark.gpu
def test_unmodified_callers_reach_several_devices_through_vx_devices(binaries):
    """VX_DEVICES makes every CVoxelyze of an unmodified caller run slabbed (three slabs on device 0 here) where the model can be
    cut, and moves it to one device where a call needs that (collisions, Poisson materials, stateInfo, the mesh, the solve):
    the whole drop-in source and the reference's own gtests pass unchanged."""
    env = dict(os.environ, VX_DEVICES="0,0,0")
    r = subprocess.run([binaries["b200"]], capture_output=True, text=True, timeout=1200, env=env)
    assert r.returncode == 0 and "0 failures" in r.stdout and "31 tests" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]
