"""CPU-only checks of the drop-in boundary: every symbol declared in include/voxelyze_b200.h
is exported by the product library and by both oracles; no compute is issued."""
import os
import re
import subprocess

from voxelyze_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "voxelyze_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vx_[a-z0-9_]+)\s*\(", text)))


def exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_header_declares_the_binding_surface():
    syms = declared_symbols()
    assert "vx_step" in syms and "vx_create" in syms and len(syms) >= 30


def test_product_exports_every_symbol(built):
    have = exported(capi.PRODUCT_SO)
    missing = [s for s in declared_symbols() if s not in have]
    assert not missing, missing


def test_oracles_export_every_symbol(built):
    for path in (capi.ORACLE_SO, capi.REF_SO):
        if not os.path.exists(path):
            continue
        have = exported(path)
        missing = [s for s in declared_symbols() if s not in have]
        assert not missing, (path, missing)


def test_ctypes_binding_matches_header(built):
    lib = capi.load_oracle()
    assert sorted(lib.symbols) == declared_symbols()
    assert lib.lib.vx_abi_version() == 4
    assert lib.backend == "oracle-port"


def test_product_library_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "--list-elf", capi.PRODUCT_SO], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_does_not_reference_the_oracle(built):
    """The product path must not link or load anything under oracle/."""
    out = subprocess.run(["ldd", capi.PRODUCT_SO], capture_output=True, text=True).stdout
    assert "oracle" not in out and "vxref" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "voxelyze_b200", "csrc")):
        for f in files:
            assert "oracle" not in open(os.path.join(root, f)).read().replace("Not used by oracle/", "").lower(), f
