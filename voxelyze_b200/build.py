"""In-tree builds: the CUDA product library and (separately) the oracle checkers.

    python -m voxelyze_b200.build            # product + oracles
    python -m voxelyze_b200.build product    # libvoxelyze_b200.so only

nvcc cross-compiles for sm_100a without a GPU; the resulting .so files are git-ignored but
travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "voxelyze_b200", "csrc")
LIBDIR = os.path.join(ROOT, "voxelyze_b200", "lib")
PRODUCT_SO = os.path.join(LIBDIR, "libvoxelyze_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # the reference is built without FMA contraction (x86-64 baseline); parity needs the same
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
    "-diag-suppress", "177",
]


def _host_cxx() -> str:
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def product_sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    srcs.append(os.path.join(ROOT, "include", "voxelyze_b200.h"))
    return srcs


def build_product(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and _newer(PRODUCT_SO, product_sources()):
        return PRODUCT_SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, "-ccbin", _host_cxx(), "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           "-o", PRODUCT_SO, os.path.join(CSRC, "vx_capi.cu"), "-ldl"]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    subprocess.run(cmd, check=True)
    return PRODUCT_SO


FACADE_SO = os.path.join(LIBDIR, "libvoxelyze_facade.so")
FACADE_DIR = os.path.join(ROOT, "voxelyze_b200", "facade")


def build_facade(force: bool = False) -> str:
    """The C++ drop-in class API (CVoxelyze, CVX_*) on top of the C-ABI library."""
    src = os.path.join(FACADE_DIR, "src", "voxelyze_facade.cpp")
    src_json = os.path.join(FACADE_DIR, "src", "voxelyze_json.cpp")
    src_mesh = os.path.join(FACADE_DIR, "src", "voxelyze_mesh.cpp")
    src_solver = os.path.join(FACADE_DIR, "src", "voxelyze_solver.cpp")
    inc = os.path.join(FACADE_DIR, "include")
    deps = [src, src_json, src_mesh, src_solver, PRODUCT_SO] + [os.path.join(inc, f) for f in os.listdir(inc)] + [os.path.join(CSRC, "vx_material.hpp")]
    if not force and _newer(FACADE_SO, deps):
        return FACADE_SO
    cmd = [_host_cxx(), "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-variable", "-Wno-overloaded-virtual",
           "-I", inc, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", FACADE_SO, src, src_json, src_mesh, src_solver,
           "-L", LIBDIR, "-lvoxelyze_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    return FACADE_SO


def build_cpp_tests() -> dict:
    """tests/cpp/dropin_tests.cpp compiled twice: against the facade (runs on the GPU) and, where
    /root/reference exists, against the unmodified reference (runs on the CPU) - same source."""
    out = {}
    tdir = os.path.join(ROOT, "tests", "cpp")
    bdir = os.path.join(tdir, "_build")
    os.makedirs(bdir, exist_ok=True)
    src = os.path.join(tdir, "dropin_tests.cpp")
    inc = os.path.join(FACADE_DIR, "include")
    exe = os.path.join(bdir, "dropin_b200")
    if not _newer(exe, [src, FACADE_SO]):
        subprocess.run([_host_cxx(), "-O2", "-std=c++17", "-Wno-overloaded-virtual", "-I", inc, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                        "-o", exe, src, "-L", LIBDIR, "-lvoxelyze_facade", "-lvoxelyze_b200",
                        "-Wl,-rpath," + LIBDIR], check=True)
    out["b200"] = exe
    e2e = os.path.join(bdir, "facade_e2e")              # the headline workload through the C++ class API (bench.py's e2e.facade leg)
    e2e_src = os.path.join(ROOT, "tools", "facade_e2e.cpp")
    if not _newer(e2e, [e2e_src, FACADE_SO]):
        subprocess.run([_host_cxx(), "-O2", "-std=c++17", "-I", inc, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", e2e, e2e_src,
                        "-L", LIBDIR, "-lvoxelyze_facade", "-lvoxelyze_b200", "-Wl,-rpath," + LIBDIR], check=True)
    out["facade_e2e"] = e2e
    ref = "/root/reference"
    exe_ref = os.path.join(bdir, "dropin_ref")
    if os.path.isdir(os.path.join(ref, "src")):
        if not _newer(exe_ref, [src]):
            names = ["Voxelyze", "VX_Link", "VX_Voxel", "VX_External", "VX_Material", "VX_MaterialVoxel",
                     "VX_MaterialLink", "VX_Collision", "VX_LinearSolver", "VX_MeshRender"]
            subprocess.run([_host_cxx(), "-O3", "-std=c++11", "-w", "-DDROPIN_REFERENCE", "-I", os.path.join(ref, "include"),
                            "-o", exe_ref, src] + [os.path.join(ref, "src", n + ".cpp") for n in names], check=True)
    if os.path.exists(exe_ref):
        out["ref"] = exe_ref
    out.update(build_ref_gtests())
    return out


def build_ref_gtests() -> dict:
    """The reference's OWN gtest headers (test/tVX_Material.h, tVX_MaterialLink.h, tVX_Voxel.h, tVoxelyze.h), unmodified and
    not copied: compiled where they lie under /root/reference through per-file symbolic links (tests/cpp/_build/rt_*/reftests/
    test/) whose sibling `include` link selects the implementation under test -- the reference's headers (+ its sources) or the
    facade's.  GoogleTest is not in the image: tests/cpp/gtest_shim/gtest/gtest.h stands in.  Only where /root/reference
    exists (this container); the GPU box runs the prebuilt binaries."""
    out = {}
    tdir = os.path.join(ROOT, "tests", "cpp")
    bdir = os.path.join(tdir, "_build")
    ref = "/root/reference"
    main = os.path.join(tdir, "ref_gtests_main.cpp")
    shim = os.path.join(tdir, "gtest_shim")
    exes = {"gtests_ref": os.path.join(bdir, "ref_gtests_ref"), "gtests_b200": os.path.join(bdir, "ref_gtests_b200")}
    if os.path.isdir(os.path.join(ref, "test")):
        tests = sorted(f for f in os.listdir(os.path.join(ref, "test")) if f.startswith("t") and f.endswith(".h"))
        for variant, inc in (("ref", os.path.join(ref, "include")), ("b200", os.path.join(FACADE_DIR, "include"))):
            d = os.path.join(bdir, "rt_" + variant, "reftests")
            os.makedirs(os.path.join(d, "test"), exist_ok=True)
            for f in tests:                              # file links, not a directory link: `..` must stay inside rt_*/reftests
                link = os.path.join(d, "test", f)
                if not os.path.islink(link):
                    os.symlink(os.path.join(ref, "test", f), link)
            link = os.path.join(d, "include")
            if not os.path.islink(link):
                os.symlink(inc, link)
        deps = [main, os.path.join(shim, "gtest", "gtest.h")]
        if not _newer(exes["gtests_ref"], deps):
            names = ["Voxelyze", "VX_Link", "VX_Voxel", "VX_External", "VX_Material", "VX_MaterialVoxel",
                     "VX_MaterialLink", "VX_Collision", "VX_LinearSolver"]
            subprocess.run([_host_cxx(), "-O3", "-std=c++11", "-w", "-I", shim, "-I", os.path.join(bdir, "rt_ref"), "-I", os.path.join(ref, "include"),
                            "-o", exes["gtests_ref"], main] + [os.path.join(ref, "src", n + ".cpp") for n in names], check=True)
        inc = os.path.join(FACADE_DIR, "include")
        if not _newer(exes["gtests_b200"], deps + [FACADE_SO] + [os.path.join(inc, f) for f in os.listdir(inc)]):
            subprocess.run([_host_cxx(), "-O2", "-std=c++17", "-w", "-I", shim, "-I", os.path.join(bdir, "rt_b200"), "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                            "-o", exes["gtests_b200"], main, "-L", LIBDIR, "-lvoxelyze_facade", "-lvoxelyze_b200", "-Wl,-rpath," + LIBDIR], check=True)
    for k, exe in exes.items():
        if os.path.exists(exe):
            out[k] = exe
    return out


def build_oracles() -> None:
    """Builds the checkers (oracle port always; oracle/_ref only where /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port", "ref"], check=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "product"):
        print(build_product(force="--force" in sys.argv, verbose="-v" in sys.argv))
        print(build_facade(force="--force" in sys.argv))
    if what in ("all", "oracle"):
        build_oracles()
