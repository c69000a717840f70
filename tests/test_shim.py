"""The C++ host side: include/pcs_b200_shim.hpp keeps the reference's call shapes.
CPU: it compiles as C++11 against rs2-like frame types and links the C ABI.
GPU: the binary runs and is bit-exact against the oracle (tests/cpp/shim_main.cpp)."""
import os
import subprocess

import pytest

import pointcloud_stitching_b200 as pcs
from conftest import ROOT

BIN = os.path.join(ROOT, "tests", "cpp", "_build", "shim_main")


def _build():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    cmd = ["/usr/bin/g++", "-std=c++11", "-O1", "-Wall", "-pthread", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "oracle", "stubs"), "-I" + os.path.join(ROOT, "oracle"),
           os.path.join(ROOT, "tests", "cpp", "shim_main.cpp"), "-o", BIN,
           pcs.LIB_PATH, os.path.join(ROOT, "oracle", "_build", "libpcs_oracle.so"),
           "-Wl,-rpath," + os.path.dirname(pcs.LIB_PATH), "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_shim_compiles_and_links_as_cxx11():
    _build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_shim_binary_bit_exact_on_gpu():
    _build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("OK")


MG_BIN = os.path.join(ROOT, "tests", "cpp", "_build", "multigpu_pull_main")


def _build_multigpu():
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cuda = os.path.dirname(os.path.dirname(os.path.realpath(nvcc)))
    os.makedirs(os.path.dirname(MG_BIN), exist_ok=True)
    cmd = ["/usr/bin/g++", "-std=c++11", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "oracle"), "-I" + os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "cpp", "multigpu_pull_main.cpp"), "-o", MG_BIN,
           pcs.LIB_PATH, os.path.join(ROOT, "oracle", "_build", "libpcs_oracle.so"),
           "-L" + os.path.join(cuda, "lib64"), "-lcudart",
           "-Wl,-rpath," + os.path.dirname(pcs.LIB_PATH), "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build"),
           "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_multigpu_cxx_host_compiles_and_links():
    """tests/cpp/multigpu_pull_main.cpp: a single-process C++ host, one pcs_ctx per GPU, pull exchange
    over pcs_b200_enable_peer + z-slab sharded voxel merge, all through the C ABI."""
    _build_multigpu()
    assert os.path.exists(MG_BIN)


@pytest.mark.gpu
def test_multigpu_cxx_host_bit_exact_on_gpu():
    """Runs on two GPUs when the box has them (prints SKIP and passes on one)."""
    _build_multigpu()
    r = subprocess.run([MG_BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("OK")


MP_BIN = os.path.join(ROOT, "tests", "cpp", "_build", "multiproc_pull_main")


def _build_multiproc():
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cuda = os.path.dirname(os.path.dirname(os.path.realpath(nvcc)))
    os.makedirs(os.path.dirname(MP_BIN), exist_ok=True)
    cmd = ["/usr/bin/g++", "-std=c++11", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "oracle"), "-I" + os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "cpp", "multiproc_pull_main.cpp"), "-o", MP_BIN,
           pcs.LIB_PATH, os.path.join(ROOT, "oracle", "_build", "libpcs_oracle.so"),
           "-L" + os.path.join(cuda, "lib64"), "-lcudart",
           "-Wl,-rpath," + os.path.dirname(pcs.LIB_PATH), "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build"),
           "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_multiprocess_cxx_host_compiles_and_links():
    """tests/cpp/multiproc_pull_main.cpp: one PROCESS per GPU (fork + socketpair, no torch / NCCL), the peers'
    frames mapped with pcs_b200_ipc_export / pcs_b200_ipc_open, pull exchange through the C ABI."""
    _build_multiproc()
    assert os.path.exists(MP_BIN)


@pytest.mark.gpu
def test_multiprocess_cxx_host_bit_exact_on_gpu():
    """Runs on two GPUs when the box has them (prints SKIP and passes on one)."""
    _build_multiproc()
    r = subprocess.run([MP_BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("OK"), r.stdout


def test_division_by_constant_is_correctly_rounded():
    """The pipelined kernel divides by the colour width with a precomputed reciprocal and two
    Markstein corrections; check it against IEEE division over 2^-12..2^24 for common widths."""
    src = os.path.join(ROOT, "tests", "cpp", "constdiv_check.c")
    exe = os.path.join(os.path.dirname(BIN), "constdiv_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["/usr/bin/gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", src, "-o", exe, "-lm"],
                   check=True, capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    assert " 0 mismatches" in r.stdout


def test_guard_bound_dominates_the_distance_between_the_two_tap_chains():
    """Rotated depth->colour calibrations: the pipelined kernel accepts the tap of a cheap projection chain where it
    stays eps away from every integer (pcs_guard.h).  Sample the exact and the cheap chain on random rigs and check
    that their distance stays below the bound the host derives (largest observed fraction well under 1)."""
    src = os.path.join(ROOT, "tests", "cpp", "guard_check.cpp")
    exe = os.path.join(os.path.dirname(BIN), "guard_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["/usr/bin/g++", "-O2", "-mfma", "-ffp-contract=off", "-I",
                    os.path.join(ROOT, "pointcloud_stitching_b200", "csrc"), src, "-o", exe],
                   check=True, capture_output=True, text=True)
    r = subprocess.run([exe, "60", "300000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    frac = float(r.stdout.strip().splitlines()[-1].split("largest fraction of the bound")[1].split()[0])
    assert 0.1 < frac < 0.7, r.stdout       # neither violated nor vacuous


def test_segment_windows_hold_every_tap_of_a_rotated_rig():
    """pcs_guard.h: the colour rows staged per depth row and 128-px segment must hold the tap of every pixel at every depth
    beyond the near limit (8 M random pixels over four geometries, the rig of `bench.py --tex rotated`), and stay as short
    as DESIGN.md says (3-4 rows per depth row)."""
    src = os.path.join(ROOT, "tests", "cpp", "guard_window_main.cpp")
    exe = os.path.join(os.path.dirname(BIN), "guard_window")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++11", "-I", os.path.join(ROOT, "pointcloud_stitching_b200", "csrc"), src, "-o", exe],
                   check=True, capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert r.stdout.count("misses 0 of") == 4, r.stdout
