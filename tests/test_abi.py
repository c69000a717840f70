"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every
symbol include/pcs_b200.h declares; no compute is attempted without a GPU."""
import ctypes as C
import os
import subprocess

import pytest

import pointcloud_stitching_b200 as pcs


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_exported():
    syms = pcs.declared_symbols()
    assert len(syms) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", pcs.LIB_PATH], capture_output=True, text=True,
                         check=True).stdout
    rows = [line.split() for line in out.splitlines() if line.strip()]
    exported = {r[-1] for r in rows}
    assert not [s for s in syms if s not in exported]
    # and nothing but the ABI leaks out of the library (weak libstdc++ template bodies aside)
    assert not [r[-1] for r in rows if r[-2] == "T" and not r[-1].startswith("pcs_b200_")]


def test_abi_version_and_status_strings():
    assert pcs.lib.pcs_b200_abi_version() == 2
    assert pcs.lib.pcs_b200_status_string(0) == b"ok"
    assert pcs.lib.pcs_b200_status_string(pcs.PCS_ERR_CUDA) == b"CUDA error"
    assert pcs.lib.pcs_b200_status_string(123) == b"ok"          # counts / byte sizes are successes


def test_struct_layouts_match_header(tmp_path):
    # sizes the C side was compiled with (include/pcs_b200.h)
    assert C.sizeof(pcs.Intrinsics) == 48
    assert C.sizeof(pcs.StreamDesc) == 2 * 48 + 9 * 4 + 3 * 4 + 4 + 4 + 4 + 16 * 4 + 4 + 4 * 4 + 4
    assert C.sizeof(pcs.Config) == 16
    assert C.sizeof(pcs.FrameJob) == 8 + 5 * 8
    # and what a C compiler makes of the header itself
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "pcs_b200.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", '
                   'sizeof(pcs_intrinsics), sizeof(pcs_stream_desc), sizeof(pcs_config), sizeof(pcs_frame_job), '
                   'sizeof(pcs_ipc_handle), sizeof(pcs_shard_peers)); return 0; }\n')
    exe = tmp_path / "sizes"
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-I" + os.path.join(pcs.ROOT, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True, text=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(pcs.Intrinsics), C.sizeof(pcs.StreamDesc), C.sizeof(pcs.Config), C.sizeof(pcs.FrameJob),
                   C.sizeof(pcs.IpcHandle), C.sizeof(pcs.ShardPeers)]


def test_library_contains_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", pcs.LIB_PATH], capture_output=True, text=True).stdout
    archs = {tok for line in out.splitlines() for tok in line.replace(".", " ").split() if tok.startswith("sm_")}
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(pcs.PcsError) as e:
        pcs.Context()
    assert e.value.status == pcs.PCS_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_bad_arguments_are_errors_not_exits():
    assert pcs.lib.pcs_b200_create(None, None) == pcs.PCS_ERR_INVALID
    cfg = pcs.Config(0, 0, 0, 0)
    h = C.c_void_p()
    assert pcs.lib.pcs_b200_create(C.byref(cfg), C.byref(h)) == pcs.PCS_ERR_INVALID
    assert pcs.lib.pcs_b200_set_stream(None, 0, None) == pcs.PCS_ERR_INVALID
    assert pcs.lib.pcs_b200_send_xyzrgb(None, 0, None, None, None, 0) == pcs.PCS_ERR_INVALID
    assert b"null" in pcs.lib.pcs_b200_last_error(None)


def test_product_does_not_touch_the_oracle():
    # the package (and the library sources) must not import, link or execute oracle/
    for dirpath, _, files in os.walk(os.path.dirname(pcs.__file__)):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "pcs_oracle" not in text, f
    deps = subprocess.run(["ldd", pcs.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps
