"""the UNMODIFIED reference stitcher (src/pcs-multicamera-optimized.cpp built as
a program with its own main(), oracle/_ref/pcs-multicamera-optimized) consumes a camera over the
reference's wire protocol -- one-byte 'Z' pull, [int32 bytes][records] reply
(src/pcs-multicamera-optimized.cpp:268-297,342-345) -- runs its own convert -> transform -> `+=`
loop and hands the stitched cloud to its viewer (`-v`).  The stub viewer (oracle/stubs) dumps what it
is given; that dump must equal oracle unpack -> transform of the payload, field for field.

* CPU: a Python fake camera serves oracle-made payloads (checks the harness and the reference binary);
* GPU: the camera is pointcloud_stitching_b200/pcs_camera_node, i.e. our library behind the
  reference's sendXYZRGBPointcloud call shape -- the drop-in claim end to end.

`-v` rather than the :9000 server path because initServerSocket() (:167-197) is declared int and has
no return statement: g++ 13 ends it in `ud2` at -O0 and lets it fall through at -O3.
The reference hard-codes port 8000 and "localhost" (:42,457)."""
import os
import socket
import subprocess
import threading
import time

import numpy as np
import pytest

from conftest import ROOT
from pointcloud_stitching_b200 import synth

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "pcs-multicamera-optimized")
NODE_BIN = os.path.join(ROOT, "pointcloud_stitching_b200", "pcs_camera_node")
W, H, FRAMES = 1280, 720, 3


def _port_free(port):
    s = socket.socket()
    try:
        s.bind(("0.0.0.0", port))
        return True
    except OSError:
        return False
    finally:
        s.close()


def _expected(restatement):
    import oracle
    cal = oracle.make_calib(W, H, translation=synth.D2C_BASELINE)
    frames = [(synth.depth_frame(W, H, 0, f), synth.color_frame(W, H, 0, f)) for f in range(FRAMES)]
    payloads = [restatement.frame(cal, z, c, 3, W * 3, synth.TF_CAMERA) for z, c in frames]
    clouds = [restatement.transform_cloud(restatement.unpack(p), synth.TF_STITCH[0]) for p in payloads]
    return frames, payloads, clouds


def _run_stitcher(n_frames, dump):
    """Run the reference stitcher with -v until its (stub) viewer has been given n_frames clouds."""
    import oracle
    assert _wait_listening_camera(), "camera never listened on :8000"
    env = dict(os.environ, PCS_STUB_VIEWER_DUMP=str(dump), PCS_STUB_VIEWER_FRAMES=str(n_frames))
    r = subprocess.run([REF_BIN, "-v"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env,
                       timeout=180)
    assert r.returncode == 0, r.stdout[-2000:]
    raw = np.fromfile(str(dump), np.uint8)
    out, off = [], 0
    for _ in range(n_frames):
        n = int(raw[off:off + 4].view(np.int32)[0])
        out.append(raw[off + 4:off + 4 + 32 * n].view(oracle.PCLPOINT))
        off += 4 + 32 * n
    return out


def _same_cloud(got, want):
    return len(got) == len(want) and all(np.array_equal(got[k], want[k]) for k in ("x", "y", "z", "w", "b", "g", "r", "a"))


def _wait_listening_camera():
    # do not connect to :8000 here (the camera accepts exactly one client); poll /proc instead
    t0 = time.time()
    while time.time() - t0 < 90:
        with open("/proc/net/tcp") as f:
            if any(line.split()[1].endswith(":1F40") and line.split()[3] == "0A" for line in f.readlines()[1:]):
                return True
        time.sleep(0.1)
    return False


needs_ref = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/pcs-multicamera-optimized not built")
needs_port = pytest.mark.skipif(not _port_free(8000), reason="port 8000 busy")


@needs_ref
@needs_port
def test_reference_stitcher_with_fake_camera(restatement, tmp_path):
    _, payloads, want = _expected(restatement)

    def camera():
        srv = socket.socket()
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind(("0.0.0.0", 8000))
        srv.listen(1)
        c, _ = srv.accept()
        f = 0
        try:
            while True:
                z = c.recv(1)
                if z != b"Z":
                    break
                p = payloads[f % FRAMES].reshape(-1).view(np.uint8).tobytes()
                c.sendall(np.int32(len(p)).tobytes() + p)
                f += 1
        except OSError:
            pass
        c.close()
        srv.close()

    th = threading.Thread(target=camera, daemon=True)
    th.start()
    got = _run_stitcher(FRAMES, tmp_path / "viewer.bin")
    for f in range(FRAMES):
        assert _same_cloud(got[f], want[f]), "frame %d" % f
    th.join(timeout=10)


@pytest.mark.gpu
@needs_ref
@needs_port
def test_reference_stitcher_with_b200_camera_node(restatement, tmp_path):
    frames, _, want = _expected(restatement)
    d, c = tmp_path / "depth.raw", tmp_path / "color.raw"
    np.stack([z for z, _ in frames]).tofile(d)
    np.stack([col for _, col in frames]).tofile(c)
    node = subprocess.Popen([NODE_BIN, "--depth", str(d), "--color", str(c), "--w", str(W), "--h", str(H),
                             "--frames", str(FRAMES), "--tx", "0.015", "--tf", "0"],
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        got = _run_stitcher(FRAMES, tmp_path / "viewer.bin")
    finally:
        try:
            out, _ = node.communicate(timeout=20)
        except subprocess.TimeoutExpired:
            node.kill()
            out, _ = node.communicate()
    for f in range(FRAMES):
        assert _same_cloud(got[f], want[f]), "frame %d\n%s" % (f, out)
    assert "frames sent" in out
