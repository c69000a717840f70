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


# ---- the rest of SURVEY s8(f) rank 1: push-mode camera, and the stitcher tier serving a viewer on its own port --------
STITCH_BIN = os.path.join(ROOT, "pointcloud_stitching_b200", "pcs_stitch_node")


def _free_ports(k):
    socks = [socket.socket() for _ in range(k)]
    for s in socks:
        s.bind(("127.0.0.1", 0))
    ports = [s.getsockname()[1] for s in socks]
    for s in socks:
        s.close()
    return ports


def _consecutive_free_ports(k):
    for base in range(18000, 30000, 16):
        if all(_port_free(base + i) for i in range(k)):
            return base
    raise RuntimeError("no free port range")


def _connect(port, timeout=90):
    t0 = time.time()
    while time.time() - t0 < timeout:
        try:
            return socket.create_connection(("127.0.0.1", port), timeout=60)
        except OSError:
            time.sleep(0.1)
    raise RuntimeError("nobody listens on :%d" % port)


def _read_exact(sock, n):
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(min(1 << 20, n - len(buf)))
        if not chunk:
            raise EOFError("peer closed after %d of %d bytes" % (len(buf), n))
        buf += chunk
    return bytes(buf)


def _read_frame(sock):
    n = int(np.frombuffer(_read_exact(sock, 4), np.int32)[0])
    return np.frombuffer(_read_exact(sock, n), np.int16)


def _camera_files(tmp_path, cam, w, h, frames):
    d, c = tmp_path / ("depth%d.raw" % cam), tmp_path / ("color%d.raw" % cam)
    np.stack([synth.depth_frame(w, h, cam, f) for f in range(frames)]).tofile(d)
    np.stack([synth.color_frame(w, h, cam, f) for f in range(frames)]).tofile(c)
    return str(d), str(c)


def _finish(proc, what):
    try:
        out, _ = proc.communicate(timeout=30)
    except subprocess.TimeoutExpired:
        proc.kill()
        out, _ = proc.communicate()
    return out


@pytest.mark.gpu
def test_push_mode_camera_node(restatement, tmp_path):
    """`pcs-camera-optimized -f X -s` (src/pcs-camera-optimized.cpp:264-302): every frame is sent without waiting
    for a pull.  The reader below never writes a byte -- what pcs-multicamera-client's readCloud does before
    its first pull (src/pcs-multicamera-client.cpp:363-371)."""
    import oracle
    w, h, frames, loops = 848, 480, 2, 5
    port, = _free_ports(1)
    d, c = _camera_files(tmp_path, 3, w, h, frames)
    node = subprocess.Popen([NODE_BIN, "--depth", d, "--color", c, "--w", str(w), "--h", str(h), "--frames", str(frames),
                             "--tx", "0.015", "--tf", "0", "--port", str(port), "--push", "--loops", str(loops)],
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    cal = oracle.make_calib(w, h, translation=synth.D2C_BASELINE)
    want = [restatement.frame(cal, synth.depth_frame(w, h, 3, f), synth.color_frame(w, h, 3, f), 3, w * 3, synth.TF_CAMERA)
            for f in range(frames)]
    try:
        s = _connect(port)
        for f in range(loops):
            assert np.array_equal(_read_frame(s), want[f % frames].reshape(-1)), "frame %d" % f
        assert s.recv(1) == b""            # the node closes after --loops frames
        s.close()
    finally:
        out = _finish(node, "camera")
    assert "%d frames sent" % loops in out, out


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["push_raw_d1", "pull_raw_d3", "pull_pcl_d2"])
def test_stitch_node_serves_the_viewer(restatement, tmp_path, mode):
    """The stitcher tier end to end: two GPU camera nodes -> pcs_stitch_node (readCloud threads, GPU stitch) -> a viewer
    that pulls [int32][stitched] with 'Z' on the server port (src/pcs-multicamera-client.cpp:398-403;
    src/pcs-multicamera-optimized.cpp:299-313).  Both flow-control modes of SURVEY F11."""
    import oracle
    push, pcl, d = mode.startswith("push"), "pcl" in mode, int(mode[-1])
    geoms = [(848, 480), (640, 480)]
    frames, served = 2, 3
    base = _consecutive_free_ports(3)
    cam_port, viewer_port = base, base + 2
    nodes, want_pay = [], []
    for cam, (w, h) in enumerate(geoms):
        dp, cp = _camera_files(tmp_path, cam, w, h, frames)
        args = [NODE_BIN, "--depth", dp, "--color", cp, "--w", str(w), "--h", str(h), "--frames", str(frames), "--tx", "0.015",
                "--tf", "0", "--port", str(cam_port + cam)]
        if push:
            args += ["--push", "--loops", str(served)]
        nodes.append(subprocess.Popen(args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        cal = oracle.make_calib(w, h, translation=synth.D2C_BASELINE)
        want_pay.append([restatement.frame(cal, synth.depth_frame(w, h, cam, f), synth.color_frame(w, h, cam, f), 3, w * 3,
                                           synth.TF_CAMERA) for f in range(frames)])
    sargs = [STITCH_BIN, "--cameras", "2", "--camera-port", str(cam_port), "--viewer-port", str(viewer_port), "--downsample",
             str(d), "--frames", str(served)]
    if not push:
        sargs.append("--prime-pull")
    tfs = [synth.TF_STITCH[2], synth.TF_STITCH[5]]
    if pcl:
        import json
        tf_file = tmp_path / "rig.json"
        tf_file.write_text(json.dumps({"left": np.asarray(tfs[0], np.float64).reshape(4, 4).tolist(),
                                       "right": np.asarray(tfs[1], np.float64).reshape(4, 4).tolist()}))
        sargs += ["--pcl", "--tf-file", str(tf_file), "--names", "left,right"]
    stitcher = subprocess.Popen(sargs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    problem = None
    try:
        v = _connect(viewer_port)
        for f in range(served):
            v.sendall(b"Z")
            got = _read_frame(v)
            pays = [want_pay[cam][f % frames] for cam in range(2)]
            if pcl:
                want = restatement.pcl_stitch(pays, tfs, d)
            else:
                want = restatement.concat([p.reshape(-1) for p in pays], d)
            if not np.array_equal(got, np.frombuffer(want[4:].tobytes(), np.int16)):
                problem = "stitched frame %d differs (%s)" % (f, mode)
                break
        v.close()
    except (EOFError, OSError, RuntimeError) as e:
        problem = repr(e)
    finally:
        sout = _finish(stitcher, "stitcher")
        outs = [_finish(n, "camera") for n in nodes]
    log = "\n--- stitcher ---\n" + sout + "".join("\n--- camera %d ---\n%s" % (i, o) for i, o in enumerate(outs))
    assert problem is None, problem + log
    assert "%d stitched frames served" % served in sout, log
