"""marker survey -> runtime-loadable transforms.  The only known-answer data
in the reference repository is calibration/pcs4.csv: marker rows :2-11 (copied to
tests/golden/pcs4_markers.csv) and the two 4x4 results its authors wrote underneath (:14-24)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from pointcloud_stitching_b200 import calibration

# calibration/pcs4.csv:14-24, as printed by the reference script (8 decimals)
DEXTRO = [[-0.99574067, 0.02631655, -0.08836260, 0.02300000],
          [0.09219821, 0.28421870, -0.95431610, 2.05300000],
          [0.00000000, -0.95839823, -0.28543446, 1.84600000],
          [0.00000000, 0.00000000, 0.00000000, 1.00000000]]
LEVO = [[0.99056815, -0.04332334, 0.12999166, -0.51300000],
        [-0.13694491, -0.34463012, 0.92869595, -1.86300000],
        [0.00456483, -0.93773833, -0.34731253, 1.90300000],
        [0.00000000, 0.00000000, 0.00000000, 1.00000000]]


def test_known_answer_pcs4():
    tfs = calibration.transforms_from_csv(os.path.join(GOLDEN, "pcs4_markers.csv"))
    assert sorted(tfs) == ["DEXTRO", "LEVO"]
    assert np.allclose(tfs["DEXTRO"], DEXTRO, atol=5e-9, rtol=0)
    assert np.allclose(tfs["LEVO"], LEVO, atol=5e-9, rtol=0)
    for t in tfs.values():                      # rigid: orthonormal rotation, det +1
        r = t[:3, :3]
        assert np.allclose(r.T @ r, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(r), 1.0)


def test_reference_label_convention_and_roundtrip(tmp_path):
    # A + A1..A4 (calibration/camera_alignment.py:66-69); a missing marker is the reference's KeyError
    rows = ["1,0,0,0,", "2,1.0,2.0,0.5,A", "3,0.9,2.1,0.6,A1", "4,1.1,2.1,0.6,A2", "5,1.1,1.9,0.4,A3", "6,0.9,1.9,0.4,A4",
            "7,5,5,5,B", "8,5,5,6,B1"]
    p = tmp_path / "rig.csv"
    p.write_text("\n".join(rows) + "\n")
    tfs = calibration.transforms_from_csv(str(p))
    assert list(tfs) == ["A"] and np.allclose(tfs["A"][:3, 3], [1.0, 2.0, 0.5])
    with pytest.raises(KeyError, match="B2"):
        calibration.camera_transform(calibration.load_markers(str(p)), "B")
    out = tmp_path / "rig.json"
    calibration.save_transforms(str(out), tfs)
    back = calibration.load_transforms(str(out))
    assert back["A"].dtype == np.float32 and back["A"].shape == (16,)
    assert np.allclose(back["A"].reshape(4, 4), tfs["A"], atol=1e-6)


def test_cxx_host_loads_the_same_transforms(tmp_path):
    """include/pcs_b200_shim.hpp::load_transform (used by pcs_camera_node --tf-file/--camera) reads the
    JSON calibration.py writes and yields the same float32 values as load_transforms."""
    import subprocess
    from conftest import ROOT
    tfs = calibration.transforms_from_csv(os.path.join(GOLDEN, "pcs4_markers.csv"))
    tfs["E-7"] = np.array([[1e-9, -2.5e3, 3, 4], [5, 6e-12, 7, -8], [9, 10, 11, 12.125], [0, 0, 0, 1]])
    js = tmp_path / "rig.json"
    calibration.save_transforms(str(js), tfs)
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "tf_json_main")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["/usr/bin/g++", "-std=c++11", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "tf_json_main.cpp"), "-o", exe],
                   check=True, capture_output=True, text=True)
    want = calibration.load_transforms(str(js))
    for name in tfs:
        r = subprocess.run([exe, str(js), name], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout
        got = np.array([int(x, 16) for x in r.stdout.split()], np.uint32).view(np.float32)
        assert np.array_equal(got, want[name]), name
    assert subprocess.run([exe, str(js), "NOPE"], capture_output=True, text=True).returncode == 1
    # files other tools write: one line, 3 x 4 (last row implied), a short matrix followed by more numbers, a camera
    # name that also occurs as a value
    other = tmp_path / "other.json"
    other.write_text('{"note":"LEVO","short":[[1,2,3],[4,5,6]],"n":[7,8,9,10,11,12,13,14,15,16],'
                     '"LEVO":[[1,0,0,0.5],[0,1,0,-2e0],[0,0,1,3]],\n"flat" : [1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16],'
                     '"text":["a"]}')
    r = subprocess.run([exe, str(other), "LEVO"], capture_output=True, text=True)
    got = np.array([int(x, 16) for x in r.stdout.split()], np.uint32).view(np.float32)
    assert r.returncode == 0 and np.array_equal(got, np.float32([1, 0, 0, .5, 0, 1, 0, -2, 0, 0, 1, 3, 0, 0, 0, 1]))
    r = subprocess.run([exe, str(other), "flat"], capture_output=True, text=True)
    got = np.array([int(x, 16) for x in r.stdout.split()], np.uint32).view(np.float32)
    assert r.returncode == 0 and np.array_equal(got, np.arange(1, 17, dtype=np.float32))
    for name in ("short", "n", "text", "note"):       # 6 numbers, 10 numbers, not numbers, not a matrix
        assert subprocess.run([exe, str(other), name], capture_output=True, text=True).returncode == 1, name
    assert subprocess.run([exe, str(tmp_path / "missing.json"), "LEVO"], capture_output=True).returncode == 1
