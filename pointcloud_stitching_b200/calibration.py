"""Camera registration: theodolite marker survey -> per-camera 4x4 rigid transforms, as a
runtime-loadable file.

The reference computes these with calibration/camera_alignment.py and the result is pasted
into the C++ sources by hand (`transform[k] << ...`, src/pcs-multicamera-optimized.cpp:417-455,
`tf_mat`, src/pcs-camera-optimized.cpp:64-67), i.e. a recompile per rig.  Same formula here
(calibration/camera_alignment.py:22-51), but the transforms go to a JSON file that feeds
``pcs_stream_desc.tf`` at run time.

Markers around camera C:      1-----2          X = (m1 - m2) + (m4 - m3)
                              ---C---          Y' = (m4 - m1) + (m3 - m2)
                              4-----3          Z = X x Y',  Y = Z x X   (columns = X, Y, Z, normalised)
The translation column is the surveyed position of the camera itself (:66-69).
"""
from __future__ import annotations

import csv
import json

import numpy as np


def load_markers(path):
    """CSV rows ``pointNo, X, Y, Z, Label`` -> {label: xyz}; rows without a label are skipped."""
    out = {}
    with open(path, newline="") as f:
        for row in csv.reader(f):
            row = [c.strip() for c in row]
            if len(row) < 5 or not row[4]:
                continue
            try:
                out[row[4]] = np.array([float(row[1]), float(row[2]), float(row[3])])
            except ValueError:
                continue
    return out


def camera_rotation(markers, prefix):
    try:
        m1, m2, m3, m4 = (markers["%s%d" % (prefix, k)] for k in (1, 2, 3, 4))
    except KeyError as e:
        raise KeyError("Label %s not found" % e.args[0])
    x = (m1 - m2) + (m4 - m3)
    y_intermediate = (m4 - m1) + (m3 - m2)
    z = np.cross(x, y_intermediate)
    y = np.cross(z, x)
    return np.array([x / np.linalg.norm(x), y / np.linalg.norm(y), z / np.linalg.norm(z)]).T


def camera_transform(markers, origin_label, prefix=None):
    """4x4 float64: rotation from markers ``<prefix>1..4``, translation = point ``origin_label``."""
    prefix = origin_label if prefix is None else prefix
    if origin_label not in markers:
        raise KeyError("Label %s not found" % origin_label)
    t = np.eye(4)
    t[:3, :3] = camera_rotation(markers, prefix)
    t[:3, 3] = markers[origin_label]
    return t


def cameras_in(markers):
    """(origin_label, marker_prefix) for every camera the survey holds: a label C with C1..C4
    (the reference's A..H convention), or a longer name whose initial has the markers
    (calibration/pcs4.csv: DEXTRO + D1..D4, LEVO + L1..L4)."""
    cams = []
    for label in markers:
        if label[-1:].isdigit():
            continue
        for prefix in (label, label[0]):
            if all("%s%d" % (prefix, k) in markers for k in (1, 2, 3, 4)):
                cams.append((label, prefix))
                break
    return cams


def transforms_from_csv(path):
    m = load_markers(path)
    return {origin: camera_transform(m, origin, prefix) for origin, prefix in cameras_in(m)}


def save_transforms(path, transforms):
    with open(path, "w") as f:
        json.dump({k: np.asarray(v, np.float64).reshape(4, 4).tolist() for k, v in transforms.items()}, f, indent=1)


def load_transforms(path):
    """{camera: float32[16] row-major}, ready for ``stream_desc(tf=...)`` / ``pcs_stream_desc.tf``."""
    with open(path) as f:
        return {k: np.asarray(v, np.float32).reshape(-1) for k, v in json.load(f).items()}


if __name__ == "__main__":
    import sys
    np.set_printoptions(precision=8, suppress=True, floatmode="fixed")
    tfs = transforms_from_csv(sys.argv[1])
    for name, t in tfs.items():
        print(name)
        print(t)
    if len(sys.argv) > 2:
        save_transforms(sys.argv[2], tfs)
