#!/usr/bin/env python
"""Golden vectors for SURVEY §8 row f4 (DTU on-disk formats): writes a tiny synthetic DTU-shaped tree under tests/golden/dtu/
(camera files, pair.txt, PFM depth maps) and records what the UNMODIFIED reference (/root/reference/scripts/data.py: Cameras,
load_depth, CustomSampler) reads from it into tests/golden/dtu_formats.npz.  Test infrastructure only; run in the container that
holds /root/reference:   python oracle/make_golden_dtu.py"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = os.path.join(ROOT, "tests", "golden", "dtu")
N_CAM = 4


def write_tree():
    rng = np.random.default_rng(20260)
    cam_dir = os.path.join(BASE, "Cameras", "train")
    os.makedirs(cam_dir, exist_ok=True)
    for i in range(N_CAM):
        ang = 0.1 * (i + 1)
        R = np.array([[np.cos(ang), -np.sin(ang), 0.0], [np.sin(ang), np.cos(ang), 0.0], [0.0, 0.0, 1.0]])
        T = rng.normal(size=3) * 100.0
        K = np.array([[361.54125 + i, 0.0, 82.900625], [0.0, 360.3975 + i, 66.383875], [0.0, 0.0, 1.0]])
        with open(os.path.join(cam_dir, "{:0>8}".format(str(i)) + "_cam.txt"), "w") as f:
            f.write("extrinsic\n")
            for r in range(3):
                f.write(" ".join(repr(float(v)) for v in list(R[r]) + [T[r]]) + "\n")
            f.write("0.0 0.0 0.0 1.0\n\nintrinsic\n")
            for r in range(3):
                f.write(" ".join(repr(float(v)) for v in K[r]) + "\n")
            f.write("\n{} {}\n".format(425.0 + 5 * i, 2.5 + 0.1 * i))
    with open(os.path.join(BASE, "Cameras", "pair.txt"), "w") as f:
        f.write(f"{N_CAM}\n")
        for i in range(N_CAM):
            others = [j for j in range(N_CAM) if j != i]
            f.write(f"{i}\n{len(others)} " + " ".join(f"{j} {100.0 / (1 + abs(i - j)):.3f}" for j in others) + " \n")
    dep_dir = os.path.join(BASE, "Depths", "scan1_train")
    os.makedirs(dep_dir, exist_ok=True)
    for i in range(2):
        h, w = 5, 7
        img = (425.0 + 50.0 * rng.random((h, w))).astype(np.float32)
        img[0, :2] = 0.0                                              # invalid pixels
        scale = 1.0 if i == 0 else -1.0                               # both branches of the reference's endianness rule
        with open(os.path.join(dep_dir, "depth_map_" + "{:0>4}".format(str(i)) + ".pfm"), "wb") as f:
            f.write(b"Pf\n")
            f.write(f"{w} {h}\n".encode())
            f.write(f"{scale}\n".encode())
            f.write(np.flipud(img).astype("<f4").tobytes())          # little-endian bytes on disk in both files
    with open(os.path.join(dep_dir, "colour_0000.pfm"), "wb") as f:   # 3-channel PFM
        img3 = rng.random((3, 4, 3)).astype(np.float32)
        f.write(b"PF\n4 3\n1.0\n")
        f.write(np.flipud(img3).astype("<f4").tobytes())


def main():
    write_tree()
    sys.path.insert(0, "/root/reference/scripts")
    import data as ref                                                # the unmodified reference

    out = {}
    for tag, cams in (("all", [0, 1, 2, 3]), ("sub02", [0, 2]), ("sub03", [0, 3])):
        c = ref.Cameras(BASE, cams)
        out[f"{tag}_K"] = np.stack(c.K); out[f"{tag}_R"] = np.stack(c.R); out[f"{tag}_T"] = np.stack(c.T)
        out[f"{tag}_d"] = np.stack(c.d); out[f"{tag}_d_int"] = np.stack(c.d_int)
        out[f"{tag}_n_pairs"] = np.int64(len(c.pairs))
        for k, p in enumerate(c.pairs):
            out[f"{tag}_pair{k}"] = np.asarray(p, dtype=np.int64)
    dep_dir = os.path.join(BASE, "Depths", "scan1_train")
    for name in ("depth_map_0000.pfm", "depth_map_0001.pfm", "colour_0000.pfm"):
        out["pfm_" + name] = np.ascontiguousarray(ref.load_depth(os.path.join(dep_dir, name)))
    d = ref.Depths(BASE, [0, 1], scan_idx=[1])
    out["depth_file_names"] = np.array([os.path.relpath(p, BASE) for p in d.file_names[0]])
    random.seed(1234)
    items = list(range(40))
    s = ref.CustomSampler(items, i=2, batch_size=7)
    out["sampler_seq"] = np.int64(list(s)); out["sampler_items_after"] = np.int64(items); out["sampler_len"] = np.int64(len(s))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dtu_formats.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
