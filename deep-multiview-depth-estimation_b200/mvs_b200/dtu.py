"""DTU on-disk formats and the training checkpoint dictionary -- host-side Python, as the reference's is.

What the reference reads and writes around the path (/root/reference/scripts):
  data.py:40-69     `Cameras.load`   `<8 digits>_cam.txt`: "extrinsic" + 4x4 world-to-camera matrix, "intrinsic" + 3x3 K, then
                                     `d_min d_interval` -> K [3,3], R [3,3], T [3,1], d [1,1], d_int [1,1] (float64)
  data.py:71-80     `Cameras.pair`   `pair.txt`: per reference view the ids of its best source views (every second token of the
                                     score line)
  data.py:327-358   `load_depth`     PFM depth maps (`Pf` / `PF`), flipped to top-down rows
  data.py:446-460   `CustomSampler`  resumable sampler: shuffle once, skip the first i * batch_size indices
  train.py:111-121  checkpoint dict  keys epoch, batch_idx, model_state_dict, optimizer_state_dict, scheduler_state_dict, loss,
  train.py:167-187  and its reload   acc_1, acc_2

These functions reproduce the reference's results ON THE SAME FILES, quirks included (each is named where it occurs); the
fixtures under tests/golden/dtu/ were produced by the unmodified reference (oracle/make_golden_dtu.py).  Nothing here touches
the GPU."""
from __future__ import annotations

import os
import random
import re

import numpy as np
import torch


def read_cam_file(path: str):
    """One `*_cam.txt` -> (K [3,3], R [3,3], T [3,1], d_min [1,1], d_int [1,1]), float64 (data.py:40-69).  The file layout is
    positional, as the reference reads it: line 1 a label, lines 2-5 the 4x4 extrinsic matrix, two skipped lines, lines 8-10 K, one
    skipped line, line 12 `d_min d_int [...]`."""
    with open(path) as f:
        f.readline()
        rows = [np.float64(f.readline().split()) for _ in range(4)]
        f.readline(); f.readline()
        k = [np.float64(f.readline().split()) for _ in range(3)]
        f.readline()
        d = np.float64(f.readline().split())
    K = np.vstack(k)
    R = np.vstack([r[0:3] for r in rows[:3]])
    T = np.vstack([r[-1] for r in rows[:3]])
    return K, R, T, np.array([d[0]]).reshape(-1, 1), np.array([d[1]]).reshape(-1, 1)


def cam_file_names(base_path: str, cam_list):
    """Paths of the camera files of `cam_list` (data.py:30-36): <base>/Cameras/train/<idx, 8 digits>_cam.txt."""
    folder = os.path.join(base_path, "Cameras", "train")
    return [os.path.join(folder, "{:0>8}".format(str(i)) + "_cam.txt") for i in cam_list]


def read_cameras(base_path: str, cam_list):
    """`Cameras(path, cam_list)` of the reference without the class: dict of lists K, R, T, d, d_int (one entry per camera of
    cam_list, data.py:40-69) and pairs (data.py:71-80)."""
    out = {"K": [], "R": [], "T": [], "d": [], "d_int": []}
    for p in cam_file_names(base_path, cam_list):
        K, R, T, d, di = read_cam_file(p)
        out["K"].append(K); out["R"].append(R); out["T"].append(T); out["d"].append(d); out["d_int"].append(di)
    out["pairs"] = read_pairs(os.path.join(base_path, "Cameras", "train", "..", "pair.txt"), cam_list)
    return out


def read_pairs(pair_txt: str, cam_list):
    """`pair.txt` -> list of int64 arrays, the source-view ids of every reference view found in cam_list (data.py:71-80).
    The reference's control flow is kept verbatim because its result depends on it: the first view line is tested by its FIRST
    CHARACTER (it is still a string), later ones by their first token; a view outside cam_list leaves its score line to be read
    as the next "view line" (its first token, the number of pairs, is then what is looked up in cam_list)."""
    pairs = []
    with open(pair_txt) as f:
        f.readline()                                   # header: number of views
        line = f.readline()
        while line:
            if int(line[0]) in cam_list:
                pair_line = f.readline().split()
                pairs.append(np.int64(pair_line[1::2]))
            line = f.readline().split()
    return pairs


def depth_file_names(base_path: str, cam_list, scan_idx):
    """Per scan the ground-truth depth files of cam_list (data.py:82-100): <base>/Depths/scan<k>_train/depth_map_<idx, 4 digits>.pfm."""
    return [[os.path.join(base_path, "Depths", "scan" + str(s) + "_train", "depth_map_" + "{:0>4}".format(str(i)) + ".pfm")
             for i in cam_list] for s in scan_idx]


def load_pfm(path: str) -> np.ndarray:
    """PFM file -> float32 rows top-down, [height, width] for `Pf` and [height, width, 3] for `PF` (data.py:327-358,
    `load_depth`: cv2.flip returns a single-channel [h, w, 1] array as [h, w]).  Quirk kept: the reference
    reads the samples LITTLE-endian when the scale line is POSITIVE and big-endian otherwise -- the reverse of the PFM
    convention (negative = little-endian) -- so a file is decoded exactly as the reference decodes it."""
    with open(path, "rb") as f:
        header = f.readline().decode("UTF-8").rstrip()
        dim_match = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("UTF-8"))
        scale = float(f.readline().decode("UTF-8").rstrip())
        data_string = f.read()
    if header == "PF":
        ch_dim = 3
    elif header == "Pf":
        ch_dim = 1
    else:
        raise Exception("Invalid Header for PFM file.")
    if not dim_match:
        raise Exception("PFM header gives no dimensions.")
    width, height = map(int, dim_match.groups())
    data = np.frombuffer(data_string, "<f" if scale > 0 else ">f")
    data = np.reshape(data, (height, width, ch_dim))
    data = np.flipud(data)                             # cv2.flip(data, 0): the file stores rows bottom-up ...
    return data[:, :, 0] if ch_dim == 1 else data      # ... and cv2 hands a one-channel image back without its channel axis


def write_pfm(path: str, image: np.ndarray, scale: float = 1.0):
    """Inverse of load_pfm under the same (reversed) endianness rule: load_pfm(write_pfm(x)) == x."""
    image = np.asarray(image, dtype=np.float32)
    if image.ndim == 2:
        image = image[:, :, None]
    h, w, c = image.shape
    if c not in (1, 3):
        raise ValueError("PFM holds 1 or 3 channels")
    with open(path, "wb") as f:
        f.write(("PF\n" if c == 3 else "Pf\n").encode())
        f.write(f"{w} {h}\n".encode())
        f.write(f"{float(scale)}\n".encode())
        f.write(np.flipud(image).astype("<f" if scale > 0 else ">f").tobytes())


class ResumableSampler(torch.utils.data.Sampler):
    """`CustomSampler` (data.py:446-460): shuffles `data` IN PLACE once with the global `random` state, then yields the indices
    i * batch_size .. len(data) - 1 in order -- a loader pickled mid-epoch resumes where batch i would have started."""

    def __init__(self, data, i=0, batch_size=14):
        random.shuffle(data)
        self.seq = list(range(len(data)))[i * batch_size:]

    def __iter__(self):
        return iter(self.seq)

    def __len__(self):
        return len(self.seq)


CHECKPOINT_KEYS = ("epoch", "batch_idx", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "loss", "acc_1", "acc_2")


def checkpoint_dict(epoch, batch_idx, model, optimizer, scheduler, loss, acc_1, acc_2):
    """The dictionary train.py:111-121 hands to torch.save (same keys, same order)."""
    return {"epoch": epoch, "batch_idx": batch_idx, "model_state_dict": model.state_dict(),
            "optimizer_state_dict": optimizer.state_dict(), "scheduler_state_dict": scheduler.state_dict(),
            "loss": loss, "acc_1": acc_1, "acc_2": acc_2}


def checkpoint_name(save_path, id_str, epoch, batch_idx):
    """File name of a checkpoint (train.py:121): <save_path>/<id>_<epoch>_<batch_idx>."""
    return os.path.join(save_path, id_str + "_" + str(epoch) + "_" + str(batch_idx))


def load_checkpoint(ckpt, model, optimizer, scheduler, epochs, map_location=None):
    """train.py:167-187 (`load_from_ckpt`) on already-built objects: restores the three state dicts and returns
    (start_epoch, batch_idx, loss, acc_1, acc_2); raises ValueError when the checkpoint already covers `epochs`."""
    checkpoint = torch.load(ckpt, map_location=map_location, weights_only=False) if isinstance(ckpt, (str, os.PathLike)) else ckpt
    if epochs - (checkpoint["epoch"] + 1) <= 0:
        raise ValueError("Epochs provided: {:d}, epochs completed in ckpt: {:d}".format(epochs, checkpoint["epoch"] + 1))
    model.load_state_dict(checkpoint["model_state_dict"])
    optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
    scheduler.load_state_dict(checkpoint["scheduler_state_dict"])
    return checkpoint["epoch"] + 1, checkpoint["batch_idx"], checkpoint["loss"], checkpoint["acc_1"], checkpoint["acc_2"]
