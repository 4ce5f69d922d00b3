"""Drop-in for /root/reference/scripts/depthmap.py (`model.py:7` binds this `extract_depth_map`)."""
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import mvs_b200  # noqa: E402


def extract_depth_map(prob_volume, d_batch):
    import config
    return mvs_b200.extract_depth_map(prob_volume, d_batch, int(config.N_DEPTH_EST))
