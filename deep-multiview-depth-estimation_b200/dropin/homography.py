"""Drop-in for /root/reference/scripts/homography.py -- same module name, same function, same signature.

Put this directory ahead of the reference's scripts/ on sys.path; `model.py:5` then binds
this `homography_warping`.  Constants come from the host application's `config` module, read at call time.
"""
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import mvs_b200  # noqa: E402


def homography_warping(K_batch, R_batch, T_batch, d_min, d_int, feature_maps, batch_size, n_views, d_num=None):
    import config
    return mvs_b200.homography_warping(K_batch, R_batch, T_batch, d_min, d_int, feature_maps, batch_size, n_views,
                                       config.D_NUM if d_num is None else d_num, config.D_SCALE)
