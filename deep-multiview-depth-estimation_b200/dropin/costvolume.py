"""Drop-in for /root/reference/scripts/costvolume.py (`model.py:6` binds this `assemble_cost_volume`)."""
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import mvs_b200  # noqa: E402


def assemble_cost_volume(warped_feature_maps, n_views: int):
    return mvs_b200.assemble_cost_volume(warped_feature_maps, n_views)
