"""Drop-in for /root/reference/scripts/loss.py (`train.py` / `test.py` bind this `loss_fcn`): the masked L1 training loss on the
fused kernels of libmvs_b200.so -- one launch forward, one backward."""
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import mvs_b200  # noqa: E402


def loss_fcn(gt, initial, refined):
    """Compute loss from initial and refined depth maps: (loss, initial_acc, refined_acc), scripts/loss.py:4-41."""
    return mvs_b200.ops.masked_l1_loss(gt, initial, refined)
