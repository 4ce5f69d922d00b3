"""Host logic of the CostVolumeReg drop-in (central-region algebra + analytic BatchNorm statistics) checked on
CPU against the reference's goldens.  The conv backend here is torch's CPU conv; the product forward()
itself refuses CPU tensors (checked below)."""
import os

import numpy as np
import pytest
import torch

import mvs_b200
from mvs_b200 import conv3d as conv_backends


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


def _reg(golden_dir):
    reg = mvs_b200.CostVolumeReg(device="cpu", precision="fp32")
    w0 = np.load(os.path.join(golden_dir, "reg_weights.npz"))
    reg.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in w0.items()})     # reference key names
    return reg


def _relmax(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max() / np.abs(b).max())


def test_state_dict_keys_and_init_match_reference(golden_dir):
    w0 = np.load(os.path.join(golden_dir, "reg_weights.npz"))
    torch.manual_seed(1234)                                   # seed used by oracle/make_golden.py
    reg = mvs_b200.CostVolumeReg(device="cpu", precision="fp32")
    sd = reg.state_dict()
    assert sorted(sd.keys()) == sorted(w0.keys())
    for k in w0:
        assert np.array_equal(sd[k].numpy(), w0[k]), k        # same construction order => same RNG draws
    assert sum(p.numel() for p in reg.parameters()) == 321864  # report Table 1


@pytest.mark.parametrize("n,expect", [(192, (48, 144, 1)), (8, (2, 6, 1)), (15, (3, 11, 2)), (7, (1, 5, 2)),
                                      (128, (32, 96, 1)), (160, (40, 120, 1)), (2, (0, 1, 2))])
def test_central_region(n, expect):
    assert mvs_b200.central_region(n) == expect
    # brute force: which outputs of the reference's stride-2 conv can see any real input
    p = n // 2 + 1
    live = [o for o in range(n) if any(0 <= 2 * o - p + k < n for k in range(3))]
    assert (live[0], live[-1]) == expect[:2]
    read = sorted({(o + p - k) // 2 for o in range(n) for k in range(3) if (o + p - k) % 2 == 0 and 0 <= (o + p - k) // 2 < n})
    assert read[0] >= expect[0] and read[-1] <= expect[1]      # a transposed conv reads only the box


@pytest.mark.parametrize("name", ["tiny_b1v3", "b2v3", "v7_odd"])
def test_logits_and_running_stats_match_reference(golden_dir, name):
    g = _load(golden_dir, name)
    reg = _reg(golden_dir).train()
    logits = reg.logits(torch.from_numpy(g["cost"]), conv_backends.get("cudnn"))
    assert logits.shape == g["logits"].shape
    assert _relmax(logits.detach().numpy(), g["logits"]) < 1e-4
    prob = torch.softmax(logits, 2)
    assert _relmax(prob.detach().numpy(), g["prob"]) < 1e-4
    sd = reg.state_dict()
    for k, v in g.items():
        if k.startswith("bn_after/"):
            assert np.allclose(sd[k[len("bn_after/"):]].numpy(), v, rtol=2e-5, atol=1e-6), k


def test_gradients_match_reference(golden_dir):
    g = _load(golden_dir, "tiny_b1v3")
    reg = _reg(golden_dir).train()
    cv = torch.from_numpy(g["cost"]).requires_grad_(True)
    logits = reg.logits(cv, conv_backends.get("cudnn"))
    # reference loss: sum(depth * gdepth) with depth = extract_depth_map(softmax(logits)); restate with torch ops
    prob = torch.softmax(logits, 2)
    _, order = prob.sort(dim=2, descending=True, stable=True)
    filt = prob * (order < 5).float()
    depth = (torch.from_numpy(g["d_batch"]).unsqueeze(1) * filt).sum(2).squeeze(2) / filt.sum(2)
    names = [n for n, _ in reg.named_parameters()]
    grads = torch.autograd.grad((depth * torch.from_numpy(g["gdepth"])).sum(), [cv] + list(reg.parameters()))
    assert _relmax(grads[0].numpy(), g["gcv"]) < 2e-4
    for n, gr in zip(names, grads[1:]):
        ref = g["gparam/" + n]
        assert np.abs(gr.numpy() - ref).max() <= 2e-4 * max(np.abs(ref).max(), 1e-3), n


def test_eval_mode_uses_running_stats(golden_dir):
    g = _load(golden_dir, "tiny_b1v3")
    reg = _reg(golden_dir).train()
    cv = torch.from_numpy(g["cost"])
    reg.logits(cv, conv_backends.get("cudnn"))                 # one train pass moves the running stats
    reg.eval()
    ours = reg.logits(cv, conv_backends.get("cudnn")).detach()
    # plain dense-canvas evaluation with the same (updated) statistics
    import plane_sweep as ps
    sd = {k: v.clone() for k, v in reg.state_dict().items()}
    _, ref = ps.reg_forward(sd, cv, train_bn=False, return_logits=True)
    assert _relmax(ours.numpy(), ref.numpy()) < 1e-4


def test_forward_refuses_cpu(golden_dir):
    reg = _reg(golden_dir)
    with pytest.raises(mvs_b200.MvsB200Error):
        reg(torch.zeros(1, 32, 4, 4, 4))
