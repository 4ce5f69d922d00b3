#!/usr/bin/env python
"""Golden vectors for SURVEY §8 row f2 (depth refinement and the glue around it): the UNMODIFIED reference's `DepthRefinement`
(/root/reference/scripts/model.py:129-152) and the lines of `MVSNet.forward` around it (model.py:190-205: normalise, resize the
reference image, concatenate, refine, de-normalise), run on the CPU in fp32 on seeded inputs, forward and backward, recorded
into tests/golden/refine.npz together with the weights.  Test infrastructure only; run in the container that holds
/root/reference:   python oracle/make_golden_refine.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402

B, V, D, D_SCALE, IN_H, IN_W = 2, 3, 8, 2.5, 48, 64


def main():
    ref = load_reference(D, D_SCALE, IN_H, IN_W)
    cfg, model = ref["config"], ref["model"]
    torch.manual_seed(77)
    net = model.DepthRefinement(device=torch.device("cpu")).train()
    with torch.no_grad():                                            # BatchNorm affine parameters away from (1, 0)
        for m in net.model:
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0.0, 0.3)
    weights = {k: v.detach().clone().numpy() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(78)
    h, w = cfg.FEAT_H, cfg.FEAT_W
    nn_input = torch.rand(B * V, 3, IN_H, IN_W, generator=gen)
    d_min = torch.tensor([425.0, 431.5]).reshape(B, 1, 1, 1)
    d_int = torch.tensor([1.0, 1.25]).reshape(B, 1, 1, 1)
    initial = (d_min + d_int * D * D_SCALE * torch.rand(B, 1, h, w, generator=gen)).requires_grad_(True)
    ref_views = torch.arange(0, B * V, V)

    # model.py:190-205, verbatim in meaning (DEVICE = cpu here)
    d_trans = d_min
    d_scale = d_int.mul(cfg.D_NUM).mul(cfg.D_SCALE)
    norm = torch.div(torch.subtract(initial, d_trans), d_scale)
    refine_input = torch.cat((norm, torch.nn.functional.interpolate(nn_input[ref_views], (cfg.FEAT_H, cfg.FEAT_W), mode="bilinear")), dim=1)
    refined = net(refine_input).mul(d_scale).add(d_trans)

    g = torch.randn(refined.shape, generator=gen)
    refined.backward(g)
    out = {"nn_input": nn_input.numpy(), "d_min": d_min.numpy(), "d_int": d_int.numpy(), "initial": initial.detach().numpy(),
           "refine_input": refine_input.detach().numpy(), "refined": refined.detach().numpy(), "g_refined": g.numpy(),
           "g_initial": initial.grad.numpy(), "d_num": np.int64(D), "d_scale": np.float64(D_SCALE), "n_views": np.int64(V)}
    for k, v in weights.items():
        out["w0." + k] = v
    for k, v in net.state_dict().items():                            # running statistics after the one train-mode pass
        if "running" in k or "num_batches" in k:
            out["w1." + k] = v.detach().numpy()
    for k, p in net.named_parameters():
        out["g." + k] = p.grad.numpy()
    path = os.path.join(HERE, "..", "tests", "golden", "refine.npz")
    np.savez_compressed(path, **out)
    print("wrote", os.path.normpath(path), {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
