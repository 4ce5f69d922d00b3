#!/usr/bin/env python
"""Golden vectors for SURVEY §8 row f1 (the 2D feature encoder): the UNMODIFIED reference's `FeatureEncoder`
(/root/reference/scripts/model.py:20-65) run on the CPU in fp32, train-mode BatchNorm, on seeded images -- forward, the
gradients of every parameter for a seeded upstream gradient, and the running statistics after the pass -- recorded with the
weights into tests/golden/encoder.npz.  Test infrastructure only; run in the container that holds /root/reference:
    python oracle/make_golden_encoder.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402

N, IN_H, IN_W = 3, 24, 40


def main():
    ref = load_reference(8, 2.5, IN_H, IN_W)
    torch.manual_seed(91)
    net = ref["model"].FeatureEncoder(device=torch.device("cpu")).train()
    with torch.no_grad():                                            # BatchNorm affine parameters away from (1, 0)
        for m in net.model:
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0.0, 0.3)
    out = {"w0." + k: v.detach().clone().numpy() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(92)
    images = torch.rand(N, 3, IN_H, IN_W, generator=gen)
    feats = net(images)
    g = torch.randn(feats.shape, generator=gen)
    feats.backward(g)
    out.update({"images": images.numpy(), "features": feats.detach().numpy(), "g_features": g.numpy()})
    for k, v in net.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["w1." + k] = v.detach().numpy()
    for k, p in net.named_parameters():
        out["g." + k] = p.grad.numpy()
    path = os.path.join(HERE, "..", "tests", "golden", "encoder.npz")
    np.savez_compressed(path, **out)
    print("wrote", os.path.normpath(path), feats.shape, len(out), "arrays")


if __name__ == "__main__":
    main()
