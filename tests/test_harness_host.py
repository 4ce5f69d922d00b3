"""Host-side checks of the harness pieces that wrap the hot path (no GPU): the fused BatchNorm2d+ReLU module keeps
nn.BatchNorm2d's parameters / buffers and its stock behaviour off the GPU, and the harness nets keep the reference's
state_dict layout (Conv2d / BatchNorm2d / ReLU triples, model.py:22-65, :129-152)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from mvs_b200.harness import BatchNormReLU2d, FeatureEncoder, DepthRefinement, loss_fcn


def test_batchnorm_relu_2d_is_batchnorm_then_relu_off_the_gpu():
    torch.manual_seed(0)
    fused, bn = BatchNormReLU2d(8), nn.BatchNorm2d(8)
    with torch.no_grad():
        fused.weight.uniform_(0.5, 1.5); fused.bias.uniform_(-0.5, 0.5)
    bn.load_state_dict(fused.state_dict())
    assert sorted(fused.state_dict()) == sorted(bn.state_dict())
    x = torch.randn(3, 8, 6, 7)
    for train in (True, False):
        fused.train(train); bn.train(train)
        assert torch.allclose(fused(x), F.relu(bn(x)), atol=1e-6)
    assert torch.allclose(fused.running_mean, bn.running_mean) and torch.allclose(fused.running_var, bn.running_var)


def test_harness_nets_keep_the_triples_layout_of_the_reference():
    enc, ref = FeatureEncoder(), DepthRefinement()
    keys = list(enc.state_dict())
    # 7 conv+BN triples and a final conv: indices 0,1 | 3,4 | ... | 18,19 | 21 (ReLU / Identity slots carry no state)
    assert keys[0] == "model.0.weight" and "model.1.running_mean" in keys and "model.21.weight" in keys
    assert not any(k.startswith("model.2.") or k.startswith("model.20.") for k in keys)
    assert sum(p.numel() for p in enc.parameters()) + sum(p.numel() for p in ref.parameters()) > 0
    x = torch.randn(2, 3, 32, 40)
    assert enc.eval()(x).shape == (2, 32, 8, 10)
    assert ref.eval()(torch.randn(2, 4, 8, 10)).shape == (2, 1, 8, 10)


def test_loss_matches_the_reference_expression():
    """scripts/loss.py:4-41: masked L1 of both maps, per-item mean over valid pixels, summed over the batch."""
    g = torch.Generator().manual_seed(1)
    gt = 425 + 480 * torch.rand(3, 1, 8, 10, generator=g)
    gt = gt * (torch.rand(3, 1, 8, 10, generator=g) > 0.3)
    a, b = 600 + torch.randn(3, 1, 8, 10, generator=g), 600 + torch.randn(3, 1, 8, 10, generator=g)
    loss, acc0, acc1 = loss_fcn(gt, a, b)
    mask = (gt != 0).float()
    n = mask.sum((1, 2, 3))
    l0 = (mask * (gt - a).abs()).sum((1, 2, 3)) / n
    l1 = (mask * (gt - b).abs()).sum((1, 2, 3)) / n
    assert torch.allclose(loss, (l0 + l1).sum()) and torch.allclose(acc0, l0.mean()) and torch.allclose(acc1, l1.mean())
