"""Generate golden vectors from the UNMODIFIED reference  --  TEST INFRASTRUCTURE.

Runs only in the authoring container, where /root/reference exists.  It imports the real
``/root/reference/scripts/{config,homography,costvolume,depthmap,model}.py`` (kornia is
provided by oracle/kornia_shim, see its docstring), patches ``config`` *before* the other
modules bind its constants, pushes seeded inputs through the reference's own
functions on CPU and stores inputs + outputs as ``tests/golden/*.npz``.

    python oracle/make_golden.py            # rewrites tests/golden/ (small cases, seconds)
    python oracle/make_golden.py --fullsize # tests/golden/cfg1_digest.npz: the reference at BASELINE size (~1-2 min CPU)

Nothing under tests/ or the product reads /root/reference at run time; the committed vectors are
what travels to the GPU box.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/scripts"
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, HERE)
import plane_sweep as ps  # noqa: E402  (fixtures only: DTU cameras)

_REF_MODULES = ("config", "homography", "costvolume", "depthmap", "model", "utils", "loss")


def load_reference(D, d_scale, in_h, in_w):
    """Fresh import of the reference with config patched for this problem size."""
    for m in _REF_MODULES:
        sys.modules.pop(m, None)
    for p in (REF, os.path.join(HERE, "kornia_shim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    config = importlib.import_module("config")
    config.D_NUM, config.D_SCALE = D, d_scale
    config.IN_H, config.IN_W = in_h, in_w
    config.FEAT_H, config.FEAT_W = in_h // config.DIM_REDUCE, in_w // config.DIM_REDUCE
    dims = np.array([D, config.FEAT_H, config.FEAT_W])
    config.PAD = tuple(int(x) for x in (np.floor(dims / 2) + 1))
    config.OUTPAD = tuple(int(x) for x in ((dims + 1) % 2))
    mods = {m: importlib.import_module(m) for m in ("homography", "costvolume", "depthmap", "model")}
    mods["config"] = config
    return mods


def smooth_features(n, c, h, w, gen):
    """conv-like (smooth) features: white noise through a 3x3 box blur, unit-ish scale."""
    x = torch.randn(n, c, h + 2, w + 2, generator=gen)
    k = torch.ones(c, 1, 3, 3) / 9.0
    return torch.nn.functional.conv2d(x, k, groups=c) * 3.0


def run_case(name, B, V, D, h, w, d_scale, seed, keep_warped=False, with_reg=True, d_min_val=425.0, keep_gparam=None):
    ref = load_reference(D, d_scale, h * 4, w * 4)
    gen = torch.Generator().manual_seed(seed)
    K, R, T = ps.synthetic_cameras(B, V, h, w, seed=seed)
    keep_gparam = keep_warped if keep_gparam is None else keep_gparam
    if isinstance(d_min_val, (list, tuple)):                          # unequal d_min per batch item (homography.py:26 quirk)
        d_min = torch.tensor(d_min_val, dtype=torch.float32).reshape(B, 1, 1, 1)
    else:
        d_min = torch.full((B, 1, 1, 1), d_min_val)
    d_int = torch.ones(B, 1, 1, 1)
    feat = smooth_features(B * V, 32, h, w, gen).requires_grad_(True)

    # record the homographies the reference hands to kornia (homography.py:85-86)
    rec = []
    real_warp = ref["homography"].warp_perspective
    ref["homography"].warp_perspective = lambda src, M, *a, **k: (rec.append(M.detach().clone()), real_warp(src, M, *a, **k))[1]
    warped, d_batch, ref_idx = ref["homography"].homography_warping(K, R, T, d_min, d_int, feat, B, V)
    H = torch.stack(rec, 1)                                            # N,D,3,3
    cost = ref["costvolume"].assemble_cost_volume(warped, V)

    out = dict(B=B, V=V, D=D, h=h, w=w, d_scale=d_scale, K=K, R=R, T=T, d_min=d_min, d_int=d_int,
               feat=feat.detach(), H=H, d_batch=d_batch, ref_idx=ref_idx, cost=cost.detach())
    if keep_warped:
        out["warped"] = warped.detach()

    # gradient of a seeded weighted sum of the cost volume wrt the features (K2 parity)
    gw = torch.randn(cost.shape, generator=gen)
    (gfeat,) = torch.autograd.grad((cost * gw).sum(), feat, retain_graph=False)
    out["gcost"] = gw; out["gfeat"] = gfeat

    if with_reg:
        torch.manual_seed(1234)                                        # same weights in every case
        reg = ref["model"].CostVolumeReg()
        reg.train()
        logit_box = []
        reg.conv_out.register_forward_hook(lambda m, i, o: logit_box.append(o.detach()))
        cv = cost.detach().clone().requires_grad_(True)
        prob = reg(cv)
        depth = ref["depthmap"].extract_depth_map(prob, d_batch)
        gd = torch.randn(depth.shape, generator=gen)
        params = [p for p in reg.parameters()]
        grads = torch.autograd.grad((depth * gd).sum(), [cv] + params)
        out.update(prob=prob.detach(), logits=logit_box[0], depth=depth.detach(), gdepth=gd, gcv=grads[0])
        sd = {k: v.detach().clone() for k, v in reg.state_dict().items()}       # after 1 train fwd
        out.update({"bn_after/" + k: v for k, v in sd.items() if "running" in k or "tracked" in k})
        if keep_gparam:   # parameter gradients only in the tiny case (1.3 MB per copy)
            out.update({"gparam/" + n: g for (n, _), g in zip(reg.named_parameters(), grads[1:])})
        # initial weights (identical for every case thanks to the fixed seed) are stored once
        wpath = os.path.join(OUT, "reg_weights.npz")
        if not os.path.exists(wpath):
            torch.manual_seed(1234)
            w0 = ref["model"].CostVolumeReg().state_dict()
            np.savez(wpath, **{k: v.numpy() for k, v in w0.items()})
    np.savez(os.path.join(OUT, name + ".npz"),
             **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()})
    print(f"{name}: cost {tuple(cost.shape)} max {float(cost.detach().max()):.4f}"
          + (f" depth [{float(out['depth'].min()):.2f}, {float(out['depth'].max()):.2f}]" if with_reg else ""))


def run_depth_ties(name, seed=7):
    """extract_depth_map on a probability volume with planted exact ties."""
    D, h, w = 12, 6, 10
    ref = load_reference(D, 40, 4 * h, 4 * w)
    gen = torch.Generator().manual_seed(seed)
    logits = torch.randn(2, 1, D, h, w, generator=gen)
    logits[:, :, 3, :, :5] = logits[:, :, 1, :, :5]          # tie between two of planes 0..4
    logits[:, :, 9, 2:4] = logits[:, :, 0, 2:4]              # tie between plane 0 and a later plane
    logits[0, :, :, 5, 9] = 0.25                              # all planes equal
    prob = torch.softmax(logits, 2)
    d_batch = ps.depth_table(torch.full((2, 1, 1, 1), 425.0), torch.ones(2, 1, 1, 1), D, 40)
    depth = ref["depthmap"].extract_depth_map(prob, d_batch)
    np.savez(os.path.join(OUT, name + ".npz"), prob=prob.numpy(), d_batch=d_batch.numpy(), depth=depth.numpy())
    print(f"{name}: depth [{float(depth.min()):.2f}, {float(depth.max()):.2f}]")


DIGEST_PLANES = [0, 1, 47, 96, 143, 191]
DIGEST_COST_CH = [0, 13, 31]


def run_fullsize_digest(name="cfg1_digest", seed=11):
    """BASELINE.json configs[0]/[1] size (B=1, V=3, D=192, 128x160x32): the unmodified reference's whole hot path on CPU
    (~1 min), stored as a DIGEST -- the depth map, the kept-plane ranks and their stability margin, sampled planes of the cost /
    logit / probability volumes, BatchNorm running statistics -- because the full tensors (503 MB cost volume) cannot be
    committed.  Inputs are rebuilt from the seed on the GPU box (ps.smooth_features_exact, checksum stored)."""
    B, V, D, h, w = 1, 3, 192, 128, 160
    d_scale = 480.0 / D
    ref = load_reference(D, d_scale, h * 4, w * 4)
    K, R, T = ps.synthetic_cameras(B, V, h, w, seed=seed)
    d_min, d_int = torch.full((B, 1, 1, 1), 425.0), torch.ones(B, 1, 1, 1)
    feat = ps.smooth_features_exact(B * V, 32, h, w, seed)
    with torch.no_grad():
        warped, d_batch, ref_idx = ref["homography"].homography_warping(K, R, T, d_min, d_int, feat, B, V)
        cost = ref["costvolume"].assemble_cost_volume(warped, V)
        del warped
        torch.manual_seed(1234)
        reg = ref["model"].CostVolumeReg()
        reg.train()
        box = []
        reg.conv_out.register_forward_hook(lambda m, i, o: box.append(o.detach()))
        prob = reg(cost)
        depth = ref["depthmap"].extract_depth_map(prob, d_batch)
    logits = box[0]
    pn = prob.numpy()
    odepth, ranks = ps.extract_depth(pn, d_batch.numpy())
    ties = ps.tie_pixels(pn)
    margin = ps.rank_margin(pn)
    err = np.abs(odepth[:, 0] - depth.numpy()[:, 0])[~ties].max()
    oprob = ps.reg_forward(_fresh_reg_state(ref), cost, train_bn=True)
    perr = float((oprob - prob).abs().max() / prob.abs().max())
    print(f"{name}: oracle-vs-reference at full size: depth (ties excluded) {err:.3e}, prob rel {perr:.3e}; ties {int(ties.sum())} px")
    sd = reg.state_dict()
    out = dict(B=B, V=V, D=D, h=h, w=w, d_scale=d_scale, seed=seed, K=K, R=R, T=T, d_min=d_min, d_int=d_int,
               d_batch=d_batch, planes=np.array(DIGEST_PLANES), cost_ch=np.array(DIGEST_COST_CH),
               feat_sum=np.float64(feat.double().sum().item()), feat_abs_sum=np.float64(feat.double().abs().sum().item()),
               feat_probe=feat[:, ::11, ::37, ::41].numpy(),
               cost=cost[:, DIGEST_COST_CH][:, :, DIGEST_PLANES].numpy(), cost_absmax=np.float32(cost.abs().max().item()),
               logits=logits[:, :, DIGEST_PLANES].numpy(), logits_absmax=np.float32(logits.abs().max().item()),
               prob=prob[:, :, DIGEST_PLANES].numpy(), prob_absmax=np.float32(prob.abs().max().item()),
               depth=depth.numpy(), ranks=ranks.astype(np.int16), ties=ties, margin=margin.astype(np.float32),
               oracle_prob_relerr=np.float64(perr), oracle_depth_abserr=np.float64(err))
    out.update({"bn_after/" + k: v.numpy() for k, v in sd.items() if "running" in k or "tracked" in k})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: depth [{float(depth.min()):.2f}, {float(depth.max()):.2f}], median rank margin {float(np.median(margin)):.2e}")


def _fresh_reg_state(ref):
    torch.manual_seed(1234)
    return {k: v.detach().clone() for k, v in ref["model"].CostVolumeReg().state_dict().items()}


def main():
    if "--fullsize" in sys.argv:                   # ~1-2 min of CPU: kept out of the default regeneration
        os.makedirs(OUT, exist_ok=True)
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        run_fullsize_digest()
        return
    os.makedirs(OUT, exist_ok=True)
    wpath = os.path.join(OUT, "reg_weights.npz")
    if os.path.exists(wpath):
        os.remove(wpath)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    run_case("tiny_b1v3", B=1, V=3, D=8, h=16, w=20, d_scale=60, seed=0, keep_warped=True)
    run_case("b2v3", B=2, V=3, D=8, h=16, w=20, d_scale=60, seed=1)
    run_case("v5", B=1, V=5, D=6, h=16, w=20, d_scale=80, seed=2, with_reg=False)
    run_case("v7_odd", B=1, V=7, D=7, h=15, w=19, d_scale=70, seed=3)
    run_case("val_dmin0", B=1, V=3, D=8, h=16, w=20, d_scale=60, seed=4, with_reg=False, d_min_val=0.0)
    run_depth_ties("depth_ties")
    # batch quirk (homography.py:26) with UNEQUAL d_min: flat view i reads depth row i mod B
    run_case("bquirk_b2v3", B=2, V=3, D=8, h=16, w=20, d_scale=60, seed=5, keep_warped=True, keep_gparam=False,
             d_min_val=[425.0, 520.0])


if __name__ == "__main__":
    main()
