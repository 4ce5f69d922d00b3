"""CPU oracle for MVSNet's plane-sweep hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``deep-multiview-depth-estimation_b200/``) never imports anything from ``oracle/`` and
raises if its CUDA library is missing.

What it restates (citations are into /root/reference/):
  * depth table + per-plane homographies ........ scripts/homography.py:23-75
  * kornia 0.6.3 ``warp_perspective`` (third-party, un-vendored, pinned in
    requirements.txt:1; single call site scripts/homography.py:85-86) .... SURVEY App. A.2
  * variance cost volume ........................ scripts/costvolume.py:7-16
  * CostVolumeReg forward (3D "U-Net" + softmax). scripts/model.py:70-126, :223-247
  * depth extraction ............................ scripts/depthmap.py:11-19

PARITY PINNING.  The reference has no test, golden vector or fixture that pins numeric
results for this path, and its warp arithmetic lives in kornia, which is absent
(=> at the kornia boundary: "parity unpinned" by reference-owned vectors).  The oracle is
instead pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the authoring
container by importing the unmodified /root/reference/scripts modules with the kornia
stand-in in oracle/kornia_shim (oracle/make_golden.py; vectors in tests/golden/).
tests/test_oracle_golden.py checks every function here against those vectors.

Two independent formulations are provided for the geometry:
  * ``homographies_chain32``  -- the reference's own fp32 torch op chain (bit-faithful)
  * ``homographies_closed64`` -- closed form in fp64 (rank-one / Sherman-Morrison), the
    formulation the CUDA kernel uses (19 numbers per view)
and two for the sampler: an explicit numpy bilinear gather and torch ``grid_sample``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

N_DEPTH_EST = 5  # scripts/config.py:9


# --------------------------------------------------------------------------------------
# a1  depth table                                            scripts/homography.py:23-26
# --------------------------------------------------------------------------------------
def depth_table(d_min: torch.Tensor, d_int: torch.Tensor, d_num: int, d_scale) -> torch.Tensor:
    """``d_batch_0[B,D,1,1] = d_min + D_SCALE*d_int*k`` in fp32, exactly as the reference."""
    k = torch.arange(d_num).reshape(1, d_num, 1, 1)
    return d_min + d_scale * d_int * k


def view_depth_rows(batch_size: int, n_views: int, bug_compatible: bool = True) -> np.ndarray:
    """Which row of ``d_batch_0`` flat view ``i = b*V+v`` uses.

    The reference tiles the table V times along dim 0 (scripts/homography.py:26), i.e. row
    order [b0,b1,..,b0,b1,..] while matrices are ordered b*V+v => view i reads row
    ``i mod B`` ("batch quirk").  ``bug_compatible=False`` gives the
    geometrically intended ``i div V``.
    """
    i = np.arange(batch_size * n_views)
    return (i % batch_size) if bug_compatible else (i // n_views)


# --------------------------------------------------------------------------------------
# a3  homographies                                           scripts/homography.py:40-75
# --------------------------------------------------------------------------------------
def homographies_chain32(K, R, T, d_batch_0, batch_size, n_views):
    """H_i[N,D,3,3] through the reference's own fp32 op order (for "table mode" parity)."""
    N = batch_size * n_views
    d_num = d_batch_0.shape[1]
    d_batch = torch.tile(d_batch_0, (n_views, 1, 1, 1))
    ref_idx = torch.arange(0, N, n_views).repeat_interleave(n_views)
    rep = lambda m: m.unsqueeze(1).repeat(1, d_num, 1, 1)
    eye = rep(torch.eye(3).unsqueeze(0))
    K_ref, R_ref0, T_ref0 = rep(K[ref_idx]), R[ref_idx], T[ref_idx]
    R_ref = rep(R_ref0)
    C_ref = rep(-torch.matmul(R_ref0.transpose(-2, -1), T_ref0))
    n_ref = R_ref[:, :, :, 2].unsqueeze(2)           # 3rd COLUMN of R_ref, as a 1x3 row (:49)
    K_v, R_v = rep(K), rep(R)
    C_v = rep(-torch.matmul(R.transpose(-2, -1), T))
    RK = torch.matmul(K_v, R_v)
    RK_ref = torch.matmul(R_ref.transpose(-2, -1), torch.inverse(K_ref))
    mid = eye - torch.matmul(C_v - C_ref, n_ref) / d_batch
    return torch.matmul(RK, torch.matmul(mid, RK_ref))


def view_params_closed64(K, R, T, batch_size, n_views, h, w):
    """Per-view closed form of the *sampling* map, fp64.

    H_i(d) = A - u w^T / d  with  A = K_i R_i R_ref^T K_ref^-1,  u = K_i R_i (C_i - C_ref),
    w^T = n R_ref^T K_ref^-1,  n = R_ref[:,2].  Sherman-Morrison:
        H_i(d)^-1 = A^-1 + g r^T / (d - s),   g = A^-1 u,  r^T = w^T A^-1,  s = w^T A^-1 u.
    kornia's align_corners mismatch (App. A.2) is a fixed affine map on the result:
        ix = px * w/(w-1) - 0.5,  iy = py * h/(h-1) - 0.5,
    folded here into rows 0/1 of A^-1 and g.  Returns dict of [N,...] float64 arrays:
    ``Ainv`` [N,3,3], ``g`` [N,3], ``r`` [N,3], ``s`` [N].
    """
    K = np.asarray(K, dtype=np.float64).reshape(-1, 3, 3)
    R = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
    T = np.asarray(T, dtype=np.float64).reshape(-1, 3, 1)
    N = batch_size * n_views
    out = dict(Ainv=np.zeros((N, 3, 3)), g=np.zeros((N, 3)), r=np.zeros((N, 3)), s=np.zeros(N))
    S = np.diag([w / (w - 1.0) if w > 1 else 1.0, h / (h - 1.0) if h > 1 else 1.0, 1.0])
    for i in range(N):
        j = (i // n_views) * n_views
        C_i = -R[i].T @ T[i]
        C_r = -R[j].T @ T[j]
        RrKr = R[j].T @ np.linalg.inv(K[j])
        A = K[i] @ R[i] @ RrKr
        u = K[i] @ R[i] @ (C_i - C_r)              # 3x1
        wT = R[j][:, 2].reshape(1, 3) @ RrKr        # 1x3
        Ainv = np.linalg.inv(A)
        g = Ainv @ u
        r = wT @ Ainv
        out["Ainv"][i] = S @ Ainv
        out["g"][i] = (S @ g).ravel()
        out["r"][i] = r.ravel()
        out["s"][i] = (wT @ Ainv @ u).item()
    return out


def sample_positions_closed64(params, depths_per_view, h, w):
    """(ix, iy)[N,D,h,w] fp64 from the closed form.  ``depths_per_view`` is [N,D]."""
    N, D = depths_per_view.shape
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    p = np.stack([xs, ys, np.ones_like(xs)], 0).reshape(3, -1)            # 3,hw
    ix = np.empty((N, D, h, w)); iy = np.empty((N, D, h, w))
    for i in range(N):
        a = params["Ainv"][i] @ p                                         # 3,hw
        c = params["r"][i] @ p                                            # hw
        dv = depths_per_view[i].astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            t = 1.0 / (dv - params["s"][i])                                   # D
        # the reference divides by d itself (homography.py:70): a plane at d == 0 is NaN for EVERY view
        t = np.where(dv == 0.0, np.nan, t)
        q = a[None] + params["g"][i][None, :, None] * (c[None, None, :] * t[:, None, None])
        ix[i] = (q[:, 0] / q[:, 2] - 0.5).reshape(D, h, w)
        iy[i] = (q[:, 1] / q[:, 2] - 0.5).reshape(D, h, w)
    return ix, iy


def sample_positions_from_H(H, h, w):
    """(ix, iy)[N,D,h,w] fp64 from explicit homographies H[N,D,3,3] (kornia inverts them)."""
    Hn = np.asarray(H, dtype=np.float64)
    N, D = Hn.shape[:2]
    Hinv = np.linalg.inv(Hn)
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    p = np.stack([xs, ys, np.ones_like(xs)], 0).reshape(3, -1)
    q = Hinv @ p                                                           # N,D,3,hw
    ix = (q[:, :, 0] / q[:, :, 2]) * (w / (w - 1.0)) - 0.5
    iy = (q[:, :, 1] / q[:, :, 2]) * (h / (h - 1.0)) - 0.5
    return ix.reshape(N, D, h, w), iy.reshape(N, D, h, w)


# --------------------------------------------------------------------------------------
# a4  bilinear warp, zero padding  (grid_sample(bilinear, zeros, align_corners=False))
# --------------------------------------------------------------------------------------
def bilinear_gather(feat: np.ndarray, ix: np.ndarray, iy: np.ndarray) -> np.ndarray:
    """feat[N,C,h,w], ix/iy[N,D,h,w] -> warped[N,C,D,h,w]; explicit 4-tap gather in fp64."""
    feat = np.asarray(feat, dtype=np.float64)
    N, C, h, w = feat.shape
    D = ix.shape[1]
    out = np.zeros((N, C, D, h, w))
    bad = ~(np.isfinite(ix) & np.isfinite(iy))                              # NaN positions -> NaN samples
    ix = np.where(bad, 0.0, ix); iy = np.where(bad, 0.0, iy)
    x0 = np.floor(ix); y0 = np.floor(iy)
    fx = ix - x0; fy = iy - y0
    x0 = x0.astype(np.int64); y0 = y0.astype(np.int64)
    for n in range(N):
        f = feat[n]
        for dy, dx, wt in ((0, 0, (1 - fx[n]) * (1 - fy[n])), (0, 1, fx[n] * (1 - fy[n])),
                           (1, 0, (1 - fx[n]) * fy[n]), (1, 1, fx[n] * fy[n])):
            xx = x0[n] + dx; yy = y0[n] + dy
            ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            v = f[:, np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]         # C,D,h,w
            out[n] += v * (wt * ok)[None]
        out[n][:, bad[n]] = np.nan
    return out


def bilinear_grid_sample(feat: torch.Tensor, ix: torch.Tensor, iy: torch.Tensor) -> torch.Tensor:
    """Same result through torch's CPU ``grid_sample`` (fp32), one call per plane like the reference."""
    N, C, h, w = feat.shape
    gx = (2.0 * ix + 1.0) / w - 1.0
    gy = (2.0 * iy + 1.0) / h - 1.0
    planes = []
    for d in range(ix.shape[1]):
        g = torch.stack((gx[:, d], gy[:, d]), -1).to(feat.dtype)
        planes.append(F.grid_sample(feat, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    return torch.stack(planes, 2)


def warp_reference_chain(feat, K, R, T, d_min, d_int, batch_size, n_views, d_num, d_scale):
    """Whole of scripts/homography.py:23-92 restated with torch CPU ops, fp32, reference op order."""
    d0 = depth_table(d_min, d_int, d_num, d_scale)
    H = homographies_chain32(K, R, T, d0, batch_size, n_views)
    h, w = feat.shape[-2:]
    # kornia: unit-normalise, invert in fp32, transform the unit grid, grid_sample (App. A.2)
    def unit(hh, ww):
        return torch.tensor([[2.0 / (ww - 1), 0, -1.0], [0, 2.0 / (hh - 1), -1.0], [0, 0, 1.0]])
    Nu = unit(h, w)
    xs = (torch.linspace(0, w - 1, w) / (w - 1) - 0.5) * 2
    ys = (torch.linspace(0, h - 1, h) / (h - 1) - 0.5) * 2
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    pts = torch.stack((gx, gy, torch.ones_like(gx)), -1).reshape(1, h * w, 3)
    planes = []
    for d in range(d_num):
        Mu = Nu @ (H[:, d] @ torch.inverse(Nu.unsqueeze(0)))
        q = torch.bmm(pts.expand(H.shape[0], -1, -1), torch.inverse(Mu).transpose(1, 2))
        z = q[..., 2:]
        sc = torch.where(z.abs() > 1e-8, 1.0 / (z + 1e-8), torch.ones_like(z))
        g = (sc * q[..., :2]).reshape(-1, h, w, 2)
        planes.append(F.grid_sample(feat, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    return torch.stack(planes, 2), d0


# --------------------------------------------------------------------------------------
# a5  variance cost volume                                   scripts/costvolume.py:7-16
# --------------------------------------------------------------------------------------
def variance_cost(warped, n_views: int):
    """Population variance over the V views (reference view included)."""
    bn, c, d, h, w = warped.shape
    x = warped.reshape(bn // n_views, n_views, c, d, h, w)
    mean = x.sum(1) / n_views
    if isinstance(x, np.ndarray):
        return ((x - mean[:, None]) ** 2).sum(1) / n_views
    return (x - mean.unsqueeze(1)).pow(2).sum(1) / n_views


def plane_sweep_cost(feat, K, R, T, d_min, d_int, batch_size, n_views, d_num, d_scale,
                     bug_compatible=True, sampler="numpy"):
    """features -> (cost[B,C,D,h,w], warped[N,C,D,h,w], depths[B,D]) via the closed form (fp64 geometry)."""
    N, C, h, w = feat.shape
    d0 = depth_table(d_min, d_int, d_num, d_scale).reshape(batch_size, d_num).numpy()
    rows = view_depth_rows(batch_size, n_views, bug_compatible)
    params = view_params_closed64(K, R, T, batch_size, n_views, h, w)
    ix, iy = sample_positions_closed64(params, d0[rows], h, w)
    if sampler == "numpy":
        warped = bilinear_gather(feat.detach().numpy(), ix, iy)
    else:
        warped = bilinear_grid_sample(feat, torch.from_numpy(ix), torch.from_numpy(iy))
    return variance_cost(warped, n_views), warped, d0


# --------------------------------------------------------------------------------------
# a6  CostVolumeReg                                          scripts/model.py:70-126
# --------------------------------------------------------------------------------------
REG_CONVS = {  # name: (cin, cout, stride, transposed)
    "conv_0_0": (32, 8, 1, False), "conv_1_0": (32, 16, 2, False), "conv_2_0": (32, 32, 2, False),
    "conv_3_0": (32, 64, 2, False), "conv_1_1": (16, 16, 1, False), "conv_2_1": (32, 32, 1, False),
    "conv_3_1": (64, 64, 1, False), "deconv_3_0": (64, 32, 2, True), "deconv_2_0": (32, 16, 2, True),
    "deconv_1_0": (16, 8, 2, True), "conv_out": (8, 1, 1, False),
}
REG_BN = {"BN_0": 8, "BN_1": 16, "BN_2": 32, "BN_3": 64}


def reg_pad(D, h, w):
    """PAD = floor(dim/2)+1, OUTPAD = (dim+1) mod 2            scripts/config.py:20-21"""
    return tuple(int(x) // 2 + 1 for x in (D, h, w)), tuple((int(x) + 1) % 2 for x in (D, h, w))


def reg_init(seed=0, dtype=torch.float32):
    """State dict with the reference's key names and PyTorch default init."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for name, (cin, cout, _, tr) in REG_CONVS.items():
        shape = (cin, cout, 3, 3, 3) if tr else (cout, cin, 3, 3, 3)
        fan_in = shape[1] * 27
        bound = 1.0 / fan_in ** 0.5           # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +)
        sd[name + ".weight"] = ((torch.rand(shape, generator=gen) * 2 - 1) * bound).to(dtype)
    for name, c in REG_BN.items():
        sd[name + ".weight"] = torch.ones(c, dtype=dtype); sd[name + ".bias"] = torch.zeros(c, dtype=dtype)
        sd[name + ".running_mean"] = torch.zeros(c, dtype=dtype); sd[name + ".running_var"] = torch.ones(c, dtype=dtype)
        sd[name + ".num_batches_tracked"] = torch.tensor(0)
    return sd


def reg_forward(sd, cv, train_bn=True, update_running=False, eps=1e-5, momentum=0.1, return_logits=False):
    """CostVolumeReg.forward (scripts/model.py:100-126) with functional torch CPU ops.

    BN modules are shared between layers (BN_0 x2, BN_1 x3, BN_2 x3, BN_3 x2); in train mode the
    statistics are over (B,D,h,w) including the exact-zero regions the stride-2 convs leave.
    """
    D, h, w = cv.shape[-3:]
    pad, outpad = reg_pad(D, h, w)

    def conv(name, x):
        cin, cout, s, tr = REG_CONVS[name]
        wgt = sd[name + ".weight"]
        if tr:
            return F.conv_transpose3d(x, wgt, stride=2, padding=pad, output_padding=outpad)
        return F.conv3d(x, wgt, stride=s, padding=(1 if s == 1 else pad))

    def bn_relu(name, x):
        rm, rv = sd[name + ".running_mean"], sd[name + ".running_var"]
        if update_running and train_bn:
            sd[name + ".num_batches_tracked"] += 1
        y = F.batch_norm(x, rm if (update_running or not train_bn) else None,
                         rv if (update_running or not train_bn) else None,
                         sd[name + ".weight"], sd[name + ".bias"], train_bn, momentum, eps)
        return F.relu(y)

    y0 = bn_relu("BN_0", conv("conv_0_0", cv))
    y1 = bn_relu("BN_1", conv("conv_1_0", cv))
    y2 = bn_relu("BN_2", conv("conv_2_0", cv))
    y3 = bn_relu("BN_3", conv("conv_3_0", cv))
    y1 = bn_relu("BN_1", conv("conv_1_1", y1))
    y2 = bn_relu("BN_2", conv("conv_2_1", y2))
    y3 = bn_relu("BN_3", conv("conv_3_1", y3))
    y3 = bn_relu("BN_2", conv("deconv_3_0", y3))
    y2 = bn_relu("BN_1", conv("deconv_2_0", y3 + y2))
    y1 = bn_relu("BN_0", conv("deconv_1_0", y2 + y1))
    logits = conv("conv_out", y1 + y0)
    prob = torch.softmax(logits, dim=2)
    return (prob, logits) if return_logits else prob


# --------------------------------------------------------------------------------------
# a7  depth extraction                                       scripts/depthmap.py:11-19
# --------------------------------------------------------------------------------------
def kept_planes(prob: np.ndarray, n_est: int = N_DEPTH_EST) -> np.ndarray:
    """ranks[B,n_est,h,w]: position of plane j (j < n_est) in the *stable* descending sort over D.

    The reference masks the UNSORTED volume with ``argsort_desc(P) < 5`` (depthmap.py:11-15), so the
    planes it keeps are exactly {rank(j): j = 0..4}.
    """
    P = np.asarray(prob)[:, 0]                                           # B,D,h,w
    B, D, h, w = P.shape
    ranks = np.empty((B, n_est, h, w), dtype=np.int64)
    for j in range(min(n_est, D)):
        gt = (P > P[:, j:j + 1]).sum(1)
        eq_before = (P[:, :j] == P[:, j:j + 1]).sum(1)
        ranks[:, j] = gt + eq_before
    return ranks[:, :min(n_est, D)]


def extract_depth(prob, d_batch, n_est: int = N_DEPTH_EST):
    """prob[B,1,D,h,w], d_batch[B,D,1,1] -> depth[B,1,h,w] (fp64 accumulate) and the kept-plane ranks."""
    P = np.asarray(prob, dtype=np.float64)[:, 0]
    d = np.asarray(d_batch, dtype=np.float64).reshape(P.shape[0], P.shape[1])
    ranks = kept_planes(np.asarray(prob), n_est)
    pk = np.take_along_axis(P, ranks, axis=1)                            # B,5,h,w
    dk = np.take_along_axis(np.broadcast_to(d[:, :, None, None], P.shape), ranks, axis=1)
    return ((dk * pk).sum(1) / pk.sum(1))[:, None], ranks


def extract_depth_torch(prob: torch.Tensor, d_batch: torch.Tensor, n_est: int = N_DEPTH_EST):
    """Literal restatement of scripts/depthmap.py:11-19 (differentiable; uses torch.sort)."""
    _, order = prob.sort(dim=2, descending=True, stable=True)
    mask = torch.less(order, n_est).float()
    filt = prob * mask
    return (d_batch.unsqueeze(1) * filt).sum(2).squeeze(2).div(filt.sum(2))


def tie_pixels(prob: np.ndarray, n_est: int = N_DEPTH_EST) -> np.ndarray:
    """Mask[B,h,w] of pixels whose kept set depends on sort tie-breaking (reference result unpinned there)."""
    P = np.asarray(prob)[:, 0]
    B, D, h, w = P.shape
    bad = np.zeros((B, h, w), dtype=bool)
    for j in range(min(n_est, D)):
        bad |= ((P == P[:, j:j + 1]).sum(1) > 1)
    return bad


# --------------------------------------------------------------------------------------
# fixtures                                                   SURVEY App. C (DTU cameras)
# --------------------------------------------------------------------------------------
DTU_K = [[361.54126, 0.0, 82.90063], [0.0, 360.3975, 66.38387], [0.0, 0.0, 1.0]]
DTU_R = [
    [[0.970263, 0.00748, 0.241939], [-0.014743, 0.999493, 0.028223], [-0.241605, -0.030951, 0.969881]],
    [[0.885052, -0.307962, 0.34906], [0.220575, 0.937798, 0.268109], [-0.409915, -0.160296, 0.897928]],
    [[0.802256, -0.439347, 0.404178], [0.427993, 0.895282, 0.123659], [-0.416183, 0.073779, 0.906283]],
]
DTU_T = [[-191.02, 3.28832, 22.5401], [-258.497, -156.493, 71.838], [-291.419, -77.0495, 71.2762]]


def _rot(axis, ang):
    axis = np.asarray(axis, dtype=np.float64); axis /= np.linalg.norm(axis)
    Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx


def smooth_features_exact(n, c, h, w, seed):
    """Seeded conv-like (smooth) feature maps that regenerate BIT-IDENTICALLY on any host: CPU-generator white noise through a
    3x3 box blur written as nine shifted adds in a fixed order (elementwise IEEE fp32 -- no library convolution whose summation
    order could differ between CPUs).  Full-size fixtures (tests/golden/cfg1_digest.npz) store only digests of the reference's
    outputs; the inputs are rebuilt from the seed by this function and verified against a stored checksum."""
    gen = torch.Generator().manual_seed(int(seed))
    x = torch.randn(n, c, h + 2, w + 2, generator=gen)
    acc = torch.zeros(n, c, h, w)
    for dy in range(3):
        for dx in range(3):
            acc = acc + x[:, :, dy:dy + h, dx:dx + w]
    return acc * (1.0 / 3.0)


def rank_margin(prob: np.ndarray, n_est: int = N_DEPTH_EST) -> np.ndarray:
    """[B,h,w]: smallest relative gap between the probability of any plane j < n_est and any OTHER plane of the pixel.  The kept
    set of scripts/depthmap.py:11-15 is {rank(j)}: a perturbation of every probability by less than margin/2 (relative) cannot
    change it.  Parity tests of reduced-precision paths compare depth only where the reference's kept set is that stable."""
    P = np.asarray(prob, dtype=np.float64)[:, 0]
    B, D, h, w = P.shape
    m = np.full((B, h, w), np.inf)
    for j in range(min(n_est, D)):
        gap = np.abs(P - P[:, j:j + 1])
        gap[:, j] = np.inf
        m = np.minimum(m, gap.min(1) / np.maximum(P[:, j], 1e-300))
    return m


def synthetic_cameras(batch_size, n_views, h=128, w=160, seed=0):
    """DTU-shaped cameras: views 0..2 are the real DTU cams of SURVEY App. C; further views are small
    seeded perturbations of them (the 49-camera table lives in a pickle that does not travel).
    K is scaled from the 160x128 feature grid to (w,h).  Returns CPU fp32 K[N,3,3], R[N,3,3], T[N,3,1]."""
    rng = np.random.RandomState(seed)
    Ks, Rs, Ts = [], [], []
    sx, sy = w / 160.0, h / 128.0
    K = np.array(DTU_K); K[0] *= sx; K[1] *= sy
    for b in range(batch_size):
        for v in range(n_views):
            R = np.array(DTU_R[v % 3]); T = np.array(DTU_T[v % 3])
            if v >= 3 or b > 0:
                dR = _rot(rng.randn(3), 0.04 * rng.randn())
                R = dR @ R
                T = T + rng.randn(3) * 8.0
            Ks.append(K); Rs.append(R); Ts.append(T.reshape(3, 1))
    f = lambda a: torch.tensor(np.stack(a), dtype=torch.float32)
    return f(Ks), f(Rs), f(Ts)
