"""Depth-slab sharding of ONE plane sweep across the GPUs of a box (BASELINE.json configs[3]; SURVEY §8e, second bullet).

One process per GPU.  Rank r owns the canvas planes [a_r, b_r) of every full-size tensor of the regulariser
(scripts/model.py:100-126) and the box planes [ja_r, jb_r) of every central-box tensor (regulariser.py: the stride-2
layers of the reference only carry data on a central box, plane j of the box being centred on canvas plane 2j - L + 1).

  * K1 (warp + variance) shards trivially by plane: every rank has all V feature maps and sweeps its own planes PLUS the
    one or two halo planes its convolutions read -- recomputing a halo plane costs 1/(b_r - a_r) of the slab, exchanging
    it would cost a 32-channel plane over NVLink and a synchronisation.
  * stride-1 convolutions (conv_0_0, conv_{1,2,3}_1, conv_out) run on a slab extended by one plane per side with their
    ordinary padding; the two outermost output planes of an interior edge are wrong by construction and are dropped.  The
    halo planes of the small tensors (box planes of the stride-2 branches, the 8-channel canvas in front of conv_out) are
    exchanged point to point (`reslab`).
  * a stride-2 convolution reads canvas planes 2j - L + k, a stride-2 transposed convolution writes canvas planes
    2j - L + k: with the box partition derived from the canvas partition both are LOCAL up to one halo plane; what a
    rank computes of a transposed convolution's canvas that lies inside the central box is handed to the box owners
    (`reslab` again: a fixed pattern, up to three peers).
  * train-mode BatchNorm (the reference keeps it on at test time, test.py:61) needs the per-channel sums over the whole
    canvas: one all-reduce of 2·C doubles per BatchNorm; every rank then applies the same affine map and performs the
    same running-statistics update, so the replicas of the 382 k parameters / buffers stay identical without a broadcast.
  * the depth softmax and the rank-based depth extraction need all D planes of a pixel: the 1-channel logits are
    re-sharded from plane slabs to row slabs (`reshard_rows`, all-to-all pattern), K4 runs on rows, the depth rows are
    all-gathered.

All exchanges are `torch.distributed` point-to-point batches / all-reduces on the default stream (NCCL over NVLink on the
box, gloo in tests/test_depth_slab_gloo.py).  The per-slab compute is the same set of libmvs_b200.so kernels the
single-GPU module uses.  Forward only (inference); training shards by scene/batch instead (harness.FlatGradAllReduce).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib, ops
from . import conv3d as conv_backends
from .regulariser import CostVolumeReg, central_region

_CL = torch.channels_last_3d


# --------------------------------------------------------------------------------------------------
# who owns which planes
# --------------------------------------------------------------------------------------------------
class SlabPlan:
    """Pure arithmetic (no communication): the plane ranges every rank owns and needs, as half-open (lo, hi) pairs."""

    def __init__(self, D: int, world: int):
        if world < 1 or D < 2 * world:
            raise ValueError(f"cannot split D = {D} planes over {world} ranks (need at least two planes per rank)")
        self.D, self.R = int(D), int(world)
        self.lo, self.hi, self.L = central_region(D)
        self.nC = self.hi - self.lo + 1
        cuts = [0] + [2 * int(round(r * D / (2.0 * world))) for r in range(1, world)] + [D]
        if any(b - a < 2 for a, b in zip(cuts, cuts[1:])):
            raise ValueError(f"depth slabs of D = {D} over {world} ranks would be thinner than two planes")
        self.canvas = [(cuts[r], cuts[r + 1]) for r in range(world)]
        # box plane j belongs to the rank that owns its central input plane 2j - L + 1
        jc = [0] + [min(self.nC, max(0, -(-(cuts[r] + self.L - 1) // 2))) for r in range(1, world)] + [self.nC]
        if any(b - a < 1 for a, b in zip(jc, jc[1:])):
            raise ValueError(f"too many ranks ({world}) for the {self.nC}-plane central box of D = {D}")
        self.box = [(jc[r], jc[r + 1]) for r in range(world)]
        # E = the box dilated by one plane where the canvas allows (output range of conv_k_1 that sees real data)
        self.e0 = -1 if self.lo >= 1 else 0
        self.e1 = self.nC + 1 if self.hi + 1 <= D - 1 else self.nC
        self.tbox = [(self.e0 if r == 0 else ja, self.e1 if r == world - 1 else jb) for r, (ja, jb) in enumerate(self.box)]

    # ---- per-rank derived ranges ---------------------------------------------------------------
    def s2_input(self, r):
        """(c0, c1, garbage, n_out): canvas planes the stride-2 branch convolution of rank r reads, how many leading output
        planes are wrong by construction (0 or 1) and how many box planes the call produces."""
        ja, jb = self.box[r]
        g = 0 if ja == 0 else 1
        c0 = 2 * (ja - g)
        c1 = min(self.D, 2 * (jb - 1) - self.L + 3)
        return c0, c1, g, jb - ja + g

    def conv0_input(self, r):
        a, b = self.canvas[r]
        return max(0, a - 1), min(self.D, b + 1)

    def cost_planes(self, r):
        """Canvas planes of the cost volume rank r sweeps (its own + halo)."""
        c0, c1, _, _ = self.s2_input(r)
        d0, d1 = self.conv0_input(r)
        return min(c0, d0), max(c1, d1)

    def s_halo(self, r):
        """Box planes of S = conv_k_0(cv) needed to produce T = conv_k_1(..) on tbox[r]."""
        ta, tb = self.tbox[r]
        return max(0, ta - 1), min(self.nC, tb + 1)

    def x_range(self, r):
        """(xa, xb, zlo, zhi): box-frame planes of the conv_k_1 input that exist on the canvas, and how many zero planes
        (the convolution's own padding at the canvas border) go in front of / behind them."""
        ta, tb = self.tbox[r]
        xa, xb = max(ta - 1, -self.lo), min(tb + 1, self.D - self.lo)
        return xa, xb, xa - (ta - 1), (tb + 1) - xb

    def up_input(self, r):
        """(ua, ub, L_loc): box planes a transposed convolution needs for canvas planes canvas[r], and the padding of the
        equivalent local transposed convolution (out o - a = 2 (j - ua) - L_loc + k)."""
        a, b = self.canvas[r]
        ua = max(0, (self.L + a - 1) // 2)
        ub = min(self.nC, (b - 1 + self.L) // 2 + 1)
        return ua, ub, self.L + a - 2 * ua

    def box_piece(self, r):
        """Box-frame planes of the central box that lie in rank r's canvas slab (possibly empty)."""
        a, b = self.canvas[r]
        clamp = lambda v: min(self.nC, max(0, v - self.lo))
        return clamp(a), clamp(b)                            # monotone over ranks: an ordered partition of [0, nC)


def row_partition(h: int, world: int):
    cuts = [int(round(r * h / float(world))) for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


# --------------------------------------------------------------------------------------------------
# exchanges
# --------------------------------------------------------------------------------------------------
class TorchDistComm:
    """The exchanges the slab path needs, on torch.distributed (NCCL over NVLink on the box, gloo in the CPU tests).
    tests/test_gpu_depth_slab.py substitutes an in-process implementation to drive R slabs on one GPU."""

    def __init__(self, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def _peer(self, r):
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def exchange(self, sends, recvs):
        """sends: [(peer, contiguous tensor)], recvs: [(peer, contiguous buffer)] -- at most one message per ordered pair;
        one batched isend/irecv (a single NCCL group call: no ordering deadlocks)."""
        p2p = [dist.P2POp(dist.irecv, buf, self._peer(s), self.group) for s, buf in recvs]
        p2p += [dist.P2POp(dist.isend, t, self._peer(d), self.group) for d, t in sends]
        if p2p:
            for req in dist.batch_isend_irecv(p2p):
                req.wait()

    def all_reduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather(self, t):
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t, group=self.group)
        return parts

    def broadcast(self, t, src):
        dist.broadcast(t, self._peer(src), group=self.group)
        return t


def reslab(x: torch.Tensor, have, want, comm, fill: float = 0.0) -> torch.Tensor:
    """Re-distribute a plane-sharded [B,C,n,h,w] tensor: this rank holds the global planes have[rank] in `x` and gets
    back the global planes want[rank].  `have` must be an ordered partition of a contiguous plane range (every rank
    evaluates the same lists); planes of `want` nobody owns are set to `fill`.  Halo exchange, box hand-over and
    identity are all instances.  One batched isend/irecv; zero-copy for B = 1 (a plane range of a channel-last volume is
    one contiguous chunk)."""
    rank = comm.rank
    h0, h1 = have[rank]
    w0, w1 = want[rank]
    B, C, n, h, w = x.shape
    if n != h1 - h0:
        raise ValueError(f"reslab: tensor holds {n} planes, plan says {h1 - h0}")
    x = x.contiguous(memory_format=_CL) if n > 0 else x
    out = torch.empty((B, w1 - w0, h, w, C), dtype=x.dtype, device=x.device)         # memory order of channels_last_3d
    xv = x.permute(0, 2, 3, 4, 1)
    H0, H1 = have[0][0], have[-1][1]
    for lo_, hi_ in ((w0, min(w1, H0)), (max(w0, H1), w1)):
        if hi_ > lo_:
            out[:, lo_ - w0:hi_ - w0].fill_(fill)
    sends, recvs, staged = [], [], []
    for s, (s0, s1) in enumerate(have):                   # what I receive
        lo_, hi_ = max(s0, w0), min(s1, w1)
        if hi_ <= lo_:
            continue
        dst = out[:, lo_ - w0:hi_ - w0]
        if s == rank:
            dst.copy_(xv[:, lo_ - h0:hi_ - h0])
            continue
        buf = dst if dst.is_contiguous() else torch.empty(dst.shape, dtype=x.dtype, device=x.device)
        if buf is not dst:
            staged.append((buf, dst))
        recvs.append((s, buf))
    for d, (d0, d1) in enumerate(want):                   # what I send
        if d == rank:
            continue
        lo_, hi_ = max(h0, d0), min(h1, d1)
        if hi_ <= lo_:
            continue
        src = xv[:, lo_ - h0:hi_ - h0]
        sends.append((d, src if src.is_contiguous() else src.contiguous()))
    comm.exchange(sends, recvs)
    for buf, dst in staged:
        dst.copy_(buf)
    return out.permute(0, 4, 1, 2, 3)                     # [B,C,n',h,w] with channels_last_3d strides


def reshard_rows(x: torch.Tensor, planes, rows, comm) -> torch.Tensor:
    """[B,1,n_own,h,w] plane slab of a 1-channel volume -> [B,1,D,rows_own,w] row slab holding ALL planes (all-to-all
    pattern as one batched isend/irecv: gloo has no all_to_all, and the pieces are uneven anyway)."""
    rank = comm.rank
    B, C, n, h, w = x.shape
    a, b = planes[rank]
    if n != b - a or C != 1:
        raise ValueError("reshard_rows: expects the rank's own 1-channel plane slab")
    r0, r1 = rows[rank]
    D = planes[-1][1]
    out = torch.empty((B, 1, D, r1 - r0, w), dtype=x.dtype, device=x.device)
    sends, recvs, staged = [], [], []
    for s, (s0, s1) in enumerate(planes):
        if s == rank:
            out[:, :, s0:s1].copy_(x[:, :, :, r0:r1])
            continue
        if s1 > s0 and r1 > r0:
            buf = torch.empty((B, 1, s1 - s0, r1 - r0, w), dtype=x.dtype, device=x.device)
            staged.append((buf, out[:, :, s0:s1]))
            recvs.append((s, buf))
    for d, (d0, d1) in enumerate(rows):
        if d != rank and d1 > d0 and n > 0:
            sends.append((d, x[:, :, :, d0:d1].contiguous()))
    comm.exchange(sends, recvs)
    for buf, dst in staged:
        dst.copy_(buf)
    return out


def gather_rows(x_rows: torch.Tensor, rows, comm) -> torch.Tensor:
    """[B,1,rows_own,w] on every rank -> the full [B,1,h,w] map on every rank (padded all_gather)."""
    B, C, _, w = x_rows.shape
    hmax = max(b - a for a, b in rows)
    pad = torch.zeros((B, C, hmax, w), dtype=x_rows.dtype, device=x_rows.device)
    pad[:, :, :x_rows.shape[2]] = x_rows
    parts = comm.all_gather(pad)
    return torch.cat([p[:, :, :b - a] for p, (a, b) in zip(parts, rows)], 2)


# --------------------------------------------------------------------------------------------------
# per-slab arithmetic: kernels of libmvs_b200.so on the GPU; the torch expressions serve the CPU (gloo) tests of the
# slab algebra, exactly like regulariser.py's CPU branch -- the product entry point refuses CPU tensors
# --------------------------------------------------------------------------------------------------
def _sums(x: torch.Tensor) -> torch.Tensor:
    """[2, C] fp64: per-channel (sum x, sum x^2) over a [B,C,D,h,w] box view."""
    C = x.shape[1]
    if x.numel() == 0:
        return torch.zeros(2, C, dtype=torch.float64, device=x.device)
    if x.is_cuda:
        s1, s2 = ops.channel_sums(x)
        return torch.stack([s1.double(), s2.double()])
    xd = x.double()
    return torch.stack([xd.sum((0, 2, 3, 4)), (xd * xd).sum((0, 2, 3, 4))])


def _affine_geo(x, scale, shift, in_origin, out_origin, out_dims):
    """y = relu(xv*scale + shift) on the output box; xv = x inside its box (origin in_origin), 0 outside (one frame)."""
    B, C = x.shape[:2]
    if min(out_dims) <= 0:
        return torch.empty((B, C) + tuple(max(0, n) for n in out_dims), dtype=x.dtype, device=x.device)
    if x.is_cuda and x.numel() > 0:
        return ops.affine_relu_geo(x, scale, shift, in_origin, out_origin, out_dims)
    bv = lambda v: v.to(torch.float32).view(1, -1, 1, 1, 1)
    y = F.relu(bv(shift)).expand(B, C, *out_dims).clone()
    src, dst = [], []
    for ax in range(3):
        lo_ = max(in_origin[ax], out_origin[ax])
        hi_ = min(in_origin[ax] + x.shape[2 + ax], out_origin[ax] + out_dims[ax])
        if hi_ <= lo_:
            return y.to(x.dtype)
        src.append(slice(lo_ - in_origin[ax], hi_ - in_origin[ax]))
        dst.append(slice(lo_ - out_origin[ax], hi_ - out_origin[ax]))
    ix = (slice(None), slice(None))
    y[ix + tuple(dst)] = F.relu(x[ix + tuple(src)].float() * bv(scale) + bv(shift))
    return y.to(x.dtype)


def _s2_box(be, x, w_cat, pads, out_dims):
    """out(o) = sum_k W[k] x(2o - pads + k), o in [0, out_dims), zero outside x."""
    if hasattr(be, "conv3d_s2_box"):
        y = be.conv3d_s2_box(x, w_cat, pads, out_dims)
        if y is not None:
            return y
    P = tuple(q if q >= 2 else q + 2 for q in pads)        # symmetric padding of the same parity
    nat = tuple((n + 2 * p - 3) // 2 + 1 for n, p in zip(x.shape[2:], P))
    off = tuple((p - q) // 2 for p, q in zip(P, pads))
    extra = [max(0, o + n - m) for o, n, m in zip(off, out_dims, nat)]               # outputs past the natural extent
    if any(extra):
        x = F.pad(x, (0, 2 * extra[2], 0, 2 * extra[1], 0, 2 * extra[0]))
    y = be.conv3d(x, w_cat, 2, P)
    return y[(slice(None), slice(None)) + tuple(slice(o, o + n) for o, n in zip(off, out_dims))]


# --------------------------------------------------------------------------------------------------
# the sharded regulariser
# --------------------------------------------------------------------------------------------------
class DepthSlabCostVolumeReg:
    """Runs a CostVolumeReg (same Parameter / buffer objects, replicated on every rank) on depth slabs.

    slab_logits(cost_fn, B, D, h, w) -> this rank's planes of the logits, [B,1,b_r-a_r,h,w] fp32
        cost_fn(c0, c1) returns the cost-volume planes [c0, c1) as [B,32,c1-c0,h,w] (K1 on the rank's own planes + halo).
    forward(cost_fn, d_batch, B, D, h, w) -> (depth [B,1,h,w] on every rank, prob rows [B,1,D,rows,w], (row0, row1))
    """

    def __init__(self, reg: CostVolumeReg, comm=None):
        self.reg = reg
        self.comm = TorchDistComm() if comm is None else comm
        self.rank, self.world = self.comm.rank, self.comm.world

    # ---- helpers ----
    def plan(self, D):
        return SlabPlan(D, self.world)

    def _allsum(self, t):
        return self.comm.all_reduce_sum(t)

    def _moments(self, s, n_full):
        mean = s[0] / n_full
        var = (s[1] / n_full - mean * mean).clamp_min(0)
        return mean.float(), var.float()

    @torch.no_grad()
    def slab_logits(self, cost_fn, B, D, h, w, be=None):
        reg, r, R = self.reg, self.rank, self.world
        if be is None:        # as CostVolumeReg.forward: precision="fp32" owns its precision (TF32 off inside the calls)
            be = conv_backends.ExactTorchConvBackend if reg.precision == "fp32" else conv_backends.get(reg.conv_backend)
        plan = SlabPlan(D, R)
        dt = torch.bfloat16 if reg.precision == "bf16" else torch.float32
        train = reg.BN_0.training
        n_full = B * D * h * w
        a, b = plan.canvas[r]
        ja, jb = plan.box[r]
        ta, tb = plan.tbox[r]
        rg = [(plan.lo, plan.hi, plan.L), central_region(h), central_region(w)]      # per-axis box geometry
        dims = (D, h, w)
        # a backend that packs its own filters takes the fp32 parameters as they are (as CostVolumeReg.logits does)
        wdt = torch.float32 if getattr(be, "fp32_weights", False) else dt
        W_ = lambda name: reg._w(name, dt if name == "conv_out" else wdt)

        # ---- cost-volume planes of this rank (own + halo), swept locally
        k0, k1 = plan.cost_planes(r)
        cv = cost_fn(k0, k1)
        cv = cv if cv.dtype == dt else cv.to(dt)
        cv = cv.contiguous(memory_format=_CL)

        # ---- conv_0_0 on the canvas slab and the three stride-2 branches on box planes [ja, jb), both from the cost volume
        d0, d1 = plan.conv0_input(r)
        y0_raw = be.conv3d(cv[:, :, d0 - k0:d1 - k0], W_("conv_0_0"), 1, (1, 1, 1))
        own = y0_raw[:, :, a - d0:b - d0]
        c0, c1, g, n_out = plan.s2_input(r)
        w_cat = torch.cat([W_(f"conv_{k}_0") for k in (1, 2, 3)], 0)
        widths = [reg.conv_1_0.out_channels, reg.conv_2_0.out_channels, reg.conv_3_0.out_channels]
        S_all = _s2_box(be, cv[:, :, c0 - k0:c1 - k0], w_cat, tuple(L for _, _, L in rg),
                        (n_out,) + tuple(hi - lo + 1 for lo, hi, _ in rg[1:]))[:, :, g:]
        del cv
        S_parts = torch.split(S_all, widths, 1)
        # ONE all-reduce for the statistics of BN_0(conv_0_0) and of the three branch outputs (they are independent)
        if train:
            st = self._allsum(torch.cat([_sums(own)] + [_sums(S) for S in S_parts], 1))
            st = torch.split(st, [own.shape[1]] + widths, 1)
        mean, var = self._moments(st[0], n_full) if train else (None, None)
        scale, shift = reg._bn_affine(reg.BN_0, mean, var, n_full)
        y0 = _affine_geo(y0_raw, scale, shift, (d0, 0, 0), (a, 0, 0), (b - a, h, w))

        # in-plane geometry exactly as regulariser.py: C (box), E = C+1 ring, F = C+2 ring, clipped to the canvas
        C_lo = [lo for lo, _, _ in rg]
        C_dims = [hi - lo + 1 for lo, hi, _ in rg]
        E_lo = [max(0, lo - 1) for lo, _, _ in rg]
        E_hi = [min(n - 1, hi + 1) for (_, hi, _), n in zip(rg, dims)]
        F_lo = [max(0, e - 1) for e in E_lo]
        F_hi = [min(n - 1, e + 1) for e, n in zip(E_hi, dims)]
        sa, sb = plan.s_halo(r)
        xa, xb, zlo, zhi = plan.x_range(r)
        zpad = [1 if E_lo[2] == 0 else 0, 1 if E_hi[2] == w - 1 else 0,
                1 if E_lo[1] == 0 else 0, 1 if E_hi[1] == h - 1 else 0, zlo, zhi]
        # ONE halo exchange for the three branches (they are channel groups of one tensor)
        S_ext_parts = torch.split(reslab(S_all, plan.box, [plan.s_halo(q) for q in range(R)], self.comm), widths, 1)
        bns = (reg.BN_1, reg.BN_2, reg.BN_3)
        Ts, bgs = [], []
        for idx, (k, bn) in enumerate(zip((1, 2, 3), bns)):
            mean, var = self._moments(st[1 + idx], n_full) if train else (None, None)
            scale, shift = reg._bn_affine(bn, mean, var, n_full)
            bgs.append(F.relu(shift))
            X = _affine_geo(S_ext_parts[idx], scale, shift, (plan.lo + sa, C_lo[1], C_lo[2]), (plan.lo + xa, F_lo[1], F_lo[2]),
                            (xb - xa, F_hi[1] - F_lo[1] + 1, F_hi[2] - F_lo[2] + 1))
            if any(zpad):
                X = F.pad(X, zpad)
            X = (X if X.dtype == dt else X.to(dt)).contiguous(memory_format=_CL)
            Ts.append(be.conv3d(X, W_(f"conv_{k}_1"), 1, (0, 0, 0)))                   # planes [ta, tb) x E_h x E_w
        if train:                                                                     # ONE all-reduce for the three conv_k_1 outputs
            tt = torch.split(self._allsum(torch.cat([_sums(T) for T in Ts], 1)), widths, 1)
        enc = {}
        for idx, (k, bn) in enumerate(zip((1, 2, 3), bns)):
            mean = var = None
            if train:
                mean, var = reg._stats_from_sums_with_constant_outside(tt[idx][0], tt[idx][1], W_(f"conv_{k}_1").float(), bgs[idx],
                                                                       dims, E_lo, E_hi, B, n_full)
            scale, shift = reg._bn_affine(bn, mean, var, n_full)
            enc[k] = _affine_geo(Ts[idx], scale, shift, (plan.lo + ta, E_lo[1], E_lo[2]), (plan.lo + ja, C_lo[1], C_lo[2]),
                                 (jb - ja, C_dims[1], C_dims[2]))
        del Ts

        # ---- decoder: transposed convolutions from box planes to canvas planes
        Lhw = (rg[1][2], rg[2][2])
        up_want = [plan.up_input(q)[:2] for q in range(R)]
        pieces = [plan.box_piece(q) for q in range(R)]

        def up(z, name, bn, to_box):
            ua, ub, L_loc = plan.up_input(r)
            z_ext = reslab(z.to(dt), plan.box, up_want, self.comm).contiguous(memory_format=_CL)
            U = be.conv_transpose3d_alloc(z_ext, W_(name), 2, (L_loc,) + Lhw, (b - a, h, w))
            mean, var = self._moments(self._allsum(_sums(U[:, :, :b - a, :h, :w])), n_full) if train else (None, None)
            scale, shift = reg._bn_affine(bn, mean, var, n_full)
            if not to_box:
                return _affine_geo(U, scale, shift, (0, 0, 0), (0, 0, 0), (b - a, h, w))
            pa, pb = pieces[r]
            Uc = U[:, :, :b - a, :h, :w]
            piece = _affine_geo(Uc, scale, shift, (a, 0, 0), (plan.lo + pa, C_lo[1], C_lo[2]), (pb - pa, C_dims[1], C_dims[2]))
            return reslab(piece, pieces, plan.box, self.comm)

        c3 = up(enc[3], "deconv_3_0", reg.BN_2, True)
        c2 = up(c3 + enc[2].to(c3.dtype), "deconv_2_0", reg.BN_1, True)
        y1 = up(c2 + enc[1].to(c2.dtype), "deconv_1_0", reg.BN_0, False)
        z = (y1 + y0).contiguous(memory_format=_CL)

        # ---- conv_out on the slab + one halo plane per side
        z_ext = reslab(z, plan.canvas, [plan.conv0_input(q) for q in range(R)], self.comm)
        if z_ext.is_cuda and dt == torch.bfloat16 and z_ext.shape[1] == 8 and reg.conv_out.out_channels == 1:
            lg = ops.conv_out(z_ext, reg.conv_out.weight)
        else:
            lg = be.conv3d(z_ext, W_("conv_out"), 1, (1, 1, 1)).float()
        return lg[:, :, a - d0:b - d0].contiguous()

    @torch.no_grad()
    def forward(self, cost_fn, d_batch, B, D, h, w, be=None):
        logits = self.slab_logits(cost_fn, B, D, h, w, be)
        if not logits.is_cuda:
            raise _lib.MvsB200Error("DepthSlabCostVolumeReg.forward needs CUDA tensors; mvs_b200 has no CPU path")
        planes, rows = SlabPlan(D, self.world).canvas, row_partition(h, self.world)
        self.last_logits = logits                                 # this rank's planes (parity checks)
        lr = reshard_rows(logits, planes, rows, self.comm)
        prob_rows, depth_rows = ops.softmax_depth(lr, d_batch, self.reg.n_depth_est)
        return gather_rows(depth_rows, rows, self.comm), prob_rows, rows[self.rank]


class GraphedSlabForward:
    """A rank's whole sharded forward -- K1 on its planes, the regulariser slabs, every NCCL exchange, K4, the depth
    all-gather -- captured once and replayed as ONE CUDA graph.  At 4-8 ranks a slab's ~250 kernels last tens of
    microseconds each; issued eagerly the pass is bound by the host (8 GPUs at cfg4: 14.2 ms eager, 4.15 ms replayed; 1 GPU
    unsharded: 12.2 ms -- profiles/r01_scaling.md).  Inputs live in static buffers: `features` is copied into the graph's own
    feature tensor, cameras / depth range are re-targeted with `sweep.update(...)`.  Call `release()` before tearing the
    process group down (a live graph that holds NCCL work blocks the communicator's destruction)."""

    def __init__(self, sharded: "DepthSlabCostVolumeReg", sweep, feature_shape, device, out_dtype=torch.bfloat16, warmup=2):
        self.sharded, self.sweep, self.out_dtype, self.warmup = sharded, sweep, out_dtype, warmup
        N, C, h, w = feature_shape
        self.features = torch.zeros((N, h, w, C), dtype=torch.float32, device=device).permute(0, 3, 1, 2)   # channel-last
        self.graph, self.out = None, None

    def _run(self):
        sw = self.sweep
        return self.sharded.forward(slab_cost_fn(self.features, sw, self.out_dtype), sw.d_batch_dev, sw.B, sw.D, sw.h, sw.w)

    def __call__(self, features):
        self.features.copy_(features)
        if self.graph is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(self.warmup):           # lazy workspaces, NCCL connections, tile plans -- outside the capture
                    self._run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.out = self._run()
            self.graph = g
        self.graph.replay()
        return self.out

    def release(self):
        self.graph, self.out = None, None
        torch.cuda.synchronize()


def slab_cost_fn(feature_maps: torch.Tensor, sweep: "ops.PlaneSweep", out_dtype=torch.bfloat16):
    """cost_fn for DepthSlabCostVolumeReg: K1 on a plane range of a full sweep (every rank holds all V feature maps and
    the whole 1/(d - s) table; it launches the fused kernel on the planes it needs)."""
    class _Sub:
        pass

    def cost_fn(c0, c1):
        sub = _Sub()
        sub.B, sub.V, sub.D, sub.h, sub.w = sweep.B, sweep.V, c1 - c0, sweep.h, sweep.w
        sub.view_params = sweep.view_params
        sub.tinv = sweep.tinv[:, c0:c1].contiguous()
        return ops.warp_variance(feature_maps, sub, out_dtype)

    return cost_fn
