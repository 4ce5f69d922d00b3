"""kornia 0.6.3 ``warp_perspective`` semantics, restated.

Chain executed by 0.6.3 for ``warp_perspective(src[N,C,H,W], M[N,3,3], dsize, align_corners=...)``:
  normalize_homography -> normal_transform_pixel, _torch_inverse_cast
  create_meshgrid(normalized_coordinates=True)
  transform_points -> convert_points_{to,from}_homogeneous (eps = 1e-8)
  torch.nn.functional.grid_sample(mode, padding_mode, align_corners)
Everything stays in the dtype of ``src`` (fp32 on the reference path) so the fp32
rounding of the real dependency is reproduced, not just its maths.
"""
import torch
import torch.nn.functional as F


def _pixel_to_unit(height, width, device, dtype):
    # pixel index -> [-1, 1] with the *align_corners=True* convention (0 -> -1, size-1 -> +1);
    # a singleton dimension divides by 1e-14 instead of 0.
    eps = 1e-14
    wd = eps if width == 1 else float(width - 1)
    hd = eps if height == 1 else float(height - 1)
    m = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]], device=device, dtype=dtype)
    m[0, 0] = m[0, 0] * 2.0 / wd
    m[1, 1] = m[1, 1] * 2.0 / hd
    return m.unsqueeze(0)


def _inverse(m):
    # 0.6.3 inverts in fp32 unless the input already is fp64
    dt = m.dtype
    work = m if dt in (torch.float32, torch.float64) else m.to(torch.float32)
    return torch.inverse(work).to(dt)


def _unit_grid(height, width, device, dtype):
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    xs = (xs / (width - 1) - 0.5) * 2
    ys = (ys / (height - 1) - 0.5) * 2
    gx, gy = torch.meshgrid(xs, ys, indexing="ij")
    return torch.stack((gx, gy), dim=-1).permute(1, 0, 2).unsqueeze(0)  # 1,H,W,2


def warp_perspective(src, M, dsize, mode="bilinear", padding_mode="zeros", align_corners=True):
    if src.dim() != 4:
        raise ValueError(f"Input src must be a BxCxHxW tensor. Got {src.shape}")
    if M.dim() != 3 or M.shape[-2:] != (3, 3):
        raise ValueError(f"Input M must be a Bx3x3 tensor. Got {M.shape}")
    n, _, h_in, w_in = src.shape
    h_out, w_out = dsize
    to_unit_src = _pixel_to_unit(h_in, w_in, M.device, M.dtype)
    to_unit_dst = _pixel_to_unit(h_out, w_out, M.device, M.dtype)
    m_unit = to_unit_dst @ (M @ _inverse(to_unit_src))          # dst_unit <- src_unit
    m_unit_inv = _inverse(m_unit)                                 # src_unit <- dst_unit
    grid = _unit_grid(h_out, w_out, src.device, src.dtype).repeat(n, 1, 1, 1)
    pts = grid.reshape(n, h_out * w_out, 2)
    pts_h = F.pad(pts, (0, 1), value=1.0)
    q = torch.bmm(pts_h, m_unit_inv.transpose(1, 2))
    z = q[..., -1:]
    scale = torch.where(z.abs() > 1e-8, 1.0 / (z + 1e-8), torch.ones_like(z))
    g = (scale * q[..., :-1]).reshape(n, h_out, w_out, 2)
    return F.grid_sample(src, g, mode=mode, padding_mode=padding_mode, align_corners=align_corners)
