// K1 / K2: fused homography warp + variance cost volume, forward and backward (sm_100a).
//
// Reference semantics (citations into /root/reference/scripts):
//   homography.py:40-75   H_i(d) = K_i R_i (I - (C_i - C_ref) n / d) R_ref^T K_ref^-1
//   homography.py:78-90   per plane: kornia.warp_perspective(features, H_i(d))  -> bilinear, zero pad,
//                         sampling position inv(H_i(d)) p, then ix = px*w/(w-1) - 0.5
//   costvolume.py:10-14   mean over the V views, population variance over the V views
//
// Design:
//   * geometry: host folds everything into 16 floats per view (include/mvs_b200.h); the per-plane
//     inverse homography is a rank-one update evaluated in registers: q = a + g * (c * tinv[d]).
//   * one thread owns CPL channels of one (y,x) pixel and walks a run of depth planes.  Adjacent planes
//     move the sampling position by a fraction of a pixel, so the 2x2 tap footprint of every source
//     view is kept in registers and re-fetched (vectorised channel-last 16 B loads through L1) only when
//     floor(ix) or floor(iy) changes.  The reference view (H = I) is sampled once per pixel.
//   * all V samples of a voxel are in registers => mean, then sum (f - mean)^2, exactly the reference's
//     two-pass variance (the sum f / sum f^2 moment form loses the 1e-4 target).
//   * the [B,D,h,w,C] volume is written once with streaming 16 B stores; warped volumes never exist.
#include "common.cuh"
#include <limits.h>
#include <stdlib.h>

using namespace mvsb200;

namespace {

struct __align__(16) ViewParams {
    float A[9];
    float g[3];
    float r[3];
    float pad;
};
static_assert(sizeof(ViewParams) == MVSB200_VIEW_PARAM_FLOATS * 4, "view param size");

constexpr int kC = 32;          // channels (CostVolumeReg in_ch, scripts/model.py:70)
constexpr int kSlots = kC / 4;  // float4 slots per voxel row
constexpr int kWX = 2, kWY = 4; // warps per CTA along x / y
constexpr int kThreads = 32 * kWX * kWY;

struct PixelView {  // per (pixel, view) constants of the rank-one form
    float a0, a1, a2, c;
};

__device__ __forceinline__ PixelView pixel_view(const ViewParams& p, float x, float y) {
    PixelView o;
    o.a0 = fmaf(p.A[0], x, fmaf(p.A[1], y, p.A[2]));
    o.a1 = fmaf(p.A[3], x, fmaf(p.A[4], y, p.A[5]));
    o.a2 = fmaf(p.A[6], x, fmaf(p.A[7], y, p.A[8]));
    o.c = fmaf(p.r[0], x, fmaf(p.r[1], y, p.r[2]));
    return o;
}

struct Sample {  // bilinear footprint of one sampling position
    int x0, y0;
    float w00, w01, w10, w11;
};

// q = a + g*(c*t);  (ix,iy) = q.xy/q.z - 0.5; clamp so that far-out / non-finite positions land on an
// all-out-of-bounds footprint (grid_sample zero padding) and int conversion is always defined.
__device__ __forceinline__ Sample sample_at(const PixelView& pv, float gx, float gy, float gz, float t, int h, int w) {
    const float m = pv.c * t;
    const float qx = fmaf(gx, m, pv.a0), qy = fmaf(gy, m, pv.a1), qz = fmaf(gz, m, pv.a2);
    const float rz = __frcp_rn(qz);
    float ix = fmaf(qx, rz, -0.5f), iy = fmaf(qy, rz, -0.5f);
    ix = fminf(fmaxf(ix, -2.0f), (float)(w + 1));   // NaN -> -2
    iy = fminf(fmaxf(iy, -2.0f), (float)(h + 1));
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    Sample s;
    s.x0 = (int)fx0;
    s.y0 = (int)fy0;
    const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
    s.w00 = wx0 * wy0; s.w01 = wx1 * wy0; s.w10 = wx0 * wy1; s.w11 = wx1 * wy1;
    return s;
}

template <int NV4>
__device__ __forceinline__ void load_taps(const float4* __restrict__ fview, int x0, int y0, int h, int w,
                                          const int (&slot)[NV4], float4 (&t)[4][NV4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int xx = x0 + (j & 1), yy = y0 + (j >> 1);
        const bool ok = (unsigned)xx < (unsigned)w && (unsigned)yy < (unsigned)h;
        const float4* p = fview + ((size_t)yy * w + xx) * kSlots;
#pragma unroll
        for (int k = 0; k < NV4; ++k) t[j][k] = ok ? __ldg(p + slot[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__device__ __forceinline__ float4 blend(const Sample& s, float4 t00, float4 t01, float4 t10, float4 t11) {
    float4 o;
    o.x = fmaf(s.w11, t11.x, fmaf(s.w10, t10.x, fmaf(s.w01, t01.x, s.w00 * t00.x)));
    o.y = fmaf(s.w11, t11.y, fmaf(s.w10, t10.y, fmaf(s.w01, t01.y, s.w00 * t00.y)));
    o.z = fmaf(s.w11, t11.z, fmaf(s.w10, t10.z, fmaf(s.w01, t01.z, s.w00 * t00.z)));
    o.w = fmaf(s.w11, t11.w, fmaf(s.w10, t10.w, fmaf(s.w01, t01.w, s.w00 * t00.w)));
    return o;
}

// Which float4 slots of the 32-channel row a lane owns.  fp32 rows: interleaved so that each store
// instruction of a pixel's lane group covers a contiguous 16*LPP bytes; bf16 rows: contiguous channels
// per lane so that one lane emits one 8/16-byte store.
template <int NV4, bool CONTIG>
__device__ __forceinline__ void lane_slots(int cg, int (&slot)[NV4]) {
    constexpr int LPP = kSlots / NV4;
#pragma unroll
    for (int k = 0; k < NV4; ++k) slot[k] = CONTIG ? cg * NV4 + k : cg + k * LPP;
}

struct TileCoord {
    int b, d0, nd, x, y, cg;
    bool active;
};

template <int NV4>
__device__ __forceinline__ TileCoord tile_coord(int D, int h, int w, int dchunk, int tiles_x) {
    constexpr int LPP = kSlots / NV4, PPW = 32 / LPP;
    TileCoord t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    t.b = blockIdx.z;
    t.d0 = blockIdx.y * dchunk;
    t.nd = min(dchunk, D - t.d0);
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    t.x = (tx * kWX + (warp % kWX)) * PPW + lane / LPP;
    t.y = ty * kWY + warp / kWX;
    t.cg = lane % LPP;
    t.active = t.x < w && t.y < h;
    return t;
}

template <int V>
__device__ __forceinline__ void stage_tinv(float* s_tinv, const float* __restrict__ tinv, int b, int D, int d0, int nd,
                                           int dchunk) {
    for (int i = threadIdx.x; i < (V - 1) * nd; i += kThreads) {
        const int v = i / nd, dd = i - v * nd;
        s_tinv[v * dchunk + dd] = tinv[(size_t)(b * V + v + 1) * D + d0 + dd];
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// K1 forward
// ------------------------------------------------------------------------------------------------
template <int V, int CPL, bool BF16OUT>
__global__ void __launch_bounds__(kThreads) warp_variance_fwd_kernel(const float4* __restrict__ feat,
                                                                     const ViewParams* __restrict__ vp,
                                                                     const float* __restrict__ tinv,
                                                                     void* __restrict__ cost, int D, int h, int w,
                                                                     int dchunk, int tiles_x) {
    constexpr int NV4 = CPL / 4;
    extern __shared__ float s_tinv[];
    const TileCoord tc = tile_coord<NV4>(D, h, w, dchunk, tiles_x);
    stage_tinv<V>(s_tinv, tinv, tc.b, D, tc.d0, tc.nd, dchunk);
    if (!tc.active) return;

    int slot[NV4];
    lane_slots<NV4, BF16OUT>(tc.cg, slot);
    const float xf = (float)tc.x, yf = (float)tc.y;
    const size_t plane = (size_t)h * w;
    const ViewParams* vpb = vp + (size_t)tc.b * V;

    // reference view: H = I for every plane => one sample per pixel
    float4 ref[NV4];
    {
        const PixelView pv = pixel_view(vpb[0], xf, yf);
        const Sample s = sample_at(pv, 0.f, 0.f, 0.f, 0.f, h, w);
        float4 t[4][NV4];
        load_taps<NV4>(feat + (size_t)(tc.b * V) * plane * kSlots, s.x0, s.y0, h, w, slot, t);
#pragma unroll
        for (int k = 0; k < NV4; ++k) ref[k] = blend(s, t[0][k], t[1][k], t[2][k], t[3][k]);
    }

    PixelView pv[V];
    float gx[V], gy[V], gz[V];
    int cx[V], cy[V];
    float4 taps[V][4][NV4];
#pragma unroll
    for (int v = 1; v < V; ++v) {
        pv[v] = pixel_view(vpb[v], xf, yf);
        gx[v] = vpb[v].g[0]; gy[v] = vpb[v].g[1]; gz[v] = vpb[v].g[2];
        cx[v] = INT_MIN; cy[v] = INT_MIN;
    }

    const float invV = 1.0f / (float)V;
    size_t vox = ((size_t)(tc.b * D + tc.d0) * h + tc.y) * w + tc.x;
    for (int dd = 0; dd < tc.nd; ++dd, vox += plane) {
        float4 val[V][NV4];
#pragma unroll
        for (int k = 0; k < NV4; ++k) val[0][k] = ref[k];
        bool nan_plane = false;
#pragma unroll
        for (int v = 1; v < V; ++v) {
            const float t = s_tinv[(v - 1) * dchunk + dd];
            nan_plane |= (t != t);
            const Sample s = sample_at(pv[v], gx[v], gy[v], gz[v], t, h, w);
            if (s.x0 != cx[v] || s.y0 != cy[v]) {
                load_taps<NV4>(feat + (size_t)(tc.b * V + v) * plane * kSlots, s.x0, s.y0, h, w, slot, taps[v]);
                cx[v] = s.x0; cy[v] = s.y0;
            }
#pragma unroll
            for (int k = 0; k < NV4; ++k) val[v][k] = blend(s, taps[v][0][k], taps[v][1][k], taps[v][2][k], taps[v][3][k]);
        }
        float4 res[NV4];
#pragma unroll
        for (int k = 0; k < NV4; ++k) {
            float4 sum = val[0][k];
#pragma unroll
            for (int v = 1; v < V; ++v) { sum.x += val[v][k].x; sum.y += val[v][k].y; sum.z += val[v][k].z; sum.w += val[v][k].w; }
            const float4 mean = make_float4(sum.x * invV, sum.y * invV, sum.z * invV, sum.w * invV);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float dx = val[v][k].x - mean.x, dy = val[v][k].y - mean.y, dz = val[v][k].z - mean.z, dw = val[v][k].w - mean.w;
                acc.x = fmaf(dx, dx, acc.x); acc.y = fmaf(dy, dy, acc.y); acc.z = fmaf(dz, dz, acc.z); acc.w = fmaf(dw, dw, acc.w);
            }
            res[k] = make_float4(acc.x * invV, acc.y * invV, acc.z * invV, acc.w * invV);
            if (nan_plane) res[k] = make_float4(NAN, NAN, NAN, NAN);   // d == 0 plane (reference divides by d)
        }
        if (BF16OUT) {
            __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(cost) + vox * kC + tc.cg * CPL;
            if (NV4 == 1) {
                uint2 o = make_uint2(pack_bf16x2(res[0].x, res[0].y), pack_bf16x2(res[0].z, res[0].w));
                asm volatile("st.global.cs.v2.b32 [%0], {%1,%2};" ::"l"(row), "r"(o.x), "r"(o.y) : "memory");
            } else {
                uint4 o = make_uint4(pack_bf16x2(res[0].x, res[0].y), pack_bf16x2(res[0].z, res[0].w),
                                     pack_bf16x2(res[NV4 - 1].x, res[NV4 - 1].y), pack_bf16x2(res[NV4 - 1].z, res[NV4 - 1].w));
                st_cs_u4(reinterpret_cast<uint4*>(row), o);
            }
        } else {
            float4* row = reinterpret_cast<float4*>(cost) + vox * kSlots;
#pragma unroll
            for (int k = 0; k < NV4; ++k) st_cs_f4(row + slot[k], res[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2 backward: d cost / d features
//   d cost / d f_v = (2/V) (f_v - mean) * gcost   (the mean term cancels), chained
//   through the bilinear taps.  Tap gradients of a footprint are accumulated in registers across the
//   planes that share it and flushed with 16-byte vector atomics when the footprint moves.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_f4(float4* p, float4 v) {
    atomicAdd(p, v);   // sm_90+: single 16-byte reduction at L2
}

template <int NV4>
__device__ __forceinline__ void flush_taps(float4* __restrict__ gview, int x0, int y0, int h, int w,
                                           const int (&slot)[NV4], float4 (&acc)[4][NV4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int xx = x0 + (j & 1), yy = y0 + (j >> 1);
        const bool ok = (unsigned)xx < (unsigned)w && (unsigned)yy < (unsigned)h;
        float4* p = gview + ((size_t)yy * w + xx) * kSlots;
#pragma unroll
        for (int k = 0; k < NV4; ++k) {
            if (ok) red_add_f4(p + slot[k], acc[j][k]);
            acc[j][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

__device__ __forceinline__ void axpy4(float4& a, float s, const float4& g) {
    a.x = fmaf(s, g.x, a.x); a.y = fmaf(s, g.y, a.y); a.z = fmaf(s, g.z, a.z); a.w = fmaf(s, g.w, a.w);
}

template <int V, int CPL, bool BF16G>
__global__ void __launch_bounds__(kThreads) warp_variance_bwd_kernel(const float4* __restrict__ feat,
                                                                     const ViewParams* __restrict__ vp,
                                                                     const float* __restrict__ tinv,
                                                                     const void* __restrict__ gcost,
                                                                     float4* __restrict__ gfeat, int D, int h, int w,
                                                                     int dchunk, int tiles_x) {
    constexpr int NV4 = CPL / 4;
    extern __shared__ float s_tinv[];
    const TileCoord tc = tile_coord<NV4>(D, h, w, dchunk, tiles_x);
    stage_tinv<V>(s_tinv, tinv, tc.b, D, tc.d0, tc.nd, dchunk);
    if (!tc.active) return;

    int slot[NV4];
    lane_slots<NV4, BF16G>(tc.cg, slot);
    const float xf = (float)tc.x, yf = (float)tc.y;
    const size_t plane = (size_t)h * w;
    const ViewParams* vpb = vp + (size_t)tc.b * V;

    float4 ref[NV4], gref[NV4];
    Sample sref;
    {
        const PixelView pv0 = pixel_view(vpb[0], xf, yf);
        sref = sample_at(pv0, 0.f, 0.f, 0.f, 0.f, h, w);
        float4 t[4][NV4];
        load_taps<NV4>(feat + (size_t)(tc.b * V) * plane * kSlots, sref.x0, sref.y0, h, w, slot, t);
#pragma unroll
        for (int k = 0; k < NV4; ++k) {
            ref[k] = blend(sref, t[0][k], t[1][k], t[2][k], t[3][k]);
            gref[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }

    PixelView pv[V];
    float gx[V], gy[V], gz[V];
    int cx[V], cy[V];
    float4 taps[V][4][NV4];
    float4 acc[V][4][NV4];
#pragma unroll
    for (int v = 1; v < V; ++v) {
        pv[v] = pixel_view(vpb[v], xf, yf);
        gx[v] = vpb[v].g[0]; gy[v] = vpb[v].g[1]; gz[v] = vpb[v].g[2];
        cx[v] = INT_MIN; cy[v] = INT_MIN;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < NV4; ++k) acc[v][j][k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    const float invV = 1.0f / (float)V, twoV = 2.0f / (float)V;
    size_t vox = ((size_t)(tc.b * D + tc.d0) * h + tc.y) * w + tc.x;
    for (int dd = 0; dd < tc.nd; ++dd, vox += plane) {
        // upstream gradient row slice
        float4 g[NV4];
        if (BF16G) {
            const __nv_bfloat16* row = reinterpret_cast<const __nv_bfloat16*>(gcost) + vox * kC + tc.cg * CPL;
#pragma unroll
            for (int k = 0; k < NV4; ++k) {
                const uint2 u = *reinterpret_cast<const uint2*>(row + 4 * k);
                const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
                const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
                g[k] = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
            }
        } else {
            const float4* row = reinterpret_cast<const float4*>(gcost) + vox * kSlots;
#pragma unroll
            for (int k = 0; k < NV4; ++k) g[k] = ld_cs_f4(row + slot[k]);
        }

        float4 val[V][NV4];
        Sample smp[V];
#pragma unroll
        for (int k = 0; k < NV4; ++k) val[0][k] = ref[k];
        bool nan_plane = false;
#pragma unroll
        for (int v = 1; v < V; ++v) {
            const float t = s_tinv[(v - 1) * dchunk + dd];
            nan_plane |= (t != t);
            smp[v] = sample_at(pv[v], gx[v], gy[v], gz[v], t, h, w);
            if (smp[v].x0 != cx[v] || smp[v].y0 != cy[v]) {
                if (cx[v] != INT_MIN)
                    flush_taps<NV4>(gfeat + (size_t)(tc.b * V + v) * plane * kSlots, cx[v], cy[v], h, w, slot, acc[v]);
                load_taps<NV4>(feat + (size_t)(tc.b * V + v) * plane * kSlots, smp[v].x0, smp[v].y0, h, w, slot, taps[v]);
                cx[v] = smp[v].x0; cy[v] = smp[v].y0;
            }
#pragma unroll
            for (int k = 0; k < NV4; ++k)
                val[v][k] = blend(smp[v], taps[v][0][k], taps[v][1][k], taps[v][2][k], taps[v][3][k]);
        }
        if (nan_plane) continue;   // d == 0 plane: the reference's gradient is NaN there; contribute nothing
#pragma unroll
        for (int k = 0; k < NV4; ++k) {
            float4 sum = val[0][k];
#pragma unroll
            for (int v = 1; v < V; ++v) { sum.x += val[v][k].x; sum.y += val[v][k].y; sum.z += val[v][k].z; sum.w += val[v][k].w; }
            const float4 mean = make_float4(sum.x * invV, sum.y * invV, sum.z * invV, sum.w * invV);
            const float4 gs = make_float4(g[k].x * twoV, g[k].y * twoV, g[k].z * twoV, g[k].w * twoV);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 gv = make_float4((val[v][k].x - mean.x) * gs.x, (val[v][k].y - mean.y) * gs.y,
                                              (val[v][k].z - mean.z) * gs.z, (val[v][k].w - mean.w) * gs.w);
                if (v == 0) {
                    gref[k].x += gv.x; gref[k].y += gv.y; gref[k].z += gv.z; gref[k].w += gv.w;
                } else {
                    axpy4(acc[v][0][k], smp[v].w00, gv);
                    axpy4(acc[v][1][k], smp[v].w01, gv);
                    axpy4(acc[v][2][k], smp[v].w10, gv);
                    axpy4(acc[v][3][k], smp[v].w11, gv);
                }
            }
        }
    }
    // final flush: source views, then the reference view's constant footprint
#pragma unroll
    for (int v = 1; v < V; ++v)
        if (cx[v] != INT_MIN)
            flush_taps<NV4>(gfeat + (size_t)(tc.b * V + v) * plane * kSlots, cx[v], cy[v], h, w, slot, acc[v]);
    {
        float4 a0[4][NV4];
#pragma unroll
        for (int k = 0; k < NV4; ++k) {
            a0[0][k] = make_float4(sref.w00 * gref[k].x, sref.w00 * gref[k].y, sref.w00 * gref[k].z, sref.w00 * gref[k].w);
            a0[1][k] = make_float4(sref.w01 * gref[k].x, sref.w01 * gref[k].y, sref.w01 * gref[k].z, sref.w01 * gref[k].w);
            a0[2][k] = make_float4(sref.w10 * gref[k].x, sref.w10 * gref[k].y, sref.w10 * gref[k].z, sref.w10 * gref[k].w);
            a0[3][k] = make_float4(sref.w11 * gref[k].x, sref.w11 * gref[k].y, sref.w11 * gref[k].z, sref.w11 * gref[k].w);
        }
        flush_taps<NV4>(gfeat + (size_t)(tc.b * V) * plane * kSlots, sref.x0, sref.y0, h, w, slot, a0);
    }
}

// ------------------------------------------------------------------------------------------------
// parity/debug: the warped volumes as homography_warping returns them, [N, C, D, h, w] contiguous
// ------------------------------------------------------------------------------------------------
__global__ void warp_materialize_kernel(const float4* __restrict__ feat, const ViewParams* __restrict__ vp,
                                        const float* __restrict__ tinv, float* __restrict__ warped, int V, int D, int h,
                                        int w) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int n = blockIdx.z / D, d = blockIdx.z % D;
    if (x >= w) return;
    const int v = n % V;
    const ViewParams p = vp[n];
    const PixelView pv = pixel_view(p, (float)x, (float)y);
    const float t = (v == 0) ? 0.f : tinv[(size_t)n * D + d];
    const bool nan_plane = tinv[(size_t)(n - v + (V > 1 ? 1 : 0)) * D + d] != tinv[(size_t)(n - v + (V > 1 ? 1 : 0)) * D + d];
    const Sample s = (v == 0) ? sample_at(pv, 0.f, 0.f, 0.f, 0.f, h, w) : sample_at(pv, p.g[0], p.g[1], p.g[2], t, h, w);
    const size_t plane = (size_t)h * w;
    const float4* fview = feat + (size_t)n * plane * kSlots;
    for (int k = 0; k < kSlots; ++k) {
        int slot[1] = {k};
        float4 tp[4][1];
        load_taps<1>(fview, s.x0, s.y0, h, w, slot, tp);
        float4 o = blend(s, tp[0][0], tp[1][0], tp[2][0], tp[3][0]);
        if (nan_plane) o = make_float4(NAN, NAN, NAN, NAN);
        float* dst = warped + (((size_t)n * kC + 4 * k) * D + d) * plane + (size_t)y * w + x;
        dst[0] = o.x; dst[(size_t)D * plane] = o.y; dst[2 * (size_t)D * plane] = o.z; dst[3 * (size_t)D * plane] = o.w;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int check_common(const void* feat, const void* vp, const void* tinv, const void* vol, int B, int V, int C, int D, int h,
                 int w, const char* name) {
    MVS_REQUIRE(feat && vp && tinv && vol, "%s: null pointer", name);
    MVS_REQUIRE(aligned16(feat) && aligned16(vp) && aligned16(vol), "%s: pointers must be 16-byte aligned", name);
    MVS_REQUIRE(C == kC, "%s: C must be 32 (got %d)", name, C);
    MVS_REQUIRE(V >= 2 && V <= 8, "%s: 2 <= V <= 8 (got %d)", name, V);
    MVS_REQUIRE(B >= 1 && D >= 1 && h >= 1 && w >= 1, "%s: bad shape B=%d D=%d h=%d w=%d", name, B, D, h, w);
    MVS_REQUIRE(B <= 65535, "%s: B too large", name);
    return MVSB200_OK;
}

struct Plan {
    dim3 grid;
    int dchunk, tiles_x;
    size_t smem;
};

Plan make_plan(int B, int V, int D, int h, int w, int cpl) {
    const int ppw = 32 / (kSlots / (cpl / 4));
    Plan p;
    p.tiles_x = (w + kWX * ppw - 1) / (kWX * ppw);
    const int tiles_y = (h + kWY - 1) / kWY;
    const long tiles = (long)p.tiles_x * tiles_y * B;
    // enough CTAs for >= 4 waves of 2 CTAs/SM on 148 SMs, but keep depth runs long (tap reuse across planes)
    long nchunks = (4L * 2 * 148 + tiles - 1) / tiles;
    long maxchunks = D >= 32 ? D / 16 : 1;
    if (nchunks > maxchunks) nchunks = maxchunks;
    if (nchunks < 1) nchunks = 1;
    p.dchunk = (int)((D + nchunks - 1) / nchunks);
    if (const char* e = getenv("MVSB200_DCHUNK")) {
        int v = atoi(e);
        if (v > 0) p.dchunk = v < D ? v : D;
    }
    p.grid = dim3((unsigned)(p.tiles_x * tiles_y), (unsigned)((D + p.dchunk - 1) / p.dchunk), (unsigned)B);
    p.smem = (size_t)(V - 1) * p.dchunk * sizeof(float);
    return p;
}

template <int V, int CPL>
int launch_fwd(const float* feat, const float* vp, const float* tinv, void* cost, int dtype, int B, int D, int h, int w,
               cudaStream_t st) {
    const Plan p = make_plan(B, V, D, h, w, CPL);
    MVS_REQUIRE(p.smem <= 48 * 1024 && p.grid.y <= 65535, "warp_variance_fwd: depth run too long");
    if (dtype == MVSB200_BF16)
        warp_variance_fwd_kernel<V, CPL, true><<<p.grid, kThreads, p.smem, st>>>(
            (const float4*)feat, (const ViewParams*)vp, tinv, cost, D, h, w, p.dchunk, p.tiles_x);
    else
        warp_variance_fwd_kernel<V, CPL, false><<<p.grid, kThreads, p.smem, st>>>(
            (const float4*)feat, (const ViewParams*)vp, tinv, cost, D, h, w, p.dchunk, p.tiles_x);
    MVS_CHECK_LAUNCH("warp_variance_fwd");
    return MVSB200_OK;
}

template <int V, int CPL>
int launch_bwd(const float* feat, const float* vp, const float* tinv, const void* gcost, int dtype, float* gfeat, int B,
               int D, int h, int w, cudaStream_t st) {
    const Plan p = make_plan(B, V, D, h, w, CPL);
    MVS_REQUIRE(p.smem <= 48 * 1024 && p.grid.y <= 65535, "warp_variance_bwd: depth run too long");
    if (dtype == MVSB200_BF16)
        warp_variance_bwd_kernel<V, CPL, true><<<p.grid, kThreads, p.smem, st>>>(
            (const float4*)feat, (const ViewParams*)vp, tinv, gcost, (float4*)gfeat, D, h, w, p.dchunk, p.tiles_x);
    else
        warp_variance_bwd_kernel<V, CPL, false><<<p.grid, kThreads, p.smem, st>>>(
            (const float4*)feat, (const ViewParams*)vp, tinv, gcost, (float4*)gfeat, D, h, w, p.dchunk, p.tiles_x);
    MVS_CHECK_LAUNCH("warp_variance_bwd");
    return MVSB200_OK;
}

}  // namespace

extern "C" int mvsb200_warp_variance_fwd(const float* feat, const float* view_params, const float* tinv, void* cost,
                                         int cost_dtype, int B, int V, int C, int D, int h, int w, void* stream) {
    if (int rc = check_common(feat, view_params, tinv, cost, B, V, C, D, h, w, "warp_variance_fwd")) return rc;
    MVS_REQUIRE(cost_dtype == MVSB200_F32 || cost_dtype == MVSB200_BF16, "warp_variance_fwd: bad cost dtype %d", cost_dtype);
    cudaStream_t st = (cudaStream_t)stream;
    switch (V) {
        case 2: return launch_fwd<2, 8>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
        case 3: return launch_fwd<3, 8>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
        case 4: return launch_fwd<4, 4>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
        case 5: return launch_fwd<5, 4>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
        case 6: return launch_fwd<6, 4>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
        case 7: return launch_fwd<7, 4>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
        case 8: return launch_fwd<8, 4>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
    }
    MVS_FAIL(MVSB200_E_UNSUPPORTED, "warp_variance_fwd: V=%d", V);
}

extern "C" int mvsb200_warp_variance_bwd(const float* feat, const float* view_params, const float* tinv,
                                         const void* gcost, int gcost_dtype, float* gfeat, int B, int V, int C, int D,
                                         int h, int w, void* stream) {
    if (int rc = check_common(feat, view_params, tinv, gcost, B, V, C, D, h, w, "warp_variance_bwd")) return rc;
    MVS_REQUIRE(gfeat && aligned16(gfeat), "warp_variance_bwd: gfeat null or misaligned");
    MVS_REQUIRE(gcost_dtype == MVSB200_F32 || gcost_dtype == MVSB200_BF16, "warp_variance_bwd: bad gcost dtype %d", gcost_dtype);
    cudaStream_t st = (cudaStream_t)stream;
    MVS_CUDA(cudaMemsetAsync(gfeat, 0, (size_t)B * V * h * w * kC * sizeof(float), st));
    switch (V) {
        case 2: return launch_bwd<2, 4>(feat, view_params, tinv, gcost, gcost_dtype, gfeat, B, D, h, w, st);
        case 3: return launch_bwd<3, 4>(feat, view_params, tinv, gcost, gcost_dtype, gfeat, B, D, h, w, st);
        case 4: return launch_bwd<4, 4>(feat, view_params, tinv, gcost, gcost_dtype, gfeat, B, D, h, w, st);
        case 5: return launch_bwd<5, 4>(feat, view_params, tinv, gcost, gcost_dtype, gfeat, B, D, h, w, st);
        case 6: return launch_bwd<6, 4>(feat, view_params, tinv, gcost, gcost_dtype, gfeat, B, D, h, w, st);
        case 7: return launch_bwd<7, 4>(feat, view_params, tinv, gcost, gcost_dtype, gfeat, B, D, h, w, st);
        case 8: return launch_bwd<8, 4>(feat, view_params, tinv, gcost, gcost_dtype, gfeat, B, D, h, w, st);
    }
    MVS_FAIL(MVSB200_E_UNSUPPORTED, "warp_variance_bwd: V=%d", V);
}

extern "C" int mvsb200_warp_materialize(const float* feat, const float* view_params, const float* tinv, float* warped,
                                        int B, int V, int C, int D, int h, int w, void* stream) {
    if (int rc = check_common(feat, view_params, tinv, warped, B, V, C, D, h, w, "warp_materialize")) return rc;
    MVS_REQUIRE((long)B * V * D <= 65535 && h <= 65535, "warp_materialize: debug path, volume too large");
    dim3 grid((w + 63) / 64, h, B * V * D);
    warp_materialize_kernel<<<grid, 64, 0, (cudaStream_t)stream>>>((const float4*)feat, (const ViewParams*)view_params,
                                                                   tinv, warped, V, D, h, w);
    MVS_CHECK_LAUNCH("warp_materialize");
    return MVSB200_OK;
}
