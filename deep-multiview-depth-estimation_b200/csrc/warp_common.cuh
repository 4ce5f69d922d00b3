// Device helpers shared by the warp kernels (K1 / K2 / the parity materialiser): per-view geometry in the rank-one
// (Sherman-Morrison) form, the bilinear footprint record, 128/256-bit predicated loads.
//
// Reference semantics (citations into /root/reference/scripts):
//   homography.py:40-75   H_i(d) = K_i R_i (I - (C_i - C_ref) n / d) R_ref^T K_ref^-1
//   homography.py:78-90   per plane: kornia.warp_perspective(features, H_i(d))  -> bilinear, zero pad,
//                         sampling position inv(H_i(d)) p, then ix = px*w/(w-1) - 0.5
#pragma once
#include "common.cuh"

namespace mvsb200 {
namespace warp {

struct __align__(16) ViewParams {
    float A[9];
    float g[3];
    float r[3];
    float pad;
};
static_assert(sizeof(ViewParams) == MVSB200_VIEW_PARAM_FLOATS * 4, "view param size");

constexpr int kC = 32;          // channels (CostVolumeReg in_ch, scripts/model.py:70)
constexpr int kSlots = kC / 4;  // float4 slots per voxel row

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

// 1/x to <= 1 ulp (MUFU.RCP): sampling-position error ~2e-5 px, below the 9e-5 px by which the reference's own
// fp32 matrix chain deviates from exact arithmetic.  Every kernel uses this same reciprocal so
// that forward, backward and the materialised volumes agree on the footprints.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Footprint record of one (pixel, plane, view): the 2x2 footprint is addressed from ONE clamped base (xc, yc) in
// [0, w-2] x [0, h-2]: the taps are base, base + one voxel row, base + one line, base + line + row.  Zero padding and the
// clamping are folded into the four weights (a footprint hanging over the left/top edge hands its in-bounds weight to the
// first tap, over the right/bottom edge to the second).
struct __align__(16) FootRec1 {
    float w00, w01, w10, w11;     // weights of the taps at base, base+1, base+line, base+line+1
    int off;                      // byte offset of the base tap inside the view's feature map (fp32 rows of kC channels)
};

// q = a + g*(c*t);  (ix,iy) = q.xy/q.z - 0.5; clamped so that far-out / non-finite positions land on an
// all-out-of-bounds footprint (grid_sample zero padding) and the int conversion is always defined.
__device__ __forceinline__ FootRec1 make_record1(const PixelView& pv, float gx, float gy, float gz, float t, int h, int w) {
    const float m = pv.c * t;
    const float qx = fmaf(gx, m, pv.a0), qy = fmaf(gy, m, pv.a1), qz = fmaf(gz, m, pv.a2);
    const float rz = rcp_approx(qz);
    float ix = fmaf(qx, rz, -0.5f), iy = fmaf(qy, rz, -0.5f);
    ix = fminf(fmaxf(ix, -2.0f), (float)(w + 1));   // NaN -> -2: footprint entirely out of bounds
    iy = fminf(fmaxf(iy, -2.0f), (float)(h + 1));
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float wx1 = ix - fx0, wy1 = iy - fy0;
    float ax = ((unsigned)x0 < (unsigned)w) ? 1.0f - wx1 : 0.f, bx = ((unsigned)(x0 + 1) < (unsigned)w) ? wx1 : 0.f;
    float ay = ((unsigned)y0 < (unsigned)h) ? 1.0f - wy1 : 0.f, by = ((unsigned)(y0 + 1) < (unsigned)h) ? wy1 : 0.f;
    if (x0 < 0) { ax = bx; bx = 0.f; } else if (x0 > w - 2) { bx = ax; ax = 0.f; }
    if (y0 < 0) { ay = by; by = 0.f; } else if (y0 > h - 2) { by = ay; ay = 0.f; }
    const int xc = min(max(x0, 0), w - 2), yc = min(max(y0, 0), h - 2);
    FootRec1 r;
    r.w00 = ax * ay; r.w01 = bx * ay; r.w10 = ax * by; r.w11 = bx * by;
    if (t != t) { r.w00 = NAN; r.w01 = NAN; r.w10 = NAN; r.w11 = NAN; }      // d == 0 plane: the reference's whole plane is NaN
    r.off = (yc * w + xc) * (kC * 4);
    return r;
}

__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }

// predicated 16-byte read-only load that keeps the old register contents when the predicate is off
__device__ __forceinline__ void ldg_f4_if_b(float4& t, const char* p, int pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "+f"(t.x), "+f"(t.y), "+f"(t.z), "+f"(t.w) : "l"(p), "r"(pred));
}

// 8 consecutive channels of a voxel row in registers
struct F8 {
    float2 v[4];
};

// predicated 32-byte (256-bit, sm_100) read-only load: four lanes cover one full 128-byte voxel row per instruction
__device__ __forceinline__ void ldg_f8_if(F8& t, const char* p, int pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %9, 0;\n\t@p ld.global.nc.L1::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t}"
                 : "+f"(t.v[0].x), "+f"(t.v[0].y), "+f"(t.v[1].x), "+f"(t.v[1].y), "+f"(t.v[2].x), "+f"(t.v[2].y), "+f"(t.v[3].x),
                   "+f"(t.v[3].y)
                 : "l"(p), "r"(pred));
}
__device__ __forceinline__ void ldg_f8(F8& t, const char* p) {
    asm volatile("ld.global.nc.L1::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(t.v[0].x), "=f"(t.v[0].y), "=f"(t.v[1].x), "=f"(t.v[1].y), "=f"(t.v[2].x), "=f"(t.v[2].y), "=f"(t.v[3].x),
                   "=f"(t.v[3].y)
                 : "l"(p));
}
// streaming 32-byte store that does not allocate in L1
__device__ __forceinline__ void st_cs_f8(void* p, const F8& t) {
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(t.v[0].x), "f"(t.v[0].y),
                 "f"(t.v[1].x), "f"(t.v[1].y), "f"(t.v[2].x), "f"(t.v[2].y), "f"(t.v[3].x), "f"(t.v[3].y)
                 : "memory");
}

// out = w00*t00 + w01*t01 + w10*t10 + w11*t11 on 8 channels, packed fp32 (same operation order as the scalar blend:
// fma(w11,t11, fma(w10,t10, fma(w01,t01, w00*t00))))
__device__ __forceinline__ void blend8(const float4& wt, const F8& t00, const F8& t01, const F8& t10, const F8& t11, F8& o) {
    const float2 a = make_float2(wt.x, wt.x), b = make_float2(wt.y, wt.y), c = make_float2(wt.z, wt.z), d = make_float2(wt.w, wt.w);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        o.v[k] = __ffma2_rn(d, t11.v[k], __ffma2_rn(c, t10.v[k], __ffma2_rn(b, t01.v[k], __fmul2_rn(a, t00.v[k]))));
}

}  // namespace warp
}  // namespace mvsb200
