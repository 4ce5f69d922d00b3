// K1: fused homography warp + variance cost volume, forward (sm_100a); the parity materialiser; the C entry points of K1 / K2.
// The backward kernel (K2) lives in warp_variance_bwd3.cu.
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
//   * 8 lanes own a pixel (4 channels each) and walk a run of depth planes.  Adjacent planes move the sampling position
//     by a fraction of a pixel, so the 2x2 tap footprint of every source view is kept in registers and re-fetched
//     (vectorised channel-last 16 B loads through L1, one full 128-byte line per pixel and tap) only when it moves.
//     The reference view (H = I) is sampled once per pixel.
//   * all V samples of a voxel are in registers => mean, then sum (f - mean)^2, exactly the reference's
//     two-pass variance (the sum f / sum f^2 moment form loses the 1e-4 target); V = 3 uses the
//     algebraically identical two-difference form.
//   * the [B,D,h,w,C] volume is written once with streaming 16 B stores; warped volumes never exist.
#include "warp_common.cuh"
#include <limits.h>
#include <stdlib.h>

using namespace mvsb200;
using namespace mvsb200::warp;

namespace mvsb200 {
namespace warp {
int warp_variance_bwd3(const float* feat, const float* vp, const float* tinv, const void* gcost, int dtype, float* gfeat, int B,
                       int V, int D, int h, int w, cudaStream_t st);
}  // namespace warp
}  // namespace mvsb200

namespace {

constexpr int kWX = 2, kWY = 4; // warps per CTA along x / y
constexpr int kThreads = 32 * kWX * kWY;
constexpr int kTX = 8, kTY = 4;                      // pixel tile of a CTA
constexpr int kRun = 8;                              // planes per staged run (shared memory spent here is L1 lost)

struct Sample {  // bilinear footprint of one sampling position
    int x0, y0;
    float w00, w01, w10, w11;
};

// q = a + g*(c*t);  (ix,iy) = q.xy/q.z - 0.5; clamp so that far-out / non-finite positions land on an
// all-out-of-bounds footprint (grid_sample zero padding) and int conversion is always defined.
__device__ __forceinline__ Sample sample_at(const PixelView& pv, float gx, float gy, float gz, float t, int h, int w) {
    const float m = pv.c * t;
    const float qx = fmaf(gx, m, pv.a0), qy = fmaf(gy, m, pv.a1), qz = fmaf(gz, m, pv.a2);
    const float rz = rcp_approx(qz);
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


// population variance over V samples, two-pass as costvolume.py:12-14 (mean, then sum (x-mean)^2, / V).  Scalar math: plain
// FADD/FFMA can issue on either FMA sub-pipe, the packed f32x2 forms only on the heavy one (profiles/r01_k1_notes.md)
template <int V>
__device__ __forceinline__ float variance1(const float (&x)[V], float ninv, float inv_out) {
    float sum = x[0];
#pragma unroll
    for (int v = 1; v < V; ++v) sum += x[v];
    const float nmean = sum * ninv;
    float d = x[0] + nmean;
    float acc = d * d;
#pragma unroll
    for (int v = 1; v < V; ++v) {
        d = x[v] + nmean;
        acc = fmaf(d, d, acc);
    }
    return acc * inv_out;
}

// ------------------------------------------------------------------------------------------------
// K1 forward ("one offset" form): two phases per run of kRun planes.
//   phase 1  a warp computes the footprint records of ITS four pixels for the run's planes and source views once (homography in
//            registers, ONE clamped base offset per footprint, zero padding and clamping folded into the four weights:
//            warp_common.cuh) into shared memory, double buffered, __syncwarp only -- warps drift apart instead of
//            convoying through a CTA barrier;
//   phase 2  8 lanes own one pixel (4 channels each, so a tap load of a pixel is one full 128-byte line and a warp's
//            load instruction touches only the lines of pixels whose footprint moved); per plane a lane reads the
//            record (one broadcast LDS.128 + LDS.32), refreshes its cached 2x2 taps with predicated loads, blends with
//            packed f32x2 math, takes the variance over the V samples it holds and streams one 16-byte piece of the
//            [B,D,h,w,32] row.
// Round-2 measurements that shaped what is NOT here (profiles/r02_k1_k2_notes.md): a view-outer form with 8 channels per lane
// and 256-bit loads (1.5x fewer instructions) ran 30 % slower -- long/short-scoreboard stalls of its serial per-view loops;
// packed vs scalar FP mixes of this kernel all land within 2 %.
__device__ __forceinline__ void blend2w(float w00, float w01, float w10, float w11, const float4& t00, const float4& t01,
                                        const float4& t10, const float4& t11, float2& lo, float2& hi) {
    const float2 a = make_float2(w00, w00), b = make_float2(w01, w01), c = make_float2(w10, w10), d = make_float2(w11, w11);
    lo = __ffma2_rn(d, lo2(t11), __ffma2_rn(c, lo2(t10), __ffma2_rn(b, lo2(t01), __fmul2_rn(a, lo2(t00)))));
    hi = __ffma2_rn(d, hi2(t11), __ffma2_rn(c, hi2(t10), __ffma2_rn(b, hi2(t01), __fmul2_rn(a, hi2(t00)))));
}

template <int V>
struct Fwd2Cfg {
    static constexpr int kRunV = V <= 4 ? kRun : kRun / 2;
    static constexpr int kWRec = (V - 1) * kRunV * 4;                                    // records per warp and buffer
    static constexpr size_t kSmem = (size_t)(kThreads / 32) * 2 * kWRec * (sizeof(float4) + sizeof(int));
};

// two-difference form of the same variance: d1 = b - a, d2 = c - a, var = (2/9)(d1 (d1 - d2) + d2^2)  (6 operations)
__device__ __forceinline__ float2 variance3_two_diff(float2 a, float2 b, float2 c) {
    const float2 m1 = make_float2(-1.f, -1.f);
    const float2 d1 = __ffma2_rn(a, m1, b), d2 = __ffma2_rn(a, m1, c);
    const float2 t = __ffma2_rn(d1, __ffma2_rn(d2, m1, d1), __fmul2_rn(d2, d2));
    return __fmul2_rn(t, make_float2(2.0f / 9.0f, 2.0f / 9.0f));
}
template <int V, bool BF16OUT>
__global__ void __launch_bounds__(kThreads, V <= 3 ? 3 : (V <= 5 ? 2 : 1))
warp_variance_fwd2_kernel(const float4* __restrict__ feat, const ViewParams* __restrict__ vp, const float* __restrict__ tinv,
                          void* __restrict__ cost, int D, int h, int w, int dchunk, int tiles_x) {
    constexpr int RUN = Fwd2Cfg<V>::kRunV, WREC = Fwd2Cfg<V>::kWRec, NW = kThreads / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* rec_w = reinterpret_cast<float4*>(smem_raw) + warp * 2 * WREC;              // [buffer][V-1][RUN][4 pixels]
    int* rec_o = reinterpret_cast<int*>(reinterpret_cast<float4*>(smem_raw) + NW * 2 * WREC) + warp * 2 * WREC;

    const int b = blockIdx.z, d0 = blockIdx.y * dchunk, nd = min(dchunk, D - d0);
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const unsigned plane = (unsigned)h * (unsigned)w;
    const ViewParams* vpb = vp + (size_t)b * V;

    // ---- phase-1 identity: pixel p4 of this warp's four (fixed per lane: 32 is a multiple of 4)
    const int slot1 = warp * 4 + (lane & 3);
    PixelView pv1[V];
    {
        const float x1 = (float)(tx * kTX + (slot1 & (kTX - 1))), y1 = (float)(ty * kTY + slot1 / kTX);
#pragma unroll
        for (int v = 1; v < V; ++v) pv1[v] = pixel_view(vpb[v], x1, y1);
    }
    auto stage_run = [&](int run0, int buf) {
        const int nrun = min(RUN, nd - run0);
#pragma unroll
        for (int i0 = 0; i0 < WREC; i0 += 32) {
            const int i = i0 + lane;                     // record (v, dd, p4): p4 = i & 3 == lane & 3
            const int v = i / (RUN * 4) + 1, dd = (i >> 2) % RUN;
            if (i < WREC && dd < nrun) {
                PixelView pv = pv1[1];
#pragma unroll
                for (int u = 2; u < V; ++u) if (u == v) pv = pv1[u];
                const ViewParams& q = vpb[v];
                const FootRec1 r = make_record1(pv, q.g[0], q.g[1], q.g[2], __ldg(tinv + (size_t)(b * V + v) * D + d0 + run0 + dd), h, w);
                rec_w[buf * WREC + i] = make_float4(r.w00, r.w01, r.w10, r.w11);
                rec_o[buf * WREC + i] = r.off;
            }
        }
    };

    // ---- phase-2 identity: pixel lane/8 of the warp's four, channels 4*cg .. 4*cg+3
    const int p4 = lane >> 3, cg = lane & 7;
    const int pl = warp * 4 + p4;
    const int px = tx * kTX + (pl & (kTX - 1)), py = ty * kTY + pl / kTX;
    const bool active = px < w && py < h;
    const char* fb = reinterpret_cast<const char*>(feat + (size_t)(b * V) * plane * kSlots + cg);
    const size_t view_bytes = (size_t)plane * kC * 4, line_bytes = (size_t)w * kC * 4;

    stage_run(0, 0);

    float2 ref[2];                                   // reference view: H = I on every plane => one sample per pixel
    {
        const PixelView pv = pixel_view(vpb[0], (float)px, (float)py);
        const FootRec1 r = make_record1(pv, 0.f, 0.f, 0.f, 0.f, h, w);
        const char* pa = fb + (unsigned)r.off;
        const float4 t00 = __ldg(reinterpret_cast<const float4*>(pa)), t01 = __ldg(reinterpret_cast<const float4*>(pa + kC * 4)),
                     t10 = __ldg(reinterpret_cast<const float4*>(pa + line_bytes)),
                     t11 = __ldg(reinterpret_cast<const float4*>(pa + line_bytes + kC * 4));
        blend2w(r.w00, r.w01, r.w10, r.w11, t00, t01, t10, t11, ref[0], ref[1]);
    }

    float4 taps[V][4];
    int key[V];
    const char* fv[V];                               // this lane's 16 bytes of every voxel row of view v
#pragma unroll
    for (int v = 1; v < V; ++v) {
        key[v] = -1;
        fv[v] = fb + (size_t)v * view_bytes;
#pragma unroll
        for (int j = 0; j < 4; ++j) taps[v][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float invV = 1.0f / (float)V;
    __syncwarp();

    int buf = 0;
    for (int run0 = 0; run0 < nd; run0 += RUN, buf ^= 1) {
        const int nrun = min(RUN, nd - run0);
        if (run0 + RUN < nd) stage_run(run0 + RUN, buf ^ 1);
        if (active) {
            size_t vox = ((size_t)(b * D + d0 + run0) * h + py) * w + px;
            const float4* rw = rec_w + buf * WREC + p4;
            const int* ro = rec_o + buf * WREC + p4;
            // (Requesting the taps of plane dd+1 before the variance of plane dd -- a software pipeline -- needs the next
            // plane's weights live across the blend: 120 registers => 2 CTAs/SM => 217 us instead of 160; with 80 registers
            // it spills => 390 us.  64 registers / 4 CTAs per SM: 166 us.  Measured, profiles/r01_k1_notes.md.)
            float4 wt[V];
            auto fetch = [&](int dd) {
#pragma unroll
                for (int v = 1; v < V; ++v) {
                    wt[v] = rw[((v - 1) * RUN + dd) * 4];
                    const int of = ro[((v - 1) * RUN + dd) * 4];
                    const int changed = of != key[v];
                    key[v] = of;
                    const char* pa = fv[v] + (unsigned)of;
                    const char* pb = pa + line_bytes;
                    ldg_f4_if_b(taps[v][0], pa, changed);
                    ldg_f4_if_b(taps[v][1], pa + kC * 4, changed);
                    ldg_f4_if_b(taps[v][2], pb, changed);
                    ldg_f4_if_b(taps[v][3], pb + kC * 4, changed);
                }
            };
            for (int dd = 0; dd < nrun; ++dd, vox += plane) {
                float2 val[2][V];
                val[0][0] = ref[0]; val[1][0] = ref[1];
                fetch(dd);
#pragma unroll
                for (int v = 1; v < V; ++v)
                    blend2w(wt[v].x, wt[v].y, wt[v].z, wt[v].w, taps[v][0], taps[v][1], taps[v][2], taps[v][3], val[0][v], val[1][v]);
                float2 r0, r1;
                if (V == 3) {
                    r0 = variance3_two_diff(val[0][0], val[0][1], val[0][2]);
                    r1 = variance3_two_diff(val[1][0], val[1][1], val[1][2]);
                } else {
                    float xs[4][V];
#pragma unroll
                    for (int v = 0; v < V; ++v) { xs[0][v] = val[0][v].x; xs[1][v] = val[0][v].y; xs[2][v] = val[1][v].x; xs[3][v] = val[1][v].y; }
                    r0 = make_float2(variance1<V>(xs[0], -invV, invV), variance1<V>(xs[1], -invV, invV));
                    r1 = make_float2(variance1<V>(xs[2], -invV, invV), variance1<V>(xs[3], -invV, invV));
                }
                if (BF16OUT) {
                    __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(cost) + vox * kC + 4 * cg;
                    st_cs_u2(row, pack_bf16x2(r0.x, r0.y), pack_bf16x2(r1.x, r1.y));
                } else {
                    st_cs_f4(reinterpret_cast<float4*>(cost) + vox * kSlots + cg, make_float4(r0.x, r0.y, r1.x, r1.y));
                }
            }
        }
        __syncwarp();                                // next buffer staged by every lane, this buffer consumed
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

struct FwdPlan {
    dim3 grid;
    int dchunk, tiles_x;
    size_t smem;
};

FwdPlan make_fwd_plan(int B, int V, int D, int h, int w) {
    FwdPlan p;
    p.tiles_x = (w + kTX - 1) / kTX;
    const int tiles_y = (h + kTY - 1) / kTY;
    const long tiles = (long)p.tiles_x * tiles_y * B;
    // >= 4 waves of 3 CTAs/SM on 148 SMs, in depth runs that are multiples of kRun
    long nchunks = (4L * 3 * 148 + tiles - 1) / tiles;
    const long maxchunks = (D + kRun - 1) / kRun;
    if (nchunks > maxchunks) nchunks = maxchunks;
    if (nchunks < 1) nchunks = 1;
    p.dchunk = (int)((D + nchunks - 1) / nchunks);
    p.dchunk = (p.dchunk + kRun - 1) / kRun * kRun;
    if (const char* e = getenv("MVSB200_DCHUNK")) {
        int v = atoi(e);
        if (v > 0) p.dchunk = v < D ? v : D;         // clamped to [1, D]
    }
    p.grid = dim3((unsigned)(p.tiles_x * tiles_y), (unsigned)((D + p.dchunk - 1) / p.dchunk), (unsigned)B);
    return p;
}

template <int V>
int launch_fwd2(const float* feat, const float* vp, const float* tinv, void* cost, int dtype, int B, int D, int h, int w,
                cudaStream_t st) {
    FwdPlan p = make_fwd_plan(B, V, D, h, w);
    p.smem = Fwd2Cfg<V>::kSmem;
    MVS_REQUIRE(p.grid.y <= 65535 && (long)h * w < (1L << 24), "warp_variance_fwd: volume too large");
    if (dtype == MVSB200_BF16)
        warp_variance_fwd2_kernel<V, true><<<p.grid, kThreads, p.smem, st>>>((const float4*)feat, (const ViewParams*)vp, tinv, cost, D, h,
                                                                             w, p.dchunk, p.tiles_x);
    else
        warp_variance_fwd2_kernel<V, false><<<p.grid, kThreads, p.smem, st>>>((const float4*)feat, (const ViewParams*)vp, tinv, cost, D, h,
                                                                              w, p.dchunk, p.tiles_x);
    MVS_CHECK_LAUNCH("warp_variance_fwd2");
    return MVSB200_OK;
}

}  // namespace

extern "C" int mvsb200_warp_variance_fwd(const float* feat, const float* view_params, const float* tinv, void* cost,
                                         int cost_dtype, int B, int V, int C, int D, int h, int w, void* stream) {
    if (int rc = check_common(feat, view_params, tinv, cost, B, V, C, D, h, w, "warp_variance_fwd")) return rc;
    MVS_REQUIRE(cost_dtype == MVSB200_F32 || cost_dtype == MVSB200_BF16, "warp_variance_fwd: bad cost dtype %d", cost_dtype);
    MVS_REQUIRE(h >= 2 && w >= 2, "warp_variance_fwd: feature maps must be at least 2x2 (the reference's pixel normalisation "
                                  "divides by h-1 and w-1)");
    cudaStream_t st = (cudaStream_t)stream;
    switch (V) {
        case 2: return launch_fwd2<2>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
        case 3: return launch_fwd2<3>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
        case 4: return launch_fwd2<4>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
        case 5: return launch_fwd2<5>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
        case 6: return launch_fwd2<6>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
        case 7: return launch_fwd2<7>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
        case 8: return launch_fwd2<8>(feat, view_params, tinv, cost, cost_dtype, B, D, h, w, st);
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
    MVS_REQUIRE(h >= 2 && w >= 2, "warp_variance_bwd: feature maps must be at least 2x2");
    MVS_CUDA(cudaMemsetAsync(gfeat, 0, (size_t)B * V * h * w * kC * sizeof(float), st));
    return warp_variance_bwd3(feat, view_params, tinv, gcost, gcost_dtype, gfeat, B, V, D, h, w, st);    // warp_variance_bwd3.cu
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
