// K1 / K2, third form ("view-outer"): fused homography warp + variance cost volume, forward and backward (sm_100a).
//
// Why a third form.  The second form (warp_variance.cu) keeps the 2x2 taps of EVERY source view of a pixel in registers and
// therefore gives a lane only 4 channels: the per-plane control work of a (pixel, view) -- footprint record, footprint test,
// address arithmetic, predicated reloads -- is replicated in 8 lanes and outweighs the arithmetic 2.5 : 1 (ncu, round 1:
// 106 warp instructions per 128 channel-voxels, 30 of them packed FMAs).  Here a lane owns 8 channels (4 lanes per pixel,
// one 256-bit load per tap, 4 lanes = one full 128-byte row) and walks ONE source view at a time over a short run of planes:
// only that view's taps are live in registers, the blended samples of the earlier views of the run wait in a thread-private
// slice of shared memory (each lane re-reads exactly what it wrote: no barrier, no bank conflict) and the pass over the last
// view takes the variance.  Half the control instructions per channel-voxel at the register footprint of the second form.
//
// Reference semantics: warp_common.cuh (homography.py:40-90) and costvolume.py:10-14 (mean over the V views, population
// variance over the V views, the reference view included).
#include "warp_common.cuh"
#include <stdlib.h>

using namespace mvsb200;
using namespace mvsb200::warp;

namespace {

constexpr int kTX = 16, kTY = 4;                   // pixel tile of a CTA: 8 warps, a warp = 8 pixels of a line x 4 lanes
constexpr int kThreads = 256;
constexpr int kPW = 8;                             // pixels per warp

template <int V>
struct Cfg3 {
    // planes per run: the stored samples of a run cost (V-2) * RUN * 8 KB of shared memory per CTA
    static constexpr int kRun = V <= 4 ? 4 : 2;
    static constexpr int kRec = (V - 1) * kRun * kPW;                        // records per warp and run
    static constexpr size_t kRecBytes = (size_t)(kThreads / 32) * kRec * (sizeof(float4) + sizeof(int));
    static constexpr size_t kPvBytes = (size_t)(kThreads / 32) * (V - 1) * kPW * sizeof(float4);
    static constexpr size_t kValBytes = (size_t)(V - 2) * kRun * kThreads * 32;
    static constexpr size_t kSmemFwd = kRecBytes + kPvBytes + kValBytes;
};

__device__ __forceinline__ float2 sub2(float2 a, float2 b) {        // a - b, one packed FMA (b * -1 is exact)
    return __ffma2_rn(b, make_float2(-1.f, -1.f), a);
}

// population variance of three samples from two differences: with d1 = b - a, d2 = c - a,
// sum (x - mean)^2 = (2/3)(d1^2 + d2^2 - d1 d2)  =>  var = (2/9) (d1 (d1 - d2) + d2^2); no cancellation (positive definite form)
__device__ __forceinline__ float2 variance3(float2 a, float2 b, float2 c) {
    const float2 d1 = sub2(b, a), d2 = sub2(c, a);
    const float2 t = __ffma2_rn(d1, sub2(d1, d2), __fmul2_rn(d2, d2));
    return __fmul2_rn(t, make_float2(2.0f / 9.0f, 2.0f / 9.0f));
}

// ------------------------------------------------------------------------------------------------
// K1 forward
// ------------------------------------------------------------------------------------------------
template <int V, bool BF16OUT, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
warp_variance_fwd3_kernel(const char* __restrict__ feat, const ViewParams* __restrict__ vp, const float* __restrict__ tinv,
                          void* __restrict__ cost, int D, int h, int w, int dchunk, int tiles_x) {
    constexpr int RUN = Cfg3<V>::kRun, NREC = Cfg3<V>::kRec, NW = kThreads / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* rec_w = reinterpret_cast<float4*>(smem_raw) + warp * NREC;                                   // [view-1][plane][pixel]
    int* rec_o = reinterpret_cast<int*>(smem_raw + (size_t)NW * NREC * sizeof(float4)) + warp * NREC;
    float4* pvs = reinterpret_cast<float4*>(smem_raw + Cfg3<V>::kRecBytes) + warp * (V - 1) * kPW;       // [view-1][pixel]
    float4* vals = reinterpret_cast<float4*>(smem_raw + Cfg3<V>::kRecBytes + Cfg3<V>::kPvBytes) + threadIdx.x;
    // vals[((slot * RUN + plane) * 2 + half) * kThreads]: thread-private, a warp's access is 512 contiguous bytes

    const int b = blockIdx.z, d0 = blockIdx.y * dchunk, nd = min(dchunk, D - d0);
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const unsigned plane = (unsigned)h * (unsigned)w;
    const ViewParams* vpb = vp + (size_t)b * V;
    const int x_w = tx * kTX + (warp & 1) * kPW, py = ty * kTY + (warp >> 1);        // the warp's 8 pixels: (x_w .. x_w+7, py)

    // per (pixel, view) constants of the warp's pixels, once
    for (int i = lane; i < (V - 1) * kPW; i += 32) {
        const PixelView pv = pixel_view(vpb[i / kPW + 1], (float)(x_w + (i & (kPW - 1))), (float)py);
        pvs[i] = make_float4(pv.a0, pv.a1, pv.a2, pv.c);
    }

    const int p8 = lane >> 2, cq = lane & 3;
    const int px = x_w + p8;
    const bool active = px < w && py < h;
    const size_t view_bytes = (size_t)plane * kC * 4, line_bytes = (size_t)w * kC * 4;
    const char* fb = feat + (size_t)(b * V) * view_bytes + cq * 32;                  // this lane's 32 bytes of every voxel row

    F8 ref;                                          // reference view: H = I on every plane => one sample per pixel
    {
        const PixelView pv = pixel_view(vpb[0], (float)px, (float)py);
        const FootRec1 r = make_record1(pv, 0.f, 0.f, 0.f, 0.f, h, w);
        const char* pa = fb + (unsigned)r.off;
        F8 t00, t01, t10, t11;
        ldg_f8(t00, pa); ldg_f8(t01, pa + kC * 4); ldg_f8(t10, pa + line_bytes); ldg_f8(t11, pa + line_bytes + kC * 4);
        blend8(make_float4(r.w00, r.w01, r.w10, r.w11), t00, t01, t10, t11, ref);
    }
    const float invV = 1.0f / (float)V;

    F8 taps[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) taps[j].v[k] = make_float2(0.f, 0.f);

    for (int run0 = 0; run0 < nd; run0 += RUN) {
        const int nrun = min(RUN, nd - run0);
        __syncwarp();                                // the previous run's records are consumed (and pvs written, first time)
        for (int i = lane; i < NREC; i += 32) {      // footprint records of this run: (view, plane, pixel), one per lane and step
            const int p = i & (kPW - 1), dd = (i / kPW) & (RUN - 1), v = i / (kPW * RUN) + 1;
            if (dd < nrun) {
                const float4 f = pvs[(v - 1) * kPW + p];
                PixelView pv;
                pv.a0 = f.x; pv.a1 = f.y; pv.a2 = f.z; pv.c = f.w;
                const ViewParams& q = vpb[v];
                const FootRec1 r = make_record1(pv, q.g[0], q.g[1], q.g[2], __ldg(tinv + (size_t)(b * V + v) * D + d0 + run0 + dd), h, w);
                rec_w[i] = make_float4(r.w00, r.w01, r.w10, r.w11);
                rec_o[i] = r.off;
            }
        }
        __syncwarp();
        if (!active) continue;

        // ---- views 1 .. V-2: blend, park the samples in this thread's slice of shared memory
#pragma unroll
        for (int v = 1; v < V - 1; ++v) {
            const char* fv = fb + (size_t)v * view_bytes;
            int key = -1;
            for (int dd = 0; dd < nrun; ++dd) {
                const float4 wt = rec_w[((v - 1) * RUN + dd) * kPW + p8];
                const int of = rec_o[((v - 1) * RUN + dd) * kPW + p8];
                const int changed = of != key;
                key = of;
                const char* pa = fv + (unsigned)of;
                const char* pb = pa + line_bytes;
                ldg_f8_if(taps[0], pa, changed); ldg_f8_if(taps[1], pa + kC * 4, changed);
                ldg_f8_if(taps[2], pb, changed); ldg_f8_if(taps[3], pb + kC * 4, changed);
                F8 val;
                blend8(wt, taps[0], taps[1], taps[2], taps[3], val);
                float4* dst = vals + (size_t)(((v - 1) * RUN + dd) * 2) * kThreads;
                dst[0] = make_float4(val.v[0].x, val.v[0].y, val.v[1].x, val.v[1].y);
                dst[kThreads] = make_float4(val.v[2].x, val.v[2].y, val.v[3].x, val.v[3].y);
            }
        }
        // ---- last view: blend, variance over all V samples, stream the row out
        {
            constexpr int v = V - 1;
            const char* fv = fb + (size_t)v * view_bytes;
            int key = -1;
            size_t vox = ((size_t)(b * D + d0 + run0) * h + py) * w + px;
            for (int dd = 0; dd < nrun; ++dd, vox += plane) {
                const float4 wt = rec_w[((v - 1) * RUN + dd) * kPW + p8];
                const int of = rec_o[((v - 1) * RUN + dd) * kPW + p8];
                const int changed = of != key;
                key = of;
                const char* pa = fv + (unsigned)of;
                const char* pb = pa + line_bytes;
                ldg_f8_if(taps[0], pa, changed); ldg_f8_if(taps[1], pa + kC * 4, changed);
                ldg_f8_if(taps[2], pb, changed); ldg_f8_if(taps[3], pb + kC * 4, changed);
                F8 val, out;
                blend8(wt, taps[0], taps[1], taps[2], taps[3], val);
                if (V == 2) {
                    // two samples: var = ((a - b) / 2)^2
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 hd = __fmul2_rn(sub2(val.v[k], ref.v[k]), make_float2(0.5f, 0.5f));
                        out.v[k] = __fmul2_rn(hd, hd);
                    }
                } else if (V == 3) {
                    const float4* src = vals + (size_t)(dd * 2) * kThreads;
                    const float4 s0 = src[0], s1 = src[kThreads];
                    out.v[0] = variance3(ref.v[0], lo2(s0), val.v[0]);
                    out.v[1] = variance3(ref.v[1], hi2(s0), val.v[1]);
                    out.v[2] = variance3(ref.v[2], lo2(s1), val.v[2]);
                    out.v[3] = variance3(ref.v[3], hi2(s1), val.v[3]);
                } else {
                    // two-pass as costvolume.py:12-14 (mean, then sum (x - mean)^2, / V); the parked samples are read twice
                    F8 sum;
#pragma unroll
                    for (int k = 0; k < 4; ++k) sum.v[k] = __fadd2_rn(ref.v[k], val.v[k]);
#pragma unroll
                    for (int u = 1; u < V - 1; ++u) {
                        const float4* src = vals + (size_t)(((u - 1) * RUN + dd) * 2) * kThreads;
                        const float4 s0 = src[0], s1 = src[kThreads];
                        sum.v[0] = __fadd2_rn(sum.v[0], lo2(s0)); sum.v[1] = __fadd2_rn(sum.v[1], hi2(s0));
                        sum.v[2] = __fadd2_rn(sum.v[2], lo2(s1)); sum.v[3] = __fadd2_rn(sum.v[3], hi2(s1));
                    }
                    const float2 ninv = make_float2(-invV, -invV);
                    F8 nmean, acc;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        nmean.v[k] = __fmul2_rn(sum.v[k], ninv);
                        const float2 da = __fadd2_rn(ref.v[k], nmean.v[k]), db = __fadd2_rn(val.v[k], nmean.v[k]);
                        acc.v[k] = __ffma2_rn(db, db, __fmul2_rn(da, da));
                    }
#pragma unroll
                    for (int u = 1; u < V - 1; ++u) {
                        const float4* src = vals + (size_t)(((u - 1) * RUN + dd) * 2) * kThreads;
                        const float4 s0 = src[0], s1 = src[kThreads];
                        float2 d;
                        d = __fadd2_rn(lo2(s0), nmean.v[0]); acc.v[0] = __ffma2_rn(d, d, acc.v[0]);
                        d = __fadd2_rn(hi2(s0), nmean.v[1]); acc.v[1] = __ffma2_rn(d, d, acc.v[1]);
                        d = __fadd2_rn(lo2(s1), nmean.v[2]); acc.v[2] = __ffma2_rn(d, d, acc.v[2]);
                        d = __fadd2_rn(hi2(s1), nmean.v[3]); acc.v[3] = __ffma2_rn(d, d, acc.v[3]);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) out.v[k] = __fmul2_rn(acc.v[k], make_float2(invV, invV));
                }
                if (BF16OUT) {
                    uint4 u;
                    u.x = pack_bf16x2(out.v[0].x, out.v[0].y); u.y = pack_bf16x2(out.v[1].x, out.v[1].y);
                    u.z = pack_bf16x2(out.v[2].x, out.v[2].y); u.w = pack_bf16x2(out.v[3].x, out.v[3].y);
                    st_cs_u4(reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(cost) + vox * kC + 8 * cq), u);
                } else {
                    st_cs_f8(reinterpret_cast<float*>(cost) + vox * kC + 8 * cq, out);
                }
            }
        }
    }
}

struct Plan3 {
    dim3 grid;
    int dchunk, tiles_x;
};

Plan3 make_plan3(int B, int D, int h, int w, int run, int ctas_per_sm) {
    Plan3 p;
    p.tiles_x = (w + kTX - 1) / kTX;
    const int tiles_y = (h + kTY - 1) / kTY;
    const long tiles = (long)p.tiles_x * tiles_y * B;
    // >= 4 waves of resident CTAs on 148 SMs, in depth runs that are multiples of the staged run
    long nchunks = (4L * ctas_per_sm * 148 + tiles - 1) / tiles;
    const long maxchunks = (D + 4 * run - 1) / (4 * run);
    if (nchunks > maxchunks) nchunks = maxchunks;
    if (nchunks < 1) nchunks = 1;
    p.dchunk = (int)((D + nchunks - 1) / nchunks);
    p.dchunk = (p.dchunk + run - 1) / run * run;
    if (const char* e = getenv("MVSB200_DCHUNK")) {
        const int v = atoi(e);
        if (v > 0) p.dchunk = v < D ? v : D;           // clamped to [1, D]
    }
    p.grid = dim3((unsigned)(p.tiles_x * tiles_y), (unsigned)((D + p.dchunk - 1) / p.dchunk), (unsigned)B);
    return p;
}

template <int V, int MINB>
int launch_fwd3(const float* feat, const float* vp, const float* tinv, void* cost, int dtype, int B, int D, int h, int w,
                cudaStream_t st) {
    const Plan3 p = make_plan3(B, D, h, w, Cfg3<V>::kRun, MINB);
    const size_t smem = Cfg3<V>::kSmemFwd;
    MVS_REQUIRE(p.grid.y <= 65535 && (long)h * w < (1L << 24), "warp_variance_fwd: volume too large");
    // per device / context attribute: set on every launch (cheap)
    if (dtype == MVSB200_BF16) {
        MVS_CUDA(cudaFuncSetAttribute(warp_variance_fwd3_kernel<V, true, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        warp_variance_fwd3_kernel<V, true, MINB><<<p.grid, kThreads, smem, st>>>((const char*)feat, (const ViewParams*)vp, tinv, cost, D,
                                                                                 h, w, p.dchunk, p.tiles_x);
    } else {
        MVS_CUDA(cudaFuncSetAttribute(warp_variance_fwd3_kernel<V, false, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        warp_variance_fwd3_kernel<V, false, MINB><<<p.grid, kThreads, smem, st>>>((const char*)feat, (const ViewParams*)vp, tinv, cost, D,
                                                                                  h, w, p.dchunk, p.tiles_x);
    }
    MVS_CHECK_LAUNCH("warp_variance_fwd3");
    return MVSB200_OK;
}

}  // namespace

namespace mvsb200 {
namespace warp {

// dispatch entry of the third-form forward kernel (called by mvsb200_warp_variance_fwd); 32-byte aligned feature / volume
// pointers (256-bit loads and stores), h, w >= 2
int warp_variance_fwd3(const float* feat, const float* vp, const float* tinv, void* cost, int dtype, int B, int V, int D, int h,
                       int w, cudaStream_t st) {
    MVS_REQUIRE(((uintptr_t)feat & 31u) == 0 && ((uintptr_t)cost & 31u) == 0, "warp_variance_fwd: pointers must be 32-byte aligned");
    switch (V) {
        case 2: return launch_fwd3<2, 3>(feat, vp, tinv, cost, dtype, B, D, h, w, st);
        case 3: return launch_fwd3<3, 3>(feat, vp, tinv, cost, dtype, B, D, h, w, st);
        case 4: return launch_fwd3<4, 3>(feat, vp, tinv, cost, dtype, B, D, h, w, st);
        case 5: return launch_fwd3<5, 3>(feat, vp, tinv, cost, dtype, B, D, h, w, st);
        case 6: return launch_fwd3<6, 2>(feat, vp, tinv, cost, dtype, B, D, h, w, st);
        case 7: return launch_fwd3<7, 2>(feat, vp, tinv, cost, dtype, B, D, h, w, st);
        case 8: return launch_fwd3<8, 2>(feat, vp, tinv, cost, dtype, B, D, h, w, st);
    }
    MVS_FAIL(MVSB200_E_UNSUPPORTED, "warp_variance_fwd: V=%d", V);
}

}  // namespace warp
}  // namespace mvsb200
