// K2, third form: d cost / d features of the fused warp + variance (sm_100a).
//
//   d cost / d f_v = (2/V) (f_v - mean) * gcost   (the mean term cancels), chained through the bilinear taps
//   (autograd of scripts/homography.py:85-90 and scripts/costvolume.py:10-14 w.r.t. the feature maps).
//
// Why a third form.  ncu of the second-form kernel (round 2, profiles/r02_k2_notes.md): 263 warp instructions per warp-plane
// of which 50 are the packed FMAs of the math -- four tap offsets per footprint with 64-bit multiply-add address arithmetic for
// each (54 IMAD), every vector reduction its own predicated branch, 16 predicated moves to clear the accumulators of a moved
// footprint, and the upstream gradient loaded right where it is needed (DRAM latency exposed once per plane).  With the
// reductions and the tap loads switched off it still took 86 % of its time: bound by its own instruction stream.  This form
// keeps the one-pass structure (a lane holds the 2x2 taps AND the 2x2 gradient accumulators of every source view; 8 lanes own a
// pixel, 4 channels each) and removes the overhead:
//   * one clamped base offset per footprint (warp_common.cuh); the gradient map has the feature map's layout, so the same
//     32-bit offset addresses the taps (loads) and the accumulators' home (reductions);
//   * ONE divergent branch per (view, plane) for "the footprint moved": flush the four accumulators (16-byte vector reductions;
//     8 lanes = one full 128-byte line per tap), clear them, reload the four taps -- nothing of it is issued while the
//     footprint stands still (adjacent planes move it by ~0.26 px);
//   * the upstream gradient of plane d+1 is requested before the arithmetic of plane d (prefetching the lines of a footprint
//     one plane before it moves was measured too: 8 % slower -- profiles/r02_k1_k2_notes.md);
//   * footprint records staged per warp (no CTA barrier), weights already 0 on a d == 0 plane (the reference's gradient is NaN
//     there: contribute nothing).
#include "warp_common.cuh"
#include <stdlib.h>

using namespace mvsb200;
using namespace mvsb200::warp;

namespace {

constexpr int kTX = 8, kTY = 4;                    // pixel tile of a CTA: 8 warps x 4 pixels
constexpr int kThreads = 256;

template <int V>
struct Bwd3Cfg {
    static constexpr int kRun = V <= 4 ? 8 : 4;                                          // planes per staged run
    static constexpr int kWRec = (V - 1) * kRun * 4;                                     // records per warp and buffer
    static constexpr size_t kSmem = (size_t)(kThreads / 32) * 2 * kWRec * (sizeof(float4) + sizeof(int));
    static constexpr int kMinBlocks = V <= 3 ? 2 : 1;
};

__device__ __forceinline__ void red_add_f4(char* p, float2 lo, float2 hi) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(lo.x), "f"(lo.y), "f"(hi.x), "f"(hi.y) : "memory");
}

__device__ __forceinline__ float4 ldg_nc_f4(const char* p) {
    float4 t;
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(p));
    return t;
}

// clear a channel pair through a 64-bit zero (one CS2R per pair instead of two moves)
__device__ __forceinline__ void zero2(float2& a) {
    double z;
    asm volatile("mov.f64 %0, 0d0000000000000000;" : "=d"(z));
    const long long b = __double_as_longlong(z);
    a.x = __int_as_float((int)(b & 0xffffffffll));
    a.y = __int_as_float((int)(b >> 32));
}

// w (scalar, broadcast) * t + acc on two channels
__device__ __forceinline__ float2 fma_s(float w, float2 t, float2 acc) { return __ffma2_rn(make_float2(w, w), t, acc); }

__device__ __forceinline__ void blend4(const float4& wt, const float4 (&t)[4], float2& lo, float2& hi) {
    lo = fma_s(wt.w, lo2(t[3]), fma_s(wt.z, lo2(t[2]), fma_s(wt.y, lo2(t[1]), __fmul2_rn(make_float2(wt.x, wt.x), lo2(t[0])))));
    hi = fma_s(wt.w, hi2(t[3]), fma_s(wt.z, hi2(t[2]), fma_s(wt.y, hi2(t[1]), __fmul2_rn(make_float2(wt.x, wt.x), hi2(t[0])))));
}

template <bool BF16G>
struct GRaw {                                        // this lane's 4 channels of one upstream-gradient row, as loaded
    uint4 u;
};
template <bool BF16G>
__device__ __forceinline__ GRaw<BF16G> load_g(const void* gcost, size_t vox, int cg) {
    GRaw<BF16G> r;
    if (BF16G) {
        const uint2 u = __ldcs(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(gcost) + vox * kC + 4 * cg));
        r.u = make_uint4(u.x, u.y, 0u, 0u);
    } else {
        const float4 f = ld_cs_f4(reinterpret_cast<const float4*>(gcost) + vox * kSlots + cg);
        r.u = make_uint4(__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z), __float_as_uint(f.w));
    }
    return r;
}
template <bool BF16G>
__device__ __forceinline__ void unpack_g(const GRaw<BF16G>& r, float2& lo, float2& hi) {
    if (BF16G) {
        lo = make_float2(__uint_as_float(r.u.x << 16), __uint_as_float(r.u.x & 0xffff0000u));
        hi = make_float2(__uint_as_float(r.u.y << 16), __uint_as_float(r.u.y & 0xffff0000u));
    } else {
        lo = make_float2(__uint_as_float(r.u.x), __uint_as_float(r.u.y));
        hi = make_float2(__uint_as_float(r.u.z), __uint_as_float(r.u.w));
    }
}

template <int V, bool BF16G>
__global__ void __launch_bounds__(kThreads, Bwd3Cfg<V>::kMinBlocks)
warp_variance_bwd3_kernel(const char* __restrict__ feat, const ViewParams* __restrict__ vp, const float* __restrict__ tinv,
                          const void* __restrict__ gcost, char* __restrict__ gfeat, int D, int h, int w, int dchunk, int tiles_x) {
    constexpr int RUN = Bwd3Cfg<V>::kRun, WREC = Bwd3Cfg<V>::kWRec, NW = kThreads / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* rec_w = reinterpret_cast<float4*>(smem_raw) + warp * 2 * WREC;              // [buffer][view-1][plane][4 pixels]
    int* rec_o = reinterpret_cast<int*>(reinterpret_cast<float4*>(smem_raw) + NW * 2 * WREC) + warp * 2 * WREC;

    const int b = blockIdx.z, d0 = blockIdx.y * dchunk, nd = min(dchunk, D - d0);
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const unsigned plane = (unsigned)h * (unsigned)w;
    const ViewParams* vpb = vp + (size_t)b * V;

    // ---- staging identity: pixel lane & 3 of this warp's four (fixed per lane: 32 is a multiple of 4)
    const int slot1 = warp * 4 + (lane & 3);
    const float x1 = (float)(tx * kTX + (slot1 & (kTX - 1))), y1 = (float)(ty * kTY + slot1 / kTX);
    auto stage_run = [&](int run0, int buf) {
        const int nrun = min(RUN, nd - run0);
#pragma unroll
        for (int i0 = 0; i0 < WREC; i0 += 32) {
            const int i = i0 + lane;                 // record (v, dd, pixel): pixel = i & 3 == lane & 3
            const int v = i / (RUN * 4) + 1, dd = (i >> 2) % RUN;
            if (i < WREC && dd < nrun) {
                const ViewParams& q = vpb[v];
                const PixelView pv = pixel_view(q, x1, y1);
                const float t = __ldg(tinv + (size_t)(b * V + v) * D + d0 + run0 + dd);
                FootRec1 r = make_record1(pv, q.g[0], q.g[1], q.g[2], t, h, w);
                const bool nan_plane = t != t;       // d == 0 plane
                if (nan_plane) { r.w00 = 0.f; r.w01 = 0.f; r.w10 = 0.f; r.w11 = 0.f; }
                rec_w[buf * WREC + i] = make_float4(r.w00, r.w01, r.w10, r.w11);
                rec_o[buf * WREC + i] = r.off | (nan_plane ? 1 : 0);       // offsets are multiples of 128: bit 0 = plane flag
            }
        }
    };

    // ---- main identity: pixel lane / 8 of the warp's four, channels 4*cg .. 4*cg+3
    const int p4 = lane >> 3, cg = lane & 7;
    const int pl = warp * 4 + p4;
    const int px = tx * kTX + (pl & (kTX - 1)), py = ty * kTY + pl / kTX;
    const bool active = px < w && py < h;
    // 32-bit byte offsets from the tensor base (the launcher checks B*V*h*w*128 < 2^32): the gradient map has the feature map's
    // layout, so ONE offset addresses a tap (feat + o) and its gradient's home (gfeat + o); 64-bit arithmetic only where an
    // address is formed, inside the rarely taken "footprint moved" block
    const unsigned view_bytes = plane * (kC * 4), line_bytes = (unsigned)w * (kC * 4);
    unsigned lane_off = (unsigned)(b * V) * view_bytes + cg * 16;  // this lane's 16 bytes of row 0 of the sample's view 0

    stage_run(0, 0);

    float2 ref[2], gref[2];                          // reference view: H = I on every plane => one sample, one footprint per pixel
    FootRec1 rref;
    {
        const PixelView pv = pixel_view(vpb[0], (float)px, (float)py);
        rref = make_record1(pv, 0.f, 0.f, 0.f, 0.f, h, w);
        const char* pa = feat + (lane_off + (unsigned)rref.off);
        float4 t[4];
        t[0] = ldg_nc_f4(pa); t[1] = ldg_nc_f4(pa + kC * 4); t[2] = ldg_nc_f4(pa + line_bytes); t[3] = ldg_nc_f4(pa + line_bytes + kC * 4);
        blend4(make_float4(rref.w00, rref.w01, rref.w10, rref.w11), t, ref[0], ref[1]);
        gref[0] = gref[1] = make_float2(0.f, 0.f);
    }

    float4 taps[V][4];
    float2 acc[V][4][2];                             // gradient of the cached footprint of view v: [tap][channel pair]
    int key[V];                                      // its record offset (< 0: none yet)
#pragma unroll
    for (int v = 1; v < V; ++v) {
        key[v] = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            taps[v][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            acc[v][j][0] = acc[v][j][1] = make_float2(0.f, 0.f);
        }
    }
    const float invV = 1.0f / (float)V, twoV = 2.0f / (float)V;
    const float2 ninv = make_float2(-invV, -invV);
    __syncwarp();

    size_t vox = ((size_t)(b * D + d0) * h + py) * w + px;
    GRaw<BF16G> gnext;
    if (active) gnext = load_g<BF16G>(gcost, vox, cg);

    int buf = 0;
    for (int run0 = 0; run0 < nd; run0 += RUN, buf ^= 1) {
        const int nrun = min(RUN, nd - run0);
        if (run0 + RUN < nd) stage_run(run0 + RUN, buf ^ 1);
        if (active) {
            const float4* rw = rec_w + buf * WREC + p4;
            const int* ro = rec_o + buf * WREC + p4;
            for (int dd = 0; dd < nrun; ++dd) {
                asm volatile("" : "+r"(lane_off));   // keep the lane's base offset in its register (not re-derived in the block below)
                const GRaw<BF16G> gcur = gnext;
                vox += plane;
                if (run0 + dd + 1 < nd) gnext = load_g<BF16G>(gcost, vox, cg);      // next plane's row: in flight during this plane's math
                float2 val[V][2];
                float4 wt[V];
                int nan_plane = 0;
#pragma unroll
                for (int v = 1; v < V; ++v) {
                    wt[v] = rw[((v - 1) * RUN + dd) * 4];
                    const int of = ro[((v - 1) * RUN + dd) * 4];
                    nan_plane = of & 1;
                    if (of != key[v]) {              // the footprint moved: fetch the new taps, flush its gradient, clear
                        const unsigned vo = lane_off + (unsigned)v * view_bytes;
                        const char* pa = feat + (vo + (unsigned)(of & ~1));
                        const char* pb = pa + line_bytes;
                        taps[v][0] = ldg_nc_f4(pa); taps[v][1] = ldg_nc_f4(pa + kC * 4);
                        taps[v][2] = ldg_nc_f4(pb); taps[v][3] = ldg_nc_f4(pb + kC * 4);
                        if (key[v] >= 0) {
                            char* qa = gfeat + (vo + (unsigned)(key[v] & ~1));
                            char* qb = qa + line_bytes;
                            red_add_f4(qa, acc[v][0][0], acc[v][0][1]);
                            red_add_f4(qa + kC * 4, acc[v][1][0], acc[v][1][1]);
                            red_add_f4(qb, acc[v][2][0], acc[v][2][1]);
                            red_add_f4(qb + kC * 4, acc[v][3][0], acc[v][3][1]);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) { zero2(acc[v][j][0]); zero2(acc[v][j][1]); }
                        key[v] = of;
                    }
                    blend4(wt[v], taps[v], val[v][0], val[v][1]);
                }
                float2 g[2];
                unpack_g<BF16G>(gcur, g[0], g[1]);
                const float gsc = nan_plane ? 0.f : twoV;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    float2 sum = ref[k];
#pragma unroll
                    for (int v = 1; v < V; ++v) sum = __fadd2_rn(sum, val[v][k]);
                    const float2 gk = __fmul2_rn(g[k], make_float2(gsc, gsc));
                    const float2 ng = __fmul2_rn(__fmul2_rn(sum, ninv), gk);            // -mean * gk
                    gref[k] = __fadd2_rn(gref[k], __ffma2_rn(ref[k], gk, ng));
#pragma unroll
                    for (int v = 1; v < V; ++v) {
                        const float2 gv = __ffma2_rn(val[v][k], gk, ng);               // (f_v - mean) * gk
                        acc[v][0][k] = fma_s(wt[v].x, gv, acc[v][0][k]);
                        acc[v][1][k] = fma_s(wt[v].y, gv, acc[v][1][k]);
                        acc[v][2][k] = fma_s(wt[v].z, gv, acc[v][2][k]);
                        acc[v][3][k] = fma_s(wt[v].w, gv, acc[v][3][k]);
                    }
                }
            }
        }
        __syncwarp();                                // next buffer staged by every lane, this buffer consumed
    }
    if (!active) return;
    // final flush: the live source-view footprints, then the reference view's constant footprint
#pragma unroll
    for (int v = 1; v < V; ++v) {
        if (key[v] >= 0) {
            char* qa = gfeat + (lane_off + (unsigned)v * view_bytes + (unsigned)(key[v] & ~1));
            char* qb = qa + line_bytes;
            red_add_f4(qa, acc[v][0][0], acc[v][0][1]);
            red_add_f4(qa + kC * 4, acc[v][1][0], acc[v][1][1]);
            red_add_f4(qb, acc[v][2][0], acc[v][2][1]);
            red_add_f4(qb + kC * 4, acc[v][3][0], acc[v][3][1]);
        }
    }
    {
        char* qa = gfeat + (lane_off + (unsigned)rref.off);
        char* qb = qa + line_bytes;
        const float wr[4] = {rref.w00, rref.w01, rref.w10, rref.w11};
        char* q[4] = {qa, qa + kC * 4, qb, qb + kC * 4};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 s = make_float2(wr[j], wr[j]);
            red_add_f4(q[j], __fmul2_rn(s, gref[0]), __fmul2_rn(s, gref[1]));
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
    // >= 4 waves of resident CTAs on 148 SMs; long depth chunks keep the footprint accumulators alive (one flush of the
    // reference view's footprint and of the live source footprints per pixel and chunk)
    long nchunks = (4L * ctas_per_sm * 148 + tiles - 1) / tiles;
    const long maxchunks = (D + 2 * run - 1) / (2 * run);
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

template <int V>
int launch_bwd3(const float* feat, const float* vp, const float* tinv, const void* gcost, int dtype, float* gfeat, int B, int D,
                int h, int w, cudaStream_t st) {
    const Plan3 p = make_plan3(B, D, h, w, Bwd3Cfg<V>::kRun, Bwd3Cfg<V>::kMinBlocks);
    const size_t smem = Bwd3Cfg<V>::kSmem;
    MVS_REQUIRE(p.grid.y <= 65535 && (long)h * w < (1L << 24), "warp_variance_bwd: volume too large");
    MVS_REQUIRE((unsigned long long)B * V * h * w * kC * 4 < (1ull << 32), "warp_variance_bwd: feature maps of one call must stay below 4 GB");
    // per device / context attribute: set on every launch (cheap)
    if (dtype == MVSB200_BF16) {
        MVS_CUDA(cudaFuncSetAttribute(warp_variance_bwd3_kernel<V, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        warp_variance_bwd3_kernel<V, true><<<p.grid, kThreads, smem, st>>>((const char*)feat, (const ViewParams*)vp, tinv, gcost,
                                                                          (char*)gfeat, D, h, w, p.dchunk, p.tiles_x);
    } else {
        MVS_CUDA(cudaFuncSetAttribute(warp_variance_bwd3_kernel<V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        warp_variance_bwd3_kernel<V, false><<<p.grid, kThreads, smem, st>>>((const char*)feat, (const ViewParams*)vp, tinv, gcost,
                                                                           (char*)gfeat, D, h, w, p.dchunk, p.tiles_x);
    }
    MVS_CHECK_LAUNCH("warp_variance_bwd3");
    return MVSB200_OK;
}

}  // namespace

namespace mvsb200 {
namespace warp {

// dispatch entry of the third-form backward kernel (called by mvsb200_warp_variance_bwd after gfeat was zeroed); h, w >= 2
int warp_variance_bwd3(const float* feat, const float* vp, const float* tinv, const void* gcost, int dtype, float* gfeat, int B,
                       int V, int D, int h, int w, cudaStream_t st) {
    switch (V) {
        case 2: return launch_bwd3<2>(feat, vp, tinv, gcost, dtype, gfeat, B, D, h, w, st);
        case 3: return launch_bwd3<3>(feat, vp, tinv, gcost, dtype, gfeat, B, D, h, w, st);
        case 4: return launch_bwd3<4>(feat, vp, tinv, gcost, dtype, gfeat, B, D, h, w, st);
        case 5: return launch_bwd3<5>(feat, vp, tinv, gcost, dtype, gfeat, B, D, h, w, st);
        case 6: return launch_bwd3<6>(feat, vp, tinv, gcost, dtype, gfeat, B, D, h, w, st);
        case 7: return launch_bwd3<7>(feat, vp, tinv, gcost, dtype, gfeat, B, D, h, w, st);
        case 8: return launch_bwd3<8>(feat, vp, tinv, gcost, dtype, gfeat, B, D, h, w, st);
    }
    MVS_FAIL(MVSB200_E_UNSUPPORTED, "warp_variance_bwd: V=%d", V);
}

}  // namespace warp
}  // namespace mvsb200
