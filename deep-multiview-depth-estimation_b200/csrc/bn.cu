// K3b: train-mode BatchNorm3d (+ReLU) over channel-last volumes, forward and backward (sm_100a).
//
// Reference semantics (citations into /root/reference/scripts):
//   model.py:241-247   BatchNorm3d(eps=1e-5, momentum=0.1), ReLU
//   model.py:101-121   y = ReLU(BN(conv(x))) after every convolution of CostVolumeReg except conv_out
//   train.py:61 / test.py:61   the modules run in train mode => batch statistics over (B, D, h, w)
//
// Layout: a [B, C, D, h, w] volume in channels_last_3d strides is M = B*D*h*w rows of C contiguous channels.  Every
// kernel below is one streaming pass (HBM-bound): a thread owns one 8-channel chunk position (so its channel group is
// fixed over a grid-stride loop whose stride is a multiple of C/8) and moves 16 bytes (bf16) / 32 bytes (fp32) per row.
//   bn_stats          sum, sum of squares per channel  -> per-CTA partials -> fixed-order fp64 finalize (deterministic)
//   bn_relu_fwd       y = max(x*scale + shift, 0)
//   bn_relu_bwd_reduce   dbeta = sum g, dgamma = sum g*xhat with g = gy * [x*scale+shift > 0]
//   bn_relu_bwd_apply    dx = gamma*invstd * (g - dbeta/M - xhat*dgamma/M)
#include "common.cuh"

using namespace mvsb200;

namespace {

constexpr int kBnThreads = 256;
constexpr int kBnBlocks = 148 * 4;      // persistent grid: 4 CTAs of 256 threads per SM
constexpr int kMaxC = 64;

template <typename T>
struct Chunk;                            // 8 consecutive channels of one row
// Raw form: the loads of several chunks are issued back to back and unpacked afterwards (kBnUnroll chunks in flight per thread:
// with one 16-byte load in flight per thread, 1024 threads per SM keep 16-32 KB in flight -- ~3 TB/s at HBM latency; measured)
constexpr int kBnUnroll = 4;
struct RawF { float4 a, b; };
template <>
struct Chunk<float> {
    typedef RawF Raw;
    static __device__ __forceinline__ Raw ldraw(const float* p) {
        Raw r; r.a = __ldcs(reinterpret_cast<const float4*>(p)); r.b = __ldcs(reinterpret_cast<const float4*>(p) + 1); return r;
    }
    static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
        v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
    }
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(p)), b = __ldcs(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ float round(float v) { return v; }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
        __stcs(reinterpret_cast<float4*>(p) + 1, make_float4(v[4], v[5], v[6], v[7]));
    }
};
template <>
struct Chunk<__nv_bfloat16> {
    typedef uint4 Raw;
    static __device__ __forceinline__ Raw ldraw(const __nv_bfloat16* p) { return __ldcs(reinterpret_cast<const uint4*>(p)); }
    static __device__ __forceinline__ void unpack(const Raw& u, float (&v)[8]) {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 u = __ldcs(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ float round(float v) { return __bfloat162float(__float2bfloat16(v)); }   // the value a bf16 store keeps
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        __stcs(reinterpret_cast<uint4*>(p), make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                        pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
    }
};

// Sum each thread's 2 x 8 accumulators over the CTA per channel and write the CTA's partial row [2][C].
// Lanes with equal (lane % CPR) hold the same channel group.
__device__ __forceinline__ void block_reduce_to_partial(float (&a)[8], float (&b)[8], int cpr, int C, float* partial_row) {
    __shared__ float s_part[kBnThreads / 32][2][kMaxC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        for (int o = 16; o >= cpr; o >>= 1) {
            a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
            b[i] += __shfl_xor_sync(0xffffffffu, b[i], o);
        }
    }
    if (lane < cpr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s_part[warp][0][lane * 8 + i] = a[i];
            s_part[warp][1][lane * 8 + i] = b[i];
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < 2 * C) {
        const int which = threadIdx.x / C, c = threadIdx.x % C;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kBnThreads / 32; ++w) s += s_part[w][which][c];
        partial_row[which * C + c] = s;
    }
}

template <typename T>
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const T* __restrict__ x, long long n_chunks, int C,
                                                              float* __restrict__ partials) {
    const int cpr = C / 8;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * kBnThreads;
    for (long long i = (long long)blockIdx.x * kBnThreads + threadIdx.x; i < n_chunks; i += kBnUnroll * stride) {
        typename Chunk<T>::Raw raw[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u)
            if (i + u * stride < n_chunks) raw[u] = Chunk<T>::ldraw(x + (i + u * stride) * 8);
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            if (i + u * stride < n_chunks) {
                float v[8];
                Chunk<T>::unpack(raw[u], v);
#pragma unroll
                for (int k = 0; k < 8; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
            }
        }
    }
    block_reduce_to_partial(s, q, cpr, C, partials + (size_t)blockIdx.x * 2 * C);
}

// one CTA of 8 x 2C threads: out[0][c] = sum_a / M, out[1][c] = sum_b / M - (sum_a / M)^2   (mode 0: mean, biased variance)
//                            out[0][c] = sum_a,     out[1][c] = sum_b                        (mode 1: raw sums)
// slice s of 8 adds the partial rows k = s, s+8, ... in fp64; the 8 slices are combined in fixed order (deterministic)
constexpr int kFinSlices = 8;
__global__ void __launch_bounds__(kFinSlices * 2 * kMaxC) bn_finalize_kernel(const float* __restrict__ partials, int nblocks, int C,
                                                                            double inv_m, int mode, float* __restrict__ out0,
                                                                            float* __restrict__ out1) {
    __shared__ double s_acc[kFinSlices][2 * kMaxC];
    const int col = threadIdx.x % (2 * kMaxC), slice = threadIdx.x / (2 * kMaxC);
    if (col < 2 * C) {
        double a = 0.0;
        for (int k = slice; k < nblocks; k += kFinSlices) a += (double)partials[(size_t)k * 2 * C + col];
        s_acc[slice][col] = a;
    }
    __syncthreads();
    const int c = threadIdx.x;
    if (c >= C) return;
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int s = 0; s < kFinSlices; ++s) { a += s_acc[s][c]; b += s_acc[s][C + c]; }
    if (mode == 0) {
        const double mean = a * inv_m;
        double var = b * inv_m - mean * mean;
        out0[c] = (float)mean;
        out1[c] = (float)(var > 0.0 ? var : 0.0);
    } else {
        out0[c] = (float)a;
        out1[c] = (float)b;
    }
}

// Statistics finalize + everything per-channel that follows it in a train-mode BatchNorm (scripts/model.py:242-247,
// torch.nn.BatchNorm semantics): mean, biased variance, invstd, the affine map y = x*scale + shift, and the running
// statistics update (momentum, unbiased variance, num_batches_tracked) -- one launch instead of a dozen [C]-sized
// elementwise kernels per BatchNorm (the step had ~1000 launches of a few microseconds each around 160 real ones).
struct BnAffineOut {
    float *mean, *var, *invstd, *scale, *shift;
    float *running_mean, *running_var;       // may be null
    long long* num_batches_tracked;          // may be null
};
__global__ void __launch_bounds__(kFinSlices * 2 * kMaxC) bn_finalize_affine_kernel(const float* __restrict__ partials, int nblocks,
                                                                                   int C, double inv_m, double unbias,
                                                                                   const float* __restrict__ gamma,
                                                                                   const float* __restrict__ beta, double eps,
                                                                                   double momentum, BnAffineOut o) {
    __shared__ double s_acc[kFinSlices][2 * kMaxC];
    const int col = threadIdx.x % (2 * kMaxC), slice = threadIdx.x / (2 * kMaxC);
    if (col < 2 * C) {
        double a = 0.0;
        for (int k = slice; k < nblocks; k += kFinSlices) a += (double)partials[(size_t)k * 2 * C + col];
        s_acc[slice][col] = a;
    }
    __syncthreads();
    const int c = threadIdx.x;
    if (c == 0 && o.num_batches_tracked) *o.num_batches_tracked += 1;
    if (c >= C) return;
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int s = 0; s < kFinSlices; ++s) { a += s_acc[s][c]; b += s_acc[s][C + c]; }
    const double mean = a * inv_m;
    double var = b * inv_m - mean * mean;
    var = var > 0.0 ? var : 0.0;
    // the values downstream kernels read are the fp32 ones: derive invstd / scale / shift from the ROUNDED mean and variance,
    // as the elementwise fp32 expressions this replaces did
    const float meanf = (float)mean, varf = (float)var;
    const float invstd = rsqrtf(varf + (float)eps);
    const float scale = gamma[c] * invstd;
    o.mean[c] = meanf;
    o.var[c] = varf;
    o.invstd[c] = invstd;
    o.scale[c] = scale;
    o.shift[c] = beta[c] - meanf * scale;
    if (o.running_mean) {
        o.running_mean[c] = (float)((1.0 - momentum) * (double)o.running_mean[c] + momentum * (double)meanf);
        o.running_var[c] = (float)((1.0 - momentum) * (double)o.running_var[c] + momentum * (double)varf * unbias);
    }
}

// Space-to-depth addressing: rows of N maps [N, H, W, C] ("shallow") against [N, H/2, W/2, 4C] ("deep", channel
// (py*2 + px)*C + c of pixel (j, i) = channel c of pixel (2j + py, 2i + px)) -- the form in which a 5x5 stride-2 convolution is a 3x3
// one.  The BatchNorm that precedes such a convolution writes its output straight in the deep form (S2D forward kernel) and its
// backward reads the incoming gradient from it, so the permutation costs no pass of its own.
struct S2dGeo { unsigned W, H, qshift; };           // shallow map size, log2(16-byte chunks per shallow pixel)
__device__ __forceinline__ long long s2d_chunk(long long i, const S2dGeo& g) {
    const unsigned k = (unsigned)i & ((1u << g.qshift) - 1u);
    const unsigned pix = (unsigned)(i >> g.qshift);
    const unsigned t = pix / g.W, x = pix - t * g.W;
    const unsigned n = t / g.H, y = t - n * g.H;
    const unsigned deep_pix = (n * (g.H >> 1) + (y >> 1)) * (g.W >> 1) + (x >> 1);
    return ((((long long)deep_pix << 2) + ((y & 1u) * 2u + (x & 1u))) << g.qshift) | k;
}

// ADD: y = round_T(max(x*scale + shift, 0)) + add -- the skip additions of the decoder (scripts/model.py:117-123) folded into the
// apply pass; the normalised value is rounded to the storage type first, so the result is bit-identical to a separate addition of
// the stored tensor (the depth-slab path and the unfused form keep agreeing exactly).
template <typename T, bool ADD = false, bool S2D = false>
__global__ void __launch_bounds__(kBnThreads) bn_relu_fwd_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift, T* __restrict__ y,
                                                                 long long n_chunks, int C, int relu, const T* __restrict__ add = nullptr,
                                                                 S2dGeo geo = S2dGeo{}) {
    const int cpr = C / 8;
    const long long i0 = (long long)blockIdx.x * kBnThreads + threadIdx.x;
    const int cg = (int)(i0 % cpr);
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = scale[cg * 8 + k]; sh[k] = shift[cg * 8 + k]; }
    const long long stride = (long long)gridDim.x * kBnThreads;
    for (long long i = i0; i < n_chunks; i += kBnUnroll * stride) {
        typename Chunk<T>::Raw raw[kBnUnroll], rawa[ADD ? kBnUnroll : 1];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u)
            if (i + u * stride < n_chunks) {
                raw[u] = Chunk<T>::ldraw(x + (i + u * stride) * 8);
                if (ADD) rawa[ADD ? u : 0] = Chunk<T>::ldraw(add + (i + u * stride) * 8);
            }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            if (i + u * stride < n_chunks) {
                float v[8], a[8];
                Chunk<T>::unpack(raw[u], v);
                if (ADD) Chunk<T>::unpack(rawa[ADD ? u : 0], a);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    v[k] = fmaf(v[k], sc[k], sh[k]);
                    if (relu) v[k] = fmaxf(v[k], 0.f);
                    if (ADD) v[k] = Chunk<T>::round(v[k]) + a[k];
                }
                Chunk<T>::store(y + (S2D ? s2d_chunk(i + u * stride, geo) : i + u * stride) * 8, v);
            }
        }
    }
}

template <typename TX, typename TG, bool S2D = false>
__global__ void __launch_bounds__(kBnThreads) bn_relu_bwd_reduce_kernel(const TX* __restrict__ x, const TG* __restrict__ gy,
                                                                        const float* __restrict__ scale,
                                                                        const float* __restrict__ shift,
                                                                        const float* __restrict__ mean,
                                                                        const float* __restrict__ invstd, long long n_chunks,
                                                                        int C, int relu, float* __restrict__ partials,
                                                                        S2dGeo geo = S2dGeo{}) {
    const int cpr = C / 8;
    const long long i0 = (long long)blockIdx.x * kBnThreads + threadIdx.x;
    const int cg = (int)(i0 % cpr);
    float sc[8], sh[8], mu[8], is[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        sc[k] = scale[cg * 8 + k]; sh[k] = shift[cg * 8 + k]; mu[k] = mean[cg * 8 + k]; is[k] = invstd[cg * 8 + k];
    }
    float sg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sgx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * kBnThreads;
    for (long long i = i0; i < n_chunks; i += kBnUnroll * stride) {
        typename Chunk<TX>::Raw rx[kBnUnroll];
        typename Chunk<TG>::Raw rg[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u)
            if (i + u * stride < n_chunks) {
                rx[u] = Chunk<TX>::ldraw(x + (i + u * stride) * 8);
                rg[u] = Chunk<TG>::ldraw(gy + (S2D ? s2d_chunk(i + u * stride, geo) : i + u * stride) * 8);
            }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            if (i + u * stride < n_chunks) {
                float v[8], g[8];
                Chunk<TX>::unpack(rx[u], v);
                Chunk<TG>::unpack(rg[u], g);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float gk = (!relu || fmaf(v[k], sc[k], sh[k]) > 0.f) ? g[k] : 0.f;
                    sg[k] += gk;
                    sgx[k] = fmaf(gk, (v[k] - mu[k]) * is[k], sgx[k]);
                }
            }
        }
    }
    block_reduce_to_partial(sg, sgx, cpr, C, partials + (size_t)blockIdx.x * 2 * C);
}

// dx = gamma*invstd * (g - dbeta/M - xhat*dgamma/M) with g = gy * [x*scale+shift > 0], rewritten per channel as
//   dx = t + [mask] * a1 * gy,   t = a2 * x + a3,   a1 = gamma*invstd,  a2 = -a1 * (dgamma/M) * invstd,
//   a3 = a1 * ((dgamma/M) * mean * invstd - dbeta/M)
// (two FMAs and a select per element instead of seven operations).
struct BwdCoef {
    float sc[8], sh[8], a1[8], a2[8], a3[8];
};
__device__ __forceinline__ void load_bwd_coef(BwdCoef& k, int cg, const float* scale, const float* shift, const float* mean,
                                              const float* invstd, const float* gamma, const float* dbeta, const float* dgamma,
                                              float inv_m) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cg * 8 + j;
        const float is = invstd[c], k0 = gamma[c] * is, k1 = dbeta[c] * inv_m, k2 = dgamma[c] * inv_m;
        k.sc[j] = scale[c]; k.sh[j] = shift[c];
        k.a1[j] = k0;
        k.a2[j] = -k0 * k2 * is;
        k.a3[j] = k0 * (k2 * mean[c] * is - k1);
    }
}
__device__ __forceinline__ void bwd_apply8(const BwdCoef& k, float (&v)[8], const float (&g)[8], int relu, bool has_g) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float t = fmaf(k.a2[j], v[j], k.a3[j]);
        const bool on = has_g && (!relu || fmaf(v[j], k.sc[j], k.sh[j]) > 0.f);
        v[j] = on ? fmaf(k.a1[j], g[j], t) : t;
    }
}

template <typename TX, typename TG, bool S2D = false>
__global__ void __launch_bounds__(kBnThreads) bn_relu_bwd_apply_kernel(const TX* __restrict__ x, const TG* __restrict__ gy,
                                                                       const float* __restrict__ scale,
                                                                       const float* __restrict__ shift,
                                                                       const float* __restrict__ mean,
                                                                       const float* __restrict__ invstd,
                                                                       const float* __restrict__ gamma,
                                                                       const float* __restrict__ dbeta,
                                                                       const float* __restrict__ dgamma, TX* __restrict__ dx,
                                                                       long long n_chunks, int C, int relu, float inv_m,
                                                                       S2dGeo geo = S2dGeo{}) {
    const int cpr = C / 8;
    const long long i0 = (long long)blockIdx.x * kBnThreads + threadIdx.x;
    const int cg = (int)(i0 % cpr);
    BwdCoef k;
    load_bwd_coef(k, cg, scale, shift, mean, invstd, gamma, dbeta, dgamma, inv_m);
    const long long stride = (long long)gridDim.x * kBnThreads;
    for (long long i = i0; i < n_chunks; i += kBnUnroll * stride) {
        typename Chunk<TX>::Raw rx[kBnUnroll];
        typename Chunk<TG>::Raw rg[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u)
            if (i + u * stride < n_chunks) {
                rx[u] = Chunk<TX>::ldraw(x + (i + u * stride) * 8);
                rg[u] = Chunk<TG>::ldraw(gy + (S2D ? s2d_chunk(i + u * stride, geo) : i + u * stride) * 8);
            }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            if (i + u * stride < n_chunks) {
                float v[8], g[8];
                Chunk<TX>::unpack(rx[u], v);
                Chunk<TG>::unpack(rg[u], g);
                bwd_apply8(k, v, g, relu, true);
                Chunk<TX>::store(dx + (i + u * stride) * 8, v);
            }
        }
    }
}


// ---- crop-aware variants -------------------------------------------------------------------------------------------
// The transposed convolutions of the regulariser produce dense canvases whose BatchNorm statistics are over the full
// canvas, but only the central box of the normalised result is ever read (regulariser.py).  These variants normalise
// the full canvas' statistics-wise while writing / reading gradients only on the box [d0,d0+dc) x [h0,h0+hc) x [w0,w0+wc).
struct CropBox {
    int Da, ha, wa;         // allocation that holds the canvas at its origin (the library's transposed conv returns one
                            // extra plane/line/column that the reference crops away)
    int D, h, w;            // canvas: the volume the statistics are taken over
    int d0, h0, w0;         // box origin
    int dc, hc, wc;         // box size
    FastDiv f_wa, f_ha, f_Da, f_w, f_h, f_D, f_wc, f_hc, f_dc;   // the same extents as division constants
    int cshift;             // log2(channel groups per row): C/8 is 1, 2, 4 or 8
};

// chunk index inside the box volume -> chunk index inside the canvas volume
// (32-bit index arithmetic: the host checks that every chunk count is below 2^31)
__device__ __forceinline__ long long box_to_canvas(long long i, int cpr, const CropBox& c) {
    const unsigned iu = (unsigned)i;
    const int cg = (int)(iu & (unsigned)(cpr - 1));
    unsigned r = iu >> c.cshift, ux, uy, ud;
    r = fd_divmod(r, c.f_wc, ux);
    r = fd_divmod(r, c.f_hc, uy);
    const long long b = fd_divmod(r, c.f_dc, ud);
    const int x = (int)ux, y = (int)uy, d = (int)ud;
    return ((((b * c.Da + d + c.d0) * c.ha + y + c.h0) * c.wa + x + c.w0)) * cpr + cg;
}
// chunk index inside the canvas -> chunk index inside the allocation
__device__ __forceinline__ long long canvas_to_alloc(long long i, int cpr, const CropBox& c) {
    const unsigned iu = (unsigned)i;
    const int cg = (int)(iu & (unsigned)(cpr - 1));
    unsigned r = iu >> c.cshift, ux, uy, ud;
    r = fd_divmod(r, c.f_w, ux);
    r = fd_divmod(r, c.f_h, uy);
    const long long b = fd_divmod(r, c.f_D, ud);
    const int x = (int)ux, y = (int)uy, d = (int)ud;
    return ((((b * c.Da + d) * c.ha + y) * c.wa + x)) * cpr + cg;
}
// chunk index inside the allocation -> chunk index inside the box (>= 0), -1 inside the canvas but outside the box,
// -2 outside the canvas
__device__ __forceinline__ long long alloc_to_box(long long i, int cpr, const CropBox& c) {
    const unsigned iu = (unsigned)i;
    const int cg = (int)(iu & (unsigned)(cpr - 1));
    unsigned r = iu >> c.cshift, ux, uy, ud;
    r = fd_divmod(r, c.f_wa, ux);
    r = fd_divmod(r, c.f_ha, uy);
    const long long b = fd_divmod(r, c.f_Da, ud);
    const int xa = (int)ux, ya = (int)uy, da = (int)ud;
    if (xa >= c.w || ya >= c.h || da >= c.D) return -2;
    const int x = xa - c.w0, y = ya - c.h0, d = da - c.d0;
    if ((unsigned)x >= (unsigned)c.wc || (unsigned)y >= (unsigned)c.hc || (unsigned)d >= (unsigned)c.dc) return -1;
    return ((((b * c.dc + d) * c.hc + y) * c.wc + x)) * cpr + cg;
}

template <typename T>
__global__ void __launch_bounds__(kBnThreads) bn_stats_geo_kernel(const T* __restrict__ x, long long n_canvas_chunks, int C,
                                                                  float* __restrict__ partials, CropBox cb) {
    const int cpr = C / 8;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * kBnThreads;
    for (long long i = (long long)blockIdx.x * kBnThreads + threadIdx.x; i < n_canvas_chunks; i += kBnUnroll * stride) {
        typename Chunk<T>::Raw raw[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u)
            if (i + u * stride < n_canvas_chunks) raw[u] = Chunk<T>::ldraw(x + canvas_to_alloc(i + u * stride, cpr, cb) * 8);
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            if (i + u * stride < n_canvas_chunks) {
                float v[8];
                Chunk<T>::unpack(raw[u], v);
#pragma unroll
                for (int k = 0; k < 8; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
            }
        }
    }
    block_reduce_to_partial(s, q, cpr, C, partials + (size_t)blockIdx.x * 2 * C);
}

template <typename T, bool ADD = false>
__global__ void __launch_bounds__(kBnThreads) bn_relu_fwd_crop_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                                      const float* __restrict__ shift, T* __restrict__ y,
                                                                      long long n_box_chunks, int C, int relu, CropBox cb,
                                                                      const T* __restrict__ add = nullptr) {
    const int cpr = C / 8;
    const long long i0 = (long long)blockIdx.x * kBnThreads + threadIdx.x;
    const int cg = (int)(i0 % cpr);
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = scale[cg * 8 + k]; sh[k] = shift[cg * 8 + k]; }
    const long long stride = (long long)gridDim.x * kBnThreads;
    for (long long i = i0; i < n_box_chunks; i += kBnUnroll * stride) {
        typename Chunk<T>::Raw raw[kBnUnroll], rawa[ADD ? kBnUnroll : 1];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u)
            if (i + u * stride < n_box_chunks) {
                raw[u] = Chunk<T>::ldraw(x + box_to_canvas(i + u * stride, cpr, cb) * 8);
                if (ADD) rawa[ADD ? u : 0] = Chunk<T>::ldraw(add + (i + u * stride) * 8);        // the addend lives on the box, as y does
            }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            if (i + u * stride < n_box_chunks) {
                float v[8], a[8];
                Chunk<T>::unpack(raw[u], v);
                if (ADD) Chunk<T>::unpack(rawa[ADD ? u : 0], a);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    v[k] = fmaf(v[k], sc[k], sh[k]);
                    if (relu) v[k] = fmaxf(v[k], 0.f);
                    if (ADD) v[k] = Chunk<T>::round(v[k]) + a[k];
                }
                Chunk<T>::store(y + (i + u * stride) * 8, v);
            }
        }
    }
}

// Line-major (see the apply kernel below): a warp takes one line (b, d, y) of the BOX at a time; the gradient is zero outside.
template <typename TX, typename TG>
__global__ void __launch_bounds__(kBnThreads) bn_relu_bwd_reduce_crop_kernel(
    const TX* __restrict__ x, const TG* __restrict__ gy, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ mean, const float* __restrict__ invstd, long long n_box_lines, int C, int relu,
    float* __restrict__ partials, CropBox cb) {
    const int cpr = C / 8;
    const int lane = threadIdx.x & 31;
    const int cg = lane & (cpr - 1);
    float sc[8], sh[8], mu[8], is[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        sc[k] = scale[cg * 8 + k]; sh[k] = shift[cg * 8 + k]; mu[k] = mean[cg * 8 + k]; is[k] = invstd[cg * 8 + k];
    }
    float sg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sgx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long warps = (long long)gridDim.x * (kBnThreads / 32);
    const int line_chunks = cb.wc * cpr;
    for (long long line = (long long)blockIdx.x * (kBnThreads / 32) + (threadIdx.x >> 5); line < n_box_lines; line += warps) {
        unsigned uy, ud;
        unsigned r = fd_divmod((unsigned)line, cb.f_hc, uy);
        const unsigned b = fd_divmod(r, cb.f_dc, ud);
        const size_t xbase = ((((size_t)b * cb.Da + ud + cb.d0) * cb.ha + uy + cb.h0) * cb.wa + cb.w0) * cpr;
        const size_t gbase = (size_t)line * line_chunks;
        for (int j0 = lane; j0 < line_chunks; j0 += 32 * kBnUnroll) {
            typename Chunk<TX>::Raw rx[kBnUnroll];
            typename Chunk<TG>::Raw rg[kBnUnroll];
#pragma unroll
            for (int u = 0; u < kBnUnroll; ++u) {
                const int j = j0 + 32 * u;
                if (j < line_chunks) { rx[u] = Chunk<TX>::ldraw(x + (xbase + j) * 8); rg[u] = Chunk<TG>::ldraw(gy + (gbase + j) * 8); }
            }
#pragma unroll
            for (int u = 0; u < kBnUnroll; ++u) {
                if (j0 + 32 * u < line_chunks) {
                    float v[8], g[8];
                    Chunk<TX>::unpack(rx[u], v);
                    Chunk<TG>::unpack(rg[u], g);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float gk = (!relu || fmaf(v[k], sc[k], sh[k]) > 0.f) ? g[k] : 0.f;
                        sg[k] += gk;
                        sgx[k] = fmaf(gk, (v[k] - mu[k]) * is[k], sgx[k]);
                    }
                }
            }
        }
    }
    block_reduce_to_partial(sg, sgx, cpr, C, partials + (size_t)blockIdx.x * 2 * C);
}

// Line-major form: a warp takes one line (b, d, y) of the ALLOCATION at a time -- where the line sits relative to the canvas
// and to the box is decided once per line, a chunk inside the line costs a shift, a mask and two compares (the chunk-major
// form decoded every 16-byte chunk with six multiply-high divisions: ~80 of its ~220 instructions per chunk).
template <typename TX, typename TG>
__global__ void __launch_bounds__(kBnThreads) bn_relu_bwd_apply_crop_kernel(
    const TX* __restrict__ x, const TG* __restrict__ gy, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
    const float* __restrict__ dbeta, const float* __restrict__ dgamma, TX* __restrict__ dx, long long n_lines, int C, int relu,
    float inv_m, CropBox cb) {
    const int cpr = C / 8;
    const int lane = threadIdx.x & 31;
    const int cg = lane & (cpr - 1);                 // 32 % cpr == 0: a lane keeps its channel group along a line
    BwdCoef k;
    load_bwd_coef(k, cg, scale, shift, mean, invstd, gamma, dbeta, dgamma, inv_m);
    const long long warps = (long long)gridDim.x * (kBnThreads / 32);
    const int line_chunks = cb.wa * cpr;
    for (long long line = (long long)blockIdx.x * (kBnThreads / 32) + (threadIdx.x >> 5); line < n_lines; line += warps) {
        unsigned uy, ud;
        unsigned r = fd_divmod((unsigned)line, cb.f_ha, uy);
        const unsigned b = fd_divmod(r, cb.f_Da, ud);
        const int ya = (int)uy, da = (int)ud;
        const bool in_canvas = ya < cb.h && da < cb.D;
        const int yb = ya - cb.h0, db = da - cb.d0;
        const bool in_box_line = in_canvas && (unsigned)yb < (unsigned)cb.hc && (unsigned)db < (unsigned)cb.dc;
        const size_t xbase = (size_t)line * line_chunks;                                     // chunk index of the line's first chunk
        const size_t gbase = (((size_t)b * cb.dc + (in_box_line ? db : 0)) * cb.hc + (in_box_line ? yb : 0)) * cb.wc * cpr;
        for (int j0 = lane; j0 < line_chunks; j0 += 32 * kBnUnroll) {
            typename Chunk<TX>::Raw rx[kBnUnroll];
            typename Chunk<TG>::Raw rg[kBnUnroll];
            int kind[kBnUnroll];                                  // 0: beyond the line, 1: allocation slack (zero), 2: x only, 3: x and gy
#pragma unroll
            for (int u = 0; u < kBnUnroll; ++u) {
                const int j = j0 + 32 * u, xa = j >> cb.cshift;
                kind[u] = j >= line_chunks ? 0 : ((!in_canvas || xa >= cb.w) ? 1 : 2);
                if (kind[u] == 2) {
                    rx[u] = Chunk<TX>::ldraw(x + (xbase + j) * 8);
                    const int xb = xa - cb.w0;
                    if (in_box_line && (unsigned)xb < (unsigned)cb.wc) {
                        rg[u] = Chunk<TG>::ldraw(gy + (gbase + (size_t)xb * cpr + cg) * 8);
                        kind[u] = 3;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kBnUnroll; ++u) {
                const int j = j0 + 32 * u;
                if (kind[u] == 0) continue;
                float v[8], g[8];
                if (kind[u] == 1) {                               // allocation slack outside the canvas: no gradient
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = 0.f;
                } else {
                    Chunk<TX>::unpack(rx[u], v);
                    if (kind[u] == 3) {
                        Chunk<TG>::unpack(rg[u], g);
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) g[q] = 0.f;
                    }
                    bwd_apply8(k, v, g, relu, kind[u] == 3);
                }
                Chunk<TX>::store(dx + (xbase + j) * 8, v);
            }
        }
    }
}

int check_bn(const void* x, long long M, int C, const char* name) {
    MVS_REQUIRE(x && aligned16(x), "%s: null or misaligned volume", name);
    MVS_REQUIRE(C == 8 || C == 16 || C == 32 || C == 64, "%s: C must be 8, 16, 32 or 64 (got %d)", name, C);
    MVS_REQUIRE(M >= 1, "%s: empty volume", name);
    return MVSB200_OK;
}

int grid_for(long long n_chunks) {
    long long b = (n_chunks + kBnThreads - 1) / kBnThreads;
    return (int)(b < kBnBlocks ? b : kBnBlocks);
}

}  // namespace

extern "C" int64_t mvsb200_bn_workspace_floats(void) { return (int64_t)kBnBlocks * 2 * kMaxC; }

extern "C" int mvsb200_bn_stats(const void* x, int dtype, int64_t M, int C, float* workspace, float* mean, float* var,
                                void* stream) {
    if (int rc = check_bn(x, M, C, "bn_stats")) return rc;
    MVS_REQUIRE(workspace && mean && var, "bn_stats: null output");
    MVS_REQUIRE(dtype == MVSB200_F32 || dtype == MVSB200_BF16, "bn_stats: bad dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_chunks = (long long)M * C / 8;
    const int grid = grid_for(n_chunks);
    if (dtype == MVSB200_BF16)
        bn_stats_kernel<__nv_bfloat16><<<grid, kBnThreads, 0, st>>>((const __nv_bfloat16*)x, n_chunks, C, workspace);
    else
        bn_stats_kernel<float><<<grid, kBnThreads, 0, st>>>((const float*)x, n_chunks, C, workspace);
    MVS_CHECK_LAUNCH("bn_stats");
    bn_finalize_kernel<<<1, kFinSlices * 2 * kMaxC, 0, st>>>(workspace, grid, C, 1.0 / (double)M, 0, mean, var);
    MVS_CHECK_LAUNCH("bn_finalize");
    return MVSB200_OK;
}

extern "C" int mvsb200_bn_relu_fwd(const void* x, int dtype, const float* scale, const float* shift, void* y, int relu,
                                   int64_t M, int C, void* stream) {
    if (int rc = check_bn(x, M, C, "bn_relu_fwd")) return rc;
    MVS_REQUIRE(y && aligned16(y) && scale && shift, "bn_relu_fwd: null or misaligned argument");
    MVS_REQUIRE(dtype == MVSB200_F32 || dtype == MVSB200_BF16, "bn_relu_fwd: bad dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_chunks = (long long)M * C / 8;
    const int grid = grid_for(n_chunks);
    if (dtype == MVSB200_BF16)
        bn_relu_fwd_kernel<__nv_bfloat16><<<grid, kBnThreads, 0, st>>>((const __nv_bfloat16*)x, scale, shift,
                                                                       (__nv_bfloat16*)y, n_chunks, C, relu);
    else
        bn_relu_fwd_kernel<float><<<grid, kBnThreads, 0, st>>>((const float*)x, scale, shift, (float*)y, n_chunks, C, relu);
    MVS_CHECK_LAUNCH("bn_relu_fwd");
    return MVSB200_OK;
}

template <typename TX, typename TG>
static int bn_bwd_impl(const void* x, const void* gy, const float* scale, const float* shift, const float* mean,
                       const float* invstd, const float* gamma, float* workspace, float* dbeta, float* dgamma, void* dx,
                       int relu, int64_t M, int C, cudaStream_t st, const S2dGeo* s2d = nullptr) {
    const long long n_chunks = (long long)M * C / 8;
    const int grid = grid_for(n_chunks);
    if (s2d)
        bn_relu_bwd_reduce_kernel<TX, TG, true><<<grid, kBnThreads, 0, st>>>((const TX*)x, (const TG*)gy, scale, shift, mean, invstd,
                                                                             n_chunks, C, relu, workspace, *s2d);
    else
        bn_relu_bwd_reduce_kernel<TX, TG><<<grid, kBnThreads, 0, st>>>((const TX*)x, (const TG*)gy, scale, shift, mean, invstd,
                                                                       n_chunks, C, relu, workspace);
    MVS_CHECK_LAUNCH("bn_relu_bwd_reduce");
    bn_finalize_kernel<<<1, kFinSlices * 2 * kMaxC, 0, st>>>(workspace, grid, C, 1.0, 1, dbeta, dgamma);
    MVS_CHECK_LAUNCH("bn_finalize");
    if (s2d)
        bn_relu_bwd_apply_kernel<TX, TG, true><<<grid, kBnThreads, 0, st>>>((const TX*)x, (const TG*)gy, scale, shift, mean, invstd,
                                                                            gamma, dbeta, dgamma, (TX*)dx, n_chunks, C, relu,
                                                                            (float)(1.0 / (double)M), *s2d);
    else
        bn_relu_bwd_apply_kernel<TX, TG><<<grid, kBnThreads, 0, st>>>((const TX*)x, (const TG*)gy, scale, shift, mean, invstd,
                                                                      gamma, dbeta, dgamma, (TX*)dx, n_chunks, C, relu,
                                                                      (float)(1.0 / (double)M));
    MVS_CHECK_LAUNCH("bn_relu_bwd_apply");
    return MVSB200_OK;
}

static int make_s2d(int64_t M, int C, int H, int W, S2dGeo* g, const char* name) {
    MVS_REQUIRE(H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && M % ((int64_t)H * W) == 0, "%s: maps of %d x %d do not tile M", name, H, W);
    MVS_REQUIRE(C == 8 || C == 16 || C == 32 || C == 64, "%s: C must be 8, 16, 32 or 64 (got %d)", name, C);
    MVS_REQUIRE(M * (C / 8) < (1LL << 31), "%s: too many rows for 32-bit pixel indices", name);
    g->W = (unsigned)W; g->H = (unsigned)H;
    g->qshift = C == 8 ? 0u : (C == 16 ? 1u : (C == 32 ? 2u : 3u));
    return MVSB200_OK;
}

/* BatchNorm apply + ReLU whose OUTPUT is written in the space-to-depth form: x rows [N, H, W, C] -> y rows [N, H/2, W/2, 4C]
 * (M = N*H*W).  The backward reads its incoming gradient from that form. */
extern "C" int mvsb200_bn_relu_fwd_s2d(const void* x, int dtype, const float* scale, const float* shift, void* y, int relu,
                                       int64_t M, int C, int H, int W, void* stream) {
    if (int rc = check_bn(x, M, C, "bn_relu_fwd_s2d")) return rc;
    MVS_REQUIRE(y && aligned16(y) && scale && shift, "bn_relu_fwd_s2d: null or misaligned argument");
    MVS_REQUIRE(dtype == MVSB200_F32 || dtype == MVSB200_BF16, "bn_relu_fwd_s2d: bad dtype %d", dtype);
    S2dGeo g;
    if (int rc = make_s2d(M, C, H, W, &g, "bn_relu_fwd_s2d")) return rc;
    MVS_REQUIRE(dtype == MVSB200_BF16, "bn_relu_fwd_s2d: bf16 rows only (16-byte chunks of 8 channels)");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_chunks = (long long)M * C / 8;
    bn_relu_fwd_kernel<__nv_bfloat16, false, true><<<grid_for(n_chunks), kBnThreads, 0, st>>>(
        (const __nv_bfloat16*)x, scale, shift, (__nv_bfloat16*)y, n_chunks, C, relu, nullptr, g);
    MVS_CHECK_LAUNCH("bn_relu_fwd_s2d");
    return MVSB200_OK;
}

extern "C" int mvsb200_bn_relu_bwd_s2d(const void* x, const void* gy, const float* scale, const float* shift, const float* mean,
                                       const float* invstd, const float* gamma, float* workspace, float* dbeta, float* dgamma,
                                       void* dx, int relu, int64_t M, int C, int H, int W, void* stream) {
    if (int rc = check_bn(x, M, C, "bn_relu_bwd_s2d")) return rc;
    MVS_REQUIRE(gy && aligned16(gy) && dx && aligned16(dx), "bn_relu_bwd_s2d: null or misaligned volume");
    MVS_REQUIRE(scale && shift && mean && invstd && gamma && workspace && dbeta && dgamma, "bn_relu_bwd_s2d: null vector");
    S2dGeo g;
    if (int rc = make_s2d(M, C, H, W, &g, "bn_relu_bwd_s2d")) return rc;
    return bn_bwd_impl<__nv_bfloat16, __nv_bfloat16>(x, gy, scale, shift, mean, invstd, gamma, workspace, dbeta, dgamma, dx, relu, M, C,
                                                     (cudaStream_t)stream, &g);
}

extern "C" int mvsb200_bn_relu_bwd(const void* x, int x_dtype, const void* gy, int g_dtype, const float* scale,
                                   const float* shift, const float* mean, const float* invstd, const float* gamma,
                                   float* workspace, float* dbeta, float* dgamma, void* dx, int relu, int64_t M, int C,
                                   void* stream) {
    if (int rc = check_bn(x, M, C, "bn_relu_bwd")) return rc;
    MVS_REQUIRE(gy && aligned16(gy) && dx && aligned16(dx), "bn_relu_bwd: null or misaligned volume");
    MVS_REQUIRE(scale && shift && mean && invstd && gamma && workspace && dbeta && dgamma, "bn_relu_bwd: null vector");
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == MVSB200_BF16 && g_dtype == MVSB200_BF16)
        return bn_bwd_impl<__nv_bfloat16, __nv_bfloat16>(x, gy, scale, shift, mean, invstd, gamma, workspace, dbeta, dgamma, dx, relu, M, C, st);
    if (x_dtype == MVSB200_BF16 && g_dtype == MVSB200_F32)
        return bn_bwd_impl<__nv_bfloat16, float>(x, gy, scale, shift, mean, invstd, gamma, workspace, dbeta, dgamma, dx, relu, M, C, st);
    if (x_dtype == MVSB200_F32 && g_dtype == MVSB200_F32)
        return bn_bwd_impl<float, float>(x, gy, scale, shift, mean, invstd, gamma, workspace, dbeta, dgamma, dx, relu, M, C, st);
    if (x_dtype == MVSB200_F32 && g_dtype == MVSB200_BF16)
        return bn_bwd_impl<float, __nv_bfloat16>(x, gy, scale, shift, mean, invstd, gamma, workspace, dbeta, dgamma, dx, relu, M, C, st);
    MVS_FAIL(MVSB200_E_BADARG, "bn_relu_bwd: bad dtypes %d / %d", x_dtype, g_dtype);
}

// ---- crop-aware entry points: geo = {Da, ha, wa, D, h, w, d0, h0, w0, dc, hc, wc} (host ints) --------------------------
static int make_box(const int* g, int64_t M, int64_t* Bout, CropBox* cb) {
    MVS_REQUIRE(g != nullptr, "bn crop: null geometry");
    cb->Da = g[0]; cb->ha = g[1]; cb->wa = g[2]; cb->D = g[3]; cb->h = g[4]; cb->w = g[5];
    cb->d0 = g[6]; cb->h0 = g[7]; cb->w0 = g[8]; cb->dc = g[9]; cb->hc = g[10]; cb->wc = g[11];
    const int64_t per = (int64_t)cb->D * cb->h * cb->w;
    MVS_REQUIRE(per > 0 && M % per == 0, "bn crop: canvas %dx%dx%d does not divide M", cb->D, cb->h, cb->w);
    MVS_REQUIRE(cb->Da >= cb->D && cb->ha >= cb->h && cb->wa >= cb->w, "bn crop: canvas larger than its allocation");
    MVS_REQUIRE((M / per) * (int64_t)cb->Da * cb->ha * cb->wa * 8 < (1LL << 31), "bn crop: volume too large for 32-bit chunk indices");
    MVS_REQUIRE(cb->d0 >= 0 && cb->h0 >= 0 && cb->w0 >= 0 && cb->dc >= 1 && cb->hc >= 1 && cb->wc >= 1 &&
                cb->d0 + cb->dc <= cb->D && cb->h0 + cb->hc <= cb->h && cb->w0 + cb->wc <= cb->w, "bn crop: box outside the canvas");
    *Bout = M / per;
    cb->f_wa = make_fastdiv((unsigned)cb->wa); cb->f_ha = make_fastdiv((unsigned)cb->ha); cb->f_Da = make_fastdiv((unsigned)cb->Da);
    cb->f_w = make_fastdiv((unsigned)cb->w); cb->f_h = make_fastdiv((unsigned)cb->h); cb->f_D = make_fastdiv((unsigned)cb->D);
    cb->f_wc = make_fastdiv((unsigned)cb->wc); cb->f_hc = make_fastdiv((unsigned)cb->hc); cb->f_dc = make_fastdiv((unsigned)cb->dc);
    cb->cshift = 0;                                   // set per call from C (set_cshift)
    return MVSB200_OK;
}

static void set_cshift(CropBox* cb, int C) { cb->cshift = C == 8 ? 0 : (C == 16 ? 1 : (C == 32 ? 2 : 3)); }

extern "C" int mvsb200_bn_stats_geo(const void* x, int dtype, int64_t M, int C, float* workspace, float* mean, float* var,
                                    const int* geo12, void* stream) {
    if (int rc = check_bn(x, M, C, "bn_stats_geo")) return rc;
    MVS_REQUIRE(workspace && mean && var, "bn_stats_geo: null output");
    MVS_REQUIRE(dtype == MVSB200_F32 || dtype == MVSB200_BF16, "bn_stats_geo: bad dtype %d", dtype);
    CropBox cb; int64_t B;
    if (int rc = make_box(geo12, M, &B, &cb)) return rc;
    set_cshift(&cb, C);
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_chunks = (long long)M * C / 8;
    const int grid = grid_for(n_chunks);
    if (dtype == MVSB200_BF16)
        bn_stats_geo_kernel<__nv_bfloat16><<<grid, kBnThreads, 0, st>>>((const __nv_bfloat16*)x, n_chunks, C, workspace, cb);
    else
        bn_stats_geo_kernel<float><<<grid, kBnThreads, 0, st>>>((const float*)x, n_chunks, C, workspace, cb);
    MVS_CHECK_LAUNCH("bn_stats_geo");
    bn_finalize_kernel<<<1, kFinSlices * 2 * kMaxC, 0, st>>>(workspace, grid, C, 1.0 / (double)M, 0, mean, var);
    MVS_CHECK_LAUNCH("bn_finalize");
    return MVSB200_OK;
}

extern "C" int mvsb200_bn_relu_fwd_crop(const void* x, int dtype, const float* scale, const float* shift, void* y, int relu,
                                        int64_t M, int C, const int* geo12, void* stream) {
    if (int rc = check_bn(x, M, C, "bn_relu_fwd_crop")) return rc;
    MVS_REQUIRE(y && aligned16(y) && scale && shift, "bn_relu_fwd_crop: null or misaligned argument");
    MVS_REQUIRE(dtype == MVSB200_F32 || dtype == MVSB200_BF16, "bn_relu_fwd_crop: bad dtype %d", dtype);
    CropBox cb; int64_t B;
    if (int rc = make_box(geo12, M, &B, &cb)) return rc;
    set_cshift(&cb, C);
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_box = (long long)B * cb.dc * cb.hc * cb.wc * C / 8;
    const int grid = grid_for(n_box);
    if (dtype == MVSB200_BF16)
        bn_relu_fwd_crop_kernel<__nv_bfloat16><<<grid, kBnThreads, 0, st>>>((const __nv_bfloat16*)x, scale, shift,
                                                                            (__nv_bfloat16*)y, n_box, C, relu, cb);
    else
        bn_relu_fwd_crop_kernel<float><<<grid, kBnThreads, 0, st>>>((const float*)x, scale, shift, (float*)y, n_box, C, relu, cb);
    MVS_CHECK_LAUNCH("bn_relu_fwd_crop");
    return MVSB200_OK;
}

/* y = round(max(x*scale + shift, 0)) + add: BatchNorm apply + ReLU + skip addition in one pass (SURVEY §8b `bn_relu_add_apply`;
 * scripts/model.py:117-123).  geo12 == NULL: x, add and y are [M, C] rows; otherwise y and add live on the crop box of x's canvas
 * (geometry as mvsb200_bn_relu_fwd_crop).  The normalised value is rounded to the storage type before the addition, so the result
 * equals a separate addition of the stored tensor bit for bit. */
extern "C" int mvsb200_bn_relu_add_apply(const void* x, int dtype, const float* scale, const float* shift, const void* add, void* y,
                                         int relu, int64_t M, int C, const int* geo12, void* stream) {
    if (int rc = check_bn(x, M, C, "bn_relu_add_apply")) return rc;
    MVS_REQUIRE(y && aligned16(y) && add && aligned16(add) && scale && shift, "bn_relu_add_apply: null or misaligned argument");
    MVS_REQUIRE(dtype == MVSB200_F32 || dtype == MVSB200_BF16, "bn_relu_add_apply: bad dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    if (!geo12) {
        const long long n_chunks = (long long)M * C / 8;
        const int grid = grid_for(n_chunks);
        if (dtype == MVSB200_BF16)
            bn_relu_fwd_kernel<__nv_bfloat16, true><<<grid, kBnThreads, 0, st>>>((const __nv_bfloat16*)x, scale, shift, (__nv_bfloat16*)y,
                                                                                 n_chunks, C, relu, (const __nv_bfloat16*)add);
        else
            bn_relu_fwd_kernel<float, true><<<grid, kBnThreads, 0, st>>>((const float*)x, scale, shift, (float*)y, n_chunks, C, relu,
                                                                         (const float*)add);
    } else {
        CropBox cb; int64_t B;
        if (int rc = make_box(geo12, M, &B, &cb)) return rc;
        set_cshift(&cb, C);
        const long long n_box = (long long)B * cb.dc * cb.hc * cb.wc * C / 8;
        const int grid = grid_for(n_box);
        if (dtype == MVSB200_BF16)
            bn_relu_fwd_crop_kernel<__nv_bfloat16, true><<<grid, kBnThreads, 0, st>>>((const __nv_bfloat16*)x, scale, shift,
                                                                                      (__nv_bfloat16*)y, n_box, C, relu, cb,
                                                                                      (const __nv_bfloat16*)add);
        else
            bn_relu_fwd_crop_kernel<float, true><<<grid, kBnThreads, 0, st>>>((const float*)x, scale, shift, (float*)y, n_box, C, relu,
                                                                              cb, (const float*)add);
    }
    MVS_CHECK_LAUNCH("bn_relu_add_apply");
    return MVSB200_OK;
}

template <typename TX, typename TG>
static int bn_bwd_crop_impl(const void* x, const void* gy, const float* scale, const float* shift, const float* mean,
                            const float* invstd, const float* gamma, float* workspace, float* dbeta, float* dgamma, void* dx,
                            int relu, int64_t M, int C, int64_t B, const CropBox& cb, cudaStream_t st) {
    const long long n_chunks = (long long)B * cb.Da * cb.ha * cb.wa * C / 8;       // dx covers the whole allocation
    const long long n_box = (long long)B * cb.dc * cb.hc * cb.wc * C / 8;
    const int grid_box = grid_for(n_box), grid = grid_for(n_chunks);
    const long long n_box_lines = (long long)B * cb.dc * cb.hc;
    bn_relu_bwd_reduce_crop_kernel<TX, TG><<<grid_box, kBnThreads, 0, st>>>((const TX*)x, (const TG*)gy, scale, shift, mean,
                                                                            invstd, n_box_lines, C, relu, workspace, cb);
    MVS_CHECK_LAUNCH("bn_relu_bwd_reduce_crop");
    bn_finalize_kernel<<<1, kFinSlices * 2 * kMaxC, 0, st>>>(workspace, grid_box, C, 1.0, 1, dbeta, dgamma);
    MVS_CHECK_LAUNCH("bn_finalize");
    const long long n_lines = (long long)B * cb.Da * cb.ha;
    MVS_REQUIRE(n_lines < (1LL << 31), "bn_relu_bwd_crop: too many lines");
    const long long want = (n_lines + kBnThreads / 32 - 1) / (kBnThreads / 32);
    const int grid_lines = (int)(want < kBnBlocks ? want : kBnBlocks);
    (void)grid;
    bn_relu_bwd_apply_crop_kernel<TX, TG><<<grid_lines, kBnThreads, 0, st>>>((const TX*)x, (const TG*)gy, scale, shift, mean, invstd,
                                                                             gamma, dbeta, dgamma, (TX*)dx, n_lines, C, relu,
                                                                             (float)(1.0 / (double)M), cb);
    MVS_CHECK_LAUNCH("bn_relu_bwd_apply_crop");
    return MVSB200_OK;
}

extern "C" int mvsb200_bn_relu_bwd_crop(const void* x, int x_dtype, const void* gy, int g_dtype, const float* scale,
                                        const float* shift, const float* mean, const float* invstd, const float* gamma,
                                        float* workspace, float* dbeta, float* dgamma, void* dx, int relu, int64_t M, int C,
                                        const int* geo12, void* stream) {
    if (int rc = check_bn(x, M, C, "bn_relu_bwd_crop")) return rc;
    MVS_REQUIRE(gy && aligned16(gy) && dx && aligned16(dx), "bn_relu_bwd_crop: null or misaligned volume");
    MVS_REQUIRE(scale && shift && mean && invstd && gamma && workspace && dbeta && dgamma, "bn_relu_bwd_crop: null vector");
    CropBox cb; int64_t B;
    if (int rc = make_box(geo12, M, &B, &cb)) return rc;
    set_cshift(&cb, C);
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == MVSB200_BF16 && g_dtype == MVSB200_BF16)
        return bn_bwd_crop_impl<__nv_bfloat16, __nv_bfloat16>(x, gy, scale, shift, mean, invstd, gamma, workspace, dbeta, dgamma, dx, relu, M, C, B, cb, st);
    if (x_dtype == MVSB200_BF16 && g_dtype == MVSB200_F32)
        return bn_bwd_crop_impl<__nv_bfloat16, float>(x, gy, scale, shift, mean, invstd, gamma, workspace, dbeta, dgamma, dx, relu, M, C, B, cb, st);
    if (x_dtype == MVSB200_F32 && g_dtype == MVSB200_F32)
        return bn_bwd_crop_impl<float, float>(x, gy, scale, shift, mean, invstd, gamma, workspace, dbeta, dgamma, dx, relu, M, C, B, cb, st);
    if (x_dtype == MVSB200_F32 && g_dtype == MVSB200_BF16)
        return bn_bwd_crop_impl<float, __nv_bfloat16>(x, gy, scale, shift, mean, invstd, gamma, workspace, dbeta, dgamma, dx, relu, M, C, B, cb, st);
    MVS_FAIL(MVSB200_E_BADARG, "bn_relu_bwd_crop: bad dtypes %d / %d", x_dtype, g_dtype);
}

/* bn_stats (+ geometry) followed by the whole per-channel algebra of a train-mode BatchNorm in ONE finalize launch: mean, biased
 * variance, invstd = 1/sqrt(var + eps), scale = gamma*invstd, shift = beta - mean*scale, and -- when running_mean is given --
 * the running statistics update of torch.nn.BatchNorm (momentum, unbiased variance M/(M-1), num_batches_tracked += 1). */
extern "C" int mvsb200_bn_stats_affine(const void* x, int dtype, int64_t M, int C, float* workspace, const int* geo12,
                                       const float* gamma, const float* beta, double eps, double momentum, float* running_mean,
                                       float* running_var, int64_t* num_batches_tracked, float* mean, float* var, float* invstd,
                                       float* scale, float* shift, void* stream) {
    if (int rc = check_bn(x, M, C, "bn_stats_affine")) return rc;
    MVS_REQUIRE(workspace && mean && var && invstd && scale && shift && gamma && beta, "bn_stats_affine: null vector");
    MVS_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_stats_affine: running_mean and running_var go together");
    MVS_REQUIRE(dtype == MVSB200_F32 || dtype == MVSB200_BF16, "bn_stats_affine: bad dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_chunks = (long long)M * C / 8;
    const int grid = grid_for(n_chunks);
    if (geo12) {
        CropBox cb; int64_t B;
        if (int rc = make_box(geo12, M, &B, &cb)) return rc;
        set_cshift(&cb, C);
        if (dtype == MVSB200_BF16)
            bn_stats_geo_kernel<__nv_bfloat16><<<grid, kBnThreads, 0, st>>>((const __nv_bfloat16*)x, n_chunks, C, workspace, cb);
        else
            bn_stats_geo_kernel<float><<<grid, kBnThreads, 0, st>>>((const float*)x, n_chunks, C, workspace, cb);
    } else {
        if (dtype == MVSB200_BF16)
            bn_stats_kernel<__nv_bfloat16><<<grid, kBnThreads, 0, st>>>((const __nv_bfloat16*)x, n_chunks, C, workspace);
        else
            bn_stats_kernel<float><<<grid, kBnThreads, 0, st>>>((const float*)x, n_chunks, C, workspace);
    }
    MVS_CHECK_LAUNCH("bn_stats");
    BnAffineOut o{mean, var, invstd, scale, shift, running_mean, running_var, reinterpret_cast<long long*>(num_batches_tracked)};
    bn_finalize_affine_kernel<<<1, kFinSlices * 2 * kMaxC, 0, st>>>(workspace, grid, C, 1.0 / (double)M,
                                                                     M > 1 ? (double)M / (double)(M - 1) : 1.0, gamma, beta, eps,
                                                                     momentum, o);
    MVS_CHECK_LAUNCH("bn_finalize_affine");
    return MVSB200_OK;
}

/* The finalize launch of mvsb200_bn_stats_affine on per-CTA partial sums [n_blocks][2][C] that a producer kernel already wrote
 * (the transposed convolution's epilogue, mvsb200_deconv3d_s2_fwd_stats): the statistics pass over the canvas disappears. */
extern "C" int mvsb200_bn_finalize_affine(const float* partials, int n_blocks, int64_t M, int C, const float* gamma, const float* beta,
                                          double eps, double momentum, float* running_mean, float* running_var,
                                          int64_t* num_batches_tracked, float* mean, float* var, float* invstd, float* scale,
                                          float* shift, void* stream) {
    MVS_REQUIRE(partials && mean && var && invstd && scale && shift && gamma && beta, "bn_finalize_affine: null vector");
    MVS_REQUIRE(n_blocks >= 1 && M >= 1 && (C == 8 || C == 16 || C == 32 || C == 64), "bn_finalize_affine: bad shape");
    MVS_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_finalize_affine: running_mean and running_var go together");
    BnAffineOut o{mean, var, invstd, scale, shift, running_mean, running_var, reinterpret_cast<long long*>(num_batches_tracked)};
    bn_finalize_affine_kernel<<<1, kFinSlices * 2 * kMaxC, 0, (cudaStream_t)stream>>>(partials, n_blocks, C, 1.0 / (double)M,
                                                                                     M > 1 ? (double)M / (double)(M - 1) : 1.0, gamma,
                                                                                     beta, eps, momentum, o);
    MVS_CHECK_LAUNCH("bn_finalize_affine");
    return MVSB200_OK;
}
