// K3d: per-channel sums and affine+ReLU over BOXES of channel-last volumes, with their gradients (sm_100a).
//
// Why: the reference's stride-2 branches (scripts/model.py:104-113, padding dim/2+1 from scripts/config.py:20) only carry
// information on a central box of the canvas; outside it their BatchNorm+ReLU output is one constant per channel.
// regulariser.py therefore evaluates those layers on boxes and folds the remainder into the batch statistics
// analytically.  What is left per branch is streaming work on box-shaped tensors:
//   channel_sums      s1[c] = sum x, s2[c] = sum x^2 over a (strided) box            -> statistics
//   affine_relu_geo   y(p) = max(xv(p)*scale[c] + shift[c], 0) for p in an OUTPUT box, with xv(p) = x(p) inside the INPUT
//                     box and 0 outside it (both boxes given in one coordinate frame).  Output box larger than the input
//                     box = "data on the box, BatchNorm'd zero around it" (the next convolution's operand); output box
//                     inside the input box = crop.
// Each is one pass; gradients: d/dx of the sums, and (dx, dscale, dshift) of the affine map.  x may be any view with
// unit channel stride (e.g. a channel slice of the stacked branch convolution); outputs are dense channel-last.
#include "common.cuh"

using namespace mvsb200;

namespace {

constexpr int kThreadsA = 256;
constexpr int kBlocksA = 148 * 4;
constexpr int kMaxCA = 64;

struct View {                 // a [B, C, D, h, w] view with unit channel stride; strides in elements
    long long sb, sd, sh, sw;
    int B, D, h, w;
    FastDiv fD, fh, fw;       // the extents as division constants (common.cuh)
};
struct Geo {
    View in;                  // x
    int io[3];                // origin of the input box in the common frame
    int oo[3];                // origin of the output box
    int od[3];                // size of the output box (D, h, w)
    FastDiv fod[3];
};

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                               pack_bf16x2(v[6], v[7]));
}

// Raw 8-channel chunks: the loads of kUnrollA independent chunks are issued back to back and unpacked afterwards (one 16-byte
// load in flight per thread keeps ~16 KB per SM in flight: ~3 TB/s at HBM latency, measured; the index decode of the next
// chunks overlaps the loads of the first).
constexpr int kUnrollA = 4;
template <typename T> struct Raw8;
template <> struct Raw8<float> { float4 a, b; };
template <> struct Raw8<__nv_bfloat16> { uint4 u; };
__device__ __forceinline__ Raw8<float> ldraw8(const float* p) {
    Raw8<float> r; r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4); return r;
}
__device__ __forceinline__ Raw8<__nv_bfloat16> ldraw8(const __nv_bfloat16* p) {
    Raw8<__nv_bfloat16> r; r.u = *reinterpret_cast<const uint4*>(p); return r;
}
__device__ __forceinline__ void zero8(Raw8<float>& r) { r.a = make_float4(0.f, 0.f, 0.f, 0.f); r.b = r.a; }
__device__ __forceinline__ void zero8(Raw8<__nv_bfloat16>& r) { r.u = make_uint4(0u, 0u, 0u, 0u); }
__device__ __forceinline__ void unpack8(const Raw8<float>& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
__device__ __forceinline__ void unpack8(const Raw8<__nv_bfloat16>& r, float (&v)[8]) {
    const uint32_t w[4] = {r.u.x, r.u.y, r.u.z, r.u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}

// dense row index -> coordinates (32-bit arithmetic: rows < 2^31 is checked on the host)
__device__ __forceinline__ void decode_row(unsigned r, const FastDiv& fD, const FastDiv& fh, const FastDiv& fw, int& b, int& d,
                                           int& y, int& x) {
    unsigned ux, uy, ud;
    r = fd_divmod(r, fw, ux);
    r = fd_divmod(r, fh, uy);
    b = (int)fd_divmod(r, fD, ud);
    x = (int)ux; y = (int)uy; d = (int)ud;
}

__device__ __forceinline__ void reduce_to_partial(float (&a)[8], float (&b)[8], int cpr, int C, float* partial_row) {
    __shared__ float s_part[kThreadsA / 32][2][kMaxCA];
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
        for (int i = 0; i < 8; ++i) { s_part[warp][0][lane * 8 + i] = a[i]; s_part[warp][1][lane * 8 + i] = b[i]; }
    }
    __syncthreads();
    if ((int)threadIdx.x < 2 * C) {
        const int which = threadIdx.x / C, c = threadIdx.x % C;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kThreadsA / 32; ++w) s += s_part[w][which][c];
        partial_row[which * C + c] = s;
    }
}

// one CTA of 8 x 2C threads; slice s adds the partial rows k = s, s+8, ... in fp64, slices combined in fixed order
constexpr int kFinSlicesA = 8;
__global__ void __launch_bounds__(kFinSlicesA * 2 * kMaxCA) finalize_sums_kernel(const float* __restrict__ partials, int nblocks, int C,
                                                                                float* __restrict__ o0, float* __restrict__ o1) {
    __shared__ double s_acc[kFinSlicesA][2 * kMaxCA];
    const int col = threadIdx.x % (2 * kMaxCA), slice = threadIdx.x / (2 * kMaxCA);
    if (col < 2 * C) {
        double a = 0.0;
        for (int k = slice; k < nblocks; k += kFinSlicesA) a += (double)partials[(size_t)k * 2 * C + col];
        s_acc[slice][col] = a;
    }
    __syncthreads();
    const int c = threadIdx.x;
    if (c >= C) return;
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int s = 0; s < kFinSlicesA; ++s) { a += s_acc[s][c]; b += s_acc[s][C + c]; }
    o0[c] = (float)a;
    o1[c] = (float)b;
}

template <typename T>
__global__ void __launch_bounds__(kThreadsA) channel_sums_kernel(const T* __restrict__ x, View v, unsigned n_chunks, int C,
                                                                 float* __restrict__ partials) {
    const int cpr = C / 8;
    const unsigned i0 = blockIdx.x * kThreadsA + threadIdx.x;
    const int cshift = cpr == 1 ? 0 : (cpr == 2 ? 1 : (cpr == 4 ? 2 : 3));
    const int cg = (int)(i0 & (unsigned)(cpr - 1));
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned stride = gridDim.x * kThreadsA;
    for (unsigned i = i0; i < n_chunks; i += kUnrollA * stride) {
        Raw8<T> raw[kUnrollA];
#pragma unroll
        for (int u = 0; u < kUnrollA; ++u) {
            const unsigned j = i + u * stride;             // (j < i: wrapped past 2^32, beyond the end)
            if (j < n_chunks && j >= i) {
                int b, d, y, xx;
                decode_row(j >> cshift, v.fD, v.fh, v.fw, b, d, y, xx);
                raw[u] = ldraw8(x + b * v.sb + d * v.sd + y * v.sh + xx * v.sw + cg * 8);
            } else {
                zero8(raw[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnrollA; ++u) {
            float val[8];
            unpack8(raw[u], val);
#pragma unroll
            for (int k = 0; k < 8; ++k) { s[k] += val[k]; q[k] = fmaf(val[k], val[k], q[k]); }
        }
    }
    reduce_to_partial(s, q, cpr, C, partials + (size_t)blockIdx.x * 2 * C);
}

template <typename T>
__global__ void __launch_bounds__(kThreadsA) channel_sums_bwd_kernel(const T* __restrict__ x, View v, unsigned n_chunks, int C,
                                                                     const float* __restrict__ g1, const float* __restrict__ g2,
                                                                     T* __restrict__ gx) {
    const int cpr = C / 8;
    const unsigned i0 = blockIdx.x * kThreadsA + threadIdx.x;
    const int cshift = cpr == 1 ? 0 : (cpr == 2 ? 1 : (cpr == 4 ? 2 : 3));
    const int cg = (int)(i0 & (unsigned)(cpr - 1));
    float a[8], b2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = g1[cg * 8 + k]; b2[k] = 2.f * g2[cg * 8 + k]; }
    for (unsigned i = i0; i < n_chunks; i += gridDim.x * kThreadsA) {
        int b, d, y, xx;
        decode_row(i >> cshift, v.fD, v.fh, v.fw, b, d, y, xx);
        float val[8];
        load8(x + b * v.sb + d * v.sd + y * v.sh + xx * v.sw + cg * 8, val);
#pragma unroll
        for (int k = 0; k < 8; ++k) val[k] = fmaf(val[k], b2[k], a[k]);
        store8(gx + (size_t)i * 8, val);
    }
}

// value of x at output row (b, od, oy, ox) of the output box, 0 outside the input box
template <typename T>
__device__ __forceinline__ bool load_in(const T* __restrict__ x, const Geo& g, int b, int od, int oy, int ox, int cg, float (&val)[8]) {
    const int id = od + g.oo[0] - g.io[0], iy = oy + g.oo[1] - g.io[1], ix = ox + g.oo[2] - g.io[2];
    if ((unsigned)id < (unsigned)g.in.D && (unsigned)iy < (unsigned)g.in.h && (unsigned)ix < (unsigned)g.in.w) {
        load8(x + b * g.in.sb + id * g.in.sd + iy * g.in.sh + ix * g.in.sw + cg * 8, val);
        return true;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) val[k] = 0.f;
    return false;
}

template <typename T>
__device__ __forceinline__ bool ldraw_in(const T* __restrict__ x, const Geo& g, int b, int od, int oy, int ox, int cg, Raw8<T>& r) {
    const int id = od + g.oo[0] - g.io[0], iy = oy + g.oo[1] - g.io[1], ix = ox + g.oo[2] - g.io[2];
    if ((unsigned)id < (unsigned)g.in.D && (unsigned)iy < (unsigned)g.in.h && (unsigned)ix < (unsigned)g.in.w) {
        r = ldraw8(x + b * g.in.sb + id * g.in.sd + iy * g.in.sh + ix * g.in.sw + cg * 8);
        return true;
    }
    zero8(r);
    return false;
}

template <typename T>
__global__ void __launch_bounds__(kThreadsA) affine_relu_geo_fwd_kernel(const T* __restrict__ x, Geo g, unsigned n_out_chunks, int C,
                                                                        const float* __restrict__ scale,
                                                                        const float* __restrict__ shift, T* __restrict__ y, int relu) {
    const int cpr = C / 8;
    const unsigned i0 = blockIdx.x * kThreadsA + threadIdx.x;
    const int cshift = cpr == 1 ? 0 : (cpr == 2 ? 1 : (cpr == 4 ? 2 : 3));
    const int cg = (int)(i0 & (unsigned)(cpr - 1));
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = scale[cg * 8 + k]; sh[k] = shift[cg * 8 + k]; }
    const unsigned stride = gridDim.x * kThreadsA;
    for (unsigned i = i0; i < n_out_chunks; i += kUnrollA * stride) {
        Raw8<T> raw[kUnrollA];
#pragma unroll
        for (int u = 0; u < kUnrollA; ++u) {
            const unsigned j = i + u * stride;
            if (j < n_out_chunks && j >= i) {
                int b, d, yy, xx;
                decode_row(j >> cshift, g.fod[0], g.fod[1], g.fod[2], b, d, yy, xx);
                ldraw_in(x, g, b, d, yy, xx, cg, raw[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnrollA; ++u) {
            const unsigned j = i + u * stride;
            if (j < n_out_chunks && j >= i) {
                float val[8];
                unpack8(raw[u], val);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    val[k] = fmaf(val[k], sc[k], sh[k]);
                    if (relu) val[k] = fmaxf(val[k], 0.f);
                }
                store8(y + (size_t)j * 8, val);
            }
        }
    }
}

// dshift[c] = sum_O g, dscale[c] = sum_O g * xv   with g = gy * [xv*scale+shift > 0]
template <typename T, typename TG>
__global__ void __launch_bounds__(kThreadsA) affine_relu_geo_bwd_reduce_kernel(const T* __restrict__ x, const TG* __restrict__ gy,
                                                                               Geo g, unsigned n_out_chunks, int C,
                                                                               const float* __restrict__ scale,
                                                                               const float* __restrict__ shift, int relu,
                                                                               float* __restrict__ partials) {
    const int cpr = C / 8;
    const unsigned i0 = blockIdx.x * kThreadsA + threadIdx.x;
    const int cshift = cpr == 1 ? 0 : (cpr == 2 ? 1 : (cpr == 4 ? 2 : 3));
    const int cg = (int)(i0 & (unsigned)(cpr - 1));
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = scale[cg * 8 + k]; sh[k] = shift[cg * 8 + k]; }
    float sg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sgx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned stride = gridDim.x * kThreadsA;
    for (unsigned i = i0; i < n_out_chunks; i += kUnrollA * stride) {
        Raw8<T> rx[kUnrollA];
        Raw8<TG> rg[kUnrollA];
#pragma unroll
        for (int u = 0; u < kUnrollA; ++u) {
            const unsigned j = i + u * stride;
            if (j < n_out_chunks && j >= i) {
                int b, d, yy, xx;
                decode_row(j >> cshift, g.fod[0], g.fod[1], g.fod[2], b, d, yy, xx);
                ldraw_in(x, g, b, d, yy, xx, cg, rx[u]);
                rg[u] = ldraw8(gy + (size_t)j * 8);
            } else {
                zero8(rx[u]);
                zero8(rg[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnrollA; ++u) {
            float val[8], gv[8];
            unpack8(rx[u], val);
            unpack8(rg[u], gv);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float gk = (!relu || fmaf(val[k], sc[k], sh[k]) > 0.f) ? gv[k] : 0.f;
                sg[k] += gk;
                sgx[k] = fmaf(gk, val[k], sgx[k]);
            }
        }
    }
    reduce_to_partial(sg, sgx, cpr, C, partials + (size_t)blockIdx.x * 2 * C);
}

// gx over the input box (dense): g * scale where the voxel is inside the output box, else 0
template <typename T, typename TG>
__global__ void __launch_bounds__(kThreadsA) affine_relu_geo_bwd_dx_kernel(const T* __restrict__ x, const TG* __restrict__ gy, Geo g,
                                                                           unsigned n_in_chunks, int C,
                                                                           const float* __restrict__ scale,
                                                                           const float* __restrict__ shift, int relu,
                                                                           T* __restrict__ gx) {
    const int cpr = C / 8;
    const unsigned i0 = blockIdx.x * kThreadsA + threadIdx.x;
    const int cshift = cpr == 1 ? 0 : (cpr == 2 ? 1 : (cpr == 4 ? 2 : 3));
    const int cg = (int)(i0 & (unsigned)(cpr - 1));
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = scale[cg * 8 + k]; sh[k] = shift[cg * 8 + k]; }
    for (unsigned i = i0; i < n_in_chunks; i += gridDim.x * kThreadsA) {
        int b, d, yy, xx;
        decode_row(i >> cshift, g.in.fD, g.in.fh, g.in.fw, b, d, yy, xx);
        const int od = d + g.io[0] - g.oo[0], oy = yy + g.io[1] - g.oo[1], ox = xx + g.io[2] - g.oo[2];
        float out[8];
        if ((unsigned)od < (unsigned)g.od[0] && (unsigned)oy < (unsigned)g.od[1] && (unsigned)ox < (unsigned)g.od[2]) {
            float val[8], gv[8];
            load8(x + b * g.in.sb + d * g.in.sd + yy * g.in.sh + xx * g.in.sw + cg * 8, val);
            const size_t orow = (((size_t)b * g.od[0] + od) * g.od[1] + oy) * g.od[2] + ox;
            load8(gy + (orow * cpr + cg) * 8, gv);
#pragma unroll
            for (int k = 0; k < 8; ++k) out[k] = (!relu || fmaf(val[k], sc[k], sh[k]) > 0.f) ? gv[k] * sc[k] : 0.f;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) out[k] = 0.f;
        }
        store8(gx + (size_t)i * 8, out);
    }
}

// Backward APPLY of train-mode BatchNorm + ReLU on a box (the stride-2 branches): with the per-channel vectors of the
// statistics' backward (a = dL/d(sum x), b2 = 2 dL/d(sum x^2)),
//   gx = gy * [xv*scale+shift > 0] * scale + a + b2 * xv      over the input box (gy = 0 outside the output box),
// one pass instead of "affine dx" + "channel sums backward" + an addition -- and written through arbitrary row strides, so that
// the result can land directly inside the padded channel-stacked buffer the strided convolution's backward reads.
template <typename T, typename TG>
__global__ void __launch_bounds__(kThreadsA) box_bn_relu_bwd_apply_kernel(const T* __restrict__ x, const TG* __restrict__ gy, Geo g,
                                                                          unsigned n_in_chunks, int C,
                                                                          const float* __restrict__ scale,
                                                                          const float* __restrict__ shift,
                                                                          const float* __restrict__ a, const float* __restrict__ b2,
                                                                          int relu, T* __restrict__ gx, long long osb,
                                                                          long long osd, long long osh, long long osw) {
    const int cpr = C / 8;
    const unsigned i0 = blockIdx.x * kThreadsA + threadIdx.x;
    const int cshift = cpr == 1 ? 0 : (cpr == 2 ? 1 : (cpr == 4 ? 2 : 3));
    const int cg = (int)(i0 & (unsigned)(cpr - 1));
    float sc[8], sh[8], av[8], bv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = scale[cg * 8 + k]; sh[k] = shift[cg * 8 + k]; av[k] = a[cg * 8 + k]; bv[k] = b2[cg * 8 + k]; }
    const unsigned stride = gridDim.x * kThreadsA;
    for (unsigned i = i0; i < n_in_chunks; i += kUnrollA * stride) {
        Raw8<T> rx[kUnrollA];
        Raw8<TG> rg[kUnrollA];
        long long dst[kUnrollA];                         // element offset of the chunk in gx, < 0: beyond the end
        bool has_g[kUnrollA];
#pragma unroll
        for (int u = 0; u < kUnrollA; ++u) {
            const unsigned j = i + u * stride;
            dst[u] = -1;
            has_g[u] = false;
            if (j < n_in_chunks && j >= i) {
                int b, d, yy, xx;
                decode_row(j >> cshift, g.in.fD, g.in.fh, g.in.fw, b, d, yy, xx);
                const int od = d + g.io[0] - g.oo[0], oy = yy + g.io[1] - g.oo[1], ox = xx + g.io[2] - g.oo[2];
                rx[u] = ldraw8(x + b * g.in.sb + d * g.in.sd + yy * g.in.sh + xx * g.in.sw + cg * 8);
                dst[u] = b * osb + d * osd + yy * osh + xx * osw + cg * 8;
                if ((unsigned)od < (unsigned)g.od[0] && (unsigned)oy < (unsigned)g.od[1] && (unsigned)ox < (unsigned)g.od[2]) {
                    const size_t orow = (((size_t)b * g.od[0] + od) * g.od[1] + oy) * g.od[2] + ox;
                    rg[u] = ldraw8(gy + (orow * cpr + cg) * 8);
                    has_g[u] = true;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnrollA; ++u) {
            if (dst[u] < 0) continue;
            float val[8], out[8];
            unpack8(rx[u], val);
#pragma unroll
            for (int k = 0; k < 8; ++k) out[k] = fmaf(bv[k], val[k], av[k]);
            if (has_g[u]) {
                float gv[8];
                unpack8(rg[u], gv);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (!relu || fmaf(val[k], sc[k], sh[k]) > 0.f) out[k] = fmaf(gv[k], sc[k], out[k]);
            }
            store8(gx + dst[u], out);
        }
    }
}

int make_view(const int64_t* strides4, const int* dims4, int C, View* v, const char* name) {
    MVS_REQUIRE(strides4 && dims4, "%s: null geometry", name);
    MVS_REQUIRE(C == 8 || C == 16 || C == 32 || C == 64, "%s: C must be 8, 16, 32 or 64 (got %d)", name, C);
    v->sb = strides4[0]; v->sd = strides4[1]; v->sh = strides4[2]; v->sw = strides4[3];
    v->B = dims4[0]; v->D = dims4[1]; v->h = dims4[2]; v->w = dims4[3];
    MVS_REQUIRE(v->B >= 1 && v->D >= 1 && v->h >= 1 && v->w >= 1, "%s: empty box", name);
    MVS_REQUIRE(v->sb % 8 == 0 && v->sd % 8 == 0 && v->sh % 8 == 0 && v->sw % 8 == 0, "%s: strides must be multiples of 8 elements", name);
    MVS_REQUIRE((long long)v->B * v->D * v->h * v->w * (C / 8) < (1LL << 31), "%s: box too large", name);
    v->fD = make_fastdiv((unsigned)v->D); v->fh = make_fastdiv((unsigned)v->h); v->fw = make_fastdiv((unsigned)v->w);
    return MVSB200_OK;
}

int grid_of(unsigned n_chunks) {
    const unsigned b = (n_chunks + kThreadsA - 1) / kThreadsA;
    return (int)(b < (unsigned)kBlocksA ? (b ? b : 1) : kBlocksA);
}

int make_geo(const int64_t* strides4, const int* geo13, int C, Geo* g, const char* name) {
    MVS_REQUIRE(geo13 != nullptr, "%s: null geometry", name);
    if (int rc = make_view(strides4, geo13, C, &g->in, name)) return rc;
    for (int i = 0; i < 3; ++i) { g->io[i] = geo13[4 + i]; g->oo[i] = geo13[7 + i]; g->od[i] = geo13[10 + i]; }
    MVS_REQUIRE(g->od[0] >= 1 && g->od[1] >= 1 && g->od[2] >= 1, "%s: empty output box", name);
    MVS_REQUIRE((long long)g->in.B * g->od[0] * g->od[1] * g->od[2] * (C / 8) < (1LL << 31), "%s: output box too large", name);
    for (int i = 0; i < 3; ++i) g->fod[i] = make_fastdiv((unsigned)g->od[i]);
    return MVSB200_OK;
}

}  // namespace

extern "C" int64_t mvsb200_affine_workspace_floats(void) { return (int64_t)kBlocksA * 2 * kMaxCA; }

extern "C" int mvsb200_channel_sums(const void* x, int dtype, const int64_t* strides4, const int* dims4, int C, float* workspace,
                                    float* s1, float* s2, void* stream) {
    MVS_REQUIRE(x && aligned16(x) && workspace && s1 && s2, "channel_sums: null or misaligned pointer");
    View v;
    if (int rc = make_view(strides4, dims4, C, &v, "channel_sums")) return rc;
    const unsigned n = (unsigned)((long long)v.B * v.D * v.h * v.w * (C / 8));
    const int grid = grid_of(n);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MVSB200_BF16) channel_sums_kernel<__nv_bfloat16><<<grid, kThreadsA, 0, st>>>((const __nv_bfloat16*)x, v, n, C, workspace);
    else if (dtype == MVSB200_F32) channel_sums_kernel<float><<<grid, kThreadsA, 0, st>>>((const float*)x, v, n, C, workspace);
    else MVS_FAIL(MVSB200_E_BADARG, "channel_sums: bad dtype %d", dtype);
    MVS_CHECK_LAUNCH("channel_sums");
    finalize_sums_kernel<<<1, kFinSlicesA * 2 * kMaxCA, 0, st>>>(workspace, grid, C, s1, s2);
    MVS_CHECK_LAUNCH("channel_sums_finalize");
    return MVSB200_OK;
}

extern "C" int mvsb200_channel_sums_bwd(const void* x, int dtype, const int64_t* strides4, const int* dims4, int C, const float* g1,
                                        const float* g2, void* gx, void* stream) {
    MVS_REQUIRE(x && aligned16(x) && g1 && g2 && gx && aligned16(gx), "channel_sums_bwd: null or misaligned pointer");
    View v;
    if (int rc = make_view(strides4, dims4, C, &v, "channel_sums_bwd")) return rc;
    const unsigned n = (unsigned)((long long)v.B * v.D * v.h * v.w * (C / 8));
    const int grid = grid_of(n);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MVSB200_BF16) channel_sums_bwd_kernel<__nv_bfloat16><<<grid, kThreadsA, 0, st>>>((const __nv_bfloat16*)x, v, n, C, g1, g2, (__nv_bfloat16*)gx);
    else if (dtype == MVSB200_F32) channel_sums_bwd_kernel<float><<<grid, kThreadsA, 0, st>>>((const float*)x, v, n, C, g1, g2, (float*)gx);
    else MVS_FAIL(MVSB200_E_BADARG, "channel_sums_bwd: bad dtype %d", dtype);
    MVS_CHECK_LAUNCH("channel_sums_bwd");
    return MVSB200_OK;
}

extern "C" int mvsb200_affine_relu_geo_fwd(const void* x, int dtype, const int64_t* strides4, const int* geo13, int C,
                                           const float* scale, const float* shift, void* y, int relu, void* stream) {
    MVS_REQUIRE(x && aligned16(x) && y && aligned16(y) && scale && shift, "affine_relu_geo_fwd: null or misaligned pointer");
    Geo g;
    if (int rc = make_geo(strides4, geo13, C, &g, "affine_relu_geo_fwd")) return rc;
    const unsigned n = (unsigned)((long long)g.in.B * g.od[0] * g.od[1] * g.od[2] * (C / 8));
    const int grid = grid_of(n);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MVSB200_BF16)
        affine_relu_geo_fwd_kernel<__nv_bfloat16><<<grid, kThreadsA, 0, st>>>((const __nv_bfloat16*)x, g, n, C, scale, shift, (__nv_bfloat16*)y, relu);
    else if (dtype == MVSB200_F32)
        affine_relu_geo_fwd_kernel<float><<<grid, kThreadsA, 0, st>>>((const float*)x, g, n, C, scale, shift, (float*)y, relu);
    else MVS_FAIL(MVSB200_E_BADARG, "affine_relu_geo_fwd: bad dtype %d", dtype);
    MVS_CHECK_LAUNCH("affine_relu_geo_fwd");
    return MVSB200_OK;
}

template <typename T, typename TG>
static int affine_bwd_impl(const void* x, const void* gy, const Geo& g, int C, const float* scale, const float* shift, int relu,
                           float* workspace, float* gscale, float* gshift, void* gx, cudaStream_t st) {
    const unsigned n_out = (unsigned)((long long)g.in.B * g.od[0] * g.od[1] * g.od[2] * (C / 8));
    const unsigned n_in = (unsigned)((long long)g.in.B * g.in.D * g.in.h * g.in.w * (C / 8));
    const int grid_out = grid_of(n_out), grid_in = grid_of(n_in);
    affine_relu_geo_bwd_reduce_kernel<T, TG><<<grid_out, kThreadsA, 0, st>>>((const T*)x, (const TG*)gy, g, n_out, C, scale, shift, relu, workspace);
    MVS_CHECK_LAUNCH("affine_relu_geo_bwd_reduce");
    finalize_sums_kernel<<<1, kFinSlicesA * 2 * kMaxCA, 0, st>>>(workspace, grid_out, C, gshift, gscale);
    MVS_CHECK_LAUNCH("affine_relu_geo_bwd_finalize");
    affine_relu_geo_bwd_dx_kernel<T, TG><<<grid_in, kThreadsA, 0, st>>>((const T*)x, (const TG*)gy, g, n_in, C, scale, shift, relu, (T*)gx);
    MVS_CHECK_LAUNCH("affine_relu_geo_bwd_dx");
    return MVSB200_OK;
}

extern "C" int mvsb200_affine_relu_geo_bwd(const void* x, int x_dtype, const int64_t* strides4, const int* geo13, int C,
                                           const float* scale, const float* shift, const void* gy, int g_dtype, float* workspace,
                                           float* gscale, float* gshift, void* gx, int relu, void* stream) {
    MVS_REQUIRE(x && aligned16(x) && gy && aligned16(gy) && gx && aligned16(gx), "affine_relu_geo_bwd: null or misaligned volume");
    MVS_REQUIRE(scale && shift && workspace && gscale && gshift, "affine_relu_geo_bwd: null vector");
    Geo g;
    if (int rc = make_geo(strides4, geo13, C, &g, "affine_relu_geo_bwd")) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == MVSB200_BF16 && g_dtype == MVSB200_BF16) return affine_bwd_impl<__nv_bfloat16, __nv_bfloat16>(x, gy, g, C, scale, shift, relu, workspace, gscale, gshift, gx, st);
    if (x_dtype == MVSB200_BF16 && g_dtype == MVSB200_F32) return affine_bwd_impl<__nv_bfloat16, float>(x, gy, g, C, scale, shift, relu, workspace, gscale, gshift, gx, st);
    if (x_dtype == MVSB200_F32 && g_dtype == MVSB200_F32) return affine_bwd_impl<float, float>(x, gy, g, C, scale, shift, relu, workspace, gscale, gshift, gx, st);
    if (x_dtype == MVSB200_F32 && g_dtype == MVSB200_BF16) return affine_bwd_impl<float, __nv_bfloat16>(x, gy, g, C, scale, shift, relu, workspace, gscale, gshift, gx, st);
    MVS_FAIL(MVSB200_E_BADARG, "affine_relu_geo_bwd: bad dtypes %d / %d", x_dtype, g_dtype);
}

template <typename T, typename TG>
static int box_bn_reduce_impl(const void* x, const void* gy, const Geo& g, int C, const float* scale, const float* shift, int relu,
                              float* workspace, float* gscale, float* gshift, cudaStream_t st) {
    const unsigned n_out = (unsigned)((long long)g.in.B * g.od[0] * g.od[1] * g.od[2] * (C / 8));
    const int grid_out = grid_of(n_out);
    affine_relu_geo_bwd_reduce_kernel<T, TG><<<grid_out, kThreadsA, 0, st>>>((const T*)x, (const TG*)gy, g, n_out, C, scale, shift, relu, workspace);
    MVS_CHECK_LAUNCH("affine_relu_geo_bwd_reduce");
    finalize_sums_kernel<<<1, kFinSlicesA * 2 * kMaxCA, 0, st>>>(workspace, grid_out, C, gshift, gscale);
    MVS_CHECK_LAUNCH("affine_relu_geo_bwd_finalize");
    return MVSB200_OK;
}

/* First half of the box BatchNorm+ReLU backward: gshift[c] = sum g, gscale[c] = sum g * xv with g = gy * [xv*scale+shift > 0]
 * over the output box (deterministic two-stage reduction). */
extern "C" int mvsb200_box_bn_relu_bwd_reduce(const void* x, int x_dtype, const int64_t* strides4, const int* geo13, int C,
                                              const float* scale, const float* shift, const void* gy, int g_dtype, float* workspace,
                                              float* gscale, float* gshift, int relu, void* stream) {
    MVS_REQUIRE(x && aligned16(x) && gy && aligned16(gy), "box_bn_relu_bwd_reduce: null or misaligned volume");
    MVS_REQUIRE(scale && shift && workspace && gscale && gshift, "box_bn_relu_bwd_reduce: null vector");
    Geo g;
    if (int rc = make_geo(strides4, geo13, C, &g, "box_bn_relu_bwd_reduce")) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == MVSB200_BF16 && g_dtype == MVSB200_BF16) return box_bn_reduce_impl<__nv_bfloat16, __nv_bfloat16>(x, gy, g, C, scale, shift, relu, workspace, gscale, gshift, st);
    if (x_dtype == MVSB200_BF16 && g_dtype == MVSB200_F32) return box_bn_reduce_impl<__nv_bfloat16, float>(x, gy, g, C, scale, shift, relu, workspace, gscale, gshift, st);
    if (x_dtype == MVSB200_F32 && g_dtype == MVSB200_F32) return box_bn_reduce_impl<float, float>(x, gy, g, C, scale, shift, relu, workspace, gscale, gshift, st);
    if (x_dtype == MVSB200_F32 && g_dtype == MVSB200_BF16) return box_bn_reduce_impl<float, __nv_bfloat16>(x, gy, g, C, scale, shift, relu, workspace, gscale, gshift, st);
    MVS_FAIL(MVSB200_E_BADARG, "box_bn_relu_bwd_reduce: bad dtypes %d / %d", x_dtype, g_dtype);
}

/* Second half: gx = gy * [xv*scale+shift > 0] * scale + a + b2 * xv over the input box, written through out_strides4 (elements;
 * gx has x's dtype and may be a view into a larger, channel-stacked buffer). */
extern "C" int mvsb200_box_bn_relu_bwd_apply(const void* x, int x_dtype, const int64_t* strides4, const int* geo13, int C,
                                             const float* scale, const float* shift, const float* a, const float* b2, const void* gy,
                                             int g_dtype, void* gx, const int64_t* out_strides4, int relu, void* stream) {
    MVS_REQUIRE(x && aligned16(x) && gy && aligned16(gy) && gx && aligned16(gx), "box_bn_relu_bwd_apply: null or misaligned volume");
    MVS_REQUIRE(scale && shift && a && b2 && out_strides4, "box_bn_relu_bwd_apply: null vector");
    for (int i = 0; i < 4; ++i) MVS_REQUIRE(out_strides4[i] % 8 == 0, "box_bn_relu_bwd_apply: output strides must be multiples of 8 elements");
    Geo g;
    if (int rc = make_geo(strides4, geo13, C, &g, "box_bn_relu_bwd_apply")) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned n_in = (unsigned)((long long)g.in.B * g.in.D * g.in.h * g.in.w * (C / 8));
    const int grid_in = grid_of(n_in);
    const long long o0 = out_strides4[0], o1 = out_strides4[1], o2 = out_strides4[2], o3 = out_strides4[3];
#define MVS_APPLY(T, TG)                                                                                                        \
    box_bn_relu_bwd_apply_kernel<T, TG><<<grid_in, kThreadsA, 0, st>>>((const T*)x, (const TG*)gy, g, n_in, C, scale, shift, a, b2, relu, \
                                                                       (T*)gx, o0, o1, o2, o3)
    if (x_dtype == MVSB200_BF16 && g_dtype == MVSB200_BF16) MVS_APPLY(__nv_bfloat16, __nv_bfloat16);
    else if (x_dtype == MVSB200_BF16 && g_dtype == MVSB200_F32) MVS_APPLY(__nv_bfloat16, float);
    else if (x_dtype == MVSB200_F32 && g_dtype == MVSB200_F32) MVS_APPLY(float, float);
    else if (x_dtype == MVSB200_F32 && g_dtype == MVSB200_BF16) MVS_APPLY(float, __nv_bfloat16);
    else MVS_FAIL(MVSB200_E_BADARG, "box_bn_relu_bwd_apply: bad dtypes %d / %d", x_dtype, g_dtype);
#undef MVS_APPLY
    MVS_CHECK_LAUNCH("box_bn_relu_bwd_apply");
    return MVSB200_OK;
}

// ---- per-channel algebra of the box BatchNorm (the stride-2 branches) in one launch each ---------------------------------
// forward:  mean = s1/n, var = s2/n - mean^2 (the canvas outside the box is zero and counts), r = 1/sqrt(var + eps),
//           scale = gamma r, shift = beta - mean scale (fp64 inside), running statistics as torch.nn.BatchNorm updates them.
// backward: from G_scale = sum over the box + external, G_shift likewise:
//           g_beta = G_shift, G = G_scale - mean G_shift, g_gamma = G r, g_var = -G gamma r^3 / 2,
//           g_mean = -gamma r G_shift - 2 mean g_var;  a = g_mean / n = dL/d(sum S),  b2 = 2 g_var / n = 2 dL/d(sum S^2).
// These were ~12 + ~20 [C]-sized torch launches per branch and step.
namespace {
__global__ void box_bn_algebra_fwd_kernel(const float* __restrict__ s1, const float* __restrict__ s2,
                                          const double* __restrict__ add1, const double* __restrict__ add2, int C, double n,
                                          const float* __restrict__ gamma, const float* __restrict__ beta, double eps,
                                          double momentum, float* running_mean, float* running_var, long long* nbt,
                                          float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean,
                                          float* __restrict__ var, double* __restrict__ stat64) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt) *nbt += 1;
    if (c >= C) return;
    // add1 / add2: the sums of the values (and their squares) the tensor takes OUTSIDE the box it was summed over (conv_k_1: 27
    // closed-form border classes of a convolution of a constant, regulariser.py)
    const double m = ((double)s1[c] + (add1 ? add1[c] : 0.0)) / n;
    double v = ((double)s2[c] + (add2 ? add2[c] : 0.0)) / n - m * m;
    v = v > 0.0 ? v : 0.0;
    const double r = 1.0 / sqrt(v + eps);
    const double g = (double)gamma[c];
    scale[c] = (float)(g * r);
    shift[c] = (float)((double)beta[c] - m * g * r);
    mean[c] = (float)m;
    var[c] = (float)v;
    stat64[c] = m;
    stat64[C + c] = r;
    if (running_mean) {
        running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * (double)(float)m);
        running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * (double)(float)v * (n > 1.0 ? n / (n - 1.0) : 1.0));
    }
}
__global__ void box_bn_algebra_bwd_kernel(const float* __restrict__ gscale, const float* __restrict__ gshift,
                                          const float* __restrict__ gscale_ext, const float* __restrict__ gshift_ext,
                                          const double* __restrict__ stat64, const float* __restrict__ gamma, int C, double n,
                                          float* __restrict__ a, float* __restrict__ b2, float* __restrict__ g_gamma,
                                          float* __restrict__ g_beta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = stat64[c], r = stat64[C + c], g = (double)gamma[c];
    const double Gsc = (double)gscale[c] + (gscale_ext ? (double)gscale_ext[c] : 0.0);
    const double Gsh = (double)gshift[c] + (gshift_ext ? (double)gshift_ext[c] : 0.0);
    const double G = Gsc - m * Gsh;
    const double g_var = G * g * (-0.5) * r * r * r;
    const double g_mean = -(g * r) * Gsh - 2.0 * m * g_var;
    a[c] = (float)(g_mean / n);
    b2[c] = (float)(2.0 * g_var / n);
    g_gamma[c] = (float)(G * r);
    g_beta[c] = (float)Gsh;
}
}  // namespace

// ---- the closed-form part of the second-stage box BatchNorms (scripts/model.py:105-110: conv_{1,2,3}_1) --------------------
// Outside its computed box E the output of a stride-1, padding-1 convolution whose input is the per-channel constant bg takes one of
// 27 values per output channel -- which taps fall off the canvas along each axis (edge class a in {low face, interior, high face}:
// tap d valid iff not (a = 0 and d = 0) and not (a = 2 and d = 2)):
//     val[co][a,b,c] = sum_{d,h,w valid for (a,b,c)} sum_ci W[co][ci][d,h,w] bg[ci]
// and what the BatchNorm statistics need of them is A1[co] = sum_cls cnt[cls] val, A2[co] = sum_cls cnt[cls] val^2 (cnt = voxels
// per class, geometry only).  Written with torch this was a five-operand einsum (permutes + batched GEMMs), fp64 casts, products
// and reductions forward and their autograd graph backward: ~30 launches per layer and step.  Here one launch forward (a CTA per
// output channel, fp64 inside) and two backward:  dval = cnt (gA1 + 2 val gA2),  dt[co][tap] = sum_cls valid(cls,tap) dval,
// gW[co][ci][tap] = dt[co][tap] bg[ci],  gbg[ci] = sum_{co,tap} dt[co][tap] W[co][ci][tap] (second launch, fixed order).
namespace {

__device__ __forceinline__ bool tap_valid(int cls, int tap) {
    const int a = cls / 9, b = (cls / 3) % 3, c = cls % 3, d = tap / 9, h = (tap / 3) % 3, w = tap % 3;
    return !((a == 0 && d == 0) || (a == 2 && d == 2) || (b == 0 && h == 0) || (b == 2 && h == 2) || (c == 0 && w == 0) || (c == 2 && w == 2));
}

__global__ void __launch_bounds__(32) outside_sums_fwd_kernel(const float* __restrict__ W, const float* __restrict__ bg,
                                                              const float* __restrict__ cnt, int Cin, double* __restrict__ val,
                                                              double* __restrict__ A1, double* __restrict__ A2) {
    __shared__ double t[27], v[27];
    const int co = blockIdx.x, lane = threadIdx.x;
    if (lane < 27) {
        const float* w = W + (size_t)co * Cin * 27 + lane;
        double acc = 0.0;
        for (int ci = 0; ci < Cin; ++ci) acc += (double)w[(size_t)ci * 27] * (double)bg[ci];
        t[lane] = acc;
    }
    __syncwarp();
    if (lane < 27) {
        double acc = 0.0;
        for (int tap = 0; tap < 27; ++tap)
            if (tap_valid(lane, tap)) acc += t[tap];
        v[lane] = acc;
        val[(size_t)co * 27 + lane] = acc;
    }
    __syncwarp();
    if (lane == 0) {
        double a1 = 0.0, a2 = 0.0;
        for (int cls = 0; cls < 27; ++cls) { const double c = (double)cnt[cls]; a1 += c * v[cls]; a2 += c * v[cls] * v[cls]; }
        A1[co] = a1; A2[co] = a2;
    }
}

__global__ void __launch_bounds__(64) outside_sums_bwd_w_kernel(const float* __restrict__ bg, const float* __restrict__ cnt,
                                                                const double* __restrict__ val, const double* __restrict__ gA1,
                                                                const double* __restrict__ gA2, int Cin, float* __restrict__ dt_out,
                                                                float* __restrict__ gW) {
    __shared__ double dv[27];
    __shared__ float dt[27];
    const int co = blockIdx.x;
    if (threadIdx.x < 27) {
        const int cls = threadIdx.x;
        dv[cls] = (double)cnt[cls] * ((gA1 ? gA1[co] : 0.0) + 2.0 * val[(size_t)co * 27 + cls] * (gA2 ? gA2[co] : 0.0));
    }
    __syncthreads();
    if (threadIdx.x < 27) {
        const int tap = threadIdx.x;
        double acc = 0.0;
        for (int cls = 0; cls < 27; ++cls)
            if (tap_valid(cls, tap)) acc += dv[cls];
        dt[tap] = (float)acc;
        dt_out[(size_t)co * 27 + tap] = (float)acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Cin * 27; i += blockDim.x) gW[(size_t)co * Cin * 27 + i] = dt[i % 27] * bg[i / 27];
}

__global__ void __launch_bounds__(32) outside_sums_bwd_bg_kernel(const float* __restrict__ W, const float* __restrict__ dt, int Cout,
                                                                 int Cin, float* __restrict__ gbg) {
    const int ci = blockIdx.x, lane = threadIdx.x;
    double acc = 0.0;
    if (lane < 27)
        for (int co = 0; co < Cout; ++co) acc += (double)dt[(size_t)co * 27 + lane] * (double)W[((size_t)co * Cin + ci) * 27 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) gbg[ci] = (float)acc;
}

}  // namespace

extern "C" int mvsb200_outside_sums_fwd(const float* W, const float* bg, const float* cnt27, int Cout, int Cin, double* val,
                                        double* A1, double* A2, void* stream) {
    MVS_REQUIRE(W && bg && cnt27 && val && A1 && A2, "outside_sums_fwd: null pointer");
    MVS_REQUIRE(Cout >= 1 && Cout <= 4096 && Cin >= 1 && Cin <= 4096, "outside_sums_fwd: bad shape");
    outside_sums_fwd_kernel<<<Cout, 32, 0, (cudaStream_t)stream>>>(W, bg, cnt27, Cin, val, A1, A2);
    MVS_CHECK_LAUNCH("outside_sums_fwd");
    return MVSB200_OK;
}

extern "C" int mvsb200_outside_sums_bwd(const float* W, const float* bg, const float* cnt27, const double* val, const double* gA1,
                                        const double* gA2, int Cout, int Cin, float* dt_workspace, float* gW, float* gbg, void* stream) {
    MVS_REQUIRE(W && bg && cnt27 && val && dt_workspace && gW && gbg, "outside_sums_bwd: null pointer");
    MVS_REQUIRE(Cout >= 1 && Cout <= 4096 && Cin >= 1 && Cin <= 4096, "outside_sums_bwd: bad shape");
    outside_sums_bwd_w_kernel<<<Cout, 64, 0, (cudaStream_t)stream>>>(bg, cnt27, val, gA1, gA2, Cin, dt_workspace, gW);
    MVS_CHECK_LAUNCH("outside_sums_bwd_w");
    outside_sums_bwd_bg_kernel<<<Cin, 32, 0, (cudaStream_t)stream>>>(W, dt_workspace, Cout, Cin, gbg);
    MVS_CHECK_LAUNCH("outside_sums_bwd_bg");
    return MVSB200_OK;
}

extern "C" int mvsb200_box_bn_algebra_fwd(const float* s1, const float* s2, const double* add1, const double* add2, int C, double n_full,
                                          const float* gamma, const float* beta,
                                          double eps, double momentum, float* running_mean, float* running_var,
                                          int64_t* num_batches_tracked, float* scale, float* shift, float* mean, float* var,
                                          double* stat64, void* stream) {
    MVS_REQUIRE(s1 && s2 && gamma && beta && scale && shift && mean && var && stat64, "box_bn_algebra_fwd: null pointer");
    MVS_REQUIRE(C >= 1 && C <= 4096 && n_full >= 1.0, "box_bn_algebra_fwd: bad shape");
    MVS_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "box_bn_algebra_fwd: running_mean and running_var go together");
    box_bn_algebra_fwd_kernel<<<(C + 63) / 64, 64, 0, (cudaStream_t)stream>>>(s1, s2, add1, add2, C, n_full, gamma, beta, eps, momentum, running_mean,
                                                                             running_var, reinterpret_cast<long long*>(num_batches_tracked),
                                                                             scale, shift, mean, var, stat64);
    MVS_CHECK_LAUNCH("box_bn_algebra_fwd");
    return MVSB200_OK;
}

extern "C" int mvsb200_box_bn_algebra_bwd(const float* gscale, const float* gshift, const float* gscale_ext, const float* gshift_ext,
                                          const double* stat64, const float* gamma, int C, double n_full, float* a, float* b2,
                                          float* g_gamma, float* g_beta, void* stream) {
    MVS_REQUIRE(gscale && gshift && stat64 && gamma && a && b2 && g_gamma && g_beta, "box_bn_algebra_bwd: null pointer");
    MVS_REQUIRE(C >= 1 && C <= 4096 && n_full >= 1.0, "box_bn_algebra_bwd: bad shape");
    box_bn_algebra_bwd_kernel<<<(C + 63) / 64, 64, 0, (cudaStream_t)stream>>>(gscale, gshift, gscale_ext, gshift_ext, stat64, gamma, C, n_full,
                                                                             a, b2, g_gamma, g_beta);
    MVS_CHECK_LAUNCH("box_bn_algebra_bwd");
    return MVSB200_OK;
}
