// NCHW <-> NHWC transposes for the feature maps (7.9 MB per sample: negligible next to the volumes).
#include "common.cuh"
using namespace mvsb200;

// [N][R][Cc] -> [N][Cc][R] tiled transpose through shared memory (R, Cc arbitrary)
__global__ void transpose_last2_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int Cc) {
    __shared__ float tile[32][33];
    const size_t base = (size_t)blockIdx.z * R * Cc;
    int c = blockIdx.x * 32 + threadIdx.x;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int r = blockIdx.y * 32 + j;
        if (r < R && c < Cc) tile[j][threadIdx.x] = src[base + (size_t)r * Cc + c];
    }
    __syncthreads();
    int r = blockIdx.y * 32 + threadIdx.x;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int cc = blockIdx.x * 32 + j;
        if (r < R && cc < Cc) dst[base + (size_t)cc * R + r] = tile[threadIdx.x][j];
    }
}

static int transpose_last2(const float* src, float* dst, int N, int R, int Cc, void* stream, const char* name) {
    MVS_REQUIRE(src && dst, "%s: null pointer", name);
    MVS_REQUIRE(N > 0 && R > 0 && Cc > 0 && N <= 65535, "%s: bad shape N=%d R=%d C=%d", name, N, R, Cc);
    dim3 grid((Cc + 31) / 32, (R + 31) / 32, N), block(32, 8);
    MVS_REQUIRE(grid.y <= 65535, "%s: plane too large", name);
    transpose_last2_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, dst, R, Cc);
    MVS_CHECK_LAUNCH(name);
    return MVSB200_OK;
}

extern "C" int mvsb200_nchw_to_nhwc_f32(const float* src, float* dst, int N, int C, int H, int W, void* stream) {
    return transpose_last2(src, dst, N, C, H * W, stream, "nchw_to_nhwc");   // [N][C][HW] -> [N][HW][C]
}
extern "C" int mvsb200_nhwc_to_nchw_f32(const float* src, float* dst, int N, int C, int H, int W, void* stream) {
    return transpose_last2(src, dst, N, H * W, C, stream, "nhwc_to_nchw");   // [N][HW][C] -> [N][C][HW]
}

// 8-channel bf16 voxel rows widened to 16 channels (upper half zero): the data gradient of conv_0_0 (scripts/model.py:101)
// contracts over its 8 output channels, UMMA needs K = 16.  One 16-byte load, one 32-byte store per row.
__global__ void __launch_bounds__(256) widen_rows_8to16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long M) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const uint4 v = __ldg(src + i);
    dst[2 * i] = v;
    dst[2 * i + 1] = make_uint4(0u, 0u, 0u, 0u);
}

extern "C" int mvsb200_widen_rows_8to16_bf16(const void* src, void* dst, int64_t M, void* stream) {
    MVS_REQUIRE(src && dst && aligned16(src) && aligned16(dst), "widen_rows_8to16: null or misaligned pointer");
    MVS_REQUIRE(M > 0 && M < ((int64_t)1 << 38), "widen_rows_8to16: bad row count");
    const long long blocks = (M + 255) / 256;
    widen_rows_8to16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, (uint4*)dst, (long long)M);
    MVS_CHECK_LAUNCH("widen_rows_8to16");
    return MVSB200_OK;
}

// ---- 2D feature encoder: maps as channel-last bf16 rows ---------------------------------------------
// images fp32 [N, 3, H, W] (any element strides) -> bf16 rows [N, H, W, 8]: channels 0..2 the colours, 3..7 zero -- what the
// K = 16 tensor-core convolution reads through an 8-channel TMA map (scripts/model.py:27, the encoder's first layer).
__global__ void __launch_bounds__(256) image_to_rows8_kernel(const float* __restrict__ img, long long s_n, long long s_c, long long s_h,
                                                             long long s_w, int N, int H, int W, uint4* __restrict__ rows) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * H * W) return;
    const int x = (int)(i % W), y = (int)((i / W) % H), n = (int)(i / ((long long)W * H));
    const float* p = img + n * s_n + y * s_h + x * s_w;
    const __nv_bfloat162 a = __floats2bfloat162_rn(p[0], p[s_c]);
    const __nv_bfloat162 b = __floats2bfloat162_rn(p[2 * s_c], 0.f);
    uint4 v;
    v.x = *reinterpret_cast<const unsigned*>(&a);
    v.y = *reinterpret_cast<const unsigned*>(&b);
    v.z = 0u; v.w = 0u;
    rows[i] = v;
}

extern "C" int mvsb200_image_to_rows8(const float* images, const int64_t* strides4_host, int N, int H, int W, void* rows, void* stream) {
    MVS_REQUIRE(images && strides4_host && rows && aligned16(rows), "image_to_rows8: null or misaligned pointer");
    MVS_REQUIRE(N >= 1 && H >= 1 && W >= 1, "image_to_rows8: bad shape");
    const long long tot = (long long)N * H * W;
    image_to_rows8_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        images, strides4_host[0], strides4_host[1], strides4_host[2], strides4_host[3], N, H, W, reinterpret_cast<uint4*>(rows));
    MVS_CHECK_LAUNCH("image_to_rows8");
    return MVSB200_OK;
}

// Space-to-depth of channel-last rows and its inverse: [N, 2h, 2w, C] <-> [N, h, w, 4C], channel (py*2 + px)*C + c of pixel (j, i) =
// channel c of pixel (2j + py, 2i + px).  A 5x5 stride-2 convolution (scripts/model.py:31,39) of the left form is a 3x3 stride-1
// convolution of the right one (tap k = 2t + p of an axis reads parity class p at j - 1 + t), i.e. the tensor-core kernel the
// regulariser already has.  One thread per 16-byte chunk, coalesced on the deep side.
__global__ void __launch_bounds__(256) s2d_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int N, int h, int w,
                                                       int q, int inverse) {
    // q = 16-byte chunks per shallow pixel (C / 8); deep pixel = 4q chunks
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)N * h * w * 4 * q;
    if (i >= tot) return;
    const int k = (int)(i % q), cls = (int)((i / q) % 4), x = (int)((i / (4 * q)) % w), y = (int)((i / ((long long)4 * q * w)) % h);
    const long long n = i / ((long long)4 * q * w * h);
    const long long shallow = ((n * (2 * h) + 2 * y + (cls >> 1)) * (2 * w) + 2 * x + (cls & 1)) * q + k;
    if (inverse) dst[shallow] = __ldg(src + i);
    else dst[i] = __ldg(src + shallow);
}

extern "C" int mvsb200_s2d_rows_bf16(const void* src, void* dst, int N, int h, int w, int C, int inverse, void* stream) {
    MVS_REQUIRE(src && dst && aligned16(src) && aligned16(dst), "s2d_rows: null or misaligned pointer");
    MVS_REQUIRE(N >= 1 && h >= 1 && w >= 1 && C >= 8 && C % 8 == 0, "s2d_rows: bad shape (C must be a multiple of 8)");
    const long long tot = (long long)N * h * w * 4 * (C / 8);
    MVS_REQUIRE(tot < ((long long)1 << 40), "s2d_rows: too large");
    s2d_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), N, h, w, C / 8, inverse);
    MVS_CHECK_LAUNCH("s2d_rows");
    return MVSB200_OK;
}

// ---- filter packing ---------------------------------------------------------------------------------------------------
// The tensor-core kernels want their 3x3x3 filters as bf16 [slot][row][col] matrices (K-major B operands): tap-major, rows =
// output channels padded to a multiple of 16, columns = contraction channels, in the tap order of the kernel (natural, depth
// tap innermost for the kdn form, flipped for data gradients, a trailing all-zero tap for the transposed convolutions), from
// the fp32 parameter [dim0][dim1][3][3][3] of the layer (scripts/model.py:223-234).  Written with torch this is 3..6
// elementwise launches per filter and step (zeros, permute-copy, cast, slice-assign, re-order); here ONE launch:
//     out[s][r][c] = w[r*sr + (c + c0)*sc + tap[s]]   for r < rows_real, c < cols_real, tap[s] >= 0;  0 otherwise.
struct PackTaps { int tap[32]; };
__global__ void __launch_bounds__(256) pack_filter_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int n_slots,
                                                          int n_rows, int n_cols, int rows_real, int cols_real, int c0, int sr,
                                                          int sc, const __grid_constant__ PackTaps taps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots * n_rows * n_cols) return;
    const int c = i % n_cols, r = (i / n_cols) % n_rows, s = i / (n_cols * n_rows);
    const int t = taps.tap[s];
    float v = 0.f;
    if (t >= 0 && r < rows_real && c < cols_real) v = w[(size_t)r * sr + (size_t)(c + c0) * sc + t];
    out[i] = __float2bfloat16(v);
}

extern "C" int mvsb200_pack_filter(const float* w, void* out, int n_slots, int n_rows, int n_cols, int rows_real, int cols_real,
                                   int c0, int sr, int sc, const int* taps_host, void* stream) {
    MVS_REQUIRE(w && out && taps_host, "pack_filter: null pointer");
    MVS_REQUIRE(n_slots >= 1 && n_slots <= 32 && n_rows >= 1 && n_cols >= 1 && rows_real >= 0 && rows_real <= n_rows &&
                cols_real >= 0 && cols_real <= n_cols && c0 >= 0 && sr >= 1 && sc >= 1, "pack_filter: bad shape");
    PackTaps t;
    for (int s = 0; s < 32; ++s) {
        t.tap[s] = s < n_slots ? taps_host[s] : -1;
        MVS_REQUIRE(t.tap[s] >= -1 && t.tap[s] < 27, "pack_filter: tap %d out of range", t.tap[s]);
    }
    const int n = n_slots * n_rows * n_cols;
    pack_filter_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, reinterpret_cast<__nv_bfloat16*>(out), n_slots, n_rows,
                                                                         n_cols, rows_real, cols_real, c0, sr, sc, t);
    MVS_CHECK_LAUNCH("pack_filter");
    return MVSB200_OK;
}
