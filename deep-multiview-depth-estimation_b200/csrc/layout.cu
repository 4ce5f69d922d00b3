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
