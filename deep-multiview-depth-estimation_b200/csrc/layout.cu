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
