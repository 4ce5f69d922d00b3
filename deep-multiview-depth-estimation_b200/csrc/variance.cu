// assemble_cost_volume on an already-materialised warped tensor (scripts/costvolume.py:3-16).
// The fused K1 kernel is the product path; this covers callers that hand the drop-in a plain
// [B*V, C, D, h, w] tensor (viewed as [B][V][M], M = C*D*h*w).  Pure streaming: reads V*M, writes M.
#include "common.cuh"
using namespace mvsb200;

namespace {
constexpr int kMaxV = 16;

__global__ void variance_views_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int V, size_t M) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= M) return;
    const float* p = x + (size_t)b * V * M + i;
    float sum = 0.f;
    for (int v = 0; v < V; ++v) sum += p[(size_t)v * M];
    const float mean = sum / (float)V;
    float acc = 0.f;
    for (int v = 0; v < V; ++v) { const float d = p[(size_t)v * M] - mean; acc = fmaf(d, d, acc); }
    out[(size_t)b * M + i] = acc / (float)V;
}

__global__ void variance_views_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gout,
                                          float* __restrict__ gx, int V, size_t M) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= M) return;
    const float* p = x + (size_t)b * V * M + i;
    float sum = 0.f;
    for (int v = 0; v < V; ++v) sum += p[(size_t)v * M];
    const float mean = sum / (float)V, gs = gout[(size_t)b * M + i] * (2.0f / (float)V);
    for (int v = 0; v < V; ++v) gx[(size_t)b * V * M + (size_t)v * M + i] = (p[(size_t)v * M] - mean) * gs;
}
}  // namespace

extern "C" int mvsb200_variance_views_fwd(const float* x, float* out, int B, int V, int64_t M, void* stream) {
    MVS_REQUIRE(x && out, "variance_views_fwd: null pointer");
    MVS_REQUIRE(B >= 1 && B <= 65535 && V >= 1 && V <= kMaxV && M >= 1, "variance_views_fwd: bad shape");
    dim3 grid((unsigned)((M + 255) / 256), B);
    variance_views_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, V, (size_t)M);
    MVS_CHECK_LAUNCH("variance_views_fwd");
    return MVSB200_OK;
}

extern "C" int mvsb200_variance_views_bwd(const float* x, const float* gout, float* gx, int B, int V, int64_t M, void* stream) {
    MVS_REQUIRE(x && gout && gx, "variance_views_bwd: null pointer");
    MVS_REQUIRE(B >= 1 && B <= 65535 && V >= 1 && V <= kMaxV && M >= 1, "variance_views_bwd: bad shape");
    dim3 grid((unsigned)((M + 255) / 256), B);
    variance_views_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, gout, gx, V, (size_t)M);
    MVS_CHECK_LAUNCH("variance_views_bwd");
    return MVSB200_OK;
}
