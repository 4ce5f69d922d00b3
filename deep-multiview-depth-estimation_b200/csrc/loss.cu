// Masked L1 training loss and its gradient, one launch each.
//
// Reference semantics (/root/reference/scripts/loss.py:4-41): with mask = (gt != 0) and n_valid[b] = sum of the mask over sample b,
//   l0[b] = sum mask |gt - initial| / n_valid[b],  l1[b] likewise for the refined map,
//   loss = sum_b (l0[b] + l1[b]),  initial_acc = mean_b l0[b],  refined_acc = mean_b l1[b].
// Written with torch these are ~25 elementwise / reduction launches on [B,1,h,w] maps plus as many in the backward; the maps are
// tiny (80 KB each at 160x128), so the step paid launches, not bytes.  Forward: 32 CTAs per sample reduce (n_valid, sum0, sum1) of
// their chunk, the last CTA to finish (ticket counter) combines chunks and samples in a fixed order (deterministic).  Backward: one elementwise pass,
//   d loss / d initial = -(g_loss + g_acc0 / B) * mask * sign(gt - initial) / n_valid[b]      (sign(0) = 0, as torch.abs' gradient).
// A sample without valid pixels gives 0 / 0 = NaN, as the reference does.
#include "common.cuh"
using namespace mvsb200;

namespace {

constexpr int kLossThreads = 256;
constexpr int kLossParts = 32;       // CTAs per sample (a [B,1,128,160] map is 20 K pixels per sample: one CTA per sample was latency-bound)

// workspace layout: [3][B] per-sample (n_valid, l0, l1) | [B][kLossParts][3] partial sums | ticket counter
__global__ void __launch_bounds__(kLossThreads) masked_l1_fwd_kernel(const float* __restrict__ gt, const float* __restrict__ a0,
                                                                     const float* __restrict__ a1, int B, int n,
                                                                     float* __restrict__ ws, float* __restrict__ out3) {
    __shared__ float s_red[3][kLossThreads / 32];
    __shared__ bool s_last;
    float* per_sample = ws;
    float* partial = ws + 3 * (size_t)B;
    unsigned* ticket = reinterpret_cast<unsigned*>(ws + 3 * (size_t)B + 3 * (size_t)B * kLossParts);
    const int b = blockIdx.x / kLossParts, part = blockIdx.x % kLossParts;
    const int chunk = (n + kLossParts - 1) / kLossParts;
    const int i0 = part * chunk, i1 = min(n, i0 + chunk);
    const float* g = gt + (size_t)b * n;
    const float* p0 = a0 + (size_t)b * n;
    const float* p1 = a1 + (size_t)b * n;
    float nv = 0.f, s0 = 0.f, s1 = 0.f;
    for (int i = i0 + threadIdx.x; i < i1; i += kLossThreads) {
        const float t = g[i];
        if (t != 0.f) { nv += 1.f; s0 += fabsf(t - p0[i]); s1 += fabsf(t - p1[i]); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nv += __shfl_xor_sync(0xffffffffu, nv, o);
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[0][warp] = nv; s_red[1][warp] = s0; s_red[2][warp] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        nv = s0 = s1 = 0.f;
        for (int w = 0; w < kLossThreads / 32; ++w) { nv += s_red[0][w]; s0 += s_red[1][w]; s1 += s_red[2][w]; }
        float* pp = partial + ((size_t)b * kLossParts + part) * 3;
        pp[0] = nv; pp[1] = s0; pp[2] = s1;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == (unsigned)(B * kLossParts - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {                // the last CTA combines everything in a fixed order (deterministic)
        __threadfence();
        const volatile float* vp = partial;
        float l0 = 0.f, l1 = 0.f;
        for (int k = 0; k < B; ++k) {
            float n_k = 0.f, s0_k = 0.f, s1_k = 0.f;
            for (int q = 0; q < kLossParts; ++q) {
                const volatile float* pp = vp + ((size_t)k * kLossParts + q) * 3;
                n_k += pp[0]; s0_k += pp[1]; s1_k += pp[2];
            }
            per_sample[k] = n_k;                     // n_valid
            per_sample[B + k] = s0_k / n_k;          // l0
            per_sample[2 * B + k] = s1_k / n_k;      // l1
            l0 += s0_k / n_k; l1 += s1_k / n_k;
        }
        out3[0] = l0 + l1;
        out3[1] = l0 / (float)B;
        out3[2] = l1 / (float)B;
        *ticket = 0u;                                // ready for the next launch (graph replays included)
    }
}

__global__ void __launch_bounds__(kLossThreads) masked_l1_bwd_kernel(const float* __restrict__ gt, const float* __restrict__ a0,
                                                                     const float* __restrict__ a1,
                                                                     const float* __restrict__ per_sample, const float* __restrict__ g3,
                                                                     int B, int n, float* __restrict__ ga0, float* __restrict__ ga1) {
    const long long i = (long long)blockIdx.x * kLossThreads + threadIdx.x;
    if (i >= (long long)B * n) return;
    const int b = (int)(i / n);
    const float inv = 1.f / per_sample[b];
    const float w0 = (g3[0] + g3[1] / (float)B) * inv, w1 = (g3[0] + g3[2] / (float)B) * inv;
    const float t = gt[i];
    const bool valid = t != 0.f;
    const float d0 = t - a0[i], d1 = t - a1[i];
    const float sg0 = d0 > 0.f ? 1.f : (d0 < 0.f ? -1.f : 0.f), sg1 = d1 > 0.f ? 1.f : (d1 < 0.f ? -1.f : 0.f);
    ga0[i] = valid ? -w0 * sg0 : 0.f * w0;       // 0 * w keeps the reference's NaN for a sample without valid pixels
    ga1[i] = valid ? -w1 * sg1 : 0.f * w1;
}

}  // namespace

extern "C" int64_t mvsb200_masked_l1_workspace_floats(int B) { return 3 * (int64_t)B + 3 * (int64_t)B * kLossParts + 1; }

extern "C" int mvsb200_masked_l1_fwd(const float* gt, const float* initial, const float* refined, int B, int n, float* workspace,
                                     float* out3, void* stream) {
    MVS_REQUIRE(gt && initial && refined && workspace && out3, "masked_l1_fwd: null pointer");
    MVS_REQUIRE(B >= 1 && B <= 65535 / kLossParts && n >= 1, "masked_l1_fwd: bad shape");
    masked_l1_fwd_kernel<<<B * kLossParts, kLossThreads, 0, (cudaStream_t)stream>>>(gt, initial, refined, B, n, workspace, out3);
    MVS_CHECK_LAUNCH("masked_l1_fwd");
    return MVSB200_OK;
}

extern "C" int mvsb200_masked_l1_bwd(const float* gt, const float* initial, const float* refined, const float* workspace,
                                     const float* g3, int B, int n, float* g_initial, float* g_refined, void* stream) {
    MVS_REQUIRE(gt && initial && refined && workspace && g3 && g_initial && g_refined, "masked_l1_bwd: null pointer");
    MVS_REQUIRE(B >= 1 && B <= 65535 && n >= 1, "masked_l1_bwd: bad shape");
    const long long tot = (long long)B * n;
    masked_l1_bwd_kernel<<<(unsigned)((tot + kLossThreads - 1) / kLossThreads), kLossThreads, 0, (cudaStream_t)stream>>>(
        gt, initial, refined, workspace, g3, B, n, g_initial, g_refined);
    MVS_CHECK_LAUNCH("masked_l1_bwd");
    return MVSB200_OK;
}
