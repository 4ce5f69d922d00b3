// K3 (backward of the stride-2 layers): weight gradient of the stride-2 convolutions and of the stride-2 transposed
// convolutions of the regulariser on tcgen05 tensor cores, ONE launch per layer (sm_100a).
//
// Reference semantics (citations into /root/reference/scripts):
//   model.py:104-110   conv_{1,2,3}_0 = Conv3d(k=3, stride=2, padding=dim/2+1)          big = layer input, small = output gradient
//   model.py:115-121   deconv_{3,2,1}_0 = ConvTranspose3d(k=3, stride=2, ...)           big = output gradient, small = layer input
//   model.py:223-234   the factories; autograd's weight gradient of both is
//       gw[k][cb][cs] = sum_{b,o} big(2o - pad + k)[cb] * small(o)[cs]     k = (kd,kh,kw), big zero outside its volume.
//
// GEMM view: the reduction runs over voxels, so both operands are MN-major (voxel = K row, channels contiguous).  The
// stride-2 read of `big` along w is removed WITHOUT de-interleaving copies: a TMA tensor map presents each line of `big` as
// rows of voxel PAIRS (2*Cb channels per row = one swizzle atom); output voxel ox needs the pairs ox-1 and ox, and an
// MN-major operand may be assembled from swizzle atoms that start at ANY row, so
//     A[K = ox, M = (pair a in {ox-1, ox}, w parity, cb)]   = the pair-row line read from row ox, atoms one row apart
//     B[K = ox, N = cs]                                     = the matching line of `small`
// and one tcgen05.mma chain over a line accumulates all three kw taps of one (kd, kh):  three of the four (a, parity)
// row groups of D are the taps kw = 2(a-1) + parity + pad, the fourth is dropped.  Lines of `big` are addressed directly
// (y = 2 oy - pad + kh, d = 2 od - pad + kd): no de-interleaving along h or depth is needed because a line is the unit.
// A CTA owns ONE depth tap kd (blockIdx % 3) and keeps its three kh accumulator blocks (3 x Cs fp32 columns) in TMEM over
// its whole life.  The TMA producer and the MMA issuer are single-warp serial code, so the pipeline is organised to cost ONE
// barrier wait, ONE expect_tx and ONE commit per line of small (measured: with a barrier per loaded line and run-time
// modulo ring indices the kernel took 0.9 ms whatever the channel counts, MMAs and loads switched off -- the skeleton was
// the bound): a ring STAGE holds a line of small and the two new lines of big it needs (kh = 1, 2; kh = 0 is the previous
// stage's kh = 2 line, or a third line loaded with the first stage of an item).  The three kd CTAs of the same rank walk the
// same lines of `small` at the same time (L2 hits).  Epilogue: fp32 vector reductions into gw.
#include "tc_common.cuh"
#include <stdlib.h>

using namespace mvsb200;

namespace {

constexpr int kWgThreads = 192;
constexpr int kMaxRing = 8;
constexpr int kZeroA = 4096;        // zero operand for the accumulator-initialising MMA: 16 + 8 rows of <= 128 bytes ...
constexpr int kZeroB = 4096;        // ... and two N atoms of 16 rows x 128 bytes, 2048 bytes apart

struct S2WgLineParams {
    int B, Ds, Hs, Ws;              // small volume
    int Db, Hb;                     // planes / lines of big
    int pad_d, pad_h, pad_w;
    int Cs;                         // channels of small = MMA N
    int n_chunks, chunk_ch;         // small is loaded as n_chunks boxes of chunk_ch channels (one swizzle atom each)
    int ksteps;                     // ceil(Ws / 16)
    int line_a_bytes, line_b_bytes; // ring slot of a big line; one chunk of a small line (= stride between its N atoms)
    int stage_bytes, n_stages;      // stage = [small: n_chunks x line_b_bytes][big kh=0][big kh=1][big kh=2]
    int ychunk, nychunks;           // lines of small per item
    float* gw;                      // [27][CB][Cs] fp32, accumulated into
    int dbg;                        // diagnostics (MVSB200_S2WG_DBG): 1 = no MMAs issued, 2 = no TMA loads issued
};

// MN-major operand descriptor with the row pitch (= swizzle span) given at run time
__device__ __forceinline__ uint64_t umma_desc_mn_rt(uint32_t saddr, uint32_t rowb, uint32_t atom_stride) {
    const uint64_t layout = rowb == 128 ? 2 : (rowb == 64 ? 4 : 6);
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((atom_stride >> 4) & 0x3fff) << 16;   // LBO = stride between MN atoms
    d |= (uint64_t)((8 * rowb) >> 4) << 32;               // SBO = stride between 8-row K groups
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int CB>
__global__ void __launch_bounds__(kWgThreads, 1)
conv3d_s2_wgrad_lines_kernel(const __grid_constant__ CUtensorMap tm_big, const __grid_constant__ CUtensorMap tm_small,
                             const __grid_constant__ S2WgLineParams p) {
    constexpr int ROWA = 4 * CB;                        // bytes of a voxel-pair row == swizzle span of A
    constexpr int ATOM_ROWS = 2 * CB;                   // D rows per atom: (w parity, cb)
    const int ROWG = 2 * p.chunk_ch;                    // bytes of a row of a small chunk == swizzle span of B
    // D fp32, A/B bf16, A and B MN-major, N = Cs, M = 128
    const uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.Cs >> 3) << 17) | ((128u >> 4) << 24);

    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    unsigned char* zero_a = smem;
    unsigned char* zero_b = smem + kZeroA;
    unsigned char* stages = zero_b + kZeroB;
    const int slot_b_bytes = p.n_chunks * p.line_b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)p.n_stages * p.stage_bytes);
    uint64_t* full = bars;                      // [kMaxRing]  stage landed
    uint64_t* empty = bars + kMaxRing;          // [kMaxRing]  stage no longer read by any MMA
    uint64_t* done = bars + 2 * kMaxRing;       // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 3 * p.Cs) tmem_cols <<= 1;

    // zero everything the MMAs may read and TMA never writes: the zero operands, the row tails of every ring slot
    {
        uint4* z = reinterpret_cast<uint4*>(smem);
        const int n = (int)((reinterpret_cast<unsigned char*>(bars) - smem) / 16);
        for (int i = threadIdx.x; i < n; i += kWgThreads) z[i] = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_big) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_small) : "memory");
        for (int i = 0; i < kMaxRing; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's depth tap and its share of the (batch, plane of small, line chunk) items of that tap
    const int kd = blockIdx.x % 3, rank = blockIdx.x / 3;
    const int nrank = ((int)gridDim.x - kd + 2) / 3;
    const int od_lo = max(0, (p.pad_d - kd + 1) >> 1);
    const int od_hi = min(p.Ds - 1, (p.Db - 1 + p.pad_d - kd) >> 1);
    const int nod = max(0, od_hi - od_lo + 1);
    const int n_items = p.B * nod * p.nychunks;

    // item -> (b, od, d, [oy0, oy1))
    auto decode = [&](int item, int& b, int& od, int& d, int& oy0, int& oy1) {
        const int yc = item % p.nychunks, r = item / p.nychunks;
        od = od_lo + r % nod;
        b = r / nod;
        d = 2 * od - p.pad_d + kd;
        oy0 = yc * p.ychunk;
        oy1 = min(p.Hs, oy0 + p.ychunk);
    };

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        const uint32_t bytes_a = (uint32_t)ROWA * (p.Ws + 1);
        const uint32_t bytes_b = (uint32_t)ROWG * p.Ws * p.n_chunks;
        int slot = 0, round = 0;                         // ring position of the next stage, times the ring has wrapped
        for (int item = rank; item < n_items; item += nrank) {
            int b, od, d, oy0, oy1;
            decode(item, b, od, d, oy0, oy1);
            for (int oy = oy0; oy < oy1; ++oy) {
                if (round > 0) mbar_wait(empty + slot, (round - 1) & 1);
                if (elect_one()) {
                    unsigned char* st = stages + (size_t)slot * p.stage_bytes;
                    const int y0 = 2 * oy - p.pad_h;
                    const int kh_first = oy == oy0 ? 0 : 1;          // kh = 0 of later lines is the previous stage's kh = 2 line
                    uint32_t bytes = bytes_b;
                    for (int kh = kh_first; kh < 3; ++kh) bytes += (y0 + kh >= 0 && y0 + kh < p.Hb) ? bytes_a : 0u;
                    if (p.dbg & 2) { mbar_arrive(full + slot); }
                    else {
                        mbar_expect_tx(full + slot, bytes);
                        for (int c = 0; c < p.n_chunks; ++c)
                            tma_load_5d(st + (size_t)c * p.line_b_bytes, &tm_small, full + slot, c * p.chunk_ch, 0, oy, od, b);
                        // pairs -1 .. Ws-1 of a line of big: smem row r holds pair r - 1 (pair -1 and pairs beyond the line: zeros)
                        for (int kh = kh_first; kh < 3; ++kh)
                            if (y0 + kh >= 0 && y0 + kh < p.Hb)
                                tma_load_5d(st + slot_b_bytes + (size_t)kh * p.line_a_bytes, &tm_big, full + slot, 0, -1, y0 + kh, d, b);
                    }
                }
                __syncwarp();
                if (++slot == p.n_stages) { slot = 0; ++round; }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        const uint64_t da0 = umma_desc_mn<ROWA>(0, ROWA);                                   // A atoms: one pair row apart
        const uint64_t db0 = umma_desc_mn_rt(0, (uint32_t)ROWG, (uint32_t)p.line_b_bytes);  // B atoms: one chunk apart
        const uint64_t dz0 = umma_desc_mn_rt(0, (uint32_t)ROWG, 2048u);
        const uint32_t a_hi = (uint32_t)(da0 >> 32), b_hi = (uint32_t)(db0 >> 32), z_hi = (uint32_t)(dz0 >> 32);
        const uint32_t sa_lo = (uint32_t)da0 | (smem_u32(stages) >> 4);
        const uint32_t sb_lo = (uint32_t)db0 | (smem_u32(stages) >> 4);
        const uint32_t st16 = (uint32_t)p.stage_bytes >> 4, la16 = (uint32_t)p.line_a_bytes >> 4, big16 = (uint32_t)slot_b_bytes >> 4;
        // the three accumulator blocks start at zero: one MMA each on the all-zero operands
        if (elect_one()) {
            for (int kh = 0; kh < 3; ++kh)
                umma_bf16_lohi(tmem_base + (uint32_t)(kh * p.Cs), (uint32_t)da0 | (smem_u32(zero_a) >> 4), a_hi,
                               (uint32_t)dz0 | (smem_u32(zero_b) >> 4), z_hi, IDESC, 0u);
        }
        __syncwarp();
        int slot = 0, round = 0, prev = 0;
        for (int item = rank; item < n_items; item += nrank) {
            int b, od, d, oy0, oy1;
            decode(item, b, od, d, oy0, oy1);
            for (int oy = oy0; oy < oy1; ++oy) {
                mbar_wait(full + slot, round & 1);
                tc_fence_after();
                if (elect_one()) {
                    const int y0 = 2 * oy - p.pad_h;
                    const uint32_t g_lo = sb_lo + (uint32_t)slot * st16;
                    for (int kh = 0; kh < 3; ++kh) {
                        if (y0 + kh < 0 || y0 + kh >= p.Hb) continue;
                        const uint32_t d_tmem = tmem_base + (uint32_t)(kh * p.Cs);
                        uint32_t a_lo = (kh == 0 && oy != oy0) ? sa_lo + (uint32_t)prev * st16 + big16 + 2u * la16
                                                               : sa_lo + (uint32_t)slot * st16 + big16 + (uint32_t)kh * la16;
                        uint32_t b_lo = g_lo;
                        if (p.dbg & 1) continue;
#pragma unroll 2
                        for (int ks = 0; ks < p.ksteps; ++ks) {
                            umma_bf16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, IDESC, 1u);
                            a_lo += (uint32_t)(16 * ROWA) >> 4;
                            b_lo += (uint32_t)(16 * ROWG) >> 4;
                        }
                    }
                    if (oy != oy0) umma_commit(empty + prev);        // its kh = 2 line was this line's kh = 0
                    if (oy == oy1 - 1) umma_commit(empty + slot);
                }
                __syncwarp();
                prev = slot;
                if (++slot == p.n_stages) { slot = 0; ++round; }
            }
        }
        if (elect_one()) umma_commit(done);
        __syncwarp();
    } else {
        // ===================================== epilogue: TMEM -> gw ==============================
        mbar_wait(done, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int m = q * 32 + lane;                     // D row = (pair atom a, w parity, cb)
        const int a = m / ATOM_ROWS, par = (m / CB) & 1, cb = m % CB;
        const int kw = 2 * (a - 1) + par + p.pad_w;      // pair ox + a - 1, voxel 2(ox + a - 1) + par = 2 ox - pad + kw
        const bool real = kw >= 0 && kw <= 2 && n_items > 0;
        for (int kh = 0; kh < 3; ++kh) {
            float* dst = p.gw + ((size_t)(((kd * 3 + kh) * 3 + (real ? kw : 0)) * CB + cb)) * p.Cs;
            for (int c0 = 0; c0 < p.Cs; c0 += 16) {
                uint32_t v[16];
                tmem_ld<16>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kh * p.Cs + c0), v);
                tmem_ld_wait();
                if (real) {
#pragma unroll
                    for (int k = 0; k < 16; k += 4)
                        red_add_v4(dst + c0 + k, __uint_as_float(v[k]), __uint_as_float(v[k + 1]), __uint_as_float(v[k + 2]),
                                   __uint_as_float(v[k + 3]));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

template <int CB>
int launch_s2_wgrad_lines(const void* big, const void* small, float* gw, int B, int Db, int Hb, int Wb, int Ds, int Hs, int Ws,
                          int Cs, int pad_d, int pad_h, int pad_w, const long long* ss, cudaStream_t st) {
    constexpr int ROWA = 4 * CB;
    constexpr int ATOMS = 128 / (2 * CB);
    EncodeTiledFn enc = encode_fn();
    MVS_REQUIRE(enc != nullptr, "conv3d_s2_wgrad_lines: cuTensorMapEncodeTiled is not available from the driver");
    S2WgLineParams p;
    p.B = B; p.Ds = Ds; p.Hs = Hs; p.Ws = Ws; p.Db = Db; p.Hb = Hb;
    p.pad_d = pad_d; p.pad_h = pad_h; p.pad_w = pad_w;
    p.Cs = Cs;
    p.n_chunks = Cs > 64 ? 2 : 1;
    p.chunk_ch = Cs > 64 ? 64 : Cs;
    const int ROWG = 2 * p.chunk_ch;
    p.ksteps = (Ws + 15) / 16;
    const int rows_a = (Ws + 1 > 16 * p.ksteps + ATOMS ? Ws + 1 : 16 * p.ksteps + ATOMS);
    p.line_a_bytes = (rows_a * ROWA + 1023) / 1024 * 1024;
    p.line_b_bytes = (16 * p.ksteps * ROWG + 1023) / 1024 * 1024;
    const size_t budget = 227 * 1024 - 1024 - kZeroA - kZeroB - 512;
    const size_t slot_b = (size_t)p.n_chunks * p.line_b_bytes;
    p.stage_bytes = (int)slot_b + 3 * p.line_a_bytes;
    p.n_stages = (int)(budget / (size_t)p.stage_bytes);
    if (p.n_stages > kMaxRing) p.n_stages = kMaxRing;
    MVS_REQUIRE(p.n_stages >= 2, "conv3d_s2_wgrad_lines: lines of %d voxels do not fit shared memory (Cb=%d, Cs=%d)", Ws, CB, Cs);

    CUtensorMap tm_big, tm_small;
    {
        const cuuint64_t rowb = (cuuint64_t)2 * CB;          // bytes of one voxel of big
        const cuuint64_t dims[5] = {(cuuint64_t)2 * CB, (cuuint64_t)Wb / 2, (cuuint64_t)Hb, (cuuint64_t)Db, (cuuint64_t)B};
        const cuuint64_t strides[4] = {2 * rowb, rowb * Wb, rowb * Wb * Hb, rowb * Wb * Hb * Db};
        const cuuint32_t box[5] = {(cuuint32_t)2 * CB, (cuuint32_t)(Ws + 1), 1, 1, 1};
        const cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tm_big, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(big), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(ROWA), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MVS_REQUIRE(r == CUDA_SUCCESS, "conv3d_s2_wgrad_lines: cuTensorMapEncodeTiled(big) failed (%d)", (int)r);
    }
    {
        // voxel-row strides of small in bytes: dense channel-last, or those of the allocation the box lives in (ss: elements)
        const cuuint64_t rowb = (cuuint64_t)2 * Cs;
        const cuuint64_t dims[5] = {(cuuint64_t)Cs, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)Ds, (cuuint64_t)B};
        const cuuint64_t dense[4] = {rowb, rowb * Ws, rowb * Ws * Hs, rowb * Ws * Hs * Ds};
        const cuuint64_t strides[4] = {ss ? (cuuint64_t)ss[3] * 2 : dense[0], ss ? (cuuint64_t)ss[2] * 2 : dense[1],
                                       ss ? (cuuint64_t)ss[1] * 2 : dense[2], ss ? (cuuint64_t)ss[0] * 2 : dense[3]};
        const cuuint32_t box[5] = {(cuuint32_t)p.chunk_ch, (cuuint32_t)Ws, 1, 1, 1};
        const cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tm_small, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(small), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(ROWG), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MVS_REQUIRE(r == CUDA_SUCCESS, "conv3d_s2_wgrad_lines: cuTensorMapEncodeTiled(small) failed (%d)", (int)r);
    }
    int sms = 148;
    {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) sms = n;
    }
    if (sms < 3) sms = 3;
    // lines of small per item: balance (items per CTA of a depth tap) x (lines + the extra big lines at an item's start)
    const long nrank = sms / 3;
    long best_cost = -1;
    int best_nyc = 1;
    for (int nyc = 1; nyc <= Hs; ++nyc) {
        const int yc = (Hs + nyc - 1) / nyc;
        if ((long)(nyc - 1) * yc >= Hs) continue;
        const long items = (long)B * Ds * nyc;
        const long cost = ((items + nrank - 1) / nrank) * (yc + 2);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_nyc = nyc; }
    }
    p.nychunks = best_nyc;
    p.ychunk = (Hs + best_nyc - 1) / best_nyc;
    p.gw = gw;
    p.dbg = 0;
    if (const char* e = getenv("MVSB200_S2WG_DBG")) p.dbg = atoi(e);
    if (p.dbg & 4) fprintf(stderr, "s2wg: CB=%d Cs=%d ksteps=%d line_a=%d line_b=%d stage=%d n_stages=%d nyc=%d ychunk=%d\n", CB, Cs, p.ksteps, p.line_a_bytes, p.line_b_bytes, p.stage_bytes, p.n_stages, p.nychunks, p.ychunk);
    const size_t smem = 1024 + kZeroA + kZeroB + (size_t)p.n_stages * p.stage_bytes + 512;
    MVS_CUDA(cudaFuncSetAttribute(conv3d_s2_wgrad_lines_kernel<CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3d_s2_wgrad_lines_kernel<CB><<<dim3((unsigned)sms), kWgThreads, smem, st>>>(tm_big, tm_small, p);
    MVS_CHECK_LAUNCH("conv3d_s2_wgrad_lines");
    return MVSB200_OK;
}

}  // namespace

/* Weight gradient of the stride-2 layers, one launch (see the header of this file):
 *   gw[k][cb][cs] = sum_o big(2o - pad + k)[cb] * small(o)[cs]. */
extern "C" int mvsb200_conv3d_s2_wgrad_lines(const void* big, const void* small, float* gw, int B, int Db, int Hb, int Wb, int Cb,
                                             int Ds, int Hs, int Ws, int Cs, int pad_d, int pad_h, int pad_w,
                                             const int64_t* small_strides4, void* stream) {
    const char* name = "conv3d_s2_wgrad_lines";
    MVS_REQUIRE(big && small && gw, "%s: null pointer", name);
    MVS_REQUIRE(aligned16(big) && aligned16(small) && aligned16(gw), "%s: pointers must be 16-byte aligned", name);
    MVS_REQUIRE(B >= 1 && B <= 65535 && Db >= 1 && Hb >= 1 && Wb >= 2 && Ds >= 1 && Hs >= 1 && Ws >= 1, "%s: bad shape", name);
    MVS_REQUIRE(Wb % 2 == 0, "%s: the strided operand needs an even number of voxels per line (got %d)", name, Wb);
    MVS_REQUIRE(Ws + 1 <= 256, "%s: lines of the dense operand are limited to 255 voxels (got %d)", name, Ws);
    MVS_REQUIRE(Cb == 8 || Cb == 16 || Cb == 32, "%s: the strided operand must have 8, 16 or 32 channels (got %d)", name, Cb);
    MVS_REQUIRE(Cs == 16 || Cs == 32 || Cs == 64 || (Cs > 64 && Cs <= 128 && Cs % 16 == 0),
                "%s: the dense operand needs 16, 32, 64 or 80..128 (multiple of 16) channels (got %d)", name, Cs);
    MVS_REQUIRE(pad_d >= 1 && pad_d <= 2 && pad_h >= 1 && pad_h <= 2 && pad_w >= 1 && pad_w <= 2, "%s: pad must be 1 or 2 per axis", name);
    long long ssv[4];
    const long long* ss = nullptr;
    if (small_strides4) {
        for (int i = 0; i < 4; ++i) {
            ssv[i] = (long long)small_strides4[i];
            MVS_REQUIRE(ssv[i] > 0 && ssv[i] % 8 == 0, "%s: strides of the dense operand must be positive multiples of 8 elements", name);
        }
        ss = ssv;
    }
    cudaStream_t st = (cudaStream_t)stream;
    MVS_CUDA(cudaMemsetAsync(gw, 0, (size_t)27 * Cb * Cs * sizeof(float), st));
    if (Cb == 8) return launch_s2_wgrad_lines<8>(big, small, gw, B, Db, Hb, Wb, Ds, Hs, Ws, Cs, pad_d, pad_h, pad_w, ss, st);
    if (Cb == 16) return launch_s2_wgrad_lines<16>(big, small, gw, B, Db, Hb, Wb, Ds, Hs, Ws, Cs, pad_d, pad_h, pad_w, ss, st);
    return launch_s2_wgrad_lines<32>(big, small, gw, B, Db, Hb, Wb, Ds, Hs, Ws, Cs, pad_d, pad_h, pad_w, ss, st);
}
