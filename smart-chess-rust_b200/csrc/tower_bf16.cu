// bf16 throughput mode of the policy/value network: tcgen05 / TMEM / TMA implicit-GEMM
// convolutions with the bias + LayerNorm(C) (+ReLU) epilogue fused in, for sm_100a.
//
// Network arithmetic restated from py/module.py:120-126 (stem), :38-46 (ResBlockSE),
// :70-76 / :89-93 (head 1x1 convs); LayerNorm2d = LN over channels, eps 1e-6 (timm).
//
// GEMM view of one 3x3 layer: M = 64 * boards, N = 256 output channels, K = 9 taps x Cin.
//   A (activations, bf16 NHWC [board][rank][file][C]) is never materialised as im2col: for
//   tap (dy,dx) and channel chunk kc the TMA loads the 4-D box {64 ch, 8 files, 8 ranks,
//   2 boards} at coordinates (64*kc, dx, dy, board0); ranks/files outside [0,8) are
//   zero-filled by the TMA unit, which IS the conv padding.  The box lands in shared memory
//   as 128 rows x 128 bytes, 128B-swizzled = the canonical K-major UMMA operand layout.
//   B (weights, bf16 [tap][cout][cin]) is a plain 2-D box {64, 256}.
//   D accumulates in TMEM (128 lanes x 256 fp32 columns per tile, two tiles = all 512
//   columns, so the epilogue of tile i overlaps the MMAs of tile i+1).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (each owns the TMEM lane quadrant warp_id % 4; one thread = one
// (board, square) row, so LayerNorm over channels is a per-thread reduction).
#include <cuda.h>

#include <vector>

#include "common.cuh"

namespace scb {

constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_BK = 64;
constexpr int TC_STAGES = 4;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;  // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 2;  // 32 KB
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_THREADS = 192;
constexpr int TC_SMEM_PARAMS = 3 * TC_BN * 4;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + TC_SMEM_PARAMS + 256 + 1024;

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                            int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
            "r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (sm_100 UMMA):
//   bits 0-13 start address >> 4, 16-29 leading byte offset >> 4 (unused for swizzled
//   K-major, 1), 32-45 stride byte offset >> 4 (8 rows x 128 B = 1024 B between 8-row core
//   groups), 46-47 descriptor version 1, 61-63 layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor: fp32 accumulate (bit 4), A/B bf16 (bits 7, 10), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- the kernel ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_ln_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                  __nv_bfloat16 *__restrict__ out, const float *__restrict__ bias, const float *__restrict__ gamma,
                  const float *__restrict__ beta, int n_tiles, int taps, int kchunks, int relu)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float *s_bias = reinterpret_cast<float *>(smem + TC_STAGES * TC_STAGE_BYTES);
    float *s_gamma = s_bias + TC_BN;
    float *s_beta = s_gamma + TC_BN;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_beta + TC_BN);
    // bars: [0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * TC_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * TC_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * TC_STAGES + 2 + s); };

    for (int i = threadIdx.x; i < TC_BN; i += TC_THREADS) {
        s_bias[i] = bias[i];
        s_gamma[i] = gamma[i];
        s_beta[i] = beta[i];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nkb = taps * kchunks;

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int board0 = tile * 2;
                for (int tap = 0; tap < taps; tap++) {
                    const int dy = taps == 9 ? tap / 3 - 1 : 0;
                    const int dx = taps == 9 ? tap % 3 - 1 : 0;
                    for (int kc = 0; kc < kchunks; kc++) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        const uint32_t a_dst = smem_base + stage * TC_STAGE_BYTES;
                        const uint32_t b_dst = a_dst + TC_A_BYTES;
                        mbar_expect_tx(full_bar(stage), TC_STAGE_BYTES);
                        tma_load_4d(a_dst, &map_a, full_bar(stage), kc * TC_BK, dx, dy, board0);
                        tma_load_2d(b_dst, &map_w, full_bar(stage), kc * TC_BK, tap * TC_BN);
                        if (++stage == TC_STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(TC_BM, TC_BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
                const int as = it & 1;
                const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
                mbar_wait(tempty_bar(as), aphase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * TC_BN);
                for (int kb = 0; kb < nkb; kb++) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * TC_STAGE_BYTES;
                    const uint64_t da = umma_desc_sw128(a_addr);
                    const uint64_t db = umma_desc_sw128(a_addr + TC_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; k++) {
                        // +32 bytes per K=16 slice inside the 128-byte swizzle row
                        tc_mma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                                    (uint32_t)((kb | k) != 0));
                    }
                    tc_commit(empty_bar(stage));
                    if (++stage == TC_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                tc_commit(tfull_bar(as));
            }
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const int as = it & 1;
            const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * TC_BN);
            uint32_t r[32];
            float sum = 0.f;
#pragma unroll 1
            for (int ch = 0; ch < TC_BN / 32; ch++) {
                tmem_ld32(taddr + ch * 32, r);
#pragma unroll
                for (int j = 0; j < 32; j++) sum += __uint_as_float(r[j]) + s_bias[ch * 32 + j];
            }
            const float mean = sum * (1.f / TC_BN);
            float sq = 0.f;
#pragma unroll 1
            for (int ch = 0; ch < TC_BN / 32; ch++) {
                tmem_ld32(taddr + ch * 32, r);
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    float d = __uint_as_float(r[j]) + s_bias[ch * 32 + j] - mean;
                    sq = fmaf(d, d, sq);
                }
            }
            const float rstd = rsqrtf(sq * (1.f / TC_BN) + LN_EPS);
            uint4 *orow = reinterpret_cast<uint4 *>(out + ((size_t)tile * TC_BM + row) * TC_BN);
#pragma unroll 1
            for (int ch = 0; ch < TC_BN / 32; ch++) {
                tmem_ld32(taddr + ch * 32, r);
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int c = ch * 32 + 2 * j;
                    float y0 = (__uint_as_float(r[2 * j]) + s_bias[c] - mean) * rstd * s_gamma[c] + s_beta[c];
                    float y1 = (__uint_as_float(r[2 * j + 1]) + s_bias[c + 1] - mean) * rstd * s_gamma[c + 1] + s_beta[c + 1];
                    if (relu) {
                        y0 = fmaxf(y0, 0.f);
                        y1 = fmaxf(y1, 0.f);
                    }
                    __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
                    pk[j] = *reinterpret_cast<uint32_t *>(&h);
                }
#pragma unroll
                for (int q = 0; q < 4; q++) orow[ch * 4 + q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(as));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct ActMap {
    const void *ptr;
    int boards, c;
    CUtensorMap map;
};

struct TcConv {
    CUtensorMap map_w;
    int taps, cin_pad;
    const float *bias, *gamma, *beta;
    std::vector<ActMap> act_maps;
};

static int make_act_map(const void *ptr, int boards, int c, CUtensorMap *m)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return SC_E_CUDA;
    }
    cuuint64_t dims[4] = {(cuuint64_t)c, 8, 8, (cuuint64_t)boards};
    cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)c * 16, (cuuint64_t)c * 128};
    cuuint32_t box[4] = {TC_BK, 8, 8, 2};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(activations) failed: " + std::to_string((int)r));
        return SC_E_CUDA;
    }
    return SC_OK;
}

int tc_conv_create(TcConv **out, const __nv_bfloat16 *w, int taps, int cin_pad, const float *bias,
                   const float *gamma, const float *beta)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return SC_E_CUDA;
    }
    TcConv *c = new TcConv();
    c->taps = taps;
    c->cin_pad = cin_pad;
    c->bias = bias;
    c->gamma = gamma;
    c->beta = beta;
    cuuint64_t dims[2] = {(cuuint64_t)cin_pad, (cuuint64_t)taps * TC_BN};
    cuuint64_t strides[1] = {(cuuint64_t)cin_pad * 2};
    cuuint32_t box[2] = {TC_BK, TC_BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&c->map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16 *>(w), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        delete c;
        set_error("cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
        return SC_E_CUDA;
    }
    static bool attr_set = false;
    if (!attr_set) {
        SCB_CUDA(cudaFuncSetAttribute(tc_conv_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        attr_set = true;
    }
    *out = c;
    return SC_OK;
}

void tc_conv_destroy(TcConv *c) { delete c; }

int tc_conv_launch(TcConv *c, const __nv_bfloat16 *in, int n_boards_alloc, int n_boards, __nv_bfloat16 *out,
                   int relu, int num_sms, cudaStream_t st)
{
    if (n_boards <= 0) return SC_OK;
    const CUtensorMap *ma = nullptr;
    for (auto &m : c->act_maps)
        if (m.ptr == in && m.boards == n_boards_alloc && m.c == c->cin_pad) ma = &m.map;
    if (!ma) {
        ActMap am;
        am.ptr = in;
        am.boards = n_boards_alloc;
        am.c = c->cin_pad;
        SCB_CHECK(make_act_map(in, n_boards_alloc, c->cin_pad, &am.map));
        c->act_maps.push_back(am);
        ma = &c->act_maps.back().map;
    }
    const int n_tiles = (n_boards + 1) / 2;
    const int grid = n_tiles < num_sms ? n_tiles : num_sms;
    tc_conv_ln_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(*ma, c->map_w, out, c->bias, c->gamma, c->beta, n_tiles,
                                                               c->taps, c->cin_pad / TC_BK, relu);
    SCB_CUDA(cudaGetLastError());
    return SC_OK;
}

// ---- squeeze-excitation + residual + ReLU on bf16 activations (py/module.py:43-45) ---------
__global__ void __launch_bounds__(256) se_res_bf16_kernel(const __nv_bfloat16 *__restrict__ y,
                                                          const __nv_bfloat16 *__restrict__ x,
                                                          __nv_bfloat16 *__restrict__ out,
                                                          const float *__restrict__ w1t, const float *__restrict__ b1,
                                                          const float *__restrict__ w2t, const float *__restrict__ b2)
{
    __shared__ float s_mean[C_TOWER];
    __shared__ float s_hid[C_SE];
    const int b = blockIdx.x, c = threadIdx.x;
    const __nv_bfloat16 *yb = y + (size_t)b * 64 * C_TOWER;
    float yv[64];
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < 64; s++) {
        yv[s] = __bfloat162float(yb[s * C_TOWER + c]);
        sum += yv[s];
    }
    s_mean[c] = sum * (1.f / 64.f);
    __syncthreads();
    if (c < C_SE) {
        float a = b1[c];
        for (int k = 0; k < C_TOWER; k++) a = fmaf(w1t[k * C_SE + c], s_mean[k], a);
        s_hid[c] = fmaxf(a, 0.f);
    }
    __syncthreads();
    float g = b2[c];
    for (int k = 0; k < C_SE; k++) g = fmaf(w2t[k * C_TOWER + c], s_hid[k], g);
    g = 1.f / (1.f + __expf(-g));
    const __nv_bfloat16 *xb = x + (size_t)b * 64 * C_TOWER;
    __nv_bfloat16 *ob = out + (size_t)b * 64 * C_TOWER;
#pragma unroll
    for (int s = 0; s < 64; s++)
        ob[s * C_TOWER + c] = __float2bfloat16(fmaxf(fmaf(g, yv[s], __bfloat162float(xb[s * C_TOWER + c])), 0.f));
}

int launch_se_res_bf16(const __nv_bfloat16 *y, const __nv_bfloat16 *x, __nv_bfloat16 *out, int n, const float *w1t,
                       const float *b1, const float *w2t, const float *b2, cudaStream_t st)
{
    if (n <= 0) return SC_OK;
    se_res_bf16_kernel<<<n, 256, 0, st>>>(y, x, out, w1t, b1, w2t, b2);
    SCB_CUDA(cudaGetLastError());
    return SC_OK;
}

// ---- policy conv 256 -> 73 + LayerNorm(73) (py/module.py:73-74), CUDA-core version ----------
// one block per board: 64 rows x 80 (73 valid) outputs, thread = (row, 20-column quarter)
__global__ void __launch_bounds__(256) policy_conv2_bf16_kernel(const __nv_bfloat16 *__restrict__ p1,
                                                                const float *__restrict__ w /*[256][80]*/,
                                                                const float *__restrict__ bias,
                                                                const float *__restrict__ gamma,
                                                                const float *__restrict__ beta,
                                                                float *__restrict__ logits)
{
    __shared__ __align__(16) float s_w[32][LD_POLICY];
    __shared__ __nv_bfloat16 s_a[64][C_TOWER + 2];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int r = tid >> 2, q = tid & 3;
    const __nv_bfloat16 *src = p1 + (size_t)b * 64 * C_TOWER;
    for (int i = tid; i < 64 * C_TOWER; i += 256) s_a[i >> 8][i & 255] = src[i];
    float acc[20];
#pragma unroll
    for (int j = 0; j < 20; j++) acc[j] = 0.f;
    for (int k0 = 0; k0 < C_TOWER; k0 += 32) {
        __syncthreads();
        for (int i = tid; i < 32 * LD_POLICY; i += 256) s_w[i / LD_POLICY][i % LD_POLICY] = w[(size_t)k0 * LD_POLICY + i];
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < 32; k++) {
            const float a = __bfloat162float(s_a[r][k0 + k]);
            const float4 *wp = reinterpret_cast<const float4 *>(&s_w[k][q * 20]);
#pragma unroll
            for (int j = 0; j < 5; j++) {
                float4 v = wp[j];
                acc[4 * j + 0] = fmaf(a, v.x, acc[4 * j + 0]);
                acc[4 * j + 1] = fmaf(a, v.y, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(a, v.z, acc[4 * j + 2]);
                acc[4 * j + 3] = fmaf(a, v.w, acc[4 * j + 3]);
            }
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 20; j++) {
        int c = q * 20 + j;
        acc[j] = c < C_POLICY ? acc[j] + bias[c] : 0.f;
        sum += acc[j];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float mean = sum * (1.f / C_POLICY);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 20; j++) {
        int c = q * 20 + j;
        float d = c < C_POLICY ? acc[j] - mean : 0.f;
        sq = fmaf(d, d, sq);
    }
    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
    const float rstd = rsqrtf(sq * (1.f / C_POLICY) + LN_EPS);
    float *o = logits + ((size_t)b * 64 + r) * LD_POLICY + q * 20;
#pragma unroll
    for (int j = 0; j < 20; j++) {
        int c = q * 20 + j;
        o[j] = c < C_POLICY ? (acc[j] - mean) * rstd * gamma[c] + beta[c] : 0.f;
    }
}

int launch_policy_conv2_bf16(const __nv_bfloat16 *p1, int n, const float *w, const float *bias, const float *gamma,
                             const float *beta, float *logits, cudaStream_t st)
{
    if (n <= 0) return SC_OK;
    policy_conv2_bf16_kernel<<<n, 256, 0, st>>>(p1, w, bias, gamma, beta, logits);
    SCB_CUDA(cudaGetLastError());
    return SC_OK;
}

// ---- value FC 16384 -> 128 (py/module.py:95), CUDA-core split-K version ----------------------
// grid (ceil(n/32), n_split); block 256 = 32 boards x 8 column groups of 16
__global__ void __launch_bounds__(256) value_fc_bf16_kernel(const __nv_bfloat16 *__restrict__ v1, int n,
                                                            const __nv_bfloat16 *__restrict__ w,
                                                            float *__restrict__ pre, int k_per_split)
{
    __shared__ __nv_bfloat16 s_a[32][64 + 2];
    __shared__ __align__(16) __nv_bfloat16 s_w[64][N_VALUE_HIDDEN];
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * 32, sp = blockIdx.y;
    const int rb = tid >> 3, cg = tid & 7;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0.f;
    const int kbeg = sp * k_per_split;
    for (int k0 = kbeg; k0 < kbeg + k_per_split; k0 += 64) {
        __syncthreads();
        for (int i = tid; i < 32 * 64; i += 256) {
            int rr = i >> 6, kk = i & 63;
            s_a[rr][kk] = (b0 + rr < n) ? v1[(size_t)(b0 + rr) * (64 * C_TOWER) + k0 + kk] : __float2bfloat16(0.f);
        }
        const uint4 *wsrc = reinterpret_cast<const uint4 *>(w + (size_t)k0 * N_VALUE_HIDDEN);
        uint4 *wdst = reinterpret_cast<uint4 *>(&s_w[0][0]);
        for (int i = tid; i < 64 * N_VALUE_HIDDEN / 8; i += 256) wdst[i] = wsrc[i];
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < 64; k++) {
            const float a = __bfloat162float(s_a[rb][k]);
            const __nv_bfloat162 *wp = reinterpret_cast<const __nv_bfloat162 *>(&s_w[k][cg * 16]);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                float2 f = __bfloat1622float2(wp[j]);
                acc[2 * j] = fmaf(a, f.x, acc[2 * j]);
                acc[2 * j + 1] = fmaf(a, f.y, acc[2 * j + 1]);
            }
        }
    }
    if (b0 + rb < n) {
        float *o = pre + ((size_t)sp * n + b0 + rb) * N_VALUE_HIDDEN + cg * 16;
#pragma unroll
        for (int j = 0; j < 16; j++) o[j] = acc[j];
    }
}

int launch_value_fc_bf16(const __nv_bfloat16 *v1, int n, const __nv_bfloat16 *w, float *hidden_pre, int n_split,
                         cudaStream_t st)
{
    if (n <= 0) return SC_OK;
    const int K = 64 * C_TOWER;
    dim3 grid((n + 31) / 32, n_split);
    value_fc_bf16_kernel<<<grid, 256, 0, st>>>(v1, n, w, hidden_pre, K / n_split);
    SCB_CUDA(cudaGetLastError());
    return SC_OK;
}

}  // namespace scb
