// PTX wrappers (mbarrier, TMA, tcgen05 / TMEM, cluster) and small warp-level helpers shared by the tcgen05 kernels
// of tower_bf16.cu (throughput: CTA pairs, whole tower per launch) and tower_lat.cu (latency: one 2-board tile split
// over a cluster of 4 / 8 CTAs).  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace scb {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;

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
// ---- cta_group::2 (CTA pair) variants -------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster addresses carry the CTA rank in bit 24; clearing it addresses the even (leader) CTA
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                             int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
            "r"(dst),
        "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            dst),
        "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc2_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc2_commit_mc(uint32_t bar)
{
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"((uint16_t)3)
        : "memory");
}
// arrive on a barrier of the leader CTA (rank 0) of the pair
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar)
{
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
        "}" ::"r"(bar)
        : "memory");
}

// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t p;
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t"
        "}"
        : "=r"(p));
    return p != 0;
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

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32])
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
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Sum over the 32 lanes of a warp of 32 per-lane values, result for index `lane` lands in
// lane `lane` (recursive halving: 16+8+4+2+1 = 31 shuffles instead of 32 x 5).
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane)
{
    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2, b1 = lane & 1;
    float w16[16], w8[8], w4[4], w2[2];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        float send = b16 ? v[i] : v[i + 16];
        float keep = b16 ? v[i + 16] : v[i];
        w16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float send = b8 ? w16[i] : w16[i + 8];
        float keep = b8 ? w16[i + 8] : w16[i];
        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float send = b4 ? w8[i] : w8[i + 4];
        float keep = b4 ? w8[i + 4] : w8[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        float send = b2 ? w4[i] : w4[i + 2];
        float keep = b2 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    float send = b1 ? w2[0] : w2[1];
    float keep = b1 ? w2[1] : w2[0];
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

__device__ __forceinline__ void bf16x8_to_float(const uint4 &u, float (&f)[8])
{
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// Column sums over the 16 lanes of each board for rows ordered (rank, board, file): lane bit 3 is the board, so
// the recursive halving skips that bit.  Lane l ends with the sums of its board for indices i0, i0 + 1,
// i0 = 16 b16 + 8 b4 + 4 b2 + 2 b1 (30 shuffles).
__device__ __forceinline__ float2 warp_transpose_reduce_2boards(float (&v)[32], int lane, int &i0)
{
    const bool b16 = lane & 16, b4 = lane & 4, b2 = lane & 2, b1 = lane & 1;
    float w16[16], w8[8], w4[4], w2[2];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        float send = b16 ? v[i] : v[i + 16];
        float keep = b16 ? v[i + 16] : v[i];
        w16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float send = b4 ? w16[i] : w16[i + 8];
        float keep = b4 ? w16[i + 8] : w16[i];
        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float send = b2 ? w8[i] : w8[i + 4];
        float keep = b2 ? w8[i + 4] : w8[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        float send = b1 ? w4[i] : w4[i + 2];
        float keep = b1 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    i0 = (b16 ? 16 : 0) + (b4 ? 8 : 0) + (b2 ? 4 : 0) + (b1 ? 2 : 0);
    return make_float2(w2[0], w2[1]);
}


// ---- LayerNorm arithmetic shared by both kernels -------------------------------------------------------------
// Written with explicit round-to-nearest intrinsics so that the compiler cannot contract / reassociate it
// differently in different kernels: the latency kernel must produce the bits of the throughput kernel.
// Statistics: per 32-channel chunk k the sums (s_k, q_k) of a = acc + bias and a * a (ln_chunk_stats: two 16-channel
// halves, each four interleaved chains and a tree); combined as
//   half0 = (p0 + p1) + (p2 + p3), half1 = (p4 + p5) + (p6 + p7), total = half0 + half1
// -- a fixed tree, so the chunks may be computed by different threads / CTAs.
// half a chunk: 16 channels as four interleaved sequential chains (j mod 4) and a fixed tree
__device__ __forceinline__ void ln_half_chunk_stats(const float *a /*[16]*/, float &s, float &q)
{
    float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 16; j++) {
        s4[j & 3] = __fadd_rn(s4[j & 3], a[j]);
        q4[j & 3] = __fmaf_rn(a[j], a[j], q4[j & 3]);
    }
    s = __fadd_rn(__fadd_rn(s4[0], s4[1]), __fadd_rn(s4[2], s4[3]));
    q = __fadd_rn(__fadd_rn(q4[0], q4[1]), __fadd_rn(q4[2], q4[3]));
}
// a 32-channel chunk = its two halves added (the latency kernel computes the halves in two threads)
__device__ __forceinline__ float2 ln_join_halves(const float2 h0, const float2 h1)
{
    return make_float2(__fadd_rn(h0.x, h1.x), __fadd_rn(h0.y, h1.y));
}
__device__ __forceinline__ void ln_chunk_stats(const float (&a)[32], float &s, float &q)
{
    float2 h0, h1;
    ln_half_chunk_stats(a, h0.x, h0.y);
    ln_half_chunk_stats(a + 16, h1.x, h1.y);
    const float2 c = ln_join_halves(h0, h1);
    s = c.x;
    q = c.y;
}
__device__ __forceinline__ float2 ln_half(const float2 p0, const float2 p1, const float2 p2, const float2 p3)
{
    return make_float2(__fadd_rn(__fadd_rn(p0.x, p1.x), __fadd_rn(p2.x, p3.x)),
                       __fadd_rn(__fadd_rn(p0.y, p1.y), __fadd_rn(p2.y, p3.y)));
}
__device__ __forceinline__ void ln_finish(const float2 h0, const float2 h1, float eps, float &mean, float &rstd)
{
    mean = __fmul_rn(__fadd_rn(h0.x, h1.x), 1.f / 256.f);
    const float var = fmaxf(__fsub_rn(__fmul_rn(__fadd_rn(h0.y, h1.y), 1.f / 256.f), __fmul_rn(mean, mean)), 0.f);
    rstd = rsqrtf(__fadd_rn(var, eps));
}
// (a - mean) * rstd * gamma + beta with a = acc + bias already formed
__device__ __forceinline__ float ln_apply(float a, float mean, float rstd, float gamma, float beta)
{
    return __fmaf_rn(__fmul_rn(__fsub_rn(a, mean), rstd), gamma, beta);
}

// squeeze-excitation scalar pieces, shared for the same reason
__device__ __forceinline__ float se_hidden(float b1, float half0, float half1)
{
    return fmaxf(__fadd_rn(__fadd_rn(b1, half0), half1), 0.f);
}
// Both SE matrix-vector products are defined chunk-wise so that the chunks can be computed by different threads / CTAs:
// per 32-input chunk one sequential fma chain from 0 (se_chain8 = 8 inputs of it), the chunk sums combined in fixed trees.
__device__ __forceinline__ float se_chain8(const float (&w)[8], const float4 a, const float4 b, float acc)
{
    acc = fmaf(w[0], a.x, acc); acc = fmaf(w[1], a.y, acc); acc = fmaf(w[2], a.z, acc); acc = fmaf(w[3], a.w, acc);
    acc = fmaf(w[4], b.x, acc); acc = fmaf(w[5], b.y, acc); acc = fmaf(w[6], b.z, acc); acc = fmaf(w[7], b.w, acc);
    return acc;
}
__device__ __forceinline__ float se_tree4(float p0, float p1, float p2, float p3)
{
    return __fadd_rn(__fadd_rn(p0, p1), __fadd_rn(p2, p3));
}
// FC2 pre-activation: bias + the four 32-hidden-unit chunk sums
__device__ __forceinline__ float se_fc2_sum(float b2, float c0, float c1, float c2, float c3)
{
    return __fadd_rn(b2, se_tree4(c0, c1, c2, c3));
}
__device__ __forceinline__ float se_sigmoid(float g) { return __fdiv_rn(1.f, __fadd_rn(1.f, __expf(-g))); }

}  // namespace scb
