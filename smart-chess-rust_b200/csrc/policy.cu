// Move index + legal-move gather + renormalisation.
//
// Replaces, per leaf:
//   `Move::rotate` + `Move::encode`           reference src/chess.rs:533-550
//   queenmoves / knightmoves / underpromotions src/queenmoves.rs:3-34, knightmoves.rs:7-31,
//                                              underpromotions.rs:6-33
//   `log_softmax` over all 4672 logits         py/module.py:78-80
//   `_get_move_distribution` (take + exp)      src/backends/torch.rs:148-175, onnx.rs:81-91
//   `post_process_distr` (p / (sum + 1e-5))    src/chess.rs:879-903
//
// One warp per leaf.  The 8x8x73 index is pure integer arithmetic on (from, to, promo); the
// direction tables are 3x3 / 5x5 lookups held in shared memory.  The policy map lives in
// HBM as [leaf][square][LD_POLICY] (square-major, the layout the head's GEMM writes), while
// the reference's flat index is NCHW:  flat = channel*64 + square  (py/module.py:75).  The
// move index i in [0,4672) is applied to that flat vector as-is, i.e. it reads
// channel i/64, square i%64.
#include "common.cuh"
#include "move_index.cuh"

namespace scb {

constexpr int POL_WARPS = 4;

__global__ void __launch_bounds__(POL_WARPS * 32) move_index_kernel(const sc_position *__restrict__ pos,
                                                                    const sc_move *__restrict__ moves,
                                                                    const int32_t *__restrict__ off, int n,
                                                                    int32_t *__restrict__ index_out)
{
    __shared__ int8_t s_q[9], s_k[25];
    if (threadIdx.x < 9) s_q[threadIdx.x] = c_queen_dir[threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 57) s_k[threadIdx.x - 32] = c_knight_type[threadIdx.x - 32];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * POL_WARPS + warp;
    if (b >= n) return;
    const int turn = pos[b].meta[0];
    const int beg = off[b], end = off[b + 1];
    for (int k = beg + lane; k < end; k += 32) index_out[k] = move_index_dev(moves[k], turn, s_q, s_k);
}

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// log-sum-exp of the 4672 logits of one leaf: returns (max, log(sum exp(x - max)))
__device__ __forceinline__ void leaf_lse(const float *__restrict__ lg, int lane, float &mx, float &lsum)
{
    // row s has 73 valid floats at stride LD_POLICY; walk the whole [64][80] tile with
    // coalesced float4 loads and mask the pad columns
    const float4 *p = reinterpret_cast<const float4 *>(lg);
    float m = -INFINITY;
    constexpr int NV = 64 * LD_POLICY / 4;  // 1280 float4
    for (int i = lane; i < NV; i += 32) {
        float4 v = p[i];
        int c = (i * 4) % LD_POLICY;
        if (c + 0 < C_POLICY) m = fmaxf(m, v.x);
        if (c + 1 < C_POLICY) m = fmaxf(m, v.y);
        if (c + 2 < C_POLICY) m = fmaxf(m, v.z);
        if (c + 3 < C_POLICY) m = fmaxf(m, v.w);
    }
    m = warp_max(m);
    float s = 0.f;
    for (int i = lane; i < NV; i += 32) {
        float4 v = p[i];
        int c = (i * 4) % LD_POLICY;
        if (c + 0 < C_POLICY) s += expf(v.x - m);
        if (c + 1 < C_POLICY) s += expf(v.y - m);
        if (c + 2 < C_POLICY) s += expf(v.z - m);
        if (c + 3 < C_POLICY) s += expf(v.w - m);
    }
    s = warp_sum(s);
    mx = m;
    lsum = logf(s);
}

__global__ void __launch_bounds__(POL_WARPS * 32) policy_gather_kernel(const float *__restrict__ logits,
                                                                       const sc_position *__restrict__ pos,
                                                                       const sc_move *__restrict__ moves,
                                                                       const int32_t *__restrict__ off,
                                                                       const int32_t *__restrict__ cnts, int n,
                                                                       float *__restrict__ priors)
{
    __shared__ int8_t s_q[9], s_k[25];
    __shared__ float s_p[POL_WARPS][SC_MAX_MOVES];
    if (threadIdx.x < 9) s_q[threadIdx.x] = c_queen_dir[threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 57) s_k[threadIdx.x - 32] = c_knight_type[threadIdx.x - 32];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * POL_WARPS + warp;
    if (b >= n) return;
    const float *lg = logits + (size_t)b * 64 * LD_POLICY;
    float mx, lsum;
    leaf_lse(lg, lane, mx, lsum);
    const int turn = pos[b].meta[0];
    const int beg = off ? off[b] : b * SC_MAX_MOVES;
    int cnt = off ? off[b + 1] - beg : cnts[b];
    if (cnt > SC_MAX_MOVES) cnt = SC_MAX_MOVES;
    for (int k = lane; k < cnt; k += 32) {
        int idx = move_index_dev(moves[beg + k], turn, s_q, s_k);
        float p = 0.f;
        if (idx >= 0) {
            float x = lg[(idx & 63) * LD_POLICY + (idx >> 6)];
            p = expf((x - mx) - lsum);  // exp(log_softmax(x)[idx])
        }
        s_p[warp][k] = p;
    }
    __syncwarp();
    // `distr.iter().sum::<f32>() + 1e-5` is a left-to-right f32 sum (chess.rs:891)
    float sum = 0.f;
    if (lane == 0) {
        for (int k = 0; k < cnt; k++) sum += s_p[warp][k];
        sum += 1e-5f;
    }
    sum = __shfl_sync(0xffffffffu, sum, 0);
    for (int k = lane; k < cnt; k += 32) priors[beg + k] = __fdiv_rn(s_p[warp][k], sum);
}

// full log_softmax in the reference's flatten order (tolerance gate / debugging)
__global__ void __launch_bounds__(POL_WARPS * 32) policy_logp_full_kernel(const float *__restrict__ logits, int n,
                                                                          float *__restrict__ logp)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * POL_WARPS + warp;
    if (b >= n) return;
    const float *lg = logits + (size_t)b * 64 * LD_POLICY;
    float mx, lsum;
    leaf_lse(lg, lane, mx, lsum);
    float *o = logp + (size_t)b * SC_N_POLICY;
    for (int i = lane; i < SC_N_POLICY; i += 32) o[i] = (lg[(i & 63) * LD_POLICY + (i >> 6)] - mx) - lsum;
}

// training target of `chess_encode_steps` (src/lib.rs:104-112): one warp per ply
__global__ void __launch_bounds__(POL_WARPS * 32) dist_scatter_kernel(const sc_position *__restrict__ pos,
                                                                      const sc_move *__restrict__ moves,
                                                                      const uint32_t *__restrict__ counts,
                                                                      const int32_t *__restrict__ off, int n,
                                                                      float *__restrict__ dist)
{
    __shared__ int8_t s_q[9], s_k[25];
    if (threadIdx.x < 9) s_q[threadIdx.x] = c_queen_dir[threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 57) s_k[threadIdx.x - 32] = c_knight_type[threadIdx.x - 32];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * POL_WARPS + warp;
    if (b >= n) return;
    float4 *o4 = reinterpret_cast<float4 *>(dist + (size_t)b * SC_N_POLICY);
    for (int i = lane; i < SC_N_POLICY / 4; i += 32) o4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int beg = off[b], end = off[b + 1];
    uint32_t sum = 0;
    for (int k = beg + lane; k < end; k += 32) sum += counts[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float denom = (float)sum + 1e-5f;
    const int turn = pos[b].meta[0];
    __syncwarp();
    for (int k = beg + lane; k < end; k += 32) {
        int idx = move_index_dev(moves[k], turn, s_q, s_k);
        if (idx >= 0) dist[(size_t)b * SC_N_POLICY + idx] = __fdiv_rn((float)counts[k], denom);
    }
}

int launch_dist_scatter(const sc_position *d_pos, const sc_move *d_moves, const uint32_t *d_counts,
                        const int32_t *d_off, int n, float *d_dist, cudaStream_t st)
{
    if (n <= 0) return SC_OK;
    dist_scatter_kernel<<<(n + POL_WARPS - 1) / POL_WARPS, POL_WARPS * 32, 0, st>>>(d_pos, d_moves, d_counts, d_off, n,
                                                                                     d_dist);
    SCB_CUDA(cudaGetLastError());
    return SC_OK;
}

int launch_move_index(const sc_position *d_pos, const sc_move *d_moves, const int32_t *d_off, int n,
                      int32_t *d_index, cudaStream_t st)
{
    if (n <= 0) return SC_OK;
    move_index_kernel<<<(n + POL_WARPS - 1) / POL_WARPS, POL_WARPS * 32, 0, st>>>(d_pos, d_moves, d_off, n, d_index);
    SCB_CUDA(cudaGetLastError());
    return SC_OK;
}

int launch_policy_gather(const float *logits, const sc_position *d_pos, const sc_move *d_moves,
                         const int32_t *d_off, const int32_t *d_cnt, int n, float *d_priors, cudaStream_t st)
{
    if (n <= 0) return SC_OK;
    policy_gather_kernel<<<(n + POL_WARPS - 1) / POL_WARPS, POL_WARPS * 32, 0, st>>>(logits, d_pos, d_moves, d_off,
                                                                                      d_cnt, n, d_priors);
    SCB_CUDA(cudaGetLastError());
    return SC_OK;
}

int launch_policy_logp_full(const float *logits, int n, float *d_logp, cudaStream_t st)
{
    if (n <= 0) return SC_OK;
    policy_logp_full_kernel<<<(n + POL_WARPS - 1) / POL_WARPS, POL_WARPS * 32, 0, st>>>(logits, n, d_logp);
    SCB_CUDA(cudaGetLastError());
    return SC_OK;
}

}  // namespace scb
