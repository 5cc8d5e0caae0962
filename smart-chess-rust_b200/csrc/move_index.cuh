// The 8x8x73 move index (`Move::rotate` + `Move::encode`, reference src/chess.rs:533-550;
// src/queenmoves.rs:3-34, src/knightmoves.rs:7-31, src/underpromotions.rs:6-33) as integer arithmetic
// on (from, to, promo) with two small direction tables.  Shared by the stand-alone kernels (policy.cu)
// and the policy-head epilogue that gathers the legal-move priors (tower_bf16.cu).
#pragma once
#include "common.cuh"

namespace scb {

// queen direction by (sign(d_rank)+1)*3 + (sign(d_file)+1); centre is impossible
static __constant__ int8_t c_queen_dir[9] = {5, 4, 3, 6, -1, 2, 7, 0, 1};
// knight type by (d_rank+2)*5 + (d_file+2)
static __constant__ int8_t c_knight_type[25] = {-1, 4,  -1, 3,  -1,   // d_rank = -2: (-2,-1)=4, (-2,1)=3
                                         5,  -1, -1, -1, 2,    // d_rank = -1: (-1,-2)=5, (-1,2)=2
                                         -1, -1, -1, -1, -1,
                                         6,  -1, -1, -1, 1,    // d_rank = +1: (1,-2)=6, (1,2)=1
                                         -1, 7,  -1, 0,  -1};  // d_rank = +2: (2,-1)=7, (2,1)=0

__device__ __forceinline__ int move_index_dev(sc_move m, int turn, const int8_t *qdir, const int8_t *ktype)
{
    if ((m.from | m.to) & 0xC0) return -1;  // not a square: no index, prior 0 (keeps every later lookup in range)
    int fr = m.from >> 3, ff = m.from & 7, tr = m.to >> 3, tf = m.to & 7;
    if (!turn) {  // Move::rotate for Black to move
        fr = 7 - fr;
        tr = 7 - tr;
    }
    const int d0 = tr - fr, d1 = tf - ff;
    const int a0 = abs(d0), a1 = abs(d1);
    const int base = fr * 584 + ff * 73;
    const bool queen_promo = (m.promo == 0) || (m.promo == 5);
    if ((d0 == 0 || d1 == 0 || a0 == a1) && queen_promo) {
        int s0 = (d0 > 0) - (d0 < 0), s1 = (d1 > 0) - (d1 < 0);
        int dir = qdir[(s0 + 1) * 3 + (s1 + 1)];
        if (dir < 0) return -1;
        return base + dir * 7 + (max(a0, a1) - 1);
    }
    if (a0 <= 2 && a1 <= 2) {
        int k = ktype[(d0 + 2) * 5 + (d1 + 2)];
        if (k >= 0) return base + 56 + k;
    }
    if (m.promo >= 2 && m.promo <= 4 && fr == 6 && tr == 7 && a1 <= 1)
        return base + 64 + (d1 + 1) * 3 + (m.promo - 2);
    return -1;
}

}  // namespace scb
