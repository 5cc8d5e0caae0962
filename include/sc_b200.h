/*
 * sc_b200.h -- C ABI of the B200 leaf-evaluation backend for smart-chess-rust.
 *
 * This is what a new `src/backends/b200.rs` would bind.  Every entry
 * point names the reference interface it replaces (paths relative to the reference repo).
 * Plain pointers and sizes only; no torch types.  All functions return 0 on success or a
 * negative SC_E_* code; `sc_last_error()` gives the message.  The reference has no error
 * channel (every failure is unwrap()/panic!, e.g. src/backends/torch.rs:30-31), so the Rust
 * shim panics on a non-zero status.
 *
 * Threading: a handle is single-threaded like the reference backends (src/game.rs takes
 * &self but every backend is !Sync in practice: RefCell at src/backends/onnx.rs:9-11).
 * Use one handle per worker thread / per GPU.
 */
#ifndef SC_B200_H
#define SC_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_OK 0
#define SC_E_INVAL (-1)   /* bad argument                         */
#define SC_E_CUDA (-2)    /* CUDA runtime / driver error          */
#define SC_E_IO (-3)      /* weight blob unreadable / malformed   */
#define SC_E_NOGPU (-4)   /* no sm_100 device: there is NO CPU fallback */
#define SC_E_STATE (-5)

#define SC_MODE_FP32 0 /* parity mode: FP32 FFMA tower (1e-4 gate)                    */
#define SC_MODE_BF16 1 /* throughput mode: tcgen05 bf16 operands, fp32 accumulate      */

#define SC_LOOKBACK 8       /* src/chess.rs:23 */
#define SC_N_PLANES 112     /* 8 x 14, py/module.py:118-121 */
#define SC_N_META 7         /* src/chess.rs:652-662 */
#define SC_N_POLICY 4672    /* 8 x 8 x 73 */
#define SC_MAX_MOVES 256

/*
 * One leaf as `_encode` (src/chess.rs:845-877) sees it: up to 8 boards of history, newest
 * first, each as python-chess bitboards (a1 = bit 0, h8 = bit 63, white-oriented, NOT
 * rotated -- the rotation of `Board::rotate`, src/chess.rs:594-621, happens on the device),
 * plus the `encode_meta` vector of the current board (src/chess.rs:652-662).
 *
 *   slot[t][0..5] = pawns, knights, bishops, rooks, queens, kings (both colours)
 *   slot[t][6]    = white occupancy
 *   slot[t][7]    = bit 0: is_repetition(2), bit 1: is_repetition(3)   (src/chess.rs:375-380)
 *   meta          = [turn(1=white), fullmove_number, K-castle(stm), Q-castle(stm),
 *                    K-castle(opp), Q-castle(opp), halfmove_clock]
 *   n_hist        = number of valid slots, 1..8 (history stops at the tree root)
 */
typedef struct sc_position {
    uint64_t slot[SC_LOOKBACK][8];
    int32_t meta[SC_N_META];
    int32_t n_hist;
} sc_position; /* 544 bytes */

/* `chess::Move` (src/chess.rs:36-41): squares are rank*8+file; promo 0 or 2..5 (N,B,R,Q). */
typedef struct sc_move {
    uint8_t from, to, promo, pad;
} sc_move;

typedef struct sc_engine sc_engine;

/* -------- lifecycle: replaces tch::CModule::load_on_device (src/main.rs:88-94,
 *          src/play.rs:356-365) / ort Session (src/main.rs:107-113) ------------------------- */
int sc_create(const char *weights_blob_path, int device, int mode, int max_batch, sc_engine **out);
int sc_destroy(sc_engine *e);
const char *sc_last_error(void);
/* number of residual blocks found in the blob; max batch; mode */
int sc_info(const sc_engine *e, int *n_res_blocks, int *max_batch, int *mode);

/* -------- the hot path: replaces chess_tch_predict (src/backends/torch.rs:89-146) and
 *          ChessOnnx::predict (src/backends/onnx.rs:13-56) for n non-terminal leaves ---------
 * pos[n], moves = CSR over leaves (move_off[n+1]), priors_out CSR by move_off, value_out[n]
 * (White's perspective, py/module.py:147-149).  Host pointers; pinned memory makes the copies
 * asynchronous.  `stream` is a cudaStream_t (NULL = the engine's own stream); the call
 * returns after the results are in the host buffers. */
int sc_eval(sc_engine *e, int n, const sc_position *pos, const sc_move *moves, const int32_t *move_off,
            float *priors_out, float *value_out, void *stream);

/* same computation, inputs/outputs already resident in device memory, asynchronous on
 * `stream` (no host copies, no sync): the leg bench.py reports as `value`. */
int sc_eval_device(sc_engine *e, int n, const void *d_pos, const void *d_moves, const void *d_move_off,
                   int n_moves_total, void *d_priors_out, void *d_value_out, void *stream);

/* -------- bit-exact gates ------------------------------------------------------------------ */
/* `_encode` (src/chess.rs:845-877): planes_out int8 [n][8][8][112] (rank, file, channel), the
 * `Array3<i8>` the reference builds; meta_out int32 [n][7].  Host pointers. */
int sc_encode_only(sc_engine *e, int n, const sc_position *pos, int8_t *planes_out, int32_t *meta_out);
/* `Move::rotate` + `Move::encode` (src/chess.rs:533-550, queenmoves.rs, knightmoves.rs,
 * underpromotions.rs): index_out CSR by move_off; turn comes from pos[i].meta[0]. */
int sc_move_index_only(sc_engine *e, int n, const sc_position *pos, const sc_move *moves,
                       const int32_t *move_off, int32_t *index_out);

/* -------- tolerance gate: ChessModule.forward (py/module.py:135-154) -------------------------
 * planes float [n][112][8][8] (NCHW, what the backends feed), meta float [n][7];
 * logp_out float [n][4672] in the reference's flatten order, value_out float [n]. Host pointers. */
int sc_forward_only(sc_engine *e, int n, const float *planes, const float *meta, float *logp_out,
                    float *value_out);

/* -------- accounting ------------------------------------------------------------------------ */
/* kernels launched by this engine since creation (bench.py's gpu_launches) */
int64_t sc_launch_count(const sc_engine *e);
/* average device time in ms of the conv tower kernels / all kernels of the most recent
 * sc_eval* call, measured with CUDA events on the launching stream (profiling aid) */
int sc_last_timing(sc_engine *e, float *tower_ms, float *total_ms);
/* per-call event timing: 0 off, 1 = tower/total of a call (two event records + one sync per
 * call), 2 = additionally one event pair around every 3x3 tower convolution launch */
int sc_set_timing(sc_engine *e, int enabled);
/* level 2: average device time (ms) of the 3x3 256->256 tower convolution launches of the
 * most recent call and how many there were (the roofline kernel of bench.py) */
int sc_kernel_timing(sc_engine *e, float *conv3x3_avg_ms, int *n_launches);

#ifdef __cplusplus
}
#endif
#endif /* SC_B200_H */
