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

#define SC_MODE_FP32 0 /* parity mode (1e-4 gate): fp32 arithmetic; the 256-wide convolutions run on the tensor
                          cores as bf16x3 split operands (6 bf16 products per fp32 product) with fp32 accumulate  */
#define SC_MODE_BF16 1 /* throughput mode: tcgen05 bf16 operands, fp32 accumulate      */
#define SC_MODE_FP32_FFMA 2 /* FP32 on the CUDA cores only (FFMA): the referee of the parity mode                */

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
/* CUDA device ordinal the engine lives on and its SM count (one engine per worker thread / per GPU,
 * the in-process form of scripts/run_batch:21's one process per job) */
int sc_device_info(const sc_engine *e, int *device, int *num_sms);

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

/* asynchronous form of sc_eval for double-buffered callers: submit returns as soon as the work is
 * queued on `stream` (host buffers must be pinned and stay untouched until the wait returns);
 * moves and priors are STRIDED: leaf i owns moves[i*SC_MAX_MOVES .. +move_cnt[i]) and the same
 * range of priors_out.  At most SC_MAX_INFLIGHT tickets may be outstanding. */
#define SC_MAX_INFLIGHT 4
int sc_eval_submit(sc_engine *e, int n, const sc_position *pos, const sc_move *moves_strided,
                   const int32_t *move_cnt, float *priors_out_strided, float *value_out, void *stream,
                   int *ticket);
int sc_eval_wait(sc_engine *e, int ticket);

/* -------- bit-exact gates ------------------------------------------------------------------ */
/* `_encode` (src/chess.rs:845-877): planes_out int8 [n][8][8][112] (rank, file, channel), the
 * `Array3<i8>` the reference builds; meta_out int32 [n][7].  Host pointers. */
int sc_encode_only(sc_engine *e, int n, const sc_position *pos, int8_t *planes_out, int32_t *meta_out);
/* `Move::rotate` + `Move::encode` (src/chess.rs:533-550, queenmoves.rs, knightmoves.rs,
 * underpromotions.rs): index_out CSR by move_off; turn comes from pos[i].meta[0]. */
int sc_move_index_only(sc_engine *e, int n, const sc_position *pos, const sc_move *moves,
                       const int32_t *move_off, int32_t *index_out);

/* -------- training-data batch encoder: `chess_encode_steps` (src/lib.rs:47-128), the call
 * py/dataset.py:47-87 makes once per trace.  The game is replayed from the start position with the
 * native rules; for every ply i the outputs are what the reference returns for that step:
 *   planes_out  int8 [n][8][8][112]   `history.view(...)` (identical with and without apply_mirror)
 *   meta_out    int32 [n][7]          `step.encode_meta()`; with apply_mirror it is the ROTATED board's
 *                                     meta (turn flipped, fullmove+1 on White's plies, castling pairs swapped)
 *   dist_out    float [n][4672]       visit counts / (sum + 1e-5) scattered at the move indices
 *   index_out   int32 CSR             move indices of the LEGAL moves in generation order; index_off[n+1]
 * `played[i]` is the move made at ply i, children (CSR by child_off) are the (move, visit count) pairs of
 * the trace row.  Like the reference (which panics) the call fails with SC_E_INVAL if the children of a
 * ply are not exactly its legal moves or the played move is not among them. n <= max_batch. */
int sc_encode_steps(sc_engine *e, int n, const sc_move *played, const sc_move *child_moves,
                    const uint32_t *child_counts, const int32_t *child_off, int apply_mirror, int8_t *planes_out,
                    int32_t *meta_out, float *dist_out, int32_t *index_out, int32_t *index_off);

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
/* level 2: average device time (ms) of the timed tower-convolution launches of the most recent call and
 * how many there were (the roofline kernel of bench.py): one launch of the whole-tower kernel, or the 38
 * per-layer 3x3 256->256 launches with SCB200_TOWER=0; sc_timed_flops_per_leaf = algorithmic FLOPs per
 * leaf those launches cover together */
int sc_kernel_timing(sc_engine *e, float *conv3x3_avg_ms, int *n_launches);
double sc_timed_flops_per_leaf(const sc_engine *e);

/* ===================== batched self-play driver (host side, C++) ================================
 * Replaces the single-tree loop of the `selfplay` binary (src/main.rs:153-238) and `mcts::mcts` /
 * `mcts::step` (src/mcts.rs:237-328) for thousands of concurrent games: every search tree keeps the
 * reference's sequential semantics (one rollout at a time per tree: select by PUCT, expand all legal
 * moves, back up the white-perspective value), and the leaves of all trees are evaluated as one
 * device batch per step.  Priors are stored on the children at expansion, which removes the
 * reference's re-evaluation of every node on each descent (src/mcts.rs:149-153) without changing
 * any result.  Field names follow the reference's CLI (src/main.rs:25-60). */
typedef struct sc_selfplay_config {
    int32_t n_trees;            /* games in flight (BASELINE configs[2]: 2048) */
    int32_t rollout_num;        /* --rollout-num; 0 together with rollout_factor 0 = the reference's default of 300 */
    int32_t num_steps;          /* --num-steps: maximum plies per game */
    float cpuct;                /* --cpuct */
    float epsilon;              /* --epsilon: Dirichlet(0.3) mix at the root (src/mcts.rs:171-184) */
    int32_t with_noise;         /* main.rs passes true; parity runs pass 0 */
    int32_t temperature_switch; /* --temperature-switch: plies played at temperature 1 */
    float temperature;          /* --temperature after the switch (0 = first max-N child) */
    uint64_t seed;              /* per-tree RNG streams derive from this */
    int32_t n_threads;          /* host worker threads walking the trees */
    int32_t evaluator;          /* 0 = the engine (GPU); 1 = position-hash stand-in (test hook, no engine) */
    int32_t pipeline_groups;    /* 1 or 2: trees are split in groups that alternate between host and device */
    int32_t keep_traces;        /* keep the traces of finished games in memory (sc_selfplay_trace_json) */
    int32_t leaves_per_tree;    /* 0/1: one leaf per tree per batch, the reference's exact sequential search.
                                   K > 1: up to K leaves per tree per batch; in-flight paths carry a virtual loss
                                   (one visit lost by the mover).  Not visit-count identical to the reference.
                                   -1: 1 while the trees fill the batch, then trees-of-the-group / trees-still-playing
                                   (<= 16) so that the tail of a finite run keeps the device busy. */
    float rollout_factor;       /* --rollout-factor (src/main.rs:175-180): > 0 (then rollout_num must be 0): the rollouts
                                   of a move are min(300, (int)(legal moves at the root * rollout_factor)) */
} sc_selfplay_config;

typedef struct sc_selfplay_stats {
    int64_t leaf_evals;     /* network-evaluated leaves */
    int64_t terminal_evals; /* rollouts that ended in a position without legal moves (no network call) */
    int64_t rollouts;
    int64_t moves;          /* plies played (search results consumed) */
    int64_t games_finished;
    int64_t white_wins, black_wins, draws, unfinished; /* unfinished = hit num_steps without an outcome */
    int64_t batches;
    double seconds;         /* wall time of sc_selfplay_run */
    double wait_seconds;    /* of which: host waiting for the device */
    int64_t games_dropped;  /* games abandoned because the network returned a non-finite prior / value for one of
                               their leaves (the reference only prints a warning, src/backends/torch.rs:129-135);
                               the slot starts the run's next game, the other games are not affected */
} sc_selfplay_stats;

typedef struct sc_selfplay sc_selfplay;

int sc_selfplay_create(sc_engine *e, const sc_selfplay_config *cfg, sc_selfplay **out);
/* plays until `max_games` games have finished (each tree slot starts a new game when one ends), or
 * until `max_moves` plies were played in total (<= 0: no limit), or `max_seconds` elapsed (<= 0: none).
 * ONE run per driver object: a second call returns SC_E_STATE (create a new driver instead). */
int sc_selfplay_run(sc_selfplay *sp, int64_t max_games, int64_t max_moves, double max_seconds,
                    sc_selfplay_stats *stats);
/* In-process multi-GPU: runs n drivers (each created on its own engine, normally one engine per GPU) on n host
 * threads of THIS process and returns when all have finished -- the in-process form of scripts/run_batch:21
 * (the reference scales by independent processes; a Rust host has no torchrun).  Every driver gets the same
 * limits; stats[i] belongs to sps[i].  Games are sharded by the caller (seeds / max_games per driver); there
 * is no exchange between the drivers.  Returns the first non-zero status. */
int sc_selfplay_run_many(sc_selfplay **sps, int n, int64_t max_games, int64_t max_moves, double max_seconds,
                         sc_selfplay_stats *stats);
/* trace of the k-th finished game in the format of src/trace.rs:23-32
 * ({"steps": [[uci, q, [[uci, n, q, uct], ...]], ...], "outcome": {...} | null}); returns the number
 * of bytes needed (including the NUL); copies at most `cap`. */
int64_t sc_selfplay_trace_json(sc_selfplay *sp, int64_t k, char *buf, int64_t cap);
/* which game the k-th finished trace is: the game's start order within the run (0-based; the games sitting in the tree
 * slots at the start are 0 .. n_trees-1, every later game takes the next number when its slot frees up); -1 if k is out
 * of range.  Lets a caller name trace files by game instead of by finishing order (scripts/run_batch: trace{k}.json). */
int64_t sc_selfplay_trace_game(sc_selfplay *sp, int64_t k);
int sc_selfplay_destroy(sc_selfplay *sp);

/* One self-play game through the one-leaf-at-a-time interface of the reference: the C++ mirror of
 * `trait Game` / `mcts::mcts` / `mcts::step` (src/game.rs:3-21, src/mcts.rs:132-328; csrc/host/game.hpp)
 * and the move loop of src/main.rs:153-238, with `predict` = sc_eval of one leaf (predict runs at every
 * level of every descent, as in the reference).  Uses rollout_num, num_steps, cpuct, epsilon, with_noise,
 * temperature_switch, temperature, seed, evaluator of `cfg`.  Writes the trace (format above) to `buf`;
 * returns the bytes needed including the NUL, or -1 (sc_last_error). */
int64_t sc_game_selfplay(sc_engine *e, const sc_selfplay_config *cfg, char *buf, int64_t cap);

/* ---- arena: two networks play each other (the `play` binary with --black-type nn, src/play.rs:241-343,
 * and scripts/leader-board).  Both players share one configuration, as the reference's CLI does
 * (src/play.rs:43-81): rollout_num = --rollout, cpuct, temperature, temperature_switch; noise is off
 * (play.rs:250), temperature 0 picks uniformly among the most visited children (play.rs:269-278), a game
 * ends when outcome(claim_draw=true) is set after a ply or after num_steps (200) plies.  `white` moves on
 * even plies.  Within a pipeline group every game is at a ply of the same parity (a slot whose game ended
 * starts the run's next game at the next even group ply), so each device batch belongs to one network and
 * the batches stay full.  Colour-swapped games = a second arena with the engines exchanged, as
 * scripts/leader-board:49-54 does. */
int sc_arena_create(sc_engine *white, sc_engine *black, const sc_selfplay_config *cfg, sc_selfplay **out);

/* Synthetic workload: seeded uniform random play from the start position with the driver's
 * native rules, restart on mate/stalemate or at `max_ply`; every non-terminal position met is emitted
 * with its true history (packed leaf, node depth = ply) and its legal moves (CSR).  Fills exactly n. */
int sc_random_positions(int n, uint64_t seed, int max_ply, sc_position *pos_out, sc_move *moves_out,
                        int32_t *move_off /* n+1 */, int max_moves_total);

/* Test hook (bf16 engines): encodes n leaves and runs only the first n_layers (0 = all 41: stem, 19 x (conv1, conv2),
 * the two 256-wide head 1x1 convolutions) of the convolution tower with the throughput kernel (which = 0) or the
 * small-batch latency kernel (which = 1); copies the three activation buffers (bf16 bit patterns, [n][64][256]:
 * x = block input / output, t = conv1 output / policy head input, y = value head input) to the host.  Lets the tests
 * compare the two kernels layer by layer. */
int sc_debug_tower(sc_engine *e, int n, const sc_position *pos, int n_layers, int which, uint16_t *x_out,
                   uint16_t *t_out, uint16_t *y_out);

/* Test hook: one Dirichlet(alpha) sample of size n from the driver's root-noise sampler (src/mcts.rs:123-130
 * uses rand_distr::Dirichlet(0.3)); lets the tests check its moments. */
int sc_test_dirichlet(uint64_t seed, float alpha, int n, float *out);

/* perft (number of leaf nodes of the legal-move tree of the given depth <= 7) of the driver's native rules from a
 * FEN (NULL = start position): pins the move generator to the published perft tables. */
int sc_rules_perft(const char *fen, int depth, uint64_t *nodes);

/* Host rules probe: replays `n_history` moves from the start position with the driver's native rules
 * (the replacement of the python-chess calls at src/chess.rs:665-788) and reports what the reference
 * would see there: legal moves in python-chess generation order, the packed leaf `_encode` would be
 * given (node depth = ply), and outcome(claim_draw=True) (termination code of src/chess.rs:87-105 or
 * 0, winner 1/0/-1).  Returns SC_E_INVAL if a history move is not legal. */
int sc_rules_probe(const sc_move *history, int n_history, sc_move *legal_out, int *n_legal, sc_position *packed_out,
                   int *termination, int *winner);
/* the same from an arbitrary start position (`chess.Board(fen)`; NULL = the initial position) */
int sc_rules_probe_fen(const char *fen, const sc_move *history, int n_history, sc_move *legal_out, int *n_legal,
                       sc_position *packed_out, int *termination, int *winner);

#ifdef __cplusplus
}
#endif
#endif /* SC_B200_H */
