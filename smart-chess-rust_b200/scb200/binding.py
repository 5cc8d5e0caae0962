"""ctypes binding of include/sc_b200.h."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

SC_MODE_FP32 = 0
SC_MODE_BF16 = 1
SC_MODE_FP32_FFMA = 2
SC_N_PLANES = 112
SC_N_META = 7
SC_N_POLICY = 4672

# struct sc_position { uint64_t slot[8][8]; int32_t meta[7]; int32_t n_hist; }  (544 bytes)
POSITION_DTYPE = np.dtype([("slot", np.uint64, (8, 8)), ("meta", np.int32, (7,)), ("n_hist", np.int32)])
# struct sc_move { uint8_t from, to, promo, pad; }
MOVE_DTYPE = np.dtype([("from", np.uint8), ("to", np.uint8), ("promo", np.uint8), ("pad", np.uint8)])
assert POSITION_DTYPE.itemsize == 544 and MOVE_DTYPE.itemsize == 4

DECLARED_SYMBOLS = [
    "sc_create", "sc_destroy", "sc_last_error", "sc_info", "sc_eval", "sc_eval_device", "sc_encode_only",
    "sc_move_index_only", "sc_forward_only", "sc_launch_count", "sc_last_timing", "sc_set_timing", "sc_kernel_timing",
    "sc_eval_submit", "sc_eval_wait", "sc_selfplay_create", "sc_selfplay_run", "sc_selfplay_trace_json",
    "sc_selfplay_destroy", "sc_rules_probe", "sc_arena_create", "sc_encode_steps", "sc_timed_flops_per_leaf",
    "sc_random_positions", "sc_test_dirichlet", "sc_game_selfplay", "sc_rules_perft", "sc_device_info",
    "sc_selfplay_run_many", "sc_debug_tower", "sc_rules_probe_fen", "sc_selfplay_trace_game",
]


class SelfPlayConfig(C.Structure):
    """struct sc_selfplay_config (field names follow the reference CLI, src/main.rs:25-60)"""
    _fields_ = [("n_trees", C.c_int32), ("rollout_num", C.c_int32), ("num_steps", C.c_int32), ("cpuct", C.c_float),
                ("epsilon", C.c_float), ("with_noise", C.c_int32), ("temperature_switch", C.c_int32),
                ("temperature", C.c_float), ("seed", C.c_uint64), ("n_threads", C.c_int32), ("evaluator", C.c_int32),
                ("pipeline_groups", C.c_int32), ("keep_traces", C.c_int32),
                ("leaves_per_tree", C.c_int32), ("rollout_factor", C.c_float)]


class SelfPlayStats(C.Structure):
    _fields_ = [("leaf_evals", C.c_int64), ("terminal_evals", C.c_int64), ("rollouts", C.c_int64), ("moves", C.c_int64),
                ("games_finished", C.c_int64), ("white_wins", C.c_int64), ("black_wins", C.c_int64),
                ("draws", C.c_int64), ("unfinished", C.c_int64), ("batches", C.c_int64), ("seconds", C.c_double),
                ("wait_seconds", C.c_double), ("games_dropped", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_LIB = None


class SCError(RuntimeError):
    pass


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libscb200.so")


def load_library():
    """Loads libscb200.so; raises loudly if the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise SCError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for this backend)")
        L = C.CDLL(p)
        L.sc_last_error.restype = C.c_char_p
        L.sc_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.sc_destroy.argtypes = [C.c_void_p]
        L.sc_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.sc_device_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.sc_selfplay_run_many.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int64, C.c_double,
                                           C.POINTER(SelfPlayStats)]
        L.sc_debug_tower.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sc_eval.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        L.sc_eval_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
        L.sc_encode_only.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sc_move_index_only.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sc_forward_only.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sc_encode_steps.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 5
        L.sc_timed_flops_per_leaf.restype = C.c_double
        L.sc_timed_flops_per_leaf.argtypes = [C.c_void_p]
        L.sc_launch_count.restype = C.c_int64
        L.sc_launch_count.argtypes = [C.c_void_p]
        L.sc_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.sc_set_timing.argtypes = [C.c_void_p, C.c_int]
        L.sc_kernel_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.sc_eval_submit.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.POINTER(C.c_int)]
        L.sc_eval_wait.argtypes = [C.c_void_p, C.c_int]
        L.sc_selfplay_create.argtypes = [C.c_void_p, C.POINTER(SelfPlayConfig), C.POINTER(C.c_void_p)]
        L.sc_selfplay_run.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.POINTER(SelfPlayStats)]
        L.sc_game_selfplay.restype = C.c_int64
        L.sc_game_selfplay.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int64]
        L.sc_selfplay_trace_json.restype = C.c_int64
        L.sc_selfplay_trace_json.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int64]
        L.sc_selfplay_trace_game.restype = C.c_int64
        L.sc_selfplay_trace_game.argtypes = [C.c_void_p, C.c_int64]
        L.sc_selfplay_destroy.argtypes = [C.c_void_p]
        L.sc_arena_create.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(SelfPlayConfig), C.POINTER(C.c_void_p)]
        L.sc_random_positions.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.sc_test_dirichlet.argtypes = [C.c_uint64, C.c_float, C.c_int, C.c_void_p]
        L.sc_rules_perft.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_uint64)]
        L.sc_rules_probe.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.c_void_p,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.sc_rules_probe_fen.argtypes = [C.c_char_p] + L.sc_rules_probe.argtypes
        _LIB = L
    return _LIB


def _check(rc: int, what: str):
    if rc != 0:
        raise SCError(f"{what} failed ({rc}): {load_library().sc_last_error().decode()}")


def pack_positions(slots, metas, n_hists) -> np.ndarray:
    """Builds an sc_position array from per-leaf (slot[8,8] u64, meta[7] i32, n_hist)."""
    n = len(metas)
    out = np.zeros(n, dtype=POSITION_DTYPE)
    for i in range(n):
        out["slot"][i] = slots[i]
        out["meta"][i] = metas[i]
        out["n_hist"][i] = n_hists[i]
    return out


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):  # torch tensor (device or pinned host)
        return a.data_ptr()
    return a.ctypes.data


class Engine:
    """`ChessTS` / `ChessOnnx` counterpart (src/backends/torch.rs:14-17, onnx.rs:8-11): owns the
    model for the lifetime of the process and evaluates batches of leaves."""

    def __init__(self, blob_path: str, device: int = 0, mode: int = SC_MODE_BF16, max_batch: int = 2048):
        L = load_library()
        h = C.c_void_p()
        _check(L.sc_create(blob_path.encode(), device, mode, max_batch, C.byref(h)), "sc_create")
        self._h = h
        self.mode = mode
        self.max_batch = max_batch
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            load_library().sc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def info(self):
        a, b, c, d, s = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(load_library().sc_info(self._h, C.byref(a), C.byref(b), C.byref(c)), "sc_info")
        _check(load_library().sc_device_info(self._h, C.byref(d), C.byref(s)), "sc_device_info")
        return {"n_res_blocks": a.value, "max_batch": b.value, "mode": c.value, "device": d.value, "num_sms": s.value}

    def eval(self, positions: np.ndarray, moves: np.ndarray, move_off: np.ndarray, priors_out=None, value_out=None,
             stream: int = 0):
        """`predict` for n leaves: returns (priors CSR float32, values float32[n])."""
        n = len(positions)
        # numpy inputs are coerced to the ABI's element types (an int64 move_off or a strided view would otherwise be
        # reinterpreted) unless they already have them; torch tensors (pinned host buffers of bench.py) are taken as they are
        def _as(a, dt):
            if hasattr(a, "data_ptr") or (a.dtype == dt and a.flags.c_contiguous):
                return a
            return np.ascontiguousarray(a, dtype=dt)

        positions, moves, move_off = _as(positions, POSITION_DTYPE), _as(moves, MOVE_DTYPE), _as(move_off, np.int32)
        total = int(move_off[n]) if n else 0
        if priors_out is None:
            priors_out = np.zeros(max(total, 1), dtype=np.float32)
        if value_out is None:
            value_out = np.zeros(max(n, 1), dtype=np.float32)
        _check(load_library().sc_eval(self._h, n, _ptr(positions), _ptr(moves), _ptr(move_off), _ptr(priors_out),
                                      _ptr(value_out), stream or None), "sc_eval")
        return priors_out[:total], value_out[:n]

    def eval_device(self, n, d_pos, d_moves, d_off, n_moves_total, d_priors, d_value, stream: int = 0):
        _check(load_library().sc_eval_device(self._h, n, _ptr(d_pos), _ptr(d_moves), _ptr(d_off), n_moves_total,
                                             _ptr(d_priors), _ptr(d_value), stream or None), "sc_eval_device")

    def submit(self, positions, moves_strided, move_cnt, priors_out_strided, value_out, stream: int = 0) -> int:
        """Asynchronous `predict` (sc_eval_submit): moves / priors strided by 256 per leaf, buffers must stay alive
        (and should be pinned) until `wait(ticket)` returns.  At most SC_MAX_INFLIGHT = 4 tickets in flight."""
        t = C.c_int(-1)
        _check(load_library().sc_eval_submit(self._h, len(positions), _ptr(positions), _ptr(moves_strided), _ptr(move_cnt),
                                             _ptr(priors_out_strided), _ptr(value_out), stream or None, C.byref(t)),
               "sc_eval_submit")
        return t.value

    def wait(self, ticket: int):
        _check(load_library().sc_eval_wait(self._h, ticket), "sc_eval_wait")

    def encode_only(self, positions: np.ndarray):
        n = len(positions)
        positions = np.ascontiguousarray(positions, dtype=POSITION_DTYPE)
        planes = np.zeros((n, 8, 8, SC_N_PLANES), dtype=np.int8)
        meta = np.zeros((n, SC_N_META), dtype=np.int32)
        _check(load_library().sc_encode_only(self._h, n, _ptr(positions), _ptr(planes), _ptr(meta)), "sc_encode_only")
        return planes, meta

    def move_index_only(self, positions: np.ndarray, moves: np.ndarray, move_off: np.ndarray):
        n = len(positions)
        positions = np.ascontiguousarray(positions, dtype=POSITION_DTYPE)
        moves = np.ascontiguousarray(moves, dtype=MOVE_DTYPE)
        move_off = np.ascontiguousarray(move_off, dtype=np.int32)
        out = np.full(int(move_off[n]) if n else 0, -7, dtype=np.int32)
        _check(load_library().sc_move_index_only(self._h, n, _ptr(positions), _ptr(moves), _ptr(move_off), _ptr(out)),
               "sc_move_index_only")
        return out

    def forward_only(self, planes: np.ndarray, meta: np.ndarray):
        """planes float32 [n,112,8,8], meta float32 [n,7] -> (logp [n,4672], value [n])."""
        planes = np.ascontiguousarray(planes, dtype=np.float32)
        meta = np.ascontiguousarray(meta, dtype=np.float32)
        n = planes.shape[0]
        logp = np.zeros((n, SC_N_POLICY), dtype=np.float32)
        value = np.zeros(n, dtype=np.float32)
        _check(load_library().sc_forward_only(self._h, n, _ptr(planes), _ptr(meta), _ptr(logp), _ptr(value)),
               "sc_forward_only")
        return logp, value

    def encode_steps(self, steps, apply_mirror: bool = False):
        """`libsmartchess.chess_encode_steps(steps, apply_mirror)` (src/lib.rs:47-128): steps is the list
        [((from, to, promo), [((from, to, promo), visit_count), ...]), ...] a trace yields (py/dataset.py:70-77).
        Returns a list of (planes int8[8,8,112], meta int32[7], dist float32[4672], [move indices])."""
        n = len(steps)
        played = np.zeros(n, dtype=MOVE_DTYPE)
        off = np.zeros(n + 1, dtype=np.int32)
        cm, cc = [], []
        for i, (mv, children) in enumerate(steps):
            played[i] = (int(mv[0]), int(mv[1]), int(mv[2]), 0)
            for m, c in children:
                cm.append((int(m[0]), int(m[1]), int(m[2]), 0))
                cc.append(int(c))
            off[i + 1] = len(cm)
        cm = np.array(cm, dtype=MOVE_DTYPE) if cm else np.zeros(1, dtype=MOVE_DTYPE)
        cc = np.array(cc, dtype=np.uint32) if cc else np.zeros(1, dtype=np.uint32)
        planes = np.zeros((n, 8, 8, SC_N_PLANES), dtype=np.int8)
        meta = np.zeros((n, SC_N_META), dtype=np.int32)
        dist = np.zeros((n, SC_N_POLICY), dtype=np.float32)
        index = np.zeros(max(1, n * 256), dtype=np.int32)
        ioff = np.zeros(n + 1, dtype=np.int32)
        _check(load_library().sc_encode_steps(self._h, n, _ptr(played), _ptr(cm), _ptr(cc), _ptr(off), int(apply_mirror),
                                              _ptr(planes), _ptr(meta), _ptr(dist), _ptr(index), _ptr(ioff)),
               "sc_encode_steps")
        return [(planes[i], meta[i], dist[i], index[ioff[i]:ioff[i + 1]].tolist()) for i in range(n)]

    def debug_tower(self, positions, n_layers: int, which: int):
        """test hook sc_debug_tower: (x, t, y) activation buffers as uint16 [n, 64, 256] after n_layers layers"""
        n = len(positions)
        positions = np.ascontiguousarray(positions, dtype=POSITION_DTYPE)
        out = [np.zeros((n, 64, 256), dtype=np.uint16) for _ in range(3)]
        _check(load_library().sc_debug_tower(self._h, n, _ptr(positions), n_layers, which, _ptr(out[0]), _ptr(out[1]),
                                             _ptr(out[2])), "sc_debug_tower")
        return out

    def launch_count(self) -> int:
        return int(load_library().sc_launch_count(self._h))

    def set_timing(self, level: int):
        _check(load_library().sc_set_timing(self._h, int(level)), "sc_set_timing")

    def kernel_timing(self):
        """(average ms of a 3x3 tower conv launch in the last call, number of launches)"""
        a, n = C.c_float(), C.c_int()
        _check(load_library().sc_kernel_timing(self._h, C.byref(a), C.byref(n)), "sc_kernel_timing")
        return a.value, n.value

    def timed_flops_per_leaf(self) -> float:
        return float(load_library().sc_timed_flops_per_leaf(self._h))

    def last_timing(self):
        a, b = C.c_float(), C.c_float()
        _check(load_library().sc_last_timing(self._h, C.byref(a), C.byref(b)), "sc_last_timing")
        return a.value, b.value


class SelfPlay:
    """Batched counterpart of the `selfplay` binary (src/main.rs:153-238): n_trees games in flight,
    one leaf per tree per device batch.  evaluator="hash" needs no engine (test hook)."""

    def __init__(self, engine, n_trees=2048, rollout_num=180, num_steps=150, cpuct=2.5, epsilon=0.15,
                 with_noise=True, temperature_switch=4, temperature=0.0, seed=0, n_threads=0, evaluator="engine",
                 pipeline_groups=2, keep_traces=False, black_engine=None, arena=False, leaves_per_tree=1,
                 rollout_factor=0.0):
        import json as _json

        self._json = _json
        L = load_library()
        cfg = SelfPlayConfig(n_trees, rollout_num, num_steps, cpuct, epsilon, int(with_noise), temperature_switch,
                             temperature, seed, n_threads, 0 if evaluator == "engine" else 1, pipeline_groups,
                             int(keep_traces), int(leaves_per_tree), float(rollout_factor))
        h = C.c_void_p()
        self._engine = (engine, black_engine)  # keep alive
        if arena:
            _check(L.sc_arena_create(engine.handle if engine is not None else None,
                                     black_engine.handle if black_engine is not None else None, C.byref(cfg),
                                     C.byref(h)), "sc_arena_create")
        else:
            _check(L.sc_selfplay_create(engine.handle if engine is not None else None, C.byref(cfg), C.byref(h)),
                   "sc_selfplay_create")
        self._h = h

    def run(self, max_games=0, max_moves=0, max_seconds=0.0):
        st = SelfPlayStats()
        _check(load_library().sc_selfplay_run(self._h, max_games, max_moves, max_seconds, C.byref(st)), "sc_selfplay_run")
        return st.as_dict()

    @staticmethod
    def run_many(drivers, max_games=0, max_moves=0, max_seconds=0.0):
        """sc_selfplay_run_many: every driver on its own host thread of this process (one engine per GPU)."""
        n = len(drivers)
        hs = (C.c_void_p * n)(*[d._h for d in drivers])
        st = (SelfPlayStats * n)()
        _check(load_library().sc_selfplay_run_many(hs, n, max_games, max_moves, max_seconds, st), "sc_selfplay_run_many")
        return [s.as_dict() for s in st]

    def trace(self, k: int):
        L = load_library()
        n = L.sc_selfplay_trace_json(self._h, k, None, 0)
        if n < 0:
            return None
        buf = C.create_string_buffer(n)
        L.sc_selfplay_trace_json(self._h, k, buf, n)
        return self._json.loads(buf.value.decode())

    def trace_game(self, k: int) -> int:
        """start order (0-based) of the game the k-th finished trace belongs to; -1 if there is no such trace"""
        return int(load_library().sc_selfplay_trace_game(self._h, k))

    def close(self):
        if getattr(self, "_h", None):
            load_library().sc_selfplay_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def game_selfplay(engine, rollout_num=20, num_steps=150, cpuct=2.5, epsilon=0.15, with_noise=False,
                  temperature_switch=0, temperature=0.0, seed=0, evaluator="engine"):
    """One game of the `selfplay` binary (src/main.rs:153-238) through the C++ mirror of the reference's
    `Game` / `mcts` interface (csrc/host/game.hpp), one leaf per `predict`.  Returns the trace dict."""
    import json

    L = load_library()
    cfg = SelfPlayConfig(1, rollout_num, num_steps, cpuct, epsilon, int(with_noise), temperature_switch, temperature, seed,
                         1, 0 if evaluator == "engine" else 1, 1, 1, 1, 0.0)
    cap = 1 << 24
    buf = C.create_string_buffer(cap)
    n = L.sc_game_selfplay(engine.handle if engine is not None else None, C.byref(cfg), buf, cap)
    if n < 0:
        raise SCError("sc_game_selfplay failed: " + L.sc_last_error().decode())
    return json.loads(buf.value.decode())


def rules_perft(fen, depth: int) -> int:
    """perft of the driver's native rules (fen None = start position)."""
    n = C.c_uint64()
    _check(load_library().sc_rules_perft(fen.encode() if fen else None, depth, C.byref(n)), "sc_rules_perft")
    return int(n.value)


def rules_probe(history, fen=None):
    """history: list of (from, to, promo) [from `fen`, default the initial position] -> (legal moves [n,3] uint8,
    sc_position record, (termination, winner))."""
    h = np.zeros(len(history), dtype=MOVE_DTYPE)
    for i, m in enumerate(history):
        h[i] = (int(m[0]), int(m[1]), int(m[2]), 0)
    legal = np.zeros(256, dtype=MOVE_DTYPE)
    n, term, win = C.c_int(), C.c_int(), C.c_int()
    pos = np.zeros(1, dtype=POSITION_DTYPE)
    _check(load_library().sc_rules_probe_fen(fen.encode() if fen else None, _ptr(h) if len(h) else None, len(h), _ptr(legal),
                                             C.byref(n), _ptr(pos), C.byref(term), C.byref(win)), "sc_rules_probe")
    mv = np.stack([legal["from"][: n.value], legal["to"][: n.value], legal["promo"][: n.value]], axis=1)
    return mv, pos[0], (term.value, win.value)


class Arena(SelfPlay):
    """`play --black-type nn` for n_trees games at once (src/play.rs:241-343): `white` moves on even plies."""

    def __init__(self, white, black, n_trees=256, rollout=100, cpuct=1.5, temperature=0.0, temperature_switch=8,
                 max_plies=200, seed=0, n_threads=0, evaluator="engine", pipeline_groups=2, keep_traces=False,
                 leaves_per_tree=1):
        super().__init__(white, n_trees=n_trees, rollout_num=rollout, num_steps=max_plies, cpuct=cpuct,
                         with_noise=False, temperature_switch=temperature_switch, temperature=temperature, seed=seed,
                         n_threads=n_threads, evaluator=evaluator, pipeline_groups=pipeline_groups,
                         keep_traces=keep_traces, black_engine=black, arena=True, leaves_per_tree=leaves_per_tree)


def elo(total: int, wins: int, losses: int) -> float:
    """scripts/elo.py:17-19: Elo difference from a Total/Win/Loss tally."""
    import math

    s = (wins + (total - wins - losses) / 2) / total
    return 400 * math.log(s / (1 - s), 10)


def random_positions(n: int, seed: int = 1, max_ply: int = 150):
    """Synthetic workload from the driver's native rules: (sc_position[n], sc_move[total], move_off[n+1])."""
    pos = np.zeros(n, dtype=POSITION_DTYPE)
    moves = np.zeros(max(1, n * 64), dtype=MOVE_DTYPE)
    off = np.zeros(n + 1, dtype=np.int32)
    _check(load_library().sc_random_positions(n, seed, max_ply, _ptr(pos), _ptr(moves), _ptr(off), len(moves)),
           "sc_random_positions")
    return pos, moves[: int(off[n])].copy(), off
