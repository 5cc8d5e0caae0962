"""ctypes face of oracle/chess_oracle.c (TEST INFRASTRUCTURE -- never on the product path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference leg may import
this module.  See the header of chess_oracle.c for the reference file:line each function
restates and for what is / is not pinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PIECE_SYMBOLS = {1: "p", 2: "n", 3: "b", 4: "r", 5: "q", 6: "k"}
SYMBOL_PIECES = {v: k for k, v in PIECE_SYMBOLS.items()}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "chess_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oc_game_new.restype = C.c_void_p
        L.oc_game_from_fen.restype = C.c_void_p
        L.oc_game_from_fen.argtypes = [C.c_char_p]
        L.oc_game_dup.restype = C.c_void_p
        L.oc_game_dup.argtypes = [C.c_void_p]
        L.oc_game_free.argtypes = [C.c_void_p]
        L.oc_game_push.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oc_game_pop.argtypes = [C.c_void_p]
        L.oc_game_ply.argtypes = [C.c_void_p]
        L.oc_game_turn.argtypes = [C.c_void_p]
        L.oc_game_legal_moves.argtypes = [C.c_void_p, C.c_void_p]
        L.oc_game_is_check.argtypes = [C.c_void_p]
        L.oc_game_piece_at.argtypes = [C.c_void_p, C.c_int]
        L.oc_game_is_repetition.argtypes = [C.c_void_p, C.c_int]
        L.oc_game_encode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oc_game_pack.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oc_game_outcome.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.oc_move_index.argtypes = [C.c_int] * 4
        L.oc_perft.restype = C.c_uint64
        L.oc_perft.argtypes = [C.c_void_p, C.c_int]
        L.oc_post_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.oc_position_hash.restype = C.c_uint64
        L.oc_position_hash.argtypes = [C.c_void_p]
        L.oc_tree_new.restype = C.c_void_p
        L.oc_tree_new.argtypes = [C.c_void_p, C.c_void_p]
        L.oc_tree_new_hash.restype = C.c_void_p
        L.oc_tree_free.argtypes = [C.c_void_p]
        L.oc_tree_game.restype = C.c_void_p
        L.oc_tree_game.argtypes = [C.c_void_p]
        L.oc_tree_search.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.oc_tree_root_children.argtypes = [C.c_void_p] * 5
        L.oc_tree_root_q.restype = C.c_float
        L.oc_tree_root_q.argtypes = [C.c_void_p]
        L.oc_tree_root_n.argtypes = [C.c_void_p]
        L.oc_tree_step_argmax.argtypes = [C.c_void_p]
        L.oc_tree_step_index.argtypes = [C.c_void_p, C.c_int]
        L.oc_tree_n_evals.restype = C.c_long
        L.oc_tree_n_evals.argtypes = [C.c_void_p]
        L.oc_tree_n_predicts.restype = C.c_long
        L.oc_tree_n_predicts.argtypes = [C.c_void_p]
        L.oc_init()
        _LIB = L
    return _LIB


def sq_name(sq: int) -> str:
    return "abcdefgh"[sq & 7] + str((sq >> 3) + 1)


def sq_parse(s: str) -> int:
    return (int(s[1]) - 1) * 8 + "abcdefgh".index(s[0])


def uci(m) -> str:
    f, t, p = int(m[0]), int(m[1]), int(m[2])
    return sq_name(f) + sq_name(t) + (PIECE_SYMBOLS[p] if p else "")


def parse_uci(s: str):
    return (sq_parse(s[0:2]), sq_parse(s[2:4]), SYMBOL_PIECES[s[4]] if len(s) > 4 else 0)


class Game:
    """python-chess `Board` restated (see chess_oracle.c)."""

    def __init__(self, fen: str | None = None, _ptr=None):
        L = lib()
        if _ptr is not None:
            self._p = _ptr
        elif fen is None:
            self._p = L.oc_game_new()
        else:
            self._p = L.oc_game_from_fen(fen.encode())
            if not self._p:
                raise ValueError("bad fen")
        self._own = True

    def __del__(self):
        if getattr(self, "_own", False) and self._p:
            lib().oc_game_free(self._p)
            self._p = None

    def dup(self) -> "Game":
        return Game(_ptr=lib().oc_game_dup(self._p))

    @property
    def ply(self) -> int:
        return lib().oc_game_ply(self._p)

    @property
    def turn(self) -> int:
        return lib().oc_game_turn(self._p)

    def push(self, m):
        if isinstance(m, str):
            m = parse_uci(m)
        if lib().oc_game_push(self._p, int(m[0]), int(m[1]), int(m[2])):
            raise RuntimeError("game too long")

    def pop(self):
        lib().oc_game_pop(self._p)

    def legal_moves(self) -> np.ndarray:
        buf = np.zeros((256, 3), dtype=np.uint8)
        n = lib().oc_game_legal_moves(self._p, buf.ctypes.data)
        return buf[:n].copy()

    def legal_uci(self):
        return [uci(m) for m in self.legal_moves()]

    def is_check(self) -> bool:
        return bool(lib().oc_game_is_check(self._p))

    def piece_at(self, sq: int) -> int:
        return lib().oc_game_piece_at(self._p, sq)

    def is_repetition(self, count: int) -> bool:
        return bool(lib().oc_game_is_repetition(self._p, count))

    def encode(self, node_depth: int | None = None):
        """`_encode`: (planes int8 [8,8,112] HWC, meta int32 [7])."""
        if node_depth is None:
            node_depth = self.ply
        planes = np.zeros((8, 8, 112), dtype=np.int8)
        meta = np.zeros(7, dtype=np.int32)
        lib().oc_game_encode(self._p, node_depth, planes.ctypes.data, meta.ctypes.data)
        return planes, meta

    def pack(self, node_depth: int | None = None):
        """The sc_position a Rust shim would build: (slot u64[8,8], meta i32[7], n_hist)."""
        if node_depth is None:
            node_depth = self.ply
        slot = np.zeros((8, 8), dtype=np.uint64)
        meta = np.zeros(7, dtype=np.int32)
        nh = C.c_int32(0)
        lib().oc_game_pack(self._p, node_depth, slot.ctypes.data, meta.ctypes.data, C.byref(nh))
        return slot, meta, nh.value

    def outcome(self, claim_draw: bool = True):
        w = C.c_int(-1)
        t = lib().oc_game_outcome(self._p, int(claim_draw), C.byref(w))
        return (t, w.value) if t else None

    def move_indices(self, moves=None) -> np.ndarray:
        if moves is None:
            moves = self.legal_moves()
        t = self.turn
        return np.array([move_index(m, t) for m in moves], dtype=np.int32)

    def perft(self, depth: int) -> int:
        return lib().oc_perft(self._p, depth)

    def position_hash(self) -> int:
        return lib().oc_position_hash(self._p)

    # ---- SAN (only what py/validation/sample.csv needs) -------------------------------
    def parse_san(self, san: str):
        s = san.rstrip("+#!?")
        legal = self.legal_moves()
        if s in ("O-O", "0-0", "O-O-O", "0-0-0"):
            want_df = 2 if s in ("O-O", "0-0") else -2
            c = [m for m in legal if abs(self.piece_at(int(m[0]))) == 6 and (int(m[1]) & 7) - (int(m[0]) & 7) == want_df]
            assert len(c) == 1, (san, [uci(m) for m in legal])
            return c[0]
        promo = 0
        if "=" in s:
            s, pr = s.split("=")
            promo = SYMBOL_PIECES[pr.lower()]
        pt = 1
        if s[0] in "NBRQK":
            pt = SYMBOL_PIECES[s[0].lower()]
            s = s[1:]
        s = s.replace("x", "")
        to = sq_parse(s[-2:])
        dis = s[:-2]
        c = []
        for m in legal:
            f, t, p = int(m[0]), int(m[1]), int(m[2])
            if t != to or p != promo or abs(self.piece_at(f)) != pt:
                continue
            if any((ch in "abcdefgh" and "abcdefgh"[f & 7] != ch) or (ch in "12345678" and str((f >> 3) + 1) != ch) for ch in dis):
                continue
            c.append(m)
        assert len(c) == 1, (san, [uci(m) for m in legal])
        return c[0]


def move_index(m, turn: int) -> int:
    return lib().oc_move_index(int(m[0]), int(m[1]), int(m[2]), int(turn))


def post_process(logp_row: np.ndarray, idx: np.ndarray) -> np.ndarray:
    logp_row = np.ascontiguousarray(logp_row, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    out = np.zeros(len(idx), dtype=np.float32)
    lib().oc_post_process(logp_row.ctypes.data, idx.ctypes.data, len(idx), out.ctypes.data)
    return out


EVAL_FN = C.CFUNCTYPE(C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_float))


class Tree:
    """Sequential PUCT search of mcts.rs over the oracle rules; evaluator is either the
    built-in position-hash stand-in (evaluator=None) or a Python callable
    f(game: Game, node_depth, moves[n,3]) -> (priors[n], value)."""

    def __init__(self, evaluator=None):
        L = lib()
        if evaluator is None:
            self._cb = None
            self._t = L.oc_tree_new_hash()
        else:
            def _cb(ctx, gptr, depth, moves, n, priors):
                g = Game(_ptr=gptr)
                g._own = False
                mv = np.ctypeslib.as_array(moves, shape=(n * 3,)).reshape(n, 3).copy()
                p, v = evaluator(g, depth, mv)
                for i in range(n):
                    priors[i] = float(p[i])
                return float(v)

            self._cb = EVAL_FN(_cb)
            self._t = L.oc_tree_new(C.cast(self._cb, C.c_void_p), None)

    def __del__(self):
        if getattr(self, "_t", None):
            lib().oc_tree_free(self._t)
            self._t = None

    @property
    def game(self) -> Game:
        g = Game(_ptr=lib().oc_tree_game(self._t))
        g._own = False
        return g

    def search(self, n_rollout: int, cpuct: float):
        lib().oc_tree_search(self._t, n_rollout, cpuct)

    def root_children(self):
        mv = np.zeros((256, 3), dtype=np.uint8)
        n_act = np.zeros(256, dtype=np.int32)
        q = np.zeros(256, dtype=np.float32)
        u = np.zeros(256, dtype=np.float32)
        n = lib().oc_tree_root_children(self._t, mv.ctypes.data, n_act.ctypes.data, q.ctypes.data, u.ctypes.data)
        return mv[:n].copy(), n_act[:n].copy(), q[:n].copy(), u[:n].copy()

    def root_q(self) -> float:
        return lib().oc_tree_root_q(self._t)

    def root_n(self) -> int:
        return lib().oc_tree_root_n(self._t)

    def step_argmax(self) -> int:
        return lib().oc_tree_step_argmax(self._t)

    def step_index(self, i: int) -> int:
        return lib().oc_tree_step_index(self._t, i)

    @property
    def n_evals(self) -> int:
        return lib().oc_tree_n_evals(self._t)

    @property
    def n_predicts(self) -> int:
        return lib().oc_tree_n_predicts(self._t)


def random_play_positions(n: int, seed: int = 1, max_ply: int = 150):
    """SURVEY 8(d) synthetic positions: seeded uniform random play from the start
    position, restart on mate/stalemate or at ply `max_ply`; every position keeps its
    true history.  Returns a list of Game objects (copies)."""
    rng = np.random.RandomState(seed)
    out = []
    g = Game()
    while len(out) < n:
        mv = g.legal_moves()
        if len(mv) == 0 or g.ply >= max_ply:
            g = Game()
            continue
        out.append(g.dup())
        g.push(mv[rng.randint(len(mv))])
    return out


# ----------------------------------------------------------------------------------------------
# `chess_encode_steps` (/root/reference/src/lib.rs:47-128), restated LITERALLY: per step the board
# is extracted (`Board::extract`, chess.rs:356-412), optionally rotated (`Board::rotate`,
# chess.rs:594-621), pushed to the FRONT of an 8-deep history, and the whole history is viewed with
# `rotate = (step.turn == Black)` (`BoardHistory::view`, chess.rs:828-842) -- so with apply_mirror
# boards may be rotated twice.  Pure Python on purpose (small cases only).
# ----------------------------------------------------------------------------------------------
def _extract_board(g: "Game") -> dict:
    pm = {}
    for sq in range(64):
        p = g.piece_at(sq)
        if p:
            pm[(sq >> 3, sq & 7)] = (abs(p), 1 if p > 0 else 0)
    _, meta, _ = g.pack()
    return {"piece_map": pm, "turn": int(meta[0]), "fullmove": int(meta[1]), "K": (int(meta[2]), int(meta[4])),
            "Q": (int(meta[3]), int(meta[5])), "halfmove": int(meta[6]), "rep2": g.is_repetition(2),
            "rep3": g.is_repetition(3)}


def _rotate_board(b: dict) -> dict:
    return {"piece_map": {(7 - r, f): (pt, 1 - c) for (r, f), (pt, c) in b["piece_map"].items()},
            "turn": 1 - b["turn"], "fullmove": b["fullmove"] + (1 if b["turn"] == 1 else 0),
            "K": (b["K"][1], b["K"][0]), "Q": (b["Q"][1], b["Q"][0]), "halfmove": b["halfmove"], "rep2": b["rep2"],
            "rep3": b["rep3"]}


def _encode_pieces(b: dict) -> np.ndarray:
    a = np.zeros((8, 8, 14), dtype=np.int8)
    for (r, f), (pt, c) in b["piece_map"].items():
        a[r, f, (pt - 1) + (0 if c == 1 else 6)] = 1
    a[:, :, 12] = int(b["rep2"])
    a[:, :, 13] = int(b["rep3"])
    return a


def encode_steps(steps, apply_mirror: bool):
    """steps: [((from,to,promo), [((from,to,promo), count), ...]), ...] -> list of
    (planes int8[8,8,112], meta int32[7], dist float32[4672], [legal move indices])."""
    from collections import deque

    g = Game()
    hist = deque(maxlen=8)
    out = []
    for mv, children in steps:
        legal = g.legal_moves()
        assert {tuple(int(x) for x in m) for m in legal} == {tuple(int(x) for x in m) for m, _ in children}, "inconsistent moves"
        assert tuple(int(x) for x in mv) in {tuple(int(x) for x in m) for m in legal}
        b = _extract_board(g)
        step = _rotate_board(b) if apply_mirror else b
        ori_turn = (1 - step["turn"]) if apply_mirror else step["turn"]
        hist.appendleft(step)
        rot = step["turn"] == 0
        planes = np.zeros((8, 8, 112), dtype=np.int8)
        for i, hb in enumerate(hist):
            planes[:, :, 14 * i:14 * i + 14] = _encode_pieces(_rotate_board(hb) if rot else hb)
        meta = np.array([step["turn"], step["fullmove"], step["K"][0], step["Q"][0], step["K"][1], step["Q"][1],
                         step["halfmove"]], dtype=np.int32)
        idx = [move_index(m, ori_turn) for m in legal]
        total = np.float32(sum(int(c) for _, c in children))
        dist = np.zeros(4672, dtype=np.float32)
        for m, c in children:
            dist[move_index(m, ori_turn)] = np.float32(c) / (total + np.float32(1e-5))
        out.append((planes, meta, dist, idx))
        g.push(mv)
    return out
