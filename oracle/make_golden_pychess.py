#!/usr/bin/env python
"""Pins the rules half of the oracle (and the product's native rules) to python-chess itself.

    pip install chess==1.11.1        # the version the reference locks (pyproject.toml:10, uv.lock:309-310)
    python oracle/make_golden_pychess.py [--sample-csv /path/to/reference/py/validation/sample.csv]

python-chess is NOT installable in the offline build container of this repository, which is why rows a6 / c of
SURVEY.md section 8 are "parity unpinned".  Anybody WITH python-chess runs this script once; it writes
tests/golden/pychess_golden.json.gz, and tests/test_oracle_chess.py::test_rules_against_python_chess_goldens then
holds BOTH rule engines (oracle/chess_oracle.c and smart-chess-rust_b200/csrc/host/chess_rules.hpp) to it:
everything the reference takes from python-chess on the hot path (src/chess.rs:356-412, 665-788) --

  * `board.legal_moves` in generation order (child order = search tie-breaks, src/mcts.rs:78-88)
  * `board.is_repetition(2)` / `(3)`                       (planes 12/13 of every history slot)
  * `has_kingside/queenside_castling_rights` for both colours, `halfmove_clock`, `fullmove_number`, `turn` (meta)
  * `board.outcome(claim_draw=True)`: termination + winner  (src/chess.rs:719-729)

for every ply of (a) the 60 SAN games of py/validation/sample.csv (3 539 plies), (b) seeded uniform random play from
the initial position until >= 10 000 plies are recorded, (c) seeded random play from edge-case FENs (en passant with
a pinned capturer / discovered check, castling through and out of check, rights lost by rook capture, under-promotion,
half-move clock at 99 / 100 / 149, bare kings, K+B vs K, the 218-move position) and (d) shuffle games that walk through
two-, three- and five-fold repetition and past the 50- and 75-move marks.
"""
import argparse
import csv
import gzip
import json
import os
import random

import chess

HERE = os.path.dirname(os.path.abspath(__file__))

EDGE_FENS = [
    # tests/test_gpu_parity.py:220-222
    "R6R/3Q4/1Q4Q1/4Q3/2Q4Q/Q4Q2/pp1Q4/kBNN1KB1 w - - 0 1",
    "7k/8/8/8/8/8/5q2/7K w - - 0 1",
    "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 b kq - 0 1",
    # en passant: legal, pinned capturer (horizontal discovered check), capturer pinned on the diagonal
    "rnbqkbnr/ppp1p1pp/8/3pPp2/8/8/PPPP1PPP/RNBQKBNR w KQkq f6 0 3",
    "8/8/8/K2pP2r/8/8/8/4k3 w - d6 0 1",
    "8/8/8/8/k2Pp2Q/8/8/4K3 b - d3 0 1",
    "4k3/8/8/2b5/3Pp3/8/8/6K1 b - d3 0 1",
    # castling: through check, out of check, rook attacked (still legal), rights lost by capture
    "r3k2r/8/8/8/4r3/8/8/R3K2R w KQkq - 0 1",
    "r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 4 10",
    "r3k2r/8/8/8/8/5b2/8/R3K2R w KQkq - 0 1",
    "r3k2r/1B6/8/8/8/8/8/R3K2R w KQkq - 0 1",
    # promotions / under-promotions, both colours
    "3n1n1k/4P3/8/8/8/8/4p3/3N1N1K w - - 0 1",
    "3n1n2/4P3/8/8/8/8/4p3/k2N1N1K b - - 0 1",
    # half-move clock around the 50- and 75-move marks
    "8/8/8/4k3/8/8/3RK3/8 w - - 99 80",
    "8/8/8/4k3/8/8/3RK3/8 b - - 100 80",
    "8/8/8/4k3/8/8/3RK3/8 w - - 149 120",
    # insufficient material and its neighbours
    "8/8/8/4k3/8/8/4K3/8 w - - 0 1",
    "8/8/8/4k3/8/8/3BK3/8 w - - 0 1",
    "8/8/8/4k3/8/8/3NK3/8 b - - 0 1",
    "8/8/3b4/4k3/8/8/3BK3/8 w - - 0 1",
    "8/8/8/4k3/8/8/3NKN2/8 w - - 0 1",
    # double check, stalemate, smothered mate
    "4k3/8/8/8/8/5n2/4r3/4K3 w - - 0 1",
    "7k/5Q2/6K1/8/8/8/8/8 b - - 0 1",
    "6rk/5Npp/8/8/8/8/8/6K1 b - - 0 1",
]

TERMINATION = {  # src/chess.rs:87-105 == python-chess' enum values
    chess.Termination.CHECKMATE: 1, chess.Termination.STALEMATE: 2, chess.Termination.INSUFFICIENT_MATERIAL: 3,
    chess.Termination.SEVENTYFIVE_MOVES: 4, chess.Termination.FIVEFOLD_REPETITION: 5, chess.Termination.FIFTY_MOVES: 6,
    chess.Termination.THREEFOLD_REPETITION: 7,
}


def record(board: chess.Board) -> dict:
    oc = board.outcome(claim_draw=True)
    t = board.turn
    return {
        "legal": [m.uci() for m in board.legal_moves],
        "rep2": bool(board.is_repetition(2)), "rep3": bool(board.is_repetition(3)),
        # encode_meta order (src/chess.rs:652-662): turn, fullmove, K(stm), Q(stm), K(opp), Q(opp), halfmove
        "meta": [int(t), board.fullmove_number, int(board.has_kingside_castling_rights(t)),
                 int(board.has_queenside_castling_rights(t)), int(board.has_kingside_castling_rights(not t)),
                 int(board.has_queenside_castling_rights(not t)), board.halfmove_clock],
        "outcome": None if oc is None else [TERMINATION[oc.termination], -1 if oc.winner is None else int(oc.winner)],
    }


def play(board: chess.Board, moves, fen=None) -> dict:
    """One game: the record of every position met (before each move and after the last)."""
    plies = [record(board)]
    ucis = []
    for m in moves(board):
        ucis.append(m.uci())
        board.push(m)
        plies.append(record(board))
    return {"fen": fen, "moves": ucis, "plies": plies}


def san_moves(sans):
    def gen(board):
        for s in sans:
            yield board.parse_san(s)
    return gen


def random_moves(rng, max_plies):
    def gen(board):
        for _ in range(max_plies):
            legal = list(board.legal_moves)
            if not legal or board.is_game_over(claim_draw=False):
                return
            yield legal[rng.randrange(len(legal))]
    return gen


def uci_moves(ucis):
    def gen(board):
        for u in ucis:
            yield chess.Move.from_uci(u)
    return gen


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sample-csv", default="/root/reference/py/validation/sample.csv")
    ap.add_argument("--random-plies", type=int, default=10000)
    ap.add_argument("--seed", type=int, default=20261018)
    a = ap.parse_args()
    rng = random.Random(a.seed)
    games = []
    with open(a.sample_csv) as f:
        for row in csv.DictReader(f):
            games.append(dict(play(chess.Board(), san_moves(row["moves"].split())), source="sample.csv:" + row["id"]))
    n = 0
    while n < a.random_plies:
        g = play(chess.Board(), random_moves(rng, 220))
        n += len(g["plies"])
        games.append(dict(g, source="random"))
    for fen in EDGE_FENS:
        for k in range(3):
            games.append(dict(play(chess.Board(fen), random_moves(rng, 60), fen), source="edge"))
    shuffle = ["g1f3", "g8f6", "f3g1", "f6g8"]
    games.append(dict(play(chess.Board(), uci_moves(shuffle * 5)), source="repetition shuffle"))
    # 160 quiet plies without an early repetition: the rook visits a fresh square of the d-file / second rank pattern
    kr = chess.Board("8/8/8/4k3/8/8/3RK3/8 w - - 0 1")
    games.append(dict(play(kr, random_moves(rng, 320), kr.fen()), source="rook ending, random"))
    out = {"python_chess": chess.__version__, "seed": a.seed, "games": games,
           "n_plies": sum(len(g["plies"]) for g in games)}
    dst = os.path.join(HERE, "..", "tests", "golden", "pychess_golden.json.gz")
    with gzip.open(dst, "wt") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", dst, out["n_plies"], "plies from", len(games), "games; python-chess", chess.__version__)


if __name__ == "__main__":
    main()
