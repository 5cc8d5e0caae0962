import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "smart-chess-rust_b200"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def co():
    import chess_oracle

    chess_oracle.lib()
    return chess_oracle


@pytest.fixture(scope="session")
def sample_games():
    with open(os.path.join(GOLDEN, "sample_games.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def net_golden():
    d = dict(np.load(os.path.join(GOLDEN, "net_golden.npz")))
    with open(os.path.join(GOLDEN, "net_golden.json")) as f:
        d["info"] = json.load(f)
    return d


def games_to_batch(games, node_depths=None):
    """oracle Game objects -> (sc_position array, sc_move array, move_off, list of legal move arrays)."""
    import scb200

    n = len(games)
    pos = np.zeros(n, dtype=scb200.POSITION_DTYPE)
    off = np.zeros(n + 1, dtype=np.int32)
    mv_all = []
    for i, g in enumerate(games):
        slot, meta, nh = g.pack(None if node_depths is None else node_depths[i])
        pos["slot"][i] = slot
        pos["meta"][i] = meta
        pos["n_hist"][i] = nh
        mv = g.legal_moves()
        mv_all.append(mv)
        off[i + 1] = off[i] + len(mv)
    moves = np.zeros(int(off[n]), dtype=scb200.MOVE_DTYPE)
    k = 0
    for mv in mv_all:
        for m in mv:
            moves[k] = (m[0], m[1], m[2], 0)
            k += 1
    return pos, moves, off, mv_all
