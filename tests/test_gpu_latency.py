"""Latency path for small batches (tower_lat.cu: one 2-board tile per cluster of 8 / 4 CTAs) -- the `Game::predict`
drop-in is a batch of one (src/backends/torch.rs:115-125).  It must produce the bits of the throughput kernel
(tower_bf16.cu), so that a leaf's result does not depend on the batch it travels in."""
import time

import numpy as np
import pytest
import torch

from conftest import games_to_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net19(tmp_path_factory):
    import net
    import scb200

    sd = net.init_state_dict(19, 0)
    p = str(tmp_path_factory.mktemp("w19") / "n19.scw")
    scb200.write_blob(sd, p)
    return sd, p


@pytest.fixture(scope="module")
def leaves(co):
    games = co.random_play_positions(96, seed=77)
    return games, games_to_batch(games)


def _eval_sizes(blob, pos, moves, off, sizes, env, monkeypatch):
    import scb200

    for k in ("SCB200_LATENCY", "SCB200_LAT_CLUSTER", "SCB200_LAT_ONE_BOARD"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    e = scb200.Engine(blob, 0, scb200.SC_MODE_BF16, 128)      # switches are read at creation
    out = {}
    try:
        for n in sizes:
            l0 = e.launch_count()
            pri, val = e.eval(pos[:n], moves[: off[n]], off[: n + 1])
            out[n] = (pri.copy(), val.copy(), e.launch_count() - l0)
    finally:
        e.close()
    return out


def test_latency_kernel_bit_identical_to_throughput_kernel(net19, leaves, monkeypatch):
    games, (pos, moves, off, mv_all) = leaves
    sizes = (1, 2, 3, 7, 16, 33, 36, 37, 64, 74, 75, 96)      # both cluster sizes, odd tiles, and past the switch-over
    ref = _eval_sizes(net19[1], pos, moves, off, sizes, {"SCB200_LATENCY": "0"}, monkeypatch)
    # default policy; clusters of 4 only; clusters of 8 only; two-board tiles only
    for env in ({}, {"SCB200_LAT_CLUSTER": "4"}, {"SCB200_LAT_CLUSTER": "8"}, {"SCB200_LAT_ONE_BOARD": "0"}):
        got = _eval_sizes(net19[1], pos, moves, off, sizes, env, monkeypatch)
        for n in sizes:
            assert np.array_equal(got[n][1], ref[n][1]), (env, n, np.abs(got[n][1] - ref[n][1]).max())
            assert np.array_equal(got[n][0], ref[n][0]), (env, n, np.abs(got[n][0] - ref[n][0]).max())
    # and a leaf evaluated alone equals the same leaf inside the big batch
    assert np.array_equal(ref[1][1], ref[96][1][:1])


def test_latency_kernel_vs_oracle(co, net19, leaves):
    import net
    import scb200

    games, (pos, moves, off, mv_all) = leaves
    n = 9
    planes = np.stack([g.encode()[0] for g in games[:n]])
    meta = np.stack([g.encode()[1] for g in games[:n]])
    lp, v = net.forward(net19[0], net.planes_i8_hwc_to_nchw(planes), torch.from_numpy(meta).float())
    ref = np.concatenate([co.post_process(lp[i].numpy(), games[i].move_indices(mv_all[i])) for i in range(n)])
    e = scb200.Engine(net19[1], 0, scb200.SC_MODE_BF16, 16)
    try:
        pri, val = e.eval(pos[:n], moves[: off[n]], off[: n + 1])
        assert np.abs(pri - ref).max() < 2e-2 and np.abs(val - v.numpy().reshape(-1)).max() < 2e-2
        # latency of the one-leaf call through the C ABI with host buffers
        for nb in (1, 8):
            for it in range(40):
                if it == 10:
                    t0 = time.perf_counter()
                e.eval(pos[:nb], moves[: off[nb]], off[: nb + 1])
            print(f"sc_eval n={nb}: {(time.perf_counter() - t0) / 30 * 1e3:.3f} ms")
    finally:
        e.close()


def test_latency_kernel_repeatable_under_random_batch_sizes(net19, leaves, monkeypatch):
    """Race hunt: 300 calls with random batch sizes (every cluster shape, back to back on one engine, so the exchange
    buffers, mbarrier phases and the SE weight staging are re-used across launches); each leaf's result must be the
    bits the throughput kernel produced for it, every time."""
    import scb200

    games, (pos, moves, off, mv_all) = leaves
    ref = _eval_sizes(net19[1], pos, moves, off, (96,), {"SCB200_LATENCY": "0"}, monkeypatch)[96]
    for k in ("SCB200_LATENCY", "SCB200_LAT_CLUSTER", "SCB200_LAT_ONE_BOARD"):
        monkeypatch.delenv(k, raising=False)
    e = scb200.Engine(net19[1], 0, scb200.SC_MODE_BF16, 128)
    rng = np.random.RandomState(5)
    try:
        for it in range(300):
            n = int(rng.choice([1, 2, 3, 5, 8, 15, 16, 17, 31, 33, 34, 40, 63, 66]))
            lo = int(rng.randint(0, 96 - n + 1))
            sub_off = (off[lo:lo + n + 1] - off[lo]).astype(np.int32)
            pri, val = e.eval(pos[lo:lo + n], moves[off[lo]:off[lo + n]], sub_off)
            assert np.array_equal(val, ref[1][lo:lo + n]), (it, n, lo)
            assert np.array_equal(pri, ref[0][off[lo]:off[lo + n]]), (it, n, lo)
    finally:
        e.close()
