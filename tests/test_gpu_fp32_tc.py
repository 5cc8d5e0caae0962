"""FP32 parity mode on the tensor cores (SC_MODE_FP32: bf16x3 split operands, 6 bf16 products per fp32 product, fp32
accumulate in TMEM) against its referee, the same network in FP32 FFMA on the CUDA cores (SC_MODE_FP32_FFMA), and
against the CPU oracle.  Every other fp32 test of the suite (1e-4 gates vs oracle / reference goldens, the configs[0]
golden game, CPU-reference search) runs on SC_MODE_FP32, i.e. on this path."""
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


def test_tensor_core_fp32_mode_vs_ffma_referee_and_oracle(co, net19):
    import net
    import scb200

    games = co.random_play_positions(192, seed=31)
    pos, moves, off, mv_all = games_to_batch(games)
    planes = np.stack([g.encode()[0] for g in games])
    meta = np.stack([g.encode()[1] for g in games])
    x = net.planes_i8_hwc_to_nchw(planes)
    lp, v = net.forward(net19[0], x, torch.from_numpy(meta).float())
    lp, v = lp.numpy(), v.numpy().reshape(-1)
    out = {}
    for mode in (scb200.SC_MODE_FP32, scb200.SC_MODE_FP32_FFMA):
        e = scb200.Engine(net19[1], 0, mode, 192)
        try:
            assert e.info()["mode"] == mode
            lp_g, v_g = e.forward_only(x.numpy(), meta.astype(np.float32))
            pri, val = e.eval(pos, moves, off)
            t0 = time.perf_counter()
            for _ in range(3):
                e.eval(pos, moves, off)
            ms = (time.perf_counter() - t0) / 3 * 1e3
            out[mode] = (lp_g, v_g, pri.copy(), val.copy(), ms)
        finally:
            e.close()
    for mode, (lp_g, v_g, pri, val, ms) in out.items():
        print(f"mode {mode}: max|d logp| {np.abs(lp_g - lp).max():.3e} max|d value| {np.abs(v_g - v).max():.3e} "
              f"vs oracle; 192 leaves in {ms:.2f} ms")
        assert np.abs(lp_g - lp).max() < 1e-4 and np.abs(v_g - v).max() < 1e-4
    a, b = out[scb200.SC_MODE_FP32], out[scb200.SC_MODE_FP32_FFMA]
    d_lp, d_v, d_p = np.abs(a[0] - b[0]).max(), np.abs(a[1] - b[1]).max(), np.abs(a[2] - b[2]).max()
    print(f"tensor-core split vs FFMA: max|d logp| {d_lp:.3e} max|d value| {d_v:.3e} max|d prior| {d_p:.3e}")
    assert d_lp < 5e-5 and d_v < 5e-5 and d_p < 5e-5
    assert a[4] < b[4]                      # and it is the faster one
