"""Network variants of the reference's module.py: `norm="BatchNorm"` (folded into the convolutions by the
exporter) and `use_se=False` (residual-only blocks), in every engine mode, against the oracle -- whose forward for these
weights equals the reference's own module.py bit for bit (tests/test_oracle_net.py, oracle/make_golden_net.py)."""
import numpy as np
import pytest
import torch

from conftest import games_to_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["bn_se", "ln_nose", "bn_nose"])
def test_variant_networks_all_modes(co, net_golden, tmp_path, tag, monkeypatch):
    import net
    import scb200

    info = net_golden["info"]["variants"][tag]
    sd = net.init_variant_state_dict(info["n_res_blocks"], info["seed"], info["norm"], info["use_se"])
    blob = str(tmp_path / f"{tag}.scw")
    scb200.write_blob(sd, blob)
    games = co.random_play_positions(70, seed=41)
    pos, moves, off, mv_all = games_to_batch(games)
    planes = np.stack([g.encode()[0] for g in games])
    meta = np.stack([g.encode()[1] for g in games])
    x = net.planes_i8_hwc_to_nchw(planes)
    lp, v = net.forward(sd, x, torch.from_numpy(meta).float())
    lp, v = lp.numpy(), v.numpy().reshape(-1)
    ref = np.concatenate([co.post_process(lp[i], g.move_indices(mv)) for i, (g, mv) in enumerate(zip(games, mv_all))])
    # the reference's own outputs on the golden positions, through the engine
    xg = net.planes_i8_hwc_to_nchw(net_golden["planes_i8"]).numpy()
    mg = net_golden["meta_i32"].astype(np.float32)
    same_net = abs(net.state_dict_digest({k: t for k, t in sd.items() if t.ndim > 0})["sum"] - info["digest"]["sum"]) < 1e-6
    for mode, tol in ((scb200.SC_MODE_FP32, 1e-4), (scb200.SC_MODE_FP32_FFMA, 1e-4), (scb200.SC_MODE_BF16, 2e-2)):
        monkeypatch.delenv("SCB200_LATENCY", raising=False)
        e = scb200.Engine(blob, 0, mode, 128)
        try:
            lp_g, v_g = e.forward_only(x.numpy(), meta.astype(np.float32))
            pri, val = e.eval(pos, moves, off)
            d = (np.abs(np.exp(lp_g) - np.exp(lp)).max() if mode == scb200.SC_MODE_BF16 else np.abs(lp_g - lp).max(),
                 np.abs(v_g - v).max(), np.abs(pri - ref).max(), np.abs(val - v).max())
            print(tag, "mode", mode, "max diffs (logp|p, value, prior, value)", ["%.2e" % t for t in d])
            assert max(d) < tol
            if same_net:
                lg, vg = e.forward_only(xg, mg)
                dg = np.abs(np.exp(lg) - np.exp(net_golden["logp_" + tag])).max() if mode == scb200.SC_MODE_BF16 else \
                    np.abs(lg - net_golden["logp_" + tag]).max()
                assert dg < tol and np.abs(vg - net_golden["value_" + tag]).max() < tol
            if mode == scb200.SC_MODE_BF16:
                # the small-batch latency kernel and the throughput kernel agree bit for bit on the variants too
                p3, v3 = e.eval(pos[:3], moves[: off[3]], off[:4])
                p3, v3 = p3.copy(), v3.copy()
                monkeypatch.setenv("SCB200_LATENCY", "0")
                e2 = scb200.Engine(blob, 0, mode, 128)
                try:
                    q3, w3 = e2.eval(pos[:3], moves[: off[3]], off[:4])
                    assert np.array_equal(p3, q3) and np.array_equal(v3, w3)
                    assert np.array_equal(v3, val[:3]) and np.array_equal(p3, pri[: off[3]])
                finally:
                    e2.close()
        finally:
            e.close()
