"""CPU tests: the network oracle against the golden vectors produced by the reference's own
module.py (oracle/make_golden_net.py)."""
import numpy as np
import pytest
import torch


def _digest_matches(d, ref):
    return d["numel"] == ref["numel"] and abs(d["sum"] - ref["sum"]) < 1e-6 and abs(d["wsum"] - ref["wsum"]) < 1e-4


def test_small_net_matches_reference_golden(net_golden):
    import net

    info = net_golden["info"]["net2"]
    sd = net.perturb_norm_params(net.init_state_dict(info["n_res_blocks"], info["seed"]), info["perturb_seed"])
    if not _digest_matches(net.state_dict_digest(sd), net_golden["info"]["digest2"]):
        pytest.skip("torch CPU RNG stream differs from the build container; golden net not reproducible here")
    x = net.planes_i8_hwc_to_nchw(net_golden["planes_i8"])
    meta = torch.from_numpy(net_golden["meta_i32"]).float()
    lp, v = net.forward(sd, x, meta)
    assert np.abs(lp.numpy() - net_golden["logp2"]).max() < 2e-5
    assert np.abs(v.numpy().reshape(-1) - net_golden["value2"]).max() < 2e-6
    assert np.allclose(np.exp(lp.numpy()).sum(1), 1.0, atol=1e-4)


def test_seed0_net_matches_reference_golden(net_golden):
    import net

    sd = net.init_state_dict(19, 0)
    assert sum(v.numel() for v in sd.values()) == 26203612
    if not _digest_matches(net.state_dict_digest(sd), net_golden["info"]["digest19"]):
        pytest.skip("torch CPU RNG stream differs from the build container; golden net not reproducible here")
    x = net.planes_i8_hwc_to_nchw(net_golden["planes_i8"][:4])
    meta = torch.from_numpy(net_golden["meta_i32"][:4]).float()
    lp, v = net.forward(sd, x, meta)
    assert np.abs(lp.numpy() - net_golden["logp19"][:4]).max() < 2e-5
    assert np.abs(v.numpy().reshape(-1) - net_golden["value19"][:4]).max() < 2e-6


def test_value_is_white_perspective(net_golden):
    # py/module.py:147-149: v * (2*turn - 1): flipping meta[0] alone flips the sign contribution
    import net

    sd = net.init_state_dict(1, 3)
    x = net.planes_i8_hwc_to_nchw(net_golden["planes_i8"][:1])
    m1 = torch.tensor([[1.0, 5, 1, 1, 1, 1, 0]])
    m0 = torch.tensor([[0.0, 5, 1, 1, 1, 1, 0]])
    sd["value_head.ffn.0.weight"][:, -7] = 0.0                   # make the FC ignore the turn input
    _, v1 = net.forward(sd, x, m1)
    _, v0 = net.forward(sd, x, m0)
    assert torch.allclose(v1, -v0, atol=1e-6)


def test_blob_roundtrip(tmp_path):
    import struct

    import net
    import scb200

    sd = net.init_state_dict(1, 0)
    p = str(tmp_path / "w.scw")
    nbytes = scb200.write_blob(sd, p)
    raw = open(p, "rb").read()
    assert raw[:8] == b"SCB2WTS1"
    nb, nt = struct.unpack("<II", raw[8:16])
    assert nb == 1 and nt == len(sd) + 1 and nbytes >= sum(v.numel() for v in sd.values()) * 4   # + `__config__`
    # lightning-style checkpoints are unwrapped like _load_ckpt (py/module.py:170-176)
    wrapped = {"pytorch-lightning_version": "2", "state_dict": {"model." + k: v for k, v in sd.items()}}
    p2 = str(tmp_path / "w2.scw")
    scb200.write_blob(wrapped, p2)
    assert open(p2, "rb").read() == raw
    # a torch.compile'd module's names carry `_orig_mod.` (load_model(compile=True), py/module.py:207-208)
    p3 = str(tmp_path / "w3.scw")
    scb200.write_blob({"_orig_mod." + k: v for k, v in sd.items()}, p3)
    assert open(p3, "rb").read() == raw
    # torch.save'd checkpoint through the CLI path
    ck = str(tmp_path / "c.ckpt")
    torch.save(wrapped, ck)
    p4 = str(tmp_path / "w4.scw")
    scb200.export_checkpoint(ck, p4)
    assert open(p4, "rb").read() == raw
    # inputs the engine cannot run are refused at export time, not at load time
    with pytest.raises(ValueError, match="missing"):
        scb200.write_blob({"foo": torch.zeros(1)}, p3)


def test_oracle_variants_match_reference_goldens(net_golden):
    """`NormTable["BatchNorm"]` and `use_se=False` (py/module.py:6-9, 28-36): the oracle's forward for the three variant
    nets equals the outputs the reference's own module.py produced for the same weights (oracle/make_golden_net.py)."""
    import net

    x = net.planes_i8_hwc_to_nchw(net_golden["planes_i8"])
    meta = torch.from_numpy(net_golden["meta_i32"]).float()
    for tag, info in net_golden["info"]["variants"].items():
        sd = net.init_variant_state_dict(info["n_res_blocks"], info["seed"], info["norm"], info["use_se"])
        d = net.state_dict_digest({k: t for k, t in sd.items() if t.ndim > 0})
        if abs(d["sum"] - info["digest"]["sum"]) > 1e-6:
            pytest.skip("seeded variant net differs on this machine; golden vectors not comparable")
        assert ("res_blocks.0.se.fc1.weight" in sd) == info["use_se"]
        assert ("conv_block.1.running_mean" in sd) == (info["norm"] == "BatchNorm")
        assert ("conv_block.0.bias" in sd) == (info["norm"] != "BatchNorm")
        lp, v = net.forward(sd, x, meta)
        assert np.abs(lp.numpy() - net_golden["logp_" + tag]).max() < 1e-6
        assert np.abs(v.numpy().reshape(-1) - net_golden["value_" + tag]).max() < 1e-6


def test_exporter_folds_batchnorm_and_marks_variants(tmp_path):
    """scb200.export: a BatchNorm checkpoint becomes conv + bias tensors without norm layers (folding identity checked
    in fp64 on one layer), and `__config__` = [norm folded, use_se]."""
    import struct

    import net
    import scb200
    from scb200 import export

    sd = net.init_variant_state_dict(1, 5, "BatchNorm", False)
    folded = export._fold_batchnorm(sd, 1)
    assert "conv_block.1.weight" not in folded and "conv_block.0.bias" in folded
    xin = torch.randn(2, 112, 8, 8, dtype=torch.float64)
    ref = torch.nn.functional.batch_norm(
        torch.nn.functional.conv2d(xin, sd["conv_block.0.weight"].double(), None, padding=1),
        sd["conv_block.1.running_mean"].double(), sd["conv_block.1.running_var"].double(),
        sd["conv_block.1.weight"].double(), sd["conv_block.1.bias"].double(), training=False, eps=1e-5)
    got = torch.nn.functional.conv2d(xin, folded["conv_block.0.weight"].double(), folded["conv_block.0.bias"].double(), padding=1)
    assert (ref - got).abs().max() < 1e-5
    p = str(tmp_path / "v.scw")
    scb200.write_blob(sd, p)
    raw = open(p, "rb").read()
    assert b"__config__" in raw and b"running_mean" not in raw and b"se.fc1" not in raw
