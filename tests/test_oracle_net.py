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
    assert nb == 1 and nt == len(sd) and nbytes >= sum(v.numel() for v in sd.values()) * 4
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
    bn = dict(sd)
    bn["conv_block.1.running_mean"] = torch.zeros(256)
    with pytest.raises(ValueError, match="BatchNorm"):
        scb200.write_blob(bn, p3)
    with pytest.raises(ValueError, match="missing"):
        scb200.write_blob({"foo": torch.zeros(1)}, p3)
