"""Pins oracle/net.py to the reference's own network and writes tests/golden/net_golden.npz.

Run in the build container only (needs /root/reference):  python oracle/make_golden_net.py

What it does:
  1. imports the UNMODIFIED /root/reference/py/module.py behind a shim of the one missing
     third-party class (timm.layers.norm.LayerNorm2d: LayerNorm over C of an NCHW map,
     eps 1e-6 -- restated from memory of timm 1.0.26, the only unpinned piece);
  2. asserts that oracle.net.init_state_dict(19, seed 0) == load_model(...).state_dict()
     tensor for tensor, and that oracle.net.forward == the reference forward bit for bit;
  3. stores logp/value of the reference for 8 real positions (taken from the sample.csv
     games via the chess oracle) for the seed-0 19-block net and for a 2-block net with
     perturbed LayerNorm parameters, plus the weight digest used to recognise the net.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import chess_oracle as co  # noqa: E402
import net  # noqa: E402


def import_reference_module():
    timm = types.ModuleType("timm")
    layers = types.ModuleType("timm.layers")
    norm = types.ModuleType("timm.layers.norm")

    class LayerNorm2d(torch.nn.LayerNorm):
        def __init__(self, num_channels, eps=1e-6, affine=True):
            super().__init__(num_channels, eps=eps, elementwise_affine=affine)

        def forward(self, x):
            x = x.permute(0, 2, 3, 1)
            x = torch.nn.functional.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
            return x.permute(0, 3, 1, 2)

    norm.LayerNorm2d = LayerNorm2d
    sys.modules["timm"] = timm
    sys.modules["timm.layers"] = layers
    sys.modules["timm.layers.norm"] = norm
    sys.path.insert(0, "/root/reference/py")
    import module  # the reference's file, unmodified

    return module


def golden_positions():
    with open(os.path.join(HERE, "..", "tests", "golden", "sample_games.json")) as f:
        games = json.load(f)["games"]
    picks = [(1, 0), (1, 1), (1, 17), (2, 30), (3, 44), (3, 45), (5, 12), (1, 60)]
    planes, metas = [], []
    for gi, ply in picks:
        g = co.Game()
        for u in games[gi]["uci"].split()[:ply]:
            g.push(u)
        p, m = g.encode()
        planes.append(p)
        metas.append(m)
    return np.stack(planes), np.stack(metas), picks


def main():
    module = import_reference_module()
    planes_i8, meta_i32, picks = golden_positions()
    x = net.planes_i8_hwc_to_nchw(planes_i8)
    meta = torch.from_numpy(meta_i32).float()
    out = {"planes_i8": planes_i8, "meta_i32": meta_i32, "picks": np.array(picks)}

    ref = module.load_model(n_res_blocks=19, device="cpu", compile=False)
    sd = net.init_state_dict(19, 0)
    rsd = ref.state_dict()
    assert list(rsd.keys()) == list(sd.keys())
    for k in rsd:
        assert torch.equal(rsd[k], sd[k]), k
    with torch.no_grad():
        lp, v = ref(x, meta)
    lp2, v2 = net.forward(sd, x, meta)
    assert torch.equal(lp, lp2) and torch.equal(v, v2)
    out["logp19"] = lp.numpy()
    out["value19"] = v.numpy().reshape(-1)
    digest19 = net.state_dict_digest(sd)

    # 2-block net with non-trivial LayerNorm parameters, loaded INTO the reference module
    ref2 = module.ChessModule(n_res_blocks=2).eval()
    sd2 = net.perturb_norm_params(net.init_state_dict(2, 7))
    ref2.load_state_dict(sd2, strict=True)
    with torch.no_grad():
        lp, v = ref2(x, meta)
    lp2, v2 = net.forward(sd2, x, meta)
    assert torch.equal(lp, lp2) and torch.equal(v, v2)
    out["logp2"] = lp.numpy()
    out["value2"] = v.numpy().reshape(-1)
    digest2 = net.state_dict_digest(sd2)

    # NormTable / use_se variants (module.py:6-9, 28-36): 2-block nets, weights from oracle.net.init_variant_state_dict
    # loaded INTO the reference module (strict: keys and shapes are the reference's), outputs from the reference forward
    variants = {}
    for tag, norm, use_se, seed in (("bn_se", "BatchNorm", True, 21), ("ln_nose", "LayerNorm", False, 22),
                                    ("bn_nose", "BatchNorm", False, 23)):
        refv = module.ChessModule(n_res_blocks=2, use_se=use_se, norm=norm).eval()
        sdv = net.init_variant_state_dict(2, seed, norm, use_se)
        refv.load_state_dict(sdv, strict=True)
        with torch.no_grad():
            lp, v = refv(x, meta)
        lpo, vo = net.forward(sdv, x, meta)
        assert torch.equal(lp, lpo) and torch.equal(v, vo), tag
        out["logp_" + tag] = lp.numpy()
        out["value_" + tag] = v.numpy().reshape(-1)
        variants[tag] = {"norm": norm, "use_se": use_se, "seed": seed, "n_res_blocks": 2,
                         "digest": net.state_dict_digest({k: t for k, t in sdv.items() if t.ndim > 0})}

    dst = os.path.join(HERE, "..", "tests", "golden", "net_golden.npz")
    np.savez_compressed(dst, **out)
    with open(os.path.join(HERE, "..", "tests", "golden", "net_golden.json"), "w") as f:
        json.dump({"torch": torch.__version__, "digest19": digest19, "digest2": digest2,
                   "net2": {"n_res_blocks": 2, "seed": 7, "perturb_seed": 1234}, "variants": variants}, f, indent=1)
    print("ok", os.path.getsize(dst), digest19, digest2)


if __name__ == "__main__":
    main()
