"""Weight-blob exporter for the B200 backend.

Plays the role of `export_pt_f32` / `export_pt_bf16` (/root/reference/py/export.py:36-65)
and of `scripts/export_model.py` for the new backend: instead of a TorchScript file the
engine loads a flat blob of the state_dict tensors under the reference's own parameter
names (py/module.py:109-133).  Accepts a raw state_dict or a Lightning checkpoint, like
`_load_ckpt` (py/module.py:157-181).

Blob layout (little endian):
    8s   magic  b"SCB2WTS1"
    u32  n_res_blocks
    u32  n_tensors
    per tensor: u32 name_len, name, u32 ndim, u32 dims[ndim], u64 offset, u64 numel
    u64  data_bytes
    data: fp32, each tensor 64-byte aligned, `offset` relative to the start of data

Network variants (`NormTable`, `use_se`, py/module.py:6-9, 14-36): a BatchNorm network is exported with every
BatchNorm folded into the convolution before it (inference uses the running statistics, so
`bn(conv(x)) = conv(x) * g / sqrt(var + eps) + (beta - mean * g / sqrt(var + eps))` is a convolution with scaled
weights and a bias); the norm tensors are dropped and the engine runs those layers without normalisation.  A
`use_se=False` network simply has no `se.*` tensors.  The tensor `__config__` = [norm folded, use_se] tells the engine.
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"SCB2WTS1"


def _normalise_state_dict(obj):
    """Raw state_dict, Lightning checkpoint (py/module.py:172-177 chops the first name component),
    or the state_dict of a torch.compile'd / scripted module (`_orig_mod.` prefix)."""
    if hasattr(obj, "state_dict") and not isinstance(obj, dict):
        obj = obj.state_dict()
    if isinstance(obj, dict) and "pytorch-lightning_version" in obj:
        obj = {k.split(".", 1)[1]: v for k, v in obj["state_dict"].items()}
    sd = {(k[len("_orig_mod."):] if k.startswith("_orig_mod.") else k): v for k, v in obj.items()}
    required = ["conv_block.0.weight", "conv_block.1.weight", "policy_head.model.2.weight",
                "value_head.ffn.0.weight", "value_head.ffn.2.weight"]
    missing = [k for k in required if k not in sd]
    if missing:
        raise ValueError("not a ChessModule state_dict, missing: " + ", ".join(missing))
    return sd


BN_EPS = 1e-5  # torch.nn.BatchNorm2d default


def _fold_batchnorm(sd, n_blocks):
    """NormTable["BatchNorm"] (py/module.py:6-9): conv (bias=False, py/module.py:18) + BatchNorm2d in eval mode ->
    conv with bias.  Returns a new dict without the norm tensors."""
    import torch

    pairs = [("conv_block.0", "conv_block.1"), ("value_head.conv.0", "value_head.conv.1"),
             ("policy_head.model.0", "policy_head.model.1"), ("policy_head.model.2", "policy_head.model.3")]
    for i in range(n_blocks):
        pairs += [(f"res_blocks.{i}.conv1", f"res_blocks.{i}.bn1"), (f"res_blocks.{i}.conv2", f"res_blocks.{i}.bn2")]
    out = {k: v for k, v in sd.items()}
    for conv, norm in pairs:
        w = sd[conv + ".weight"].detach().double()
        g, b = sd[norm + ".weight"].detach().double(), sd[norm + ".bias"].detach().double()
        mean, var = sd[norm + ".running_mean"].detach().double(), sd[norm + ".running_var"].detach().double()
        scale = g / torch.sqrt(var + BN_EPS)
        bias = sd[conv + ".bias"].detach().double() if conv + ".bias" in sd else torch.zeros_like(mean)
        out[conv + ".weight"] = (w * scale.view(-1, 1, 1, 1)).float()
        out[conv + ".bias"] = ((bias - mean) * scale + b).float()
        for suffix in (".weight", ".bias", ".running_mean", ".running_var", ".num_batches_tracked"):
            out.pop(norm + suffix, None)
    return out


def write_blob(state_dict, path: str) -> int:
    import numpy as _np

    sd = _normalise_state_dict(state_dict)
    n_blocks = 0
    while f"res_blocks.{n_blocks}.conv1.weight" in sd:
        n_blocks += 1
    folded = any(k.endswith("running_mean") for k in sd)
    use_se = "res_blocks.0.se.fc1.weight" in sd or n_blocks == 0
    if folded:
        sd = _fold_batchnorm(sd, n_blocks)
    sd = dict(sd)
    sd["__config__"] = _np.array([1.0 if folded else 0.0, 1.0 if use_se else 0.0], dtype=_np.float32)
    entries = []
    chunks = []
    off = 0
    for name, t in sd.items():
        a = np.ascontiguousarray(t.detach().cpu().float().numpy() if hasattr(t, "detach") else np.asarray(t, dtype=np.float32))
        pad = (-off) % 64
        if pad:
            chunks.append(b"\0" * pad)
            off += pad
        entries.append((name.encode(), a.shape, off, a.size))
        chunks.append(a.tobytes())
        off += a.nbytes
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", n_blocks, len(entries)))
        for name, shape, o, numel in entries:
            f.write(struct.pack("<I", len(name)))
            f.write(name)
            f.write(struct.pack("<I", len(shape)))
            for d in shape:
                f.write(struct.pack("<I", d))
            f.write(struct.pack("<QQ", o, numel))
        f.write(struct.pack("<Q", off))
        for c in chunks:
            f.write(c)
    return off


def export_checkpoint(checkpoint_path: str, out_path: str) -> int:
    """`scripts/export_model.py -c ckpt` equivalent for the .scw blob."""
    import torch

    try:
        ckpt = torch.load(checkpoint_path, weights_only=True, map_location="cpu")
    except Exception:
        # a TorchScript export of the reference (py/export.py:36-65) carries the same parameter names
        ckpt = torch.jit.load(checkpoint_path, map_location="cpu").state_dict()
    return write_blob(ckpt, out_path)


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description="export a checkpoint to the B200 backend's weight blob")
    ap.add_argument("-c", "--checkpoint", required=True)
    ap.add_argument("-o", "--output", required=True)
    a = ap.parse_args()
    print(export_checkpoint(a.checkpoint, a.output), "bytes of tensor data")
