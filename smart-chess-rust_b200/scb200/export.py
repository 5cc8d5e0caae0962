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
    if any(k.endswith("running_mean") for k in sd):
        # NormTable["BatchNorm"] (py/module.py:6-9) is never built by load_model (py/module.py:199)
        raise ValueError("BatchNorm checkpoints are not supported: the engine implements the LayerNorm network "
                         "that the reference's load_model builds")
    required = ["conv_block.0.weight", "conv_block.0.bias", "conv_block.1.weight", "policy_head.model.2.weight",
                "value_head.ffn.0.weight", "value_head.ffn.2.weight"]
    missing = [k for k in required if k not in sd]
    if missing:
        raise ValueError("not a ChessModule state_dict, missing: " + ", ".join(missing))
    return sd


def write_blob(state_dict, path: str) -> int:
    sd = _normalise_state_dict(state_dict)
    n_blocks = 0
    while f"res_blocks.{n_blocks}.conv1.weight" in sd:
        n_blocks += 1
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
