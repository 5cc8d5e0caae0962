"""Random-init weights of the reference architecture for benchmarks without checkpoints.

`load_model` (/root/reference/py/module.py:184-212) seeds torch with 0 and builds
`ChessModule(n_res_blocks=...)`; this produces a state_dict of the same tensors (same parameter names,
shapes, torch default initialisers drawn in the same construction order) without the reference tree.
"""
from __future__ import annotations

from collections import OrderedDict


def random_init_state_dict(n_res_blocks: int = 19, seed: int = 0):
    import torch

    torch.manual_seed(seed)
    sd = OrderedDict()

    def conv(name, cin, cout, k):
        m = torch.nn.Conv2d(cin, cout, kernel_size=k, padding=k // 2)
        sd[name + ".weight"], sd[name + ".bias"] = m.weight.detach().clone(), m.bias.detach().clone()

    def norm(name, c):
        sd[name + ".weight"], sd[name + ".bias"] = torch.ones(c), torch.zeros(c)

    def linear(name, din, dout):
        m = torch.nn.Linear(din, dout)
        sd[name + ".weight"], sd[name + ".bias"] = m.weight.detach().clone(), m.bias.detach().clone()

    conv("conv_block.0", 112, 256, 3)
    norm("conv_block.1", 256)
    for i in range(n_res_blocks):
        p = f"res_blocks.{i}."
        conv(p + "conv1", 256, 256, 3)
        norm(p + "bn1", 256)
        conv(p + "conv2", 256, 256, 3)
        norm(p + "bn2", 256)
        conv(p + "se.fc1", 256, 128, 1)
        conv(p + "se.fc2", 128, 256, 1)
    conv("value_head.conv.0", 256, 256, 1)
    norm("value_head.conv.1", 256)
    linear("value_head.ffn.0", 64 * 256 + 7, 128)
    linear("value_head.ffn.2", 128, 1)
    conv("policy_head.model.0", 256, 256, 1)
    norm("policy_head.model.1", 256)
    conv("policy_head.model.2", 256, 73, 1)
    norm("policy_head.model.3", 73)
    return sd
