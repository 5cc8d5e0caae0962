"""CPU oracle for the policy/value network (TEST INFRASTRUCTURE -- never on the product path).

Plain-torch fp32 restatement of the arithmetic of the reference network:

* `ChessModule.forward`            /root/reference/py/module.py:135-154
* stem `conv_block`                /root/reference/py/module.py:120-126
* `ResBlockSE.forward`             /root/reference/py/module.py:38-46
* `PolicyHead`                     /root/reference/py/module.py:65-80
* `ValueHead`                      /root/reference/py/module.py:83-106
* `load_model` (seed-0 init)       /root/reference/py/module.py:184-212

Third-party arithmetic restated here because the packages are not vendored in the
reference tree:

* `timm.layers.norm.LayerNorm2d` (timm==1.0.26, uv.lock): LayerNorm over the channel
  axis of an NCHW map, eps = 1e-6, affine.                         [parity: pinned only
  through `oracle/make_golden_net.py`, which runs the reference's own module.py with a
  10-line shim of that class -- the eps value itself is restated from memory]
* `torchvision.ops.SqueezeExcitation(C, C//2)` (torchvision 0.26, installed here):
  scale = sigmoid(fc2(relu(fc1(avgpool(x))))), out = scale * x, fc1/fc2 are 1x1 convs.

Pinning: `oracle/make_golden_net.py` imports the UNMODIFIED reference module.py in the
build container, checks (a) that `init_state_dict` reproduces `load_model`'s seed-0
tensors bit-for-bit and (b) that `forward` equals the reference forward, and writes the
fixtures under tests/golden/.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this file.
"""
from __future__ import annotations

import math
import struct
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-6          # timm LayerNorm2d default
N_PLANES = 112         # 8 history slots x 14 planes   (module.py:118-121)
N_META = 7
N_POLICY = 73 * 64     # 4672
C = 256


# --------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------
def _conv(sd, name, cin, cout, k):
    m = torch.nn.Conv2d(cin, cout, kernel_size=k, padding=k // 2, bias=True)
    sd[name + ".weight"] = m.weight.detach().clone()
    sd[name + ".bias"] = m.bias.detach().clone()


def _ln(sd, name, c):
    sd[name + ".weight"] = torch.ones(c)
    sd[name + ".bias"] = torch.zeros(c)


def _linear(sd, name, din, dout):
    m = torch.nn.Linear(din, dout)
    sd[name + ".weight"] = m.weight.detach().clone()
    sd[name + ".bias"] = m.bias.detach().clone()


def init_state_dict(n_res_blocks: int = 19, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Random-init weights, consuming the torch CPU RNG in the same order as
    `load_model` (module.py:197 `torch.manual_seed(0)` then `ChessModule(...)`,
    construction order module.py:119-133, 19-36, 93-104, 69-76), so that seed 0
    reproduces the reference's random-init network tensor for tensor."""
    torch.manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    _conv(sd, "conv_block.0", N_PLANES, C, 3)
    _ln(sd, "conv_block.1", C)
    for i in range(n_res_blocks):
        p = f"res_blocks.{i}."
        _conv(sd, p + "conv1", C, C, 3)
        _ln(sd, p + "bn1", C)
        _conv(sd, p + "conv2", C, C, 3)
        _ln(sd, p + "bn2", C)
        _conv(sd, p + "se.fc1", C, C // 2, 1)
        _conv(sd, p + "se.fc2", C // 2, C, 1)
    _conv(sd, "value_head.conv.0", C, C, 1)
    _ln(sd, "value_head.conv.1", C)
    _linear(sd, "value_head.ffn.0", 64 * C + N_META, 128)
    _linear(sd, "value_head.ffn.2", 128, 1)
    _conv(sd, "policy_head.model.0", C, C, 1)
    _ln(sd, "policy_head.model.1", C)
    _conv(sd, "policy_head.model.2", C, 73, 1)
    _ln(sd, "policy_head.model.3", 73)
    return sd


def perturb_norm_params(sd, seed: int = 1234, scale: float = 0.25):
    """Random-init leaves every LayerNorm at gamma=1, beta=0, which would let a kernel
    that ignores gamma/beta pass.  Tests use this to make them non-trivial."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, v in sd.items():
        is_norm = v.ndim == 1 and (".bn" in k or "conv_block.1" in k or "conv.1." in k
                                   or "model.1." in k or "model.3." in k)
        if is_norm and k.endswith(".weight"):
            out[k] = 1.0 + scale * (2 * torch.rand(v.shape, generator=g) - 1)
        elif is_norm and k.endswith(".bias"):
            out[k] = scale * (2 * torch.rand(v.shape, generator=g) - 1)
        else:
            out[k] = v.clone()
    return out


def init_variant_state_dict(n_res_blocks: int, seed: int, norm: str = "LayerNorm", use_se: bool = True):
    """Weights of `ChessModule(n_res_blocks, use_se=use_se, norm=norm)` (module.py:109-133): with norm="BatchNorm" the
    convolutions carry no bias (module.py:18, 69, 88, 117) and every norm layer has running statistics; with
    use_se=False the blocks have no SqueezeExcitation (module.py:28-36).  Keys and shapes are the reference module's
    (`oracle/make_golden_net.py` loads the result INTO it with strict=True); the values are a seeded draw of this
    function's own (norm parameters and running statistics non-trivial), not the reference's initialisation order."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    bn = norm == "BatchNorm"

    def conv(name, cin, cout, k):
        bound = 1.0 / math.sqrt(cin * k * k)
        sd[name + ".weight"] = (2 * torch.rand((cout, cin, k, k), generator=g) - 1) * bound
        if not bn:
            sd[name + ".bias"] = (2 * torch.rand((cout,), generator=g) - 1) * bound

    def nrm(name, c):
        sd[name + ".weight"] = 1.0 + 0.25 * (2 * torch.rand((c,), generator=g) - 1)
        sd[name + ".bias"] = 0.25 * (2 * torch.rand((c,), generator=g) - 1)
        if bn:
            sd[name + ".running_mean"] = 0.1 * (2 * torch.rand((c,), generator=g) - 1)
            sd[name + ".running_var"] = 0.05 + 0.2 * torch.rand((c,), generator=g)
            sd[name + ".num_batches_tracked"] = torch.tensor(100)

    def lin(name, din, dout):
        bound = 1.0 / math.sqrt(din)
        sd[name + ".weight"] = (2 * torch.rand((dout, din), generator=g) - 1) * bound
        sd[name + ".bias"] = (2 * torch.rand((dout,), generator=g) - 1) * bound

    def se_conv(name, cin, cout):      # SqueezeExcitation's 1x1 convs always have a bias (torchvision)
        bound = 1.0 / math.sqrt(cin)
        sd[name + ".weight"] = (2 * torch.rand((cout, cin, 1, 1), generator=g) - 1) * bound
        sd[name + ".bias"] = (2 * torch.rand((cout,), generator=g) - 1) * bound

    conv("conv_block.0", N_PLANES, C, 3)
    nrm("conv_block.1", C)
    for i in range(n_res_blocks):
        p = f"res_blocks.{i}."
        conv(p + "conv1", C, C, 3)
        nrm(p + "bn1", C)
        conv(p + "conv2", C, C, 3)
        nrm(p + "bn2", C)
        if use_se:
            se_conv(p + "se.fc1", C, C // 2)
            se_conv(p + "se.fc2", C // 2, C)
    conv("value_head.conv.0", C, C, 1)
    nrm("value_head.conv.1", C)
    lin("value_head.ffn.0", 64 * C + N_META, 128)
    lin("value_head.ffn.2", 128, 1)
    conv("policy_head.model.0", C, C, 1)
    nrm("policy_head.model.1", C)
    conv("policy_head.model.2", C, 73, 1)
    nrm("policy_head.model.3", 73)
    return sd


def n_res_blocks_of(sd) -> int:
    n = 0
    while f"res_blocks.{n}.conv1.weight" in sd:
        n += 1
    return n


def state_dict_digest(sd) -> dict:
    """Small, order-sensitive fingerprint used to check that the seed-0 net built on
    another machine is the same net the golden vectors were made with."""
    tot = 0.0
    wsum = 0.0
    n = 0
    for i, (k, v) in enumerate(sd.items()):
        d = v.double().flatten()
        tot += float(d.sum())
        wsum += float((d * torch.arange(1, d.numel() + 1, dtype=torch.float64)).sum()) / d.numel() * (i + 1)
        n += d.numel()
    return {"numel": n, "sum": tot, "wsum": wsum}


# --------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------
def layer_norm_2d(x, w, b):
    """timm LayerNorm2d: normalise each (n, :, h, w) channel vector."""
    y = F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, LN_EPS)
    return y.permute(0, 3, 1, 2)


BN_EPS = 1e-5          # torch.nn.BatchNorm2d default


def norm_2d(sd, name, x):
    """NormTable (module.py:6-9): timm LayerNorm2d, or BatchNorm2d in eval mode (running statistics)."""
    if name + ".running_mean" in sd:
        return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                            sd[name + ".bias"], training=False, eps=BN_EPS)
    return layer_norm_2d(x, sd[name + ".weight"], sd[name + ".bias"])


def squeeze_excite(sd, p, x):
    if p + "fc1.weight" not in sd:      # use_se=False: torch.nn.Identity (module.py:35)
        return x
    s = F.adaptive_avg_pool2d(x, 1)
    s = F.relu(F.conv2d(s, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
    s = torch.sigmoid(F.conv2d(s, sd[p + "fc2.weight"], sd[p + "fc2.bias"]))
    return s * x


def res_block(sd, p, x):
    """module.py:38-46."""
    out = F.conv2d(x, sd[p + "conv1.weight"], sd.get(p + "conv1.bias"), padding=1)
    out = F.relu(norm_2d(sd, p + "bn1", out))
    out = F.conv2d(out, sd[p + "conv2.weight"], sd.get(p + "conv2.bias"), padding=1)
    out = norm_2d(sd, p + "bn2", out)
    out = squeeze_excite(sd, p + "se.", out)
    return F.relu(out + x)


def tower(sd, planes):
    x = F.conv2d(planes, sd["conv_block.0.weight"], sd.get("conv_block.0.bias"), padding=1)
    x = F.relu(norm_2d(sd, "conv_block.1", x))
    for i in range(n_res_blocks_of(sd)):
        x = res_block(sd, f"res_blocks.{i}.", x)
    return x


def policy_head(sd, latent):
    """module.py:70-80: conv1x1 -> LN(256) -> conv1x1(->73) -> LN(73) -> Flatten (NCHW:
    index = c*64 + h*8 + w) -> log_softmax over all 4672."""
    p = "policy_head.model."
    x = F.conv2d(latent, sd[p + "0.weight"], sd.get(p + "0.bias"))
    x = norm_2d(sd, p + "1", x)
    x = F.conv2d(x, sd[p + "2.weight"], sd.get(p + "2.bias"))
    x = norm_2d(sd, p + "3", x)
    return F.log_softmax(x.flatten(1), dim=1)


def value_head(sd, latent, meta):
    """module.py:89-106 and the white-perspective flip at module.py:147-149."""
    p = "value_head."
    x = F.conv2d(latent, sd[p + "conv.0.weight"], sd.get(p + "conv.0.bias"))
    x = F.relu(norm_2d(sd, p + "conv.1", x)).flatten(1)
    x = torch.cat((x, meta), dim=1)
    x = F.relu(F.linear(x, sd[p + "ffn.0.weight"], sd[p + "ffn.0.bias"]))
    v = torch.tanh(F.linear(x, sd[p + "ffn.2.weight"], sd[p + "ffn.2.bias"]))
    turn = meta[:, 0:1]
    return v * (turn * 2 - 1)


@torch.no_grad()
def forward(sd, planes, meta):
    """planes [B,112,8,8] float, meta [B,7] float -> (logp [B,4672], value [B,1])."""
    latent = tower(sd, planes.float())
    return policy_head(sd, latent), value_head(sd, latent, meta.float())


@torch.no_grad()
def forward_bf16_operands(sd, planes, meta):
    """Emulation of the tensor-core mode: every conv/linear rounds its input and its
    weight to bf16, accumulates in fp32; everything else stays fp32.  Used to size the
    bf16 tolerance, not as a parity target."""
    q = OrderedDict()
    for k, v in sd.items():
        q[k] = v.bfloat16().float() if (v.ndim > 1) else v

    def rnd(t):
        return t.bfloat16().float()

    class _Q:
        pass

    def conv(x, w, b, pad=0):
        return F.conv2d(rnd(x), w, b, padding=pad)

    x = conv(planes.float(), q["conv_block.0.weight"], q["conv_block.0.bias"], 1)
    x = F.relu(layer_norm_2d(x, q["conv_block.1.weight"], q["conv_block.1.bias"]))
    for i in range(n_res_blocks_of(sd)):
        p = f"res_blocks.{i}."
        o = conv(x, q[p + "conv1.weight"], q[p + "conv1.bias"], 1)
        o = F.relu(layer_norm_2d(o, q[p + "bn1.weight"], q[p + "bn1.bias"]))
        o = conv(o, q[p + "conv2.weight"], q[p + "conv2.bias"], 1)
        o = layer_norm_2d(o, q[p + "bn2.weight"], q[p + "bn2.bias"])
        s = F.adaptive_avg_pool2d(o, 1)
        s = F.relu(F.conv2d(s, sd[p + "se.fc1.weight"], sd[p + "se.fc1.bias"]))
        s = torch.sigmoid(F.conv2d(s, sd[p + "se.fc2.weight"], sd[p + "se.fc2.bias"]))
        x = F.relu(s * o + x)
    ph = "policy_head.model."
    y = conv(x, q[ph + "0.weight"], q[ph + "0.bias"])
    y = layer_norm_2d(y, q[ph + "1.weight"], q[ph + "1.bias"])
    y = conv(y, q[ph + "2.weight"], q[ph + "2.bias"])
    y = layer_norm_2d(y, q[ph + "3.weight"], q[ph + "3.bias"])
    logp = F.log_softmax(y.flatten(1), dim=1)
    vh = "value_head."
    z = conv(x, q[vh + "conv.0.weight"], q[vh + "conv.0.bias"])
    z = F.relu(layer_norm_2d(z, q[vh + "conv.1.weight"], q[vh + "conv.1.bias"])).flatten(1)
    z = torch.cat((rnd(z), meta.float()), dim=1)
    z = F.relu(F.linear(z, sd[vh + "ffn.0.weight"].bfloat16().float(), sd[vh + "ffn.0.bias"]))
    v = torch.tanh(F.linear(z, sd[vh + "ffn.2.weight"], sd[vh + "ffn.2.bias"]))
    return logp, v * (meta[:, 0:1].float() * 2 - 1)


def planes_i8_hwc_to_nchw(planes_hwc: np.ndarray) -> torch.Tensor:
    """`Array3<i8>` (8,8,112) HWC per position -> float NCHW, as the backends do
    (/root/reference/src/backends/onnx.rs:66-69, torch.rs:115-119)."""
    t = torch.from_numpy(np.ascontiguousarray(planes_hwc)).float()
    return t.permute(0, 3, 1, 2).contiguous()


def priors_from_logp(logp_row: np.ndarray, move_index: np.ndarray) -> np.ndarray:
    """`_get_move_distribution` + `post_process_distr`
    (/root/reference/src/backends/onnx.rs:81-91, /root/reference/src/chess.rs:891-901):
    p_k = exp(logp[idx_k]); p_k / (sum_k p_k + 1e-5), f32, sum taken left to right."""
    p = np.exp(logp_row.astype(np.float32)[move_index]).astype(np.float32)
    s = np.float32(0.0)
    for x in p:
        s = np.float32(s + x)
    s = np.float32(s + np.float32(1e-5))
    return (p / s).astype(np.float32)
