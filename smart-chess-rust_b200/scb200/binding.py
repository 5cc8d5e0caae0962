"""ctypes binding of include/sc_b200.h."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

SC_MODE_FP32 = 0
SC_MODE_BF16 = 1
SC_N_PLANES = 112
SC_N_META = 7
SC_N_POLICY = 4672

# struct sc_position { uint64_t slot[8][8]; int32_t meta[7]; int32_t n_hist; }  (544 bytes)
POSITION_DTYPE = np.dtype([("slot", np.uint64, (8, 8)), ("meta", np.int32, (7,)), ("n_hist", np.int32)])
# struct sc_move { uint8_t from, to, promo, pad; }
MOVE_DTYPE = np.dtype([("from", np.uint8), ("to", np.uint8), ("promo", np.uint8), ("pad", np.uint8)])
assert POSITION_DTYPE.itemsize == 544 and MOVE_DTYPE.itemsize == 4

DECLARED_SYMBOLS = [
    "sc_create", "sc_destroy", "sc_last_error", "sc_info", "sc_eval", "sc_eval_device", "sc_encode_only",
    "sc_move_index_only", "sc_forward_only", "sc_launch_count", "sc_last_timing", "sc_set_timing", "sc_kernel_timing",
]

_LIB = None


class SCError(RuntimeError):
    pass


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libscb200.so")


def load_library():
    """Loads libscb200.so; raises loudly if the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise SCError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for this backend)")
        L = C.CDLL(p)
        L.sc_last_error.restype = C.c_char_p
        L.sc_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.sc_destroy.argtypes = [C.c_void_p]
        L.sc_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.sc_eval.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        L.sc_eval_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
        L.sc_encode_only.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sc_move_index_only.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sc_forward_only.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sc_launch_count.restype = C.c_int64
        L.sc_launch_count.argtypes = [C.c_void_p]
        L.sc_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.sc_set_timing.argtypes = [C.c_void_p, C.c_int]
        L.sc_kernel_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        _LIB = L
    return _LIB


def _check(rc: int, what: str):
    if rc != 0:
        raise SCError(f"{what} failed ({rc}): {load_library().sc_last_error().decode()}")


def pack_positions(slots, metas, n_hists) -> np.ndarray:
    """Builds an sc_position array from per-leaf (slot[8,8] u64, meta[7] i32, n_hist)."""
    n = len(metas)
    out = np.zeros(n, dtype=POSITION_DTYPE)
    for i in range(n):
        out["slot"][i] = slots[i]
        out["meta"][i] = metas[i]
        out["n_hist"][i] = n_hists[i]
    return out


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):  # torch tensor (device or pinned host)
        return a.data_ptr()
    return a.ctypes.data


class Engine:
    """`ChessTS` / `ChessOnnx` counterpart (src/backends/torch.rs:14-17, onnx.rs:8-11): owns the
    model for the lifetime of the process and evaluates batches of leaves."""

    def __init__(self, blob_path: str, device: int = 0, mode: int = SC_MODE_BF16, max_batch: int = 2048):
        L = load_library()
        h = C.c_void_p()
        _check(L.sc_create(blob_path.encode(), device, mode, max_batch, C.byref(h)), "sc_create")
        self._h = h
        self.mode = mode
        self.max_batch = max_batch
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            load_library().sc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def info(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        _check(load_library().sc_info(self._h, C.byref(a), C.byref(b), C.byref(c)), "sc_info")
        return {"n_res_blocks": a.value, "max_batch": b.value, "mode": c.value}

    def eval(self, positions: np.ndarray, moves: np.ndarray, move_off: np.ndarray, priors_out=None, value_out=None,
             stream: int = 0):
        """`predict` for n leaves: returns (priors CSR float32, values float32[n])."""
        n = len(positions)
        positions = np.ascontiguousarray(positions, dtype=POSITION_DTYPE) if not hasattr(positions, "data_ptr") else positions
        total = int(move_off[n]) if n else 0
        if priors_out is None:
            priors_out = np.zeros(max(total, 1), dtype=np.float32)
        if value_out is None:
            value_out = np.zeros(max(n, 1), dtype=np.float32)
        _check(load_library().sc_eval(self._h, n, _ptr(positions), _ptr(moves), _ptr(move_off), _ptr(priors_out),
                                      _ptr(value_out), stream or None), "sc_eval")
        return priors_out[:total], value_out[:n]

    def eval_device(self, n, d_pos, d_moves, d_off, n_moves_total, d_priors, d_value, stream: int = 0):
        _check(load_library().sc_eval_device(self._h, n, _ptr(d_pos), _ptr(d_moves), _ptr(d_off), n_moves_total,
                                             _ptr(d_priors), _ptr(d_value), stream or None), "sc_eval_device")

    def encode_only(self, positions: np.ndarray):
        n = len(positions)
        positions = np.ascontiguousarray(positions, dtype=POSITION_DTYPE)
        planes = np.zeros((n, 8, 8, SC_N_PLANES), dtype=np.int8)
        meta = np.zeros((n, SC_N_META), dtype=np.int32)
        _check(load_library().sc_encode_only(self._h, n, _ptr(positions), _ptr(planes), _ptr(meta)), "sc_encode_only")
        return planes, meta

    def move_index_only(self, positions: np.ndarray, moves: np.ndarray, move_off: np.ndarray):
        n = len(positions)
        positions = np.ascontiguousarray(positions, dtype=POSITION_DTYPE)
        moves = np.ascontiguousarray(moves, dtype=MOVE_DTYPE)
        move_off = np.ascontiguousarray(move_off, dtype=np.int32)
        out = np.full(int(move_off[n]) if n else 0, -7, dtype=np.int32)
        _check(load_library().sc_move_index_only(self._h, n, _ptr(positions), _ptr(moves), _ptr(move_off), _ptr(out)),
               "sc_move_index_only")
        return out

    def forward_only(self, planes: np.ndarray, meta: np.ndarray):
        """planes float32 [n,112,8,8], meta float32 [n,7] -> (logp [n,4672], value [n])."""
        planes = np.ascontiguousarray(planes, dtype=np.float32)
        meta = np.ascontiguousarray(meta, dtype=np.float32)
        n = planes.shape[0]
        logp = np.zeros((n, SC_N_POLICY), dtype=np.float32)
        value = np.zeros(n, dtype=np.float32)
        _check(load_library().sc_forward_only(self._h, n, _ptr(planes), _ptr(meta), _ptr(logp), _ptr(value)),
               "sc_forward_only")
        return logp, value

    def launch_count(self) -> int:
        return int(load_library().sc_launch_count(self._h))

    def set_timing(self, level: int):
        _check(load_library().sc_set_timing(self._h, int(level)), "sc_set_timing")

    def kernel_timing(self):
        """(average ms of a 3x3 tower conv launch in the last call, number of launches)"""
        a, n = C.c_float(), C.c_int()
        _check(load_library().sc_kernel_timing(self._h, C.byref(a), C.byref(n)), "sc_kernel_timing")
        return a.value, n.value

    def last_timing(self):
        a, b = C.c_float(), C.c_float()
        _check(load_library().sc_last_timing(self._h, C.byref(a), C.byref(b)), "sc_last_timing")
        return a.value, b.value
