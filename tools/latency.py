"""Latency of one sc_eval call (host buffers in, host buffers out) against the batch size, with the latency kernel
(default) and with SCB200_LATENCY=0 (throughput kernel only).  python tools/latency.py [--blocks 19]"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "smart-chess-rust_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=19)
    ap.add_argument("--iters", type=int, default=200)
    a = ap.parse_args()
    import numpy as np
    import scb200

    tmp = tempfile.mkdtemp()
    blob = os.path.join(tmp, "w.scw")
    scb200.write_blob(scb200.random_init_state_dict(a.blocks, 0), blob)
    pos, moves, off = scb200.random_positions(512, seed=3)
    out = {}
    for lat in ("1", "0"):
        os.environ["SCB200_LATENCY"] = lat
        e = scb200.Engine(blob, 0, scb200.SC_MODE_BF16, 512)
        e.set_timing(1)
        res = {}
        for n in (1, 2, 4, 8, 16, 32, 36, 48, 64, 74, 96, 128, 256, 512):
            pri = np.zeros(int(off[n]), np.float32)
            val = np.zeros(n, np.float32)
            for it in range(a.iters + 20):
                if it == 20:
                    t0 = time.perf_counter()
                e.eval(pos[:n], moves[: off[n]], off[: n + 1], pri, val)
            ms = (time.perf_counter() - t0) / a.iters * 1e3
            tower_ms, total_ms = e.last_timing()
            res[n] = {"call_ms": round(ms, 4), "tower_ms": round(tower_ms, 4), "kernels_ms": round(total_ms, 4)}
        e.close()
        out["latency_kernel" if lat == "1" else "throughput_kernel_only"] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
