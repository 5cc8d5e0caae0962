"""Layer-by-layer comparison of the latency kernel with the throughput kernel (sc_debug_tower)."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "smart-chess-rust_b200"))
import numpy as np
import scb200

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tmp = tempfile.mkdtemp()
blob = os.path.join(tmp, "w.scw")
scb200.write_blob(scb200.random_init_state_dict(blocks, 0), blob)
pos, moves, off = scb200.random_positions(n, seed=3)
e = scb200.Engine(blob, 0, scb200.SC_MODE_BF16, 16)


def f32(a):
    return (a.astype(np.uint32) << 16).view(np.float32)


for k in range(1, 2 * blocks + 4):
    a = e.debug_tower(pos, k, 0)
    b = e.debug_tower(pos, k, 1)
    msg = []
    for name, u, v in zip("xty", a, b):
        d = np.abs(f32(u) - f32(v))
        bad = np.argwhere(u != v)
        msg.append(f"{name}: diff {int((u != v).sum())}/{u.size} max {np.nanmax(d):.3g}" + (f" first {bad[0].tolist()}" if len(bad) else ""))
    print(f"layers={k}: " + " | ".join(msg))
    if k <= 3:
        for name, u, v in zip("xty", a, b):
            if (u != v).any():
                bad = np.argwhere(u != v)
                print("   boards", np.unique(bad[:, 0]), "squares", np.unique(bad[:, 1])[:16], "channels", np.unique(bad[:, 2])[:40])
                i = tuple(bad[0])
                print("   ref", f32(u)[i[0], i[1], i[2] // 8 * 8: i[2] // 8 * 8 + 8], "\n   lat", f32(v)[i[0], i[1], i[2] // 8 * 8: i[2] // 8 * 8 + 8])
e.close()
