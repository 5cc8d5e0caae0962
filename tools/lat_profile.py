"""SCB200_PHASE_PROFILE=1 python tools/lat_profile.py [n]: phase cycle counters of the latency kernel for one call."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "smart-chess-rust_b200"))
import scb200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
tmp = tempfile.mkdtemp()
blob = os.path.join(tmp, "w.scw")
scb200.write_blob(scb200.random_init_state_dict(19, 0), blob)
pos, moves, off = scb200.random_positions(max(n, 4), seed=3)
e = scb200.Engine(blob, 0, scb200.SC_MODE_BF16, 128)
for _ in range(3):
    e.eval(pos[:n], moves[: off[n]], off[: n + 1])
e.close()
