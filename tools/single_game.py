"""ONE self-play game at 180 rollouts/move (a single search tree) against leaves_per_tree: with K > 1 the tree puts up to
K leaves into one device batch (virtual loss), i.e. the batch sizes of the small-batch latency kernel.
python tools/single_game.py [--plies 24]"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "smart-chess-rust_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plies", type=int, default=24)
    a = ap.parse_args()
    import scb200

    tmp = tempfile.mkdtemp()
    blob = os.path.join(tmp, "w.scw")
    scb200.write_blob(scb200.random_init_state_dict(19, 0), blob)
    out = {}
    for k in (1, 4, 16, 32, 64):
        eng = scb200.Engine(blob, 0, scb200.SC_MODE_BF16, 64)
        sp = scb200.SelfPlay(eng, n_trees=1, leaves_per_tree=k, rollout_num=180, num_steps=a.plies, cpuct=2.5, with_noise=True,
                             temperature_switch=4, n_threads=1, pipeline_groups=1, seed=3)
        st = sp.run(max_games=1)
        out[k] = {"leaf_evals_per_s": round(st["leaf_evals"] / st["seconds"]), "plies_per_s": round(st["moves"] / st["seconds"], 1),
                  "ms_per_batch": round(st["seconds"] / max(st["batches"], 1) * 1e3, 3)}
        sp.close()
        eng.close()
    print(json.dumps({"what": "one game, one tree, 180 rollouts/move, 19 blocks bf16, by leaves_per_tree", "result": out}))


if __name__ == "__main__":
    main()
