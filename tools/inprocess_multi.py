"""In-process multi-GPU self-play (north_star: one worker per GPU, each with its own tree pool and CUDA streams, no
torchrun): one engine + one batched driver per visible GPU, all driven from host threads of THIS process through
sc_selfplay_run_many.  Prints one JSON line with the aggregate and the per-GPU rates, and the single-GPU rate of the
same configuration for comparison.   python tools/inprocess_multi.py [--moves 8192]"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "smart-chess-rust_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--moves", type=int, default=8192, help="plies per GPU")
    ap.add_argument("--trees", type=int, default=2048)
    a = ap.parse_args()
    import torch

    import scb200

    ndev = torch.cuda.device_count()
    tmp = tempfile.mkdtemp()
    blob = os.path.join(tmp, "w.scw")
    scb200.write_blob(scb200.random_init_state_dict(19, 0), blob)
    threads = max(1, (os.cpu_count() or 8) // ndev)
    kw = dict(n_trees=a.trees, rollout_num=180, num_steps=150, cpuct=2.5, epsilon=0.15, with_noise=True,
              temperature_switch=4, temperature=0.0, n_threads=threads, pipeline_groups=2)

    def run(devs):
        engs = [scb200.Engine(blob, d, scb200.SC_MODE_BF16, a.trees) for d in devs]
        sps = [scb200.SelfPlay(e, seed=100 + d, **kw) for e, d in zip(engs, devs)]
        stats = scb200.SelfPlay.run_many(sps, max_moves=a.moves)
        for sp in sps:
            sp.close()
        for e in engs:
            e.close()
        secs = max(s["seconds"] for s in stats)
        return {"gpus": len(devs), "leaf_evals_per_s": sum(s["leaf_evals"] for s in stats) / secs,
                "plies_per_s": sum(s["moves"] for s in stats) / secs, "seconds": secs,
                "per_gpu_leaf_evals_per_s": [s["leaf_evals"] / s["seconds"] for s in stats],
                "device_wait_frac": [s["wait_seconds"] / s["seconds"] for s in stats]}

    one = run([0])
    allg = run(list(range(ndev)))
    print(json.dumps({"what": "in-process multi-GPU self-play (sc_selfplay_run_many), 180 rollouts/move, %d trees per GPU, "
                              "%d host threads per GPU" % (a.trees, threads),
                      "single_gpu": one, "all_gpus": allg,
                      "scaling_efficiency": allg["leaf_evals_per_s"] / (one["leaf_evals_per_s"] * ndev)}))


if __name__ == "__main__":
    main()
