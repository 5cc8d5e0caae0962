"""Writes tests/golden/search_cfg1.json: the visit-count tables of BASELINE configs[0]
(`selfplay --rollout-num 20 --num-steps 150 --cpuct 2.5`, noise off, temperature 0) played by the
sequential CPU search (oracle restatement of src/mcts.rs) with the reference's OWN network
(/root/reference/py/module.py, load_model(n_res_blocks=19) seed-0 init, fp32, batch 1 per predict
as src/backends/torch.rs:119 does).

Run in the build container only (needs /root/reference):  python oracle/make_golden_search.py
The GPU test replays the same game through the batched driver + CUDA engine in fp32 mode and compares
every table (tests/test_gpu_selfplay.py).
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import chess_oracle as co  # noqa: E402
import net  # noqa: E402
from make_golden_net import import_reference_module  # noqa: E402

ROLLOUTS, PLIES, CPUCT = 20, 150, 2.5


def main():
    module = import_reference_module()
    model = module.load_model(n_res_blocks=19, device="cpu", compile=False)
    digest = net.state_dict_digest(model.state_dict())

    def ev(game, depth, moves):
        planes, meta = game.encode(depth)
        with torch.no_grad():
            lp, v = model(net.planes_i8_hwc_to_nchw(planes[None]), torch.from_numpy(meta[None]).float())
        pri = co.post_process(lp[0].numpy(), game.move_indices(moves))
        return pri, float(v[0, 0])

    t = co.Tree(ev)
    steps = []
    t0 = time.time()
    for ply in range(PLIES):
        t.search(ROLLOUTS, CPUCT)
        mv, n_act, q, u = t.root_children()
        if len(mv) == 0:
            break
        root_q = t.root_q()
        i = t.step_argmax()
        steps.append({"move": co.uci(mv[i]), "root_q": float(np.float32(root_q)),
                      "children": [[co.uci(m), int(n), float(np.float32(qq))] for m, n, qq in zip(mv, n_act, q)]})
        if ply % 10 == 0:
            print(ply, steps[-1]["move"], "%.0fs" % (time.time() - t0), flush=True)
    out = {"config": {"rollout_num": ROLLOUTS, "num_steps": PLIES, "cpuct": CPUCT, "with_noise": False,
                      "temperature": 0.0, "temperature_switch": 0, "n_res_blocks": 19, "init": "load_model seed 0"},
           "weights_digest": digest, "steps": steps}
    p = os.path.join(HERE, "..", "tests", "golden", "search_cfg1.json")
    with open(p, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", p, len(steps), "plies", os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
