#!/bin/bash
# BASELINE.json configs[2] and configs[4] run literally through the library's orchestration (one GPU):
#   [2] self-play 500 games, rollout-num 180, temperature-switch 4, 2048 concurrent trees  (scb200.run_batch = scripts/run_batch)
#   [4] leader-board: two random-init nets, rollout-num 100, temperature-switch 8, both colours (scb200.leader_board)
# usage (under gpurun, from the repo root): bash tools/run_configs.sh > gpurun_out/configs.log
set -e
export PYTHONPATH=$PWD/smart-chess-rust_b200
T=$(mktemp -d)
python - <<PY
import scb200
scb200.write_blob(scb200.random_init_state_dict(19, 0), "$T/seed0.scw")
scb200.write_blob(scb200.random_init_state_dict(19, 1), "$T/seed1.scw")
PY
wall() { python -c "import time,sys; print('wall %.1f s' % (time.time()-float(sys.argv[1])))" $1; }
for K in 1 4; do
  echo "== configs[2]: 500 games, leaves per tree $K"
  T0=$(date +%s.%N)
  python -m scb200.run_batch -c $T/seed0.scw -N 500 --prefix $T/traces_$K --rollout-num 180 --temperature-switch 4 --cpuct 2.5 -n 150 --trees 2048 --leaves-per-tree $K 2>&1 | tail -1
  wall $T0
  python - <<PY
import json, glob, collections
c = collections.Counter(); plies = 0; files = sorted(glob.glob("$T/traces_$K/trace*.json"))
for f in files:
    tr = json.load(open(f)); plies += len(tr["steps"])
    c[(tr["outcome"] or {}).get("termination", "none(150 plies)")] += 1
print(len(files), "trace files,", plies, "plies;", dict(c))
PY
done
echo "== configs[4]: leader-board, 2 x 1024 games on one GPU"
T0=$(date +%s.%N)
python -m scb200.leader_board -W $T/seed0.scw -B $T/seed1.scw -N 1024 --prefix $T/replay --rollout 100 --temperature-switch 8 --cpuct 1.5 2>&1 | grep -v "json," | tail -4
wall $T0
ls $T/replay | wc -l
rm -rf $T
