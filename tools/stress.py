"""Soak test on one GPU: long self-play / arena runs in every driver mode; checks counters and that nothing hangs."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "smart-chess-rust_b200"))
import scb200

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
tmp = tempfile.mkdtemp()
blobs = []
for seed in (0, 1):
    p = os.path.join(tmp, f"s{seed}.scw")
    scb200.write_blob(scb200.random_init_state_dict(19, seed), p)
    blobs.append(p)
ea = scb200.Engine(blobs[0], 0, scb200.SC_MODE_BF16, 2048)
eb = scb200.Engine(blobs[1], 0, scb200.SC_MODE_BF16, 2048)
for name, mk in (
    ("selfplay K=1", lambda: scb200.SelfPlay(ea, n_trees=2048, rollout_num=180, num_steps=150, temperature_switch=4, n_threads=8)),
    ("selfplay K=4", lambda: scb200.SelfPlay(ea, n_trees=512, rollout_num=180, num_steps=150, temperature_switch=4, n_threads=8, leaves_per_tree=4)),
    ("selfplay short games", lambda: scb200.SelfPlay(ea, n_trees=1500, rollout_num=12, num_steps=40, temperature_switch=4, n_threads=8, keep_traces=True)),
    ("arena", lambda: scb200.Arena(ea, eb, n_trees=2048, rollout=100, n_threads=8)),
    ("arena short", lambda: scb200.Arena(ea, eb, n_trees=777, rollout=10, max_plies=30, n_threads=8, keep_traces=True)),
):
    sp = mk()
    t0 = time.time()
    st = sp.run(max_seconds=secs)
    dt = time.time() - t0
    inflight = st["leaf_evals"] + st["terminal_evals"] - st["rollouts"]   # leaves submitted when the clock ran out
    assert 0 <= inflight <= 2048, st
    assert dt < secs + 10, (name, dt)
    print(f"{name}: {st['leaf_evals'] / st['seconds']:.0f} leaf evals/s, {st['moves']} plies, {st['games_finished']} games, "
          f"W/B/D {st['white_wins']}/{st['black_wins']}/{st['draws']}, batches {st['batches']}", flush=True)
    tr = sp.trace(0) if st["games_finished"] else None      # None unless the run keeps traces
    assert tr is None or len(tr["steps"]) > 0
    sp.close()
ea.close(); eb.close()
print("stress ok")
