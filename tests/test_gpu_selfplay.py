"""GPU tests of the batched self-play driver fed by the engine (run with -m gpu)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_net(tmp_path_factory):
    import net
    import scb200

    sd = net.perturb_norm_params(net.init_state_dict(2, 7), 1234)
    p = str(tmp_path_factory.mktemp("w") / "n2.scw")
    scb200.write_blob(sd, p)
    return sd, p


def _oracle_net_evaluator(co, sd):
    import net

    def ev(game, depth, moves):
        planes, meta = game.encode(depth)
        lp, v = net.forward(sd, net.planes_i8_hwc_to_nchw(planes[None]), torch.from_numpy(meta[None]).float())
        pri = co.post_process(lp[0].numpy(), game.move_indices(moves))
        return pri, float(v[0, 0])

    return ev


def test_gpu_selfplay_visit_counts_match_cpu_reference_search(co, small_net):
    """north_star: at temperature 0 with noise off, fp32 mode yields identical MCTS visit counts and
    chosen moves.  CPU side = sequential oracle search (mcts.rs restated) fed by the oracle network;
    GPU side = batched driver + CUDA engine.  A near-tie in uct can legitimately flip (1e-4 gate, not
    bit equality), so the test reports the first divergence and requires the first plies to agree."""
    import scb200

    sd, blob = small_net
    plies, rollouts, cpuct = 6, 24, 2.5
    t = co.Tree(_oracle_net_evaluator(co, sd))
    ref = []
    for _ in range(plies):
        t.search(rollouts, cpuct)
        mv, n_act, q, u = t.root_children()
        i = t.step_argmax()
        ref.append((co.uci(mv[i]), [int(x) for x in n_act], [float(x) for x in q]))
    eng = scb200.Engine(blob, 0, scb200.SC_MODE_FP32, 8)
    sp = scb200.SelfPlay(eng, n_trees=4, rollout_num=rollouts, num_steps=plies, cpuct=cpuct, with_noise=False,
                         temperature_switch=0, temperature=0.0, keep_traces=True, pipeline_groups=2, n_threads=2)
    st = sp.run(max_games=4)
    assert st["games_finished"] == 4 and st["leaf_evals"] > 0 and st["batches"] > 0
    traces = [sp.trace(k) for k in range(4)]
    assert all(tr == traces[0] for tr in traces)           # identical games in every slot / group
    first_div = None
    for ply, ((mv, q, ch), (rmv, rn, rq)) in enumerate(zip(traces[0]["steps"], ref)):
        if [c[1] for c in ch] != rn or mv != rmv:
            first_div = ply
            break
        assert np.allclose([c[2] for c in ch], rq, atol=5e-3)
    print("first divergence ply:", first_div)
    # measured on B200 (rounds 1-2): no divergence over the tested plies; the kernels are deterministic, so anything
    # else is a regression (a legitimate near-tie flip would show up here with its ply)
    assert first_div is None, f"visit counts diverge from the CPU reference search at ply {first_div}"
    sp.close()
    eng.close()


def test_batched_equals_single_tree_same_evaluator(co, small_net):
    """Search semantics must not depend on how many trees share a batch: with the same (GPU, bf16)
    evaluator one tree alone and 64 trees in two pipeline groups play the same game."""
    import scb200

    sd, blob = small_net
    out = []
    for n_trees, groups, thr in ((1, 1, 0), (64, 2, 4)):
        eng = scb200.Engine(blob, 0, scb200.SC_MODE_BF16, 64)
        sp = scb200.SelfPlay(eng, n_trees=n_trees, rollout_num=30, num_steps=10, cpuct=2.5, with_noise=False,
                             temperature_switch=0, temperature=0.0, keep_traces=True, pipeline_groups=groups,
                             n_threads=thr)
        st = sp.run(max_games=n_trees)
        assert st["games_finished"] == n_trees
        out.append([sp.trace(k) for k in range(n_trees)])
        sp.close()
        eng.close()
    for tr in out[1]:
        assert tr == out[0][0]


def test_selfplay_config2_shape_smoke(co, small_net):
    """README recipe shape (cpuct 2.5, temperature-switch 4, noise on) on a small net: games progress,
    counters are consistent, traces replay legally."""
    import scb200

    sd, blob = small_net
    eng = scb200.Engine(blob, 0, scb200.SC_MODE_BF16, 128)
    sp = scb200.SelfPlay(eng, n_trees=256, rollout_num=16, num_steps=30, cpuct=2.5, epsilon=0.15, with_noise=True,
                         temperature_switch=4, temperature=0.0, keep_traces=True, pipeline_groups=2, n_threads=4, seed=3)
    st = sp.run(max_games=256)
    assert st["games_finished"] == 256 and st["moves"] >= 256 * 20
    assert st["rollouts"] == st["leaf_evals"] + st["terminal_evals"]
    openings = set()
    for k in range(0, 256, 17):
        tr = sp.trace(k)
        g = co.Game()
        for mv, q, ch in tr["steps"]:
            assert mv in g.legal_uci()
            assert sum(c[1] for c in ch) == 15
            g.push(mv)
        openings.add(tuple(s[0] for s in tr["steps"][:4]))
    assert len(openings) > 3
    sp.close()
    eng.close()


def test_arena_two_networks(co, small_net, tmp_path):
    """BASELINE configs[4] shape on small nets: two different networks, both colour assignments; each
    ply's batch must come from the network of the side to move."""
    import net
    import scb200

    sd_a, blob_a = small_net
    sd_b = net.perturb_norm_params(net.init_state_dict(2, 11), 99)
    blob_b = str(tmp_path / "b.scw")
    scb200.write_blob(sd_b, blob_b)
    ea = scb200.Engine(blob_a, 0, scb200.SC_MODE_BF16, 64)
    eb = scb200.Engine(blob_b, 0, scb200.SC_MODE_BF16, 64)

    def first_plies(white, black):
        a = scb200.Arena(white, black, n_trees=16, rollout=20, cpuct=1.5, temperature=0.0, temperature_switch=0,
                         max_plies=6, seed=1, keep_traces=True, n_threads=2)
        st = a.run(max_games=16)
        assert st["games_finished"] == 16 and st["moves"] == 16 * 6
        trs = [a.trace(k) for k in range(16)]
        a.close()
        return trs

    def selfplay_root(engine):
        sp = scb200.SelfPlay(engine, n_trees=1, rollout_num=20, num_steps=1, cpuct=1.5, with_noise=False,
                             temperature_switch=0, temperature=0.0, keep_traces=True, pipeline_groups=1)
        sp.run(max_games=1)
        tr = sp.trace(0)
        sp.close()
        return tr["steps"][0][2]

    ab, ba = first_plies(ea, eb), first_plies(eb, ea)
    root_a, root_b = selfplay_root(ea), selfplay_root(eb)
    assert root_a != root_b
    for tr in ab:
        assert tr["steps"][0][2] == root_a           # White's search used network A
    for tr in ba:
        assert tr["steps"][0][2] == root_b           # ... and network B after the swap
    # ply 1 is searched by the other network: same position => same stats as that network's own search
    by_first = {}
    for tr in ab:
        by_first.setdefault(tr["steps"][0][0], []).append(tr["steps"][1][2])
    for tr in first_plies(eb, eb):
        mv = tr["steps"][0][0]
        if mv in by_first:
            assert tr["steps"][1][2] in by_first[mv]
    ea.close()
    eb.close()


def test_virtual_loss_mode_on_engine(co, small_net):
    """leaves_per_tree = 4: 64 trees fill a 256-row batch.  Same invariants as the CPU test, now with the
    engine in the loop and dense batch rows handed out across host threads; the result must not
    depend on the thread count (rows are assigned in a different order, results per row are not)."""
    import scb200

    sd, blob = small_net

    def run(threads):
        eng = scb200.Engine(blob, 0, scb200.SC_MODE_BF16, 256)
        sp = scb200.SelfPlay(eng, n_trees=64, rollout_num=33, num_steps=12, cpuct=2.5, with_noise=False,
                             temperature_switch=0, temperature=0.0, keep_traces=True, pipeline_groups=2,
                             n_threads=threads, leaves_per_tree=4)
        st = sp.run(max_games=64)
        trs = [sp.trace(k) for k in range(64)]
        sp.close()
        eng.close()
        return st, trs

    st, trs = run(1)
    assert st["games_finished"] == 64 and st["moves"] == 64 * 12
    assert st["rollouts"] == 64 * 12 * 33 == st["leaf_evals"] + st["terminal_evals"]
    # one leaf per tree per batch would take rollouts / (32 trees per pipeline group) batches
    assert st["batches"] < 64 * 12 * 33 / 32 / 2
    for tr in trs[::7]:
        g = co.Game()
        for mv, q, ch in tr["steps"]:
            assert [c[0] for c in ch] == g.legal_uci()
            assert sum(c[1] for c in ch) == 32
            assert all(c[1] >= 0 and abs(c[2]) <= c[1] + 1e-4 for c in ch)
            g.push(mv)
    st4, trs4 = run(4)
    assert trs4 == trs

    with pytest.raises(RuntimeError):
        eng = scb200.Engine(blob, 0, scb200.SC_MODE_BF16, 64)
        try:
            scb200.SelfPlay(eng, n_trees=64, rollout_num=8, num_steps=2, pipeline_groups=1, leaves_per_tree=4)
        finally:
            eng.close()


def test_cfg1_game_reproduces_reference_visit_tables(tmp_path):
    """BASELINE configs[0] (`--rollout-num 20 --num-steps 150 --cpuct 2.5`, noise off, temperature 0):
    tests/golden/search_cfg1.json holds every root's visit table of the game the sequential CPU search
    plays with the reference's own 19-block seed-0 network (oracle/make_golden_search.py).  The fp32
    engine behind the batched driver must play the same game with the same tables."""
    import json
    import os

    import net
    import scb200

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "search_cfg1.json")))
    cfg = gold["config"]
    sd = scb200.random_init_state_dict(cfg["n_res_blocks"], 0)
    d = net.state_dict_digest(sd)
    assert d["numel"] == gold["weights_digest"]["numel"] and abs(d["sum"] - gold["weights_digest"]["sum"]) < 1e-6
    blob = str(tmp_path / "seed0.scw")
    scb200.write_blob(sd, blob)
    eng = scb200.Engine(blob, 0, scb200.SC_MODE_FP32, 4)
    sp = scb200.SelfPlay(eng, n_trees=2, rollout_num=cfg["rollout_num"], num_steps=cfg["num_steps"], cpuct=cfg["cpuct"],
                         with_noise=False, temperature_switch=0, temperature=0.0, keep_traces=True, pipeline_groups=2)
    sp.run(max_games=2)
    tr = sp.trace(0)
    assert sp.trace(1) == tr
    sp.close()
    eng.close()
    first_div, max_dq = None, 0.0
    for ply, (got, ref) in enumerate(zip(tr["steps"], gold["steps"])):
        mv, q, ch = got
        assert [c[0] for c in ch] == [c[0] for c in ref["children"]]           # legal moves, python-chess order
        if mv != ref["move"] or [c[1] for c in ch] != [c[1] for c in ref["children"]]:
            first_div = ply
            break
        max_dq = max(max_dq, max(abs(c[2] - r[2]) for c, r in zip(ch, ref["children"])), abs(q - ref["root_q"]))
    print("plies compared:", min(len(tr["steps"]), len(gold["steps"])), "first divergence:", first_div, "max |dQ|:", max_dq)
    assert max_dq < 1e-3
    # measured on B200 (rounds 1-2): all 82 plies identical, max |dQ| 5.2e-6; deterministic kernels -> pinned to that
    assert first_div is None, f"golden configs[0] game diverges at ply {first_div} (max |dQ| so far {max_dq:.2e})"
    if first_div is None:
        assert len(tr["steps"]) == len(gold["steps"])
        # the golden game ends with no legal moves: the driver must have seen the same terminal position
        assert tr["outcome"] is not None or len(gold["steps"]) == cfg["num_steps"]


def test_game_interface_mirror_on_engine(co, small_net, tmp_path):
    """The one-leaf-at-a-time path (`Game::predict` mirror, csrc/host/game.hpp -> sc_eval with n = 1):
    (1) with the fp32 engine and the seed-0 19-block net it replays the configs[0] golden game;
    (2) with the bf16 engine it plays exactly the game of the batched driver (same evaluator, batch
    composition does not change a leaf's result)."""
    import json
    import os
    import time

    import scb200

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "search_cfg1.json")))
    cfg = gold["config"]
    blob19 = str(tmp_path / "seed0.scw")
    scb200.write_blob(scb200.random_init_state_dict(cfg["n_res_blocks"], 0), blob19)
    eng = scb200.Engine(blob19, 0, scb200.SC_MODE_FP32, 1)
    t0 = time.time()
    tr = scb200.game_selfplay(eng, rollout_num=cfg["rollout_num"], num_steps=cfg["num_steps"], cpuct=cfg["cpuct"])
    print("one-leaf path, fp32, 19 blocks: %.1f plies/s" % (len(tr["steps"]) / (time.time() - t0)))
    eng.close()
    assert len(tr["steps"]) == len(gold["steps"])
    for (mv, q, ch), ref in zip(tr["steps"], gold["steps"]):
        assert mv == ref["move"]
        assert [(c[0], c[1]) for c in ch] == [(c[0], c[1]) for c in ref["children"]]
        assert max(abs(c[2] - r[2]) for c, r in zip(ch, ref["children"])) < 1e-3
    assert tr["outcome"] is not None and tr["outcome"]["termination"] in ("Checkmate", "Stalemate")

    sd, blob = small_net
    eng = scb200.Engine(blob, 0, scb200.SC_MODE_BF16, 64)
    sp = scb200.SelfPlay(eng, n_trees=32, rollout_num=30, num_steps=12, cpuct=2.5, with_noise=False, temperature_switch=3,
                         temperature=0.0, keep_traces=True, pipeline_groups=2, seed=0)
    sp.run(max_games=32)
    batched = sp.trace(0)          # tree 0 uses the RNG stream of seed 0, like the mirror
    sp.close()
    mirror = scb200.game_selfplay(eng, rollout_num=30, num_steps=12, cpuct=2.5, temperature_switch=3, seed=0)
    eng.close()
    assert mirror == batched


def test_run_batch_writes_traces_the_reference_pipeline_reads(co, small_net, tmp_path, capsys):
    """`python -m scb200.run_batch` (scripts/run_batch for this backend): trace{k}.json files in the format of
    src/trace.rs that py/dataset.py:47-87 consumes: {"steps": [[uci, q, [[uci, n, q, uct], ...]], ...], "outcome"}."""
    import json

    from scb200 import run_batch

    sd, blob = small_net
    out = tmp_path / "traces"
    rc = run_batch.main(["-c", blob, "-N", "5", "--prefix", str(out), "--trees", "4", "--rollout-num", "12", "-n", "9",
                         "--temperature-switch", "3", "--cpuct", "2.5", "--threads", "2"])
    assert rc == 0
    lines = [l for l in capsys.readouterr().out.splitlines() if l and not l.startswith("#")]
    assert len(lines) == 5 and lines[0].startswith("01, null, num-steps: 9")
    for k in range(1, 6):
        tr = json.load(open(out / f"trace{k}.json"))
        assert set(tr) == {"steps", "outcome"} and len(tr["steps"]) == 9
        g = co.Game()
        for mv, q, ch in tr["steps"]:
            assert isinstance(q, float) and mv in g.legal_uci()
            assert [c[0] for c in ch] == g.legal_uci() and sum(c[1] for c in ch) == 11
            assert all(len(c) == 4 for c in ch)
            g.push(mv)


def test_leader_board_script(co, small_net, tmp_path, capsys):
    """`python -m scb200.leader_board` (scripts/leader-board + show-result + elo.py): both colour assignments,
    trace files per game, the `total / white / black` tallies and the Elo line."""
    import json

    import net
    import scb200
    from scb200 import leader_board

    sd_a, blob_a = small_net
    blob_b = str(tmp_path / "b.scw")
    scb200.write_blob(net.perturb_norm_params(net.init_state_dict(2, 11), 99), blob_b)
    out = tmp_path / "replay"
    rc = leader_board.main(["-W", blob_a, "-B", blob_b, "-N", "6", "--prefix", str(out), "--rollout", "10", "--max-plies", "40",
                            "--temperature-switch", "4", "--threads", "2", "--trees", "6"])
    assert rc == 0
    text = capsys.readouterr().out
    assert "Swapping the players" in text and text.count(" / ") >= 2
    for tag in ("w", "b"):
        for k in range(1, 7):
            tr = json.load(open(out / f"{tag}_{k}.json"))
            g = co.Game()
            for mv, q, ch in tr["steps"]:
                assert mv in g.legal_uci()
                g.push(mv)
            assert len(tr["steps"]) <= 40


def test_two_engines_two_drivers_one_process(co, small_net):
    """In-process multi-GPU (north_star: one worker per GPU, each with its own tree pool and CUDA streams; reference
    analogue scripts/run_batch:21): N engines on N devices, N drivers on N host threads of ONE process
    (sc_selfplay_run_many), results identical to the runs of the same seeds one after another on one GPU.
    With a single visible GPU both engines live on device 0 (still two engines, two streams, two host threads)."""
    import scb200

    sd, blob = small_net
    ndev = torch.cuda.device_count()
    devs = [0, 1 % ndev]
    kw = dict(n_trees=16, rollout_num=24, num_steps=8, cpuct=2.5, with_noise=True, temperature_switch=3, temperature=0.0,
              keep_traces=True, pipeline_groups=2, n_threads=2)

    def traces(sp, n):
        return sorted(str(sp.trace(k)) for k in range(n))

    solo = []
    for i, seed in enumerate((11, 12)):
        eng = scb200.Engine(blob, 0, scb200.SC_MODE_BF16, 16)
        sp = scb200.SelfPlay(eng, seed=seed, **kw)
        st = sp.run(max_games=16)
        assert st["games_finished"] == 16
        solo.append(traces(sp, 16))
        sp.close()
        eng.close()
    engs = [scb200.Engine(blob, d, scb200.SC_MODE_BF16, 16) for d in devs]
    assert [e.info()["device"] for e in engs] == devs
    sps = [scb200.SelfPlay(e, seed=seed, **kw) for e, seed in zip(engs, (11, 12))]
    stats = scb200.SelfPlay.run_many(sps, max_games=16)
    assert [s["games_finished"] for s in stats] == [16, 16]
    for sp, want in zip(sps, solo):
        assert traces(sp, 16) == want
    print("in-process drivers on devices", devs, "leaf evals", [s["leaf_evals"] for s in stats])
    for sp in sps:
        sp.close()
    for e in engs:
        e.close()


def test_non_finite_leaf_drops_only_its_own_game(co, tmp_path):
    """Per-game failure isolation: a network that returns NaN for SOME positions must cost exactly the games that
    run into such a leaf; the others finish.  The NaN is built from finite weights: hidden unit 0 of the value FC
    gets the weight 4.5e37 on the half-move clock and the weight 0 in the last layer, so its pre-activation overflows
    to +inf once the clock reaches 8 (8 quiet plies in a row) and 0 * inf = NaN; below that 0 * finite = 0."""
    import net
    import scb200

    sd = net.perturb_norm_params(net.init_state_dict(2, 7), 1234)
    w = sd["value_head.ffn.0.weight"].clone()
    w[0, 64 * 256 + 6] = 4.5e37
    sd["value_head.ffn.0.weight"] = w
    w2 = sd["value_head.ffn.2.weight"].clone()
    w2[0, 0] = 0.0
    sd["value_head.ffn.2.weight"] = w2
    blob = str(tmp_path / "nan.scw")
    scb200.write_blob(sd, blob)
    eng = scb200.Engine(blob, 0, scb200.SC_MODE_BF16, 64)
    sp = scb200.SelfPlay(eng, n_trees=64, rollout_num=16, num_steps=24, cpuct=2.5, with_noise=True, temperature_switch=24,
                         temperature=1.0, keep_traces=True, pipeline_groups=2, n_threads=2, seed=5)
    st = sp.run(max_games=64)
    print("dropped", st["games_dropped"], "finished", st["games_finished"])
    assert st["games_dropped"] > 0 and st["games_finished"] > 0
    assert st["games_dropped"] + st["games_finished"] == 64
    for k in range(st["games_finished"]):
        tr = sp.trace(k)
        assert all(np.isfinite(s[1]) and all(np.isfinite(c[2]) for c in s[2]) for s in tr["steps"])
    sp.close()
    eng.close()
