"""CPU tests of the product's host side (native rules + batched search driver) against the oracle.
The driver runs with its position-hash stand-in evaluator (a test hook with the same specification
as the oracle's), so search semantics are compared without network numerics in the loop:
identical visit counts, Q sums, uct values and chosen moves, always."""
import numpy as np
import pytest


def _oracle_selfplay(co, n_plies, rollouts, cpuct):
    t = co.Tree()
    steps = []
    for _ in range(n_plies):
        t.search(rollouts, cpuct)
        mv, n_act, q, u = t.root_children()
        if len(mv) == 0:
            break
        root_q = t.root_q()
        i = t.step_argmax()
        steps.append((co.uci(mv[i]), root_q, [(co.uci(m), int(n), float(qq), float(uu)) for m, n, qq, uu in zip(mv, n_act, q, u)]))
    return steps


def test_hash_selfplay_matches_oracle_sequential_search(co):
    import scb200

    plies, rollouts, cpuct = 24, 60, 2.5
    ref = _oracle_selfplay(co, plies, rollouts, cpuct)
    sp = scb200.SelfPlay(None, n_trees=3, rollout_num=rollouts, num_steps=plies, cpuct=cpuct, with_noise=False,
                         temperature_switch=0, temperature=0.0, evaluator="hash", keep_traces=True, n_threads=0)
    st = sp.run(max_games=3)
    assert st["games_finished"] == 3 and st["moves"] == 3 * plies
    assert st["rollouts"] == 3 * plies * rollouts
    for k in range(3):
        tr = sp.trace(k)
        assert tr["outcome"] is None                     # num_steps reached, no outcome (main.rs:235-238)
        assert len(tr["steps"]) == len(ref)
        for (mv, q, ch), (rmv, rq, rch) in zip(tr["steps"], ref):
            assert mv == rmv
            assert np.float32(q) == np.float32(rq)
            assert [c[0] for c in ch] == [c[0] for c in rch]          # python-chess move order
            assert [c[1] for c in ch] == [c[1] for c in rch]          # visit counts
            assert np.array_equal(np.float32([c[2] for c in ch]), np.float32([c[2] for c in rch]))
            assert np.array_equal(np.float32([c[3] for c in ch]), np.float32([c[3] for c in rch]))
    sp.close()


def test_native_rules_match_oracle_on_sample_games(co, sample_games):
    """The driver's rules engine is independent of the oracle's; replaying the sample games through a
    1-rollout hash search exercises movegen order at every position: with one rollout per move and
    temperature 0 the chosen move is the FIRST legal move, so instead we compare full child lists
    along forced lines: run the driver with 2 rollouts and check every child list against the oracle."""
    import scb200

    sp = scb200.SelfPlay(None, n_trees=1, rollout_num=2, num_steps=120, cpuct=1.0, with_noise=False,
                         temperature_switch=0, temperature=0.0, evaluator="hash", keep_traces=True)
    sp.run(max_games=1)
    tr = sp.trace(0)
    g = co.Game()
    for mv, q, ch in tr["steps"]:
        assert [c[0] for c in ch] == g.legal_uci()
        g.push(mv)
    assert len(tr["steps"]) == 120
    sp.close()


def test_threads_and_tree_count_do_not_change_results(co):
    import scb200

    out = []
    for n_trees, n_threads in ((1, 0), (5, 3)):
        sp = scb200.SelfPlay(None, n_trees=n_trees, rollout_num=40, num_steps=12, cpuct=2.5, with_noise=False,
                             temperature_switch=0, temperature=0.0, evaluator="hash", keep_traces=True,
                             n_threads=n_threads)
        sp.run(max_games=n_trees)
        out.append([sp.trace(k) for k in range(n_trees)])
        sp.close()
    for tr in out[1]:
        assert tr == out[0][0]


def test_game_termination_and_trace_format(co):
    import scb200

    # with temperature 1 and noise the games differ per tree; play until every game ends on its own
    sp = scb200.SelfPlay(None, n_trees=8, rollout_num=8, num_steps=400, cpuct=2.5, with_noise=True, epsilon=0.15,
                         temperature_switch=400, temperature=1.0, evaluator="hash", keep_traces=True, seed=7,
                         n_threads=2)
    st = sp.run(max_games=8)
    assert st["games_finished"] == 8
    assert st["white_wins"] + st["black_wins"] + st["draws"] + st["unfinished"] == 8
    names = {"Checkmate", "Stalemate", "InsufficientMaterial", "SeventyfiveMoves", "FivefoldRepetition", "FiftyMoves",
             "ThreefoldRepetition"}
    seen_moves = set()
    for k in range(8):
        tr = sp.trace(k)
        assert set(tr.keys()) == {"steps", "outcome"}
        g = co.Game()
        for mv, q, ch in tr["steps"]:
            assert mv in g.legal_uci()
            assert sum(c[1] for c in ch) == 8 - 1          # root N = R, children sum to R - 1
            g.push(mv)
        seen_moves.add(tuple(s[0] for s in tr["steps"][:6]))
        if tr["outcome"] is not None:
            assert tr["outcome"]["termination"] in names
            oc = g.outcome(claim_draw=True)
            assert oc is not None
            code = {1: "Checkmate", 2: "Stalemate", 3: "InsufficientMaterial", 4: "SeventyfiveMoves",
                    5: "FivefoldRepetition", 6: "FiftyMoves", 7: "ThreefoldRepetition"}[oc[0]]
            assert code == tr["outcome"]["termination"]
            assert {1: "White", 0: "Black", -1: None}[oc[1]] == tr["outcome"]["winner"]
            # main.rs:223: outcome is only looked at after move index 100 (or when no legal move is left)
            assert len(tr["steps"]) > 100 or len(g.legal_moves()) == 0
    assert len(seen_moves) > 1                             # the per-tree RNG streams differ
    sp.close()


def test_native_rules_probe_matches_oracle_on_real_games(co, sample_games):
    """Every position of 20 of the sample.csv games (castling, promotions, en passant, checks, mates):
    legal moves in the same ORDER, identical packed leaf (bitboards, repetition flags, meta, n_hist),
    identical outcome(claim_draw=True)."""
    import scb200

    for game in sample_games["games"][:20]:
        g = co.Game()
        hist = []
        ucis = game["uci"].split()
        for ply in range(len(ucis) + 1):
            mv, pos, (term, win) = scb200.rules_probe(hist)
            assert np.array_equal(mv, g.legal_moves()), (game["id"], ply)
            slot, meta, nh = g.pack()
            assert np.array_equal(pos["slot"], slot) and np.array_equal(pos["meta"], meta) and pos["n_hist"] == nh
            oc = g.outcome(claim_draw=True)
            assert (term, win) == (oc if oc is not None else (0, -1))
            if ply < len(ucis):
                m = co.parse_uci(ucis[ply])
                hist.append(m)
                g.push(m)


def test_native_rules_probe_random_play_with_repetitions(co):
    import scb200

    rng = np.random.RandomState(5)
    for _ in range(6):
        g = co.Game()
        hist = []
        for ply in range(160):
            mv, pos, oc = scb200.rules_probe(hist)
            ref = g.legal_moves()
            assert np.array_equal(mv, ref)
            slot, meta, nh = g.pack()
            assert np.array_equal(pos["slot"], slot) and np.array_equal(pos["meta"], meta)
            o = g.outcome(claim_draw=True)
            assert oc == (o if o is not None else (0, -1))
            if len(ref) == 0:
                break
            # bias towards shuffling pieces back and forth so that repetitions occur
            m = ref[rng.randint(len(ref))] if rng.rand() < 0.5 or ply < 2 else None
            if m is None:
                back = [x for x in ref if len(hist) >= 2 and x[0] == hist[-2][1] and x[1] == hist[-2][0]]
                m = back[0] if back else ref[rng.randint(len(ref))]
            hist.append((int(m[0]), int(m[1]), int(m[2])))
            g.push(m)
    with pytest.raises(scb200.SCError):
        scb200.rules_probe([(0, 63, 0)])


def test_arena_rounds_hash_evaluator(co):
    """play.rs semantics on the hash evaluator: rounds of n_trees games, outcome checked after every
    ply, 200-ply cap, tally consistent, every trace replays legally and ends where the oracle says."""
    import scb200

    a = scb200.Arena(None, None, n_trees=6, rollout=10, cpuct=1.5, temperature=0.0, temperature_switch=8,
                     max_plies=200, seed=2, evaluator="hash", keep_traces=True, n_threads=2)
    st = a.run(max_games=9)
    assert st["games_finished"] == 9
    assert st["white_wins"] + st["black_wins"] + st["draws"] + st["unfinished"] == 9
    for k in range(9):
        tr = a.trace(k)
        g = co.Game()
        for i, (mv, q, ch) in enumerate(tr["steps"]):
            assert g.outcome(claim_draw=True) is None        # the game would have stopped earlier
            assert mv in g.legal_uci()
            assert sum(c[1] for c in ch) == 9
            g.push(mv)
        oc = g.outcome(claim_draw=True)
        if tr["outcome"] is None:
            assert oc is None and len(tr["steps"]) == 200
        else:
            assert oc is not None and {1: "White", 0: "Black", -1: None}[oc[1]] == tr["outcome"]["winner"]
    a.close()
    assert abs(scb200.elo(200, 120, 60) - 107.538) < 1e-2     # scripts/elo.py


def test_random_positions_generator_matches_oracle_encoding(co):
    """bench.py's workload generator (native rules) emits legal, well-formed leaves: replaying its games
    is impossible from outside, so check each leaf's packed history against itself -- n_hist grows 1..8,
    slot 0 meta consistent, moves are exactly the legal moves of the position rebuilt from slot 0."""
    import scb200

    pos, moves, off = scb200.random_positions(400, seed=3, max_ply=60)
    assert len(pos) == 400 and off[0] == 0 and off[-1] == len(moves)
    assert pos["n_hist"].min() == 1 and pos["n_hist"].max() == 8
    # plies restart at 60: meta fullmove stays small, both colours appear
    assert pos["meta"][:, 1].max() <= 31 and set(pos["meta"][:, 0].tolist()) == {0, 1}
    # first leaf is the start position with the 20 standard moves in python-chess order
    g = co.Game()
    first = [(int(m["from"]), int(m["to"]), int(m["promo"])) for m in moves[off[0]:off[1]]]
    assert first == [tuple(int(x) for x in m) for m in g.legal_moves()]
    slot, meta, nh = g.pack()
    assert np.array_equal(pos["slot"][0], slot) and np.array_equal(pos["meta"][0], meta)
    # determinism
    p2, m2, o2 = scb200.random_positions(400, seed=3, max_ply=60)
    assert np.array_equal(p2, pos) and np.array_equal(o2, off)


def test_root_noise_sampler_moments():
    """Dirichlet(0.3) root noise (mcts.rs:123-130): components sum to 1, mean 1/n, and the variance of a
    component is (1/n)(1-1/n)/(n*alpha+1)."""
    import ctypes

    import scb200

    L = scb200.load_library()
    n, alpha, reps = 30, 0.3, 4000
    xs = np.zeros((reps, n), dtype=np.float32)
    for r in range(reps):
        assert L.sc_test_dirichlet(1000 + r, ctypes.c_float(alpha), n, xs[r].ctypes.data) == 0
    assert np.allclose(xs.sum(1), 1.0, atol=1e-4) and (xs >= 0).all()
    assert abs(xs.mean() - 1.0 / n) < 1e-6
    var_expected = (1 / n) * (1 - 1 / n) / (n * alpha + 1)
    assert abs(xs.var(axis=0).mean() / var_expected - 1.0) < 0.1
    # same shape as numpy's Dirichlet(0.3): compare a tail statistic (the largest component)
    ref = np.random.RandomState(0).dirichlet([alpha] * n, size=reps)
    assert abs(np.median(xs.max(1)) - np.median(ref.max(1))) < 0.02
    assert abs(np.mean(xs < 1e-3) - np.mean(ref < 1e-3)) < 0.02


def test_virtual_loss_mode_invariants(co):
    """leaves_per_tree > 1 is not the reference's sequential search (in-flight paths carry a virtual
    loss), so it is checked through invariants: every move still gets exactly rollout_num rollouts,
    no virtual visit survives into the recorded statistics, |q| <= n, the children are the legal moves
    in python-chess order, and the run is deterministic and independent of threads."""
    import scb200

    plies, rollouts = 16, 50

    def run(k, threads, trees=4):
        sp = scb200.SelfPlay(None, n_trees=trees, rollout_num=rollouts, num_steps=plies, cpuct=2.5, with_noise=False,
                             temperature_switch=0, temperature=0.0, evaluator="hash", keep_traces=True,
                             n_threads=threads, leaves_per_tree=k)
        st = sp.run(max_games=trees)
        tr = [sp.trace(i) for i in range(trees)]
        sp.close()
        return st, tr

    st1, tr1 = run(1, 0)
    for k in (2, 4, 8):
        st, tr = run(k, 0)
        assert st["games_finished"] == 4 and st["moves"] == 4 * plies
        assert st["rollouts"] == 4 * plies * rollouts
        for t in tr:
            g = co.Game()
            for mv, q, ch in t["steps"]:
                assert [c[0] for c in ch] == g.legal_uci()
                assert sum(c[1] for c in ch) == rollouts - 1        # first rollout expands the root
                assert all(c[1] >= 0 and abs(c[2]) <= c[1] + 1e-4 for c in ch)
                g.push(mv)
        st_b, tr_b = run(k, 3)
        assert tr_b == tr
        # the search differs from K = 1 only through the order in which leaves are evaluated
        same = sum(a["steps"][0][0] == b["steps"][0][0] for a, b in zip(tr1, tr))
        assert same >= 1


def test_game_interface_mirror_equals_batched_driver_and_oracle(co):
    """csrc/host/game.hpp mirrors the reference's `Game::predict` / `mcts::mcts` / `mcts::step` interface
    (one leaf per predict, predict at every level).  With the same stand-in evaluator it must produce
    the trace of the batched driver and of the oracle's sequential search, bit for bit — including a
    game with temperature sampling, where both draw one uniform number per move from the same stream."""
    import scb200

    plies, rollouts, cpuct = 24, 60, 2.5
    ref = _oracle_selfplay(co, plies, rollouts, cpuct)
    tr = scb200.game_selfplay(None, rollout_num=rollouts, num_steps=plies, cpuct=cpuct, evaluator="hash")
    assert tr["outcome"] is None and len(tr["steps"]) == len(ref)
    for (mv, q, ch), (rmv, rq, rch) in zip(tr["steps"], ref):
        assert mv == rmv and np.float32(q) == np.float32(rq)
        assert [(c[0], c[1]) for c in ch] == [(c[0], c[1]) for c in rch]
        assert np.array_equal(np.float32([c[2] for c in ch]), np.float32([c[2] for c in rch]))
        assert np.array_equal(np.float32([c[3] for c in ch]), np.float32([c[3] for c in rch]))
    for seed, tswitch in ((0, 0), (5, 6)):
        sp = scb200.SelfPlay(None, n_trees=1, rollout_num=40, num_steps=130, cpuct=1.7, with_noise=False,
                             temperature_switch=tswitch, temperature=0.0, evaluator="hash", keep_traces=True, seed=seed)
        sp.run(max_games=1)
        batched = sp.trace(0)
        sp.close()
        mirror = scb200.game_selfplay(None, rollout_num=40, num_steps=130, cpuct=1.7, temperature_switch=tswitch,
                                      seed=seed, evaluator="hash")
        assert mirror == batched
    with pytest.raises(scb200.SCError):
        scb200.game_selfplay(None, rollout_num=-1, evaluator="hash")


def test_rollout_factor_follows_the_reference_cli():
    """`--rollout-factor v` (src/main.rs:175-180): the rollouts of a move are min(300, (legal moves at the root * v) as
    i32); both given is an error (the reference panics); neither = 300.  Batched driver and one-leaf mirror agree."""
    import scb200

    for v in (1.5, 20.0):
        sp = scb200.SelfPlay(None, n_trees=1, rollout_num=0, rollout_factor=v, num_steps=12, cpuct=2.0, with_noise=False,
                             temperature_switch=0, temperature=0.0, evaluator="hash", keep_traces=True, seed=3)
        st = sp.run(max_games=1)
        tr = sp.trace(0)
        sp.close()
        assert st["moves"] == 12
        for mv, q, ch in tr["steps"]:
            want = min(300, int(np.float32(len(ch)) * np.float32(v)))
            assert sum(c[1] for c in ch) == want - 1          # the first rollout expands the (reset) root
        cfgd = dict(rollout_num=0, num_steps=12, cpuct=2.0, temperature_switch=0, seed=3, evaluator="hash")
        L = scb200.binding.load_library()
        import ctypes as C
        cfg = scb200.binding.SelfPlayConfig(1, 0, 12, 2.0, 0.15, 0, 0, 0.0, 3, 1, 1, 1, 1, 1, v)
        buf = C.create_string_buffer(1 << 22)
        assert L.sc_game_selfplay(None, C.byref(cfg), buf, 1 << 22) > 0
        import json
        assert json.loads(buf.value.decode()) == tr
    with pytest.raises(scb200.SCError):
        scb200.SelfPlay(None, n_trees=1, rollout_num=10, rollout_factor=2.0, evaluator="hash")
    sp = scb200.SelfPlay(None, n_trees=1, rollout_num=0, num_steps=1, with_noise=False, temperature_switch=0,
                         evaluator="hash", keep_traces=True)
    sp.run(max_games=1)
    assert sum(c[1] for c in sp.trace(0)["steps"][0][2]) == 299
    with pytest.raises(scb200.SCError):
        sp.run(max_games=1)                                   # one run per driver object
    sp.close()


def test_native_rules_perft_matches_published_tables():
    """The product's own move generator (csrc/host/chess_rules.hpp, independent of the oracle's) against the
    published perft tables: start position to depth 5, Kiwipete and four more standard test positions
    (castling through check, en-passant pins, promotions with check, discovered checks)."""
    import scb200

    assert [scb200.rules_perft(None, d) for d in range(1, 6)] == [20, 400, 8902, 197281, 4865609]
    cases = [
        ("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", [48, 2039, 97862, 4085603]),
        ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", [14, 191, 2812, 43238, 674624]),
        ("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", [6, 264, 9467, 422333]),
        ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", [44, 1486, 62379, 2103487]),
        ("r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", [46, 2079, 89890, 3894594]),
        ("1k1r4/1r5p/p4n1P/1ppP1P2/PP6/4PP1b/3B4/R1N1K3 b - - 0 39", None),     # src/chess_fast.rs:84-98
    ]
    for fen, exp in cases:
        if exp is None:
            assert scb200.rules_perft(fen, 1) > 0
            continue
        assert [scb200.rules_perft(fen, d + 1) for d in range(len(exp))] == exp, fen
    with pytest.raises(scb200.SCError):
        scb200.rules_perft("not a fen", 1)


def test_auto_leaves_per_tree_fills_the_tail(co):
    """leaves_per_tree = -1: one leaf per tree while the trees fill the batch; when the run's games run out, the trees
    still playing share the rows.  12 games on 8 trees: the last 4 run with two leaves each.  Same invariants as the
    fixed multi-leaf mode, and fewer batches than the one-leaf mode needs for the same games."""
    import scb200

    def run(k):
        sp = scb200.SelfPlay(None, n_trees=8, rollout_num=40, num_steps=10, cpuct=2.5, with_noise=False,
                             temperature_switch=0, temperature=0.0, evaluator="hash", keep_traces=True, leaves_per_tree=k,
                             pipeline_groups=1)
        st = sp.run(max_games=12)
        tr = [sp.trace(i) for i in range(12)]
        sp.close()
        return st, tr

    st1, tr1 = run(1)
    sta, tra = run(-1)
    for st in (st1, sta):
        assert st["games_finished"] == 12 and st["moves"] == 120 and st["rollouts"] == 120 * 40
    assert tra[:8] == tr1[:8]                       # the first 8 games never share rows: identical to the one-leaf mode
    for t in tra:
        g = co.Game()
        for mv, q, ch in t["steps"]:
            assert [c[0] for c in ch] == g.legal_uci() and sum(c[1] for c in ch) == 39
            g.push(mv)


def test_host_search_is_clean_under_thread_sanitizer():
    """`make tsan`: the batched driver (self-play with 1 and 4 leaves per tree, arena; 8 worker threads, stand-in
    evaluator) built with -fsanitize=thread; any data race makes the harness exit non-zero.  (The reference's tree is
    `unsafe impl Send` and single-threaded, src/mcts.rs:26.)"""
    import os
    import shutil
    import subprocess

    if not shutil.which("g++"):
        pytest.skip("no g++")
    csrc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "smart-chess-rust_b200", "csrc")
    r = subprocess.run(["make", "-C", csrc, "tsan"], capture_output=True, text=True, timeout=600)
    if r.returncode != 0 and "unrecognized" in r.stderr and "sanitize" in r.stderr:
        pytest.skip("toolchain without ThreadSanitizer")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "tsan harness ok" in r.stdout and "WARNING: ThreadSanitizer" not in r.stderr


def test_trace_game_numbers_are_a_permutation_of_the_started_games():
    """sc_selfplay_trace_game: every kept trace knows which game of the run it is (start order), so files can be named
    by game instead of by finishing order (scripts/run_batch writes trace{k}.json for game k)."""
    import scb200

    sp = scb200.SelfPlay(None, n_trees=4, rollout_num=8, num_steps=12, evaluator="hash", keep_traces=True,
                         temperature_switch=12, temperature=1.0, seed=2, n_threads=2)
    st = sp.run(max_games=10)
    ids = [sp.trace_game(k) for k in range(st["games_finished"])]
    assert st["games_finished"] == 10 and sorted(ids) == list(range(10))
    assert sp.trace_game(10) == -1 and sp.trace_game(-1) == -1
    sp.close()
