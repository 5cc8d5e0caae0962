// ThreadSanitizer harness for the host-side search (`make -C smart-chess-rust_b200/csrc tsan`).
//
// The reference marks its tree `unsafe impl Send` (src/mcts.rs:26) and never runs it on more than one thread; the
// batched driver walks thousands of trees with a worker pool (csrc/host/search.cpp), so the sharing discipline is
// checked here: self-play and arena runs with the stand-in (position-hash) evaluator, 8 worker threads, under TSAN.
// The engine entry points search.cpp links against are stubbed -- no GPU and no CUDA call is involved.
#include <cstdio>
#include <cstring>
#include <string>

#include "../common.cuh"

namespace scb {
static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
}  // namespace scb

extern "C" {
const char *sc_last_error(void) { return scb::g_err.c_str(); }
int sc_eval(sc_engine *, int, const sc_position *, const sc_move *, const int32_t *, float *, float *, void *) { return SC_E_NOGPU; }
int sc_eval_submit(sc_engine *, int, const sc_position *, const sc_move *, const int32_t *, float *, float *, void *, int *)
{
    return SC_E_NOGPU;
}
int sc_eval_wait(sc_engine *, int) { return SC_E_NOGPU; }
int sc_info(const sc_engine *, int *, int *, int *) { return SC_E_NOGPU; }
int sc_device_info(const sc_engine *, int *, int *) { return SC_E_NOGPU; }
}

static int run(bool arena, int threads, int leaves_per_tree)
{
    sc_selfplay_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.n_trees = 96;
    cfg.rollout_num = 24;
    cfg.num_steps = 40;
    cfg.cpuct = 2.5f;
    cfg.epsilon = 0.15f;
    cfg.with_noise = 1;
    cfg.temperature_switch = 6;
    cfg.seed = 9;
    cfg.n_threads = threads;
    cfg.evaluator = 1;
    cfg.pipeline_groups = 2;
    cfg.keep_traces = 1;
    cfg.leaves_per_tree = leaves_per_tree;
    sc_selfplay *sp = nullptr;
    int rc = arena ? sc_arena_create(nullptr, nullptr, &cfg, &sp) : sc_selfplay_create(nullptr, &cfg, &sp);
    if (rc != SC_OK) {
        fprintf(stderr, "create failed: %s\n", sc_last_error());
        return 1;
    }
    sc_selfplay_stats st;
    rc = sc_selfplay_run(sp, 192, 0, 0.0, &st);
    printf("%s threads=%d leaves_per_tree=%d: rc=%d games=%lld plies=%lld rollouts=%lld\n", arena ? "arena" : "selfplay", threads,
           leaves_per_tree, rc, (long long)st.games_finished, (long long)st.moves, (long long)st.rollouts);
    const bool ok = rc == SC_OK && st.games_finished == 192 && sc_selfplay_trace_json(sp, 191, nullptr, 0) > 0;
    sc_selfplay_destroy(sp);
    return ok ? 0 : 1;
}

int main(int argc, char **argv)
{
    int bad = 0;
    if (argc > 1) {  // `prof`: one single-threaded self-play run (the gprof target of `make hostprof`)
        bad = run(false, 1, 1);
        printf(bad ? "FAILED\n" : "host profile run ok\n");
        return bad;
    }
    bad += run(false, 8, 1);
    bad += run(false, 8, 4);
    bad += run(true, 8, 1);
    printf(bad ? "FAILED\n" : "tsan harness ok\n");
    return bad;
}
