// Batched self-play driver: thousands of concurrent PUCT search trees whose leaves are evaluated
// as one device batch per step.
//
// Per tree the semantics are those of the reference, restated from:
//   `Node`, `uct`, `find_max`, `backward`, `select`, `mcts`   src/mcts.rs:16-24, 61-98, 132-289
//   `step` (move choice + reset of the chosen child)          src/mcts.rs:292-328
//   the self-play move loop and its termination rules         src/main.rs:153-238
//   `predict` terminal rule (no legal moves -> +1/-1/0)       src/backends/torch.rs:96-106
//   `reverse_q` (node's side to move is Black)                src/backends/torch.rs:49-52
//   trace format                                              src/trace.rs:5-42
// What is new (north_star): the trees advance in lock-step, each contributing one leaf per step
// (so every tree still runs its rollouts strictly one after another, exactly like `mcts::mcts`),
// priors are stored on the children at expansion instead of re-running `predict` at every level of
// every descent, positions are native bitboards (host/chess_rules.hpp) instead of python-chess
// objects, and trees are split in two groups that alternate between host work and device work.
// Batch rows are handed out densely (one atomic counter per group).  Optional, not reference-exact:
// leaves_per_tree > 1 lets a tree contribute several leaves per batch, in-flight paths carrying a
// virtual loss.  Arena mode (two networks, src/play.rs) keeps every game of a group at a ply of the
// same parity, so a batch belongs to one network; finished games are replaced at the next even ply.
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../common.cuh"
#include "chess_rules.hpp"
#include "host_util.hpp"

using namespace scb;
using namespace scb::chess;
using namespace scb::host;

namespace {

// ---- tree ------------------------------------------------------------------------------------------
struct Node {
    Move mv;
    uint8_t step_color;  // Step.1: the side to move AFTER the move (src/chess.rs:65-66)
    int32_t depth;
    float q;             // sum of backed-up rewards, White's perspective
    int32_t n;
    float uct;
    float prior;         // stored at expansion (the reference recomputes it on every descent)
    int32_t parent;
    int32_t first_child;
    int32_t n_children;
    uint8_t pending;     // an evaluation of this (unexpanded) node is in flight
};

struct Tree {
    std::vector<Node> nodes;
    int root = 0;
    Game game;         // state at the search root + scratch pushes along the current path
    int root_ply = 0;
    Rng rng;
    // rollouts whose leaf is waiting for the network: one per step in the reference-exact mode,
    // up to leaves_per_tree with virtual loss
    struct Pending {
        int leaf = -1;
        std::vector<int> path;
        MoveList moves;
        int child_color = 0;
        int slot = -1;     // row of the group's batch buffers
        bool vloss = false;
        float value = 0.f;  // hash evaluator: result computed at collection time
        float pri[256];
    };
    std::vector<Pending> pend;
    int n_pend = 0;
    // current search / game
    int rollouts_done = 0;
    int rollout_target = 0;  // rollouts of the current move (--rollout-num, or --rollout-factor x legal moves at the root)
    int64_t game_no = 0;     // start order of this slot's current game within the run (0-based)
    bool failed = false;     // the network returned a non-finite prior / value for one of this game's leaves
    int move_index = 0;      // `i` of the self-play loop (main.rs:168)
    int ply_offset = 0;      // arena: the (even) group ply at which this slot's current game started
    bool active = true;
    bool game_over = false;  // arena: this slot's game ended (the next one starts at the next even group ply)
    TraceRec trace;
    std::vector<float> noisy;  // scratch for root priors mixed with Dirichlet noise

    void new_game(uint64_t seed)
    {
        game = Game();
        nodes.clear();
        Node r{};
        r.step_color = WHITE;  // Step(None, White), main.rs:157-165
        r.parent = -1;
        r.first_child = -1;
        nodes.push_back(r);
        root = 0;
        root_ply = 0;
        rollouts_done = 0;
        move_index = 0;
        n_pend = 0;
        game_over = false;
        failed = false;
        rollout_target = -1;  // set before the first descent of every move (move_rollouts)
        trace = TraceRec();
        (void)seed;
    }
};

}  // namespace

struct sc_selfplay {
    sc_engine *eng = nullptr;
    sc_engine *eng_black = nullptr;  // arena mode: eng plays White, eng_black plays Black
    bool arena = false;
    int group_ply[2] = {0, 0};       // arena: the ply every tree of the group is searching
    sc_selfplay_config cfg{};
    std::vector<Tree> trees;
    // pinned batch buffers per pipeline group
    struct Group {
        int first = 0, count = 0;
        sc_position *pos = nullptr;
        sc_move *moves = nullptr;
        int32_t *cnt = nullptr;
        float *priors = nullptr, *value = nullptr;
        int ticket = -1;
        bool inflight = false;
        sc_engine *eng = nullptr;  // engine the in-flight batch was submitted to
        std::atomic<int> n_used{0};  // rows of the batch buffers filled by the current step (dense)
        int capacity = 0;            // trees of the group x leaves_per_tree
        int k_now = 1;               // leaves per tree of the current step (leaves_per_tree = -1: follows the
                                     // number of trees still playing, so that the batch stays full)
    } groups[2];
    int n_groups = 1;
    // stats
    std::atomic<int64_t> leaf_evals{0}, terminal_evals{0}, rollouts{0}, moves{0}, games_finished{0}, white{0}, black{0},
        draws{0}, unfinished{0}, games_started{0}, games_dropped{0};
    bool ran = false;  // sc_selfplay_run is single-shot
    int64_t batches = 0;
    int64_t max_games = 0, max_moves = 0;
    std::mutex trace_mu;
    std::vector<std::string> traces;
    std::vector<int64_t> trace_games;  // start order of the game each kept trace belongs to
    // worker pool
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_start, cv_done;
    int job_gen = 0, job_group = -1, job_pending = 0;
    bool quit = false;
    std::string err;
};

namespace {

// `uct` (src/mcts.rs:61-76), same f32 operation order
inline float uct_score(float sqrt_total, float prior, float q, int n_act, bool reverse_q, float cpuct)
{
    const float avg = q / ((float)n_act + 1e-4f) * (reverse_q ? -1.f : 1.f);
    const float expl = (sqrt_total + 0.01f) / (1.f + (float)n_act) * cpuct * prior;
    return avg + expl;
}

inline void pack_leaf(const Tree &t, int node_depth, sc_position *out)
{
    pack_position(t.game, node_depth + 1, out);  // history stops at the game root (src/chess.rs:851-867)
}

// virtual loss: while an evaluation is in flight every node of its path (below the root) looks like one more
// visit that ended in a loss for the player who chose it, so the next descent of the same step goes elsewhere
inline void apply_vloss(Tree &t, const std::vector<int> &path, float sign)
{
    for (size_t i = 1; i < path.size(); i++) {
        Node &c = t.nodes[path[i]];
        c.n += (int)sign;
        c.q += sign * (c.step_color == WHITE ? 1.f : -1.f);  // mover was Black iff White is to move after it
    }
}

// rollouts of the move the tree is about to search (src/main.rs:175-180)
inline int move_rollouts(const sc_selfplay_config &cfg, Tree &t)
{
    if (cfg.rollout_factor > 0.f) {
        MoveList l;
        t.game.cur.legal_moves(l);
        return std::min(300, (int)((float)l.n * cfg.rollout_factor));
    }
    return cfg.rollout_num > 0 ? cfg.rollout_num : 300;
}

// expansion (src/mcts.rs:269-283) + backward (src/mcts.rs:90-98) for one pending leaf
void finish_rollout(sc_selfplay *sp, Tree &t, Tree::Pending &P, const float *priors, float value)
{
    const int leaf = P.leaf;
    const int n = P.moves.n;
    if (P.vloss) apply_vloss(t, P.path, -1.f);
    t.nodes[leaf].pending = 0;
    if (n > 0) {
        const int first = (int)t.nodes.size();
        const int depth = t.nodes[leaf].depth + 1;
        t.nodes.resize(first + n);
        for (int i = 0; i < n; i++) {
            Node &c = t.nodes[first + i];
            c.mv = P.moves.m[i];
            c.step_color = (uint8_t)P.child_color;
            c.depth = depth;
            c.q = 0.f;
            c.n = 0;
            c.uct = 0.f;
            c.prior = priors[i];
            c.parent = leaf;
            c.first_child = -1;
            c.n_children = 0;
            c.pending = 0;
        }
        t.nodes[leaf].first_child = first;
        t.nodes[leaf].n_children = n;
    }
    for (int idx : P.path) {
        t.nodes[idx].n += 1;
        t.nodes[idx].q += value;
    }
    t.rollouts_done++;
    sp->rollouts.fetch_add(1, std::memory_order_relaxed);
}

// one `select` descent (src/mcts.rs:132-227). Returns true if the leaf needs the network.
// Returns 1 if the leaf needs the network, 0 if the rollout finished at a terminal leaf, -1 if it ran into a
// leaf whose evaluation is already in flight (virtual-loss mode only; nothing was changed).
int descend(sc_selfplay *sp, Tree &t, Tree::Pending &P)
{
    const sc_selfplay_config &cfg = sp->cfg;
    P.path.clear();
    int node = t.root;
    P.path.push_back(node);
    for (;;) {
        Node &nd = t.nodes[node];
        if (nd.n_children == 0) break;  // leaf: unexplored or terminal
        int best;
        if (nd.n_children == 1)
            best = nd.first_child;
        else {
            const bool reverse_q = nd.step_color == BLACK;
            const Node *ch = &t.nodes[nd.first_child];
            int tot = 0;
            for (int i = 0; i < nd.n_children; i++) tot += ch[i].n;
            const float sq = std::sqrt((float)tot);
            const float *pri = nullptr;
            const bool is_root = P.path.size() == 1;
            if (is_root && cfg.with_noise && nd.n_children >= 2) {
                // fresh Dirichlet(0.3) sample on every rollout (src/mcts.rs:123-130, 171-184)
                t.noisy.resize(nd.n_children);
                float g[256];
                float tot_g = 0.f;
                for (int i = 0; i < nd.n_children; i++) {
                    g[i] = t.rng.gammaf(0.3f);
                    tot_g += g[i];
                }
                const float inv = tot_g > 0.f ? 1.0f / tot_g : 0.f;
                for (int i = 0; i < nd.n_children; i++)
                    t.noisy[i] = ch[i].prior * (1.0f - cfg.epsilon) + (g[i] * inv) * cfg.epsilon;
                pri = t.noisy.data();
            }
            int bi = 0;
            float bu = 0.f;
            for (int i = 0; i < nd.n_children; i++) {
                Node &c = t.nodes[nd.first_child + i];
                const float u = uct_score(sq, pri ? pri[i] : c.prior, c.q, c.n, reverse_q, cfg.cpuct);
                c.uct = u;
                if (i == 0 || u >= bu) {  // Iterator::max_by keeps the LAST maximum
                    bu = u;
                    bi = i;
                }
            }
            best = nd.first_child + bi;
        }
        t.game.push(t.nodes[best].mv);
        P.path.push_back(best);
        node = best;
    }
    // `predict` at the leaf: legal moves; none -> terminal value without the network
    auto unwind = [&]() {
        while (t.game.ply() > t.root_ply) t.game.pop();
    };
    if (t.nodes[node].pending) {
        unwind();
        return -1;
    }
    P.leaf = node;
    P.vloss = false;
    P.child_color = !t.game.cur.turn;
    t.game.cur.legal_moves(P.moves);
    if (P.moves.n == 0) {
        float v = 0.f;
        if (t.game.cur.in_check()) v = t.game.cur.turn == WHITE ? -1.f : 1.f;  // winner = side that mated
        sp->terminal_evals.fetch_add(1, std::memory_order_relaxed);
        unwind();
        finish_rollout(sp, t, P, nullptr, v);
        return 0;
    }
    return 1;  // the game is left AT the leaf: the caller packs the position, then unwinds
}

// the move choice of `mcts::step` (src/mcts.rs:292-328) + the loop body of main.rs:198-228.
// Returns false when the game is over.
bool play_move(sc_selfplay *sp, Tree &t)
{
    const sc_selfplay_config &cfg = sp->cfg;
    Node &root = t.nodes[t.root];
    bool over = false;
    if (root.n_children == 0) {
        // step() == None: no legal move at the root -> outcome, break (main.rs:212-216)
        int w;
        int term = t.game.outcome(true, &w);
        t.trace.has_outcome = term != T_NONE;
        t.trace.termination = term;
        t.trace.winner = w;
        over = true;
    } else {
        const float temperature = t.move_index < cfg.temperature_switch ? 1.0f : cfg.temperature;
        const Node *ch = &t.nodes[root.first_child];
        int choice = 0;
        if (temperature == 0.0f) {
            for (int i = 1; i < root.n_children; i++)
                if (ch[i].n > ch[choice].n) choice = i;  // position() of the FIRST maximum
        } else {
            // WeightedIndex over N^(1/T)
            const float power = 1.0f / temperature;
            double tot = 0.0;
            std::vector<double> w(root.n_children);
            for (int i = 0; i < root.n_children; i++) {
                w[i] = std::pow((float)ch[i].n, power);
                tot += w[i];
            }
            double r = t.rng.uniform() * tot, acc = 0.0;
            choice = root.n_children - 1;
            for (int i = 0; i < root.n_children; i++) {
                acc += w[i];
                if (r < acc) {
                    choice = i;
                    break;
                }
            }
        }
        TraceStep st;
        st.mv = ch[choice].mv;
        st.q = root.q;
        if (cfg.keep_traces) {
            st.cmv.resize(root.n_children);
            st.cn.resize(root.n_children);
            st.cq.resize(root.n_children);
            st.cu.resize(root.n_children);
            for (int i = 0; i < root.n_children; i++) {
                st.cmv[i] = ch[i].mv;
                st.cn[i] = ch[i].n;
                st.cq[i] = ch[i].q;
                st.cu[i] = ch[i].uct;
            }
            t.trace.steps.push_back(std::move(st));
        }
        // navigate_down + reset(): the chosen child becomes a fresh root (q = 0, n = 0, no children)
        Node nr = ch[choice];
        nr.q = 0.f;
        nr.n = 0;
        nr.first_child = -1;
        nr.n_children = 0;
        nr.parent = -1;
        t.game.push(nr.mv);
        t.root_ply = t.game.ply();
        t.nodes.clear();
        t.nodes.push_back(nr);
        t.root = 0;
        sp->moves.fetch_add(1, std::memory_order_relaxed);
        // main.rs:223-228: the outcome is only looked at after move index 100
        if (t.move_index > 100) {
            int w;
            int term = t.game.outcome(true, &w);
            if (term != T_NONE) {
                t.trace.has_outcome = true;
                t.trace.termination = term;
                t.trace.winner = w;
                over = true;
            }
        }
        t.move_index++;
        if (!over && t.move_index >= cfg.num_steps) over = true;  // loop bound: trace saved without outcome
    }
    t.rollouts_done = 0;
    t.rollout_target = -1;
    if (over) {
        if (t.trace.has_outcome) {
            if (t.trace.winner == WHITE) sp->white++;
            else if (t.trace.winner == BLACK) sp->black++;
            else sp->draws++;
        } else
            sp->unfinished++;
        if (cfg.keep_traces) {
            std::string js = trace_to_json(t.trace);
            std::lock_guard<std::mutex> lk(sp->trace_mu);
            sp->traces.push_back(std::move(js));
            sp->trace_games.push_back(t.game_no);
        }
        sp->games_finished++;
    }
    return !over;
}

// arena: `NNPlayer::bestmove` choice + `step` + the loop test of `play_loop` (src/play.rs:253-343)
void play_move_arena(sc_selfplay *sp, Tree &t)
{
    const sc_selfplay_config &cfg = sp->cfg;
    Node &root = t.nodes[t.root];
    const Node *ch = &t.nodes[root.first_child];
    const float temperature = root.depth < cfg.temperature_switch ? 1.0f : cfg.temperature;
    int choice = 0;
    if (temperature == 0.0f) {
        int mx = 0, cnt = 0;
        for (int i = 0; i < root.n_children; i++) mx = std::max(mx, ch[i].n);
        for (int i = 0; i < root.n_children; i++) cnt += ch[i].n == mx;
        int pick = (int)(t.rng.uniform() * cnt);  // uniform among the most visited (play.rs:269-278)
        if (pick >= cnt) pick = cnt - 1;
        for (int i = 0; i < root.n_children; i++)
            if (ch[i].n == mx && pick-- == 0) {
                choice = i;
                break;
            }
    } else {
        const float power = 1.0f / temperature;
        double tot = 0.0;
        std::vector<double> w(root.n_children);
        for (int i = 0; i < root.n_children; i++) {
            w[i] = std::pow((float)ch[i].n, power);
            tot += w[i];
        }
        double r = t.rng.uniform() * tot, acc = 0.0;
        choice = root.n_children - 1;
        for (int i = 0; i < root.n_children; i++) {
            acc += w[i];
            if (r < acc) {
                choice = i;
                break;
            }
        }
    }
    if (cfg.keep_traces) {
        TraceStep st;
        st.mv = ch[choice].mv;
        st.q = root.q;
        for (int i = 0; i < root.n_children; i++) {
            st.cmv.push_back(ch[i].mv);
            st.cn.push_back(ch[i].n);
            st.cq.push_back(ch[i].q);
            st.cu.push_back(ch[i].uct);
        }
        t.trace.steps.push_back(std::move(st));
    }
    Node nr = ch[choice];
    nr.q = 0.f;
    nr.n = 0;
    nr.first_child = -1;
    nr.n_children = 0;
    nr.parent = -1;
    t.game.push(nr.mv);
    t.root_ply = t.game.ply();
    t.nodes.clear();
    t.nodes.push_back(nr);
    t.root = 0;
    t.rollouts_done = 0;
    t.rollout_target = -1;
    t.move_index++;
    sp->moves.fetch_add(1, std::memory_order_relaxed);
    int w;
    const int term = t.game.outcome(true, &w);
    if (term != T_NONE || t.move_index >= cfg.num_steps) {
        t.trace.has_outcome = term != T_NONE;
        t.trace.termination = term;
        t.trace.winner = w;
        t.game_over = true;
        if (term == T_NONE) sp->unfinished++;
        else if (w == WHITE) sp->white++;
        else if (w == BLACK) sp->black++;
        else sp->draws++;
        if (cfg.keep_traces) {
            std::string js = trace_to_json(t.trace);
            std::lock_guard<std::mutex> lk(sp->trace_mu);
            sp->traces.push_back(std::move(js));
            sp->trace_games.push_back(t.game_no);
        }
        sp->games_finished++;
    }
}

// Collects up to leaves_per_tree leaves of one tree into the group's batch buffers.  Returns the number
// collected.  `may_search` is evaluated before every descent.
struct BatchOut {
    sc_position *pos;
    sc_move *moves;
    int32_t *cnt;
    std::atomic<int> *n_used;
    int k;  // leaves per tree of this step
};

template <typename Ready>
int collect_leaves(sc_selfplay *sp, Tree &t, const BatchOut &out, Ready finish_move)
{
    const int K = out.k;
    if ((int)t.pend.size() < K) t.pend.resize(K);
    int collected = 0;
    for (;;) {
        if (t.rollout_target < 0) t.rollout_target = move_rollouts(sp->cfg, t);
        if (t.rollouts_done + collected >= t.rollout_target) {
            if (collected > 0) break;           // the move's last evaluations are in flight
            if (!finish_move()) break;          // move played / game over / tree waits
            continue;
        }
        if (collected >= K) break;
        Tree::Pending &P = t.pend[collected];
        const int r = descend(sp, t, P);
        if (r == 0) continue;                    // terminal leaf: rollout already backed up
        if (r < 0) break;                        // ran into an in-flight leaf: wait for it
        if (sp->cfg.evaluator == 1) {
            P.value = hash_eval(t.game.cur, P.moves.m, P.moves.n, P.pri);
        } else {
            const int s = out.n_used->fetch_add(1, std::memory_order_relaxed);
            P.slot = s;
            pack_leaf(t, t.nodes[P.leaf].depth, out.pos + s);
            sc_move *mv = out.moves + (size_t)s * SC_MAX_MOVES;
            for (int i = 0; i < P.moves.n; i++) mv[i] = sc_move{P.moves.m[i].from, P.moves.m[i].to, P.moves.m[i].promo, 0};
            out.cnt[s] = P.moves.n;
        }
        while (t.game.ply() > t.root_ply) t.game.pop();
        sp->leaf_evals.fetch_add(1, std::memory_order_relaxed);
        t.nodes[P.leaf].pending = 1;
        if (K > 1) {
            apply_vloss(t, P.path, 1.f);
            P.vloss = true;
        }
        collected++;
        if (sp->cfg.evaluator == 1 && K == 1) {
            // reference-exact mode on the stand-in evaluator: finish at once, keep going
            finish_rollout(sp, t, P, P.pri, P.value);
            collected = 0;
        }
    }
    if (sp->cfg.evaluator == 1) {
        for (int i = 0; i < collected; i++) finish_rollout(sp, t, t.pend[i], t.pend[i].pri, t.pend[i].value);
        if (collected > 0) return -1;            // stand-in evaluator: call again, nothing is in flight
        return 0;
    }
    t.n_pend = collected;
    return collected;
}

// Per-game failure isolation: a leaf whose priors / value came back non-finite poisons only its own game.  The
// reference prints a warning and plays on with NaNs in the tree (src/backends/torch.rs:129-135); here the game is
// dropped and counted, the slot starts the run's next game, every other game of the batch is untouched.
void finish_pending(sc_selfplay *sp, Tree &t, const float *priors, const float *value)
{
    for (int i = 0; i < t.n_pend; i++) {
        Tree::Pending &P = t.pend[i];
        const float *pr = priors + (size_t)P.slot * SC_MAX_MOVES;
        bool ok = std::isfinite(value[P.slot]);
        for (int k = 0; ok && k < P.moves.n; k++) ok = std::isfinite(pr[k]);
        if (!ok) t.failed = true;
        if (!t.failed) finish_rollout(sp, t, P, pr, value[P.slot]);
    }
    t.n_pend = 0;
    if (t.failed) {
        sp->games_dropped++;
        const int64_t started = sp->games_started.fetch_add(1) + 1;
        if (sp->max_games > 0 && started > sp->max_games) {
            t.new_game(0);
            t.active = false;
        } else {
            const int off = t.ply_offset, mi = t.move_index;
            t.new_game(0);
            t.game_no = started - 1;
            // arena: the replacement game starts at the next even group ply after the one the slot was searching
            if (sp->arena) t.ply_offset = (off + mi + 2) & ~1;
        }
    }
}

// arena: a tree only searches the ply its pipeline group is at, then waits for the others.
// A finished game's slot starts the run's next game at the next EVEN group ply: White (the first network) is
// to move at even group plies in every game of the group, so one batch never mixes the two networks.
void advance_tree_arena(sc_selfplay *sp, Tree &t, int group_ply, const BatchOut &out, const float *priors,
                        const float *value)
{
    finish_pending(sp, t, priors, value);
    auto next_game = [&]() {
        const int64_t k = sp->games_started.fetch_add(1) + 1;
        if (sp->max_games > 0 && k > sp->max_games) {
            t.active = false;
            return;
        }
        t.new_game(0);
        t.game_no = k - 1;
        t.ply_offset = (group_ply + 2) & ~1;  // > group_ply and even
    };
    for (;;) {
        if (t.active && t.game_over) next_game();
        if (!t.active || t.move_index + t.ply_offset > group_ply) return;
        const int r = collect_leaves(sp, t, out, [&]() {
            play_move_arena(sp, t);
            if (t.game_over) next_game();
            return t.active && t.move_index + t.ply_offset <= group_ply;
        });
        if (r >= 0) return;
    }
}

// advance one tree until it has leaves for the network or has no more work
void advance_tree(sc_selfplay *sp, Tree &t, const BatchOut &out, const float *priors, const float *value)
{
    finish_pending(sp, t, priors, value);
    for (;;) {
        if (!t.active) return;
        const int r = collect_leaves(sp, t, out, [&]() {
            if (!play_move(sp, t)) {
                // start the next game in this slot if the run still needs games
                int64_t started = sp->games_started.fetch_add(1) + 1;
                if (sp->max_games > 0 && started > sp->max_games) {
                    t.active = false;
                    return false;
                }
                t.new_game(0);
                t.game_no = started - 1;
            }
            if (sp->max_moves > 0 && sp->moves.load(std::memory_order_relaxed) >= sp->max_moves) {
                t.active = false;
                return false;
            }
            return true;
        });
        if (r >= 0) return;
    }
}

void run_group_slice(sc_selfplay *sp, int g, int worker, int n_workers)
{
    sc_selfplay::Group &G = sp->groups[g];
    const int per = (G.count + n_workers - 1) / n_workers;
    const int lo = worker * per, hi = std::min(G.count, lo + per);
    const BatchOut out{G.pos, G.moves, G.cnt, &G.n_used, G.k_now};
    for (int i = lo; i < hi; i++) {
        Tree &t = sp->trees[G.first + i];
        if (sp->arena)
            advance_tree_arena(sp, t, sp->group_ply[g], out, G.priors, G.value);
        else
            advance_tree(sp, t, out, G.priors, G.value);
    }
}

void worker_main(sc_selfplay *sp, int worker, int n_workers)
{
    int seen = 0;
    for (;;) {
        int g;
        {
            std::unique_lock<std::mutex> lk(sp->mu);
            sp->cv_start.wait(lk, [&] { return sp->quit || sp->job_gen != seen; });
            if (sp->quit) return;
            seen = sp->job_gen;
            g = sp->job_group;
        }
        run_group_slice(sp, g, worker + 1, n_workers + 1);
        {
            std::lock_guard<std::mutex> lk(sp->mu);
            if (--sp->job_pending == 0) sp->cv_done.notify_one();
        }
    }
}

// n_threads = T: the calling thread takes slice 0, T - 1 pool threads take the others
void parallel_advance(sc_selfplay *sp, int g)
{
    {
        // leaves per tree of this step
        sc_selfplay::Group &G = sp->groups[g];
        const int k = sp->cfg.leaves_per_tree;
        if (k >= 0)
            G.k_now = k > 1 ? k : 1;
        else {
            // auto: one leaf per tree (the reference's search) while most trees are playing; when games run out
            // the remaining trees share the batch rows, up to 16 leaves each, with virtual loss
            int active = 0;
            for (int i = 0; i < G.count; i++) active += sp->trees[G.first + i].active ? 1 : 0;
            int kk = active > 0 ? G.count / active : 1;
            G.k_now = kk < 1 ? 1 : (kk > 16 ? 16 : kk);
        }
    }
    const int nw = (int)sp->workers.size();
    if (nw == 0) {
        run_group_slice(sp, g, 0, 1);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(sp->mu);
        sp->job_group = g;
        sp->job_pending = nw;
        sp->job_gen++;
    }
    sp->cv_start.notify_all();
    run_group_slice(sp, g, 0, nw + 1);
    std::unique_lock<std::mutex> lk(sp->mu);
    sp->cv_done.wait(lk, [&] { return sp->job_pending == 0; });
}

}  // namespace

extern "C" {

int sc_selfplay_create(sc_engine *e, const sc_selfplay_config *cfg, sc_selfplay **out)
{
    if (cfg && cfg->rollout_factor > 0.f && cfg->rollout_num > 0) {
        set_error("sc_selfplay_create: both rollout_factor and rollout_num are specified");  // main.rs:179 panics
        return SC_E_INVAL;
    }
    if (!cfg || !out || cfg->n_trees <= 0 || cfg->rollout_num < 0 || cfg->rollout_factor < 0.f || cfg->num_steps <= 0 ||
        (cfg->evaluator == 0 && !e) || (cfg->evaluator != 0 && cfg->evaluator != 1)) {
        set_error("sc_selfplay_create: bad argument");
        return SC_E_INVAL;
    }
    if (cfg->leaves_per_tree < -1 || cfg->leaves_per_tree > 64) {
        set_error("sc_selfplay_create: leaves_per_tree must be -1 (auto) or 0..64");
        return SC_E_INVAL;
    }
    const int kleaves = cfg->leaves_per_tree > 1 ? cfg->leaves_per_tree : 1;
    sc_selfplay *sp = new sc_selfplay();
    sp->eng = e;
    sp->cfg = *cfg;
    sp->n_groups = cfg->pipeline_groups >= 2 && cfg->n_trees >= 2 && cfg->evaluator == 0 ? 2 : 1;
    // Group sizes: the conv kernels run one 2-board tile per SM per wave, so a batch costs
    // ceil(boards / (2 * SMs)) waves.  Splitting 2048 trees as 1024 + 1024 costs 4 + 4 waves; sizing the
    // first group to a whole number of waves (888 = 3 * 296 boards on 148 SMs) costs 3 + 4 = the 7 waves
    // of a single 2048 batch.
    int size0 = sp->n_groups == 1 ? cfg->n_trees : (cfg->n_trees + 1) / 2;
    if (sp->n_groups == 2 && e) {
        int sms = 148, dev = 0;
        sc_device_info(e, &dev, &sms);  // the engine's device, not device 0
        const int wave = 2 * sms;
        if (size0 * kleaves >= wave) size0 = std::max(1, size0 * kleaves / wave * wave / kleaves);
    }
    if (e) {
        int mb = 0;
        sc_info(e, nullptr, &mb, nullptr);
        if (std::max(size0, cfg->n_trees - size0) * kleaves > mb) {
            delete sp;
            set_error("sc_selfplay_create: trees per pipeline group x leaves_per_tree exceed the engine's max_batch");
            return SC_E_INVAL;
        }
    }
    sp->trees.resize(cfg->n_trees);
    for (int i = 0; i < cfg->n_trees; i++) {
        sp->trees[i].rng.seed(cfg->seed * 0x9E3779B97F4A7C15ULL + (uint64_t)i + 1);
        sp->trees[i].new_game(0);
        sp->trees[i].game_no = i;
    }
    for (int g = 0; g < sp->n_groups; g++) {
        sc_selfplay::Group &G = sp->groups[g];
        G.first = g == 0 ? 0 : size0;
        G.count = sp->n_groups == 1 ? cfg->n_trees : (g == 0 ? size0 : cfg->n_trees - size0);
        G.capacity = G.count * kleaves;
        const size_t nm = (size_t)G.capacity * SC_MAX_MOVES;
        if (cfg->evaluator == 0) {
            if (cudaMallocHost(&G.pos, sizeof(sc_position) * G.capacity) != cudaSuccess ||
                cudaMallocHost(&G.moves, sizeof(sc_move) * nm) != cudaSuccess ||
                cudaMallocHost(&G.cnt, sizeof(int32_t) * G.capacity) != cudaSuccess ||
                cudaMallocHost(&G.priors, sizeof(float) * nm) != cudaSuccess ||
                cudaMallocHost(&G.value, sizeof(float) * G.capacity) != cudaSuccess) {
                set_error("sc_selfplay_create: pinned allocation failed");
                sc_selfplay_destroy(sp);
                return SC_E_CUDA;
            }
        } else {
            G.pos = new sc_position[G.capacity];
            G.moves = new sc_move[nm];
            G.cnt = new int32_t[G.capacity];
            G.priors = new float[nm];
            G.value = new float[G.capacity];
        }
        memset(G.cnt, 0, sizeof(int32_t) * G.capacity);
    }
    const int nt = cfg->n_threads > 1 ? cfg->n_threads - 1 : 0;  // pool threads next to the calling thread
    for (int w = 0; w < nt; w++) sp->workers.emplace_back(worker_main, sp, w, nt);
    *out = sp;
    return SC_OK;
}

int sc_arena_create(sc_engine *white, sc_engine *black, const sc_selfplay_config *cfg, sc_selfplay **out)
{
    if (!cfg || (cfg->evaluator == 0 && (!white || !black))) {
        set_error("sc_arena_create: bad argument");
        return SC_E_INVAL;
    }
    sc_selfplay_config c = *cfg;
    c.with_noise = 0;  // play.rs:250 turns the Dirichlet noise off
    int rc = sc_selfplay_create(white, &c, out);
    if (rc != SC_OK) return rc;
    if (black) {
        int mb = 0;
        sc_info(black, nullptr, &mb, nullptr);
        int need = 0;
        for (int g = 0; g < (*out)->n_groups; g++) need = std::max(need, (*out)->groups[g].capacity);
        if (need > mb) {
            sc_selfplay_destroy(*out);
            *out = nullptr;
            set_error("sc_arena_create: trees per pipeline group exceed the black engine's max_batch");
            return SC_E_INVAL;
        }
    }
    (*out)->arena = true;
    (*out)->eng_black = black;
    return SC_OK;
}

int sc_selfplay_run(sc_selfplay *sp, int64_t max_games, int64_t max_moves, double max_seconds, sc_selfplay_stats *stats)
{
    if (!sp) return SC_E_INVAL;
    if (sp->ran) {
        set_error("sc_selfplay_run: a driver object plays one run; create a new one");
        return SC_E_STATE;
    }
    sp->ran = true;
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    sp->max_games = max_games;
    sp->max_moves = max_moves;
    // the games already sitting in the tree slots count as started
    if (sp->games_started.load() == 0) sp->games_started = (int64_t)sp->trees.size();
    if (max_games > 0)
        for (size_t i = 0; i < sp->trees.size(); i++)
            if ((int64_t)i >= max_games && sp->trees[i].move_index == 0 && sp->trees[i].rollouts_done == 0)
                sp->trees[i].active = false;
    double wait_s = 0.0;
    int rc = SC_OK;
    bool any = true;
    while (any && rc == SC_OK) {
        any = false;
        for (int g = 0; g < sp->n_groups && rc == SC_OK; g++) {
            sc_selfplay::Group &G = sp->groups[g];
            if (G.inflight) {
                const auto w0 = clk::now();
                rc = sc_eval_wait(G.eng ? G.eng : sp->eng, G.ticket);
                wait_s += std::chrono::duration<double>(clk::now() - w0).count();
                G.inflight = false;
                if (rc != SC_OK) break;
            }
            G.n_used.store(0, std::memory_order_relaxed);
            parallel_advance(sp, g);
            int n_leaves = G.n_used.load(std::memory_order_relaxed);
            if (sp->arena) {
                // nobody produced a leaf: every game of the group made its move of this ply -> next ply
                for (int guard = 0; n_leaves == 0 && guard < 1000000; guard++) {
                    bool playing = false;
                    for (int i = 0; i < G.count; i++) playing = playing || sp->trees[G.first + i].active;
                    if (!playing) break;
                    sp->group_ply[g]++;
                    parallel_advance(sp, g);
                    n_leaves = G.n_used.load(std::memory_order_relaxed);
                    if (sp->cfg.evaluator == 1) {
                        // hash evaluator never leaves a leaf pending: keep stepping plies until every game is played
                        bool left = false;
                        for (int i = 0; i < G.count; i++) left = left || sp->trees[G.first + i].active;
                        if (!left) break;
                    }
                }
            }
            if (sp->cfg.evaluator == 0 && n_leaves > 0) {
                sc_engine *use = sp->arena && (sp->group_ply[g] & 1) ? sp->eng_black : sp->eng;
                rc = sc_eval_submit(use, n_leaves, G.pos, G.moves, G.cnt, G.priors, G.value, nullptr, &G.ticket);
                G.inflight = rc == SC_OK;
                G.eng = use;
                sp->batches++;
                any = true;
            }
        }
        if (max_seconds > 0 && std::chrono::duration<double>(clk::now() - t0).count() > max_seconds) {
            // time is up: take the results of the batches in flight into the trees (no virtual visit, no pending
            // flag is left behind), then stop
            for (int g = 0; g < sp->n_groups; g++) {
                sc_selfplay::Group &G = sp->groups[g];
                if (!G.inflight) continue;
                rc = sc_eval_wait(G.eng ? G.eng : sp->eng, G.ticket);
                G.inflight = false;
                if (rc != SC_OK) break;
                for (int i = 0; i < G.count; i++) finish_pending(sp, sp->trees[G.first + i], G.priors, G.value);
            }
            for (auto &t : sp->trees) t.active = false;
            break;
        }
    }
    if (stats) {
        stats->leaf_evals = sp->leaf_evals;
        stats->terminal_evals = sp->terminal_evals;
        stats->rollouts = sp->rollouts;
        stats->moves = sp->moves;
        stats->games_finished = sp->games_finished;
        stats->white_wins = sp->white;
        stats->black_wins = sp->black;
        stats->draws = sp->draws;
        stats->unfinished = sp->unfinished;
        stats->batches = sp->batches;
        stats->seconds = std::chrono::duration<double>(clk::now() - t0).count();
        stats->wait_seconds = wait_s;
        stats->games_dropped = sp->games_dropped;
    }
    return rc;
}

int sc_selfplay_run_many(sc_selfplay **sps, int n, int64_t max_games, int64_t max_moves, double max_seconds,
                         sc_selfplay_stats *stats)
{
    if (!sps || n <= 0 || !stats) {
        set_error("sc_selfplay_run_many: bad argument");
        return SC_E_INVAL;
    }
    std::vector<int> rcs((size_t)n, SC_OK);
    std::vector<std::string> errs((size_t)n);
    std::vector<std::thread> th;
    for (int i = 0; i < n; i++)
        th.emplace_back([&, i] {
            rcs[i] = sc_selfplay_run(sps[i], max_games, max_moves, max_seconds, stats + i);
            if (rcs[i] != SC_OK) errs[i] = sc_last_error();  // the message is thread-local
        });
    for (auto &t : th) t.join();
    for (int i = 0; i < n; i++)
        if (rcs[i] != SC_OK) {
            set_error("driver " + std::to_string(i) + ": " + errs[i]);
            return rcs[i];
        }
    return SC_OK;
}

int64_t sc_selfplay_trace_json(sc_selfplay *sp, int64_t k, char *buf, int64_t cap)
{
    if (!sp) return -1;
    std::lock_guard<std::mutex> lk(sp->trace_mu);
    if (k < 0 || k >= (int64_t)sp->traces.size()) return -1;
    const std::string &s = sp->traces[(size_t)k];
    if (buf && cap > 0) {
        const int64_t n = std::min<int64_t>(cap - 1, (int64_t)s.size());
        memcpy(buf, s.data(), (size_t)n);
        buf[n] = 0;
    }
    return (int64_t)s.size() + 1;
}

int64_t sc_selfplay_trace_game(sc_selfplay *sp, int64_t k)
{
    if (!sp) return -1;
    std::lock_guard<std::mutex> lk(sp->trace_mu);
    if (k < 0 || k >= (int64_t)sp->trace_games.size()) return -1;
    return sp->trace_games[(size_t)k];
}

int sc_random_positions(int n, uint64_t seed, int max_ply, sc_position *pos_out, sc_move *moves_out, int32_t *move_off,
                        int max_moves_total)
{
    if (n < 0 || (n > 0 && (!pos_out || !moves_out || !move_off)) || max_ply <= 0) {
        set_error("sc_random_positions: bad argument");
        return SC_E_INVAL;
    }
    Rng rng;
    rng.seed(seed);
    Tree t;
    t.new_game(0);
    int k = 0, total = 0;
    if (move_off) move_off[0] = 0;
    while (k < n) {
        MoveList l;
        t.game.cur.legal_moves(l);
        if (l.n == 0 || t.game.ply() >= max_ply) {
            t.new_game(0);
            continue;
        }
        if (total + l.n > max_moves_total) {
            set_error("sc_random_positions: move buffer too small");
            return SC_E_INVAL;
        }
        pack_leaf(t, t.game.ply(), pos_out + k);
        for (int i = 0; i < l.n; i++) moves_out[total + i] = sc_move{l.m[i].from, l.m[i].to, l.m[i].promo, 0};
        total += l.n;
        move_off[++k] = total;
        t.game.push(l.m[(int)(rng.next() % (uint64_t)l.n)]);
    }
    return SC_OK;
}

int sc_test_dirichlet(uint64_t seed, float alpha, int n, float *out)
{
    if (n <= 0 || !out || !(alpha > 0.f)) return SC_E_INVAL;
    Rng rng;
    rng.seed(seed);
    float tot = 0.f;
    for (int i = 0; i < n; i++) {
        out[i] = rng.gammaf(alpha);
        tot += out[i];
    }
    for (int i = 0; i < n; i++) out[i] /= tot;
    return SC_OK;
}

static uint64_t perft_nodes(const Position &p, int depth)
{
    MoveList l;
    p.legal_moves(l);
    if (depth <= 1) return depth == 1 ? (uint64_t)l.n : 1;
    uint64_t n = 0;
    for (int i = 0; i < l.n; i++) {
        Position c = p;
        c.push(l.m[i]);
        n += perft_nodes(c, depth - 1);
    }
    return n;
}

int sc_rules_perft(const char *fen, int depth, uint64_t *nodes)
{
    Position p;
    if (!nodes || depth < 0 || depth > 7 || (fen ? !p.set_fen(fen) : (p.set_start(), false))) {
        set_error("sc_rules_perft: bad argument");
        return SC_E_INVAL;
    }
    *nodes = perft_nodes(p, depth);
    return SC_OK;
}

int sc_rules_probe(const sc_move *history, int n_history, sc_move *legal_out, int *n_legal, sc_position *packed_out,
                   int *termination, int *winner)
{
    return sc_rules_probe_fen(nullptr, history, n_history, legal_out, n_legal, packed_out, termination, winner);
}

int sc_rules_probe_fen(const char *fen, const sc_move *history, int n_history, sc_move *legal_out, int *n_legal,
                       sc_position *packed_out, int *termination, int *winner)
{
    if (n_history < 0 || (n_history > 0 && !history)) {
        set_error("sc_rules_probe: bad argument");
        return SC_E_INVAL;
    }
    Tree t;
    t.new_game(0);
    if (fen && !t.game.cur.set_fen(fen)) {
        set_error("sc_rules_probe_fen: bad FEN");
        return SC_E_INVAL;
    }
    MoveList l;
    for (int i = 0; i < n_history; i++) {
        t.game.cur.legal_moves(l);
        Move m{history[i].from, history[i].to, history[i].promo};
        bool ok = false;
        for (int k = 0; k < l.n; k++) ok = ok || (l.m[k] == m);
        if (!ok) {
            set_error("sc_rules_probe: illegal move in history at ply " + std::to_string(i));
            return SC_E_INVAL;
        }
        t.game.push(m);
    }
    t.game.cur.legal_moves(l);
    if (n_legal) *n_legal = l.n;
    if (legal_out)
        for (int k = 0; k < l.n; k++) legal_out[k] = sc_move{l.m[k].from, l.m[k].to, l.m[k].promo, 0};
    if (packed_out) pack_leaf(t, t.game.ply(), packed_out);
    if (termination) {
        int w = -1;
        *termination = t.game.outcome(true, &w);
        if (winner) *winner = w;
    }
    return SC_OK;
}

int sc_selfplay_destroy(sc_selfplay *sp)
{
    if (!sp) return SC_OK;
    {
        std::lock_guard<std::mutex> lk(sp->mu);
        sp->quit = true;
    }
    sp->cv_start.notify_all();
    for (auto &w : sp->workers) w.join();
    for (int g = 0; g < sp->n_groups; g++) {
        sc_selfplay::Group &G = sp->groups[g];
        if (sp->cfg.evaluator == 0) {
            if (G.inflight && G.eng) sc_eval_wait(G.eng, G.ticket);
            cudaFreeHost(G.pos);
            cudaFreeHost(G.moves);
            cudaFreeHost(G.cnt);
            cudaFreeHost(G.priors);
            cudaFreeHost(G.value);
        } else {
            delete[] G.pos;
            delete[] G.moves;
            delete[] G.cnt;
            delete[] G.priors;
            delete[] G.value;
        }
    }
    delete sp;
    return SC_OK;
}

}  // extern "C"
