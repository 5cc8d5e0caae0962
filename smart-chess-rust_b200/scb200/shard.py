"""Multi-GPU sharding of self-play by game: games are independent units, the reference
scales by running independent processes (`parallel -j$P`, scripts/run_batch:21).  Here: one process
per GPU, game g -> rank g mod world, no exchange step and no collective on the data path; only the
end-of-run counters are reduced (SUM) and the elapsed time (MAX)."""
from __future__ import annotations


def games_for_rank(n_games: int, rank: int, world: int) -> int:
    """How many of games 0..n_games-1 the rank owns under g -> g mod world."""
    return n_games // world + (1 if rank < n_games % world else 0)


def game_ids_for_rank(n_games: int, rank: int, world: int):
    return list(range(rank, n_games, world))


def rank_seed(base_seed: int, rank: int) -> int:
    """Per-rank RNG stream base; tree i of rank r seeds from (rank_seed, i) inside the driver."""
    return base_seed * 1000003 + rank


SUM_KEYS = ("leaf_evals", "terminal_evals", "rollouts", "moves", "games_finished", "white_wins", "black_wins",
            "draws", "unfinished", "batches")


def reduce_stats(stats: dict, dist=None, device="cpu") -> dict:
    """SUM the counters and MAX the times over ranks (torch.distributed, any backend)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(stats)
    import torch

    s = torch.tensor([float(stats[k]) for k in SUM_KEYS], dtype=torch.float64, device=device)
    t = torch.tensor([float(stats["seconds"]), float(stats["wait_seconds"])], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = {k: int(v) for k, v in zip(SUM_KEYS, s.tolist())}
    out["seconds"], out["wait_seconds"] = t.tolist()
    return out
