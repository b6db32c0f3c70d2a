"""Scene -> GPU partitioning for batched independent scenes (BASELINE.json config 4).

A planning cycle runs on ONE GPU; only independent scenes (own World, costmap, MapGrids) are sharded. There
is no collective on the data path: every rank plans its own scenes and the per-scene argmins (16 bytes each)
are gathered on the host. torch.distributed is used for the gather only.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple


def scenes_for_rank(n_scenes: int, rank: int, world_size: int) -> List[int]:
    """Scene i is planned by rank (i mod world_size)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    return list(range(rank, n_scenes, world_size))


def gather_scene_results(local: Dict[int, Tuple[int, float]], n_scenes: int, group=None) -> List[Tuple[int, float]]:
    """Host-side gather of {scene id: (best_index, best_total)} from every rank; returns the list ordered by
    scene id on every rank. Works with any torch.distributed backend (gloo on CPU, nccl on GPUs)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        merged = dict(local)
    else:
        parts: List[Dict[int, Tuple[int, float]]] = [None] * dist.get_world_size(group)  # type: ignore
        dist.all_gather_object(parts, local, group=group)
        merged = {}
        for p in parts:
            for k, v in p.items():
                if k in merged:
                    raise RuntimeError(f"scene {k} planned by two ranks")
                merged[k] = v
    missing = [s for s in range(n_scenes) if s not in merged]
    if missing:
        raise RuntimeError(f"scenes {missing[:8]} were planned by no rank")
    return [merged[s] for s in range(n_scenes)]


def plan_scenes_sharded(plan_fn, n_scenes: int, rank: int, world_size: int, group=None):
    """Runs `plan_fn(list of scene ids) -> {scene id: (best_index, best_total)}` on this rank's share and gathers."""
    mine = scenes_for_rank(n_scenes, rank, world_size)
    local = plan_fn(mine) if mine else {}
    return gather_scene_results(local, n_scenes, group=group)
