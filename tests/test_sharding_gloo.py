"""world_size-2 gloo test of the multi-GPU path's host logic: scenes are partitioned over ranks, each rank
plans its share (the CPU oracle stands in for the GPU planner here -- tests may use it), the per-scene
argmins are gathered on the host and equal the single-process result. No collective touches the data path."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plan_fn_factory():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from humap_local_planner_b200 import scenes
    cfg = scenes.CONFIGS["cfg0"]
    params = scenes.make_params(cfg)
    smp = scenes.make_sampling(cfg)

    def plan_fn(ids):
        out = {}
        for s in ids:
            sc = scenes.make_scene(cfg, 200 + s)
            r = ob.plan(params, sc, smp, early_exit=True, want=())["result"]
            out[s] = (int(r.best_index), float(r.best_total))
        return out
    return plan_fn


def _worker(rank, world, port, n_scenes, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from humap_local_planner_b200.sharding import plan_scenes_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = plan_scenes_sharded(_plan_fn_factory(), n_scenes, rank, world)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_partition_covers_every_scene_once():
    sys.path.insert(0, ROOT)
    from humap_local_planner_b200.sharding import scenes_for_rank
    for world in (1, 2, 4, 8):
        seen = sorted(s for r in range(world) for s in scenes_for_rank(37, r, world))
        assert seen == list(range(37))
    with pytest.raises(ValueError):
        scenes_for_rank(4, 2, 2)


def test_two_rank_gather_equals_single_process():
    n_scenes = 5
    single = _plan_fn_factory()(list(range(n_scenes)))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_scenes, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [single[s] for s in range(n_scenes)]
    assert got[0] == expect and got[1] == expect
