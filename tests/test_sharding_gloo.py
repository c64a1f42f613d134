"""Multi-rank host logic on CPU: world_size-2 gloo.  Each rank solves its contiguous shard of the pose
batch (here with the CPU oracle standing in for the GPU engine -- the host-side sharding, packing and
gather code is what is under test) and the all-gathered result must equal the unsharded solve."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from pick_ik_b200 import sharding  # noqa: E402


def test_shard_range_partitions():
    for total in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_pack_roundtrip():
    rng = np.random.default_rng(0)
    sol = rng.normal(size=(5, 7))
    err = np.array([1, -31, 1, 1, -31], dtype=np.int32)
    cost = rng.random(5)
    its = np.array([0, 100, 3, 7, 100], dtype=np.int32)
    un = sharding.unpack_results(sharding.pack_results(sol, err, cost, its))
    np.testing.assert_array_equal(un["solution"], sol)
    np.testing.assert_array_equal(un["error_code"], err)
    np.testing.assert_array_equal(un["cost"], cost)
    np.testing.assert_array_equal(un["iterations"], its)


def _worker(rank, world, port, total, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import orc
    from pick_ik_b200 import robots, sharding as sh

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chain = robots.panda()
    orobot = orc.build_robot(chain.joint_desc())
    params = orc.default_params(mode="global", memetic_population_size=16, memetic_max_generations=20)
    goal = orc.make_targets(orobot, total)
    home = np.array(robots.PANDA_HOME)
    a, b = sh.shard_range(total, rank, world)
    res = orc.solve_batch(orobot, params, goal[a:b], home, first_problem_index=a, n_threads=2)
    packed = torch.from_numpy(sh.pack_results(res["solution"], res["error_code"], res["cost"], res["iterations"]))
    max_rows = max(sh.shard_range(total, r, world)[1] - sh.shard_range(total, r, world)[0] for r in range(world))
    gathered = sh.all_gather_results(packed, world, max_rows)
    full = sh.concat_shards(gathered, total, world).numpy()
    if rank == 0:
        ref = orc.solve_batch(orobot, params, goal, home, first_problem_index=0, n_threads=2)
        un = sh.unpack_results(full)
        ok = all(np.array_equal(un[k], ref[k]) for k in ("solution", "error_code", "cost", "iterations"))
        with open(os.path.join(out_dir, "ok"), "w") as f:
            f.write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_unsharded(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, 37, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "1"
