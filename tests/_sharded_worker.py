"""Worker of tests/test_gpu_sharded_nccl.py (one process per GPU, launched with torch.distributed.run):
pik_solve_batch_sharded on this rank's shard must reproduce, on every rank, the unsharded solve."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pick_ik_b200 import capi, robots, sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")  # plumbing only: the data-path collective is the library's own NCCL
    torch.cuda.set_device(local)
    chain = robots.panda()
    n = chain.num_variables
    solver = capi.Solver(capi.Robot(chain), device=local)
    params = capi.default_params(mode="global", memetic_population_size=32, memetic_max_generations=30)
    per_rank = 96
    total = per_rank * world
    home = np.array(robots.PANDA_HOME)
    rng = np.random.default_rng(7)
    jd = chain.joint_desc()
    mv = jd[jd["type"] != 0]
    q = rng.uniform(mv["min_position"], mv["max_position"], size=(total, n))
    ident = np.zeros((total, 7))
    ident[:, 3] = 1.0
    _, _, goal = solver.eval_cost(params, ident, home, q)
    whole = solver.solve_batch(params, goal, home)  # every rank solves the whole batch as the reference

    ids = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = capi.Comm(ids[0], world, rank, local)
    a, b = sharding.shard_range(total, rank, world)
    assert b - a == per_rank
    shard_goal = np.ascontiguousarray(goal[a:b])
    for memory in (capi.MEM_HOST, capi.MEM_DEVICE):
        if memory == capi.MEM_HOST:
            gathered = np.zeros((world, per_rank, n + 3))
            solver.solve_batch_sharded_ptr(comm, params, per_rank, a, shard_goal.ctypes.data, home.ctypes.data, 0,
                                           gathered.ctypes.data, memory)
        else:
            dev = torch.device("cuda", local)
            d_goal, d_seed = torch.from_numpy(shard_goal).to(dev), torch.from_numpy(home).to(dev)
            d_g = torch.zeros((world, per_rank, n + 3), dtype=torch.float64, device=dev)
            solver.solve_batch_sharded_ptr(comm, params, per_rank, a, d_goal.data_ptr(), d_seed.data_ptr(), 0,
                                           d_g.data_ptr(), memory)
            gathered = d_g.cpu().numpy()
        got = sharding.unpack_results(gathered.reshape(total, n + 3))
        for k in ("solution", "cost", "error_code", "iterations"):
            np.testing.assert_array_equal(got[k], whole[k], err_msg=f"rank {rank} memory {memory}: {k}")
    assert 0 < (whole["error_code"] == 1).sum()

    # uneven shards (pick_ik_b200/sharding.py::shard_range on a batch that does not divide) gathered on rank 1 only
    total_u = total - 5
    counts = np.array([sharding.shard_range(total_u, r, world)[1] - sharding.shard_range(total_u, r, world)[0]
                       for r in range(world)], dtype=np.int64)
    a, b = sharding.shard_range(total_u, rank, world)
    shard_goal = np.ascontiguousarray(goal[a:b])
    root = world - 1
    block = np.full((total_u, n + 3), -7.0)
    solver.solve_batch_gather_ptr(comm, params, b - a, a, shard_goal.ctypes.data, home.ctypes.data, 0,
                                  block.ctypes.data if rank == root else 0, capi.MEM_HOST, counts=counts, root=root)
    if rank == root:
        got = sharding.unpack_results(block)
        for k in ("solution", "cost", "error_code", "iterations"):
            np.testing.assert_array_equal(got[k], whole[k][:total_u], err_msg=f"uneven gather on root: {k}")
    # an empty shard on rank 0, every rank receives
    counts0 = np.array([0] + [per_rank] * (world - 1), dtype=np.int64)
    mine = int(counts0[rank])
    a0 = int(counts0[:rank].sum())
    g0 = np.ascontiguousarray(goal[a0:a0 + mine]) if mine else np.zeros((1, 7))
    block0 = np.zeros((int(counts0.sum()), n + 3))
    solver.solve_batch_gather_ptr(comm, params, mine, a0, g0.ctypes.data, home.ctypes.data, 0, block0.ctypes.data,
                                  capi.MEM_HOST, counts=counts0, root=-1)
    got = sharding.unpack_results(block0)
    for k in ("solution", "cost", "error_code", "iterations"):
        np.testing.assert_array_equal(got[k], whole[k][:int(counts0.sum())], err_msg=f"empty shard: {k}")

    # a rank whose own solve fails (invalid parameters) still takes part: its call returns the error, the others
    # receive its rows as NaN instead of waiting forever
    bad = capi.default_params(mode="global", memetic_population_size=32, memetic_max_generations=30)
    if rank == 0:
        bad.gd_step_size = -1.0
    gathered = np.zeros((world, per_rank, n + 3))
    a, b = sharding.shard_range(total, rank, world)
    shard_goal = np.ascontiguousarray(goal[a:b])
    try:
        solver.solve_batch_sharded_ptr(comm, bad, per_rank, a, shard_goal.ctypes.data, home.ctypes.data, 0,
                                       gathered.ctypes.data, capi.MEM_HOST)
        failed = False
    except capi.PikError as e:
        failed = True
        assert e.status == -3
    assert failed == (rank == 0)
    assert np.isnan(gathered[0]).all()
    got = sharding.unpack_results(gathered[1:].reshape(-1, n + 3))
    np.testing.assert_array_equal(got["solution"], whole["solution"][per_rank:])
    comm.close()
    solver.close()
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
