"""BASELINE.json's full-size configurations on the GPU, checked through properties that do not need the
oracle to solve the whole batch (it would take minutes on CPU), plus an oracle spot check of a random
subset of problems (bit-exact).

  configs[1]  Panda 7-DoF, 65 536 random reachable poses, memetic pop=128, 100 generations
  configs[2]  UR5 6-DoF, 262 144 poses, GD-only
  configs[3]  Fetch 8-DoF (arm + torso), joint-centering + avoid-limits costs, memetic pop=256
"""
import numpy as np
import pytest

from oracle import orc
from pick_ik_b200 import capi, robots

pytestmark = pytest.mark.gpu


def uniform_configs(chain, B, seed):
    jd = chain.joint_desc()
    mv = jd[jd["type"] != 0]
    lo = np.where(mv["bounded"] != 0, mv["min_position"], -np.pi)
    hi = np.where(mv["bounded"] != 0, mv["max_position"], np.pi)
    return lo + (hi - lo) * np.random.default_rng(seed).random((B, len(mv)))


def quat_angle(qa, qb):
    d = np.abs(np.sum(qa * qb, axis=1)) / (np.linalg.norm(qa, axis=1) * np.linalg.norm(qb, axis=1))
    return 2.0 * np.arccos(np.clip(d, -1.0, 1.0))


def check_properties(solver, params, goal, seed, res, min_solved):
    B = len(goal)
    ok = res["error_code"] == 1
    assert set(np.unique(res["error_code"])) <= {1, -31}
    assert ok.mean() >= min_solved, f"solved fraction {ok.mean():.4f}"
    seed_b = np.broadcast_to(seed, (B, solver.n))
    # failures hand the seed back (src/pick_ik_plugin.cpp:216)
    np.testing.assert_array_equal(res["solution"][~ok], seed_b[~ok])
    # every reported solution passes the engine's own solution test and an independent numpy check of
    # the frame thresholds (src/goal.cpp:27-36) on the FK of the solution
    cost, is_sol, tip = solver.eval_cost(params, goal, seed, res["solution"])
    assert (is_sol[ok] == 1).all()
    dist = np.linalg.norm(tip[:, :3] - goal[:, :3], axis=1)
    assert (dist[ok] <= params.position_threshold * (1 + 1e-9)).all()
    ang = quat_angle(tip[:, 3:], goal[:, 3:])
    assert (ang[ok] <= params.orientation_threshold + 1e-7).all()
    # the reported cost is the cost of the reported solution
    np.testing.assert_array_equal(cost[ok], res["cost"][ok])
    assert np.isfinite(res["solution"]).all()
    return ok


def spot_check(orobot, oparams, goal, seed, res, first, picks):
    for b in picks:
        sd = seed if seed.ndim == 1 else seed[b]
        ref = orc.solve_batch(orobot, oparams, goal[b:b + 1], sd, first_problem_index=first + int(b), n_threads=1)
        assert ref["error_code"][0] == res["error_code"][b]
        assert ref["iterations"][0] == res["iterations"][b]
        np.testing.assert_array_equal(ref["solution"][0], res["solution"][b])
        assert ref["cost"][0] == res["cost"][b]


def test_panda_65536_memetic_pop128():
    chain = robots.panda()
    solver = capi.Solver(capi.Robot(chain))
    kw = dict(mode="global", memetic_population_size=128)
    params = capi.default_params(**kw)
    B = 65536
    home = np.array(robots.PANDA_HOME)
    ident = np.zeros((B, 7)); ident[:, 3] = 1.0
    _, _, goal = solver.eval_cost(params, ident, home, uniform_configs(chain, B, 1))
    res = solver.solve_batch(params, goal, home)
    ok = check_properties(solver, params, goal, home, res, 0.98)
    st = solver.stats()
    assert st.solved == ok.sum() and st.problems == B
    # (100 generations + the launches that may pass part of their list on: PIK_DEFER_LAUNCHES, 12)
    assert st.generation_launches <= 112 and res["iterations"].max() <= 100
    picks = np.random.default_rng(0).choice(B, 96, replace=False)
    picks = np.concatenate([picks, np.flatnonzero(~ok)[:4]])  # a few that ran all 100 generations too
    spot_check(orc.build_robot(chain.joint_desc()), orc.default_params(**kw), goal, home, res, 0, picks)
    # idempotence: the same call again gives the same bits
    again = solver.solve_batch(params, goal, home)
    for k in ("solution", "error_code", "cost", "iterations"):
        np.testing.assert_array_equal(again[k], res[k])
    solver.close()


def test_ur5_262144_gd_only():
    chain = robots.ur5()
    solver = capi.Solver(capi.Robot(chain))
    kw = dict(mode="local")
    params = capi.default_params(**kw)
    B = 262144
    seed = uniform_configs(chain, B, 2)
    target_q = seed + np.random.default_rng(3).uniform(-0.1, 0.1, seed.shape)
    ident = np.zeros((B, 7)); ident[:, 3] = 1.0
    _, _, goal = solver.eval_cost(params, ident, seed, target_q)
    res = solver.solve_batch(params, goal, seed)
    check_properties(solver, params, goal, seed, res, 0.5)
    assert res["iterations"].max() <= params.gd_max_iters
    picks = np.random.default_rng(1).choice(B, 256, replace=False)
    spot_check(orc.build_robot(chain.joint_desc()), orc.default_params(**kw), goal, seed, res, 0, picks)
    solver.close()


def test_fetch_memetic_pop256_with_joint_costs():
    chain = robots.fetch()
    solver = capi.Solver(capi.Robot(chain))
    kw = dict(mode="global", memetic_population_size=256, center_joints_weight=0.01, avoid_joint_limits_weight=0.01,
              cost_threshold=0.01, position_threshold=0.01)
    params = capi.default_params(**kw)
    B = 65536
    jd = chain.joint_desc(); mv = jd[jd["type"] != 0]
    seed = np.where(mv["bounded"] != 0, 0.5 * (mv["min_position"] + mv["max_position"]), 0.0)  # mid-range, unbounded: 0
    ident = np.zeros((B, 7)); ident[:, 3] = 1.0
    _, _, goal = solver.eval_cost(params, ident, seed, uniform_configs(chain, B, 4))
    res = solver.solve_batch(params, goal, seed)
    check_properties(solver, params, goal, seed, res, 0.9)
    picks = np.random.default_rng(2).choice(B, 32, replace=False)
    spot_check(orc.build_robot(chain.joint_desc()), orc.default_params(**kw), goal, seed, res, 0, picks)
    solver.close()


def test_config5_shape_131072_poses_per_gpu_sharded_with_oracle_spot_check():
    """BASELINE.json configs[4] as written -- 131 072 Panda poses per GPU, sharded over the GPUs of the box, NCCL
    exchange of the packed solutions -- through bench.py, which checks problems of other ranks' shards (one GPU:
    of the batch) bit for bit against the CPU oracle and fails on any mismatch."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = min(capi.device_count(), 8)
    n = 1 << (n.bit_length() - 1)  # 1, 2, 4 or 8
    tail = ["--gpus", str(n), "--batch", "131072", "--steps", "1", "--warmup", "1", "--no-cpu-baseline",
            "--no-other-configs", "--no-pipelined", "--spot-check", "48"]
    if n == 1:
        cmd = [sys.executable, os.path.join(root, "bench.py")] + tail
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
               "127.0.0.1", "--master-port", "29533", os.path.join(root, "bench.py")] + tail
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == n and line["config"]["poses_per_gpu"] == 131072
    assert line["parity_spot_check"]["mismatches"] == 0 and line["parity_spot_check"]["problems_per_rank"] == 48
    assert line["solved_frac"] > 0.98
