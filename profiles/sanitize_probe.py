"""Small solves for compute-sanitizer (memcheck / racecheck): throughput mapping with whole-wave launches on a tiny wave,
the wide mapping, a robot with unbounded variables (full sort), a two-tip tree, local mode.
usage: compute-sanitizer --tool memcheck python profiles/sanitize_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pick_ik_b200 import capi, robots  # noqa: E402


def run(name, B, env, **kw):
    for k, v in env.items():
        os.environ[k] = v
    robot = capi.Robot(robots.ROBOTS[name]() if name in robots.ROBOTS else robots.TREES[name]())
    solver = capi.Solver(robot)
    params = capi.default_params(**kw)
    q = robot.random_configurations(B, 0xC0FFEE)
    seed = robot.random_configurations(1, 7)[0]
    T = robot.n_tips
    ident = np.zeros((B, T, 7))
    ident[..., 3] = 1.0
    goal = solver.eval_cost(params, ident.reshape(B, -1) if T > 1 else ident.reshape(B, 7), seed, q)[2]
    res = solver.solve_batch(params, goal, seed)
    print(name, env, "solved", int((res["error_code"] == 1).sum()), "of", B, flush=True)
    solver.close()
    for k in env:
        del os.environ[k]


G = dict(mode="global", memetic_population_size=16, memetic_max_generations=6)
ONLY = set(sys.argv[1:])  # case numbers to run (default: all)
_case = [0]
_run = run


def run(*a, **k):  # noqa: F811
    _case[0] += 1
    if not ONLY or str(_case[0]) in ONLY:
        _run(*a, **k)


run("panda", 300, {"PIK_WIDE_WARPS_PER_SM": "0", "PIK_WAVE_CTAS": "1"}, **G)
run("panda", 300, {}, **G)
run("ur5", 200, {"PIK_WIDE_WARPS_PER_SM": "0"}, mode="global", memetic_population_size=20, memetic_elite_size=5, memetic_max_generations=4)
run("fetch", 150, {"PIK_WIDE_WARPS_PER_SM": "0"}, **G)
run("fetch", 150, {}, **G)
run("two_arm", 100, {}, **G)
run("panda", 400, {}, mode="local", gd_max_iters=20)
