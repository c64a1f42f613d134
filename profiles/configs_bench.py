"""Times the other BASELINE.json configurations (parity-test cases, not bench lines) through the C-ABI with
host buffers: configs[2] UR5 262 144 poses GD-only, configs[3] Fetch arm+torso 65 536 poses memetic pop=256
with joint-centering + avoid-limits costs, and configs[1] for reference.  usage: python profiles/configs_bench.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pick_ik_b200 import capi, robots  # noqa: E402


def uniform_configs(chain, B, seed):
    jd = chain.joint_desc()
    mv = jd[jd["type"] != 0]
    lo = np.where(mv["bounded"] != 0, mv["min_position"], -np.pi)
    hi = np.where(mv["bounded"] != 0, mv["max_position"], np.pi)
    return lo + (hi - lo) * np.random.default_rng(seed).random((B, len(mv)))


def run(name, chain, kw, B, seed, target_q):
    solver = capi.Solver(capi.Robot(chain))
    params = capi.default_params(**kw)
    ident = np.zeros((B, 7)); ident[:, 3] = 1.0
    _, _, goal = solver.eval_cost(params, ident, seed, target_q)
    solver.solve_batch(params, goal, seed)
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        res = solver.solve_batch(params, goal, seed)
        times.append(time.perf_counter() - t0)
    st = solver.stats()
    print(json.dumps({"config": name, "poses": B, "solves_per_s": B / min(times), "ms": 1e3 * min(times),
                      "device_ms": st.device_ms, "solved_frac": float((res["error_code"] == 1).mean()),
                      "kernel_launches": int(st.kernel_launches), "gd_steps": int(st.gd_steps),
                      "problem_generations": int(st.problem_generations)}))
    solver.close()


def main():
    panda = robots.panda()
    B = 65536
    run("configs[1] panda memetic pop128", panda, dict(mode="global", memetic_population_size=128), B,
        np.array(robots.PANDA_HOME), uniform_configs(panda, B, 1))
    ur5 = robots.ur5()
    B = 262144
    seed = uniform_configs(ur5, B, 2)
    run("configs[2] ur5 gd-only", ur5, dict(mode="local"), B, seed, seed + np.random.default_rng(3).uniform(-0.1, 0.1, seed.shape))
    fetch = robots.fetch()
    B = 65536
    jd = fetch.joint_desc(); mv = jd[jd["type"] != 0]
    seed = np.where(mv["bounded"] != 0, 0.5 * (mv["min_position"] + mv["max_position"]), 0.0)
    run("configs[3] fetch memetic pop256 + joint costs", fetch,
        dict(mode="global", memetic_population_size=256, center_joints_weight=0.01, avoid_joint_limits_weight=0.01,
             cost_threshold=0.01, position_threshold=0.01), B, seed, uniform_configs(fetch, B, 4))


if __name__ == "__main__":
    main()
