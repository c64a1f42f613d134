"""Generates the golden fixtures in this directory from the CPU oracle (oracle/pik_oracle.c).

The reference (pick_ik) cannot be built or imported here (it needs ROS 2 / MoveIt / Eigen), so these
vectors pin the ORACLE's behaviour -- which is itself pinned to the reference by the known-answer tests
in tests/test_oracle_reference_vectors.py -- so that (a) a change to the oracle that alters results is
caught on CPU and (b) the GPU tests can check the CUDA path against files that do not depend on the
oracle being rebuilt on the GPU box.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import orc  # noqa: E402
from pick_ik_b200 import robots  # noqa: E402

CASES = {
    # name: (robot, params, B, seed kind)
    "panda_memetic_p128": ("panda", dict(mode="global", memetic_population_size=128), 96, "home"),
    "panda_memetic_p16": ("panda", dict(mode="global", memetic_population_size=16), 128, "home"),
    "fetch_memetic_goals": ("fetch", dict(mode="global", memetic_population_size=64, center_joints_weight=0.01,
                                          avoid_joint_limits_weight=0.01, cost_threshold=0.01, position_threshold=0.01,
                                          memetic_max_generations=30), 64, "random"),
    "ur5_local": ("ur5", dict(mode="local"), 256, "perturbed"),
    "rr_local": ("rr", dict(mode="local", rotation_scale=1.0, position_threshold=1e-4), 64, "perturbed"),
    # 16 variables (the table limit), prismatic + continuous + general-axis joints, minimal-displacement goal
    "snake16_memetic": ("snake16", dict(mode="global", memetic_population_size=48, memetic_elite_size=5,
                                        memetic_max_generations=12, memetic_gd_max_iters=10, position_threshold=0.02,
                                        orientation_threshold=0.05, minimal_displacement_weight=0.01,
                                        cost_threshold=0.2), 48, "random"),
}


def inputs(name):
    robot_name, kw, B, seed_kind = CASES[name]
    chain = robots.ROBOTS[robot_name]()
    orobot = orc.build_robot(chain.joint_desc())
    if seed_kind == "home":
        seed = np.array(robots.PANDA_HOME)
        goal = orc.make_targets(orobot, B, gen_seed=0xC0FFEE)
    elif seed_kind == "random":
        seed = np.stack([orc.random_configuration(orobot, 77, b) for b in range(B)])
        goal = orc.make_targets(orobot, B, gen_seed=0xC0FFEE)
    else:
        seed = np.stack([orc.random_configuration(orobot, 78, b) for b in range(B)])
        delta = np.random.default_rng(1234).uniform(-0.1, 0.1, seed.shape)
        goal = np.stack([orc.pose_from_fk(orobot, q) for q in seed + delta])
    return chain, orobot, kw, goal, seed


def main():
    only = set(sys.argv[1:])
    for name in CASES:
        if only and name not in only:
            continue
        chain, orobot, kw, goal, seed = inputs(name)
        res = orc.solve_batch(orobot, orc.default_params(**kw), goal, seed, first_problem_index=0)
        cost, is_sol, tip = orc.eval_cost_batch(orobot, orc.default_params(**kw), goal,
                                                seed, np.broadcast_to(seed, (len(goal), orobot.n)).copy())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), goal=goal, seed=seed, solution=res["solution"],
                            error_code=res["error_code"], cost=res["cost"], iterations=res["iterations"],
                            seed_cost=cost, seed_is_solution=is_sol, seed_tip=tip)
        print(name, "solved", int((res["error_code"] == 1).sum()), "/", len(goal))


if __name__ == "__main__":
    main()
