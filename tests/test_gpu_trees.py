"""GPU parity for kinematic trees, several tips, floating / planar / mimic joints (the tree kernels) against the
CPU oracle: bit-equal costs, flags, tip poses, joint values and iteration counts."""
import numpy as np
import pytest

from oracle import orc
from pick_ik_b200 import capi, robots

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trees():
    cache = {}

    def get(name):
        if name not in cache:
            tree = robots.TREES[name]()
            cache[name] = (tree, orc.build_robot_tree(*tree.tree_arrays()), capi.Solver(capi.Robot(tree)))
        return cache[name]

    yield get
    for _, _, s in cache.values():
        s.close()


def random_configs(orobot, B, seed):
    return np.stack([orc.random_configuration(orobot, seed, b) for b in range(B)])


def check(got, ref, label):
    for k in ("error_code", "iterations", "solution", "cost"):
        np.testing.assert_array_equal(got[k], ref[k], err_msg=f"{label}: {k}")


@pytest.mark.parametrize("name", list(robots.TREES))
@pytest.mark.parametrize("goals", [False, True])
def test_tree_eval_cost_bit_exact(trees, name, goals):
    tree, orobot, solver = trees(name)
    kw = dict(center_joints_weight=0.3, avoid_joint_limits_weight=0.7, minimal_displacement_weight=0.2,
              cost_threshold=0.05, position_threshold=0.2, orientation_threshold=0.3) if goals else {}
    op, gp = orc.default_params(**kw), capi.default_params(**kw)
    B = 1500
    q = random_configs(orobot, B, 11)
    tq = random_configs(orobot, B, 12)
    near = np.arange(B) % 3 == 0
    tq[near] = q[near] + (0.01 if goals else 1e-6)
    goal = np.stack([orc.poses_from_fk(orobot, t) for t in tq])
    if orobot.n_tips == 1:
        goal = goal[:, 0]
    seed = random_configs(orobot, B, 13)
    c_ref, s_ref, tip_ref = orc.eval_cost_batch(orobot, op, goal, seed, q)
    c, s, tip = solver.eval_cost(gp, goal, seed, q)
    np.testing.assert_array_equal(c, c_ref)
    np.testing.assert_array_equal(s, s_ref)
    np.testing.assert_array_equal(tip, tip_ref)
    assert s_ref.sum() < B
    if not goals:
        assert 0 < s_ref.sum()


@pytest.mark.parametrize("name", list(robots.TREES))
def test_tree_gd_local_parity(trees, name):
    tree, orobot, solver = trees(name)
    op, gp = orc.default_params(mode="local"), capi.default_params(mode="local")
    B = 600
    seed = random_configs(orobot, B, 21)
    tq = seed + np.random.default_rng(5).uniform(-0.05, 0.05, seed.shape)
    goal = np.stack([orc.poses_from_fk(orobot, t) for t in tq])
    if orobot.n_tips == 1:
        goal = goal[:, 0]
    ref = orc.solve_batch(orobot, op, goal, seed)
    check(solver.solve_batch(gp, goal, seed), ref, name)
    assert 0 < (ref["error_code"] == 1).sum()


TREE_MEMETIC = [
    ("two_arm", dict(memetic_population_size=24, memetic_max_generations=25), 120),
    ("three_tip", dict(memetic_population_size=16, memetic_elite_size=3, memetic_max_generations=15), 100),
    ("floating_arm", dict(memetic_population_size=16, memetic_max_generations=8, position_threshold=0.01,
                          orientation_threshold=0.02), 60),
    ("planar_arm", dict(memetic_population_size=32, memetic_max_generations=20, minimal_displacement_weight=0.01,
                        cost_threshold=0.1), 100),
    ("mimic_arm", dict(memetic_population_size=16, memetic_elite_size=2, memetic_max_generations=20, rotation_scale=0.0), 100),
    ("two_arm", dict(memetic_population_size=16, memetic_max_generations=12, memetic_num_threads=3), 60),
]


@pytest.mark.parametrize("mapping", ["throughput", "wide", "default"])
@pytest.mark.parametrize("name,kw,B", TREE_MEMETIC)
def test_tree_memetic_parity(trees, name, kw, B, mapping, monkeypatch):
    if mapping != "default":
        monkeypatch.setenv("PIK_WIDE_WARPS_PER_SM", "0" if mapping == "throughput" else "1000000000")
    tree, orobot, solver = trees(name)
    op, gp = orc.default_params(mode="global", **kw), capi.default_params(mode="global", **kw)
    goal = orc.make_targets(orobot, B)
    seed = random_configs(orobot, B, 41)
    ref = orc.solve_batch(orobot, op, goal, seed, first_problem_index=300)
    check(solver.solve_batch(gp, goal, seed, first_problem_index=300), ref, f"{name} {kw}")
    assert 0 < (ref["error_code"] == 1).sum()
