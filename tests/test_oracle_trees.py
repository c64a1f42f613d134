"""Kinematic trees, several tips, floating / planar / mimic joints in the oracle (CPU).  The reference evaluates one
pose cost and one frame test per tip over MoveIt's whole-tree FK (src/goal.cpp:80-89,163-175,188-203,
src/fk_moveit.cpp:20-34) and documents the joint frames in src/forward_kinematics.cpp:39-80; it holds no test vectors
for them, so the tree path is pinned against the serial-chain path (itself pinned to the reference's vectors) and
against closed forms."""
import math

import numpy as np

from oracle import orc
from pick_ik_b200 import capi, robots
from pick_ik_b200.robots import JOINT_FIXED, JOINT_REVOLUTE, Joint, RobotChain


def _chain_of(tree, tip_index, name):
    """The serial chain model root -> tip of a tree, as a RobotChain."""
    j = tree.tip_joints[tip_index]
    path = []
    while j >= 0:
        path.append(j)
        j = tree.joints[j].parent
    joints = [tree.joints[k] for k in reversed(path)]
    return RobotChain(name, "root", "tip", joints), list(reversed(path))


def test_every_tip_of_a_tree_equals_its_serial_chain_bit_for_bit():
    for name in ("two_arm", "three_tip"):
        tree = robots.TREES[name]()
        o = orc.build_robot_tree(*tree.tree_arrays())
        var_of_joint, v = {}, 0
        for k, j in enumerate(tree.joints):
            if j.type != JOINT_FIXED:
                var_of_joint[k] = v
                v += 1
        rng = np.random.default_rng(3)
        for tip in range(tree.num_tips):
            chain, path = _chain_of(tree, tip, f"{name}_{tip}")
            oc = orc.build_robot(chain.joint_desc())
            cols = [var_of_joint[k] for k in path if k in var_of_joint]
            for _ in range(50):
                q = rng.uniform(-2.0, 2.0, o.n)
                poses = orc.poses_from_fk(o, q)
                np.testing.assert_array_equal(poses[tip], orc.pose_from_fk(oc, q[cols]), err_msg=f"{name} tip {tip}")


def test_multi_tip_cost_is_the_sum_of_the_tips_pose_costs_and_every_frame_test_must_pass():
    tree = robots.two_arm()
    o = orc.build_robot_tree(*tree.tree_arrays())
    p = orc.default_params(position_threshold=0.05, orientation_threshold=0.05)
    rng = np.random.default_rng(4)
    for _ in range(30):
        q, qt = rng.uniform(-1.4, 1.4, o.n), rng.uniform(-1.4, 1.4, o.n)
        goal = orc.poses_from_fk(o, qt)
        pb = orc.make_problem(o, p, goal, q)
        tips = orc.poses_from_fk(o, q)
        total = 0.0
        for t in range(2):
            total = total + orc.pose_cost(orc.frame(goal[t, :3], goal[t, 3:]), orc.frame(tips[t, :3], tips[t, 3:]), 1.0, 0.5)
        # (the tips' frames went through a quaternion on the way here: equal up to rounding)
        assert math.isclose(orc.cost(pb, q), total, rel_tol=1e-12)
        assert orc.is_solution(pb, qt)
        # left arm at its goal, right arm not: not a solution
        q_half = qt.copy()
        q_half[4:] = q[4:]
        assert not orc.is_solution(pb, q_half)


def test_floating_and_planar_joint_frames():
    """src/forward_kinematics.cpp:64-79: floating = Translation3d(v0 v1 v2) * Quaterniond(w = v6, x = v3, y = v4, z = v5);
    planar = Translation3d(x, y, 0) * rotation about z."""
    fl = orc.build_robot_tree(*robots.floating_arm().tree_arrays())
    assert fl.n == 9
    q = np.zeros(9)
    q[:3] = (0.3, -0.2, 0.5)
    c, s = math.cos(0.35), math.sin(0.35)
    q[3:7] = (0.0, 0.0, s, c)  # x y z w: a rotation of 0.7 rad about z
    pose = orc.pose_from_fk(fl, q)
    # chain: origin (0,0,0.1) * floating * (0,0,0.2) * Ry(0) * (0.3,0,0) * Rz(0) * (0.2,0,0)
    reach = 0.5
    np.testing.assert_allclose(pose[:3], (0.3 + reach * math.cos(0.7), -0.2 + reach * math.sin(0.7), 0.1 + 0.5 + 0.2), atol=1e-12)
    np.testing.assert_allclose(pose[3:], (c, 0, 0, s), atol=1e-12)
    for i in range(3):
        assert not fl.vars[i].bounded or (fl.vars[i].min, fl.vars[i].max) == (-1.0, 1.0)
    for i in range(3, 7):
        assert fl.vars[i].bounded and (fl.vars[i].min, fl.vars[i].max) == (-1.0, 1.0)

    pl = orc.build_robot_tree(*robots.planar_arm().tree_arrays())
    assert pl.n == 6 and not pl.vars[0].bounded and not pl.vars[2].bounded and pl.vars[2].half_span == math.pi
    q = np.array([1.0, 2.0, 0.5, 0.0, 0.0, 0.0])
    pose = orc.pose_from_fk(pl, q)
    reach = 0.1 + 0.3 + 0.25 + 0.1
    np.testing.assert_allclose(pose[:3], (1.0 + reach * math.cos(0.5), 2.0 + reach * math.sin(0.5), 0.05 + 0.3), atol=1e-12)
    np.testing.assert_allclose(pose[3:], (math.cos(0.25), 0, 0, math.sin(0.25)), atol=1e-12)


def test_mimic_joint_follows_its_master_and_owns_no_variable():
    tree = robots.mimic_arm()
    o = orc.build_robot_tree(*tree.tree_arrays())
    assert o.n == 3 and o.n_steps == 4
    # the same arm with the mimic joint as an ordinary joint driven by hand
    plain = RobotChain("plain", "root", "tip", [Joint(j.name, j.type, j.xyz, j.rpy, j.axis, j.lower, j.upper, j.velocity)
                                               for j in tree.joints])
    oc = orc.build_robot(plain.joint_desc())
    rng = np.random.default_rng(5)
    for _ in range(40):
        q = rng.uniform(-1.5, 1.5, 3)
        q4 = np.array([q[0], q[1], q[1] * -0.5 + 0.1, q[2]])
        np.testing.assert_array_equal(orc.pose_from_fk(o, q), orc.pose_from_fk(oc, q4))


def test_tree_robot_tables_match_between_library_and_oracle():
    for name, make in robots.TREES.items():
        tree = make()
        r = capi.Robot(tree)
        o = orc.build_robot_tree(*tree.tree_arrays())
        assert r.n == o.n == tree.num_variables and r.n_tips == o.n_tips == tree.num_tips
        assert r.chain_signature() == "tree"
        for i in range(r.n):
            a, b = r.variable(i), o.vars[i]
            for f in ("min", "max", "mid", "half_span", "max_velocity_rcp", "minimal_displacement_factor", "bounded"):
                assert getattr(a, f) == getattr(b, f), (name, i, f)
        np.testing.assert_array_equal(r.random_configurations(9, 7, 100),
                                      np.stack([orc.random_configuration(o, 7, 100 + b) for b in range(9)]))


def test_trees_solve_on_the_oracle():
    for name in ("two_arm", "three_tip", "planar_arm", "mimic_arm"):
        o = orc.build_robot_tree(*robots.TREES[name]().tree_arrays())
        goal = orc.make_targets(o, 16)
        seed = np.stack([orc.random_configuration(o, 5, b) for b in range(16)])
        res = orc.solve_batch(o, orc.default_params(mode="global", memetic_max_generations=60), goal, seed)
        assert (res["error_code"] == 1).mean() >= 0.8, name
        pr = orc.default_params(mode="global")
        for b in np.flatnonzero(res["error_code"] == 1):
            assert orc.is_solution(orc.make_problem(o, pr, goal[b], seed[b]), res["solution"][b])
