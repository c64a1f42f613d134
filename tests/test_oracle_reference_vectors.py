"""Pins the CPU oracle (oracle/pik_oracle.c) to every known-answer the reference's own tests hold for
the hot path.  Each test names the reference test it ports (paths relative to pick_ik @ 8c99999).
Catch::Approx's default tolerance is eps = 100 * FLT_EPSILON relative (scale 0)."""
import math

import numpy as np
import pytest

from oracle import orc
from pick_ik_b200 import robots

APPROX_EPS = 100 * 1.1920929e-07


def approx(value, expected, margin=0.0):
    return abs(value - expected) <= max(margin, APPROX_EPS * abs(expected))


def quat_aa(angle, axis):
    """Eigen::AngleAxisd -> quaternion (w, x, y, z)."""
    ax = np.asarray(axis, dtype=float)
    return np.concatenate([[math.cos(angle / 2)], math.sin(angle / 2) * ax])


ZERO = orc.frame((0, 0, 0), (1, 0, 0, 0))


# ---------------------------------------------------------------------------------------------
# tests/goal_tests.cpp
# ---------------------------------------------------------------------------------------------
class TestMakeFrameTests:
    """tests/goal_tests.cpp:9-71 (pick_ik::make_frame_tests)"""

    pe, oe = 0.00001, 0.001

    def test_zero_threshold(self):  # :20-23
        assert orc.frame_test(ZERO, ZERO, 0.0, 0.0)

    def test_almost_but_not_quite(self):  # :25-33
        f = orc.frame((self.pe,) * 3, (1 - self.oe, 0.0, 0.0, self.oe))
        assert not orc.frame_test(ZERO, f, self.pe, self.oe)

    def test_within_position_not_orientation(self):  # :35-42
        f = orc.frame((0.0, 0.000009, 0.0), (0.707, 0.0, 0.707, 0.0))
        assert not orc.frame_test(ZERO, f, self.pe, self.oe)

    def test_within_threshold(self):  # :44-51
        f = orc.frame((0.0, 0.000009, 0.0), (0.99999, 0.0, 0.0, 0.00001))
        assert orc.frame_test(ZERO, f, self.pe, self.oe)

    def test_goal_is_frame(self):  # :53-57
        assert orc.frame_test(ZERO, ZERO, self.pe, self.oe)

    def test_orientation_differs(self):  # :59-70
        rot = orc.frame((0, 0, 0), quat_aa(math.pi / 4, (0, 0, 1)))
        assert not orc.frame_test(ZERO, rot, self.pe, self.oe)
        assert orc.frame_test(ZERO, rot, self.pe, None)  # position-only frame test


class TestMakePoseCostFn:
    """tests/goal_tests.cpp:73-226 (pick_ik::make_pose_cost_fn)"""

    y2 = orc.frame((0, 2, 0), (1, 0, 0, 0))
    xy1 = orc.frame((1, 1, 0), (1, 0, 0, 0))
    xyz1 = orc.frame((1, 1, 1), (1, 0, 0, 0))
    rx1 = orc.frame((0, 0, 0), quat_aa(1.0, (1, 0, 0)))
    ry2 = orc.frame((0, 0, 0), quat_aa(2.0, (0, 1, 0)))

    @pytest.mark.parametrize("ps,rs", [(0.0, 0.0), (1.0, 0.0), (1.0, 0.5)])
    def test_goal_is_frame(self, ps, rs):  # :87-100
        assert approx(orc.pose_cost(ZERO, ZERO, ps, rs), 0.0)

    def test_goal_is_second_index(self):  # :102-105 (the frame at the goal's index is compared)
        assert approx(orc.pose_cost(self.y2, self.y2, 1.0, 0.0), 0.0)

    def test_translation_squares(self):  # :107-124
        assert approx(orc.pose_cost(ZERO, self.y2, 1.0, 0.5), 4.0)
        assert approx(orc.pose_cost(ZERO, self.xy1, 1.0, 0.5), 2.0)
        assert approx(orc.pose_cost(ZERO, self.xyz1, 1.0, 0.5), 3.0)

    def test_zero_scales(self):  # :126-134
        assert approx(orc.pose_cost(ZERO, self.xyz1, 0.0, 0.5), 0.0)
        assert approx(orc.pose_cost(ZERO, self.rx1, 1.0, 0.0), 0.0)

    def test_negative_scales_equal_zero_scales(self):  # :136-146
        assert orc.pose_cost(ZERO, self.xyz1, 0.0, 0.5) == orc.pose_cost(ZERO, self.xyz1, -1.0, 0.5)
        assert orc.pose_cost(ZERO, self.rx1, 1.0, 0.0) == orc.pose_cost(ZERO, self.rx1, 1.0, -0.5)

    def test_rotation_one_axis(self):  # :148-154
        assert approx(orc.pose_cost(ZERO, self.ry2, 1.0, 1.0), 4.0)

    def test_rotation_one_axis_scaled(self):  # :156-169
        assert approx(orc.pose_cost(ZERO, self.ry2, 1.0, 0.5), 4.0 * 0.25)

    Q_GOAL = (3.2004117980888137e-12, 0.9239557003781338, -0.38249949508300274, 1.324932598914536e-12)
    BIO_IK = [
        # (goal t, frame q, frame t)   tests/goal_tests.cpp:171-198 ("Test 0") and :200-226 ("Test 2")
        ((0.3548182547092438, -0.04776066541671753, 0.5902695655822754),
         (-0.0033032628064278945, 0.9163043570028795, -0.40044067474764505, -0.004762331364117075),
         (0.3363926217416014, -0.043807946580255344, 0.5864240526436293)),
        ((0.3327501714229584, -0.025710120797157288, 0.5902695655822754),
         (2.1223489422435532e-07, 0.9239554647443051, -0.38250006378889556, 1.925047999919496e-05),
         (0.3327318727877646, -0.02570328270961634, 0.5900141633600922)),
    ]

    @pytest.mark.parametrize("case", [0, 1])
    def test_bio_ik_goldens(self, case):
        goal_t, q_frame, frame_t = self.BIO_IK[case]
        goal = orc.frame(goal_t, self.Q_GOAL)
        frame = orc.frame(frame_t, q_frame)
        ps, rs = 1.0, 0.5
        dot = float(np.dot(self.Q_GOAL, q_frame))
        expected = float(np.sum((np.array(goal_t) - np.array(frame_t)) ** 2)) + (2.0 * math.acos(dot) * rs) ** 2
        assert approx(orc.pose_cost(goal, frame, ps, rs), expected)

    def test_pose_cost_functions(self):  # :228-275: one function per goal, each on the frame of its index
        goal = orc.frame(self.BIO_IK[1][0], self.Q_GOAL)
        frame = orc.frame(self.BIO_IK[1][2], self.BIO_IK[1][1])
        assert approx(orc.pose_cost(goal, goal, 1.0, 0.5), 0.0, margin=1e-15)
        assert approx(orc.pose_cost(frame, frame, 1.0, 0.5), 0.0, margin=1e-15)


# ---------------------------------------------------------------------------------------------
# tests/robot_tests.cpp
# ---------------------------------------------------------------------------------------------
def test_robot_variable_counts():
    """tests/robot_tests.cpp:85-108: the RR robot has two variables, the Panda arm seven."""
    assert orc.build_robot(robots.rr(1.0).joint_desc()).n == 2
    assert orc.build_robot(robots.panda().joint_desc()).n == 7


# ---------------------------------------------------------------------------------------------
# tests/ik_tests.cpp
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def rr():
    return orc.build_robot(robots.rr(2.0).joint_desc())


@pytest.fixture(scope="module")
def panda():
    return orc.build_robot(robots.panda().joint_desc())


def test_rr_fk(rr):
    """tests/ik_tests.cpp:50-75 (RR model FK)"""
    _, t = orc.fk(rr, [0.0, 0.0])
    assert approx(t[0], 3.0) and approx(t[1], 0.0)
    _, t = orc.fk(rr, [math.pi / 4, -math.pi / 4])
    assert approx(t[0], 2.0 * math.cos(math.pi / 4) + 1.0, margin=0.001)
    assert approx(t[1], 2.0 * math.sin(math.pi / 4), margin=0.001)


def ik_test_params(**kw):
    """IkTestParams, tests/ik_tests.cpp:78-86, with GradientIkParams defaults (ik_gradient.hpp:15-23)."""
    base = dict(mode="local", position_threshold=0.0001, orientation_threshold=0.001, cost_threshold=0.0001,
                position_scale=1.0, rotation_scale=1.0, gd_max_iters=100, gd_step_size=0.0001, gd_min_cost_delta=1e-12)
    base.update(kw)
    return orc.default_params(**base)


def solve_ik_test(robot, goal_pose, guess, **kw):
    p = ik_test_params(**kw)
    return orc.ik_gradient(orc.make_problem(robot, p, goal_pose, guess), guess)


S45 = math.sin(math.pi / 4)
RR_GOAL_A = [3.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0]
RR_GOAL_B = [S45, 3.0 * S45, 0.0] + list(quat_aa(0.75 * math.pi, (0, 0, 1)))


@pytest.mark.parametrize("goal,expected,guess", [
    (RR_GOAL_A, (0.0, 0.0), (0.1, -0.1)),                                # :140-151
    (RR_GOAL_A, (0.0, 0.0), (math.pi / 2, -math.pi / 2)),                # :153-164
    (RR_GOAL_B, (math.pi / 4, math.pi / 2), (math.pi / 4 + 0.1, math.pi / 2 - 0.1)),  # :166-178
    (RR_GOAL_B, (math.pi / 4, math.pi / 2), (0.0, 0.0)),                # :180-192
])
def test_rr_ik(rr, goal, expected, guess):
    res = solve_ik_test(rr, goal, list(guess))
    assert res.found
    assert approx(res.solution[0], expected[0], margin=0.01)
    assert approx(res.solution[1], expected[1], margin=0.01)


def test_rr_unreachable(rr):  # :194-203
    assert not solve_ik_test(rr, [0, 0, 0, 1, 0, 0, 0], [0.0, 0.0]).found


def test_rr_reachable_position_not_orientation(rr):  # :205-216
    assert not solve_ik_test(rr, [S45, 3.0 * S45, 0.0, 1, 0, 0, 0], [0.0, 0.0]).found


def test_rr_zero_rotation_scale(rr):  # :218-233
    res = solve_ik_test(rr, [S45, 3.0 * S45, 0.0, 1, 0, 0, 0], [math.pi / 4 + 0.1, math.pi / 2 - 0.1],
                        rotation_scale=0.0)
    assert res.found
    assert approx(res.solution[0], math.pi / 4, margin=0.01)
    assert approx(res.solution[1], math.pi / 2, margin=0.01)


HOME = list(robots.PANDA_HOME)


def test_panda_ik_exact_home(panda):  # :252-268
    res = solve_ik_test(panda, orc.pose_from_fk(panda, HOME), HOME, rotation_scale=0.5)
    assert res.found
    for i in range(7):
        assert approx(res.solution[i], HOME[i], margin=0.01)


def test_panda_ik_perturbed_home(panda):  # :270-292
    actual = [0.1, -math.pi / 4 - 0.1, 0.1, -3.0 * math.pi / 4 - 0.1, 0.1, math.pi / 2 - 0.1, math.pi / 4 + 0.1]
    res = solve_ik_test(panda, orc.pose_from_fk(panda, actual), HOME, rotation_scale=0.5)
    assert res.found
    for i in range(7):
        assert approx(res.solution[i], actual[i], margin=0.025)  # "Note the extra tolerance..."


def test_panda_home_pose_matches_bio_ik_goal(panda):
    """The bio_ik-instrumented goal of tests/goal_tests.cpp:172-177 is the Panda home pose of panda_link8
    (z = 0.59027, q ~ (0, 0.92396, -0.38250, 0)): pins the chain table itself."""
    l8 = orc.build_robot(robots.panda("panda_link8").joint_desc())
    pose = orc.pose_from_fk(l8, HOME)
    assert abs(pose[2] - 0.5902695655822754) < 2e-5
    assert abs(abs(pose[4]) - 0.9239557003781338) < 5e-4 and abs(abs(pose[5]) - 0.38249949508300274) < 5e-4


# ---------------------------------------------------------------------------------------------
# tests/ik_memetic_tests.cpp
# ---------------------------------------------------------------------------------------------
def isometry_is_approx(pose_a, pose_b, prec):
    """Eigen isApprox on the 4x4: |A - B|_F^2 <= prec^2 * min(|A|_F^2, |B|_F^2)."""
    def mat(p):
        m = np.eye(4)
        m[:3, :3] = orc.quat_to_matrix(p[3:])
        m[:3, 3] = p[:3]
        return m
    a, b = mat(pose_a), mat(pose_b)
    return np.sum((a - b) ** 2) <= prec ** 2 * min(np.sum(a ** 2), np.sum(b ** 2))


def memetic_params(**kw):
    """MemeticIkTestParams (tests/ik_memetic_tests.cpp:14-33) over MemeticIkParams / GradientIkParams
    defaults (ik_memetic.hpp:26-45, ik_gradient.hpp:15-23)."""
    base = dict(mode="global", position_threshold=0.001, orientation_threshold=0.001, cost_threshold=0.001,
                position_scale=1.0, rotation_scale=0.5, memetic_population_size=16, memetic_elite_size=4,
                memetic_max_generations=100, memetic_gd_max_iters=100, memetic_wipeout_fitness_tol=0.00001)
    base.update(kw)
    return orc.default_params(**base)


MEMETIC_SECTIONS = [
    ("home", HOME, {}),                                                                      # :121-135
    ("near home", [h + o for h, o in zip(HOME, [0.1, -0.1, 0.0, 0.1, -0.1, 0.0, 0.1])], {}),  # :137-155
    ("zeros", [0.0] * 7, {}),                                                                # :157-172, :174-190
    # :192-213; the test weights all three goals with center_joints_weight (copy-paste in the reference, :55-63)
    ("zeros with goals", [0.0] * 7, dict(center_joints_weight=0.01, avoid_joint_limits_weight=0.01,
                                         cost_threshold=0.01, position_threshold=0.01)),
]


@pytest.mark.parametrize("label,guess,kw", MEMETIC_SECTIONS, ids=[s[0] for s in MEMETIC_SECTIONS])
@pytest.mark.parametrize("stream", [0, 1, 2, 3])  # the reference's RNG is unseeded: any stream must succeed
def test_panda_memetic(panda, label, guess, kw, stream):
    p = memetic_params(**kw)
    goal = orc.pose_from_fk(panda, HOME)
    res = orc.ik_memetic(orc.make_problem(panda, p, goal, guess), guess, problem_index=stream)
    assert res.found
    final = orc.pose_from_fk(panda, list(res.solution)[:7])
    assert isometry_is_approx(goal, final, p.position_threshold)
