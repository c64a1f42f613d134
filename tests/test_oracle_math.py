"""The oracle's own building blocks: deterministic sincos / atan2 against libm, Philox4x32-10 against
the published Random123 known-answer vectors, variable-table semantics of src/robot.cpp."""
import ctypes as C
import math

import numpy as np

from oracle import orc
from pick_ik_b200 import robots


def ulp_diff(a, b):
    if a == b:
        return 0.0
    return abs(a - b) / math.ulp(max(abs(a), abs(b), 5e-324))


def test_sincos_within_ulps_of_libm():
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-10, 10, 4000), rng.uniform(-1e4, 1e4, 2000), [0.0, 1e-300, -1e-9, math.pi, -math.pi / 2]])
    worst = 0.0
    for x in xs:
        s, c = orc.sincos(float(x))
        worst = max(worst, ulp_diff(s, math.sin(x)), ulp_diff(c, math.cos(x)))
    assert worst <= 2.0
    s, c = orc.sincos(float("nan"))
    assert math.isnan(s) and math.isnan(c)
    s, c = orc.sincos(float("inf"))
    assert math.isnan(s) and math.isnan(c)


def test_atan2_within_ulps_of_libm():
    rng = np.random.default_rng(1)
    worst = 0.0
    for _ in range(5000):
        y, x = rng.normal(size=2) * 10 ** rng.uniform(-6, 3)
        worst = max(worst, ulp_diff(orc.atan2(y, x), math.atan2(y, x)))
    assert worst <= 2.0
    assert orc.atan2(0.0, 0.0) == 0.0
    assert orc.atan2(1.0, 0.0) == math.pi / 2
    assert orc.atan2(0.0, 1.0) == 0.0
    assert math.isnan(orc.atan2(float("nan"), 1.0))


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10."""
    assert orc.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert orc.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert orc.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_random_configuration_respects_limits_and_streams():
    for name in ("panda", "fetch", "ur5"):
        rb = orc.build_robot(robots.ROBOTS[name]().joint_desc())
        a = orc.random_configuration(rb, 5, 0)
        b = orc.random_configuration(rb, 5, 1)
        assert not np.array_equal(a, b)
        np.testing.assert_array_equal(a, orc.random_configuration(rb, 5, 0))
        for i in range(rb.n):
            v = rb.vars[i]
            lo, hi = (v.min, v.max) if v.bounded else (-math.pi, math.pi)
            assert lo <= a[i] <= hi


def test_variable_table_follows_robot_from():
    """src/robot.cpp:44-85: half_span, minimal_displacement_factor normalisation, unbounded variables."""
    rb = orc.build_robot(robots.fetch().joint_desc())
    vs = [rb.vars[i] for i in range(rb.n)]
    assert [v.bounded for v in vs] == [1, 1, 1, 0, 1, 0, 1, 0]
    assert abs(sum(v.minimal_displacement_factor for v in vs) - 1.0) < 1e-12
    assert vs[3].half_span == math.pi and vs[1].half_span == (1.6056 - -1.6056) / 2.0
    rr = orc.build_robot(robots.rr().joint_desc())
    assert [rr.vars[i].minimal_displacement_factor for i in range(2)] == [0.5, 0.5]  # no velocity limits: 1 / n
    # clamp_to_limits (robot.cpp:36-42): bounded clamps to [min, max]; unbounded leaves the value alone
    assert orc.lib().orc_clamp_to_limits(C.byref(vs[1]), 9.0) == 1.6056
    assert orc.lib().orc_clamp_to_limits(C.byref(vs[3]), 9.0) == 9.0
    assert orc.lib().orc_is_valid_configuration(C.byref(rb), np.zeros(8).ctypes.data_as(C.POINTER(C.c_double))) == 1


def test_goal_cost_formulas():
    """src/goal.cpp:91-144 on the Fetch table (bounded-only loops for center / avoid-limits)."""
    rb = orc.build_robot(robots.fetch().joint_desc())
    q = np.array([0.1, 1.5, -1.0, 2.0, 2.2, -3.0, 0.3, 1.0])
    seed = np.zeros(8)
    dp = C.POINTER(C.c_double)
    center = orc.lib().orc_center_joints_cost(C.byref(rb), q.ctypes.data_as(dp))
    avoid = orc.lib().orc_avoid_joint_limits_cost(C.byref(rb), q.ctypes.data_as(dp))
    mind = orc.lib().orc_minimal_displacement_cost(C.byref(rb), q.ctypes.data_as(dp), seed.ctypes.data_as(dp))
    vs = [rb.vars[i] for i in range(8)]
    exp_center = sum(((q[i] - (v.min + v.max) / 2) * v.minimal_displacement_factor) ** 2 for i, v in enumerate(vs) if v.bounded)
    exp_avoid = sum((max(0.0, abs(q[i] - v.mid) * 2 - v.half_span) * v.minimal_displacement_factor) ** 2
                    for i, v in enumerate(vs) if v.bounded)
    exp_mind = sum(((q[i] - seed[i]) * v.minimal_displacement_factor) ** 2 for i, v in enumerate(vs))
    assert math.isclose(center, exp_center, rel_tol=1e-13)
    assert math.isclose(avoid, exp_avoid, rel_tol=1e-13) and avoid > 0
    assert math.isclose(mind, exp_mind, rel_tol=1e-13)


def test_axis_is_normalised():
    """moveit_core's setAxis normalises the joint axis; a description whose axes are 4e-7 too long must give an
    orthonormal tip rotation (and the pose of the unit-axis robot to rounding)."""
    desc = robots.skew6().joint_desc().copy()
    unit = orc.build_robot(desc)
    for j in range(len(desc)):
        desc[j]["axis"] = tuple(np.array(desc[j]["axis"]) * (1.0 + 4.0e-7))
    scaled = orc.build_robot(desc)
    q = np.array([0.3, -0.7, 1.1, 0.4, -1.3, 0.9])
    R, t = orc.fk(scaled, q)
    R0, t0 = orc.fk(unit, q)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-14
    assert np.abs(R - R0).max() < 1e-14 and np.abs(t - t0).max() < 1e-14
