"""The C-ABI library without a GPU: it loads, exports every symbol include/pik.h declares, and its
host-side entry points (parameters, robot table, status codes) behave like the reference's host code.
No compute entry point is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import orc
from pick_ik_b200 import capi, robots

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pik.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pik_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(L, name), f"libpik_b200.so does not export {name}"
    assert sorted(capi.EXPORTS) == names
    assert L.pik_version() >= 100


def test_params_defaults_match_yaml():
    """src/pick_ik_parameters.yaml defaults."""
    p = capi.default_params()
    expect = dict(mode=0, gd_step_size=0.0001, gd_max_iters=100, gd_min_cost_delta=1.0e-12, position_threshold=0.001,
                  orientation_threshold=0.001, approximate_solution_position_threshold=0.05,
                  approximate_solution_orientation_threshold=0.05, approximate_solution_joint_threshold=0.0,
                  approximate_solution_cost_threshold=0.0, cost_threshold=0.001, position_scale=1.0, rotation_scale=0.5,
                  center_joints_weight=0.0, avoid_joint_limits_weight=0.0, minimal_displacement_weight=0.0,
                  stop_optimization_on_valid_solution=1, memetic_num_threads=1, memetic_stop_on_first_solution=1,
                  memetic_population_size=16, memetic_elite_size=4, memetic_wipeout_fitness_tol=0.00001,
                  memetic_max_generations=100, memetic_gd_max_iters=25, memetic_gd_max_time=0.005)
    for k, v in expect.items():
        assert getattr(p, k) == v, k
    assert capi.validate_params(p) == capi.PIK_OK


@pytest.mark.parametrize("field,value", [
    ("mode", 2), ("gd_step_size", 1e-13), ("gd_max_iters", 0), ("gd_min_cost_delta", 0.0), ("position_threshold", -1.0),
    ("orientation_threshold", -0.1), ("cost_threshold", -1.0), ("position_scale", -1.0), ("rotation_scale", -1.0),
    ("center_joints_weight", -1.0), ("memetic_num_threads", 0), ("memetic_population_size", 0), ("memetic_elite_size", 0),
    ("memetic_wipeout_fitness_tol", -1.0), ("memetic_max_generations", 0), ("memetic_gd_max_iters", 0),
    ("memetic_gd_max_time", -1.0), ("position_threshold", float("nan")),
])
def test_params_validators(field, value):
    """the gt_eq / one_of validators of src/pick_ik_parameters.yaml"""
    p = capi.default_params(**{field: value})
    assert capi.validate_params(p) == -3  # PIK_E_INVALID_PARAMS


def test_params_elite_must_fit_population():
    assert capi.validate_params(capi.default_params(memetic_elite_size=17)) == -3
    assert capi.validate_params(capi.default_params(memetic_elite_size=40, memetic_population_size=64)) == -7
    with pytest.raises(ValueError):
        capi.default_params(mode="sideways")


@pytest.mark.parametrize("name", ["panda", "ur5", "fetch", "rr", "skew6"])
def test_robot_table_matches_oracle(name):
    """Robot::from (src/robot.cpp:44-85) as the library derives it vs the oracle's restatement."""
    chain = robots.ROBOTS[name]()
    r = capi.Robot(chain)
    o = orc.build_robot(chain.joint_desc())
    assert r.n == o.n == chain.num_variables
    for i in range(r.n):
        a, b = r.variable(i), o.vars[i]
        for f in ("min", "max", "mid", "half_span", "max_velocity_rcp", "minimal_displacement_factor", "bounded"):
            assert getattr(a, f) == getattr(b, f), (name, i, f)


def test_robot_validity_and_errors():
    r = capi.Robot(robots.panda())
    assert r.is_valid_configuration(np.array(robots.PANDA_HOME))
    assert not r.is_valid_configuration(np.zeros(7))  # joint4's upper limit is -0.0698
    with pytest.raises(ValueError):
        robots.panda("no_such_link")  # get_link_indices: unknown link (src/robot.cpp:107-120)
    h = C.c_void_p()
    assert capi.lib().pik_robot_create(None, 0, C.byref(h)) == -2
    bad = robots.panda().joint_desc().copy()
    bad[0]["axis"] = (0.0, 0.0, 0.0)
    assert capi.lib().pik_robot_create(bad.ctypes.data_as(C.c_void_p), len(bad), C.byref(h)) == -2
    only_fixed = robots.panda().joint_desc()[-2:].copy()
    assert capi.lib().pik_robot_create(only_fixed.ctypes.data_as(C.c_void_p), 2, C.byref(h)) == -2


def test_status_strings_and_no_device():
    assert capi.status_string(0) == "ok"
    assert "parameter" in capi.status_string(-3)
    if capi.device_count() == 0:
        with pytest.raises(capi.PikError) as e:
            capi.Solver(capi.Robot(robots.panda()))
        assert e.value.status == -5  # PIK_E_NO_DEVICE: the product path fails loudly without a GPU


def test_chain_signatures():
    """select_spec: the most specific compiled chain signature each fixture robot matches (host-side)."""
    expect = {"panda": "all-z 7R, x-rotation origins (static)", "ur5": "y-rotation origins", "fetch": "identity origins",
              "rr": "identity origins", "skew6": "generic", "snake16": "generic",
              "single": "generic",  # a general-axis joint: the origin-pattern kernels carry no out-of-line joint path
              "single_prismatic": "x-rotation origins"}  # no origin after the first joint: vacuously any pattern
    for name, sig in expect.items():
        assert capi.Robot(robots.ROBOTS[name]()).chain_signature() == sig, name


def test_random_configurations_match_the_oracle_stream():
    """pik_random_configurations (host) draws the bits of orc_random_configuration: the bench targets are
    reproducible from the C side and from the oracle alike."""
    from oracle import orc

    for name in ("panda", "fetch", "ur5", "snake16"):
        chain = robots.ROBOTS[name]()
        robot = capi.Robot(chain)
        orobot = orc.build_robot(chain.joint_desc())
        q = robot.random_configurations(37, 0xC0FFEE, 1000)
        ref = np.stack([orc.random_configuration(orobot, 0xC0FFEE, 1000 + b) for b in range(37)])
        np.testing.assert_array_equal(q, ref)
        assert all(robot.is_valid_configuration(row) for row in q)
