"""The multi-species path of ik_memetic (src/ik_memetic.cpp:315-370) in the oracle: lockstep schedule, `terminate`,
and the min-fitness pick in arrival order.  CPU only."""
import ctypes as C

import numpy as np

from oracle import orc
from pick_ik_b200 import robots


def _species(pb, seed, index, n_species, stop_on_first):
    res = orc.Result()
    g = np.ascontiguousarray(seed, dtype=np.float64)
    orc.lib().orc_ik_memetic_species(C.byref(pb), g.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint32(index),
                                     C.c_int(n_species), C.c_int(stop_on_first), C.byref(res))
    return res


def test_one_species_equals_the_single_species_solver():
    chain = robots.panda()
    r = orc.build_robot(chain.joint_desc())
    home = np.array(robots.PANDA_HOME)
    p = orc.default_params(mode="global", memetic_max_generations=30)
    for b, goal in enumerate(orc.make_targets(r, 12)):
        pb = orc.make_problem(r, p, goal, home)
        a = orc.ik_memetic(pb, home, b)
        s = _species(pb, home, b, 1, 1)
        assert (a.found, a.iterations, a.cost) == (s.found, s.iterations, s.cost)
        assert list(a.solution) == list(s.solution)


def test_species_pick_rules():
    chain = robots.panda()
    r = orc.build_robot(chain.joint_desc())
    home = np.array(robots.PANDA_HOME)
    kw = dict(mode="global", memetic_max_generations=40)
    goal = orc.make_targets(r, 64)
    one = orc.solve_batch(r, orc.default_params(**kw), goal, home)
    first = orc.solve_batch(r, orc.default_params(memetic_num_threads=4, **kw), goal, home)
    all_ = orc.solve_batch(r, orc.default_params(memetic_num_threads=4, memetic_stop_on_first_solution=0, **kw), goal, home)
    ok = one["error_code"] == 1
    # species 0 is the single-species run: the race ends no later than it would have, and never fails where it solved
    assert (first["iterations"][ok] <= one["iterations"][ok]).all()
    assert (first["error_code"][ok] == 1).all()
    # without stop_on_first every species runs to its own end and the cheapest value wins
    both = (first["error_code"] == 1) & (all_["error_code"] == 1)
    assert (all_["cost"][both] <= first["cost"][both]).all()
    assert (all_["cost"][both] < first["cost"][both]).any()
    # failures hand the seed back
    for res in (first, all_):
        bad = res["error_code"] != 1
        np.testing.assert_array_equal(res["solution"][bad], np.broadcast_to(home, (int(bad.sum()), 7)))


def test_terminated_species_return_approximate_solutions():
    """With return_approximate_solution the species that are terminated return their best individuals, and a
    cheaper approximate one replaces the valid first arrival (the reference compares fitness only)."""
    chain = robots.panda()
    r = orc.build_robot(chain.joint_desc())
    home = np.array(robots.PANDA_HOME)
    kw = dict(mode="global", memetic_max_generations=40, memetic_num_threads=5, return_approximate_solution=1)
    goal = orc.make_targets(r, 48)
    res = orc.solve_batch(r, orc.default_params(**kw), goal, home)
    assert (res["error_code"] == 1).all()
    strict = orc.solve_batch(r, orc.default_params(**dict(kw, return_approximate_solution=0)), goal, home)
    ok = strict["error_code"] == 1
    assert (res["cost"][ok] <= strict["cost"][ok]).all()
    np.testing.assert_array_equal(res["iterations"][ok], strict["iterations"][ok])
