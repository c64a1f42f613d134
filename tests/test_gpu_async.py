"""Asynchronous and concurrent use of the C-ABI: pik_solve_batch_async / pik_solver_wait, several solvers on one
device (same parameters: they share the device's constant tables and overlap; different parameters: the second
call waits for the device work that still reads the old tables), host threads solving at the same time."""
import threading

import numpy as np
import pytest

from oracle import orc
from pick_ik_b200 import capi, robots

pytestmark = pytest.mark.gpu


def _pinned(shape, dtype):
    import torch

    return torch.empty(shape, dtype=dtype).pin_memory()


def test_async_two_solvers_overlap_and_match_sync():
    import torch

    chain = robots.panda()
    orobot = orc.build_robot(chain.joint_desc())
    robot = capi.Robot(chain)
    n = robot.n
    params = capi.default_params(mode="global", memetic_population_size=32, memetic_max_generations=40)
    home = np.array(robots.PANDA_HOME)
    B = 700
    goals = [orc.make_targets(orobot, B, first=k * B) for k in range(2)]
    solvers = [capi.Solver(robot) for _ in range(2)]
    sync = [solvers[0].solve_batch(params, goals[k], home, first_problem_index=k * B) for k in range(2)]
    h_goal = [_pinned((B, 7), torch.float64) for _ in range(2)]
    h_seed = _pinned((n,), torch.float64)
    h_seed.numpy()[:] = home
    outs = []
    for k in range(2):
        h_goal[k].numpy()[:] = goals[k]
        outs.append(dict(solution=_pinned((B, n), torch.float64), error_code=_pinned((B,), torch.int32),
                         cost=_pinned((B,), torch.float64), iterations=_pinned((B,), torch.int32)))
    for rep in range(3):
        for k in range(2):
            o = outs[k]
            solvers[k].solve_batch_async_ptr(params, B, k * B, h_goal[k].data_ptr(), h_seed.data_ptr(), 0,
                                             o["solution"].data_ptr(), o["error_code"].data_ptr(),
                                             o["cost"].data_ptr(), o["iterations"].data_ptr(), capi.MEM_HOST)
        # one solve in flight per solver
        with pytest.raises(capi.PikError) as ei:
            solvers[0].solve_batch_async_ptr(params, B, 0, h_goal[0].data_ptr(), h_seed.data_ptr(), 0,
                                             outs[0]["solution"].data_ptr(), outs[0]["error_code"].data_ptr(),
                                             outs[0]["cost"].data_ptr(), outs[0]["iterations"].data_ptr(), capi.MEM_HOST)
        assert ei.value.status == -9
        for k in range(2):
            solvers[k].wait()
            assert solvers[k].query()
            for key in ("solution", "error_code", "cost", "iterations"):
                np.testing.assert_array_equal(outs[k][key].numpy(), sync[k][key], err_msg=f"rep {rep} solver {k} {key}")
            st = solvers[k].stats()
            assert st.problems == B and st.solved == (sync[k]["error_code"] == 1).sum()
    for s in solvers:
        s.close()


def test_host_threads_with_different_parameters():
    """Two host threads, each with its own solver and its own parameters and robot, on the same device: the
    constant tables are replaced back and forth; every result must equal the oracle's."""
    cases = [("panda", dict(mode="global", memetic_population_size=16, memetic_max_generations=25), 150),
             ("ur5", dict(mode="local"), 400)]
    refs, errors = {}, []
    setups = []
    for name, kw, B in cases:
        chain = robots.ROBOTS[name]()
        orobot = orc.build_robot(chain.joint_desc())
        goal = orc.make_targets(orobot, B)
        seed = np.array(robots.PANDA_HOME) if name == "panda" else np.stack(
            [orc.random_configuration(orobot, 77, b) for b in range(B)])
        refs[name] = orc.solve_batch(orobot, orc.default_params(**kw), goal, seed)
        setups.append((name, chain, kw, goal, seed))

    def work(name, chain, kw, goal, seed):
        try:
            solver = capi.Solver(capi.Robot(chain))
            params = capi.default_params(**kw)
            for _ in range(6):
                got = solver.solve_batch(params, goal, seed)
                for key in ("solution", "error_code", "cost", "iterations"):
                    np.testing.assert_array_equal(got[key], refs[name][key], err_msg=f"{name} {key}")
            solver.close()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=su) for su in setups]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_host_threads_one_per_device():
    """A host thread per GPU in one process (include/pik.h): every device has its own constant tables."""
    n_dev = capi.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    chain = robots.panda()
    orobot = orc.build_robot(chain.joint_desc())
    kw = dict(mode="global", memetic_population_size=16, memetic_max_generations=25)
    B = 120
    goal = orc.make_targets(orobot, B)
    home = np.array(robots.PANDA_HOME)
    ref = orc.solve_batch(orobot, orc.default_params(**kw), goal, home)
    errors = []

    def work(dev):
        try:
            solver = capi.Solver(capi.Robot(chain), device=dev)
            got = solver.solve_batch(capi.default_params(**kw), goal, home)
            for key in ("solution", "error_code", "cost", "iterations"):
                np.testing.assert_array_equal(got[key], ref[key])
            solver.close()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(d,)) for d in range(min(n_dev, 4))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
