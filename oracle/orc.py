"""ctypes wrapper around the CPU oracle (oracle/pik_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by pick_ik_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
MAX_VARS = 16
MAX_STEPS = 16
MAX_TIPS = 4


def _source_hash() -> str:
    import hashlib

    h = hashlib.sha256()
    for f in ("pik_oracle.c", "pik_oracle.h", "Makefile"):
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False) -> str:
    """Compiles the oracle when its sources changed (content hash: mtimes do not survive a copy)."""
    stamp = LIB_PATH + ".hash"
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == _source_hash():
                return LIB_PATH
    subprocess.check_call(["make", "-C", HERE, "-B", "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    with open(stamp, "w") as fh:
        fh.write(_source_hash())
    return LIB_PATH


class Variable(C.Structure):
    _fields_ = [("min", C.c_double), ("max", C.c_double), ("mid", C.c_double),
                ("half_span", C.c_double), ("max_velocity_rcp", C.c_double),
                ("minimal_displacement_factor", C.c_double), ("bounded", C.c_int32), ("pad_", C.c_int32)]


class Step(C.Structure):
    _fields_ = [("kind", C.c_int32), ("parent", C.c_int32), ("var0", C.c_int32), ("pad_", C.c_int32),
                ("sign", C.c_double), ("R", C.c_double * 9), ("t", C.c_double * 3), ("axis", C.c_double * 3),
                ("axis_sq", C.c_double * 6), ("mimic_factor", C.c_double), ("mimic_offset", C.c_double)]


class Robot(C.Structure):
    _fields_ = [("n", C.c_int32), ("has_tip", C.c_int32), ("n_steps", C.c_int32), ("n_tips", C.c_int32),
                ("steps", Step * MAX_STEPS), ("vars", Variable * MAX_VARS), ("tip_step", C.c_int32 * MAX_TIPS),
                ("tip_has", C.c_int32 * MAX_TIPS), ("tip_R", (C.c_double * 9) * MAX_TIPS),
                ("tip_t", (C.c_double * 3) * MAX_TIPS)]


class Params(C.Structure):
    _fields_ = [("mode", C.c_int32), ("gd_max_iters", C.c_int32), ("gd_step_size", C.c_double),
                ("gd_min_cost_delta", C.c_double), ("position_threshold", C.c_double),
                ("orientation_threshold", C.c_double), ("cost_threshold", C.c_double),
                ("position_scale", C.c_double), ("rotation_scale", C.c_double),
                ("center_joints_weight", C.c_double), ("avoid_joint_limits_weight", C.c_double),
                ("minimal_displacement_weight", C.c_double), ("memetic_wipeout_fitness_tol", C.c_double),
                ("stop_optimization_on_valid_solution", C.c_int32), ("memetic_population_size", C.c_int32),
                ("memetic_elite_size", C.c_int32), ("memetic_max_generations", C.c_int32),
                ("memetic_gd_max_iters", C.c_int32), ("return_approximate_solution", C.c_int32),
                ("memetic_num_threads", C.c_int32), ("memetic_stop_on_first_solution", C.c_int32),
                ("rng_seed", C.c_uint64)]


class Problem(C.Structure):
    _fields_ = [("robot", C.POINTER(Robot)), ("params", C.POINTER(Params)), ("goal_t", (C.c_double * 3) * MAX_TIPS),
                ("goal_R", (C.c_double * 9) * MAX_TIPS), ("goal_q", (C.c_double * 4) * MAX_TIPS),
                ("seed", C.c_double * MAX_VARS)]


class Result(C.Structure):
    _fields_ = [("found", C.c_int32), ("iterations", C.c_int32), ("cost", C.c_double),
                ("evals", C.c_uint64), ("wipeouts", C.c_uint32), ("gd_steps", C.c_uint32),
                ("solution", C.c_double * MAX_VARS)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        d = C.c_double
        dp = C.POINTER(C.c_double)
        _lib.orc_atan2.restype = d
        _lib.orc_atan2.argtypes = [d, d]
        _lib.orc_sincos.argtypes = [d, dp, dp]
        _lib.orc_linear_distance.restype = d
        _lib.orc_angular_distance.restype = d
        _lib.orc_angular_distance_q.restype = d
        _lib.orc_pose_cost.restype = d
        _lib.orc_pose_cost.argtypes = [dp, dp, dp, dp, d, d]
        _lib.orc_frame_test.argtypes = [dp, dp, dp, dp, d, d]
        _lib.orc_cost.restype = d
        _lib.orc_clamp_to_limits.restype = d
        _lib.orc_clamp_to_limits.argtypes = [C.POINTER(Variable), d]
        _lib.orc_center_joints_cost.restype = d
        _lib.orc_avoid_joint_limits_cost.restype = d
        _lib.orc_minimal_displacement_cost.restype = d
        _lib.orc_random_configuration.argtypes = [C.POINTER(Robot), C.c_uint64, C.c_uint32, dp]
        _lib.orc_solve_batch.argtypes = [C.POINTER(Robot), C.POINTER(Params), C.c_int64, C.c_int64, dp, dp,
                                         C.c_int64, dp, C.POINTER(C.c_int32), dp, C.POINTER(C.c_int32),
                                         C.POINTER(C.c_uint64), C.c_int]
        _lib.orc_eval_cost_batch.argtypes = [C.POINTER(Robot), C.POINTER(Params), C.c_int64, dp, dp, C.c_int64,
                                             dp, dp, C.POINTER(C.c_int32), dp]
        _lib.orc_ik_memetic.argtypes = [C.POINTER(Problem), dp, C.c_uint32, C.POINTER(Result)]
        _lib.orc_ik_gradient.argtypes = [C.POINTER(Problem), dp, C.POINTER(Result)]
    return _lib


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def default_params(**kw) -> Params:
    p = Params()
    lib().orc_params_default(C.byref(p))
    for k, v in kw.items():
        if k == "mode" and isinstance(v, str):
            v = {"global": 0, "local": 1}[v]
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def build_robot(joint_desc: np.ndarray) -> Robot:
    r = Robot()
    jd = np.ascontiguousarray(joint_desc)
    rc = lib().orc_robot_build(jd.ctypes.data_as(C.c_void_p), C.c_int(len(jd)), C.byref(r))
    if rc != 0:
        raise ValueError(f"orc_robot_build failed: {rc}")
    return r


def build_robot_tree(joint_desc: np.ndarray, parent, tip_joint, mimic_of=None, mimic_factor=None,
                     mimic_offset=None) -> Robot:
    """Kinematic tree with several tips (and mimic joints): orc_robot_build_tree."""
    r = Robot()
    jd = np.ascontiguousarray(joint_desc)
    n = len(jd)
    par = np.ascontiguousarray(parent, dtype=np.int32)
    tips = np.ascontiguousarray(tip_joint, dtype=np.int32)
    assert par.shape == (n,)
    mo = mf = mb = None
    if mimic_of is not None:
        mo = np.ascontiguousarray(mimic_of, dtype=np.int32)
        mf = np.ascontiguousarray(mimic_factor, dtype=np.float64)
        mb = np.ascontiguousarray(mimic_offset, dtype=np.float64)
    vp = C.c_void_p
    rc = lib().orc_robot_build_tree(jd.ctypes.data_as(vp), C.c_int(n), par.ctypes.data_as(vp), tips.ctypes.data_as(vp),
                                    C.c_int(len(tips)), None if mo is None else mo.ctypes.data_as(vp),
                                    None if mf is None else mf.ctypes.data_as(vp),
                                    None if mb is None else mb.ctypes.data_as(vp), C.byref(r))
    if rc != 0:
        raise ValueError(f"orc_robot_build_tree failed: {rc}")
    return r


def sincos(x: float):
    s, c = C.c_double(), C.c_double()
    lib().orc_sincos(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def atan2(y: float, x: float) -> float:
    return lib().orc_atan2(float(y), float(x))


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def quat_to_matrix(q_wxyz) -> np.ndarray:
    q = np.asarray(q_wxyz, dtype=np.float64).copy()
    R = np.zeros(9)
    lib().orc_quat_to_matrix(_dp(q), _dp(R))
    return R.reshape(3, 3)


def matrix_to_quat(R) -> np.ndarray:
    Rm = np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(9))
    q = np.zeros(4)
    lib().orc_matrix_to_quat(_dp(Rm), _dp(q))
    return q


def frame(t, q_wxyz):
    """(t[3], R[9]) of Translation3d(t) * Quaterniond(w,x,y,z) (no normalisation)."""
    return np.asarray(t, dtype=np.float64).copy(), quat_to_matrix(q_wxyz).reshape(9).copy()


def pose_cost(goal, tip, position_scale, rotation_scale) -> float:
    return lib().orc_pose_cost(_dp(goal[0]), _dp(goal[1]), _dp(tip[0]), _dp(tip[1]),
                               float(position_scale), float(rotation_scale))


def frame_test(goal, tip, position_threshold=None, orientation_threshold=None) -> bool:
    pt = -1.0 if position_threshold is None else float(position_threshold)
    ot = -1.0 if orientation_threshold is None else float(orientation_threshold)
    return bool(lib().orc_frame_test(_dp(goal[0]), _dp(goal[1]), _dp(tip[0]), _dp(tip[1]), pt, ot))


def fk(robot: Robot, q):
    qa = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
    R, t = np.zeros(9), np.zeros(3)
    lib().orc_fk(C.byref(robot), _dp(qa), _dp(R), _dp(t))
    return R.reshape(3, 3), t


def pose_from_fk(robot: Robot, q) -> np.ndarray:
    qa = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
    pose = np.zeros(7)
    lib().orc_pose_from_fk(C.byref(robot), _dp(qa), _dp(pose))
    return pose


def poses_from_fk(robot: Robot, q) -> np.ndarray:
    """[n_tips, 7] poses of all tips."""
    qa = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
    pose = np.zeros((robot.n_tips, 7))
    lib().orc_poses_from_fk(C.byref(robot), _dp(qa), _dp(pose.reshape(-1)))
    return pose


def random_configuration(robot: Robot, gen_seed: int, problem_index: int) -> np.ndarray:
    q = np.zeros(robot.n)
    lib().orc_random_configuration(C.byref(robot), C.c_uint64(gen_seed), C.c_uint32(problem_index), _dp(q))
    return q


def make_problem(robot: Robot, params: Params, goal_pose, seed) -> Problem:
    pb = Problem()
    gp = np.ascontiguousarray(np.asarray(goal_pose, dtype=np.float64))
    sd = np.ascontiguousarray(np.asarray(seed, dtype=np.float64))
    lib().orc_problem_init(C.byref(pb), C.byref(robot), C.byref(params), _dp(gp), _dp(sd))
    pb._keep = (robot, params)
    return pb


def cost(pb: Problem, q) -> float:
    qa = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
    return lib().orc_cost(C.byref(pb), _dp(qa))


def is_solution(pb: Problem, q) -> bool:
    qa = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
    return bool(lib().orc_is_solution(C.byref(pb), _dp(qa)))


def ik_gradient(pb: Problem, initial_guess) -> Result:
    g = np.ascontiguousarray(np.asarray(initial_guess, dtype=np.float64))
    res = Result()
    lib().orc_ik_gradient(C.byref(pb), _dp(g), C.byref(res))
    return res


def ik_memetic(pb: Problem, initial_guess, problem_index: int = 0) -> Result:
    g = np.ascontiguousarray(np.asarray(initial_guess, dtype=np.float64))
    res = Result()
    lib().orc_ik_memetic(C.byref(pb), _dp(g), C.c_uint32(problem_index), C.byref(res))
    return res


def solve_batch(robot: Robot, params: Params, goal_pose: np.ndarray, seed: np.ndarray,
                first_problem_index: int = 0, n_threads: int = 0):
    """Returns dict(solution [B,n], error_code [B], cost [B], iterations [B], evals)."""
    goal_pose = np.ascontiguousarray(goal_pose, dtype=np.float64)
    B = goal_pose.shape[0]
    assert goal_pose.size == B * robot.n_tips * 7, "goal_pose must be [B, n_tips, 7]"
    n = robot.n
    seed = np.ascontiguousarray(seed, dtype=np.float64)
    stride = 0 if seed.ndim == 1 or seed.shape[0] == 1 else n
    sol = np.zeros((B, n))
    err = np.zeros(B, dtype=np.int32)
    cst = np.zeros(B)
    its = np.zeros(B, dtype=np.int32)
    ev = C.c_uint64(0)
    lib().orc_solve_batch(C.byref(robot), C.byref(params), B, first_problem_index, _dp(goal_pose.reshape(-1)),
                          _dp(seed.reshape(-1)), stride, _dp(sol.reshape(-1)), _ip(err), _dp(cst), _ip(its),
                          C.byref(ev), n_threads)
    return dict(solution=sol, error_code=err, cost=cst, iterations=its, evals=ev.value)


def solve_batch_reference_structure(robot: Robot, params: Params, goal_pose: np.ndarray, seed: np.ndarray,
                                    first_problem_index: int = 0):
    """Baseline A (timing only): one solve at a time, E gradient-descent threads per generation behind one FK mutex."""
    goal_pose = np.ascontiguousarray(goal_pose, dtype=np.float64)
    B = goal_pose.shape[0]
    n = robot.n
    seed = np.ascontiguousarray(seed, dtype=np.float64)
    stride = 0 if seed.ndim == 1 or seed.shape[0] == 1 else n
    sol = np.zeros((B, n))
    err = np.zeros(B, dtype=np.int32)
    cst = np.zeros(B)
    its = np.zeros(B, dtype=np.int32)
    lib().orc_solve_batch_reference_structure(C.byref(robot), C.byref(params), C.c_int64(B), C.c_int64(first_problem_index),
                                              _dp(goal_pose.reshape(-1)), _dp(seed.reshape(-1)), C.c_int64(stride),
                                              _dp(sol.reshape(-1)), _ip(err), _dp(cst), _ip(its))
    return dict(solution=sol, error_code=err, cost=cst, iterations=its)


def eval_cost_batch(robot: Robot, params: Params, goal_pose: np.ndarray, seed: np.ndarray, q: np.ndarray):
    goal_pose = np.ascontiguousarray(goal_pose, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    B, n = q.shape
    seed = np.ascontiguousarray(seed, dtype=np.float64)
    stride = 0 if seed.ndim == 1 or seed.shape[0] == 1 else n
    cst = np.zeros(B)
    sol = np.zeros(B, dtype=np.int32)
    assert goal_pose.size == B * robot.n_tips * 7, "goal_pose must be [B, n_tips, 7]"
    tip = np.zeros((B, 7)) if robot.n_tips == 1 else np.zeros((B, robot.n_tips, 7))
    lib().orc_eval_cost_batch(C.byref(robot), C.byref(params), B, _dp(goal_pose.reshape(-1)), _dp(seed.reshape(-1)), stride,
                              _dp(q.reshape(-1)), _dp(cst), _ip(sol), _dp(tip.reshape(-1)))
    return cst, sol, tip


def make_targets(robot: Robot, B: int, gen_seed: int = 0xC0FFEE, first: int = 0) -> np.ndarray:
    """target b = FK(q*_b), q*_b ~ U(limits) from Philox stream (gen_seed, b)."""
    if robot.n_tips == 1:
        out = np.zeros((B, 7))
        for b in range(B):
            out[b] = pose_from_fk(robot, random_configuration(robot, gen_seed, first + b))
        return out
    out = np.zeros((B, robot.n_tips, 7))
    for b in range(B):
        out[b] = poses_from_fk(robot, random_configuration(robot, gen_seed, first + b))
    return out
