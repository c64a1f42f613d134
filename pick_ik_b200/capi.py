"""ctypes binding of the C-ABI in include/pik.h (libpik_b200.so).

This is the only compute path of the package: there is no CPU fallback.  Importing works without a
GPU (so the symbol table can be checked); creating a ``Solver`` without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .robots import JOINT_DESC_DTYPE, RobotChain, RobotTree

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PIK_LIB_PATH") or os.path.join(HERE, "libpik_b200.so")  # override: A/B experiments only

PIK_OK = 0
PIK_SUCCESS = 1
PIK_NO_IK_SOLUTION = -31
MODE_GLOBAL, MODE_LOCAL = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1

# every symbol include/pik.h declares
EXPORTS = [
    "pik_version", "pik_status_string", "pik_params_default", "pik_params_validate", "pik_robot_create",
    "pik_robot_create_tree", "pik_robot_num_tips", "pik_robot_destroy", "pik_robot_num_variables", "pik_robot_get_variable",
    "pik_robot_is_valid_configuration", "pik_robot_chain_signature", "pik_random_configurations", "pik_solver_create", "pik_solver_destroy", "pik_solve_batch",
    "pik_solve_batch_async", "pik_solver_wait", "pik_solver_query", "pik_eval_cost", "pik_solver_synchronize", "pik_solver_get_stats", "pik_solver_last_error",
    "pik_device_count", "pik_host_alloc", "pik_host_free", "pik_measure_fp64_peak",
    "pik_urdf_chain", "pik_urdf_tree", "pik_srdf_group", "pik_comm_unique_id", "pik_comm_create", "pik_comm_destroy", "pik_comm_last_error", "pik_solve_batch_sharded",
    "pik_solve_batch_gather",
]
COMM_ID_BYTES = 128
URDF_NAME_BYTES = 64


class PikError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        msg = f"{where}: {status_string(status)} ({status})"
        if detail:
            msg += f": {detail}"
        super().__init__(msg)


class Params(C.Structure):
    """pik_params: src/pick_ik_parameters.yaml names and defaults (+ rng_seed)."""

    _fields_ = [
        ("mode", C.c_int32), ("gd_max_iters", C.c_int32), ("gd_step_size", C.c_double),
        ("gd_min_cost_delta", C.c_double), ("position_threshold", C.c_double),
        ("orientation_threshold", C.c_double), ("approximate_solution_position_threshold", C.c_double),
        ("approximate_solution_orientation_threshold", C.c_double),
        ("approximate_solution_joint_threshold", C.c_double),
        ("approximate_solution_cost_threshold", C.c_double), ("cost_threshold", C.c_double),
        ("position_scale", C.c_double), ("rotation_scale", C.c_double), ("center_joints_weight", C.c_double),
        ("avoid_joint_limits_weight", C.c_double), ("minimal_displacement_weight", C.c_double),
        ("memetic_wipeout_fitness_tol", C.c_double), ("memetic_gd_max_time", C.c_double),
        ("stop_optimization_on_valid_solution", C.c_int32), ("memetic_num_threads", C.c_int32),
        ("memetic_stop_on_first_solution", C.c_int32), ("memetic_population_size", C.c_int32),
        ("memetic_elite_size", C.c_int32), ("memetic_max_generations", C.c_int32),
        ("memetic_gd_max_iters", C.c_int32), ("return_approximate_solution", C.c_int32),
        ("rng_seed", C.c_uint64),
    ]


class Variable(C.Structure):
    _fields_ = [("min", C.c_double), ("max", C.c_double), ("mid", C.c_double), ("half_span", C.c_double),
                ("max_velocity_rcp", C.c_double), ("minimal_displacement_factor", C.c_double),
                ("bounded", C.c_int32), ("pad_", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("problems", C.c_int64), ("solved", C.c_int64), ("generation_launches", C.c_int64),
                ("kernel_launches", C.c_int64), ("problem_generations", C.c_int64), ("gd_steps", C.c_int64),
                ("device_ms", C.c_double), ("generation_ms", C.c_double)]


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Loads libpik_b200.so; raises if it has not been built (python -m pick_ik_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m pick_ik_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses: host or device
    L.pik_version.restype = C.c_int
    L.pik_status_string.restype = C.c_char_p
    L.pik_status_string.argtypes = [C.c_int]
    L.pik_params_default.restype = None
    L.pik_params_default.argtypes = [C.POINTER(Params)]
    L.pik_params_validate.argtypes = [C.POINTER(Params)]
    L.pik_robot_create.argtypes = [vp, C.c_int32, C.POINTER(vp)]
    L.pik_robot_create_tree.argtypes = [vp, C.c_int32, vp, vp, C.c_int32, vp, vp, vp, C.POINTER(vp)]
    L.pik_robot_num_tips.restype = C.c_int32
    L.pik_robot_num_tips.argtypes = [vp]
    L.pik_robot_destroy.restype = None
    L.pik_robot_destroy.argtypes = [vp]
    L.pik_robot_num_variables.restype = C.c_int32
    L.pik_robot_num_variables.argtypes = [vp]
    L.pik_robot_get_variable.argtypes = [vp, C.c_int32, C.POINTER(Variable)]
    L.pik_robot_is_valid_configuration.argtypes = [vp, dp]
    L.pik_robot_chain_signature.restype = C.c_char_p
    L.pik_robot_chain_signature.argtypes = [vp]
    L.pik_random_configurations.argtypes = [vp, C.c_uint64, C.c_int64, C.c_int64, dp]
    L.pik_solver_create.argtypes = [vp, C.c_int32, vp, C.POINTER(vp)]
    L.pik_solver_destroy.restype = None
    L.pik_solver_destroy.argtypes = [vp]
    L.pik_solve_batch.argtypes = [vp, C.POINTER(Params), C.c_int64, C.c_int64, dp, dp, C.c_int64, dp, ip, dp, ip,
                                  C.c_int32]
    L.pik_solve_batch_async.argtypes = L.pik_solve_batch.argtypes
    L.pik_solver_wait.argtypes = [vp]
    L.pik_solver_query.argtypes = [vp]
    L.pik_eval_cost.argtypes = [vp, C.POINTER(Params), C.c_int64, dp, dp, C.c_int64, dp, dp, ip, dp, C.c_int32]
    L.pik_solver_synchronize.argtypes = [vp]
    L.pik_solver_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.pik_solver_last_error.restype = C.c_char_p
    L.pik_solver_last_error.argtypes = [vp]
    L.pik_device_count.restype = C.c_int
    L.pik_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.pik_host_free.restype = None
    L.pik_host_free.argtypes = [vp]
    L.pik_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.pik_urdf_chain.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, vp, C.c_int32, C.POINTER(C.c_int32), vp, vp]
    L.pik_urdf_tree.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int32, vp, C.c_int32, C.POINTER(C.c_int32),
                                vp, vp, vp, vp, vp, vp, vp]
    L.pik_srdf_group.argtypes = [C.c_char_p, C.c_char_p, vp, vp, C.c_int32, C.POINTER(C.c_int32)]
    L.pik_comm_unique_id.argtypes = [vp]
    L.pik_comm_create.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.pik_comm_destroy.restype = None
    L.pik_comm_destroy.argtypes = [vp]
    L.pik_comm_last_error.restype = C.c_char_p
    L.pik_solve_batch_sharded.argtypes = [vp, vp, C.POINTER(Params), C.c_int64, C.c_int64, dp, dp, C.c_int64, dp,
                                          C.c_int32]
    L.pik_solve_batch_gather.argtypes = [vp, vp, C.POINTER(Params), C.c_int64, C.c_int64, dp, dp, C.c_int64, vp,
                                         C.c_int32, dp, C.c_int32]
    _lib = L
    return L


def status_string(status: int) -> str:
    return lib().pik_status_string(int(status)).decode()


def default_params(**kw) -> Params:
    p = Params()
    lib().pik_params_default(C.byref(p))
    for k, v in kw.items():
        if k == "mode" and isinstance(v, str):
            if v not in ("global", "local"):
                raise ValueError("mode must be one of ['global', 'local']")
            v = {"global": MODE_GLOBAL, "local": MODE_LOCAL}[v]
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def validate_params(p: Params) -> int:
    return lib().pik_params_validate(C.byref(p))


def device_count() -> int:
    return lib().pik_device_count()


class Robot:
    """pik_robot: the flattened chain + Robot::Variable table (src/robot.cpp:44-85)."""

    def __init__(self, chain_or_desc):
        h = C.c_void_p()
        if isinstance(chain_or_desc, RobotTree):
            desc, parent, tips, mimic_of, factor, offset = chain_or_desc.tree_arrays()
            desc = np.ascontiguousarray(desc, dtype=JOINT_DESC_DTYPE)
            vp = C.c_void_p
            rc = lib().pik_robot_create_tree(desc.ctypes.data_as(vp), len(desc), parent.ctypes.data_as(vp),
                                             tips.ctypes.data_as(vp), len(tips), mimic_of.ctypes.data_as(vp),
                                             factor.ctypes.data_as(vp), offset.ctypes.data_as(vp), C.byref(h))
            where = "pik_robot_create_tree"
        else:
            desc = chain_or_desc.joint_desc() if isinstance(chain_or_desc, RobotChain) else chain_or_desc
            desc = np.ascontiguousarray(desc, dtype=JOINT_DESC_DTYPE)
            rc = lib().pik_robot_create(desc.ctypes.data_as(C.c_void_p), len(desc), C.byref(h))
            where = "pik_robot_create"
        self.desc = desc
        if rc != PIK_OK:
            raise PikError(rc, where)
        self.handle = h
        self.n = lib().pik_robot_num_variables(h)
        self.n_tips = lib().pik_robot_num_tips(h)

    def variable(self, i: int) -> Variable:
        v = Variable()
        rc = lib().pik_robot_get_variable(self.handle, i, C.byref(v))
        if rc != PIK_OK:
            raise PikError(rc, "pik_robot_get_variable")
        return v

    def chain_signature(self) -> str:
        return lib().pik_robot_chain_signature(self.handle).decode()

    def is_valid_configuration(self, q) -> bool:
        qa = np.ascontiguousarray(q, dtype=np.float64)
        assert qa.shape == (self.n,)
        return bool(lib().pik_robot_is_valid_configuration(self.handle, qa.ctypes.data_as(C.c_void_p)))

    def random_configurations(self, B: int, gen_seed: int = 0xC0FFEE, first_problem_index: int = 0) -> np.ndarray:
        """pik_random_configurations: q [B, n] from the Philox stream (gen_seed, first_problem_index + b)."""
        q = np.empty((B, self.n))
        rc = lib().pik_random_configurations(self.handle, gen_seed, first_problem_index, B, q.ctypes.data)
        if rc != PIK_OK:
            raise PikError(rc, "pik_random_configurations")
        return q

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None:
            _lib.pik_robot_destroy(h)


def _addr(a) -> Optional[int]:
    return None if a is None else a.ctypes.data


class Solver:
    """pik_solver on one CUDA device."""

    def __init__(self, robot: Robot, device: int = 0, stream: int = 0):
        self.robot = robot
        self.n = robot.n
        self.n_tips = robot.n_tips
        h = C.c_void_p()
        rc = lib().pik_solver_create(robot.handle, device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != PIK_OK:
            raise PikError(rc, "pik_solver_create")
        self.handle = h

    def _check(self, rc: int, where: str):
        if rc != PIK_OK:
            raise PikError(rc, where, lib().pik_solver_last_error(self.handle).decode())

    # -- host-memory API (numpy) --------------------------------------------------------------
    def solve_batch(self, params: Params, goal_pose: np.ndarray, seed: np.ndarray, first_problem_index: int = 0,
                    out: Optional[dict] = None) -> dict:
        goal_pose = np.ascontiguousarray(goal_pose, dtype=np.float64)
        if goal_pose.shape[1:] not in ((self.n_tips, 7),) + (((7,),) if self.n_tips == 1 else ()):
            raise ValueError("goal_pose must be [B, n_tips, 7] ([B, 7] for one tip)")
        B = goal_pose.shape[0]
        seed = np.ascontiguousarray(seed, dtype=np.float64)
        if seed.shape == (self.n,) or seed.shape == (1, self.n):
            stride = 0
        elif seed.shape == (B, self.n):
            stride = self.n
        else:
            raise ValueError("seed must be [n] or [B, n]")
        if out is None:
            out = dict(solution=np.empty((B, self.n)), error_code=np.empty(B, dtype=np.int32), cost=np.empty(B),
                       iterations=np.empty(B, dtype=np.int32))
        rc = lib().pik_solve_batch(self.handle, C.byref(params), B, first_problem_index, _addr(goal_pose),
                                   _addr(seed), stride, _addr(out["solution"]), _addr(out["error_code"]),
                                   _addr(out["cost"]), _addr(out["iterations"]), MEM_HOST)
        self._check(rc, "pik_solve_batch")
        return out

    def eval_cost(self, params: Params, goal_pose: np.ndarray, seed: np.ndarray, q: np.ndarray):
        goal_pose = np.ascontiguousarray(goal_pose, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        B = q.shape[0]
        assert q.shape == (B, self.n) and goal_pose.size == B * self.n_tips * 7
        seed = np.ascontiguousarray(seed, dtype=np.float64)
        stride = 0 if seed.shape in ((self.n,), (1, self.n)) else self.n
        cost = np.empty(B)
        sol = np.empty(B, dtype=np.int32)
        tip = np.empty((B, 7)) if self.n_tips == 1 else np.empty((B, self.n_tips, 7))
        rc = lib().pik_eval_cost(self.handle, C.byref(params), B, _addr(goal_pose), _addr(seed), stride, _addr(q),
                                 _addr(cost), _addr(sol), _addr(tip), MEM_HOST)
        self._check(rc, "pik_eval_cost")
        return cost, sol, tip

    # -- raw-pointer API (device or pinned host addresses) ---------------------------------------
    def solve_batch_ptr(self, params: Params, B: int, first_problem_index: int, goal_pose: int, seed: int,
                        seed_stride: int, solution: int, error_code: int, cost: int, iterations: int,
                        memory: int = MEM_DEVICE):
        rc = lib().pik_solve_batch(self.handle, C.byref(params), B, first_problem_index, goal_pose, seed, seed_stride,
                                   solution, error_code, cost or None, iterations or None, memory)
        self._check(rc, "pik_solve_batch")

    def solve_batch_async_ptr(self, params: Params, B: int, first_problem_index: int, goal_pose: int, seed: int,
                              seed_stride: int, solution: int, error_code: int, cost: int, iterations: int,
                              memory: int = MEM_DEVICE):
        """pik_solve_batch_async: enqueue the whole solve on the solver's stream and return; wait() completes it."""
        rc = lib().pik_solve_batch_async(self.handle, C.byref(params), B, first_problem_index, goal_pose, seed,
                                         seed_stride, solution, error_code, cost or None, iterations or None, memory)
        self._check(rc, "pik_solve_batch_async")

    def wait(self):
        self._check(lib().pik_solver_wait(self.handle), "pik_solver_wait")

    def query(self) -> bool:
        rc = lib().pik_solver_query(self.handle)
        if rc < 0:
            self._check(rc, "pik_solver_query")
        return rc == 1

    def solve_batch_sharded_ptr(self, comm: "Comm", params: Params, B_local: int, first_problem_index: int,
                                goal_pose: int, seed: int, seed_stride: int, gathered: int, memory: int = MEM_DEVICE):
        """This rank's shard + the NCCL all-gather of the packed results: gathered [n_ranks][B_local][n + 3]."""
        rc = lib().pik_solve_batch_sharded(self.handle, comm.handle, C.byref(params), B_local, first_problem_index,
                                           goal_pose, seed, seed_stride, gathered, memory)
        if rc != PIK_OK:
            raise PikError(rc, "pik_solve_batch_sharded",
                           lib().pik_solver_last_error(self.handle).decode() or lib().pik_comm_last_error().decode())

    def solve_batch_gather_ptr(self, comm: "Comm", params: Params, B_local: int, first_problem_index: int,
                               goal_pose: int, seed: int, seed_stride: int, gathered: int, memory: int = MEM_DEVICE,
                               counts=None, root: int = -1):
        """pik_solve_batch_gather: uneven shards (counts [n_ranks]) and / or a single receiving rank (root)."""
        c = None
        if counts is not None:
            c = np.ascontiguousarray(counts, dtype=np.int64)
            assert c.shape == (comm.n_ranks,)
        rc = lib().pik_solve_batch_gather(self.handle, comm.handle, C.byref(params), B_local, first_problem_index,
                                          goal_pose, seed, seed_stride, None if c is None else c.ctypes.data, root,
                                          gathered or None, memory)
        if rc != PIK_OK:
            raise PikError(rc, "pik_solve_batch_gather",
                           lib().pik_solver_last_error(self.handle).decode() or lib().pik_comm_last_error().decode())

    def synchronize(self):
        self._check(lib().pik_solver_synchronize(self.handle), "pik_solver_synchronize")

    def stats(self) -> Stats:
        s = Stats()
        self._check(lib().pik_solver_get_stats(self.handle, C.byref(s)), "pik_solver_get_stats")
        return s

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        self._check(lib().pik_measure_fp64_peak(self.handle, C.byref(v)), "pik_measure_fp64_peak")
        return v.value

    def close(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None:
            _lib.pik_solver_destroy(h)

    def __del__(self):
        self.close()


def urdf_chain(urdf_xml: str, base_link: str, tip_link: str):
    """pik_urdf_chain: (joint_desc array in root-to-tip order, joint names) of the chain base_link -> tip_link."""
    n = C.c_int32()
    xml, base, tip = urdf_xml.encode(), base_link.encode(), tip_link.encode()
    rc = lib().pik_urdf_chain(xml, base, tip, None, 0, C.byref(n), None, None)
    if rc != PIK_OK:
        raise PikError(rc, "pik_urdf_chain")
    desc = np.zeros(n.value, dtype=JOINT_DESC_DTYPE)
    names = C.create_string_buffer(max(1, n.value) * URDF_NAME_BYTES)
    rc = lib().pik_urdf_chain(xml, base, tip, desc.ctypes.data_as(C.c_void_p), n.value, C.byref(n), names, None)
    if rc != PIK_OK:
        raise PikError(rc, "pik_urdf_chain")
    raw = names.raw
    return desc, [raw[k * URDF_NAME_BYTES:(k + 1) * URDF_NAME_BYTES].split(b"\0")[0].decode() for k in range(n.value)]


def _names(buf, n):
    raw = buf.raw
    return [raw[k * URDF_NAME_BYTES:(k + 1) * URDF_NAME_BYTES].split(b"\0")[0].decode() for k in range(n)]


def urdf_tree(urdf_xml: str, base_link: str, tip_links):
    """pik_urdf_tree: dict(desc, parent, tip_joint, mimic_of, mimic_factor, mimic_offset, joint_names, link_names) of the
    joints between base_link and the tip links, parents first (the arguments of pik_robot_create_tree)."""
    n = C.c_int32()
    xml, base = urdf_xml.encode(), base_link.encode()
    tips = (C.c_char_p * len(tip_links))(*[t.encode() for t in tip_links])
    rc = lib().pik_urdf_tree(xml, base, tips, len(tip_links), None, 0, C.byref(n), None, None, None, None, None, None, None)
    if rc != PIK_OK:
        raise PikError(rc, "pik_urdf_tree")
    k = n.value
    desc = np.zeros(k, dtype=JOINT_DESC_DTYPE)
    parent, mimic_of = np.zeros(k, dtype=np.int32), np.zeros(k, dtype=np.int32)
    tip_joint = np.zeros(len(tip_links), dtype=np.int32)
    factor, offset = np.zeros(k), np.zeros(k)
    jn, ln = C.create_string_buffer(max(1, k) * URDF_NAME_BYTES), C.create_string_buffer(max(1, k) * URDF_NAME_BYTES)
    vp = C.c_void_p
    rc = lib().pik_urdf_tree(xml, base, tips, len(tip_links), desc.ctypes.data_as(vp), k, C.byref(n), parent.ctypes.data_as(vp),
                             tip_joint.ctypes.data_as(vp), mimic_of.ctypes.data_as(vp), factor.ctypes.data_as(vp),
                             offset.ctypes.data_as(vp), jn, ln)
    if rc != PIK_OK:
        raise PikError(rc, "pik_urdf_tree")
    return dict(desc=desc, parent=parent, tip_joint=tip_joint, mimic_of=mimic_of, mimic_factor=factor, mimic_offset=offset,
                joint_names=_names(jn, k), link_names=_names(ln, k))


def srdf_group(srdf_xml: str, group: str):
    """pik_srdf_group: (base_link, [tip_links]) of a planning group defined by <chain> entries."""
    n = C.c_int32()
    base = C.create_string_buffer(URDF_NAME_BYTES)
    rc = lib().pik_srdf_group(srdf_xml.encode(), group.encode(), base, None, 0, C.byref(n))
    if rc != PIK_OK:
        raise PikError(rc, "pik_srdf_group")
    tips = C.create_string_buffer(max(1, n.value) * URDF_NAME_BYTES)
    rc = lib().pik_srdf_group(srdf_xml.encode(), group.encode(), base, tips, n.value, C.byref(n))
    if rc != PIK_OK:
        raise PikError(rc, "pik_srdf_group")
    return base.raw.split(b"\0")[0].decode(), _names(tips, n.value)


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0); hand the bytes to every rank."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = lib().pik_comm_unique_id(buf)
    if rc != PIK_OK:
        raise PikError(rc, "pik_comm_unique_id", lib().pik_comm_last_error().decode())
    return buf.raw


class Comm:
    """pik_comm: the library's own NCCL communicator for the sharded solve (one rank per GPU)."""

    def __init__(self, unique_id: bytes, n_ranks: int, rank: int, device: int):
        if len(unique_id) != COMM_ID_BYTES:
            raise ValueError("unique_id must be COMM_ID_BYTES long")
        h = C.c_void_p()
        rc = lib().pik_comm_create(C.create_string_buffer(unique_id, COMM_ID_BYTES), n_ranks, rank, device, C.byref(h))
        if rc != PIK_OK:
            raise PikError(rc, "pik_comm_create", lib().pik_comm_last_error().decode())
        self.handle = h
        self.n_ranks, self.rank = n_ranks, rank

    def close(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None:
            _lib.pik_comm_destroy(h)

    def __del__(self):
        self.close()
