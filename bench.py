#!/usr/bin/env python
"""bench.py -- IK solves/sec on BASELINE.json's headline workload.

A step = one pass of the hot path over one batch: 65 536 random reachable Franka Panda poses per GPU, memetic
solver (population 128, 4 elites, 25 GD iterations per elite, <= 100 generations), seed state = Panda home
(BASELINE.json configs[1]).  Targets: FK of configurations drawn from the Philox4x32-10 stream (0xC0FFEE, global
problem index) -- pik_random_configurations in the C-ABI, orc_random_configuration in the oracle: the same bits.

  value      device-resident inputs -> device-resident outputs, CUDA events on the solver's stream
  e2e        the same call through the C-ABI with pinned HOST buffers (H2D and D2H inside the timed region)
  pipelined  (extra) two solvers on two streams, one batch in flight on each: the throughput-bound head of one
             batch runs in the shadow of the latency-bound tail of the other (pik_solve_batch_async)
  N > 1      (torchrun) one process per GPU; weak scaling by default (each rank its own 65 536 poses), or
             --scaling strong (65 536 poses in total); RNG keyed by the global problem index; one NCCL exchange
             of the packed solutions inside the timed region (pik_solve_batch_sharded: the library's own
             communicator, torch.distributed only carries the unique id and the barriers), max over ranks.
             After the timed region every rank checks problems of OTHER ranks' shards, taken from the gathered
             block, bit for bit against the CPU oracle.

`--impl reference` times the reference's CPU algorithm (oracle/pik_oracle.c, a restatement: the reference
itself needs ROS 2 / MoveIt / Eigen and cannot be built here) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "IK solves/sec (Panda 7-DoF, memetic pop=128) at 1/2/4/8 B200 vs ref CPU"
UNIT = "solves/s"
WORKLOAD = dict(robot="panda", poses_per_gpu=65536, population=128, elites=4, gd_iters=25, max_generations=100)
TARGET_SEED = 0xC0FFEE
# FP64 operations per cost evaluation, SURVEY.md 8(d): 63 per constant origin + 18 per axis-aligned revolute
# joint (21 prismatic) + 60 pose cost + ~45 per sincos + ~50 atan2 (+ ~6 n for the joint-limit goals)
FLOP_PER_EVAL = {"panda": 1050.0, "ur5": 930.0, "fetch": 1190.0}


def solver_kwargs():
    return dict(mode="global", memetic_population_size=WORKLOAD["population"],
                memetic_elite_size=WORKLOAD["elites"], memetic_gd_max_iters=WORKLOAD["gd_iters"],
                memetic_max_generations=WORKLOAD["max_generations"])


def config_dict(n_gpus: int, B: int, scaling: str):
    return {
        "workload": "Panda 7-DoF, %d random reachable poses per GPU, memetic pop=128 elite=4 gd_iters=25, "
                    "<=100 generations, seed=home (BASELINE.json configs[1]); targets = FK of Philox(0xC0FFEE, b) "
                    "configurations" % B,
        "robot": "panda", "poses_per_gpu": B, "global_poses": B * n_gpus, "population": WORKLOAD["population"],
        "elites": WORKLOAD["elites"], "max_generations": WORKLOAD["max_generations"],
        "parallelism": "pose-batch sharded x%d (%s scaling)" % (n_gpus, scaling),
        "l2": "inputs larger than L2: population state 2 x %.2f GB per GPU streams through HBM every generation"
              % (B * 16 * 128 * 8 / 1e9),
    }


# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and clock-event reasons with NVML during the timed region."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, device_index: int):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.handle = None
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                import torch

                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.handle = None

    def run(self):
        if self.handle is None:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's algorithm (test infrastructure, executed here only)
# ---------------------------------------------------------------------------------------------------
def cpu_oracle_run(B: int, threads: int, first: int = 0):
    """Times the CPU restatement (oracle) on poses first .. first + B of the workload, one pose per thread
    (baseline B of SURVEY.md 8d); returns (solves/s, dict)."""
    from oracle import orc
    from pick_ik_b200 import robots

    chain = robots.panda()
    orobot = orc.build_robot(chain.joint_desc())
    op = orc.default_params(**solver_kwargs())
    goal = orc.make_targets(orobot, B, TARGET_SEED, first)
    home = np.array(robots.PANDA_HOME)
    t0 = time.perf_counter()
    res = orc.solve_batch(orobot, op, goal, home, first_problem_index=first, n_threads=threads)
    dt = time.perf_counter() - t0
    return B / dt, dict(seconds=dt, solved=int((res["error_code"] == 1).sum()), evals=int(res["evals"]))


def cpu_reference_structure_run(B: int):
    """Baseline A of SURVEY.md 8(d): the reference's own structure -- one solve at a time, the elites' gradient
    descents of every generation in E threads that share one FK mutex (src/ik_memetic.cpp:230-239,
    src/fk_moveit.cpp:21)."""
    from oracle import orc
    from pick_ik_b200 import robots

    if not hasattr(orc, "solve_batch_reference_structure"):
        return None
    chain = robots.panda()
    orobot = orc.build_robot(chain.joint_desc())
    op = orc.default_params(**solver_kwargs())
    goal = orc.make_targets(orobot, B, TARGET_SEED, 0)
    home = np.array(robots.PANDA_HOME)
    t0 = time.perf_counter()
    res = orc.solve_batch_reference_structure(orobot, op, goal, home)
    dt = time.perf_counter() - t0
    return {"value": B / dt, "unit": UNIT, "threads": WORKLOAD["elites"], "sample": "first %d poses, %.1f s" % (B, dt),
            "solved": int((res["error_code"] == 1).sum()),
            "structure": "one solve at a time; per generation E=4 gradient-descent threads behind one FK mutex"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    sample = args.ref_sample or (WORKLOAD["poses_per_gpu"] if args.steps <= 6 else 16384)
    for _ in range(args.warmup):
        cpu_oracle_run(min(sample, 256), cores)
    times, solved = [], 0
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        v, info = cpu_oracle_run(sample, cores)
        times.append(info["seconds"])
        solved = info["solved"]
    total = time.perf_counter() - t_all0
    value = sample * args.steps / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus, WORKLOAD["poses_per_gpu"], args.scaling),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "each step solves the first %d poses of the workload (of %d) with %d host threads, "
                                   "one pose per thread; oracle/pik_oracle.c restates the reference's CPU algorithm "
                                   "(the reference needs ROS 2/MoveIt/Eigen and does not build here)"
                                   % (sample, WORKLOAD["poses_per_gpu"], cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "solved_frac": solved / sample, "wall_s": total,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes of generation-kernel launches from the committed ncu --set full captures, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def make_goals(robot, solver, params, B: int, first: int, seed_state):
    """Targets of problems first .. first + B: FK (by the engine) of the Philox configurations."""
    q = robot.random_configurations(B, TARGET_SEED, first)
    ident = np.zeros((B, 7))
    ident[:, 3] = 1.0
    _, _, goal = solver.eval_cost(params, ident, seed_state, q)
    return goal


def oracle_spot_check(robot, solver, params, home, rows, global_indices):
    """rows [k][n + 3] (joints, cost, error_code, iterations) of problems global_indices against the CPU oracle,
    bit for bit.  Returns the number of mismatching problems."""
    from oracle import orc
    from pick_ik_b200 import robots

    orobot = orc.build_robot(robots.panda().joint_desc())
    op = orc.default_params(**solver_kwargs())
    n = robot.n
    bad = 0
    for row, g in zip(rows, global_indices):
        goal = make_goals(robot, solver, params, 1, int(g), home)
        ref = orc.solve_batch(orobot, op, goal, home, first_problem_index=int(g), n_threads=1)
        same = (np.array_equal(ref["solution"][0], row[:n]) and ref["cost"][0] == row[n]
                and ref["error_code"][0] == int(row[n + 1]) and ref["iterations"][0] == int(row[n + 2]))
        bad += 0 if same else 1
    return bad


def other_config_lines(torch, capi, robots, dev, local_rank, peak_hbm, fp64_peak, steps):
    """BASELINE.json configs[2] (UR5 262 144 poses, GD-only) and configs[3] (Fetch arm+torso, pop 256, joint costs,
    65 536 poses) on this GPU, device-resident, each with its own roofline fractions."""
    out = []
    stream = torch.cuda.current_stream()

    def run(label, name, kw, B, goal_np, seed_np, per_problem_seed):
        robot = capi.Robot(robots.ROBOTS[name]())
        n = robot.n
        solver = capi.Solver(robot, device=local_rank, stream=stream.cuda_stream)
        params = capi.default_params(**kw)
        d_goal = torch.from_numpy(goal_np(robot, solver, params)).to(dev)
        seed_arr = seed_np(robot)
        d_seed = torch.from_numpy(seed_arr).to(dev)
        d_sol = torch.empty((B, n), dtype=torch.float64, device=dev)
        d_err = torch.empty(B, dtype=torch.int32, device=dev)
        d_cost = torch.empty(B, dtype=torch.float64, device=dev)
        d_its = torch.empty(B, dtype=torch.int32, device=dev)
        ms, gd_steps, prob_gens = [], 0, 0
        for it in range(steps + 1):
            solver.solve_batch_ptr(params, B, 0, d_goal.data_ptr(), d_seed.data_ptr(), n if per_problem_seed else 0,
                                   d_sol.data_ptr(), d_err.data_ptr(), d_cost.data_ptr(), d_its.data_ptr(), capi.MEM_DEVICE)
            st = solver.stats()
            if it > 0:
                ms.append(st.device_ms)
                gd_steps, prob_gens = st.gd_steps, st.problem_generations
        t = float(np.mean(ms))
        P, E = kw.get("memetic_population_size", 16), kw.get("memetic_elite_size", 4)
        if kw["mode"] == "global":
            evals = gd_steps * (2 * n + 3) + prob_gens * (2 * E + (P - E) + 1)
            algo_bytes = 2 * P * (2 * n + 2) * 8 * prob_gens
        else:
            evals = gd_steps * (2 * n + 3) + B
            algo_bytes = B * (2 * n + 10) * 8
        tf = evals * FLOP_PER_EVAL[name] / (t * 1e-3) / 1e12
        out.append({"config": label, "poses": B, "ms": t, "solves_per_s": B / (t * 1e-3),
                    "solved_frac": float((d_err == 1).sum().item()) / B, "chain_signature": robot.chain_signature(),
                    "fp64_tflops": tf, "fp64_frac": tf / fp64_peak if fp64_peak else None,
                    "hbm_gbs": algo_bytes / (t * 1e-3) / 1e9, "hbm_frac": algo_bytes / (t * 1e-3) / 1e9 / peak_hbm})
        solver.close()

    # configs[2]: per-problem seed q0 ~ U(limits), target = FK(q0 + delta), delta ~ U(-0.1, 0.1)^n (as tests/ik_tests.cpp:272-273)
    B2 = 262144
    ur_state = {}

    def ur_seed(robot):
        if "q0" not in ur_state:
            ur_state["q0"] = robot.random_configurations(B2, TARGET_SEED + 1, 0)
        return ur_state["q0"]

    def ur_goal(robot, solver, params):
        q0 = ur_seed(robot)
        lo = np.array([robot.variable(i).min for i in range(robot.n)])
        hi = np.array([robot.variable(i).max for i in range(robot.n)])
        u = (robot.random_configurations(B2, TARGET_SEED + 2, 0) - lo) / (hi - lo)
        ident = np.zeros((B2, 7)); ident[:, 3] = 1.0
        return solver.eval_cost(params, ident, q0, q0 + (0.2 * u - 0.1))[2]

    run("configs[2] UR5 6-DoF, 262144 poses, GD-only", "ur5", dict(mode="local"), B2, ur_goal, ur_seed, True)

    # configs[3]: seed = mid-range (unbounded: 0), tests/ik_memetic_tests.cpp:186-206 goal weights
    B3 = 65536

    def fetch_seed(robot):
        return np.array([robot.variable(i).mid if robot.variable(i).bounded else 0.0 for i in range(robot.n)])

    def fetch_goal(robot, solver, params):
        return make_goals(robot, solver, params, B3, 0, fetch_seed(robot))

    run("configs[3] Fetch 8-DoF arm+torso, 65536 poses, memetic pop=256, joint-centering + avoid-limits", "fetch",
        dict(mode="global", memetic_population_size=256, center_joints_weight=0.01, avoid_joint_limits_weight=0.01,
             cost_threshold=0.01, position_threshold=0.01), B3, fetch_goal, fetch_seed, False)
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    from pick_ik_b200 import capi, robots

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    strong = args.scaling == "strong"
    B = args.batch // world if strong else args.batch
    chain = robots.panda()
    n = chain.num_variables
    robot = capi.Robot(chain)
    stream = torch.cuda.current_stream()
    solver = capi.Solver(robot, device=local_rank, stream=stream.cuda_stream)
    params = capi.default_params(**solver_kwargs())
    home = np.array(robots.PANDA_HOME)

    first = rank * B
    goal_np = make_goals(robot, solver, params, B, first, home)
    d_goal = torch.from_numpy(goal_np).to(dev)
    d_seed = torch.from_numpy(home).to(dev)
    d_sol = torch.empty((B, n), dtype=torch.float64, device=dev)
    d_err = torch.empty(B, dtype=torch.int32, device=dev)
    d_cost = torch.empty(B, dtype=torch.float64, device=dev)
    d_its = torch.empty(B, dtype=torch.int32, device=dev)
    gathered = torch.empty((world, B, n + 3), dtype=torch.float64, device=dev) if world > 1 else None
    comm = None
    if world > 1:
        # the library's own NCCL communicator: rank 0 draws the id, torch.distributed only carries the bytes
        ids = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = capi.Comm(ids[0], world, rank, local_rank)

    def step_device():
        if world == 1:
            solver.solve_batch_ptr(params, B, first, d_goal.data_ptr(), d_seed.data_ptr(), 0, d_sol.data_ptr(),
                                   d_err.data_ptr(), d_cost.data_ptr(), d_its.data_ptr(), capi.MEM_DEVICE)
        else:
            # shard solve + NCCL exchange of the packed results on the solver's stream (pik_solve_batch_sharded)
            solver.solve_batch_sharded_ptr(comm, params, B, first, d_goal.data_ptr(), d_seed.data_ptr(), 0,
                                           gathered.data_ptr(), capi.MEM_DEVICE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    gen_ms = 0.0
    prob_gens = 0
    gd_steps = 0
    gen_launches = 0
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
        st = solver.stats()
        launches += st.kernel_launches
        gen_ms += st.generation_ms
        prob_gens += st.problem_generations
        gd_steps += st.gd_steps
        gen_launches += st.generation_launches
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    if world == 1:
        solved = int((d_err == 1).sum().item())
    else:
        solved = int((gathered[rank, :, n + 1] == 1).sum().item())

    # ---- end to end through the C-ABI with pinned host buffers
    h_goal = torch.from_numpy(goal_np).pin_memory()
    h_seed = torch.from_numpy(home.copy()).pin_memory()
    h_out = dict(solution=torch.empty((B, n), dtype=torch.float64).pin_memory().numpy(),
                 error_code=torch.empty(B, dtype=torch.int32).pin_memory().numpy(),
                 cost=torch.empty(B, dtype=torch.float64).pin_memory().numpy(),
                 iterations=torch.empty(B, dtype=torch.int32).pin_memory().numpy())
    h2d = h_goal.numel() * 8 + h_seed.numel() * 8
    e2e_steps = args.steps
    if world == 1:
        d2h = sum(v.nbytes for v in h_out.values())

        def step_e2e():
            solver.solve_batch(params, h_goal.numpy(), h_seed.numpy(), first, out=h_out)
    else:
        # host goal poses in; the packed results of every rank gathered on rank 0 and copied to its (pinned) host
        # block: the one device-to-host copy of the job
        h_gathered = torch.empty((world, B, n + 3) if rank == 0 else (1,), dtype=torch.float64).pin_memory()
        d2h = world * B * (n + 3) * 8

        def step_e2e():
            solver.solve_batch_gather_ptr(comm, params, B, first, h_goal.data_ptr(), h_seed.data_ptr(), 0,
                                          h_gathered.data_ptr() if rank == 0 else 0, capi.MEM_HOST, root=0)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * B * e2e_steps / e2e_s
    if world == 1:
        assert int((h_out["error_code"] == 1).sum()) == solved, "e2e and device-resident runs disagree"
    elif rank == 0:
        # rank 0 holds every shard: the same bits as the device-resident all-gather of the timed region
        assert np.array_equal(gathered.cpu().numpy(), h_gathered.numpy()), "device- and host-memory gathers disagree"

    # ---- parity under the bench: gathered results (other ranks' shards when N > 1) against the CPU oracle
    n_check = args.spot_check
    rng = np.random.default_rng(1000 + rank)
    if world == 1:
        picks = rng.choice(B, min(n_check, B), replace=False)
        rows = np.concatenate([h_out["solution"][picks], h_out["cost"][picks, None],
                               h_out["error_code"][picks, None].astype(np.float64),
                               h_out["iterations"][picks, None].astype(np.float64)], axis=1)
        global_idx = first + picks
    else:
        others = np.array([r for r in range(world) if r != rank])
        pr = rng.choice(others, n_check)
        pb = rng.integers(0, B, n_check)
        host_g = gathered.cpu().numpy()  # the all-gather of the timed region: every rank holds every shard
        rows = host_g[pr, pb]
        global_idx = pr * B + pb
    mismatches = oracle_spot_check(robot, solver, params, home, rows, global_idx) if n_check > 0 else 0
    mm = torch.tensor([float(mismatches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(mm)
        # every rank holds every shard: the solved count over the whole job must agree across ranks
        tot = torch.tensor([float((gathered[:, :, n + 1] == 1).sum().item())], dtype=torch.float64, device=dev)
        lo, hi = tot.clone(), tot.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert lo.item() == hi.item(), "ranks disagree on the gathered results"
    assert mm.item() == 0, "%d spot-checked problems differ from the CPU oracle" % int(mm.item())

    # ---- several batches in flight (extra figure; N = 1)
    pipelined = None
    if world == 1 and not args.no_pipelined:
        M = max(2, args.in_flight)
        outs1 = [d_sol, d_err, d_cost, d_its]
        pairs = [(solver, outs1)]
        streams = []
        for _ in range(M - 1):
            sx = torch.cuda.Stream(device=dev)
            streams.append(sx)
            pairs.append((capi.Solver(robot, device=local_rank, stream=sx.cuda_stream),
                          [torch.empty((B, n), dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.int32, device=dev),
                           torch.empty(B, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.int32, device=dev)]))

        def launch(sv, outs):
            sv.solve_batch_async_ptr(params, B, first, d_goal.data_ptr(), d_seed.data_ptr(), 0, outs[0].data_ptr(),
                                     outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), capi.MEM_DEVICE)

        n_batches = M * max(2, args.steps)
        for pr in pairs:  # warm-up of the solvers' buffers
            launch(*pr)
        for sv, _ in pairs:
            sv.wait()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for pr in pairs:
            launch(*pr)
        for k in range(M, n_batches):
            sv, outs = pairs[k % M]
            sv.wait()
            launch(sv, outs)
        for sv, _ in pairs:
            sv.wait()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        for _, outs in pairs:
            assert int((outs[1] == 1).sum().item()) == solved
        pipelined = {"value": n_batches * B / dt, "unit": UNIT, "batches_in_flight": M, "batches": n_batches,
                     "ms_per_batch": 1e3 * dt / n_batches,
                     "how": "%d solvers on %d streams sharing one constant table, pik_solve_batch_async / "
                            "pik_solver_wait, device-resident buffers, wall clock over %d batches" % (M, M, n_batches)}
        for sv, _ in pairs[1:]:
            sv.close()

    if rank == 0:
        n_evals_gd = 2 * n + 3
        P, E = WORKLOAD["population"], WORKLOAD["elites"]
        bytes_gen = 2 * P * (2 * n + 2) * 8
        peak, peak_src = load_peaks()
        achieved = bytes_gen * prob_gens / (gen_ms * 1e-3) / 1e9 if gen_ms > 0 else 0.0
        traffic = load_traffic()
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic.get("dram_bytes_per_step") if traffic else None,
                    "traffic_note": traffic.get("note") if traffic else None,
                    "kernel": "memetic_generation_kernel", "peak_source": peak_src,
                    "algorithmic_bytes_per_problem_generation": bytes_gen,
                    "algorithmic_bytes_per_step": bytes_gen * prob_gens / args.steps,
                    "problem_generations_per_step": prob_gens / args.steps,
                    "generations_per_step": gen_launches / args.steps,
                    "kernel_ms_per_step": gen_ms / args.steps,
                    "kernel_share_of_step": gen_ms / ms if world == 1 else None}
        # the honest bound of this path is FP64 issue, not DRAM: report it beside the contract figure
        evals = gd_steps * n_evals_gd + prob_gens * (2 * E + (P - E) + 1)
        flop_per_eval = FLOP_PER_EVAL["panda"]
        fp64_peak = solver.measure_fp64_peak()
        fp64_ach = evals * flop_per_eval / (gen_ms * 1e-3) / 1e12 if gen_ms > 0 else 0.0
        fp64 = {"achieved_tflops": fp64_ach, "peak_tflops": fp64_peak, "frac": fp64_ach / fp64_peak if fp64_peak else None,
                "peak_source": "measured DFMA microbenchmark (pik_measure_fp64_peak)",
                "cost_evals_per_step": evals / args.steps, "flop_per_eval": flop_per_eval}
        cpu = None
        cpu_a = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, info = cpu_oracle_run(args.cpu_sample, cores)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "first %d poses of the workload (of %d), %d host threads, one pose per thread, %.1f s; "
                             "oracle/pik_oracle.c (restatement; the reference does not build here)"
                             % (args.cpu_sample, B, cores, info["seconds"])}
            cpu_a = cpu_reference_structure_run(args.cpu_sample_a)
        others = None
        if world == 1 and not args.no_other_configs:
            others = other_config_lines(torch, capi, robots, dev, local_rank, peak, fp64_peak, max(2, min(args.steps, 5)))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(world, B, args.scaling),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": int(launches), "roofline": roofline, "fp64": fp64, "cpu_baseline": cpu,
            "cpu_baseline_reference_structure": cpu_a,
            "pipelined": pipelined, "other_configs": others,
            "parity_spot_check": {"problems_per_rank": int(n_check), "ranks": world, "mismatches": int(mm.item()),
                                  "against": "oracle.solve_batch(first_problem_index = global index), bit for bit; "
                                             + ("problems of other ranks' shards, from the gathered block"
                                                if world > 1 else "random problems of the batch")},
            "solved_frac": solved / B, "mean_generations": prob_gens / args.steps / B,
        }
        print(json.dumps(line))
    if comm is not None:
        comm.close()
    solver.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch poses per GPU; strong: --batch poses in total")
    ap.add_argument("--batch", type=int, default=WORKLOAD["poses_per_gpu"], help="poses per GPU (weak) / in total (strong)")
    ap.add_argument("--cpu-sample", type=int, default=WORKLOAD["poses_per_gpu"])
    ap.add_argument("--cpu-sample-a", type=int, default=1024)
    ap.add_argument("--ref-sample", type=int, default=0, help="poses per step of --impl reference (0: the whole batch "
                    "for up to 6 steps, 16384 beyond)")
    ap.add_argument("--spot-check", type=int, default=64, help="problems per rank checked against the CPU oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-pipelined", action="store_true")
    ap.add_argument("--in-flight", type=int, default=2, help="batches in flight of the extra `pipelined` figure")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
