#!/usr/bin/env python
"""bench.py -- IK solves/sec on BASELINE.json's headline workload.

A step = one pass of the hot path over one batch: 65 536 random reachable Franka Panda poses per GPU,
memetic solver (population 128, 4 elites, 25 GD iterations per elite, <= 100 generations), seed state =
Panda home.  `value` times device-resident inputs -> device-resident outputs with CUDA events; `e2e` times
the same call through the C-ABI with pinned HOST buffers (H2D and D2H inside).  N > 1 (torchrun): one
process per GPU, weak scaling (each rank its own 65 536 poses, RNG keyed by the global problem index), one
NCCL all-gather of the packed solutions inside the timed region (pik_solve_batch_sharded: the library's own
communicator, torch.distributed only carries the unique id and the barriers), max over ranks.

`--impl reference` times the reference's CPU algorithm (oracle/pik_oracle.c, a restatement: the reference
itself needs ROS 2 / MoveIt / Eigen and cannot be built here) on all host cores on a bounded sample.
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


def solver_kwargs():
    return dict(mode="global", memetic_population_size=WORKLOAD["population"],
                memetic_elite_size=WORKLOAD["elites"], memetic_gd_max_iters=WORKLOAD["gd_iters"],
                memetic_max_generations=WORKLOAD["max_generations"])


def random_reachable_configs(chain, B: int, block: int) -> np.ndarray:
    """q*_b ~ U(limits) per variable (continuous joints: U(-pi, pi)); block = shard index."""
    jd = chain.joint_desc()
    mv = jd[jd["type"] != 0]
    lo = np.where(mv["bounded"] != 0, mv["min_position"], -np.pi)
    hi = np.where(mv["bounded"] != 0, mv["max_position"], np.pi)
    u = np.random.default_rng([TARGET_SEED, block]).random((B, len(mv)))
    return lo + (hi - lo) * u


def config_dict(n_gpus: int, B: int):
    return {
        "workload": "Panda 7-DoF, %d random reachable poses per GPU, memetic pop=128 elite=4 gd_iters=25, "
                    "<=100 generations, seed=home (BASELINE.json configs[1])" % B,
        "robot": "panda", "poses_per_gpu": B, "global_poses": B * n_gpus, "population": WORKLOAD["population"],
        "elites": WORKLOAD["elites"], "max_generations": WORKLOAD["max_generations"],
        "parallelism": "pose-batch sharded x%d" % n_gpus,
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
def cpu_oracle_run(B: int, threads: int, block: int = 0):
    """Times the CPU restatement (oracle) on the first B poses of the workload; returns (solves/s, dict)."""
    from oracle import orc
    from pick_ik_b200 import robots

    chain = robots.panda()
    orobot = orc.build_robot(chain.joint_desc())
    op = orc.default_params(**solver_kwargs())
    q = random_reachable_configs(chain, B, block)
    zeros = np.zeros((B, 7))
    zeros[:, 3] = 1.0
    home = np.array(robots.PANDA_HOME)
    _, _, goal = orc.eval_cost_batch(orobot, op, zeros, home, q)
    t0 = time.perf_counter()
    res = orc.solve_batch(orobot, op, goal, home, first_problem_index=0, n_threads=threads)
    dt = time.perf_counter() - t0
    return B / dt, dict(seconds=dt, solved=int((res["error_code"] == 1).sum()), evals=int(res["evals"]))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    sample = args.ref_sample
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
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus, WORKLOAD["poses_per_gpu"]),
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
    """dram bytes per generation-kernel launch from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


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

    B = args.batch
    chain = robots.panda()
    n = chain.num_variables
    robot = capi.Robot(chain)
    stream = torch.cuda.current_stream()
    solver = capi.Solver(robot, device=local_rank, stream=stream.cuda_stream)
    params = capi.default_params(**solver_kwargs())
    home = np.array(robots.PANDA_HOME)

    # synthetic reachable targets: FK of random valid configurations (computed by the engine itself)
    qstar = random_reachable_configs(chain, B, rank)
    ident = np.zeros((B, 7))
    ident[:, 3] = 1.0
    _, _, goal_np = solver.eval_cost(params, ident, home, qstar)

    first = rank * B
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
            # shard solve + ncclAllGather of the packed results on the solver's stream (pik_solve_batch_sharded)
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
    e2e_steps = max(1, min(args.steps, 5))
    if world == 1:
        d2h = sum(v.nbytes for v in h_out.values())

        def step_e2e():
            solver.solve_batch(params, h_goal.numpy(), h_seed.numpy(), first, out=h_out)
    else:
        # host goal poses in, the gathered results of every rank out (pinned), the NCCL all-gather in between
        h_gathered = torch.empty((world, B, n + 3), dtype=torch.float64).pin_memory()
        d2h = h_gathered.numel() * 8

        def step_e2e():
            solver.solve_batch_sharded_ptr(comm, params, B, first, h_goal.data_ptr(), h_seed.data_ptr(), 0,
                                           h_gathered.data_ptr(), capi.MEM_HOST)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * B * e2e_steps / e2e_s
    e2e_solved = int((h_out["error_code"] == 1).sum()) if world == 1 else int((h_gathered[rank, :, n + 1] == 1).sum())
    assert e2e_solved == solved, "e2e and device-resident runs disagree"
    if world > 1:
        # every rank holds every shard: the solved count over the whole job must agree across ranks
        tot = torch.tensor([float((h_gathered[:, :, n + 1] == 1).sum())], dtype=torch.float64, device=dev)
        lo, hi = tot.clone(), tot.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert lo.item() == hi.item(), "ranks disagree on the gathered results"

    if rank == 0:
        n_evals_gd = 2 * n + 3
        P, E = WORKLOAD["population"], WORKLOAD["elites"]
        bytes_gen = 2 * P * (2 * n + 2) * 8
        peak, peak_src = load_peaks()
        achieved = bytes_gen * prob_gens / (gen_ms * 1e-3) / 1e9 if gen_ms > 0 else 0.0
        traffic = load_traffic()
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                    "kernel": "memetic_generation_kernel", "peak_source": peak_src,
                    "algorithmic_bytes_per_problem_generation": bytes_gen,
                    "problem_generations_per_step": prob_gens / args.steps,
                    "launches_per_step": gen_launches / args.steps,
                    "kernel_ms_per_step": gen_ms / args.steps,
                    "kernel_share_of_step": gen_ms / ms if world == 1 else None}
        # the honest bound of this path is FP64 issue, not DRAM: report it beside the contract figure
        evals = gd_steps * n_evals_gd + prob_gens * (2 * E + (P - E) + 1)
        flop_per_eval = 1050.0
        fp64_peak = solver.measure_fp64_peak()
        fp64_ach = evals * flop_per_eval / (gen_ms * 1e-3) / 1e12 if gen_ms > 0 else 0.0
        fp64 = {"achieved_tflops": fp64_ach, "peak_tflops": fp64_peak, "frac": fp64_ach / fp64_peak if fp64_peak else None,
                "peak_source": "measured DFMA microbenchmark (pik_measure_fp64_peak)",
                "cost_evals_per_step": evals / args.steps, "flop_per_eval": flop_per_eval}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, info = cpu_oracle_run(args.cpu_sample, cores)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "first %d poses of the workload, %d host threads, one pose per thread, %.1f s; "
                             "oracle/pik_oracle.c (restatement; the reference does not build here)"
                             % (args.cpu_sample, cores, info["seconds"])}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(world, B),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": int(launches), "roofline": roofline, "fp64": fp64, "cpu_baseline": cpu,
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
    ap.add_argument("--batch", type=int, default=WORKLOAD["poses_per_gpu"], help="poses per GPU")
    ap.add_argument("--cpu-sample", type=int, default=16384)
    ap.add_argument("--ref-sample", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
