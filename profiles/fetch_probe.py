"""Fetch-256 (BASELINE.json configs[3]) probe: one solve of B poses; with PIK_TRACE / a -DPIK_PHASE_TRACE build
(PIK_LIB_PATH, PIK_DEBUG_PHASES=1) prints the per-generation / per-phase picture.  usage: python profiles/fetch_probe.py [B] [robot]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pick_ik_b200 import capi, robots  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
name = sys.argv[2] if len(sys.argv) > 2 else "fetch"
robot = capi.Robot(robots.ROBOTS[name]())
solver = capi.Solver(robot)
if name == "fetch":
    kw = dict(mode="global", memetic_population_size=256, center_joints_weight=0.01, avoid_joint_limits_weight=0.01,
              cost_threshold=0.01, position_threshold=0.01)
    if os.environ.get("PROBE_NO_GOALS"):
        kw.update(center_joints_weight=0.0, avoid_joint_limits_weight=0.0)
    seed = np.array([robot.variable(i).mid if robot.variable(i).bounded else 0.0 for i in range(robot.n)])
elif name == "ur5":
    kw = dict(mode="local")
    seed = robot.random_configurations(B, 0xC0FFEE + 1)
else:
    kw = dict(mode="global", memetic_population_size=128)
    seed = np.array(robots.PANDA_HOME)
params = capi.default_params(**kw)
ident = np.zeros((B, 7)); ident[:, 3] = 1.0
q = robot.random_configurations(B, 0xC0FFEE)
if name == "ur5":
    q = seed + 0.1 * np.sin(q * 7.0)
goal = solver.eval_cost(params, ident, seed, q)[2]
for _ in range(2):
    res = solver.solve_batch(params, goal, seed)
st = solver.stats()
print("signature", robot.chain_signature(), "device_ms", st.device_ms, "solved", st.solved, "gd_steps", st.gd_steps,
      "problem_generations", st.problem_generations, file=sys.stderr)
