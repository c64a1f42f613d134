"""GPU parity: the CUDA path through the C-ABI (libpik_b200.so) against the CPU oracle on the same
seeded inputs.  Bar (BASELINE.json north_star): bit-exact converged / not-converged flags, converged
joint angles within 1e-5 rad; because both sides share one arithmetic contract we additionally require
bit-equal joint values, costs and iteration counts."""
import numpy as np
import pytest

from oracle import orc
from pick_ik_b200 import capi, robots

pytestmark = pytest.mark.gpu

TOL_RAD = 1e-5  # north_star tolerance for converged joint angles


def both_params(**kw):
    return orc.default_params(**kw), capi.default_params(**kw)


@pytest.fixture(scope="module")
def solvers():
    cache = {}

    def get(name):
        if name not in cache:
            chain = robots.ROBOTS[name]()
            cache[name] = (chain, orc.build_robot(chain.joint_desc()), capi.Solver(capi.Robot(chain)))
        return cache[name]

    yield get
    for _, _, s in cache.values():
        s.close()


def random_configs(orobot, B, seed):
    return np.stack([orc.random_configuration(orobot, seed, b) for b in range(B)])


def check_solve(got, ref, label):
    np.testing.assert_array_equal(got["error_code"], ref["error_code"], err_msg=f"{label}: flags")
    ok = ref["error_code"] == 1
    if ok.any():
        assert np.abs(got["solution"][ok] - ref["solution"][ok]).max() <= TOL_RAD, f"{label}: joints"
    np.testing.assert_array_equal(got["iterations"], ref["iterations"], err_msg=f"{label}: iterations")
    np.testing.assert_array_equal(got["solution"], ref["solution"], err_msg=f"{label}: joints bit-equal")
    np.testing.assert_array_equal(got["cost"], ref["cost"], err_msg=f"{label}: cost bit-equal")


@pytest.mark.parametrize("name", ["rr", "panda", "ur5", "fetch", "skew6", "snake16"])
@pytest.mark.parametrize("goals", [False, True])
def test_eval_cost_bit_exact(solvers, name, goals):
    chain, orobot, solver = solvers(name)
    kw = dict(center_joints_weight=0.3, avoid_joint_limits_weight=0.7, minimal_displacement_weight=0.2,
              cost_threshold=0.05, position_threshold=0.2, orientation_threshold=0.3) if goals else {}
    op, gp = both_params(**kw)
    B = 3000
    q = random_configs(orobot, B, 11)
    q[::7] *= 1.7  # outside the limits as well
    tq = random_configs(orobot, B, 12)
    near = np.arange(B) % 3 == 0
    tq[near] = q[near] + (0.01 if goals else 1e-5)  # some configurations that pass the solution test
    goal = np.stack([orc.pose_from_fk(orobot, t) for t in tq])
    seed = random_configs(orobot, B, 13)
    c_ref, s_ref, tip_ref = orc.eval_cost_batch(orobot, op, goal, seed, q)
    c, s, tip = solver.eval_cost(gp, goal, seed, q)
    np.testing.assert_array_equal(c, c_ref)
    np.testing.assert_array_equal(s, s_ref)
    np.testing.assert_array_equal(tip, tip_ref)
    if not goals:
        assert 0 < s_ref.sum() < B


@pytest.mark.parametrize("name", ["single", "single_prismatic", "rr", "panda", "ur5", "fetch", "skew6", "snake16"])
def test_gd_local_parity(solvers, name):
    chain, orobot, solver = solvers(name)
    op, gp = both_params(mode="local")
    B = 1500
    seed = random_configs(orobot, B, 21)
    rng = np.random.default_rng(5)
    tq = seed + rng.uniform(-0.1, 0.1, seed.shape)
    goal = np.stack([orc.pose_from_fk(orobot, t) for t in tq])
    ref = orc.solve_batch(orobot, op, goal, seed)
    got = solver.solve_batch(gp, goal, seed)
    check_solve(got, ref, name)
    assert 0 < (ref["error_code"] == 1).sum()


def test_gd_local_options(solvers):
    chain, orobot, solver = solvers("fetch")
    B = 600
    seed = random_configs(orobot, B, 31)
    tq = seed + np.random.default_rng(6).uniform(-0.05, 0.05, seed.shape)
    goal = np.stack([orc.pose_from_fk(orobot, t) for t in tq])
    for kw in (dict(stop_optimization_on_valid_solution=0),
               dict(return_approximate_solution=1, gd_max_iters=7),
               dict(rotation_scale=0.0),
               dict(position_scale=0.0),
               dict(center_joints_weight=0.01, avoid_joint_limits_weight=0.01, minimal_displacement_weight=0.02,
                    cost_threshold=0.01, position_threshold=0.01)):
        op, gp = both_params(mode="local", **kw)
        check_solve(solver.solve_batch(gp, goal, seed), orc.solve_batch(orobot, op, goal, seed), str(kw))


MEMETIC_CASES = [
    ("panda", dict(memetic_population_size=16), 300),
    ("panda", dict(memetic_population_size=128), 200),
    ("rr", dict(memetic_population_size=8, memetic_elite_size=2), 200),
    ("ur5", dict(memetic_population_size=32, memetic_elite_size=5, memetic_max_generations=20), 150),
    ("fetch", dict(memetic_population_size=64, center_joints_weight=0.01, avoid_joint_limits_weight=0.01,
                   cost_threshold=0.01, position_threshold=0.01, memetic_max_generations=30), 150),
    ("skew6", dict(memetic_population_size=24, memetic_elite_size=3, memetic_max_generations=25,
                   minimal_displacement_weight=0.01, cost_threshold=0.05), 150),
    ("panda", dict(memetic_population_size=4, memetic_elite_size=4, memetic_max_generations=10), 100),
    ("panda", dict(memetic_population_size=20, memetic_elite_size=1, memetic_max_generations=10), 100),
    ("fetch", dict(memetic_population_size=32, stop_optimization_on_valid_solution=0, memetic_max_generations=6), 100),
    ("panda", dict(memetic_population_size=16, return_approximate_solution=1, memetic_max_generations=3), 100),
    # table limits: 16 variables (kMaxVars), 32 elites (kMaxElites), population 1024 (kMaxPopulation)
    ("snake16", dict(memetic_population_size=40, memetic_elite_size=6, memetic_max_generations=8,
                     memetic_gd_max_iters=6, position_threshold=0.02, orientation_threshold=0.05), 40),
    ("panda", dict(memetic_population_size=64, memetic_elite_size=32, memetic_max_generations=4,
                   memetic_gd_max_iters=5), 24),
    ("rr", dict(memetic_population_size=1024, memetic_elite_size=3, memetic_max_generations=3,
                memetic_gd_max_iters=4), 12),
    # one moving joint (n = 1): every gene mutates with probability 1
    ("single", dict(memetic_population_size=12, memetic_elite_size=2, memetic_max_generations=6, rotation_scale=0.0), 40),
    ("single_prismatic", dict(memetic_population_size=8, memetic_elite_size=1, memetic_max_generations=6,
                              rotation_scale=0.0), 40),
]


@pytest.mark.parametrize("mapping", ["throughput", "wide", "wide-per-generation"])
@pytest.mark.parametrize("name,kw,B", MEMETIC_CASES)
def test_memetic_parity(solvers, name, kw, B, mapping, monkeypatch):
    # the two lane mappings of the generation kernel (1 lane per elite / the warp spread over one
    # problem's GD evaluations, which also keeps a problem for all its generations in one launch)
    monkeypatch.setenv("PIK_WIDE_WARPS_PER_SM", "0" if mapping == "throughput" else "1000000000")
    if mapping == "wide-per-generation":
        monkeypatch.setenv("PIK_TRACE_GENERATIONS", "1")
    chain, orobot, solver = solvers(name)
    op, gp = both_params(mode="global", **kw)
    goal = orc.make_targets(orobot, B)
    if name == "panda":
        seed = np.array(robots.PANDA_HOME)
    else:
        seed = random_configs(orobot, B, 41)
    first = 1000
    ref = orc.solve_batch(orobot, op, goal, seed, first_problem_index=first)
    got = solver.solve_batch(gp, goal, seed, first_problem_index=first)
    check_solve(got, ref, f"{name} {kw}")
    st = solver.stats()
    assert st.solved == (ref["error_code"] == 1).sum()
    assert st.kernel_launches >= 1


@pytest.mark.parametrize("env", ["PIK_NO_STATIC", "PIK_GENERIC_ONLY"])
@pytest.mark.parametrize("mode", ["global", "local"])
def test_less_specific_chain_signatures(env, mode, monkeypatch):
    """The Panda through the kernels of the less specific chain signatures it also matches: origin pattern
    'rotation about x' with run-time n / kinds (PIK_NO_STATIC) and the fully generic kernels."""
    monkeypatch.setenv(env, "1")
    chain = robots.panda()
    orobot = orc.build_robot(chain.joint_desc())
    solver = capi.Solver(capi.Robot(chain))  # the signature is chosen when the solver is created
    kw = dict(mode=mode, memetic_population_size=32, memetic_max_generations=40)
    op, gp = both_params(**kw)
    B = 160
    goal = orc.make_targets(orobot, B)
    seed = np.array(robots.PANDA_HOME) if mode == "global" else random_configs(orobot, B, 61)
    for wide in ("0", "1000000000"):
        monkeypatch.setenv("PIK_WIDE_WARPS_PER_SM", wide)
        check_solve(solver.solve_batch(gp, goal, seed), orc.solve_batch(orobot, op, goal, seed), f"{env} {mode} {wide}")
    solver.close()


def test_memetic_seed_already_valid(solvers):
    chain, orobot, solver = solvers("panda")
    op, gp = both_params(mode="global")
    q = random_configs(orobot, 64, 51)
    goal = np.stack([orc.pose_from_fk(orobot, t) for t in q])
    ref = orc.solve_batch(orobot, op, goal, q)
    got = solver.solve_batch(gp, goal, q)
    check_solve(got, ref, "valid seed")
    assert (got["iterations"] == 0).all() and (got["error_code"] == 1).all()
    np.testing.assert_array_equal(got["solution"], q)


def test_sharding_invariance(solvers):
    """A shard solved with first_problem_index reproduces the unsharded result (RNG keyed by the
    global problem index)."""
    chain, orobot, solver = solvers("panda")
    _, gp = both_params(mode="global", memetic_population_size=32)
    B = 128
    goal = orc.make_targets(orobot, B)
    seed = np.array(robots.PANDA_HOME)
    whole = solver.solve_batch(gp, goal, seed)
    lo = solver.solve_batch(gp, goal[:50], seed, first_problem_index=0)
    hi = solver.solve_batch(gp, goal[50:], seed, first_problem_index=50)
    for k in ("solution", "error_code", "cost", "iterations"):
        np.testing.assert_array_equal(np.concatenate([lo[k], hi[k]]), whole[k])


def test_empty_batch_and_errors(solvers):
    chain, orobot, solver = solvers("panda")
    gp = capi.default_params()
    out = solver.solve_batch(gp, np.zeros((0, 7)), np.array(robots.PANDA_HOME))
    assert out["solution"].shape == (0, 7)
    bad = capi.default_params(memetic_elite_size=40, memetic_population_size=16)
    with pytest.raises(capi.PikError):
        solver.solve_batch(bad, np.zeros((1, 7)), np.array(robots.PANDA_HOME))


SPECIES_CASES = [
    ("panda", dict(memetic_population_size=16, memetic_num_threads=3, memetic_max_generations=40), 120),
    ("panda", dict(memetic_population_size=16, memetic_num_threads=4, memetic_stop_on_first_solution=0,
                   memetic_max_generations=25), 80),
    ("panda", dict(memetic_population_size=16, memetic_num_threads=5, return_approximate_solution=1,
                   memetic_max_generations=30), 80),
    ("fetch", dict(memetic_population_size=32, memetic_num_threads=2, stop_optimization_on_valid_solution=0,
                   memetic_max_generations=6), 60),
    ("rr", dict(memetic_population_size=8, memetic_elite_size=2, memetic_num_threads=7, memetic_max_generations=12), 60),
]


@pytest.mark.parametrize("mapping", ["throughput", "wide", "default"])
@pytest.mark.parametrize("name,kw,B", SPECIES_CASES)
def test_memetic_species_parity(solvers, name, kw, B, mapping, monkeypatch):
    """memetic_num_threads species per problem (src/ik_memetic.cpp:315-370): lockstep schedule, `terminate`, and the
    min-fitness pick in arrival order, against the oracle's restatement."""
    if mapping != "default":
        monkeypatch.setenv("PIK_WIDE_WARPS_PER_SM", "0" if mapping == "throughput" else "1000000000")
    chain, orobot, solver = solvers(name)
    op, gp = both_params(mode="global", **kw)
    goal = orc.make_targets(orobot, B)
    seed = np.array(robots.PANDA_HOME) if name == "panda" else random_configs(orobot, B, 43)
    first = 500
    ref = orc.solve_batch(orobot, op, goal, seed, first_problem_index=first)
    got = solver.solve_batch(gp, goal, seed, first_problem_index=first)
    check_solve(got, ref, f"{name} {kw}")
    assert 0 < (ref["error_code"] == 1).sum()


def test_sub_batches_reproduce_the_undivided_solve(solvers, monkeypatch):
    """PIK_SUB_BATCHES (experiment): the batch cut into sub-batches on streams of their own gives the bits of the
    undivided batch (RNG keyed by the global problem index), species included."""
    chain, orobot, solver = solvers("panda")
    B = 310
    goal = orc.make_targets(orobot, B)
    seed = np.array(robots.PANDA_HOME)
    for kw in (dict(memetic_population_size=16, memetic_max_generations=30),
               dict(memetic_population_size=16, memetic_max_generations=20, memetic_num_threads=3)):
        op, gp = both_params(mode="global", **kw)
        ref = orc.solve_batch(orobot, op, goal, seed, first_problem_index=77)
        for k in ("3", "8"):
            monkeypatch.setenv("PIK_SUB_BATCHES", k)
            check_solve(solver.solve_batch(gp, goal, seed, first_problem_index=77), ref, f"sub-batches {k} {kw}")
            assert solver.stats().solved == (ref["error_code"] == 1).sum()


def test_axis_within_tolerance_is_normalised():
    """pik_robot_create admits |axis|^2 within 1e-6 of 1; the table is built from the unit axis
    (RevoluteJointModel::setAxis normalises), in the library and in the oracle alike."""
    desc = robots.skew6().joint_desc().copy()
    for j in range(len(desc)):
        desc[j]["axis"] = tuple(np.array(desc[j]["axis"]) * (1.0 + 4.0e-7))
    orobot = orc.build_robot(desc)
    solver = capi.Solver(capi.Robot(desc))
    op, gp = both_params(mode="local")
    B = 300
    seed = random_configs(orobot, B, 3)
    goal = np.stack([orc.pose_from_fk(orobot, s + 0.05) for s in seed])
    check_solve(solver.solve_batch(gp, goal, seed), orc.solve_batch(orobot, op, goal, seed), "scaled axes")
    solver.close()


@pytest.mark.parametrize("name,kw", [
    ("panda", dict(memetic_population_size=128)),
    ("ur5", dict(memetic_population_size=48, memetic_elite_size=5, memetic_max_generations=40)),
    ("skew6", dict(memetic_population_size=30, memetic_elite_size=3, memetic_max_generations=40)),
])
def test_reproduce_side_by_side_equals_one_at_a_time(solvers, name, kw, monkeypatch):
    """The throughput mapping lets the problems of a warp reproduce side by side (32 / PW lanes each, running
    E best + worst); PIK_LOCKSTEP_MASK bit 3 selects the one-problem-at-a-time form (windows of 32, fitness array).
    Same bits, a batch large enough for multi-round launches, with and without the CTA-wide round barrier."""
    chain, orobot, solver = solvers(name)
    B = 6000
    goal = orc.make_targets(orobot, B)
    seed = np.array(robots.PANDA_HOME) if name == "panda" else random_configs(orobot, 1, 9)[0]
    gp = capi.default_params(mode="global", **kw)
    monkeypatch.setenv("PIK_LOCKSTEP_MASK", "11")
    ref = solver.solve_batch(gp, goal, seed)
    for mask in ("3", "1", "0"):
        monkeypatch.setenv("PIK_LOCKSTEP_MASK", mask)
        got = solver.solve_batch(gp, goal, seed)
        for key in ("solution", "error_code", "cost", "iterations"):
            np.testing.assert_array_equal(got[key], ref[key], err_msg=f"{name} mask {mask} {key}")
    # and the one-at-a-time form against the oracle on a slice
    op = orc.default_params(mode="global", **kw)
    oref = orc.solve_batch(orobot, op, goal[:48], seed)
    for key in ("solution", "error_code", "cost", "iterations"):
        np.testing.assert_array_equal(ref[key][:48], oref[key], err_msg=f"{name} oracle {key}")


@pytest.mark.parametrize("kw,B", [
    (dict(memetic_population_size=16, memetic_max_generations=30), 1500),
    # nothing converges: every problem runs all its generations, many of them a launch or more late
    (dict(memetic_population_size=8, memetic_max_generations=9, position_threshold=1e-9, orientation_threshold=1e-9), 700),
    (dict(memetic_population_size=12, memetic_max_generations=14, memetic_num_threads=2,
          memetic_stop_on_first_solution=0), 400),
])
def test_whole_wave_launches(solvers, kw, B, monkeypatch):
    """A throughput launch processes whole waves of its resident CTAs and passes the rest of its list on to the next
    launch, so a problem's generation lags behind the launch index (PIK_WAVE_CTAS shrinks the wave so that batches
    of test size take that path; PIK_DEFER_LAUNCHES=0 turns it off).  Same bits as the oracle either way."""
    chain, orobot, solver = solvers("panda")
    goal = orc.make_targets(orobot, B)
    seed = np.array(robots.PANDA_HOME)
    op, gp = both_params(mode="global", **kw)
    ref = orc.solve_batch(orobot, op, goal, seed, first_problem_index=5)
    monkeypatch.setenv("PIK_WIDE_WARPS_PER_SM", "0")  # throughput mapping throughout
    for ctas, defer in (("3", None), ("1", "40"), ("2", "0")):
        monkeypatch.setenv("PIK_WAVE_CTAS", ctas)
        if defer is None:
            monkeypatch.delenv("PIK_DEFER_LAUNCHES", raising=False)
        else:
            monkeypatch.setenv("PIK_DEFER_LAUNCHES", defer)
        check_solve(solver.solve_batch(gp, goal, seed, first_problem_index=5), ref, f"wave of {ctas} CTAs, defer {defer}")
        st = solver.stats()
        assert st.problems == B and st.solved == (ref["error_code"] == 1).sum()
    # default wide mapping behind the whole-wave launches
    monkeypatch.delenv("PIK_WIDE_WARPS_PER_SM")
    monkeypatch.setenv("PIK_WAVE_CTAS", "2")
    monkeypatch.delenv("PIK_DEFER_LAUNCHES", raising=False)
    check_solve(solver.solve_batch(gp, goal, seed, first_problem_index=5), ref, "wave of 2 CTAs, wide tail")
