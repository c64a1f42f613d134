/*
 * pik_oracle.h -- CPU ORACLE for the pick_ik hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a dependency-free plain-C restatement of the reference's algorithm for the
 * memetic + numeric-gradient IK path.  It is the checker the CUDA product is compared
 * against.  Nothing under pick_ik_b200/ may include, link or call it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Reference files restated (paths relative to the pick_ik source tree @ 8c99999):
 *   src/robot.cpp:23-105           variable table, clamp, validity, random valid configuration
 *   src/fk_moveit.cpp:20-34        what FK returns (tip frames); the chain-walk arithmetic is
 *                                  moveit_core's RobotState (absent from the tree, unpinned) and is
 *                                  restated from its published algorithm;
 *                                  per-joint-type semantics also follow src/forward_kinematics.cpp:39-80
 *   src/goal.cpp:17-144,163-203    distances, frame tests, pose cost, joint goals, solution test, cost
 *   src/ik_gradient.cpp:14-139     GradientIk::from, step, ik_gradient
 *   src/ik_memetic.cpp:18-373      MemeticIk, ik_memetic_impl, ik_memetic (one species or several, lockstep)
 *   src/pick_ik_plugin.cpp:88-217  goal assembly, threshold enabling, YAML -> solver param mapping
 *
 * PARITY PINNING.  The reference cannot be built in this environment (needs ROS 2, MoveIt, Eigen,
 * rsl, urdfdom, Catch2).  The oracle is pinned against every known-answer the reference's own tests
 * hold for this path (tests/goal_tests.cpp, tests/ik_tests.cpp, tests/ik_memetic_tests.cpp) -- see
 * tests/test_oracle_reference_vectors.py.  The reference's RNG stream is unseeded
 * (rsl thread-local mt19937 seeded from random_device), so memetic joint values are
 * PARITY-UNPINNED at the RNG boundary: the oracle defines the seeded counter-based stream
 * (Philox4x32-10) that the CUDA product must reproduce.
 *
 * Arithmetic contract (what makes CPU and GPU bit-identical): IEEE-754 binary64 only;
 * + - * / sqrt fma, each correctly rounded; no contraction other than the fma() calls written
 * out here (compile with -ffp-contract=off); own sincos / atan2 built from those operations.
 * Wall-clock limits of the reference are replaced by iteration caps (max_time = +inf).
 */
#ifndef PIK_ORACLE_H
#define PIK_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_VARS 16
#define ORC_MAX_STEPS 16
#define ORC_MAX_TIPS 4

/* joint kinds in a robot description (before folding of fixed joints).  FLOATING (7 variables: x y z, quaternion
 * x y z w) and PLANAR (3 variables: x y theta) follow forward_kinematics.cpp:64-79. */
enum { ORC_JOINT_FIXED = 0, ORC_JOINT_REVOLUTE = 1, ORC_JOINT_PRISMATIC = 2, ORC_JOINT_FLOATING = 3,
       ORC_JOINT_PLANAR = 4 };

/* Same memory layout as pik_joint_desc in include/pik.h (both are filled from the same fixtures). */
typedef struct {
    int32_t type;        /* ORC_JOINT_* */
    int32_t bounded;     /* position_bounded_ (continuous revolute = 0) */
    double origin_R[9];  /* parent-link -> joint frame rotation, row-major */
    double origin_t[3];
    double axis[3];      /* unit axis for revolute / prismatic */
    double min_position; /* bounds.min_position_ / max_position_ (continuous: -pi / +pi) */
    double max_position;
    double max_velocity; /* bounds.max_velocity_ (0 = none) */
} orc_joint_desc;

/* Robot::Variable, include/pick_ik/robot.hpp:15-37 */
typedef struct {
    double min, max, mid;
    double half_span;
    double max_velocity_rcp;
    double minimal_displacement_factor;
    int32_t bounded;
    int32_t pad_;
} orc_variable;

/* step kinds after folding: which joint motion follows the folded constant origin */
enum { ORC_STEP_REV_X = 0, ORC_STEP_REV_Y = 1, ORC_STEP_REV_Z = 2, ORC_STEP_REV_GENERAL = 3,
       ORC_STEP_PRISMATIC = 4, ORC_STEP_FLOATING = 8, ORC_STEP_PLANAR = 9 };

/* One moving joint of the kinematic tree.  What FK returns for the tips is what MoveIt's
 * RobotState::updateLinkTransforms computes (fk_moveit.cpp:20-34): the frame of a joint's child link =
 * frame of the parent link * joint origin * joint motion, walked root to leaf. */
typedef struct {
    int32_t kind;
    int32_t parent; /* step whose frame the folded origin is applied to; -1 = the model root */
    int32_t var0;   /* first variable of the joint (a mimic joint: the variable of the joint it follows) */
    int32_t pad_;
    double sign;  /* +1 / -1 for axis-aligned revolute (rotation about -axis = angle negated) */
    double R[9];  /* folded constant origin: product of the fixed transforms since the parent  */
    double t[3];  /* moving joint, times this joint's origin                                   */
    double axis[3];
    double axis_sq[6]; /* xx, yy, zz, xy, xz, yz (cached like RevoluteJointModel::setAxis) */
    double mimic_factor, mimic_offset; /* joint value = q[var0] * factor + offset (1, 0 unless a mimic joint) */
} orc_step;

typedef struct {
    int32_t n;       /* active variables (robot.cpp:122-160: the non-mimic joints below the tips) */
    int32_t has_tip; /* serial chains: tip_has[0] */
    int32_t n_steps; /* moving joints, parents before children */
    int32_t n_tips;
    orc_step steps[ORC_MAX_STEPS];
    orc_variable vars[ORC_MAX_VARS];
    int32_t tip_step[ORC_MAX_TIPS]; /* step the tip link hangs on (-1: the model root) */
    int32_t tip_has[ORC_MAX_TIPS];  /* fixed transform(s) between that step and the tip link */
    double tip_R[ORC_MAX_TIPS][9];
    double tip_t[ORC_MAX_TIPS][3];
} orc_robot;

/* Mirror of the YAML parameter set (src/pick_ik_parameters.yaml) + solver structs. */
typedef struct {
    int32_t mode; /* 0 = global (memetic), 1 = local (gradient) */
    int32_t gd_max_iters;
    double gd_step_size;
    double gd_min_cost_delta;
    double position_threshold;
    double orientation_threshold;
    double cost_threshold;
    double position_scale;
    double rotation_scale;
    double center_joints_weight;
    double avoid_joint_limits_weight;
    double minimal_displacement_weight;
    double memetic_wipeout_fitness_tol;
    int32_t stop_optimization_on_valid_solution;
    int32_t memetic_population_size;
    int32_t memetic_elite_size;
    int32_t memetic_max_generations;
    int32_t memetic_gd_max_iters;
    int32_t return_approximate_solution;
    int32_t memetic_num_threads;            /* species (ik_memetic.cpp:315-370) */
    int32_t memetic_stop_on_first_solution;
    uint64_t rng_seed;
} orc_params;

/* One IK problem: robot + one goal frame per tip (goal.cpp:80-89, assert at goal.cpp:169) + seed state (for
 * minimal displacement) + params */
typedef struct {
    const orc_robot* robot;
    const orc_params* params;
    double goal_t[ORC_MAX_TIPS][3];
    double goal_R[ORC_MAX_TIPS][9];
    double goal_q[ORC_MAX_TIPS][4]; /* Quaterniond(goal_R): w,x,y,z */
    double seed[ORC_MAX_VARS];
} orc_problem;

/* ---- deterministic math ---- */
void orc_sincos(double x, double* s, double* c);
double orc_atan2(double y, double x);
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* ---- frames ---- */
void orc_quat_to_matrix(const double q_wxyz[4], double R[9]);  /* Eigen toRotationMatrix, no normalisation */
void orc_matrix_to_quat(const double R[9], double q_wxyz[4]);  /* Eigen Quaterniond(Matrix3d) */
double orc_linear_distance(const double t1[3], const double t2[3]);
double orc_angular_distance_q(const double q_goal[4], const double R_tip[9]);
double orc_angular_distance(const double R_goal[9], const double R_tip[9]);
/* goal.cpp:27-36; threshold < 0 means "disabled" (std::nullopt) */
int orc_frame_test(const double goal_t[3], const double goal_R[9], const double tip_t[3],
                   const double tip_R[9], double position_threshold, double orientation_threshold);
/* goal.cpp:51-78 */
double orc_pose_cost(const double goal_t[3], const double goal_R[9], const double tip_t[3],
                     const double tip_R[9], double position_scale, double rotation_scale);

/* ---- robot ---- */
/* serial chain model root -> tip: joints in chain order, one tip behind the last joint */
int orc_robot_build(const orc_joint_desc* joints, int n_joints, orc_robot* out);
/* kinematic tree with several tips.  parent[j]: the joint whose child link joint j hangs on (-1: the model root;
 * parents precede children); tip_joint[t]: the joint whose child link is tip t; mimic_of[j] (or NULL): the joint a
 * mimic joint follows, value = factor * master + offset (mimic joints contribute no variable, robot.cpp:145-147).
 * Variables are numbered in joint order. */
int orc_robot_build_tree(const orc_joint_desc* joints, int n_joints, const int32_t* parent,
                         const int32_t* tip_joint, int n_tips, const int32_t* mimic_of,
                         const double* mimic_factor, const double* mimic_offset, orc_robot* out);
/* frame of tip 0 */
void orc_fk(const orc_robot* robot, const double* q, double R[9], double t[3]);
/* frames of all tips: R [n_tips][9], t [n_tips][3] */
void orc_fk_tips(const orc_robot* robot, const double* q, double* R, double* t);
double orc_clamp_to_limits(const orc_variable* v, double val);
int orc_is_valid_configuration(const orc_robot* robot, const double* q);

/* ---- goals / cost ---- */
void orc_params_default(orc_params* p);
void orc_problem_init(orc_problem* pb, const orc_robot* robot, const orc_params* params,
                      const double* goal_pose /* [n_tips][7]: px py pz qw qx qy qz */, const double* seed);
double orc_center_joints_cost(const orc_robot* robot, const double* q);
double orc_avoid_joint_limits_cost(const orc_robot* robot, const double* q);
double orc_minimal_displacement_cost(const orc_robot* robot, const double* q, const double* seed);
double orc_cost(const orc_problem* pb, const double* q);
int orc_is_solution(const orc_problem* pb, const double* q);

/* ---- solvers ---- */
/* One GD step on caller-provided state (tests). state arrays have n entries. Returns improved. */
int orc_gd_step(const orc_problem* pb, double* gradient, double* working, double* local,
                double* best, double* local_cost, double* best_cost);

typedef struct {
    int32_t found;        /* optional has_value */
    int32_t iterations;   /* GD iterations (local) or generations executed (global) */
    double cost;          /* best cost / fitness */
    uint64_t evals;       /* cost-function evaluations (as the reference would execute them) */
    uint32_t wipeouts;
    uint32_t gd_steps;    /* total GD step() calls */
    double solution[ORC_MAX_VARS];
} orc_result;

void orc_ik_gradient(const orc_problem* pb, const double* initial_guess, orc_result* out);
void orc_ik_memetic(const orc_problem* pb, const double* initial_guess, uint32_t problem_index,
                    orc_result* out);
/* ik_memetic with num_threads species in the lockstep schedule (see pik_oracle.c) */
void orc_ik_memetic_species(const orc_problem* pb, const double* initial_guess, uint32_t problem_index,
                            int n_species, int stop_on_first, orc_result* out);

/* Batch driver = plugin mapping (pick_ik_plugin.cpp:209-217): error_code 1 / -31, solution = seed on
 * failure.  goal_pose [B][n_tips][7].  seed_stride = 0 broadcasts one seed.  n_threads <= 0: all cores. */
void orc_solve_batch(const orc_robot* robot, const orc_params* params, int64_t B,
                     int64_t first_problem_index, const double* goal_pose, const double* seed,
                     int64_t seed_stride, double* solution, int32_t* error_code, double* cost,
                     int32_t* iterations, uint64_t* evals_total, int n_threads);

/* Baseline A (timing only): the reference's execution structure -- one solve at a time, the elites' gradient descents
 * of every generation in E new threads behind one FK mutex (ik_memetic.cpp:230-239, fk_moveit.cpp:21).  Global mode,
 * one species.  Same results as orc_solve_batch. */
void orc_solve_batch_reference_structure(const orc_robot* robot, const orc_params* params, int64_t B,
                                         int64_t first_problem_index, const double* goal_pose, const double* seed,
                                         int64_t seed_stride, double* solution, int32_t* error_code, double* cost,
                                         int32_t* iterations);

/* Batched FK + cost (checker for pik_eval_cost): q [B][n], goal_pose [B][n_tips][7], tip_pose [B][n_tips][7] */
void orc_eval_cost_batch(const orc_robot* robot, const orc_params* params, int64_t B,
                         const double* goal_pose, const double* seed, int64_t seed_stride,
                         const double* q, double* cost, int32_t* is_solution, double* tip_pose);

/* Synthetic target generator: q* ~ U(limits) from Philox stream (gen_seed, b). */
void orc_random_configuration(const orc_robot* robot, uint64_t gen_seed, uint32_t problem_index,
                              double* q);
void orc_pose_from_fk(const orc_robot* robot, const double* q, double pose[7]);        /* tip 0 */
void orc_poses_from_fk(const orc_robot* robot, const double* q, double* pose /* [n_tips][7] */);

#ifdef __cplusplus
}
#endif
#endif
