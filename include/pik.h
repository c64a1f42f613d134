/*
 * pik.h -- C-ABI of the B200-native batched IK engine (libpik_b200.so).
 *
 * This is the drop-in boundary for pick_ik's memetic + numeric-gradient hot path.  Every entry
 * point cites the reference interface it replaces (paths relative to pick_ik @ 8c99999).  Plain
 * pointers and sizes only; no C++ / torch types; functions return 0 (PIK_OK) or a negative
 * PIK_E_* status and never throw.  Handles are thread-compatible: one in-flight call per solver.
 *
 * Arithmetic: IEEE binary64 on the device (the reference computes in `double` everywhere).
 * Wall-clock limits of the reference (max_time, memetic_gd_max_time, the plugin timeout) are
 * replaced by the iteration caps of the same parameter set; see DESIGN.md.
 */
#ifndef PIK_H
#define PIK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIK_MAX_VARS 16
#define PIK_MAX_ELITES 32

enum {
    PIK_OK = 0,
    PIK_E_INVALID_ARGUMENT = -1,
    PIK_E_INVALID_ROBOT = -2,
    PIK_E_INVALID_PARAMS = -3, /* a pick_ik_parameters.yaml validator failed */
    PIK_E_CUDA = -4,
    PIK_E_NO_DEVICE = -5,
    PIK_E_OUT_OF_MEMORY = -6,
    PIK_E_UNSUPPORTED = -7,
    PIK_E_NCCL = -8, /* NCCL missing or a collective failed: pik_comm_last_error() */
    PIK_E_BUSY = -9  /* a pik_solve_batch_async is in flight on this solver: pik_solver_wait first */
};

/* moveit_msgs::msg::MoveItErrorCodes values written by pick_ik_plugin.cpp:212,215 */
enum { PIK_SUCCESS = 1, PIK_NO_IK_SOLUTION = -31 };

/* FLOATING: 7 variables x y z, quaternion x y z w; PLANAR: 3 variables x y theta (src/forward_kinematics.cpp:64-79) */
enum { PIK_JOINT_FIXED = 0, PIK_JOINT_REVOLUTE = 1, PIK_JOINT_PRISMATIC = 2, PIK_JOINT_FLOATING = 3, PIK_JOINT_PLANAR = 4 };
#define PIK_MAX_TIPS 4
enum { PIK_MODE_GLOBAL = 0, PIK_MODE_LOCAL = 1 }; /* yaml `mode`: "global" | "local" */
enum { PIK_MEM_HOST = 0, PIK_MEM_DEVICE = 1 };

/* One joint of the robot (serial chain: in chain order, model root -> tip link).  Replaces what
 * Robot::from (src/robot.cpp:44-85) and make_fk_fn (src/fk_moveit.cpp:11-35) read from the MoveIt
 * RobotModel: joint type, LinkModel::getJointOriginTransform, joint axis, VariableBounds. */
typedef struct pik_joint_desc {
    int32_t type;        /* PIK_JOINT_* */
    int32_t bounded;     /* VariableBounds::position_bounded_ (URDF continuous = 0) */
    double origin_R[9];  /* parent link -> joint frame rotation, row-major */
    double origin_t[3];
    double axis[3];      /* unit axis (revolute / prismatic) */
    double min_position; /* VariableBounds::min_position_ / max_position_ (continuous: -pi / pi) */
    double max_position;
    double max_velocity; /* VariableBounds::max_velocity_, 0 = none */
    /* FLOATING / PLANAR joints: bounded / min / max apply to the translation variables (MoveIt's default: unbounded);
     * the quaternion components are bounded to [-1, 1] and the planar angle is unbounded, as MoveIt's joint models
     * set them; max_velocity applies to every variable of the joint */
} pik_joint_desc;

/* Robot::Variable, include/pick_ik/robot.hpp:15-37 */
typedef struct pik_variable {
    double min, max, mid, half_span, max_velocity_rcp, minimal_displacement_factor;
    int32_t bounded;
    int32_t pad_;
} pik_variable;

/* src/pick_ik_parameters.yaml, same names and defaults; the YAML -> solver mapping of
 * pick_ik_plugin.cpp:166-196 is applied inside the library. */
typedef struct pik_params {
    int32_t mode;
    int32_t gd_max_iters;
    double gd_step_size;
    double gd_min_cost_delta;
    double position_threshold;
    double orientation_threshold;
    double approximate_solution_position_threshold;    /* used by the plugin shim only */
    double approximate_solution_orientation_threshold; /* used by the plugin shim only */
    double approximate_solution_joint_threshold;       /* used by the plugin shim only */
    double approximate_solution_cost_threshold;        /* used by the plugin shim only */
    double cost_threshold;
    double position_scale;
    double rotation_scale;
    double center_joints_weight;
    double avoid_joint_limits_weight;
    double minimal_displacement_weight;
    double memetic_wipeout_fitness_tol;
    double memetic_gd_max_time; /* accepted, ignored: replaced by memetic_gd_max_iters */
    int32_t stop_optimization_on_valid_solution;
    int32_t memetic_num_threads;            /* species per problem (ik_memetic.cpp:315-370), run in lockstep */
    int32_t memetic_stop_on_first_solution; /* the first species to return a value terminates the others */
    int32_t memetic_population_size;
    int32_t memetic_elite_size;
    int32_t memetic_max_generations;
    int32_t memetic_gd_max_iters;
    int32_t return_approximate_solution; /* KinematicsQueryOptions::return_approximate_solution */
    uint64_t rng_seed; /* the reference RNG is unseeded; ours is Philox4x32-10 keyed by this */
} pik_params;

/* counters of the last pik_solve_batch on a solver */
typedef struct pik_stats {
    int64_t problems;
    int64_t solved;
    int64_t generation_launches;  /* memetic generation kernel launches */
    int64_t kernel_launches;      /* all kernels launched by the call */
    int64_t problem_generations;  /* sum over problems of generations executed */
    int64_t gd_steps;             /* GD step() executions */
    double device_ms;             /* CUDA-event time of the call's device work */
    double generation_ms;         /* CUDA-event time summed over the generation kernel launches */
} pik_stats;

typedef struct pik_robot pik_robot;
typedef struct pik_solver pik_solver;

int pik_version(void);
const char* pik_status_string(int status);

/* defaults of src/pick_ik_parameters.yaml */
void pik_params_default(pik_params* p);
/* the YAML validators (one_of / gt_eq) plus elite <= population and table limits */
int pik_params_validate(const pik_params* p);

/* Robot::from + chain flattening (src/robot.cpp:44-85; src/pick_ik_plugin.cpp:65-68): the serial chain model root ->
 * tip link, joints in chain order, one tip behind the last joint */
int pik_robot_create(const pik_joint_desc* joints, int32_t n_joints, pik_robot** out);
/*
 * The general form: a kinematic tree with n_tips tip links (src/pick_ik_plugin.cpp:60-68: one goal pose per tip frame,
 * src/goal.cpp:80-89,163-175; FK of every tip as src/fk_moveit.cpp:20-34 returns it), floating / planar joints and mimic
 * joints.
 *   parent    [n_joints]  the joint whose child link joint j hangs on, -1 = the model root; parents precede children
 *                         (NULL: a serial chain in joint order)
 *   tip_joint [n_tips]    the joint whose child link is tip t (n_tips <= PIK_MAX_TIPS)
 *   mimic_of  [n_joints]  the joint a mimic joint follows (value = mimic_factor * master + mimic_offset), -1 = none;
 *                         NULL: no mimic joints.  Mimic joints own no variable (src/robot.cpp:145-147).
 * Variables are numbered in joint order (floating: 7, planar: 3, revolute / prismatic: 1, fixed / mimic: 0).
 * Robots the serial-chain kernels cannot express run on the tree kernels (pik_robot_chain_signature: "tree").
 */
int pik_robot_create_tree(const pik_joint_desc* joints, int32_t n_joints, const int32_t* parent, const int32_t* tip_joint,
                          int32_t n_tips, const int32_t* mimic_of, const double* mimic_factor, const double* mimic_offset,
                          pik_robot** out);
int32_t pik_robot_num_tips(const pik_robot* robot);
void pik_robot_destroy(pik_robot* robot);
int32_t pik_robot_num_variables(const pik_robot* robot);
int pik_robot_get_variable(const pik_robot* robot, int32_t i, pik_variable* out);
/* name of the compiled chain signature the kernels will use for this robot (host; informational):
 * "generic", "all-z 7R, x-rotation origins (static)", "identity origins", "x-rotation origins", "y-rotation origins",
 * "tree" */
const char* pik_robot_chain_signature(const pik_robot* robot);
/* Robot::is_valid_configuration, src/robot.cpp:97-105 (host) */
int pik_robot_is_valid_configuration(const pik_robot* robot, const double* q);

/*
 * URDF -> chain table (host only; replaces the MoveIt RobotModel the reference reads in Robot::from,
 * src/robot.cpp:44-85, and get_active_variable_indices, src/robot.cpp:122-160): the serial chain
 * base_link -> tip_link of a URDF document as pik_joint_desc[] in root-to-tip order.  urdfdom / MoveIt
 * semantics: rpy -> normalised quaternion -> rotation, default axis (1,0,0) normalised, <limit> intersected
 * with <safety_controller> soft limits, continuous joints unbounded with the nominal range -pi..pi.
 * *n_joints receives the chain length (also when capacity == 0 and out == NULL: size query).
 * joint_names, link_names (optional): capacity * PIK_URDF_NAME_BYTES chars each, one NUL-terminated name per
 * joint / per child link of that joint (the last link is the tip).
 * PIK_E_INVALID_ROBOT: malformed XML / tip not below base; PIK_E_UNSUPPORTED: floating, planar or mimic
 * joints on the chain (pik_urdf_tree takes them); PIK_E_INVALID_ARGUMENT: a name longer than PIK_URDF_NAME_BYTES - 1
 * (names are never truncated); PIK_E_OUT_OF_MEMORY.  No exception crosses the boundary.
 */
#define PIK_URDF_NAME_BYTES 64
int pik_urdf_chain(const char* urdf_xml, const char* base_link, const char* tip_link, pik_joint_desc* out,
                   int32_t capacity, int32_t* n_joints, char* joint_names, char* link_names);
/*
 * The general form, for pik_robot_create_tree: the joints between base_link and every one of the n_tips tip links
 * (what get_active_variable_indices walks, src/robot.cpp:122-143), parents first and every branch contiguous
 * (depth-first, children in document order), with
 *   parent [capacity], tip_joint [n_tips], mimic_of / mimic_factor / mimic_offset [capacity]
 * as pik_robot_create_tree takes them (<mimic joint multiplier offset>; a mimic joint whose master is not among
 * these joints: PIK_E_UNSUPPORTED).  floating and planar joints are emitted with MoveIt's default bounds.
 * Size query as above (capacity == 0, out == NULL).
 */
int pik_urdf_tree(const char* urdf_xml, const char* base_link, const char* const* tip_links, int32_t n_tips,
                  pik_joint_desc* out, int32_t capacity, int32_t* n_joints, int32_t* parent, int32_t* tip_joint,
                  int32_t* mimic_of, double* mimic_factor, double* mimic_offset, char* joint_names, char* link_names);
/*
 * SRDF planning group -> base link and tip links (the part of the JointModelGroup the plugin needs,
 * src/pick_ik_plugin.cpp:42-49): the <chain base_link tip_link/> entries of <group name=GROUP>, sub-groups
 * (<group name=.../> inside a group) included.  base_link: PIK_URDF_NAME_BYTES chars; tip_links: capacity *
 * PIK_URDF_NAME_BYTES chars; *n_tips receives the number of chains (size query with capacity == 0).
 * PIK_E_UNSUPPORTED: the group lists joints / links only, or its chains do not share one base link.
 */
int pik_srdf_group(const char* srdf_xml, const char* group, char* base_link, char* tip_links, int32_t capacity,
                   int32_t* n_tips);

/* stream: a cudaStream_t (or NULL for a stream owned by the solver) */
int pik_solver_create(const pik_robot* robot, int32_t device, void* stream, pik_solver** out);
void pik_solver_destroy(pik_solver* solver);

/*
 * Batched replacement for the solve in PickIKPlugin::searchPositionIK
 * (src/pick_ik_plugin.cpp:162-217): ik_memetic (src/ik_memetic.cpp:285-373, one species) when
 * mode == global, ik_gradient (src/ik_gradient.cpp:96-139) when mode == local, for B independent
 * problems.
 *   goal_pose  [B][n_tips][7]  px py pz qw qx qy qz of every tip in the model frame (the goal_frames of
 *                      pick_ik_plugin.cpp:88-94; quaternion used un-normalised like tf2::fromMsg)
 *   seed       [B][n] (seed_stride = n) or [n] (seed_stride = 0): ik_seed_state
 *   solution   [B][n]  genes on success, the seed on failure (pick_ik_plugin.cpp:213,216)
 *   error_code [B]     PIK_SUCCESS / PIK_NO_IK_SOLUTION
 *   cost       [B]     best cost found (may be NULL)
 *   iterations [B]     generations (global) / GD iterations (local) executed (may be NULL)
 * first_problem_index offsets the RNG stream key so a shard of a larger batch reproduces the
 * unsharded result.  memory: PIK_MEM_HOST (host pointers; the copies to and from the device are
 * part of the call) or PIK_MEM_DEVICE (device pointers on the solver's device).  The call returns
 * when the results are complete.
 */
int pik_solve_batch(pik_solver* solver, const pik_params* params, int64_t B,
                    int64_t first_problem_index, const double* goal_pose, const double* seed,
                    int64_t seed_stride, double* solution, int32_t* error_code, double* cost,
                    int32_t* iterations, int32_t memory);

/*
 * The same solve without waiting for it: every copy and kernel of the call is enqueued on the solver's stream
 * (the kernels size themselves from device-side counters, so the host reads nothing back while the solve runs)
 * and the call returns.  The buffers must stay valid and the results may be read only after pik_solver_wait
 * (host outputs should be page-locked, pik_host_alloc, for the copies to be asynchronous).  One solve in flight
 * per solver: a second call returns PIK_E_BUSY; use one solver (own stream) per batch in flight -- solvers of
 * the same robot and parameters run concurrently on one device, e.g. the head of batch k + 1 in the shadow of
 * the latency-bound tail of batch k.  pik_solver_query: 1 = nothing in flight or finished, 0 = still running.
 */
int pik_solve_batch_async(pik_solver* solver, const pik_params* params, int64_t B,
                          int64_t first_problem_index, const double* goal_pose, const double* seed,
                          int64_t seed_stride, double* solution, int32_t* error_code, double* cost,
                          int32_t* iterations, int32_t memory);
int pik_solver_wait(pik_solver* solver);
int pik_solver_query(pik_solver* solver);

/*
 * Batched FK + cost + solution test: make_cost_fn (src/goal.cpp:188-203),
 * make_is_solution_test_fn (src/goal.cpp:163-186) and the tip frame of make_fk_fn
 * (src/fk_moveit.cpp:20-34) for B configurations q [B][n].  Outputs may be NULL.
 * goal_pose [B][n_tips][7]; tip_pose [B][n_tips][7] = px py pz qw qx qy qz of every tip.
 */
int pik_eval_cost(pik_solver* solver, const pik_params* params, int64_t B, const double* goal_pose,
                  const double* seed, int64_t seed_stride, const double* q, double* cost,
                  int32_t* is_solution, double* tip_pose, int32_t memory);

int pik_solver_synchronize(pik_solver* solver);

/*
 * Multi-GPU: one process (or host thread) per GPU, the pose batch sharded in contiguous ranges
 * (SURVEY.md 8e).  Problems are independent, so the only exchange is one NCCL all-gather of the packed
 * per-problem results at the end, issued on the solver's stream right behind the last kernel.
 *
 * pik_comm wraps an ncclComm_t that the library creates itself (NCCL is loaded with dlopen at the
 * first call: single-GPU users need no NCCL).  Rank 0 calls pik_comm_unique_id and hands the 128
 * bytes to the other ranks by any means (MPI, torch.distributed, a file); every rank then calls
 * pik_comm_create.
 */
typedef struct pik_comm pik_comm;
#define PIK_COMM_ID_BYTES 128
int pik_comm_unique_id(void* id_out /* PIK_COMM_ID_BYTES */);
int pik_comm_create(const void* id, int32_t n_ranks, int32_t rank, int32_t device, pik_comm** out);
void pik_comm_destroy(pik_comm* comm);
/* text of the last NCCL / loader error of this process (empty string if none) */
const char* pik_comm_last_error(void);
/*
 * pik_solve_batch on this rank's shard (B_local problems, the same on every rank; RNG streams keyed
 * by first_problem_index + local index = the global problem index), then the all-gather:
 *   gathered [n_ranks][B_local][n + 3] doubles on every rank: joints[n], cost, error_code,
 *   iterations (integers are exact in binary64), rank-major = global problem order when rank r owns
 *   [r * B_local, (r + 1) * B_local).
 * memory applies to goal_pose / seed / gathered alike.
 */
int pik_solve_batch_sharded(pik_solver* solver, pik_comm* comm, const pik_params* params, int64_t B_local,
                            int64_t first_problem_index, const double* goal_pose, const double* seed,
                            int64_t seed_stride, double* gathered, int32_t memory);
/*
 * The general form.  counts [n_ranks]: problems of every rank's shard (NULL: B_local on every rank; otherwise
 * counts[rank] == B_local); shards may be uneven or empty (pick_ik_b200/sharding.py::shard_range).  root: the one
 * rank that receives the block, or -1 for every rank; only receiving ranks need `gathered`
 * ([sum counts][n + 3], rank order) and only they pay for a device-to-host copy of it (PIK_MEM_HOST).  Even
 * shards to every rank travel as one ncclAllGather, everything else as grouped ncclSend / ncclRecv.  Every rank
 * takes part in the exchange even if its own solve fails (its rows arrive as NaN and its call returns the
 * error), so a failing rank does not leave the others waiting.
 */
int pik_solve_batch_gather(pik_solver* solver, pik_comm* comm, const pik_params* params, int64_t B_local,
                           int64_t first_problem_index, const double* goal_pose, const double* seed,
                           int64_t seed_stride, const int64_t* counts, int32_t root, double* gathered, int32_t memory);

/*
 * Synthetic workload generator of the benchmarks: q [B][n] (host), configuration b drawn uniformly
 * within the limits of every variable (unbounded: -pi .. pi) from the Philox4x32-10 stream (gen_seed,
 * first_problem_index + b).  Host arithmetic only: the same bits from C, from the Python harness and from the CPU
 * oracle (orc_random_configuration).  Targets are the FK of these configurations (pik_eval_cost's tip_pose).
 */
int pik_random_configurations(const pik_robot* robot, uint64_t gen_seed, int64_t first_problem_index, int64_t B,
                              double* q);

/* number of CUDA devices visible (0 if none) */
int pik_device_count(void);
/* page-locked host buffers for PIK_MEM_HOST calls (optional; any host memory is accepted) */
int pik_host_alloc(void** out, size_t bytes);
void pik_host_free(void* p);
/* FP64 FMA throughput of the solver's device in TFLOP/s, measured with independent DFMA chains on
 * every SM; bench.py reports the FP64 roofline fraction against it. */
int pik_measure_fp64_peak(pik_solver* solver, double* tflops);
int pik_solver_get_stats(pik_solver* solver, pik_stats* out);
/* last CUDA error text of this solver (empty string if none) */
const char* pik_solver_last_error(const pik_solver* solver);

#ifdef __cplusplus
}
#endif
#endif
