// pik_types.h -- plain structs shared by the host runtime (pik_api.cu) and the kernels (pik_kernels.cu).
#pragma once

#include <stdint.h>

namespace pik {

constexpr int kMaxVars = 16;          // variables, and moving joints (steps)
constexpr int kMaxTips = 4;
constexpr int kMaxSavedFrames = 4;    // frames of branch points kept while a tree is walked
constexpr int kMaxElites = 32;       // the elites of one problem live in one warp
constexpr int kMaxPopulation = 1024;
constexpr int kSmDenseSize = 256;    // %smid values are below this (a power of two)

// kPrisX/Y/Z: prismatic along +-x / y / z (sign in DevRobot::sign): t += column * (sign * q), which is what the general
// form computes up to the sign of exact zeros (fma(x, +-0, t) == t)
// kFloating (7 variables: x y z, quaternion x y z w) and kPlanar (x y theta): src/forward_kinematics.cpp:64-79; tree
// kernels only
enum StepKind : int { kRevX = 0, kRevY = 1, kRevZ = 2, kRevGeneral = 3, kPrismatic = 4, kPrisX = 5, kPrisY = 6, kPrisZ = 7,
                      kFloating = 8, kPlanar = 9 };

// Sparsity pattern of a constant rotation (URDF origins are mostly rotations about one coordinate axis, often
// by multiples of pi/2): entries that are exactly 0 or 1 need no arithmetic -- x * 1 == x and fma(x, 0, y) == y
// exactly (up to the sign of an exact zero, which nothing on this path can observe) -- so a kernel compiled
// for a chain signature whose origins all share a pattern skips them.  The skipped terms are the ones a full
// product would add as exact zeros: results are bit-identical to the full left-to-right product of the
// oracle.  (A run-time dispatch on the pattern was measured too: its branches cost what the skipped
// arithmetic saves.)
enum OriginClass : int { kOrgGeneral = 0, kOrgIdentity = 1, kOrgRotX = 2, kOrgRotY = 3, kOrgRotZ = 4 };

// Flattened chain + variable table (constant memory on the device).  Fixed joints are folded into the
// constant origin that precedes each moving joint, or into the tip transform.
struct DevRobot {
    int n;
    int has_tip;
    int any_unbounded;  // some variable has no position bounds (URDF continuous joint)
    int pad_;
    int kind[kMaxVars];
    int bounded[kMaxVars];
    double sign[kMaxVars];
    double R[kMaxVars + 1][9];  // folded constant origin preceding each moving joint, row-major; [n] = the tip transform
    double t[kMaxVars + 1][3];
    int ocls[kMaxVars + 1];     // OriginClass of R[j] (most specific pattern; used by select_spec)
    double axis[kMaxVars][3];
    double axis_sq[kMaxVars][6];  // xx yy zz xy xz yz
    double tip_R[9];
    double tip_t[3];
    double vmin[kMaxVars], vmax[kMaxVars], vmid[kMaxVars], vhalf[kMaxVars], vfac[kMaxVars];
    // ---- kinematic trees, several tips, multi-variable and mimic joints (is_tree: served by the tree kernels only).
    // Step k (a moving joint, parents before children): frame = frame of step parent[k] (-1: the model root) * folded
    // origin R[k], t[k] * joint motion on variables var0[k].. (value * mimic_factor + mimic_offset for a mimic joint).
    // A serial single-tip chain of one-variable joints has parent[k] = k - 1, var0[k] = k and its tip in R[n], t[n].
    int is_tree;
    int n_steps;
    int n_tips;
    int pad2_;
    int parent[kMaxVars];
    int var0[kMaxVars];
    int load_slot[kMaxVars];  // where the walk finds the parent frame: -1 the previous step, -2 the model root, >= 0 a saved frame
    int save_slot[kMaxVars];  // >= 0: the frame of this step is kept in that slot for a later branch
    double mimic_factor[kMaxVars], mimic_offset[kMaxVars];
    int tip_step[kMaxTips];   // step the tip link hangs on (-1: the model root)
    int tip_has[kMaxTips];    // fixed transform between that step and the tip link
    double tips_R[kMaxTips][9];
    double tips_t[kMaxTips][3];
};

// Solver parameters as the kernels see them (pick_ik_plugin.cpp:97-129,166-196 applied).
struct DevParams {
    double step_size, min_cost_delta;
    double position_threshold, orientation_threshold, cost_threshold_sq;
    double position_scale, rotation_scale;
    double w2_center, w2_avoid, w2_mindisp;  // weight^2, 0 = goal absent
    double wipeout_tol;
    int gd_max_iters;  // local: gd_max_iters; global: memetic_gd_max_iters
    int stop_on_valid, approx;
    int P, E, max_generations;
    int lockstep;  // block barrier per GD step in throughput mode (PIK_NO_LOCKSTEP disables)
    int debug;  // PIK_DEBUG_PHASES: warp 0 of CTA 0 prints the cycle count of each phase of a generation
    // Launch policy of the generation kernels, evaluated ON THE DEVICE from the size of the active list (the host
    // enqueues every generation of a solve up front and never reads a count back):
    int sm_count;
    int lanes_max;               // memetic_max_lanes_per_elite(E)
    long long wide_capacity_lanes;  // the lane mapping doubles L while n_active * E * 2L stays below this
    int wide_units_max;          // warps a wide launch provides (grid * warps per CTA)
    int persistent_units_max;    // a wide launch with one problem per warp and at most this many problems keeps
                                 // every problem for all its remaining generations (0: never)
    int wave_ctas;               // the resident CTAs of a throughput launch (its grid, when the batch fills them)
    int defer_launches;          // throughput launches 0 .. defer_launches - 1 process whole waves of resident CTAs only
                                 // and pass the rest of their list on untouched (0: every launch processes its list)
    unsigned short sm_dense[kSmDenseSize];  // %smid -> 0 .. sm_count - 1 (%smid has holes where SMs are fused off)
};

#ifdef __CUDACC__
#define PIK_HD __host__ __device__
#else
#define PIK_HD
#endif

// Lanes per elite (L) for a generation over n_active problems.  A lone warp of this code issues ~0.2 instructions
// per cycle (dependent FP64 chains: 8-cycle DFMA latency), so an SM sub-partition needs 4-5 resident warps to stay
// busy and a launch with fewer is latency-bound.  Throughput mode (L = 1: a whole GD instance per lane, 32 / E
// problems per warp) executes the fewest instructions per problem but has the longest serial path per generation;
// as the batch drains, the warp's lanes are spread over the evaluations of each GD step instead: the largest L
// whose launch still fits the lanes the wide flavour keeps resident (capacity_lanes) and the warps its grid
// provides (wide_units_max).  Evaluated on the device by every generation launch and on the host when a solve is
// planned (monotone: fewer problems never give a smaller L).
PIK_HD inline int wide_problems_per_warp_max(int E) { return 32 / (2 * E) > 1 ? 32 / (2 * E) : 1; }
PIK_HD inline int lanes_for(long long n_active, int E, int lanes_max, long long capacity_lanes, int wide_units_max) {
    int L = 1;
    while (L * 2 <= lanes_max && n_active * E * (L * 2) <= capacity_lanes) {
        const int pw = 32 / (E * L * 2);
        if ((n_active + pw - 1) / pw > wide_units_max) break;
        L *= 2;
    }
    return L;
}

// status codes in meta[b].status; kSolvedTerminated: a species that returned its best individual because another
// species of the problem had returned a value (src/ik_memetic.cpp:264-282)
enum : int { kActive = 0, kSolved = 1, kFailed = 2, kSolvedTerminated = 3, kFailedTerminated = 4 };

// Per-problem solver state that is not an individual (MemeticIk members, ik_memetic.hpp:47-85)
struct ProblemMeta {
    int has_prev;    // previous_fitness_.has_value()
    int iter;        // generations completed
    int status;
    int init_epoch;  // number of initPopulation calls so far (RNG stream epoch)
};

// Per-solve device buffers (all device pointers).  Every per-problem solver buffer is indexed by the SUB-problem
// sp (one per species of a problem; n_species == 1: sp == problem); goal_pose and seed by the problem sp / n_species.
//
// Population layout (Individual, ik_memetic.hpp:19-24, as structure-of-arrays): for problem b and
// buffer s in {0,1}: pop[((s * B + b) * (2n+2) + row) * P + slot], rows 0..n-1 genes, n..2n-1
// gradient, 2n fitness, 2n+1 extinction.  Generation g reads buffer g & 1 and writes buffer
// (g + 1) & 1, so the previous occupants of the child slots stay readable while children are
// produced speculatively (src/ik_memetic.cpp:181-188 seeds a random individual from them).
// Individuals are never moved: order[s][b][i] is the slot of population_[i] (the sort of
// src/ik_memetic.cpp:200-203 permutes this index row instead of the individuals).
struct SolveBuffers {
    const double* goal_pose;  // [B][7]
    const double* seed;       // [B][n] or [n]
    int64_t seed_stride;      // n or 0
    double* solution;         // [B][n]
    int32_t* error_code;      // [B]
    double* cost;             // [B] or null
    int32_t* iterations;      // [B] or null
    double* pop;              // [2][B][2n+2][P]
    uint16_t* order;          // [2][B][P]
    double* hdr;              // [B][n+2]: best genes[n], best fitness, previous fitness
    ProblemMeta* meta;        // [B]
    int32_t* active;          // [2][B] compacted lists of active problems
    int32_t* counters;        // [max_generations + 2]: counters[g] = problems active in generation g
    int32_t* sched;           // [max_generations + 1][sm_count + 2]: per generation launch, the heads of the per-SM unit
                              // queues, the number of CTAs that have left, the units taken (memetic_generation_kernel)
    int32_t* group_term;      // [B / n_species] or null: generation + 1 at which a species of the problem returned a
                              // value (the `terminate` flag of ik_memetic, src/ik_memetic.cpp:334-346), 0 = none
    unsigned long long* stats;  // [8]: problem_generations, gd_steps, solved, finished, problems with a value (species pick)
    int64_t B;                // sub-problems = IK problems * n_species
    int64_t first_problem_index;
    int32_t n_species;        // MemeticIkParams::num_threads: sub-problem sp = problem sp / n_species, species sp % n_species
    int32_t stop_on_first;    // MemeticIkParams::stop_on_first_soln
    int32_t sm_rotation;      // first SM queue of this sub-batch's unit dealing (memetic_generation_kernel)
    // Philox4x32-10 key schedule of rng_seed: k0_r, k1_r for round r.  Per call (kernel parameter space, i.e.
    // constant-bank operands like the tables above), so that calls that differ only in the seed share one
    // parameter upload.
    uint32_t round_key[20];
};

}  // namespace pik
