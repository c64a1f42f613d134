// pik_types.h -- plain structs shared by the host runtime (pik_api.cu) and the kernels (pik_kernels.cu).
#pragma once

#include <stdint.h>

namespace pik {

constexpr int kMaxVars = 16;
constexpr int kMaxElites = 32;       // the elites of one problem live in one warp
constexpr int kMaxPopulation = 1024;

enum StepKind : int { kRevX = 0, kRevY = 1, kRevZ = 2, kRevGeneral = 3, kPrismatic = 4 };

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
    uint32_t seed_lo, seed_hi;
    uint32_t round_key[20];  // Philox4x32-10 key schedule of (seed_lo, seed_hi): k0_r, k1_r for round r
    int lockstep;  // block barrier per GD step in throughput mode (PIK_NO_LOCKSTEP disables)
    int debug;  // PIK_DEBUG_PHASES: warp 0 of CTA 0 prints the cycle count of each phase of a generation
};

// status codes in meta[b].status
enum : int { kActive = 0, kSolved = 1, kFailed = 2 };

// Per-problem solver state that is not an individual (MemeticIk members, ik_memetic.hpp:47-85)
struct ProblemMeta {
    int has_prev;    // previous_fitness_.has_value()
    int iter;        // generations completed
    int status;
    int init_epoch;  // number of initPopulation calls so far (RNG stream epoch)
};

// Per-solve device buffers (all device pointers).
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
    int32_t* counters;        // [2] sizes of the two lists
    unsigned long long* stats;  // [4]: problem_generations, gd_steps, solved, finished
    int64_t B;
    int64_t first_problem_index;
};

}  // namespace pik
