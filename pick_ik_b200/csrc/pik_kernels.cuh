// pik_kernels.cuh -- kernel launchers and device-buffer layouts shared by pik_kernels.cu and
// pik_api.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pik_device.cuh"

namespace pik {

constexpr int kMaxElites = 32;
constexpr int kMaxPopulation = 1024;

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
// Individuals are never moved: order[b][i] is the slot of population_[i] (the sort of
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
    uint16_t* order;          // [B][P]
    double* hdr;              // [B][n+2]: best genes[n], best fitness, previous fitness
    ProblemMeta* meta;        // [B]
    int32_t* active;          // [2][B] compacted lists of active problems
    int32_t* counters;        // [2] sizes of the two lists
    unsigned long long* stats;  // [4]: problem_generations, gd_steps, solved, finished
    int64_t B;
    int64_t first_problem_index;
};

// Launch shape of the memetic kernels: T threads per CTA, G = T / E problems per CTA.
struct MemeticShape {
    int threads;
    int group;
    size_t smem;
};
MemeticShape memetic_shape(int n, int P, int E);
size_t gd_local_smem_bytes(int n, int threads);

cudaError_t launch_eval_cost(cudaStream_t stream, const DevRobot& robot, const DevParams& pr, int64_t B,
                             const double* goal_pose, const double* seed, int64_t seed_stride, const double* q,
                             double* cost, int32_t* is_solution, double* tip_pose);
cudaError_t launch_gd_local(cudaStream_t stream, const DevRobot& robot, const DevParams& pr, const SolveBuffers& sb);
cudaError_t launch_memetic_init(cudaStream_t stream, const DevRobot& robot, const DevParams& pr,
                                const SolveBuffers& sb);
// One launch advances every problem of active list `list_in` by one generation and appends the
// still-active ones to the other list.  n_active bounds the grid.
cudaError_t launch_memetic_generation(cudaStream_t stream, const DevRobot& robot, const DevParams& pr,
                                      const SolveBuffers& sb, int list_in, int64_t n_active);
// FP64 FMA throughput microbenchmark (roofline denominator for the FP64 bound): returns FLOP count
cudaError_t launch_fp64_peak(cudaStream_t stream, double* sink, int blocks, int threads, int iters);
cudaError_t configure_kernels();

}  // namespace pik
