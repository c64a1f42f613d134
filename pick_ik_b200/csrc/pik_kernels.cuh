// pik_kernels.cuh -- kernel launchers shared by pik_kernels.cu and pik_api.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pik_types.h"

namespace pik {

// Launch shape of the memetic kernels.  Every warp is autonomous: it owns `problems_per_warp` problems
// and gives each of their E elites `lanes_per_elite` lanes (1 = throughput mode: a whole GD instance per
// lane; > 1 = wide mode for a draining batch: the evaluations of one GD step run on different lanes).
struct MemeticShape {
    int threads;            // per CTA
    int warps;              // per CTA
    int lanes_per_elite;    // L
    int problems_per_warp;  // PW = 32 / (E * L)
    size_t smem;            // per CTA
};
MemeticShape memetic_shape(int n, int n_tips, int P, int E, int lanes_per_elite);
int memetic_max_lanes_per_elite(int E);
size_t gd_local_smem_bytes(int n, int n_tips);

// Copies the robot table and the solver parameters to constant memory (ordered on `stream`).
cudaError_t upload_constants(cudaStream_t stream, const DevRobot& robot, const DevParams& pr);

cudaError_t launch_eval_cost(cudaStream_t stream, int n, int64_t B, const double* goal_pose, const double* seed,
                             int64_t seed_stride, const double* q, double* cost, int32_t* is_solution,
                             double* tip_pose);
// Chain signatures the kernels are compiled for; select_spec() picks the most specific one a robot table
// matches.  kSpecAllZ7: n = 7, every joint about z, x-rotation origins, z-rotation tool frame (Franka Panda and
// every DH-style 7R arm): n is a constant, the kind dispatch folds away.  kSpecOrg*: any n and kinds, but every
// joint origin (from the second joint on) has the named sparsity pattern (OriginClass) -- identity (URDFs whose
// origins only translate: Fetch), rotation about x, rotation about y (the UR family) -- so the chain walk
// skips the terms that are exact zeros.  kSpecGeneric: anything.
// kSpecTree: kinematic trees, several tips, floating / planar / mimic joints (DevRobot::is_tree).
enum : int { kSpecGeneric = 0, kSpecAllZ7 = 1, kSpecOrgIdentity = 2, kSpecOrgRotX = 3, kSpecOrgRotY = 4, kSpecTree = 5, kSpecCount = 6 };
int select_spec(const DevRobot& robot);

cudaError_t launch_gd_local(cudaStream_t stream, int spec, int n, int n_tips, const SolveBuffers& sb, int sm_count);
cudaError_t launch_memetic_init(cudaStream_t stream, int spec, int n, int n_tips, int P, int E, const SolveBuffers& sb);
// The launches of a global-mode solve over n_sub sub-problems, planned once per call (nothing is read back from the
// device while the solve runs).  Per generation: one launch of the throughput flavour (blocks_t) and one of the
// wide flavour (blocks_w); each reads the size of the generation's active list and returns at once unless the lane
// mapping that size calls for is its own (memetic_generation_kernel).  use_throughput / use_wide: a flavour that
// can never run for this batch size is not enqueued; first_launch_runs_all: the batch is so small that the first
// wide launch keeps every problem for all its generations.
struct GenerationPlan {
    unsigned blocks_t, blocks_w;
    int threads_t, threads_w;
    size_t smem_t, smem_w;
    int lanes_max, wide_units_max, persistent_units_max, wave_ctas;
    long long wide_capacity_lanes;
    bool use_throughput, use_wide, first_launch_runs_all;
};
GenerationPlan plan_generations(int n, int n_tips, int P, int E, int64_t n_sub, int sm_count, long long wide_warps_per_sm,
                                bool allow_persistent);
cudaError_t launch_memetic_generation(cudaStream_t stream, int spec, const GenerationPlan& plan, const SolveBuffers& sb,
                                      int gen, bool wide);
// ik_memetic's pick over the species of every problem (src/ik_memetic.cpp:334-370): sb holds the per-species results
cudaError_t launch_species_pick(cudaStream_t stream, const SolveBuffers& sb, int n, int64_t n_problems, double* solution,
                                int32_t* error_code, double* cost, int32_t* iterations);
// packed [B][n + 3] = joints, cost, error_code, iterations (the all-gather payload of the sharded solve)
cudaError_t launch_pack_results(cudaStream_t stream, int64_t B, int n, const double* solution, const double* cost,
                                const int32_t* error_code, const int32_t* iterations, double* packed);
// FP64 FMA throughput microbenchmark (roofline denominator for the FP64 bound)
cudaError_t launch_fp64_peak(cudaStream_t stream, double* sink, int blocks, int threads, int iters);
cudaError_t configure_kernels();
// dense [kSmDenseSize]: %smid -> 0 .. sm_count - 1 on the current device (synchronises `stream`)
cudaError_t discover_sm_ids(cudaStream_t stream, int sm_count, unsigned short* dense);

}  // namespace pik
