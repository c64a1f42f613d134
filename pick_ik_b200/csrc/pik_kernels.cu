// pik_kernels.cu -- sm_100a kernels of the batched IK engine.
//
//   eval_cost_kernel           make_cost_fn / make_is_solution_test_fn / FK    (src/goal.cpp:163-203, src/fk_moveit.cpp:20-34)
//   gd_local_kernel            ik_gradient                                      (src/ik_gradient.cpp:96-139)
//   memetic_init_kernel        ik_memetic early-out + MemeticIk::from + initPopulation (src/ik_memetic.cpp:18-41,93-117,285-296)
//   memetic_generation_kernel  one iteration of ik_memetic_impl's loop          (src/ik_memetic.cpp:228-269):
//                              gradientDescent on the elites, reproduce, sortPopulation, solution test, checkWipeout
//
// Mapping.  The unit of work is one cost evaluation (FK chain walk + pose/goal costs): a serial FP64
// dependency chain of ~10^3 operations.  Every WARP is autonomous -- it owns PW problems end to end (the only
// block barriers are the optional lockstep ones of the throughput mapping, which exchange no data):
//   * elite local search: one GD instance per lane (L = 1, throughput mapping), every lane carrying whole
//     finite-difference + line-search steps of its own elite as frame pairs (gd_step_compact); or, once the
//     batch has drained so far that a launch is bound by the serial path of a generation, L lanes per elite
//     (wide mapping, gd_elite_wide): the finite-difference pairs and the accepted point of the previous step
//     in one round, the two line-search points in a second;
//   * reproduction: children are made in windows, one child per lane; the sequential mating-pool semantics of
//     the reference are kept by committing a window only up to the first child that removes a parent and
//     restarting after it with the shrunken pool.  One problem per warp: windows of 32; PW > 1 problems per
//     warp: side by side, 32 / PW lanes each, so that a removal throws away at most 32 / PW - 1 children;
//   * the sort permutes a slot-index row, never the individuals (the E best and the worst: a running list kept
//     while children are committed, or warp reductions over the fitness array; a full sort only for robots with
//     unbounded variables, which need the whole order).
// Scheduling.  Grids of resident CTAs take warp-units from one queue per SM; a launch is a scheduling step, not a
// generation: problems carry their own generation count, and a throughput launch processes whole waves of its
// resident CTAs only, passing the rest of its list on (see memetic_generation_kernel).
// All arithmetic is binary64 with --fmad=false (see pik_device.cuh) and is bit-identical to
// oracle/pik_oracle.c whatever the mapping.
#include "pik_kernels.cuh"

#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <cstdlib>

#include "pik_device.cuh"
#include "pik_host_robot.h"

namespace pik {

namespace {

constexpr int kWarpsPerBlock = 4;       // every kernel but the throughput-mode generation launches
constexpr int kWarpsPerBlockBulk = 16;  // throughput-mode generation launches: ONE CTA per SM, its 16 warps in step (see
                                        // `lockstep`; two CTAs of 8 warps: +3 % per wave, four of 4: +11 %)
constexpr int kThreads = 32 * kWarpsPerBlock;
constexpr unsigned kFull = 0xffffffffu;

// Per-warp shared memory.  Columns (stride kS = 32): q, g, best [n] each, sc [2n].
struct WarpSmem {
    double* q;
    double* g;
    double* best;
    double* sc;
    double* cs;    // [2n][16] finite-difference / line-search costs of the lane-parallel GD step: columns 16..31 of
                   // the sc rows, which a lane-parallel step (at most 16 elite groups per warp) never uses
    double* fit;   // [P] fitness by population position of the problem being reproduced / sorted
    double* goal;  // [PW][7]
    double* efit;  // [32] elite fitness by column
    double* eext;  // [32] elite extinction by column
    double* f0s;   // [32] best_curr fitness per problem
    int* pool;     // [32]
    int* top;      // [33] positions of rank 0..E-1 and (at [E]) rank P-1
    int* pidx;     // [32] problem index per problem slot, or -1
    int* flag;     // [32]
    int* ctl;      // [4]
    uint16_t* perm;  // [next power of two >= P] positions being sorted (robots with unbounded variables)
    // side-by-side reproduce (robots without unbounded variables; aliases perm):
    double* rtf;   // [64] per problem slot: fitness of the E best so far, then of the worst
    int* rti;      // [64] ... and their positions
    int* grp;      // [64] per problem slot: mating pool size, [32 + slot]: next child
};

// fit aliases the sc rows (dead once the elite searches are done) when the population fits there
__host__ __device__ inline bool fit_in_sc(int n, int P) { return P <= 2 * n * kS; }

__host__ __device__ inline int pow2_at_least(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

__host__ __device__ inline size_t warp_smem_bytes(int n, int P, int PW, int T) {
    size_t d = (size_t)5 * n * kS + (fit_in_sc(n, P) ? 0 : (size_t)P) + (size_t)PW * 7 * T + 3 * 32;
    size_t i = 32 + 33 + 32 + 32 + 4;
    size_t tail = (size_t)pow2_at_least(P) * 2;  // perm, or rtf + rti + grp
    if (tail < 64 * 8 + 128 * 4) tail = 64 * 8 + 128 * 4;
    return ((d * 8 + i * 4 + 4 + tail) + 15) & ~size_t(15);
}

__device__ __forceinline__ WarpSmem carve_warp(unsigned char* base, int n, int P, int PW, int T) {
    WarpSmem W;
    double* d = reinterpret_cast<double*>(base);
    W.q = d; d += (size_t)n * kS;
    W.g = d; d += (size_t)n * kS;
    W.best = d; d += (size_t)n * kS;
    W.sc = d; d += (size_t)2 * n * kS;
    W.cs = W.sc + 16;
    if (fit_in_sc(n, P)) {
        W.fit = W.sc;
    } else {
        W.fit = d; d += P;
    }
    W.goal = d; d += PW * 7 * T;  // [PW][T][7]: the goal frame of every tip
    W.efit = d; d += 32;
    W.eext = d; d += 32;
    W.f0s = d; d += 32;
    int* ip = reinterpret_cast<int*>(d);
    W.pool = ip; ip += 32;
    W.top = ip; ip += 33;
    W.pidx = ip; ip += 32;
    W.flag = ip; ip += 32;
    W.ctl = ip; ip += 4;
    if ((reinterpret_cast<uintptr_t>(ip) & 7) != 0) ++ip;
    W.perm = reinterpret_cast<uint16_t*>(ip);
    W.rtf = reinterpret_cast<double*>(ip);
    W.rti = ip + 128;
    W.grp = ip + 192;
    return W;
}

__device__ __forceinline__ double* pop_ptr(const SolveBuffers& sb, int buf, int64_t b, int n, int P) {
    return sb.pop + (((size_t)buf * (size_t)sb.B + (size_t)b) * (size_t)(2 * n + 2)) * (size_t)P;
}

__device__ __forceinline__ uint16_t* order_ptr(const SolveBuffers& sb, int buf, int64_t b, int P) {
    return sb.order + ((size_t)buf * (size_t)sb.B + (size_t)b) * (size_t)P;
}

// sub-problem sp -> IK problem (goal, seed, RNG problem word) and species (RNG individual word)
__device__ __forceinline__ int problem_of(const SolveBuffers& sb, int64_t sp) {
    return sb.n_species == 1 ? (int)sp : (int)(sp / sb.n_species);
}
__device__ __forceinline__ uint32_t species_of(const SolveBuffers& sb, int64_t sp) {
    return sb.n_species == 1 ? 0u : (uint32_t)(sp % sb.n_species);
}

// Plugin output mapping (src/pick_ik_plugin.cpp:209-217): genes on success, the seed otherwise.
__device__ __forceinline__ void write_result(const SolveBuffers& sb, int n, int64_t b, bool found, const double* genes,
                                             int genes_stride, const double* seed, double cost, int iterations) {
    sb.error_code[b] = found ? 1 : -31;
    double* out = sb.solution + (size_t)b * n;
    for (int j = 0; j < n; ++j) out[j] = found ? genes[j * genes_stride] : seed[j];
    if (sb.cost) sb.cost[b] = cost;
    if (sb.iterations) sb.iterations[b] = iterations;
}

// -----------------------------------------------------------------------------------------------
// Batched FK + cost + solution test, one configuration per lane
// -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) eval_cost_kernel(int64_t B, const double* __restrict__ goal_pose,
                                                             const double* __restrict__ seed, int64_t seed_stride,
                                                             const double* __restrict__ q, double* cost,
                                                             int32_t* is_solution, double* tip_pose) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = c_rb.n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* col = reinterpret_cast<double*>(smem_raw) + (size_t)warp * (n + 7) * kS + lane;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int T = c_rb.n_tips;
    double g7[7 * kMaxTips];
    goals_from_poses(goal_pose + (size_t)(7 * T) * (size_t)b, g7);
    const double* sd = seed + b * seed_stride;
    for (int j = 0; j < n; ++j) col[j * kS] = q[b * n + j];
    double aux[5];
    if (c_rb.is_tree) {
        const double c = eval_tree(col, nullptr, kViewPlain, -1, 0.0, g7, sd, aux, tip_pose ? tip_pose + (size_t)(7 * T) * (size_t)b : nullptr);
        if (cost) cost[b] = c;
        if (is_solution) is_solution[b] = solution_from_aux(aux) ? 1 : 0;
        return;
    }
    // tip frame: walk here (the tip pose is an output of this kernel only)
    const ConfigView cv{col, nullptr, kViewPlain, -1, 0.0};
    Frame F;
    frame_load_origin(F, 0);
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
        double sj, cj;
        joint_sincos<GenericSpec>(j, cv.at(j), sj, cj);
        walk_joint<GenericSpec>(F, j, j > 0, cv.at(j), sj, cj);
    }
    if (c_rb.has_tip) frame_mul_const(F, c_rb.tip_R, c_rb.tip_t);
    const double c = total_cost(g7, F, cv, sd, aux);
    if (cost) cost[b] = c;
    if (is_solution) is_solution[b] = solution_from_aux(aux) ? 1 : 0;
    if (tip_pose) {
        double* tp = tip_pose + 7 * b;
        tp[0] = F.t[0]; tp[1] = F.t[1]; tp[2] = F.t[2];
        matrix_to_quat(F.r, tp[3], tp[4], tp[5], tp[6]);
    }
}

// -----------------------------------------------------------------------------------------------
// ik_gradient (src/ik_gradient.cpp:96-139), one problem per lane, whole loop on chip
// -----------------------------------------------------------------------------------------------
// Problems converge after very different numbers of iterations (a few ... gd_max_iters), so a lane that has finished
// its problem takes the next one from a device-wide counter instead of idling until the slowest lane of its warp is
// done: the warp keeps all its lanes in the same step() code whatever problem each of them is on.
template <class S>
__global__ void __launch_bounds__(kThreads) gd_local_kernel(const __grid_constant__ SolveBuffers sb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = c_rb.n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = c_rb.n_tips;
    double* base = reinterpret_cast<double*>(smem_raw) + (size_t)warp * (5 * n + 12 + 7 * T) * kS;
    GdState st{base + lane, base + (size_t)n * kS + lane, base + (size_t)2 * n * kS + lane,
               base + (size_t)3 * n * kS + lane, base + (size_t)5 * n * kS + lane, 0.0, 0.0};
    double* g7 = base + (size_t)(5 * n + 12) * kS + lane * 7 * T;  // 7 contiguous doubles per lane and tip
    unsigned long long* next_problem = reinterpret_cast<unsigned long long*>(sb.stats + 5);
    int64_t b = -1;
    const double* sd = nullptr;
    bool running = false;   // this lane is inside the iteration loop of problem b
    bool drained = false;   // the counter has run past the batch
    int iters = 0;
    double previous_cost = 0.0;
    unsigned long long steps = 0, solved = 0, finished = 0;
    double aux[5];
    for (;;) {
        // lanes without a problem take the next ones (one atomic per warp)
        const unsigned want = __ballot_sync(kFull, !running && !drained);
        if (want) {
            unsigned long long first = 0;
            if (lane == __ffs(want) - 1) first = atomicAdd(next_problem, (unsigned long long)__popc(want));
            first = __shfl_sync(kFull, first, __ffs(want) - 1);
            if (!running && !drained) {
                b = (int64_t)first + __popc(want & ((1u << lane) - 1u));
                if (b >= sb.B) {
                    drained = true;
                } else {
                    sd = sb.seed + b * sb.seed_stride;
                    for (int j = 0; j < n; ++j) {
                        st.q[j * kS] = sd[j];
                        st.best[j * kS] = sd[j];
                        st.g[j * kS] = 0.0;
                    }
                    goals_from_poses(sb.goal_pose + (size_t)(7 * c_rb.n_tips) * (size_t)b, g7);
                    const double c0 = eval_chain<S>(st.q, nullptr, kViewPlain, -1, 0.0, nullptr, st.sc, g7, sd, aux);
                    if (c_pr.stop_on_valid && solution_from_aux(aux)) {  // ik_gradient.cpp:102-104
                        write_result(sb, n, b, true, st.best, kS, sd, c0, 0);
                        ++solved;
                        ++finished;
                    } else {
                        st.local_cost = st.best_cost = c0;  // GradientIk::from
                        previous_cost = 0.0;
                        iters = 0;
                        running = true;
                    }
                }
            }
        }
        if (!__any_sync(kFull, running)) {
            if (__all_sync(kFull, drained)) break;
            continue;
        }
        if (running) {
            bool found = false, stop = iters >= c_pr.gd_max_iters;
            if (!stop) {
                const bool improved = gd_step<S, true>(st, g7, sd, aux);
                ++steps;
                // best == local when improved, so aux describes best (ik_gradient.cpp:117-121)
                if (improved && c_pr.stop_on_valid && solution_from_aux(aux)) {
                    found = true;
                    stop = true;
                } else if (fabs(st.local_cost - previous_cost) <= c_pr.min_cost_delta) {  // ik_gradient.cpp:123-125
                    stop = true;
                } else {
                    previous_cost = st.local_cost;
                    ++iters;
                    stop = iters >= c_pr.gd_max_iters;
                }
            }
            if (stop) {
                if (!found && !c_pr.stop_on_valid) {  // ik_gradient.cpp:130-132
                    eval_chain<S>(st.best, nullptr, kViewPlain, -1, 0.0, nullptr, nullptr, g7, sd, aux);
                    found = solution_from_aux(aux);
                }
                if (!found && c_pr.approx) found = true;  // ik_gradient.cpp:134-136
                write_result(sb, n, b, found, st.best, kS, sd, st.best_cost, iters);
                if (found) ++solved;
                ++finished;
                running = false;
            }
        }
    }
    if (sb.stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            steps += __shfl_xor_sync(kFull, steps, o);
            solved += __shfl_xor_sync(kFull, solved, o);
            finished += __shfl_xor_sync(kFull, finished, o);
        }
        if (lane == 0) {
            atomicAdd(&sb.stats[1], steps);
            atomicAdd(&sb.stats[2], solved);
            atomicAdd(&sb.stats[3], finished);
        }
    }
}

// -----------------------------------------------------------------------------------------------
// initPopulation (src/ik_memetic.cpp:93-117) for every problem slot k of the warp with flag[k] != 0, from
// the best genes in hdr[b], into the population buffer the problem's next generation reads.  Elite 0
// keeps the genes; elites 1..E-1 are random valid configurations seeded from them; children are copies
// whose fitness equals the best fitness (identical genes, deterministic cost), so only E - 1 evaluations
// are new.  Extinctions are computed on this unsorted population as the reference does.  Called by all
// lanes of the warp.
// -----------------------------------------------------------------------------------------------
template <class S>
__device__ __forceinline__ void init_population_warp(const SolveBuffers& sb, const WarpSmem& W, int PW, int lane) {
    const int n = c_rb.n, P = c_pr.P, E = c_pr.E;
    const int k_e = lane / E, e = lane % E;
    const bool elite_lane = lane < PW * E && W.flag[k_e] != 0;
    int64_t b = -1;
    double* dst = nullptr;
    const double* hdr = nullptr;
    if (elite_lane) {
        b = W.pidx[k_e];
        const ProblemMeta m = sb.meta[b];
        dst = pop_ptr(sb, m.iter & 1, b, n, P);
        hdr = sb.hdr + (size_t)b * (n + 2);
        double* col = W.q + lane;
        for (int j = 0; j < n; ++j) col[j * kS] = hdr[j];
        double f = hdr[n];
        if (e > 0) {
            const Stream st = make_stream((uint32_t)(sb.first_problem_index + problem_of(sb, b)), kStreamInit,
                                          (uint32_t)m.init_epoch, (uint32_t)e, species_of(sb, b));
            random_valid_configuration(sb, st, col);
            f = eval_chain<S>(col, nullptr, kViewPlain, -1, 0.0, nullptr, nullptr, W.goal + (7 * c_rb.n_tips) * k_e,
                           sb.seed + problem_of(sb, b) * sb.seed_stride, nullptr);
        }
        W.efit[lane] = f;
        for (int j = 0; j < n; ++j) {
            dst[(size_t)j * P + e] = col[j * kS];
            dst[(size_t)(n + j) * P + e] = 0.0;
        }
        dst[(size_t)(2 * n) * P + e] = f;
    }
    __syncwarp();
    for (int k = 0; k < PW; ++k) {
        if (!W.flag[k]) continue;
        const int64_t bb = W.pidx[k];
        const int buf = sb.meta[bb].iter & 1;
        uint16_t* ord = order_ptr(sb, buf, bb, P);
        for (int i = lane; i < P; i += 32) ord[i] = (uint16_t)i;
        if (c_rb.any_unbounded) {  // children = copies of the genes; only ever read to seed a random individual
            const double* h = sb.hdr + (size_t)bb * (n + 2);
            double* d = pop_ptr(sb, buf, bb, n, P);
            for (int i = E + lane; i < P; i += 32)
                for (int j = 0; j < n; ++j) d[(size_t)j * P + i] = h[j];
        }
    }
    if (elite_lane) {
        const double f0 = W.efit[k_e * E];
        const double fl = (P > E) ? hdr[n] : W.efit[k_e * E + E - 1];
        const double grading = (double)e / (double)(P - 1);  // ik_memetic.cpp:36-39
        dst[(size_t)(2 * n + 1) * P + e] = (W.efit[lane] + f0 * (grading - 1.0)) / fl;
        if (e == 0) {
            sb.meta[b].has_prev = 0;  // previous_fitness_.reset()
            sb.meta[b].init_epoch += 1;
        }
    }
    __syncwarp();
}

template <class S>
__global__ void __launch_bounds__(kThreads) memetic_init_kernel(const __grid_constant__ SolveBuffers sb, int PW) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = c_rb.n, P = c_pr.P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const WarpSmem W = carve_warp(smem_raw + (size_t)warp * warp_smem_bytes(n, P, PW, c_rb.n_tips), n, P, PW, c_rb.n_tips);
    const int64_t base = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * PW;
    if (base >= sb.B) return;
    bool keep = false;
    int64_t b = -1;
    if (lane < PW) {
        b = base + lane;
        W.pidx[lane] = b < sb.B ? (int)b : -1;
        W.flag[lane] = 0;
        if (b < sb.B) {
            const int pb = problem_of(sb, b);
            const double* sd = sb.seed + pb * sb.seed_stride;
            double* g7 = W.goal + (7 * c_rb.n_tips) * lane;
            goals_from_poses(sb.goal_pose + (size_t)(7 * c_rb.n_tips) * (size_t)pb, g7);
            double* col = W.q + lane;
            double* hdr = sb.hdr + (size_t)b * (n + 2);
            for (int j = 0; j < n; ++j) {
                col[j * kS] = sd[j];
                hdr[j] = sd[j];  // best_ = {seed, cost(seed)}, ik_memetic.cpp:18-22
            }
            double aux[5];
            const double c = eval_chain<S>(col, nullptr, kViewPlain, -1, 0.0, nullptr, nullptr, g7, sd, aux);
            hdr[n] = c;
            hdr[n + 1] = 0.0;
            ProblemMeta m{0, 0, kActive, 0};
            const bool s = solution_from_aux(aux);
            bool done = false, found = false;
            if (c_pr.stop_on_valid && s) {  // ik_memetic.cpp:294-296
                done = found = true;
            } else if (c_pr.max_generations <= 0) {
                done = true;
                found = (!c_pr.stop_on_valid && s) || c_pr.approx;
            }
            if (done) {
                m.status = found ? kSolved : kFailed;
                write_result(sb, n, b, found, sd, 1, sd, c, 0);
                if (sb.stats) {
                    if (found) atomicAdd(&sb.stats[2], 1ull);
                    atomicAdd(&sb.stats[3], 1ull);
                }
            } else {
                W.flag[lane] = 1;
                keep = true;
            }
            sb.meta[b] = m;
        }
    }
    {
        const unsigned mask = __ballot_sync(kFull, keep);
        int basepos = 0;
        if (lane == 0 && mask) basepos = atomicAdd(&sb.counters[0], __popc(mask));
        basepos = __shfl_sync(kFull, basepos, 0);
        if (keep) sb.active[basepos + __popc(mask & ((1u << lane) - 1u))] = (int32_t)b;
    }
    __syncwarp();
    init_population_warp<S>(sb, W, PW, lane);
}

// A species that sees `terminate` (src/ik_memetic.cpp:264-282): it has completed its generation, leaves the loop
// before iter++ and returns its best individual if that passes the solution test (only looked at when the
// optimisation does not stop on valid solutions) or approximate solutions are allowed.  One lane.
template <class S>
__device__ __forceinline__ void finish_terminated(const SolveBuffers& sb, const WarpSmem& W, int b, int lane) {
    const int n = c_rb.n;
    ProblemMeta m = sb.meta[b];
    const int pb = problem_of(sb, b);
    const double* sd = sb.seed + (size_t)pb * sb.seed_stride;
    const double* hdr = sb.hdr + (size_t)b * (n + 2);
    bool found = false;
    if (!c_pr.stop_on_valid) {
        double* g7 = W.goal + (7 * c_rb.n_tips) * lane;
        goals_from_poses(sb.goal_pose + (size_t)(7 * c_rb.n_tips) * (size_t)pb, g7);
        double* col = W.q + lane;
        for (int j = 0; j < n; ++j) col[j * kS] = hdr[j];
        double aux[5];
        eval_chain<S>(col, nullptr, kViewPlain, -1, 0.0, nullptr, nullptr, g7, sd, aux);
        found = solution_from_aux(aux);
    }
    if (!found && c_pr.approx) found = true;
    m.status = found ? kSolvedTerminated : kFailedTerminated;
    m.iter = m.iter - 1;
    write_result(sb, n, b, found, hdr, 1, sd, hdr[n], m.iter);
    if (sb.stats) {
        if (found) atomicAdd(&sb.stats[2], 1ull);
        atomicAdd(&sb.stats[3], 1ull);
    }
    sb.meta[b] = m;
}

// -----------------------------------------------------------------------------------------------
// gradientDescent(i) (src/ik_memetic.cpp:66-91) with L lanes per elite.  A GD step is two rounds of
// single-frame evaluations, every one through the same eval_chain call site:
//   round A   the 2n finite-difference points of the current configuration AND the current configuration
//             itself (2n + 1 tasks spread over the L lanes of the group), all reading the sin/cos cache the
//             group filled cooperatively just before (one joint per lane).  Evaluating the accepted point
//             of step s together with the finite differences of step s + 1 removes a whole serial
//             evaluation from every step; the finite differences are speculative (discarded if the
//             termination test of step s fires).
//   round B   the two line-search points on two lanes.
// Column c holds the GD state of group c.  Returns the number of step() executions of this lane's group.
// Results are bit-identical to the serial loop: every evaluation is a full chain walk of its configuration.
// -----------------------------------------------------------------------------------------------
template <class S>
__device__ __forceinline__ int gd_elite_wide(const WarpSmem& W, int L, int lane, bool valid, const double* g7,
                                             const double* sd, double& best_cost_out, long long* ph = nullptr) {
    // ph (PIK_PHASE_TRACE builds): cycles of lane 0 in [0] sin/cos cache, [1] round A, [2] control + gradient,
    // [3] round B, [4] accepted step
    long long tph = 0;
    auto mark = [&](int k) {
#ifdef PIK_PHASE_TRACE
        if (ph) {
            const long long now = clock64();
            if (k >= 0) ph[k] += now - tph;
            tph = now;
        }
#else
        (void)k; (void)tph;
#endif
    };
    const int n = c_rb.n;
    const int c = lane / L, gl = lane % L;
    const bool leader = gl == 0;
    double* q = W.q + c;
    double* g = W.g + c;
    double* best = W.best + c;
    double* sc = W.sc + c;
    const double h = c_pr.step_size;
    const int acc_lane = c * L + n % L;  // lane that evaluates the current configuration
    double local_cost = 0.0, best_cost = 0.0, previous_cost = 0.0;
    int it = 0, steps = 0;
    bool act_l = valid && leader;  // the group still runs (leader's copy)
    bool first = true;
    for (;;) {
        const bool act = __shfl_sync(kFull, act_l ? 1 : 0, c * L) != 0 && valid;
        if (!__any_sync(kFull, act)) break;
        mark(-1);
        // sin/cos cache of the current configuration, one joint per lane
        for (int j = gl; j < n; j += L) {
            if (act) {
                double sj, cj;
                joint_sincos<S>(j, q[j * kS], sj, cj);
                sc[(2 * j) * kS] = sj;
                sc[(2 * j + 1) * kS] = cj;
            }
        }
        __syncwarp();
        mark(0);
        // round A: n finite-difference pairs + the current configuration, one task per lane and sub-round
        double cur = 0.0;
        for (int k = gl; k <= n; k += L) {
            if (act) {
                const CostPair cp = pair_costs_from_origin<S>(kPairFd, k < n ? k : -1, q, nullptr, sc, g7, sd);
                if (k < n) {
                    W.cs[(2 * k) * kS + c] = cp.m;
                    W.cs[(2 * k + 1) * kS + c] = cp.p;
                } else {
                    cur = cp.m;
                }
            }
        }
        cur = __shfl_sync(kFull, cur, acc_lane);
        __syncwarp();
        mark(1);
        bool improved = false;
        if (act && leader) {
            if (first) {
                local_cost = best_cost = cur;  // GradientIk::from
                if (c_pr.gd_max_iters <= 0) act_l = false;
            } else {
                // the tail of step(): the accepted point's cost, best update (ik_gradient.cpp:88-93), then the
                // loop control of gradientDescent (ik_memetic.cpp:75-86)
                local_cost = cur;
                if (local_cost < best_cost) {
                    improved = true;
                    best_cost = local_cost;
                }
                ++steps;
                if (fabs(local_cost - previous_cost) <= c_pr.min_cost_delta) {
                    act_l = false;
                } else {
                    previous_cost = local_cost;
                    ++it;
                    if (it >= c_pr.gd_max_iters) act_l = false;
                }
            }
        }
        first = false;
        const unsigned ctl = __shfl_sync(kFull, (act_l ? 1u : 0u) | (improved ? 2u : 0u), c * L);
        const bool go = (ctl & 1u) != 0 && valid;
        // one joint per lane: best <- local when improved; the gradient g_i = C(q + h e_i) - C(q - h e_i), its
        // normalisation by h / (h + sum |g_i|) (ik_gradient.cpp:42-54; the sum runs in joint order on every lane)
        if (act) {
            double f = 0.0;
            if (go) {
                double sum = h;
                for (int i = 0; i < n; ++i) sum = sum + fabs(W.cs[(2 * i + 1) * kS + c] - W.cs[(2 * i) * kS + c]);
                f = 1.0 / sum * h;
            }
            for (int j = gl; j < n; j += L) {
                if (ctl & 2u) best[j * kS] = q[j * kS];
                if (go) g[j * kS] = (W.cs[(2 * j + 1) * kS + c] - W.cs[(2 * j) * kS + c]) * f;
            }
        }
        __syncwarp();
        mark(2);
        // round B: the line search.  With at least four lanes per elite: cooperative sin/cos and a row-parallel chain
        // walk (line_search_rows; every lane of the warp takes part); with two: one evaluation per lane.
        if (L >= 4) {
            const double cost = line_search_rows<S>(L, gl, go, q, g, sc, W.cs + c, g7, sd);
            if (go && gl < 2) W.cs[gl * kS + c] = cost;
        } else if (go && gl < 2) {
            W.cs[gl * kS + c] = eval_chain<S>(q, g, gl == 0 ? kViewMinus : kViewPlus, -1, 0.0, nullptr, nullptr, g7, sd, nullptr);
        }
        __syncwarp();
        mark(3);
        // the always-accepted step (ik_gradient.cpp:67-85), one joint per lane
        if (go) {
            const double p1 = W.cs[c], p3 = W.cs[kS + c];
            const double p2 = (p1 + p3) * 0.5;
            const double cost_diff = (p3 - p1) * 0.5;
            double joint_diff = p2 / cost_diff;
            if (!(fabs(joint_diff) <= 0x1.fffffffffffffp+1023)) joint_diff = 0.0;  // !isfinite
            for (int j = gl; j < n; j += L) q[j * kS] = clamp_to_limits(j, q[j * kS] - g[j * kS] * joint_diff);
        }
        __syncwarp();
        mark(4);
    }
    best_cost_out = best_cost;
    return steps;
}

// gd_elite_wide for chains whose n finite-difference pairs fill the L lanes of an elite exactly (n a multiple of L, L >=
// 4: the Fetch's 8 variables on 8 or 4 lanes): the accepted point of step s - 1, which gd_elite_wide evaluates beside the
// finite differences of step s, would cost a whole extra sub-round there, so it moves into the line-search round of step
// s, onto the row lanes (line_search_rows3).  step() then learns the cost of its point one round later: the gradient
// and the line search of step s are computed before the termination test of step s - 1 is known and are discarded when
// it fires (the previous gradient, which gradientDescent returns, is kept aside in registers until then).  Same
// results as the serial loop: every evaluation is a full chain walk of its configuration.
template <class S>
__device__ __forceinline__ int gd_elite_wide_deferred(const WarpSmem& W, int L, int lane, bool valid, const double* g7,
                                                      const double* sd, double& best_cost_out) {
    const int n = c_rb.n;
    const int c = lane / L, gl = lane % L;
    const bool leader = gl == 0;
    double* q = W.q + c;
    double* g = W.g + c;
    double* best = W.best + c;
    double* sc = W.sc + c;
    double* csM = W.cs + c;       // finite-difference costs, then sin/cos of q - g
    double* csP = W.cs + c + 8;   // sin/cos of q + g (L >= 4: at most 8 groups per warp, columns 8..15 are free)
    const double h = c_pr.step_size;
    double local_cost = 0.0, best_cost = 0.0, previous_cost = 0.0;
    int it = 0, steps = 0;
    bool act_l = valid && leader;  // the group still runs (leader's copy)
    bool first = true;
    for (;;) {
        // fin: this round only settles the last step (its cost is all that is missing)
        const unsigned st = __shfl_sync(kFull, (act_l ? 1u : 0u) | ((!first && it + 1 >= c_pr.gd_max_iters) ? 2u : 0u), c * L);
        const bool act = (st & 1u) != 0 && valid;
        const bool fin = (st & 2u) != 0;
        if (!__any_sync(kFull, act)) break;
        for (int j = gl; j < n; j += L) {
            if (act) {
                double sj, cj;
                joint_sincos<S>(j, q[j * kS], sj, cj);
                sc[(2 * j) * kS] = sj;
                sc[(2 * j + 1) * kS] = cj;
            }
        }
        __syncwarp();
        // round A: the n finite-difference pairs, one per lane
        const bool spec = act && !fin;
        for (int k = gl; k < n; k += L) {
            if (spec) {
                const CostPair cp = pair_costs_from_origin<S>(kPairFd, k, q, nullptr, sc, g7, sd);
                csM[(2 * k) * kS] = cp.m;
                csM[(2 * k + 1) * kS] = cp.p;
            }
        }
        __syncwarp();
        // the gradient of this step (ik_gradient.cpp:42-54), speculative; the previous one is kept in registers
        double g_old[4] = {0.0, 0.0, 0.0, 0.0};
        if (spec) {
            double sum = h;
            for (int i = 0; i < n; ++i) sum = sum + fabs(csM[(2 * i + 1) * kS] - csM[(2 * i) * kS]);
            const double f = 1.0 / sum * h;
            int slot = 0;
            for (int j = gl; j < n; j += L, ++slot) {
                const double gj = (csM[(2 * j + 1) * kS] - csM[(2 * j) * kS]) * f;
                if (slot < 4) g_old[slot] = g[j * kS];
                g[j * kS] = gj;
            }
        }
        __syncwarp();
        // round B: the line search of this step and the cost of its starting point, on the row lanes
        const double cost = line_search_rows3<S>(L, gl, spec, act, q, g, sc, csM, csP, g7, sd);
        const double p1 = __shfl_sync(kFull, cost, c * L), p3 = __shfl_sync(kFull, cost, c * L + 1);
        const double cur = __shfl_sync(kFull, cost, c * L + 2);
        bool improved = false;
        if (act && leader) {
            if (first) {
                local_cost = best_cost = cur;  // GradientIk::from
                if (c_pr.gd_max_iters <= 0) act_l = false;
            } else {
                // the tail of step(): the accepted point's cost, best update (ik_gradient.cpp:88-93), then the
                // loop control of gradientDescent (ik_memetic.cpp:75-86)
                local_cost = cur;
                if (local_cost < best_cost) {
                    improved = true;
                    best_cost = local_cost;
                }
                ++steps;
                if (fabs(local_cost - previous_cost) <= c_pr.min_cost_delta) {
                    act_l = false;
                } else {
                    previous_cost = local_cost;
                    ++it;
                    if (it >= c_pr.gd_max_iters) act_l = false;
                }
            }
        }
        first = false;
        const unsigned ctl = __shfl_sync(kFull, (act_l ? 1u : 0u) | (improved ? 2u : 0u), c * L);
        const bool go = (ctl & 1u) != 0 && valid;
        __syncwarp();  // the row lanes have read q and g (line_search_rows3)
        if (act) {
            int slot = 0;
            for (int j = gl; j < n; j += L, ++slot) {
                if (ctl & 2u) best[j * kS] = q[j * kS];
                if (!go && spec && slot < 4) g[j * kS] = g_old[slot];  // the step ended here: its gradient stands
            }
        }
        // the always-accepted step (ik_gradient.cpp:67-85), one joint per lane
        if (go) {
            const double p2 = (p1 + p3) * 0.5;
            const double cost_diff = (p3 - p1) * 0.5;
            double joint_diff = p2 / cost_diff;
            if (!(fabs(joint_diff) <= 0x1.fffffffffffffp+1023)) joint_diff = 0.0;  // !isfinite
            for (int j = gl; j < n; j += L) q[j * kS] = clamp_to_limits(j, q[j * kS] - g[j * kS] * joint_diff);
        }
        __syncwarp();
    }
    best_cost_out = best_cost;
    return steps;
}

// -----------------------------------------------------------------------------------------------
// One generation of ik_memetic_impl (src/ik_memetic.cpp:228-269) for PW problems per warp.
// -----------------------------------------------------------------------------------------------
// Register budgets: the throughput flavour (and the generic kernel, which serves both modes) runs 1 CTA of
// 16 warps per SM at 128 registers; the wide flavour (several lanes per elite) 3 CTAs of 4 warps at 168 registers
// (no spills: a lone warp pays the full latency of every local-memory access).
//
// Launch policy, evaluated on the device: the host enqueues, for every generation g, one launch of the throughput
// flavour and one of the wide flavour -- each a grid of RESIDENT CTAs (1 resp. 3 per SM) -- and never reads a count
// back.  Each launch reads the size of generation g's active list, derives the lane mapping from it (lanes_for, the
// function the host used to call) and returns at once unless the mapping is its own.
//
// Work distribution.  The work units (a warp's PW problems) are dealt round-robin into one queue per SM (unit u ->
// queue u mod n_sm); a CTA claims, one warp-load at a time, from the queue of the SM it RUNS ON (%smid through the
// dense table sm_dense), so that every SM receives the same number of warps whatever the size of the list and
// wherever the hardware placed the CTAs -- what used to be a host-side choice of grid and CTA size.  A CTA that has
// worked and finds its queue empty takes from the other SMs' queues (the tail of a launch of several rounds
// balances like the hardware's own CTA scheduling); one that finds its queue empty at once leaves (the launch fills
// less than a wave and the other queues belong to CTAs that are just starting), and the last CTA to leave sweeps up
// whatever no SM-local CTA has taken (possible only when other kernels occupy part of the device).
//
// A wide launch with one problem per warp and at most persistent_units_max problems keeps every problem for all
// its remaining generations: problems are independent, so the tail of the batch needs no global step between
// generations; the launches enqueued behind it find an empty list.
// Block barriers: the lanes of a warp reconverge first (bar.sync is an aligned barrier: every thread of a warp must
// execute it together, which code that has just left a divergent region does not guarantee by itself).
__device__ __forceinline__ void block_barrier() {
    __syncwarp();
    __syncthreads();
}
__device__ __forceinline__ int block_barrier_or(int predicate) {
    __syncwarp();
    return __syncthreads_or(predicate);
}

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

__device__ __forceinline__ unsigned sm_id() {
    unsigned v;
    asm("mov.u32 %0, %%smid;" : "=r"(v));
    return v;
}

template <class S>
__global__ void __launch_bounds__(S::kWide ? 128 : 32 * kWarpsPerBlockBulk, S::kWide ? 3 : 1) memetic_generation_kernel(const __grid_constant__ SolveBuffers sb,
                                                                      int gen) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_claim[4];
    const int n_listed = sb.counters[gen];
    if (n_listed <= 0) return;
    const int n = c_rb.n, P = c_pr.P, E = c_pr.E;
    const int L = lanes_for(n_listed, E, c_pr.lanes_max, c_pr.wide_capacity_lanes, c_pr.wide_units_max);
    if ((L > 1) != S::kWide) return;
    const int PW = 32 / (E * L);
    const int list_in = gen & 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int nsm = c_pr.sm_count;
    // Whole waves only: a throughput launch whose list is longer than what the resident CTAs hold at a time processes
    // a whole number of such waves and passes the rest of the list on to the next launch untouched (problems are
    // independent and carry their own generation count, so a launch is a scheduling step, not a generation: the
    // part-filled last round of a launch runs at a fraction of the throughput of a full one).  The entries passed on
    // are copied before anything else, so they lead the next list and are processed by the next launch.
    int n_active = n_listed;
    if (!S::kWide && gen < c_pr.defer_launches) {
        const int wave = (int)gridDim.x * wpb * PW;
        if ((int)gridDim.x == c_pr.wave_ctas && n_listed > wave) n_active = (n_listed / wave) * wave;
    }
    const int units = (n_active + PW - 1) / PW;
    const int sigma = c_pr.sm_dense[sm_id() & (kSmDenseSize - 1)];
    int* qhead = sb.sched + (size_t)gen * (size_t)(nsm + 2);  // [nsm] queue heads, CTAs that have left, units taken
    // Units are dealt to the queues in chunks: one unit at a time for the throughput flavour (every SM the same number
    // of warps, +-1), a CTA's worth for the wide flavour (a launch bound by one warp's latency loses nothing when its
    // warps share an SM, and the fewer CTAs it keeps resident the more room there is for the launches of the other
    // sub-batches of the solve, which run beside it).  Dealing starts at queue sm_rotation, so that the sub-batches
    // of a solve do not all begin on the same SMs.
    const int chunk = S::kWide ? wpb : 1;
    const int n_chunks = (units + chunk - 1) / chunk;
    const int rot = sb.sm_rotation % nsm;
    const bool persistent = S::kWide && PW == 1 && units <= c_pr.persistent_units_max;  // (0 for species in lockstep)
    // (with launches that pass problems on, an upper bound: a problem leaves when ITS generation count is up)
    const int max_gens = persistent ? (c_pr.defer_launches > 0 ? c_pr.max_generations : c_pr.max_generations - gen) : 1;
    // Throughput mode keeps the warps of a CTA in step through the GD phase with one block barrier per
    // GD step (below): warps that run the same code at the same time share their instruction-cache fills.
    // Every warp of the CTA executes the same number of these barriers before it reaches the claim barrier again
    // (idle warps run them empty below; nobody leaves the CTA early); a persistent launch, whose warps run different
    // numbers of generations, takes none.
    const bool lockstep = L == 1 && !persistent && (c_pr.lockstep & 1) != 0;
    const bool lockstep_rep = L == 1 && !persistent && (c_pr.lockstep & 2) != 0;
    const int pw_carve = S::kWide ? wide_problems_per_warp_max(E) : PW;
    const WarpSmem W = carve_warp(smem_raw + (size_t)warp * warp_smem_bytes(n, P, pw_carve, c_rb.n_tips), n, P, pw_carve, c_rb.n_tips);
    const int32_t* act_in = sb.active + (size_t)list_in * (size_t)sb.B;
    int32_t* act_out = sb.active + (size_t)(list_in ^ 1) * (size_t)sb.B;
    if (n_active < n_listed) {
        const int rest = n_listed - n_active;
        const int per = (rest + (int)gridDim.x - 1) / (int)gridDim.x;
        const int lo = n_active + (int)blockIdx.x * per;
        const int hi = lo + per < n_listed ? lo + per : n_listed;
        if (lo < hi) {
            if (threadIdx.x == 0) s_claim[0] = atomicAdd(&sb.counters[gen + 1], hi - lo);
            block_barrier();
            const int at = s_claim[0];
            for (int i = lo + (int)threadIdx.x; i < hi; i += (int)blockDim.x) act_out[at + (i - lo)] = act_in[i];
            block_barrier();  // s_claim is rewritten by the first claim
        }
    }
    bool worked = false, sweeping = false;
  for (;;) {
    if (warp == 0) {
        // entries of queue t: units t, t + nsm, t + 2 nsm, ...
        auto queue_len = [&](int t) {
            const int tq = t - rot >= 0 ? t - rot : t - rot + nsm;  // the queue's rank in dealing order
            return tq < n_chunks ? ((n_chunks - tq + nsm - 1) / nsm) * chunk : 0;
        };
        int got_q = -1, got_e = 0, got_n = 0;
        auto try_queue = [&](int t) {  // one lane
            const int ql = queue_len(t);
            const int old = atomicAdd(&qhead[t], wpb);
            if (old < ql) {
                got_q = t;
                got_e = old;
                got_n = ql - old < wpb ? ql - old : wpb;
                atomicAdd(&qhead[nsm + 1], got_n);  // units taken so far, over all queues
            }
        };
        if (lane == 0 && ld_volatile(&qhead[sigma]) < queue_len(sigma)) try_queue(sigma);
        got_q = __shfl_sync(kFull, got_q, 0);
        if (got_q < 0 && (worked || sweeping) && ld_volatile(&qhead[nsm + 1]) < n_chunks * chunk) {
            // the other SMs' queues, 32 at a time, nearest first
            for (int d0 = 1; d0 < nsm && got_q < 0; d0 += 32) {
                const int d = d0 + lane;
                const int t = sigma + d < nsm ? sigma + d : sigma + d - nsm;
                const bool open = d < nsm && ld_volatile(&qhead[t]) < queue_len(t);
                unsigned mask = __ballot_sync(kFull, open);
                while (mask != 0 && got_q < 0) {
                    const int l = __ffs(mask) - 1;
                    if (lane == l) try_queue(t);
                    got_q = __shfl_sync(kFull, got_q, l);
                    if (got_q >= 0) {
                        got_e = __shfl_sync(kFull, got_e, l);
                        got_n = __shfl_sync(kFull, got_n, l);
                    }
                    mask &= mask - 1;
                }
            }
        }
        if (lane == 0) {
            s_claim[0] = got_q;
            s_claim[1] = got_e;
            s_claim[2] = got_n;
        }
    }
    block_barrier();
    const int claim_q = s_claim[0], claim_e = s_claim[1], claim_n = s_claim[2];
    if (claim_n == 0) {
        if (sweeping) return;
        if (threadIdx.x == 0)
            s_claim[3] = (atomicAdd(&qhead[nsm], 1) == (int)gridDim.x - 1 && ld_volatile(&qhead[nsm + 1]) < n_chunks * chunk) ? 1 : 0;
        block_barrier();
        if (!s_claim[3]) return;
        sweeping = true;
        continue;
    }
    worked = true;
    // wide launches rotate the warps over the entries from claim to claim so that partly filled CTAs of one SM do
    // not pile their warps on the same sub-partitions
    const int r_unit = S::kWide ? (warp + wpb - (claim_e / wpb) % wpb) % wpb : warp;
    int64_t base = n_active;
    if (r_unit < claim_n) {
        // entry e of queue t: chunk (e / chunk) * nsm + rank of t, unit e % chunk of it
        const int e = claim_e + r_unit;
        const int tq = claim_q - rot >= 0 ? claim_q - rot : claim_q - rot + nsm;
        const int64_t unit = ((int64_t)(e / chunk) * nsm + tq) * chunk + e % chunk;
        if (unit < units) base = unit * PW;
    }
    // A warp without a unit has nothing to do -- unless the CTA's warps walk in step: then it goes through the code
    // below with no problem in any slot, so that every warp of the CTA executes the SAME barrier instructions the
    // same number of times (a block barrier in conditional code is defined only if the whole block takes it).
    if (base >= n_active && !lockstep && !lockstep_rep) {
    } else {
    if (lane < PW) {
        const int64_t idx = base + lane;
        int b = idx < n_active ? act_in[idx] : -1;
        if (b >= 0 && sb.group_term) {
            // `terminate` (src/ik_memetic.cpp:264-268): a species of this problem returned a value in an EARLIER
            // generation (a flag set in this launch is not looked at: the species run in lockstep)
            const int term = sb.group_term[problem_of(sb, b)];
            if (term > 0 && term <= gen) {
                finish_terminated<S>(sb, W, b, lane);
                b = -1;
            }
        }
        W.pidx[lane] = b;
        W.flag[lane] = 0;
        if (b >= 0) goals_from_poses(sb.goal_pose + (size_t)(7 * c_rb.n_tips) * (size_t)problem_of(sb, b), W.goal + (7 * c_rb.n_tips) * lane);
    }
    __syncwarp();

  for (int gen_here = 0;; ++gen_here) {
#ifdef PIK_PHASE_TRACE  // build with -DPIK_PHASE_TRACE: PIK_DEBUG_PHASES=1 prints per-phase cycle counts (costs a stack frame)
    const bool dbg = c_pr.debug && blockIdx.x == 0 && warp == 0 && lane == 0;
#else
    constexpr bool dbg = false;
#endif
    long long t_start = 0, t_gd = 0, t_rep = 0, t_sort = 0, t_book = 0;
    if (dbg) t_start = clock64();
    // ---- gradientDescent(i) for every elite (src/ik_memetic.cpp:66-91, 230-239)
    int gd_steps = 0;
    {
        const int c = lane / L;            // GD column = elite group
        const int k = c / E, e = c % E;    // problem slot, elite
        const int b = c < PW * E ? W.pidx[k] : -1;
        const bool valid = b >= 0;
        const bool leader = (lane % L) == 0;
        const double* sd = nullptr;
        const double* src = nullptr;
        double* dst = nullptr;
        int slot = 0;
        if (valid) {
            const int iter = sb.meta[b].iter;
            slot = order_ptr(sb, iter & 1, b, P)[e];
            src = pop_ptr(sb, iter & 1, b, n, P);
            dst = pop_ptr(sb, (iter & 1) ^ 1, b, n, P);
            sd = sb.seed + (size_t)problem_of(sb, b) * sb.seed_stride;
            if (leader) {
                for (int j = 0; j < n; ++j) {
                    const double v = src[(size_t)j * P + slot];
                    W.q[j * kS + c] = v;
                    W.best[j * kS + c] = v;
                    W.g[j * kS + c] = 0.0;  // GradientIk::from: zero gradient
                }
            }
        }
        __syncwarp();
        double best_cost = 0.0;
        if (L == 1) {
            GdState st{W.q + c, W.g + c, W.best + c, W.sc + c, nullptr, 0.0, 0.0};
            const double* g7 = W.goal + (7 * c_rb.n_tips) * (valid ? k : 0);
            if (valid) st.local_cost = st.best_cost = eval_chain<S>(st.q, nullptr, kViewPlain, -1, 0.0, nullptr, st.sc, g7, sd, nullptr);
            bool going = valid;
            double previous_cost = 0.0;
            for (int step = 0; step < c_pr.gd_max_iters; ++step) {
                if (lockstep) {
                    block_barrier();
                } else if (!__any_sync(kFull, going)) {
                    break;
                }
                if (going) {
                    gd_step<S>(st, g7, sd, nullptr);
                    ++gd_steps;
                    if (fabs(st.local_cost - previous_cost) <= c_pr.min_cost_delta) going = false;
                    previous_cost = st.local_cost;
                }
            }
            best_cost = st.best_cost;
        } else {
#ifdef PIK_PHASE_TRACE
            long long ph[5] = {0, 0, 0, 0, 0};
            const int steps = gd_elite_wide<S>(W, L, lane, valid, W.goal + (7 * c_rb.n_tips) * (valid ? k : 0), sd, best_cost, dbg ? ph : nullptr);
            if (dbg)
                printf("pik wide gd phases: sincos %lld  roundA %lld  control %lld  roundB %lld  accept %lld\n", ph[0], ph[1],
                       ph[2], ph[3], ph[4]);
#else
            // (n a multiple of L: the accepted point would cost the finite-difference round an extra sub-round)
            const bool deferred = !S::kTree && L >= 4 && n % L == 0 && n <= 4 * L && (c_pr.lockstep & 4) == 0;
            const int steps = deferred
                                  ? gd_elite_wide_deferred<S>(W, L, lane, valid, W.goal + (7 * c_rb.n_tips) * (valid ? k : 0), sd, best_cost)
                                  : gd_elite_wide<S>(W, L, lane, valid, W.goal + (7 * c_rb.n_tips) * (valid ? k : 0), sd, best_cost);
#endif
            if (leader) gd_steps = steps;
        }
        // genes <- best, fitness <- cost_fn(best) (== best_cost: same genes, deterministic cost),
        // gradient <- the last normalised gradient (ik_memetic.cpp:88-90)
        if (valid && leader) {
            for (int j = 0; j < n; ++j) {
                dst[(size_t)j * P + slot] = W.best[j * kS + c];
                dst[(size_t)(n + j) * P + slot] = W.g[j * kS + c];
            }
            const double ext = src[(size_t)(2 * n + 1) * P + slot];
            W.efit[c] = best_cost;
            W.eext[c] = ext;
            dst[(size_t)(2 * n) * P + slot] = best_cost;
            dst[(size_t)(2 * n + 1) * P + slot] = ext;
        }
    }
    __syncwarp();

    if (dbg) t_gd = clock64();
    // ---- per problem: reproduce, sortPopulation, best update
    const double inv_n = 1.0 / (double)n;
    int n_problems = 0;
    // Robots without unbounded variables need only the E best and the worst of a generation, which are kept as the
    // children are committed: the G = PW problems of the warp reproduce side by side, 32 / G lanes each (a child
    // that removes a parent costs its problem the rest of a window of 32 / G children instead of 32).  Otherwise
    // (whole order needed, or one problem per warp) one problem at a time over all 32 lanes.
    const int G = (PW > 1 && !c_rb.any_unbounded && (c_pr.lockstep & 8) == 0) ? PW : 1;
    const int LG = 32 / G;
    for (int k0 = 0; k0 < PW; k0 += G) {
        if (lockstep_rep && G == 1) block_barrier();  // re-align the CTA's warps at every problem (instruction-cache sharing)
        const bool in_group = lane / LG < G;  // (32 / G lanes per problem: E = 5 leaves lanes 30 and 31 out)
        const int kk = in_group ? k0 + lane / LG : k0;  // this lane's problem slot
        const int gl = lane % LG;
        const int b = in_group ? W.pidx[kk] : -1;
        const int c0 = kk * E;  // first elite column of the problem
        const int t0 = kk * (E + 1);
        int iter = 0;
        const double* src = nullptr;
        double* dst = nullptr;
        const uint16_t* ord_in = nullptr;
        const double* sd = nullptr;
        uint32_t rng_problem = 0, rng_species = 0;
        if (b >= 0) {
            iter = sb.meta[b].iter;
            src = pop_ptr(sb, iter & 1, b, n, P);
            dst = pop_ptr(sb, (iter & 1) ^ 1, b, n, P);
            ord_in = order_ptr(sb, iter & 1, b, P);
            sd = sb.seed + (size_t)problem_of(sb, b) * sb.seed_stride;
            rng_problem = (uint32_t)(sb.first_problem_index + problem_of(sb, b));
            rng_species = species_of(sb, b);
        }
        const double* g7 = W.goal + (7 * c_rb.n_tips) * kk;
        if (gl < E && in_group) {
            const double fe = W.efit[c0 + gl];
            W.pool[c0 + gl] = gl;
            if (G == 1) {
                W.fit[gl] = fe;
            } else {
                // running order of the E best (the elites to begin with) and the worst, key (fitness, position)
                int rank = 0;
                for (int e = 0; e < E; ++e)
                    if (key_less(W.efit[c0 + e], e, fe, gl)) ++rank;
                W.rtf[t0 + rank] = fe;
                W.rti[t0 + rank] = gl;
                if (rank == E - 1) {
                    W.rtf[t0 + E] = fe;
                    W.rti[t0 + E] = gl;
                }
            }
        }
        if (gl == 0 && in_group) {
            W.grp[kk] = E;                  // mating pool size
            W.grp[32 + kk] = b >= 0 ? E : P;  // next child
        }
        __syncwarp();

        // reproduce (src/ik_memetic.cpp:119-190).  A child that beats a parent removes it from the mating
        // pool, which changes the parent draws of every later child: a window is committed up to the first
        // such child and the walk resumes after it.  Each child's random stream is keyed by (generation,
        // child), so its draws do not depend on the history.
        double* col = W.q + lane;
        if (dbg) t_rep -= clock64();
        for (;;) {
            const int s = in_group ? W.grp[32 + kk] : P;
            const int i = s + gl;
            const bool actv = i < P;
            const bool any = __any_sync(kFull, actv);
            if (lockstep_rep && G > 1) {
                if (!block_barrier_or(any)) break;  // the CTA's warps walk their windows in step
                if (!any) continue;
            } else if (!any) {
                break;
            }
            const int ps = W.grp[kk];
            double f = 0.0;
            int ia = 0, ib = 0;
            bool removes = false;
            if (actv) {
                const int slot = ord_in[i];
                if (ps > 0) {
                    const Stream st = make_stream(rng_problem, kStreamReproduce, (uint32_t)iter, (uint32_t)i, rng_species);
                    uint32_t m0, m1, m2, m3;
                    IndexWords iw;
                    iw.st = st;
                    philox_block(sb, st, 0, iw.h0, iw.h1, iw.h2, iw.h3);
                    philox_block(sb, st, 1, m0, m1, m2, m3);
                    iw.h4 = m2;
                    iw.h5 = m3;
                    iw.v0 = iw.v1 = iw.v2 = iw.v3 = 0;
                    iw.ovf_block = (uint32_t)(2 * n + 2);
                    iw.pos = 0;
                    const uint32_t idxA = uniform_int_words(sb, iw, (uint32_t)ps);
                    uint32_t idxB = idxA;
                    while (idxB == idxA && ps > 1) idxB = uniform_int_words(sb, iw, (uint32_t)ps);
                    ia = W.pool[c0 + idxA];
                    ib = W.pool[c0 + idxB];
                    const int ca = c0 + ia, cb = c0 + ib;
                    const double extinction = 0.5 * (W.eext[ca] + W.eext[cb]);
                    const double mutation_prob = extinction * (1.0 - inv_n) + inv_n;
                    const double mix = uniform_real_words(0.0, 1.0, m0, m1);
#pragma unroll 1
                    for (int j = 0; j < n; ++j) {
                        uint32_t a0, a1, a2, a3, u0, u1, u2, u3;
                        philox_block(sb, st, (uint32_t)(2 + 2 * j), a0, a1, a2, a3);
                        philox_block(sb, st, (uint32_t)(3 + 2 * j), u0, u1, u2, u3);
                        double gene = mix * W.best[j * kS + ca] + (1.0 - mix) * W.best[j * kS + cb];
                        const double rA = uniform_real_words(0.0, 1.0, a0, a1);
                        const double rB = uniform_real_words(0.0, 1.0, a2, a3);
                        gene = gene + (rA * W.g[j * kS + ca] + rB * W.g[j * kS + cb]);
                        const double original = gene;
                        if (uniform_real_words(0.0, 1.0, u0, u1) < mutation_prob)
                            gene = gene + extinction * c_rb.vhalf[j] * uniform_real_words(-1.0, 1.0, u2, u3);
                        gene = clamp_to_limits(j, gene);
                        col[j * kS] = gene;
                        dst[(size_t)j * P + slot] = gene;
                        dst[(size_t)(n + j) * P + slot] = gene - original;
                    }
                } else {
                    // empty pool: a random individual seeded from the slot's previous occupant
                    const Stream st = make_stream(rng_problem, kStreamRandomChild, (uint32_t)iter, (uint32_t)i, rng_species);
                    // (only an unbounded variable reads its previous value, robot.cpp:23-30; the child slots of a fresh
                    // population are written for robots with such variables only)
                    for (int j = 0; j < n; ++j) col[j * kS] = c_rb.bounded[j] ? 0.0 : src[(size_t)j * P + slot];
                    random_valid_configuration(sb, st, col);
                    for (int j = 0; j < n; ++j) {
                        dst[(size_t)j * P + slot] = col[j * kS];
                        dst[(size_t)(n + j) * P + slot] = 0.0;
                    }
                }
                f = eval_chain<S>(col, nullptr, kViewPlain, -1, 0.0, nullptr, nullptr, g7, sd, nullptr);
                dst[(size_t)(2 * n) * P + slot] = f;
                if (ps > 0) removes = f < W.efit[c0 + ia] || f < W.efit[c0 + ib];
            }
            const unsigned ballot = __ballot_sync(kFull, removes);
            const unsigned gmask = LG == 32 ? ballot : (ballot >> (lane - gl)) & ((1u << LG) - 1u);  // this problem's lanes
            const int l = gmask ? __ffs(gmask) - 1 : LG - 1;  // last lane of the group whose child is committed
            const int from = (lane - gl + l) & 31;
            const double fl = __shfl_sync(kFull, f, from);
            const int a = __shfl_sync(kFull, ia, from), bb = __shfl_sync(kFull, ib, from);
            const bool commit = actv && gl <= l;
            __syncwarp();  // every lane has read the group's pool and next-child entries
            if (gl == 0 && actv) {
                if (gmask) {
                    int p2 = ps;
                    // parents are referenced by identity; A first, then B (ik_memetic.cpp:170-177)
                    for (int which = 0; which < 2; ++which) {
                        const int target = which == 0 ? a : bb;
                        if (fl < W.efit[c0 + target]) {
                            for (int x = 0; x < p2; ++x)
                                if (W.pool[c0 + x] == target) {
                                    for (int y = x; y + 1 < p2; ++y) W.pool[c0 + y] = W.pool[c0 + y + 1];
                                    --p2;
                                    break;
                                }
                        }
                    }
                    W.grp[kk] = p2;
                }
                W.grp[32 + kk] = s + l + 1;
            }
            if (G == 1) {
                if (commit) W.fit[i] = f;
            } else {
                // the committed children that enter the E best or become the worst, in position order
                const bool enters = commit && (key_less(f, i, W.rtf[t0 + E - 1], W.rti[t0 + E - 1]) ||
                                               key_less(W.rtf[t0 + E], W.rti[t0 + E], f, i));
                const unsigned eb = __ballot_sync(kFull, enters);
                if (eb) {
                    W.fit[lane] = f;  // (the sc rows: free during reproduce)
                    __syncwarp();
                    unsigned em = LG == 32 ? eb : (eb >> (lane - gl)) & ((1u << LG) - 1u);
                    if (gl == 0) {
                        while (em) {
                            const int x = __ffs(em) - 1;
                            em &= em - 1;
                            const double fx = W.fit[lane + x];
                            const int ix = s + x;
                            if (key_less(W.rtf[t0 + E], W.rti[t0 + E], fx, ix)) {
                                W.rtf[t0 + E] = fx;
                                W.rti[t0 + E] = ix;
                            }
                            int r = E;
                            while (r > 0 && key_less(fx, ix, W.rtf[t0 + r - 1], W.rti[t0 + r - 1])) --r;
                            if (r < E) {
                                for (int y = E - 1; y > r; --y) {
                                    W.rtf[t0 + y] = W.rtf[t0 + y - 1];
                                    W.rti[t0 + y] = W.rti[t0 + y - 1];
                                }
                                W.rtf[t0 + r] = fx;
                                W.rti[t0 + r] = ix;
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (dbg) t_rep += clock64();

      for (int k = k0; k < k0 + G; ++k) {
        const int b = W.pidx[k];
        if (b < 0) continue;
        ++n_problems;
        const int iter = sb.meta[b].iter;
        double* dst = pop_ptr(sb, (iter & 1) ^ 1, b, n, P);
        const uint16_t* ord_in = order_ptr(sb, iter & 1, b, P);
        uint16_t* ord_out = order_ptr(sb, (iter & 1) ^ 1, b, P);
        // sortPopulation (src/ik_memetic.cpp:200-209) under the total order (fitness, position), NaN last.
        if (G > 1) {
            // the running order kept while the children were committed
            for (int i = lane; i < P; i += 32) ord_out[i] = ord_in[i];
            for (int r = lane; r <= E; r += 32) W.top[r] = W.rti[k * (E + 1) + r];
            __syncwarp();
        } else if (c_rb.any_unbounded) {
            // whole order needed: the previous occupant of every position may seed a random individual.  Bitonic
            // sort of the positions (padded to a power of two; padding sorts last) in shared memory: the key is a
            // total order, so any sorting network gives the permutation of the reference's stable sort.
            const int PP = pow2_at_least(P);
            for (int i = lane; i < PP; i += 32) W.perm[i] = (uint16_t)i;
            __syncwarp();
            auto pos_less = [&](int a, int b) {  // positions >= P are padding
                if (a >= P || b >= P) return a < b;
                return key_less(W.fit[a], a, W.fit[b], b);
            };
            for (int k = 2; k <= PP; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = lane; t < (PP >> 1); t += 32) {
                        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // the pair (i, i + j) of this stage
                        const int x = i | j;
                        const int a = W.perm[i], b = W.perm[x];
                        const bool ascending = (i & k) == 0;
                        if (pos_less(b, a) == ascending) {
                            W.perm[i] = (uint16_t)b;
                            W.perm[x] = (uint16_t)a;
                        }
                    }
                    __syncwarp();
                }
            }
            for (int r = lane; r < P; r += 32) {
                const int i = W.perm[r];
                ord_out[r] = ord_in[i];
                if (r < E) W.top[r] = i;
                if (r == P - 1) W.top[E] = i;
            }
            __syncwarp();
        } else {
            // only the E best (the next parents) and the worst (extinction scale) matter: E + 1 warp reductions
            for (int i = lane; i < P; i += 32) ord_out[i] = ord_in[i];
            double pf = -INFINITY;
            int pi = -1;
            for (int r = 0; r <= E; ++r) {
                const bool want_max = r == E;
                double mf = want_max ? -INFINITY : make_nan();
                int mi = want_max ? -1 : INT_MAX;
                for (int i = lane; i < P; i += 32) {
                    const double fi = W.fit[i];
                    if (want_max) {
                        if (key_less(mf, mi, fi, i)) { mf = fi; mi = i; }
                    } else if (key_less(pf, pi, fi, i) && key_less(fi, i, mf, mi)) {
                        mf = fi; mi = i;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double of = __shfl_xor_sync(kFull, mf, o);
                    const int oi = __shfl_xor_sync(kFull, mi, o);
                    const bool take = want_max ? key_less(mf, mi, of, oi) : key_less(of, oi, mf, mi);
                    if (take) { mf = of; mi = oi; }
                }
                if (lane == 0) W.top[r] = mi;
                pf = mf;
                pi = mi;
            }
            __syncwarp();
        }
        if (!c_rb.any_unbounded) {
            if (lane == 0) {
                // bring the E best to the front of the slot-index row with swaps
                for (int r = 0; r < E; ++r) W.pool[r] = W.top[r];
                for (int r = 0; r < E; ++r) {
                    const int at = W.pool[r];
                    if (at != r) {
                        const uint16_t tmp = ord_out[r];
                        ord_out[r] = ord_out[at];
                        ord_out[at] = tmp;
                        for (int r2 = r + 1; r2 < E; ++r2)
                            if (W.pool[r2] == r) W.pool[r2] = at;
                    }
                }
            }
            __syncwarp();
        }
        // computeExtinctions (ik_memetic.cpp:57-64); only the next generation's parents are ever read
        auto top_fitness = [&](int r) { return G > 1 ? W.rtf[k * (E + 1) + r] : W.fit[W.top[r]]; };
        const double f0 = top_fitness(0);
        const double fmax = top_fitness(E);
        if (lane < E) {
            const double grading = (double)lane / (double)(P - 1);
            dst[(size_t)(2 * n + 1) * P + ord_out[lane]] = (top_fitness(lane) + f0 * (grading - 1.0)) / fmax;
        }
        double* hdr = sb.hdr + (size_t)b * (n + 2);
        if (f0 < hdr[n]) {  // best_ = best_curr_, ik_memetic.cpp:206-208
            const int slot0 = ord_out[0];
            __syncwarp();
            if (lane < n) hdr[lane] = dst[(size_t)lane * P + slot0];
            if (lane == 0) hdr[n] = f0;
        }
        if (lane == 0) W.f0s[k] = f0;
        __syncwarp();
      }
    }

    if (dbg) t_sort = clock64();
    // ---- per problem, one lane each: solution test, wipeout check, bookkeeping
    {
        bool keep = false;
        int b = -1;
        if (lane < PW && W.pidx[lane] >= 0) {
            const int k = lane;
            b = W.pidx[k];
            ProblemMeta m = sb.meta[b];
            const int iter = m.iter;
            const double* sd = sb.seed + (size_t)problem_of(sb, b) * sb.seed_stride;
            double* hdr = sb.hdr + (size_t)b * (n + 2);
            const double f0 = W.f0s[k];
            const bool final_gen = iter + 1 >= c_pr.max_generations;
            bool s = false;
            if (c_pr.stop_on_valid || final_gen) {
                double* col = W.q + lane;
                for (int j = 0; j < n; ++j) col[j * kS] = hdr[j];
                double aux[5];
                eval_chain<S>(col, nullptr, kViewPlain, -1, 0.0, nullptr, nullptr, W.goal + (7 * c_rb.n_tips) * k, sd, aux);
                s = solution_from_aux(aux);
            }
            bool done = false, found = false;
            int its = iter;
            if (c_pr.stop_on_valid && s) {  // ik_memetic.cpp:252-255
                done = found = true;
            } else if (final_gen) {  // ik_memetic.cpp:272-282
                done = true;
                found = (!c_pr.stop_on_valid && s) || c_pr.approx;
                its = iter + 1;
            }
            if (done) {
                m.status = found ? kSolved : kFailed;
                m.iter = its;
                write_result(sb, n, b, found, hdr, 1, sd, hdr[n], its);
                // the first species to return a value sets `terminate` for the others (src/ik_memetic.cpp:334-346)
                if (found && sb.group_term && sb.stop_on_first) sb.group_term[problem_of(sb, b)] = iter + 1;
                if (sb.stats) {
                    if (found) atomicAdd(&sb.stats[2], 1ull);
                    atomicAdd(&sb.stats[3], 1ull);
                }
            } else {
                // checkWipeout (ik_memetic.cpp:43-55)
                bool wipe = false;
                if (m.has_prev) wipe = !(f0 < hdr[n + 1] - c_pr.wipeout_tol);
                if (!wipe) {
                    m.has_prev = 1;
                    hdr[n + 1] = f0;
                }
                W.flag[k] = wipe ? 1 : 0;
                m.iter = iter + 1;
                keep = true;
            }
            sb.meta[b] = m;
        }
        if (sb.stats) {
            int steps = gd_steps;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(kFull, steps, o);
            if (lane == 0 && n_problems > 0) {
                atomicAdd(&sb.stats[0], (unsigned long long)n_problems);
                atomicAdd(&sb.stats[1], (unsigned long long)steps);
            }
        }
        const unsigned mask = __ballot_sync(kFull, keep);
        // the warp keeps its (single) problem for another generation
        const bool again = PW == 1 && mask != 0 && gen_here + 1 < max_gens;
        if (!again) {
            int basepos = 0;
            if (lane == 0 && mask) basepos = atomicAdd(&sb.counters[gen + gen_here + 1], __popc(mask));
            basepos = __shfl_sync(kFull, basepos, 0);
            if (keep) act_out[basepos + __popc(mask & ((1u << lane) - 1u))] = b;
        }
        __syncwarp();
        if (dbg) t_book = clock64();
        init_population_warp<S>(sb, W, PW, lane);
        if (dbg)
            printf("pik phases L=%d PW=%d: gd %lld  reproduce %lld  sort+best %lld  bookkeeping %lld  init %lld cycles (gd steps %d)\n", L,
                   PW, t_gd - t_start, t_rep, t_sort - t_gd - t_rep, t_book - t_sort, clock64() - t_book, gd_steps);
        if (!again) break;
        if (lane < PW) W.flag[lane] = 0;
        __syncwarp();
    }
  }
    }  // this warp's unit
    block_barrier();  // s_claim is rewritten by the next claim
  }
}

// ik_memetic's pick over the species of a problem (src/ik_memetic.cpp:334-370), one thread per problem.  The
// reference takes the species' results in the order they are pushed: here the order of the lockstep schedule --
// (generations executed, returned at the solution test before returned on `terminate`, species index).  With
// stop_on_first_soln the first result, if it holds a value, is taken unconditionally; every later value replaces the
// pick only when its fitness is strictly lower.  No value at all: NO_IK_SOLUTION, the seed, the lowest best fitness
// and the largest generation count of the species.
__global__ void species_pick_kernel(const __grid_constant__ SolveBuffers sb, int n, int64_t n_problems, double* solution,
                                    int32_t* error_code, double* cost, int32_t* iterations) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_problems) return;
    const int S = sb.n_species;
    auto phase_of = [](int status) { return (status == kSolvedTerminated || status == kFailedTerminated) ? 1 : 0; };
    int pick = -1;
    double min_cost = 1.7976931348623157e308;  // std::numeric_limits<double>::max()
    int prev_it = -1, prev_ph = -1, prev_s = -1;
    double fail_cost = 0.0;
    int fail_it = 0;
    for (int k = 0; k < S; ++k) {
        // k-th result in push order: the smallest (iterations, phase, species) above the previous one
        int cur = -1, cit = 0, cph = 0;
        for (int s = 0; s < S; ++s) {
            const int64_t sp = b * S + s;
            const int it = sb.iterations[sp], ph = phase_of(sb.meta[sp].status);
            const bool after_prev = it > prev_it || (it == prev_it && (ph > prev_ph || (ph == prev_ph && s > prev_s)));
            const bool before_cur = cur < 0 || it < cit || (it == cit && (ph < cph || (ph == cph && s < cur)));
            if (after_prev && before_cur) { cur = s; cit = it; cph = ph; }
        }
        prev_it = cit; prev_ph = cph; prev_s = cur;
        const int64_t sp = b * S + cur;
        const double c = sb.cost[sp];
        const bool has_value = sb.error_code[sp] == 1;
        if (k == 0 || fit_less(c, fail_cost)) fail_cost = c;
        if (cit > fail_it) fail_it = cit;
        if (has_value && ((k == 0 && sb.stop_on_first) || c < min_cost)) {
            pick = cur;
            min_cost = c;
        }
    }
    const double* sd = sb.seed + (size_t)b * sb.seed_stride;
    double* out = solution + (size_t)b * n;
    if (pick >= 0) {
        const int64_t sp = b * S + pick;
        for (int j = 0; j < n; ++j) out[j] = sb.solution[sp * n + j];
        error_code[b] = 1;
        if (sb.stats) atomicAdd(&sb.stats[4], 1ull);  // problems with a value (stats[2] counts species)
        if (cost) cost[b] = sb.cost[sp];
        if (iterations) iterations[b] = sb.iterations[sp];
    } else {
        for (int j = 0; j < n; ++j) out[j] = sd[j];
        error_code[b] = -31;
        if (cost) cost[b] = fail_cost;
        if (iterations) iterations[b] = fail_it;
    }
}

// packed[b] = joints[n], cost, error_code, iterations: the row each rank contributes to the all-gather
__global__ void pack_results_kernel(int64_t B, int n, const double* __restrict__ solution, const double* __restrict__ cost,
                                    const int32_t* __restrict__ error_code, const int32_t* __restrict__ iterations,
                                    double* __restrict__ packed) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = n + 3;
    if (idx >= B * w) return;
    const int64_t b = idx / w;
    const int k = (int)(idx - b * w);
    double v;
    if (k < n) v = solution[b * n + k];
    else if (k == n) v = cost[b];
    else if (k == n + 1) v = (double)error_code[b];
    else v = (double)iterations[b];
    packed[idx] = v;
}

// which %smid values exist on this device (they have holes where SMs are fused off)
__global__ void sm_discover_kernel(int* seen, int spin_cycles) {
    if (threadIdx.x == 0) {
        seen[sm_id() & (kSmDenseSize - 1)] = 1;
        const long long t0 = clock64();
        while (clock64() - t0 < spin_cycles) {}
    }
}

__global__ void fp64_peak_kernel(double* sink, int iters) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3, a4 = a0 + 4e-3,
           a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 0.9999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) sink[0] = r;
}

}  // namespace

int memetic_max_lanes_per_elite(int E) {
    int L = 32 / E;
    if (L > 16) L = 16;
    return L < 1 ? 1 : L;
}

MemeticShape memetic_shape(int n, int T, int P, int E, int lanes_per_elite) {
    MemeticShape s;
    int L = lanes_per_elite < 1 ? 1 : lanes_per_elite;
    if (L > memetic_max_lanes_per_elite(E)) L = memetic_max_lanes_per_elite(E);
    s.lanes_per_elite = L;
    s.problems_per_warp = 32 / (E * L);
    s.warps = L == 1 ? kWarpsPerBlockBulk : kWarpsPerBlock;
    // the wide flavour carves its shared memory for the largest number of problems per warp it may be given
    const int pw_carve = L == 1 ? s.problems_per_warp : wide_problems_per_warp_max(E);
    // a large population may not leave room for 16 (resp. 4) warps' worth of shared memory
    while (s.warps > 1 && s.warps * warp_smem_bytes(n, P, pw_carve, T) > (L == 1 ? 200 : 112) * 1024) --s.warps;
    s.threads = 32 * s.warps;
    s.smem = s.warps * warp_smem_bytes(n, P, pw_carve, T);
    return s;
}

GenerationPlan plan_generations(int n, int T, int P, int E, int64_t n_sub, int sm_count, long long wide_warps_per_sm,
                                bool allow_persistent) {
    GenerationPlan g;
    const MemeticShape t = memetic_shape(n, T, P, E, 1);
    g.threads_t = t.threads;
    g.smem_t = t.smem;
    // grids of resident CTAs (1 per SM for the throughput flavour, 3 for the wide one): the CTAs claim their work from
    // the per-SM queues, in as many rounds as it takes.  The units are dealt over all the SMs, so a batch of less than
    // a wave still gets a CTA on every SM it has a unit for (whose surplus warps idle).
    int64_t blocks_t = (n_sub + t.problems_per_warp - 1) / t.problems_per_warp;
    g.wave_ctas = sm_count;
    if (const char* env = std::getenv("PIK_WAVE_CTAS")) g.wave_ctas = std::atoi(env) > 0 ? std::atoi(env) : g.wave_ctas;  // (tests)
    if (blocks_t > (int64_t)g.wave_ctas) blocks_t = (int64_t)g.wave_ctas;
    g.blocks_t = (unsigned)blocks_t;
    g.lanes_max = memetic_max_lanes_per_elite(E);
    // largest power of two the doubling of lanes_for can reach
    int lw = 1;
    while (lw * 2 <= g.lanes_max) lw *= 2;
    const MemeticShape w = memetic_shape(n, T, P, E, lw);
    g.threads_w = w.threads;
    g.smem_w = w.smem;
    const int64_t units_w = (n_sub + w.problems_per_warp - 1) / w.problems_per_warp;
    int64_t blocks_w = (units_w + w.warps - 1) / w.warps;
    if (blocks_w > (int64_t)3 * sm_count) blocks_w = (int64_t)3 * sm_count;
    g.blocks_w = (unsigned)blocks_w;
    g.wide_units_max = (int)(blocks_w * w.warps);
    g.wide_capacity_lanes = (long long)sm_count * wide_warps_per_sm * 32;
    if (g.lanes_max < 2) g.wide_capacity_lanes = 0;
    // (only once every warp has an SM sub-partition to itself: a launch per generation re-spreads the
    // survivors over the SMs, a persistent warp stays where it started)
    g.persistent_units_max = allow_persistent ? sm_count * 4 : 0;
    const int l0 = lanes_for(n_sub, E, g.lanes_max, g.wide_capacity_lanes, g.wide_units_max);
    g.use_wide = lanes_for(1, E, g.lanes_max, g.wide_capacity_lanes, g.wide_units_max) > 1;
    g.use_throughput = l0 == 1;
    g.first_launch_runs_all = l0 > 1 && 32 / (E * l0) == 1 && n_sub <= g.persistent_units_max;
    return g;
}

size_t gd_local_smem_bytes(int n, int T) { return (size_t)kWarpsPerBlock * (5 * n + 12 + 7 * T) * kS * sizeof(double); }

// Compiled chain signatures (see StaticSpec).  kinds nibble: X 0, Y 1, Z 2, general 3, prismatic 4.
// Every signature is compiled in two flavours: throughput (generation launches with one lane per elite,
// gd_local, init) and wide (generation launches with several lanes per elite).
// Franka Panda and every 7-joint DH-style arm: all joints about +z, every joint origin a rotation about x
// (alpha_i = +-pi/2 or 0), tool frame a rotation about z
constexpr unsigned long long kKindsAllZ7 = 0x2222222ull;
using SpecAllZ7T = StaticSpec<7, kKindsAllZ7, true, false, kOrgRotX, kOrgRotZ, true>;
using SpecAllZ7W = StaticSpec<7, kKindsAllZ7, true, true, kOrgRotX, kOrgRotZ, true>;
using GenericSpecW = PatternSpec<kOrgGeneral, kOrgGeneral, true>;

static bool origins_have_pattern(const DevRobot& rb, int origin_cls, int tip_cls) {
    bool ok = origin_has_pattern(rb.R[rb.n], tip_cls);
    for (int j = 1; j < rb.n; ++j) ok = ok && origin_has_pattern(rb.R[j], origin_cls);
    return ok;
}

int select_spec(const DevRobot& rb) {
    if (rb.is_tree) return kSpecTree;
    // PIK_GENERIC_ONLY: the generic kernels; PIK_NO_STATIC: no compile-time n / kinds (pattern kernels only)
    if (std::getenv("PIK_GENERIC_ONLY")) return kSpecGeneric;
    unsigned long long kinds = 0;
    for (int j = 0; j < rb.n; ++j) kinds |= (unsigned long long)(rb.kind[j] & 15) << (4 * j);
    bool unit_sign = true;
    for (int j = 0; j < rb.n; ++j) unit_sign = unit_sign && rb.sign[j] == 1.0;
    if (rb.n == 7 && kinds == kKindsAllZ7 && rb.has_tip && unit_sign && !std::getenv("PIK_NO_STATIC") &&
        origins_have_pattern(rb, SpecAllZ7T::origin_cls, SpecAllZ7T::tip_cls))
        return kSpecAllZ7;
    // the origin-pattern kernels carry no general-axis joint path
    for (int j = 0; j < rb.n; ++j)
        if (rb.kind[j] == kRevGeneral || rb.kind[j] == kPrismatic) return kSpecGeneric;
    if (origins_have_pattern(rb, kOrgIdentity, kOrgIdentity)) return kSpecOrgIdentity;
    if (origins_have_pattern(rb, kOrgRotX, kOrgGeneral)) return kSpecOrgRotX;
    if (origins_have_pattern(rb, kOrgRotY, kOrgGeneral)) return kSpecOrgRotY;
    return kSpecGeneric;
}

// CALL sees S = the throughput flavour of the signature (wide = false) or the wide flavour
#define PIK_DISPATCH_SPEC(spec, wide, CALL)                                                            \
    switch ((spec) * 2 + ((wide) ? 1 : 0)) {                                                           \
        case kSpecAllZ7 * 2: { using S = SpecAllZ7T; CALL; break; }                                    \
        case kSpecAllZ7 * 2 + 1: { using S = SpecAllZ7W; CALL; break; }                                \
        case kSpecOrgIdentity * 2: { using S = PatternSpec<kOrgIdentity, kOrgIdentity, false>; CALL; break; } \
        case kSpecOrgIdentity * 2 + 1: { using S = PatternSpec<kOrgIdentity, kOrgIdentity, true>; CALL; break; } \
        case kSpecOrgRotX * 2: { using S = PatternSpec<kOrgRotX, kOrgGeneral, false>; CALL; break; }    \
        case kSpecOrgRotX * 2 + 1: { using S = PatternSpec<kOrgRotX, kOrgGeneral, true>; CALL; break; } \
        case kSpecOrgRotY * 2: { using S = PatternSpec<kOrgRotY, kOrgGeneral, false>; CALL; break; }    \
        case kSpecOrgRotY * 2 + 1: { using S = PatternSpec<kOrgRotY, kOrgGeneral, true>; CALL; break; } \
        case kSpecGeneric * 2 + 1: { using S = GenericSpecW; CALL; break; }                            \
        case kSpecTree * 2: { using S = TreeSpec<false>; CALL; break; }                                \
        case kSpecTree * 2 + 1: { using S = TreeSpec<true>; CALL; break; }                             \
        default: { using S = GenericSpec; CALL; break; }                                               \
    }

template <class S>
static cudaError_t configure_spec(bool wide) {
    cudaError_t e = cudaFuncSetAttribute(memetic_generation_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess || wide) return e;
    e = cudaFuncSetAttribute(memetic_init_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gd_local_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

cudaError_t configure_kernels() {
    cudaError_t e = cudaFuncSetAttribute(eval_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int spec = 0; spec < kSpecCount && e == cudaSuccess; ++spec) {
        PIK_DISPATCH_SPEC(spec, false, e = configure_spec<S>(false));
        if (e == cudaSuccess) PIK_DISPATCH_SPEC(spec, true, e = configure_spec<S>(true));
    }
    return e;
}

cudaError_t upload_constants(cudaStream_t stream, const DevRobot& robot, const DevParams& pr) {
    cudaError_t e = cudaMemcpyToSymbolAsync(c_rb, &robot, sizeof(DevRobot), 0, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbolAsync(c_pr, &pr, sizeof(DevParams), 0, cudaMemcpyHostToDevice, stream);
}

cudaError_t launch_eval_cost(cudaStream_t stream, int n, int64_t B, const double* goal_pose, const double* seed,
                             int64_t seed_stride, const double* q, double* cost, int32_t* is_solution,
                             double* tip_pose) {
    if (B <= 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((B + kThreads - 1) / kThreads);
    const size_t smem = (size_t)kWarpsPerBlock * (n + 7) * kS * sizeof(double);
    eval_cost_kernel<<<blocks, kThreads, smem, stream>>>(B, goal_pose, seed, seed_stride, q, cost, is_solution, tip_pose);
    return cudaGetLastError();
}

cudaError_t launch_gd_local(cudaStream_t stream, int spec, int n, int T, const SolveBuffers& sb, int sm_count) {
    if (sb.B <= 0) return cudaSuccess;
    // resident CTAs only (3 per SM at 168 registers): the lanes pull their problems from a counter
    int64_t blocks64 = (sb.B + kThreads - 1) / kThreads;
    if (blocks64 > (int64_t)3 * sm_count) blocks64 = (int64_t)3 * sm_count;
    const unsigned blocks = (unsigned)blocks64;
    PIK_DISPATCH_SPEC(spec, false, (gd_local_kernel<S><<<blocks, kThreads, gd_local_smem_bytes(n, T), stream>>>(sb)));
    return cudaGetLastError();
}

cudaError_t launch_memetic_init(cudaStream_t stream, int spec, int n, int T, int P, int E, const SolveBuffers& sb) {
    if (sb.B <= 0) return cudaSuccess;
    MemeticShape s = memetic_shape(n, T, P, E, 1);
    s.warps = kWarpsPerBlock;
    s.threads = kThreads;
    s.smem = s.warps * warp_smem_bytes(n, P, s.problems_per_warp, T);
    const int64_t per_block = (int64_t)s.problems_per_warp * s.warps;
    const unsigned blocks = (unsigned)((sb.B + per_block - 1) / per_block);
    PIK_DISPATCH_SPEC(spec, false, (memetic_init_kernel<S><<<blocks, s.threads, s.smem, stream>>>(sb, s.problems_per_warp)));
    return cudaGetLastError();
}

cudaError_t launch_memetic_generation(cudaStream_t stream, int spec, const GenerationPlan& g, const SolveBuffers& sb,
                                      int gen, bool wide) {
    if (wide) {
        PIK_DISPATCH_SPEC(spec, true, (memetic_generation_kernel<S><<<g.blocks_w, g.threads_w, g.smem_w, stream>>>(sb, gen)));
    } else {
        PIK_DISPATCH_SPEC(spec, false, (memetic_generation_kernel<S><<<g.blocks_t, g.threads_t, g.smem_t, stream>>>(sb, gen)));
    }
    return cudaGetLastError();
}

cudaError_t launch_species_pick(cudaStream_t stream, const SolveBuffers& sb, int n, int64_t n_problems, double* solution,
                                int32_t* error_code, double* cost, int32_t* iterations) {
    if (n_problems <= 0) return cudaSuccess;
    species_pick_kernel<<<(unsigned)((n_problems + 127) / 128), 128, 0, stream>>>(sb, n, n_problems, solution, error_code,
                                                                                  cost, iterations);
    return cudaGetLastError();
}

cudaError_t launch_pack_results(cudaStream_t stream, int64_t B, int n, const double* solution, const double* cost,
                                const int32_t* error_code, const int32_t* iterations, double* packed) {
    if (B <= 0) return cudaSuccess;
    const int64_t total = B * (n + 3);
    pack_results_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(B, n, solution, cost, error_code, iterations, packed);
    return cudaGetLastError();
}

cudaError_t discover_sm_ids(cudaStream_t stream, int sm_count, unsigned short* dense) {
    for (int i = 0; i < kSmDenseSize; ++i) dense[i] = (unsigned short)(i % sm_count);  // fallback: still correct, less even
    int* d_seen = nullptr;
    cudaError_t e = cudaMalloc(&d_seen, kSmDenseSize * sizeof(int));
    if (e != cudaSuccess) return e;
    int seen[kSmDenseSize];
    for (int attempt = 0; attempt < 3 && e == cudaSuccess; ++attempt) {
        e = cudaMemsetAsync(d_seen, 0, kSmDenseSize * sizeof(int), stream);
        if (e != cudaSuccess) break;
        sm_discover_kernel<<<sm_count * 32, 32, 0, stream>>>(d_seen, 20000 << attempt);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(seen, d_seen, sizeof(seen), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) break;
        int count = 0;
        for (int i = 0; i < kSmDenseSize; ++i) count += seen[i] ? 1 : 0;
        if (count == sm_count) {
            int next = 0;
            for (int i = 0; i < kSmDenseSize; ++i) dense[i] = (unsigned short)(seen[i] ? next++ : 0);
            break;
        }
    }
    cudaFree(d_seen);
    return e;
}

cudaError_t launch_fp64_peak(cudaStream_t stream, double* sink, int blocks, int threads, int iters) {
    fp64_peak_kernel<<<blocks, threads, 0, stream>>>(sink, iters);
    return cudaGetLastError();
}

}  // namespace pik
