// pik_api.cu -- the C-ABI of include/pik.h: robot tables, solver handles, device buffers, launches.
// No CPU solve path exists here: every entry point that computes runs the CUDA kernels or fails.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/pik.h"
#include "pik_host_robot.h"
#include "pik_internal.h"
#include "pik_kernels.cuh"

using namespace pik;

struct pik_robot {
    DevRobot dev;
    pik_variable vars[kMaxVars];
};

namespace {

struct DeviceArray {
    void* ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct pik_solver {
    pik_robot robot;
    int device = 0;
    int sm_count = 148;
    int spec = 0;  // compiled chain signature matching the robot (select_spec)
    unsigned short sm_dense[kSmDenseSize];  // %smid -> dense SM index on this device
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    int32_t* h_counters = nullptr;            // pinned [2] (PIK_TRACE only)
    unsigned long long* h_stats = nullptr;    // pinned [8]
    // staging for PIK_MEM_HOST calls
    DeviceArray d_goal, d_seed, d_q, d_solution, d_error, d_cost, d_iters, d_issol, d_tip, d_packed, d_gather;
    // solver state
    DeviceArray d_pop, d_order, d_hdr, d_meta, d_active, d_counters, d_sched, d_stats;
    // per-species results and the terminate flags of a multi-species solve (memetic_num_threads > 1)
    DeviceArray d_sub_solution, d_sub_error, d_sub_cost, d_sub_iters, d_group_term;
    pik_stats stats{};
    // the solve in flight between pik_solve_batch_async and pik_solver_wait
    bool in_flight = false;
    bool timed_generations = false;
    bool in_flight_species = false;
    int in_flight_slices = 1;
    // sub-batches of a large global-mode solve: streams of descending priority and their completion events
    std::vector<cudaStream_t> sub_streams;
    std::vector<cudaEvent_t> sub_done;
    int in_flight_status = PIK_OK;
    std::string last_error;
};

namespace {

int fail_cuda(pik_solver* s, cudaError_t e, const char* what) {
    if (s) s->last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? PIK_E_OUT_OF_MEMORY : PIK_E_CUDA;
}

#define PIK_CUDA(s, call)                                         \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return fail_cuda((s), e__, #call); \
    } while (0)

int ensure(pik_solver* s, DeviceArray& a, size_t bytes) {
    if (bytes <= a.bytes) return PIK_OK;
    if (a.ptr) {
        PIK_CUDA(s, cudaStreamSynchronize(s->stream));
        PIK_CUDA(s, cudaFree(a.ptr));
        a.ptr = nullptr;
        a.bytes = 0;
    }
    PIK_CUDA(s, cudaMalloc(&a.ptr, bytes));
    a.bytes = bytes;
    return PIK_OK;
}

void release(DeviceArray& a) {
    if (a.ptr) cudaFree(a.ptr);
    a.ptr = nullptr;
    a.bytes = 0;
}

bool finite_ge(double v, double lo) { return v == v && v >= lo; }

// ---------------------------------------------------------------------------------------------------
// The robot table and the solver parameters are read by the kernels as constant-bank operands of the FP64
// instructions (c_rb / c_pr, pik_device.cuh): one copy per DEVICE.  Calls on different devices never meet.
// Calls on the same device SHARE the copy when their robot table and parameters are identical (two solvers
// pipelining batches on two streams, concurrent searchPositionIK calls of one plugin: the per-call data --
// buffers, RNG key schedule, species count -- travel as kernel arguments); a call with different contents waits,
// on the host, for the device work that still reads the old contents, then replaces them.  No lock is held while
// a solve runs.
// ---------------------------------------------------------------------------------------------------
struct DeviceConstants {
    std::mutex mu;
    std::condition_variable cv;
    bool valid = false;
    DevRobot rb;
    DevParams pr;
    int enqueuing = 0;                  // calls between acquire and commit (their launches are being issued)
    std::vector<cudaEvent_t> readers;   // completion events of committed calls that read the current contents
    cudaEvent_t uploaded = nullptr;     // recorded behind the last upload
};
constexpr int kMaxDevices = 64;
constexpr int kMaxSubBatches = 8;
DeviceConstants g_constants[kMaxDevices];

// Makes (rb, pr) the constants of s->device for the work s is about to enqueue on its stream.
int constants_acquire(pik_solver* s, const DevRobot& rb, const DevParams& pr) {
    if (s->device < 0 || s->device >= kMaxDevices) return PIK_E_INVALID_ARGUMENT;
    DeviceConstants& dc = g_constants[s->device];
    std::unique_lock<std::mutex> lock(dc.mu);
    const bool same = dc.valid && std::memcmp(&dc.rb, &rb, sizeof(rb)) == 0 && std::memcmp(&dc.pr, &pr, sizeof(pr)) == 0;
    if (!same) {
        dc.cv.wait(lock, [&] { return dc.enqueuing == 0; });
        for (cudaEvent_t ev : dc.readers) PIK_CUDA(s, cudaEventSynchronize(ev));
        dc.readers.clear();
        dc.valid = false;
        if (!dc.uploaded) PIK_CUDA(s, cudaEventCreateWithFlags(&dc.uploaded, cudaEventDisableTiming));
        dc.rb = rb;
        dc.pr = pr;
        PIK_CUDA(s, upload_constants(s->stream, dc.rb, dc.pr));
        PIK_CUDA(s, cudaEventRecord(dc.uploaded, s->stream));
        dc.valid = true;
    }
    dc.enqueuing += 1;
    lock.unlock();
    // launches on other streams than the uploading one must not overtake the upload
    const cudaError_t e = cudaStreamWaitEvent(s->stream, dc.uploaded, 0);
    if (e != cudaSuccess) {
        lock.lock();
        dc.enqueuing -= 1;
        dc.cv.notify_all();
        return fail_cuda(s, e, "cudaStreamWaitEvent");
    }
    return PIK_OK;
}

// The call has issued its last launch (done: an event recorded behind it on the call's stream).
void constants_commit(pik_solver* s, cudaEvent_t done) {
    DeviceConstants& dc = g_constants[s->device];
    std::lock_guard<std::mutex> lock(dc.mu);
    dc.enqueuing -= 1;
    if (done) {
        bool present = false;
        for (cudaEvent_t ev : dc.readers) present = present || ev == done;
        if (!present) dc.readers.push_back(done);
    }
    dc.cv.notify_all();
}

void constants_forget(pik_solver* s, cudaEvent_t ev) {
    if (s->device < 0 || s->device >= kMaxDevices) return;
    DeviceConstants& dc = g_constants[s->device];
    std::lock_guard<std::mutex> lock(dc.mu);
    for (size_t i = 0; i < dc.readers.size(); ++i)
        if (dc.readers[i] == ev) {
            dc.readers.erase(dc.readers.begin() + (long)i);
            break;
        }
}

bool trace_enabled() { return std::getenv("PIK_TRACE") != nullptr; }

// PIK_TRACE_GENERATIONS: one launch per generation even for a handful of problems (no persistent launch)
bool trace_each_generation() { return std::getenv("PIK_TRACE_GENERATIONS") != nullptr; }

// PIK_WIDE_WARPS_PER_SM overrides the warps per SM the wide lane mapping aims for (read per call: the parity tests
// run both mappings; 0 = always throughput mode, huge = always the widest mapping the launch can hold).  12 = 3 CTAs
// of the wide flavour, which runs at 168 registers without spills.
long long wide_warps_per_sm() {
    const char* env = std::getenv("PIK_WIDE_WARPS_PER_SM");
    return env ? std::atoll(env) : 12ll;
}

}  // namespace

extern "C" {

int pik_version(void) { return 100; }

const char* pik_status_string(int status) {
    switch (status) {
        case PIK_OK: return "ok";
        case PIK_E_INVALID_ARGUMENT: return "invalid argument";
        case PIK_E_INVALID_ROBOT: return "invalid robot description";
        case PIK_E_INVALID_PARAMS: return "parameter validation failed";
        case PIK_E_CUDA: return "CUDA error";
        case PIK_E_NO_DEVICE: return "no CUDA device";
        case PIK_E_OUT_OF_MEMORY: return "out of device memory";
        case PIK_E_UNSUPPORTED: return "unsupported";
        case PIK_E_NCCL: return "NCCL error (pik_comm_last_error)";
        case PIK_E_BUSY: return "a solve is in flight on this solver";
        default: return "unknown status";
    }
}

void pik_params_default(pik_params* p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->mode = PIK_MODE_GLOBAL;
    p->gd_step_size = 0.0001;
    p->gd_max_iters = 100;
    p->gd_min_cost_delta = 1.0e-12;
    p->position_threshold = 0.001;
    p->orientation_threshold = 0.001;
    p->approximate_solution_position_threshold = 0.05;
    p->approximate_solution_orientation_threshold = 0.05;
    p->approximate_solution_joint_threshold = 0.0;
    p->approximate_solution_cost_threshold = 0.0;
    p->cost_threshold = 0.001;
    p->position_scale = 1.0;
    p->rotation_scale = 0.5;
    p->center_joints_weight = 0.0;
    p->avoid_joint_limits_weight = 0.0;
    p->minimal_displacement_weight = 0.0;
    p->stop_optimization_on_valid_solution = 1;
    p->memetic_num_threads = 1;
    p->memetic_stop_on_first_solution = 1;
    p->memetic_population_size = 16;
    p->memetic_elite_size = 4;
    p->memetic_wipeout_fitness_tol = 0.00001;
    p->memetic_max_generations = 100;
    p->memetic_gd_max_iters = 25;
    p->memetic_gd_max_time = 0.005;
    p->return_approximate_solution = 0;
    p->rng_seed = 0x5EED;
}

int pik_params_validate(const pik_params* p) {
    if (!p) return PIK_E_INVALID_ARGUMENT;
    // src/pick_ik_parameters.yaml validators
    if (p->mode != PIK_MODE_GLOBAL && p->mode != PIK_MODE_LOCAL) return PIK_E_INVALID_PARAMS;
    if (!finite_ge(p->gd_step_size, 1.0e-12)) return PIK_E_INVALID_PARAMS;
    if (p->gd_max_iters < 1) return PIK_E_INVALID_PARAMS;
    if (!finite_ge(p->gd_min_cost_delta, 1.0e-64)) return PIK_E_INVALID_PARAMS;
    const double ge0[] = {p->position_threshold, p->orientation_threshold,
                          p->approximate_solution_position_threshold, p->approximate_solution_orientation_threshold,
                          p->approximate_solution_joint_threshold, p->approximate_solution_cost_threshold,
                          p->cost_threshold, p->position_scale, p->rotation_scale, p->center_joints_weight,
                          p->avoid_joint_limits_weight, p->minimal_displacement_weight,
                          p->memetic_wipeout_fitness_tol, p->memetic_gd_max_time};
    for (double v : ge0)
        if (!finite_ge(v, 0.0)) return PIK_E_INVALID_PARAMS;
    if (p->memetic_num_threads < 1 || p->memetic_population_size < 1 || p->memetic_elite_size < 1 ||
        p->memetic_max_generations < 1 || p->memetic_gd_max_iters < 1)
        return PIK_E_INVALID_PARAMS;
    // beyond the YAML: the reference leaves elite > population unguarded; table limits
    if (p->memetic_elite_size > p->memetic_population_size) return PIK_E_INVALID_PARAMS;
    if (p->memetic_elite_size > kMaxElites || p->memetic_population_size > kMaxPopulation) return PIK_E_UNSUPPORTED;
    return PIK_OK;
}

int pik_robot_create_tree(const pik_joint_desc* joints, int32_t n_joints, const int32_t* parent, const int32_t* tip_joint,
                          int32_t n_tips, const int32_t* mimic_of, const double* mimic_factor, const double* mimic_offset,
                          pik_robot** out) {
    if (!out) return PIK_E_INVALID_ARGUMENT;
    *out = nullptr;
    if (!joints || n_joints <= 0 || !tip_joint || n_tips <= 0) return PIK_E_INVALID_ROBOT;
    if (mimic_of && (!mimic_factor || !mimic_offset)) return PIK_E_INVALID_ARGUMENT;
    for (int j = 0; j < n_joints; ++j) {
        const pik_joint_desc& jd = joints[j];
        if (jd.type == PIK_JOINT_FIXED) continue;
        if (jd.type < PIK_JOINT_FIXED || jd.type > PIK_JOINT_PLANAR) return PIK_E_INVALID_ROBOT;
        if (jd.type == PIK_JOINT_REVOLUTE || jd.type == PIK_JOINT_PRISMATIC) {
            const double a2 = jd.axis[0] * jd.axis[0] + jd.axis[1] * jd.axis[1] + jd.axis[2] * jd.axis[2];
            if (!(std::fabs(a2 - 1.0) < 1e-6)) return PIK_E_INVALID_ROBOT;
        }
        if (!(jd.min_position <= jd.max_position)) return PIK_E_INVALID_ROBOT;
    }
    pik_robot* r = new (std::nothrow) pik_robot;
    if (!r) return PIK_E_OUT_OF_MEMORY;
    double rcp[kMaxVars] = {0};
    const int rc = build_dev_robot_tree(joints, n_joints, parent, tip_joint, n_tips, mimic_of, mimic_factor, mimic_offset,
                                        &r->dev, rcp);
    if (rc != PIK_OK) {
        delete r;
        return rc;
    }
    for (int i = 0; i < r->dev.n; ++i) {
        pik_variable& v = r->vars[i];
        v.min = r->dev.vmin[i];
        v.max = r->dev.vmax[i];
        v.mid = r->dev.vmid[i];
        v.half_span = r->dev.vhalf[i];
        v.max_velocity_rcp = rcp[i];
        v.minimal_displacement_factor = r->dev.vfac[i];
        v.bounded = r->dev.bounded[i];
        v.pad_ = 0;
    }
    *out = r;
    return PIK_OK;
}

int pik_robot_create(const pik_joint_desc* joints, int32_t n_joints, pik_robot** out) {
    if (!out) return PIK_E_INVALID_ARGUMENT;
    *out = nullptr;
    if (!joints || n_joints <= 0) return PIK_E_INVALID_ROBOT;
    const int32_t tip = n_joints - 1;
    return pik_robot_create_tree(joints, n_joints, nullptr, &tip, 1, nullptr, nullptr, nullptr, out);
}

int32_t pik_robot_num_tips(const pik_robot* robot) { return robot ? robot->dev.n_tips : 0; }

void pik_robot_destroy(pik_robot* robot) { delete robot; }

int32_t pik_robot_num_variables(const pik_robot* robot) { return robot ? robot->dev.n : 0; }

int pik_robot_get_variable(const pik_robot* robot, int32_t i, pik_variable* out) {
    if (!robot || !out || i < 0 || i >= robot->dev.n) return PIK_E_INVALID_ARGUMENT;
    *out = robot->vars[i];
    return PIK_OK;
}

const char* pik_robot_chain_signature(const pik_robot* robot) {
    if (!robot) return "";
    switch (select_spec(robot->dev)) {
        case kSpecAllZ7: return "all-z 7R, x-rotation origins (static)";
        case kSpecOrgIdentity: return "identity origins";
        case kSpecOrgRotX: return "x-rotation origins";
        case kSpecOrgRotY: return "y-rotation origins";
        case kSpecTree: return "tree";
        default: return "generic";
    }
}

int pik_robot_is_valid_configuration(const pik_robot* robot, const double* q) {
    if (!robot || !q) return 0;
    for (int i = 0; i < robot->dev.n; ++i) {
        const pik_variable& v = robot->vars[i];
        if (!(!v.bounded || (q[i] <= v.max && q[i] >= v.min))) return 0;  // robot.cpp:32-34
    }
    return 1;
}

// Synthetic workload generator: q*_b ~ U(min_i, max_i) per variable (unbounded: U(-pi, pi)) from the
// Philox4x32-10 stream (key gen_seed; counter = (block i >> 1, 0, kStreamTarget << 28, b), words 2 (i & 1), +1),
// host arithmetic only (+ - *), so a C caller, the Python harness and the CPU oracle draw the same targets.
int pik_random_configurations(const pik_robot* robot, uint64_t gen_seed, int64_t first_problem_index, int64_t B,
                              double* q) {
    if (!robot || !q || B < 0 || first_problem_index < 0 || first_problem_index + B > (int64_t)0xffffffffll)
        return PIK_E_INVALID_ARGUMENT;
    const int n = robot->dev.n;
    const uint32_t k0 = (uint32_t)gen_seed, k1 = (uint32_t)(gen_seed >> 32);
    for (int64_t b = 0; b < B; ++b) {
        uint32_t w[4] = {0, 0, 0, 0};
        for (int i = 0; i < n; ++i) {
            if ((i & 1) == 0) {
                uint32_t c0 = (uint32_t)(i >> 1), c1 = 0, c2 = 3u << 28, c3 = (uint32_t)(first_problem_index + b);
                uint32_t ka = k0, kb = k1;
                for (int round = 0; round < 10; ++round) {
                    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
                    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ ka, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ kb;
                    c1 = (uint32_t)p1;
                    c3 = (uint32_t)p0;
                    c0 = n0;
                    c2 = n2;
                    ka += 0x9E3779B9u;
                    kb += 0xBB67AE85u;
                }
                w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
            }
            const uint32_t lo = w[2 * (i & 1)], hi = w[2 * (i & 1) + 1];
            const double u = (double)(((((uint64_t)hi) << 32) | (uint64_t)lo) >> 11) * 0x1.0p-53;
            const pik_variable& v = robot->vars[i];
            const double a = v.bounded ? v.min : -M_PI, bb = v.bounded ? v.max : M_PI;
            q[b * n + i] = a + (bb - a) * u;
        }
    }
    return PIK_OK;
}

int pik_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int pik_solver_create(const pik_robot* robot, int32_t device, void* stream, pik_solver** out) {
    if (!out) return PIK_E_INVALID_ARGUMENT;
    *out = nullptr;
    if (!robot) return PIK_E_INVALID_ARGUMENT;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return PIK_E_NO_DEVICE;
    if (device < 0 || device >= count) return PIK_E_INVALID_ARGUMENT;
    pik_solver* s = new (std::nothrow) pik_solver;
    if (!s) return PIK_E_OUT_OF_MEMORY;
    s->robot = *robot;
    s->device = device;
    s->spec = select_spec(robot->dev);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) {
        if (stream) {
            s->stream = static_cast<cudaStream_t>(stream);
        } else {
            e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
            s->own_stream = true;
        }
    }
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev2);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev3);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&s->h_counters), 2 * sizeof(int32_t), cudaHostAllocDefault);
    if (e == cudaSuccess)
        e = cudaHostAlloc(reinterpret_cast<void**>(&s->h_stats), kMaxSubBatches * 8 * sizeof(unsigned long long), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = configure_kernels();
    if (e == cudaSuccess) {
        // one discovery per device and process
        static std::mutex mu;
        static bool known[kMaxDevices];
        static unsigned short table[kMaxDevices][kSmDenseSize];
        std::lock_guard<std::mutex> lock(mu);
        if (device < kMaxDevices && !known[device]) {
            e = discover_sm_ids(s->stream, s->sm_count, table[device]);
            known[device] = e == cudaSuccess;
        }
        if (e == cudaSuccess && device < kMaxDevices) std::memcpy(s->sm_dense, table[device], sizeof(s->sm_dense));
    }
    if (e != cudaSuccess) {
        std::fprintf(stderr, "pik_solver_create: %s\n", cudaGetErrorString(e));
        pik_solver_destroy(s);
        return PIK_E_CUDA;
    }
    *out = s;
    return PIK_OK;
}

void pik_solver_destroy(pik_solver* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->ev1) constants_forget(s, s->ev1);
    for (cudaStream_t st : s->sub_streams) {
        cudaStreamSynchronize(st);
        cudaStreamDestroy(st);
    }
    for (cudaEvent_t ev : s->sub_done) cudaEventDestroy(ev);
    DeviceArray* arrays[] = {&s->d_goal, &s->d_seed, &s->d_q, &s->d_solution, &s->d_error, &s->d_cost, &s->d_iters,
                             &s->d_issol, &s->d_tip, &s->d_packed, &s->d_gather, &s->d_pop, &s->d_order, &s->d_hdr, &s->d_meta, &s->d_active,
                             &s->d_counters, &s->d_sched, &s->d_stats, &s->d_sub_solution, &s->d_sub_error, &s->d_sub_cost, &s->d_sub_iters,
                             &s->d_group_term};
    for (DeviceArray* a : arrays) release(*a);
    if (s->h_counters) cudaFreeHost(s->h_counters);
    if (s->h_stats) cudaFreeHost(s->h_stats);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->ev2) cudaEventDestroy(s->ev2);
    if (s->ev3) cudaEventDestroy(s->ev3);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

}  // extern "C"

namespace {

// Collects the outcome of the solve in flight: waits for the stream, reads the counters and the event times.
int finish_solve(pik_solver* s) {
    if (!s->in_flight) return PIK_OK;
    s->in_flight = false;
    if (s->in_flight_status != PIK_OK) return s->in_flight_status;
    PIK_CUDA(s, cudaSetDevice(s->device));
    PIK_CUDA(s, cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    PIK_CUDA(s, cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->stats.device_ms = ms;
    if (s->timed_generations) {
        PIK_CUDA(s, cudaEventElapsedTime(&ms, s->ev2, s->ev3));
        s->stats.generation_ms = ms;
    }
    s->stats.problem_generations = s->stats.gd_steps = s->stats.solved = 0;
    for (int k = 0; k < s->in_flight_slices; ++k) {
        s->stats.problem_generations += (int64_t)s->h_stats[8 * k + 0];
        s->stats.gd_steps += (int64_t)s->h_stats[8 * k + 1];
        s->stats.solved += (int64_t)s->h_stats[8 * k + 2];
    }
    // (the pick over species counts the problems with a value in slot 4 of the first block)
    if (s->in_flight_species) s->stats.solved = (int64_t)s->h_stats[4];
    return PIK_OK;
}

// Sub-batches of a global-mode solve (EXPERIMENT, off by default: PIK_SUB_BATCHES=K).  A batch spends most of its time
// in generations whose active list no longer fills the device, so the batch can be cut into K contiguous sub-batches,
// each a complete solve of its own (own active lists, counters and launches; RNG streams keyed by the global
// problem index, so the results are those of the undivided batch bit for bit) on its own stream of descending
// priority, in the hope that the throughput-bound head of one sub-batch runs in the shadow of the under-filled
// generations of another.  Measured on B200 (profiles/r02_notes.md): it does not pay -- 68.3 ms undivided, 76.2 ms
// for K = 2, 100.6 ms for K = 4 per 65 536 Panda poses.  CTAs are not preempted: the latency-critical launches of the
// sub-batch that is ahead queue behind the resident (persistent) CTAs of the others, and every launch, including
// the ones that find nothing to do, needs SM slots to run its grid.  Two whole batches in flight on two solvers
// (pik_solve_batch_async), whose phases drift apart by themselves, do gain (bench.py "pipelined").
int sub_batch_count(int64_t n_sub, bool trace) {
    (void)n_sub;
    if (trace) return 1;
    if (const char* env = std::getenv("PIK_SUB_BATCHES")) {
        const long v = std::atol(env);
        if (v >= 1) return (int)(v > kMaxSubBatches ? kMaxSubBatches : v);
    }
    return 1;
}

int ensure_sub_streams(pik_solver* s, int K) {
    if ((int)s->sub_streams.size() >= K) return PIK_OK;
    int lo = 0, hi = 0;  // numerically lower = higher priority
    PIK_CUDA(s, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    while ((int)s->sub_streams.size() < K) {
        const int k = (int)s->sub_streams.size();
        int prio = hi + k;
        if (prio > lo) prio = lo;
        cudaStream_t st = nullptr;
        cudaEvent_t ev = nullptr;
        PIK_CUDA(s, cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio));
        const cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            cudaStreamDestroy(st);
            return fail_cuda(s, e, "cudaEventCreateWithFlags");
        }
        s->sub_streams.push_back(st);
        s->sub_done.push_back(ev);
    }
    return PIK_OK;
}

// Enqueues a whole solve on the solver's stream: H2D copies (PIK_MEM_HOST), the init kernel, every generation
// launch (the kernels size and skip themselves from the device-side counters: nothing is read back in between),
// the pick over species, the D2H copies.  keep_on_device: the results stay in the solver's device buffers
// (d_solution, d_error, d_cost, d_iters) for a follow-up on the same stream (the sharded gather) and the output
// pointers are ignored.
int enqueue_solve(pik_solver* s, const pik_params* params, int64_t B, int64_t first_problem_index,
                  const double* goal_pose, const double* seed, int64_t seed_stride, double* solution,
                  int32_t* error_code, double* cost, int32_t* iterations, int32_t memory, bool keep_on_device) {
    if (!s) return PIK_E_INVALID_ARGUMENT;
    if (s->in_flight) return PIK_E_BUSY;
    s->last_error.clear();
    if (!params || B < 0 || !goal_pose || !seed) return PIK_E_INVALID_ARGUMENT;
    if (!keep_on_device && (!solution || !error_code)) return PIK_E_INVALID_ARGUMENT;
    const int n = s->robot.dev.n;
    const int T = s->robot.dev.n_tips;
    if (seed_stride != 0 && seed_stride != n) return PIK_E_INVALID_ARGUMENT;
    if (memory != PIK_MEM_HOST && memory != PIK_MEM_DEVICE) return PIK_E_INVALID_ARGUMENT;
    if (B > (int64_t)1 << 30 || first_problem_index < 0 || first_problem_index + B > (int64_t)0xffffffffll)
        return PIK_E_INVALID_ARGUMENT;
    int rc = pik_params_validate(params);
    if (rc != PIK_OK) return rc;
    const bool global = params->mode == PIK_MODE_GLOBAL;
    const int S = global ? params->memetic_num_threads : 1;  // species; ik_gradient has none
    const int64_t n_sub = B * S;
    if (n_sub > (int64_t)1 << 30) return PIK_E_INVALID_ARGUMENT;
    std::memset(&s->stats, 0, sizeof(s->stats));
    s->stats.problems = B;
    s->timed_generations = false;
    s->in_flight_species = S > 1;
    s->in_flight_slices = 1;
    s->in_flight_status = PIK_OK;
    if (B == 0) return PIK_OK;
    PIK_CUDA(s, cudaSetDevice(s->device));
    DevParams pr = make_dev_params(*params);
    const int P = pr.P;
    const size_t seed_elems = seed_stride ? (size_t)B * n : (size_t)n;
    const bool trace = trace_enabled();
    const int K = global ? sub_batch_count(n_sub, trace) : 1;
    if (K > 1 && (rc = ensure_sub_streams(s, K)) != PIK_OK) return rc;
    s->in_flight_slices = K;

    // the whole batch as the kernels address it; the sub-batches below are windows into these arrays
    SolveBuffers sb;
    std::memset(&sb, 0, sizeof(sb));
    sb.B = n_sub;
    sb.first_problem_index = first_problem_index;
    sb.seed_stride = seed_stride;
    sb.n_species = S;
    sb.stop_on_first = params->memetic_stop_on_first_solution ? 1 : 0;
    make_round_keys(params->rng_seed, sb.round_key);
    if (memory == PIK_MEM_HOST) {
        if ((rc = ensure(s, s->d_goal, (size_t)B * 7 * T * 8)) || (rc = ensure(s, s->d_seed, seed_elems * 8))) return rc;
        sb.goal_pose = static_cast<double*>(s->d_goal.ptr);
        sb.seed = static_cast<double*>(s->d_seed.ptr);
    } else {
        sb.goal_pose = goal_pose;
        sb.seed = seed;
    }
    // where the per-problem results go on the device
    double* r_solution = solution;
    int32_t* r_error = error_code;
    double* r_cost = cost;
    int32_t* r_iters = iterations;
    if (memory == PIK_MEM_HOST || keep_on_device) {
        if ((rc = ensure(s, s->d_solution, (size_t)B * n * 8)) || (rc = ensure(s, s->d_error, (size_t)B * 4)) ||
            (rc = ensure(s, s->d_cost, (size_t)B * 8)) || (rc = ensure(s, s->d_iters, (size_t)B * 4)))
            return rc;
        r_solution = static_cast<double*>(s->d_solution.ptr);
        r_error = static_cast<int32_t*>(s->d_error.ptr);
        r_cost = static_cast<double*>(s->d_cost.ptr);
        r_iters = static_cast<int32_t*>(s->d_iters.ptr);
    }
    if (S > 1) {
        // the kernels write per-species results; species_pick_kernel reduces them to the per-problem results
        if ((rc = ensure(s, s->d_sub_solution, (size_t)n_sub * n * 8)) || (rc = ensure(s, s->d_sub_error, (size_t)n_sub * 4)) ||
            (rc = ensure(s, s->d_sub_cost, (size_t)n_sub * 8)) || (rc = ensure(s, s->d_sub_iters, (size_t)n_sub * 4)) ||
            (rc = ensure(s, s->d_group_term, (size_t)B * 4)))
            return rc;
        sb.solution = static_cast<double*>(s->d_sub_solution.ptr);
        sb.error_code = static_cast<int32_t*>(s->d_sub_error.ptr);
        sb.cost = static_cast<double*>(s->d_sub_cost.ptr);
        sb.iterations = static_cast<int32_t*>(s->d_sub_iters.ptr);
        sb.group_term = static_cast<int32_t*>(s->d_group_term.ptr);
    } else {
        sb.solution = r_solution;
        sb.error_code = r_error;
        sb.cost = r_cost;
        sb.iterations = r_iters;
    }
    if ((rc = ensure(s, s->d_stats, (size_t)K * 8 * sizeof(unsigned long long)))) return rc;
    sb.stats = static_cast<unsigned long long*>(s->d_stats.ptr);
    GenerationPlan plan;
    std::memset(&plan, 0, sizeof(plan));
    // Launches of the throughput flavour may pass part of their list on (memetic_generation_kernel): a problem then
    // lags behind the launch index, by at most the number of launches allowed to do so, and the solve gets that
    // many launches more.  Species that may terminate one another advance strictly one generation per launch.
    int defer_launches = 12;
    if (const char* env = std::getenv("PIK_DEFER_LAUNCHES")) defer_launches = std::atoi(env);
    if (defer_launches < 0 || !global || trace_each_generation() || (S > 1 && params->memetic_stop_on_first_solution)) defer_launches = 0;
    if (defer_launches > pr.max_generations) defer_launches = pr.max_generations;
    pr.defer_launches = defer_launches;
    const int n_launches = pr.max_generations + defer_launches;
    const size_t n_counters = (size_t)n_launches + 2;
    const size_t n_sched = ((size_t)n_launches + 1) * ((size_t)s->sm_count + 2);
    const int64_t slice_problems = (B + K - 1) / K;  // problems per sub-batch (the last one may be shorter)
    if (global) {
        const size_t F = 2 * (size_t)n + 2;
        if ((rc = ensure(s, s->d_pop, 2 * (size_t)n_sub * F * P * 8)) || (rc = ensure(s, s->d_order, 2 * (size_t)n_sub * P * 2)) ||
            (rc = ensure(s, s->d_hdr, (size_t)n_sub * (n + 2) * 8)) || (rc = ensure(s, s->d_meta, (size_t)n_sub * sizeof(ProblemMeta))) ||
            (rc = ensure(s, s->d_active, 2 * (size_t)n_sub * 4)) || (rc = ensure(s, s->d_counters, (size_t)K * n_counters * 4)) ||
            (rc = ensure(s, s->d_sched, (size_t)K * n_sched * 4)))
            return rc;
        sb.pop = static_cast<double*>(s->d_pop.ptr);
        sb.order = static_cast<uint16_t*>(s->d_order.ptr);
        sb.hdr = static_cast<double*>(s->d_hdr.ptr);
        sb.meta = static_cast<ProblemMeta*>(s->d_meta.ptr);
        sb.active = static_cast<int32_t*>(s->d_active.ptr);
        sb.counters = static_cast<int32_t*>(s->d_counters.ptr);
        sb.sched = static_cast<int32_t*>(s->d_sched.ptr);
        // one plan, sized for the largest sub-batch, serves every sub-batch
        // (species that may be terminated by one another advance one generation per launch: no persistent launch)
        const bool lockstep_species = S > 1 && params->memetic_stop_on_first_solution;
        plan = plan_generations(n, T, P, pr.E, slice_problems * S, s->sm_count, wide_warps_per_sm(),
                                !trace_each_generation() && !lockstep_species);
        pr.sm_count = s->sm_count;
        pr.lanes_max = plan.lanes_max;
        pr.wide_capacity_lanes = plan.wide_capacity_lanes;
        pr.wide_units_max = plan.wide_units_max;
        pr.persistent_units_max = plan.persistent_units_max;
        pr.wave_ctas = plan.wave_ctas;
        std::memcpy(pr.sm_dense, s->sm_dense, sizeof(pr.sm_dense));
    }

    // window k of the batch: problems [p0, p0 + pk), sub-problems [p0 S, (p0 + pk) S)
    auto slice = [&](int k) {
        SolveBuffers w = sb;
        const int64_t p0 = (int64_t)k * slice_problems;
        const int64_t pk = (p0 + slice_problems <= B ? slice_problems : B - p0);
        const int64_t s0 = p0 * S;
        const size_t F = 2 * (size_t)n + 2;
        w.B = pk * S;
        w.first_problem_index = first_problem_index + p0;
        w.goal_pose = sb.goal_pose + 7 * (size_t)T * (size_t)p0;
        w.seed = sb.seed + (size_t)seed_stride * (size_t)p0;
        w.solution = sb.solution + (size_t)s0 * n;
        w.error_code = sb.error_code + s0;
        w.cost = sb.cost ? sb.cost + s0 : nullptr;
        w.iterations = sb.iterations ? sb.iterations + s0 : nullptr;
        if (global) {
            w.pop = sb.pop + 2 * (size_t)s0 * F * P;   // [2][w.B][F][P] of its own
            w.order = sb.order + 2 * (size_t)s0 * P;  // [2][w.B][P]
            w.hdr = sb.hdr + (size_t)s0 * (n + 2);
            w.meta = sb.meta + s0;
            w.active = sb.active + 2 * (size_t)s0;     // [2][w.B]
            w.counters = sb.counters + (size_t)k * n_counters;
            w.sched = sb.sched + (size_t)k * n_sched;
            if (sb.group_term) w.group_term = sb.group_term + p0;
        }
        w.stats = sb.stats + (size_t)k * 8;
        w.sm_rotation = K > 1 ? (int32_t)(((int64_t)k * s->sm_count) / K) : 0;
        return w;
    };

    cudaStream_t st = s->stream;
    if ((rc = constants_acquire(s, s->robot.dev, pr)) != PIK_OK) return rc;
    // from here on every exit goes through constants_commit
    auto issue = [&]() -> int {
        PIK_CUDA(s, cudaEventRecord(s->ev0, st));
        if (memory == PIK_MEM_HOST) {
            PIK_CUDA(s, cudaMemcpyAsync(s->d_goal.ptr, goal_pose, (size_t)B * 7 * T * 8, cudaMemcpyHostToDevice, st));
            PIK_CUDA(s, cudaMemcpyAsync(s->d_seed.ptr, seed, seed_elems * 8, cudaMemcpyHostToDevice, st));
        }
        PIK_CUDA(s, cudaMemsetAsync(sb.stats, 0, (size_t)K * 8 * sizeof(unsigned long long), st));
        if (!global) {
            PIK_CUDA(s, launch_gd_local(st, s->spec, n, T, sb, s->sm_count));
            s->stats.kernel_launches += 1;
        } else {
            PIK_CUDA(s, cudaMemsetAsync(sb.counters, 0, (size_t)K * n_counters * sizeof(int32_t), st));
            PIK_CUDA(s, cudaMemsetAsync(sb.sched, 0, (size_t)K * n_sched * sizeof(int32_t), st));
            if (sb.group_term) PIK_CUDA(s, cudaMemsetAsync(sb.group_term, 0, (size_t)B * sizeof(int32_t), st));
            PIK_CUDA(s, cudaEventRecord(s->ev2, st));
            for (int k = 0; k < K; ++k) {
                const SolveBuffers w = slice(k);
                if (w.B <= 0) continue;
                cudaStream_t sk = st;
                if (K > 1) {
                    sk = s->sub_streams[k];
                    PIK_CUDA(s, cudaStreamWaitEvent(sk, s->ev2, 0));  // inputs and cleared counters are in place
                }
                PIK_CUDA(s, launch_memetic_init(sk, s->spec, n, T, P, pr.E, w));
                s->stats.kernel_launches += 1;
                if (K == 1) PIK_CUDA(s, cudaEventRecord(s->ev2, st));
                const int n_gens = plan.first_launch_runs_all ? 1 : n_launches;
                for (int gen = 0; gen < n_gens; ++gen) {
                    if (trace) PIK_CUDA(s, cudaEventRecord(s->ev2, st));
                    if (plan.use_throughput) {
                        PIK_CUDA(s, launch_memetic_generation(sk, s->spec, plan, w, gen, false));
                        s->stats.kernel_launches += 1;
                    }
                    if (plan.use_wide) {
                        PIK_CUDA(s, launch_memetic_generation(sk, s->spec, plan, w, gen, true));
                        s->stats.kernel_launches += 1;
                    }
                    if (k == 0) s->stats.generation_launches += 1;
                    if (trace) {
                        // PIK_TRACE: host-synchronous, one line per generation (experiments only; one sub-batch)
                        PIK_CUDA(s, cudaEventRecord(s->ev3, st));
                        PIK_CUDA(s, cudaMemcpyAsync(s->h_counters, w.counters + gen, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
                        PIK_CUDA(s, cudaStreamSynchronize(st));
                        float gms = 0.f;
                        PIK_CUDA(s, cudaEventElapsedTime(&gms, s->ev2, s->ev3));
                        s->stats.generation_ms += gms;
                        const int n_active = s->h_counters[0];
                        std::fprintf(stderr, "pik gen %3d active %8d lanes %2d  %8.3f ms\n", gen, n_active,
                                     lanes_for(n_active, pr.E, pr.lanes_max, pr.wide_capacity_lanes, pr.wide_units_max), gms);
                        if (s->h_counters[1] == 0) break;
                    }
                }
                if (K > 1) {
                    PIK_CUDA(s, cudaEventRecord(s->sub_done[k], sk));
                    PIK_CUDA(s, cudaStreamWaitEvent(st, s->sub_done[k], 0));
                }
            }
            if (!trace) {
                PIK_CUDA(s, cudaEventRecord(s->ev3, st));
                s->timed_generations = true;
            }
            if (S > 1) {
                PIK_CUDA(s, launch_species_pick(st, sb, n, B, r_solution, r_error, r_cost, r_iters));
                s->stats.kernel_launches += 1;
            }
        }
        if (memory == PIK_MEM_HOST && !keep_on_device) {
            PIK_CUDA(s, cudaMemcpyAsync(solution, r_solution, (size_t)B * n * 8, cudaMemcpyDeviceToHost, st));
            PIK_CUDA(s, cudaMemcpyAsync(error_code, r_error, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
            if (cost) PIK_CUDA(s, cudaMemcpyAsync(cost, r_cost, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
            if (iterations) PIK_CUDA(s, cudaMemcpyAsync(iterations, r_iters, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        }
        PIK_CUDA(s, cudaMemcpyAsync(s->h_stats, sb.stats, (size_t)K * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        PIK_CUDA(s, cudaEventRecord(s->ev1, st));
        return PIK_OK;
    };
    rc = issue();
    if (rc != PIK_OK) {
        // kernels already issued may still read the constants: drain the streams before anyone replaces them
        cudaStreamSynchronize(st);
        for (int k = 0; k < K && K > 1; ++k) cudaStreamSynchronize(s->sub_streams[k]);
        constants_commit(s, nullptr);
        return rc;
    }
    constants_commit(s, s->ev1);
    s->in_flight = true;
    return PIK_OK;
}

}  // namespace

extern "C" {

int pik_solve_batch_async(pik_solver* s, const pik_params* params, int64_t B, int64_t first_problem_index,
                          const double* goal_pose, const double* seed, int64_t seed_stride, double* solution,
                          int32_t* error_code, double* cost, int32_t* iterations, int32_t memory) {
    return enqueue_solve(s, params, B, first_problem_index, goal_pose, seed, seed_stride, solution, error_code, cost,
                         iterations, memory, false);
}

int pik_solver_wait(pik_solver* s) {
    if (!s) return PIK_E_INVALID_ARGUMENT;
    return finish_solve(s);
}

int pik_solver_query(pik_solver* s) {
    if (!s) return PIK_E_INVALID_ARGUMENT;
    if (!s->in_flight) return 1;
    if (cudaSetDevice(s->device) != cudaSuccess) return PIK_E_CUDA;
    const cudaError_t e = cudaEventQuery(s->ev1);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) return 0;
    return fail_cuda(s, e, "cudaEventQuery");
}

int pik_solve_batch(pik_solver* s, const pik_params* params, int64_t B, int64_t first_problem_index,
                    const double* goal_pose, const double* seed, int64_t seed_stride, double* solution,
                    int32_t* error_code, double* cost, int32_t* iterations, int32_t memory) {
    const int rc = enqueue_solve(s, params, B, first_problem_index, goal_pose, seed, seed_stride, solution, error_code,
                                 cost, iterations, memory, false);
    if (rc != PIK_OK) return rc;
    return finish_solve(s);
}

int pik_eval_cost(pik_solver* s, const pik_params* params, int64_t B, const double* goal_pose, const double* seed,
                  int64_t seed_stride, const double* q, double* cost, int32_t* is_solution, double* tip_pose,
                  int32_t memory) {
    if (!s) return PIK_E_INVALID_ARGUMENT;
    if (s->in_flight) return PIK_E_BUSY;
    s->last_error.clear();
    if (!params || B < 0 || !goal_pose || !seed || !q) return PIK_E_INVALID_ARGUMENT;
    const int n = s->robot.dev.n;
    const int T = s->robot.dev.n_tips;
    if (seed_stride != 0 && seed_stride != n) return PIK_E_INVALID_ARGUMENT;
    if (memory != PIK_MEM_HOST && memory != PIK_MEM_DEVICE) return PIK_E_INVALID_ARGUMENT;
    int rc = pik_params_validate(params);
    if (rc != PIK_OK) return rc;
    if (B == 0) return PIK_OK;
    PIK_CUDA(s, cudaSetDevice(s->device));
    const DevParams pr = make_dev_params(*params);
    cudaStream_t st = s->stream;
    const size_t seed_elems = seed_stride ? (size_t)B * n : (size_t)n;
    if (memory == PIK_MEM_HOST &&
        ((rc = ensure(s, s->d_goal, (size_t)B * 7 * T * 8)) || (rc = ensure(s, s->d_seed, seed_elems * 8)) ||
         (rc = ensure(s, s->d_q, (size_t)B * n * 8)) || (rc = ensure(s, s->d_cost, (size_t)B * 8)) ||
         (rc = ensure(s, s->d_issol, (size_t)B * 4)) || (rc = ensure(s, s->d_tip, (size_t)B * 7 * T * 8))))
        return rc;
    if ((rc = constants_acquire(s, s->robot.dev, pr)) != PIK_OK) return rc;
    auto issue = [&]() -> int {
        if (memory == PIK_MEM_DEVICE) {
            PIK_CUDA(s, launch_eval_cost(st, n, B, goal_pose, seed, seed_stride, q, cost, is_solution, tip_pose));
            return PIK_OK;
        }
        PIK_CUDA(s, cudaMemcpyAsync(s->d_goal.ptr, goal_pose, (size_t)B * 7 * T * 8, cudaMemcpyHostToDevice, st));
        PIK_CUDA(s, cudaMemcpyAsync(s->d_seed.ptr, seed, seed_elems * 8, cudaMemcpyHostToDevice, st));
        PIK_CUDA(s, cudaMemcpyAsync(s->d_q.ptr, q, (size_t)B * n * 8, cudaMemcpyHostToDevice, st));
        PIK_CUDA(s, launch_eval_cost(st, n, B, static_cast<double*>(s->d_goal.ptr),
                                     static_cast<double*>(s->d_seed.ptr), seed_stride, static_cast<double*>(s->d_q.ptr),
                                     cost ? static_cast<double*>(s->d_cost.ptr) : nullptr,
                                     is_solution ? static_cast<int32_t*>(s->d_issol.ptr) : nullptr,
                                     tip_pose ? static_cast<double*>(s->d_tip.ptr) : nullptr));
        if (cost) PIK_CUDA(s, cudaMemcpyAsync(cost, s->d_cost.ptr, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
        if (is_solution) PIK_CUDA(s, cudaMemcpyAsync(is_solution, s->d_issol.ptr, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        if (tip_pose) PIK_CUDA(s, cudaMemcpyAsync(tip_pose, s->d_tip.ptr, (size_t)B * 7 * T * 8, cudaMemcpyDeviceToHost, st));
        return PIK_OK;
    };
    rc = issue();
    // the call is synchronous: no reader is left behind when it returns
    const cudaError_t e = cudaStreamSynchronize(st);
    constants_commit(s, nullptr);
    if (rc != PIK_OK) return rc;
    if (e != cudaSuccess) return fail_cuda(s, e, "cudaStreamSynchronize");
    return PIK_OK;
}

int pik_solver_synchronize(pik_solver* s) {
    if (!s) return PIK_E_INVALID_ARGUMENT;
    PIK_CUDA(s, cudaSetDevice(s->device));
    PIK_CUDA(s, cudaStreamSynchronize(s->stream));
    return PIK_OK;
}

int pik_solver_get_stats(pik_solver* s, pik_stats* out) {
    if (!s || !out) return PIK_E_INVALID_ARGUMENT;
    *out = s->stats;
    return PIK_OK;
}

const char* pik_solver_last_error(const pik_solver* s) { return s ? s->last_error.c_str() : ""; }

int pik_host_alloc(void** out, size_t bytes) {
    if (!out) return PIK_E_INVALID_ARGUMENT;
    *out = nullptr;
    const cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    return e == cudaSuccess ? PIK_OK : (e == cudaErrorMemoryAllocation ? PIK_E_OUT_OF_MEMORY : PIK_E_CUDA);
}

void pik_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int pik_measure_fp64_peak(pik_solver* s, double* tflops) {
    if (!s || !tflops) return PIK_E_INVALID_ARGUMENT;
    if (s->in_flight) return PIK_E_BUSY;
    PIK_CUDA(s, cudaSetDevice(s->device));
    cudaDeviceProp prop;
    PIK_CUDA(s, cudaGetDeviceProperties(&prop, s->device));
    int rc = ensure(s, s->d_stats, 8 * sizeof(unsigned long long));
    if (rc) return rc;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        PIK_CUDA(s, cudaEventRecord(s->ev0, s->stream));
        PIK_CUDA(s, launch_fp64_peak(s->stream, static_cast<double*>(s->d_stats.ptr), blocks, threads, iters));
        PIK_CUDA(s, cudaEventRecord(s->ev1, s->stream));
        PIK_CUDA(s, cudaStreamSynchronize(s->stream));
        float ms = 0.f;
        PIK_CUDA(s, cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        const double flops = 2.0 * 8.0 * (double)iters * threads * (double)blocks;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    *tflops = best;
    return PIK_OK;
}

}  // extern "C"

// ---- hooks for pik_comm.cu (pik_internal.h)
int pik_internal_solver_device(const pik_solver* s) { return s ? s->device : -1; }
int pik_internal_solver_num_variables(const pik_solver* s) { return s ? s->robot.dev.n : 0; }
void* pik_internal_solver_stream(const pik_solver* s) { return s ? static_cast<void*>(s->stream) : nullptr; }

int pik_internal_solve_keep(pik_solver* s, const pik_params* params, int64_t B, int64_t first_problem_index,
                            const double* goal_pose, const double* seed, int64_t seed_stride, int32_t memory) {
    return enqueue_solve(s, params, B, first_problem_index, goal_pose, seed, seed_stride, nullptr, nullptr, nullptr, nullptr,
                         memory, true);
}

int pik_internal_finish(pik_solver* s) { return s ? finish_solve(s) : PIK_E_INVALID_ARGUMENT; }

int pik_internal_pack(pik_solver* s, int64_t B, size_t packed_elems, size_t gather_elems, double** packed, double** gather) {
    if (!s || !packed || !gather) return PIK_E_INVALID_ARGUMENT;
    int rc;
    if ((rc = ensure(s, s->d_packed, packed_elems * sizeof(double)))) return rc;
    if (gather_elems && (rc = ensure(s, s->d_gather, gather_elems * sizeof(double)))) return rc;
    *packed = static_cast<double*>(s->d_packed.ptr);
    *gather = gather_elems ? static_cast<double*>(s->d_gather.ptr) : nullptr;
    if (B < 0) {
        // the shard failed: rows of NaN (all bits set) for the ranks that wait for them
        PIK_CUDA(s, cudaMemsetAsync(*packed, 0xff, packed_elems * sizeof(double), s->stream));
        return PIK_OK;
    }
    if (B == 0) return PIK_OK;
    PIK_CUDA(s, launch_pack_results(s->stream, B, s->robot.dev.n, static_cast<double*>(s->d_solution.ptr),
                                    static_cast<double*>(s->d_cost.ptr), static_cast<int32_t*>(s->d_error.ptr),
                                    static_cast<int32_t*>(s->d_iters.ptr), *packed));
    s->stats.kernel_launches += 1;
    return PIK_OK;
}
