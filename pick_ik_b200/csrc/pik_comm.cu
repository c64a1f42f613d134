// pik_comm.cu -- multi-GPU entry points of include/pik.h: the pose batch shards trivially across ranks
// (problems are independent, RNG streams are keyed by the global problem index), so the only exchange is
// one NCCL all-gather of the packed per-problem results, issued on the solver's stream right behind the
// last kernel of the shard.  NCCL is loaded with dlopen at the first call so that single-GPU users need no
// NCCL at all; a process that already carries an NCCL (e.g. PyTorch's bundled one) gets that instance.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "../../include/pik.h"
#include "pik_internal.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

std::mutex g_mutex;
std::string g_error;
NcclApi g_nccl;

void set_error(const std::string& e) {
    std::lock_guard<std::mutex> lock(g_mutex);
    g_error = e;
}

template <class F>
bool load_sym(void* h, const char* name, F& out) {
    out = reinterpret_cast<F>(dlsym(h, name));
    return out != nullptr;
}

const NcclApi* nccl() {
    static std::once_flag once;
    std::call_once(once, [] {
        // an NCCL already in the process first (same SONAME), then the system library
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            const char* e = dlerror();
            set_error(std::string("cannot load libnccl.so.2: ") + (e ? e : "?"));
            return;
        }
        g_nccl.handle = h;
        g_nccl.ok = load_sym(h, "ncclGetUniqueId", g_nccl.GetUniqueId) && load_sym(h, "ncclCommInitRank", g_nccl.CommInitRank) &&
                    load_sym(h, "ncclCommDestroy", g_nccl.CommDestroy) && load_sym(h, "ncclAllGather", g_nccl.AllGather) &&
                    load_sym(h, "ncclGetErrorString", g_nccl.GetErrorString) && load_sym(h, "ncclSend", g_nccl.Send) &&
                    load_sym(h, "ncclRecv", g_nccl.Recv) && load_sym(h, "ncclGroupStart", g_nccl.GroupStart) &&
                    load_sym(h, "ncclGroupEnd", g_nccl.GroupEnd);
        if (!g_nccl.ok) set_error("libnccl.so.2 lacks a required symbol");
    });
    return g_nccl.ok ? &g_nccl : nullptr;
}

int fail_nccl(const NcclApi* api, ncclResult_t r, const char* what) {
    set_error(std::string(what) + ": " + (api && api->GetErrorString ? api->GetErrorString(r) : "NCCL error"));
    return PIK_E_NCCL;
}

}  // namespace

struct pik_comm {
    ncclComm_t comm = nullptr;
    int n_ranks = 1;
    int rank = 0;
    int device = 0;
};

extern "C" {

static_assert(sizeof(ncclUniqueId) <= PIK_COMM_ID_BYTES, "ncclUniqueId does not fit PIK_COMM_ID_BYTES");

const char* pik_comm_last_error(void) {
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lock(g_mutex);
    copy = g_error;
    return copy.c_str();
}

int pik_comm_unique_id(void* id_out) {
    if (!id_out) return PIK_E_INVALID_ARGUMENT;
    const NcclApi* api = nccl();
    if (!api) return PIK_E_NCCL;
    ncclUniqueId id;
    const ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) return fail_nccl(api, r, "ncclGetUniqueId");
    std::memset(id_out, 0, PIK_COMM_ID_BYTES);
    std::memcpy(id_out, &id, sizeof(id));
    return PIK_OK;
}

int pik_comm_create(const void* id, int32_t n_ranks, int32_t rank, int32_t device, pik_comm** out) {
    if (!out) return PIK_E_INVALID_ARGUMENT;
    *out = nullptr;
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return PIK_E_INVALID_ARGUMENT;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return PIK_E_NO_DEVICE;
    if (device < 0 || device >= count) return PIK_E_INVALID_ARGUMENT;
    const NcclApi* api = nccl();
    if (!api) return PIK_E_NCCL;
    if (cudaSetDevice(device) != cudaSuccess) return PIK_E_CUDA;
    pik_comm* c = new (std::nothrow) pik_comm;
    if (!c) return PIK_E_OUT_OF_MEMORY;
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    const ncclResult_t r = api->CommInitRank(&c->comm, n_ranks, uid, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail_nccl(api, r, "ncclCommInitRank");
    }
    c->n_ranks = n_ranks;
    c->rank = rank;
    c->device = device;
    *out = c;
    return PIK_OK;
}

void pik_comm_destroy(pik_comm* comm) {
    if (!comm) return;
    const NcclApi* api = nccl();
    if (api && comm->comm) {
        cudaSetDevice(comm->device);
        api->CommDestroy(comm->comm);
    }
    delete comm;
}

int pik_solve_batch_gather(pik_solver* solver, pik_comm* comm, const pik_params* params, int64_t B_local,
                           int64_t first_problem_index, const double* goal_pose, const double* seed,
                           int64_t seed_stride, const int64_t* counts, int32_t root, double* gathered, int32_t memory) {
    if (!solver || !comm) return PIK_E_INVALID_ARGUMENT;
    if (pik_internal_solver_device(solver) != comm->device) return PIK_E_INVALID_ARGUMENT;
    if (root < -1 || root >= comm->n_ranks) return PIK_E_INVALID_ARGUMENT;
    if (memory != PIK_MEM_HOST && memory != PIK_MEM_DEVICE) return PIK_E_INVALID_ARGUMENT;
    const bool receives = root < 0 || root == comm->rank;
    if (receives && !gathered) return PIK_E_INVALID_ARGUMENT;
    if (B_local < 0 || (counts && counts[comm->rank] != B_local)) return PIK_E_INVALID_ARGUMENT;
    const NcclApi* api = nccl();
    if (!api) return PIK_E_NCCL;
    const int n = pik_internal_solver_num_variables(solver);
    const size_t w = (size_t)(n + 3);
    bool even = true;
    int64_t total = 0, max_count = 0;
    for (int r = 0; r < comm->n_ranks; ++r) {
        const int64_t c = counts ? counts[r] : B_local;
        if (c < 0) return PIK_E_INVALID_ARGUMENT;
        even = even && c == B_local;
        total += c;
        if (c > max_count) max_count = c;
    }
    if (total == 0) return PIK_OK;
    // From here on this rank takes part in the exchange whatever happens to its own shard: a rank that returned
    // early would leave the others waiting in the collective.  A failed shard travels as rows of NaN.
    const int solve_rc =
        B_local > 0 ? pik_internal_solve_keep(solver, params, B_local, first_problem_index, goal_pose, seed, seed_stride, memory)
                    : PIK_OK;
    double* packed = nullptr;
    double* dst = nullptr;
    // packed shard and (for host callers) the gathered block live in the solver's staging buffers
    const size_t gather_elems = (receives && memory == PIK_MEM_HOST) ? (size_t)total * w : 0;
    int rc = pik_internal_pack(solver, solve_rc == PIK_OK ? B_local : -B_local, (size_t)(B_local > 0 ? B_local : 1) * w,
                               gather_elems, &packed, &dst);
    if (rc != PIK_OK) {
        pik_internal_finish(solver);
        return solve_rc != PIK_OK ? solve_rc : rc;  // no buffer to send from: the other ranks are on their own
    }
    if (memory == PIK_MEM_DEVICE) dst = gathered;
    cudaStream_t st = static_cast<cudaStream_t>(pik_internal_solver_stream(solver));
    ncclResult_t r = ncclSuccess;
    if (root < 0 && even) {
        r = api->AllGather(packed, dst, (size_t)B_local * w, ncclDouble, comm->comm, st);
    } else {
        // uneven shards, or only the root wants the block: grouped point-to-point (NCCL has no gatherv)
        r = api->GroupStart();
        if (receives && r == ncclSuccess) {
            size_t off = 0;
            for (int src = 0; src < comm->n_ranks && r == ncclSuccess; ++src) {
                const size_t c = (size_t)(counts ? counts[src] : B_local);
                if (c > 0) r = api->Recv(dst + off * w, c * w, ncclDouble, src, comm->comm, st);
                off += c;
            }
        }
        if (B_local > 0 && r == ncclSuccess) {
            if (root >= 0) {
                r = api->Send(packed, (size_t)B_local * w, ncclDouble, root, comm->comm, st);
            } else {
                for (int to = 0; to < comm->n_ranks && r == ncclSuccess; ++to)
                    r = api->Send(packed, (size_t)B_local * w, ncclDouble, to, comm->comm, st);
            }
        }
        const ncclResult_t re = api->GroupEnd();
        if (r == ncclSuccess) r = re;
    }
    if (r != ncclSuccess) {
        pik_internal_finish(solver);
        return fail_nccl(api, r, "NCCL exchange");
    }
    if (receives && memory == PIK_MEM_HOST &&
        cudaMemcpyAsync(gathered, dst, (size_t)total * w * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) {
        pik_internal_finish(solver);
        return PIK_E_CUDA;
    }
    rc = pik_internal_finish(solver);
    if (rc == PIK_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = PIK_E_CUDA;
    return solve_rc != PIK_OK ? solve_rc : rc;
}

int pik_solve_batch_sharded(pik_solver* solver, pik_comm* comm, const pik_params* params, int64_t B_local,
                            int64_t first_problem_index, const double* goal_pose, const double* seed,
                            int64_t seed_stride, double* gathered, int32_t memory) {
    return pik_solve_batch_gather(solver, comm, params, B_local, first_problem_index, goal_pose, seed, seed_stride, nullptr,
                                  -1, gathered, memory);
}

}  // extern "C"
