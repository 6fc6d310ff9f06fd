// NCCL communicator owned by a handle: the exchange steps of the path enqueued from the
// library itself, so a data-parallel train step needs no host callback and can be captured into a CUDA graph.
//   - BatchNorm statistics of every layer (forward: sum / sum of squares; backward: sum g / sum g*xhat), fp64
//   - the flat gradient buffer, once per step (SUM: the loss is a sum, model_builder.py:42)
//   - NAP fit statistics (column sums, Gram matrix, rotated sums)
// libnccl is resolved at run time from the process image (PyTorch ships and loads libnccl.so.2; there is no NCCL
// header or link dependency at build time): ncclGetUniqueId on one rank, the 128-byte id travels through the host
// layer's own rendezvous (torch.distributed broadcast), every rank calls ncclCommInitRank.
#include <dlfcn.h>

#include "mmad_internal.cuh"

using namespace mmad;

namespace {

struct NcclId { char internal[128]; };
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclId, int);
typedef int (*CommDestroyFn)(NcclComm);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);

constexpr int kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0;

struct NcclApi {
    void* lib = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    AllReduceFn all_reduce = nullptr;
    GetErrorStringFn error_string = nullptr;
    bool tried = false;
} g_nccl;

int load_nccl() {
    if (g_nccl.all_reduce) return MMAD_OK;
    if (!g_nccl.tried) {
        g_nccl.tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.lib) break;
        }
        if (g_nccl.lib) {
            g_nccl.get_unique_id = (GetUniqueIdFn)dlsym(g_nccl.lib, "ncclGetUniqueId");
            g_nccl.comm_init_rank = (CommInitRankFn)dlsym(g_nccl.lib, "ncclCommInitRank");
            g_nccl.comm_destroy = (CommDestroyFn)dlsym(g_nccl.lib, "ncclCommDestroy");
            g_nccl.all_reduce = (AllReduceFn)dlsym(g_nccl.lib, "ncclAllReduce");
            g_nccl.error_string = (GetErrorStringFn)dlsym(g_nccl.lib, "ncclGetErrorString");
        }
    }
    if (!g_nccl.lib || !g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.comm_destroy || !g_nccl.all_reduce) {
        g_nccl.all_reduce = nullptr;
        set_error("libnccl.so.2 not found in the process (import torch first, or put NCCL on the library path)");
        return MMAD_E_UNSUPPORTED;
    }
    return MMAD_OK;
}

int nccl_ok(int r, const char* what) {
    if (r == 0) return MMAD_OK;
    set_error("%s failed: %s", what, g_nccl.error_string ? g_nccl.error_string(r) : "NCCL error");
    return MMAD_E_CUDA;
}

}  // namespace

namespace mmad {

void handle_set_grad_allreduce(mmad_t h, bool on);

int comm_allreduce(mmad_t h, void* d_buf, long long count, bool f64, cudaStream_t s) {
    void* comm = nullptr; int world = 1;
    handle_comm(h, &comm, &world);
    if (world <= 1 || count <= 0) return MMAD_OK;
    // small fp64 vectors (BatchNorm statistics, NAP column sums): one-kernel exchange over NVLink peer memory when it is open
    if (f64 && count <= peer_max_doubles() && peer_ready(h)) return peer_allreduce_f64(h, static_cast<double*>(d_buf), count, s);
    // the flat gradient buffer, when it is the library-owned peer-mapped one: in-place sum over NVLink peer memory
    if (!f64 && peer_grads_match(h, d_buf, count)) return peer_allreduce_grads(h, s);
    if (!comm) { set_error("no communicator (mmad_comm_init)"); return MMAD_E_STATE; }
    return nccl_ok(g_nccl.all_reduce(d_buf, d_buf, (size_t)count, f64 ? kNcclFloat64 : kNcclFloat32, kNcclSum, comm, s), "ncclAllReduce");
}

}  // namespace mmad

extern "C" {

int mmad_comm_unique_id(unsigned char* h_id) {
    if (!h_id) { set_error("null argument"); return MMAD_E_ARG; }
    int rc = load_nccl();
    if (rc) return rc;
    NcclId id;
    rc = nccl_ok(g_nccl.get_unique_id(&id), "ncclGetUniqueId");
    if (rc) return rc;
    memcpy(h_id, id.internal, sizeof id.internal);
    return MMAD_OK;
}

int mmad_comm_init(mmad_t h, const unsigned char* h_id, int rank, int world) {
    if (!h || !h_id || world < 1 || rank < 0 || rank >= world) { set_error("bad argument"); return MMAD_E_ARG; }
    int rc = load_nccl();
    if (rc) return rc;
    mmad_comm_destroy(h);
    if (world == 1) { handle_set_comm(h, nullptr, 1, 0); return MMAD_OK; }
    NcclId id;
    memcpy(id.internal, h_id, sizeof id.internal);
    NcclComm comm = nullptr;
    rc = nccl_ok(g_nccl.comm_init_rank(&comm, world, id, rank), "ncclCommInitRank");
    if (rc) return rc;
    handle_set_comm(h, comm, world, rank);
    handle_graph_clear(h);
    return MMAD_OK;
}

int mmad_comm_destroy(mmad_t h) {
    if (!h) return MMAD_OK;
    void* comm = nullptr; int world = 1;
    handle_comm(h, &comm, &world);
    if (comm && g_nccl.comm_destroy) {
        handle_graph_clear(h);          // captured graphs hold collectives of this communicator
        cudaDeviceSynchronize();
        g_nccl.comm_destroy(comm);
    }
    handle_set_comm(h, nullptr, 1, 0);
    return MMAD_OK;
}

int mmad_comm_set_grad_allreduce(mmad_t h, int on) {
    if (!h) { set_error("null handle"); return MMAD_E_ARG; }
    handle_set_grad_allreduce(h, on != 0);
    handle_graph_clear(h);
    return MMAD_OK;
}

int mmad_comm_world(mmad_t h) {
    if (!h) return 1;
    void* comm = nullptr; int world = 1;
    handle_comm(h, &comm, &world);
    return comm ? world : 1;
}

int mmad_comm_allreduce_f32(mmad_t h, float* d_buf, long long count, void* stream) {
    if (!h || (!d_buf && count > 0)) { set_error("bad argument"); return MMAD_E_ARG; }
    return comm_allreduce(h, d_buf, count, false, (cudaStream_t)stream);
}

int mmad_comm_allreduce_f64(mmad_t h, double* d_buf, long long count, void* stream) {
    if (!h || (!d_buf && count > 0)) { set_error("bad argument"); return MMAD_E_ARG; }
    return comm_allreduce(h, d_buf, count, true, (cudaStream_t)stream);
}

}  // extern "C"
