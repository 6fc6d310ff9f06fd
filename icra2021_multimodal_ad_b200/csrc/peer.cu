// Latency-bound all-reduce of the BatchNorm statistic vectors over NVLink peer memory.
//
// A data-parallel train step needs the column sums of every BatchNorm layer combined across the ranks BEFORE the layer can be
// normalised -- 16 forward and 16 backward exchanges of <= 2 x 1408 doubles, strictly serialised with the layer chain.
// Through ncclAllReduce each of them costs ~15-20 us at 8 ranks (protocol set-up dominates 22 KB), 0.5 ms per step; the step
// itself is 0.8 ms.  Here every rank owns a buffer that all peers map (cudaIpc), and ONE small kernel per exchange
//   1. pushes its vector into its slot of EVERY rank's buffer (one CTA per destination) as 8-byte (32 data bits, sequence
//      number) pairs -- an 8-byte
//      store is single-copy atomic over NVLink, so the flag travels with the data and no fence or separate signal is needed
//      (the low-latency protocol of collective libraries);
//   2. polls its own buffer until every rank's pairs carry this exchange's sequence number;
//   3. adds the world vectors in RANK ORDER in fp64 and writes the result in place.
// Every rank adds the same numbers in the same order, so the replicas stay bit-identical.  The sequence number lives in
// device memory and is advanced by the kernel, so the exchange replays inside a captured CUDA graph; two buffer sets alternate
// by its parity (a rank can be at most one exchange ahead of the slowest peer: finishing exchange k needs every peer's push
// of exchange k).  All polls are bounded: a lost peer traps instead of hanging the device.
#include "mmad_internal.cuh"

namespace mmad {

namespace {

constexpr int kPeerThreads = 1024;

struct PeerState {
    PeerPtrs p;
    uint2* local = nullptr;
    unsigned long long* d_seq = nullptr;     // [0]: exchange sequence number, [1]: CTA completion counter (as unsigned int)
    double* d_out = nullptr;                 // sums before they are copied back in place
    bool open = false;
    void* mapped[kPeerMaxWorld] = {nullptr};
    // gradient all-reduce: this rank's flat gradient buffer (library-owned so that it can be mapped by the peers)
    float* gbuf = nullptr;
    long long gcount = 0;
    bool gopen = false;
    void* gmapped[kPeerMaxWorld] = {nullptr};
    float* gptr[kPeerMaxWorld] = {nullptr};
    unsigned long long* g_seq = nullptr;     // [0]: sequence number of the gradient all-reduce, [1]: CTA completion counter
};

struct GradPtrs {
    float* buf[kPeerMaxWorld];               // every rank's gradient buffer as mapped here, ROTATED: entry k is rank (rank + 1 + k) % world
                                             // (own buffer last) -- at any moment the 8 ranks talk to 8 different peers instead of all
                                             // starting with rank 0's memory.  Chunk r is summed by rank r alone and broadcast, so the
                                             // order of ITS additions need not match any other rank's: the replicas stay identical.
    uint2* flags[kPeerMaxWorld];             // every rank's flag area: [2 phases][world] (unused, tag) pairs
    int world, rank;
};
constexpr size_t kPeerExchangeBytes = (size_t)2 * kPeerMaxWorld * (2 * kPeerMaxDoubles) * sizeof(uint2);     // 2 MB, flags follow
constexpr int kGradThreads = 512;

// Sum of the flat fp32 gradient buffers of all ranks, in place on every rank, over NVLink peer memory:
//   entry barrier  every rank flags "my backward pass is done" into every peer's flag area and waits for all flags;
//   slice          rank r owns chunk r of the vector: it LOADS that chunk from every rank's buffer (16-byte P2P loads, L1
//                  bypassed), adds them in a fixed order (peers r+1, r+2, ..., itself) and STORES the sum into chunk r of every rank's buffer -- nobody else reads
//                  or writes chunk r, so the two phases of a textbook reduce-scatter + all-gather need no barrier between them;
//   exit barrier   system-scope fence, the last CTA flags "my chunk is everywhere" to all peers and waits for theirs.
// Each rank moves (world-1)/world of the vector in each direction once: 35.8 MB at 8 GPUs for the 40.9 MB of the benchmark
// model.  Measured in the data-parallel step at B = 256 per GPU: +14 us at 2 GPUs (ncclAllReduce +32 us), +139 us at 8
// (ncclAllReduce +251 us) before the peer order was rotated, +44 us after (0.81 TB/s per direction: the NVLink 5 rate).
// Every rank receives the same sums: the replicas stay
// bit-identical.  All polls are bounded (trap, not hang).
__global__ void __launch_bounds__(kGradThreads)
peer_allreduce_f32_kernel(const GradPtrs G, long long n4, unsigned long long* __restrict__ seq_ctr, unsigned int* __restrict__ done_ctr) {
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(seq_ctr);
    const uint32_t tag = (uint32_t)seq + 1u;
    const int world = G.world, rank = G.rank, tid = threadIdx.x;
    auto poll = [&](const uint2* flag, int what) {
        const long long t0 = clock64();
        while (peer_ld(flag).y != tag) {
            if (clock64() - t0 > 4000000000LL) {
                printf("mmad gradient all-reduce: rank %d: %s flag of rank %d never arrived (exchange %llu)\n", rank, what ? "exit" : "entry", tid, seq);
                __trap();
            }
        }
    };
    if (blockIdx.x == 0 && tid < world) peer_st(G.flags[tid] + rank, 0u, tag);
    if (tid < world) poll(G.flags[rank] + tid, 0);
    __syncthreads();
    const long long chunk = (n4 + world - 1) / world;
    const long long lo = (long long)rank * chunk, hi = lo + chunk < n4 ? lo + chunk : n4;
    for (long long i = lo + (long long)blockIdx.x * kGradThreads + tid; i < hi; i += (long long)gridDim.x * kGradThreads) {
        float4 v[kPeerMaxWorld];
#pragma unroll
        for (int r = 0; r < kPeerMaxWorld; ++r)
            if (r < world) v[r] = __ldcg(reinterpret_cast<const float4*>(G.buf[r]) + i);
        float4 a = v[0];
#pragma unroll
        for (int r = 1; r < kPeerMaxWorld; ++r)
            if (r < world) { a.x += v[r].x; a.y += v[r].y; a.z += v[r].z; a.w += v[r].w; }
#pragma unroll
        for (int r = 0; r < kPeerMaxWorld; ++r)
            if (r < world) __stcg(reinterpret_cast<float4*>(G.buf[r]) + i, a);
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    if (tid == 0) s_last = atomicAdd(done_ctr, 1u) + 1u == gridDim.x;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (tid < world) {
        peer_st(G.flags[tid] + world + rank, 0u, tag);
        poll(G.flags[rank] + world + tid, 1);
    }
    __syncthreads();
    if (tid == 0) { *done_ctr = 0; __threadfence(); *reinterpret_cast<volatile unsigned long long*>(seq_ctr) = seq + 1; }
}


// grid = world CTAs: CTA r pushes this rank's vector to rank r; the CTAs then share the polling / summation of the elements.
// One CTA for everything was push-bound at 8 ranks (45 000 remote 8-byte stores through one SM).
__global__ void __launch_bounds__(kPeerThreads)
peer_allreduce_f64_kernel(const PeerPtrs P, double* __restrict__ buf, double* __restrict__ out, int count,
                          unsigned long long* __restrict__ seq_ctr, unsigned int* __restrict__ done_ctr) {
    const unsigned long long seq = *seq_ctr;          // every CTA reads it before the last one to finish advances it
    const uint32_t tag = (uint32_t)seq + 1u;          // never 0 (the buffers start zeroed)
    const size_t set = (size_t)(seq & 1) * P.world * (2 * kPeerMaxDoubles);
    const int world = P.world, rank = P.rank;
    // 1. push to peer blockIdx.x (one destination per CTA)
    for (int r = blockIdx.x; r < world; r += gridDim.x) {
        uint2* dst = P.buf[r] + set + (size_t)rank * (2 * kPeerMaxDoubles);
        for (int i = threadIdx.x; i < count; i += kPeerThreads) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(buf[i]);
            peer_st(dst + 2 * (size_t)i, (uint32_t)bits, tag);
            peer_st(dst + 2 * (size_t)i + 1, (uint32_t)(bits >> 32), tag);
        }
    }
    // 2. + 3. poll the local buffer, add in rank order: elements are dealt round-robin to the CTAs.  The sums go to `out`
    // (not in place: other CTAs may still be pushing `buf`) and are copied back by the last CTA to finish.
    const uint2* mine = P.buf[rank] + set;
    const long long t0 = clock64();
    for (int i = blockIdx.x * kPeerThreads + threadIdx.x; i < count; i += gridDim.x * kPeerThreads) {
        double sum = 0.0;
        for (int r = 0; r < world; ++r) {
            const uint2* q = mine + (size_t)r * (2 * kPeerMaxDoubles) + 2 * (size_t)i;
            uint2 a = peer_ld(q), b = peer_ld(q + 1);
            while (a.y != tag || b.y != tag) {
                if (clock64() - t0 > 4000000000LL) {
                    printf("mmad peer all-reduce: rank %d never received element %d of rank %d (exchange %llu)\n", rank, i, r, seq);
                    __trap();
                }
                if (a.y != tag) a = peer_ld(q);
                if (b.y != tag) b = peer_ld(q + 1);
            }
            sum += __longlong_as_double((long long)(((unsigned long long)b.x << 32) | a.x));
        }
        out[i] = sum;
    }
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(done_ctr, 1u) + 1u == gridDim.x;
        if (s_last) __threadfence();
    }
    __syncthreads();
    if (!s_last) return;
    for (int i = threadIdx.x; i < count; i += kPeerThreads) buf[i] = __ldcg(out + i);
    __syncthreads();
    if (threadIdx.x == 0) { *done_ctr = 0; *seq_ctr = seq + 1; }
}

}  // namespace

void peer_state_free(void* p) {
    PeerState* S = static_cast<PeerState*>(p);
    if (!S) return;
    for (int r = 0; r < kPeerMaxWorld; ++r) if (S->mapped[r]) cudaIpcCloseMemHandle(S->mapped[r]);
    for (int r = 0; r < kPeerMaxWorld; ++r) if (S->gmapped[r]) cudaIpcCloseMemHandle(S->gmapped[r]);
    cudaFree(S->local); cudaFree(S->d_seq); cudaFree(S->d_out); cudaFree(S->gbuf); cudaFree(S->g_seq);
    delete S;
}

bool peer_ready(mmad_t h) {
    PeerState* S = static_cast<PeerState*>(handle_peer_get(h));
    return S && S->open;
}

int peer_max_doubles() { return kPeerMaxDoubles; }

// the library-owned gradient buffer, when d_buf / count are exactly it and every peer has mapped it
bool peer_grads_match(mmad_t h, const void* d_buf, long long count) {
    PeerState* S = static_cast<PeerState*>(handle_peer_get(h));
    return S && S->open && S->gopen && d_buf == S->gbuf && count == S->gcount;
}

int peer_allreduce_grads(mmad_t h, cudaStream_t s) {
    PeerState* S = static_cast<PeerState*>(handle_peer_get(h));
    if (!S || !S->gopen) { set_error("peer gradient buffers are not open"); return MMAD_E_STATE; }
    GradPtrs G;
    memset(&G, 0, sizeof G);
    G.world = S->p.world; G.rank = S->p.rank;
    for (int r = 0; r < G.world; ++r) {
        G.buf[r] = S->gptr[(G.rank + 1 + r) % G.world];
        G.flags[r] = reinterpret_cast<uint2*>(reinterpret_cast<char*>(S->p.buf[r]) + kPeerExchangeBytes);
    }
    const long long n4 = S->gcount / 4;
    const long long per_rank = (n4 + G.world - 1) / G.world;
    int grid = (int)std::min<long long>(128, (per_rank + kGradThreads - 1) / kGradThreads);
    if (grid < 1) grid = 1;
    peer_allreduce_f32_kernel<<<grid, kGradThreads, 0, s>>>(G, n4, S->g_seq, reinterpret_cast<unsigned int*>(S->g_seq + 1));
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

bool peer_kernel_args(mmad_t h, PeerPtrs* ptrs, unsigned long long** seq_ctr, unsigned int** done_ctr) {
    PeerState* S = static_cast<PeerState*>(handle_peer_get(h));
    if (!S || !S->open) return false;
    *ptrs = S->p;
    *seq_ctr = S->d_seq;
    *done_ctr = reinterpret_cast<unsigned int*>(S->d_seq + 1);
    return true;
}

int peer_allreduce_f64(mmad_t h, double* d_buf, long long count, cudaStream_t s) {
    PeerState* S = static_cast<PeerState*>(handle_peer_get(h));
    if (!S || !S->open) { set_error("peer buffers are not open"); return MMAD_E_STATE; }
    if (count <= 0) return MMAD_OK;
    if (count > kPeerMaxDoubles) { set_error("peer all-reduce: %lld doubles > %d", count, kPeerMaxDoubles); return MMAD_E_ARG; }
    peer_allreduce_f64_kernel<<<S->p.world, kPeerThreads, 0, s>>>(S->p, d_buf, S->d_out, (int)count, S->d_seq,
                                                                  reinterpret_cast<unsigned int*>(S->d_seq + 1));
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

}  // namespace mmad

using namespace mmad;

extern "C" {

int mmad_peer_create(mmad_t h, unsigned char* h_handle) {
    if (!h || !h_handle) { set_error("null argument"); return MMAD_E_ARG; }
    static_assert(sizeof(cudaIpcMemHandle_t) == MMAD_IPC_HANDLE_BYTES, "IPC handle size");
    mmad_peer_close(h);
    PeerState* S = new PeerState();
    handle_peer_set(h, S);
    const size_t bytes = kPeerExchangeBytes + 4096;       // 2 MB of exchange slots + the flags of the gradient all-reduce
    MMAD_CUDA_OK(cudaMalloc(&S->local, bytes));
    MMAD_CUDA_OK(cudaMemset(S->local, 0, bytes));
    MMAD_CUDA_OK(cudaMalloc(&S->g_seq, 16));
    MMAD_CUDA_OK(cudaMemset(S->g_seq, 0, 16));
    MMAD_CUDA_OK(cudaMalloc(&S->d_seq, 16));
    MMAD_CUDA_OK(cudaMemset(S->d_seq, 0, 16));
    MMAD_CUDA_OK(cudaMalloc(&S->d_out, (size_t)kPeerMaxDoubles * 8));
    MMAD_CUDA_OK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t hd;
    MMAD_CUDA_OK(cudaIpcGetMemHandle(&hd, S->local));
    memcpy(h_handle, &hd, sizeof hd);
    return MMAD_OK;
}

int mmad_peer_open(mmad_t h, const unsigned char* h_handles, int rank, int world) {
    if (!h || !h_handles || world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world) { set_error("bad argument"); return MMAD_E_ARG; }
    PeerState* S = static_cast<PeerState*>(handle_peer_get(h));
    if (!S || !S->local) { set_error("mmad_peer_create first"); return MMAD_E_STATE; }
    S->p.world = world; S->p.rank = rank;
    for (int r = 0; r < world; ++r) {
        if (r == rank) { S->p.buf[r] = S->local; continue; }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, h_handles + (size_t)r * sizeof hd, sizeof hd);
        void* q = nullptr;
        MMAD_CUDA_OK(cudaIpcOpenMemHandle(&q, hd, cudaIpcMemLazyEnablePeerAccess));
        S->mapped[r] = q;
        S->p.buf[r] = static_cast<uint2*>(q);
    }
    S->open = true;
    handle_graph_clear(h);          // captured steps hold the exchange they were captured with
    return MMAD_OK;
}

int mmad_peer_close(mmad_t h) {
    if (!h) return MMAD_OK;
    void* p = handle_peer_get(h);
    if (p) {
        handle_graph_clear(h);
        cudaDeviceSynchronize();
        peer_state_free(p);
        handle_peer_set(h, nullptr);
    }
    return MMAD_OK;
}

int mmad_peer_grad_alloc(mmad_t h, long long n_floats, float** d_ptr, unsigned char* h_handle) {
    if (!h || !d_ptr || !h_handle || n_floats <= 0) { set_error("bad argument"); return MMAD_E_ARG; }
    PeerState* S = static_cast<PeerState*>(handle_peer_get(h));
    if (!S || !S->local) { set_error("mmad_peer_create first"); return MMAD_E_STATE; }
    if (S->gbuf) { set_error("the gradient buffer of this handle is already allocated"); return MMAD_E_STATE; }
    const long long padded = (n_floats + 3) / 4 * 4;
    MMAD_CUDA_OK(cudaMalloc(&S->gbuf, (size_t)padded * 4));
    MMAD_CUDA_OK(cudaMemset(S->gbuf, 0, (size_t)padded * 4));
    MMAD_CUDA_OK(cudaDeviceSynchronize());
    S->gcount = padded;
    cudaIpcMemHandle_t hd;
    MMAD_CUDA_OK(cudaIpcGetMemHandle(&hd, S->gbuf));
    memcpy(h_handle, &hd, sizeof hd);
    *d_ptr = S->gbuf;
    return MMAD_OK;
}

int mmad_peer_grad_open(mmad_t h, const unsigned char* h_handles) {
    if (!h || !h_handles) { set_error("bad argument"); return MMAD_E_ARG; }
    PeerState* S = static_cast<PeerState*>(handle_peer_get(h));
    if (!S || !S->open || !S->gbuf) { set_error("mmad_peer_open and mmad_peer_grad_alloc first"); return MMAD_E_STATE; }
    for (int r = 0; r < S->p.world; ++r) {
        if (r == S->p.rank) { S->gptr[r] = S->gbuf; continue; }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, h_handles + (size_t)r * sizeof hd, sizeof hd);
        void* q = nullptr;
        MMAD_CUDA_OK(cudaIpcOpenMemHandle(&q, hd, cudaIpcMemLazyEnablePeerAccess));
        S->gmapped[r] = q;
        S->gptr[r] = static_cast<float*>(q);
    }
    S->gopen = true;
    return MMAD_OK;
}

int mmad_peer_allreduce_f64(mmad_t h, double* d_buf, long long count, void* stream) {
    if (!h || (!d_buf && count > 0)) { set_error("bad argument"); return MMAD_E_ARG; }
    return peer_allreduce_f64(h, d_buf, count, (cudaStream_t)stream);
}

}  // extern "C"
