// Internal declarations shared by the libmmad.so translation units.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>      // header-only; ranges are no-ops unless a profiler injects itself

#include "../../include/mmad.h"

namespace mmad {

void set_error(const char* fmt, ...);
extern unsigned long long g_launches;   // kernels launched by this library (bench.py's gpu_launches)

#define MMAD_LAUNCHED() (++mmad::g_launches)
#define MMAD_CUDA_OK(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            mmad::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return MMAD_E_CUDA;                                                         \
        }                                                                               \
    } while (0)

// NVTX range around an entry point of the C ABI (nsys / ncu --nvtx show the path's phases by name)
struct NvtxScope {
    explicit NvtxScope(const char* name) { nvtxRangePushA(name); }
    ~NvtxScope() { nvtxRangePop(); }
    NvtxScope(const NvtxScope&) = delete;
    NvtxScope& operator=(const NvtxScope&) = delete;
};

// Programmatic dependent launch (PDL).  A train step at B = 256 is ~55 dependent kernels of 3-6 us: the launch latency
// between two of them and the next kernel's prologue (barrier init, TMEM allocation, tensor-map fetch) are a third of the
// step.  Kernels launched through launch_k while a PdlScope is active carry the programmatic-stream-serialization
// attribute: their CTAs may become resident while the previous kernel is still running, and they call pdl_wait() before
// touching any global memory (it returns when the previous grid has completed and its writes are visible).  Every kernel
// signals pdl_trigger() at its start.  Only kernels that contain pdl_wait() may be launched through launch_k.
extern thread_local bool g_pdl;
struct PdlScope {
    bool prev;
    explicit PdlScope(bool on) : prev(g_pdl) { g_pdl = on; }
    ~PdlScope() { g_pdl = prev; }
    PdlScope(const PdlScope&) = delete;
    PdlScope& operator=(const PdlScope&) = delete;
};
template <typename... KA, typename... A>
inline cudaError_t launch_k(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    if (g_pdl) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline size_t round_up_sz(size_t x, size_t m) { return (x + m - 1) / m * m; }

constexpr int kPad = 64;   // every activation / weight row is padded to a multiple of 64 elements

// ---------------------------------------------------------------------------------------
// Fused GEMM epilogue description (shared by the CUDA-core and the tcgen05 kernels).
//   acc = sum_k A[m,k] * B[n,k]                 (B rows are output features: Y = X W^T)
//   v   = acc * acc_scale * col_scale[n] + bias[n]
//   if (act) v = leaky_relu(v) * bn_scale[n] + bn_shift[n]
//   Y[m,n] = v (and fp16 hi/lo split copies for the tensor-core path)
//   if (ref)  d = v - ref[m,n];  dout[m,n] = d;  rowpart[tile_n][m] = sum_n d^2
//   if (sq_self) rowpart[tile_n][m] = sum_n v^2             (NAP rotation epilogue)
// ---------------------------------------------------------------------------------------
struct Epilogue {
    const float* bias = nullptr;
    const float* bn_scale = nullptr;   // non-null => LeakyReLU + BatchNorm affine
    const float* bn_shift = nullptr;
    const float* bn_mean = nullptr;    // non-null => bn_scale/bn_shift hold raw gamma/beta, folded in the epilogue
    const float* bn_var = nullptr;
    float bn_eps = 1e-5f;
    float slope = 0.2f;
    float acc_scale = 1.0f;
    // tensor-core kernels: first-order compensation of the truncating fp32 accumulation -- the accumulator of a tile that
    // went through n k-blocks is multiplied by (1 + acc_comp * n) on top of acc_scale; 0 = off
    float acc_comp = 0.0f;
    const float* col_scale = nullptr;  // optional per-column multiplier applied with acc_scale
    float* Y = nullptr;  int ldy = 0;  int y_cols = 0;   // writes cols [0,y_cols); cols >= N are zero-filled
    __half* Yh = nullptr; __half* Yl = nullptr; int ldh = 0;
    const float* ref = nullptr; int ldref = 0;
    float* dout = nullptr; int lddout = 0;
    __half* Dh = nullptr; __half* Dl = nullptr; int lddh = 0;   // fp16 split of d * d_scale (NAP operand)
    float d_scale = 1.0f;
    int d_cols = 0;                                      // dout/Dh/Dl columns written (>= N: zero filled)
    float* rowpart = nullptr; int rowpart_stride = 0;   // [tiles_n][rowpart_stride]
    int sq_self = 0;
    float* pre = nullptr; int ldpre = 0;                // train: pre-activation (bias added, before act)
    int plain = 0;                                      // Y = acc * acc_scale + bias only, rows of Y may be unaligned
    int b_upper_tri = 0;                                // number of leading rows n of B with B[n, k] == 0 for k < n (tensor-core kernels skip those k-blocks)
    int pre_zeroed = 0;                                 // split-K: the caller already zeroed Y (no memset inside gemm_tc)
    int split_k_ok = 0;                                 // plain mode: split-K with atomic accumulation allowed
    long long slab_stride = 0;                          // plain split-K WITHOUT atomics: split sp stores its partial tile at Y + sp * slab_stride
                                                        // (elements); the consumer adds the slabs in order -- deterministic, no zeroing
    float y_split_scale = 1.f;                          // Yh/Yl hold the split of (value * y_split_scale)
    // MMAD_PREC_F16F8: Yl / Dl hold the fp8 twin instead of the fp16 lo part -- per 4 columns 8 bytes:
    // 4 x e4m3((v - hi) * 2^11) then 4 x e4m3(v)  (v = value * split scale).  Same bytes per row as the fp16 lo; a
    // 64-column k-block of the fp16 operand corresponds to one 128-byte block of the twin.
    int lo_f8 = 0;
    int d_lo_f8 = 0;             // the DIFF twins (Dh / Dl) in the fp8 twin layout (lo_f8 describes the activation twins Yh / Yl)
};

constexpr float kF8LoScale = 2048.f;   // the residual v - fp16(v) is at most 2^-11 |v|: scaled to |lo8| <= |v|

struct GemmShape {
    int M, N, K;
    const float* A; int lda;     // [M, K] (or [K, M] when transA)
    const float* B; int ldb;     // [N, K] (or [K, N] when transB)
    bool transA = false, transB = false;
};

// Column map of the concatenated diffs: the workspace copy pads every layer segment to a
// multiple of kPad columns (aligned vector stores, TMA-legal strides); 'tight' is the reference's
// concatenation (utils/metric.py:166-167).
struct SegMap {
    int n = 0;
    int tight_off[MMAD_MAX_LAYERS + 2] = {0};
    int pad_off[MMAD_MAX_LAYERS + 2] = {0};
};
__host__ __device__ inline int seg_pad_col(const SegMap& m, int c) {
    int i = 0;
    while (i + 1 < m.n && c >= m.tight_off[i + 1]) ++i;
    return m.pad_off[i] + (c - m.tight_off[i]);
}

// small batches on the tensor cores: plain split-K GEMM into `acc` (zeroed by the caller) + the epilogue as its own kernel (gemm_tc.cu)
struct TcOperand;
int gemm_tc_small(const TcOperand& A, const TcOperand& B, int M, int N, int K, int passes, const Epilogue& e, float* slabs, int ldacc,
                  long long slab_stride, int max_slabs, cudaStream_t s);

// CUDA-core fp32 GEMM with the fused epilogue (gemm_simt.cu)
int gemm_simt(const GemmShape& g, const Epilogue& e, cudaStream_t s);
int gemm_simt_tile_n();   // columns covered by one rowpart slot

// weight-streaming fp32 GEMM for <= 64 rows (gemm_skinny.cu): one 16-column row-partial slot per CTA
int gemm_skinny(const GemmShape& g, const Epilogue& e, cudaStream_t s);
bool gemm_skinny_ok(const GemmShape& g, const Epilogue& e);
int gemm_skinny_max_rows();
int gemm_skinny_tile_n();

// tcgen05 GEMM (gemm_tc.cu): operands are fp16 hi/lo pairs described by TMA maps.
struct TcOperand {
    CUtensorMap hi, lo;      // 2-D maps, box = [64 (K) x rows], 128B swizzle
    int rows = 0, k = 0;
    bool mn = false;         // MN-major: the matrix is stored [contraction, M or N] (boxes of 64 x 64)
};
int tc_available();
int tc_make_operand_map(CUtensorMap* map, const __half* base, int rows, int kp, int ld, int box_rows);
// fp8 twin ([rows, 2 * round_up(k, 64)] bytes, row stride ld_bytes): box = [128 bytes x box_rows]
int tc_make_operand_map_f8(CUtensorMap* map, const void* base, int rows, int k, int ld_bytes, int box_rows);
int gemm_tc(const TcOperand& A, const TcOperand& B, int M, int N, int K, int passes,
            const Epilogue& e, cudaStream_t s, int bn = 256,    // bn: CTA tile width; K-major B maps need box_rows == bn
            int* splits_out = nullptr, int max_splits = 16);    // split-K factor chosen (plain epilogue), its upper bound
int gemm_tc_tile_n();                    // the default (256)
int gemm_tc_rowpart_cols();              // columns per row-partial slot written by the tensor-core epilogues (128)
int gemm_tc_pick_bn(int M, int N);
// CTA-pair (cta_group::2) kernel for bulk scoring (gemm_tc2.cu): K-major operands, fused epilogue;
// B maps with box_rows == 128 (each CTA of the pair stages half of the 256-wide tile)
int tc2_available();
int gemm_tc2(const TcOperand& A, const TcOperand& B, int M, int N, int K, int passes, const Epilogue& e, cudaStream_t s);       // 256, or 128 when 256 would leave SMs idle

// handle internals shared with train.cu (defined in mmad_api.cu)
struct LayerView { int K, N, Kp, Np; __half* Wh; __half* Wl; float wscale; };
const mmad_desc_t* handle_desc(mmad_t h);
LayerView handle_layer(mmad_t h, int module, int index);

// fp32 view of a packed layer (stream.cu): W [N, Kp] zero padded, bias / eval-BN scale / shift [Np]
struct LayerF32 { int K, N, Kp, Np; bool has_bn; const float* W; const float* bias; const float* scale; const float* shift; };
LayerF32 handle_layer_f32(mmad_t h, int module, int index);
unsigned long long handle_weights_gen(mmad_t h);      // incremented by every mmad_set_layer
void* handle_stream_get(mmad_t h);                    // opaque state of the one-launch realtime kernel (stream.cu)
void handle_stream_set(mmad_t h, void* state);
void stream_state_free(void* state);
bool stream_enabled();                                // MMAD_NO_STREAM_KERNEL=1 disables
int stream_max_rows();
int stream_score(mmad_t h, const float* h_x, int ldx, int rows, int lo, int hi, float* h_base, float* h_sap);
float* stream_input_buffer(mmad_t h, int lo, int hi);
// fused whole-chain kernel for models whose widths are all <= 128 (smallnet.cu)
void* handle_smallnet_get(mmad_t h);
void handle_smallnet_set(mmad_t h, void* state);
void smallnet_state_free(void* state);
bool smallnet_enabled();                              // MMAD_NO_SMALLNET=1 disables
bool smallnet_fits(mmad_t h);
// tensor-core variant (smallnet_tc.cu): TMA maps of a layer's fp16 twins (box 64 halfs x 128 rows) and their scale
int handle_layer_tcmaps(mmad_t h, int module, int index, CUtensorMap* wh, CUtensorMap* wl, float* wscale);
int smallnet_tc_score(mmad_t h, const float* d_x, int ldx, int n, int lo, int hi, float* d_base, float* d_sap, float acc_comp,
                      cudaStream_t s);
int smallnet_prepare(mmad_t h, int lo, int hi, cudaStream_t s);     // (re)builds the packed plan; synchronises s
int smallnet_score(mmad_t h, const float* d_x, int ldx, int n, int lo, int hi, float* d_base, float* d_sap, cudaStream_t s);

// CUDA-graph cache of a handle (launch-bound sequences: the train step, small-batch scoring)
bool graphs_enabled();                       // MMAD_NO_GRAPHS=1 disables
cudaGraphExec_t handle_graph_find(mmad_t h, const std::string& key, unsigned long long* launches);
void handle_graph_put(mmad_t h, const std::string& key, cudaGraphExec_t g, unsigned long long launches);
cudaStream_t handle_capture_stream(mmad_t h);
void handle_graph_clear(mmad_t h);
// second stream + fork/join events of a handle (independent branches of a launch sequence); 0 on success
int handle_aux(mmad_t h, cudaStream_t* s2, cudaEvent_t* ev_fork, cudaEvent_t* ev_join);

// NVLink peer-memory exchange of small fp64 vectors (peer.cu)
void* handle_peer_get(mmad_t h);
void handle_peer_set(mmad_t h, void* state);
void peer_state_free(void* state);
bool peer_ready(mmad_t h);
int peer_max_doubles();
int peer_allreduce_f64(mmad_t h, double* d_buf, long long count, cudaStream_t s);
bool peer_grads_match(mmad_t h, const void* d_buf, long long count);
// train step: early publication of the loss to mapped pinned memory (mmad_api.cu)
int handle_loss_doorbell(mmad_t h, uint2** d_pair, unsigned long long** d_seq, bool allocate);   // 0: available
void handle_loss_published(mmad_t h);
int handle_loss_read(mmad_t h, float* out);
int peer_allreduce_grads(mmad_t h, cudaStream_t s);

// Every rank's exchange buffer as mapped by this rank (cudaIpc), [2 sets][world slots][2 * kPeerMaxDoubles] (data, tag) pairs
constexpr int kPeerMaxWorld = 16;
constexpr int kPeerMaxDoubles = 4096;                 // 2 x the widest BatchNorm layer (padded) fits with room to spare
struct PeerPtrs {
    uint2* buf[kPeerMaxWorld];
    int world, rank;
};
// kernel arguments of an exchange fused into another kernel (peer_exchange_cta): false when the peer buffers are not open
bool peer_kernel_args(mmad_t h, PeerPtrs* ptrs, unsigned long long** seq_ctr, unsigned int** done_ctr);

#ifdef __CUDACC__
__device__ __forceinline__ uint2 peer_ld(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void peer_st(uint2* p, uint32_t data, uint32_t seq) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(seq) : "memory");
}

// All-reduce (sum over ranks, rank order, fp64) of NV <= 64 doubles held in shared memory by ONE CTA, as part of a grid-wide
// exchange in which every CTA of the grid does the same for its own elements: CTA-local element e is element elem_of(e) of the
// exchange vector.  Protocol of peer.cu: each value travels as two 8-byte (32 data bits, tag) pairs into the sender's slot of
// EVERY rank's buffer, the receiver polls its own buffer; the tag is the exchange's sequence number + 1, the buffer set its
// parity.  The CTA that finishes last advances the sequence number (all CTAs have read it by then).  Call with all threads of
// the CTA (>= 2 * NV threads); `vals` holds the local sums on entry and the global sums on return.
template <int NV, typename ElemOf>
__device__ __forceinline__ void peer_exchange_cta(const PeerPtrs& P, unsigned long long* seq_ctr, unsigned int* done_ctr, double* vals,
                                                  ElemOf elem_of) {
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(seq_ctr);
    const uint32_t tag = (uint32_t)seq + 1u;
    const size_t set = (size_t)(seq & 1) * P.world * (2 * kPeerMaxDoubles);
    const int t = threadIdx.x;
    if (t < 2 * NV) {
        const int e = t >> 1, part = t & 1;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[e]);
        const uint32_t w = part ? (uint32_t)(bits >> 32) : (uint32_t)bits;
        const size_t off = set + (size_t)P.rank * (2 * kPeerMaxDoubles) + 2 * (size_t)elem_of(e) + part;
        for (int k = 0; k < P.world; ++k) {       // own rank last: the ranks start with different peers
            int r = P.rank + 1 + k;
            if (r >= P.world) r -= P.world;
            peer_st(P.buf[r] + off, w, tag);
        }
    }
    __syncthreads();            // every thread has read vals[]
    if (t < NV) {
        const uint2* mine = P.buf[P.rank] + set + 2 * (size_t)elem_of(t);
        const long long t0 = clock64();
        double sum = 0.0;
        for (int r = 0; r < P.world; ++r) {
            const uint2* q = mine + (size_t)r * (2 * kPeerMaxDoubles);
            uint2 a = peer_ld(q), b = peer_ld(q + 1);
            while (a.y != tag || b.y != tag) {
                if (clock64() - t0 > 4000000000LL) {
                    printf("mmad fused exchange: rank %d never received element %d of rank %d (exchange %llu)\n", P.rank, elem_of(t), r, seq);
                    __trap();
                }
                if (a.y != tag) a = peer_ld(q);
                if (b.y != tag) b = peer_ld(q + 1);
            }
            sum += __longlong_as_double((long long)(((unsigned long long)b.x << 32) | a.x));
        }
        vals[t] = sum;
    }
    __syncthreads();
    if (t == 0) {
        __threadfence();
        if (atomicAdd(done_ctr, 1u) + 1u == gridDim.x) {
            *done_ctr = 0;
            __threadfence();
            *reinterpret_cast<volatile unsigned long long*>(seq_ctr) = seq + 1;
        }
    }
}
#endif

// NCCL communicator of a handle (comm.cu)
void handle_comm(mmad_t h, void** comm, int* world);
void handle_set_comm(mmad_t h, void* comm, int world, int rank);
int comm_allreduce(mmad_t h, void* d_buf, long long count, bool f64, cudaStream_t s);
bool handle_grad_allreduce(mmad_t h);     // all-reduce every layer's gradients inside the train step

// elementwise helpers (elementwise.cu)
int pad_split(const float* x, int ldx, int n, int D, float* xp, int ldp, __half* xh, __half* xl, int ldh,
              cudaStream_t s, int lo_f8 = 0);
// MMAD_PREC_F16F8 weight twins: Wh = fp16(W * scale), W8 = per 4 columns [4 x e4m3(Wh / 2^11) | 4 x e4m3(W * scale - Wh)];
// *d_scale (device) receives the power-of-two scale that puts max|W| * scale in [2^13, 2^14)
int split_weights_f8(const float* W, int N, int K, int ldw, int Kp, __half* Wh, uint8_t* W8, float* d_scale, cudaStream_t s);
int split_weights(const float* W, int N, int K, int Kp, float scale, __half* Wh, __half* Wl, cudaStream_t s);
int split_weights_multi(int n, const float* const* W, const int* N, const int* K, const int* Kp, float scale, __half* const* Wh,
                        __half* const* Wl, cudaStream_t s);
int finalize_scores(const float* rowpart, int stride, int n, int slot_lo_base, int slot_hi_base,
                    int sap_slot_lo, int sap_slot_hi, float inv_base, float inv_sap,
                    float* base, float* sap, cudaStream_t s);
int finalize_sum(const float* rowpart, int stride, int n, int slot_lo, int slot_hi, float scale, float* out,
                 cudaStream_t s);
int reduce_sum_all(const float* rowpart, int stride, int n, int slot_lo, int slot_hi, float* acc, cudaStream_t s);
int fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int N, int Np,
            float* scale, float* shift, cudaStream_t s);
int copy_pad_vec(const float* src, int N, int Np, float* dst, cudaStream_t s);
int colsum_f64(const float* d, int ld, int n, int cols, const SegMap& sm, double* sum, cudaStream_t s);
int gram_f64_accumulate(const float* g32, int ld32, int D, const SegMap& sm, double* g64, cudaStream_t s);
int center_rows(float* d, int ld, int n, int cols, const SegMap& sm, const float* mu, cudaStream_t s);
int gram_f64_direct(const float* dc, int ld, int rows, int D, const SegMap& sm, double* g64, cudaStream_t s);
int nap_pack(const float* mu, const float* vt, const float* var, const float* mu2, int K, int D, int Dp,
             const SegMap& sm, float* B, float* colscale, float* bias, float* bias_rot, cudaStream_t s);
int nap_restandardize(const float* var, const float* mu2, const float* bias_rot, int K, float* colscale, float* bias,
                      cudaStream_t s);
int col_sum_sq_f64(const float* y, int ld, int n, int cols, double* sum, double* sq, cudaStream_t s);

}  // namespace mmad
