// Weight-streaming fp32 GEMM for the realtime caller (test_file/realtime_tester.py:291-309: batch_size = 10 windows
// per call; used up to 32 rows, taller calls go to the tensor-core kernels).
//
//   Y[M <= 32, N] = epilogue(A[M, K] . B[N, K]^T)        exact fp32 FMA arithmetic
//
// With M this small the layer is bound by streaming the weights once (41 MB of distinct fp32 weights per call,
// L2-resident between calls) and by launch latency, not by math: a 128-row tensor-core tile would leave all but
// N/256 SMs idle and serialise the K loop.  Here every CTA owns 16 output columns and the full K extent, so a
// 1728-wide layer runs on 108 SMs at once; the whole 15-layer chain is replayed from one CUDA graph.
// CTA: 256 threads = 16 columns x 16 row groups of 4; A and B slabs of 64 k staged in shared memory
// (double buffered, 16-byte loads, padded rows: conflict-free), float4 inner product over k.
// Epilogue: the Epilogue subset the scoring chain needs (bias, LeakyReLU + BN affine, fp32 output with zero
// padding, diff against a reference, row partial sums of squares -- one 16-column slot per CTA).
#include "mmad_internal.cuh"

namespace mmad {

namespace {

constexpr int SK_MAX_M = 32;  // rows handled by this kernel family
constexpr int SK_N = 16;      // output columns per CTA
constexpr int SK_T = 256;
// cp.async ring depth by staged row count: the fewer rows, the smaller a stage and the more weight slabs are in
// flight per CTA (15 x 4 KB at <= 4 rows) -- the K loop of a CTA is otherwise one L2 round trip per slab
// ... and the longer a slab (k per stage), the fewer barrier rounds per layer: 256 / 128 / 64 k per slab
template <int ROWS> struct SkCfg {
    static constexpr int kK = ROWS <= 4 ? 256 : (ROWS <= 16 ? 128 : 64);
    static constexpr int kLd = kK + 4;
    static constexpr int kStages = ROWS <= 4 ? 4 : (ROWS <= 16 ? 5 : 6);
    static constexpr int kAFloats = ROWS * kLd, kBFloats = SK_N * kLd;
    static constexpr int kSmem = kStages * (kAFloats + kBFloats) * 4 + 16 * 4 * 16 * 4;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Thread roles: tx = column (16), ty (16) = row group rg (4 rows each) x k split ks.  With M rows only
// RG = ceil(M/4) row groups exist (rounded to a power of two), so the remaining 16/RG thread rows split each
// 64-wide k slab among themselves and their partial sums are combined through shared memory at the end:
// one window costs 1/16 of the inner-product work of 64.
template <int ROWS>
__global__ void __launch_bounds__(SK_T)
gemm_skinny_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, Epilogue e) {
    constexpr int SK_STAGES = SkCfg<ROWS>::kStages;
    constexpr int SK_K = SkCfg<ROWS>::kK, SK_LD = SkCfg<ROWS>::kLd;
    constexpr int SK_A_FLOATS = SkCfg<ROWS>::kAFloats, SK_B_FLOATS = SkCfg<ROWS>::kBFloats;
    extern __shared__ __align__(16) float sk_smem[];
    float* As = sk_smem;                                       // [stage][64][68]
    float* Bs = sk_smem + SK_STAGES * SK_A_FLOATS;             // [stage][16][68]
    float* red = Bs + SK_STAGES * SK_B_FLOATS;                 // [16 ty][4][16 tx]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * SK_N;
    int RG = 1;
    while (RG * 4 < M) RG <<= 1;                               // 1, 2, 4, 8, 16
    const int KS = 16 / RG;
    const int rg = ty % RG, ks = ty / RG;
    const int kspan = SK_K / KS;                               // k per thread per slab (multiple of 4)
    const int Kr = (K + 3) & ~3;                               // rows are padded to a multiple of 4 (zeros in B beyond K)
    const int nk = (Kr + SK_K - 1) / SK_K;
    const int rows_ld = RG * 4 < ROWS ? RG * 4 : ROWS;         // A rows staged (rows >= M are zero filled once)
    // rows of A in [M, rows_ld) and k beyond Kr are never loaded: clear every stage once
    for (int i = tid; i < SK_STAGES * (SK_A_FLOATS + SK_B_FLOATS); i += SK_T) sk_smem[i] = 0.f;
    __syncthreads();
    auto issue = [&](int kt) {
        if (kt < nk) {
            const int st = kt % SK_STAGES;
            float* a_st = As + (size_t)st * SK_A_FLOATS;
            float* b_st = Bs + (size_t)st * SK_B_FLOATS;
            constexpr int V = SK_K / 4;                        // float4 per row of a slab
            for (int i = tid; i < (rows_ld + SK_N) * V; i += SK_T) {
                const int row = i / V, kk = (i % V) * 4;
                const int k = kt * SK_K + kk;
                if (row < rows_ld) {
                    if (k >= Kr) *reinterpret_cast<float4*>(a_st + row * SK_LD + kk) = make_float4(0.f, 0.f, 0.f, 0.f);   // tail of the last slab
                    else if (row < M) cp_async16(a_st + row * SK_LD + kk, A + (size_t)row * lda + k);
                } else {
                    const int br = row - rows_ld, c = n0 + br;
                    if (k >= Kr) *reinterpret_cast<float4*>(b_st + br * SK_LD + kk) = make_float4(0.f, 0.f, 0.f, 0.f);
                    else if (c < N) cp_async16(b_st + br * SK_LD + kk, B + (size_t)c * ldb + k);
                }
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < SK_STAGES - 1; ++i) issue(i);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<SK_STAGES - 2>();
        __syncthreads();                  // slab kt has landed for every thread; slab kt-1's buffer is free again
        issue(kt + SK_STAGES - 1);
        const int st = kt % SK_STAGES;
        const float* a_s = As + (size_t)st * SK_A_FLOATS + (rg * 4) * SK_LD;
        const float* b_s = Bs + (size_t)st * SK_B_FLOATS + tx * SK_LD;
        const int k_lo = ks * kspan;
        for (int k = k_lo; k < k_lo + kspan; k += 4) {
            const float4 b = *reinterpret_cast<const float4*>(b_s + k);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(a_s + i * SK_LD + k);
                acc[i] = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc[i]))));
            }
        }
    }
    if (KS > 1) {                         // combine the k splits
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) red[(ty * 4 + i) * 16 + tx] = acc[i];
        __syncthreads();
        if (ks == 0) {
            for (int s2 = 1; s2 < KS; ++s2)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] += red[((rg + s2 * RG) * 4 + i) * 16 + tx];
        }
    }
    // ---- fused epilogue (threads of k split 0 own the outputs; everybody takes part in the shuffles) ----
    const int c = n0 + tx;
    const bool cok = c < N;
    const float bias = (e.bias && cok) ? e.bias[c] : 0.f;
    const float cs = e.acc_scale * ((e.col_scale && cok) ? e.col_scale[c] : 1.f);
    const float sc = (e.bn_scale && cok) ? e.bn_scale[c] : 1.f;
    const float sh = (e.bn_scale && cok) ? e.bn_shift[c] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = rg * 4 + i;
        const bool rok = r < M && ks == 0;
        float v = fmaf(acc[i], cs, bias);
        if (e.bn_scale) {
            v = v > 0.f ? v : v * e.slope;
            v = fmaf(v, sc, sh);
        }
        if (!cok) v = 0.f;
        float sq = 0.f;
        if (rok) {
            if (e.Y && c < e.y_cols) e.Y[(size_t)r * e.ldy + c] = v;
            if (e.ref) {
                const float d = cok ? v - e.ref[(size_t)r * e.ldref + c] : 0.f;
                if (e.dout && c < e.d_cols) e.dout[(size_t)r * e.lddout + c] = d;
                sq = d * d;
            } else if (e.sq_self) {
                sq = v * v;
            }
        }
        if (e.rowpart) {   // the 16 columns of this CTA live in one half-warp
            sq += __shfl_xor_sync(0xffffffffu, sq, 1);
            sq += __shfl_xor_sync(0xffffffffu, sq, 2);
            sq += __shfl_xor_sync(0xffffffffu, sq, 4);
            sq += __shfl_xor_sync(0xffffffffu, sq, 8);
            if (tx == 0 && rok) e.rowpart[(size_t)blockIdx.x * e.rowpart_stride + r] = sq;
        }
    }
}

bool g_sk_init = false;

}  // namespace

int gemm_skinny_max_rows() { return SK_MAX_M; }
int gemm_skinny_tile_n() { return SK_N; }

bool gemm_skinny_ok(const GemmShape& g, const Epilogue& e) {
    auto al = [](const void* q, int ld) { return ((reinterpret_cast<uintptr_t>(q) & 15) == 0) && ld % 4 == 0; };
    return g.M <= SK_MAX_M && !g.transA && !g.transB && al(g.A, g.lda) && al(g.B, g.ldb) && !e.Yh && !e.Dh && !e.pre && !e.bn_mean;
}

int gemm_skinny(const GemmShape& g, const Epilogue& e, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0) return MMAD_OK;
    if (!gemm_skinny_ok(g, e)) { set_error("gemm_skinny: unsupported shape/epilogue"); return MMAD_E_ARG; }
    if (!g_sk_init) {
        MMAD_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<4>::kSmem));
        MMAD_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<16>::kSmem));
        MMAD_CUDA_OK(cudaFuncSetAttribute(gemm_skinny_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<32>::kSmem));
        g_sk_init = true;
    }
    const int cols = std::max(g.N, std::max(e.Y ? e.y_cols : 0, e.ref ? e.d_cols : 0));
    const int grid = (cols + SK_N - 1) / SK_N;
    if (g.M <= 4) gemm_skinny_kernel<4><<<grid, SK_T, SkCfg<4>::kSmem, s>>>(g.M, g.N, g.K, g.A, g.lda, g.B, g.ldb, e);
    else if (g.M <= 16) gemm_skinny_kernel<16><<<grid, SK_T, SkCfg<16>::kSmem, s>>>(g.M, g.N, g.K, g.A, g.lda, g.B, g.ldb, e);
    else gemm_skinny_kernel<32><<<grid, SK_T, SkCfg<32>::kSmem, s>>>(g.M, g.N, g.K, g.A, g.lda, g.B, g.ldb, e);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

}  // namespace mmad
