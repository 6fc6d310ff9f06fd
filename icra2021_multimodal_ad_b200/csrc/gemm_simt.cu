// CUDA-core fp32 GEMM with the fused layer epilogue (MMAD_PREC_FP32).
//
// This is the bit-faithful-fp32 arithmetic mode: every product is an fp32 FMA, so the
// scores match the reference's CPU sgemm path to rounding-order noise (~1e-6).  It is
// also the kernel for the shapes the tensor-core kernel does not take (transposed
// operands of the backward pass, tiny batches).  The tcgen05 kernel (gemm_tc.cu) is the
// throughput path.
//
// Tiling: CTA tile BM x 128 x 16, 256 threads, each thread a (BM/16) x 8 register
// tile; A/B tiles are staged k-major in shared memory (double-buffered, register
// prefetch of the next k-slab) so the inner loop reads conflict-free float4s.
#include "mmad_internal.cuh"

namespace mmad {

namespace {

constexpr int BN = 128;
constexpr int BK = 16;
constexpr int NTHREADS = 256;
constexpr int SPAD = 4;

template <int BM>
__device__ __forceinline__ int row_of(int ty, int i) {
    constexpr int TM = BM / 16;
    if (TM == 8) return i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4);
    return ty * TM + i;
}
__device__ __forceinline__ int col_of(int tx, int j) { return j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

// Load a [ROWS x BK] slab of an operand stored [rows, K] (k contiguous), or when TRANS
// stored [K, rows] (rows contiguous), into registers as float4s.
template <int ROWS, bool TRANS>
__device__ __forceinline__ void load_slab(const float* __restrict__ P, int ld, int rows_total, int K,
                                          int row0, int k0, int tid, bool vec, float4 (&r)[ROWS * BK / 4 / NTHREADS]) {
    constexpr int NV = ROWS * BK / 4 / NTHREADS;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        int idx = tid + i * NTHREADS;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!TRANS) {
            int row = row0 + idx / (BK / 4);
            int k = k0 + (idx % (BK / 4)) * 4;
            if (row < rows_total) {
                const float* p = P + (size_t)row * ld + k;
                if (vec && k + 3 < K) {
                    v = *reinterpret_cast<const float4*>(p);
                } else {
                    if (k < K) v.x = p[0];
                    if (k + 1 < K) v.y = p[1];
                    if (k + 2 < K) v.z = p[2];
                    if (k + 3 < K) v.w = p[3];
                }
            }
        } else {
            int k = k0 + idx / (ROWS / 4);
            int row = row0 + (idx % (ROWS / 4)) * 4;
            if (k < K) {
                const float* p = P + (size_t)k * ld + row;
                if (vec && row + 3 < rows_total) {
                    v = *reinterpret_cast<const float4*>(p);
                } else {
                    if (row < rows_total) v.x = p[0];
                    if (row + 1 < rows_total) v.y = p[1];
                    if (row + 2 < rows_total) v.z = p[2];
                    if (row + 3 < rows_total) v.w = p[3];
                }
            }
        }
        r[i] = v;
    }
}

template <int ROWS, bool TRANS>
__device__ __forceinline__ void store_slab(float (*S)[ROWS + SPAD], int tid, const float4 (&r)[ROWS * BK / 4 / NTHREADS]) {
    constexpr int NV = ROWS * BK / 4 / NTHREADS;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        int idx = tid + i * NTHREADS;
        if (!TRANS) {
            int row = idx / (BK / 4);
            int k = (idx % (BK / 4)) * 4;
            S[k + 0][row] = r[i].x;
            S[k + 1][row] = r[i].y;
            S[k + 2][row] = r[i].z;
            S[k + 3][row] = r[i].w;
        } else {
            int k = idx / (ROWS / 4);
            int row = (idx % (ROWS / 4)) * 4;
            *reinterpret_cast<float4*>(&S[k][row]) = r[i];
        }
    }
}

template <int BM, bool TA, bool TB>
__global__ void __launch_bounds__(NTHREADS)
gemm_simt_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                 bool vecA, bool vecB, Epilogue e) {
    constexpr int TM = BM / 16;
    __shared__ __align__(16) float As[2][BK][BM + SPAD];
    __shared__ __align__(16) float Bs[2][BK][BN + SPAD];

    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;

    float acc[TM][8];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    static_assert(BM * BK / 4 % NTHREADS == 0, "slab split");
    float4 ra[BM * BK / 4 / NTHREADS];
    float4 rb[BN * BK / 4 / NTHREADS];

    const int nk = (K + BK - 1) / BK;
    auto loadA = [&](int kt) { load_slab<BM, TA>(A, lda, M, K, m0, kt * BK, tid, vecA, ra); };
    auto loadB = [&](int kt) { load_slab<BN, TB>(B, ldb, N, K, n0, kt * BK, tid, vecB, rb); };

    loadA(0);
    loadB(0);
    store_slab<BM, TA>(As[0], tid, ra);
    store_slab<BN, TB>(Bs[0], tid, rb);
    __syncthreads();

    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            loadA(kt + 1);
            loadB(kt + 1);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[8];
            if constexpr (TM == 8) {
                float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
                float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            } else {
                float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            }
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_slab<BM, TA>(As[cur ^ 1], tid, ra);
            store_slab<BN, TB>(Bs[cur ^ 1], tid, rb);
        }
        __syncthreads();
    }

    // ---- fused epilogue ----
    float bias[8], sc[8], sh[8], cs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int c = n0 + col_of(tx, j);
        bool ok = c < N;
        bias[j] = (e.bias && ok) ? e.bias[c] : 0.f;
        cs[j] = e.acc_scale * ((e.col_scale && ok) ? e.col_scale[c] : 1.f);
        sc[j] = (e.bn_scale && ok) ? e.bn_scale[c] : 1.f;
        sh[j] = (e.bn_scale && ok) ? e.bn_shift[c] : 0.f;
        if (e.bn_mean && ok) {   // raw BatchNorm tensors: fold here (stand-alone FCLayer op)
            sc[j] = sc[j] / sqrtf(e.bn_var[c] + e.bn_eps);
            sh[j] = sh[j] - e.bn_mean[c] * sc[j];
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = m0 + row_of<BM>(ty, i);
        const bool rok = r < M;
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = n0 + col_of(tx, j);
            float v = fmaf(acc[i][j], cs[j], bias[j]);
            if (rok && e.pre && c < N) e.pre[(size_t)r * e.ldpre + c] = v;
            if (e.bn_scale) {
                v = v > 0.f ? v : v * e.slope;
                v = fmaf(v, sc[j], sh[j]);
            }
            if (c >= N) v = 0.f;
            if (rok) {
                if (e.Y && c < e.y_cols) e.Y[(size_t)r * e.ldy + c] = v;
                if (e.Yh && c < e.y_cols) {
                    const float vs = v * e.y_split_scale;
                    __half h = __float2half_rn(vs);
                    e.Yh[(size_t)r * e.ldh + c] = h;
                    if (e.Yl) e.Yl[(size_t)r * e.ldh + c] = __float2half_rn(vs - __half2float(h));
                }
                if (e.ref) {
                    float d = 0.f;
                    if (c < N) d = v - e.ref[(size_t)r * e.ldref + c];
                    if (e.dout && c < e.d_cols) e.dout[(size_t)r * e.lddout + c] = d;
                    if (e.Dh && c < e.d_cols) {
                        const float ds = d * e.d_scale;
                        __half h = __float2half_rn(ds);
                        e.Dh[(size_t)r * e.lddh + c] = h;
                        e.Dl[(size_t)r * e.lddh + c] = __float2half_rn(ds - __half2float(h));
                    }
                    sq = fmaf(d, d, sq);
                } else if (e.sq_self) {
                    sq = fmaf(v, v, sq);
                }
            }
        }
        if (e.rowpart) {
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            if (tx == 0 && rok) e.rowpart[(size_t)blockIdx.x * e.rowpart_stride + r] = sq;
        }
    }
}

template <int BM>
int launch(const GemmShape& g, const Epilogue& e, cudaStream_t s) {
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
    const bool vecA = (g.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
    const bool vecB = (g.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);
    if (!g.transA && !g.transB)
        gemm_simt_kernel<BM, false, false><<<grid, NTHREADS, 0, s>>>(g.M, g.N, g.K, g.A, g.lda, g.B, g.ldb, vecA, vecB, e);
    else if (!g.transA && g.transB)
        gemm_simt_kernel<BM, false, true><<<grid, NTHREADS, 0, s>>>(g.M, g.N, g.K, g.A, g.lda, g.B, g.ldb, vecA, vecB, e);
    else if (g.transA && !g.transB)
        gemm_simt_kernel<BM, true, false><<<grid, NTHREADS, 0, s>>>(g.M, g.N, g.K, g.A, g.lda, g.B, g.ldb, vecA, vecB, e);
    else
        gemm_simt_kernel<BM, true, true><<<grid, NTHREADS, 0, s>>>(g.M, g.N, g.K, g.A, g.lda, g.B, g.ldb, vecA, vecB, e);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

}  // namespace

int gemm_simt_tile_n() { return BN; }

int gemm_simt(const GemmShape& g, const Epilogue& e, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0) return MMAD_OK;
    if (g.M <= 64) return launch<64>(g, e, s);
    return launch<128>(g, e, s);
}

}  // namespace mmad
