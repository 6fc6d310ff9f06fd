// AutoEncoder.step (models/auto_encoder.py:57-77) on the device: train-mode forward with BatchNorm batch
// statistics (layers/fc_layer.py:37-48), summed-MSE loss (modules/loss.py:31-32), hand-written backward,
// BatchNorm running-stat update, and a multi-tensor Adam (novelty_detection.py:90).
//
// Per layer forward:   pre = in W^T + b  (GEMM)  ->  a = lrelu(pre)  ->  column sum / sum of squares over the
//                      batch (fp64)  [-> all-reduce hook: N-GPU data parallel == 1 GPU on the concatenated batch]
//                      ->  mean, inv = rsqrt(var_b + eps), running stats  ->  out = (a - mean) inv gamma + beta
// Per layer backward:  s1 = sum g, s2 = sum g xhat (fp64)  [-> hook]  ->  g_a = gamma inv (g - s1/B - xhat s2/B)
//                      ->  g_pre = g_a lrelu'(pre);  gb = sum g_pre;  gW = g_pre^T in  (GEMM);  g_in = g_pre W  (GEMM)
// Saved for backward: pre and out of every layer (xhat is recomputed from pre, mean, inv).
// Parameters / gradients / running stats are the caller's fp32 tensors (state_dict layout), used in place.
#include <algorithm>
#include <math.h>
#include <string>

#include "mmad_internal.cuh"

using namespace mmad;

namespace {

// ---- column-statistic kernels --------------------------------------------------------------------------
// One pattern for all of them: a CTA of 8 warps owns 128 columns x RS rows; lane <-> 4 consecutive columns
// (one 16-byte load per row: 512 B per warp-row, fully coalesced), warp w walks rows w, w+8, ...;
// per-thread fp64 accumulators, cross-warp reduction through shared memory, one atomic per column per CTA.
constexpr int kCB = 128;      // columns per CTA
constexpr int kCT = 256;      // threads per CTA

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

inline int row_slab(int rows) { return rows >= 2048 ? 128 : (rows >= 512 ? 64 : 32); }
inline dim3 col_grid(int n_cols, int rows) { return dim3((n_cols + kCB - 1) / kCB, (rows + row_slab(rows) - 1) / row_slab(rows)); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// sum the per-warp partials of NS statistics; thread t < 128 returns the totals of column t of this CTA
template <int NS>
__device__ __forceinline__ void cta_col_reduce(const double (&acc)[NS][4], double* sm, double (&tot)[NS]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) sm[(warp * NS + s) * kCB + lane * 4 + j] = acc[s][j];
    __syncthreads();
    if (threadIdx.x < kCB) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kCT / 32; ++w) t += sm[(w * NS + s) * kCB + threadIdx.x];
            tot[s] = t;
        }
    }
}

__device__ __forceinline__ void split4_store(const float (&v)[4], float scale, __half* __restrict__ h, __half* __restrict__ l,
                                             size_t i) {
    // saturating hi (a dead BatchNorm column can produce huge gradients; inf would poison the MMA)
    __half hh[4], ll[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float vc = fminf(fmaxf(v[j] * scale, -65504.f), 65504.f);
        hh[j] = __float2half_rn(vc);
        ll[j] = __float2half_rn(vc - __half2float(hh[j]));
    }
    *reinterpret_cast<uint2*>(h + i) = *reinterpret_cast<const uint2*>(hh);
    *reinterpret_cast<uint2*>(l + i) = *reinterpret_cast<const uint2*>(ll);
}

// st[0][c] += sum_r a, st[1][c] += sum_r a^2 with a = lrelu(pre[r,c])
__global__ void __launch_bounds__(kCT) bn_fwd_stats_kernel(const float* __restrict__ pre, int ld, int B, int N, float slope,
                                                           double* __restrict__ st, int st_stride, int RS) {
    __shared__ double sm[(kCT / 32) * 2 * kCB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * kCB + lane * 4;
    const int r_lo = blockIdx.y * RS, r_hi = min(B, r_lo + RS);
    double acc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
    if (c0 < N) {
#pragma unroll 4
        for (int r = r_lo + warp; r < r_hi; r += kCT / 32) {
            const float4 v = ld4(pre + (size_t)r * ld + c0);
            const float a[4] = {lrelu(v.x, slope), lrelu(v.y, slope), lrelu(v.z, slope), lrelu(v.w, slope)};
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[0][j] += (double)a[j]; acc[1][j] = fma((double)a[j], (double)a[j], acc[1][j]); }
        }
    }
    double tot[2];
    cta_col_reduce<2>(acc, sm, tot);
    const int c = blockIdx.x * kCB + threadIdx.x;
    if (threadIdx.x < kCB && c < N) {
        atomicAdd(&st[c], tot[0]);
        atomicAdd(&st[st_stride + c], tot[1]);
    }
}

// mean / inv-std of the (global) batch from the statistics, running-stat update by the first row slab
// (torch BatchNorm1d: momentum, unbiased running var), then out = (lrelu(pre) - mean) inv gamma + beta as fp32
// and/or fp16 hi/lo twins (tensor-core operands of the next layer and of dW).
__global__ void __launch_bounds__(kCT) bn_fwd_apply_kernel(const float* __restrict__ pre, int ld, int B, int N, float slope,
                                                           const double* __restrict__ st, int st_stride, double Bg, float eps,
                                                           float momentum, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float* __restrict__ run_mean,
                                                           float* __restrict__ run_var, long long* __restrict__ nbt,
                                                           float* __restrict__ mean_out, float* __restrict__ inv_out,
                                                           float* __restrict__ out, int ldo, __half* __restrict__ oh,
                                                           __half* __restrict__ ol, int RS) {
    __shared__ float s_mean[kCB], s_inv[kCB], s_g[kCB], s_b[kCB];
    if (threadIdx.x < kCB) {
        const int c = blockIdx.x * kCB + threadIdx.x;
        float m = 0.f, iv = 0.f, g = 0.f, b = 0.f;
        if (c < N) {
            const double md = st[c] / Bg;
            double var_b = st[st_stride + c] / Bg - md * md;
            if (var_b < 0.0) var_b = 0.0;
            m = (float)md;
            iv = (float)(1.0 / sqrt(var_b + (double)eps));
            g = gamma[c]; b = beta[c];
            if (blockIdx.y == 0) {
                mean_out[c] = m;
                inv_out[c] = iv;
                if (run_mean) {
                    const double unb = Bg > 1.0 ? var_b * (Bg / (Bg - 1.0)) : var_b;
                    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * m;
                    run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)unb;
                }
                if (c == 0 && nbt) *nbt += 1;
            }
        }
        s_mean[threadIdx.x] = m; s_inv[threadIdx.x] = iv; s_g[threadIdx.x] = g; s_b[threadIdx.x] = b;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * kCB + lane * 4;
    if (c0 >= N) return;
    float m[4], iv[4], g[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { m[j] = s_mean[lane * 4 + j]; iv[j] = s_inv[lane * 4 + j]; g[j] = s_g[lane * 4 + j]; b[j] = s_b[lane * 4 + j]; }
    const int r_lo = blockIdx.y * RS, r_hi = min(B, r_lo + RS);
#pragma unroll 4
    for (int r = r_lo + warp; r < r_hi; r += kCT / 32) {
        const float4 v = ld4(pre + (size_t)r * ld + c0);
        const float p4[4] = {v.x, v.y, v.z, v.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (c0 + j < N) ? fmaf((lrelu(p4[j], slope) - m[j]) * iv[j], g[j], b[j]) : 0.f;
        if (out) *reinterpret_cast<float4*>(out + (size_t)r * ldo + c0) = make_float4(o[0], o[1], o[2], o[3]);
        if (oh) split4_store(o, 1.f, oh, ol, (size_t)r * ldo + c0);
    }
}

// st[0][c] += sum_r g, st[1][c] += sum_r g * xhat ; the first row slab also clears the bias-gradient accumulator
__global__ void __launch_bounds__(kCT) bn_bwd_reduce_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ pre,
                                                            int ld, int B, int N, float slope, float gscale,
                                                            const float* __restrict__ mean, const float* __restrict__ inv,
                                                            double* __restrict__ st, int st_stride, float* __restrict__ gb,
                                                            int RS) {
    __shared__ double sm[(kCT / 32) * 2 * kCB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * kCB + lane * 4;
    const int r_lo = blockIdx.y * RS, r_hi = min(B, r_lo + RS);
    double acc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
    if (c0 < N) {
        float m[4], iv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const bool ok = c0 + j < N; m[j] = ok ? mean[c0 + j] : 0.f; iv[j] = ok ? inv[c0 + j] : 0.f; }
#pragma unroll 4
        for (int r = r_lo + warp; r < r_hi; r += kCT / 32) {
            const float4 gv4 = ld4(g + (size_t)r * ldg + c0);
            const float4 pv4 = ld4(pre + (size_t)r * ld + c0);
            const float gv[4] = {gv4.x * gscale, gv4.y * gscale, gv4.z * gscale, gv4.w * gscale};
            const float pv[4] = {pv4.x, pv4.y, pv4.z, pv4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float xh = (lrelu(pv[j], slope) - m[j]) * iv[j];
                acc[0][j] += (double)gv[j];
                acc[1][j] = fma((double)gv[j], (double)xh, acc[1][j]);
            }
        }
    }
    double tot[2];
    cta_col_reduce<2>(acc, sm, tot);
    const int c = blockIdx.x * kCB + threadIdx.x;
    if (threadIdx.x < kCB && c < N) {
        atomicAdd(&st[c], tot[0]);
        atomicAdd(&st[st_stride + c], tot[1]);
        if (blockIdx.y == 0) gb[c] = 0.f;
    }
}

// local parameter gradients of BatchNorm (data parallel: before the cross-rank combination of the statistics)
__global__ void bn_bwd_param_kernel(const double* __restrict__ st, int st_stride, int N, float* __restrict__ ggamma,
                                    float* __restrict__ gbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    gbeta[c] = (float)st[c];
    ggamma[c] = (float)st[st_stride + c];
}

// g_pre = gamma inv (g - s1/Bg - xhat s2/Bg) * lrelu'(pre), written as fp32 and/or fp16 twins (x twin_scale);
// gb[c] += sum_r g_pre (bias gradient); ggamma/gbeta written here when write_params (single GPU)
__global__ void __launch_bounds__(kCT) bn_bwd_apply_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ pre,
                                                           int ld, int B, int N, float slope, float gscale,
                                                           const float* __restrict__ mean, const float* __restrict__ inv,
                                                           const float* __restrict__ gamma, const double* __restrict__ st,
                                                           int st_stride, double Bg, float* __restrict__ gpre, int ldo,
                                                           __half* __restrict__ gh, __half* __restrict__ gl, float twin_scale,
                                                           float* __restrict__ gb, float* __restrict__ ggamma,
                                                           float* __restrict__ gbeta, int write_params, int RS) {
    __shared__ double sm[(kCT / 32) * 1 * kCB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * kCB + lane * 4;
    const int r_lo = blockIdx.y * RS, r_hi = min(B, r_lo + RS);
    double acc[1][4] = {{0.0, 0.0, 0.0, 0.0}};
    if (c0 < N) {
        float m[4], iv[4], k[4], a1[4], a2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool ok = c0 + j < N;
            m[j] = ok ? mean[c0 + j] : 0.f;
            iv[j] = ok ? inv[c0 + j] : 0.f;
            k[j] = ok ? gamma[c0 + j] * iv[j] : 0.f;
            a1[j] = ok ? (float)(st[c0 + j] / Bg) : 0.f;
            a2[j] = ok ? (float)(st[st_stride + c0 + j] / Bg) : 0.f;
        }
#pragma unroll 4
        for (int r = r_lo + warp; r < r_hi; r += kCT / 32) {
            const float4 gv4 = ld4(g + (size_t)r * ldg + c0);
            const float4 pv4 = ld4(pre + (size_t)r * ld + c0);
            const float gv[4] = {gv4.x * gscale, gv4.y * gscale, gv4.z * gscale, gv4.w * gscale};
            const float pv[4] = {pv4.x, pv4.y, pv4.z, pv4.w};
            float ga[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float xh = (lrelu(pv[j], slope) - m[j]) * iv[j];
                float t = k[j] * (gv[j] - a1[j] - xh * a2[j]);
                t = pv[j] > 0.f ? t : t * slope;
                ga[j] = (c0 + j < N) ? t : 0.f;
                acc[0][j] += (double)ga[j];
            }
            if (gpre) *reinterpret_cast<float4*>(gpre + (size_t)r * ldo + c0) = make_float4(ga[0], ga[1], ga[2], ga[3]);
            if (gh) split4_store(ga, twin_scale, gh, gl, (size_t)r * ldo + c0);
        }
    }
    double tot[1];
    cta_col_reduce<1>(acc, sm, tot);
    const int c = blockIdx.x * kCB + threadIdx.x;
    if (threadIdx.x < kCB && c < N) {
        atomicAdd(&gb[c], (float)tot[0]);
        if (write_params && blockIdx.y == 0) {
            gbeta[c] = (float)st[c];
            ggamma[c] = (float)st[st_stride + c];
        }
    }
}

// ---- one-kernel BatchNorm for small batches (B <= 512, single rank) -----------------------------------------------
// At B = 256 the statistics kernel and the apply kernel are each pure latency (launch + one L2 round trip + a reduction:
// 4 - 7 us for 1.4 MB), four of them per BatchNorm layer.  Columns are independent, so a CTA that owns 16 columns over ALL
// rows can do both: every thread keeps its RPT rows x 4 columns in registers (all loads in flight at once), the column sums
// are reduced through shuffles (row lanes of a warp) and shared memory (warps) in a fixed order -- no atomics, no zeroed
// accumulators -- and the normalised values are written from the registers.  Thread t: column quad t & 3, rows t >> 2 + 64 i.
constexpr int kFB = 16;       // columns per CTA of the fused kernels

template <int NS>
__device__ __forceinline__ void fused_col_reduce(double (&s)[NS][4], double* sm /* [8][NS][kFB] */, double (&tot)[NS][4]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, cq = threadIdx.x & 3;
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double v = s[k][j];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (lane < 4) sm[(warp * NS + k) * kFB + lane * 4 + j] = v;
        }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kCT / 32; ++w) t += sm[(w * NS + k) * kFB + cq * 4 + j];
            tot[k][j] = t;
        }
}

// DP: the column sums are combined across the ranks INSIDE the kernel (peer_exchange_cta: NVLink peer memory, no separate
// all-reduce launch, no second pass over the layer) -- the data-parallel step keeps the single-GPU kernel count.
struct PeerArgs { PeerPtrs P; unsigned long long* seq; unsigned int* done; int stride; double Bg; };

template <int RPT, bool DP>
__global__ void __launch_bounds__(kCT) bn_fwd_fused_kernel(const float* __restrict__ pre, int ld, int B, int N, float slope, float eps,
                                                           float momentum, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float* __restrict__ run_mean,
                                                           float* __restrict__ run_var, long long* __restrict__ nbt,
                                                           float* __restrict__ mean_out, float* __restrict__ inv_out,
                                                           float* __restrict__ out, int ldo, __half* __restrict__ oh,
                                                           __half* __restrict__ ol, const PeerArgs pa) {
    __shared__ double sm[(kCT / 32) * 2 * kFB];
    __shared__ double s_tot[2 * kFB];
    pdl_trigger();
    pdl_wait();
    const int cq = threadIdx.x & 3, rl = threadIdx.x >> 2;
    const int c0 = blockIdx.x * kFB + cq * 4;
    const bool col_ok = c0 < N;
    float a[RPT][4];
    double s[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = rl + 64 * i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok && r < B) v = ld4(pre + (size_t)r * ld + c0);
        a[i][0] = lrelu(v.x, slope); a[i][1] = lrelu(v.y, slope); a[i][2] = lrelu(v.z, slope); a[i][3] = lrelu(v.w, slope);
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[0][j] += (double)a[i][j]; s[1][j] = fma((double)a[i][j], (double)a[i][j], s[1][j]); }
    double tot[2][4];
    fused_col_reduce<2>(s, sm, tot);
    if (DP) {
        if (rl == 0) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j) s_tot[k * kFB + cq * 4 + j] = tot[k][j];
        }
        __syncthreads();
        const int col_base = blockIdx.x * kFB, stride = pa.stride;
        peer_exchange_cta<2 * kFB>(pa.P, pa.seq, pa.done, s_tot, [=](int e) { return (e / kFB) * stride + col_base + (e % kFB); });
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) tot[k][j] = s_tot[k * kFB + cq * 4 + j];
    }
    if (!col_ok) return;
    const double Bg = DP ? pa.Bg : (double)B;
    float m[4], iv[4], g[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        const bool ok = c < N;
        const double md = tot[0][j] / Bg;
        double var_b = tot[1][j] / Bg - md * md;
        if (var_b < 0.0) var_b = 0.0;
        m[j] = (float)md;
        iv[j] = (float)(1.0 / sqrt(var_b + (double)eps));
        g[j] = ok ? gamma[c] : 0.f; b[j] = ok ? beta[c] : 0.f;
        if (ok && rl == 0) {
            mean_out[c] = m[j];
            inv_out[c] = iv[j];
            if (run_mean) {
                const double unb = Bg > 1.0 ? var_b * (Bg / (Bg - 1.0)) : var_b;
                run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * m[j];
                run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)unb;
            }
            if (c == 0 && nbt) *nbt += 1;
        }
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = rl + 64 * i;
        if (r >= B) continue;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (c0 + j < N) ? fmaf((a[i][j] - m[j]) * iv[j], g[j], b[j]) : 0.f;
        if (out) *reinterpret_cast<float4*>(out + (size_t)r * ldo + c0) = make_float4(o[0], o[1], o[2], o[3]);
        if (oh) split4_store(o, 1.f, oh, ol, (size_t)r * ldo + c0);
    }
}

// backward twin: s1 = sum g, s2 = sum g xhat, g_pre = gamma inv (g - s1/B - xhat s2/B) lrelu'(pre), gb = sum g_pre,
// ggamma = s2, gbeta = s1 -- everything bn_bwd_reduce_kernel + bn_bwd_apply_kernel produce, from one read of g and pre
template <int RPT, bool DP>
__global__ void __launch_bounds__(kCT) bn_bwd_fused_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ pre, int ld,
                                                           int B, int N, float slope, float gscale, const float* __restrict__ mean,
                                                           const float* __restrict__ inv, const float* __restrict__ gamma,
                                                           float* __restrict__ gpre, int ldo, __half* __restrict__ gh,
                                                           __half* __restrict__ gl, float twin_scale, float* __restrict__ gb,
                                                           float* __restrict__ ggamma, float* __restrict__ gbeta, const PeerArgs pa) {
    __shared__ double sm[(kCT / 32) * 2 * kFB];
    __shared__ double s_tot[2 * kFB];
    pdl_trigger();
    pdl_wait();
    const int cq = threadIdx.x & 3, rl = threadIdx.x >> 2;
    const int c0 = blockIdx.x * kFB + cq * 4;
    const bool col_ok = c0 < N;
    float m[4], iv[4], k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool ok = col_ok && c0 + j < N;
        m[j] = ok ? mean[c0 + j] : 0.f;
        iv[j] = ok ? inv[c0 + j] : 0.f;
        k[j] = ok ? gamma[c0 + j] * iv[j] : 0.f;
    }
    float gv[RPT][4], xh[RPT][4];
    unsigned pos[RPT];                   // bit j: pre > 0
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = rl + 64 * i;
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), p4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok && r < B) { g4 = ld4(g + (size_t)r * ldg + c0); p4 = ld4(pre + (size_t)r * ld + c0); }
        const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
        gv[i][0] = g4.x * gscale; gv[i][1] = g4.y * gscale; gv[i][2] = g4.z * gscale; gv[i][3] = g4.w * gscale;
        pos[i] = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            xh[i][j] = (lrelu(pv[j], slope) - m[j]) * iv[j];
            pos[i] |= (pv[j] > 0.f ? 1u : 0u) << j;
        }
    }
    double s[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        if (rl + 64 * i >= B) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[0][j] += (double)gv[i][j]; s[1][j] = fma((double)gv[i][j], (double)xh[i][j], s[1][j]); }
    }
    double tot[2][4];
    fused_col_reduce<2>(s, sm, tot);
    // BatchNorm parameter gradients from the LOCAL sums (the flat gradient buffer is all-reduced as a whole afterwards)
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (col_ok && c0 + j < N && rl == 0) { gbeta[c0 + j] = (float)tot[0][j]; ggamma[c0 + j] = (float)tot[1][j]; }
    if (DP) {
        if (rl == 0) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j) s_tot[k * kFB + cq * 4 + j] = tot[k][j];
        }
        __syncthreads();
        const int col_base = blockIdx.x * kFB, stride = pa.stride;
        peer_exchange_cta<2 * kFB>(pa.P, pa.seq, pa.done, s_tot, [=](int e) { return (e / kFB) * stride + col_base + (e % kFB); });
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) tot[k][j] = s_tot[k * kFB + cq * 4 + j];
    }
    const double Bg = DP ? pa.Bg : (double)B;
    float a1[4], a2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { a1[j] = (float)(tot[0][j] / Bg); a2[j] = (float)(tot[1][j] / Bg); }
    double acc[1][4] = {{0.0, 0.0, 0.0, 0.0}};
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = rl + 64 * i;
        if (r >= B || !col_ok) continue;
        float ga[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t = k[j] * (gv[i][j] - a1[j] - xh[i][j] * a2[j]);
            t = ((pos[i] >> j) & 1u) ? t : t * slope;
            ga[j] = (c0 + j < N) ? t : 0.f;
            acc[0][j] += (double)ga[j];
        }
        if (gpre) *reinterpret_cast<float4*>(gpre + (size_t)r * ldo + c0) = make_float4(ga[0], ga[1], ga[2], ga[3]);
        if (gh) split4_store(ga, twin_scale, gh, gl, (size_t)r * ldo + c0);
    }
    __syncthreads();                     // the first reduction's partials have been read by everyone
    double tb[1][4];
    fused_col_reduce<1>(acc, sm, tb);
    if (col_ok && rl == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (c0 + j < N) gb[c0 + j] = (float)tb[0][j];
    }
}

// ---- follow-ups of the bare Linear layers at small batch --------------------------------------------------------
// With one tile per CTA the GEMM's fused epilogue is all exposed latency (12 us for the loss epilogue of the output layer on
// 28 CTAs with no split-K); a plain split-K GEMM plus one of these small kernels is half of that.
// fp16 hi / lo twins of scale * src[:, :cols] (zero padded to ldh columns... up to cols_p)
__global__ void __launch_bounds__(256) twin_split_kernel(const float* __restrict__ src, int lds, int B, int cols, int cols_p, float scale,
                                                         __half* __restrict__ h, __half* __restrict__ l, int ldh) {
    pdl_trigger();
    pdl_wait();
    const int q4 = cols_p >> 2;
    const size_t total = (size_t)B * q4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / q4), c = (int)(i % q4) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (c + 3 < cols) {
            const float4 t = ld4(src + (size_t)r * lds + c);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (c + j < cols) v[j] = src[(size_t)r * lds + c + j];
        }
        split4_store(v, scale, h, l, (size_t)r * ldh + c);
    }
}

// output layer: d = xhat - x (fp32, zero padded to cols_p), twins of twin_scale * d, rowsum[r] = sum_c d^2.
// One row per 128 threads.
__global__ void __launch_bounds__(256) loss_diff_kernel(const float* __restrict__ xhat, int ldx, const float* __restrict__ ref, int ldref,
                                                        int B, int cols, int cols_p, float* __restrict__ dout, int ldd,
                                                        __half* __restrict__ dh, __half* __restrict__ dl, int lddh, float twin_scale,
                                                        float* __restrict__ rowsum) {
    __shared__ float s_part[8];
    pdl_trigger();
    pdl_wait();
    const int half = threadIdx.x >> 7, t = threadIdx.x & 127;
    const int r = blockIdx.x * 2 + half;
    float sq = 0.f;
    if (r < B) {
        for (int c = t * 4; c < cols_p; c += 512) {
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            if (c + 3 < cols) {
                const float4 a = ld4(xhat + (size_t)r * ldx + c), b = ld4(ref + (size_t)r * ldref + c);
                d[0] = a.x - b.x; d[1] = a.y - b.y; d[2] = a.z - b.z; d[3] = a.w - b.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (c + j < cols) d[j] = xhat[(size_t)r * ldx + c + j] - ref[(size_t)r * ldref + c + j];
            }
            sq = fmaf(d[0], d[0], fmaf(d[1], d[1], fmaf(d[2], d[2], fmaf(d[3], d[3], sq))));
            *reinterpret_cast<float4*>(dout + (size_t)r * ldd + c) = make_float4(d[0], d[1], d[2], d[3]);
            if (dh) split4_store(d, twin_scale, dh, dl, (size_t)r * lddh + c);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (t == 0 && r < B) rowsum[r] = ((s_part[half * 4] + s_part[half * 4 + 1]) + s_part[half * 4 + 2]) + s_part[half * 4 + 3];
}

// The step's loss goes to the host as soon as the forward pass has it: an 8-byte (value, sequence) pair in mapped pinned
// memory (single-copy atomic; the sequence number lives on the device so that the kernel replays inside a graph).
// AutoEncoder.step returns float(loss) every step (models/auto_encoder.py:77): read through the stream that is a wait for
// backward + Adam as well, and everything the host does before the next launch is GPU idle time (~80 us of a 0.47-ms step);
// read through this pair (mmad_train_loss) the host returns while the backward pass is still running and the next step's
// launch is queued behind it.
__global__ void loss_publish_kernel(const float* __restrict__ d_loss, uint2* __restrict__ out, unsigned long long* __restrict__ seq) {
    pdl_trigger();
    pdl_wait();
    const unsigned long long s = *seq + 1;
    *seq = s;
    const float v = *d_loss;
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(out), "r"(__float_as_uint(v)), "r"((uint32_t)s) : "memory");
    __threadfence_system();
}

// gb[c] += scale * sum_r g[r,c]   (bias gradient of a bare Linear layer; gb zeroed by the caller)
__global__ void __launch_bounds__(kCT) col_sum_scaled_kernel(const float* __restrict__ g, int ldg, int B, int N, float scale,
                                                             float* __restrict__ gb, int RS) {
    __shared__ double sm[(kCT / 32) * 1 * kCB];
    pdl_trigger();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * kCB + lane * 4;
    const int r_lo = blockIdx.y * RS, r_hi = min(B, r_lo + RS);
    double acc[1][4] = {{0.0, 0.0, 0.0, 0.0}};
    if (c0 < N) {
#pragma unroll 4
        for (int r = r_lo + warp; r < r_hi; r += kCT / 32) {
            const float4 v = ld4(g + (size_t)r * ldg + c0);
            acc[0][0] += (double)v.x; acc[0][1] += (double)v.y; acc[0][2] += (double)v.z; acc[0][3] += (double)v.w;
        }
    }
    double tot[1];
    cta_col_reduce<1>(acc, sm, tot);
    const int c = blockIdx.x * kCB + threadIdx.x;
    if (threadIdx.x < kCB && c < N) atomicAdd(&gb[c], (float)(tot[0] * (double)scale));
}

// VIB forward (decorators/variational_info_bottleneck.py:19-42, k = 1): enc_out [B, 2h] -> z = eps exp(logvar/2) + mu
// (zero padded to ldz columns); kl_acc += -1/2 sum (1 + logvar - mu^2 - exp(logvar))
__global__ void vib_train_fwd_kernel(const float* __restrict__ o, int ldo, int B, int h, const float* __restrict__ eps,
                                     float* __restrict__ z, int ldz, __half* __restrict__ zh, __half* __restrict__ zl,
                                     double* __restrict__ kl_acc) {
    const size_t total = (size_t)B * ldz;
    double kl = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / ldz), j = (int)(i % ldz);
        float v = 0.f;
        if (j < h) {
            const float m = o[(size_t)b * ldo + j], lv = o[(size_t)b * ldo + h + j];
            v = fmaf(eps[(size_t)b * h + j], expf(0.5f * lv), m);
            kl += -0.5 * (1.0 + (double)lv - (double)m * (double)m - exp((double)lv));
        }
        z[i] = v;
        if (zh) {
            const __half hh = __float2half_rn(v);
            zh[i] = hh;
            zl[i] = __float2half_rn(v - __half2float(hh));
        }
    }
    __shared__ double sm[256];
    sm[threadIdx.x] = kl;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0 && sm[0] != 0.0) atomicAdd(kl_acc, sm[0]);
}

// VIB backward: g_out[:, :h] = g_z + beta mu ;  g_out[:, h:] = g_z eps exp(logvar/2)/2 - beta (1 - exp(logvar))/2
__global__ void vib_train_bwd_kernel(const float* __restrict__ gz, int ldg, const float* __restrict__ o, int ldo, int B, int h,
                                     const float* __restrict__ eps, float beta, float* __restrict__ gout, int ldgo,
                                     __half* __restrict__ gh, __half* __restrict__ gl, float twin_scale) {
    const size_t total = (size_t)B * h;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / h), j = (int)(i % h);
        const float m = o[(size_t)b * ldo + j], lv = o[(size_t)b * ldo + h + j];
        const float g = gz[(size_t)b * ldg + j];
        const float sg = expf(0.5f * lv);
        const float g_mu = fmaf(beta, m, g);
        const float g_lv = g * eps[i] * 0.5f * sg - 0.5f * beta * (1.f - expf(lv));
        gout[(size_t)b * ldgo + j] = g_mu;
        gout[(size_t)b * ldgo + h + j] = g_lv;
        if (gh) {
            const float a = fminf(fmaxf(g_mu * twin_scale, -65504.f), 65504.f), c = fminf(fmaxf(g_lv * twin_scale, -65504.f), 65504.f);
            const __half ha = __float2half_rn(a), hc = __float2half_rn(c);
            gh[(size_t)b * ldgo + j] = ha; gl[(size_t)b * ldgo + j] = __float2half_rn(a - __half2float(ha));
            gh[(size_t)b * ldgo + h + j] = hc; gl[(size_t)b * ldgo + h + j] = __float2half_rn(c - __half2float(hc));
        }
    }
}

__global__ void loss_finish_kernel(float* __restrict__ loss, const double* __restrict__ kl_acc, float beta) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *loss += beta * (float)(*kl_acc);
}

// ---- multi-tensor Adam ----------------------------------------------------------------------------
constexpr int kAdamMaxTensors = 48;
constexpr int kAdamChunk = 4096;     // elements per block
struct AdamArgs {
    float* p[kAdamMaxTensors];
    const float* g[kAdamMaxTensors];
    float* m[kAdamMaxTensors];
    float* v[kAdamMaxTensors];
    long long numel[kAdamMaxTensors];
    int block_start[kAdamMaxTensors + 1];   // first block of every tensor
    int n;
};

// torch.optim.Adam single-tensor formulation (torch/optim/adam.py _single_tensor_adam):
//   m.lerp_(g, 1-b1);  v = b2 v + (1-b2) g g;  denom = sqrt(v)/sqrt(bc2) + eps;  p += -(lr/bc1) * (m/denom)
__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamArgs a, float w1, float b2, float omb2,
                                                   float bc2_sqrt, float eps, float neg_step, float gscale) {
    int t = 0;
    while (t + 1 < a.n && (int)blockIdx.x >= a.block_start[t + 1]) ++t;
    const long long base = (long long)(blockIdx.x - a.block_start[t]) * kAdamChunk;
    const long long end = min(a.numel[t], base + kAdamChunk);
    float* __restrict__ p = a.p[t];
    const float* __restrict__ g = a.g[t];
    float* __restrict__ m = a.m[t];
    float* __restrict__ v = a.v[t];
    auto upd = [&](float pv, float gr, float& mv, float& vv) -> float {
        const float gv = gr * gscale;
        mv = mv + w1 * (gv - mv);
        vv = vv * b2 + omb2 * gv * gv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        return pv + neg_step * (mv / denom);
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec) {      // 16-byte accesses: four independent element updates per thread per iteration
        const long long end4 = base + ((end - base) & ~3LL);
        // two groups per thread per trip, all eight loads issued before the first update: the kernel is one pass over
        // 7 x 4 bytes per parameter and lives on the number of bytes in flight
        const long long stride = blockDim.x * 4LL;
        long long i = base + threadIdx.x * 4LL;
        for (; i + stride < end4; i += 2 * stride) {
            float4 pv0 = *reinterpret_cast<const float4*>(p + i), pv1 = *reinterpret_cast<const float4*>(p + i + stride);
            const float4 gr0 = *reinterpret_cast<const float4*>(g + i), gr1 = *reinterpret_cast<const float4*>(g + i + stride);
            float4 mv0 = *reinterpret_cast<const float4*>(m + i), mv1 = *reinterpret_cast<const float4*>(m + i + stride);
            float4 vv0 = *reinterpret_cast<const float4*>(v + i), vv1 = *reinterpret_cast<const float4*>(v + i + stride);
            pv0.x = upd(pv0.x, gr0.x, mv0.x, vv0.x); pv0.y = upd(pv0.y, gr0.y, mv0.y, vv0.y);
            pv0.z = upd(pv0.z, gr0.z, mv0.z, vv0.z); pv0.w = upd(pv0.w, gr0.w, mv0.w, vv0.w);
            pv1.x = upd(pv1.x, gr1.x, mv1.x, vv1.x); pv1.y = upd(pv1.y, gr1.y, mv1.y, vv1.y);
            pv1.z = upd(pv1.z, gr1.z, mv1.z, vv1.z); pv1.w = upd(pv1.w, gr1.w, mv1.w, vv1.w);
            *reinterpret_cast<float4*>(m + i) = mv0; *reinterpret_cast<float4*>(m + i + stride) = mv1;
            *reinterpret_cast<float4*>(v + i) = vv0; *reinterpret_cast<float4*>(v + i + stride) = vv1;
            *reinterpret_cast<float4*>(p + i) = pv0; *reinterpret_cast<float4*>(p + i + stride) = pv1;
        }
        for (; i < end4; i += stride) {
            float4 pv = *reinterpret_cast<const float4*>(p + i);
            const float4 gr = *reinterpret_cast<const float4*>(g + i);
            float4 mv = *reinterpret_cast<const float4*>(m + i);
            float4 vv = *reinterpret_cast<const float4*>(v + i);
            pv.x = upd(pv.x, gr.x, mv.x, vv.x); pv.y = upd(pv.y, gr.y, mv.y, vv.y);
            pv.z = upd(pv.z, gr.z, mv.z, vv.z); pv.w = upd(pv.w, gr.w, mv.w, vv.w);
            *reinterpret_cast<float4*>(m + i) = mv;
            *reinterpret_cast<float4*>(v + i) = vv;
            *reinterpret_cast<float4*>(p + i) = pv;
        }
        for (long long i = end4 + threadIdx.x; i < end; i += blockDim.x) {
            float mv = m[i], vv = v[i];
            p[i] = upd(p[i], g[i], mv, vv);
            m[i] = mv; v[i] = vv;
        }
        return;
    }
    for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
        float mv = m[i], vv = v[i];
        p[i] = upd(p[i], g[i], mv, vv);
        m[i] = mv; v[i] = vv;
    }
}

struct TrainPlan {
    size_t total = 0;
    int B = 0;
    size_t pre[2][MMAD_MAX_LAYERS] = {{0}};
    size_t out[2][MMAD_MAX_LAYERS] = {{0}};
    size_t mean[2][MMAD_MAX_LAYERS] = {{0}};
    size_t inv[2][MMAD_MAX_LAYERS] = {{0}};
    size_t st[2][MMAD_MAX_LAYERS] = {{0}};    // [4][Np] doubles per layer: forward sum / sum sq, backward sum g / sum g xhat
    size_t st_all = 0, st_bytes = 0;
    size_t pre_all = 0, pre_bytes = 0;
    size_t g[2] = {0, 0};                      // gradient ping-pong [B, maxNp]
    size_t gz[2 * MMAD_MAX_LAYERS] = {0};      // B <= 512 (tensor-core path): one input-gradient buffer per layer inside the zeroed region,
    bool has_gz = false;                       // so the split-K dX GEMMs need no memset node between the kernels of the step
    size_t z = 0, genc = 0, eps = 0;           // VIB: sampled code, gradient wrt the encoder output, staged noise
    size_t rowpart = 0;
    size_t kl = 0;
    int maxNp = 0;
    // tensor-core modes: fp16 hi/lo twins
    size_t xh = 0, xl = 0, xp = 0;
    size_t outh[2][MMAD_MAX_LAYERS] = {{0}}, outl[2][MMAD_MAX_LAYERS] = {{0}};
    size_t gth[2 * MMAD_MAX_LAYERS] = {0}, gtl[2 * MMAD_MAX_LAYERS] = {0};   // g_pre twins of every layer (forward order index):
                                                                              // own buffers, so dW can run on a second stream
    size_t zh = 0, zl = 0, gench = 0, gencl = 0;   // VIB twins: sampled code, gradient wrt the encoder output
};

int np_of(int n) { return round_up(n, kPad); }

TrainPlan make_train_plan(const mmad_desc_t& d, int B, bool tc) {
    TrainPlan p;
    p.B = B;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t at = off; off += round_up_sz(bytes, 256); return at; };
    int maxNp = np_of(d.enc_widths[0]);
    for (int i = 1; i <= d.n_enc; ++i) maxNp = std::max(maxNp, np_of(d.enc_widths[i]));
    for (int i = 0; i <= d.n_dec; ++i) maxNp = std::max(maxNp, np_of(d.dec_widths[i]));
    p.maxNp = maxNp;
    p.st_all = off;
    for (int m = 0; m < 2; ++m) {
        const int n = m == 0 ? d.n_enc : d.n_dec;
        const int* w = m == 0 ? d.enc_widths : d.dec_widths;
        for (int i = 0; i < n; ++i) p.st[m][i] = take((size_t)4 * np_of(w[i + 1]) * 8);
    }
    p.kl = take(8);
    p.st_bytes = off - p.st_all;
    p.pre_all = off;                    // every layer's pre-activation buffer, contiguous: one memset zeroes them
    for (int m = 0; m < 2; ++m) {
        const int n = m == 0 ? d.n_enc : d.n_dec;
        const int* w = m == 0 ? d.enc_widths : d.dec_widths;
        for (int i = 0; i < n; ++i) p.pre[m][i] = take((size_t)B * np_of(w[i + 1]) * 4);
    }
    if (tc && B <= 512) {
        for (int i = 0; i < d.n_enc + d.n_dec; ++i) p.gz[i] = take((size_t)B * maxNp * 4);
        p.has_gz = true;
    }
    p.pre_bytes = off - p.pre_all;
    for (int m = 0; m < 2; ++m) {
        const int n = m == 0 ? d.n_enc : d.n_dec;
        const int* w = m == 0 ? d.enc_widths : d.dec_widths;
        for (int i = 0; i < n; ++i) {
            const int Np = np_of(w[i + 1]);
            const bool bn = i < n - 1;
            p.out[m][i] = bn ? take((size_t)B * Np * 4) : p.pre[m][i];
            if (bn) { p.mean[m][i] = take((size_t)Np * 4); p.inv[m][i] = take((size_t)Np * 4); }
        }
    }
    p.g[0] = take((size_t)B * maxNp * 4);
    p.g[1] = take((size_t)B * maxNp * 4);
    p.z = take((size_t)B * np_of(d.dec_widths[0]) * 4);
    p.genc = take((size_t)B * np_of(d.enc_widths[d.n_enc]) * 4);
    p.eps = take((size_t)B * np_of(d.dec_widths[0]) * 4);
    p.rowpart = take((size_t)((d.enc_widths[0] + 63) / 64 + 1) * B * 4);
    p.xp = take((size_t)B * np_of(d.enc_widths[0]) * 4);
    if (tc) {
        p.xh = take((size_t)B * np_of(d.enc_widths[0]) * 2);
        p.xl = take((size_t)B * np_of(d.enc_widths[0]) * 2);
        for (int m = 0; m < 2; ++m) {
            const int n = m == 0 ? d.n_enc : d.n_dec;
            const int* w = m == 0 ? d.enc_widths : d.dec_widths;
            for (int i = 0; i < n; ++i) {
                p.outh[m][i] = take((size_t)B * np_of(w[i + 1]) * 2);
                p.outl[m][i] = take((size_t)B * np_of(w[i + 1]) * 2);
            }
        }
        for (int i = 0; i < d.n_enc + d.n_dec; ++i) { p.gth[i] = take((size_t)B * maxNp * 2); p.gtl[i] = take((size_t)B * maxNp * 2); }
        p.zh = take((size_t)B * np_of(d.dec_widths[0]) * 2);
        p.zl = take((size_t)B * np_of(d.dec_widths[0]) * 2);
        p.gench = take((size_t)B * np_of(d.enc_widths[d.n_enc]) * 2);
        p.gencl = take((size_t)B * np_of(d.enc_widths[d.n_enc]) * 2);
    }
    p.total = off;
    return p;
}

inline int ew_grid(size_t total) {
    size_t g = (total + 255) / 256;
    return (int)(g > 148 * 8 ? 148 * 8 : (g == 0 ? 1 : g));
}

}  // namespace

extern "C" {

size_t mmad_train_workspace_bytes(mmad_t h, int batch) {
    if (!h || batch < 1) return 0;
    return make_train_plan(*handle_desc(h), batch, tc_available() != 0).total;
}

// The whole step as a sequence of launches on stream s.  Inputs were staged into the workspace by the caller
// (x as padded fp32 + fp16 twins, VIB noise), so every pointer used here is stable from step to step and the
// sequence can be captured into a CUDA graph.
static int train_body(mmad_t h, const mmad_desc_t& d, const TrainPlan& p, bool tc, bool vib, int batch, long long global_batch,
                      const mmad_train_layer_t* enc, const mmad_train_layer_t* dec, float beta_kl, float bn_momentum,
                      float* d_loss, char* ws, mmad_allreduce_fn allreduce, void* allreduce_ctx, cudaStream_t s,
                      cudaStream_t s2, cudaEvent_t ev_fork, cudaEvent_t ev_join, bool publish) {
    bool forked = false;
    // cross-rank combination of the BatchNorm statistics: caller's hook, or the handle's own NCCL communicator
    void* comm_p = nullptr; int comm_world = 1;
    handle_comm(h, &comm_p, &comm_world);
    const bool dist = allreduce != nullptr || (comm_p && comm_world > 1);
    auto stats_allreduce = [&](double* buf, long long cnt) -> int {
        if (allreduce) {
            const int rc = allreduce(allreduce_ctx, buf, cnt, s);
            if (rc) { set_error("allreduce hook failed (%d)", rc); return MMAD_E_STATE; }
            return MMAD_OK;
        }
        return comm_allreduce(h, buf, cnt, true, s);
    };
    const int D = d.enc_widths[0];
    const int enc_out = d.enc_widths[d.n_enc], dec_in = d.dec_widths[0];
    const int passes = d.precision == MMAD_PREC_F16 ? 1 : 3;   // F16F8 is a scoring mode: training keeps the full fp16 split
    const float* d_eps = vib ? (const float*)(ws + p.eps) : nullptr;
    const int B = batch;
    const double Bg = (double)global_batch;
    const float slope = d.lrelu_slope;
    constexpr float GS = 16.f;       // gradients are scaled by 2^4 before the fp16 hi/lo split
    // one-kernel BatchNorm (statistics + apply) when one rank holds the whole batch and it fits the register tile
    static const bool no_fuse_bn = getenv("MMAD_NO_FUSED_BN") != nullptr;
    // Data parallel over the handle's peer buffers: the same kernels with the cross-rank combination of the column sums inside
    // (no hook, every rank holds B rows, the statistic vectors fit the exchange buffer).
    PeerArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.Bg = Bg;
    int max_np = 0;
    for (int i = 1; i <= d.n_enc; ++i) max_np = std::max(max_np, np_of(d.enc_widths[i]));
    for (int i = 1; i <= d.n_dec; ++i) max_np = std::max(max_np, np_of(d.dec_widths[i]));
    const bool fuse_dp = dist && !allreduce && !no_fuse_bn && B <= 512 && global_batch == (long long)B * comm_world &&
                         2 * max_np <= kPeerMaxDoubles && peer_kernel_args(h, &pa.P, &pa.seq, &pa.done);
    const bool fuse_bn = fuse_dp || (!dist && B >= 128 && B <= 512 && global_batch == (long long)B && !no_fuse_bn);   // below 128 rows the two-kernel form measured faster
    if (!fuse_bn) MMAD_CUDA_OK(cudaMemsetAsync(ws + p.st_all, 0, p.st_bytes, s));
    else if (vib) MMAD_CUDA_OK(cudaMemsetAsync(ws + p.kl, 0, 8, s));        // the KL accumulator shares that region
    MMAD_CUDA_OK(cudaMemsetAsync(d_loss, 0, 4, s));

    // an activation / gradient matrix: fp32 and (tensor-core modes) fp16 hi/lo twins, same leading dimension
    struct Mat { const float* f; const __half* h; const __half* l; int ld; };
    struct Ref { int m, i, K, N, Np; const mmad_train_layer_t* L; Mat in; bool bn; };
    std::vector<Ref> order;

    // C[M,N] = A B^T-like product on the tensor cores.  a_mn / b_mn: operand stored [contraction, M|N].
    auto tc_gemm = [&](const __half* Ah, const __half* Al, int lda, bool a_mn, const __half* Bh, const __half* Bl, int ldb,
                       bool b_mn, int M, int N, int K, const Epilogue& e, cudaStream_t s) -> int {
        TcOperand A, Bo;
        int rc;
        if (B >= 2048 && tc2_available() && (a_mn ? b_mn : true)) {      // large batch: CTA pairs (3-stage pipeline)
            if (a_mn) { rc = tc_make_operand_map(&A.hi, Ah, K, M, lda, 64); if (!rc) rc = tc_make_operand_map(&A.lo, Al, K, M, lda, 64); }
            else { rc = tc_make_operand_map(&A.hi, Ah, M, K, lda, 128); if (!rc) rc = tc_make_operand_map(&A.lo, Al, M, K, lda, 128); }
            if (rc) return rc;
            if (b_mn) { rc = tc_make_operand_map(&Bo.hi, Bh, K, N, ldb, 64); if (!rc) rc = tc_make_operand_map(&Bo.lo, Bl, K, N, ldb, 64); }
            else { rc = tc_make_operand_map(&Bo.hi, Bh, N, K, ldb, 128); if (!rc) rc = tc_make_operand_map(&Bo.lo, Bl, N, K, ldb, 128); }
            if (rc) return rc;
            A.mn = a_mn; Bo.mn = b_mn;
            return gemm_tc2(A, Bo, M, N, K, passes, e, s);
        }
        const int bn = gemm_tc_pick_bn(M, N);
        if (a_mn) { rc = tc_make_operand_map(&A.hi, Ah, K, M, lda, 64); if (!rc) rc = tc_make_operand_map(&A.lo, Al, K, M, lda, 64); }
        else { rc = tc_make_operand_map(&A.hi, Ah, M, K, lda, 128); if (!rc) rc = tc_make_operand_map(&A.lo, Al, M, K, lda, 128); }
        if (rc) return rc;
        if (b_mn) { rc = tc_make_operand_map(&Bo.hi, Bh, K, N, ldb, 64); if (!rc) rc = tc_make_operand_map(&Bo.lo, Bl, K, N, ldb, 64); }
        else { rc = tc_make_operand_map(&Bo.hi, Bh, N, K, ldb, bn); if (!rc) rc = tc_make_operand_map(&Bo.lo, Bl, N, K, ldb, bn); }
        if (rc) return rc;
        A.mn = a_mn; Bo.mn = b_mn;
        return gemm_tc(A, Bo, M, N, K, passes, e, s, bn);
    };

    // ------------------------------- forward -------------------------------
    const int Dp = np_of(D);
    const float* xref = (const float*)(ws + p.xp);      // staged input: reference of the loss epilogue
    const int ldxref = Dp;
    Mat cur{xref, tc ? (const __half*)(ws + p.xh) : nullptr, tc ? (const __half*)(ws + p.xl) : nullptr, Dp};
    bool gw_zeroed = false;
    if (tc) {
        // weights changed since the last step: refresh every layer's fp16 twins in one launch
        const float* Wp[2 * MMAD_MAX_LAYERS]; __half* Whp[2 * MMAD_MAX_LAYERS]; __half* Wlp[2 * MMAD_MAX_LAYERS];
        int Ns[2 * MMAD_MAX_LAYERS], Ks[2 * MMAD_MAX_LAYERS], Kps[2 * MMAD_MAX_LAYERS], nt = 0;
        float wscale = 256.f;
        uintptr_t g_lo = ~(uintptr_t)0, g_hi = 0;
        size_t g_sum = 0;
        int g_cnt = 0;
        for (int m = 0; m < 2; ++m) {
            const int n = m == 0 ? d.n_enc : d.n_dec;
            const mmad_train_layer_t* Ls = m == 0 ? enc : dec;
            for (int i = 0; i < n; ++i, ++nt) {
                const LayerView lv = handle_layer(h, m, i);
                Wp[nt] = Ls[i].W; Whp[nt] = lv.Wh; Wlp[nt] = lv.Wl; Ns[nt] = lv.N; Ks[nt] = lv.K; Kps[nt] = lv.Kp;
                wscale = lv.wscale;
                auto span = [&](const float* q, size_t count) {
                    if (!q) return;
                    const uintptr_t a0 = reinterpret_cast<uintptr_t>(q);
                    g_lo = std::min(g_lo, a0); g_hi = std::max(g_hi, a0 + count * 4); g_sum += count * 4; ++g_cnt;
                };
                span(Ls[i].gW, (size_t)lv.N * lv.K); span(Ls[i].gb, lv.N);
                if (i < n - 1) { span(Ls[i].ggamma, lv.N); span(Ls[i].gbeta, lv.N); }
            }
        }
        // split-K outputs start from zero: the pre-activation buffers (contiguous) and, when ALL gradient tensors
        // tile one flat buffer exactly (they do in the Python layer), that buffer -- two memsets instead of one per GEMM.
        // All memsets come first: from here on the step is an unbroken chain of kernels (programmatic dependent launches).
        MMAD_CUDA_OK(cudaMemsetAsync(ws + p.pre_all, 0, p.pre_bytes, s));
        if (g_hi - g_lo <= g_sum + 12 * (size_t)g_cnt) {      // the gradient tensors tile one buffer (gaps < 16 B are alignment padding)
            MMAD_CUDA_OK(cudaMemsetAsync(reinterpret_cast<void*>(g_lo), 0, g_hi - g_lo, s));
            gw_zeroed = true;
        }
        int rc = split_weights_multi(nt, Wp, Ns, Ks, Kps, wscale, Whp, Wlp, s);
        if (rc) return rc;
    }
    // programmatic dependent launches for the kernels of the main stream from here on (mmad_internal.cuh)
    static const bool no_pdl = getenv("MMAD_NO_PDL") != nullptr;
    PdlScope pdl_scope(tc && !no_pdl);
    const int loss_tile_n = tc ? gemm_tc_rowpart_cols() : gemm_simt_tile_n();
    int loss_slots = 0;               // > 0: the loss follow-up kernel wrote whole-row sums
    for (int m = 0; m < 2; ++m) {
        const int n = m == 0 ? d.n_enc : d.n_dec;
        const int* w = m == 0 ? d.enc_widths : d.dec_widths;
        const mmad_train_layer_t* Ls = m == 0 ? enc : dec;
        if (m == 1 && vib) {
            const int hp = np_of(dec_in);
            __half* zh = tc ? (__half*)(ws + p.zh) : nullptr;
            __half* zl = tc ? (__half*)(ws + p.zl) : nullptr;
            vib_train_fwd_kernel<<<ew_grid((size_t)B * hp), 256, 0, s>>>(cur.f, cur.ld, B, dec_in, d_eps, (float*)(ws + p.z), hp, zh, zl,
                                                                        (double*)(ws + p.kl));
            MMAD_LAUNCHED();
            cur = Mat{(const float*)(ws + p.z), zh, zl, hp};
        }
        for (int i = 0; i < n; ++i) {
            const int K = w[i], N = w[i + 1], Np = np_of(N);
            const bool bn = i < n - 1;
            const mmad_train_layer_t& L = Ls[i];
            order.push_back({m, i, K, N, Np, &L, cur, bn});
            float* pre = (float*)(ws + p.pre[m][i]);
            __half* oh = tc ? (__half*)(ws + p.outh[m][i]) : nullptr;
            __half* ol = tc ? (__half*)(ws + p.outl[m][i]) : nullptr;
            Epilogue e;
            e.bias = L.b;
            e.slope = slope;
            const bool last = (m == 1 && i == n - 1);
            if (bn) {
                e.pre = pre; e.ldpre = Np;
            } else {
                e.Y = pre; e.ldy = Np; e.y_cols = Np;
                if (tc && !last) { e.Yh = oh; e.Yl = ol; e.ldh = Np; }
                if (last) {   // loss epilogue: d = xhat - x, row sums of d^2; g = 2 d enters the backward pass
                    e.ref = xref; e.ldref = ldxref;
                    e.dout = (float*)(ws + p.g[0]); e.lddout = p.maxNp; e.d_cols = tc ? Np : N;
                    if (tc) { const int li = d.n_enc + d.n_dec - 1; e.Dh = (__half*)(ws + p.gth[li]); e.Dl = (__half*)(ws + p.gtl[li]); e.lddh = p.maxNp; e.d_scale = 2.f * GS; e.Y = nullptr; }
                    e.rowpart = (float*)(ws + p.rowpart); e.rowpart_stride = B;
                }
            }
            int rc;
            if (tc) {
                const LayerView lv = handle_layer(h, m, i);
                e.acc_scale = 1.f / lv.wscale;
                const bool small = !(B >= 2048 && tc2_available());
                if (small) {   // trivial epilogue (bias + store): plain mode, split-K when the tile count is small
                    Epilogue pe;
                    pe.bias = L.b; pe.slope = slope; pe.acc_scale = e.acc_scale;
                    pe.Y = pre; pe.ldy = Np; pe.y_cols = N; pe.plain = 1; pe.split_k_ok = 1; pe.pre_zeroed = 1;
                    rc = tc_gemm(cur.h, cur.l, cur.ld, false, lv.Wh, lv.Wl, lv.Kp, false, B, N, K, pe, s);
                    if (!rc && !bn && !last) {       // twins of the bare Linear output for the next layer
                        MMAD_CUDA_OK(launch_k(twin_split_kernel, dim3(ew_grid((size_t)B * Np / 4)), dim3(256), 0, s, pre, Np, B, N, Np, 1.f, oh, ol, Np));
                        MMAD_LAUNCHED();
                    }
                    if (!rc && last) {               // d = xhat - x, its twins (x 2 GS), row sums of d^2
                        const int li = d.n_enc + d.n_dec - 1;
                        MMAD_CUDA_OK(launch_k(loss_diff_kernel, dim3((B + 1) / 2), dim3(256), 0, s, pre, Np, xref, ldxref, B, N, Np, (float*)(ws + p.g[0]),
                                              p.maxNp, (__half*)(ws + p.gth[li]), (__half*)(ws + p.gtl[li]), p.maxNp, 2.f * GS,
                                              (float*)(ws + p.rowpart)));
                        MMAD_LAUNCHED();
                        loss_slots = 1;
                    }
                } else {
                    rc = tc_gemm(cur.h, cur.l, cur.ld, false, lv.Wh, lv.Wl, lv.Kp, false, B, N, K, e, s);
                }
            } else {
                GemmShape g;
                g.M = B; g.N = N; g.K = K; g.A = cur.f; g.lda = cur.ld; g.B = L.W; g.ldb = K;
                rc = gemm_simt(g, e, s);
            }
            if (rc) return rc;
            if (bn) {
                double* st = (double*)(ws + p.st[m][i]);
                float* mean = (float*)(ws + p.mean[m][i]);
                float* inv = (float*)(ws + p.inv[m][i]);
                float* out = (float*)(ws + p.out[m][i]);
                if (fuse_bn) {
#define MMAD_BN_FWD2(R, DPV) MMAD_CUDA_OK(launch_k(bn_fwd_fused_kernel<R, DPV>, dim3((N + kFB - 1) / kFB), dim3(kCT), 0, s, pre, Np, B, N, slope, d.bn_eps, bn_momentum, \
                           L.gamma, L.beta, L.run_mean, L.run_var, L.num_batches_tracked, mean, inv, tc ? nullptr : out, Np, oh, ol, pa))
#define MMAD_BN_FWD(R) do { if (fuse_dp) MMAD_BN_FWD2(R, true); else MMAD_BN_FWD2(R, false); } while (0)
                    pa.stride = Np;
                    if (B <= 64) MMAD_BN_FWD(1); else if (B <= 128) MMAD_BN_FWD(2); else if (B <= 256) MMAD_BN_FWD(4); else MMAD_BN_FWD(8);
#undef MMAD_BN_FWD2
#undef MMAD_BN_FWD
                    MMAD_LAUNCHED();
                } else {
                    bn_fwd_stats_kernel<<<col_grid(N, B), kCT, 0, s>>>(pre, Np, B, N, slope, st, Np, row_slab(B));
                    MMAD_LAUNCHED();
                    if (dist) {
                        rc = stats_allreduce(st, 2LL * Np);
                        if (rc) return rc;
                    }
                    bn_fwd_apply_kernel<<<col_grid(N, B), kCT, 0, s>>>(pre, Np, B, N, slope, st, Np, Bg, d.bn_eps, bn_momentum, L.gamma,
                                                                       L.beta, L.run_mean, L.run_var, L.num_batches_tracked, mean, inv,
                                                                       tc ? nullptr : out, Np, oh, ol, row_slab(B));
                    MMAD_LAUNCHED();
                }
                cur = Mat{out, oh, ol, Np};
            } else {
                cur = Mat{pre, oh, ol, Np};
            }
        }
    }
    {   // loss = sum d^2 (+ beta KL)
        const int slots = loss_slots ? loss_slots : (D + loss_tile_n - 1) / loss_tile_n;
        int rc = reduce_sum_all((const float*)(ws + p.rowpart), B, B, 0, slots, d_loss, s);
        if (rc) return rc;
        if (vib) {
            loss_finish_kernel<<<1, 32, 0, s>>>(d_loss, (const double*)(ws + p.kl), beta_kl);
            MMAD_LAUNCHED();
        }
        uint2* d_pair = nullptr; unsigned long long* d_seq = nullptr;
        if (publish && !handle_loss_doorbell(h, &d_pair, &d_seq, false)) {      // (allocated by mmad_train_fwd_bwd, outside any capture)
            MMAD_CUDA_OK(launch_k(loss_publish_kernel, dim3(1), dim3(1), 0, s, (const float*)d_loss, d_pair, d_seq));
            MMAD_LAUNCHED();
        }
    }
    // Data parallel with the handle's communicator: the gradients are all-reduced in two buckets on the second stream -- the
    // decoder's parameters (one contiguous block of the flat gradient buffer) as soon as the decoder's backward pass is done,
    // overlapping the encoder's backward pass; the encoder's at the end.  (One all-reduce per LAYER was measured slower than a
    // single flat one: ten medium collectives cost more latency than they hide.)
    const bool grad_ar = !allreduce && comm_p && comm_world > 1 && handle_grad_allreduce(h);
    auto module_grad_allreduce = [&](int m, cudaStream_t st) -> int {
        const mmad_train_layer_t* Ls = m == 0 ? enc : dec;
        const int n = m == 0 ? d.n_enc : d.n_dec;
        const int* w = m == 0 ? d.enc_widths : d.dec_widths;
        uintptr_t lo = ~(uintptr_t)0, hi = 0; size_t sum = 0; int cnt = 0;
        for (int i = 0; i < n; ++i) {
            const bool bn = i < n - 1;
            const float* ptrs[4] = {Ls[i].gW, Ls[i].gb, bn ? Ls[i].ggamma : nullptr, bn ? Ls[i].gbeta : nullptr};
            const size_t cnts[4] = {(size_t)w[i + 1] * w[i], (size_t)w[i + 1], (size_t)w[i + 1], (size_t)w[i + 1]};
            for (int k = 0; k < 4; ++k) if (ptrs[k]) {
                const uintptr_t a0 = reinterpret_cast<uintptr_t>(ptrs[k]);
                lo = std::min(lo, a0); hi = std::max(hi, a0 + cnts[k] * 4); sum += cnts[k] * 4; ++cnt;
            }
        }
        if (hi - lo <= sum + 12 * (size_t)cnt)      // one contiguous block (alignment padding only)
            return comm_allreduce(h, reinterpret_cast<void*>(lo), (long long)((hi - lo) / 4), false, st);
        for (int i = 0; i < n; ++i) {               // scattered parameters: tensor by tensor
            const bool bn = i < n - 1;
            float* ptrs[4] = {Ls[i].gW, Ls[i].gb, bn ? Ls[i].ggamma : nullptr, bn ? Ls[i].gbeta : nullptr};
            const size_t cnts[4] = {(size_t)w[i + 1] * w[i], (size_t)w[i + 1], (size_t)w[i + 1], (size_t)w[i + 1]};
            for (int k = 0; k < 4; ++k) if (ptrs[k]) {
                const int rc = comm_allreduce(h, ptrs[k], (long long)cnts[k], false, st);
                if (rc) return rc;
            }
        }
        return MMAD_OK;
    };
    // ------------------------------- backward -------------------------------
    // g (ping) holds d = xhat - x in fp32 (and 2 d GS as twins); dL/dxhat = 2 d enters through gscale
    int gi = 0;
    const float* gcur = (const float*)(ws + p.g[0]);     // fp32 gradient entering the current layer
    float gscale = 2.f;
    for (int idx = (int)order.size() - 1; idx >= 0; --idx) {
        const Ref& r = order[idx];
        const mmad_train_layer_t& L = *r.L;
        double* st = (double*)(ws + p.st[r.m][r.i]);
        Mat gin{gcur, tc ? (const __half*)(ws + p.gth[idx]) : nullptr,
                tc ? (const __half*)(ws + p.gtl[idx]) : nullptr, p.maxNp};
        if (vib && r.m == 0 && r.i == d.n_enc - 1)
            gin = Mat{(const float*)(ws + p.genc), tc ? (const __half*)(ws + p.gench) : nullptr,
                      tc ? (const __half*)(ws + p.gencl) : nullptr, np_of(enc_out)};
        Mat gpre = gin;
        float gemm_scale = gscale;       // CUDA-core path: factor still to be applied to g_pre
        if (r.bn) {
            const float* pre = (const float*)(ws + p.pre[r.m][r.i]);
            const float* mean = (const float*)(ws + p.mean[r.m][r.i]);
            const float* inv = (const float*)(ws + p.inv[r.m][r.i]);
            float* go = (float*)(ws + p.g[gi ^ 1]);
            __half* goh = tc ? (__half*)(ws + p.gth[idx]) : nullptr;
            __half* gol = tc ? (__half*)(ws + p.gtl[idx]) : nullptr;
            double* stb = st + 2 * r.Np;       // backward statistics (zeroed with the forward ones at step start)
            if (fuse_bn) {
#define MMAD_BN_BWD2(R, DPV) MMAD_CUDA_OK(launch_k(bn_bwd_fused_kernel<R, DPV>, dim3((r.N + kFB - 1) / kFB), dim3(kCT), 0, s, gin.f, gin.ld, pre, r.Np, B, r.N, slope, \
                           gscale, mean, inv, L.gamma, tc ? nullptr : go, p.maxNp, goh, gol, GS, L.gb, L.ggamma, L.gbeta, pa))
#define MMAD_BN_BWD(R) do { if (fuse_dp) MMAD_BN_BWD2(R, true); else MMAD_BN_BWD2(R, false); } while (0)
                pa.stride = r.Np;
                if (B <= 64) MMAD_BN_BWD(1); else if (B <= 128) MMAD_BN_BWD(2); else if (B <= 256) MMAD_BN_BWD(4); else MMAD_BN_BWD(8);
#undef MMAD_BN_BWD2
#undef MMAD_BN_BWD
                MMAD_LAUNCHED();
            } else {
                bn_bwd_reduce_kernel<<<col_grid(r.N, B), kCT, 0, s>>>(gin.f, gin.ld, pre, r.Np, B, r.N, slope, gscale, mean, inv, stb, r.Np,
                                                                      L.gb, row_slab(B));
                MMAD_LAUNCHED();
                if (dist) {    // parameter gradients from the LOCAL sums, then the statistics are combined
                    bn_bwd_param_kernel<<<(r.N + 127) / 128, 128, 0, s>>>(stb, r.Np, r.N, L.ggamma, L.gbeta);
                    MMAD_LAUNCHED();
                    int rc = stats_allreduce(stb, 2LL * r.Np);
                    if (rc) return rc;
                }
                bn_bwd_apply_kernel<<<col_grid(r.N, B), kCT, 0, s>>>(gin.f, gin.ld, pre, r.Np, B, r.N, slope, gscale, mean, inv, L.gamma, stb,
                                                                     r.Np, Bg, tc ? nullptr : go, p.maxNp, goh, gol, GS, L.gb, L.ggamma,
                                                                     L.gbeta, dist ? 0 : 1, row_slab(B));
                MMAD_LAUNCHED();
            }
            gpre = Mat{go, goh, gol, p.maxNp};
            gi ^= 1;
            gemm_scale = 1.f;
        } else {
            if (!gw_zeroed) MMAD_CUDA_OK(cudaMemsetAsync(L.gb, 0, (size_t)r.N * 4, s));      // (else: inside the flat buffer zeroed at step start)
            MMAD_CUDA_OK(launch_k(col_sum_scaled_kernel, col_grid(r.N, B), dim3(kCT), 0, s, gin.f, gin.ld, B, r.N, gscale, L.gb, row_slab(B)));
            MMAD_LAUNCHED();
        }
        {   // gW[N,K] = g_pre^T in
            Epilogue e;
            e.Y = L.gW; e.ldy = r.K; e.y_cols = r.K;
            int rc;
            if (tc) {
                e.plain = 1; e.split_k_ok = 1; e.pre_zeroed = gw_zeroed ? 1 : 0;
                e.acc_scale = 1.f / GS;          // twins of g_pre carry GS (the 2 of dL/dxhat is inside the loss twins)
                // dW is off the critical path (nothing downstream reads it): second stream, joined at the end
                cudaStream_t sw = s;
                if (s2) {
                    MMAD_CUDA_OK(cudaEventRecord(ev_fork, s));
                    MMAD_CUDA_OK(cudaStreamWaitEvent(s2, ev_fork, 0));
                    sw = s2;
                    forked = true;
                }
                {
                    PdlScope side(sw == s ? g_pdl : false);      // second stream: ordinary (event) dependencies
                    rc = tc_gemm(gpre.h, gpre.l, gpre.ld, true, r.in.h, r.in.l, r.in.ld, true, r.N, r.K, B, e, sw);
                }
                // the module's last weight gradient is queued: its bucket follows on the same stream (the fork event above is
                // behind every bias / BatchNorm gradient of the module, which are computed on the main stream)
                if (!rc && grad_ar && r.i == 0) rc = module_grad_allreduce(r.m, sw);
            } else {
                GemmShape g;
                g.M = r.N; g.N = r.K; g.K = B;
                g.A = gpre.f; g.lda = gpre.ld; g.transA = true;
                g.B = r.in.f; g.ldb = r.in.ld; g.transB = true;
                e.acc_scale = gemm_scale;
                rc = gemm_simt(g, e, s);
                if (!rc && grad_ar && r.i == 0) rc = module_grad_allreduce(r.m, s);
            }
            if (rc) return rc;
        }
        if (idx > 0) {   // g_in[B,K] = g_pre W
            const bool into_vib = vib && r.m == 1 && r.i == 0;
            const Ref& prev = order[idx - 1];
            const bool use_gz = tc && p.has_gz;
            float* gout = use_gz ? (float*)(ws + p.gz[idx]) : (float*)(ws + p.g[gi ^ 1]);
            Epilogue e;
            e.Y = gout; e.ldy = p.maxNp;
            if (use_gz) e.pre_zeroed = 1;
            int rc;
            if (tc) {
                const LayerView lv = handle_layer(h, r.m, r.i);
                e.y_cols = np_of(r.K);
                e.acc_scale = 1.f / (GS * lv.wscale);
                const bool small = !(B >= 2048 && tc2_available());
                if (!prev.bn && !small) {   // the consumer is a bare Linear: it needs the twins of g_pre = g_in directly
                    e.Yh = (__half*)(ws + p.gth[idx - 1]); e.Yl = (__half*)(ws + p.gtl[idx - 1]); e.ldh = p.maxNp; e.y_split_scale = GS;
                } else {
                    e.plain = 1; e.split_k_ok = 1; e.y_cols = r.K;
                }
                rc = tc_gemm(gpre.h, gpre.l, gpre.ld, false, lv.Wh, lv.Wl, lv.Kp, true, B, r.K, r.N, e, s);
                if (!rc && !prev.bn && small) {   // small batch: plain split-K GEMM, twins by a follow-up kernel
                    const int Kq = np_of(r.K);
                    MMAD_CUDA_OK(launch_k(twin_split_kernel, dim3(ew_grid((size_t)B * Kq / 4)), dim3(256), 0, s, (const float*)gout, p.maxNp, B, r.K, Kq, GS,
                                          (__half*)(ws + p.gth[idx - 1]), (__half*)(ws + p.gtl[idx - 1]), p.maxNp));
                    MMAD_LAUNCHED();
                }
            } else {
                GemmShape g;
                g.M = B; g.N = r.K; g.K = r.N;
                g.A = gpre.f; g.lda = gpre.ld;
                g.B = L.W; g.ldb = r.K; g.transB = true;
                e.y_cols = r.K;
                e.acc_scale = gemm_scale;
                rc = gemm_simt(g, e, s);
            }
            if (rc) return rc;
            gi ^= 1;
            gcur = gout;
            if (into_vib) {
                vib_train_bwd_kernel<<<ew_grid((size_t)B * dec_in), 256, 0, s>>>(gout, p.maxNp, (const float*)(ws + p.pre[0][prev.i]), prev.Np,
                                                                                 B, dec_in, d_eps, beta_kl, (float*)(ws + p.genc), np_of(enc_out),
                                                                                 tc ? (__half*)(ws + p.gench) : nullptr,
                                                                                 tc ? (__half*)(ws + p.gencl) : nullptr, GS);
                MMAD_LAUNCHED();
            }
        }
        gscale = 1.f;
    }
    if (forked) {
        MMAD_CUDA_OK(cudaEventRecord(ev_join, s2));
        MMAD_CUDA_OK(cudaStreamWaitEvent(s, ev_join, 0));
    }
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int mmad_train_fwd_bwd(mmad_t h, const float* d_x, int ldx, int batch, long long global_batch,
                       const mmad_train_layer_t* enc, const mmad_train_layer_t* dec, const float* d_eps, float beta_kl,
                       float bn_momentum, float* d_loss, void* d_ws, size_t ws_bytes, mmad_allreduce_fn allreduce,
                       void* allreduce_ctx, void* stream) {
    mmad::NvtxScope nvtx_("mmad_train_fwd_bwd");
    if (!h || !d_x || !enc || !dec || !d_loss || !d_ws) { set_error("null argument"); return MMAD_E_ARG; }
    const mmad_desc_t& d = *handle_desc(h);
    const int D = d.enc_widths[0];
    if (batch < 1 || ldx < D) { set_error("bad batch/ldx"); return MMAD_E_ARG; }
    if (global_batch < batch) global_batch = batch;
    const int enc_out = d.enc_widths[d.n_enc], dec_in = d.dec_widths[0];
    const bool vib = d_eps != nullptr;
    if (vib ? (enc_out != 2 * dec_in) : (enc_out != dec_in)) {
        set_error("encoder output %d does not feed decoder input %d (%s)", enc_out, dec_in, vib ? "VIB expects 2x" : "pass eps for a VIB model");
        return MMAD_E_ARG;
    }
    // tensor-core GEMMs in the f16x3 / f16 modes, CUDA-core fp32 otherwise
    const bool tc = d.precision != MMAD_PREC_FP32 && tc_available();
    const TrainPlan p = make_train_plan(d, batch, tc_available() != 0);
    if (ws_bytes < p.total) { set_error("train workspace too small: %zu < %zu", ws_bytes, p.total); return MMAD_E_WORKSPACE; }
    for (int m = 0; m < 2; ++m) {
        const int n = m == 0 ? d.n_enc : d.n_dec;
        const mmad_train_layer_t* L = m == 0 ? enc : dec;
        for (int i = 0; i < n; ++i) {
            const bool bn = i < n - 1;
            if (!L[i].W || !L[i].b || !L[i].gW || !L[i].gb || (bn && (!L[i].gamma || !L[i].beta || !L[i].ggamma || !L[i].gbeta))) {
                set_error("layer %d.%d: missing parameter/gradient pointer", m, i);
                return MMAD_E_ARG;
            }
        }
    }
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = (char*)d_ws;
    // ---- stage the per-step inputs into the workspace (everything after this uses stable pointers) ----
    const int Dp = np_of(D);
    int rc = pad_split(d_x, ldx, batch, D, (float*)(ws + p.xp), Dp, tc ? (__half*)(ws + p.xh) : nullptr,
                       tc ? (__half*)(ws + p.xl) : nullptr, Dp, s);
    if (rc) return rc;
    if (vib) MMAD_CUDA_OK(cudaMemcpyAsync(ws + p.eps, d_eps, (size_t)batch * dec_in * 4, cudaMemcpyDeviceToDevice, s));

    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);
    const bool use_graph = !allreduce && graphs_enabled() && cap == cudaStreamCaptureStatusNone;
    // the loss doorbell (mapped pinned pair + device sequence counter) is allocated here, outside any capture
    uint2* bell_pair = nullptr; unsigned long long* bell_seq = nullptr;
    const bool bell = cap == cudaStreamCaptureStatusNone && !handle_loss_doorbell(h, &bell_pair, &bell_seq, true);
    {
        cudaStream_t s2 = nullptr; cudaEvent_t ef = nullptr, ej = nullptr;
        if (tc && handle_aux(h, &s2, &ef, &ej)) s2 = nullptr;
        if (!use_graph) {
            rc = train_body(h, d, p, tc, vib, batch, global_batch, enc, dec, beta_kl, bn_momentum, d_loss, ws, allreduce, allreduce_ctx, s, s2, ef, ej, bell);
            if (!rc && bell) handle_loss_published(h);
            return rc;
        }
    }

    // ---- CUDA graph of the step, keyed by everything the launch sequence depends on ----
    std::string key("train");
    auto add = [&](const void* q, size_t n) { key.append((const char*)q, n); };
    add(&batch, sizeof batch); add(&global_batch, sizeof global_batch); add(&d.precision, sizeof d.precision);
    add(enc, sizeof(mmad_train_layer_t) * d.n_enc); add(dec, sizeof(mmad_train_layer_t) * d.n_dec);
    add(&beta_kl, sizeof beta_kl); add(&bn_momentum, sizeof bn_momentum); add(&d_loss, sizeof d_loss); add(&d_ws, sizeof d_ws);
    add(&vib, sizeof vib);
    unsigned long long n_launch = 0;
    cudaGraphExec_t exec = handle_graph_find(h, key, &n_launch);
    if (!exec) {
        cudaStream_t cs = handle_capture_stream(h);
        if (!cs) return MMAD_E_CUDA;
        const unsigned long long l0 = g_launches;
        MMAD_CUDA_OK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
        cudaStream_t s2 = nullptr; cudaEvent_t ef = nullptr, ej = nullptr;
        if (tc && handle_aux(h, &s2, &ef, &ej)) s2 = nullptr;
        rc = train_body(h, d, p, tc, vib, batch, global_batch, enc, dec, beta_kl, bn_momentum, d_loss, ws, nullptr, nullptr, cs, s2, ef, ej, bell);
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        n_launch = g_launches - l0;
        g_launches = l0;
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess || !graph) { set_error("graph capture of the train step failed: %s", cudaGetErrorString(ce)); cudaGetLastError(); return MMAD_E_CUDA; }
        ce = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) { set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); return MMAD_E_CUDA; }
        handle_graph_put(h, key, exec, n_launch);
    }
    MMAD_CUDA_OK(cudaGraphLaunch(exec, s));
    g_launches += n_launch;
    if (bell) handle_loss_published(h);
    return MMAD_OK;
}

int mmad_train_loss(mmad_t h, float* h_loss) {
    if (!h || !h_loss) { set_error("null argument"); return MMAD_E_ARG; }
    return handle_loss_read(h, h_loss);
}

int mmad_adam_step(int n_tensors, float* const* h_params, float* const* h_grads, float* const* h_m, float* const* h_v,
                   const long long* h_numel, int step, float lr, float beta1, float beta2, float eps, float grad_scale,
                   void* stream) {
    mmad::NvtxScope nvtx_("mmad_adam_step");
    if (n_tensors < 0 || (n_tensors > 0 && (!h_params || !h_grads || !h_m || !h_v || !h_numel)) || step < 1) {
        set_error("bad argument"); return MMAD_E_ARG;
    }
    // scalars exactly as torch computes them (python doubles, then cast)
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float neg_step = (float)(-((double)lr / bc1));
    const float bc2_sqrt = (float)sqrt(bc2);
    for (int t0 = 0; t0 < n_tensors; t0 += kAdamMaxTensors) {
        AdamArgs a;
        a.n = std::min(kAdamMaxTensors, n_tensors - t0);
        int blocks = 0;
        for (int t = 0; t < a.n; ++t) {
            if (!h_params[t0 + t] || !h_grads[t0 + t] || !h_m[t0 + t] || !h_v[t0 + t] || h_numel[t0 + t] < 0) {
                set_error("adam: null tensor %d", t0 + t); return MMAD_E_ARG;
            }
            a.p[t] = h_params[t0 + t]; a.g[t] = h_grads[t0 + t]; a.m[t] = h_m[t0 + t]; a.v[t] = h_v[t0 + t];
            a.numel[t] = h_numel[t0 + t];
            a.block_start[t] = blocks;
            blocks += (int)((h_numel[t0 + t] + kAdamChunk - 1) / kAdamChunk);
        }
        a.block_start[a.n] = blocks;
        if (blocks == 0) continue;
        adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, 1.f - beta1, beta2, 1.f - beta2, bc2_sqrt, eps, neg_step, grad_scale);
        MMAD_LAUNCHED();
    }
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

}  // extern "C"
