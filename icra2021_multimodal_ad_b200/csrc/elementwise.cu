// Small HBM-bound helpers around the GEMM chain: operand packing, score finalisation,
// NAP statistics.  All are grid-stride / coalesced; none is on the critical path.
#include <cuda_fp8.h>

#include "mmad_internal.cuh"

namespace mmad {

namespace {

__device__ __forceinline__ void split_half(float v, __half& h, __half& l) {
    h = __float2half_rn(v);
    l = __float2half_rn(v - __half2float(h));
}

__device__ __forceinline__ uint8_t to_e4m3(float v) { return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3); }
// fp8 twin bytes of column c of a row (layout: Epilogue::lo_f8)
__device__ __forceinline__ void store_f8_twin1(uint8_t* row, int c, float v, float residual) {
    uint8_t* q = row + (c >> 2) * 8 + (c & 3);
    q[0] = to_e4m3(residual * kF8LoScale);
    q[4] = to_e4m3(v);
}

// x [n, D] (row stride ldx) -> zero-padded fp32 [n, ldp] and/or fp16 hi/lo [n, ldh]
__global__ void pad_split_kernel(const float* __restrict__ x, int ldx, int n, int D, float* __restrict__ xp, int ldp,
                                 __half* __restrict__ xh, __half* __restrict__ xl, int ldh, int lo_f8) {
    const int cols = xp ? ldp : ldh;
    const size_t total = (size_t)n * cols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(i / cols), c = (int)(i % cols);
        float v = c < D ? x[(size_t)r * ldx + c] : 0.f;
        if (xp && c < ldp) xp[(size_t)r * ldp + c] = v;
        if (xh && c < ldh) {
            __half h, l;
            split_half(v, h, l);
            xh[(size_t)r * ldh + c] = h;
            if (xl && lo_f8) store_f8_twin1(reinterpret_cast<uint8_t*>(xl) + (size_t)r * ldh * 2, c, v, v - __half2float(h));
            else if (xl) xl[(size_t)r * ldh + c] = l;
        }
    }
}

// same, four columns per thread (16-byte loads, 8-byte fp16 stores): needs ldx % 4 == 0, x 16-byte aligned,
// padded widths multiples of 4
__global__ void pad_split_vec4_kernel(const float* __restrict__ x, int ldx, int n, int D, float* __restrict__ xp, int ldp,
                                      __half* __restrict__ xh, __half* __restrict__ xl, int ldh, int lo_f8) {
    const int cols4 = (xp ? ldp : ldh) >> 2;
    const size_t total = (size_t)n * cols4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols4), c = (int)(i % cols4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c + 3 < D) {
            v = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * ldx + c));
        } else if (c < D) {
            const float* q = x + (size_t)r * ldx + c;
            v.x = q[0];
            if (c + 1 < D) v.y = q[1];
            if (c + 2 < D) v.z = q[2];
        }
        if (xp) *reinterpret_cast<float4*>(xp + (size_t)r * ldp + c) = v;
        if (xh) {
            const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
            uint2 hv;
            hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
            *reinterpret_cast<uint2*>(xh + (size_t)r * ldh + c) = hv;
            if (xl && lo_f8) {
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                uint8_t* q = reinterpret_cast<uint8_t*>(xl) + (size_t)r * ldh * 2 + 2 * c;
                const float ls = kF8LoScale;
                const uint32_t r01 = __nv_cvt_float2_to_fp8x2(make_float2((v.x - f01.x) * ls, (v.y - f01.y) * ls), __NV_SATFINITE, __NV_E4M3);
                const uint32_t r23 = __nv_cvt_float2_to_fp8x2(make_float2((v.z - f23.x) * ls, (v.w - f23.y) * ls), __NV_SATFINITE, __NV_E4M3);
                const uint32_t a01 = __nv_cvt_float2_to_fp8x2(make_float2(v.x, v.y), __NV_SATFINITE, __NV_E4M3);
                const uint32_t a23 = __nv_cvt_float2_to_fp8x2(make_float2(v.z, v.w), __NV_SATFINITE, __NV_E4M3);
                *reinterpret_cast<uint2*>(q) = make_uint2(r01 | (r23 << 16), a01 | (a23 << 16));
            } else if (xl) {
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
                uint2 lv;
                lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                *reinterpret_cast<uint2*>(xl + (size_t)r * ldh + c) = lv;
            }
        }
    }
}

// W [N, K] fp32 -> (W * scale) split into fp16 hi/lo [N, Kp], zero padded
__global__ void split_weights_kernel(const float* __restrict__ W, int N, int K, int Kp, float scale,
                                     __half* __restrict__ Wh, __half* __restrict__ Wl) {
    const size_t total = (size_t)N * Kp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(i / Kp), c = (int)(i % Kp);
        float v = c < K ? W[(size_t)r * K + c] * scale : 0.f;
        __half h, l;
        split_half(v, h, l);
        Wh[i] = h;
        Wl[i] = l;
    }
}

// MMAD_PREC_F16F8 weight twins ------------------------------------------------------------------
__global__ void absmax_kernel(const float* __restrict__ W, int N, int K, int ldw, unsigned int* __restrict__ out) {
    const size_t total = (size_t)N * K;
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float v = fabsf(W[(i / K) * ldw + (i % K)]);
        if (v < 3.0e38f) m = fmaxf(m, v);      // NaN / inf do not set the scale
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));   // non-negative floats order like their bits
}
// scale = the power of two that puts max|W| * scale in [2^13, 2^14): fp16 hi parts below 16384, residuals |Wl| <= 4,
// Wh / 2^11 <= 8 -- all inside e4m3's normal range for the weights that matter
__global__ void pick_pow2_scale_kernel(const unsigned int* amax, float* scale) {
    const float m = __uint_as_float(*amax);
    int ex = 0;
    if (m > 0.f) frexpf(m, &ex);             // m = f * 2^ex, f in [0.5, 1)
    *scale = m > 0.f ? exp2f((float)(14 - ex)) : 1.f;
}
__global__ void split_weights_f8_kernel(const float* __restrict__ W, int N, int K, int ldw, int Kp, const float* __restrict__ scale,
                                        __half* __restrict__ Wh, uint8_t* __restrict__ W8) {
    const size_t total = (size_t)N * Kp;
    const float sc = *scale;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / Kp), c = (int)(i % Kp);
        const float v = c < K ? W[(size_t)r * ldw + c] * sc : 0.f;
        const __half h = __float2half_rn(v);
        const float hf = __half2float(h);
        Wh[i] = h;
        uint8_t* q = W8 + (size_t)r * Kp * 2 + (c >> 2) * 8 + (c & 3);
        q[0] = to_e4m3(hf * (1.f / kF8LoScale));
        q[4] = to_e4m3(v - hf);
    }
}

// the same for up to 2 * MMAD_MAX_LAYERS matrices in ONE launch (train step: every layer's twins are refreshed)
struct SplitMulti {
    const float* W[2 * MMAD_MAX_LAYERS];
    __half* Wh[2 * MMAD_MAX_LAYERS];
    __half* Wl[2 * MMAD_MAX_LAYERS];
    int N[2 * MMAD_MAX_LAYERS], K[2 * MMAD_MAX_LAYERS], Kp[2 * MMAD_MAX_LAYERS];
    int block_start[2 * MMAD_MAX_LAYERS + 1];
    int n;
    float scale;
};
// thread <-> 4 consecutive columns of one row of the padded twin matrices (8-byte stores); the fp32 row need not be
// 16-byte aligned (K odd), so the loads are scalar -- consecutive lanes still read consecutive 16-byte groups.
// (One element per thread with a 64-bit div/mod each cost 34 us for the 5.1 M weights of the benchmark model: 6 % of a step.)
__global__ void __launch_bounds__(256) split_weights_multi_kernel(const __grid_constant__ SplitMulti a) {
    pdl_trigger();
    pdl_wait();
    int t = 0;
    while (t + 1 < a.n && (int)blockIdx.x >= a.block_start[t + 1]) ++t;
    const int Kp = a.Kp[t], K = a.K[t], q4 = Kp >> 2;
    const unsigned total4 = (unsigned)a.N[t] * (unsigned)q4;
    const unsigned base = (unsigned)(blockIdx.x - a.block_start[t]) * 1024u;       // 4096 elements per block
    const float* __restrict__ W = a.W[t];
    __half* __restrict__ Wh = a.Wh[t];
    __half* __restrict__ Wl = a.Wl[t];
    const float sc = a.scale;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const unsigned i = base + it * 256u + threadIdx.x;
        if (i >= total4) break;
        const unsigned r = i / (unsigned)q4, c = (i - r * (unsigned)q4) * 4u;
        const float* src = W + (size_t)r * K + c;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (int)(c + j) < K ? __ldg(src + j) * sc : 0.f;
        const __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v[0] - f01.x, v[1] - f01.y), l23 = __floats2half2_rn(v[2] - f23.x, v[3] - f23.y);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
        lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
        const size_t o = (size_t)r * Kp + c;
        *reinterpret_cast<uint2*>(Wh + o) = hv;
        *reinterpret_cast<uint2*>(Wl + o) = lv;
    }
}

__global__ void fold_bn_kernel(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                               int N, int Np, float* scale, float* shift) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Np) return;
    if (i < N) {
        // eval BatchNorm1d: (a - mean) / sqrt(var + eps) * gamma + beta  ==  a * s + t
        float s = gamma[i] / sqrtf(var[i] + eps);
        scale[i] = s;
        shift[i] = beta[i] - mean[i] * s;
    } else {
        scale[i] = 0.f;
        shift[i] = 0.f;
    }
}

__global__ void copy_pad_kernel(const float* src, int N, int Np, float* dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Np) dst[i] = i < N ? src[i] : 0.f;
}

// base[r] = inv_base * sum_{slots in [b_lo,b_hi)} part[slot][r];  sap likewise.
// Fixed summation order -> deterministic scores.
__global__ void finalize_scores_kernel(const float* __restrict__ part, int stride, int n, int b_lo, int b_hi,
                                       int s_lo, int s_hi, float inv_base, float inv_sap, float* __restrict__ base,
                                       float* __restrict__ sap) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (base) {
        float a = 0.f;
        for (int s = b_lo; s < b_hi; ++s) a += part[(size_t)s * stride + r];
        base[r] = a * inv_base;
    }
    if (sap) {
        float a = 0.f;
        for (int s = s_lo; s < s_hi; ++s) a += part[(size_t)s * stride + r];
        sap[r] = a * inv_sap;
    }
}

// same sums for a handful of rows and many slots (streaming calls: one 16-column slot per CTA of the skinny GEMM):
// one warp per row, lanes stride over the slots, fixed-order shuffle reduction -> deterministic
__global__ void finalize_scores_warp_kernel(const float* __restrict__ part, int stride, int n, int b_lo, int b_hi,
                                            int s_lo, int s_hi, float inv_base, float inv_sap, float* __restrict__ base,
                                            float* __restrict__ sap) {
    const int r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (r >= n) return;
    float a = 0.f, b = 0.f;
    if (base) for (int s = b_lo + lane; s < b_hi; s += 32) a += part[(size_t)s * stride + r];
    if (sap) for (int s = s_lo + lane; s < s_hi; s += 32) b += part[(size_t)s * stride + r];
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) {
        if (base) base[r] = a * inv_base;
        if (sap) sap[r] = b * inv_sap;
    }
}

__global__ void reduce_sum_all_kernel(const float* __restrict__ part, int stride, int n, int s_lo, int s_hi,
                                      float* __restrict__ acc) {
    // double accumulation inside the block, one atomic per block
    pdl_trigger();
    pdl_wait();
    double a = 0.0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
        for (int s = s_lo; s < s_hi; ++s) a += (double)part[(size_t)s * stride + r];
    __shared__ double sm[256];
    sm[threadIdx.x] = a;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(acc, (float)sm[0]);
}

// sum[c] += sum_r d[r, c]  (fp64 accumulation; grid: x over columns, y over row slabs)
__global__ void colsum_f64_kernel(const float* __restrict__ d, int ld, int n, int cols, SegMap sm,
                                  double* __restrict__ sum) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int pc = seg_pad_col(sm, c);
    int rows_per = (n + gridDim.y - 1) / gridDim.y;
    int r0 = blockIdx.y * rows_per, r1 = min(n, r0 + rows_per);
    double a = 0.0;
    for (int r = r0; r < r1; ++r) a += (double)d[(size_t)r * ld + pc];
    atomicAdd(&sum[c], a);
}

__global__ void gram_f64_accumulate_kernel(const float* __restrict__ g32, int ld32, int D, SegMap sm,
                                           double* __restrict__ g64) {
    const size_t total = (size_t)D * D;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(i / D), c = (int)(i % D);
        g64[i] += (double)g32[(size_t)seg_pad_col(sm, r) * ld32 + seg_pad_col(sm, c)];
    }
}

// NAP fit packing: B[j,:] = v_j (zero padded), colscale[j] = var_j^-1/2,
// bias[j] = -(mu . v_j + mu2_j) * var_j^-1/2   (utils/normalize.py:36-45,72-103 folded)
__global__ void nap_pack_kernel(const float* __restrict__ mu, const float* __restrict__ vt, const float* __restrict__ var,
                                const float* __restrict__ mu2, int K, int D, int Dp, SegMap seg,
                                float* __restrict__ B, float* __restrict__ colscale, float* __restrict__ bias,
                                float* __restrict__ bias_rot) {
    const int j = blockIdx.x;
    double dot = 0.0;
    for (int c = threadIdx.x; c < Dp; c += blockDim.x) B[(size_t)j * Dp + c] = 0.f;
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float v = vt[(size_t)j * D + c];
        B[(size_t)j * Dp + seg_pad_col(seg, c)] = v;
        dot += (double)v * (double)mu[c];
    }
    __shared__ double sm[256];
    sm[threadIdx.x] = dot;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float inv = 1.0f / sqrtf(var[j]);
        colscale[j] = inv;
        bias[j] = -((float)sm[0] + mu2[j]) * inv;
        bias_rot[j] = -(float)sm[0];          // Rotater.run alone: (d - mu) . v_j
    }
}

// Standardizer refit on rotated train rows (utils/normalize.py:25-34): colscale = var^-1/2,
// bias = (bias_rot - mu2) * var^-1/2
__global__ void nap_restandardize_kernel(const float* __restrict__ var, const float* __restrict__ mu2,
                                         const float* __restrict__ bias_rot, int K, float* __restrict__ colscale,
                                         float* __restrict__ bias) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= K) return;
    float inv = 1.0f / sqrtf(var[j]);
    colscale[j] = inv;
    bias[j] = (bias_rot[j] - mu2[j]) * inv;
}

// sum[c] += sum_r y[r,c];  sq[c] += sum_r y[r,c]^2   (fp64; grid: x over columns, y over row slabs)
__global__ void col_sum_sq_f64_kernel(const float* __restrict__ y, int ld, int n, int cols, double* __restrict__ sum,
                                      double* __restrict__ sq) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    int rows_per = (n + gridDim.y - 1) / gridDim.y;
    int r0 = blockIdx.y * rows_per, r1 = min(n, r0 + rows_per);
    double a = 0.0, b = 0.0;
    for (int r = r0; r < r1; ++r) { double v = (double)y[(size_t)r * ld + c]; a += v; b = fma(v, v, b); }
    atomicAdd(&sum[c], a);
    atomicAdd(&sq[c], b);
}

__global__ void center_rows_kernel(float* __restrict__ d, int ld, int n, int cols, SegMap sm,
                                   const float* __restrict__ mu) {
    const size_t total = (size_t)n * cols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(i / cols), c = (int)(i % cols);
        d[(size_t)r * ld + seg_pad_col(sm, c)] -= mu[c];
    }
}

// G64[i, j] += sum_r dc[r, pad(i)] * dc[r, pad(j)] with fp64 products and accumulation (exact products of
// fp32 inputs): the NAP fit needs eigenvalues ~1e-14 of the largest one, which an fp32 Gram
// cannot resolve.  64x64 output tile per CTA, 4x4 doubles per thread, row slabs of 16 staged in smem.
__global__ void __launch_bounds__(256)
gram_f64_direct_kernel(const float* __restrict__ dc, int ld, int rows, int D, SegMap sm, double* __restrict__ g64) {
    __shared__ float As[16][64 + 1];
    __shared__ float Bs[16][64 + 1];
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    if (j0 + 63 < i0) return;                      // lower triangle is mirrored below
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    __shared__ int pa[64], pb[64];
    if (threadIdx.x < 64) {
        pa[threadIdx.x] = (i0 + threadIdx.x < D) ? seg_pad_col(sm, i0 + threadIdx.x) : -1;
        pb[threadIdx.x] = (j0 + threadIdx.x < D) ? seg_pad_col(sm, j0 + threadIdx.x) : -1;
    }
    __syncthreads();
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int r0 = 0; r0 < rows; r0 += 16) {
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int rr = e / 64, c = e % 64;
            const int r = r0 + rr;
            As[rr][c] = (r < rows && pa[c] >= 0) ? dc[(size_t)r * ld + pa[c]] : 0.f;
            Bs[rr][c] = (r < rows && pb[c] >= 0) ? dc[(size_t)r * ld + pb[c]] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            double a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { a[q] = (double)As[rr][ty * 4 + q]; b[q] = (double)Bs[rr][tx * 4 + q]; }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int i = i0 + ty * 4 + x, j = j0 + tx * 4 + y;
            if (i < D && j < D && i <= j) {
                g64[(size_t)i * D + j] += acc[x][y];
                if (i != j) g64[(size_t)j * D + i] += acc[x][y];
            }
        }
}

inline int grid_for(size_t total, int block = 256) {
    size_t g = (total + block - 1) / block;
    return (int)(g > 148 * 16 ? 148 * 16 : (g == 0 ? 1 : g));
}

}  // namespace

int pad_split(const float* x, int ldx, int n, int D, float* xp, int ldp, __half* xh, __half* xl, int ldh,
              cudaStream_t s, int lo_f8) {
    if (n <= 0) return MMAD_OK;
    size_t total = (size_t)n * (xp ? ldp : ldh);
    const bool vec = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (ldp % 4 == 0) && (ldh % 4 == 0) &&
                     (!xp || !xh || ldp == ldh);
    if (vec) pad_split_vec4_kernel<<<grid_for(total / 4), 256, 0, s>>>(x, ldx, n, D, xp, ldp, xh, xl, ldh, lo_f8);
    else pad_split_kernel<<<grid_for(total), 256, 0, s>>>(x, ldx, n, D, xp, ldp, xh, xl, ldh, lo_f8);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int split_weights(const float* W, int N, int K, int Kp, float scale, __half* Wh, __half* Wl, cudaStream_t s) {
    split_weights_kernel<<<grid_for((size_t)N * Kp), 256, 0, s>>>(W, N, K, Kp, scale, Wh, Wl);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int split_weights_f8(const float* W, int N, int K, int ldw, int Kp, __half* Wh, uint8_t* W8, float* d_scale, cudaStream_t s) {
    // d_scale doubles as the atomicMax cell of the first pass (two floats: [scale, amax bits])
    unsigned int* amax = reinterpret_cast<unsigned int*>(d_scale) + 1;
    MMAD_CUDA_OK(cudaMemsetAsync(amax, 0, 4, s));
    absmax_kernel<<<grid_for((size_t)N * K), 256, 0, s>>>(W, N, K, ldw, amax);
    MMAD_LAUNCHED();
    pick_pow2_scale_kernel<<<1, 1, 0, s>>>(amax, d_scale);
    MMAD_LAUNCHED();
    split_weights_f8_kernel<<<grid_for((size_t)N * Kp), 256, 0, s>>>(W, N, K, ldw, Kp, d_scale, Wh, W8);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int split_weights_multi(int n, const float* const* W, const int* N, const int* K, const int* Kp, float scale, __half* const* Wh,
                        __half* const* Wl, cudaStream_t s) {
    if (n <= 0) return MMAD_OK;
    if (n > 2 * MMAD_MAX_LAYERS) { set_error("split_weights_multi: too many tensors"); return MMAD_E_ARG; }
    SplitMulti a;
    a.n = n; a.scale = scale;
    int blocks = 0;
    for (int t = 0; t < n; ++t) {
        a.W[t] = W[t]; a.Wh[t] = Wh[t]; a.Wl[t] = Wl[t]; a.N[t] = N[t]; a.K[t] = K[t]; a.Kp[t] = Kp[t];
        a.block_start[t] = blocks;
        blocks += (int)(((size_t)N[t] * Kp[t] + 4095) / 4096);
    }
    a.block_start[n] = blocks;
    MMAD_CUDA_OK(launch_k(split_weights_multi_kernel, dim3(blocks), dim3(256), 0, s, a));
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int N, int Np,
            float* scale, float* shift, cudaStream_t s) {
    fold_bn_kernel<<<(Np + 255) / 256, 256, 0, s>>>(gamma, beta, mean, var, eps, N, Np, scale, shift);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int copy_pad_vec(const float* src, int N, int Np, float* dst, cudaStream_t s) {
    copy_pad_kernel<<<(Np + 255) / 256, 256, 0, s>>>(src, N, Np, dst);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int finalize_scores(const float* rowpart, int stride, int n, int b_lo, int b_hi, int s_lo, int s_hi, float inv_base,
                    float inv_sap, float* base, float* sap, cudaStream_t s) {
    if (n <= 0) return MMAD_OK;
    if (n <= 64)
        finalize_scores_warp_kernel<<<(n + 7) / 8, 256, 0, s>>>(rowpart, stride, n, b_lo, b_hi, s_lo, s_hi, inv_base, inv_sap,
                                                                base, sap);
    else
        finalize_scores_kernel<<<(n + 255) / 256, 256, 0, s>>>(rowpart, stride, n, b_lo, b_hi, s_lo, s_hi, inv_base,
                                                               inv_sap, base, sap);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int finalize_sum(const float* rowpart, int stride, int n, int slot_lo, int slot_hi, float scale, float* out,
                 cudaStream_t s) {
    return finalize_scores(rowpart, stride, n, slot_lo, slot_hi, 0, 0, scale, 0.f, out, nullptr, s);
}

int reduce_sum_all(const float* rowpart, int stride, int n, int slot_lo, int slot_hi, float* acc, cudaStream_t s) {
    if (n <= 0) return MMAD_OK;
    int g = (n + 255) / 256;
    if (g > 64) g = 64;
    MMAD_CUDA_OK(launch_k(reduce_sum_all_kernel, dim3(g), dim3(256), 0, s, rowpart, stride, n, slot_lo, slot_hi, acc));
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int colsum_f64(const float* d, int ld, int n, int cols, const SegMap& sm, double* sum, cudaStream_t s) {
    if (n <= 0) return MMAD_OK;
    dim3 grid((cols + 127) / 128, n >= 4096 ? 32 : (n >= 256 ? 8 : 1));
    colsum_f64_kernel<<<grid, 128, 0, s>>>(d, ld, n, cols, sm, sum);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int nap_pack(const float* mu, const float* vt, const float* var, const float* mu2, int K, int D, int Dp,
             const SegMap& sm, float* B, float* colscale, float* bias, float* bias_rot, cudaStream_t s) {
    nap_pack_kernel<<<K, 256, 0, s>>>(mu, vt, var, mu2, K, D, Dp, sm, B, colscale, bias, bias_rot);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int nap_restandardize(const float* var, const float* mu2, const float* bias_rot, int K, float* colscale, float* bias,
                      cudaStream_t s) {
    nap_restandardize_kernel<<<(K + 255) / 256, 256, 0, s>>>(var, mu2, bias_rot, K, colscale, bias);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int col_sum_sq_f64(const float* y, int ld, int n, int cols, double* sum, double* sq, cudaStream_t s) {
    if (n <= 0) return MMAD_OK;
    dim3 grid((cols + 127) / 128, n >= 4096 ? 32 : (n >= 256 ? 8 : 1));
    col_sum_sq_f64_kernel<<<grid, 128, 0, s>>>(y, ld, n, cols, sum, sq);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int gram_f64_accumulate(const float* g32, int ld32, int D, const SegMap& sm, double* g64, cudaStream_t s) {
    size_t total = (size_t)D * D;
    gram_f64_accumulate_kernel<<<grid_for(total), 256, 0, s>>>(g32, ld32, D, sm, g64);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int gram_f64_direct(const float* dc, int ld, int rows, int D, const SegMap& sm, double* g64, cudaStream_t s) {
    if (rows <= 0) return MMAD_OK;
    dim3 grid((D + 63) / 64, (D + 63) / 64);
    gram_f64_direct_kernel<<<grid, 256, 0, s>>>(dc, ld, rows, D, sm, g64);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int center_rows(float* d, int ld, int n, int cols, const SegMap& sm, const float* mu, cudaStream_t s) {
    if (n <= 0) return MMAD_OK;
    center_rows_kernel<<<grid_for((size_t)n * cols), 256, 0, s>>>(d, ld, n, cols, sm, mu);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

}  // namespace mmad
