// Stand-alone HBM-bound ops of the path exposed through the C ABI: summed squared error,
// per-row mean of squares, VIB reparameterisation.
#include "mmad_internal.cuh"

using namespace mmad;

namespace {

__global__ void sq_diff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                   float* __restrict__ out) {
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float d = a[i] - b[i];
        acc += (double)(d * d);
    }
    __shared__ double sm[256];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(out, (float)sm[0]);
}

// one warp per row, coalesced over columns
__global__ void row_mean_sq_kernel(const float* __restrict__ d, int ld, int n, int cols, float* __restrict__ out) {
    int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x % 32;
    if (row >= n) return;
    const float* p = d + (size_t)row * ld;
    float acc = 0.f;
    for (int c = lane; c < cols; c += 32) acc = fmaf(p[c], p[c], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[row] = acc / (float)cols;
}

__global__ void vib_kernel(const float* __restrict__ o, int ld, int B, int h, int k, const float* __restrict__ eps,
                           float* __restrict__ z, float* __restrict__ mu, float* __restrict__ logvar) {
    const long long total = (long long)B * h;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int b = (int)(i / h), j = (int)(i % h);
        float m = o[(size_t)b * ld + j];
        float lv = o[(size_t)b * ld + h + j];
        mu[i] = m;
        logvar[i] = lv;
        float sigma = expf(lv * 0.5f);
        for (int kk = 0; kk < k; ++kk) {
            size_t zi = (size_t)kk * total + i;
            z[zi] = eps ? fmaf(eps[zi], sigma, m) : m;
        }
    }
}

// buf *= *scale unless *scale == 1 (the autograd seed of a plain loss.backward()): every thread reads the scalar
// first, so the common case is a launch that touches 4 bytes
__global__ void scale_unless_one_kernel(float* __restrict__ buf, long long n4, const float* __restrict__ scale) {
    const float s = __ldg(scale);
    if (s == 1.f) return;
    float4* b4 = reinterpret_cast<float4*>(buf);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = b4[i];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        b4[i] = v;
    }
}

}  // namespace

extern "C" {

int mmad_scale_unless_one(float* d_buf, long long n, const float* d_scale, void* stream) {
    if (!d_buf || !d_scale || n < 0 || (n & 3) || (reinterpret_cast<uintptr_t>(d_buf) & 15)) {
        set_error("mmad_scale_unless_one: buffer must be 16-byte aligned with a multiple of 4 elements");
        return MMAD_E_ARG;
    }
    if (n == 0) return MMAD_OK;
    scale_unless_one_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(d_buf, n / 4, d_scale);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int mmad_sq_diff_sum(const float* d_a, const float* d_b, long long n, float* d_out, void* stream) {
    if (!d_a || !d_b || !d_out || n < 0) { set_error("bad argument"); return MMAD_E_ARG; }
    if (n == 0) return MMAD_OK;
    long long g = (n + 255) / 256;
    if (g > 148 * 4) g = 148 * 4;
    sq_diff_sum_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(d_a, d_b, n, d_out);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int mmad_row_mean_sq(const float* d_d, int ld, int n, int cols, float* d_out, void* stream) {
    if (!d_d || !d_out || n < 0 || cols < 1 || ld < cols) { set_error("bad argument"); return MMAD_E_ARG; }
    if (n == 0) return MMAD_OK;
    row_mean_sq_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(d_d, ld, n, cols, d_out);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int mmad_vib_reparam(const float* d_out, int ld, int B, int h, int k, const float* d_eps, float* d_z, float* d_mu,
                     float* d_logvar, void* stream) {
    if (!d_out || !d_z || !d_mu || !d_logvar || B < 0 || h < 1 || ld < 2 * h) { set_error("bad argument"); return MMAD_E_ARG; }
    if (k < 1) { set_error("k should be >= 1"); return MMAD_E_ARG; }
    if (B == 0) return MMAD_OK;
    long long total = (long long)B * h;
    long long g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    vib_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(d_out, ld, B, h, k, d_eps, d_z, d_mu, d_logvar);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

}  // extern "C"
