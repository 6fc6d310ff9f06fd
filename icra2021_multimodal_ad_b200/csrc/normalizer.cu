// Stand-alone Rotater / Standardizer device ops (utils/normalize.py) for the API-compatible NAP path
// get_d_norm_loss(train_diffs, valid_diffs, test_diffs, ...).  The fused scorer (mmad_score) does not
// use these; they serve callers that hand over materialised diff matrices like the reference does.
#include "mmad_internal.cuh"

using namespace mmad;

namespace {

constexpr int kRows = 8192;   // rows per chunk of the centred scratch copy

__global__ void col_sq_dev_kernel(const float* __restrict__ d, int ld, long long n, int cols, const double* __restrict__ mean,
                                  double* __restrict__ acc) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    long long rows_per = (n + gridDim.y - 1) / gridDim.y;
    long long r0 = blockIdx.y * rows_per, r1 = min(n, r0 + rows_per);
    const double m = mean[c];
    double a = 0.0;
    for (long long r = r0; r < r1; ++r) { double v = (double)d[(size_t)r * ld + c] - m; a += v * v; }
    atomicAdd(&acc[c], a);
}

__global__ void col_sum_kernel(const float* __restrict__ d, int ld, long long n, int cols, double* __restrict__ acc) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    long long rows_per = (n + gridDim.y - 1) / gridDim.y;
    long long r0 = blockIdx.y * rows_per, r1 = min(n, r0 + rows_per);
    double a = 0.0;
    for (long long r = r0; r < r1; ++r) a += (double)d[(size_t)r * ld + c];
    atomicAdd(&acc[c], a);
}

__global__ void finish_stats_kernel(double* sum, double* sq, long long n, int cols, float* mean, float* var, int phase) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    if (phase == 0) {
        // the reference subtracts the fp32 mean and np.cov re-centres in fp64: use the fp32-rounded
        // mean for the output and the exact fp64 mean of the data for the variance
        sum[c] = sum[c] / (double)n;
        mean[c] = (float)sum[c];
    } else if (var) {
        var[c] = (float)(sq[c] / (double)(n - 1));
    }
}

// out[r, c] = (d[r, c] - mu[c]) * (var ? rsqrt-free 1/sqrt(var[c]) : 1), zero padded to ldo columns when pad
__global__ void center_copy_kernel(const float* __restrict__ d, int ld, int rows, int cols, const float* __restrict__ mu,
                                   const float* __restrict__ var, float* __restrict__ out, int ldo, int out_cols) {
    const size_t total = (size_t)rows * out_cols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(i / out_cols), c = (int)(i % out_cols);
        float v = 0.f;
        if (c < cols) {
            v = d[(size_t)r * ld + c] - mu[c];
            if (var) v = v / sqrtf(var[c]);
        }
        out[(size_t)r * ldo + c] = v;
    }
}

inline int grid_for(size_t total) {
    size_t g = (total + 255) / 256;
    return (int)(g > 148 * 16 ? 148 * 16 : (g == 0 ? 1 : g));
}

size_t ws_need(int cols) {
    const size_t cp = round_up(cols, kPad);
    return round_up_sz((size_t)kRows * cp * 4, 256) + round_up_sz(cp * cp * 4, 256) + 4 * round_up_sz((size_t)cols * 8, 256) + 1024;
}

// upper triangle (row-major, row i holds columns i..D-1) <-> full symmetric matrix: the cross-rank exchange of the NAP Gram
// matrix moves D(D+1)/2 doubles instead of D^2
__global__ void tri_pack_kernel(const double* __restrict__ full, int D, double* __restrict__ packed) {
    const int i = blockIdx.x;
    const size_t off = (size_t)i * D - (size_t)i * (i - 1) / 2 - i;      // packed index of (i, j) = off + j
    for (int j = i + threadIdx.x; j < D; j += blockDim.x) packed[off + j] = full[(size_t)i * D + j];
}
__global__ void tri_unpack_kernel(const double* __restrict__ packed, int D, double* __restrict__ full) {
    const int i = blockIdx.x;
    const size_t off = (size_t)i * D - (size_t)i * (i - 1) / 2 - i;
    for (int j = i + threadIdx.x; j < D; j += blockDim.x) {
        const double v = packed[off + j];
        full[(size_t)i * D + j] = v;
        full[(size_t)j * D + i] = v;
    }
}

}  // namespace

extern "C" {

size_t mmad_normalizer_workspace_bytes(int cols) { return cols < 1 ? 0 : ws_need(cols); }

int mmad_col_stats(const float* d_d, int ld, long long n, int cols, float* d_mean, float* d_var, void* d_ws,
                   size_t ws_bytes, void* stream) {
    if (!d_d || !d_mean || n < 1 || cols < 1 || ld < cols) { set_error("bad argument"); return MMAD_E_ARG; }
    if (!d_ws || ws_bytes < ws_need(cols)) { set_error("normalizer workspace too small"); return MMAD_E_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    double* sum = (double*)d_ws;
    double* sq = sum + round_up(cols, 32);
    MMAD_CUDA_OK(cudaMemsetAsync(sum, 0, (size_t)2 * round_up(cols, 32) * 8, s));
    dim3 grid((cols + 127) / 128, n >= 4096 ? 32 : (n >= 256 ? 8 : 1));
    col_sum_kernel<<<grid, 128, 0, s>>>(d_d, ld, n, cols, sum);
    MMAD_LAUNCHED();
    finish_stats_kernel<<<(cols + 127) / 128, 128, 0, s>>>(sum, sq, n, cols, d_mean, d_var, 0);
    MMAD_LAUNCHED();
    if (d_var) {
        col_sq_dev_kernel<<<grid, 128, 0, s>>>(d_d, ld, n, cols, sum, sq);
        MMAD_LAUNCHED();
        finish_stats_kernel<<<(cols + 127) / 128, 128, 0, s>>>(sum, sq, n, cols, d_mean, d_var, 1);
        MMAD_LAUNCHED();
    }
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int mmad_gram_accumulate(const float* d_d, int ld, long long n, int cols, const float* d_mu, double* d_gram, void* d_ws,
                         size_t ws_bytes, void* stream) {
    if (!d_d || !d_mu || !d_gram || n < 0 || cols < 1 || ld < cols) { set_error("bad argument"); return MMAD_E_ARG; }
    if (!d_ws || ws_bytes < ws_need(cols)) { set_error("normalizer workspace too small"); return MMAD_E_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    const int cp = round_up(cols, kPad);
    float* cen = (float*)d_ws;
    float* g32 = (float*)((char*)d_ws + round_up_sz((size_t)kRows * cp * 4, 256));
    SegMap ident;
    ident.n = 1; ident.tight_off[1] = cols; ident.pad_off[1] = cp;
    for (long long r0 = 0; r0 < n; r0 += kRows) {
        const int rows = (int)std::min<long long>(kRows, n - r0);
        center_copy_kernel<<<grid_for((size_t)rows * cp), 256, 0, s>>>(d_d + (size_t)r0 * ld, ld, rows, cols, d_mu, nullptr, cen, cp, cp);
        MMAD_LAUNCHED();
        int rc = gram_f64_direct(cen, cp, rows, cols, ident, d_gram, s);
        if (rc) return rc;
    }
    return MMAD_OK;
}

int mmad_rotate(const float* d_d, int ld, long long n, int cols, const float* d_mu, const float* d_vt, int K,
                float* d_out, int ldo, void* d_ws, size_t ws_bytes, void* stream) {
    if (!d_d || !d_mu || !d_vt || !d_out || n < 0 || cols < 1 || K < 1 || ld < cols || ldo < K) { set_error("bad argument"); return MMAD_E_ARG; }
    if (!d_ws || ws_bytes < ws_need(cols)) { set_error("normalizer workspace too small"); return MMAD_E_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    const int cp = round_up(cols, kPad);
    float* cen = (float*)d_ws;
    for (long long r0 = 0; r0 < n; r0 += kRows) {
        const int rows = (int)std::min<long long>(kRows, n - r0);
        center_copy_kernel<<<grid_for((size_t)rows * cp), 256, 0, s>>>(d_d + (size_t)r0 * ld, ld, rows, cols, d_mu, nullptr, cen, cp, cp);
        MMAD_LAUNCHED();
        GemmShape g;
        g.M = rows; g.N = K; g.K = cols;
        g.A = cen; g.lda = cp;
        g.B = d_vt; g.ldb = cols;
        Epilogue e;
        e.Y = d_out + (size_t)r0 * ldo; e.ldy = ldo; e.y_cols = K;
        int rc = gemm_simt(g, e, s);
        if (rc) return rc;
    }
    return MMAD_OK;
}

int mmad_standardize(const float* d_d, int ld, long long n, int cols, const float* d_mu, const float* d_var, float* d_out,
                     int ldo, void* stream) {
    if (!d_d || !d_mu || !d_var || !d_out || n < 0 || cols < 1 || ld < cols || ldo < cols) { set_error("bad argument"); return MMAD_E_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    for (long long r0 = 0; r0 < n; r0 += (1 << 20)) {
        const int rows = (int)std::min<long long>(1 << 20, n - r0);
        center_copy_kernel<<<grid_for((size_t)rows * cols), 256, 0, s>>>(d_d + (size_t)r0 * ld, ld, rows, cols, d_mu, d_var,
                                                                         d_out + (size_t)r0 * ldo, ldo, cols);
        MMAD_LAUNCHED();
    }
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int mmad_tri_pack(const double* d_full, int D, double* d_packed, void* stream) {
    if (!d_full || !d_packed || D < 1) { set_error("bad argument"); return MMAD_E_ARG; }
    tri_pack_kernel<<<D, 256, 0, (cudaStream_t)stream>>>(d_full, D, d_packed);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int mmad_tri_unpack(const double* d_packed, int D, double* d_full, void* stream) {
    if (!d_full || !d_packed || D < 1) { set_error("bad argument"); return MMAD_E_ARG; }
    tri_unpack_kernel<<<D, 256, 0, (cudaStream_t)stream>>>(d_packed, D, d_full);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

}  // extern "C"
