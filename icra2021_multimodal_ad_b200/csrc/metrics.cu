// Device AUROC / AUPRC / quantile / confusion counts, bit-identical to the reference's
// scikit-learn + NumPy path (utils/metric.py:29-130) given identical fp32 scores.
//
// sklearn recipe restated (sklearn/metrics/_ranking.py 878-1047, 1160-1205, 1317-1372, 95-115):
//   sort scores descending; thresholds = last index of every run of equal scores;
//   tps = cumsum(label)[idx], fps = 1 + idx - tps (exact integers held in fp64);
//   ROC: drop points whose second difference of fps and tps is zero (drop_intermediate),
//        prepend (0,0), divide by the totals;  PR: precision = tps/(tps+fps), recall = tps/tps[-1],
//        reversed, (1,0) appended;
//   area = sum_i (x[i+1]-x[i]) * (y[i+1]+y[i]) / 2.0 summed in NumPy's pairwise order
//        (blocks of <=128 with 8 interleaved partial sums, split n/2 rounded down to 8),
//   every operation a separately rounded fp64 op (no FMA contraction).
// Sorting/scans/compaction use CUB (part of the CUDA toolkit); the arithmetic kernels are ours.
#include <cub/cub.cuh>

#include "mmad_internal.cuh"

using namespace mmad;

namespace {

struct Bump {
    char* base; size_t off = 0;
    explicit Bump(void* b) : base((char*)b) {}
    template <class T> T* take(size_t n) {
        size_t at = off;
        off += round_up_sz(n * sizeof(T) + 16, 256);
        return base ? reinterpret_cast<T*>(base + at) : nullptr;
    }
};

struct Bufs {
    float* keys_in; float* keys; uint8_t* labs; int* tps_all; uint8_t* flags; int* idx; int* tps;
    uint8_t* keep; int* idx2; int* tps2; double* terms; int* leaf_off; int* leaf_len; double* leaf_sum;
    int* counters;        // [0] n_thresholds, [1] n_kept, [2] nonfinite flag, [3] n_leaves
    double* result;       // [0] area
    long long* counts;    // confusion
    void* cub_tmp; size_t cub_bytes;
    size_t total;
};

size_t cub_temp_bytes(long long n) {
    size_t a = 0, b = 0, c = 0, d = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, a, (const float*)nullptr, (float*)nullptr, (const uint8_t*)nullptr,
                                              (uint8_t*)nullptr, (int)n);
    cub::DeviceScan::InclusiveSum(nullptr, b, (const uint8_t*)nullptr, (int*)nullptr, (int)n);
    cub::DeviceSelect::Flagged(nullptr, c, (const int*)nullptr, (const uint8_t*)nullptr, (int*)nullptr, (int*)nullptr, (int)n);
    cub::DeviceRadixSort::SortKeys(nullptr, d, (const float*)nullptr, (float*)nullptr, (int)n);
    return std::max(std::max(a, b), std::max(c, d));
}

Bufs layout(void* ws, long long n) {
    Bump b(ws);
    Bufs o;
    const size_t N = (size_t)std::max<long long>(n, 1);
    o.keys_in = b.take<float>(N);
    o.keys = b.take<float>(N);
    o.labs = b.take<uint8_t>(N);
    o.tps_all = b.take<int>(N);
    o.flags = b.take<uint8_t>(N);
    o.idx = b.take<int>(N);
    o.tps = b.take<int>(N);
    o.keep = b.take<uint8_t>(N);
    o.idx2 = b.take<int>(N);
    o.tps2 = b.take<int>(N);
    o.terms = b.take<double>(N + 2);
    o.leaf_off = b.take<int>(N / 32 + 64);
    o.leaf_len = b.take<int>(N / 32 + 64);
    o.leaf_sum = b.take<double>(N / 32 + 64);
    o.counters = b.take<int>(8);
    o.result = b.take<double>(2);
    o.counts = b.take<long long>(4);
    o.cub_bytes = cub_temp_bytes((long long)N);
    o.cub_tmp = b.take<char>(o.cub_bytes);
    o.total = b.off;
    return o;
}

__global__ void nonfinite_kernel(const float* __restrict__ s, long long n, int* flag) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (!isfinite(s[i])) *flag = 1;
}

__global__ void flag_kernel(const float* __restrict__ s, int n, uint8_t* __restrict__ flags) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        flags[i] = (i == n - 1) || (s[i] != s[i + 1]);
}

// ROC drop_intermediate: keep the ends and every point where the second difference of fps or tps
// is non-zero (np.diff(.,2) on exact integers).
__global__ void keep_kernel(const int* __restrict__ idx, const int* __restrict__ tps, int m, uint8_t* __restrict__ keep) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        bool k = true;
        if (m > 2 && j > 0 && j < m - 1) {
            const long long f0 = 1LL + idx[j - 1] - tps[j - 1], f1 = 1LL + idx[j] - tps[j], f2 = 1LL + idx[j + 1] - tps[j + 1];
            const long long t0 = tps[j - 1], t1 = tps[j], t2 = tps[j + 1];
            k = ((f2 - f1) - (f1 - f0)) != 0 || ((t2 - t1) - (t1 - t0)) != 0;
        }
        keep[j] = k;
    }
}

__device__ __forceinline__ double trap_term(double x0, double x1, double y0, double y1) {
    // (x1 - x0) * (y1 + y0) / 2.0 with individually rounded operations
    return __ddiv_rn(__dmul_rn(__dsub_rn(x1, x0), __dadd_rn(y1, y0)), 2.0);
}

// ROC terms over the kept points with (0,0) prepended.  m = number of kept points.
__global__ void roc_terms_kernel(const int* __restrict__ idx, const int* __restrict__ tps, const int* m_ptr,
                                 double* __restrict__ terms) {
    const int m = *m_ptr;
    const double ftot = (double)(1LL + idx[m - 1] - tps[m - 1]);
    const double ttot = (double)tps[m - 1];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        // point i-1 (or the prepended origin) -> point i
        const double f1 = (double)(1LL + idx[i] - tps[i]), t1 = (double)tps[i];
        const double f0 = i ? (double)(1LL + idx[i - 1] - tps[i - 1]) : 0.0, t0 = i ? (double)tps[i - 1] : 0.0;
        const double x0 = ftot <= 0 ? nan : __ddiv_rn(f0, ftot), x1 = ftot <= 0 ? nan : __ddiv_rn(f1, ftot);
        const double y0 = ttot <= 0 ? nan : __ddiv_rn(t0, ttot), y1 = ttot <= 0 ? nan : __ddiv_rn(t1, ttot);
        terms[i] = trap_term(x0, x1, y0, y1);
    }
}

// PR terms: x = recall reversed then 0, y = precision reversed then 1.  m thresholds -> m terms.
__global__ void pr_terms_kernel(const int* __restrict__ idx, const int* __restrict__ tps, const int* m_ptr,
                                double* __restrict__ terms, int* any_decrease) {
    const int m = *m_ptr;
    const double ttot = (double)tps[m - 1];
    auto prec = [&](int j) {
        const double t = (double)tps[j], f = (double)(1LL + idx[j] - tps[j]);
        const double ps = __dadd_rn(t, f);
        return ps != 0.0 ? __ddiv_rn(t, ps) : 0.0;
    };
    auto rec = [&](int j) { return ttot == 0.0 ? 1.0 : __ddiv_rn((double)tps[j], ttot); };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const int j0 = m - 1 - i;               // reversed index of point i
        const double x0 = rec(j0), y0 = prec(j0);
        const double x1 = (i + 1 < m) ? rec(j0 - 1) : 0.0;
        const double y1 = (i + 1 < m) ? prec(j0 - 1) : 1.0;
        if (__dsub_rn(x1, x0) < 0.0) *any_decrease = 1;
        terms[i] = trap_term(x0, x1, y0, y1);
    }
}

// ---- NumPy pairwise summation (numpy/_core/src/umath/loops_utils.h.src) ----
__global__ void pw_leaves_kernel(const int* n_ptr, int* leaf_off, int* leaf_len, int* n_leaves) {
    int so[64], sl[64], sp = 0, cnt = 0;
    so[0] = 0; sl[0] = *n_ptr; sp = 1;
    while (sp) {
        --sp;
        const int o = so[sp], l = sl[sp];
        if (l <= 128) { leaf_off[cnt] = o; leaf_len[cnt] = l; ++cnt; continue; }
        int n2 = l / 2; n2 -= n2 % 8;
        so[sp] = o + n2; sl[sp] = l - n2; ++sp;     // right, processed after
        so[sp] = o; sl[sp] = n2; ++sp;              // left first
    }
    *n_leaves = cnt;
}

__global__ void pw_leaf_sum_kernel(const double* __restrict__ a, const int* __restrict__ leaf_off,
                                   const int* __restrict__ leaf_len, const int* n_leaves, double* __restrict__ leaf_sum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_leaves) return;
    const double* p = a + leaf_off[i];
    const int n = leaf_len[i];
    double res;
    if (n < 8) {
        res = 0.0;
        for (int k = 0; k < n; ++k) res = __dadd_rn(res, p[k]);
    } else {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = p[j];
        const int m = n - (n % 8);
        for (int k = 8; k < m; k += 8)
            for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], p[k + j]);
        res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                        __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (int k = m; k < n; ++k) res = __dadd_rn(res, p[k]);
    }
    leaf_sum[i] = res;
}

// post-order combine of the leaf sums in the same tree; result scaled by `direction`
__global__ void pw_combine_kernel(const int* n_ptr, const double* __restrict__ leaf_sum, const int* any_decrease,
                                  double* result) {
    // frames: (len, state, left value)
    int fl[64]; int fs[64]; double fv[64];
    int sp = 0, next_leaf = 0;
    double ret = 0.0;
    fl[0] = *n_ptr; fs[0] = 0; sp = 1;
    while (sp) {
        const int t = sp - 1;
        if (fl[t] <= 128) { ret = leaf_sum[next_leaf++]; --sp; continue; }
        int n2 = fl[t] / 2; n2 -= n2 % 8;
        if (fs[t] == 0) { fs[t] = 1; fl[sp] = n2; fs[sp] = 0; ++sp; }
        else if (fs[t] == 1) { fv[t] = ret; fs[t] = 2; fl[sp] = fl[t] - n2; fs[sp] = 0; ++sp; }
        else { ret = __dadd_rn(fv[t], ret); --sp; }
    }
    const double direction = (any_decrease && *any_decrease) ? -1.0 : 1.0;
    result[0] = __dmul_rn(direction, ret);
}

// np.quantile(fp32 array, q) with NumPy-2 fp32 arithmetic, 'linear' method
__global__ void quantile_kernel(const float* __restrict__ sorted, int n, float q, float* out) {
    if (isnan(sorted[n - 1])) { *out = sorted[n - 1]; return; }
    const float vi = __fmul_rn((float)(n - 1), q);
    const float lof = floorf(vi);
    int lo = (int)lof;
    int hi = lo + 1 < n ? lo + 1 : n - 1;
    const float g = __fsub_rn(vi, lof);
    const float a = sorted[lo], b = sorted[hi];
    const float diff = __fsub_rn(b, a);
    *out = g >= 0.5f ? __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, g))) : __fadd_rn(a, __fmul_rn(diff, g));
}

__global__ void confusion_kernel(const float* __restrict__ s, const uint8_t* __restrict__ y, long long n, float thr,
                                 int strict, unsigned long long* counts) {
    unsigned long long tp = 0, fp = 0, fn = 0, tn = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const bool pred = strict ? (s[i] > thr) : (s[i] >= thr);
        const bool lab = y[i] != 0;
        tp += pred && lab; fp += pred && !lab; fn += !pred && lab; tn += !pred && !lab;
    }
    for (int o = 16; o > 0; o >>= 1) {
        tp += __shfl_xor_sync(0xffffffffu, tp, o); fp += __shfl_xor_sync(0xffffffffu, fp, o);
        fn += __shfl_xor_sync(0xffffffffu, fn, o); tn += __shfl_xor_sync(0xffffffffu, tn, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&counts[0], tp); atomicAdd(&counts[1], fp); atomicAdd(&counts[2], fn); atomicAdd(&counts[3], tn);
    }
}

inline int grid_for(long long n) {
    long long g = (n + 255) / 256;
    return (int)std::min<long long>(std::max<long long>(g, 1), 148 * 8);
}

// common front end: finite check, sort, thresholds.  Returns 1 if a non-finite score was found.
int curve_front(const float* d_score, const uint8_t* d_label, long long n, Bufs& b, cudaStream_t s, int* nonfinite) {
    MMAD_CUDA_OK(cudaMemsetAsync(b.counters, 0, 8 * sizeof(int), s));
    nonfinite_kernel<<<grid_for(n), 256, 0, s>>>(d_score, n, b.counters + 2);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    size_t tmp = b.cub_bytes;
    MMAD_CUDA_OK(cub::DeviceRadixSort::SortPairsDescending(b.cub_tmp, tmp, d_score, b.keys, d_label, b.labs, (int)n, 0, 32, s));
    flag_kernel<<<grid_for(n), 256, 0, s>>>(b.keys, (int)n, b.flags);
    MMAD_LAUNCHED();
    tmp = b.cub_bytes;
    MMAD_CUDA_OK(cub::DeviceScan::InclusiveSum(b.cub_tmp, tmp, b.labs, b.tps_all, (int)n, s));
    tmp = b.cub_bytes;
    MMAD_CUDA_OK(cub::DeviceSelect::Flagged(b.cub_tmp, tmp, cub::CountingInputIterator<int>(0), b.flags, b.idx, b.counters + 0, (int)n, s));
    tmp = b.cub_bytes;
    MMAD_CUDA_OK(cub::DeviceSelect::Flagged(b.cub_tmp, tmp, b.tps_all, b.flags, b.tps, b.counters + 0, (int)n, s));
    g_launches += 4;
    int h = 0;
    MMAD_CUDA_OK(cudaMemcpyAsync(&h, b.counters + 2, sizeof(int), cudaMemcpyDeviceToHost, s));
    MMAD_CUDA_OK(cudaStreamSynchronize(s));
    *nonfinite = h;
    return MMAD_OK;
}

int finish_area(Bufs& b, const int* m_ptr, const int* any_decrease, double* h_out, cudaStream_t s) {
    pw_leaves_kernel<<<1, 1, 0, s>>>(m_ptr, b.leaf_off, b.leaf_len, b.counters + 3);
    MMAD_LAUNCHED();
    int m = 0;
    MMAD_CUDA_OK(cudaMemcpyAsync(&m, m_ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
    MMAD_CUDA_OK(cudaStreamSynchronize(s));
    const int max_leaves = m / 64 + 2;
    pw_leaf_sum_kernel<<<(max_leaves + 127) / 128, 128, 0, s>>>(b.terms, b.leaf_off, b.leaf_len, b.counters + 3, b.leaf_sum);
    MMAD_LAUNCHED();
    pw_combine_kernel<<<1, 1, 0, s>>>(m_ptr, b.leaf_sum, any_decrease, b.result);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    MMAD_CUDA_OK(cudaMemcpyAsync(h_out, b.result, sizeof(double), cudaMemcpyDeviceToHost, s));
    MMAD_CUDA_OK(cudaStreamSynchronize(s));
    return MMAD_OK;
}

int check_args(const void* a, const void* lab, long long n, const void* out, void* ws, size_t ws_bytes, bool need_label) {
    if (!a || (need_label && !lab) || !out || n < 1 || n > 2000000000LL) { set_error("bad metric argument (n=%lld)", n); return MMAD_E_ARG; }
    if (!ws || ws_bytes < layout(nullptr, n).total) { set_error("metric workspace too small"); return MMAD_E_WORKSPACE; }
    return MMAD_OK;
}

}  // namespace

extern "C" {

size_t mmad_metric_workspace_bytes(long long n) { return n < 1 ? 0 : layout(nullptr, n).total; }

int mmad_auc_roc(const float* d_score, const uint8_t* d_label, long long n, double* h_out, void* d_ws, size_t ws_bytes,
                 void* stream) {
    int rc = check_args(d_score, d_label, n, h_out, d_ws, ws_bytes, true);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    Bufs b = layout(d_ws, n);
    int nonfinite = 0;
    if ((rc = curve_front(d_score, d_label, n, b, s, &nonfinite))) return rc;
    if (nonfinite) { *h_out = 0.0; return MMAD_OK; }     // sklearn raises, utils/metric.py:43-44 returns .0
    int m = 0;
    MMAD_CUDA_OK(cudaMemcpyAsync(&m, b.counters + 0, sizeof(int), cudaMemcpyDeviceToHost, s));
    MMAD_CUDA_OK(cudaStreamSynchronize(s));
    keep_kernel<<<grid_for(m), 256, 0, s>>>(b.idx, b.tps, m, b.keep);
    MMAD_LAUNCHED();
    size_t tmp = b.cub_bytes;
    MMAD_CUDA_OK(cub::DeviceSelect::Flagged(b.cub_tmp, tmp, b.idx, b.keep, b.idx2, b.counters + 1, m, s));
    tmp = b.cub_bytes;
    MMAD_CUDA_OK(cub::DeviceSelect::Flagged(b.cub_tmp, tmp, b.tps, b.keep, b.tps2, b.counters + 1, m, s));
    g_launches += 2;
    roc_terms_kernel<<<grid_for(m), 256, 0, s>>>(b.idx2, b.tps2, b.counters + 1, b.terms);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return finish_area(b, b.counters + 1, nullptr, h_out, s);
}

int mmad_auc_prc(const float* d_score, const uint8_t* d_label, long long n, double* h_out, void* d_ws, size_t ws_bytes,
                 void* stream) {
    int rc = check_args(d_score, d_label, n, h_out, d_ws, ws_bytes, true);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    Bufs b = layout(d_ws, n);
    int nonfinite = 0;
    if ((rc = curve_front(d_score, d_label, n, b, s, &nonfinite))) return rc;
    if (nonfinite) { *h_out = 0.0; return MMAD_OK; }
    int m = 0;
    MMAD_CUDA_OK(cudaMemcpyAsync(&m, b.counters + 0, sizeof(int), cudaMemcpyDeviceToHost, s));
    MMAD_CUDA_OK(cudaStreamSynchronize(s));
    pr_terms_kernel<<<grid_for(m), 256, 0, s>>>(b.idx, b.tps, b.counters + 0, b.terms, b.counters + 4);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return finish_area(b, b.counters + 0, b.counters + 4, h_out, s);
}

int mmad_quantile(const float* d_valid, long long n, float q, float* h_out, void* d_ws, size_t ws_bytes, void* stream) {
    int rc = check_args(d_valid, nullptr, n, h_out, d_ws, ws_bytes, false);
    if (rc) return rc;
    if (!(q >= 0.f && q <= 1.f)) { set_error("quantile must be in [0,1]"); return MMAD_E_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    Bufs b = layout(d_ws, n);
    size_t tmp = b.cub_bytes;
    MMAD_CUDA_OK(cub::DeviceRadixSort::SortKeys(b.cub_tmp, tmp, d_valid, b.keys, (int)n, 0, 32, s));
    quantile_kernel<<<1, 1, 0, s>>>(b.keys, (int)n, q, b.keys_in);
    g_launches += 2;
    MMAD_CUDA_OK(cudaGetLastError());
    MMAD_CUDA_OK(cudaMemcpyAsync(h_out, b.keys_in, sizeof(float), cudaMemcpyDeviceToHost, s));
    MMAD_CUDA_OK(cudaStreamSynchronize(s));
    return MMAD_OK;
}

int mmad_confusion(const float* d_score, const uint8_t* d_label, long long n, float thr, int strict, long long* h_counts,
                   void* d_ws, size_t ws_bytes, void* stream) {
    int rc = check_args(d_score, d_label, n, h_counts, d_ws, ws_bytes, true);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    Bufs b = layout(d_ws, n);
    MMAD_CUDA_OK(cudaMemsetAsync(b.counts, 0, 4 * sizeof(long long), s));
    confusion_kernel<<<grid_for(n), 256, 0, s>>>(d_score, d_label, n, thr, strict, (unsigned long long*)b.counts);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    MMAD_CUDA_OK(cudaMemcpyAsync(h_counts, b.counts, 4 * sizeof(long long), cudaMemcpyDeviceToHost, s));
    MMAD_CUDA_OK(cudaStreamSynchronize(s));
    return MMAD_OK;
}

}  // extern "C"
