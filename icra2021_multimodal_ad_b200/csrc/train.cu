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

#include "mmad_internal.cuh"

using namespace mmad;

namespace mmad {
// defined in mmad_api.cu
const mmad_desc_t* handle_desc(mmad_t h);
}

namespace {

constexpr int kColThreads = 128;

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

inline dim3 col_grid(int n_cols, int rows) {
    int slabs = rows >= 4096 ? 32 : (rows >= 1024 ? 16 : (rows >= 128 ? 8 : 1));
    return dim3((n_cols + kColThreads - 1) / kColThreads, slabs);
}

// st[0][c] += sum_r a, st[1][c] += sum_r a^2 with a = lrelu(pre[r,c])
__global__ void bn_fwd_stats_kernel(const float* __restrict__ pre, int ld, int B, int N, float slope, double* __restrict__ st,
                                    int st_stride) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const int per = (B + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * per, r1 = min(B, r0 + per);
    double s = 0.0, q = 0.0;
    for (int r = r0; r < r1; ++r) {
        const double a = (double)lrelu(pre[(size_t)r * ld + c], slope);
        s += a;
        q = fma(a, a, q);
    }
    atomicAdd(&st[c], s);
    atomicAdd(&st[st_stride + c], q);
}

// mean / inv-std of the (global) batch, running-stat update (torch BatchNorm1d: momentum, unbiased running var)
__global__ void bn_fwd_finalize_kernel(const double* __restrict__ st, int st_stride, int N, int Np, double Bg, float eps,
                                       float momentum, float* __restrict__ run_mean, float* __restrict__ run_var,
                                       long long* __restrict__ nbt, float* __restrict__ mean_out, float* __restrict__ inv_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt) *nbt += 1;
    if (c >= Np) return;
    if (c >= N) { mean_out[c] = 0.f; inv_out[c] = 0.f; return; }
    const double m = st[c] / Bg;
    double var_b = st[st_stride + c] / Bg - m * m;
    if (var_b < 0.0) var_b = 0.0;
    mean_out[c] = (float)m;
    inv_out[c] = (float)(1.0 / sqrt(var_b + (double)eps));
    if (run_mean) {
        const double unb = Bg > 1.0 ? var_b * (Bg / (Bg - 1.0)) : var_b;
        run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)m;
        run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)unb;
    }
}

// out[r,c] = (lrelu(pre) - mean) * inv * gamma + beta ; padding columns [N, Np) are zero filled
__global__ void bn_fwd_apply_kernel(const float* __restrict__ pre, int ld, int B, int N, int Np, float slope,
                                    const float* __restrict__ mean, const float* __restrict__ inv,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out,
                                    int ldo) {
    const size_t total = (size_t)B * Np;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / Np), c = (int)(i % Np);
        float v = 0.f;
        if (c < N) {
            const float a = lrelu(pre[(size_t)r * ld + c], slope);
            v = fmaf((a - mean[c]) * inv[c], gamma[c], beta[c]);
        }
        out[(size_t)r * ldo + c] = v;
    }
}

// st[0][c] += sum_r g, st[1][c] += sum_r g * xhat
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ pre, int ld, int B, int N,
                                     float slope, float gscale, const float* __restrict__ mean, const float* __restrict__ inv,
                                     double* __restrict__ st, int st_stride) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const int per = (B + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * per, r1 = min(B, r0 + per);
    const float m = mean[c], iv = inv[c];
    double s1 = 0.0, s2 = 0.0;
    for (int r = r0; r < r1; ++r) {
        const float gv = g[(size_t)r * ldg + c] * gscale;
        const float xh = (lrelu(pre[(size_t)r * ld + c], slope) - m) * iv;
        s1 += (double)gv;
        s2 = fma((double)gv, (double)xh, s2);
    }
    atomicAdd(&st[c], s1);
    atomicAdd(&st[st_stride + c], s2);
}

// local parameter gradients of BatchNorm (before any cross-rank combination of the statistics)
__global__ void bn_bwd_param_kernel(const double* __restrict__ st, int st_stride, int N, float* __restrict__ ggamma,
                                    float* __restrict__ gbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    gbeta[c] = (float)st[c];
    ggamma[c] = (float)st[st_stride + c];
}

// g_pre = gamma inv (g - s1/Bg - xhat s2/Bg) * lrelu'(pre);  st[2][c] += sum_r g_pre (bias gradient)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ pre, int ld, int B, int N,
                                    float slope, float gscale, const float* __restrict__ mean, const float* __restrict__ inv,
                                    const float* __restrict__ gamma, double* __restrict__ st, int st_stride, double Bg,
                                    float* __restrict__ gpre, int ldo) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const int per = (B + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * per, r1 = min(B, r0 + per);
    const float m = mean[c], iv = inv[c];
    const float k = gamma[c] * iv;
    const float a1 = (float)(st[c] / Bg), a2 = (float)(st[st_stride + c] / Bg);
    double sb = 0.0;
    for (int r = r0; r < r1; ++r) {
        const float p = pre[(size_t)r * ld + c];
        const float xh = (lrelu(p, slope) - m) * iv;
        float ga = k * (g[(size_t)r * ldg + c] * gscale - a1 - xh * a2);
        ga = p > 0.f ? ga : ga * slope;
        gpre[(size_t)r * ldo + c] = ga;
        sb += (double)ga;
    }
    atomicAdd(&st[2 * st_stride + c], sb);
}

// st[2][c] += scale * sum_r g[r,c]   (bias gradient of a bare Linear layer)
__global__ void col_sum_scaled_kernel(const float* __restrict__ g, int ldg, int B, int N, float scale, double* __restrict__ st,
                                      int st_stride) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const int per = (B + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * per, r1 = min(B, r0 + per);
    double s = 0.0;
    for (int r = r0; r < r1; ++r) s += (double)g[(size_t)r * ldg + c];
    atomicAdd(&st[2 * st_stride + c], s * (double)scale);
}

__global__ void bias_grad_out_kernel(const double* __restrict__ st, int st_stride, int N, float* __restrict__ gb) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < N) gb[c] = (float)st[2 * st_stride + c];
}

// VIB forward (decorators/variational_info_bottleneck.py:19-42, k = 1): enc_out [B, 2h] -> z = eps exp(logvar/2) + mu
// (zero padded to ldz columns); kl_acc += -1/2 sum (1 + logvar - mu^2 - exp(logvar))
__global__ void vib_train_fwd_kernel(const float* __restrict__ o, int ldo, int B, int h, const float* __restrict__ eps,
                                     float* __restrict__ z, int ldz, double* __restrict__ kl_acc) {
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
                                     const float* __restrict__ eps, float beta, float* __restrict__ gout, int ldgo) {
    const size_t total = (size_t)B * h;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / h), j = (int)(i % h);
        const float m = o[(size_t)b * ldo + j], lv = o[(size_t)b * ldo + h + j];
        const float g = gz[(size_t)b * ldg + j];
        const float sg = expf(0.5f * lv);
        gout[(size_t)b * ldgo + j] = fmaf(beta, m, g);
        gout[(size_t)b * ldgo + h + j] = g * eps[i] * 0.5f * sg - 0.5f * beta * (1.f - expf(lv));
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
    for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
        const float gv = g[i] * gscale;
        float mv = m[i], vv = v[i];
        mv = mv + w1 * (gv - mv);
        vv = vv * b2 + omb2 * gv * gv;      // mul_ then addcmul_ (two roundings like torch; FMA contraction is off below)
        m[i] = mv;
        v[i] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        p[i] = p[i] + neg_step * (mv / denom);
    }
}

struct TrainPlan {
    size_t total = 0;
    int B = 0;
    size_t pre[2][MMAD_MAX_LAYERS] = {{0}};
    size_t out[2][MMAD_MAX_LAYERS] = {{0}};
    size_t mean[2][MMAD_MAX_LAYERS] = {{0}};
    size_t inv[2][MMAD_MAX_LAYERS] = {{0}};
    size_t st[2][MMAD_MAX_LAYERS] = {{0}};    // [3][st_stride] doubles per layer
    size_t st_all = 0, st_bytes = 0;
    size_t g[2] = {0, 0};                      // gradient ping-pong [B, maxNp]
    size_t z = 0, genc = 0;                    // VIB: sampled code, gradient wrt the encoder output
    size_t rowpart = 0;
    size_t kl = 0;
    int maxNp = 0;
};

int np_of(int n) { return round_up(n, kPad); }

TrainPlan make_train_plan(const mmad_desc_t& d, int B) {
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
        for (int i = 0; i < n; ++i) p.st[m][i] = take((size_t)3 * np_of(w[i + 1]) * 8);
    }
    p.kl = take(8);
    p.st_bytes = off - p.st_all;
    for (int m = 0; m < 2; ++m) {
        const int n = m == 0 ? d.n_enc : d.n_dec;
        const int* w = m == 0 ? d.enc_widths : d.dec_widths;
        for (int i = 0; i < n; ++i) {
            const int Np = np_of(w[i + 1]);
            p.pre[m][i] = take((size_t)B * Np * 4);
            const bool bn = i < n - 1;
            p.out[m][i] = bn ? take((size_t)B * Np * 4) : p.pre[m][i];
            if (bn) { p.mean[m][i] = take((size_t)Np * 4); p.inv[m][i] = take((size_t)Np * 4); }
        }
    }
    p.g[0] = take((size_t)B * maxNp * 4);
    p.g[1] = take((size_t)B * maxNp * 4);
    p.z = take((size_t)B * np_of(d.dec_widths[0]) * 4);
    p.genc = take((size_t)B * np_of(d.enc_widths[d.n_enc]) * 4);
    p.rowpart = take((size_t)((d.enc_widths[0] + 63) / 64 + 1) * B * 4);
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
    return make_train_plan(*handle_desc(h), batch).total;
}

int mmad_train_fwd_bwd(mmad_t h, const float* d_x, int ldx, int batch, long long global_batch,
                       const mmad_train_layer_t* enc, const mmad_train_layer_t* dec, const float* d_eps, float beta_kl,
                       float bn_momentum, float* d_loss, void* d_ws, size_t ws_bytes, mmad_allreduce_fn allreduce,
                       void* allreduce_ctx, void* stream) {
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
    const TrainPlan p = make_train_plan(d, batch);
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
    const int B = batch;
    const double Bg = (double)global_batch;
    const float slope = d.lrelu_slope;
    MMAD_CUDA_OK(cudaMemsetAsync(ws + p.st_all, 0, p.st_bytes, s));
    MMAD_CUDA_OK(cudaMemsetAsync(d_loss, 0, 4, s));

    struct Ref { int m, i, K, N, Np; const mmad_train_layer_t* L; const float* in; int ldin; bool bn; };
    std::vector<Ref> order;
    // ------------------------------- forward -------------------------------
    const float* cur = d_x;
    int ldcur = ldx;
    for (int m = 0; m < 2; ++m) {
        const int n = m == 0 ? d.n_enc : d.n_dec;
        const int* w = m == 0 ? d.enc_widths : d.dec_widths;
        const mmad_train_layer_t* Ls = m == 0 ? enc : dec;
        if (m == 1 && vib) {
            const int hp = np_of(dec_in);
            vib_train_fwd_kernel<<<ew_grid((size_t)B * hp), 256, 0, s>>>(cur, ldcur, B, dec_in, d_eps, (float*)(ws + p.z), hp,
                                                                        (double*)(ws + p.kl));
            MMAD_LAUNCHED();
            cur = (const float*)(ws + p.z);
            ldcur = hp;
        }
        for (int i = 0; i < n; ++i) {
            const int K = w[i], N = w[i + 1], Np = np_of(N);
            const bool bn = i < n - 1;
            const mmad_train_layer_t& L = Ls[i];
            order.push_back({m, i, K, N, Np, &L, cur, ldcur, bn});
            float* pre = (float*)(ws + p.pre[m][i]);
            GemmShape g;
            g.M = B; g.N = N; g.K = K; g.A = cur; g.lda = ldcur; g.B = L.W; g.ldb = K;
            Epilogue e;
            e.bias = L.b;
            e.slope = slope;
            const bool last = (m == 1 && i == n - 1);
            if (bn) {
                e.pre = pre; e.ldpre = Np;
            } else {
                e.Y = pre; e.ldy = Np; e.y_cols = Np;
                if (last) {   // loss epilogue: d = xhat - x, row sums of d^2
                    e.ref = d_x; e.ldref = ldx;
                    e.dout = (float*)(ws + p.g[0]); e.lddout = p.maxNp; e.d_cols = N;
                    e.rowpart = (float*)(ws + p.rowpart); e.rowpart_stride = B;
                }
            }
            int rc = gemm_simt(g, e, s);
            if (rc) return rc;
            if (bn) {
                double* st = (double*)(ws + p.st[m][i]);
                float* mean = (float*)(ws + p.mean[m][i]);
                float* inv = (float*)(ws + p.inv[m][i]);
                float* out = (float*)(ws + p.out[m][i]);
                bn_fwd_stats_kernel<<<col_grid(N, B), kColThreads, 0, s>>>(pre, Np, B, N, slope, st, Np);
                MMAD_LAUNCHED();
                if (allreduce) {
                    rc = allreduce(allreduce_ctx, st, 2LL * Np, s);
                    if (rc) { set_error("allreduce hook failed (%d)", rc); return MMAD_E_STATE; }
                }
                bn_fwd_finalize_kernel<<<(Np + 127) / 128, 128, 0, s>>>(st, Np, N, Np, Bg, d.bn_eps, bn_momentum, L.run_mean,
                                                                        L.run_var, L.num_batches_tracked, mean, inv);
                MMAD_LAUNCHED();
                bn_fwd_apply_kernel<<<ew_grid((size_t)B * Np), 256, 0, s>>>(pre, Np, B, N, Np, slope, mean, inv, L.gamma, L.beta,
                                                                            out, Np);
                MMAD_LAUNCHED();
                cur = out;
            } else {
                cur = pre;
            }
            ldcur = Np;
        }
    }
    {   // loss = sum d^2 (+ beta KL)
        const int slots = (D + gemm_simt_tile_n() - 1) / gemm_simt_tile_n();
        int rc = reduce_sum_all((const float*)(ws + p.rowpart), B, B, 0, slots, d_loss, s);
        if (rc) return rc;
        if (vib) {
            loss_finish_kernel<<<1, 32, 0, s>>>(d_loss, (const double*)(ws + p.kl), beta_kl);
            MMAD_LAUNCHED();
        }
    }
    // ------------------------------- backward -------------------------------
    // g (ping) holds d = xhat - x; dL/dxhat = 2 d enters through gscale of the first backward layer
    int gi = 0;
    float gscale = 2.f;
    for (int idx = (int)order.size() - 1; idx >= 0; --idx) {
        const Ref& r = order[idx];
        const mmad_train_layer_t& L = *r.L;
        double* st = (double*)(ws + p.st[r.m][r.i]);
        const float* gin = (const float*)(ws + p.g[gi]);
        int ldg = p.maxNp;
        if (vib && r.m == 0 && r.i == d.n_enc - 1) { gin = (const float*)(ws + p.genc); ldg = np_of(enc_out); }
        const float* gpre = gin;
        int ldgpre = ldg;
        float gemm_scale = gscale;
        if (r.bn) {
            const float* pre = (const float*)(ws + p.pre[r.m][r.i]);
            const float* mean = (const float*)(ws + p.mean[r.m][r.i]);
            const float* inv = (const float*)(ws + p.inv[r.m][r.i]);
            MMAD_CUDA_OK(cudaMemsetAsync(st, 0, (size_t)2 * r.Np * 8, s));
            bn_bwd_reduce_kernel<<<col_grid(r.N, B), kColThreads, 0, s>>>(gin, ldg, pre, r.Np, B, r.N, slope, gscale, mean, inv, st, r.Np);
            MMAD_LAUNCHED();
            bn_bwd_param_kernel<<<(r.N + 127) / 128, 128, 0, s>>>(st, r.Np, r.N, L.ggamma, L.gbeta);
            MMAD_LAUNCHED();
            if (allreduce) {
                int rc = allreduce(allreduce_ctx, st, 2LL * r.Np, s);
                if (rc) { set_error("allreduce hook failed (%d)", rc); return MMAD_E_STATE; }
            }
            float* go = (float*)(ws + p.g[gi ^ 1]);
            bn_bwd_apply_kernel<<<col_grid(r.N, B), kColThreads, 0, s>>>(gin, ldg, pre, r.Np, B, r.N, slope, gscale, mean, inv, L.gamma,
                                                                         st, r.Np, Bg, go, p.maxNp);
            MMAD_LAUNCHED();
            gpre = go; ldgpre = p.maxNp;
            gi ^= 1;
            gemm_scale = 1.f;
        } else {
            col_sum_scaled_kernel<<<col_grid(r.N, B), kColThreads, 0, s>>>(gin, ldg, B, r.N, gscale, st, r.Np);
            MMAD_LAUNCHED();
        }
        bias_grad_out_kernel<<<(r.N + 127) / 128, 128, 0, s>>>(st, r.Np, r.N, L.gb);
        MMAD_LAUNCHED();
        {   // gW[N,K] = g_pre^T in
            GemmShape g;
            g.M = r.N; g.N = r.K; g.K = B;
            g.A = gpre; g.lda = ldgpre; g.transA = true;
            g.B = r.in; g.ldb = r.ldin; g.transB = true;
            Epilogue e;
            e.acc_scale = gemm_scale;
            e.Y = L.gW; e.ldy = r.K; e.y_cols = r.K;
            int rc = gemm_simt(g, e, s);
            if (rc) return rc;
        }
        if (idx > 0) {   // g_in[B,K] = g_pre W
            const bool into_vib = vib && r.m == 1 && r.i == 0;
            float* gout = (float*)(ws + p.g[gi ^ 1]);
            GemmShape g;
            g.M = B; g.N = r.K; g.K = r.N;
            g.A = gpre; g.lda = ldgpre;
            g.B = L.W; g.ldb = r.K; g.transB = true;
            Epilogue e;
            e.acc_scale = gemm_scale;
            e.Y = gout; e.ldy = p.maxNp; e.y_cols = r.K;
            int rc = gemm_simt(g, e, s);
            if (rc) return rc;
            gi ^= 1;
            if (into_vib) {
                const Ref& er = order[idx - 1];
                vib_train_bwd_kernel<<<ew_grid((size_t)B * dec_in), 256, 0, s>>>(gout, p.maxNp, (const float*)(ws + p.pre[0][er.i]), er.Np,
                                                                                 B, dec_in, d_eps, beta_kl, (float*)(ws + p.genc), np_of(enc_out));
                MMAD_LAUNCHED();
            }
        }
        gscale = 1.f;
    }
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

int mmad_adam_step(int n_tensors, float* const* h_params, float* const* h_grads, float* const* h_m, float* const* h_v,
                   const long long* h_numel, int step, float lr, float beta1, float beta2, float eps, float grad_scale,
                   void* stream) {
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
