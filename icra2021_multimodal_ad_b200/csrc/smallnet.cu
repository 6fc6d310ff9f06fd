// Fused scoring chain for the per-modality autoencoders (utils/data_loaders.py:16-29: force_torque D = 64, mic D = 128; every
// width <= 128): enc(x) -> dec -> [enc(x) | enc(xhat)] with diffs, base and SAP scores, ONE kernel, exact fp32 FMA arithmetic.
//
// north_star (1): "because the MLP widths are small ... the layer chain is fused per batch tile".  The per-layer kernels move
// every activation through HBM/L2 (twins + stash, ~16 B per element per layer) and launch 15 times per chunk; here a CTA owns
// a tile of 32 windows and walks all layers with the activations in shared memory:
//   * pass A (10 layers): h = enc(x), xhat = dec(h) on the tile's 32 rows; d_0 = xhat - x -> base score;
//   * pass B (5 layers): the encoder again on the STACKED tile [x rows | xhat rows] (64 rows) -- the reference's own order
//     (reconstruction_aggregation.py:25-27 pushes x and xhat through every layer) -- so no stash of enc(x) is needed and
//     two CTAs fit an SM; a thread owns the x row AND the xhat row of the same windows, the diff is register-local.
//   * weights are read straight from global memory in a transposed, zero-padded copy Wt[k][128] made at plan time: the
//     whole model is 0.3-0.5 MB, one layer (<= 64 KB) is L1-resident while a CTA works on it, and a half-warp's loads are
//     whole 256-byte rows.
// Thread tile: R rows x 8 columns (columns 4*tx.. and 64 + 4*tx..), 256 threads = 16 column groups x 16 row groups;
// inner product over k in steps of 4 (LDS.128 of the activations, broadcast across the half-warp).
#include "mmad_internal.cuh"

namespace mmad {

namespace {

constexpr int SN_T = 256;
constexpr int SN_BM_MAX = 64;        // windows per tile: 64 (thread tile 4 / 8 rows x 8 columns) or 32 (2 / 4 rows)
constexpr int SN_W = 128;            // padded width of every layer
constexpr int SN_LD = SN_W + 4;      // shared-memory row stride (floats): rows 16-byte aligned, 4-bank skew per row
constexpr int SN_MAX_STEPS = 3 * MMAD_MAX_LAYERS;

struct SnStep {
    const float* Wt;       // [Kp4][128] transposed weights, zero padded
    const float* bias;     // [128]
    const float* scale;    // [128] eval-BN scale (nullptr: bare Linear)
    const float* shift;
    int K4;                // ceil(K / 4)
    int N;
};

struct SnPlan {
    int n_enc, n_dec, D, lo, hi, last;     // last: highest encoder layer whose diff is wanted
    float inv_base, inv_sap, slope;
    SnStep enc[MMAD_MAX_LAYERS], dec[MMAD_MAX_LAYERS];
};

// one fused layer on R rows per thread: out[r][c] = epi(sum_k in[r][k] * Wt[k][c]).  `in` / `out`: shared tiles [rows][SN_LD].
// Rows of thread ty: the first RA = BM / 16 are windows RA * ty + i (x path); the next RA, if any, the same windows' xhat rows
// (BM + RA * ty + i): a thread owns both paths of its windows, so pass B's diffs are register-local.
// The accumulators are transformed in place and handed back in v.
template <int R, int BM>
__device__ __forceinline__ void sn_layer(const SnStep& st, const float* __restrict__ in, float* __restrict__ out, float slope,
                                         int tx, int ty, float (&v)[R][8]) {
    constexpr int RA = BM / 16;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    int row[R];
#pragma unroll
    for (int i = 0; i < R; ++i) row[i] = (i < RA ? 0 : BM) + RA * ty + (i % RA);
    const float4* wt = reinterpret_cast<const float4*>(st.Wt);
#pragma unroll 2
    for (int k4 = 0; k4 < st.K4; ++k4) {
        float4 a[R];
#pragma unroll
        for (int i = 0; i < R; ++i) a[i] = *reinterpret_cast<const float4*>(in + row[i] * SN_LD + 4 * k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float4 w0 = __ldg(wt + (size_t)(4 * k4 + kk) * (SN_W / 4) + tx);
            const float4 w1 = __ldg(wt + (size_t)(4 * k4 + kk) * (SN_W / 4) + 16 + tx);
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
                v[i][0] = fmaf(av, w0.x, v[i][0]); v[i][1] = fmaf(av, w0.y, v[i][1]);
                v[i][2] = fmaf(av, w0.z, v[i][2]); v[i][3] = fmaf(av, w0.w, v[i][3]);
                v[i][4] = fmaf(av, w1.x, v[i][4]); v[i][5] = fmaf(av, w1.y, v[i][5]);
                v[i][6] = fmaf(av, w1.z, v[i][6]); v[i][7] = fmaf(av, w1.w, v[i][7]);
            }
        }
    }
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(st.bias) + tx), b1 = __ldg(reinterpret_cast<const float4*>(st.bias) + 16 + tx);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    if (st.scale) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(st.scale) + tx), s1 = __ldg(reinterpret_cast<const float4*>(st.scale) + 16 + tx);
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(st.shift) + tx), h1 = __ldg(reinterpret_cast<const float4*>(st.shift) + 16 + tx);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float x = v[i][j] + bb[j];
                x = x > 0.f ? x : x * slope;
                v[i][j] = fmaf(x, sc[j], sh[j]);          // padded columns: zero weights, bias, scale, shift -> 0
            }
    } else {
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i][j] += bb[j];
    }
    if (out) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            *reinterpret_cast<float4*>(out + row[i] * SN_LD + 4 * tx) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
            *reinterpret_cast<float4*>(out + row[i] * SN_LD + 64 + 4 * tx) = make_float4(v[i][4], v[i][5], v[i][6], v[i][7]);
        }
    }
}

// sum over the 16 column groups (the 16 lanes of a half-warp share ty)
__device__ __forceinline__ float sn_half_warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    return v;
}

template <int BM>
__global__ void __launch_bounds__(SN_T, BM == 32 ? 2 : 1)
smallnet_chain_kernel(const SnPlan* __restrict__ P, const float* __restrict__ x, int ldx, int n, float* __restrict__ base_out,
                      float* __restrict__ sap_out) {
    constexpr int RA = BM / 16;
    extern __shared__ __align__(16) float sn_smem[];
    float* xs = sn_smem;                               // [BM][SN_LD]      the tile's input rows (zero padded to 128 columns)
    float* A0 = xs + BM * SN_LD;                       // [2 BM][SN_LD]    ping
    float* A1 = A0 + 2 * BM * SN_LD;                   // [2 BM][SN_LD]    pong
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int D = P->D, L = P->n_enc, Ld = P->n_dec;
    const float slope = P->slope;
    const int tiles = (n + BM - 1) / BM;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int r0 = t * BM;
        __syncthreads();                               // the previous tile's readers are done
        for (int i = tid; i < BM * (SN_W / 4); i += SN_T) {
            const int r = i / (SN_W / 4), c = (i % (SN_W / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + r < n) {
                const float* src = x + (size_t)(r0 + r) * ldx + c;
                if (c + 3 < D) v = __ldg(reinterpret_cast<const float4*>(src));        // ldx % 4 == 0 and 16-byte base checked by the host
                else { if (c < D) v.x = src[0]; if (c + 1 < D) v.y = src[1]; if (c + 2 < D) v.z = src[2]; }
            }
            *reinterpret_cast<float4*>(xs + r * SN_LD + c) = v;
        }
        __syncthreads();
        // ---- pass A: encoder and decoder on the BM x rows ----
        const float* cur = xs;
        float* nxt = A0;
        float va[RA][8];
        for (int l = 0; l < L; ++l) {
            sn_layer<RA, BM>(P->enc[l], cur, nxt, slope, tx, ty, va);
            __syncthreads();
            cur = nxt; nxt = (nxt == A0) ? A1 : A0;
        }
        for (int l = 0; l < Ld; ++l) {
            float* out = (l == Ld - 1) ? nullptr : nxt;          // the reconstruction stays in registers
            sn_layer<RA, BM>(P->dec[l], cur, out, slope, tx, ty, va);
            if (l < Ld - 1) { __syncthreads(); cur = nxt; nxt = (nxt == A0) ? A1 : A0; }
        }
        // d_0 = xhat - x (thread-local: this thread's windows, its 8 columns), base score, stacked tile [x | xhat] for pass B
        float base_r[RA], sap_r[RA];
        float* S = nxt;                                // the free buffer
#pragma unroll
        for (int i = 0; i < RA; ++i) {
            const int r = RA * ty + i;
            const float4 x0 = *reinterpret_cast<const float4*>(xs + r * SN_LD + 4 * tx);
            const float4 x1 = *reinterpret_cast<const float4*>(xs + r * SN_LD + 64 + 4 * tx);
            const float d[8] = {va[i][0] - x0.x, va[i][1] - x0.y, va[i][2] - x0.z, va[i][3] - x0.w,
                                va[i][4] - x1.x, va[i][5] - x1.y, va[i][6] - x1.z, va[i][7] - x1.w};
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) s = fmaf(d[j], d[j], s);
            base_r[i] = sn_half_warp_sum(s);
            sap_r[i] = P->lo == 0 ? base_r[i] : 0.f;
            *reinterpret_cast<float4*>(S + r * SN_LD + 4 * tx) = x0;
            *reinterpret_cast<float4*>(S + r * SN_LD + 64 + 4 * tx) = x1;
            *reinterpret_cast<float4*>(S + (BM + r) * SN_LD + 4 * tx) = make_float4(va[i][0], va[i][1], va[i][2], va[i][3]);
            *reinterpret_cast<float4*>(S + (BM + r) * SN_LD + 64 + 4 * tx) = make_float4(va[i][4], va[i][5], va[i][6], va[i][7]);
        }
        __syncthreads();
        // ---- pass B: the encoder on [x | xhat], diffs between this thread's x rows and xhat rows ----
        cur = S; nxt = (S == A0) ? A1 : A0;
        float vb[2 * RA][8];
        for (int l = 1; l <= P->last; ++l) {
            sn_layer<2 * RA, BM>(P->enc[l - 1], cur, l == P->last ? nullptr : nxt, slope, tx, ty, vb);
            if (l >= P->lo && l < P->hi) {
#pragma unroll
                for (int i = 0; i < RA; ++i) {
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float d = vb[RA + i][j] - vb[i][j]; s = fmaf(d, d, s); }
                    sap_r[i] += sn_half_warp_sum(s);
                }
            }
            if (l < P->last) { __syncthreads(); cur = nxt; nxt = (nxt == A0) ? A1 : A0; }
        }
        if (tx == 0) {
#pragma unroll
            for (int i = 0; i < RA; ++i) {
                const int r = r0 + RA * ty + i;
                if (r < n) {
                    if (base_out) base_out[r] = base_r[i] * P->inv_base;
                    if (sap_out) sap_out[r] = sap_r[i] * P->inv_sap;
                }
            }
        }
    }
}

__global__ void sn_transpose_kernel(const float* __restrict__ W, int N, int K, int ldw, float* __restrict__ Wt, int K4) {
    // Wt[k][c] = W[c][k] for c < N, k < K; zero elsewhere (k < 4 * K4, c < 128)
    const int total = 4 * K4 * SN_W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i / SN_W, c = i % SN_W;
        Wt[i] = (c < N && k < K) ? W[(size_t)c * ldw + k] : 0.f;
    }
}

__global__ void sn_pad_vec_kernel(const float* __restrict__ v, int N, float* __restrict__ out) {
    const int c = threadIdx.x;
    if (c < SN_W) out[c] = (v && c < N) ? v[c] : 0.f;
}

struct SmallNetState {
    bool ok = false;
    int lo = -1, hi = -1;
    unsigned long long weights_gen = 0;
    SnPlan* d_plan = nullptr;
    float* d_pack = nullptr;       // transposed weights + padded vectors of every layer
    size_t pack_floats = 0;
    int grid_max[2] = {0, 0};      // BM = 32, 64
    bool attr_set = false;
};

constexpr int sn_smem_bytes(int bm) { return 5 * bm * SN_LD * 4; }    // xs + two stacked tiles: 84 480 B (BM 32) / 168 960 B (BM 64)

}  // namespace

void smallnet_state_free(void* p) {
    SmallNetState* s = static_cast<SmallNetState*>(p);
    if (!s) return;
    cudaFree(s->d_plan); cudaFree(s->d_pack);
    delete s;
}

bool smallnet_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MMAD_NO_SMALLNET"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

// every layer width (and D) <= 128
bool smallnet_fits(mmad_t h) {
    const mmad_desc_t* d = handle_desc(h);
    for (int i = 0; i <= d->n_enc; ++i) if (d->enc_widths[i] > SN_W) return false;
    for (int i = 0; i <= d->n_dec; ++i) if (d->dec_widths[i] > SN_W) return false;
    return true;
}

int smallnet_prepare(mmad_t h, int lo, int hi, cudaStream_t s) {
    SmallNetState* S = static_cast<SmallNetState*>(handle_smallnet_get(h));
    if (S && S->ok && S->lo == lo && S->hi == hi && S->weights_gen == handle_weights_gen(h)) return MMAD_OK;
    const mmad_desc_t* d = handle_desc(h);
    const int L = d->n_enc, Ld = d->n_dec;
    if (!S) {
        S = new SmallNetState();
        handle_smallnet_set(h, S);
        MMAD_CUDA_OK(cudaMalloc(&S->d_plan, sizeof(SnPlan)));
    }
    S->ok = false;
    // pack: per layer Wt [4*K4][128] + bias, scale, shift [128 each]
    size_t need = 0;
    for (int m = 0; m < 2; ++m)
        for (int i = 0; i < (m ? Ld : L); ++i) {
            const LayerF32 Lr = handle_layer_f32(h, m, i);
            need += (size_t)((Lr.K + 3) / 4 * 4) * SN_W + 3 * SN_W;
        }
    if (need > S->pack_floats) {
        MMAD_CUDA_OK(cudaStreamSynchronize(s));
        cudaFree(S->d_pack);
        MMAD_CUDA_OK(cudaMalloc(&S->d_pack, need * 4));
        S->pack_floats = need;
    }
    SnPlan P;
    memset(&P, 0, sizeof P);
    P.n_enc = L; P.n_dec = Ld; P.D = d->enc_widths[0]; P.lo = lo; P.hi = hi;
    P.last = hi > 1 ? std::min(L, hi - 1) : 0;
    P.slope = d->lrelu_slope;
    P.inv_base = 1.f / P.D;
    int dsel = 0;
    for (int l = lo; l < hi; ++l) dsel += d->enc_widths[l];
    P.inv_sap = 1.f / dsel;
    float* p = S->d_pack;
    for (int m = 0; m < 2; ++m)
        for (int i = 0; i < (m ? Ld : L); ++i) {
            const LayerF32 Lr = handle_layer_f32(h, m, i);
            SnStep& st = (m ? P.dec : P.enc)[i];
            st.K4 = (Lr.K + 3) / 4; st.N = Lr.N;
            st.Wt = p;
            sn_transpose_kernel<<<64, 256, 0, s>>>(Lr.W, Lr.N, Lr.K, Lr.Kp, p, st.K4);
            p += (size_t)4 * st.K4 * SN_W;
            st.bias = p;  sn_pad_vec_kernel<<<1, SN_W, 0, s>>>(Lr.bias, Lr.N, p); p += SN_W;
            if (Lr.has_bn) {
                st.scale = p; sn_pad_vec_kernel<<<1, SN_W, 0, s>>>(Lr.scale, Lr.N, p); p += SN_W;
                st.shift = p; sn_pad_vec_kernel<<<1, SN_W, 0, s>>>(Lr.shift, Lr.N, p); p += SN_W;
            } else {
                st.scale = nullptr; st.shift = nullptr; p += 2 * SN_W;
            }
        }
    MMAD_CUDA_OK(cudaGetLastError());
    MMAD_CUDA_OK(cudaMemcpyAsync(S->d_plan, &P, sizeof P, cudaMemcpyHostToDevice, s));
    MMAD_CUDA_OK(cudaStreamSynchronize(s));       // P lives on this stack frame
    if (!S->attr_set) {
        MMAD_CUDA_OK(cudaFuncSetAttribute(smallnet_chain_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, sn_smem_bytes(32)));
        MMAD_CUDA_OK(cudaFuncSetAttribute(smallnet_chain_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, sn_smem_bytes(64)));
        int dev = 0, sms = 148, per = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        MMAD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, smallnet_chain_kernel<32>, SN_T, sn_smem_bytes(32)));
        S->grid_max[0] = sms * std::max(per, 1);
        MMAD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, smallnet_chain_kernel<64>, SN_T, sn_smem_bytes(64)));
        S->grid_max[1] = sms * std::max(per, 1);
        S->attr_set = true;
    }
    S->lo = lo; S->hi = hi;
    S->weights_gen = handle_weights_gen(h);
    S->ok = true;
    return MMAD_OK;
}

// base / SAP scores of n rows (device pointers) for a model whose widths are all <= 128
int smallnet_score(mmad_t h, const float* d_x, int ldx, int n, int lo, int hi, float* d_base, float* d_sap, cudaStream_t s) {
    if (n <= 0) return MMAD_OK;
    SmallNetState* S = static_cast<SmallNetState*>(handle_smallnet_get(h));
    const bool fresh = S && S->ok && S->lo == lo && S->hi == hi && S->weights_gen == handle_weights_gen(h);
    if (!fresh) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(s, &cap);
        if (cap != cudaStreamCaptureStatusNone) { set_error("smallnet plan is stale while the stream is capturing"); return MMAD_E_UNSUPPORTED; }
        int rc = smallnet_prepare(h, lo, hi, s);
        if (rc) return rc;
        S = static_cast<SmallNetState*>(handle_smallnet_get(h));
    }
    // 64-window tiles (8 x 8 register tile in pass B: fewer weight loads per FMA) once they fill the SMs; 32-window tiles
    // (two CTAs per SM) for shorter calls.  MMAD_SMALLNET_BM forces one for experiments.
    static int force = -1;
    if (force < 0) { const char* e = getenv("MMAD_SMALLNET_BM"); force = e ? atoi(e) : 0; }
    const bool big = force ? force == 64 : n >= 64 * S->grid_max[1];
    if (big) {
        const int grid = std::min((n + 63) / 64, S->grid_max[1]);
        smallnet_chain_kernel<64><<<grid, SN_T, sn_smem_bytes(64), s>>>(S->d_plan, d_x, ldx, n, d_base, d_sap);
    } else {
        const int grid = std::min((n + 31) / 32, S->grid_max[0]);
        smallnet_chain_kernel<32><<<grid, SN_T, sn_smem_bytes(32), s>>>(S->d_plan, d_x, ldx, n, d_base, d_sap);
    }
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

}  // namespace mmad
