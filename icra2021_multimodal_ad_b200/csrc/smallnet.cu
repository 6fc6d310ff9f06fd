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
constexpr int SN_BM = 32;            // windows per tile
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
// Rows of thread (ty): r = rows_of(ty, i).  PASS_B: rows {2ty, 2ty+1} (x path) and {32 + 2ty, 32 + 2ty + 1} (xhat path).
template <int R>
__device__ __forceinline__ void sn_layer(const SnStep& st, const float* __restrict__ in, float* __restrict__ out, float slope,
                                         int tx, int ty, float (&v)[R][8]) {
    float acc[R][8];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    int row[R];
#pragma unroll
    for (int i = 0; i < R; ++i) row[i] = (i < 2 ? 0 : SN_BM) + 2 * ty + (i & 1);
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
                acc[i][0] = fmaf(av, w0.x, acc[i][0]); acc[i][1] = fmaf(av, w0.y, acc[i][1]);
                acc[i][2] = fmaf(av, w0.z, acc[i][2]); acc[i][3] = fmaf(av, w0.w, acc[i][3]);
                acc[i][4] = fmaf(av, w1.x, acc[i][4]); acc[i][5] = fmaf(av, w1.y, acc[i][5]);
                acc[i][6] = fmaf(av, w1.z, acc[i][6]); acc[i][7] = fmaf(av, w1.w, acc[i][7]);
            }
        }
    }
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(st.bias) + tx), b1 = __ldg(reinterpret_cast<const float4*>(st.bias) + 16 + tx);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float sc[8], sh[8];
    if (st.scale) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(st.scale) + tx), s1 = __ldg(reinterpret_cast<const float4*>(st.scale) + 16 + tx);
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(st.shift) + tx), h1 = __ldg(reinterpret_cast<const float4*>(st.shift) + 16 + tx);
        sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
        sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float x = acc[i][j] + bb[j];
            if (st.scale) {
                x = x > 0.f ? x : x * slope;
                x = fmaf(x, sc[j], sh[j]);
            }
            v[i][j] = x;            // padded columns: zero weights, zero bias, zero scale/shift -> 0
        }
        if (out) {
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

__global__ void __launch_bounds__(SN_T, 2)
smallnet_chain_kernel(const SnPlan* __restrict__ P, const float* __restrict__ x, int ldx, int n, float* __restrict__ base_out,
                      float* __restrict__ sap_out) {
    extern __shared__ __align__(16) float sn_smem[];
    float* xs = sn_smem;                               // [32][SN_LD]   the tile's input rows (zero padded to 128 columns)
    float* A0 = xs + SN_BM * SN_LD;                    // [64][SN_LD]   ping
    float* A1 = A0 + 2 * SN_BM * SN_LD;                // [64][SN_LD]   pong
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int D = P->D, L = P->n_enc, Ld = P->n_dec;
    const float slope = P->slope;
    const int tiles = (n + SN_BM - 1) / SN_BM;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int r0 = t * SN_BM;
        __syncthreads();                               // the previous tile's readers are done
        for (int i = tid; i < SN_BM * (SN_W / 4); i += SN_T) {
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
        // ---- pass A: encoder and decoder on the 32 x rows ----
        const float* cur = xs;
        float* nxt = A0;
        float v2[2][8];
        for (int l = 0; l < L; ++l) {
            sn_layer<2>(P->enc[l], cur, nxt, slope, tx, ty, v2);
            __syncthreads();
            cur = nxt; nxt = (nxt == A0) ? A1 : A0;
        }
        for (int l = 0; l < Ld; ++l) {
            // the reconstruction goes to rows 32..63 of the tile that pass B starts from: [x | xhat]
            float* out = (l == Ld - 1) ? nullptr : nxt;
            sn_layer<2>(P->dec[l], cur, out, slope, tx, ty, v2);
            if (l < Ld - 1) { __syncthreads(); cur = nxt; nxt = (nxt == A0) ? A1 : A0; }
        }
        // d_0 = xhat - x (thread-local: this thread's rows 2ty, 2ty+1, its 8 columns), base score, stacked tile for pass B
        float base2[2], sap2[2] = {0.f, 0.f};
        float* S = nxt;                                // free buffer: becomes the stacked [x | xhat] tile
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = 2 * ty + i;
            const float4 x0 = *reinterpret_cast<const float4*>(xs + r * SN_LD + 4 * tx);
            const float4 x1 = *reinterpret_cast<const float4*>(xs + r * SN_LD + 64 + 4 * tx);
            const float d[8] = {v2[i][0] - x0.x, v2[i][1] - x0.y, v2[i][2] - x0.z, v2[i][3] - x0.w,
                                v2[i][4] - x1.x, v2[i][5] - x1.y, v2[i][6] - x1.z, v2[i][7] - x1.w};
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) s = fmaf(d[j], d[j], s);
            base2[i] = sn_half_warp_sum(s);
            if (P->lo == 0) sap2[i] = base2[i];
            *reinterpret_cast<float4*>(S + r * SN_LD + 4 * tx) = x0;
            *reinterpret_cast<float4*>(S + r * SN_LD + 64 + 4 * tx) = x1;
            *reinterpret_cast<float4*>(S + (SN_BM + r) * SN_LD + 4 * tx) = make_float4(v2[i][0], v2[i][1], v2[i][2], v2[i][3]);
            *reinterpret_cast<float4*>(S + (SN_BM + r) * SN_LD + 64 + 4 * tx) = make_float4(v2[i][4], v2[i][5], v2[i][6], v2[i][7]);
        }
        __syncthreads();
        // ---- pass B: the encoder on [x | xhat], diffs between this thread's x rows and xhat rows ----
        cur = S; nxt = (S == A0) ? A1 : A0;
        float v4[4][8];
        for (int l = 1; l <= P->last; ++l) {
            sn_layer<4>(P->enc[l - 1], cur, l == P->last ? nullptr : nxt, slope, tx, ty, v4);
            if (l >= P->lo && l < P->hi) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float d = v4[2 + i][j] - v4[i][j]; s = fmaf(d, d, s); }
                    sap2[i] += sn_half_warp_sum(s);
                }
            }
            if (l < P->last) { __syncthreads(); cur = nxt; nxt = (nxt == A0) ? A1 : A0; }
        }
        if (tx == 0) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = r0 + 2 * ty + i;
                if (r < n) {
                    if (base_out) base_out[r] = base2[i] * P->inv_base;
                    if (sap_out) sap_out[r] = sap2[i] * P->inv_sap;
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
    int grid_max = 0;
    bool attr_set = false;
};

constexpr int kSnSmem = (SN_BM + 4 * SN_BM) * SN_LD * 4;    // xs + two stacked tiles = 84 480 B

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
        MMAD_CUDA_OK(cudaFuncSetAttribute(smallnet_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSnSmem));
        int dev = 0, sms = 148, per = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        MMAD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, smallnet_chain_kernel, SN_T, kSnSmem));
        S->grid_max = sms * std::max(per, 1);
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
    const int tiles = (n + SN_BM - 1) / SN_BM;
    const int grid = std::min(tiles, S->grid_max);
    smallnet_chain_kernel<<<grid, SN_T, kSnSmem, s>>>(S->d_plan, d_x, ldx, n, d_base, d_sap);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

}  // namespace mmad
