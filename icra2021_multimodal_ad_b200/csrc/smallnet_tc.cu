// Fused scoring chain of the per-modality autoencoders on the tensor cores (utils/data_loaders.py:16-29: force_torque D = 64,
// mic D = 128; every layer width <= 128), MMAD_PREC_F16X3: north_star item (1) -- "because the MLP widths are small, weights are
// held in shared memory, the layer chain is fused per batch tile, and tcgen05 tensor cores are used ... for the dense
// batch x width GEMMs".
//
// One CTA (512 threads) owns a tile of 128 windows and walks ALL layers; nothing but x and the two scores touches HBM:
//   * activations live in shared memory as fp16 hi/lo twins in the K-major, 128-byte-swizzled layout tcgen05.mma reads
//     (the epilogue writes that layout directly: 16-byte chunk j of row r goes to chunk j ^ (r & 7));
//   * one layer's weights (hi/lo twins of 256 W, <= 64 KB) sit in shared memory, fetched by TMA; the NEXT layer's weights are
//     requested as soon as the current layer's MMAs have completed, so the fetch hides behind the epilogue;
//   * the accumulators sit in TMEM (2 x 128 columns); four threads share a row (32 columns each; with ONE thread per row the
//     four warps of the CTA ran at one warp per scheduler and the epilogue's dependent-issue latency was the whole kernel):
//     tcgen05.ld hands a thread its piece of the accumulator row, it applies bias / LeakyReLU / BatchNorm, splits the result
//     into the next layer's fp16 pair and writes it back into the tile in place;
//   * pass A (enc, dec) runs on the x tile; pass B runs the encoder on the x tile AND the xhat tile with the same weights
//     (reconstruction_aggregation.py:25-27 does exactly that), so the diff d_l = enc_l(xhat) - enc_l(x) is formed from two fp32
//     accumulator rows in registers -- no stash -- and its squares are summed per row by the thread that owns the row.
// Arithmetic identical to the per-layer tensor-core kernels: three MMAs hi*lo + lo*hi + hi*hi per k-step into one fp32
// accumulator, accumulator compensation (mmad_set_option "acc_comp"), fp32 epilogue.
#include "gemm_tc_common.cuh"

namespace mmad {

using namespace tc;

namespace {

constexpr int SNT_THREADS = 512;            // 16 warps: row = 32 * (warp % 4) + lane (the TMEM lanes a warp may read), column block = warp / 4
constexpr int SNT_MAX_LAYERS = 8;
constexpr int SNT_TILE = 16384;                 // one [128 rows x 64 halfs] swizzled box
constexpr int SNT_ACT = 4 * SNT_TILE;           // hi kb0, hi kb1, lo kb0, lo kb1
constexpr int SNT_SMEM = 3 * SNT_ACT + 4096 + 1024;      // x tile, xhat tile, weights + vectors / row partials / barriers (3.6 KB) + alignment

struct alignas(64) SntLayer {
    CUtensorMap wh, wl;          // hi / lo twins of 256 W, [N rows, Kp] K-major, box 64 halfs x 128 rows
    const float* bias; const float* scale; const float* shift;      // scale == nullptr: bare Linear
    int K, N, num_kb;
    float acc_mul;               // (1 / 256) * (1 + acc_comp * k-blocks)
};

struct SntParams {
    SntLayer enc[SNT_MAX_LAYERS], dec[SNT_MAX_LAYERS];
    int n_enc, n_dec, D, lo, hi, last;
    float inv_base, inv_sap, slope;
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// byte offset of 16-byte chunk `ch` (8 halfs) of row `row` inside an activation tile half (hi or lo)
__device__ __forceinline__ uint32_t snt_chunk_off(int row, int ch) {
    return (uint32_t)((ch >> 3) * SNT_TILE + row * 128 + (((ch & 7) ^ (row & 7)) << 4));
}

// eight values -> hi / lo halves, written as one 16-byte chunk each
__device__ __forceinline__ void snt_store8(uint8_t* tile, int row, int ch, const float (&y)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(y[2 * i] - hf.x, y[2 * i + 1] - hf.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    const uint32_t off = snt_chunk_off(row, ch);
    *reinterpret_cast<uint4*>(tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(tile + 2 * SNT_TILE + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(SNT_THREADS, 1)
smallnet_tc_kernel(const __grid_constant__ SntParams P, const float* __restrict__ x, int ldx, int n, float* __restrict__ base_out,
                   float* __restrict__ sap_out) {
    extern __shared__ uint8_t snt_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(snt_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* tP = smem;                         // x tile, then the activations of pass A in place, finally xhat
    uint8_t* tQ = smem + SNT_ACT;               // x tile of pass B
    uint8_t* tW = smem + 2 * SNT_ACT;           // weights of the current layer: Wh kb0, Wh kb1, Wl kb0, Wl kb1
    float* s_vec = reinterpret_cast<float*>(smem + 3 * SNT_ACT);         // [3][128] bias, scale, shift of the current layer
    float* s_sq = s_vec + 3 * 128;                                       // [4][128] row partial sums of d^2 per column block
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(s_sq + 4 * 128);       // weights landed
    uint64_t* bar_mma = bar_w + 1;                                       // this step's MMAs completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = (warp & 3) * 32 + (tid & 31);       // this thread's row of the tile == its TMEM lane
    const int q = warp >> 2;                            // its block of 32 columns

    if (tid == 0) {
        mbar_init(smem_u32(bar_w), 1);
        mbar_init(smem_u32(bar_mma), 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);     // this warp's 32 TMEM lanes

    const int L = P.n_enc, Ld = P.n_dec, D = P.D;
    const int steps_a = L + Ld, n_steps = steps_a + P.last;
    auto layer_of = [&](int s) -> const SntLayer& { return s < L ? P.enc[s] : (s < steps_a ? P.dec[s - L] : P.enc[s - steps_a]); };
    auto fetch_weights = [&](int s) {           // thread 0: TMA of step s's weight twins into tW
        const SntLayer& Ly = layer_of(s);
        const uint32_t fb = smem_u32(bar_w);
        mbar_expect_tx(fb, (uint32_t)(2 * Ly.num_kb * SNT_TILE));
        for (int kb = 0; kb < Ly.num_kb; ++kb) {
            tma_load_2d(smem_u32(tW + kb * SNT_TILE), &Ly.wh, fb, kb * 64, 0);
            tma_load_2d(smem_u32(tW + 2 * SNT_TILE + kb * SNT_TILE), &Ly.wl, fb, kb * 64, 0);
        }
    };
    // x rows of the tile -> fp16 hi / lo twins in the swizzled layout (coalesced 32-byte reads, zero padding to 64 / 128 columns)
    auto load_x = [&](uint8_t* tile, int r0) {
        const int nch = ((D + 63) / 64) * 8;        // 16-byte chunks per row
        for (int i = tid; i < 128 * nch; i += SNT_THREADS) {
            const int row = i / nch, ch = i - row * nch;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
            if (r0 + row < n) {
                const float* src = x + (size_t)(r0 + row) * ldx + ch * 8;
                if (ch * 8 + 7 < D) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (ch * 8 + j < D) v[j] = __ldg(src + j);
                }
            }
            snt_store8(tile, row, ch, v);
        }
    };

    uint32_t g = 0;                              // global step counter: phase of both barriers
    const int tiles = (n + 127) / 128;
    if (tid == 0 && (int)blockIdx.x < tiles) fetch_weights(0);
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int r0 = t * 128;
        load_x(tP, r0);
        fence_proxy_async();
        __syncthreads();
        float base_sum = 0.f, sap_sum = 0.f;
        for (int s = 0; s < n_steps; ++s, ++g) {
            const SntLayer& Ly = layer_of(s);
            const bool pass_b = s >= steps_a;
            const int l_b = s - steps_a + 1;                 // pass B: index of the diff this step produces
            if (s == steps_a) {                              // pass B starts: x again (L2-hot) next to xhat
                load_x(tQ, r0);
                fence_proxy_async();
            }
            // this layer's epilogue vectors
            if (tid < 128) {
                const bool ok = tid < Ly.N;
                s_vec[tid] = ok ? __ldg(Ly.bias + tid) : 0.f;
                s_vec[128 + tid] = (ok && Ly.scale) ? __ldg(Ly.scale + tid) : 0.f;
                s_vec[256 + tid] = (ok && Ly.scale) ? __ldg(Ly.shift + tid) : 0.f;
            }
            __syncthreads();
            const int n_eff = (Ly.N + 15) & ~15;
            if (tid == 0) {
                mbar_wait(smem_u32(bar_w), g & 1);
                tc_fence_after();
                const uint32_t idesc = make_idesc(n_eff, false, false);
                for (int a = 0; a < (pass_b ? 2 : 1); ++a) {
                    // pass A: P -> acc0.  pass B: Q (x path) -> acc0, P (xhat path) -> acc1
                    const uint8_t* src = pass_b ? (a == 0 ? tQ : tP) : tP;
                    const uint32_t d_tmem = tmem_base + a * 128;
                    for (int kb = 0; kb < Ly.num_kb; ++kb) {
                        const uint64_t dAh = make_smem_desc<false>(smem_u32(src + kb * SNT_TILE));
                        const uint64_t dAl = make_smem_desc<false>(smem_u32(src + 2 * SNT_TILE + kb * SNT_TILE));
                        const uint64_t dBh = make_smem_desc<false>(smem_u32(tW + kb * SNT_TILE));
                        const uint64_t dBl = make_smem_desc<false>(smem_u32(tW + 2 * SNT_TILE + kb * SNT_TILE));
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                            umma_f16(d_tmem, dAh + adv, dBl + adv, idesc, (kb | k) != 0);
                            umma_f16(d_tmem, dAl + adv, dBh + adv, idesc, 1);
                            umma_f16(d_tmem, dAh + adv, dBh + adv, idesc, 1);
                        }
                    }
                }
                umma_commit(smem_u32(bar_mma));
            }
            mbar_wait(smem_u32(bar_mma), g & 1);
            tc_fence_after();
            // the weight buffer is free: the next step's weights (or the next tile's first layer) travel during the epilogue
            if (tid == 0) {
                if (s + 1 < n_steps) fetch_weights(s + 1);
                else if (t + (int)gridDim.x < tiles) fetch_weights(0);
            }
            // ---- epilogue: thread (row, q) owns columns [32 q, 32 q + 32) of its row ----
            const int kp_next = ((Ly.N + 63) / 64) * 64;
            const bool write_p = pass_b ? (l_b < P.last) : true;         // pass B's last layer feeds nothing
            const bool is_xhat = !pass_b && s == steps_a - 1;
            const int c0 = q * 32;
            float sq = 0.f;
            if (c0 < kp_next) {
                uint32_t v0[32], v1[32];
                if (c0 < n_eff) {
                    tmem_ld32(taddr + c0, v0);
                    if (pass_b) tmem_ld32(taddr + 128 + c0, v1);
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int c = c0 + i;
                    float a0 = 0.f, a1 = 0.f;
                    if (c < Ly.N) {
                        const float b = s_vec[c], sc = s_vec[128 + c], sh = s_vec[256 + c];
                        a0 = fmaf(__uint_as_float(v0[i]), Ly.acc_mul, b);
                        if (Ly.scale) { a0 = a0 > 0.f ? a0 : a0 * P.slope; a0 = fmaf(a0, sc, sh); }
                        if (pass_b) {
                            a1 = fmaf(__uint_as_float(v1[i]), Ly.acc_mul, b);
                            if (Ly.scale) { a1 = a1 > 0.f ? a1 : a1 * P.slope; a1 = fmaf(a1, sc, sh); }
                            const float d = a1 - a0;
                            sq = fmaf(d, d, sq);
                        }
                    }
                    v0[i] = __float_as_uint(a0); v1[i] = __float_as_uint(a1);
                }
                if (is_xhat && r0 + row < n) {           // d_0 = xhat - x against the fp32 input
                    const float* xr = x + (size_t)(r0 + row) * ldx;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int c = c0 + i;
                        if (c < D) { const float d = __uint_as_float(v0[i]) - __ldg(xr + c); sq = fmaf(d, d, sq); }
                    }
                }
                if (write_p) {
#pragma unroll
                    for (int k8 = 0; k8 < 4; ++k8) {
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(v0[k8 * 8 + j]);
                        snt_store8(pass_b ? tQ : tP, row, (c0 >> 3) + k8, o);
                        if (pass_b) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(v1[k8 * 8 + j]);
                            snt_store8(tP, row, (c0 >> 3) + k8, o);
                        }
                    }
                }
            }
            s_sq[q * 128 + row] = sq;
            tc_fence_before();
            fence_proxy_async();
            __syncthreads();
            if (q == 0) {                                // the row's four column blocks, fixed order
                const float t4 = ((s_sq[row] + s_sq[128 + row]) + s_sq[256 + row]) + s_sq[384 + row];
                if (is_xhat) { base_sum = t4; if (P.lo == 0) sap_sum += t4; }
                if (pass_b && l_b >= P.lo && l_b < P.hi) sap_sum += t4;
            }
        }
        if (q == 0 && r0 + row < n) {
            if (base_out) base_out[r0 + row] = base_sum * P.inv_base;
            if (sap_out) sap_out[r0 + row] = sap_sum * P.inv_sap;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
    }
}

bool g_snt_attr = false;

}  // namespace

// base / SAP scores of n rows for a model whose widths are all <= 128, F16X3 arithmetic
int smallnet_tc_score(mmad_t h, const float* d_x, int ldx, int n, int lo, int hi, float* d_base, float* d_sap, float acc_comp,
                      cudaStream_t s) {
    if (n <= 0) return MMAD_OK;
    const mmad_desc_t* d = handle_desc(h);
    if (d->n_enc > SNT_MAX_LAYERS || d->n_dec > SNT_MAX_LAYERS) { set_error("small-net tensor-core kernel: more than 8 layers per module"); return MMAD_E_UNSUPPORTED; }
    SntParams P;
    memset(&P, 0, sizeof P);
    P.n_enc = d->n_enc; P.n_dec = d->n_dec; P.D = d->enc_widths[0]; P.lo = lo; P.hi = hi;
    P.last = hi > 1 ? std::min(d->n_enc, hi - 1) : 0;
    P.slope = d->lrelu_slope;
    P.inv_base = 1.f / P.D;
    int dsel = 0;
    for (int l = lo; l < hi; ++l) dsel += d->enc_widths[l];
    P.inv_sap = 1.f / dsel;
    for (int m = 0; m < 2; ++m)
        for (int i = 0; i < (m ? d->n_dec : d->n_enc); ++i) {
            SntLayer& Ly = (m ? P.dec : P.enc)[i];
            const LayerF32 f = handle_layer_f32(h, m, i);
            float wscale = 1.f;
            if (handle_layer_tcmaps(h, m, i, &Ly.wh, &Ly.wl, &wscale)) { set_error("small-net tensor-core kernel: layer maps missing"); return MMAD_E_STATE; }
            Ly.bias = f.bias; Ly.scale = f.has_bn ? f.scale : nullptr; Ly.shift = f.has_bn ? f.shift : nullptr;
            Ly.K = f.K; Ly.N = f.N; Ly.num_kb = f.Kp / 64;
            Ly.acc_mul = (1.f / wscale) * (1.f + acc_comp * 12.f * (float)Ly.num_kb);
        }
    if (!g_snt_attr) {
        MMAD_CUDA_OK(cudaFuncSetAttribute(smallnet_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SNT_SMEM));
        g_snt_attr = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles = (n + 127) / 128;
    smallnet_tc_kernel<<<std::min(tiles, sms), SNT_THREADS, SNT_SMEM, s>>>(P, d_x, ldx, n, d_base, d_sap);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

}  // namespace mmad
