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
constexpr int SNT_SMEM = 3 * SNT_ACT + 6144 + 1024;      // x tile, xhat tile, weights + vectors (3 KB) / row partials (2 KB) / barriers + alignment

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
    unsigned long long* dbg;     // MMAD_SNT_DEBUG=1: %globaltimer stamps of CTA 0's first tile, 6 per step
};

__device__ __forceinline__ unsigned long long snt_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// byte offset of 16-byte chunk `ch` (8 halfs) of row `row` inside an activation tile half (hi or lo)
__device__ __forceinline__ uint32_t snt_chunk_off(int row, int ch) {
    return (uint32_t)((ch >> 3) * SNT_TILE + row * 128 + (((ch & 7) ^ (row & 7)) << 4));
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
    return r;
}

// eight values -> hi / lo halves, written as one 16-byte chunk each (`tile` is a shared-space address)
__device__ __forceinline__ void snt_store8(uint32_t tile, int row, int ch, const float (&y)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(y[2 * i] - hf.x, y[2 * i + 1] - hf.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    const uint32_t off = tile + snt_chunk_off(row, ch);
    sts128(off, h[0], h[1], h[2], h[3]);
    sts128(off + 2 * SNT_TILE, l[0], l[1], l[2], l[3]);
}

// the eight values of a chunk back as hi + lo
__device__ __forceinline__ void snt_load8(uint32_t tile, int row, int ch, float (&y)[8]) {
    const uint32_t off = tile + snt_chunk_off(row, ch);
    const uint4 h = lds128u(off), l = lds128u(off + 2 * SNT_TILE);
    const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hh[i]));
        const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&ll[i]));
        y[2 * i] = a.x + c.x; y[2 * i + 1] = a.y + c.y;
    }
}

// bias / LeakyReLU / BatchNorm of 32 accumulator columns in place; `sv` = shared address of the layer's vectors at column c0.
// Columns past the layer's width come out as exact zeros without a test: their weights, bias, scale and shift are all zero.
template <bool BN>
__device__ __forceinline__ void snt_transform(uint32_t (&v)[32], uint32_t sv, float mul, float slope) {
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
        const float4 b = lds128(sv + i4 * 16);
        float a[4] = {fmaf(__uint_as_float(v[4 * i4]), mul, b.x), fmaf(__uint_as_float(v[4 * i4 + 1]), mul, b.y),
                      fmaf(__uint_as_float(v[4 * i4 + 2]), mul, b.z), fmaf(__uint_as_float(v[4 * i4 + 3]), mul, b.w)};
        if (BN) {
            const float4 sc = lds128(sv + 512 + i4 * 16), sh = lds128(sv + 1024 + i4 * 16);
            const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a[j] = a[j] > 0.f ? a[j] : a[j] * slope;
                a[j] = fmaf(a[j], scv[j], shv[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) v[4 * i4 + j] = __float_as_uint(a[j]);
    }
}

__device__ __forceinline__ void snt_store32(uint32_t tile, int row, int c0, const uint32_t (&v)[32]) {
#pragma unroll
    for (int k8 = 0; k8 < 4; ++k8) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(v[k8 * 8 + j]);
        snt_store8(tile, row, (c0 >> 3) + k8, o);
    }
}

__global__ void __launch_bounds__(SNT_THREADS, 1)
smallnet_tc_kernel(const __grid_constant__ SntParams P, const float* __restrict__ x, int ldx, int n, float* __restrict__ base_out,
                   float* __restrict__ sap_out) {
    extern __shared__ uint8_t snt_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(snt_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = smem_u32(smem);
    const uint32_t tP = sb;                       // x tile, then the activations of pass A in place, finally xhat
    const uint32_t tQ = sb + SNT_ACT;             // x again: compared with xhat, then the x path of pass B in place
    const uint32_t tW = sb + 2 * SNT_ACT;         // weights of the current layer: Wh kb0, Wh kb1, Wl kb0, Wl kb1
    const uint32_t s_vec = sb + 3 * SNT_ACT;      // [2 buffers][bias, scale, shift][128]: this step's and the next step's vectors
    float* s_sq = reinterpret_cast<float*>(smem + 3 * SNT_ACT + 2 * 3 * 128 * 4);      // [4][128] row partial sums of d^2 per column block
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_sq + 4 * 128);
    const uint32_t bar_w = smem_u32(bars);        // weights landed
    const uint32_t bar_mma = smem_u32(bars + 1);  // accumulator 0 complete
    const uint32_t bar_mma2 = smem_u32(bars + 2); // pass B: accumulator 1 complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = (warp & 3) * 32 + (tid & 31);       // this thread's row of the tile == its TMEM lane
    const int q = warp >> 2;                            // its block of 32 columns
    const int c0 = q * 32;

    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        mbar_init(bar_mma2, 1);
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
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + c0;     // this warp's 32 TMEM lanes, this thread's columns

    const int L = P.n_enc, Ld = P.n_dec, D = P.D;
    const int steps_a = L + Ld, n_steps = steps_a + P.last;
    const float slope = P.slope;
    auto layer_of = [&](int s) -> const SntLayer& { return s < L ? P.enc[s] : (s < steps_a ? P.dec[s - L] : P.enc[s - steps_a]); };
    auto fetch_weights = [&](int s) {           // thread 0: TMA of step s's weight twins into tW
        const SntLayer& Ly = layer_of(s);
        const int nkb = Ly.num_kb;
        mbar_expect_tx(bar_w, (uint32_t)(2 * nkb * SNT_TILE));
        for (int kb = 0; kb < nkb; ++kb) {
            tma_load_2d(tW + kb * SNT_TILE, &Ly.wh, bar_w, kb * 64, 0);
            tma_load_2d(tW + 2 * SNT_TILE + kb * SNT_TILE, &Ly.wl, bar_w, kb * 64, 0);
        }
    };
    // threads 128..255: the epilogue vectors of step s into buffer `buf` (zero past the layer's width)
    auto fetch_vectors = [&](int s, int buf) {
        const SntLayer& Ly = layer_of(s);
        const int c = tid - 128;
        const bool ok = c < Ly.N;
        const float* sc = Ly.scale;
        const float b = ok ? __ldg(Ly.bias + c) : 0.f;
        const float a = (ok && sc) ? __ldg(sc + c) : 0.f;
        const float h = (ok && sc) ? __ldg(Ly.shift + c) : 0.f;
        const uint32_t dst = s_vec + buf * 1536 + c * 4;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst), "f"(b) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + 512), "f"(a) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + 1024), "f"(h) : "memory");
    };
    // x rows of the tile -> fp16 hi / lo twins in the swizzled layout, into BOTH tiles (coalesced 32-byte reads, zero padding
    // to 64 / 128 columns)
    auto load_x = [&](int r0) {
        const int nch = ((D + 63) / 64) * 8;        // 16-byte chunks per row
        for (int i = tid; i < 128 * nch; i += SNT_THREADS) {
            const int r = i / nch, ch = i - r * nch;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
            if (r0 + r < n) {
                const float* src = x + (size_t)(r0 + r) * ldx + ch * 8;
                if (ch * 8 + 7 < D) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (ch * 8 + j < D) v[j] = __ldg(src + j);
                }
            }
            snt_store8(tP, r, ch, v);
            snt_store8(tQ, r, ch, v);
        }
    };

    uint32_t g = 0, gb = 0;                      // step counter (phase of bar_w / bar_mma), pass-B step counter (phase of bar_mma2)
    const int tiles = (n + 127) / 128;
    if ((int)blockIdx.x < tiles) {
        if (tid == 0) fetch_weights(0);
        if (warp >= 4 && warp < 8) fetch_vectors(0, 0);
    }
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int r0 = t * 128;
        load_x(r0);
        fence_proxy_async();
        __syncthreads();
        float base_sum = 0.f, sap_sum = 0.f;
        for (int s = 0; s < n_steps; ++s, ++g) {
            const SntLayer& Ly = layer_of(s);
            const int N = Ly.N, num_kb = Ly.num_kb;
            const float mul = Ly.acc_mul;
            const bool bn = Ly.scale != nullptr;
            const bool stamp = P.dbg && tid == 0 && t == 0;
            if (stamp) P.dbg[6 * s] = snt_now();
            const bool pass_b = s >= steps_a;
            const int l_b = s - steps_a + 1;                 // pass B: index of the diff this step produces
            const int n_eff = (N + 31) & ~31;                // MMA width: whole 32-column blocks (weight rows past N are zero)
            if (tid == 0) {
                mbar_wait(bar_w, g & 1);
                tc_fence_after();
                const uint32_t idesc = make_idesc(n_eff, false, false);
                for (int a = 0; a < (pass_b ? 2 : 1); ++a) {
                    // pass A: P -> acc0.  pass B: Q (x path) -> acc0, then P (xhat path) -> acc1, each with its own barrier so
                    // that the epilogue of the x path runs under the MMAs of the xhat path
                    const uint32_t src = pass_b ? (a == 0 ? tQ : tP) : tP;
                    const uint32_t d_tmem = tmem_base + a * 128;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        const uint64_t dAh = make_smem_desc<false>(src + kb * SNT_TILE);
                        const uint64_t dAl = make_smem_desc<false>(src + 2 * SNT_TILE + kb * SNT_TILE);
                        const uint64_t dBh = make_smem_desc<false>(tW + kb * SNT_TILE);
                        const uint64_t dBl = make_smem_desc<false>(tW + 2 * SNT_TILE + kb * SNT_TILE);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                            umma_f16(d_tmem, dAh + adv, dBl + adv, idesc, (kb | k) != 0);
                            umma_f16(d_tmem, dAl + adv, dBh + adv, idesc, 1);
                            umma_f16(d_tmem, dAh + adv, dBh + adv, idesc, 1);
                        }
                    }
                    umma_commit(a == 0 ? bar_mma : bar_mma2);
                }
                if (stamp) P.dbg[6 * s + 1] = snt_now();
            }
            // the NEXT step's vectors travel under this step's MMAs (the other buffer was last read one step ago)
            if (warp >= 4 && warp < 8) {
                if (s + 1 < n_steps) fetch_vectors(s + 1, (g + 1) & 1);
                else if (t + (int)gridDim.x < tiles) fetch_vectors(0, (g + 1) & 1);
            }
            const uint32_t sv = s_vec + (g & 1) * 1536 + c0 * 4;
            const int kp_next = ((N + 63) / 64) * 64;
            const bool write_p = pass_b ? (l_b < P.last) : true;         // pass B's last layer feeds nothing
            const bool is_xhat = !pass_b && s == steps_a - 1;
            const bool active = c0 < kp_next;                            // warp-uniform
            float sq = 0.f;
            uint32_t v0[32];
            // ---- epilogue, accumulator 0: thread (row, q) owns columns [32 q, 32 q + 32) of its row ----
            mbar_wait(bar_mma, g & 1);
            tc_fence_after();
            if (stamp) P.dbg[6 * s + 2] = snt_now();
            if (!pass_b && tid == 0) {           // the weight buffer is free: the next weights travel during the epilogue
                if (s + 1 < n_steps) fetch_weights(s + 1);
                else if (t + (int)gridDim.x < tiles) fetch_weights(0);
            }
            if (active) {
                if (c0 < n_eff) tmem_ld32(taddr, v0);
                else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v0[i] = 0u;
                }
                if (bn) snt_transform<true>(v0, sv, mul, slope); else snt_transform<false>(v0, sv, mul, slope);
                if (is_xhat) {                   // d_0 = xhat - x (x as the hi + lo pair the first layer multiplied)
#pragma unroll
                    for (int k8 = 0; k8 < 4; ++k8) {
                        float xv[8];
                        snt_load8(tQ, row, (c0 >> 3) + k8, xv);
#pragma unroll
                        for (int j = 0; j < 8; ++j) { const float d = __uint_as_float(v0[k8 * 8 + j]) - xv[j]; sq = fmaf(d, d, sq); }
                    }
                }
                if (write_p) snt_store32(pass_b ? tQ : tP, row, c0, v0);
            }
            if (stamp) P.dbg[6 * s + 3] = snt_now();
            // ---- pass B, accumulator 1: the xhat path and the diff ----
            if (pass_b) {
                mbar_wait(bar_mma2, gb & 1);
                tc_fence_after();
                if (tid == 0) {
                    if (s + 1 < n_steps) fetch_weights(s + 1);
                    else if (t + (int)gridDim.x < tiles) fetch_weights(0);
                }
                if (active) {
                    uint32_t v1[32];
                    if (c0 < n_eff) tmem_ld32(taddr + 128, v1);
                    else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v1[i] = 0u;
                    }
                    if (bn) snt_transform<true>(v1, sv, mul, slope); else snt_transform<false>(v1, sv, mul, slope);
#pragma unroll
                    for (int i = 0; i < 32; ++i) { const float d = __uint_as_float(v1[i]) - __uint_as_float(v0[i]); sq = fmaf(d, d, sq); }
                    if (write_p) snt_store32(tP, row, c0, v1);
                }
                ++gb;
            }
            s_sq[q * 128 + row] = sq;
            if (stamp) P.dbg[6 * s + 4] = snt_now();
            tc_fence_before();
            fence_proxy_async();
            __syncthreads();
            if (stamp) P.dbg[6 * s + 5] = snt_now();
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
    static const bool debug = getenv("MMAD_SNT_DEBUG") != nullptr;
    static int debug_left = 3;
    const bool dbg = debug && debug_left > 0;
    if (dbg) {
        MMAD_CUDA_OK(cudaMallocManaged(&P.dbg, 6 * 32 * sizeof(unsigned long long)));
        memset(P.dbg, 0, 6 * 32 * sizeof(unsigned long long));
    }
    smallnet_tc_kernel<<<std::min(tiles, sms), SNT_THREADS, SNT_SMEM, s>>>(P, d_x, ldx, n, d_base, d_sap);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    if (dbg) {
        MMAD_CUDA_OK(cudaStreamSynchronize(s));
        const int steps = P.n_enc + P.n_dec + P.last;
        fprintf(stderr, "smallnet_tc D=%d n=%d: per step ns [mma issue, acc0 wait, epilogue 0, epilogue 1, end sync]\n", P.D, n);
        for (int i = 0; i < steps; ++i) {
            const unsigned long long* q = P.dbg + 6 * i;
            fprintf(stderr, "  step %2d: %5llu %5llu %5llu %5llu %5llu   total %llu\n", i, q[1] - q[0], q[2] - q[1], q[3] - q[2], q[4] - q[3],
                    q[5] - q[4], (i + 1 < steps ? q[6] : q[5]) - q[0]);
        }
        fprintf(stderr, "  tile total %llu ns\n", P.dbg[6 * (steps - 1) + 5] - P.dbg[0]);
        cudaFree(P.dbg);
        --debug_left;
    }
    return MMAD_OK;
}

}  // namespace mmad
