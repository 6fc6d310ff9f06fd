// Realtime scoring in ONE launch (test_file/realtime_tester.py:291-309: batch_size = 10 windows per call, base + SAP,
// nap=False): the whole 15-step chain  enc(x) -> dec -> enc(xhat) + diffs + score reduction  runs inside one cooperative
// kernel, exact fp32 FMA arithmetic, for calls of <= 64 windows.
//
// Why not the per-layer kernels: a call touches 61 MB of fp32 weights (L2-resident between calls) and 15 dependent
// layers; fifteen launches (graph replay) + one H2D + one D2H cost 127 us at one window, of which ~10 us is math.  Here:
//   * one CTA per SM (148 x 512 threads), launched cooperatively; layers are separated by a device-wide barrier
//     (one atomic arrive + acquire poll per CTA, ~1 us), not by kernel boundaries;
//   * every CTA owns ceil(N / grid) output columns of a layer (10 of 1402, 12 of 1728 ...) over the full K extent, so all SMs
//     pull the layer's weights from L2 at once; its weight slice for layer l+1 is prefetched into shared memory with
//     cp.async BEFORE it waits at the barrier of layer l (weights do not depend on the previous layer);
//   * the input comes straight from the caller's pinned, device-mapped buffer (staged once by the grid), the two score
//     vectors go straight back to mapped host memory followed by a sequence flag the host spins on -- the call is a
//     doorbell, not H2D + graph + D2H + stream synchronise.
// Inside a CTA the 16 warps form 4 column groups x 4 k-quarters; a warp accumulates NB rows x 4 columns from shared memory
// (activations re-used across its columns), lanes split k; partial sums meet in shared memory in a fixed order, so
// results are deterministic.  Diffs are taken against the stashed enc(x) activations in the epilogue; their squares are
// summed per CTA and row, and the last CTA to finish adds the per-CTA partials in a fixed order.
#include <chrono>

#include "mmad_internal.cuh"

namespace mmad {

namespace {

constexpr int ST_THREADS = 512;
constexpr int ST_CW = 4;              // columns per warp
constexpr int ST_CG = 4;              // column groups per CTA
constexpr int ST_KQ = 4;              // k quarters per CTA
constexpr int ST_CPC = ST_CW * ST_CG; // columns per CTA and layer (upper bound)
constexpr int ST_MAX_ROWS = 64;
constexpr int ST_MAX_STEPS = 3 * MMAD_MAX_LAYERS;
constexpr int ST_SMEM_CAP = 227 * 1024 - 1024;

struct StStep {
    const float* W; const float* bias; const float* scale; const float* shift;   // scale == nullptr: bare Linear
    int ibuf, obuf, rbuf;   // input / output / diff-reference buffer: -2 none, -1 the staged input x, >= 0 activation buffer
    int K;             // contraction length (valid input columns)
    int K4;            // padded contraction length / 4 (row stride of W in float4, of the shared activation tile)
    int N;
    int cpc;           // columns per CTA = ceil(N / grid)
    int diff;          // partial-sum slot of the diff this step produces, -1: none
};

struct StPlan {
    int n_steps, n_diffs, lo, hi, D, ldx, grid;
    float inv_base, inv_sap, slope;
    // every step output exists twice: as (value, sequence) pairs for calls of <= 4 rows and as plain floats (layers separated
    // by grid barriers) for taller calls; buffer i starts at base + i * bufsz elements, rows are ld (x: ldx) elements apart
    uint2* xp; uint2* pbase;
    float* xf; float* fbase;
    size_t bufsz;
    int ld;
    uint2* partial;    // [n_diffs][grid][ST_MAX_ROWS] per-CTA row sums of d^2 as (value, sequence) pairs
    int wmax4;         // largest weight slice of a CTA (float4): the pair protocol keeps TWO slices in flight
    int nact[MMAD_MAX_LAYERS + 2];   // CTAs that own columns of diff l
    StStep step[ST_MAX_STEPS];
};

__device__ __forceinline__ void st_cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void st_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void st_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// 8-byte (value, sequence) pair: single-copy atomic, so a reader that sees the call's sequence number sees the value
__device__ __forceinline__ uint2 ld_pair(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_pair(uint2* p, float value, uint32_t seq) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(value)), "r"(seq) : "memory");
}

// barrier protocol (taller calls): plain float activations, layers separated by a device-wide barrier of the cooperatively
// launched grid -- a monotonically increasing arrival counter (never reset; `target` carries the call's base), one
// release-add and one polling thread per CTA.  (One flag per producer CTA polled by every consumer was measured slower:
// 148 x 148 acquire polls crowd the L2 -- 205 us against 167 us at 10 rows.)
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar), "l"(1ULL) : "memory");
        const long long t0 = clock64();
        while (ld_acquire(bar) < target) {
            if (clock64() - t0 > 2000000000LL) {
                printf("mmad stream kernel: grid barrier timed out (block %d, target %llu)\n", blockIdx.x, target);
                __trap();
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}

// Layers are chained WITHOUT grid barriers: every activation travels as an 8-byte (value, call sequence number) pair -- the
// low-latency protocol of collective libraries.  A consumer polls the pairs it needs until they carry this call's number:
// per layer one store becomes visible and one load returns (~1.5 us), where an arrive-and-poll barrier costs a fence, an
// atomic and at least one more round trip before the activations can even be requested (3.0-5.1 us measured per layer).
// All polls are bounded: a protocol bug traps instead of hanging the device.
template <int NB, bool LL>
__global__ void __launch_bounds__(ST_THREADS, 1)
stream_chain_kernel(const StPlan* __restrict__ P, const float* __restrict__ x_src, int rows, uint2* out_host,
                    unsigned long long seq, unsigned long long* bar, unsigned long long bar2_base, unsigned long long* dbg) {
    extern __shared__ __align__(16) float st_smem[];
    __shared__ float s_red[ST_KQ][NB][ST_CPC];
    __shared__ float s_sq[NB][ST_CPC];
    __shared__ float s_vec[3][ST_CPC];        // bias, BN scale, BN shift of this CTA's columns
    __shared__ float s_fin[MMAD_MAX_LAYERS + 2][ST_MAX_ROWS];
    auto stamp = [&](int i) {                 // MMAD_STREAM_DEBUG: CTA 0's wall clock at the phase boundaries
        if (dbg && blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            dbg[i] = t;
        }
    };
    stamp(0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, grid = gridDim.x;
    const int n_steps = P->n_steps;
    const float slope = P->slope;
    const uint32_t seq32 = (uint32_t)seq;
    unsigned long long arrivals = bar2_base;      // barrier protocol: *bar counts the arrivals of every barrier of every call
    // pair protocol: two weight slices in flight (slice s in buffer s & 1, the activation tile behind both); barrier protocol:
    // one slice, the activation tile right behind it
    const int wmax4 = P->wmax4;

    auto prefetch_weights = [&](int s) -> uint32_t {    // cp.async of this CTA's weight slice; returns its float4 extent
        const StStep& st = P->step[s];
        const int c_lo = cta * st.cpc;
        int ncols = st.N - c_lo; if (ncols > st.cpc) ncols = st.cpc; if (ncols < 0) ncols = 0;
        const int n4 = ncols * st.K4;
        const float4* src = reinterpret_cast<const float4*>(st.W) + (size_t)c_lo * st.K4;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(st_smem) + (LL ? (uint32_t)(s & 1) * (uint32_t)wmax4 * 16u : 0u);
        for (int i = tid; i < n4; i += ST_THREADS) st_cp16(dst + i * 16, src + i);
        st_cp_commit();
        return (uint32_t)(st.cpc * st.K4);
    };
    auto prefetch_vec = [&](int s) -> float {           // bias / BN scale / BN shift of this CTA's columns, one per thread
        if (tid >= 3 * ST_CPC) return 0.f;
        const StStep& st = P->step[s];
        const int which = tid / ST_CPC, cl = tid % ST_CPC;
        const int c = cta * st.cpc + cl;
        if (cl >= st.cpc || c >= st.N) return 0.f;
        const float* v = which == 0 ? st.bias : (which == 1 ? st.scale : st.shift);
        return v ? __ldg(v + c) : 0.f;
    };
    uint32_t w4 = prefetch_weights(0);
    if (LL && n_steps > 1) prefetch_weights(1);
    float vpre = prefetch_vec(0);

    // ---- stage the input rows from the caller's mapped host buffer (once, by the whole grid) ----
    if constexpr (LL) {
        const int D = P->D, ldx = P->ldx;
        const int total = rows * D;
        for (int i = cta * ST_THREADS + tid; i < total; i += grid * ST_THREADS) {
            const int r = i / D, c = i - r * D;
            st_pair(P->xp + (size_t)r * ldx + c, __ldcv(x_src + i), seq32);     // host memory rewritten between calls: uncached load
        }
    } else {
        const int D4 = P->D >> 2, ld4 = P->ldx >> 2;
        const int total = rows * D4;
        float4* xd = reinterpret_cast<float4*>(P->xf);
        const float4* xs = reinterpret_cast<const float4*>(x_src);
        for (int i = cta * ST_THREADS + tid; i < total; i += grid * ST_THREADS) {
            const int r = i / D4, c = i - r * D4;
            xd[(size_t)r * ld4 + c] = __ldcv(xs + i);
        }
        arrivals += grid;
        grid_barrier(bar, arrivals);
    }
    stamp(1);

    for (int s = 0; s < n_steps; ++s) {
        const StStep st = P->step[s];
        const int K4 = st.K4, Kp = 4 * st.K4;
        const int c_lo = cta * st.cpc;
        int ncols = st.N - c_lo; if (ncols > st.cpc) ncols = st.cpc; if (ncols < 0) ncols = 0;
        if (tid < 3 * ST_CPC) s_vec[tid / ST_CPC][tid % ST_CPC] = vpre;      // visible after the __syncthreads below
        const float4* wsm = reinterpret_cast<const float4*>(st_smem) + (LL ? (size_t)(s & 1) * wmax4 : 0);
        float* asm_f = st_smem + 4 * (size_t)(LL ? 2 * wmax4 : w4);
        const float4* asm4 = reinterpret_cast<const float4*>(asm_f);
        const int cg = warp & (ST_CG - 1), kq = warp >> 2;
        const int cl0 = cg * ST_CW;
        int nj = ncols - cl0; if (nj > ST_CW) nj = ST_CW;
        const int q = K4 / ST_KQ;                                  // K4 is a multiple of 16
        const int k_lo = kq * q, k_hi = k_lo + q;
        for (int r0 = 0; r0 < rows && ncols > 0; r0 += NB) {
            int nb = rows - r0; if (nb > NB) nb = NB;
            // ---- activations of rows r0 .. r0 + nb into shared memory ----
            const long long t0 = clock64();
            if constexpr (LL) {
                // poll the (value, sequence) pairs; park the values
                const uint2* in = st.ibuf < 0 ? P->xp : P->pbase + (size_t)st.ibuf * P->bufsz;
                const int ldin = st.ibuf < 0 ? P->ldx : P->ld;
                for (int i = tid; i < nb * (Kp - st.K); i += ST_THREADS) {            // zero padding of the contraction dim
                    const int r = i / (Kp - st.K), k = st.K + i % (Kp - st.K);
                    asm_f[r * Kp + k] = 0.f;
                }
                const int total = nb * st.K;
                for (int base = 0; base < total; base += ST_THREADS * 8) {
                    unsigned pending = 0;
                    const uint2* src[8];
                    int dsti[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int i = base + j * ST_THREADS + tid;
                        if (i < total) {
                            const int r = i / st.K, k = i - r * st.K;
                            src[j] = in + (size_t)(r0 + r) * ldin + k;
                            dsti[j] = r * Kp + k;
                            pending |= 1u << j;
                        }
                    }
                    while (pending) {
                        uint2 v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) if (pending & (1u << j)) v[j] = ld_pair(src[j]);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if ((pending & (1u << j)) && v[j].y == seq32) { asm_f[dsti[j]] = __uint_as_float(v[j].x); pending &= ~(1u << j); }
                        if (pending && clock64() - t0 > 2000000000LL) {
                            printf("mmad stream kernel: activations of step %d never arrived (block %d thread %d)\n", s, cta, tid);
                            __trap();
                        }
                    }
                }
            } else {
                // the barrier behind the previous step made its outputs visible: bulk-copy the rows
                const float* in = st.ibuf < 0 ? P->xf : P->fbase + (size_t)st.ibuf * P->bufsz;
                const int ld4 = (st.ibuf < 0 ? P->ldx : P->ld) >> 2;
                const float4* src = reinterpret_cast<const float4*>(in) + (size_t)r0 * ld4;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(asm_f);
                for (int i = tid; i < nb * K4; i += ST_THREADS) {
                    const int r = i / K4, k4 = i - r * K4;
                    st_cp16(dst + i * 16, src + (size_t)r * ld4 + k4);
                }
                st_cp_commit();
            }
            // this CTA's weight slice: issued one step earlier (barrier protocol) / two steps earlier, with the next one still
            // allowed in flight (pair protocol)
            if (LL && s + 1 < n_steps) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else st_cp_wait_all();
            __syncthreads();
            if (nj > 0) {
                float acc[NB][ST_CW];
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int j = 0; j < ST_CW; ++j) acc[b][j] = 0.f;
                for (int k4 = k_lo + lane; k4 < k_hi; k4 += 32) {
                    float4 w[ST_CW];
#pragma unroll
                    for (int j = 0; j < ST_CW; ++j) w[j] = j < nj ? wsm[(cl0 + j) * K4 + k4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const float4 a = asm4[b * K4 + k4];
#pragma unroll
                        for (int j = 0; j < ST_CW; ++j) acc[b][j] = dot4(a, w[j], acc[b][j]);
                    }
                }
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int j = 0; j < ST_CW; ++j) {
                        float v = acc[b][j];
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 2);
                        v += __shfl_xor_sync(0xffffffffu, v, 1);
                        if (lane == 0) s_red[kq][b][cl0 + j] = v;
                    }
            }
            __syncthreads();
            // ---- epilogue: thread (b, cl) ----
            if (tid < NB * ST_CPC) {
                const int b = tid / ST_CPC, cl = tid % ST_CPC;
                float sq = 0.f;
                if (b < nb && cl < ncols) {
                    const int c = c_lo + cl, r = r0 + b;
                    float v = ((s_red[0][b][cl] + s_red[1][b][cl]) + s_red[2][b][cl]) + s_red[3][b][cl];
                    v += s_vec[0][cl];
                    if (st.scale) {
                        v = v > 0.f ? v : v * slope;
                        v = fmaf(v, s_vec[1][cl], s_vec[2][cl]);
                    }
                    if (st.obuf >= 0) {
                        if constexpr (LL) st_pair(P->pbase + (size_t)st.obuf * P->bufsz + (size_t)r * P->ld + c, v, seq32);
                        else P->fbase[(size_t)st.obuf * P->bufsz + (size_t)r * P->ld + c] = v;
                    }
                    if (st.rbuf > -2) {
                        // the reference is complete by now: a consumer of step s has seen every output of step s-1, recursively
                        const size_t ri = st.rbuf < 0 ? (size_t)r * P->ldx + c : (size_t)st.rbuf * P->bufsz + (size_t)r * P->ld + c;
                        float refv;
                        if constexpr (LL) refv = __uint_as_float(ld_pair((st.rbuf < 0 ? P->xp : P->pbase) + ri).x);
                        else refv = __ldcg((st.rbuf < 0 ? P->xf : P->fbase) + ri);
                        const float d = v - refv;
                        sq = d * d;
                    }
                }
                if (st.diff >= 0) s_sq[b][cl] = sq;
            }
            __syncthreads();
            if (st.diff >= 0 && tid < nb) {
                float t = 0.f;
                for (int cl = 0; cl < ncols; ++cl) t += s_sq[tid][cl];
                st_pair(P->partial + ((size_t)st.diff * grid + cta) * ST_MAX_ROWS + r0 + tid, t, seq32);
            }
        }
        if (s + 1 < n_steps) {
            __syncthreads();            // everybody is done reading the current slice and activation tile
            if constexpr (LL) {
                if (s + 2 < n_steps) prefetch_weights(s + 2);     // into the buffer this step just released
            } else {
                w4 = prefetch_weights(s + 1);      // travels while this CTA waits for the others
            }
            vpre = prefetch_vec(s + 1);
            if constexpr (!LL) {
                arrivals += grid;
                grid_barrier(bar, arrivals);
            }
        }
        stamp(2 + s);
    }

    // ---- CTA 0 collects the per-CTA partial sums -- (value, sequence) pairs again: no counter, no fence -- adds them in a fixed
    // order and writes the scores to the caller's mapped memory as pairs too: the host polls the score pairs themselves ----
    if (cta != 0) return;
    const int nd = P->n_diffs;
    const long long tf = clock64();
    for (int l = warp; l < nd; l += ST_THREADS / 32) {          // warp l: diff l, lanes split the CTAs (<= 8 loads in flight each)
        const int na = P->nact[l];
        for (int r = 0; r < rows; ++r) {
            float t[8];
            unsigned pending = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { t[j] = 0.f; if (lane + 32 * j < na) pending |= 1u << j; }
            while (pending) {
                uint2 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (pending & (1u << j)) v[j] = ld_pair(P->partial + ((size_t)l * grid + lane + 32 * j) * ST_MAX_ROWS + r);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if ((pending & (1u << j)) && v[j].y == seq32) { t[j] = __uint_as_float(v[j].x); pending &= ~(1u << j); }
                if (pending && clock64() - tf > 2000000000LL) {
                    printf("mmad stream kernel: partial sums of diff %d never arrived (lane %d)\n", l, lane);
                    __trap();
                }
            }
            float v = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if (lane == 0) s_fin[l][r] = v;
        }
    }
    __syncthreads();
    if (tid < rows) {
        float sap = 0.f;
        for (int l = P->lo; l < P->hi; ++l) sap += s_fin[l][tid];
        st_pair(out_host + tid, s_fin[0][tid] * P->inv_base, seq32);
        st_pair(out_host + ST_MAX_ROWS + tid, sap * P->inv_sap, seq32);
    }
    if (dbg) {
        __syncthreads();
        if (tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[1 + n_steps] = t; }
    }
}

struct StreamState {
    int lo = -1, hi = -1, grid = 0;
    bool ok = false;
    size_t smem[5] = {0, 0, 0, 0, 0};    // dynamic shared memory of the NB = 1 / 4 (pairs) and 8 / 12 / 16 (barriers) instantiations
    bool fits[5] = {false, false, false, false, false};
    StPlan* d_plan = nullptr;
    uint2* d_x = nullptr;                // staged input, (value, sequence) pairs
    uint2* d_act = nullptr;              // one pair buffer per step output (nothing is reused inside a call)
    float* d_xf = nullptr;               // the same as plain floats (flag protocol)
    float* d_actf = nullptr;
    uint2* d_partial = nullptr;
    unsigned long long* d_bar = nullptr;
    float* h_in = nullptr;  float* d_in = nullptr;      // pinned + mapped input [64, D]
    uint2* h_out = nullptr; uint2* d_out = nullptr;     // pinned + mapped scores [2][64] as (value, sequence) pairs
    unsigned long long seq = 0, bar2_base = 0;
    unsigned long long* d_dbg = nullptr;  // MMAD_STREAM_DEBUG=1: per-phase globaltimer stamps of CTA 0
    unsigned long long weights_gen = 0;
    cudaStream_t stream = nullptr;
    int D = 0, n_steps = 0;
};

void stream_free(StreamState* s) {
    if (!s) return;
    cudaFree(s->d_plan); cudaFree(s->d_x); cudaFree(s->d_act); cudaFree(s->d_partial); cudaFree(s->d_bar); cudaFree(s->d_dbg);
    cudaFree(s->d_xf); cudaFree(s->d_actf);
    if (s->h_in) cudaFreeHost(s->h_in);
    if (s->h_out) cudaFreeHost(s->h_out);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

constexpr int kNbOf[5] = {1, 4, 8, 12, 16};
const void* stream_kernel(int idx) {
    switch (idx) {
        case 0: return (const void*)stream_chain_kernel<1, true>;
        case 1: return (const void*)stream_chain_kernel<4, true>;
        case 2: return (const void*)stream_chain_kernel<8, false>;
        case 3: return (const void*)stream_chain_kernel<12, false>;
        default: return (const void*)stream_chain_kernel<16, false>;
    }
}

int launch(StreamState* S, int idx, int rows) {
    void* args[] = {(void*)&S->d_plan, (void*)&S->d_in, (void*)&rows, (void*)&S->d_out, (void*)&S->seq, (void*)&S->d_bar, (void*)&S->bar2_base,
                    (void*)&S->d_dbg};
    MMAD_CUDA_OK(cudaLaunchCooperativeKernel(stream_kernel(idx), dim3(S->grid), dim3(ST_THREADS), args, S->smem[idx], S->stream));
    return MMAD_OK;
}

}  // namespace

void stream_state_free(void* p) { stream_free(static_cast<StreamState*>(p)); }

bool stream_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MMAD_NO_STREAM_KERNEL"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

int stream_max_rows() { return ST_MAX_ROWS; }

// (Re)build the plan for diffs [lo, hi) of the handle's current weights.  Returns MMAD_E_UNSUPPORTED when the model does
// not fit the kernel (width > grid * 16 columns, D % 4 != 0, shared memory).
static int stream_prepare(mmad_t h, int lo, int hi) {
    StreamState* S = static_cast<StreamState*>(handle_stream_get(h));
    const mmad_desc_t* d = handle_desc(h);
    const int L = d->n_enc, Ld = d->n_dec, D = d->enc_widths[0];
    if (S && S->ok && S->lo == lo && S->hi == hi && S->weights_gen == handle_weights_gen(h)) return MMAD_OK;
    if (!S) {
        S = new StreamState();
        handle_stream_set(h, S);
        int dev = 0, sms = 0, coop = 0;
        MMAD_CUDA_OK(cudaGetDevice(&dev));
        MMAD_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        MMAD_CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        if (!coop) { set_error("device does not support cooperative launches"); return MMAD_E_UNSUPPORTED; }
        S->grid = sms;
        S->D = D;
        MMAD_CUDA_OK(cudaStreamCreateWithFlags(&S->stream, cudaStreamNonBlocking));
        MMAD_CUDA_OK(cudaMalloc(&S->d_plan, sizeof(StPlan)));
        MMAD_CUDA_OK(cudaMalloc(&S->d_bar, 8));
        MMAD_CUDA_OK(cudaMemset(S->d_bar, 0, 8));
        MMAD_CUDA_OK(cudaHostAlloc(&S->h_in, (size_t)ST_MAX_ROWS * D * 4, cudaHostAllocMapped));
        MMAD_CUDA_OK(cudaHostAlloc(&S->h_out, (size_t)2 * ST_MAX_ROWS * sizeof(uint2), cudaHostAllocMapped));
        memset(S->h_out, 0, (size_t)2 * ST_MAX_ROWS * sizeof(uint2));
        MMAD_CUDA_OK(cudaHostGetDevicePointer((void**)&S->d_in, S->h_in, 0));
        MMAD_CUDA_OK(cudaHostGetDevicePointer((void**)&S->d_out, S->h_out, 0));
        const char* dbg = getenv("MMAD_STREAM_DEBUG");
        if (dbg && dbg[0] == '1') {
            MMAD_CUDA_OK(cudaMalloc(&S->d_dbg, (ST_MAX_STEPS + 4) * 8));
            MMAD_CUDA_OK(cudaMemset(S->d_dbg, 0, (ST_MAX_STEPS + 4) * 8));
        }
    }
    S->ok = false;
    if (D % 4) { set_error("stream kernel needs D %% 4 == 0"); return MMAD_E_UNSUPPORTED; }
    // pair buffers: one per step output (L encoder + Ld decoder + L - 1 enc(xhat)), nothing is reused inside a call, so a
    // pair that carries the call's sequence number is final
    std::vector<LayerF32> enc(L), dec(Ld);
    int maxw = round_up(D, kPad);
    for (int i = 0; i < L; ++i) { enc[i] = handle_layer_f32(h, 0, i); maxw = std::max(maxw, enc[i].Np); }
    for (int i = 0; i < Ld; ++i) { dec[i] = handle_layer_f32(h, 1, i); maxw = std::max(maxw, dec[i].Np); }
    for (auto& l : enc) if (l.N > S->grid * ST_CPC) { set_error("layer too wide for the stream kernel"); return MMAD_E_UNSUPPORTED; }
    for (auto& l : dec) if (l.N > S->grid * ST_CPC) { set_error("layer too wide for the stream kernel"); return MMAD_E_UNSUPPORTED; }
    const size_t buf = (size_t)ST_MAX_ROWS * maxw;       // pairs per buffer
    const int n_buf = 2 * L + Ld;
    if (!S->d_act) {
        MMAD_CUDA_OK(cudaMalloc(&S->d_act, buf * n_buf * 8));
        MMAD_CUDA_OK(cudaMalloc(&S->d_x, buf * 8));
        MMAD_CUDA_OK(cudaMalloc(&S->d_partial, (size_t)(L + 1) * S->grid * ST_MAX_ROWS * 8));
        MMAD_CUDA_OK(cudaMemset(S->d_partial, 0, (size_t)(L + 1) * S->grid * ST_MAX_ROWS * 8));
        MMAD_CUDA_OK(cudaMemset(S->d_act, 0, buf * n_buf * 8));       // sequence 0 is never used by a call
        MMAD_CUDA_OK(cudaMemset(S->d_x, 0, buf * 8));
        MMAD_CUDA_OK(cudaMalloc(&S->d_actf, buf * n_buf * 4));
        MMAD_CUDA_OK(cudaMalloc(&S->d_xf, buf * 4));
    }
    MMAD_CUDA_OK(cudaMemset(S->d_actf, 0, buf * n_buf * 4));          // padding columns of the float buffers stay zero for ever
    MMAD_CUDA_OK(cudaMemset(S->d_xf, 0, buf * 4));
    // buffers: 0..L-1 enc(x) (the diff references), L..L+Ld-1 decoder, then enc(xhat)
    StPlan P;
    memset(&P, 0, sizeof P);
    P.lo = lo; P.hi = hi; P.D = D; P.ldx = round_up(D, kPad); P.grid = S->grid;
    P.n_diffs = L + 1;
    P.slope = d->lrelu_slope;
    P.xp = S->d_x; P.pbase = S->d_act; P.xf = S->d_xf; P.fbase = S->d_actf;
    P.bufsz = buf; P.ld = maxw;
    P.partial = S->d_partial;
    P.inv_base = 1.f / D;
    int dsel = 0;
    for (int l = lo; l < hi; ++l) dsel += d->enc_widths[l];
    P.inv_sap = 1.f / dsel;
    int ns = 0;
    size_t need[5] = {0, 0, 0, 0, 0};
    int prev_nprod = 0, wmax4 = 0, kpmax = 0;
    auto add = [&](const LayerF32& Lr, int ibuf, int obuf, int rbuf, int diff) {
        StStep& st = P.step[ns++];
        st.W = Lr.W; st.bias = Lr.bias; st.scale = Lr.has_bn ? Lr.scale : nullptr; st.shift = Lr.has_bn ? Lr.shift : nullptr;
        st.ibuf = ibuf; st.obuf = obuf; st.rbuf = rbuf;
        st.K = Lr.K; st.K4 = Lr.Kp / 4; st.N = Lr.N;
        st.cpc = (Lr.N + S->grid - 1) / S->grid;
        st.diff = diff;
        prev_nprod = (Lr.N + st.cpc - 1) / st.cpc;
        if (diff >= 0) P.nact[diff] = prev_nprod;
        wmax4 = std::max(wmax4, st.cpc * st.K4);
        kpmax = std::max(kpmax, Lr.Kp);
        for (int i = 2; i < 5; ++i) need[i] = std::max(need[i], (size_t)(st.cpc + kNbOf[i]) * Lr.Kp * 4);
    };
    int cur = -1;
    for (int l = 0; l < L; ++l) { add(enc[l], cur, l, -2, -1); cur = l; }
    const int want_enc2 = hi > 1;
    for (int l = 0; l < Ld; ++l) {
        const bool last = l == Ld - 1;
        add(dec[l], cur, L + l, last ? -1 : -2, last ? 0 : -1);
        cur = L + l;
    }
    if (want_enc2) {
        const int last = std::min(L, hi - 1);
        for (int l = 1; l <= last; ++l) {
            const int out = l == last ? -2 : L + Ld + l - 1;
            add(enc[l - 1], cur, out, l - 1, l);
            cur = out;
        }
    }
    P.wmax4 = wmax4;
    for (int i = 0; i < 2; ++i) need[i] = (size_t)2 * wmax4 * 16 + (size_t)kNbOf[i] * kpmax * 4;      // pair protocol: two slices + the tile
    P.n_steps = ns;          // diffs beyond `last` have no producers: nact == 0, their sum is 0
    S->n_steps = ns;
    MMAD_CUDA_OK(cudaMemcpy(S->d_plan, &P, sizeof P, cudaMemcpyHostToDevice));
    for (int i = 0; i < 5; ++i) {
        S->smem[i] = need[i];
        S->fits[i] = need[i] <= (size_t)ST_SMEM_CAP;
        if (S->fits[i]) {
            MMAD_CUDA_OK(cudaFuncSetAttribute(stream_kernel(i), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need[i]));
            int nblk = 0;
            MMAD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, stream_kernel(i), ST_THREADS, need[i]));
            if (nblk < 1) S->fits[i] = false;
        }
    }
    S->lo = lo; S->hi = hi;
    S->weights_gen = handle_weights_gen(h);
    S->ok = true;
    return MMAD_OK;
}

// One realtime call: rows <= 64 windows from host memory, base and SAP scores back on the host.  Returns
// MMAD_E_UNSUPPORTED (without side effects) when the model / row count does not fit, so the caller can take the graph path.
int stream_score(mmad_t h, const float* h_x, int ldx, int rows, int lo, int hi, float* h_base, float* h_sap) {
    if (rows < 1 || rows > ST_MAX_ROWS) { set_error("stream kernel: 1..64 rows"); return MMAD_E_UNSUPPORTED; }
    int rc = stream_prepare(h, lo, hi);
    if (rc) return rc;
    StreamState* S = static_cast<StreamState*>(handle_stream_get(h));
    // <= 4 rows: pair protocol; taller calls: barriers, the smallest row tile that covers the call (chunks of 16 beyond that)
    int idx = rows <= 1 ? 0 : (rows <= 4 ? 1 : (rows <= 8 ? 2 : (rows <= 12 ? 3 : 4)));
    while (idx >= 2 && !S->fits[idx]) --idx;          // a narrower row tile needs less shared memory (more chunks)
    if (idx == 1 && rows > 4) idx = 2;
    if (!S->fits[idx]) { set_error("stream kernel: shared memory"); return MMAD_E_UNSUPPORTED; }
    const int D = S->D;
    if (h_x != S->h_in) {
        if (ldx == D) memcpy(S->h_in, h_x, (size_t)rows * D * 4);
        else for (int r = 0; r < rows; ++r) memcpy(S->h_in + (size_t)r * D, h_x + (size_t)r * ldx, (size_t)D * 4);
    }
    S->seq += 1;
    rc = launch(S, idx, rows);
    if (rc) return rc;
    MMAD_LAUNCHED();
    if (idx >= 2) S->bar2_base += (unsigned long long)S->grid * (unsigned long long)S->n_steps;   // staging + n_steps - 1 barriers
    // the doorbell: every score arrives as an 8-byte (value, sequence) pair in mapped pinned memory; a pair that carries this
    // call's number is final (single-copy atomic store), so the host polls the scores themselves
    const uint32_t seq32 = (uint32_t)S->seq;
    volatile uint2* hp = S->h_out;
    auto arrived = [&]() {
        for (int r = rows - 1; r >= 0; --r)
            if (hp[r].y != seq32 || hp[ST_MAX_ROWS + r].y != seq32) return false;
        return true;
    };
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    bool ok = arrived();
    while (!ok) {
        __builtin_ia32_pause();
        ok = arrived();
        if (!ok && (++spins & 0xFFFF) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(5)) break;
    }
    if (!ok) {        // never rang: fetch the launch / execution error
        cudaError_t e = cudaStreamSynchronize(S->stream);
        if (e != cudaSuccess || !arrived()) {
            S->ok = false;
            set_error("stream kernel did not complete: %s", cudaGetErrorString(e));
            return MMAD_E_CUDA;
        }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    if (S->d_dbg) {      // per-phase times of this call (ns): staging, then every step incl. its barrier
        unsigned long long t[ST_MAX_STEPS + 4];
        const double host_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        cudaStreamSynchronize(S->stream);
        cudaMemcpy(t, S->d_dbg, (S->n_steps + 2) * 8, cudaMemcpyDeviceToHost);
        fprintf(stderr, "mmad stream rows=%d launch->doorbell %.1f us | kernel %.1f us: staging %.1f |", rows, host_us,
                (t[S->n_steps + 1] - t[0]) / 1e3, (t[1] - t[0]) / 1e3);
        for (int i = 0; i < S->n_steps; ++i) fprintf(stderr, " %.1f", (t[2 + i] - t[1 + i]) / 1e3);
        fprintf(stderr, "\n");
    }
    for (int r = 0; r < rows; ++r) {
        if (h_base) memcpy(h_base + r, (const void*)&S->h_out[r].x, 4);
        if (h_sap) memcpy(h_sap + r, (const void*)&S->h_out[ST_MAX_ROWS + r].x, 4);
    }
    return MMAD_OK;
}

float* stream_input_buffer(mmad_t h, int lo, int hi) {
    if (stream_prepare(h, lo, hi)) return nullptr;
    return static_cast<StreamState*>(handle_stream_get(h))->h_in;
}

}  // namespace mmad
