// Shared pieces of the tcgen05 GEMM kernels (gemm_tc.cu: one CTA per tile; gemm_tc2.cu: CTA pairs, cta_group::2):
// tile constants, PTX wrappers, and the fused epilogue applied to one 32-row x BN-column accumulator slice.
#pragma once
#include <cuda_fp8.h>

#include "mmad_internal.cuh"

namespace mmad {
namespace tc {

constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int STG_FLOAT4 = 32 * 4;          // float4 slots of one warp's transpose tile (32 rows x 16 columns)
constexpr int ROWPART_COLS = 128;           // columns covered by one row-partial slot
constexpr int BM = 128;
constexpr int BN_MAX = 256;            // CTA tile is 128 x BN with BN = 256 (throughput) or 128 (under-filled grids)
constexpr int BK = 64;                 // halfs per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 64 + 256;          // TMA warp, MMA warp, eight epilogue warps
constexpr int A_TILE_BYTES = BM * BK * 2;   // 16 KB

template <int PASSES, int BN> struct Cfg {
    static constexpr int kBTileBytes = BN * BK * 2;            // 32 KB / 16 KB
    // PASSES 3: A hi+lo, B hi+lo;  2: A hi+lo, B hi (lo*hi + hi*hi);  1: hi only;
    // 4: fp16 hi*hi + fp8 [lo8 | a8] . [Wh8 ; Wl8] (the fp8 twins take the place of the lo tiles, same bytes)
    static constexpr int kStageBytes = (PASSES >= 2 ? 2 : 1) * A_TILE_BYTES + (PASSES >= 3 ? 2 : 1) * kBTileBytes;
    static constexpr int kStages = (192 * 1024) / kStageBytes > 6 ? 6 : (192 * 1024) / kStageBytes;   // 2 / 3 / 4 / 6
    static constexpr int kSmemTiles = kStages * kStageBytes;   // <= 192 KB
    static constexpr int kTmemCols = 2 * BN;                   // two fp32 accumulators
    static constexpr int kSmemBytes = kSmemTiles + 4 * BN * 4 /*epilogue vectors*/ + EPI_WARPS * STG_FLOAT4 * 16 /*transpose tiles*/ +
                                      256 /*barriers*/ + 1024 /*align*/;
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (clock64() - t0 < 4000000000LL)     // ~2 s at 2 GHz
        if (mbar_try_wait(bar, parity)) return;
    printf("mmad gemm_tc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
    __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// smem matrix descriptor, 128-byte swizzle (SM100 descriptor version 1).
//   K-major  (contraction dim contiguous): rows of 64 halfs; 8-row groups SBO = 1024 B apart; LBO unused.
//   MN-major (M/N dim contiguous): one 128-byte row per k holding 64 consecutive m (or n); 8-k groups
//            SBO = 1024 B apart; the next 64 m/n start LBO = 8192 B further (one TMA box of 64 k rows).
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);                          // start address, 16-byte units
    d |= (uint64_t)(MN_MAJOR ? (8192 >> 4) : 1) << 16;               // leading byte offset
    d |= (uint64_t)(1024 >> 4) << 32;                                // stride byte offset
    d |= (uint64_t)1 << 46;                                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                          // SWIZZLE_128B
    return d;
}
// instruction descriptor: fp16 x fp16 -> fp32, M = 128, N = n; bits 15/16 select MN-major A / B
__device__ __forceinline__ uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// e4m3 x e4m3 -> fp32 (K = 32 per instruction; the instruction descriptor's format fields are 0 = E4M3, so the
// descriptor value is the one of the fp16 kind)
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


// ---- fused epilogue --------------------------------------------------------------------------------
// Per-column vectors of a tile (multiplier, bias, BatchNorm scale / shift).  Epilogue thread et owns column et of the
// tile (EPI_THREADS >= BN): it LOADS the four values of the next tile's column into registers while the current tile is
// drained (epi_vec_load), and STORES them to shared memory between the two named barriers at the top of the next
// iteration (epi_vec_store).  Loading inside the barriers put ~40 % of the epilogue warps' time on the global-load
// latency of four scalars (profiles/r1_ncu_full_gemm_tc2_f16f8.md, source page).
struct EpiVec { float mul, bias, sc, sh; };

__device__ __forceinline__ EpiVec epi_vec_load(const Epilogue& e, int N, int n0, int sp, int c, int n_kb) {
    const int gc = n0 + c;
    const bool ok = gc < N;
    EpiVec v;
    v.mul = e.acc_scale * fmaf(e.acc_comp, (float)n_kb, 1.f) * ((e.col_scale && ok) ? __ldg(e.col_scale + gc) : 1.f);
    v.bias = (e.bias && ok && sp == 0) ? __ldg(e.bias + gc) : 0.f;
    v.sc = (e.bn_scale && ok) ? __ldg(e.bn_scale + gc) : 1.f;
    v.sh = (e.bn_scale && ok) ? __ldg(e.bn_shift + gc) : 0.f;
    return v;
}
__device__ __forceinline__ void epi_vec_store(const EpiVec& v, int c, float* s_mul, float* s_bias, float* s_sc, float* s_sh) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(s_mul + c)), "f"(v.mul) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(s_bias + c)), "f"(v.bias) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(s_sc + c)), "f"(v.sc) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(s_sh + c)), "f"(v.sh) : "memory");
}

// Eight epilogue warps per CTA: warps 2-5 drain columns [0, 128) of a tile, warps 6-9 columns [128, 256); warp w may
// only touch TMEM lanes 32*(w%4).., so each lane quarter has one warp per column half.  (With four warps -- one per
// scheduler -- the epilogue is ALU-latency bound and paces the diff layers; two per scheduler hide each other.)
// A warp works in 32-row x 16-column pieces: tcgen05.ld hands each thread one accumulator ROW (16 columns), the
// piece is transposed through a 2 KB XOR-swizzled shared tile and re-read as (8 rows x 4 float4 columns), so every
// global access of the warp covers whole 64-byte (fp32) / 32-byte (fp16) row segments.
// sq[it] returns this lane's partial row sums of squares for rows row_base + it*8 + lane/4.

__device__ __forceinline__ int stg_slot(int row, int j) { return row * 4 + (j ^ ((row >> 1) & 3)); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// four e4m3 values, first in the low byte
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
    const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
    const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
    return lo | (hi << 16);
}
// fp8 twin of four consecutive columns starting at col (multiple of 4) of a row whose fp16 twin starts at `row`:
// 8 bytes at byte offset 2 * col = [4 x e4m3(residual * 2^11) | 4 x e4m3(value)]  (one 8-byte store, like the fp16 lo)
__device__ __forceinline__ void store_f8_twin(uint8_t* row, int col, float v0, float v1, float v2, float v3,
                                              float r0, float r1, float r2, float r3) {
    uint2 t;
    t.x = pack_e4m3x4(r0 * kF8LoScale, r1 * kF8LoScale, r2 * kF8LoScale, r3 * kF8LoScale);
    t.y = pack_e4m3x4(v0, v1, v2, v3);
    *reinterpret_cast<uint2*>(row + 2 * col) = t;
}

__device__ __forceinline__ uint32_t stg_off32(int row16, int j8) { return (uint32_t)(row16 * 8 + (j8 ^ (row16 & 7))) * 16u; }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
    return r;
}


// plain store mode (pre-activations of the BatchNorm layers, dX, parameter gradients; with split-K the partial sums are added
// atomically).  A train step runs ONE tile per CTA, so this epilogue is not hidden behind the next tile's MMAs: it is written
// for latency -- 32-column blocks (one tcgen05.ld each), the transposed block re-read with every load of a phase independent
// of the others, 16-byte stores / red.v4 when the rows of Y are 16-byte aligned.  (The first version walked 16-column pieces
// with a dependent load -> store chain per row pair: 6-8 us per tile, 60 % of every training GEMM.)
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr) : "memory");
    return r;
}

template <int BN>
__device__ __forceinline__ void epi_tile_plain(const Epilogue& e, int M, int N, int splits, int sp, int row_base, int n0, uint32_t taddr,
                                               float4* stg, const float* s_mul, const float* s_bias, int lane, int col_lo,
                                               int col_hi) {
    int n_cols = N - n0; if (n_cols > BN) n_cols = BN;
    int c_end = (n_cols + 31) & ~31;
    if (c_end > col_hi) c_end = col_hi;
    float* const Yb = e.Y + (e.slab_stride ? (size_t)sp * (size_t)e.slab_stride : 0);      // slab of this split, or the one output
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(Yb) & 15) == 0) && (e.ldy % 4 == 0);
    const bool atomic = splits > 1 && !e.slab_stride;
    const uint32_t stg_s = smem_u32(stg), mul_s = smem_u32(s_mul), bias_s = smem_u32(s_bias);
    for (int c0 = col_lo; c0 < c_end; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        if (vec_ok) {
            // lane = (row in a group of 4, float4 column): 4 rows x 128 contiguous bytes per instruction
            const int rsub = lane >> 3, cg = lane & 7;
            const int cc = c0 + cg * 4, gc = n0 + cc;
            const float4 mul = lds128(mul_s + cc * 4), bia = lds128(bias_s + cc * 4);
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {
                if ((lane >> 4) == ph) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) sts128(stg_s + stg_off32(lane & 15, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                __syncwarp();
                float4 a[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) a[it] = lds128(stg_s + stg_off32(it * 4 + rsub, cg));
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int r = row_base + ph * 16 + it * 4 + rsub;
                    if (r >= M || gc >= N) continue;
                    const float x0 = fmaf(a[it].x, mul.x, bia.x), x1 = fmaf(a[it].y, mul.y, bia.y);
                    const float x2 = fmaf(a[it].z, mul.z, bia.z), x3 = fmaf(a[it].w, mul.w, bia.w);
                    float* dst = Yb + (size_t)r * e.ldy + gc;
                    if (gc + 3 < N) {
                        if (atomic) red_add_v4(dst, x0, x1, x2, x3);
                        else *reinterpret_cast<float4*>(dst) = make_float4(x0, x1, x2, x3);
                    } else {
                        const float xs[3] = {x0, x1, x2};
                        for (int j = 0; j < 3 && gc + j < N; ++j) {
                            if (atomic) atomicAdd(dst + j, xs[j]); else dst[j] = xs[j];
                        }
                    }
                }
                __syncwarp();
            }
        } else {
            // rows of Y are only 4-byte aligned (parameter gradients [N, K] with odd K): lane = column, one row per instruction
            const int gc = n0 + c0 + lane;
            const float mulc = lds32(mul_s + (c0 + lane) * 4), biac = lds32(bias_s + (c0 + lane) * 4);
            const uint32_t rd = stg_s + (uint32_t)(lane & 3) * 4u;
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {
                if ((lane >> 4) == ph) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) sts128(stg_s + stg_off32(lane & 15, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                __syncwarp();
                float a[16];
#pragma unroll
                for (int rl = 0; rl < 16; ++rl) a[rl] = lds32(rd + stg_off32(rl, lane >> 2));
#pragma unroll
                for (int rl = 0; rl < 16; ++rl) {
                    const int r = row_base + ph * 16 + rl;
                    if (r >= M || gc >= N) continue;
                    const float val = fmaf(a[rl], mulc, biac);
                    if (atomic) atomicAdd(Yb + (size_t)r * e.ldy + gc, val);
                    else Yb[(size_t)r * e.ldy + gc] = val;
                }
                __syncwarp();
            }
        }
    }
}

// Fused epilogue of one warp's 32 rows x [col_lo, col_hi) slice.  The warp works in 32-column blocks: one tcgen05.ld
// hands each thread its accumulator ROW (32 columns); the block goes through a 2 KB XOR-swizzled shared tile in two
// phases of 16 rows and is re-read as (4 rows x 8 float4 columns), so every global access of the warp covers whole
// 128-byte (fp32) / 64-byte (fp16, fp8 twin) row segments.  Blocks that lie completely inside the matrix (all but the
// ragged last block of a layer / the last row tile) take the FULL instantiation: no per-element predicates or zero
// selects, row pointers advanced by constant strides -- the epilogue is issue-bound (two warps per scheduler), so the
// instruction count of this loop is what paces the chain layers once the MMA work drops to two passes.
// sq[ph * 4 + it] returns this lane's partial row sum of squares for row row_base + ph*16 + it*4 + lane/8.
struct EpiPtrs {     // per-lane pointers at (row_base + lane/8, n0 + 4 * (lane%8)); null when the output is absent
    float* Y; float* pre; float* dout; const float* ref;
    __half* Yh; uint8_t* Yl; __half* Dh; uint8_t* Dl;
};

template <bool FULL>
__device__ __forceinline__ void epi_block(const Epilogue& e, const EpiPtrs& q, int M, int row_base, int c0, int cc, int n_cols,
                                          int w_cols, int dw_cols, int pre_cols, const uint32_t (&v)[32], uint32_t stg_s,
                                          const float4& mul, const float4& bia, const float4& sc, const float4& sh, int lane,
                                          float (&sq)[8]) {
    const int rsub = lane >> 3, cg = lane & 7;
    const bool k0 = FULL || cc < n_cols, k1 = FULL || cc + 1 < n_cols, k2 = FULL || cc + 2 < n_cols, k3 = FULL || cc + 3 < n_cols;
    const bool wy = FULL || cc < w_cols, wd = FULL || cc < dw_cols, wp = FULL || cc < pre_cols;
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
        // this phase's reference loads first: 4 independent 16-byte loads per lane in flight during the transpose
        float4 rf[4];
        if (q.ref) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int rr = ph * 16 + it * 4;
                rf[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (FULL || (row_base + rr + rsub < M && k3))
                    rf[it] = __ldg(reinterpret_cast<const float4*>(q.ref + (size_t)rr * e.ldref + c0));
            }
        }
        if ((lane >> 4) == ph) {
#pragma unroll
            for (int j = 0; j < 8; ++j) sts128(stg_s + stg_off32(lane & 15, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int rl = it * 4 + rsub;
            const int rr = ph * 16 + it * 4;            // row offset from this lane's base row
            const float4 a = lds128(stg_s + stg_off32(rl, cg));
            float x0 = fmaf(a.x, mul.x, bia.x), x1 = fmaf(a.y, mul.y, bia.y);
            float x2 = fmaf(a.z, mul.z, bia.z), x3 = fmaf(a.w, mul.w, bia.w);
            if (FULL || row_base + rr + rsub < M) {
                if (q.pre && wp)     // train: pre-activation, zero padded to ldpre columns
                    *reinterpret_cast<float4*>(q.pre + (size_t)rr * e.ldpre + c0) =
                        make_float4(k0 ? x0 : 0.f, k1 ? x1 : 0.f, k2 ? x2 : 0.f, k3 ? x3 : 0.f);
                if (e.bn_scale) {
                    x0 = x0 > 0.f ? x0 : x0 * e.slope; x1 = x1 > 0.f ? x1 : x1 * e.slope;
                    x2 = x2 > 0.f ? x2 : x2 * e.slope; x3 = x3 > 0.f ? x3 : x3 * e.slope;
                    x0 = fmaf(x0, sc.x, sh.x); x1 = fmaf(x1, sc.y, sh.y);
                    x2 = fmaf(x2, sc.z, sh.z); x3 = fmaf(x3, sc.w, sh.w);
                }
                if (!FULL) { x0 = k0 ? x0 : 0.f; x1 = k1 ? x1 : 0.f; x2 = k2 ? x2 : 0.f; x3 = k3 ? x3 : 0.f; }
                if (q.Y && wy) *reinterpret_cast<float4*>(q.Y + (size_t)rr * e.ldy + c0) = make_float4(x0, x1, x2, x3);
                if (q.Yh && wy) {
                    const float ys = e.y_split_scale;
                    const float y0 = x0 * ys, y1 = x1 * ys, y2 = x2 * ys, y3 = x3 * ys;
                    const __half2 h01 = __floats2half2_rn(y0, y1), h23 = __floats2half2_rn(y2, y3);
                    uint2 hv;
                    hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                    *reinterpret_cast<uint2*>(q.Yh + (size_t)rr * e.ldh + c0) = hv;
                    if (q.Yl) {
                        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                        uint2 lv;
                        if (e.lo_f8) {
                            lv.x = pack_e4m3x4((y0 - f01.x) * kF8LoScale, (y1 - f01.y) * kF8LoScale, (y2 - f23.x) * kF8LoScale, (y3 - f23.y) * kF8LoScale);
                            lv.y = pack_e4m3x4(y0, y1, y2, y3);
                        } else {
                            const __half2 l01 = __floats2half2_rn(y0 - f01.x, y1 - f01.y);
                            const __half2 l23 = __floats2half2_rn(y2 - f23.x, y3 - f23.y);
                            lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                        }
                        *reinterpret_cast<uint2*>(q.Yl + ((size_t)rr * e.ldh + c0) * 2) = lv;
                    }
                }
                if (q.ref) {
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
                    if (k3) {
                        const float4 f = rf[it];
                        d0 = x0 - f.x; d1 = x1 - f.y; d2 = x2 - f.z; d3 = x3 - f.w;
                    } else if (k0) {
                        const float* rp = q.ref + (size_t)rr * e.ldref + c0;
                        d0 = x0 - rp[0]; if (k1) d1 = x1 - rp[1]; if (k2) d2 = x2 - rp[2];
                    }
                    float& acc = sq[ph * 4 + it];
                    acc = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, acc))));
                    if (q.dout && wd) *reinterpret_cast<float4*>(q.dout + (size_t)rr * e.lddout + c0) = make_float4(d0, d1, d2, d3);
                    if (q.Dh && wd) {
                        const float s0 = d0 * e.d_scale, s1 = d1 * e.d_scale, s2 = d2 * e.d_scale, s3 = d3 * e.d_scale;
                        const __half2 h01 = __floats2half2_rn(s0, s1), h23 = __floats2half2_rn(s2, s3);
                        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                        uint2 hv, lv;
                        hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                        *reinterpret_cast<uint2*>(q.Dh + (size_t)rr * e.lddh + c0) = hv;
                        if (e.d_lo_f8) {
                            lv.x = pack_e4m3x4((s0 - f01.x) * kF8LoScale, (s1 - f01.y) * kF8LoScale, (s2 - f23.x) * kF8LoScale, (s3 - f23.y) * kF8LoScale);
                            lv.y = pack_e4m3x4(s0, s1, s2, s3);
                        } else {
                            const __half2 l01 = __floats2half2_rn(s0 - f01.x, s1 - f01.y);
                            const __half2 l23 = __floats2half2_rn(s2 - f23.x, s3 - f23.y);
                            lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                        }
                        *reinterpret_cast<uint2*>(q.Dl + ((size_t)rr * e.lddh + c0) * 2) = lv;
                    }
                } else if (e.sq_self) {
                    float& acc = sq[ph * 4 + it];
                    acc = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, acc))));
                }
            }
        }
        __syncwarp();     // the staging tile is rewritten by the next phase
    }
}

template <int BN>
__device__ __forceinline__ void epi_tile(const Epilogue& e, int M, int N, int splits, int sp, int row_base, int n0, uint32_t taddr,
                                         float4* stg, const float* s_mul, const float* s_bias, const float* s_sc,
                                         const float* s_sh, int lane, int col_lo, int col_hi, float (&sq)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sq[i] = 0.f;
    if (e.plain) {
        epi_tile_plain<BN>(e, M, N, splits, sp, row_base, n0, taddr, stg, s_mul, s_bias, lane, col_lo, col_hi);
        return;
    }
    const int rsub = lane >> 3;             // 0..3  row inside a group of 4
    const int cg = lane & 7;                // float4 column group inside the 32-column block
    int n_cols = N - n0; if (n_cols > BN) n_cols = BN;          // valid columns of this tile
    int w_cols = e.y_cols - n0; if (w_cols > BN) w_cols = BN;      // activation columns to write (zero padded)
    int dw_cols = e.ref ? e.d_cols - n0 : 0; if (dw_cols > BN) dw_cols = BN;   // diff columns to write
    const int pre_cols = e.pre ? e.ldpre - n0 : 0;
    int c_end = (max(max(n_cols, w_cols), dw_cols) + 31) & ~31;
    if (c_end > col_hi) c_end = col_hi;
    const bool rows_full = row_base + 32 <= M;
    // per-lane base pointers at (row_base + rsub, n0 + 4 * cg); a block adds c0, a row step adds rr * ld
    const size_t r0 = (size_t)(row_base + rsub);
    const int cb = n0 + cg * 4;
    EpiPtrs q;
    q.Y = e.Y ? e.Y + r0 * e.ldy + cb : nullptr;
    q.pre = e.pre ? e.pre + r0 * e.ldpre + cb : nullptr;
    q.ref = e.ref ? e.ref + r0 * e.ldref + cb : nullptr;
    q.dout = (e.ref && e.dout) ? e.dout + r0 * e.lddout + cb : nullptr;
    q.Yh = e.Yh ? e.Yh + r0 * e.ldh + cb : nullptr;
    q.Yl = (e.Yh && e.Yl) ? reinterpret_cast<uint8_t*>(e.Yl) + (r0 * e.ldh + cb) * 2 : nullptr;
    q.Dh = (e.ref && e.Dh) ? e.Dh + r0 * e.lddh + cb : nullptr;
    q.Dl = (e.ref && e.Dh) ? reinterpret_cast<uint8_t*>(e.Dl) + (r0 * e.lddh + cb) * 2 : nullptr;
    const uint32_t stg_s = smem_u32(stg);
    const uint32_t mul_s = smem_u32(s_mul), bias_s = smem_u32(s_bias), sc_s = smem_u32(s_sc), sh_s = smem_u32(s_sh);
    for (int c0 = col_lo; c0 < c_end; c0 += 32) {
        const int cc = c0 + cg * 4;             // tile-local column of this lane's float4
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        const float4 mul = lds128(mul_s + cc * 4), bia = lds128(bias_s + cc * 4);
        const float4 sc = lds128(sc_s + cc * 4), sh = lds128(sh_s + cc * 4);
        if (rows_full && c0 + 32 <= n_cols)
            epi_block<true>(e, q, M, row_base, c0, cc, n_cols, w_cols, dw_cols, pre_cols, v, stg_s, mul, bia, sc, sh, lane, sq);
        else
            epi_block<false>(e, q, M, row_base, c0, cc, n_cols, w_cols, dw_cols, pre_cols, v, stg_s, mul, bia, sc, sh, lane, sq);
    }
}

// row partial sums of squares of this warp's column half: slot = 128-column block index
__device__ __forceinline__ void epi_rowpart(const Epilogue& e, int M, int row_base, int slot, int lane, const float (&sq)[8]) {
    if (!e.rowpart) return;
    const int rsub = lane >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = sq[i];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        const int r = row_base + (i >> 2) * 16 + (i & 3) * 4 + rsub;
        if ((lane & 7) == 0 && r < M) e.rowpart[(size_t)slot * e.rowpart_stride + r] = v;
    }
}

}  // namespace tc
}  // namespace mmad
