// tcgen05 / TMEM / TMA fused layer GEMM for sm_100a  (MMAD_PREC_F16X3, MMAD_PREC_F16)
//
//   acc[m,n] = sum_k A[m,k] * B[n,k]         A = activations [M,K], B = weights [N,K]  (Y = X W^T)
//
// Every fp32 operand is carried as an fp16 pair (hi + lo, x ~= hi + lo to ~2^-22).  PASSES == 3
// issues  hi*hi + hi*lo + lo*hi  into one fp32 TMEM accumulator (fp32-equivalent products at the
// fp16 tensor-core rate / 3); PASSES == 1 issues hi*hi only.
//
// Persistent, warp-specialised, one CTA per SM (grid = min(tiles, #SM)):
//   warp 0     TMA producer: cp.async.bulk.tensor 2-D boxes [rows x 64 halfs], 128-byte swizzle,
//              into a ring of smem stages, completion on mbarriers (expect_tx)
//   warp 1     MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::f16, M=128, N<=256,
//              K=16 per instruction, smem descriptors advanced by 32 B inside the swizzle atom;
//              tcgen05.commit releases smem stages and publishes the accumulator
//   warps 2-5  epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) from one of two
//              256-column TMEM accumulator buffers while the next tile's MMAs fill the other;
//              fused bias / LeakyReLU / BatchNorm affine / diff against the stashed activation /
//              per-row sum of squares / fp16 hi-lo re-split for the next layer.
// CTA tile 128 x 256 x 64.  Tiles are ordered n-fastest so CTAs running together share the
// activation tile in L2.  All mbarrier waits are bounded and trap instead of hanging the GPU.
#include <dlfcn.h>

#include "mmad_internal.cuh"

namespace mmad {

namespace {

constexpr int BM = 128;
constexpr int BN_MAX = 256;            // CTA tile is 128 x BN with BN = 256 (throughput) or 128 (under-filled grids)
constexpr int BK = 64;                 // halfs per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 192;
constexpr int A_TILE_BYTES = BM * BK * 2;   // 16 KB
constexpr int STG_LD = 36;                  // floats per row of an epilogue transpose tile (conflict-free float4)

template <int PASSES, int BN> struct Cfg {
    static constexpr int kOperands = PASSES == 3 ? 2 : 1;      // hi (+ lo)
    static constexpr int kBTileBytes = BN * BK * 2;            // 32 KB / 16 KB
    static constexpr int kStageBytes = kOperands * (A_TILE_BYTES + kBTileBytes);
    static constexpr int kStages = (192 * 1024) / kStageBytes > 6 ? 6 : (192 * 1024) / kStageBytes;   // 2 / 3 / 4 / 6
    static constexpr int kSmemTiles = kStages * kStageBytes;   // <= 192 KB
    static constexpr int kTmemCols = 2 * BN;                   // two fp32 accumulators
    static constexpr int kSmemBytes = kSmemTiles + 4 * BN * 4 /*epilogue vectors*/ + 4 * 32 * STG_LD * 4 /*transpose tiles*/ +
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

struct TcParams {
    int M, N, K;
    int tiles_m, tiles_n;
    int tri;             // B is upper triangular (B[n,k] = 0 for k < n): k-blocks left of the tile's first row are skipped
    int splits;          // split-K factor (plain epilogue only): work item = (tile, split), atomically accumulated
};

// ---- the kernel ---------------------------------------------------------------------------------
template <int PASSES, bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
               const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
               TcParams p, Epilogue e) {
    using C = Cfg<PASSES, BN>;
    constexpr int B_TILE_BYTES = C::kBTileBytes;
    constexpr int TMEM_COLS = C::kTmemCols;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* s_mul = reinterpret_cast<float*>(smem + C::kSmemTiles);
    float* s_bias = s_mul + BN;
    float* s_sc = s_bias + BN;
    float* s_sh = s_sc + BN;
    float* s_stage = s_sh + BN;                  // 4 warps x [32][STG_LD]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + 4 * 32 * STG_LD);
    uint64_t* full = bars;                       // [kStages]  TMA -> MMA
    uint64_t* empty = bars + C::kStages;         // [kStages]  MMA -> TMA
    uint64_t* acc_full = empty + C::kStages;     // [2]        MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;          // [2]        epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int num_kb = (p.K + BK - 1) / BK;
    const int num_tiles = p.tiles_m * p.tiles_n * p.splits;     // work items: tile-major, split fastest
    // k-block range of split sp: [sp * num_kb / splits, (sp + 1) * num_kb / splits)  (host guarantees splits <= num_kb)

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::kStages; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), 4); }
        fence_barrier_init();
        tma_prefetch_desc(&mapAh); tma_prefetch_desc(&mapBh);
        if (PASSES == 3) { tma_prefetch_desc(&mapAl); tma_prefetch_desc(&mapBl); }
    }
    if (warp == 1) {   // TMEM allocation is a warp-wide operation; the same warp frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int w = blockIdx.x; w < num_tiles; w += gridDim.x) {
                const int t = w / p.splits, sp = w % p.splits;
                const int m0 = (t / p.tiles_n) * BM, n0 = (t % p.tiles_n) * BN;
                const int kb_lo = p.tri ? n0 / BK : sp * num_kb / p.splits, kb_hi = (sp + 1) * num_kb / p.splits;
                for (int kb = kb_lo; kb < kb_hi; ++kb) {
                    mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                    const uint32_t fb = smem_u32(&full[stage]);
                    mbar_expect_tx(fb, C::kStageBytes);
                    uint8_t* st = smem + stage * C::kStageBytes;
                    auto load_a = [&](uint8_t* dst, const CUtensorMap* map) {
                        if (A_MN) {   // [k rows, m contiguous]: boxes of 64 m x 64 k
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j) tma_load_2d(smem_u32(dst + j * 8192), map, fb, m0 + j * 64, kb * BK);
                        } else {
                            tma_load_2d(smem_u32(dst), map, fb, kb * BK, m0);
                        }
                    };
                    auto load_b = [&](uint8_t* dst, const CUtensorMap* map) {
                        if (B_MN) {
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j) tma_load_2d(smem_u32(dst + j * 8192), map, fb, n0 + j * 64, kb * BK);
                        } else {
                            tma_load_2d(smem_u32(dst), map, fb, kb * BK, n0);
                        }
                    };
                    load_a(st, &mapAh);
                    load_b(st + A_TILE_BYTES, &mapBh);
                    if (PASSES == 3) {
                        load_a(st + A_TILE_BYTES + B_TILE_BYTES, &mapAl);
                        load_b(st + 2 * A_TILE_BYTES + B_TILE_BYTES, &mapBl);
                    }
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int w = blockIdx.x; w < num_tiles; w += gridDim.x) {
                const int t = w / p.splits, sp = w % p.splits;
                const int n0 = (t % p.tiles_n) * BN;
                const int kb_lo = p.tri ? n0 / BK : sp * num_kb / p.splits, kb_hi = (sp + 1) * num_kb / p.splits;
                int n_eff = p.N - n0; if (n_eff > BN) n_eff = BN;
                n_eff = (n_eff + 15) & ~15;
                const uint32_t idesc = make_idesc(n_eff, A_MN, B_MN);
                mbar_wait(smem_u32(&acc_empty[acc]), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb_lo; kb < kb_hi; ++kb) {
                    mbar_wait(smem_u32(&full[stage]), phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * C::kStageBytes);
                    const uint64_t dAh = make_smem_desc<A_MN>(st);
                    const uint64_t dBh = make_smem_desc<B_MN>(st + A_TILE_BYTES);
                    const uint64_t dAl = make_smem_desc<A_MN>(st + A_TILE_BYTES + B_TILE_BYTES);
                    const uint64_t dBl = make_smem_desc<B_MN>(st + 2 * A_TILE_BYTES + B_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // K step of 16: 32 B inside the swizzle row (K-major) or 16 rows of 128 B (MN-major), 16-byte units
                        const uint64_t advA = (uint64_t)((A_MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
                        const uint64_t advB = (uint64_t)((B_MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
                        if (PASSES == 3) {
                            umma_f16(d_tmem, dAh + advA, dBl + advB, idesc, ((kb - kb_lo) | k) != 0);
                            umma_f16(d_tmem, dAl + advA, dBh + advB, idesc, 1);
                            umma_f16(d_tmem, dAh + advA, dBh + advB, idesc, 1);
                        } else {
                            umma_f16(d_tmem, dAh + advA, dBh + advB, idesc, ((kb - kb_lo) | k) != 0);
                        }
                    }
                    umma_commit(smem_u32(&empty[stage]));                        // frees the smem stage when the MMAs retire
                    if (kb == kb_hi - 1) umma_commit(smem_u32(&acc_full[acc])); // accumulator complete
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue (warps 2..5) =================
        // TMEM hands each thread one accumulator ROW (32 columns per tcgen05.ld).  Row-per-thread
        // global accesses would touch 32 different rows per instruction, so every 32x32 chunk is
        // transposed through a warp-private smem tile and re-read as (4 rows x 8 float4 columns)
        // per instruction: every global load/store of a warp then covers whole 128-byte row segments.
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int et = threadIdx.x - 64;        // 0..127
        float* stg = s_stage + q * (32 * STG_LD);
        const int rsub = lane >> 3;             // 0..3  row inside a group of 4
        const int c4 = (lane & 7) * 4;          // first of this lane's 4 columns inside the 32-column chunk
        int acc = 0; uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < num_tiles; w += gridDim.x) {
            const int t = w / p.splits, sp = w % p.splits;
            const int m0 = (t / p.tiles_n) * BM, tn = t % p.tiles_n, n0 = tn * BN;
            // stage the per-column epilogue vectors of this tile
            asm volatile("bar.sync 1, 128;");
            for (int c = et; c < BN; c += 128) {
                const int gc = n0 + c;
                const bool ok = gc < p.N;
                s_mul[c] = e.acc_scale * ((e.col_scale && ok) ? e.col_scale[gc] : 1.f);
                s_bias[c] = (e.bias && ok && sp == 0) ? e.bias[gc] : 0.f;
                s_sc[c] = (e.bn_scale && ok) ? e.bn_scale[gc] : 1.f;
                s_sh[c] = (e.bn_scale && ok) ? e.bn_shift[gc] : 0.f;
            }
            asm volatile("bar.sync 1, 128;");
            mbar_wait(smem_u32(&acc_full[acc]), acc_phase);
            tc_fence_after();
            const int row_base = m0 + q * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            float sq[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) sq[i] = 0.f;
            int n_cols = p.N - n0; if (n_cols > BN) n_cols = BN;          // valid columns of this tile
            int w_cols = e.y_cols - n0; if (w_cols > BN) w_cols = BN;      // activation columns to write (zero padded)
            int dw_cols = e.ref ? e.d_cols - n0 : 0; if (dw_cols > BN) dw_cols = BN;   // diff columns to write
            const int c_end = (max(max(n_cols, w_cols), dw_cols) + 31) & ~31;
            for (int c0 = 0; c0 < c_end && c0 < BN; c0 += 32) {
                // issue this chunk's reference loads first: 8 independent 16-byte loads per lane in
                // flight while the accumulator chunk is fetched from TMEM and transposed
                float4 rf[8];
                if (e.ref) {
                    const int ccp = c0 + c4;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = row_base + it * 4 + rsub;
                        rf[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (r < p.M && ccp + 3 < n_cols)
                            rf[it] = __ldg(reinterpret_cast<const float4*>(e.ref + (size_t)r * e.ldref + n0 + ccp));
                    }
                }
                {
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    float4* wr = reinterpret_cast<float4*>(stg + lane * STG_LD);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        wr[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                            __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
                __syncwarp();
                if (e.plain) {
                    // plain store mode (rows of Y need not be 16-byte aligned: parameter-gradient tensors [N, K]):
                    // lane <-> column, one coalesced 128-byte row segment per instruction
                    const int c = c0 + lane;
                    const int gc = n0 + c;
                    if (gc < p.N) {
                        const float mulc = s_mul[c], biac = s_bias[c];
                        if (p.splits > 1) {     // split-K: partial products accumulate into the zeroed output
#pragma unroll 8
                            for (int rl = 0; rl < 32; ++rl) {
                                const int r = row_base + rl;
                                if (r < p.M) atomicAdd(e.Y + (size_t)r * e.ldy + gc, fmaf(stg[rl * STG_LD + lane], mulc, biac));
                            }
                        } else {
#pragma unroll 8
                            for (int rl = 0; rl < 32; ++rl) {
                                const int r = row_base + rl;
                                if (r < p.M) e.Y[(size_t)r * e.ldy + gc] = fmaf(stg[rl * STG_LD + lane], mulc, biac);
                            }
                        }
                    }
                    __syncwarp();
                    continue;
                }
                const int cc = c0 + c4;                 // tile-local column of this lane's float4
                const float4 mul = *reinterpret_cast<const float4*>(s_mul + cc);
                const float4 bia = *reinterpret_cast<const float4*>(s_bias + cc);
                const float4 sc = *reinterpret_cast<const float4*>(s_sc + cc);
                const float4 sh = *reinterpret_cast<const float4*>(s_sh + cc);
                const bool k0 = cc < n_cols, k1 = cc + 1 < n_cols, k2 = cc + 2 < n_cols, k3 = cc + 3 < n_cols;
                const bool wy = cc < w_cols, wd = cc < dw_cols;   // widths are multiples of 4 (padded to 64)
                const size_t gcol = (size_t)n0 + cc;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rl = it * 4 + rsub;
                    const int r = row_base + rl;
                    const float4 a = *reinterpret_cast<const float4*>(stg + rl * STG_LD + c4);
                    float x0 = fmaf(a.x, mul.x, bia.x), x1 = fmaf(a.y, mul.y, bia.y);
                    float x2 = fmaf(a.z, mul.z, bia.z), x3 = fmaf(a.w, mul.w, bia.w);
                    if (r < p.M) {
                        if (e.pre && cc < e.ldpre - n0)     // train: pre-activation, zero padded to ldpre columns
                            *reinterpret_cast<float4*>(e.pre + (size_t)r * e.ldpre + gcol) =
                                make_float4(k0 ? x0 : 0.f, k1 ? x1 : 0.f, k2 ? x2 : 0.f, k3 ? x3 : 0.f);
                        if (e.bn_scale) {
                            x0 = x0 > 0.f ? x0 : x0 * e.slope; x1 = x1 > 0.f ? x1 : x1 * e.slope;
                            x2 = x2 > 0.f ? x2 : x2 * e.slope; x3 = x3 > 0.f ? x3 : x3 * e.slope;
                            x0 = fmaf(x0, sc.x, sh.x); x1 = fmaf(x1, sc.y, sh.y);
                            x2 = fmaf(x2, sc.z, sh.z); x3 = fmaf(x3, sc.w, sh.w);
                        }
                        x0 = k0 ? x0 : 0.f; x1 = k1 ? x1 : 0.f; x2 = k2 ? x2 : 0.f; x3 = k3 ? x3 : 0.f;
                        if (e.Y && wy) *reinterpret_cast<float4*>(e.Y + (size_t)r * e.ldy + gcol) = make_float4(x0, x1, x2, x3);
                        if (e.Yh && wy) {
                            const float ys = e.y_split_scale;
                            const float y0 = x0 * ys, y1 = x1 * ys, y2 = x2 * ys, y3 = x3 * ys;
                            const __half2 h01 = __floats2half2_rn(y0, y1), h23 = __floats2half2_rn(y2, y3);
                            uint2 hv;
                            hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                            *reinterpret_cast<uint2*>(e.Yh + (size_t)r * e.ldh + gcol) = hv;
                            if (e.Yl) {
                                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                                const __half2 l01 = __floats2half2_rn(y0 - f01.x, y1 - f01.y);
                                const __half2 l23 = __floats2half2_rn(y2 - f23.x, y3 - f23.y);
                                uint2 lv;
                                lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                                *reinterpret_cast<uint2*>(e.Yl + (size_t)r * e.ldh + gcol) = lv;
                            }
                        }
                        if (e.ref) {
                            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
                            if (k3) {
                                const float4 f = rf[it];
                                d0 = x0 - f.x; d1 = x1 - f.y; d2 = x2 - f.z; d3 = x3 - f.w;
                            } else if (k0) {
                                const float* rp = e.ref + (size_t)r * e.ldref + gcol;
                                d0 = x0 - rp[0]; if (k1) d1 = x1 - rp[1]; if (k2) d2 = x2 - rp[2];
                            }
                            sq[it] = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, sq[it]))));
                            if (e.dout && wd) *reinterpret_cast<float4*>(e.dout + (size_t)r * e.lddout + gcol) = make_float4(d0, d1, d2, d3);
                            if (e.Dh && wd) {
                                const float s0 = d0 * e.d_scale, s1 = d1 * e.d_scale, s2 = d2 * e.d_scale, s3 = d3 * e.d_scale;
                                const __half2 h01 = __floats2half2_rn(s0, s1), h23 = __floats2half2_rn(s2, s3);
                                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                                const __half2 l01 = __floats2half2_rn(s0 - f01.x, s1 - f01.y);
                                const __half2 l23 = __floats2half2_rn(s2 - f23.x, s3 - f23.y);
                                uint2 hv, lv;
                                hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                                lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                                *reinterpret_cast<uint2*>(e.Dh + (size_t)r * e.lddh + gcol) = hv;
                                *reinterpret_cast<uint2*>(e.Dl + (size_t)r * e.lddh + gcol) = lv;
                            }
                        } else if (e.sq_self) {
                            sq[it] = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, sq[it]))));
                        }
                    }
                }
                __syncwarp();     // the staging tile is rewritten by the next chunk
            }
            // accumulator drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[acc]));
            if (e.rowpart) {
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    float v = sq[it];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    const int r = row_base + it * 4 + rsub;
                    if ((lane & 7) == 0 && r < p.M) e.rowpart[(size_t)tn * e.rowpart_stride + r] = v;
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_tc_state = -1;   // -1 unknown, 0 unavailable, 1 ok
int g_num_sms = 148;

int init_tc() {
    if (g_tc_state >= 0) return g_tc_state;
    g_tc_state = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (prop.major != 10) return 0;
    g_num_sms = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        cudaGetLastError();
        return 0;
    }
    g_encode = (EncodeTiledFn)fn;
    bool ok = true;
    auto attr = [&](auto kern, int bytes) {
        ok = ok && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
    };
#define MMAD_TC_ATTR(P, AM, BMN)                                             \
    attr(gemm_tc_kernel<P, AM, BMN, 256>, Cfg<P, 256>::kSmemBytes);           \
    attr(gemm_tc_kernel<P, AM, BMN, 128>, Cfg<P, 128>::kSmemBytes)
    MMAD_TC_ATTR(3, false, false); MMAD_TC_ATTR(1, false, false);
    MMAD_TC_ATTR(3, false, true);  MMAD_TC_ATTR(1, false, true);
    MMAD_TC_ATTR(3, true, true);   MMAD_TC_ATTR(1, true, true);
#undef MMAD_TC_ATTR
    if (!ok) {
        cudaGetLastError();
        return 0;
    }
    g_tc_state = 1;
    return 1;
}

}  // namespace

int tc_available() { return init_tc(); }
int gemm_tc_tile_n() { return BN_MAX; }

// Tile width for an M x N problem: 256 unless that leaves SMs idle, then 128 (twice the CTAs, deeper pipeline).
int gemm_tc_pick_bn(int M, int N) {
    init_tc();
    const int tiles256 = ((M + BM - 1) / BM) * ((N + 255) / 256);
    return tiles256 >= g_num_sms ? 256 : 128;
}

// 2-D map over a row-major fp16 matrix [rows, k] with row stride ld (elements): box = [64 x box_rows],
// 128-byte swizzle, zero fill outside [rows, k].  K-major operands: rows = M or N, k = contraction, box_rows =
// 128 (A) / 256 (B).  MN-major operands: rows = contraction extent, k = M or N extent, box_rows = 64.
int tc_make_operand_map(CUtensorMap* map, const __half* base, int rows, int k, int ld, int box_rows) {
    if (!init_tc()) { set_error("tcgen05 path unavailable on this device"); return MMAD_E_UNSUPPORTED; }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 8)) { set_error("TMA operand must be 16-byte aligned (ld=%d)", ld); return MMAD_E_ARG; }
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%d k=%d ld=%d", (int)r, rows, k, ld); return MMAD_E_CUDA; }
    return MMAD_OK;
}

int gemm_tc(const TcOperand& A, const TcOperand& B, int M, int N, int K, int passes, const Epilogue& e, cudaStream_t s, int bn) {
    if (!init_tc()) { set_error("tcgen05 path unavailable on this device"); return MMAD_E_UNSUPPORTED; }
    if (M <= 0 || N <= 0) return MMAD_OK;
    auto al = [](const void* q, int ld, int ldm) { return q == nullptr || (((reinterpret_cast<uintptr_t>(q) & 15) == 0) && ld % ldm == 0); };
    if (!(e.plain || al(e.Y, e.ldy, 4)) || !al(e.pre, e.ldpre, 4) || !al(e.Yh, e.ldh, 8) || !al(e.Yl, e.ldh, 8) || !al(e.ref, e.ldref, 4) || !al(e.dout, e.lddout, 4) ||
        !al(e.Dh, e.lddh, 8) || !al(e.Dl, e.lddh, 8) || (!e.plain && (e.y_cols % 4)) || (e.d_cols % 4)) {
        set_error("gemm_tc: epilogue buffers must be 16-byte aligned with padded leading dimensions");
        return MMAD_E_ARG;
    }
    TcParams p;
    p.M = M; p.N = N; p.K = K;
    p.tiles_m = (M + BM - 1) / BM;
    if (bn != 128 && bn != 256) { set_error("gemm_tc: tile width %d not instantiated", bn); return MMAD_E_ARG; }
    p.tiles_n = (N + bn - 1) / bn;
    const int tiles = p.tiles_m * p.tiles_n;
    const int num_kb = (K + BK - 1) / BK;
    p.splits = 1;
    p.tri = e.b_upper_tri ? 1 : 0;
    if (!p.tri && e.plain && e.split_k_ok && tiles * 2 <= g_num_sms) {
        p.splits = g_num_sms / tiles;
        if (p.splits > num_kb) p.splits = num_kb;
        if (p.splits < 1) p.splits = 1;
    }
    if (p.splits > 1)    // partial sums are accumulated atomically: the output starts at zero
        MMAD_CUDA_OK(cudaMemset2DAsync(e.Y, (size_t)e.ldy * 4, 0, (size_t)N * 4, M, s));
    const int items = tiles * p.splits;
    const int grid = items < g_num_sms ? items : g_num_sms;
    if (A.mn && !B.mn) { set_error("gemm_tc: MN-major A with K-major B is not instantiated"); return MMAD_E_UNSUPPORTED; }
#define MMAD_TC_LAUNCH2(P, AM, BMN, BNV)                                                                              \
    gemm_tc_kernel<P, AM, BMN, BNV><<<grid, NTHREADS, Cfg<P, BNV>::kSmemBytes, s>>>(A.hi, P == 3 ? A.lo : A.hi, B.hi, \
                                                                                   P == 3 ? B.lo : B.hi, p, e)
#define MMAD_TC_LAUNCH(P, AM, BMN)                  \
    do {                                            \
        if (bn == 256) MMAD_TC_LAUNCH2(P, AM, BMN, 256); \
        else MMAD_TC_LAUNCH2(P, AM, BMN, 128);      \
    } while (0)
    if (passes == 3) {
        if (A.mn) MMAD_TC_LAUNCH(3, true, true);
        else if (B.mn) MMAD_TC_LAUNCH(3, false, true);
        else MMAD_TC_LAUNCH(3, false, false);
    } else {
        if (A.mn) MMAD_TC_LAUNCH(1, true, true);
        else if (B.mn) MMAD_TC_LAUNCH(1, false, true);
        else MMAD_TC_LAUNCH(1, false, false);
    }
#undef MMAD_TC_LAUNCH2
#undef MMAD_TC_LAUNCH
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

}  // namespace mmad
