// tcgen05 fused layer GEMM on CTA PAIRS (cta_group::2) -- the bulk-scoring kernel.
//
// Same math and epilogue as gemm_tc.cu, but one 256 x 256 output tile is computed by two CTAs on adjacent
// SMs (a cluster of 2): each CTA stages ITS 128 rows of A and ITS half of the B tile (128 of the 256 weight
// rows), one elected thread of the even CTA issues tcgen05.mma.cta_group::2 (M = 256), and each CTA keeps its
// 128 accumulator rows in its own TMEM and runs its own epilogue.
//
// Why: in the f16x3 mode every k-block needs the hi AND lo halves of both operands.  With one CTA per tile
// that is 96 KB of shared memory per stage -- two stages, one load in flight, and a load latency of ~1 us
// against ~1 us of MMA work per k-block leaves the tensor pipe ~60-80 % busy (profiles/r1_ncu_full_gemm_tc_v3.md).
// A CTA pair needs 64 KB per CTA per stage: three stages, two loads in flight, and a third less L2->SMEM
// traffic per FLOP because B is not duplicated.
//
// Barrier protocol (s = stage, a = accumulator buffer):
//   full[s]      in the EVEN CTA: armed by its producer with expect_tx of both CTAs' bytes; both producers' TMA
//                loads complete on it (cta_group::2 loads address the even CTA's barrier)
//   empty[s]     in each CTA: tcgen05.commit.multicast from the MMA thread frees the stage in both CTAs
//   acc_full[a]  in each CTA: commit.multicast after the last k-block
//   acc_empty[a] in the EVEN CTA: 16 arrivals (8 epilogue warps x 2 CTAs; the odd CTA arrives remotely)
#include <stdlib.h>

#include "gemm_tc_common.cuh"

namespace mmad {

using namespace tc;

namespace {

constexpr int BN2 = 256;                       // tile width; each CTA holds BN2/2 rows of B
constexpr int B_HALF_BYTES = (BN2 / 2) * BK * 2;   // 16 KB
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // clears the CTA-rank bit of a shared::cluster address -> even CTA

// PASSES: 3 = hi*lo + lo*hi + hi*hi (fp32-equivalent products); 2 = lo*hi + hi*hi (A exact, B rounded to fp16:
// the NAP rotation, whose B rows are unit-max whitening vectors); 1 = hi*hi;
// 4 = hi*hi in fp16 + one fp8 pass over the twins [lo8 | a8] . [Wh8 ; Wl8] (MMAD_PREC_F16F8): per k-block four
// kind::f16 and four kind::f8f6f4 instructions into the same fp32 accumulator, against twelve kind::f16 in mode 3
template <int PASSES> struct Cfg2 {
    static constexpr int kStageBytes = (PASSES >= 2 ? 2 : 1) * A_TILE_BYTES + (PASSES >= 3 ? 2 : 1) * B_HALF_BYTES;   // 64 / 48 / 32 KB per CTA
    static constexpr int kStages = (192 * 1024) / kStageBytes > 6 ? 6 : (192 * 1024) / kStageBytes;   // 3 / 4 / 6
    static constexpr int kSmemTiles = kStages * kStageBytes;
    static constexpr int kSmemBytes = kSmemTiles + 4 * BN2 * 4 + EPI_WARPS * STG_FLOAT4 * 16 + 256 + 1024;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// L2 eviction priority of the load: the B operand (weights / the NAP whitening factor, re-read by every row tile) is kept,
// the A operand (activation rows, dead once the band of n-tiles over them is done) is evicted first.  Without the hint the
// 75 776-row A stream pushes the 122 MB NAP factor out of the 126 MB L2 and it is re-streamed from HBM (12.9 GB per launch,
// profiles/r1_ncu_full_f16f8.md) -- bandwidth that is free, but power that the capped SM clock pays for.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_even, int c0, int c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(bar_even), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f8_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // arrives on the barrier at this offset in BOTH CTAs
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_even(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
// instruction descriptor: fp16 x fp16 -> fp32, K-major A and B, M = 256 (pair), N = n
__device__ __forceinline__ uint32_t make_idesc_pair(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

struct Tc2Params {
    int M, N, K;
    int tiles_m, tiles_n;     // tiles of 256 x 256
    int tri;
    int band;                 // n-tiles per band of the tile order (see tile_coords)
    int splits;               // split-K factor (plain epilogue): work item = (tile, split), atomically accumulated
};

// Tile order: bands of `band` n-tiles, all m inside a band, n fastest.  The pairs running at any moment then cover
// ~74/band row panels of A and `band` row panels of B: with band = 8 that is ~100 MB for the NAP rotation
// (5.7 MB per 256-row fp16 hi+lo panel) and fits the 126 MB L2, so B (the 122 MB whitening factor) is read from HBM
// once per band instead of once per few row tiles.  Narrow layers (tiles_n <= band) keep the plain n-fastest order.
__device__ __forceinline__ void tile_coords(const Tc2Params& p, int t, int& tm, int& tn) {
    const int full = p.tiles_m * p.band;
    const int b = t / full;
    const int rem = t - b * full;
    int w = p.tiles_n - b * p.band;
    if (w > p.band) w = p.band;
    tm = rem / w;
    tn = b * p.band + rem % w;
}

// k-blocks accumulated into the tile at column n0, split sp (the producer's / issuer's kb_lo..kb_hi range)
__device__ __forceinline__ int tile_kb(const Tc2Params& p, int num_kb, int n0, int sp) {
    constexpr int BN = BN2;
    const int kb_lo = ((n0 + BN < p.N ? n0 + BN : p.N) <= p.tri) ? n0 / BK : sp * num_kb / p.splits, kb_hi = (sp + 1) * num_kb / p.splits;
    return kb_hi - kb_lo;
}

// A_MN / B_MN: the operand is stored [contraction, M or N] (MN-major smem descriptors, TMA boxes of 64 x 64) -- the
// training GEMMs dW = g_pre^T . in (both) and dX = g_pre . W (B) on the same row-major tensors as the forward pass.
template <int PASSES, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                Tc2Params p, Epilogue e) {
    using C = Cfg2<PASSES>;
    constexpr int BN = BN2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* s_mul = reinterpret_cast<float*>(smem + C::kSmemTiles);
    float* s_bias = s_mul + BN;
    float* s_sc = s_bias + BN;
    float* s_sh = s_sc + BN;
    float4* s_stage = reinterpret_cast<float4*>(s_sh + BN);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + EPI_WARPS * STG_FLOAT4);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::kStages;
    uint64_t* acc_full = empty + C::kStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int num_kb = (p.K + BK - 1) / BK;
    const int num_tiles = p.tiles_m * p.tiles_n * p.splits;      // work items: tile-major, split fastest

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::kStages; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), 2 * EPI_WARPS); }
        fence_barrier_init();
        tma_prefetch_desc(&mapAh); tma_prefetch_desc(&mapBh);
        if (PASSES >= 2) tma_prefetch_desc(&mapAl);
        if (PASSES >= 3) tma_prefetch_desc(&mapBl);
    }
    if (warp == 1) {   // the same logical warp of both CTAs allocates (and later frees) the pair's TMEM
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer's barriers are initialised before anything is signalled remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint64_t pol_a = l2_policy_evict_normal(), pol_b = l2_policy_evict_last();
            for (int w = pair; w < num_tiles; w += npairs) {
                const int t = w / p.splits, sp = w % p.splits;
                int tm, tn;
                tile_coords(p, t, tm, tn);
                const int n0 = tn * BN;
                int n_eff = p.N - n0; if (n_eff > BN) n_eff = BN;
                n_eff = (n_eff + 15) & ~15;
                const int m0 = tm * 256 + (int)rank * BM;                     // this CTA's 128 rows of A
                const int nb0 = n0 + (int)rank * (n_eff >> 1);                // this CTA's half of the B rows
                const int kb_lo = ((n0 + BN < p.N ? n0 + BN : p.N) <= p.tri) ? n0 / BK : sp * num_kb / p.splits, kb_hi = (sp + 1) * num_kb / p.splits;
                for (int kb = kb_lo; kb < kb_hi; ++kb) {
                    mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                    if (rank == 0) mbar_expect_tx(smem_u32(&full[stage]), 2 * C::kStageBytes);
                    const uint32_t fb = smem_u32(&full[stage]) & kPeerBitMask;
                    uint8_t* st = smem + stage * C::kStageBytes;
                    auto load_a = [&](uint8_t* dst, const CUtensorMap* map) {
                        if (A_MN) {   // [k rows, m contiguous]: two boxes of 64 m x 64 k
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j) tma_load_2d_pair(smem_u32(dst + j * 8192), map, fb, m0 + j * 64, kb * BK, pol_a);
                        } else {
                            tma_load_2d_pair(smem_u32(dst), map, fb, kb * BK, m0, pol_a);
                        }
                    };
                    auto load_b = [&](uint8_t* dst, const CUtensorMap* map) {
                        if (B_MN) {
#pragma unroll
                            for (int j = 0; j < BN / 128; ++j) tma_load_2d_pair(smem_u32(dst + j * 8192), map, fb, nb0 + j * 64, kb * BK, pol_b);
                        } else {
                            tma_load_2d_pair(smem_u32(dst), map, fb, kb * BK, nb0, pol_b);
                        }
                    };
                    load_a(st, &mapAh);
                    load_b(st + A_TILE_BYTES, &mapBh);
                    if (PASSES == 4) {   // fp8 twins: byte maps, one 128-byte block per k-block
                        tma_load_2d_pair(smem_u32(st + A_TILE_BYTES + B_HALF_BYTES), &mapAl, fb, kb * 128, m0, pol_a);
                        tma_load_2d_pair(smem_u32(st + 2 * A_TILE_BYTES + B_HALF_BYTES), &mapBl, fb, kb * 128, nb0, pol_b);
                    } else {
                        if (PASSES >= 2) load_a(st + A_TILE_BYTES + B_HALF_BYTES, &mapAl);
                        if (PASSES == 3) load_b(st + 2 * A_TILE_BYTES + B_HALF_BYTES, &mapBl);
                    }
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (even CTA only) =================
        if (lane == 0 && rank == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int w = pair; w < num_tiles; w += npairs) {
                const int t = w / p.splits, sp = w % p.splits;
                int tm, tn;
                tile_coords(p, t, tm, tn);
                const int n0 = tn * BN;
                const int kb_lo = ((n0 + BN < p.N ? n0 + BN : p.N) <= p.tri) ? n0 / BK : sp * num_kb / p.splits, kb_hi = (sp + 1) * num_kb / p.splits;
                int n_eff = p.N - n0; if (n_eff > BN) n_eff = BN;
                n_eff = (n_eff + 15) & ~15;
                const uint32_t idesc = make_idesc_pair(n_eff) | ((uint32_t)A_MN << 15) | ((uint32_t)B_MN << 16);
                mbar_wait(smem_u32(&acc_empty[acc]), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb_lo; kb < kb_hi; ++kb) {
                    mbar_wait(smem_u32(&full[stage]), phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * C::kStageBytes);
                    const uint64_t dAh = make_smem_desc<A_MN>(st);
                    const uint64_t dBh = make_smem_desc<B_MN>(st + A_TILE_BYTES);
                    const uint64_t dAl = make_smem_desc<A_MN>(st + A_TILE_BYTES + B_HALF_BYTES);
                    const uint64_t dBl = make_smem_desc<B_MN>(st + 2 * A_TILE_BYTES + B_HALF_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t advA = (uint64_t)((A_MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
                        const uint64_t advB = (uint64_t)((B_MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
                        if (PASSES == 4) {
                            umma_f8_pair(d_tmem, dAl + advA, dBl + advB, idesc, ((kb - kb_lo) | k) != 0);
                            umma_f16_pair(d_tmem, dAh + advA, dBh + advB, idesc, 1);
                        } else if (PASSES == 3) {
                            umma_f16_pair(d_tmem, dAh + advA, dBl + advB, idesc, ((kb - kb_lo) | k) != 0);
                            umma_f16_pair(d_tmem, dAl + advA, dBh + advB, idesc, 1);
                            umma_f16_pair(d_tmem, dAh + advA, dBh + advB, idesc, 1);
                        } else if (PASSES == 2) {
                            umma_f16_pair(d_tmem, dAl + advA, dBh + advB, idesc, ((kb - kb_lo) | k) != 0);
                            umma_f16_pair(d_tmem, dAh + advA, dBh + advB, idesc, 1);
                        } else {
                            umma_f16_pair(d_tmem, dAh + advA, dBh + advB, idesc, ((kb - kb_lo) | k) != 0);
                        }
                    }
                    umma_commit_pair(smem_u32(&empty[stage]));
                    if (kb == kb_hi - 1) umma_commit_pair(smem_u32(&acc_full[acc]));
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue (warps 2..9 of both CTAs): this CTA's 128 accumulator rows =================
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;       // column half of the tile this warp drains
        const int et = threadIdx.x - 64;
        float4* stg = s_stage + (warp - 2) * STG_FLOAT4;
        const int col_lo = half * 128, col_hi = col_lo + 128;
        int acc = 0; uint32_t acc_phase = 0;
        EpiVec vec;
        if (pair < num_tiles) {      // the first tile's column vectors; later ones are loaded one tile ahead
            int tm, tn;
            tile_coords(p, pair / p.splits, tm, tn);
            vec = epi_vec_load(e, p.N, tn * BN, pair % p.splits, et, tile_kb(p, num_kb, tn * BN, pair % p.splits));
        }
        for (int w = pair; w < num_tiles; w += npairs) {
            const int t = w / p.splits, sp = w % p.splits;
            int tm, tn;
            tile_coords(p, t, tm, tn);
            const int m0 = tm * 256 + (int)rank * BM, n0 = tn * BN;
            asm volatile("bar.sync 1, 256;");          // every warp is done with the previous tile's vectors
            epi_vec_store(vec, et, s_mul, s_bias, s_sc, s_sh);
            asm volatile("bar.sync 1, 256;");
            if (w + npairs < num_tiles) {
                int tm2, tn2;
                tile_coords(p, (w + npairs) / p.splits, tm2, tn2);
                vec = epi_vec_load(e, p.N, tn2 * BN, (w + npairs) % p.splits, et, tile_kb(p, num_kb, tn2 * BN, (w + npairs) % p.splits));
            }
            mbar_wait(smem_u32(&acc_full[acc]), acc_phase);
            tc_fence_after();
            const int row_base = m0 + q * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            float sq[8];
            epi_tile<BN>(e, p.M, p.N, p.splits, sp, row_base, n0, taddr, stg, s_mul, s_bias, s_sc, s_sh, lane, col_lo, col_hi, sq);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_even(smem_u32(&acc_empty[acc]));
            if (n0 + col_lo < p.N) epi_rowpart(e, p.M, row_base, (n0 + col_lo) / ROWPART_COLS, lane, sq);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // both CTAs are done with both TMEMs and no multicast arrival is in flight
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

int g_tc2_state = -1;
int g_sms = 148;

int init_tc2() {
    if (g_tc2_state >= 0) return g_tc2_state;
    g_tc2_state = 0;
    if (!tc_available()) return 0;
    const char* env = getenv("MMAD_NO_PAIR");
    if (env && env[0] == '1') return 0;
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    g_sms = prop.multiProcessorCount;
    bool ok = true;
    auto attr = [&](auto kern, int bytes) {
        ok = ok && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
    };
    attr(gemm_tc2_kernel<3, false, false>, Cfg2<3>::kSmemBytes);
    attr(gemm_tc2_kernel<4, false, false>, Cfg2<4>::kSmemBytes);
    attr(gemm_tc2_kernel<2, false, false>, Cfg2<2>::kSmemBytes);
    attr(gemm_tc2_kernel<1, false, false>, Cfg2<1>::kSmemBytes);
    attr(gemm_tc2_kernel<3, false, true>, Cfg2<3>::kSmemBytes);
    attr(gemm_tc2_kernel<1, false, true>, Cfg2<1>::kSmemBytes);
    attr(gemm_tc2_kernel<3, true, true>, Cfg2<3>::kSmemBytes);
    attr(gemm_tc2_kernel<1, true, true>, Cfg2<1>::kSmemBytes);
    if (!ok) { cudaGetLastError(); return 0; }
    g_tc2_state = 1;
    return 1;
}

}  // namespace

int tc2_available() { return init_tc2(); }

// A maps: box 64 x 128 (rows of A).  B maps: box 64 x 128 (HALF of the 256-wide tile; tc_make_operand_map(.., 128)).
int gemm_tc2(const TcOperand& A, const TcOperand& B, int M, int N, int K, int passes, const Epilogue& e, cudaStream_t s) {
    if (!init_tc2()) { set_error("cta_group::2 GEMM unavailable"); return MMAD_E_UNSUPPORTED; }
    if (M <= 0 || N <= 0) return MMAD_OK;
    if (A.mn && !B.mn) { set_error("gemm_tc2: MN-major A with K-major B is not instantiated"); return MMAD_E_UNSUPPORTED; }
    if ((A.mn || B.mn) && (passes == 2 || passes == 4)) { set_error("gemm_tc2: the 2-pass and fp8-assisted modes are instantiated for K-major operands only"); return MMAD_E_UNSUPPORTED; }
    auto al = [](const void* q, int ld, int ldm) { return q == nullptr || (((reinterpret_cast<uintptr_t>(q) & 15) == 0) && ld % ldm == 0); };
    if (!(e.plain || al(e.Y, e.ldy, 4)) || !al(e.pre, e.ldpre, 4) || !al(e.Yh, e.ldh, 8) || !al(e.Yl, e.ldh, 8) || !al(e.ref, e.ldref, 4) ||
        !al(e.dout, e.lddout, 4) || !al(e.Dh, e.lddh, 8) || !al(e.Dl, e.lddh, 8) || (!e.plain && (e.y_cols % 4)) || (e.d_cols % 4)) {
        set_error("gemm_tc2: epilogue buffers must be 16-byte aligned with padded leading dimensions");
        return MMAD_E_ARG;
    }
    Tc2Params p;
    p.M = M; p.N = N; p.K = K;
    p.tiles_m = (M + 255) / 256;
    p.tiles_n = (N + BN2 - 1) / BN2;
    p.tri = e.b_upper_tri;     // leading rows of B that are upper triangular (tiles entirely inside skip k-blocks left of them)
    // balanced bands of at most 12 n-tiles: the NAP factor's 22 n-tiles run as 2 x 11 (A is streamed twice, the 61 MB band
    // of B plus the ~7 row panels in flight stay L2 resident); MMAD_TC2_BAND overrides for experiments
    p.band = (p.tiles_n + (p.tiles_n + 11) / 12 - 1) / ((p.tiles_n + 11) / 12);
    { static int b = -1; if (b < 0) { const char* q = getenv("MMAD_TC2_BAND"); b = q ? atoi(q) : 0; } if (b > 0) p.band = b; }
    const int tiles = p.tiles_m * p.tiles_n;
    const int max_pairs = g_sms / 2;
    const int num_kb = (K + BK - 1) / BK;
    p.splits = 1;
    if (!p.tri && e.plain && e.split_k_ok) {   // split-K when it fills the pairs better (see gemm_tc)
        long best = (long)((tiles + max_pairs - 1) / max_pairs) * num_kb;
        for (int sp = 2; sp <= 16 && sp <= num_kb; ++sp) {
            const long waves = ((long)tiles * sp + max_pairs - 1) / max_pairs;
            const long cost = waves * ((num_kb + sp - 1) / sp) + waves;
            if (cost * 10 < best * 9) { best = cost; p.splits = sp; }
        }
    }
    if (p.splits > 1 && !e.pre_zeroed)
        MMAD_CUDA_OK(cudaMemset2DAsync(e.Y, (size_t)e.ldy * 4, 0, (size_t)N * 4, M, s));
    const int items = tiles * p.splits;
    const int grid = 2 * (items < max_pairs ? items : max_pairs);
#define MMAD_TC2_LAUNCH(P, AM, BMN) \
    gemm_tc2_kernel<P, AM, BMN><<<grid, NTHREADS, Cfg2<P>::kSmemBytes, s>>>(A.hi, P >= 2 ? A.lo : A.hi, B.hi, P >= 3 ? B.lo : B.hi, p, e)
    if (A.mn) { if (passes == 3) MMAD_TC2_LAUNCH(3, true, true); else MMAD_TC2_LAUNCH(1, true, true); }
    else if (B.mn) { if (passes == 3) MMAD_TC2_LAUNCH(3, false, true); else MMAD_TC2_LAUNCH(1, false, true); }
    else if (passes == 4) MMAD_TC2_LAUNCH(4, false, false);
    else if (passes == 3) MMAD_TC2_LAUNCH(3, false, false);
    else if (passes == 2) MMAD_TC2_LAUNCH(2, false, false);
    else MMAD_TC2_LAUNCH(1, false, false);
#undef MMAD_TC2_LAUNCH
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

}  // namespace mmad
