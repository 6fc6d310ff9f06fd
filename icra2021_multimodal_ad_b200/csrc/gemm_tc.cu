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

#include "gemm_tc_common.cuh"

namespace mmad {

using namespace tc;

namespace {

struct TcParams {
    int M, N, K;
    int tiles_m, tiles_n;
    int tri;             // leading rows of B that are upper triangular (B[n,k] = 0 for k < n): tiles inside skip the k-blocks left of their first row
    int splits;          // split-K factor (plain epilogue only): work item = (tile, split), atomically accumulated
    int dbg;             // MMAD_TC_DEBUG=1: CTA 0 leaves %globaltimer stamps in g_tc_stamps
};

__device__ unsigned long long g_tc_stamps[16];       // 0-7: phases of CTA 0; 8-15: end of the first tile's epilogue per epilogue warp
__device__ __forceinline__ void tc_stamp(const TcParams& p, int i) {
    if (p.dbg && blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_tc_stamps[i] = t;
    }
}

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
    float4* s_stage = reinterpret_cast<float4*>(s_sh + BN);      // 8 warps x 32 x 16 floats, XOR swizzled
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + EPI_WARPS * STG_FLOAT4);
    uint64_t* full = bars;                       // [kStages]  TMA -> MMA
    uint64_t* empty = bars + C::kStages;         // [kStages]  MMA -> TMA
    uint64_t* acc_full = empty + C::kStages;     // [2]        MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;          // [2]        epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    pdl_trigger();
    if (threadIdx.x == 0) tc_stamp(p, 0);
    const int num_kb = (p.K + BK - 1) / BK;
    const int num_tiles = p.tiles_m * p.tiles_n * p.splits;     // work items: tile-major, split fastest
    // k-block range of split sp: [sp * num_kb / splits, (sp + 1) * num_kb / splits)  (host guarantees splits <= num_kb)
    auto tile_kb = [&](int n0, int sp) {
        const int kb_lo = ((n0 + BN < p.N ? n0 + BN : p.N) <= p.tri) ? n0 / BK : sp * num_kb / p.splits, kb_hi = (sp + 1) * num_kb / p.splits;
        return kb_hi - kb_lo;
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::kStages; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), EPI_WARPS); }
        fence_barrier_init();
        tma_prefetch_desc(&mapAh); tma_prefetch_desc(&mapBh);
        if (PASSES >= 2) tma_prefetch_desc(&mapAl);
        if (PASSES >= 3) tma_prefetch_desc(&mapBl);
    }
    if (warp == 1) {   // TMEM allocation is a warp-wide operation; the same warp frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();          // everything above ran under the previous kernel's tail (PDL launches); global memory from here on
    if (threadIdx.x == 0) tc_stamp(p, 1);

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int w = blockIdx.x; w < num_tiles; w += gridDim.x) {
                const int t = w / p.splits, sp = w % p.splits;
                const int m0 = (t / p.tiles_n) * BM, n0 = (t % p.tiles_n) * BN;
                const int kb_lo = ((n0 + BN < p.N ? n0 + BN : p.N) <= p.tri) ? n0 / BK : sp * num_kb / p.splits, kb_hi = (sp + 1) * num_kb / p.splits;
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
                    if (PASSES == 4) {   // fp8 twins: byte maps, one 128-byte block per k-block
                        tma_load_2d(smem_u32(st + A_TILE_BYTES + B_TILE_BYTES), &mapAl, fb, kb * 128, m0);
                        tma_load_2d(smem_u32(st + 2 * A_TILE_BYTES + B_TILE_BYTES), &mapBl, fb, kb * 128, n0);
                    } else {
                        if (PASSES >= 2) load_a(st + A_TILE_BYTES + B_TILE_BYTES, &mapAl);
                        if (PASSES == 3) load_b(st + 2 * A_TILE_BYTES + B_TILE_BYTES, &mapBl);
                    }
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                    if (w == (int)blockIdx.x && kb == kb_lo) tc_stamp(p, 2);
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
                const int kb_lo = ((n0 + BN < p.N ? n0 + BN : p.N) <= p.tri) ? n0 / BK : sp * num_kb / p.splits, kb_hi = (sp + 1) * num_kb / p.splits;
                int n_eff = p.N - n0; if (n_eff > BN) n_eff = BN;
                n_eff = (n_eff + 15) & ~15;
                const uint32_t idesc = make_idesc(n_eff, A_MN, B_MN);
                mbar_wait(smem_u32(&acc_empty[acc]), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb_lo; kb < kb_hi; ++kb) {
                    mbar_wait(smem_u32(&full[stage]), phase);
                    tc_fence_after();
                    if (w == (int)blockIdx.x && kb == kb_lo) tc_stamp(p, 3);
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
                        if (PASSES == 4) {   // the fp8 pass covers 32 bytes of the twin per instruction too
                            umma_f8(d_tmem, dAl + advA, dBl + advB, idesc, ((kb - kb_lo) | k) != 0);
                            umma_f16(d_tmem, dAh + advA, dBh + advB, idesc, 1);
                        } else if (PASSES == 3) {
                            umma_f16(d_tmem, dAh + advA, dBl + advB, idesc, ((kb - kb_lo) | k) != 0);
                            umma_f16(d_tmem, dAl + advA, dBh + advB, idesc, 1);
                            umma_f16(d_tmem, dAh + advA, dBh + advB, idesc, 1);
                        } else if (PASSES == 2) {
                            umma_f16(d_tmem, dAl + advA, dBh + advB, idesc, ((kb - kb_lo) | k) != 0);
                            umma_f16(d_tmem, dAh + advA, dBh + advB, idesc, 1);
                        } else {
                            umma_f16(d_tmem, dAh + advA, dBh + advB, idesc, ((kb - kb_lo) | k) != 0);
                        }
                    }
                    umma_commit(smem_u32(&empty[stage]));                        // frees the smem stage when the MMAs retire
                    if (kb == kb_hi - 1) { umma_commit(smem_u32(&acc_full[acc])); if (w == (int)blockIdx.x) tc_stamp(p, 4); }   // accumulator complete
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue (warps 2..9) =================
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;       // column half of the tile this warp drains (BN = 128: half 1 idles)
        const int et = threadIdx.x - 64;        // 0..255
        float4* stg = s_stage + (warp - 2) * STG_FLOAT4;
        // BN = 128: both warps of a lane quarter share the tile's columns 64 / 64 unless the row partial sums are wanted
        // (their slots are 128 columns wide: one writer per slot)
        const bool split128 = BN == 128 && e.rowpart == nullptr;
        const int col_lo = BN == 256 ? half * 128 : (split128 ? half * 64 : 0);
        const int col_hi = BN == 256 ? col_lo + 128 : (split128 ? col_lo + 64 : (half == 0 ? 128 : 0));
        int acc = 0; uint32_t acc_phase = 0;
        EpiVec vec;
        if (et < BN && (int)blockIdx.x < num_tiles)
            vec = epi_vec_load(e, p.N, (((int)blockIdx.x / p.splits) % p.tiles_n) * BN, (int)blockIdx.x % p.splits, et,
                               tile_kb((((int)blockIdx.x / p.splits) % p.tiles_n) * BN, (int)blockIdx.x % p.splits));
        for (int w = blockIdx.x; w < num_tiles; w += gridDim.x) {
            const int t = w / p.splits, sp = w % p.splits;
            const int m0 = (t / p.tiles_n) * BM, tn = t % p.tiles_n, n0 = tn * BN;
            // stage the per-column epilogue vectors of this tile (loaded one tile ahead), load the next tile's
            asm volatile("bar.sync 1, 256;");
            if (et < BN) epi_vec_store(vec, et, s_mul, s_bias, s_sc, s_sh);
            asm volatile("bar.sync 1, 256;");
            if (et < BN && w + (int)gridDim.x < num_tiles) {
                const int w2 = w + (int)gridDim.x;
                vec = epi_vec_load(e, p.N, ((w2 / p.splits) % p.tiles_n) * BN, w2 % p.splits, et, tile_kb(((w2 / p.splits) % p.tiles_n) * BN, w2 % p.splits));
            }
            mbar_wait(smem_u32(&acc_full[acc]), acc_phase);
            tc_fence_after();
            if (threadIdx.x == 64 && w == (int)blockIdx.x) tc_stamp(p, 5);
            const int row_base = m0 + q * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            float sq[8];
            epi_tile<BN>(e, p.M, p.N, p.splits, sp, row_base, n0, taddr, stg, s_mul, s_bias, s_sc, s_sh, lane, col_lo, col_hi, sq);
            // accumulator drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[acc]));
            if (col_hi > col_lo && n0 + col_lo < p.N) epi_rowpart(e, p.M, row_base, (n0 + col_lo) / ROWPART_COLS, lane, sq);
            if (threadIdx.x == 64 && w == (int)blockIdx.x) tc_stamp(p, 6);
            if (lane == 0 && w == (int)blockIdx.x) tc_stamp(p, 6 + warp);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
    if (threadIdx.x == 0) tc_stamp(p, 7);
}

// ---- the fused epilogue as a stand-alone kernel (small batches) ------------------------------------------------
// A scoring call of 17..256 rows is one or two row tiles: 11-28 CTAs per layer, each walking all k-blocks alone (27 at
// K = 1728: 12 us) and then the fused epilogue with nothing to hide it behind (9-13 us) -- 25 us per layer, 15 layers.  The same
// layer as a PLAIN split-K GEMM over all SMs (5 us) plus this kernel (bias is already in the accumulator; LeakyReLU, BatchNorm
// affine, fp32 stash, fp16 hi/lo twins, diff against the reference, diff twins, row sums of squares per 128-column block) is a
// third of that.  The GEMM's splits do NOT add atomically: split sp stores its partial tile into slab sp and this kernel adds
// the slabs in order, so the scores stay bit-reproducible from call to call and nothing has to be zeroed.
// One warp per (row, 128-column block), lane <-> 4 columns.
__global__ void __launch_bounds__(256) epi_follow_kernel(const Epilogue e, const float* __restrict__ acc, int ldacc, int rows, int N, int cols_p,
                                                         int splits, long long slab_stride) {
    pdl_trigger();
    pdl_wait();
    const int nblk = (cols_p + 127) / 128;
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= rows * nblk) return;
    const int r = w / nblk, blk = w - r * nblk;
    const int c = blk * 128 + lane * 4;
    float sq = 0.f;
    if (c < cols_p) {
        float y[4] = {0.f, 0.f, 0.f, 0.f};
        if (c < N) {
            float4 a = *reinterpret_cast<const float4*>(acc + (size_t)r * ldacc + c);
            for (int sp = 1; sp < splits; ++sp) {          // the split-K partial tiles, added in a fixed order: deterministic
                const float4 b = *reinterpret_cast<const float4*>(acc + (size_t)sp * (size_t)slab_stride + (size_t)r * ldacc + c);
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            y[0] = a.x; y[1] = a.y; y[2] = a.z; y[3] = a.w;
            if (e.bn_scale) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j < N) {
                        const float v = y[j] > 0.f ? y[j] : y[j] * e.slope;
                        y[j] = fmaf(v, __ldg(e.bn_scale + c + j), __ldg(e.bn_shift + c + j));
                    }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) if (c + j >= N) y[j] = 0.f;
        }
        if (e.Y && c < e.y_cols) *reinterpret_cast<float4*>(e.Y + (size_t)r * e.ldy + c) = make_float4(y[0], y[1], y[2], y[3]);
        if (e.Yh && c < e.y_cols) {
            const float ys = e.y_split_scale;
            const __half2 h01 = __floats2half2_rn(y[0] * ys, y[1] * ys), h23 = __floats2half2_rn(y[2] * ys, y[3] * ys);
            uint2 hv;
            hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
            *reinterpret_cast<uint2*>(e.Yh + (size_t)r * e.ldh + c) = hv;
            if (e.Yl) {
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn(y[0] * ys - f01.x, y[1] * ys - f01.y);
                const __half2 l23 = __floats2half2_rn(y[2] * ys - f23.x, y[3] * ys - f23.y);
                uint2 lv;
                lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                *reinterpret_cast<uint2*>(e.Yl + (size_t)r * e.ldh + c) = lv;
            }
        }
        if (e.ref) {
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            if (c + 3 < N) {
                const float4 f = __ldg(reinterpret_cast<const float4*>(e.ref + (size_t)r * e.ldref + c));
                d[0] = y[0] - f.x; d[1] = y[1] - f.y; d[2] = y[2] - f.z; d[3] = y[3] - f.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (c + j < N) d[j] = y[j] - e.ref[(size_t)r * e.ldref + c + j];
            }
            sq = fmaf(d[0], d[0], fmaf(d[1], d[1], fmaf(d[2], d[2], d[3] * d[3])));
            if (e.dout && c < e.d_cols) *reinterpret_cast<float4*>(e.dout + (size_t)r * e.lddout + c) = make_float4(d[0], d[1], d[2], d[3]);
            if (e.Dh && c < e.d_cols) {
                const float s0 = d[0] * e.d_scale, s1 = d[1] * e.d_scale, s2 = d[2] * e.d_scale, s3 = d[3] * e.d_scale;
                const __half2 h01 = __floats2half2_rn(s0, s1), h23 = __floats2half2_rn(s2, s3);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn(s0 - f01.x, s1 - f01.y), l23 = __floats2half2_rn(s2 - f23.x, s3 - f23.y);
                uint2 hv, lv;
                hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                *reinterpret_cast<uint2*>(e.Dh + (size_t)r * e.lddh + c) = hv;
                *reinterpret_cast<uint2*>(e.Dl + (size_t)r * e.lddh + c) = lv;
            }
        }
    }
    if (e.rowpart && e.ref) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) e.rowpart[(size_t)blk * e.rowpart_stride + r] = sq;
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
    MMAD_TC_ATTR(3, false, false); MMAD_TC_ATTR(1, false, false); MMAD_TC_ATTR(2, false, false); MMAD_TC_ATTR(4, false, false);
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
int gemm_tc_rowpart_cols() { return ROWPART_COLS; }

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

int tc_make_operand_map_f8(CUtensorMap* map, const void* base, int rows, int k, int ld_bytes, int box_rows) {
    if (!init_tc()) { set_error("tcgen05 path unavailable on this device"); return MMAD_E_UNSUPPORTED; }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld_bytes % 16)) { set_error("TMA operand must be 16-byte aligned (ld=%d bytes)", ld_bytes); return MMAD_E_ARG; }
    cuuint64_t dims[2] = {(cuuint64_t)2 * ((k + BK - 1) / BK * BK), (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld_bytes};
    cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (fp8 twin) failed (%d) rows=%d k=%d ld=%d", (int)r, rows, k, ld_bytes); return MMAD_E_CUDA; }
    return MMAD_OK;
}

int gemm_tc(const TcOperand& A, const TcOperand& B, int M, int N, int K, int passes, const Epilogue& e, cudaStream_t s, int bn,
            int* splits_out, int max_splits) {
    if (splits_out) *splits_out = 1;
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
    const bool debug = getenv("MMAD_TC_DEBUG") != nullptr;     // read per call: a profiling script switches it on after warm-up
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (debug) cudaStreamIsCapturing(s, &cap);
    p.dbg = debug && cap == cudaStreamCaptureStatusNone;
    p.tri = e.b_upper_tri;     // leading rows of B that are upper triangular (tiles entirely inside skip k-blocks left of them)
    if (!p.tri && e.plain && e.split_k_ok) {
        // split-K when it fills the machine better: cost ~ waves x k-blocks per item; ties go to fewer splits
        // (every split adds one atomic pass over the output)
        long best = (long)((tiles + g_num_sms - 1) / g_num_sms) * num_kb;
        for (int sp = 2; sp <= max_splits && sp <= num_kb; ++sp) {
            const long waves = ((long)tiles * sp + g_num_sms - 1) / g_num_sms;
            const long cost = waves * ((num_kb + sp - 1) / sp) + waves;       // + per-item fill/epilogue overhead
            if (cost * 10 < best * 9) { best = cost; p.splits = sp; }          // require >= 10 % gain
        }
    }
    if (splits_out) *splits_out = p.splits;
    if (p.splits > 1 && !e.pre_zeroed && !e.slab_stride)    // partial sums are accumulated atomically: the output starts at zero
        MMAD_CUDA_OK(cudaMemset2DAsync(e.Y, (size_t)e.ldy * 4, 0, (size_t)N * 4, M, s));
    const int items = tiles * p.splits;
    const int grid = items < g_num_sms ? items : g_num_sms;
    if (A.mn && !B.mn) { set_error("gemm_tc: MN-major A with K-major B is not instantiated"); return MMAD_E_UNSUPPORTED; }
#define MMAD_TC_LAUNCH2(P, AM, BMN, BNV)                                                                              \
    MMAD_CUDA_OK(launch_k(gemm_tc_kernel<P, AM, BMN, BNV>, dim3(grid), dim3(NTHREADS), Cfg<P, BNV>::kSmemBytes, s, A.hi, P >= 2 ? A.lo : A.hi, B.hi, \
                          P >= 3 ? B.lo : B.hi, p, e))
#define MMAD_TC_LAUNCH(P, AM, BMN)                  \
    do {                                            \
        if (bn == 256) MMAD_TC_LAUNCH2(P, AM, BMN, 256); \
        else MMAD_TC_LAUNCH2(P, AM, BMN, 128);      \
    } while (0)
    if (passes == 2 || passes == 4) {
        if (A.mn || B.mn) { set_error("gemm_tc: the 2-pass and fp8-assisted modes are instantiated for K-major operands only"); return MMAD_E_UNSUPPORTED; }
        if (passes == 4) MMAD_TC_LAUNCH(4, false, false);
        else MMAD_TC_LAUNCH(2, false, false);
    } else if (passes == 3) {
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
    if (p.dbg) {
        static unsigned long long last_end = 0;
        unsigned long long t[16];
        MMAD_CUDA_OK(cudaStreamSynchronize(s));
        MMAD_CUDA_OK(cudaMemcpyFromSymbol(t, g_tc_stamps, sizeof t));
        fprintf(stderr, "gemm_tc M=%d N=%d K=%d passes=%d A_mn=%d B_mn=%d bn=%d splits=%d grid=%d plain=%d: prologue %llu, first TMA issued +%llu, first stage "
                "landed +%llu, last MMA issued +%llu, accumulator ready +%llu, epilogue +%llu, exit +%llu = %llu ns\n", M, N, K, passes, (int)A.mn,
                (int)B.mn, bn, p.splits, grid, (int)e.plain, t[1] - t[0], t[2] - t[1], t[3] - t[1], t[4] - t[1], t[5] - t[1], t[6] - t[5], t[7] - t[6], t[7] - t[0]);
        fprintf(stderr, "    epilogue warps 2-9 done at +%llu %llu %llu %llu | %llu %llu %llu %llu after the accumulator\n", t[8] - t[5], t[9] - t[5],
                t[10] - t[5], t[11] - t[5], t[12] - t[5], t[13] - t[5], t[14] - t[5], t[15] - t[5]);
        last_end = t[7];
    }
    return MMAD_OK;
}

// small-batch layer: plain split-K GEMM into slabs (one per split, `slab_stride` elements apart) + the follow-up kernel above
int gemm_tc_small(const TcOperand& A, const TcOperand& B, int M, int N, int K, int passes, const Epilogue& e, float* slabs, int ldacc,
                  long long slab_stride, int max_slabs, cudaStream_t s) {
    if (e.plain || e.pre || e.sq_self || e.lo_f8 || e.d_lo_f8 || e.col_scale || e.b_upper_tri || (e.rowpart && !e.ref) || ROWPART_COLS != 128) {
        set_error("gemm_tc_small: epilogue not covered by the follow-up kernel");
        return MMAD_E_UNSUPPORTED;
    }
    Epilogue pe;
    pe.bias = e.bias; pe.acc_scale = e.acc_scale; pe.acc_comp = e.acc_comp;
    pe.Y = slabs; pe.ldy = ldacc; pe.y_cols = N; pe.plain = 1; pe.split_k_ok = 1; pe.slab_stride = slab_stride;
    int splits = 1;
    int rc = gemm_tc(A, B, M, N, K, passes, pe, s, 128, &splits, std::max(1, std::min(16, max_slabs)));
    if (rc) return rc;
    int cols_p = std::max(std::max(e.Y ? e.y_cols : 0, e.Yh ? e.y_cols : 0), e.ref ? std::max(e.d_cols, N) : 0);
    if (cols_p <= 0) return MMAD_OK;
    cols_p = (cols_p + 3) & ~3;
    const int nblk = (cols_p + 127) / 128;
    MMAD_CUDA_OK(launch_k(epi_follow_kernel, dim3((M * nblk + 7) / 8), dim3(256), 0, s, e, (const float*)slabs, ldacc, M, N, cols_p, splits, slab_stride));
    MMAD_LAUNCHED();
    return MMAD_OK;
}

}  // namespace mmad
