// Multimodal feature extractor in front of the autoencoder:
// utils/data_loaders.py:152-229 (HSR_Net) / 601-674 (Multisensory_module) of the reference.
//
// Per sample the reference runs, in a Python loop, three tiny convolutions on the 32x32 RGB hand-camera image
// (3->16 k2 s2, 16->16 k3 p1, 16->16 k2 s2, ReLU after each -> 16x8x8), the same on the depth image
// (1->8->8->8 -> 8x8x8), broadcasts the force-torque scalar to 8x8, runs two 1-d convolutions on the 13 MFCC
// coefficients (1->8 k18 s9 p9, 8->16 k2 s2 -> 16 values viewed as 2x8 and repeated 8x), and concatenates the
// channels: [rgb 1024 | depth 512 | force-torque 64 | mic 128] = 1728 floats (utils/data_loaders.py:226).
// Here: ONE launch for the whole batch, one CTA per sample, every intermediate in shared memory, the input
// normalisation norm_vec (utils/data_loaders.py:703-712) folded into the loads as an affine map.
#include "mmad_internal.cuh"

using namespace mmad;

namespace {

constexpr int FT = 256;

// out[co][y][x] = relu(b[co] + sum_{ci,ky,kx} w[co][ci][ky][kx] * in[ci][y*S + ky - P][x*S + kx - P])
template <int CI, int CO, int KS, int S, int P, int HI, int HO>
__device__ __forceinline__ void conv2d_relu(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ b,
                                            float* __restrict__ out, int out_stride_is_global) {
    for (int idx = threadIdx.x; idx < CO * HO * HO; idx += FT) {
        const int co = idx / (HO * HO), y = (idx / HO) % HO, x = idx % HO;
        float acc = b[co];
#pragma unroll 4
        for (int ci = 0; ci < CI; ++ci) {
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
                const int iy = y * S + ky - P;
                if (iy < 0 || iy >= HI) continue;
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    const int ix = x * S + kx - P;
                    if (ix < 0 || ix >= HI) continue;
                    acc = fmaf(w[((co * CI + ci) * KS + ky) * KS + kx], in[(ci * HI + iy) * HI + ix], acc);
                }
            }
        }
        out[idx] = fmaxf(acc, 0.f);
    }
    (void)out_stride_is_global;
}

__device__ __forceinline__ void stage(float* dst, const float* __restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += FT) dst[i] = src[i];
}

struct FeatArgs {
    const float* r; const float* d; const float* t; const float* m;       // inputs (NULL = modality absent)
    mmad_feature_weights_t w;
    float aff[8];                                                          // (scale, shift) for r, d, t, m
    float* out; int ldo;
    int off_r, off_d, off_t, off_m;                                        // output column offsets
};

__global__ void __launch_bounds__(FT) multisensory_kernel(FeatArgs a) {
    extern __shared__ __align__(16) float fs[];
    float* img = fs;                    // [3][32][32]
    float* a1 = img + 3 * 32 * 32;      // [16][16][16]
    float* a2 = a1 + 16 * 16 * 16;      // [16][16][16]
    float* wt = a2 + 16 * 16 * 16;      // weights of the convolution in flight (<= 2304 + 16)
    const int b = blockIdx.x;
    float* out = a.out + (size_t)b * a.ldo;
    if (a.r) {
        const float* src = a.r + (size_t)b * 3 * 32 * 32;
        for (int i = threadIdx.x; i < 3 * 32 * 32; i += FT) img[i] = fmaf(src[i], a.aff[0], a.aff[1]);
        stage(wt, a.w.conv1r_w, 16 * 3 * 4); stage(wt + 192, a.w.conv1r_b, 16);
        __syncthreads();
        conv2d_relu<3, 16, 2, 2, 0, 32, 16>(img, wt, wt + 192, a1, 0);
        __syncthreads();
        stage(wt, a.w.conv2r_w, 16 * 16 * 9); stage(wt + 2304, a.w.conv2r_b, 16);
        __syncthreads();
        conv2d_relu<16, 16, 3, 1, 1, 16, 16>(a1, wt, wt + 2304, a2, 0);
        __syncthreads();
        stage(wt, a.w.conv3r_w, 16 * 16 * 4); stage(wt + 1024, a.w.conv3r_b, 16);
        __syncthreads();
        conv2d_relu<16, 16, 2, 2, 0, 16, 8>(a2, wt, wt + 1024, out + a.off_r, 1);
        __syncthreads();
    }
    if (a.d) {
        const float* src = a.d + (size_t)b * 32 * 32;
        for (int i = threadIdx.x; i < 32 * 32; i += FT) img[i] = fmaf(src[i], a.aff[2], a.aff[3]);
        stage(wt, a.w.conv1d_w, 8 * 4); stage(wt + 32, a.w.conv1d_b, 8);
        __syncthreads();
        conv2d_relu<1, 8, 2, 2, 0, 32, 16>(img, wt, wt + 32, a1, 0);
        __syncthreads();
        stage(wt, a.w.conv2d_w, 8 * 8 * 9); stage(wt + 576, a.w.conv2d_b, 8);
        __syncthreads();
        conv2d_relu<8, 8, 3, 1, 1, 16, 16>(a1, wt, wt + 576, a2, 0);
        __syncthreads();
        stage(wt, a.w.conv3d_w, 8 * 8 * 4); stage(wt + 256, a.w.conv3d_b, 8);
        __syncthreads();
        conv2d_relu<8, 8, 2, 2, 0, 16, 8>(a2, wt, wt + 256, out + a.off_d, 1);
        __syncthreads();
    }
    if (a.t) {      // t[i].repeat(1, 1, 8, 8)  (utils/data_loaders.py:645-648)
        const float v = fmaf(a.t[b], a.aff[4], a.aff[5]);
        for (int i = threadIdx.x; i < 64; i += FT) out[a.off_t + i] = v;
    }
    if (a.m) {      // conv1l (1->8, k18 s9 p9) on 13 coefficients -> [8][2]; conv2l (8->16, k2 s2) -> [16][1]
        float* mi = img;            // 13 inputs
        float* m1 = img + 16;       // [8][2]
        float* m2 = img + 32;       // [16]
        if (threadIdx.x < 13) mi[threadIdx.x] = fmaf(a.m[(size_t)b * 13 + threadIdx.x], a.aff[6], a.aff[7]);
        __syncthreads();
        if (threadIdx.x < 16) {
            const int co = threadIdx.x >> 1, p = threadIdx.x & 1;
            float acc = a.w.conv1l_b[co];
            for (int k = 0; k < 18; ++k) {
                const int i = p * 9 + k - 9;
                if (i >= 0 && i < 13) acc = fmaf(a.w.conv1l_w[co * 18 + k], mi[i], acc);
            }
            m1[co * 2 + p] = fmaxf(acc, 0.f);
        }
        __syncthreads();
        if (threadIdx.x < 16) {
            const int co = threadIdx.x;
            float acc = a.w.conv2l_b[co];
            for (int ci = 0; ci < 8; ++ci)
                for (int k = 0; k < 2; ++k) acc = fmaf(a.w.conv2l_w[(co * 8 + ci) * 2 + k], m1[ci * 2 + k], acc);
            m2[co] = fmaxf(acc, 0.f);
        }
        __syncthreads();
        // view(-1, 2, 8, 1).repeat(1, 1, 1, 8): out[j][y][x] = m2[j*8 + y]
        for (int i = threadIdx.x; i < 128; i += FT) out[a.off_m + i] = m2[i >> 3];
    }
}

constexpr int kFeatSmem = (3 * 32 * 32 + 2 * 16 * 16 * 16 + 2304 + 16) * 4;
bool g_feat_init = false;

}  // namespace

extern "C" {

int mmad_multisensory_width(int has_r, int has_d, int has_t, int has_m) {
    return (has_r ? 1024 : 0) + (has_d ? 512 : 0) + (has_t ? 64 : 0) + (has_m ? 128 : 0);
}

int mmad_multisensory_forward(const float* d_r, const float* d_d, const float* d_t, const float* d_m, int batch,
                              const mmad_feature_weights_t* w, const float* h_affine, float* d_out, int ldo, void* stream) {
    if (!w || !d_out || batch < 0) { set_error("bad argument"); return MMAD_E_ARG; }
    if (batch == 0) return MMAD_OK;
    const int width = mmad_multisensory_width(d_r != nullptr, d_d != nullptr, d_t != nullptr, d_m != nullptr);
    if (width == 0 || ldo < width) { set_error("no modality given or ldo < %d", width); return MMAD_E_ARG; }
    if ((d_r && !(w->conv1r_w && w->conv1r_b && w->conv2r_w && w->conv2r_b && w->conv3r_w && w->conv3r_b)) ||
        (d_d && !(w->conv1d_w && w->conv1d_b && w->conv2d_w && w->conv2d_b && w->conv3d_w && w->conv3d_b)) ||
        (d_m && !(w->conv1l_w && w->conv1l_b && w->conv2l_w && w->conv2l_b))) {
        set_error("missing convolution weights for a given modality"); return MMAD_E_ARG;
    }
    if (!g_feat_init) {
        MMAD_CUDA_OK(cudaFuncSetAttribute(multisensory_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFeatSmem));
        g_feat_init = true;
    }
    FeatArgs a;
    a.r = d_r; a.d = d_d; a.t = d_t; a.m = d_m; a.w = *w; a.out = d_out; a.ldo = ldo;
    for (int i = 0; i < 4; ++i) { a.aff[2 * i] = h_affine ? h_affine[2 * i] : 1.f; a.aff[2 * i + 1] = h_affine ? h_affine[2 * i + 1] : 0.f; }
    int off = 0;
    a.off_r = off; off += d_r ? 1024 : 0;
    a.off_d = off; off += d_d ? 512 : 0;
    a.off_t = off; off += d_t ? 64 : 0;
    a.off_m = off;
    multisensory_kernel<<<batch, FT, kFeatSmem, (cudaStream_t)stream>>>(a);
    MMAD_LAUNCHED();
    MMAD_CUDA_OK(cudaGetLastError());
    return MMAD_OK;
}

}  // extern "C"
