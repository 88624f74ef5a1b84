// fp32 FFMA kernels for the HiFi-GAN generator (HFG_MODE_FP32, and every layer
// shape the tensor-core path does not cover).  Activations are [B, C, T] fp32
// (time fastest), exactly the reference's layout (reference models/hifigan.py:
// 144-147), so stage dumps compare element for element.
//
// One tile kernel does Conv1d (any k, dilation) and -- launched once per output
// phase -- ConvTranspose1d as a polyphase convolution:
//
//   y[b, co, q*os + phase - p] = bias[co]
//        + sum_{m < taps(phase)} sum_{ci} W[phase][m][ci][co] * act(x[b, ci, q + m*dil - pad])
//
//   Conv1d:            os = 1, one phase, dil = d, pad = d(k-1)/2, W[m] = w[:, :, m]
//   ConvTranspose1d:   os = u, phase r in [0,u), dil = -1, pad = 0, taps = ceil((k-r)/u),
//                      W[r][m][ci][co] = w[ci, co, r + m*u]           (input index q - m)
//
// act() is leaky_relu(., slope) fused into the tile load (reference :81,83,244).
// The epilogue fuses the residual add (:85) and the MRF running sum / final
// division (:126-131) in the reference's own association order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hfg {

constexpr int kTileT = 128;     // time positions per CTA
constexpr int kTileCi = 8;      // input channels per smem stage
constexpr int kThreads = 256;   // 16 time lanes x 16 channel groups

enum : int { EPI_ACC_READ = 1, EPI_ACC_DIV = 2 };

struct ConvArgs {
    const float* x;      // [B, Cin, Tin]
    const float* w;      // [phases][taps_max][CinPad][CoutPad]
    const float* bias;   // [Cout]
    const float* res;    // [B, Cout, Tout] or nullptr
    float* y;            // [B, Cout, Tout]
    int Cin, CinPad, Tin;
    int Cout, CoutPad, Tout;
    int taps_max;        // taps of phase 0 (the longest)
    int k, u;            // ConvTranspose1d: kernel size and stride (u = 1 for Conv1d)
    int dil, pad;        // x index = q + m*dil - pad
    int out_stride, out_off;   // t = q*out_stride + phase + out_off
    int phases;
    int co_tiles;
    float slope;         // leaky_relu slope applied to x on load; 1.0f = identity
    int epi_flags;
    float div;
};

template <int RCO>
__global__ void __launch_bounds__(kThreads)
conv_tile_fp32(const ConvArgs a) {
    constexpr int TCO = RCO * 16;
    extern __shared__ __align__(16) float fp32_smem[];
    float* smem = fp32_smem;

    const int tx = threadIdx.x & 15;
    const int ty = threadIdx.x >> 4;
    const int b = blockIdx.z;
    const int phase = blockIdx.y / a.co_tiles;
    const int co0 = (blockIdx.y % a.co_tiles) * TCO;
    const int q0 = blockIdx.x * kTileT;

    int taps = a.taps_max;
    if (a.phases > 1) taps = (a.k - phase + a.u - 1) / a.u;   // taps of this phase
    if (taps < 0) taps = 0;

    // smem x window: offsets m*dil - pad for m in [0,taps_max)
    const int span = (a.taps_max - 1) * (a.dil < 0 ? -a.dil : a.dil);
    const int min_off = (a.dil < 0 ? -span : 0) - a.pad;
    const int XW = kTileT + span;
    float* xs = smem;                                  // [kTileCi][XW]
    float* ws = smem + kTileCi * XW;                   // [taps_max][kTileCi][TCO]

    float acc[RCO][8];
#pragma unroll
    for (int r = 0; r < RCO; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[r][i] = 0.f;

    const float* xb = a.x + (size_t)b * a.Cin * a.Tin;
    const float* wp = a.w + (size_t)phase * a.taps_max * a.CinPad * a.CoutPad;

    for (int ci0 = 0; ci0 < a.CinPad; ci0 += kTileCi) {
        __syncthreads();
        for (int e = threadIdx.x; e < kTileCi * XW; e += kThreads) {
            const int ci = e / XW, xo = e - ci * XW;
            const int t = q0 + min_off + xo;
            float v = 0.f;
            if (ci0 + ci < a.Cin && t >= 0 && t < a.Tin) {
                v = __ldg(xb + (size_t)(ci0 + ci) * a.Tin + t);
                v = v > 0.f ? v : v * a.slope;
            }
            xs[e] = v;
        }
        for (int e = threadIdx.x; e < taps * kTileCi * TCO; e += kThreads) {
            const int co = e % TCO;
            const int ci = (e / TCO) % kTileCi;
            const int m = e / (TCO * kTileCi);
            ws[e] = __ldg(wp + ((size_t)m * a.CinPad + ci0 + ci) * a.CoutPad + co0 + co);
        }
        __syncthreads();
        for (int m = 0; m < taps; ++m) {
            const int xo = m * a.dil - a.pad - min_off + tx;
#pragma unroll
            for (int ci = 0; ci < kTileCi; ++ci) {
                float wv[RCO], xv[8];
                const float* wrow = ws + (m * kTileCi + ci) * TCO + ty * RCO;
                if constexpr (RCO % 4 == 0) {   // 16-byte aligned: XW*kTileCi and TCO are multiples of 4
#pragma unroll
                    for (int r = 0; r < RCO; r += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(wrow + r);
                        wv[r] = w4.x; wv[r + 1] = w4.y; wv[r + 2] = w4.z; wv[r + 3] = w4.w;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < RCO; ++r) wv[r] = wrow[r];
                }
                const float* xrow = xs + ci * XW + xo;
#pragma unroll
                for (int i = 0; i < 8; ++i) xv[i] = xrow[16 * i];
#pragma unroll
                for (int r = 0; r < RCO; ++r)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[r][i] = fmaf(wv[r], xv[i], acc[r][i]);
            }
        }
    }

#pragma unroll
    for (int r = 0; r < RCO; ++r) {
        const int co = co0 + ty * RCO + r;
        if (co >= a.Cout) continue;
        const float bv = a.bias[co];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int q = q0 + tx + 16 * i;
            const int t = q * a.out_stride + phase + a.out_off;
            if (t < 0 || t >= a.Tout) continue;
            const size_t idx = ((size_t)b * a.Cout + co) * a.Tout + t;
            float v = acc[r][i] + bv;
            if (a.res) v = a.res[idx] + v;                       // x + xt          (:85)
            if (a.epi_flags & EPI_ACC_READ) v = a.y[idx] + v;    // output + rb(x)  (:129)
            if (a.epi_flags & EPI_ACC_DIV) v = v / a.div;        // / len(resblocks)(:131)
            a.y[idx] = v;
        }
    }
}

// conv_post (C_out = 1) + tanh: bandwidth kernel (reference models/hifigan.py:254-256).
// One thread per output sample; the 7-tap window re-reads hit L1.
struct PostArgs {
    const float* x;     // [B, Cin, T]
    const float* w;     // [Cin][k]
    const float* bias;  // [1]
    float* y;           // [B, 1, T]
    int Cin, T, k, pad;
    float slope;
};

__global__ void __launch_bounds__(256)
conv_post_tanh_fp32(const PostArgs a) {
    extern __shared__ float wsm[];
    for (int e = threadIdx.x; e < a.Cin * a.k; e += blockDim.x) wsm[e] = a.w[e];
    __syncthreads();
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.T) return;
    const float* xb = a.x + (size_t)b * a.Cin * a.T;
    float acc = 0.f;
    for (int ci = 0; ci < a.Cin; ++ci) {
        const float* xr = xb + (size_t)ci * a.T;
        for (int j = 0; j < a.k; ++j) {
            const int s = t + j - a.pad;
            if (s >= 0 && s < a.T) {
                float v = __ldg(xr + s);
                v = v > 0.f ? v : v * a.slope;
                acc = fmaf(wsm[ci * a.k + j], v, acc);
            }
        }
    }
    a.y[(size_t)b * a.T + t] = tanhf(acc + a.bias[0]);
}

// [B, T, C] -> [B, C, T] (fp32 mode only: the acoustic model's frames-last mel)
__global__ void transpose_btc_to_bct(const float* __restrict__ x, float* __restrict__ y, int C, int T) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (t < T && c < C) ? x[((size_t)b * T + t) * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        if (c < C && t < T) y[((size_t)b * C + c) * T + t] = tile[threadIdx.x][i];
    }
}

}  // namespace hfg
