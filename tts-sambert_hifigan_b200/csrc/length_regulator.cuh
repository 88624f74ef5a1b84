// Length regulator + duration rounding: the integer frame indexing
// that produces the generator's input.  Bit-exact by construction (integer prefix sums, a binary
// search and copies); reference: models/variance_adaptor.py:171-269 (LengthRegulator.forward),
// :746-748 (dur = clamp(round(exp(log_dur)).long(), min=1)).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hfg {

// dur = max(1, (long) rint(exp(log_dur)))   -- torch.round is round-half-to-even == rintf.
// The reference materialises exp() as a float32 tensor before rounding (models/variance_adaptor.py:746), so the
// result is DEFINED here as rintf of the correctly rounded float32 exponential: exp is evaluated in fp64 (error
// < 1 fp64 ulp, far below half an fp32 ulp except in astronomically rare double-rounding cases) and rounded to
// fp32 once.  Deterministic, independent of the device's fast-math expf, and closest to the true value; a CPU
// vector expf that is 1 ulp off can disagree only where exp(x) lies within 1 fp32 ulp of k + 0.5.
__global__ void lr_durations_from_log(const float* __restrict__ log_dur, long long n, long long* __restrict__ dur) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d = rintf((float)exp((double)log_dur[i]));
    const long long v = (long long)d;
    dur[i] = v < 1 ? 1 : v;
}

// One block per utterance: inclusive prefix sum of clamp(dur, min=0) (reference :214-219) into cum[b, :].
__global__ void lr_prefix_sum(const long long* __restrict__ dur, int Tph, int* __restrict__ cum,
                              int* __restrict__ totals) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < Tph; base += blockDim.x) {
        const int i = base + threadIdx.x;
        long long d = i < Tph ? dur[(size_t)b * Tph + i] : 0;
        int v = d < 0 ? 0 : (d > 0x3fffffff ? 0x3fffffff : (int)d);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (lane == 31) warp_sums[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int w = lane < nwarp ? warp_sums[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += n;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int prefix = carry_s + (warp > 0 ? warp_sums[warp - 1] : 0);
        if (i < Tph) cum[(size_t)b * Tph + i] = prefix + v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s += warp_sums[nwarp - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[b] = carry_s;
}

// out[b, t, :] = henc[b, p, :] with p the first phoneme whose inclusive prefix sum exceeds t
// (torch.repeat_interleave, reference :232); zero beyond sum(dur[b]) (reference :240-260).
__global__ void lr_expand(const float* __restrict__ henc, const int* __restrict__ cum, int Tph, int D,
                          int Tfrm, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.y + threadIdx.y;
    if (t >= Tfrm) return;
    const int* c = cum + (size_t)b * Tph;
    int lo = 0, hi = Tph;                       // first index with c[idx] > t
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (c[mid] > t) hi = mid; else lo = mid + 1;
    }
    float* o = out + ((size_t)b * Tfrm + t) * D;
    if (lo >= Tph) {
        for (int d = threadIdx.x; d < D; d += blockDim.x) o[d] = 0.f;
    } else {
        const float* h = henc + ((size_t)b * Tph + lo) * D;
        for (int d = threadIdx.x; d < D; d += blockDim.x) o[d] = h[d];
    }
}

}  // namespace hfg
