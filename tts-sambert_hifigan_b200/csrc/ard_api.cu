// KV-cached autoregressive mel decoder (include/hfg_ard.h).
//
// Reference: PNCAARDecoder._forward_autoregressive (models/ar_decoder.py:167-238) re-runs prenet + positional
// encoding + the whole nn.TransformerDecoder on the growing prefix for every frame.  Here one step evaluates
// ONE position per layer against cached self-attention keys / values and a once-projected encoder memory:
//
//   x = prenet(frame_{t-1}) + pe[t]                                       (:200-204)
//   per layer (torch.nn.TransformerDecoderLayer, post-norm, relu):         (:74-83, called at :211-215)
//     q,k,v = in_proj(x);  cache k,v at row t;  x = LN1(x + out_proj(softmax(q K[0..t]^T / sqrt(hd)) V[0..t]))
//     x = LN2(x + out_proj(softmax(q' Kmem^T / sqrt(hd)) Vmem))            no memory mask, as the reference
//     x = LN3(x + W2 relu(W1 x + b1) + b2)
//   frame_t = mel_proj(x)                                                  (:218)
//
// All arithmetic fp32 FFMA with a fixed summation order (deterministic).  The ~70 launches of a step are
// captured into a CUDA graph once per decode call and replayed max_len times; the step index lives in device
// memory so that one graph serves every step.
#include "../../include/hfg_ard.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "model.h"

namespace hfg {

// ------------------------------------------------------------------ kernels
// Y[m, n] = act(sum_k X[m, k] W[n, k] + bias[n] + add_row[n]) for an M x N tile of 64 x TN per block
// (TN = 64: the once-per-call memory projection with thousands of rows; TN = 16: the per-step matrices, whose
// M = batch <= 64 rows would otherwise leave all but N / 64 SMs idle).  Fixed summation order over k.
// `step` (device) offsets add_row by *step * add_row_stride (the positional-encoding row of this step).
template <bool RELU, int TN>
__global__ void __launch_bounds__(256)
ard_gemm(const float* __restrict__ X, int ldx, const float* __restrict__ W, const float* __restrict__ bias,
         const float* __restrict__ add_row, long long add_row_stride, const int* __restrict__ step,
         float* __restrict__ Y, int ldy, int M, int N, int K) {
    constexpr int RG = 256 / TN, RPT = 64 / RG;                       // row groups, rows per thread
    __shared__ float Xs[64][33];
    __shared__ float Ws[TN][33];
    const int tx = threadIdx.x % TN, ty = threadIdx.x / TN;           // column inside the tile, row group
    const int n0 = blockIdx.x * TN, m0 = blockIdx.y * 64;
    float acc[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc[r] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 32) {
        for (int e = threadIdx.x; e < 64 * 32; e += 256) {
            const int r = e >> 5, c = e & 31;
            const int m = m0 + r, k = k0 + c;
            Xs[r][c] = (m < M && k < K) ? X[(size_t)m * ldx + k] : 0.f;
        }
        for (int e = threadIdx.x; e < TN * 32; e += 256) {
            const int r = e >> 5, c = e & 31;
            const int n = n0 + r, k = k0 + c;
            Ws[r][c] = (n < N && k < K) ? W[(size_t)n * K + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) {
            const float w = Ws[tx][kk];
#pragma unroll
            for (int r = 0; r < RPT; ++r) acc[r] = fmaf(Xs[ty + RG * r][kk], w, acc[r]);
        }
        __syncthreads();
    }
    const int n = n0 + tx;
    if (n >= N) return;
    float b = bias ? bias[n] : 0.f;
    if (add_row) b += add_row[(step ? (long long)(*step) : 0ll) * add_row_stride + n];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int m = m0 + ty + RG * r;
        if (m >= M) continue;
        float v = acc[r] + b;
        if (RELU) v = fmaxf(v, 0.f);
        Y[(size_t)m * ldy + n] = v;
    }
}

// The per-step matrices (M = batch <= 64 rows per block row): a block owns 16 output columns and one 256-wide
// chunk of K, holds that whole X chunk (64 x 256) and W chunk (16 x 256) in shared memory, and runs the K loop
// without a barrier.  grid = (N / 16, M / 64, K / 256): with more than one K chunk (linear2, K = d_ff) the blocks
// write raw partial sums part[z][m][n] that ard_add_layernorm adds up in a fixed order together with the bias.
constexpr int kStepKC = 256, kStepTN = 16;
template <bool RELU>
__global__ void __launch_bounds__(256)
ard_gemm_step(const float* __restrict__ X, int ldx, const float* __restrict__ W, const float* __restrict__ bias,
              const float* __restrict__ add_row, long long add_row_stride, const int* __restrict__ step,
              float* __restrict__ Y, int ldy, long long part_stride, int M, int N, int K) {
    extern __shared__ float sm[];
    float (*Xs)[kStepKC + 1] = reinterpret_cast<float (*)[kStepKC + 1]>(sm);                       // [64][257]
    float (*Ws)[kStepKC + 1] = reinterpret_cast<float (*)[kStepKC + 1]>(sm + 64 * (kStepKC + 1));  // [16][257]
    const int tx = threadIdx.x % kStepTN, ty = threadIdx.x / kStepTN;      // column, row group (16 groups x 4 rows)
    const int n0 = blockIdx.x * kStepTN, m0 = blockIdx.y * 64;
    // this block's K range: one chunk when K is split over grid.z, all of K otherwise
    const int k_begin = gridDim.z > 1 ? blockIdx.z * kStepKC : 0;
    const int k_end = gridDim.z > 1 ? (k_begin + kStepKC < K ? k_begin + kStepKC : K) : K;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = k_begin; k0 < k_end; k0 += kStepKC) {
        const int kc = (k_end - k0) < kStepKC ? (k_end - k0) : kStepKC;
        if (k0 > k_begin) __syncthreads();                  // previous chunk consumed
        for (int e = threadIdx.x; e < 64 * kStepKC; e += 256) {
            const int r = e / kStepKC, c = e - r * kStepKC;
            Xs[r][c] = (m0 + r < M && c < kc) ? X[(size_t)(m0 + r) * ldx + k0 + c] : 0.f;
        }
        for (int e = threadIdx.x; e < kStepTN * kStepKC; e += 256) {
            const int r = e / kStepKC, c = e - r * kStepKC;
            Ws[r][c] = (n0 + r < N && c < kc) ? W[(size_t)(n0 + r) * K + k0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < kStepKC; ++kk) {
            const float w = Ws[tx][kk];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = fmaf(Xs[ty + 16 * r][kk], w, acc[r]);
        }
    }
    const int n = n0 + tx;
    if (n >= N) return;
    const bool partial = gridDim.z > 1;
    float b = 0.f;
    if (!partial) {
        b = bias ? bias[n] : 0.f;
        if (add_row) b += add_row[(step ? (long long)(*step) : 0ll) * add_row_stride + n];
    }
    float* Yp = Y + (partial ? (long long)blockIdx.z * part_stride : 0ll);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int m = m0 + ty + 16 * r;
        if (m >= M) continue;
        float v = acc[r] + b;
        if (RELU && !partial) v = fmaxf(v, 0.f);
        Yp[(size_t)m * ldy + n] = v;
    }
}

// out[m, :] = LayerNorm(x[m, :] + r[m, :]) * gamma + beta   (one warp per row; d <= 1024, eps as nn.LayerNorm).
// r may be split into n_part raw partial sums (split-K GEMM) whose bias `rbias` is added here, in a fixed order.
__global__ void ard_add_layernorm(const float* __restrict__ x, const float* __restrict__ r, int n_part, long long part_stride,
                                  const float* __restrict__ rbias, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float* __restrict__ out, int M, int d, float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    float v[32];
    const int per = (d + 31) / 32;
    float s = 0.f;
    for (int i = 0; i < per; ++i) {
        const int c = lane + 32 * i;
        float rv = 0.f;
        if (c < d) {
            rv = rbias ? rbias[c] : 0.f;
            for (int z = 0; z < n_part; ++z) rv += r[(long long)z * part_stride + (size_t)row * d + c];
        }
        v[i] = c < d ? x[(size_t)row * d + c] + rv : 0.f;
        s += v[i];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / d;
    float q = 0.f;
    for (int i = 0; i < per; ++i) {
        const int c = lane + 32 * i;
        const float dv = c < d ? v[i] - mean : 0.f;
        q += dv * dv;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float inv = rsqrtf(q / d + eps);
    for (int i = 0; i < per; ++i) {
        const int c = lane + 32 * i;
        if (c < d) out[(size_t)row * d + c] = (v[i] - mean) * inv * gamma[c] + beta[c];
    }
}

// One (utterance, head) per block, 128 threads.  Optionally appends this step's key / value rows (taken from the
// packed in_proj output) to the cache first, then out[b, h*hd : (h+1)*hd] = softmax(q K^T * scale) V over `len` rows.
//   self-attention:  len = *step + 1, Kc / Vc = per-layer caches [B, cap, d]
//   cross-attention: len = frames,    Kc / Vc = projected encoder memory [B, frames, 2d] (K | V interleaved per row)
__global__ void __launch_bounds__(128)
ard_attention(const float* __restrict__ q, int ldq, const float* __restrict__ k_new, const float* __restrict__ v_new,
              float* __restrict__ Kc, float* __restrict__ Vc, long long kv_bstride, int kv_rstride,
              const int* __restrict__ step, int fixed_len, float* __restrict__ out, int d, int hd, float scale) {
    extern __shared__ float sc[];                          // scores [len], then 4 * hd partial sums
    __shared__ float red[4];
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = step ? *step : 0;
    const int len = k_new ? t + 1 : fixed_len;
    float* Kb = Kc + (long long)b * kv_bstride + h * hd;
    float* Vb = Vc + (long long)b * kv_bstride + h * hd;
    if (k_new) {                                           // append this position to the cache
        if (tid < hd) Kb[(long long)t * kv_rstride + tid] = k_new[(size_t)b * ldq + h * hd + tid];
        else if (tid < 2 * hd && 2 * hd <= 128) Vb[(long long)t * kv_rstride + tid - hd] = v_new[(size_t)b * ldq + h * hd + tid - hd];
        if (2 * hd > 128 && tid < hd) Vb[(long long)t * kv_rstride + tid] = v_new[(size_t)b * ldq + h * hd + tid];
        __syncthreads();
    }
    const float* qp = q + (size_t)b * ldq + h * hd;
    float mx = -INFINITY;
    for (int p = tid; p < len; p += 128) {
        const float* kp = Kb + (long long)p * kv_rstride;
        float s = 0.f;
        for (int i = 0; i < hd; ++i) s = fmaf(qp[i], kp[i], s);
        s *= scale;
        sc[p] = s;
        mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.f;
    for (int p = tid; p < len; p += 128) {
        const float e = expf(sc[p] - mx);
        sc[p] = e;
        sum += e;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = (red[0] + red[1]) + (red[2] + red[3]);
    // weighted sum of values: thread = (position group g, dim); groups = 128 / hd
    const int groups = 128 / hd, g = tid / hd, dd = tid - g * hd;
    float acc = 0.f;
    for (int p = g; p < len; p += groups) acc = fmaf(sc[p], Vb[(long long)p * kv_rstride + dd], acc);
    float* part = sc + len;                                 // [groups][hd]
    part[g * hd + dd] = acc;
    __syncthreads();
    if (tid < hd) {
        float o = 0.f;
        for (int gg = 0; gg < groups; ++gg) o += part[gg * hd + tid];
        out[(size_t)b * d + h * hd + tid] = o / sum;
    }
}

// frame -> mel[b, step, :].  The step counter is advanced by the following one-thread launch (ard_advance), the last
// node of a step, so that no kernel of the step ever races with the increment.
__global__ void ard_store_frame(const float* __restrict__ frame, float* __restrict__ mel, int B, int n_mels, int max_len,
                                const int* __restrict__ step) {
    const int t = *step;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < B * n_mels; e += gridDim.x * blockDim.x) {
        const int b = e / n_mels, c = e - b * n_mels;
        mel[((size_t)b * max_len + t) * n_mels + c] = frame[e];
    }
}
__global__ void ard_advance(int* step) { *step += 1; }

}  // namespace hfg

using namespace hfg;

struct hfg_ard_handle {
    hfg_ard_config cfg{};
    int device = 0;
    bool committed = false;
    std::string last_error;
    int64_t launches = 0;
    std::map<std::string, HostTensor> sd;
    std::map<std::string, float*> dev;             // committed tensors by key
    std::vector<void*> allocs;
    // The step graph is captured and replayed on a stream of the handle's own (the caller's may be the legacy
    // default stream, which cannot be captured); two events order it after / before the caller's stream.
    cudaStream_t work = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    void ensure_stream() {
        if (work) return;
        check_cuda(cudaStreamCreateWithFlags(&work, cudaStreamNonBlocking), "cudaStreamCreate");
        check_cuda(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming), "cudaEventCreate");
        check_cuda(cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming), "cudaEventCreate");
    }
    void free_stream() {
        if (work) cudaStreamDestroy(work);
        if (ev_in) cudaEventDestroy(ev_in);
        if (ev_out) cudaEventDestroy(ev_out);
        work = nullptr; ev_in = ev_out = nullptr;
    }
    float* upload(const std::vector<float>& v) {
        float* p = nullptr;
        check_cuda(cudaMalloc((void**)&p, std::max<size_t>(1, v.size()) * sizeof(float)), "cudaMalloc(ard weights)");
        allocs.push_back(p);
        check_cuda(cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice), "cudaMemcpy(ard weights)");
        return p;
    }
    void free_weights() {
        for (void* p : allocs) cudaFree(p);
        allocs.clear();
        dev.clear();
        committed = false;
    }
};

namespace {

struct ArdPlan {
    size_t x, y, qkv, att, hid, frame, step, kcache, vcache, mem, total;
    long long cache_layer, mem_layer;
};

ArdPlan ard_plan(const hfg_ard_config& c, int B, int T, int max_len) {
    ArdPlan p{};
    size_t cur = 0;
    auto take = [&](size_t floats) { const size_t off = cur; cur += (floats * 4 + 255) / 256 * 256; return off; };
    p.x = take((size_t)B * c.d_model);
    p.y = take((size_t)B * c.d_model * std::max(1, (c.d_ff + 255) / 256));       // also the split-K partial sums of linear2
    p.qkv = take((size_t)B * 3 * c.d_model);
    p.att = take((size_t)B * c.d_model);
    p.hid = take((size_t)B * std::max(c.d_ff, c.d_model));
    p.frame = take((size_t)B * c.n_mels);
    p.step = take(64);
    p.cache_layer = (long long)B * max_len * c.d_model;
    p.kcache = take((size_t)c.n_layers * p.cache_layer);
    p.vcache = take((size_t)c.n_layers * p.cache_layer);
    p.mem_layer = (long long)B * T * 2 * c.d_model;
    p.mem = take((size_t)c.n_layers * p.mem_layer);
    p.total = cur;
    return p;
}

const HostTensor& need(const hfg_ard_handle* h, const std::string& key, std::vector<int64_t> want) {
    auto it = h->sd.find(key);
    if (it == h->sd.end()) throw StatusError(HFG_ERR_STATE, "missing state_dict key: " + key);
    if (it->second.shape != want) throw StatusError(HFG_ERR_INVALID, "shape mismatch for " + key);
    return it->second;
}

}  // namespace

#define ARD_TRY try {
#define ARD_CATCH(h)                                                         \
    } catch (const StatusError& e) {                                         \
        if (h) (h)->last_error = e.what();                                   \
        return e.code;                                                       \
    } catch (const std::exception& e) {                                      \
        if (h) (h)->last_error = e.what();                                   \
        return HFG_ERR_INVALID;                                              \
    }                                                                        \
    return HFG_OK;

extern "C" {

int hfg_ard_create(const hfg_ard_config* cfg, hfg_ard_handle** out) {
    if (!cfg || !out) return HFG_ERR_INVALID;
    *out = nullptr;
    const hfg_ard_config& c = *cfg;
    if (c.d_model <= 0 || c.d_model > 1024 || c.n_mels <= 0 || c.n_layers <= 0 || c.n_heads <= 0 || c.d_ff <= 0 ||
        c.max_pos <= 0 || c.d_model % c.n_heads != 0)
        return HFG_ERR_INVALID;
    const int hd = c.d_model / c.n_heads;
    if (hd != 16 && hd != 32 && hd != 64 && hd != 128) return HFG_ERR_UNSUPPORTED;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return HFG_ERR_CUDA; }
    hfg_ard_handle* h = new hfg_ard_handle();
    h->cfg = c;
    h->device = dev;
    *out = h;
    return HFG_OK;
}

void hfg_ard_destroy(hfg_ard_handle* h) {
    if (!h) return;
    h->free_weights();
    h->free_stream();
    delete h;
}

const char* hfg_ard_last_error(const hfg_ard_handle* h) { return h ? h->last_error.c_str() : "null handle"; }

int hfg_ard_set_weight(hfg_ard_handle* h, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
    if (!h) return HFG_ERR_INVALID;
    ARD_TRY
    if (!name || !data || !shape || ndim < 1 || ndim > 3) throw StatusError(HFG_ERR_INVALID, "hfg_ard_set_weight: bad argument");
    HostTensor t;
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] <= 0) throw StatusError(HFG_ERR_INVALID, "hfg_ard_set_weight: non-positive dim");
        t.shape.push_back(shape[i]);
        n *= shape[i];
    }
    t.data.assign(data, data + n);
    h->sd[name] = std::move(t);
    h->committed = false;
    ARD_CATCH(h)
}

int hfg_ard_commit_weights(hfg_ard_handle* h) {
    if (!h) return HFG_ERR_INVALID;
    ARD_TRY
    check_cuda(cudaSetDevice(h->device), "cudaSetDevice");
    h->free_weights();
    const hfg_ard_config& c = h->cfg;
    const int64_t d = c.d_model;
    auto put = [&](const std::string& key, std::vector<int64_t> shape) { h->dev[key] = h->upload(need(h, key, shape).data); };
    put("prenet.0.weight", {d, c.n_mels}); put("prenet.0.bias", {d});
    put("prenet.3.weight", {d, d}); put("prenet.3.bias", {d});
    put("pos_encoding.pe", {1, c.max_pos, d});
    for (int i = 0; i < c.n_layers; ++i) {
        const std::string L = "decoder.layers." + std::to_string(i) + ".";
        for (const char* a : {"self_attn.", "multihead_attn."}) {
            put(L + a + "in_proj_weight", {3 * d, d}); put(L + a + "in_proj_bias", {3 * d});
            put(L + a + "out_proj.weight", {d, d}); put(L + a + "out_proj.bias", {d});
        }
        put(L + "linear1.weight", {c.d_ff, d}); put(L + "linear1.bias", {c.d_ff});
        put(L + "linear2.weight", {d, c.d_ff}); put(L + "linear2.bias", {d});
        for (const char* nrm : {"norm1.", "norm2.", "norm3."}) { put(L + nrm + "weight", {d}); put(L + nrm + "bias", {d}); }
    }
    put("mel_proj.weight", {c.n_mels, d}); put("mel_proj.bias", {c.n_mels});
    h->committed = true;
    ARD_CATCH(h)
}

int hfg_ard_workspace_bytes(const hfg_ard_handle* hc, int32_t batch, int32_t frames, int32_t max_len, size_t* bytes) {
    hfg_ard_handle* h = const_cast<hfg_ard_handle*>(hc);
    if (!h || !bytes) return HFG_ERR_INVALID;
    ARD_TRY
    if (batch <= 0 || frames <= 0 || max_len <= 0) throw StatusError(HFG_ERR_INVALID, "batch, frames and max_len must be positive");
    *bytes = ard_plan(h->cfg, batch, frames, max_len).total;
    ARD_CATCH(h)
}

int hfg_ard_decode(hfg_ard_handle* h, const float* hvar, int32_t B, int32_t T, int32_t max_len, float* mel, void* ws_v,
                   size_t ws_bytes, void* stream) {
    if (!h) return HFG_ERR_INVALID;
    ARD_TRY
    if (!h->committed) throw StatusError(HFG_ERR_STATE, "weights not committed (call hfg_ard_commit_weights)");
    if (!hvar || !mel || B <= 0 || T <= 0 || max_len <= 0) throw StatusError(HFG_ERR_INVALID, "hfg_ard_decode: bad argument");
    if (B > 65535) throw StatusError(HFG_ERR_INVALID, "batch > 65535: split the call");
    const hfg_ard_config& c = h->cfg;
    if (max_len > c.max_pos) throw StatusError(HFG_ERR_INVALID, "max_len exceeds the positional-encoding table (reference: 5000 rows)");
    const ArdPlan p = ard_plan(c, B, T, max_len);
    if (!ws_v || ws_bytes < p.total) throw StatusError(HFG_ERR_WORKSPACE, "workspace too small");
    if (((uintptr_t)ws_v & 255) != 0) throw StatusError(HFG_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    char* ws = (char*)ws_v;
    check_cuda(cudaSetDevice(h->device), "cudaSetDevice");
    h->ensure_stream();
    cudaStream_t caller = (cudaStream_t)stream, st = h->work;
    check_cuda(cudaEventRecord(h->ev_in, caller), "cudaEventRecord(in)");
    check_cuda(cudaStreamWaitEvent(st, h->ev_in, 0), "cudaStreamWaitEvent(in)");
    const int d = c.d_model, hd = d / c.n_heads;
    const float scale = 1.0f / std::sqrt((float)hd);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    float *x = F(p.x), *y = F(p.y), *qkv = F(p.qkv), *att = F(p.att), *hid = F(p.hid), *frame = F(p.frame);
    int* step = reinterpret_cast<int*>(ws + p.step);
    auto W = [&](const std::string& k) { return h->dev.at(k); };
    int64_t per_step = 0, setup = 0;
    int64_t* counter = &setup;
    auto gemm = [&](bool relu, const float* X, int ldx, const float* Wt, const float* b, const float* add_row, long long add_stride,
                    const int* stp, float* Y, int ldy, int M, int N, int K, bool split_k = false) {
        if (M > 256) {                                     // many rows: wide tiles
            dim3 grid((N + 63) / 64, (M + 63) / 64);
            if (relu) ard_gemm<true, 64><<<grid, 256, 0, st>>>(X, ldx, Wt, b, add_row, add_stride, stp, Y, ldy, M, N, K);
            else ard_gemm<false, 64><<<grid, 256, 0, st>>>(X, ldx, Wt, b, add_row, add_stride, stp, Y, ldy, M, N, K);
        } else {                                           // a decode step: 16 columns x one 256-wide K chunk per block
            dim3 grid((N + kStepTN - 1) / kStepTN, (M + 63) / 64, split_k ? (K + kStepKC - 1) / kStepKC : 1);
            const size_t smem = (size_t)(64 + kStepTN) * (kStepKC + 1) * sizeof(float);
            const long long part_stride = (long long)M * ldy;
            if (relu) ard_gemm_step<true><<<grid, 256, smem, st>>>(X, ldx, Wt, b, add_row, add_stride, stp, Y, ldy, part_stride, M, N, K);
            else ard_gemm_step<false><<<grid, 256, smem, st>>>(X, ldx, Wt, b, add_row, add_stride, stp, Y, ldy, part_stride, M, N, K);
        }
        check_cuda(cudaGetLastError(), "ard_gemm launch");
        ++*counter;
    };
    const size_t attn_smem_self = ((size_t)max_len + 128) * sizeof(float), attn_smem_cross = ((size_t)T + 128) * sizeof(float);
    if (std::max(attn_smem_self, attn_smem_cross) > 200 * 1024)
        throw StatusError(HFG_ERR_UNSUPPORTED, "sequence too long for the attention kernel's score buffer");
    check_cuda(cudaFuncSetAttribute(ard_attention, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), "attr");
    check_cuda(cudaFuncSetAttribute(ard_gemm_step<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024), "attr");
    check_cuda(cudaFuncSetAttribute(ard_gemm_step<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024), "attr");

    // ---- once per call: start token, step counter, encoder memory projected per layer ([B*T, 2d]: K | V) ----
    check_cuda(cudaMemsetAsync(frame, 0, (size_t)B * c.n_mels * sizeof(float), st), "memset(frame)");
    check_cuda(cudaMemsetAsync(step, 0, 256, st), "memset(step)");
    for (int i = 0; i < c.n_layers; ++i) {
        const std::string L = "decoder.layers." + std::to_string(i) + ".multihead_attn.";
        gemm(false, hvar, d, W(L + "in_proj_weight") + (size_t)d * d, W(L + "in_proj_bias") + d, nullptr, 0, nullptr,
             F(p.mem) + (size_t)i * p.mem_layer, 2 * d, B * T, 2 * d, d);
    }
    // ---- one step, captured into a graph ----
    counter = &per_step;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    check_cuda(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed), "cudaStreamBeginCapture");
    bool ok = true;
    std::string why;
    try {
        gemm(true, frame, c.n_mels, W("prenet.0.weight"), W("prenet.0.bias"), nullptr, 0, nullptr, hid, d, B, d, c.n_mels);
        gemm(false, hid, d, W("prenet.3.weight"), W("prenet.3.bias"), W("pos_encoding.pe"), d, step, x, d, B, d, d);
        for (int i = 0; i < c.n_layers; ++i) {
            const std::string L = "decoder.layers." + std::to_string(i) + ".";
            float* Kc = F(p.kcache) + (size_t)i * p.cache_layer;
            float* Vc = F(p.vcache) + (size_t)i * p.cache_layer;
            float* mem = F(p.mem) + (size_t)i * p.mem_layer;
            auto add_ln = [&](const char* nrm, int n_part = 1, const float* rbias = nullptr) {
                ard_add_layernorm<<<(B + 3) / 4, 128, 0, st>>>(x, y, n_part, (long long)B * d, rbias, W(L + nrm + "weight"),
                                                               W(L + nrm + "bias"), x, B, d, 1e-5f);
                check_cuda(cudaGetLastError(), "ard_add_layernorm launch");
                ++*counter;
            };
            // self-attention over the cache
            gemm(false, x, d, W(L + "self_attn.in_proj_weight"), W(L + "self_attn.in_proj_bias"), nullptr, 0, nullptr, qkv, 3 * d, B, 3 * d, d);
            ard_attention<<<dim3(c.n_heads, B), 128, attn_smem_self, st>>>(qkv, 3 * d, qkv + d, qkv + 2 * d, Kc, Vc,
                                                                            (long long)max_len * d, d, step, 0, att, d, hd, scale);
            check_cuda(cudaGetLastError(), "ard_attention launch");
            ++*counter;
            gemm(false, att, d, W(L + "self_attn.out_proj.weight"), W(L + "self_attn.out_proj.bias"), nullptr, 0, nullptr, y, d, B, d, d);
            add_ln("norm1.");
            // cross-attention over the projected encoder memory (no mask: reference passes none)
            gemm(false, x, d, W(L + "multihead_attn.in_proj_weight"), W(L + "multihead_attn.in_proj_bias"), nullptr, 0, nullptr, qkv, d, B, d, d);
            ard_attention<<<dim3(c.n_heads, B), 128, attn_smem_cross, st>>>(qkv, d, nullptr, nullptr, mem, mem + d,
                                                                             (long long)T * 2 * d, 2 * d, nullptr, T, att, d, hd, scale);
            check_cuda(cudaGetLastError(), "ard_attention launch");
            ++*counter;
            gemm(false, att, d, W(L + "multihead_attn.out_proj.weight"), W(L + "multihead_attn.out_proj.bias"), nullptr, 0, nullptr, y, d, B, d, d);
            add_ln("norm2.");
            // feed-forward
            gemm(true, x, d, W(L + "linear1.weight"), W(L + "linear1.bias"), nullptr, 0, nullptr, hid, c.d_ff, B, c.d_ff, d);
            const int ff_parts = B <= 256 ? (c.d_ff + kStepKC - 1) / kStepKC : 1;   // split-K partial sums land in y[z]
            gemm(false, hid, c.d_ff, W(L + "linear2.weight"), W(L + "linear2.bias"), nullptr, 0, nullptr, y, d, B, d, c.d_ff,
                 /*split_k=*/ff_parts > 1);
            add_ln("norm3.", ff_parts, ff_parts > 1 ? W(L + "linear2.bias") : nullptr);
        }
        gemm(false, x, d, W("mel_proj.weight"), W("mel_proj.bias"), nullptr, 0, nullptr, frame, c.n_mels, B, c.n_mels, d);
        ard_store_frame<<<std::min(64, (B * c.n_mels + 255) / 256), 256, 0, st>>>(frame, mel, B, c.n_mels, max_len, step);
        ard_advance<<<1, 1, 0, st>>>(step);
        check_cuda(cudaGetLastError(), "ard_store_frame launch");
        per_step += 2;
    } catch (const std::exception& e) { ok = false; why = e.what(); }
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (!ok || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        throw StatusError(HFG_ERR_CUDA, "could not capture the decode step: " + (ok ? std::string(cudaGetErrorString(ce)) : why));
    }
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, nullptr, nullptr, 0);
    cudaGraphDestroy(graph);
    check_cuda(ie, "cudaGraphInstantiate");
    cudaError_t le = cudaSuccess;
    for (int t = 0; t < max_len && le == cudaSuccess; ++t) le = cudaGraphLaunch(exec, st);
    // the exec object may be destroyed once its launches are enqueued (the runtime keeps what is in flight alive)
    cudaGraphExecDestroy(exec);
    check_cuda(le, "cudaGraphLaunch");
    check_cuda(cudaEventRecord(h->ev_out, st), "cudaEventRecord(out)");
    check_cuda(cudaStreamWaitEvent(caller, h->ev_out, 0), "cudaStreamWaitEvent(out)");
    h->launches = setup + per_step * max_len;
    ARD_CATCH(h)
}

int hfg_ard_last_launch_count(const hfg_ard_handle* h, int64_t* launches) {
    if (!h || !launches) return HFG_ERR_INVALID;
    *launches = h->launches;
    return HFG_OK;
}

}  // extern "C"
