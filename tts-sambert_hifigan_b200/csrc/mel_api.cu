// On-device log-mel and log-mel L1 (include/hfg_mel.h).
// Reference: torchaudio MelSpectrogram(center = True, reflect padding, periodic Hann, power 2, slaney filterbank)
// followed by log10(. + 1e-10) (data/audio_processing.py:99-127) and the L1 between two such spectrograms
// (models/losses.py:708-797).
#include "../../include/hfg_mel.h"

#include <cuda_runtime.h>

#include <cmath>
#include <string>
#include <vector>

#include "model.h"

namespace hfg {

constexpr int kMelThreads = 256;

struct MelArgs {
    const float* wav[2];      // [B, T]; wav[1] null for plain extraction
    float* out;               // [B, n_mels, frames] (extraction) or per-frame |a - b| sums [B * frames] (L1)
    const float* window;      // [n_fft]
    const float2* twiddle;    // [n_fft / 2]: exp(-2 pi i j / n_fft)
    const float* fb;          // [n_mels][n_fft / 2 + 1]
    const int* fb_lo; const int* fb_hi;   // non-zero bin range of each filter
    long long T; int frames, n_fft, log2n, hop, n_mels;
};

__device__ __forceinline__ unsigned bitrev(unsigned v, int bits) { return __brev(v) >> (32 - bits); }

// one block per (frame, utterance): framing -> FFT -> power -> filterbank -> log10, for one or two waveforms
__global__ void __launch_bounds__(kMelThreads)
mel_frame_kernel(const MelArgs a) {
    extern __shared__ float2 xs[];                           // [n_fft] complex, then power [n_fft / 2 + 1]
    __shared__ float red[kMelThreads / 32];
    const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int N = a.n_fft, half = N >> 1;
    float* pw = reinterpret_cast<float*>(xs + N);
    float diff = 0.f;
    float first[1] = {0.f};
    for (int w = 0; w < 2; ++w) {
        if (!a.wav[w]) break;
        const float* x = a.wav[w] + (long long)b * a.T;
        for (int n = tid; n < N; n += kMelThreads) {
            long long idx = (long long)f * a.hop - half + n;  // center = True
            if (idx < 0) idx = -idx;                          // reflect (no edge repeat), as torch.nn.functional.pad
            if (idx >= a.T) idx = 2 * (a.T - 1) - idx;
            xs[bitrev((unsigned)n, a.log2n)] = make_float2(x[idx] * a.window[n], 0.f);
        }
        __syncthreads();
        for (int s = 1; s <= a.log2n; ++s) {                  // radix-2 decimation in time
            const int m = 1 << s, hm = m >> 1, tstep = N >> s;
            for (int j = tid; j < half; j += kMelThreads) {
                const int pos = j & (hm - 1), i0 = ((j >> (s - 1)) << s) + pos, i1 = i0 + hm;
                const float2 wv = a.twiddle[pos * tstep];
                const float2 u = xs[i0], v = xs[i1];
                const float2 t = make_float2(wv.x * v.x - wv.y * v.y, wv.x * v.y + wv.y * v.x);
                xs[i0] = make_float2(u.x + t.x, u.y + t.y);
                xs[i1] = make_float2(u.x - t.x, u.y - t.y);
            }
            __syncthreads();
        }
        for (int k = tid; k <= half; k += kMelThreads) pw[k] = xs[k].x * xs[k].x + xs[k].y * xs[k].y;
        __syncthreads();
        for (int m = tid; m < a.n_mels; m += kMelThreads) {   // n_mels <= 256: at most one filter per thread
            const float* fr = a.fb + (size_t)m * (half + 1);
            float acc = 0.f;
            for (int k = a.fb_lo[m]; k < a.fb_hi[m]; ++k) acc = fmaf(fr[k], pw[k], acc);
            const float lm = log10f(acc + 1e-10f);
            if (!a.wav[1]) a.out[((size_t)b * a.n_mels + m) * a.frames + f] = lm;
            else if (w == 0) first[0] = lm;
            else diff = fabsf(lm - first[0]);
        }
        __syncthreads();
    }
    if (a.wav[1]) {                                           // deterministic block sum of |a - b|
        float s = diff;
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((tid & 31) == 0) red[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int i = 0; i < kMelThreads / 32; ++i) t += red[i];
            a.out[(size_t)b * a.frames + f] = t;
        }
    }
}

// fixed-order sum of the per-frame partials -> mean
__global__ void mel_l1_finish(const float* __restrict__ part, long long n, double inv_count, float* __restrict__ loss) {
    __shared__ double red[256];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) s += (double)part[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = (float)(red[0] * inv_count);
}

// slaney mel scale (torchaudio.functional.functional._hz_to_mel / _mel_to_hz, mel_scale = "slaney")
static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

}  // namespace hfg

using namespace hfg;

struct hfg_mel_handle {
    hfg_mel_config cfg{};
    int device = 0, log2n = 0;
    std::string last_error;
    float *window = nullptr, *fb = nullptr;
    float2* twiddle = nullptr;
    int *fb_lo = nullptr, *fb_hi = nullptr;
    void release() {
        for (void* p : {(void*)window, (void*)fb, (void*)twiddle, (void*)fb_lo, (void*)fb_hi})
            if (p) cudaFree(p);
        window = fb = nullptr; twiddle = nullptr; fb_lo = fb_hi = nullptr;
    }
};

#define MEL_TRY try {
#define MEL_CATCH(h)                                                         \
    } catch (const StatusError& e) {                                         \
        if (h) (h)->last_error = e.what();                                   \
        return e.code;                                                       \
    } catch (const std::exception& e) {                                      \
        if (h) (h)->last_error = e.what();                                   \
        return HFG_ERR_INVALID;                                              \
    }                                                                        \
    return HFG_OK;

template <typename T>
static T* to_device(const std::vector<T>& v) {
    T* p = nullptr;
    check_cuda(cudaMalloc((void**)&p, v.size() * sizeof(T)), "cudaMalloc(mel tables)");
    check_cuda(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy(mel tables)");
    return p;
}

extern "C" {

int hfg_mel_create(const hfg_mel_config* cfg, hfg_mel_handle** out) {
    if (!cfg || !out) return HFG_ERR_INVALID;
    *out = nullptr;
    const hfg_mel_config& c = *cfg;
    int log2n = 0;
    while ((1 << log2n) < c.n_fft) ++log2n;
    if (c.n_fft < 64 || c.n_fft > 4096 || (1 << log2n) != c.n_fft || c.hop_length <= 0 || c.n_mels <= 0 || c.n_mels > 256 ||
        c.sample_rate <= 0 || !(c.fmax > c.fmin) || c.fmin < 0)
        return HFG_ERR_INVALID;
    if (c.win_length != c.n_fft) return HFG_ERR_UNSUPPORTED;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return HFG_ERR_CUDA; }
    hfg_mel_handle* h = new hfg_mel_handle();
    h->cfg = c; h->device = dev; h->log2n = log2n;
    try {
        const int N = c.n_fft, half = N / 2, n_freqs = half + 1;
        const double pi = 3.14159265358979323846;
        std::vector<float> win(N);
        for (int n = 0; n < N; ++n) win[n] = (float)(0.5 - 0.5 * std::cos(2.0 * pi * n / N));      // torch.hann_window(periodic=True)
        std::vector<float2> tw(half);
        for (int j = 0; j < half; ++j) tw[j] = make_float2((float)std::cos(2.0 * pi * j / N), (float)(-std::sin(2.0 * pi * j / N)));
        // torchaudio.functional.melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate, norm="slaney", mel_scale="slaney")
        std::vector<double> f_pts(c.n_mels + 2);
        const double m_min = hz_to_mel(c.fmin), m_max = hz_to_mel(c.fmax);
        for (int i = 0; i < c.n_mels + 2; ++i) f_pts[i] = mel_to_hz(m_min + (m_max - m_min) * i / (c.n_mels + 1));
        std::vector<float> fb((size_t)c.n_mels * n_freqs, 0.f);
        std::vector<int> lo(c.n_mels, n_freqs), hi(c.n_mels, 0);
        for (int m = 0; m < c.n_mels; ++m) {
            const double enorm = 2.0 / (f_pts[m + 2] - f_pts[m]);
            for (int k = 0; k < n_freqs; ++k) {
                const double fr = (double)(c.sample_rate / 2) * k / (n_freqs - 1);                  // torch.linspace(0, sr // 2, n_freqs)
                const double down = (fr - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
                const double up = (f_pts[m + 2] - fr) / (f_pts[m + 2] - f_pts[m + 1]);
                const double v = std::max(0.0, std::min(down, up)) * enorm;
                if (v > 0.0) {
                    fb[(size_t)m * n_freqs + k] = (float)v;
                    lo[m] = std::min(lo[m], k);
                    hi[m] = std::max(hi[m], k + 1);
                }
            }
            if (hi[m] == 0) lo[m] = 0;
        }
        h->window = to_device(win); h->twiddle = to_device(tw); h->fb = to_device(fb);
        h->fb_lo = to_device(lo); h->fb_hi = to_device(hi);
        check_cuda(cudaFuncSetAttribute(mel_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), "attr");
    } catch (...) {
        h->release();
        delete h;
        return HFG_ERR_CUDA;
    }
    *out = h;
    return HFG_OK;
}

void hfg_mel_destroy(hfg_mel_handle* h) {
    if (!h) return;
    h->release();
    delete h;
}

const char* hfg_mel_last_error(const hfg_mel_handle* h) { return h ? h->last_error.c_str() : "null handle"; }

int hfg_mel_frames(const hfg_mel_handle* h, int64_t samples, int64_t* frames) {
    if (!h || !frames || samples <= 0) return HFG_ERR_INVALID;
    *frames = samples / h->cfg.hop_length + 1;
    return HFG_OK;
}

static void mel_launch(hfg_mel_handle* h, const float* w0, const float* w1, int B, int64_t T, float* out, cudaStream_t st) {
    if (!w0 || !out || B <= 0 || B > 65535) throw StatusError(HFG_ERR_INVALID, "bad argument");
    if (T <= h->cfg.n_fft / 2) throw StatusError(HFG_ERR_INVALID, "waveform shorter than n_fft / 2 + 1 samples (reflect padding)");
    MelArgs a{};
    a.wav[0] = w0; a.wav[1] = w1; a.out = out;
    a.window = h->window; a.twiddle = h->twiddle; a.fb = h->fb; a.fb_lo = h->fb_lo; a.fb_hi = h->fb_hi;
    a.T = T; a.frames = (int)(T / h->cfg.hop_length + 1); a.n_fft = h->cfg.n_fft; a.log2n = h->log2n;
    a.hop = h->cfg.hop_length; a.n_mels = h->cfg.n_mels;
    const size_t smem = (size_t)a.n_fft * sizeof(float2) + (size_t)(a.n_fft / 2 + 1) * sizeof(float);
    mel_frame_kernel<<<dim3(a.frames, B), kMelThreads, smem, st>>>(a);
    check_cuda(cudaGetLastError(), "mel_frame_kernel launch");
}

int hfg_log_mel(hfg_mel_handle* h, const float* wav, int32_t B, int64_t T, float* out, void* stream) {
    if (!h) return HFG_ERR_INVALID;
    MEL_TRY
    mel_launch(h, wav, nullptr, B, T, out, (cudaStream_t)stream);
    MEL_CATCH(h)
}

int hfg_log_mel_l1(hfg_mel_handle* h, const float* ref, const float* neu, int32_t B, int64_t T, float* loss, float* scratch,
                   void* stream) {
    if (!h) return HFG_ERR_INVALID;
    MEL_TRY
    if (!neu || !loss || !scratch) throw StatusError(HFG_ERR_INVALID, "bad argument");
    mel_launch(h, ref, neu, B, T, scratch, (cudaStream_t)stream);
    const long long frames = T / h->cfg.hop_length + 1, n = (long long)B * frames;
    mel_l1_finish<<<1, 256, 0, (cudaStream_t)stream>>>(scratch, n, 1.0 / ((double)n * h->cfg.n_mels), loss);
    check_cuda(cudaGetLastError(), "mel_l1_finish launch");
    MEL_CATCH(h)
}

}  // extern "C"
