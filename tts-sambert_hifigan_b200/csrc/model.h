// Host-side model state behind an hfg_handle.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hfg.h"

namespace hfg {

struct StatusError : std::runtime_error {
    int code;
    StatusError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline void check_cuda(cudaError_t e, const char* what) {
    if (e != cudaSuccess)
        throw StatusError(HFG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

struct HostTensor {
    std::vector<int64_t> shape;
    std::vector<float> data;
    int64_t numel() const {
        int64_t n = 1;
        for (auto d : shape) n *= d;
        return n;
    }
};

// Tensor-core operand pack of one conv (tc_path.cuh): per tap, an [N=Cout][K=Cin]
// K-major matrix in the UMMA no-swizzle "interleaved" canonical layout.
struct TcPack {
    // indexed by operand precision (tc_kernels.cuh: PREC_TF32 = 0 fp32 cells consumed as tf32, PREC_BF16 = 1, PREC_FP16 = 2)
    void* w[3] = {};          // operand tiles of the plain conv kernels
    // fused-pair kernel packs: [precision][N-halves for a cta_group::2 pair ? 1 : 0][K block of 4 cells ? 1 : 0]
    void* w_pair[3][2][2] = {};
    long long half_stride[3][2] = {};     // [precision][kbc4]: bytes between the two N-halves
    // conv2 of a narrow fused pair in space-to-depth form (tc_pair_kernel.cuh): (k + 1) / 2 taps of a 2N x 2N matrix,
    // [precision][N-halves ? 1 : 0]; null where not packed
    void* w_s2d[3][2] = {};
    long long s2d_half_stride[3] = {};
    bool ok = false;          // layer shape is covered by the tensor-core path
};

struct ConvLayer {
    int cin = 0, cout = 0, k = 0, dil = 1, pad = 0;
    int cin_pad = 0, cout_pad = 0, rco = 2;
    float* w_fp32 = nullptr;  // [k][cin_pad][cout_pad]
    float* bias = nullptr;    // [cout]
    TcPack tc;
};

struct UpLayer {
    int cin = 0, cout = 0, k = 0, u = 1, p = 0, taps_max = 0;
    int cin_pad = 0, cout_pad = 0, rco = 2;
    float* w_fp32 = nullptr;  // [u][taps_max][cin_pad][cout_pad]
    float* bias = nullptr;
    TcPack tc;
    TcPack tc_stack;          // the u polyphase tap sets stacked along N (u * cout virtual output channels)
};

struct PairLayers {
    ConvLayer c1, c2;
};

}  // namespace hfg

struct hfg_handle {
    hfg_config cfg{};
    int device = 0;
    int sm_count = 0;
    int cc_major = 0;
    bool committed = false;
    std::string last_error;
    int64_t launches = 0;
    int mel_layout = 0;       // 0 = [B, n_mels, T] (reference), 1 = [B, T, n_mels] (acoustic-model output)
    unsigned long long* pair_timeline = nullptr;   // tuning only: phase stamps of the fused pair kernel (hfg_bench_layer)
    int tf32_split = -1;      // HFG_MODE_TF32 on fp16 hi + lo planes (tc_path.cuh: tc_tf32_mixed): -1 = not evaluated yet
    int pair_regs[5][3][2] = {};                   // registers per thread of tc_pair_kernel<P, P2, MINB, CTAS> (index 3: tf32 with an fp16 intermediate, 4: fp16 operands with fp32 twin planes), filled lazily

    std::map<std::string, hfg::HostTensor> sd;   // raw state_dict as set by the caller

    hfg::ConvLayer pre;
    std::vector<hfg::UpLayer> ups;
    std::vector<std::vector<std::vector<hfg::PairLayers>>> mrfs;   // [stage][resblock][pair]
    int post_cin = 0;
    float* post_w = nullptr;
    float* post_b = nullptr;
    hfg::TcPack post_tc;

    std::vector<void*> device_allocs;

    // optional per-launch timing (hfg_set_profiling): one event pair per launch
    struct ProfRec { std::string label; double flops; double bytes; cudaEvent_t e0, e1; };
    int profiling = 0;        // 0 off, 1 per launch (launches serialised), 2 per stage (streams stay on)
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> event_pool;
    size_t events_used = 0;
    cudaEvent_t next_event() {
        if (events_used == event_pool.size()) {
            cudaEvent_t e;
            hfg::check_cuda(cudaEventCreate(&e), "cudaEventCreate");
            event_pool.push_back(e);
        }
        return event_pool[events_used++];
    }
    // stage-level records (profiling == 2): events on the caller's stream around a group of launches
    void stage_begin(cudaStream_t st, const char* label) {
        if (profiling != 2) return;
        ProfRec r{label, 0.0, 0.0, next_event(), next_event()};
        hfg::check_cuda(cudaEventRecord(r.e0, st), "cudaEventRecord");
        prof.push_back(r);
    }
    void stage_end(cudaStream_t st) {
        if (profiling != 2) return;
        hfg::check_cuda(cudaEventRecord(prof.back().e1, st), "cudaEventRecord");
    }
    // call before / after a launch
    void prof_begin(cudaStream_t st, const char* label, double flops, double bytes) {
        launches++;
        if (profiling != 1) return;
        ProfRec r{label, flops, bytes, next_event(), next_event()};
        hfg::check_cuda(cudaEventRecord(r.e0, st), "cudaEventRecord");
        prof.push_back(r);
    }
    void prof_end(cudaStream_t st) {
        if (profiling != 1) return;
        hfg::check_cuda(cudaEventRecord(prof.back().e1, st), "cudaEventRecord");
    }

    // Side streams of the tensor-core path: the resblocks of one MRF are independent until their last pair
    // (reference models/hifigan.py:126-131), so resblock j runs on stream j % kStreams; fork / join and the
    // order of the running-sum updates are expressed with events (no host synchronisation).
    static constexpr int kStreams = 3;
    cudaStream_t side[kStreams - 1] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[kStreams - 1] = {nullptr, nullptr}, ev_sum[HFG_MAX_STAGES] = {};
    void ensure_streams() {
        if (ev_fork) return;
        using hfg::check_cuda;
        // side streams outrank the caller's stream: they carry the longer resblocks (the critical path), whose
        // CTAs should be placed first whenever SMs free up
        int prio_least = 0, prio_greatest = 0;
        check_cuda(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest), "cudaDeviceGetStreamPriorityRange");
        for (int i = 0; i < kStreams - 1; ++i)
            check_cuda(cudaStreamCreateWithPriority(&side[i], cudaStreamNonBlocking,
#ifdef HFG_TUNING
                                                    getenv("HFG_TC_STREAM_NOPRIO") ? prio_least :
#endif
                                                    std::max(prio_greatest, prio_least - 1 - i)),
                       "cudaStreamCreateWithPriority");
        check_cuda(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming), "cudaEventCreate");
        for (auto& e : ev_join) check_cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
        for (auto& e : ev_sum) check_cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    }
    void free_streams() {
        for (auto& s : side) { if (s) cudaStreamDestroy(s); s = nullptr; }
        if (ev_fork) cudaEventDestroy(ev_fork);
        ev_fork = nullptr;
        for (auto& e : ev_join) { if (e) cudaEventDestroy(e); e = nullptr; }
        for (auto& e : ev_sum) { if (e) cudaEventDestroy(e); e = nullptr; }
    }

    // hfg_forward_host: the library owns every device buffer and the stream, so the launch sequence of a
    // forward (about 50 kernels plus the fork / join events) is captured into a CUDA graph once per
    // (batch, frames, mode, layout, buffers) and replayed with ONE launch -- the host calls are synchronous,
    // so the ~0.4 ms of launch calls would otherwise sit on the critical path of every step.
    struct HostGraph {
        int B = 0, T = 0, mode = -1, layout = -1, calls = 0;
        const void *mel = nullptr, *wav = nullptr, *ws = nullptr;
        int64_t launches = 0;
        cudaGraphExec_t exec = nullptr;
        bool failed = false;
    } host_graph;
    // streaming host path (hfg_forward_host_submit / _wait): kHostSlots submissions in flight, each with its own
    // device mel / wav buffers and graph; one workspace and one compute stream, so forwards run back to back
    // while the neighbours' copies ride on two copy streams
    static constexpr int kHostSlots = 2;
    struct HostSlot {
        float* dev_mel = nullptr; size_t dev_mel_bytes = 0;
        float* dev_wav = nullptr; size_t dev_wav_bytes = 0;
        cudaEvent_t ev_h2d = nullptr, ev_done = nullptr, ev_d2h = nullptr;
        HostGraph graph;
        bool busy = false;
    } slots[kHostSlots];
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    void drop_host_graph() {
        if (host_graph.exec) cudaGraphExecDestroy(host_graph.exec);
        host_graph = HostGraph{};
        for (auto& s : slots) {
            if (s.graph.exec) cudaGraphExecDestroy(s.graph.exec);
            s.graph = HostGraph{};
        }
    }

    // hfg_forward_host resources
    cudaStream_t stream = nullptr;
    float* pin_mel = nullptr; size_t pin_mel_bytes = 0;
    float* pin_wav = nullptr; size_t pin_wav_bytes = 0;
    float* dev_mel = nullptr; size_t dev_mel_bytes = 0;
    float* dev_wav = nullptr; size_t dev_wav_bytes = 0;
    void* dev_ws = nullptr;   size_t dev_ws_bytes = 0;

    void* alloc_device(size_t bytes) {
        void* p = nullptr;
        hfg::check_cuda(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc(weights)");
        device_allocs.push_back(p);
        return p;
    }
    template <typename T>
    T* upload(const std::vector<T>& v) {
        T* p = (T*)alloc_device(v.size() * sizeof(T));
        hfg::check_cuda(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice),
                        "cudaMemcpy(weights)");
        return p;
    }
    void free_device_weights() {
        drop_host_graph();                                  // the captured launches point at these weights
        for (void* p : device_allocs) cudaFree(p);
        device_allocs.clear();
        committed = false;
    }
    void ensure_host_path(size_t pin_mel_need, size_t pin_wav_need, size_t mel_bytes, size_t wav_bytes,
                          size_t ws_bytes) {
        using hfg::check_cuda;
        if (!stream) check_cuda(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
        auto grow_pin = [](float*& p, size_t& have, size_t want) {
            if (have >= want) return;
            if (p) cudaFreeHost(p);
            p = nullptr; have = 0;
            check_cuda(cudaMallocHost((void**)&p, want), "cudaMallocHost");
            have = want;
        };
        auto grow_dev = [](void** p, size_t& have, size_t want) {
            if (have >= want) return;
            if (*p) cudaFree(*p);
            *p = nullptr; have = 0;
            check_cuda(cudaMalloc(p, want), "cudaMalloc");
            have = want;
        };
        if (pin_mel_need) grow_pin(pin_mel, pin_mel_bytes, pin_mel_need);
        if (pin_wav_need) grow_pin(pin_wav, pin_wav_bytes, pin_wav_need);
        grow_dev((void**)&dev_mel, dev_mel_bytes, mel_bytes);
        grow_dev((void**)&dev_wav, dev_wav_bytes, wav_bytes);
        grow_dev(&dev_ws, dev_ws_bytes, ws_bytes);
    }
    void ensure_stream_path(int slot, size_t mel_bytes, size_t wav_bytes, size_t ws_bytes) {
        using hfg::check_cuda;
        if (!stream) check_cuda(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
        if (!copy_in) check_cuda(cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking), "cudaStreamCreate");
        if (!copy_out) check_cuda(cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking), "cudaStreamCreate");
        HostSlot& s = slots[slot];
        for (cudaEvent_t* e : {&s.ev_h2d, &s.ev_done, &s.ev_d2h})
            if (!*e) check_cuda(cudaEventCreateWithFlags(e, cudaEventDisableTiming), "cudaEventCreate");
        auto grow = [&](void** p, size_t& have, size_t want, bool shared) {
            if (have >= want) return;
            // a shared buffer (the workspace) may be in use by the other slot's forward: drain the compute stream first
            if (shared) check_cuda(cudaStreamSynchronize(stream), "stream sync");
            if (*p) cudaFree(*p);
            *p = nullptr; have = 0;
            check_cuda(cudaMalloc(p, want), "cudaMalloc");
            have = want;
        };
        grow((void**)&s.dev_mel, s.dev_mel_bytes, mel_bytes, false);
        grow((void**)&s.dev_wav, s.dev_wav_bytes, wav_bytes, false);
        grow(&dev_ws, dev_ws_bytes, ws_bytes, true);
    }
    void free_host_path() {
        drop_host_graph();
        for (auto& s : slots) {
            if (s.dev_mel) cudaFree(s.dev_mel);
            if (s.dev_wav) cudaFree(s.dev_wav);
            for (cudaEvent_t e : {s.ev_h2d, s.ev_done, s.ev_d2h}) if (e) cudaEventDestroy(e);
            s = HostSlot{};
        }
        if (copy_in) cudaStreamDestroy(copy_in);
        if (copy_out) cudaStreamDestroy(copy_out);
        copy_in = copy_out = nullptr;
        if (pin_mel) cudaFreeHost(pin_mel);
        if (pin_wav) cudaFreeHost(pin_wav);
        if (dev_mel) cudaFree(dev_mel);
        if (dev_wav) cudaFree(dev_wav);
        if (dev_ws) cudaFree(dev_ws);
        if (stream) cudaStreamDestroy(stream);
        pin_mel = pin_wav = dev_mel = dev_wav = nullptr;
        dev_ws = nullptr; stream = nullptr;
        pin_mel_bytes = pin_wav_bytes = dev_mel_bytes = dev_wav_bytes = dev_ws_bytes = 0;
    }
};
