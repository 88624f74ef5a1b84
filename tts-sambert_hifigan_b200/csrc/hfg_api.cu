// C-ABI implementation (include/hfg.h): handle, state_dict ingestion, weight-norm
// folding, weight repacking, workspace planning and the per-mode launch plans.
#include "../../include/hfg.h"

#include <cuda_runtime.h>

#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "fp32_kernels.cuh"
#include "length_regulator.cuh"
#include "model.h"
#include "tc_path.cuh"

namespace hfg {

// ---------------------------------------------------------------------------
// state_dict ingestion
// ---------------------------------------------------------------------------

static const HostTensor& need(const hfg_handle* h, const std::string& key) {
    auto it = h->sd.find(key);
    if (it == h->sd.end()) throw StatusError(HFG_ERR_STATE, "missing state_dict key: " + key);
    return it->second;
}

// Plain weight of layer `prefix` ("ups.0", "mrfs.0.resblocks.1.convs1.2", ...):
// either `.weight`, or g*v/||v|| folded from `.weight_g`/`.weight_v`
// (nn.utils.weight_norm, dim=0; reference models/hifigan.py:274-283).
static HostTensor plain_weight(const hfg_handle* h, const std::string& prefix,
                               const std::vector<int64_t>& want) {
    HostTensor w;
    auto it = h->sd.find(prefix + ".weight");
    if (it != h->sd.end()) {
        w = it->second;
    } else {
        const HostTensor& g = need(h, prefix + ".weight_g");
        const HostTensor& v = need(h, prefix + ".weight_v");
        if (v.shape.size() != 3 || g.numel() != v.shape[0])
            throw StatusError(HFG_ERR_INVALID, "weight_g/weight_v shape mismatch at " + prefix);
        w = v;
        const size_t inner = (size_t)(v.shape[1] * v.shape[2]);
        for (int64_t i = 0; i < v.shape[0]; ++i) {
            double ss = 0.0;
            for (size_t e = 0; e < inner; ++e) {
                const double x = v.data[i * inner + e];
                ss += x * x;
            }
            const double s = (double)g.data[i] / std::sqrt(ss);
            for (size_t e = 0; e < inner; ++e)
                w.data[i * inner + e] = (float)(s * (double)v.data[i * inner + e]);
        }
    }
    if (w.shape != want) {
        std::string msg = "shape mismatch for " + prefix + ".weight: got [";
        for (auto d : w.shape) msg += std::to_string(d) + ",";
        msg += "] want [";
        for (auto d : want) msg += std::to_string(d) + ",";
        throw StatusError(HFG_ERR_INVALID, msg + "]");
    }
    return w;
}

static HostTensor plain_bias(const hfg_handle* h, const std::string& prefix, int64_t n) {
    const HostTensor& b = need(h, prefix + ".bias");
    if (b.numel() != n) throw StatusError(HFG_ERR_INVALID, "shape mismatch for " + prefix + ".bias");
    return b;
}

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

static int pick_rco(int cout) { return cout >= 128 ? 8 : (cout >= 64 ? 4 : 2); }

// Conv1d weight [Cout, Cin, k] -> fp32 tile layout [k][CinPad][CoutPad]
static void pack_conv_fp32(hfg_handle* h, ConvLayer& L, const HostTensor& w, const HostTensor& b) {
    L.rco = pick_rco(L.cout);
    L.cin_pad = round_up(L.cin, kTileCi);
    L.cout_pad = round_up(L.cout, L.rco * 16);
    std::vector<float> p((size_t)L.k * L.cin_pad * L.cout_pad, 0.f);
    for (int co = 0; co < L.cout; ++co)
        for (int ci = 0; ci < L.cin; ++ci)
            for (int j = 0; j < L.k; ++j)
                p[((size_t)j * L.cin_pad + ci) * L.cout_pad + co] =
                    w.data[((size_t)co * L.cin + ci) * L.k + j];
    L.w_fp32 = h->upload(p);
    L.bias = h->upload(b.data);
}

// ConvTranspose1d weight [Cin, Cout, k] -> polyphase [u][ceil(k/u)][CinPad][CoutPad]
static void pack_up_fp32(hfg_handle* h, UpLayer& L, const HostTensor& w, const HostTensor& b) {
    L.rco = pick_rco(L.cout);
    L.cin_pad = round_up(L.cin, kTileCi);
    L.cout_pad = round_up(L.cout, L.rco * 16);
    L.taps_max = (L.k + L.u - 1) / L.u;
    std::vector<float> p((size_t)L.u * L.taps_max * L.cin_pad * L.cout_pad, 0.f);
    for (int r = 0; r < L.u; ++r)
        for (int m = 0; r + m * L.u < L.k; ++m)
            for (int ci = 0; ci < L.cin; ++ci)
                for (int co = 0; co < L.cout; ++co)
                    p[(((size_t)r * L.taps_max + m) * L.cin_pad + ci) * L.cout_pad + co] =
                        w.data[((size_t)ci * L.cout + co) * L.k + r + m * L.u];
    L.w_fp32 = h->upload(p);
    L.bias = h->upload(b.data);
}

static void commit(hfg_handle* h) {
    const hfg_config& c = h->cfg;
    h->free_device_weights();
    h->ups.clear();
    h->mrfs.clear();
    h->tf32_split = -1;

    // conv_pre: Conv1d(n_mels, c0, 7, padding=3)   (reference models/hifigan.py:177-183)
    h->pre = ConvLayer{};
    h->pre.cin = c.n_mels; h->pre.cout = c.upsample_initial_channel;
    h->pre.k = 7; h->pre.dil = 1; h->pre.pad = 3;
    {
        HostTensor w = plain_weight(h, "conv_pre", {h->pre.cout, h->pre.cin, 7});
        HostTensor b = plain_bias(h, "conv_pre", h->pre.cout);
        pack_conv_fp32(h, h->pre, w, b);
        tc_pack_conv(h, h->pre, w, b);
    }
    int C = c.upsample_initial_channel;
    for (int i = 0; i < c.num_upsamples; ++i) {
        UpLayer U{};
        U.cin = C; U.cout = C / 2;
        U.k = c.upsample_kernel_sizes[i]; U.u = c.upsample_rates[i];
        U.p = (U.k - U.u) / 2;                       // reference :201
        const std::string up = "ups." + std::to_string(i);
        HostTensor w = plain_weight(h, up, {U.cin, U.cout, U.k});
        HostTensor b = plain_bias(h, up, U.cout);
        pack_up_fp32(h, U, w, b);
        tc_pack_up(h, U, w, b);
        h->ups.push_back(U);
        C = U.cout;
        std::vector<std::vector<PairLayers>> mrf;
        for (int j = 0; j < c.num_resblocks; ++j) {
            std::vector<PairLayers> rb;
            const int k = c.resblock_kernel_sizes[j];
            for (int l = 0; l < c.num_dilations[j]; ++l) {
                const int d = c.resblock_dilations[j][l];
                PairLayers P{};
                const std::string base =
                    "mrfs." + std::to_string(i) + ".resblocks." + std::to_string(j) + ".";
                // convs1[l]: dilation d, padding (k*d-d)/2   (reference :52-59, 21-23)
                P.c1.cin = P.c1.cout = C; P.c1.k = k; P.c1.dil = d; P.c1.pad = (k * d - d) / 2;
                // convs2[l]: dilation 1, padding (k-1)/2     (reference :61-69)
                P.c2.cin = P.c2.cout = C; P.c2.k = k; P.c2.dil = 1; P.c2.pad = (k - 1) / 2;
                const std::string n1 = base + "convs1." + std::to_string(l);
                const std::string n2 = base + "convs2." + std::to_string(l);
                HostTensor w1 = plain_weight(h, n1, {C, C, k}), b1 = plain_bias(h, n1, C);
                HostTensor w2 = plain_weight(h, n2, {C, C, k}), b2 = plain_bias(h, n2, C);
                pack_conv_fp32(h, P.c1, w1, b1);
                pack_conv_fp32(h, P.c2, w2, b2);
                tc_pack_conv(h, P.c1, w1, b1);
                tc_pack_conv(h, P.c2, w2, b2);
                rb.push_back(P);
            }
            mrf.push_back(rb);
        }
        h->mrfs.push_back(mrf);
    }
    // conv_post: Conv1d(C_last, 1, 7, padding=3)   (reference :215-222)
    {
        HostTensor w = plain_weight(h, "conv_post", {1, C, 7});
        HostTensor b = plain_bias(h, "conv_post", 1);
        h->post_cin = C;
        h->post_w = h->upload(w.data);     // [1][Cin][7] == [Cin][7]
        h->post_b = h->upload(b.data);
        tc_pack_post(h, w, b);
    }
    h->committed = true;
}

// ---------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------

static int64_t up_len(const UpLayer& U, int64_t t) { return (t - 1) * U.u - 2 * U.p + U.k; }

static void stage_geometry(const hfg_handle* h, int T, std::vector<int>& C, std::vector<int64_t>& L) {
    C.assign(1, h->cfg.upsample_initial_channel);
    L.assign(1, T);
    for (auto& U : h->ups) {
        C.push_back(U.cout);
        L.push_back(up_len(U, L.back()));
    }
}

// ---------------------------------------------------------------------------
// fp32 launch plan
// ---------------------------------------------------------------------------

static void launch_conv_fp32(hfg_handle* h, cudaStream_t st, const ConvArgs& a, int rco, int nq, int B,
                             const char* label) {
    const int tco = rco * 16;
    // algorithmic work: 2*Cin*Cout*k per output step (Conv1d) / per input step (ConvTranspose1d)
    const double flops = 2.0 * a.Cin * a.Cout * a.k * (double)B * (a.phases > 1 ? a.Tin : a.Tout);
    const double bytes = 4.0 * B * ((double)a.Cin * a.Tin + (double)a.Cout * a.Tout * (a.res ? 2 : 1)) +
                         4.0 * a.Cin * a.Cout * a.k;
    const int span = (a.taps_max - 1) * (a.dil < 0 ? -a.dil : a.dil);
    const size_t smem = sizeof(float) * ((size_t)kTileCi * (kTileT + span) + (size_t)a.taps_max * kTileCi * tco);
    dim3 grid((nq + kTileT - 1) / kTileT, a.phases * a.co_tiles, B);
    h->prof_begin(st, label, flops, bytes);
    switch (rco) {
        case 8: conv_tile_fp32<8><<<grid, kThreads, smem, st>>>(a); break;
        case 4: conv_tile_fp32<4><<<grid, kThreads, smem, st>>>(a); break;
        default: conv_tile_fp32<2><<<grid, kThreads, smem, st>>>(a); break;
    }
    h->prof_end(st);
    check_cuda(cudaGetLastError(), "conv_tile_fp32 launch");
}

static ConvArgs conv_args(const ConvLayer& L, const float* x, float* y, int T, float slope) {
    ConvArgs a{};
    a.x = x; a.w = L.w_fp32; a.bias = L.bias; a.res = nullptr; a.y = y;
    a.Cin = L.cin; a.CinPad = L.cin_pad; a.Tin = T;
    a.Cout = L.cout; a.CoutPad = L.cout_pad; a.Tout = T;
    a.taps_max = L.k; a.k = L.k; a.u = 1; a.dil = L.dil; a.pad = L.pad;
    a.out_stride = 1; a.out_off = 0; a.phases = 1;
    a.co_tiles = L.cout_pad / (L.rco * 16);
    a.slope = slope; a.epi_flags = 0; a.div = 1.f;
    return a;
}

static size_t workspace_fp32(const hfg_handle* h, int B, int T);
static size_t fp32_buffer_stride(const hfg_handle* h, int B, int T);

static void forward_fp32(hfg_handle* h, const float* mel, int B, int T, float* wav, char* ws,
                         cudaStream_t st, float* const* stage_out, const int* lengths, int halo) {
    std::vector<int> C;
    std::vector<int64_t> L;
    stage_geometry(h, T, C, L);
    const size_t stride = fp32_buffer_stride(h, B, T);
    float* bufX = (float*)(ws);
    float* bufR = (float*)(ws + stride);
    float* bufH = (float*)(ws + 2 * stride);
    float* bufA = (float*)(ws + 3 * stride);
    float* bufB = (float*)(ws + 4 * stride);
    const float slope = 0.1f;
    const int n_rb = h->cfg.num_resblocks;
    auto dump = [&](int idx, const float* src, size_t n) {
        if (stage_out && stage_out[idx])
            check_cuda(cudaMemcpyAsync(stage_out[idx], src, n * sizeof(float),
                                       cudaMemcpyDeviceToDevice, st), "stage dump");
    };

    // conv_pre (reference :238), no activation on the mel
    if (h->mel_layout == HFG_MEL_FRAMES_LAST) {           // acoustic-model layout: transpose into bufX first
        dim3 tgrid((T + 31) / 32, (h->cfg.n_mels + 31) / 32, B);
        h->prof_begin(st, "transpose_mel", 0, 8.0 * B * h->cfg.n_mels * T);
        transpose_btc_to_bct<<<tgrid, dim3(32, 8), 0, st>>>(mel, bufX, h->cfg.n_mels, T);
        h->prof_end(st);
        check_cuda(cudaGetLastError(), "transpose launch");
        mel = bufX;
    }
    float* cur = bufA;
    launch_conv_fp32(h, st, conv_args(h->pre, mel, cur, T, 1.0f), h->pre.rco, T, B, "conv_pre");
    dump(0, cur, (size_t)B * C[0] * L[0]);

    for (size_t i = 0; i < h->ups.size(); ++i) {
        const UpLayer& U = h->ups[i];
        const int Tin = (int)L[i], Tout = (int)L[i + 1];
        // x = ups[i](leaky_relu(x))  (reference :244-245)
        ConvArgs a{};
        a.x = cur; a.w = U.w_fp32; a.bias = U.bias; a.res = nullptr; a.y = bufX;
        a.Cin = U.cin; a.CinPad = U.cin_pad; a.Tin = Tin;
        a.Cout = U.cout; a.CoutPad = U.cout_pad; a.Tout = Tout;
        a.taps_max = U.taps_max; a.k = U.k; a.u = U.u; a.dil = -1; a.pad = 0;
        a.out_stride = U.u; a.out_off = -U.p; a.phases = U.u;
        a.co_tiles = U.cout_pad / (U.rco * 16);
        a.slope = slope; a.epi_flags = 0; a.div = 1.f;
        const int nq = (Tout - 1 + U.p) / U.u + 1;
        launch_conv_fp32(h, st, a, U.rco, nq, B, ("ups" + std::to_string(i)).c_str());
        const size_t n = (size_t)B * U.cout * Tout;
        dump(1 + 2 * (int)i, bufX, n);

        // MRF (reference :116-131)
        float* acc = (cur == bufA) ? bufB : bufA;
        for (int j = 0; j < n_rb; ++j) {
            const auto& rb = h->mrfs[i][j];
            const float* r = bufX;
            for (size_t l = 0; l < rb.size(); ++l) {
                const bool last = (l + 1 == rb.size());
                const std::string lab = "mrf" + std::to_string(i) + ".k" + std::to_string(rb[l].c1.k);
                launch_conv_fp32(h, st, conv_args(rb[l].c1, r, bufH, Tout, slope), rb[l].c1.rco, Tout, B, lab.c_str());
                ConvArgs a2 = conv_args(rb[l].c2, bufH, last ? acc : bufR, Tout, slope);
                a2.res = r;
                if (last) {
                    if (j > 0) a2.epi_flags |= EPI_ACC_READ;
                    if (j == n_rb - 1) { a2.epi_flags |= EPI_ACC_DIV; a2.div = (float)n_rb; }
                }
                launch_conv_fp32(h, st, a2, rb[l].c2.rco, Tout, B, lab.c_str());
                r = bufR;
            }
        }
        cur = acc;
        dump(2 + 2 * (int)i, cur, n);
    }
    // wav = tanh(conv_post(leaky_relu(x)))  (reference :254-256)
    PostArgs p{};
    p.x = cur; p.w = h->post_w; p.bias = h->post_b; p.y = wav;
    p.Cin = h->post_cin; p.T = (int)L.back(); p.k = 7; p.pad = 3; p.slope = slope;
    dim3 grid((p.T + 255) / 256, B);
    h->prof_begin(st, "conv_post", 2.0 * p.Cin * p.k * (double)B * p.T,
                  4.0 * B * ((double)p.Cin * p.T + p.T));
    conv_post_tanh_fp32<<<grid, 256, sizeof(float) * p.Cin * p.k, st>>>(p);
    h->prof_end(st);
    check_cuda(cudaGetLastError(), "conv_post launch");
    if (lengths) {
        // variable-length batch in the strict mode: everything is generated, samples beyond each utterance's
        // valid length are zeroed (the tensor-core modes skip those tiles instead)
        int* tab = reinterpret_cast<int*>(ws + 5 * stride);
        LenGeom g{};
        g.n_stages = (int)h->ups.size();
        for (int i = 0; i < g.n_stages; ++i) { g.u[i] = h->ups[i].u; g.k[i] = h->ups[i].k; g.p[i] = h->ups[i].p; }
        h->prof_begin(st, "len_table", 0, 0);
        tc_len_table<<<(B + 127) / 128, 128, 0, st>>>(lengths, B, T, halo, g, tab);
        h->prof_end(st);
        h->prof_begin(st, "mask_tail", 0, 0);
        tc_mask_tail<<<dim3((p.T + 255) / 256, B), 256, 0, st>>>(wav, tab + (size_t)(1 + g.n_stages) * B, p.T);
        h->prof_end(st);
        check_cuda(cudaGetLastError(), "mask_tail launch");
    }
}

// elements of the largest activation; the frames-last path also stages the transposed mel in a buffer
static size_t fp32_buffer_stride(const hfg_handle* h, int B, int T) {
    std::vector<int> C;
    std::vector<int64_t> L;
    stage_geometry(h, T, C, L);
    size_t emax = (size_t)B * h->cfg.n_mels * T;
    for (size_t i = 0; i < C.size(); ++i) emax = std::max(emax, (size_t)B * C[i] * L[i]);
    return (emax * sizeof(float) + 255) / 256 * 256;
}
static size_t fp32_lens_bytes(const hfg_handle* h, int B) {
    return ((size_t)(h->ups.size() + 2) * B * sizeof(int) + 255) / 256 * 256;
}
static size_t workspace_fp32(const hfg_handle* h, int B, int T) {
    return 5 * fp32_buffer_stride(h, B, T) + fp32_lens_bytes(h, B);
}

static void validate_config(const hfg_config& c) {
    auto bad = [](const std::string& m) { throw StatusError(HFG_ERR_INVALID, m); };
    if (c.n_mels <= 0) bad("n_mels must be positive");
    if (c.num_upsamples <= 0 || c.num_upsamples > HFG_MAX_STAGES) bad("num_upsamples out of range");
    if (c.num_resblocks <= 0 || c.num_resblocks > HFG_MAX_STAGES) bad("num_resblocks out of range");
    if (c.upsample_initial_channel <= 0 ||
        (c.upsample_initial_channel >> c.num_upsamples) <= 0 ||
        ((c.upsample_initial_channel >> c.num_upsamples) << c.num_upsamples) != c.upsample_initial_channel)
        bad("upsample_initial_channel must be divisible by 2^num_upsamples");
    for (int i = 0; i < c.num_upsamples; ++i) {
        if (c.upsample_rates[i] <= 0 || c.upsample_kernel_sizes[i] < c.upsample_rates[i])
            bad("upsample kernel size must be >= rate > 0");
    }
    for (int j = 0; j < c.num_resblocks; ++j) {
        if (c.resblock_kernel_sizes[j] <= 0 || c.resblock_kernel_sizes[j] % 2 == 0)
            bad("resblock kernel sizes must be odd (reference 'same' padding, models/hifigan.py:21-23)");
        if (c.num_dilations[j] <= 0 || c.num_dilations[j] > HFG_MAX_STAGES) bad("num_dilations out of range");
        for (int l = 0; l < c.num_dilations[j]; ++l)
            if (c.resblock_dilations[j][l] <= 0) bad("dilations must be positive");
    }
}

// Receptive radius of one output frame in mel frames: the interval of wav samples [f*hop, (f+1)*hop) is
// propagated backwards through conv_post, every MRF (widest resblock), every ConvTranspose1d and conv_pre
// (reference models/hifigan.py:224-261).  13 for the default configuration.
static int receptive_radius(const hfg_config& c) {
    long long hop = 1;
    for (int i = 0; i < c.num_upsamples; ++i) hop *= c.upsample_rates[i];
    const long long f = 1 << 20;                            // far from both edges
    long long lo = f * hop, hi = (f + 1) * hop - 1;
    lo -= 3; hi += 3;                                       // conv_post, k = 7
    auto floor_div = [](long long a, long long b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
    for (int i = c.num_upsamples - 1; i >= 0; --i) {
        long long r = 0;
        for (int j = 0; j < c.num_resblocks; ++j) {
            long long rj = 0;
            const int k = c.resblock_kernel_sizes[j];
            for (int l = 0; l < c.num_dilations[j]; ++l) rj += (long long)c.resblock_dilations[j][l] * (k - 1) / 2 + (k - 1) / 2;
            r = std::max(r, rj);
        }
        lo -= r; hi += r;
        const int u = c.upsample_rates[i], k = c.upsample_kernel_sizes[i], p = (k - u) / 2;
        // output t reads input q iff 0 <= t + p - q*u < k
        lo = -floor_div(-(lo + p - (k - 1)), u);            // ceil((lo + p - k + 1) / u)
        hi = floor_div(hi + p, u);
    }
    lo -= 3; hi += 3;                                       // conv_pre, k = 7
    return (int)std::max(f - lo, hi - f);
}

static void do_forward(hfg_handle* h, const float* mel, int B, int T, float* wav, void* ws,
                       size_t ws_bytes, int mode, cudaStream_t st, float* const* stage_out,
                       const int* lengths = nullptr, int halo = 0) {
    if (!h->committed) throw StatusError(HFG_ERR_STATE, "weights not committed (call hfg_commit_weights)");
    if (!mel || !wav) throw StatusError(HFG_ERR_INVALID, "null mel/wav pointer");
    if (B <= 0 || T <= 0) throw StatusError(HFG_ERR_INVALID, "batch and frames must be positive");
    if (B > 65535) throw StatusError(HFG_ERR_INVALID, "batch > 65535: split the call");
    size_t need_bytes = 0;
    if (mode == HFG_MODE_FP32) need_bytes = workspace_fp32(h, B, T);
    else if (tc_is_tc_mode(mode)) need_bytes = tc_workspace_bytes(h, B, T, mode);
    else throw StatusError(HFG_ERR_INVALID, "unknown mode");
    if (!ws || ws_bytes < need_bytes) throw StatusError(HFG_ERR_WORKSPACE, "workspace too small");
    if (((uintptr_t)ws & 255) != 0) throw StatusError(HFG_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    h->launches = 0;
    h->prof.clear();
    h->events_used = 0;
    if (lengths && halo < receptive_radius(h->cfg))
        throw StatusError(HFG_ERR_INVALID, "halo_frames is smaller than the receptive radius of this configuration (" +
                                               std::to_string(receptive_radius(h->cfg)) + " frames): the valid region would change");
    if (mode == HFG_MODE_FP32) forward_fp32(h, mel, B, T, wav, (char*)ws, st, stage_out, lengths, halo);
    else tc_forward(h, mel, B, T, wav, (char*)ws, mode, st, stage_out, lengths, halo);
}

}  // namespace hfg

using namespace hfg;

// Forward on the handle's own stream and buffers, replayed from a CUDA graph where possible.  First call with a
// given geometry: plain launches (lazy initialisation happens here).  Second call: the same sequence is captured
// into a graph.  From then on: one cudaGraphLaunch per forward.
static void run_graphed(hfg_handle* h, hfg_handle::HostGraph& G, float* dev_mel, float* dev_wav, int batch, int frames, int mode) {
    static const bool graphs_on = []() { const char* e = getenv("HFG_HOST_GRAPH"); return !e || atoi(e) != 0; }();
    const bool same = G.B == batch && G.T == frames && G.mode == mode && G.layout == h->mel_layout &&
                      G.mel == dev_mel && G.wav == dev_wav && G.ws == h->dev_ws;
    if (!same) {
        const bool failed = G.failed;
        if (G.exec) cudaGraphExecDestroy(G.exec);
        G = hfg_handle::HostGraph{};
        G.failed = failed;
        G.B = batch; G.T = frames; G.mode = mode; G.layout = h->mel_layout;
        G.mel = dev_mel; G.wav = dev_wav; G.ws = h->dev_ws;
    }
    const bool use_graph = graphs_on && !G.failed && h->profiling == 0;
    if (use_graph && G.exec) {
        check_cuda(cudaGraphLaunch(G.exec, h->stream), "cudaGraphLaunch");
        h->launches = G.launches;
    } else if (use_graph && G.calls >= 1) {
        cudaGraph_t graph = nullptr;
        check_cuda(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed), "cudaStreamBeginCapture");
        bool ok = true;
        std::string why;
        try {
            do_forward(h, dev_mel, batch, frames, dev_wav, h->dev_ws, h->dev_ws_bytes, mode, h->stream, nullptr);
        } catch (const std::exception& e) { ok = false; why = e.what(); }
        const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        if (ok && ce == cudaSuccess && graph &&
            cudaGraphInstantiate(&G.exec, graph, nullptr, nullptr, 0) == cudaSuccess) {
            G.launches = h->launches;
            check_cuda(cudaGraphLaunch(G.exec, h->stream), "cudaGraphLaunch");
        } else {
            // capture not possible here: remember that and run the plain sequence
            cudaGetLastError();
            G.exec = nullptr;
            G.failed = true;
            do_forward(h, dev_mel, batch, frames, dev_wav, h->dev_ws, h->dev_ws_bytes, mode, h->stream, nullptr);
        }
        if (graph) cudaGraphDestroy(graph);
    } else {
        do_forward(h, dev_mel, batch, frames, dev_wav, h->dev_ws, h->dev_ws_bytes, mode, h->stream, nullptr);
    }
    G.calls++;
}

#define HFG_TRY(h)  try {
#define HFG_CATCH(h)                                                         \
    } catch (const StatusError& e) {                                         \
        if (h) (h)->last_error = e.what();                                   \
        return e.code;                                                       \
    } catch (const std::exception& e) {                                      \
        if (h) (h)->last_error = e.what();                                   \
        return HFG_ERR_INVALID;                                              \
    }                                                                        \
    return HFG_OK;

extern "C" {

int hfg_abi_version(void) { return HFG_ABI_VERSION; }

int hfg_create(const hfg_config* cfg, hfg_handle** out) {
    if (!cfg || !out) return HFG_ERR_INVALID;
    *out = nullptr;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return HFG_ERR_CUDA; }
    hfg_handle* h = nullptr;
    try {
        validate_config(*cfg);
        h = new hfg_handle();
        h->cfg = *cfg;
        h->device = dev;
        cudaDeviceProp prop{};
        check_cuda(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
        h->sm_count = prop.multiProcessorCount;
        h->cc_major = prop.major;
        configure_kernels(h);
    } catch (const StatusError& e) {
        delete h;
        return e.code;
    } catch (...) {
        delete h;
        return HFG_ERR_INVALID;
    }
    *out = h;
    return HFG_OK;
}

void hfg_destroy(hfg_handle* h) {
    if (!h) return;
    h->free_device_weights();
    h->free_host_path();
    h->free_streams();
    delete h;
}

const char* hfg_last_error(const hfg_handle* h) { return h ? h->last_error.c_str() : "null handle"; }

int hfg_set_weight(hfg_handle* h, const char* name, const float* data, const int64_t* shape,
                   int32_t ndim) {
    if (!h) return HFG_ERR_INVALID;
    HFG_TRY(h)
    if (!name || !data || !shape || ndim < 1 || ndim > 3)
        throw StatusError(HFG_ERR_INVALID, "hfg_set_weight: bad argument");
    HostTensor t;
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] <= 0) throw StatusError(HFG_ERR_INVALID, "hfg_set_weight: non-positive dim");
        t.shape.push_back(shape[i]);
        n *= shape[i];
    }
    t.data.assign(data, data + n);
    const std::string key(name);
    // plain and weight-normed forms of the same layer are mutually exclusive
    auto ends = [&](const char* s) {
        const size_t l = strlen(s);
        return key.size() >= l && key.compare(key.size() - l, l, s) == 0;
    };
    if (ends(".weight")) {
        h->sd.erase(key + "_g");
        h->sd.erase(key + "_v");
    } else if (ends(".weight_g") || ends(".weight_v")) {
        h->sd.erase(key.substr(0, key.size() - 2));
    } else if (!ends(".bias")) {
        throw StatusError(HFG_ERR_INVALID, "unexpected state_dict key: " + key);
    }
    h->sd[key] = std::move(t);
    h->committed = false;
    HFG_CATCH(h)
}

int hfg_commit_weights(hfg_handle* h) {
    if (!h) return HFG_ERR_INVALID;
    HFG_TRY(h)
    check_cuda(cudaSetDevice(h->device), "cudaSetDevice");
    commit(h);
    HFG_CATCH(h)
}

int hfg_out_len(const hfg_handle* h, int32_t frames, int64_t* out_len) {
    if (!h || !out_len || frames <= 0) return HFG_ERR_INVALID;
    int64_t t = frames;
    const hfg_config& c = h->cfg;
    for (int i = 0; i < c.num_upsamples; ++i) {
        const int u = c.upsample_rates[i], k = c.upsample_kernel_sizes[i];
        t = (t - 1) * u - 2 * ((k - u) / 2) + k;
    }
    *out_len = t;
    return HFG_OK;
}

int hfg_workspace_bytes(const hfg_handle* hc, int32_t batch, int32_t frames, int32_t mode,
                        size_t* bytes) {
    hfg_handle* h = const_cast<hfg_handle*>(hc);
    if (!h || !bytes) return HFG_ERR_INVALID;
    HFG_TRY(h)
    if (!h->committed) throw StatusError(HFG_ERR_STATE, "weights not committed");
    if (batch <= 0 || frames <= 0) throw StatusError(HFG_ERR_INVALID, "batch and frames must be positive");
    if (mode == HFG_MODE_FP32) *bytes = workspace_fp32(h, batch, frames);
    else if (tc_is_tc_mode(mode)) *bytes = tc_workspace_bytes(h, batch, frames, mode);
    else throw StatusError(HFG_ERR_INVALID, "unknown mode");
    HFG_CATCH(h)
}

int hfg_forward(hfg_handle* h, const float* mel_dev, int32_t batch, int32_t frames, float* wav_dev,
                void* workspace_dev, size_t workspace_bytes, int32_t mode, void* stream) {
    return hfg_forward_stages(h, mel_dev, batch, frames, wav_dev, workspace_dev, workspace_bytes,
                              mode, stream, nullptr);
}

int hfg_forward_stages(hfg_handle* h, const float* mel_dev, int32_t batch, int32_t frames,
                       float* wav_dev, void* workspace_dev, size_t workspace_bytes, int32_t mode,
                       void* stream, float* const* stage_out_dev) {
    if (!h) return HFG_ERR_INVALID;
    HFG_TRY(h)
    do_forward(h, mel_dev, batch, frames, wav_dev, workspace_dev, workspace_bytes, mode,
               (cudaStream_t)stream, stage_out_dev);
    HFG_CATCH(h)
}

int hfg_forward_lengths(hfg_handle* h, const float* mel_dev, const int32_t* lengths_dev, int32_t halo_frames,
                        int32_t batch, int32_t frames, float* wav_dev, void* workspace_dev, size_t workspace_bytes,
                        int32_t mode, void* stream) {
    if (!h) return HFG_ERR_INVALID;
    HFG_TRY(h)
    if (!lengths_dev) throw StatusError(HFG_ERR_INVALID, "hfg_forward_lengths: null lengths pointer");
    do_forward(h, mel_dev, batch, frames, wav_dev, workspace_dev, workspace_bytes, mode, (cudaStream_t)stream,
               nullptr, lengths_dev, halo_frames);
    HFG_CATCH(h)
}

int hfg_receptive_radius(const hfg_config* cfg, int32_t* frames) {
    if (!cfg || !frames) return HFG_ERR_INVALID;
    try {
        validate_config(*cfg);
    } catch (...) {
        return HFG_ERR_INVALID;
    }
    *frames = receptive_radius(*cfg);
    return HFG_OK;
}

int hfg_forward_host(hfg_handle* h, const float* mel_host, int32_t batch, int32_t frames,
                     float* wav_host, int32_t mode) {
    return hfg_forward_host_ex(h, mel_host, batch, frames, wav_host, mode, 0);
}

int hfg_forward_host_ex(hfg_handle* h, const float* mel_host, int32_t batch, int32_t frames,
                        float* wav_host, int32_t mode, uint32_t flags) {
    if (!h) return HFG_ERR_INVALID;
    HFG_TRY(h)
    if (!h->committed) throw StatusError(HFG_ERR_STATE, "weights not committed");
    if (!mel_host || !wav_host || batch <= 0 || frames <= 0)
        throw StatusError(HFG_ERR_INVALID, "hfg_forward_host: bad argument");
    check_cuda(cudaSetDevice(h->device), "cudaSetDevice");
    int64_t tout = 0;
    hfg_out_len(h, frames, &tout);
    const size_t mel_bytes = sizeof(float) * (size_t)batch * h->cfg.n_mels * frames;
    const size_t wav_bytes = sizeof(float) * (size_t)batch * tout;
    size_t ws_bytes = 0;
    if (mode == HFG_MODE_FP32) ws_bytes = workspace_fp32(h, batch, frames);
    else if (tc_is_tc_mode(mode)) ws_bytes = tc_workspace_bytes(h, batch, frames, mode);
    else throw StatusError(HFG_ERR_INVALID, "unknown mode");
    const bool mel_pinned = (flags & HFG_HOST_MEL_PINNED) != 0, wav_pinned = (flags & HFG_HOST_WAV_PINNED) != 0;
    h->ensure_host_path(mel_pinned ? 0 : mel_bytes, wav_pinned ? 0 : wav_bytes, mel_bytes, wav_bytes, ws_bytes);
    const float* src = mel_host;
    if (!mel_pinned) { memcpy(h->pin_mel, mel_host, mel_bytes); src = h->pin_mel; }
    check_cuda(cudaMemcpyAsync(h->dev_mel, src, mel_bytes, cudaMemcpyHostToDevice, h->stream), "H2D mel");
    run_graphed(h, h->host_graph, h->dev_mel, h->dev_wav, batch, frames, mode);
    float* dst = wav_pinned ? wav_host : h->pin_wav;
    check_cuda(cudaMemcpyAsync(dst, h->dev_wav, wav_bytes, cudaMemcpyDeviceToHost, h->stream), "D2H wav");
    check_cuda(cudaStreamSynchronize(h->stream), "stream sync");
    if (!wav_pinned) memcpy(wav_host, h->pin_wav, wav_bytes);
    HFG_CATCH(h)
}

int hfg_forward_host_submit(hfg_handle* h, int32_t slot, const float* mel_host, int32_t batch, int32_t frames,
                            float* wav_host, int32_t mode) {
    if (!h) return HFG_ERR_INVALID;
    HFG_TRY(h)
    if (!h->committed) throw StatusError(HFG_ERR_STATE, "weights not committed");
    if (slot < 0 || slot >= hfg_handle::kHostSlots || !mel_host || !wav_host || batch <= 0 || frames <= 0)
        throw StatusError(HFG_ERR_INVALID, "hfg_forward_host_submit: bad argument");
    auto& S = h->slots[slot];
    if (S.busy) throw StatusError(HFG_ERR_STATE, "slot still in flight: call hfg_forward_host_wait first");
    check_cuda(cudaSetDevice(h->device), "cudaSetDevice");
    int64_t tout = 0;
    hfg_out_len(h, frames, &tout);
    const size_t mel_bytes = sizeof(float) * (size_t)batch * h->cfg.n_mels * frames;
    const size_t wav_bytes = sizeof(float) * (size_t)batch * tout;
    size_t ws_bytes = 0;
    if (mode == HFG_MODE_FP32) ws_bytes = workspace_fp32(h, batch, frames);
    else if (tc_is_tc_mode(mode)) ws_bytes = tc_workspace_bytes(h, batch, frames, mode);
    else throw StatusError(HFG_ERR_INVALID, "unknown mode");
    h->ensure_stream_path(slot, mel_bytes, wav_bytes, ws_bytes);
    // H2D on the copy-in stream, forward on the compute stream once it has landed, D2H on the copy-out stream once
    // the forward is done: the copies of neighbouring submissions overlap this one's kernels
    check_cuda(cudaMemcpyAsync(S.dev_mel, mel_host, mel_bytes, cudaMemcpyHostToDevice, h->copy_in), "H2D mel");
    check_cuda(cudaEventRecord(S.ev_h2d, h->copy_in), "cudaEventRecord(h2d)");
    check_cuda(cudaStreamWaitEvent(h->stream, S.ev_h2d, 0), "cudaStreamWaitEvent(h2d)");
    run_graphed(h, S.graph, S.dev_mel, S.dev_wav, batch, frames, mode);
    check_cuda(cudaEventRecord(S.ev_done, h->stream), "cudaEventRecord(done)");
    check_cuda(cudaStreamWaitEvent(h->copy_out, S.ev_done, 0), "cudaStreamWaitEvent(done)");
    check_cuda(cudaMemcpyAsync(wav_host, S.dev_wav, wav_bytes, cudaMemcpyDeviceToHost, h->copy_out), "D2H wav");
    check_cuda(cudaEventRecord(S.ev_d2h, h->copy_out), "cudaEventRecord(d2h)");
    S.busy = true;
    HFG_CATCH(h)
}

int hfg_forward_host_wait(hfg_handle* h, int32_t slot) {
    if (!h) return HFG_ERR_INVALID;
    HFG_TRY(h)
    if (slot < 0 || slot >= hfg_handle::kHostSlots) throw StatusError(HFG_ERR_INVALID, "hfg_forward_host_wait: bad slot");
    auto& S = h->slots[slot];
    if (!S.busy) throw StatusError(HFG_ERR_STATE, "nothing in flight on this slot");
    S.busy = false;
    check_cuda(cudaEventSynchronize(S.ev_d2h), "cudaEventSynchronize(d2h)");
    HFG_CATCH(h)
}

int hfg_set_mel_layout(hfg_handle* h, int32_t layout) {
    if (!h || (layout != HFG_MEL_CHANNELS_FIRST && layout != HFG_MEL_FRAMES_LAST)) return HFG_ERR_INVALID;
    h->mel_layout = layout;
    return HFG_OK;
}

int hfg_set_profiling(hfg_handle* h, int32_t enable) {
    if (!h) return HFG_ERR_INVALID;
    h->profiling = enable == 2 ? 2 : (enable != 0 ? 1 : 0);
    return HFG_OK;
}

int hfg_get_profile(hfg_handle* h, char* buf, size_t buf_bytes, size_t* needed) {
    if (!h || !needed) return HFG_ERR_INVALID;
    HFG_TRY(h)
    // aggregate by label, in first-seen order
    std::vector<std::string> order;
    std::map<std::string, std::array<double, 4>> agg;   // launches, ms, flops, bytes
    for (auto& r : h->prof) {
        check_cuda(cudaEventSynchronize(r.e1), "cudaEventSynchronize");
        float ms = 0.f;
        check_cuda(cudaEventElapsedTime(&ms, r.e0, r.e1), "cudaEventElapsedTime");
        if (!agg.count(r.label)) { order.push_back(r.label); agg[r.label] = {0, 0, 0, 0}; }
        auto& a = agg[r.label];
        a[0] += 1; a[1] += ms; a[2] += r.flops; a[3] += r.bytes;
    }
    std::string js = "[";
    for (size_t i = 0; i < order.size(); ++i) {
        const auto& a = agg[order[i]];
        char tmp[256];
        snprintf(tmp, sizeof(tmp), "%s{\"kernel\":\"%s\",\"launches\":%d,\"ms\":%.6f,\"flops\":%.6e,\"bytes\":%.6e}",
                 i ? "," : "", order[i].c_str(), (int)a[0], a[1], a[2], a[3]);
        js += tmp;
    }
    js += "]";
    *needed = js.size() + 1;
    if (buf && buf_bytes >= js.size() + 1) memcpy(buf, js.c_str(), js.size() + 1);
    HFG_CATCH(h)
}

int hfg_bench_layer(hfg_handle* h, int32_t stage, int32_t resblock, int32_t pair, int32_t which,
                    int32_t batch, int32_t rows, int32_t mode, int32_t iters, float* ms) {
    if (!h || !ms) return HFG_ERR_INVALID;
    HFG_TRY(h)
    if (!h->committed) throw StatusError(HFG_ERR_STATE, "weights not committed");
    if (!tc_supported(h)) throw StatusError(HFG_ERR_UNSUPPORTED, "tensor-core path unsupported for this config");
    if (stage < 0 || stage >= (int)h->mrfs.size() || resblock < 0 || resblock >= (int)h->mrfs[stage].size() ||
        pair < 0 || pair >= (int)h->mrfs[stage][resblock].size() || which < 0 || which > 2 || batch <= 0 ||
        rows <= 0 || iters <= 0)
        throw StatusError(HFG_ERR_INVALID, "hfg_bench_layer: bad argument");
    if (mode == HFG_MODE_BF16) *ms = tc_bench_layer_impl<PREC_BF16>(h, stage, resblock, pair, which, batch, rows, iters);
    else if (mode == HFG_MODE_FP16) *ms = tc_bench_layer_impl<PREC_FP16>(h, stage, resblock, pair, which, batch, rows, iters);
    else if (mode == HFG_MODE_TF32 && tc_tf32_mixed(h) && which == 2)
        *ms = tc_bench_layer_impl<PREC_FP16, true>(h, stage, resblock, pair, which, batch, rows, iters);
    else if (mode == HFG_MODE_TF32) *ms = tc_bench_layer_impl<PREC_TF32>(h, stage, resblock, pair, which, batch, rows, iters);
    else throw StatusError(HFG_ERR_INVALID, "hfg_bench_layer: tensor-core modes only");
    HFG_CATCH(h)
}

static int lr_fail(cudaError_t e) { (void)e; cudaGetLastError(); return HFG_ERR_CUDA; }

int hfg_durations_from_log(const float* log_dur, int64_t n, int64_t* dur, void* stream) {
    if (!log_dur || !dur || n <= 0) return HFG_ERR_INVALID;
    lr_durations_from_log<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        log_dur, (long long)n, reinterpret_cast<long long*>(dur));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HFG_OK : lr_fail(e);
}

// scratch for the prefix sums: [B*Tph + B] ints, stream-ordered allocation
static int lr_scan(const int64_t* dur, int B, int Tph, cudaStream_t st, int** cum, int** totals) {
    int* buf = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&buf, sizeof(int) * ((size_t)B * Tph + B), st);
    if (e != cudaSuccess) return lr_fail(e);
    *cum = buf;
    *totals = buf + (size_t)B * Tph;
    lr_prefix_sum<<<B, 256, 0, st>>>(reinterpret_cast<const long long*>(dur), Tph, *cum, *totals);
    e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFreeAsync(buf, st); return lr_fail(e); }
    return HFG_OK;
}

int hfg_length_regulate_frames(const int64_t* dur, int32_t batch, int32_t n_phonemes, int64_t* max_frames,
                               void* stream) {
    if (!dur || !max_frames || batch <= 0 || n_phonemes <= 0) return HFG_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int *cum = nullptr, *totals = nullptr;
    int rc = lr_scan(dur, batch, n_phonemes, st, &cum, &totals);
    if (rc != HFG_OK) return rc;
    std::vector<int> host(batch);
    cudaError_t e = cudaMemcpyAsync(host.data(), totals, sizeof(int) * batch, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(cum, st);
    if (e != cudaSuccess) return lr_fail(e);
    int64_t m = 0;
    for (int v : host) m = std::max<int64_t>(m, v);
    *max_frames = m;
    return HFG_OK;
}

int hfg_length_regulate(const float* henc, const int64_t* dur, int32_t batch, int32_t n_phonemes,
                        int32_t d_model, int32_t frames, float* out, void* stream) {
    if (!henc || !dur || !out || batch <= 0 || n_phonemes <= 0 || d_model <= 0 || frames <= 0 || batch > 65535)
        return HFG_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int *cum = nullptr, *totals = nullptr;
    int rc = lr_scan(dur, batch, n_phonemes, st, &cum, &totals);
    if (rc != HFG_OK) return rc;
    const int tx = d_model >= 128 ? 64 : 32, ty = 256 / tx;
    dim3 grid((frames + ty - 1) / ty, batch);
    lr_expand<<<grid, dim3(tx, ty), 0, st>>>(henc, cum, n_phonemes, d_model, frames, out);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(cum, st);
    return e == cudaSuccess ? HFG_OK : lr_fail(e);
}

int hfg_tf32_plan(const hfg_handle* hc, int32_t* split) {
    hfg_handle* h = const_cast<hfg_handle*>(hc);
    if (!h || !split) return HFG_ERR_INVALID;
    HFG_TRY(h)
    if (!h->committed) throw StatusError(HFG_ERR_STATE, "weights not committed");
    *split = (tc_supported(h) && tc_tf32_mixed(h)) ? 1 : 0;
    HFG_CATCH(h)
}

int hfg_last_launch_count(const hfg_handle* h, int64_t* launches) {
    if (!h || !launches) return HFG_ERR_INVALID;
    *launches = h->launches;
    return HFG_OK;
}

}  // extern "C"
