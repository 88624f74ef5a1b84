// Host side of the tensor-core path: operand packing, workspace plan, launch plan.
#pragma once
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "fp32_kernels.cuh"
#include "model.h"
#include "tc_kernels.cuh"

namespace hfg {

// --------------------------------------------------------------------------
// operand packing
// --------------------------------------------------------------------------
static inline uint16_t f2bf(float f) {          // round-to-nearest-even
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float f2tf32(float f) {           // round-to-nearest to 10 mantissa bits
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) != 0x7F800000u) u += 0x0FFFu + ((u >> 13) & 1u);
    u &= 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

static inline int tc_pick_n(int cout) {
    for (int n : {256, 128, 64, 32, 16})
        if (cout % n == 0) return n;
    return 0;
}

// A layer is covered when its channels split into 16-byte cells and UMMA shapes:
//   C_in  % 16 == 0 (bf16: two 8-channel cells per K=16 step; tf32: two 4-channel cells per K=8)
//   C_out % 16 == 0 (UMMA N granularity at M = 128)
static inline bool tc_shape_ok(int cin, int cout) { return cin % 16 == 0 && cout % 16 == 0; }

// Generic packer.  get(n, ci, phase, tap) returns the weight multiplying input
// channel ci at tap `tap` of phase `phase` for output channel n (0 outside the kernel).
// Layout: [phase][ntile][kb][tap][chunk(8)][n(N)][cell(CW)], cells of 16 bytes.
template <typename F>
static void tc_pack_generic(hfg_handle* h, TcPack& tp, int cin, int cout, int phases, int taps_max, F get) {
    tp.ok = tc_shape_ok(cin, cout);
    if (!tp.ok) return;
    const int N = tc_pick_n(cout), ntiles = cout / N;
    for (int bf = 0; bf < 2; ++bf) {
        const int CW = bf ? 8 : 4;
        const int nchunks = cin / CW, nkb = (nchunks + 7) / 8;
        const size_t block = (size_t)8 * N * 16;                       // bytes per (kb, tap)
        const size_t total = (size_t)phases * ntiles * nkb * taps_max * block;
        std::vector<uint8_t> buf(total, 0);
        for (int ph = 0; ph < phases; ++ph)
            for (int nt = 0; nt < ntiles; ++nt)
                for (int kb = 0; kb < nkb; ++kb)
                    for (int tap = 0; tap < taps_max; ++tap) {
                        uint8_t* blk = buf.data() + ((((size_t)ph * ntiles + nt) * nkb + kb) * taps_max + tap) * block;
                        const int nck = std::min(8, nchunks - 8 * kb);
                        for (int c = 0; c < nck; ++c)
                            for (int n = 0; n < N; ++n) {
                                uint8_t* cell = blk + ((size_t)c * N + n) * 16;
                                for (int e = 0; e < CW; ++e) {
                                    const int ci = (8 * kb + c) * CW + e;
                                    const float v = get(nt * N + n, ci, ph, tap);
                                    if (bf) {
                                        const uint16_t q = f2bf(v);
                                        memcpy(cell + 2 * e, &q, 2);
                                    } else {
                                        const float q = f2tf32(v);
                                        memcpy(cell + 4 * e, &q, 4);
                                    }
                                }
                            }
                    }
        void* d = h->upload(buf);
        if (bf) tp.w_bf16 = d; else tp.w_tf32 = d;
    }
}

inline void tc_pack_conv(hfg_handle* h, ConvLayer& L, const HostTensor& w, const HostTensor&) {
    const int cin = L.cin, k = L.k;
    tc_pack_generic(h, L.tc, L.cin, L.cout, 1, L.k, [&](int n, int ci, int, int tap) {
        return w.data[((size_t)n * cin + ci) * k + tap];                // Conv1d weight [C_out, C_in, k]
    });
}
inline void tc_pack_up(hfg_handle* h, UpLayer& L, const HostTensor& w, const HostTensor&) {
    const int cout = L.cout, k = L.k, u = L.u;
    tc_pack_generic(h, L.tc, L.cin, L.cout, L.u, L.taps_max, [&](int n, int ci, int ph, int tap) {
        const int j = ph + tap * u;                                     // ConvTranspose1d weight [C_in, C_out, k]
        return j < k ? w.data[((size_t)ci * cout + n) * k + j] : 0.f;
    });
}
inline void tc_pack_post(hfg_handle*, const HostTensor&, const HostTensor&) {}   // conv_post reuses post_w

inline bool tc_supported(const hfg_handle* h) {
    if (!h->pre.tc.ok) return false;
    for (auto& U : h->ups) if (!U.tc.ok) return false;
    for (auto& st : h->mrfs) for (auto& rb : st) for (auto& P : rb) if (!P.c1.tc.ok || !P.c2.tc.ok) return false;
    // conv_post reads whole cells
    if (h->post_cin % 8 != 0) return false;
    // halo must fit the zero rows in front of every plane
    for (auto& st : h->mrfs) for (auto& rb : st) for (auto& P : rb)
        if (P.c1.pad > kPadL - 1 || P.c2.pad > kPadL - 1) return false;
    for (auto& U : h->ups) if (U.taps_max - 1 > kPadL - 1) return false;
    return true;
}

// --------------------------------------------------------------------------
// workspace plan
// --------------------------------------------------------------------------
struct TcPlane {            // geometry of one activation tensor in chunk-plane layout
    int C = 0, T = 0, TP = 0, nchunks = 0;
    long long pstride = 0, bstride = 0, bytes = 0;
    size_t off = 0;
};

// rows per plane: PADL zero rows, the data rounded up so that every tile (<= 512 rows, plus up to
// 32 polyphase/halo rows) stays inside the plane, and 32 trailing zero rows
static inline int tc_tp(long long T) { return kPadL + (int)((T + 32 + 511) / 512 * 512) + 32; }

static inline TcPlane tc_plane(int B, int C, long long T, int cw, size_t& cursor) {
    TcPlane p;
    p.C = C; p.T = (int)T; p.TP = tc_tp(T); p.nchunks = C / cw;
    p.pstride = (long long)p.TP * 16;
    p.bstride = p.pstride * p.nchunks;
    p.bytes = p.bstride * B;
    p.off = cursor;
    cursor += ((size_t)p.bytes + 255) / 256 * 256;
    return p;
}

struct TcPlan {
    TcPlane mel, pre;                       // packed mel, conv_pre output
    struct Stage { TcPlane X, H, R, Y, ACC; } st[HFG_MAX_STAGES];
    size_t total = 0;
};

static inline TcPlan tc_plan(const hfg_handle* h, int B, int T, int mode) {
    const int cw = mode == HFG_MODE_BF16 ? 8 : 4;
    TcPlan p;
    size_t cur = 0;
    p.mel = tc_plane(B, h->cfg.n_mels, T, cw, cur);
    p.pre = tc_plane(B, h->cfg.upsample_initial_channel, T, cw, cur);
    long long t = T;
    for (size_t i = 0; i < h->ups.size(); ++i) {
        const UpLayer& U = h->ups[i];
        t = (t - 1) * U.u - 2 * U.p + U.k;
        p.st[i].X = tc_plane(B, U.cout, t, cw, cur);
        p.st[i].H = tc_plane(B, U.cout, t, cw, cur);
        p.st[i].R = tc_plane(B, U.cout, t, cw, cur);
        p.st[i].Y = tc_plane(B, U.cout, t, cw, cur);
        p.st[i].ACC = tc_plane(B, U.cout, t, 4, cur);          // fp32 accumulator cells
    }
    p.total = cur;
    return p;
}

inline size_t tc_workspace_bytes(const hfg_handle* h, int B, int T, int mode) {
    if (!tc_supported(h))
        throw StatusError(HFG_ERR_UNSUPPORTED,
                          "tensor-core modes need every layer's channel counts to be multiples of 16; use HFG_MODE_FP32");
    return tc_plan(h, B, T, mode).total;
}

// --------------------------------------------------------------------------
// launch plan
// --------------------------------------------------------------------------
constexpr int kTcSmemLimit = 220 * 1024;

static inline int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}

template <bool BF16>
static void tc_launch_conv(hfg_handle* h, cudaStream_t st, TcConvArgs a, int B, int cout, const char* label,
                           double flops, double bytes) {
    const int nck_max = std::min(8, a.a_nchunks);
    const int n_kb = (a.a_nchunks + 7) / 8;
    const int span = (a.taps_max - 1) * (a.dil < 0 ? -a.dil : a.dil);
    // tile shape: as many 128-row sub-tiles as TMEM (512 columns) and smem allow
    int MT = std::min(4, 512 / a.N);
    MT = std::min(MT, env_int("HFG_TC_MT", 4));
    MT = std::max(1, std::min(MT, (a.n_q + 127) / 128));
    int sa = std::min(kMaxSA, n_kb), sw = std::min(kMaxSW, std::max(2, env_int("HFG_TC_SW", 4)));
    auto smem_need = [&](int mt, int sa_, int sw_) {
        const size_t R = (size_t)mt * 128 + span;
        return (size_t)sa_ * R * nck_max * 16 + (size_t)sw_ * a.N * nck_max * 16 + (size_t)a.N * 4 + 256;
    };
    while (smem_need(MT, sa, sw) > (size_t)kTcSmemLimit) {
        if (sw > 3) --sw;
        else if (MT > 1) MT /= 2;
        else if (sw > 2) --sw;
        else throw StatusError(HFG_ERR_UNSUPPORTED, "tensor-core tile does not fit shared memory");
    }
    a.MT = MT; a.sa = sa; a.sw = sw;
    a.R = MT * 128 + span;
    a.tiles_per_batch = (a.n_q + MT * 128 - 1) / (MT * 128);
    const size_t smem = smem_need(MT, sa, sw);
    dim3 grid(B * a.tiles_per_batch, a.phases * (cout / a.N), 1);
    h->prof_begin(st, label, flops, bytes);
    tc_conv_kernel<BF16><<<grid, kTcThreads, smem, st>>>(a);
    h->prof_end(st);
    check_cuda(cudaGetLastError(), "tc_conv_kernel launch");
}

template <bool BF16>
static void tc_forward_impl(hfg_handle* h, const float* mel, int B, int T, float* wav, char* ws,
                            cudaStream_t st, float* const* stage_out) {
    constexpr int CW = BF16 ? 8 : 4;
    constexpr int ESZ = BF16 ? 2 : 4;
    const TcPlan plan = tc_plan(h, B, T, BF16 ? HFG_MODE_BF16 : HFG_MODE_TF32);
    const float slope = 0.1f;
    const int n_rb = h->cfg.num_resblocks;
    auto ptr = [&](const TcPlane& p) { return reinterpret_cast<uint8_t*>(ws + p.off); };

    // ---- zero the padding rows of every plane (one launch) ----
    {
        PadJobs jobs{};
        auto add = [&](const TcPlane& p) {
            jobs.job[jobs.n++] = PadJob{ptr(p), (long long)B * p.nchunks, p.TP, p.T};
        };
        add(plan.mel); add(plan.pre);
        for (size_t i = 0; i < h->ups.size(); ++i) {
            add(plan.st[i].X); add(plan.st[i].H); add(plan.st[i].R); add(plan.st[i].Y);
        }
        h->prof_begin(st, "zero_pads", 0, 0);
        tc_zero_pads<<<dim3(64, jobs.n), 256, 0, st>>>(jobs);
        h->prof_end(st);
        check_cuda(cudaGetLastError(), "tc_zero_pads launch");
    }
    // ---- mel -> chunk planes ----
    {
        dim3 grid((T + 127) / 128, plan.mel.nchunks, B);
        h->prof_begin(st, "pack_mel", 0, (double)B * h->cfg.n_mels * T * (4 + ESZ));
        tc_pack_input<BF16><<<grid, 128, 0, st>>>(mel, ptr(plan.mel), h->cfg.n_mels, T, plan.mel.bstride, plan.mel.pstride);
        h->prof_end(st);
        check_cuda(cudaGetLastError(), "tc_pack_input launch");
    }
    auto dump = [&](int idx, const TcPlane& p) {
        if (!stage_out || !stage_out[idx]) return;
        dim3 grid((p.T + 127) / 128, p.nchunks, B);
        tc_unpack_stage<BF16><<<grid, 128, 0, st>>>(ptr(p), stage_out[idx], p.C, p.T, p.bstride, p.pstride, 1.0f / slope);
        check_cuda(cudaGetLastError(), "tc_unpack_stage launch");
    };
    auto conv = [&](const ConvLayer& L, const TcPlane& in, const TcPlane* out, const TcPlane* res,
                    const TcPlane* acc, int acc_mode, const char* label) {
        TcConvArgs a{};
        a.a = ptr(in); a.a_bstride = in.bstride; a.a_pstride = in.pstride; a.a_nchunks = in.nchunks;
        a.N = tc_pick_n(L.cout);
        a.w = reinterpret_cast<const uint8_t*>(BF16 ? L.tc.w_bf16 : L.tc.w_tf32);
        const int n_kb = (in.nchunks + 7) / 8;
        a.w_ntile_stride = (long long)n_kb * L.k * 8 * a.N * 16;
        a.w_phase_stride = 0;
        a.bias = L.bias;
        const TcPlane& og = out ? *out : *res;               // geometry of out/res planes
        a.out = out ? ptr(*out) : nullptr;
        a.res = res ? ptr(*res) : nullptr;
        a.o_bstride = og.bstride; a.o_pstride = og.pstride;
        if (acc) { a.acc = reinterpret_cast<float*>(ptr(*acc)); a.acc_bstride = acc->bstride; a.acc_pstride = acc->pstride; }
        a.acc_mode = acc_mode; a.div = (float)n_rb;
        a.n_q = in.T; a.T_out = in.T;
        a.taps_max = L.k; a.k = L.k; a.u = 1; a.dil = L.dil; a.pad = L.pad; a.phases = 1;
        a.out_stride = 1; a.out_off = 0; a.min_off = -L.pad;
        a.slope = slope;
        const double flops = 2.0 * L.cin * L.cout * L.k * (double)B * in.T;
        const double bytes = (double)B * in.T * ESZ * (L.cin + L.cout * (res ? 2 : 1)) +
                             (acc ? 4.0 * B * in.T * L.cout * (acc_mode == TC_ACC_ADD ? 2 : 1) : 0.0) +
                             (double)ESZ * L.cin * L.cout * L.k;
        tc_launch_conv<BF16>(h, st, a, B, L.cout, label, flops, bytes);
    };

    // conv_pre (reference :238); its output is stored as leaky_relu(x) for ups[0] (:244)
    conv(h->pre, plan.mel, &plan.pre, nullptr, nullptr, TC_ACC_NONE, "conv_pre");
    dump(0, plan.pre);

    const TcPlane* cur = &plan.pre;
    for (size_t i = 0; i < h->ups.size(); ++i) {
        const UpLayer& U = h->ups[i];
        const auto& S = plan.st[i];
        {   // x = ups[i](leaky_relu(x)) as u polyphase convolutions (reference :245)
            TcConvArgs a{};
            a.a = ptr(*cur); a.a_bstride = cur->bstride; a.a_pstride = cur->pstride; a.a_nchunks = cur->nchunks;
            a.N = tc_pick_n(U.cout);
            a.w = reinterpret_cast<const uint8_t*>(BF16 ? U.tc.w_bf16 : U.tc.w_tf32);
            const int n_kb = (cur->nchunks + 7) / 8;
            a.w_ntile_stride = (long long)n_kb * U.taps_max * 8 * a.N * 16;
            a.w_phase_stride = a.w_ntile_stride * (U.cout / a.N);
            a.bias = U.bias;
            a.out = ptr(S.X); a.res = nullptr; a.o_bstride = S.X.bstride; a.o_pstride = S.X.pstride;
            a.acc_mode = TC_ACC_NONE; a.div = 1.f;
            a.T_out = S.X.T;
            a.n_q = (S.X.T - 1 + U.p) / U.u + 1;
            a.taps_max = U.taps_max; a.k = U.k; a.u = U.u; a.dil = -1; a.pad = 0; a.phases = U.u;
            a.out_stride = U.u; a.out_off = -U.p; a.min_off = -(U.taps_max - 1);
            a.slope = slope;
            const double flops = 2.0 * U.cin * U.cout * U.k * (double)B * cur->T;
            const double bytes = (double)B * ESZ * ((double)U.cin * cur->T + (double)U.cout * S.X.T) +
                                 (double)ESZ * U.cin * U.cout * U.k;
            tc_launch_conv<BF16>(h, st, a, B, U.cout, ("ups" + std::to_string(i)).c_str(), flops, bytes);
        }
        dump(1 + 2 * (int)i, S.X);

        // MRF (reference :116-131)
        const std::string lab = "mrf" + std::to_string(i);
        for (int j = 0; j < n_rb; ++j) {
            const auto& rb = h->mrfs[i][j];
            const TcPlane* r = &S.X;
            for (size_t l = 0; l < rb.size(); ++l) {
                const bool last = (l + 1 == rb.size());
                conv(rb[l].c1, *r, &S.H, nullptr, nullptr, TC_ACC_NONE, lab.c_str());
                if (!last) {
                    conv(rb[l].c2, S.H, &S.R, r, nullptr, TC_ACC_NONE, lab.c_str());
                } else {
                    int mode = n_rb == 1 ? TC_ACC_FINAL : (j == 0 ? TC_ACC_WRITE : (j == n_rb - 1 ? TC_ACC_FINAL : TC_ACC_ADD));
                    if (n_rb == 1) {
                        // single resblock: no accumulator traffic, just divide
                        conv(rb[l].c2, S.H, &S.Y, r, nullptr, TC_ACC_NONE, lab.c_str());
                    } else {
                        conv(rb[l].c2, S.H, mode == TC_ACC_FINAL ? &S.Y : nullptr, r, &S.ACC, mode, lab.c_str());
                    }
                }
                r = &S.R;
            }
        }
        cur = &S.Y;
        dump(2 + 2 * (int)i, S.Y);
    }
    // wav = tanh(conv_post(leaky_relu(x)))  (reference :254-256); planes already hold leaky_relu(x)
    {
        const int Tw = cur->T;
        dim3 grid((Tw + 255) / 256, B);
        h->prof_begin(st, "conv_post", 2.0 * h->post_cin * 7 * (double)B * Tw,
                      (double)B * Tw * (ESZ * h->post_cin + 4.0));
        tc_conv_post_tanh<BF16><<<grid, 256, sizeof(float) * h->post_cin * 7, st>>>(
            ptr(*cur), h->post_w, h->post_b, wav, h->post_cin, Tw, 7, 3, cur->bstride, cur->pstride);
        h->prof_end(st);
        check_cuda(cudaGetLastError(), "tc_conv_post_tanh launch");
    }
    (void)CW;
}

inline void tc_forward(hfg_handle* h, const float* mel, int B, int T, float* wav, char* ws, int mode,
                       cudaStream_t st, float* const* stage_out) {
    if (h->cc_major != 10)
        throw StatusError(HFG_ERR_UNSUPPORTED, "tensor-core modes need an sm_100 device (tcgen05)");
    if (mode == HFG_MODE_BF16) tc_forward_impl<true>(h, mel, B, T, wav, ws, st, stage_out);
    else tc_forward_impl<false>(h, mel, B, T, wav, ws, st, stage_out);
}

inline void configure_kernels(hfg_handle*) {
    const int smem = 100 * 1024;
    check_cuda(cudaFuncSetAttribute(conv_tile_fp32<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "attr");
    check_cuda(cudaFuncSetAttribute(conv_tile_fp32<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "attr");
    check_cuda(cudaFuncSetAttribute(conv_tile_fp32<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "attr");
    check_cuda(cudaFuncSetAttribute(tc_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "attr");
    check_cuda(cudaFuncSetAttribute(tc_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "attr");
}

}  // namespace hfg
