// Host side of the tensor-core path: operand packing, workspace plan, launch plan.
#pragma once
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "fp32_kernels.cuh"
#include "model.h"
#include "tc_kernels.cuh"
#include "tc_pair_kernel.cuh"
#include "tc_up_kernel.cuh"

namespace hfg {

// Tuning knobs (HFG_TC_*).  Production builds compile them out: env_int() is the default, always, so
// no environment variable can steer kernel selection.  Builds with -DHFG_TUNING (libhfg_b200_tuning.so, used
// by tools/ and by the kernel-variant tests through HFG_LIB_PATH) read them from the environment once per
// thread, cached by literal address.
#ifndef HFG_TUNING
static inline int env_int(const char*, int dflt) { return dflt; }
#else
static inline int env_int(const char* name, int dflt) {
    struct Ent { const char* name; bool set; int val; };
    thread_local Ent cache[48];
    thread_local int n = 0;
    for (int i = 0; i < n; ++i)
        if (cache[i].name == name) return cache[i].set ? cache[i].val : dflt;
    const char* s = getenv(name);
    Ent e{name, s != nullptr, s ? atoi(s) : 0};
    if (n < 48) cache[n++] = e;
    return e.set ? e.val : dflt;
}
#endif

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
static inline uint16_t f2h(float f) {            // fp32 -> fp16, round-to-nearest-even, saturating to +-65504
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint16_t sign = (uint16_t)((u >> 16) & 0x8000u);
    const uint32_t a = u & 0x7FFFFFFFu;
    if (a > 0x7F800000u) return (uint16_t)(sign | 0x7E00u);            // NaN
    if (a >= 0x477FF000u) return (uint16_t)(sign | 0x7BFFu);           // >= 65520 rounds past the largest finite: saturate
    if (a < 0x33000001u) return sign;                                  // < 2^-25: rounds to zero
    int e = (int)(a >> 23) - 127;
    uint32_t m = (a & 0x7FFFFFu) | 0x800000u;                          // 24-bit significand
    int shift = e < -14 ? (13 + (-14 - e)) : 13;                       // subnormal halves lose more bits
    uint32_t q = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (q & 1u))) ++q;
    uint32_t out = e < -14 ? q : (((uint32_t)(e + 15) << 10) + (q - 0x400u));   // carry into the exponent is correct by construction
    return (uint16_t)(sign | out);
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
static void tc_pack_generic(hfg_handle* h, TcPack& tp, int cin, int cout, int phases, int taps_max, F get,
                            bool pair = false, bool halves = false, int kbc = 8, int only_prec = -1) {
    if (!pair) {
        tp.ok = tc_shape_ok(cin, cout);
        if (!tp.ok) return;
    }
    // halves: two N/2-wide tiles (one per CTA of a pair) in the same [ntile][kb][tap][chunk][n] layout
    const int N = halves ? cout / 2 : (pair ? cout : tc_pick_n(cout)), ntiles = cout / N;
    for (int prec = 0; prec < 3; ++prec) {
        if (only_prec >= 0 && prec != only_prec) continue;
        const int CW = prec == PREC_TF32 ? 4 : 8;
        const int nchunks = cin / CW, nkb = (nchunks + kbc - 1) / kbc;
        // tight layout: [phase][ntile][kb][tap][chunk < nck(kb)][n][cell]
        const size_t per_tile = (size_t)nchunks * N * 16 * taps_max;
        const size_t total = (size_t)phases * ntiles * per_tile;
        std::vector<uint8_t> buf(total, 0);
        for (int ph = 0; ph < phases; ++ph)
            for (int nt = 0; nt < ntiles; ++nt)
                for (int kb = 0; kb < nkb; ++kb) {
                    const int nck = std::min(kbc, nchunks - kbc * kb);
                    for (int tap = 0; tap < taps_max; ++tap) {
                        uint8_t* blk = buf.data() + ((size_t)ph * ntiles + nt) * per_tile +
                                       ((size_t)kb * taps_max * kbc + (size_t)tap * nck) * N * 16;
                        for (int c = 0; c < nck; ++c)
                            for (int n = 0; n < N; ++n) {
                                uint8_t* cell = blk + ((size_t)c * N + n) * 16;
                                for (int e = 0; e < CW; ++e) {
                                    const int ci = (kbc * kb + c) * CW + e;
                                    const float v = get(nt * N + n, ci, ph, tap);
                                    if (prec == PREC_TF32) {
                                        const float q = f2tf32(v);
                                        memcpy(cell + 4 * e, &q, 4);
                                    } else {
                                        const uint16_t q = prec == PREC_BF16 ? f2bf(v) : f2h(v);
                                        memcpy(cell + 2 * e, &q, 2);
                                    }
                                }
                            }
                    }
                }
        void* d = h->upload(buf);
        if (pair) {
            tp.w_pair[prec][halves ? 1 : 0][kbc == 4 ? 1 : 0] = d;
            if (halves) tp.half_stride[prec][kbc == 4 ? 1 : 0] = (long long)per_tile;
        } else {
            tp.w[prec] = d;
        }
    }
}

inline void tc_pack_conv(hfg_handle* h, ConvLayer& L, const HostTensor& w, const HostTensor&) {
    const int cin = L.cin, k = L.k;
    auto get = [&](int n, int ci, int, int tap) {
        return w.data[((size_t)n * cin + ci) * k + tap];                // Conv1d weight [C_out, C_in, k]
    };
    tc_pack_generic(h, L.tc, L.cin, L.cout, 1, L.k, get);
    if (!L.tc.ok || L.cin != L.cout || L.cout > 256) return;
    // packs of the fused ResBlock kernel: whole-N (one CTA) and N-halves (cta_group::2 pair), K blocks of
    // 8 cells; tf32 additionally with K blocks of 4 cells (the fp32 H tile leaves less room for A stages)
    const bool can_halve = L.cout % 32 == 0;
    for (int prec = 0; prec < 3; ++prec) L.tc.w_pair[prec][0][0] = L.tc.w[prec];
    if (can_halve) tc_pack_generic(h, L.tc, L.cin, L.cout, 1, L.k, get, true, true, 8);
    // Space-to-depth form of a dilation-1 convolution for the narrow layers (conv2 of every pair): GEMM row m holds
    // time steps 2m and 2m+1, so input "channel" (p'', ci) and output "channel" (p, co) run over 2N values and the k
    // taps collapse into (k + 1) / 2 taps q with  W'[(p, co), (p'', ci), q] = W[co, ci, 2q + p'' - p]  (0 outside).
    if (L.dil == 1 && L.cout <= 64 && L.cout % 32 == 0 && L.k % 2 == 1) {
        const int N = L.cout, k2 = (k + 1) / 2;
        auto get2 = [&](int n, int ci, int, int q) {
            const int p = n / N, co = n % N, pi = ci / N, c = ci % N;
            const int j = 2 * q + pi - p;
            return (j >= 0 && j < k) ? w.data[((size_t)co * cin + c) * k + j] : 0.f;
        };
        for (int prec : {PREC_BF16, PREC_FP16})
            for (int halves = 0; halves < 2; ++halves) {
                TcPack tmp;
                tc_pack_generic(h, tmp, 2 * N, 2 * N, 1, k2, get2, true, halves == 1, 8, prec);
                L.tc.w_s2d[prec][halves] = tmp.w_pair[prec][halves][0];
                if (halves) L.tc.s2d_half_stride[prec] = tmp.half_stride[prec][0];
            }
    }
    // K blocks of 4 cells let tf32 C=128 run MT=2 tiles, but measured slower than MT=1 with K blocks of 8
    // (profiles/r1_tuning.md): packed only on request (HFG_TC_PACK_KBC4=1, tuning experiments)
    if (L.cin / 4 >= 16 && env_int("HFG_TC_PACK_KBC4", 0)) {
        tc_pack_generic(h, L.tc, L.cin, L.cout, 1, L.k, get, true, false, 4, /*only tf32*/ PREC_TF32);
        if (can_halve) tc_pack_generic(h, L.tc, L.cin, L.cout, 1, L.k, get, true, true, 4, PREC_TF32);
    }
}
inline void tc_pack_up(hfg_handle* h, UpLayer& L, const HostTensor& w, const HostTensor&) {
    const int cout = L.cout, k = L.k, u = L.u;
    auto get = [&](int n, int ci, int ph, int tap) {
        const int j = ph + tap * u;                                     // ConvTranspose1d weight [C_in, C_out, k]
        return j < k ? w.data[((size_t)ci * cout + n) * k + j] : 0.f;
    };
    tc_pack_generic(h, L.tc, L.cin, L.cout, L.u, L.taps_max, get);
    // Phases stacked along N: every phase multiplies the SAME activation rows (x[q - tap]) by its own tap
    // set, so the u phase-GEMMs are one GEMM with u * cout virtual output channels -- the activation tile is
    // loaded once instead of once per phase and the MMAs are u times wider.  A 32-column epilogue step must
    // not straddle two phases, hence cout % 32 == 0.
    L.tc_stack.ok = false;
    if (L.tc.ok && u > 1 && cout % 32 == 0 && tc_pick_n(u * cout) >= cout && (u * cout <= 128 || env_int("HFG_TC_UPS_STACK", 0)))
        tc_pack_generic(h, L.tc_stack, L.cin, u * cout, 1, L.taps_max,
                        [&](int nv, int ci, int, int tap) { return get(nv % cout, ci, nv / cout, tap); });
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

// rows per plane: PADL zero rows, the data, and enough trailing zero rows that a tile (<= 512 rows)
// starting anywhere below T, plus its halo / polyphase overhang (<= 32 + 25 rows), stays inside
static inline int tc_tp(long long T) { return kPadL + (int)((T + 127) / 128 * 128) + 512 + 64; }

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

static inline int tc_prec_of_mode(int mode) {
    return mode == HFG_MODE_BF16 ? PREC_BF16 : (mode == HFG_MODE_FP16 ? PREC_FP16 : PREC_TF32);
}
static inline bool tc_is_tc_mode(int mode) { return mode == HFG_MODE_TF32 || mode == HFG_MODE_BF16 || mode == HFG_MODE_FP16; }

struct PairGeom { int MT, sa, sw, G, R1, RH, TO, ctas, kbc; size_t smem; int occ; bool ok; int s2d, k2, G2; size_t stage; int groups; };
static inline PairGeom tc_pair_geometry(const hfg_handle* h, const PairLayers& P, int n_chunks, int prec);

// How the MRF sum (reference models/hifigan.py:126-131) of stage i is formed.
//   sum planes (default): every resblock but the last writes its output like any other pair; the last pair of the
//     LAST resblock adds those planes in its pre2 phase.  No fp32 accumulator plane exists.
//   fp32 ACC plane (WRITE -> ADD -> FINAL): only when that last pair does not fit the fused kernel and runs as
//     two plain convolutions.
static inline bool tc_stage_sum_planes(const hfg_handle* h, size_t i, int n_chunks, int prec) {
    const auto& mrf = h->mrfs[i];
    if (mrf.size() < 2) return false;                       // a single resblock has nothing to sum
    if ((int)mrf.size() - 1 > HFG_MAX_STAGES - 1 || !env_int("HFG_TC_SUM_PLANES", 1)) return false;
    return tc_pair_geometry(h, mrf.back().back(), n_chunks, prec).ok;
}

// The tf32 mode on fp16 operand planes ("split plan").  kind::tf32 reads 10 mantissa bits of each fp32 operand, so an
// MMA fed from fp32 planes moves 4 bytes per element through L2 and -- what bounds the narrow stages -- the 64 B/clk
// shared-memory operand read to use 10 + 8 bits of them, at half the kind::f16 rate.  In the split plan an activation
// of the residual stream is stored as TWO fp16 planes, hi = fp16(v) and lo = fp16(v - hi):
//   * hi is the round-to-nearest 10-bit-mantissa operand (tf32 truncates): every MMA reads hi planes only, at the
//     fp16 mode's rate and operand bytes;
//   * hi + lo carries 22 mantissa bits (absolute floor 3e-8): the residual x of a fused pair is formed from both --
//     the producer copies the lo cells of the tile's own rows into the (then idle) H-tile buffer, so nothing is
//     loaded synchronously and no shared memory is added -- and the pair writes both halves of its result;
//   * HBM bytes per element are those of the fp32 planes (2 + 2 read, 2 + 2 written).
// Planes that only feed MMAs (mel, conv_pre output, an MRF output followed by an upsampler) have no lo twin; the
// finished resblock outputs that only feed the MRF sum, and the last MRF output (conv_post), are plain fp32 planes.
// Used when every pair of the configuration fits the fused kernel (else the fp32-plane kind::tf32 path).
static inline bool tc_tf32_mixed_eval(const hfg_handle* h) {
    if (!env_int("HFG_TC_TF32_MIXED", 1)) return false;
    if (h->post_cin % 8 != 0) return false;
    for (size_t i = 0; i < h->ups.size(); ++i) {
        const int C = h->ups[i].cout;
        if (C % 8 != 0) return false;
        for (auto& rb : h->mrfs[i])
            for (auto& P : rb)
                if (!tc_pair_geometry(h, P, C / 8, PREC_FP16).ok) return false;
        if (h->mrfs[i].size() > 1 && !tc_stage_sum_planes(h, i, C / 8, PREC_FP16)) return false;
    }
    return true;
}

// evaluated once per committed weight set (hfg_handle::tf32_split, reset by commit): every forward asks
static inline bool tc_tf32_mixed(const hfg_handle* h) {
    if (h->tf32_split < 0) const_cast<hfg_handle*>(h)->tf32_split = tc_tf32_mixed_eval(h) ? 1 : 0;
    return h->tf32_split == 1;
}

struct TcPlan {
    bool mixed = false;                     // tf32 mode on fp16 hi + lo planes (tc_tf32_mixed)
    TcPlane mel, pre;                       // packed mel, conv_pre output
    // per stage: X = upsampler output, Y = MRF output and, per resblock (they run concurrently), two ping-pong
    // planes R, H.  Only where a pair does not fit the fused kernel: a scratch S for its intermediate, and an
    // fp32 running-sum plane ACC if that pair is the one that forms the MRF sum.
    // split plan: Xl / Rl / Hl = lo twins of X / R / H (same geometry); F32 = finished output of a resblock that
    // only feeds the MRF sum (fp32); the MRF output is Y (hi only) where an upsampler follows and Y32 (fp32) in
    // the last stage, which conv_post reads
    struct Stage {
        TcPlane X, Y, ACC, Xl, Y32; bool sum_planes = false;
        struct { TcPlane R, H, S, Rl, Hl, F32; bool fused = true; } rb[HFG_MAX_STAGES];
    } st[HFG_MAX_STAGES];
    size_t lens_off = 0;                    // int32 [(num_upsamples + 2)][B]: per-utterance row counts (tc_len_table)
    size_t total = 0;
};

static inline TcPlan tc_plan(const hfg_handle* h, int B, int T, int mode) {
    TcPlan p;
    p.mixed = mode == HFG_MODE_TF32 && tc_tf32_mixed(h);
    const int prec = p.mixed ? PREC_FP16 : tc_prec_of_mode(mode);
    const int cw = prec == PREC_TF32 ? 4 : 8;
    size_t cur = 0;
    p.mel = tc_plane(B, h->cfg.n_mels, T, cw, cur);
    p.pre = tc_plane(B, h->cfg.upsample_initial_channel, T, cw, cur);
    long long t = T;
    for (size_t i = 0; i < h->ups.size(); ++i) {
        const UpLayer& U = h->ups[i];
        t = (t - 1) * U.u - 2 * U.p + U.k;
        auto& S = p.st[i];
        const bool last_stage = i + 1 == h->ups.size();
        S.X = tc_plane(B, U.cout, t, cw, cur);
        if (p.mixed) S.Xl = tc_plane(B, U.cout, t, cw, cur);
        if (p.mixed && last_stage) S.Y32 = tc_plane(B, U.cout, t, 4, cur);
        else S.Y = tc_plane(B, U.cout, t, cw, cur);
        const int n_rb = h->cfg.num_resblocks;
        S.sum_planes = tc_stage_sum_planes(h, i, S.X.nchunks, prec);
        for (int j = 0; j < n_rb; ++j) {
            const auto& rb = h->mrfs[i][j];
            bool fused = true;
            for (auto& P : rb) fused = fused && tc_pair_geometry(h, P, S.X.nchunks, prec).ok;
            S.rb[j].fused = fused;
            // R: output of pair 0 (or, with sum planes, the finished output of a one-pair resblock); H: its partner
            // (split plan: the finished outputs go to F32 instead, so R / H only hold intermediate results)
            const bool fin_rh = S.sum_planes && j + 1 < n_rb && !p.mixed;
            const bool need_r = rb.size() > 1 || fin_rh;
            const bool need_h = rb.size() > 2 || (fin_rh && rb.size() > 1);
            if (need_r) S.rb[j].R = tc_plane(B, U.cout, t, cw, cur);
            if (need_h) S.rb[j].H = tc_plane(B, U.cout, t, cw, cur);
            if (need_r && p.mixed) S.rb[j].Rl = tc_plane(B, U.cout, t, cw, cur);
            if (need_h && p.mixed) S.rb[j].Hl = tc_plane(B, U.cout, t, cw, cur);
            if (p.mixed && j + 1 < n_rb) S.rb[j].F32 = tc_plane(B, U.cout, t, 4, cur);
            if (!fused) S.rb[j].S = tc_plane(B, U.cout, t, cw, cur);
        }
        if (!S.sum_planes && n_rb > 1) S.ACC = tc_plane(B, U.cout, t, 4, cur);          // fp32 accumulator cells
    }
    p.lens_off = cur;
    cur += ((size_t)(h->ups.size() + 2) * B * sizeof(int) + 255) / 256 * 256;
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


// taps per W stage: aim at >= 16 KB per bulk copy so tiny layers do not drown in barrier round trips
static inline int tc_tap_group(int N, int nck_max, int taps) {
    const int tap_bytes = N * nck_max * 16;
    const int g = std::max(1, env_int("HFG_TC_STAGE_BYTES", 16384) / tap_bytes);
    return std::min(g, taps);
}

template <int P>
static void tc_launch_conv(hfg_handle* h, cudaStream_t st, TcConvArgs a, int B, int cout, const char* label,
                           double flops, double bytes) {
    const int nck_max = std::min(8, a.a_nchunks);
    const int n_kb = (a.a_nchunks + 7) / 8;
    const int span = (a.taps_max - 1) * (a.dil < 0 ? -a.dil : a.dil);
    // tile shape: as many 128-row sub-tiles as TMEM (512 columns) and smem allow
    int MT = std::min(4, 512 / a.N);
    MT = std::min(MT, env_int("HFG_TC_MT", 4));
    if (a.u > 1) MT = std::min(MT, env_int("HFG_TC_UPS_MT", 4));             // polyphase upsamplers (tuning knob)
    MT = std::max(1, std::min(MT, (a.n_q + 127) / 128));
    int sa = std::min(kMaxSA, n_kb), sw = std::min(kMaxSW, std::max(2, env_int("HFG_TC_SW", 4)));
    const int G = tc_tap_group(a.N, nck_max, a.taps_max);
    a.tap_group = G;
    auto smem_need = [&](int mt, int sa_, int sw_) {
        const size_t R = (size_t)mt * 128 + span;
        return (size_t)sa_ * R * nck_max * 16 + (size_t)sw_ * G * a.N * nck_max * 16 + (size_t)a.N * 4 + 256;
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
    dim3 grid(a.phases * (cout / a.N), B * a.tiles_per_batch, 1);
    // tuning only: HFG_TC_CONV_TIMELINE="<label>:<file>" stamps the phases of the first 64 CTAs of that launch
    unsigned long long* tl_dev = nullptr;
#ifdef HFG_TUNING
    const char* tl_env = getenv("HFG_TC_CONV_TIMELINE");
#else
    const char* tl_env = nullptr;
#endif
    const char* tl_path = nullptr;
    if (tl_env) {
        const char* colon = strchr(tl_env, ':');
        if (colon && (size_t)(colon - tl_env) == strlen(label) && strncmp(tl_env, label, strlen(label)) == 0) {
            tl_path = colon + 1;
            check_cuda(cudaMalloc((void**)&tl_dev, 64 * 8 * 8), "cudaMalloc(timeline)");
            check_cuda(cudaMemsetAsync(tl_dev, 0, 64 * 8 * 8, st), "cudaMemset(timeline)");
            a.timeline = tl_dev;
        }
    }
    h->prof_begin(st, label, flops, bytes);
    // 8 epilogue warps when the CTA owns its SM anyway (big tiles: latency-bound epilogue, profiles/r1_tuning.md);
    // 4 when two CTAs can share the SM (narrow layers), which hides the epilogue better than more warps
    const int threads = (2 * (smem + 1024) <= 227 * 1024 && env_int("HFG_TC_CONV_WARPS", 0) != 8) ? 192 : kTcThreads;
    tc_conv_kernel<P><<<grid, threads, smem, st>>>(a);
    h->prof_end(st);
    check_cuda(cudaGetLastError(), "tc_conv_kernel launch");
    if (tl_dev) {
        std::vector<unsigned long long> hb(64 * 8);
        check_cuda(cudaStreamSynchronize(st), "sync(timeline)");
        check_cuda(cudaMemcpy(hb.data(), tl_dev, 64 * 8 * 8, cudaMemcpyDeviceToHost), "cudaMemcpy(timeline)");
        cudaFree(tl_dev);
        if (FILE* f = fopen(tl_path, "a")) {
            fprintf(f, "# %s N=%d MT=%d sa=%d sw=%d G=%d smem=%zu threads=%d grid=%dx%d\n", label, a.N, a.MT, a.sa, a.sw, a.tap_group,
                    smem, threads, grid.x, grid.y);
            for (int i = 0; i < 64; ++i) {
                for (int e = 0; e < 8; ++e) fprintf(f, "%llu ", hb[i * 8 + e]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
}

// ---- fused ResBlock pair ----

// Persistent, double-buffered launch of a plain-epilogue convolution (conv_pre, upsamplers): tc_up_kernel.
// Returns false when the geometry does not fit (the caller falls back to tc_launch_conv).
template <int P>
static bool tc_launch_up(hfg_handle* h, cudaStream_t st, TcConvArgs a, int B, int cout, const char* label,
                         double flops, double bytes) {
    if (!env_int("HFG_TC_UP_PERSIST", 1)) return false;
    const bool plain = !a.res && a.acc_mode == TC_ACC_NONE;
    // ResBlock convolutions (tf32 C = 256, unfused): measured slower than the one-shot kernel -- 0.77 vs 0.71 ms
    // for the stage with the resblocks concurrent (a 256-wide tile leaves 128 rows per accumulator buffer and
    // one-tap weight stages) -- so they stay on tc_conv_kernel unless asked for
    if (!plain && !env_int("HFG_TC_UP_RESBLOCK", 0)) return false;
    const int nck_max = std::min(8, a.a_nchunks);
    const int n_kb = (a.a_nchunks + 7) / 8;
    const int span = (a.taps_max - 1) * (a.dil < 0 ? -a.dil : a.dil);
    int MT = std::min(4, 256 / a.N);                                   // two accumulator buffers: 2 * MT * N <= 512 columns
    if (MT < 1) return false;
    MT = std::min(MT, env_int("HFG_TC_UP_MT", 4));
    MT = std::max(1, std::min(MT, (a.n_q + 127) / 128));
    // one barrier round trip per weight stage (profiles/r1_tuning.md section 6): stages of up to 32 KB
    const int tap_bytes = a.N * nck_max * 16;
    const int G = std::min(a.taps_max, std::max(1, env_int("HFG_TC_UP_STAGE_BYTES", 32768) / tap_bytes));
    a.tap_group = G;
    // a 256-wide tile leaves one 128-row sub-tile per accumulator buffer: twice the items of the one-shot
    // kernel and a second, mostly empty round on 148 SMs (ups0: 64 vs 52 us) -- persistent only when
    // narrower, or when everything fits one round anyway
    {
        const int mt1 = std::max(1, std::min(MT, (a.n_q + 127) / 128));
        const int items = B * ((a.n_q + mt1 * 128 - 1) / (mt1 * 128)) * a.phases * (cout / a.N);
        if (plain && a.N > 128 && items > h->sm_count && !env_int("HFG_TC_UP_WIDE", 0)) return false;
    }
    int sa = std::min(kUpMaxSA, std::max(2, env_int("HFG_TC_UP_SA", 4)));
    int sw = std::min(kUpMaxSW, std::max(2, env_int("HFG_TC_UP_SW", 4)));
    auto smem_need = [&](int mt, int sa_, int sw_) {
        const size_t R = (size_t)mt * 128 + span;
        return (size_t)sa_ * R * nck_max * 16 + (size_t)sw_ * G * a.N * nck_max * 16 + (size_t)((cout + 3) & ~3) * 4 + 256;
    };
    while (smem_need(MT, sa, sw) > (size_t)kTcSmemLimit) {
        if (sw > 3) --sw;
        else if (sa > 2) --sa;
        else if (sw > 2) --sw;
        else if (MT > 1) MT /= 2;
        else return false;
    }
    (void)n_kb;
    a.MT = MT; a.sa = sa; a.sw = sw;
    a.R = MT * 128 + span;
    a.tiles_per_batch = (a.n_q + MT * 128 - 1) / (MT * 128);
    TcUpArgs ua{};
    ua.c = a;
    ua.n_ctile = cout / a.N;
    ua.pn_per_tile = a.phases * ua.n_ctile;
    const long long n_items = (long long)B * a.tiles_per_batch * ua.pn_per_tile;
    if (n_items > 0x7fffffffLL) return false;                          // item index is an int in the kernel
    ua.n_items = (int)n_items;
    ua.cout_total = cout;
    ua.epi_sleep_ns = env_int("HFG_TC_EPI_SLEEP_NS", 0);
    const size_t smem = smem_need(MT, sa, sw);
    const int grid = std::min(ua.n_items, h->sm_count);
    // items of a CTA: strided (item, item + grid, ...) or one contiguous block -- the phases / channel tiles of one
    // activation tile are adjacent items, so a block keeps them on one SM back to back (their output cells interleave
    // in the same sectors, the tile comes from L2 again while it is hot)
    ua.contig = env_int("HFG_TC_UP_CONTIG", 0) ? (ua.n_items + grid - 1) / grid : 0;
    if (env_int("HFG_TC_VERBOSE", 0))
        fprintf(stderr, "[up] %s N=%d MT=%d G=%d sa=%d sw=%d smem=%zu grid=%d items=%d\n", label, a.N, MT, G, sa, sw, smem,
                grid, ua.n_items);
    h->prof_begin(st, label, flops, bytes);
    tc_up_kernel<P><<<grid, kTcThreads, smem, st>>>(ua);
    h->prof_end(st);
    check_cuda(cudaGetLastError(), "tc_up_kernel launch");
    return true;
}

// Operand precision of conv2 inside the fused pair (tc_pair_kernel.cuh): the on-chip intermediate of the tf32
// mode is held as fp16 -- tf32's own 10-bit mantissa -- which halves its shared-memory footprint and operand
// reads and runs conv2 at the fp16 MMA rate.
static inline int tc_pair_p2(int prec) {
    return (prec == PREC_TF32 && env_int("HFG_TC_TF32_H16", 1)) ? PREC_FP16 : prec;
}

static inline int tc_pair_ctas(const PairLayers& P, int prec) {
    const bool bf16 = prec != PREC_TF32 || tc_pair_p2(prec) != PREC_TF32;   // 2-byte H tile: the pair fits for every C
    // CTA pairs (cta_group::2: half the weight staging and B-operand reads per SM) where measured faster
    // (profiles/r1_tuning.md sections 5 and 7): every C >= 64 layer, except tf32 C = 256 whose fp32 H tile
    // leaves too little shared memory (the fused pair measured slower than two unfused launches there).
    // C = 32 stays single-CTA (0.560 vs 0.582 ms for the stage).
    const int N = P.c1.cout;
    const int dflt = (N >= 64 && (bf16 || N <= 128)) ? 2 : 1;
    const int want = env_int("HFG_TC_PAIR_CTAS", dflt);
    return (want == 2 && P.c1.tc.w_pair[prec][1][0] && P.c2.tc.w_pair[tc_pair_p2(prec)][1][0]) ? 2 : 1;
}

static inline PairGeom tc_pair_geometry(const hfg_handle* h, const PairLayers& P, int n_chunks, int prec) {
    PairGeom g{};
    const int N = P.c1.cout, k = P.c1.k, p1 = P.c1.pad, p2 = P.c2.pad;
    g.ok = false;
    if (env_int("HFG_TC_FUSE", 1) == 0) return g;
    if (N > env_int("HFG_TC_FUSE_MAXC", 256) || N % 16 != 0 || P.c1.cin != N || P.c2.cin != N || P.c2.cout != N) return g;
    if (P.c2.dil != 1 || p1 + p2 > kPadL || 2 * p2 >= 128 || p2 > p1) return g;
    const int mt_cap = env_int("HFG_TC_PAIR_MT", 4);
    const int ctas = tc_pair_ctas(P, prec);
    const int prec2 = tc_pair_p2(prec);
    const int n_chunks2 = N / (prec2 == PREC_TF32 ? 4 : 8);               // cells per row of the H tile
    const int NB = N / ctas;                                              // weight rows staged per CTA
    // Candidates from the largest tile down.  Measured rule (profiles/r1_tuning.md): two co-resident
    // CTAs per SM beat one CTA with a larger tile (one CTA's epilogue hides behind the other's MMAs),
    // so take the largest MT that still allows 2 CTAs/SM, else the largest MT that fits.  K blocks of 4
    // cells (where packed) are tried after 8: they halve the A stages and let a larger MT fit.
    PairGeom best{};
    best.ok = false;
    for (int MT : {4, 2, 1}) {
        if (MT > mt_cap || 2 * MT * N > 512) continue;
      for (int kbc : {8, 4}) {
        if (!P.c1.tc.w_pair[prec][ctas - 1][kbc == 4 ? 1 : 0] || !P.c2.tc.w_pair[prec2][ctas - 1][kbc == 4 ? 1 : 0]) continue;
        if (kbc == 4 && env_int("HFG_TC_PAIR_NO_KBC4", 0)) continue;
        const int nck_max = std::min(kbc, n_chunks), n_kb = (n_chunks + kbc - 1) / kbc;
        // W ring.  Measured (profiles/r1_tuning.md section 6): with the data always ready the kernel is still
        // bound by the MMA warp's per-stage round trip (two mbarrier waits, fence, commit), so stages are made
        // as FAT as shared memory allows -- G taps per stage with at least `sw_min` stages -- rather than deep.
        // When a second CTA can share the SM the budget is half the SM (that co-residency is worth more).
        // A ring: `sa_tiles` tiles of activations in flight (the next tile's rows stream in while this
        // tile is in its conv2 / epilogues)
        const int R1 = MT * 128 + 2 * p1;
        // conv2 in space-to-depth form (narrow layers, 2-byte intermediate): H' has MT * 64 + k2 - 1 rows of 2N channels
        // Measured per layer (profiles/r2_tuning.md section 9, bf16, us without -> with): C = 32 k = 3 43.2 -> 43.5,
        // k = 7 57.5 -> 52.5, k = 11 74.3 -> 64.3; C = 64 (CTA pair, N2 = 128) 43 -> 51 / 51 -> 60 / 62 -> 69; tf32 mode
        // (MT = 2 tiles) 0.714 -> 0.723 ms for stage 3.  Hence: 2-byte modes, C = 32, k >= 5 (HFG_TC_S2D = 2 forces it
        // wherever it is possible).
        const int s2d_want = env_int("HFG_TC_S2D", 1);
        const bool s2d_pays = prec != PREC_TF32 && N <= 32 && k >= 5;
        const bool s2d = (s2d_want == 2 || (s2d_want == 1 && s2d_pays)) && N <= 64 && N % 32 == 0 && prec2 != PREC_TF32 &&
                         MT % 2 == 0 && P.c2.tc.w_s2d[prec2][ctas - 1] != nullptr;
        const int k2 = s2d ? (k + 1) / 2 : k;
        const int RH = s2d ? (MT * 64 + k2 - 1 + 7) / 8 * 8 : (MT * 128 + 2 * p2 + 7) / 8 * 8;
        const int h_chunks = s2d ? 2 * n_chunks2 : n_chunks2;
        const size_t tap_bytes = (size_t)NB * nck_max * 16;
        const size_t tap2_bytes = s2d ? (size_t)(2 * N / ctas) * std::min(kbc, h_chunks) * 16 : tap_bytes;
        const size_t a_stage = (size_t)R1 * nck_max * 16;
        const size_t fixed0 = (size_t)h_chunks * RH * 16 + (size_t)2 * N * 4 + (size_t)env_int("HFG_TC_PAIR_PAD", 512);
        int ncols = 32;
        while (ncols < 2 * MT * N) ncols <<= 1;
        const int sw_min = std::max(2, env_int("HFG_TC_PAIR_SWMIN", 2));
        const size_t stage_cap = (size_t)env_int("HFG_TC_STAGE_BYTES", 65536);
        auto make = [&](size_t budget) {
            PairGeom c{};
            c.ok = false;
            // A ring.  One slot per K block of the tile (the whole tile resident: conv1 never waits for a slot
            // to drain) where that still leaves W stages of >= 3 taps, else two slots: tf32 C=128 gains 9 %
            // with 4 slots (G = 4), bf16 C=256 loses 7 % (G = 2 instead of 3) -- profiles/r1_tuning.md section 7.
            // More than one tile in flight (HFG_TC_PAIR_SA_TILES) measured slower everywhere.
            int sa = std::min(kPairMaxSA, n_kb * std::max(1, env_int("HFG_TC_PAIR_SA_TILES", 1)));
            sa = std::min(sa, std::max(1, env_int("HFG_TC_PAIR_SA_CAP", kPairMaxSA)));
            const int sa_floor = std::min(2, n_kb);
            auto taps_per_stage = [&](int s_a) -> long long {
                const long long room = (long long)budget - (long long)fixed0 - (long long)s_a * (long long)a_stage;
                return room <= 0 ? 0 : room / (long long)(sw_min * tap_bytes);
            };
            while (sa > sa_floor && taps_per_stage(sa) < std::min(k, 3)) --sa;
            const size_t fixed = fixed0 + sa * a_stage;
            if (fixed + sw_min * std::max(tap_bytes, tap2_bytes) > budget) return c;
            int G = (int)std::min<size_t>((size_t)k, (budget - fixed) / (sw_min * tap_bytes));
            G = (int)std::min<size_t>((size_t)G, std::max<size_t>(1, stage_cap / tap_bytes));
            // one ring serves both convolutions: a slot holds G taps of conv1 or G2 taps of conv2
            const size_t stage = std::max((size_t)G * tap_bytes, tap2_bytes);
            const int G2 = s2d ? (int)std::min<size_t>((size_t)k2, stage / tap2_bytes) : G;
            const int sw = (int)std::min<size_t>(kMaxSW, (budget - fixed) / stage);
            if (sw < 2) return c;
            c.ctas = ctas; c.kbc = kbc;
            c.MT = MT; c.sa = sa; c.sw = sw; c.G = G; c.R1 = R1; c.RH = RH; c.TO = MT * 128 - 2 * p2;
            c.s2d = s2d ? 1 : 0; c.k2 = k2; c.G2 = G2; c.stage = stage;
            // the tile as two independent halves (TcPairArgs::groups): where the tile has an even number of sub-tiles
            // in both accumulators and every activation K block is resident, for the layers whose tile time is the
            // epilogue chain rather than the MMAs (k <= 7; k = 11 is MMA-bound and would only pay the second weight pass)
            // Measured (profiles/r2_tuning.md section 13): bit-identical, but slower on every layer except C = 128 k = 3
            // (50.2 -> 48.8 us): the second weight pass doubles the MMA warp's per-stage barrier round trips.  Off.
            const int gr_want = env_int("HFG_TC_PAIR_GROUPS", 0);
            c.groups = (gr_want && MT % 2 == 0 && (!s2d || MT % 4 == 0) && sa == n_kb &&
                        (gr_want == 2 || k <= env_int("HFG_TC_PAIR_GROUPS_KMAX", 7))) ? 2 : 1;
            c.smem = fixed + (size_t)sw * stage;
            c.occ = std::max(1, std::min((int)((227 * 1024) / (c.smem + 1024)), 512 / ncols));
            c.ok = true;
            return c;
        };
        PairGeom c = make((size_t)kTcSmemLimit);
        if (!c.ok) continue;
        // a larger tile is not worth W stages thinner than 3 taps (tf32 C=128: MT=2 with G=1 1.63 ms for the
        // stage, MT=1 with G=4 1.43 ms)
        if (c.G < std::min(k, 3) && MT > 1) continue;
        // a second CTA could share the SM: worth more than fat stages for N <= 64 (stage 2: 63.6 vs 69.5 us,
        // stage 3: 73.8 vs 124 us at k = 11), not for N = 128 (115.8 vs 98.6 us)
        if (512 / ncols >= 2 && env_int("HFG_TC_PAIR_OCC2", N <= 64 ? 1 : 0)) {
            PairGeom c2 = make((227 * 1024) / 2 - 1024);
            if (c2.ok) c = c2;
        }
        if (!best.ok || (best.occ < 2 && c.occ >= 2)) best = c;
        break;                                           // this MT fits with this K block: no need for the smaller one
      }
        if (best.ok && best.occ >= 2) break;
    }
    (void)h;
    return best;
}

// LO: the kernel variant of the split plan -- in_lo / out_lo are the lo twins of in / out (same geometry; out_lo may
// be null), out32 (fp32 cells, geometry o32_b / o32_p) replaces out where the result only feeds the MRF sum or
// conv_post; sum_in then holds such fp32 planes
template <int P, bool LO = false>
static void tc_launch_pair(hfg_handle* h, cudaStream_t st, const PairLayers& L, const PairGeom& g,
                           const uint8_t* in, long long in_b, long long in_p, int n_chunks,
                           uint8_t* out, long long out_b, long long out_p,
                           float* acc, long long acc_b, long long acc_p, int acc_mode, float div,
                           const uint8_t* const* sum_in, int n_sum, const int* len_rows,
                           int B, int T, const char* label,
                           const uint8_t* in_lo = nullptr, uint8_t* out_lo = nullptr, uint8_t* out32 = nullptr,
                           long long o32_b = 0, long long o32_p = 0) {
    constexpr int ESZ = Prec<P>::ESZ;
    TcPairArgs a{};
    a.a = in; a.a_bstride = in_b; a.a_pstride = in_p;
    const int vk = g.kbc == 4 ? 1 : 0;
    const int P2 = tc_pair_p2(P);
    a.w1 = reinterpret_cast<const uint8_t*>(L.c1.tc.w_pair[P][g.ctas - 1][vk]);
    a.w2 = reinterpret_cast<const uint8_t*>(g.s2d ? L.c2.tc.w_s2d[P2][g.ctas - 1] : L.c2.tc.w_pair[P2][g.ctas - 1][vk]);
    a.w_half_stride = L.c1.tc.half_stride[P][vk];
    a.w2_half_stride = g.s2d ? L.c2.tc.s2d_half_stride[P2] : L.c2.tc.half_stride[P2][vk];
    a.s2d = g.s2d; a.k2 = g.k2; a.tap_group2 = g.G2; a.w_stage_bytes = (unsigned)g.stage;
    // Twelve epilogue warps (three per TMEM lane quarter) for the variants that own their SM (MINB = 1).  Measured
    // (profiles/r2_tuning.md section 14): bit-identical and SLOWER (C = 128 k = 3: 50.9 -> 59.6 us) -- the epilogue
    // chain is not bound by per-warp latency but shares the shared-memory bandwidth with the MMAs' operand reads.
    // The variants are instantiated in the tuning build only.
    const bool two = g.occ >= 2 && env_int("HFG_TC_PAIR_MINB", 2) >= 2;
#ifdef HFG_TUNING
    const bool ew12 = !two && env_int("HFG_TC_PAIR_EW", 8) == 12;
#else
    constexpr bool ew12 = false;
#endif
    a.groups = ew12 ? 1 : g.groups;
    a.kbc = g.kbc;
    a.poll_ns = env_int("HFG_TC_POLL_NS", 40);
    a.epi_sleep_ns = env_int("HFG_TC_EPI_SLEEP_NS", 0);
    a.pdl = env_int("HFG_TC_PDL", 0);
    a.dbg = env_int("HFG_TC_DBG", 0);
    a.b1 = L.c1.bias; a.b2 = L.c2.bias;
    a.out = out; a.o_bstride = out_b; a.o_pstride = out_p;
    a.a_lo = in_lo; a.out_lo = out_lo; a.out32 = out32; a.o32_bstride = o32_b; a.o32_pstride = o32_p;
    // the lo rows of a tile (MT * 128 per chunk) are parked in the H-tile buffer
    if (LO && (size_t)g.MT * 128 * n_chunks > (size_t)g.RH * (g.s2d ? 2 : 1) * n_chunks)
        throw StatusError(HFG_ERR_INVALID, "internal: H-tile buffer smaller than the lo rows");
    a.acc = acc; a.acc_bstride = acc_b; a.acc_pstride = acc_p; a.acc_mode = acc_mode; a.div = div;
    a.n_sum = n_sum;
    for (int s = 0; s < n_sum; ++s) a.sum_in[s] = sum_in[s];
    a.N = L.c1.cout; a.n_chunks = n_chunks; a.MT = g.MT; a.T = T;
    a.k = L.c1.k; a.dil = L.c1.dil; a.p1 = L.c1.pad; a.p2 = L.c2.pad;
    a.R1 = g.R1; a.RH = g.RH; a.TO = g.TO; a.sa = g.sa; a.sw = g.sw; a.tap_group = g.G;
    a.tiles_per_batch = (T + g.TO - 1) / g.TO;
    a.n_tiles = a.tiles_per_batch * B;
    a.len_rows = len_rows;
    a.slope = 0.1f;
    a.timeline = h->pair_timeline;
    // persistent grid: as many CTAs as are co-resident (registers, smem, TMEM columns)
    const int ctas = g.ctas;
    using KernelFn = void (*)(TcPairArgs);
    KernelFn fn = nullptr;
    if constexpr (LO) {
        static_assert(P == PREC_FP16, "the hi + lo variant runs on fp16 operand planes");
#ifdef HFG_TUNING
        if (ew12) fn = ctas == 2 ? tc_pair_kernel<P, P, 1, 2, true, 12> : tc_pair_kernel<P, P, 1, 1, true, 12>;
        else
#endif
        if (ctas == 2) fn = two ? tc_pair_kernel<P, P, 2, 2, true> : tc_pair_kernel<P, P, 1, 2, true>;
        else fn = two ? tc_pair_kernel<P, P, 2, 1, true> : tc_pair_kernel<P, P, 1, 1, true>;
    } else if (P == PREC_TF32 && P2 == PREC_FP16) {
        if (ctas == 2) fn = two ? tc_pair_kernel<PREC_TF32, PREC_FP16, 2, 2> : tc_pair_kernel<PREC_TF32, PREC_FP16, 1, 2>;
        else fn = two ? tc_pair_kernel<PREC_TF32, PREC_FP16, 2, 1> : tc_pair_kernel<PREC_TF32, PREC_FP16, 1, 1>;
    } else {
        if (ctas == 2) fn = two ? tc_pair_kernel<P, P, 2, 2> : tc_pair_kernel<P, P, 1, 2>;
        else fn = two ? tc_pair_kernel<P, P, 2, 1> : tc_pair_kernel<P, P, 1, 1>;
    }
    // co-residency: smem / TMEM columns (g.occ) and registers (64 K per SM, allocated per warp in units of 8)
    const bool ew12_used = ew12 && (LO || P != PREC_TF32);
    const int threads = pair_threads(ew12_used ? 12 : kPairEpiWarps);
#ifdef HFG_TUNING
    if constexpr (!LO && P != PREC_TF32) {
        if (ew12) fn = ctas == 2 ? tc_pair_kernel<P, P, 1, 2, false, 12> : tc_pair_kernel<P, P, 1, 1, false, 12>;
    }
#endif
    int& regs = h->pair_regs[LO ? 4 : (P == PREC_TF32 && P2 == PREC_FP16 ? 3 : P)][ew12_used ? 2 : (two ? 1 : 0)][ctas - 1];
    if (regs == 0) {
        cudaFuncAttributes fa{};
        check_cuda(cudaFuncGetAttributes(&fa, fn), "cudaFuncGetAttributes");
        regs = std::max(1, fa.numRegs);
    }
    const int occ_regs = 65536 / (((regs + 7) / 8 * 8) * threads);
    const int occ = std::max(1, std::min(occ_regs, g.occ));
    const int n_sched = (a.n_tiles + ctas - 1) / ctas;
    const int grid = ctas * std::min(n_sched, (h->sm_count / ctas) * occ);
    const double C = a.N;
    const double flops = 2.0 * 2.0 * C * C * a.k * (double)B * T;
    // algorithmic bytes: the activation in and out once, the other resblocks' outputs (or the fp32 running sum), the weights
    // (split plan: hi + lo in; hi + lo, hi only or fp32 out; fp32 sum planes)
    const double bytes = LO ? (double)B * T * C * (4.0 + (out32 ? 4.0 : (out_lo ? 4.0 : 2.0)) + 4.0 * n_sum) + 2.0 * ESZ * C * C * a.k
                             : (double)B * T * C * ESZ * ((out ? 2 : 1) + n_sum) +
                               (acc ? 4.0 * B * T * C * (acc_mode == TC_ACC_ADD ? 2 : 1) : 0.0) + 2.0 * ESZ * C * C * a.k;
    if (env_int("HFG_TC_VERBOSE", 0))
        fprintf(stderr, "[pair] ew=%d gr=%d lo=%d N=%d k=%d d=%d MT=%d G=%d sa=%d sw=%d kbc=%d ctas=%d smem=%zu occ=%d minb=%d grid=%d tiles=%d s2d=%d k2=%d G2=%d stage=%zu RH=%d\n",
                ew12_used ? 12 : 8, a.groups, (int)LO, a.N, a.k, a.dil, g.MT, g.G, g.sa, g.sw, g.kbc, ctas, g.smem, occ, two ? 2 : 1, grid, a.n_tiles, g.s2d, g.k2, g.G2, g.stage, g.RH);
    h->prof_begin(st, label, flops, bytes);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = a.pdl ? 2 : 1;
    check_cuda(cudaLaunchKernelEx(&cfg, fn, a), "tc_pair_kernel launch");
    h->prof_end(st);
    check_cuda(cudaGetLastError(), "tc_pair_kernel launch");
}

// MIXED (P = PREC_FP16): the tf32 mode on fp16 hi + lo planes (split plan, tc_tf32_mixed)
template <int P, bool MIXED = false>
static void tc_forward_impl(hfg_handle* h, const float* mel, int B, int T, float* wav, char* ws,
                            cudaStream_t st, float* const* stage_out, const int* lengths, int halo) {
    constexpr int ESZ = Prec<P>::ESZ;
    const TcPlan plan = tc_plan(h, B, T, MIXED ? HFG_MODE_TF32 : (P == PREC_BF16 ? HFG_MODE_BF16 : (P == PREC_FP16 ? HFG_MODE_FP16 : HFG_MODE_TF32)));
    if (plan.mixed != MIXED) throw StatusError(HFG_ERR_INVALID, "internal: plan / launch-path mismatch");
    const float slope = 0.1f;
    const int n_rb = h->cfg.num_resblocks;
    auto ptr = [&](const TcPlane& p) { return reinterpret_cast<uint8_t*>(ws + p.off); };
    // variable-length batch: row counts per utterance and stage, computed on the device from `lengths`
    int* len_tab = lengths ? reinterpret_cast<int*>(ws + plan.lens_off) : nullptr;
    auto lens_of = [&](int row) -> const int* { return len_tab ? len_tab + (size_t)row * B : nullptr; };

    h->stage_begin(st, "head");
    if (lengths) {
        LenGeom g{};
        g.n_stages = (int)h->ups.size();
        for (int i = 0; i < g.n_stages; ++i) { g.u[i] = h->ups[i].u; g.k[i] = h->ups[i].k; g.p[i] = h->ups[i].p; }
        h->prof_begin(st, "len_table", 0, 0);
        tc_len_table<<<(B + 127) / 128, 128, 0, st>>>(lengths, B, T, halo, g, len_tab);
        h->prof_end(st);
        check_cuda(cudaGetLastError(), "tc_len_table launch");
    }
    // ---- zero the padding rows of every plane (one launch) ----
    {
        PadJobs jobs{};
        auto flush = [&]() {
            if (!jobs.n) return;
            h->prof_begin(st, "zero_pads", 0, 0);
            tc_zero_pads<<<dim3(64, jobs.n), 256, 0, st>>>(jobs);
            h->prof_end(st);
            check_cuda(cudaGetLastError(), "tc_zero_pads launch");
            jobs.n = 0;
        };
        auto add = [&](const TcPlane& p) {
            jobs.job[jobs.n++] = PadJob{ptr(p), (long long)B * p.nchunks, p.TP, p.T};
            if (jobs.n == 40) flush();
        };
        // which planes are ever read through a convolution's halo: R/H only with more than one pair per resblock
        // (a finished resblock output that only feeds the MRF sum is read at valid rows only), S only without
        // the fused kernel
        add(plan.mel); add(plan.pre);
        for (size_t i = 0; i < h->ups.size(); ++i) {
            add(plan.st[i].X);
            // split plan: lo twins and F32 planes are read at valid rows only; the last stage's fp32 MRF output is
            // read through conv_post's halo
            if (plan.st[i].Y.bytes) add(plan.st[i].Y);
            if (plan.st[i].Y32.bytes) add(plan.st[i].Y32);
            for (int j = 0; j < n_rb; ++j) {
                if (h->mrfs[i][j].size() > 1) add(plan.st[i].rb[j].R);
                if (h->mrfs[i][j].size() > 2) add(plan.st[i].rb[j].H);
                if (!plan.st[i].rb[j].fused) add(plan.st[i].rb[j].S);
            }
        }
        flush();
    }
    // ---- mel -> chunk planes ----
    {
        dim3 grid((T + 127) / 128, plan.mel.nchunks, B);
        h->prof_begin(st, "pack_mel", 0, (double)B * h->cfg.n_mels * T * (4 + ESZ));
        tc_pack_input<P><<<grid, 128, 0, st>>>(mel, ptr(plan.mel), h->cfg.n_mels, T, plan.mel.bstride, plan.mel.pstride,
                                                  h->mel_layout);
        h->prof_end(st);
        check_cuda(cudaGetLastError(), "tc_pack_input launch");
    }
    auto dump = [&](int idx, const TcPlane& p, bool f32 = false) {
        if (!stage_out || !stage_out[idx]) return;
        dim3 grid((p.T + 127) / 128, p.nchunks, B);
        if (f32) tc_unpack_stage<PREC_TF32><<<grid, 128, 0, st>>>(ptr(p), stage_out[idx], p.C, p.T, p.bstride, p.pstride, 1.0f / slope);
        else tc_unpack_stage<P><<<grid, 128, 0, st>>>(ptr(p), stage_out[idx], p.C, p.T, p.bstride, p.pstride, 1.0f / slope);
        check_cuda(cudaGetLastError(), "tc_unpack_stage launch");
    };
    auto conv = [&](cudaStream_t st, const ConvLayer& L, const TcPlane& in, const TcPlane* out, const TcPlane* res,
                    const TcPlane* acc, int acc_mode, const char* label, const int* len_rows) {
        TcConvArgs a{};
        a.len_rows = len_rows;
        a.a = ptr(in); a.a_bstride = in.bstride; a.a_pstride = in.pstride; a.a_nchunks = in.nchunks;
        a.N = tc_pick_n(L.cout);
        a.w = reinterpret_cast<const uint8_t*>(L.tc.w[P]);
        a.w_ntile_stride = (long long)in.nchunks * L.k * a.N * 16;
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
        if (!tc_launch_up<P>(h, st, a, B, L.cout, label, flops, bytes))
            tc_launch_conv<P>(h, st, a, B, L.cout, label, flops, bytes);
    };

    // conv_pre (reference :238); its output is stored as leaky_relu(x) for ups[0] (:244)
    conv(st, h->pre, plan.mel, &plan.pre, nullptr, nullptr, TC_ACC_NONE, "conv_pre", lens_of(0));
    h->stage_end(st);
    dump(0, plan.pre);

    const TcPlane* cur = &plan.pre;
    for (size_t i = 0; i < h->ups.size(); ++i) {
        const UpLayer& U = h->ups[i];
        const auto& S = plan.st[i];
        h->stage_begin(st, ("ups" + std::to_string(i)).c_str());
        {   // x = ups[i](leaky_relu(x)) as u polyphase convolutions (reference :245)
            TcConvArgs a{};
            a.a = ptr(*cur); a.a_bstride = cur->bstride; a.a_pstride = cur->pstride; a.a_nchunks = cur->nchunks;
            // measured (bench workload, us, unstacked -> stacked): u = 2 stages tf32 71 -> 68 and 81 -> 61, bf16
            // 45 -> 41 and 38 -> 42; u = 8 stages (N tiles of 256 either way) 100 -> 104 / no change: stack where
            // the virtual width u * cout fits one N <= 128 tile
            const bool stack = U.tc_stack.ok && env_int("HFG_TC_UPS_STACK", U.u * U.cout <= 128 ? 1 : 0);
            const int cout_v = stack ? U.u * U.cout : U.cout;          // (virtual) output channels of the GEMM
            a.N = tc_pick_n(cout_v);
            a.w = reinterpret_cast<const uint8_t*>(stack ? U.tc_stack.w[P] : U.tc.w[P]);
            a.w_ntile_stride = (long long)cur->nchunks * U.taps_max * a.N * 16;
            a.w_phase_stride = stack ? 0 : a.w_ntile_stride * (U.cout / a.N);
            a.stack_cout = stack ? U.cout : 0;
            a.bias = U.bias;
            a.out = ptr(S.X); a.res = nullptr; a.o_bstride = S.X.bstride; a.o_pstride = S.X.pstride;
            if (MIXED) a.out_lo = ptr(S.Xl);
            a.acc_mode = TC_ACC_NONE; a.div = 1.f;
            a.T_out = S.X.T;
            a.n_q = (S.X.T - 1 + U.p) / U.u + 1;
            a.taps_max = U.taps_max; a.k = U.k; a.u = U.u; a.dil = -1; a.pad = 0; a.phases = stack ? 1 : U.u;
            a.out_stride = U.u; a.out_off = -U.p; a.min_off = -(U.taps_max - 1);
            a.slope = slope;
            a.len_rows = lens_of(1 + (int)i);
            const double flops = 2.0 * U.cin * U.cout * U.k * (double)B * cur->T;
            const double bytes = (double)B * ESZ * ((double)U.cin * cur->T + (double)U.cout * S.X.T) +
                                 (MIXED ? 2.0 * B * U.cout * S.X.T : 0.0) + (double)ESZ * U.cin * U.cout * U.k;
            const std::string ulab = "ups" + std::to_string(i);
            if (!tc_launch_up<P>(h, st, a, B, cout_v, ulab.c_str(), flops, bytes))
                tc_launch_conv<P>(h, st, a, B, cout_v, ulab.c_str(), flops, bytes);
        }
        h->stage_end(st);
        dump(1 + 2 * (int)i, S.X);       // (split plan: the hi half -- the stage hooks are test instrumentation)

        h->stage_begin(st, ("mrf" + std::to_string(i)).c_str());
        // MRF (reference :116-131).  The resblocks only meet in the running sum, so resblock j is enqueued on
        // stream j % n_streams: the tail of one kernel (persistent grids rarely divide evenly: 96 tile pairs
        // on 74 clusters in stage 0) is filled by the next resblock's CTAs.  Per-launch profiling serialises
        // everything on `st` so that each kernel is timed alone.
        const int n_streams = (h->profiling == 1 || n_rb < 2) ? 1 : std::max(1, std::min({env_int("HFG_TC_STREAMS", 3),
                                                                                      (int)hfg_handle::kStreams, n_rb}));
        // Stream assignment, measured at the bench workload (profiles/r1_tuning.md section 8; ms/step bf16 / tf32):
        // one stream 2.49 / 4.38; three streams, resblocks in natural order with the SHORT ones (k = 3, 7) on
        // the higher-priority side streams and the longest on the caller's stream 2.25 / 4.10; longest first
        // on the highest priority (HFG_TC_STREAM_LPT=1) 2.35 / 4.13; no priorities 2.29 / 4.16; two streams
        // 2.38 / 4.19.  The short resblocks finish their running-sum updates early and their epilogue-bound
        // kernels fill the gaps of the MMA-bound long ones.
        std::vector<int> order(n_rb), slot(n_rb, 0);
        for (int j = 0; j < n_rb; ++j) order[j] = j;
        auto cost = [&](int j) { int c = 0; for (auto& pl : h->mrfs[i][j]) c += pl.c1.k + pl.c2.k; return c; };
        if (n_streams > 1 && env_int("HFG_TC_STREAM_LPT", 0))
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost(x) > cost(y); });
        for (int r = 0; r < n_rb; ++r) slot[order[r]] = (n_streams - 1 - (r % n_streams));   // rank 0 -> last side stream
        auto stream_of = [&](int j) { return slot[j] == 0 ? st : h->side[slot[j] - 1]; };
        if (n_streams > 1) {
            h->ensure_streams();
            check_cuda(cudaEventRecord(h->ev_fork, st), "cudaEventRecord(fork)");
            for (int s = 1; s < n_streams; ++s)
                check_cuda(cudaStreamWaitEvent(h->side[s - 1], h->ev_fork, 0), "cudaStreamWaitEvent(fork)");
        }
        // pass 0: every pair but the last, costliest resblock first; pass 1: the last pairs in resblock order --
        // the pair that forms the MRF sum comes last and an event has to be recorded before the wait on it is
        // enqueued.  With sum planes (tc_stage_sum_planes) the last pairs of resblocks 0 .. n-2 are ordinary
        // pairs whose outputs the last pair of resblock n-1 adds up; otherwise the fp32 ACC plane is updated
        // WRITE -> ADD ... -> FINAL.
        const bool sum_planes = S.sum_planes;
        std::vector<const TcPlane*> src(n_rb, &S.X);
        std::vector<const uint8_t*> fin;                   // finished outputs of resblocks 0 .. n-2 (sum planes)
        for (int pass = 0; pass < 2; ++pass) {
            for (int q = 0; q < n_rb; ++q) {
                const int j = pass == 0 ? order[q] : q;
                const auto& rb = h->mrfs[i][j];
                const auto& W = S.rb[j];
                cudaStream_t sj = stream_of(j);
                const std::string lab = "mrf" + std::to_string(i) + ".k" + std::to_string(rb[0].c1.k);
                const size_t l_begin = pass == 0 ? 0 : rb.size() - 1, l_end = pass == 0 ? rb.size() - 1 : rb.size();
                for (size_t l = l_begin; l < l_end; ++l) {
                    const TcPlane* r = src[j];
                    const bool last = (l + 1 == rb.size());
                    const bool closes = last && j == n_rb - 1;          // this pair forms the MRF output
                    int mode = TC_ACC_NONE;
                    if (last && n_rb > 1) {
                        if (sum_planes) mode = closes ? TC_ACC_FINAL : TC_ACC_NONE;
                        else mode = j == 0 ? TC_ACC_WRITE : (closes ? TC_ACC_FINAL : TC_ACC_ADD);
                    }
                    if (last && n_streams > 1) {
                        if (sum_planes && closes) {
                            for (int jj = 0; jj + 1 < n_rb; ++jj)
                                if (stream_of(jj) != sj)
                                    check_cuda(cudaStreamWaitEvent(sj, h->ev_sum[jj], 0), "cudaStreamWaitEvent(sum)");
                        } else if (!sum_planes && j > 0 && stream_of(j - 1) != sj) {
                            check_cuda(cudaStreamWaitEvent(sj, h->ev_sum[j - 1], 0), "cudaStreamWaitEvent(sum)");
                        }
                    }
                    // destination of this pair: ping-pong between R and H; the closing pair writes Y; with the
                    // ACC plane the other last pairs write nothing but the running sum
                    const TcPlane* dst = (r == &W.R ? &W.H : &W.R);
                    if (closes || (last && n_rb == 1)) dst = &S.Y;
                    else if (last && !sum_planes) dst = nullptr;
                    // split plan: lo twin of a plane of this stage
                    auto twin = [&](const TcPlane* q) -> const TcPlane* {
                        return q == &S.X ? &S.Xl : (q == &W.R ? &W.Rl : (q == &W.H ? &W.Hl : nullptr));
                    };
                    (void)twin;
                    const bool use_acc = mode != TC_ACC_NONE && !sum_planes;
                    const PairGeom g = tc_pair_geometry(h, rb[l], r->nchunks, P);
                    if (MIXED) {
                        if constexpr (MIXED) {
                            // result of this pair: hi + lo where the next pair of the resblock reads it; hi only for
                            // an MRF output that an upsampler reads; fp32 where only the MRF sum (F32) or
                            // conv_post (Y32) reads it
                            const bool is_y = closes || (last && n_rb == 1);
                            const TcPlane* d32 = !last ? nullptr : (is_y ? (S.Y32.bytes ? &S.Y32 : nullptr) : &W.F32);
                            const TcPlane& g32 = S.Y32.bytes ? S.Y32 : S.rb[0].F32;     // every fp32 plane of a stage has this geometry
                            tc_launch_pair<P, true>(h, sj, rb[l], g, ptr(*r), r->bstride, r->pstride, r->nchunks,
                                                    d32 ? nullptr : ptr(*dst), S.X.bstride, S.X.pstride,
                                                    nullptr, 0, 0, mode, (float)n_rb,
                                                    (sum_planes && closes) ? fin.data() : nullptr, (sum_planes && closes) ? (int)fin.size() : 0,
                                                    lens_of(1 + (int)i), B, S.X.T, lab.c_str(),
                                                    ptr(*twin(r)), (!last) ? ptr(*twin(dst)) : nullptr, d32 ? ptr(*d32) : nullptr,
                                                    g32.bstride, g32.pstride);
                            if (last && sum_planes && !closes) fin.push_back(ptr(W.F32));
                        }
                    } else if (g.ok) {
                        tc_launch_pair<P>(h, sj, rb[l], g, ptr(*r), r->bstride, r->pstride, r->nchunks,
                                          dst ? ptr(*dst) : nullptr, S.X.bstride, S.X.pstride,
                                          use_acc ? reinterpret_cast<float*>(ptr(S.ACC)) : nullptr,
                                          S.ACC.bstride, S.ACC.pstride, mode, (float)n_rb,
                                          (sum_planes && closes) ? fin.data() : nullptr, (sum_planes && closes) ? (int)fin.size() : 0,
                                          lens_of(1 + (int)i), B, S.X.T, lab.c_str());
                    } else {
                        // unfused fallback: conv1 -> scratch, conv2 (+ residual) -> dst
                        const std::string clab = lab + ":conv";       // profile label of the unfused launches
                        conv(sj, rb[l].c1, *r, &W.S, nullptr, nullptr, TC_ACC_NONE, clab.c_str(), lens_of(1 + (int)i));
                        conv(sj, rb[l].c2, W.S, dst, r, use_acc ? &S.ACC : nullptr, use_acc ? mode : TC_ACC_NONE, clab.c_str(),
                             lens_of(1 + (int)i));
                    }
                    if (!MIXED && last && sum_planes && !closes) fin.push_back(ptr(*dst));
                    if (last && n_streams > 1 && j + 1 < n_rb)
                        check_cuda(cudaEventRecord(h->ev_sum[j], sj), "cudaEventRecord(sum)");
                    if (!last) src[j] = dst;
                }
            }
        }
        if (n_streams > 1) {
            // join: everything enqueued on the side streams precedes what follows on `st`
            for (int s = 1; s < n_streams; ++s) {
                check_cuda(cudaEventRecord(h->ev_join[s - 1], h->side[s - 1]), "cudaEventRecord(join)");
                check_cuda(cudaStreamWaitEvent(st, h->ev_join[s - 1], 0), "cudaStreamWaitEvent(join)");
            }
        }
        h->stage_end(st);
        cur = &S.Y;
        if (MIXED && S.Y32.bytes) dump(2 + 2 * (int)i, S.Y32, true); else dump(2 + 2 * (int)i, S.Y);
    }
    // wav = tanh(conv_post(leaky_relu(x)))  (reference :254-256); planes already hold leaky_relu(x)
    {
        if (MIXED) cur = &plan.st[h->ups.size() - 1].Y32;        // conv_post reads the fp32 residual stream
        const int Tw = cur->T;
        dim3 grid((Tw + kPostTile - 1) / kPostTile, B);
        const size_t post_smem = (size_t)kPostGC * post_col_cells<7>() * 16 + sizeof(float) * h->post_cin * 7;
        h->stage_begin(st, "tail");
        h->prof_begin(st, "conv_post", 2.0 * h->post_cin * 7 * (double)B * Tw,
                      (double)B * Tw * ((MIXED ? 4 : ESZ) * h->post_cin + 4.0));
        if (MIXED)
            tc_conv_post_tanh<PREC_TF32, 7><<<grid, kPostThreads, post_smem, st>>>(
                ptr(*cur), h->post_w, h->post_b, wav, h->post_cin, Tw, 3, cur->bstride, cur->pstride,
                lens_of((int)h->ups.size()), lens_of(1 + (int)h->ups.size()));
        else
            tc_conv_post_tanh<P, 7><<<grid, kPostThreads, post_smem, st>>>(
                ptr(*cur), h->post_w, h->post_b, wav, h->post_cin, Tw, 3, cur->bstride, cur->pstride,
                lens_of((int)h->ups.size()), lens_of(1 + (int)h->ups.size()));
        h->prof_end(st);
        h->stage_end(st);
        check_cuda(cudaGetLastError(), "tc_conv_post_tanh launch");
    }
}

// Micro-benchmark of ONE MRF convolution launch on scratch planes (tuning / ncu).
// which: 0 = convs1[pair], 1 = convs2[pair] (with residual).  Returns avg ms per launch.
// LO (which = 2 only): the hi + lo variant of the fused pair (split plan of the tf32 mode), on fp16 planes
template <int P, bool LO = false>
static float tc_bench_layer_impl(hfg_handle* h, int stage, int resblock, int pair, int which, int B, int T, int iters) {
    const PairLayers& PL = h->mrfs.at(stage).at(resblock).at(pair);
    const ConvLayer& L = which == 1 ? PL.c2 : PL.c1;
    size_t cur = 0;
    const int cw = Prec<P>::CW;
    TcPlane in = tc_plane(B, L.cin, T, cw, cur), out = tc_plane(B, L.cout, T, cw, cur), res = tc_plane(B, L.cout, T, cw, cur);
    TcPlane in_lo = tc_plane(B, L.cin, T, cw, cur);        // LO: lo twin of `in`; `res` serves as the lo twin of `out`
    if (LO && which != 2) throw StatusError(HFG_ERR_INVALID, "hfg_bench_layer: the tf32 split plan only has fused pairs");
    char* ws = nullptr;
    check_cuda(cudaMalloc((void**)&ws, cur), "cudaMalloc(bench)");
    check_cuda(cudaMemset(ws, 0, cur), "cudaMemset(bench)");
    cudaStream_t st = 0;
    TcConvArgs a{};
    a.a = (uint8_t*)ws + in.off; a.a_bstride = in.bstride; a.a_pstride = in.pstride; a.a_nchunks = in.nchunks;
    a.N = tc_pick_n(L.cout);
    a.w = reinterpret_cast<const uint8_t*>(L.tc.w[P]);
    a.w_ntile_stride = (long long)in.nchunks * L.k * a.N * 16;
    a.bias = L.bias;
    a.out = (uint8_t*)ws + out.off; a.res = which ? (uint8_t*)ws + res.off : nullptr;
    a.o_bstride = out.bstride; a.o_pstride = out.pstride;
    a.acc_mode = TC_ACC_NONE; a.div = 1.f;
    a.n_q = T; a.T_out = T; a.taps_max = L.k; a.k = L.k; a.u = 1; a.dil = L.dil; a.pad = L.pad; a.phases = 1;
    a.out_stride = 1; a.out_off = 0; a.min_off = -L.pad; a.slope = 0.1f;
    const int was = h->profiling;
    h->profiling = 0;
    cudaEvent_t e0, e1;
    check_cuda(cudaEventCreate(&e0), "event"); check_cuda(cudaEventCreate(&e1), "event");
    const PairGeom g = tc_pair_geometry(h, PL, in.nchunks, P);
    if (which == 2 && !g.ok) throw StatusError(HFG_ERR_UNSUPPORTED, "fused pair does not fit for this layer");
    auto launch = [&]() {
        if (which == 2)
            tc_launch_pair<P, LO>(h, st, PL, g, (uint8_t*)ws + in.off, in.bstride, in.pstride, in.nchunks,
                                  (uint8_t*)ws + out.off, out.bstride, out.pstride, nullptr, 0, 0, TC_ACC_NONE, 1.f,
                                  nullptr, 0, nullptr, B, T, "bench",
                                  LO ? (uint8_t*)ws + in_lo.off : nullptr, LO ? (uint8_t*)ws + res.off : nullptr);
        else
            tc_launch_conv<P>(h, st, a, B, L.cout, "bench", 0, 0);
    };
    for (int i = 0; i < 3; ++i) launch();
#ifdef HFG_TUNING
    const char* tl_path_env = getenv("HFG_TC_TIMELINE");
#else
    const char* tl_path_env = nullptr;
#endif
    if (const char* tl_path = tl_path_env) {
        // tuning only: one extra launch with per-phase clock stamps of the first 4 CTAs, dumped as text
        const size_t n = 4 * 16 * 16;
        unsigned long long* d = nullptr;
        check_cuda(cudaMalloc((void**)&d, n * 8), "cudaMalloc(timeline)");
        check_cuda(cudaMemset(d, 0, n * 8), "cudaMemset(timeline)");
        h->pair_timeline = d;
        launch();
        h->pair_timeline = nullptr;
        std::vector<unsigned long long> hbuf(n);
        check_cuda(cudaMemcpy(hbuf.data(), d, n * 8, cudaMemcpyDeviceToHost), "cudaMemcpy(timeline)");
        cudaFree(d);
        if (FILE* f = fopen(tl_path, "a")) {
            fprintf(f, "# stage=%d resblock=%d pair=%d N=%d k=%d d=%d MT=%d G=%d sa=%d sw=%d ctas=%d\n", stage, resblock, pair,
                    L.cout, L.k, PL.c1.dil, g.MT, g.G, g.sa, g.sw, g.ctas);
            for (size_t i = 0; i < n; i += 16) {
                for (int e = 0; e < 16; ++e) fprintf(f, "%llu ", hbuf[i + e]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    check_cuda(cudaEventRecord(e0, st), "record");
    for (int i = 0; i < iters; ++i) launch();
    check_cuda(cudaEventRecord(e1, st), "record");
    check_cuda(cudaEventSynchronize(e1), "sync");
    float ms = 0.f;
    check_cuda(cudaEventElapsedTime(&ms, e0, e1), "elapsed");
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(ws);
    h->profiling = was;
    return ms / iters;
}

inline void tc_forward(hfg_handle* h, const float* mel, int B, int T, float* wav, char* ws, int mode,
                       cudaStream_t st, float* const* stage_out, const int* lengths = nullptr, int halo = 0) {
    if (h->cc_major != 10)
        throw StatusError(HFG_ERR_UNSUPPORTED, "tensor-core modes need an sm_100 device (tcgen05)");
    if (mode == HFG_MODE_BF16) tc_forward_impl<PREC_BF16>(h, mel, B, T, wav, ws, st, stage_out, lengths, halo);
    else if (mode == HFG_MODE_FP16) tc_forward_impl<PREC_FP16>(h, mel, B, T, wav, ws, st, stage_out, lengths, halo);
    else if (tc_tf32_mixed(h)) tc_forward_impl<PREC_FP16, true>(h, mel, B, T, wav, ws, st, stage_out, lengths, halo);
    else tc_forward_impl<PREC_TF32>(h, mel, B, T, wav, ws, st, stage_out, lengths, halo);
}

inline void configure_kernels(hfg_handle*) {
    const int smem = 100 * 1024;
    check_cuda(cudaFuncSetAttribute(conv_tile_fp32<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "attr");
    check_cuda(cudaFuncSetAttribute(conv_tile_fp32<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "attr");
    check_cuda(cudaFuncSetAttribute(conv_tile_fp32<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "attr");
    auto big = [](auto fn) {
        check_cuda(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "attr");
    };
    auto each = [&](auto prec) {
        constexpr int P = decltype(prec)::value;
        big(tc_conv_kernel<P>); big(tc_up_kernel<P>);
        big(tc_pair_kernel<P, P, 1, 1>); big(tc_pair_kernel<P, P, 2, 1>);
        big(tc_pair_kernel<P, P, 1, 2>); big(tc_pair_kernel<P, P, 2, 2>);
    };
    big(tc_pair_kernel<PREC_TF32, PREC_FP16, 1, 1>); big(tc_pair_kernel<PREC_TF32, PREC_FP16, 2, 1>);
    big(tc_pair_kernel<PREC_TF32, PREC_FP16, 1, 2>); big(tc_pair_kernel<PREC_TF32, PREC_FP16, 2, 2>);
#ifdef HFG_TUNING
    // twelve epilogue warps (2-byte modes and the split plan, one CTA per SM): tuning build only
    big(tc_pair_kernel<PREC_BF16, PREC_BF16, 1, 1, false, 12>); big(tc_pair_kernel<PREC_BF16, PREC_BF16, 1, 2, false, 12>);
    big(tc_pair_kernel<PREC_FP16, PREC_FP16, 1, 1, false, 12>); big(tc_pair_kernel<PREC_FP16, PREC_FP16, 1, 2, false, 12>);
    big(tc_pair_kernel<PREC_FP16, PREC_FP16, 1, 1, true, 12>); big(tc_pair_kernel<PREC_FP16, PREC_FP16, 1, 2, true, 12>);
#endif
    // hi + lo variants (split plan of the tf32 mode)
    big(tc_pair_kernel<PREC_FP16, PREC_FP16, 1, 1, true>); big(tc_pair_kernel<PREC_FP16, PREC_FP16, 2, 1, true>);
    big(tc_pair_kernel<PREC_FP16, PREC_FP16, 1, 2, true>); big(tc_pair_kernel<PREC_FP16, PREC_FP16, 2, 2, true>);
    each(std::integral_constant<int, PREC_TF32>{});
    each(std::integral_constant<int, PREC_BF16>{});
    each(std::integral_constant<int, PREC_FP16>{});
}

}  // namespace hfg
