// tcgen05 implicit-GEMM convolution for the HiFi-GAN generator (sm_100a).
//
// Data layout ("chunk planes").  An activation tensor with C channels and T time
// steps is stored as  act[b][c / CW][PADL + t][c % CW]  where one (row, chunk)
// cell is 16 bytes: CW = 8 channels in bf16 mode, 4 channels in tf32 mode (fp32
// storage).  Each plane has TP rows; rows [0, PADL) and [PADL + T, TP) are kept
// zero, which realises every layer's own zero padding (reference Conv1d
// padding=..., models/hifigan.py:52-69) without any bounds logic in the loads.
//
// Why this layout: a [rows x 16 B] plane segment is exactly one column of UMMA
// K-major *no-swizzle* core matrices (8 rows x 16 B, rows 16 B apart, SBO = 128 B).
// A tile of R rows x 8 chunks therefore lands in shared memory with 8 plain 1-D
// bulk copies (cp.async.bulk, no tensor map), and -- because row r sits at byte
// r*16 of its chunk column -- the k taps of the convolution are the SAME smem
// tile addressed with the descriptor start address shifted by tap*dilation rows.
// No im2col, no re-load per tap.
//
//   D[128 x N] (TMEM, fp32) += A[128 x 16|8] (smem, shifted rows) * W_tap[N x 16|8]^T (smem)
//
//   M = time rows, N = output channels (<= 256 per CTA), K = input channels per tap.
//
// Warp roles (192 threads): warp 0 = bulk-copy producer, warp 1 = TMEM allocator +
// single-thread MMA issuer, warps 2..5 = epilogue (one TMEM lane = one time row per
// thread).  Rings: A stages (one per 8-chunk K block) and W stages (one per
// (K block, tap)), full/empty mbarriers; tcgen05.commit releases stages.
//
// The epilogue fuses bias, the residual add (x recovered from the stored
// leaky_relu(x) by the exact inverse), the MRF running sum / division, the NEXT
// layer's leaky_relu (every conv output in this network is consumed through
// leaky_relu(.,0.1): reference models/hifigan.py:81,83,244,254) and the
// conversion to the operand dtype.  ConvTranspose1d runs as `phases` = u
// polyphase convolutions whose outputs interleave with stride u.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hfg.h"

namespace hfg {

constexpr int kTcEpiWarps = 8;                       // two warps per TMEM lane quarter, splitting the column steps
constexpr int kTcThreads = 64 + 32 * kTcEpiWarps;
constexpr int kPadL = 32;          // zero rows in front of t = 0 in every plane
constexpr int kMaxSA = 2, kMaxSW = 8;

enum : int { TC_ACC_NONE = 0, TC_ACC_WRITE = 1, TC_ACC_ADD = 2, TC_ACC_FINAL = 3 };

// Operand precision of a tensor-core mode (template parameter P of every kernel below):
//   PREC_TF32  fp32 storage, tcgen05 kind::tf32             (4 channels per 16-byte cell)
//   PREC_BF16  bf16 storage, tcgen05 kind::f16, bf16 x bf16 (8 channels per cell)
//   PREC_FP16  fp16 storage, tcgen05 kind::f16, fp16 x fp16 (8 channels per cell): the 10-bit mantissa of
//              tf32 at the MMA rate and operand bytes of bf16; conversions saturate to +-65504
// Accumulation is fp32 in TMEM in every mode.
enum : int { PREC_TF32 = 0, PREC_BF16 = 1, PREC_FP16 = 2 };
template <int P> struct Prec {
    static constexpr int CW = P == PREC_TF32 ? 4 : 8;     // channels per 16-byte cell
    static constexpr int ESZ = P == PREC_TF32 ? 4 : 2;    // bytes per stored element
};

// Tuning instrumentation (clock64 timelines, "switch parts of the kernel off" experiments, HFG_TC_* environment
// knobs) exists only in builds with -DHFG_TUNING (lib/libhfg_b200_tuning.so); the production library has none
// of it, so a stray environment variable cannot change kernel selection or results.
#ifdef HFG_TUNING
#define HFG_DBG(a, bit) ((a).dbg & (bit))
#else
#define HFG_DBG(a, bit) 0
#endif

struct TcConvArgs {
    // A operand: plane(b, chunk) = a + b*a_bstride + chunk*a_pstride; row r at +16 r
    const uint8_t* a; long long a_bstride, a_pstride; int a_nchunks;
    // packed weights, tight: block(kb, tap) = [nck(kb) chunks][N][16 B]; taps of a K block are contiguous:
    //   offset(kb, tap) = (kb*taps_max*8 + tap*nck(kb)) * N * 16
    const uint8_t* w; long long w_phase_stride, w_ntile_stride;
    const float* bias;
    // output / residual planes (operand dtype, chunk layout, same geometry)
    uint8_t* out; const uint8_t* res; long long o_bstride, o_pstride;
    // "lo" twin of `out` (same geometry), or null: tf32 mode on fp16 operand planes (tc_path.cuh, split plan) -- the
    // value is stored as the fp16 pair hi = fp16(v) (plane `out`, what the MMAs read) and lo = fp16(v - hi)
    uint8_t* out_lo;
    // MRF accumulator planes: fp32, 4 channels per 16-byte cell
    float* acc; long long acc_bstride, acc_pstride;   // in bytes
    int acc_mode; float inv_scale;                    // TC_ACC_FINAL: v = (acc + v) / n_resblocks
    float div;
    int N;              // output channels handled by this CTA (multiple of 16, <= 256)
    int MT;             // 128-row sub-tiles per CTA (MT * N <= 512 TMEM columns)
    int n_q;            // number of q positions
    int T_out;          // valid output rows
    int taps_max, k, u, dil, pad, phases;
    int out_stride, out_off;   // t = q*out_stride + phase + out_off
    int stack_cout;     // > 0: polyphase tap sets stacked along N -- virtual channel v = phase * stack_cout + channel
    int min_off;        // smallest input row offset over taps
    int R;              // rows per A stage = MT*128 + span
    int sa, sw;         // ring depths
    int tap_group;      // taps per W stage (small layers: fewer, fatter stages)
    int tiles_per_batch;
    float slope;        // leaky_relu slope fused on the OUTPUT (and inverted on the residual)
    const int* len_rows;            // variable-length batches: output rows utterance b needs, or null (see tc_len_nq)
    unsigned long long* timeline;   // tuning only: clock64 stamps of the first 64 CTAs along grid.y, [cta][8 events]
};
#ifndef HFG_TUNING
#define HFG_CONV_TL(ev) do { } while (0)
#else
#define HFG_CONV_TL(ev)                                                                              \
    do {                                                                                             \
        if (a.timeline && blockIdx.x == 0 && blockIdx.y < 64 && lane == 0)                            \
            a.timeline[(size_t)blockIdx.y * 8 + (ev)] = (unsigned long long)clock64();                \
    } while (0)
#endif

// q positions (input rows / GEMM rows) needed for `len_out` output rows: t = q * out_stride + phase + out_off
__device__ __forceinline__ int tc_len_nq(const TcConvArgs& a, int len_out) {
    const int nq = (len_out - 1 - a.out_off) / a.out_stride + 1;
    return nq < a.n_q ? nq : a.n_q;
}

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a pipeline bug must trap (-> CUDA error on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    long long t0 = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && (++spins & 1023u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) __trap();     // ~2 s at 2 GHz
        }
    } while (!ok);
}
// Wait of the EPILOGUE warps: between polls the warp sleeps (doubling up to `max_ns`), so that eight to
// sixteen waiting warps do not spend the issue slots the working warps of the co-resident CTA need.  A
// committed ncu source view of the narrow k = 3 pair kernel showed ~40 % of all executed instructions inside
// these poll loops (profiles/r2_tuning.md).  max_ns = 0: plain spin (mbar_wait).
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t max_ns) {
    uint32_t ok, ns = 32, spins = 0;
    long long t0 = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (max_ns) {
            __nanosleep(ns);
            ns = ns < max_ns ? ns * 2 : max_ns;
        }
        if ((++spins & 1023u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) __trap();
        }
    } while (true);
}
// non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// CTA-pair commit: arrives on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit2(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
        ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the mbarrier at the same smem offset in CTA `cta` of the cluster.  RELAXED on purpose: these
// barriers only carry the event "my stage has landed / my H tile is written" to the leader, which then
// ISSUES the pair MMA; the data itself is read by the tensor core of the SM that owns it (its own smem /
// TMEM), never by the leader's threads, so no cross-CTA memory ordering is needed.  The default
// release.cluster / acquire.cluster forms compile to MEMBAR.ALL.GPU and CCTL.IVALL per stage and made the
// CTA-pair kernel 1.5x slower than the single-CTA one (profiles/r1_tuning.md).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 r;\n\tmapa.shared::cluster.u32 r, %0, %1;\n\t"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [r];\n\t}"
        ::"r"(bar), "r"(cta) : "memory");
}
// wait for an arrival that came from the peer CTA (relaxed, see above)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    long long t0 = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && (++spins & 1023u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) __trap();
        }
    } while (!ok);
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE ("interleaved" core matrices):
//   bits [0,14) start>>4, [16,30) LBO>>4 (stride between K-adjacent core matrices),
//   [32,46) SBO>>4 (stride between 8-row groups), [46,48) version = 1, [61,64) layout = 0
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// UMMA instruction descriptor: fp32 accumulate, K-major A and B, M = 128
template <int P>
__device__ __forceinline__ uint32_t umma_idesc(int N, int M = 128) {
    const uint32_t fmt = P == PREC_BF16 ? 1u : (P == PREC_FP16 ? 0u : 2u);   // kind::f16: 0 = F16, 1 = BF16; kind::tf32: 2 = TF32
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// cta_group::2: one instruction drives the tensor cores of both SMs of a CTA pair (M = 256: each CTA
// supplies its own 128 rows of A and HALF of the N rows of B from the same smem offsets)
template <int P>
__device__ __forceinline__ void umma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (P != PREC_TF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
template <int P>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (P != PREC_TF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
// One (tap, sub-tile): `ksteps` K-steps of UMMA_K (two 16-byte cells each), A and B descriptors
// advanced by 2 cells per step.  The common 4-step case is straight-line code so the
// UTCHMMAs issue back to back from uniform registers.
template <int P, int CTAS = 1>
__device__ __forceinline__ void umma_any(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    if constexpr (CTAS == 2) umma2<P>(d, ad, bd, idesc, acc);
    else umma<P>(d, ad, bd, idesc, acc);
}
template <int P, int CTAS = 1>
__device__ __forceinline__ void umma_ksteps(uint32_t d_tmem, uint32_t hi, uint32_t a_lo, uint32_t b_lo,
                                            uint32_t a_step, uint32_t b_step, uint32_t idesc, int ksteps,
                                            uint32_t acc_first) {
    if (ksteps == 4) {
        umma_any<P, CTAS>(d_tmem, ((uint64_t)hi << 32) | a_lo, ((uint64_t)hi << 32) | b_lo, idesc, acc_first);
        umma_any<P, CTAS>(d_tmem, ((uint64_t)hi << 32) | (a_lo + a_step), ((uint64_t)hi << 32) | (b_lo + b_step), idesc, 1u);
        umma_any<P, CTAS>(d_tmem, ((uint64_t)hi << 32) | (a_lo + 2 * a_step), ((uint64_t)hi << 32) | (b_lo + 2 * b_step), idesc, 1u);
        umma_any<P, CTAS>(d_tmem, ((uint64_t)hi << 32) | (a_lo + 3 * a_step), ((uint64_t)hi << 32) | (b_lo + 3 * b_step), idesc, 1u);
    } else {
        for (int s = 0; s < ksteps; ++s) {
            umma_any<P, CTAS>(d_tmem, ((uint64_t)hi << 32) | a_lo, ((uint64_t)hi << 32) | b_lo, idesc, acc_first | (uint32_t)s);
            a_lo += a_step;
            b_lo += b_step;
        }
    }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
          "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
          "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
          "r"(__float_as_uint(v[15])) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// leaky_relu with 0 < s < 1 is max(v, s v); its inverse (1/s > 1) is min(v, v/s): two instructions each
__device__ __forceinline__ float lrelu(float v, float s) { return fmaxf(v, v * s); }
__device__ __forceinline__ float lrelu_inv(float v, float inv_s) { return fminf(v, v * inv_s); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack_bf16(uint32_t u, float& lo, float& hi) {
    lo = __uint_as_float(u << 16);
    hi = __uint_as_float(u & 0xFFFF0000u);
}
// fp16 pair, round-to-nearest-even, saturating to +-65504 instead of overflowing to inf
__device__ __forceinline__ uint32_t pack_fp16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void unpack_fp16(uint32_t u, float& lo, float& hi) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u));
    lo = f.x;
    hi = f.y;
}
// two consecutive channels <-> one 32-bit word of a 2-byte-element cell
template <int P>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    if constexpr (P == PREC_FP16) return pack_fp16(lo, hi);
    else return pack_bf16(lo, hi);
}
template <int P>
__device__ __forceinline__ void unpack2(uint32_t u, float& lo, float& hi) {
    if constexpr (P == PREC_FP16) unpack_fp16(u, lo, hi);
    else unpack_bf16(u, lo, hi);
}
// one 16-byte cell <-> its Prec<P>::CW channels
template <int P>
__device__ __forceinline__ void cell_to_floats(const uint4& u, float* v) {
    if constexpr (P == PREC_TF32) {
        v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
    } else {
        unpack2<P>(u.x, v[0], v[1]); unpack2<P>(u.y, v[2], v[3]);
        unpack2<P>(u.z, v[4], v[5]); unpack2<P>(u.w, v[6], v[7]);
    }
}
template <int P>
__device__ __forceinline__ uint4 floats_to_cell(const float* v) {
    if constexpr (P == PREC_TF32)
        return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    else
        return make_uint4(pack2<P>(v[0], v[1]), pack2<P>(v[2], v[3]), pack2<P>(v[4], v[5]), pack2<P>(v[6], v[7]));
}

// 16 consecutive channels of one row <-> 16-byte cells of consecutive chunk planes
template <int P>
__device__ __forceinline__ void store_cells16(uint8_t* p, long long plane_stride, const float (&v)[16]) {
    if constexpr (P != PREC_TF32) {
#pragma unroll
        for (int g = 0; g < 2; ++g) *reinterpret_cast<uint4*>(p + g * plane_stride) = floats_to_cell<P>(v + g * 8);
    } else {
#pragma unroll
        for (int g = 0; g < 4; ++g)
            *reinterpret_cast<float4*>(p + g * plane_stride) =
                make_float4(v[g * 4 + 0], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
    }
}
template <int P>
__device__ __forceinline__ void load_cells16(const uint8_t* p, long long plane_stride, float (&v)[16]) {
    if constexpr (P != PREC_TF32) {
        uint4 u[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) u[g] = *reinterpret_cast<const uint4*>(p + g * plane_stride);
#pragma unroll
        for (int g = 0; g < 2; ++g) cell_to_floats<P>(u[g], v + g * 8);
    } else {
        float4 u[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) u[g] = *reinterpret_cast<const float4*>(p + g * plane_stride);
#pragma unroll
        for (int g = 0; g < 4; ++g) { v[g * 4 + 0] = u[g].x; v[g * 4 + 1] = u[g].y; v[g * 4 + 2] = u[g].z; v[g * 4 + 3] = u[g].w; }
    }
}
// fp32 value as an fp16 pair: hi = fp16(v), lo = fp16(v - hi).  hi + lo carries 22 bits of mantissa (absolute floor
// 3e-8 in fp16's subnormal range); hi alone is the round-to-nearest 10-bit-mantissa operand a tf32 MMA would see.
// Both conversions saturate, so the pair stays finite whatever v is.
// The fp16 halves enter the fp32 adds directly (add / sub.rn.f32.f16 -> one FHADD each, exact), which saves the
// two conversions per pair a plain unpack would cost -- these epilogues are bound by their instruction count.
__device__ __forceinline__ void split16(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack_fp16(v[2 * i], v[2 * i + 1]);
        float d0, d1;                                  // hi - v (exact in fp32); lo = fp16(-(hi - v))
        asm("{\n\t.reg .b16 h0, h1;\n\tmov.b32 {h0, h1}, %2;\n\t"
            "sub.rn.f32.f16 %0, h0, %3;\n\tsub.rn.f32.f16 %1, h1, %4;\n\t}"
            : "=f"(d0), "=f"(d1) : "r"(h[i]), "f"(v[2 * i]), "f"(v[2 * i + 1]));
        l[i] = pack_fp16(-d0, -d1);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// hi + lo of one cell (8 channels) as fp32
__device__ __forceinline__ void join16(const uint4& hi, const uint4& lo, float* v) {
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float l0, l1;
        unpack_fp16(l[i], l0, l1);
        asm("{\n\t.reg .b16 h0, h1;\n\tmov.b32 {h0, h1}, %2;\n\t"
            "add.rn.f32.f16 %0, h0, %3;\n\tadd.rn.f32.f16 %1, h1, %4;\n\t}"
            : "=f"(v[2 * i]), "=f"(v[2 * i + 1]) : "r"(h[i]), "f"(l0), "f"(l1));
    }
}
// 16 consecutive channels of one row from two cells of the hi planes and two of the lo planes
__device__ __forceinline__ void load_split16(const uint8_t* hi_p, long long hi_stride, const uint8_t* lo_p, long long lo_stride,
                                             float (&v)[16]) {
    uint4 h[2], l[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        h[g] = *reinterpret_cast<const uint4*>(hi_p + g * hi_stride);
        l[g] = *reinterpret_cast<const uint4*>(lo_p + g * lo_stride);
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) join16(h[g], l[g], v + g * 8);
}
// 16 consecutive channels of one row -> two cells of the hi planes and two of the lo planes
__device__ __forceinline__ void store_split16(uint8_t* hi_p, uint8_t* lo_p, long long plane_stride, const float (&v)[16]) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        uint4 hi, lo;
        split16(v + g * 8, hi, lo);
        *reinterpret_cast<uint4*>(hi_p + g * plane_stride) = hi;
        *reinterpret_cast<uint4*>(lo_p + g * plane_stride) = lo;
    }
}
// two 16-byte cells of consecutive rows of one chunk plane (32-byte aligned): one 256-bit store (STG.256)
__device__ __forceinline__ void st_global_256(uint8_t* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
// the same cell of two consecutive rows: one 32-byte store when both rows are valid (p = row 0, 32-byte aligned)
__device__ __forceinline__ void store_cell_rows2(uint8_t* p, const uint4& c0, const uint4& c1, bool v0, bool v1) {
    if (v0 && v1) st_global_256(p, c0, c1);
    else if (v0) *reinterpret_cast<uint4*>(p) = c0;
    else if (v1) *reinterpret_cast<uint4*>(p + 16) = c1;
}
__device__ __forceinline__ void load_f32x16(const uint8_t* p, long long plane_stride, float (&v)[16]) {
    float4 u[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) u[g] = *reinterpret_cast<const float4*>(p + g * plane_stride);
#pragma unroll
    for (int g = 0; g < 4; ++g) { v[g * 4 + 0] = u[g].x; v[g * 4 + 1] = u[g].y; v[g * 4 + 2] = u[g].z; v[g * 4 + 3] = u[g].w; }
}
__device__ __forceinline__ void store_f32x16(uint8_t* p, long long plane_stride, const float (&v)[16]) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
        *reinterpret_cast<float4*>(p + g * plane_stride) = make_float4(v[g * 4 + 0], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
}
__device__ __forceinline__ void add_bias16(float (&v)[16], const float* sb) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float4 b = *reinterpret_cast<const float4*>(sb + 4 * g);
        v[g * 4 + 0] += b.x; v[g * 4 + 1] += b.y; v[g * 4 + 2] += b.z; v[g * 4 + 3] += b.w;
    }
}

// ------------------------------------------------------------------ the kernel
template <int P>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_conv_kernel(const TcConvArgs a) {
    extern __shared__ __align__(128) uint8_t tc_smem[];
    uint8_t* smem = tc_smem;
    constexpr int CW = Prec<P>::CW;       // channels per 16-byte cell

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (a.len_rows) {                                      // tile beyond its utterance's length: nothing to do
        const int tb = blockIdx.y / a.tiles_per_batch;
        if ((int)(blockIdx.y % a.tiles_per_batch) * a.MT * 128 >= tc_len_nq(a, a.len_rows[tb])) return;
    }
    if (warp == 0) HFG_CONV_TL(0);
    const int N = a.N, MT = a.MT, R = a.R;
    const int nck_max = a.a_nchunks < 8 ? a.a_nchunks : 8;
    const uint32_t a_stage_bytes = (uint32_t)R * nck_max * 16;
    const int G = a.tap_group;
    const uint32_t w_stage_bytes = (uint32_t)G * N * nck_max * 16;
    uint8_t* sA = smem;
    uint8_t* sW = sA + (size_t)a.sa * a_stage_bytes;
    float* sBias = reinterpret_cast<float*>(sW + (size_t)a.sw * w_stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + N);
    // barrier slots: a_full[2] a_empty[2] w_full[8] w_empty[8] acc_full[1]
    const uint32_t bar0 = smem_u32(bars);
    auto A_FULL = [&](int i) { return bar0 + 8u * i; };
    auto A_EMPTY = [&](int i) { return bar0 + 8u * (kMaxSA + i); };
    auto W_FULL = [&](int i) { return bar0 + 8u * (2 * kMaxSA + i); };
    auto W_EMPTY = [&](int i) { return bar0 + 8u * (2 * kMaxSA + kMaxSW + i); };
    const uint32_t ACC_FULL = bar0 + 8u * (2 * kMaxSA + 2 * kMaxSW);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxSA + 2 * kMaxSW + 1);

    // tile coordinates
    // grid.x = (phase, channel tile) runs fastest: the CTAs that share an activation tile (and, for the
    // polyphase upsamplers, interleave their 16-byte cells in the same output sectors) are co-scheduled
    const int b = blockIdx.y / a.tiles_per_batch;
    const int q0 = (blockIdx.y % a.tiles_per_batch) * MT * 128;
    const int n_tiles = gridDim.x / a.phases;
    const int phase = blockIdx.x / n_tiles;
    const int ntile = blockIdx.x % n_tiles;
    int taps = a.taps_max;
    if (a.phases > 1) taps = (a.k - phase + a.u - 1) / a.u;
    const int n_chunks = a.a_nchunks;
    const int n_kb = (n_chunks + 7) / 8;

    uint32_t ncols = 32;
    while ((int)ncols < MT * N) ncols <<= 1;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < a.sa; ++i) { mbar_init(A_FULL(i), 1); mbar_init(A_EMPTY(i), 1); }
        for (int i = 0; i < a.sw; ++i) { mbar_init(W_FULL(i), 1); mbar_init(W_EMPTY(i), 1); }
        mbar_init(ACC_FULL, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < N; i += (int)blockDim.x - 64)
            sBias[i] = a.bias[a.stack_cout ? (ntile * N + i) % a.stack_cout : ntile * N + i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) HFG_CONV_TL(1);

    // Roles are warp-uniform: the whole warp runs the loop nest (so loop counters and
    // descriptors stay in uniform registers) and one elected lane issues the async ops.
    if (warp == 0) {
        // ===================== producer: bulk copies =====================
        const bool leader = elect_one();
        const uint8_t* ab = a.a + (long long)b * a.a_bstride + (long long)(kPadL + q0 + a.min_off) * 16;
        const uint8_t* wb = a.w + (long long)phase * a.w_phase_stride + (long long)ntile * a.w_ntile_stride;
        int sa_i = 0, sa_ph = 0, sw_i = 0, sw_ph = 0;
        auto issue_a = [&](int kb) {
            const int nck = (n_chunks - 8 * kb) < 8 ? (n_chunks - 8 * kb) : 8;
            mbar_wait(A_EMPTY(sa_i), sa_ph ^ 1);
            if (leader) {
                mbar_expect_tx(A_FULL(sa_i), (uint32_t)nck * R * 16);
                const uint32_t dst = smem_u32(sA + (size_t)sa_i * a_stage_bytes);
                for (int c = 0; c < nck; ++c)
                    bulk_g2s(dst + (uint32_t)c * R * 16, ab + (long long)(8 * kb + c) * a.a_pstride,
                             (uint32_t)R * 16, A_FULL(sa_i));
            }
            __syncwarp();
            if (++sa_i == a.sa) { sa_i = 0; sa_ph ^= 1; }
        };
        issue_a(0);
        for (int kb = 0; kb < n_kb; ++kb) {
            const int nck = (n_chunks - 8 * kb) < 8 ? (n_chunks - 8 * kb) : 8;
            for (int tap0 = 0; tap0 < taps; tap0 += G) {
                const int g = (taps - tap0) < G ? (taps - tap0) : G;
                mbar_wait(W_EMPTY(sw_i), sw_ph ^ 1);
                if (leader) {
                    mbar_expect_tx(W_FULL(sw_i), (uint32_t)g * nck * N * 16);
                    bulk_g2s(smem_u32(sW + (size_t)sw_i * w_stage_bytes),
                             wb + ((long long)kb * a.taps_max * 8 + (long long)tap0 * nck) * N * 16,
                             (uint32_t)g * nck * N * 16, W_FULL(sw_i));
                }
                __syncwarp();
                if (++sw_i == a.sw) { sw_i = 0; sw_ph ^= 1; }
                if (tap0 == 0 && kb + 1 < n_kb) issue_a(kb + 1);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one elected lane) =====================
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc<P>(N);
        // descriptor high words are loop-invariant; the low word is start>>4 | LBO>>4 << 16,
        // so stepping K chunks / sub-tiles / taps is an add of (bytes >> 4) = rows on the low word
        const uint32_t a_hi = (128u >> 4) | (1u << 14);                     // SBO = 128 B, version 1 (A and B)
        const uint32_t a_lbo = ((uint32_t)R) << 16, b_lbo = ((uint32_t)N) << 16;
        int sa_i = 0, sa_ph = 0, sw_i = 0, sw_ph = 0;
        uint32_t acc_on = 0;
        for (int kb = 0; kb < n_kb; ++kb) {
            const int nck = (n_chunks - 8 * kb) < 8 ? (n_chunks - 8 * kb) : 8;
            const int ksteps = nck >> 1;
            mbar_wait(A_FULL(sa_i), sa_ph);
            tc_fence_after();
            if (kb == 0) HFG_CONV_TL(2);
            const uint32_t a_lo0 = ((smem_u32(sA + (size_t)sa_i * a_stage_bytes) & 0x3FFFFu) >> 4) | a_lbo;
            for (int tap0 = 0; tap0 < taps; tap0 += G) {
                const int g = (taps - tap0) < G ? (taps - tap0) : G;
                mbar_wait(W_FULL(sw_i), sw_ph);
                tc_fence_after();
                const uint32_t b_stage = ((smem_u32(sW + (size_t)sw_i * w_stage_bytes) & 0x3FFFFu) >> 4) | b_lbo;
                if (leader) {
                    for (int tt = 0; tt < g; ++tt) {
                        const uint32_t b_lo = b_stage + (uint32_t)(tt * nck * N);
                        const uint32_t a_lo1 = a_lo0 + (uint32_t)((tap0 + tt) * a.dil - a.pad - a.min_off);
                        for (int mt = 0; mt < MT; ++mt)
                            umma_ksteps<P>(tmem_base + (uint32_t)(mt * N), a_hi, a_lo1 + (uint32_t)(mt * 128), b_lo,
                                              2u * (uint32_t)R, 2u * (uint32_t)N, idesc, ksteps, acc_on | (uint32_t)tt);
                    }
                    tc_commit(W_EMPTY(sw_i));
                }
                __syncwarp();
                acc_on = 1;
                if (++sw_i == a.sw) { sw_i = 0; sw_ph ^= 1; }
            }
            if (leader) tc_commit(A_EMPTY(sa_i));
            __syncwarp();
            if (++sa_i == a.sa) { sa_i = 0; sa_ph ^= 1; }
        }
        if (leader) tc_commit(ACC_FULL);
        __syncwarp();
        HFG_CONV_TL(3);
    } else {
        // ===================== epilogue: TMEM -> regs -> global =====================
        // 32 columns per step: both TMEM loads, the residual cells and the MRF partial sums are all
        // issued before the first use, so one step pays one memory latency instead of one per 16 bytes.
        mbar_wait(ACC_FULL, 0);
        tc_fence_after();
        if (warp == 2) HFG_CONV_TL(4);
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;                   // the two warps of a quarter alternate 32-column steps
        const int qlane = quarter * 32 + lane;
        const float slope = a.slope, inv_slope = 1.0f / a.slope;
        const float inv_div = 1.0f / a.div;
        const bool add_prev = a.acc_mode == TC_ACC_ADD || a.acc_mode == TC_ACC_FINAL;
        const bool acc_store = a.acc_mode == TC_ACC_WRITE || a.acc_mode == TC_ACC_ADD;
        for (int mt = 0; mt < MT; ++mt) {
            const int q = q0 + mt * 128 + qlane;
            const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * N);
            // launched with 4 or 8 epilogue warps (blockDim 192 / 320): with 8, the two warps of a TMEM lane
            // quarter alternate 32-column steps
            const int col_step = 32 * (((int)blockDim.x - 64) / 128);
            for (int c0 = 32 * half; c0 < N; c0 += col_step) {
                const bool two = c0 + 16 < N;
                uint32_t r0[16], r1[16];
                tmem_ld16(tbase + (uint32_t)c0, r0);
                if (two) tmem_ld16(tbase + (uint32_t)(c0 + 16), r1);
                float x0[16], x1[16], p0[16], p1[16];
                int ch0 = ntile * N + c0;                   // first output channel of this step
                int ph = phase;
                if (a.stack_cout) { ph = ch0 / a.stack_cout; ch0 -= ph * a.stack_cout; }   // a step never straddles phases
                const int t = q * a.out_stride + ph + a.out_off;
                const bool valid = (q < a.n_q) && (t >= 0) && (t < a.T_out);
                const long long row_bytes = (long long)(kPadL + t) * 16;
                const uint8_t* rp = a.res + (long long)b * a.o_bstride + row_bytes;
                uint8_t* op = a.out + (long long)b * a.o_bstride + row_bytes;
                uint8_t* ap = reinterpret_cast<uint8_t*>(a.acc) + (long long)b * a.acc_bstride + row_bytes;
                if (valid && a.res) {
                    load_cells16<P>(rp + (long long)(ch0 / CW) * a.o_pstride, a.o_pstride, x0);
                    if (two) load_cells16<P>(rp + (long long)((ch0 + 16) / CW) * a.o_pstride, a.o_pstride, x1);
                }
                if (valid && add_prev) {
                    load_f32x16(ap + (long long)(ch0 / 4) * a.acc_pstride, a.acc_pstride, p0);
                    if (two) load_f32x16(ap + (long long)((ch0 + 16) / 4) * a.acc_pstride, a.acc_pstride, p1);
                }
                tmem_ld_wait();
                if (!valid) continue;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    if (hh == 1 && !two) break;
                    const int cc = c0 + 16 * hh, ch = ch0 + 16 * hh;
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(hh ? r1[i] : r0[i]);
                    add_bias16(v, sBias + cc);
                    if (a.res) {                             // x + xt   (reference :85)
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += lrelu_inv(hh ? x1[i] : x0[i], inv_slope);
                    }
                    if (add_prev) {                          // output + rb(x)  (reference :129)
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += hh ? p1[i] : p0[i];
                    }
                    if (acc_store) {
                        store_f32x16(ap + (long long)(ch / 4) * a.acc_pstride, a.acc_pstride, v);
                        if (!a.out) continue;
                    }
                    if (a.acc_mode == TC_ACC_FINAL) {        // / len(resblocks)  (reference :131)
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] *= inv_div;
                    }
                    if (a.out) {                             // next layer's leaky_relu, operand dtype
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = lrelu(v[i], slope);
                        if constexpr (P == PREC_FP16) {
                            if (a.out_lo) {
                                store_split16(op + (long long)(ch / CW) * a.o_pstride,
                                              a.out_lo + (long long)b * a.o_bstride + row_bytes + (long long)(ch / CW) * a.o_pstride,
                                              a.o_pstride, v);
                                continue;
                            }
                        }
                        store_cells16<P>(op + (long long)(ch / CW) * a.o_pstride, a.o_pstride, v);
                    }
                }
            }
        }
    }
    if (warp == 2) HFG_CONV_TL(5);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
    if (warp == 0) HFG_CONV_TL(6);
}

// ------------------------------------------------------------------ small helpers
// mel [B, C, T] fp32 (reference layout) -> chunk planes in the operand dtype (no activation).
template <int P>
__global__ void tc_pack_input(const float* __restrict__ x, uint8_t* __restrict__ out, int C, int T,
                              long long bstride, long long pstride, int frames_last) {
    constexpr int CW = Prec<P>::CW;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.y, b = blockIdx.z;
    if (t >= T) return;
    float v[CW];
    if (frames_last) {                                     // x is [B, T, C]: one contiguous cell
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] = x[((size_t)b * T + t) * C + chunk * CW + i];
    } else {
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] = x[((size_t)b * C + chunk * CW + i) * T + t];
    }
    uint8_t* p = out + (long long)b * bstride + (long long)chunk * pstride + (long long)(kPadL + t) * 16;
    *reinterpret_cast<uint4*>(p) = floats_to_cell<P>(v);
}

// chunk planes holding leaky_relu(x) -> x as [B, C, T] fp32 (stage dumps for tests).
template <int P>
__global__ void tc_unpack_stage(const uint8_t* __restrict__ in, float* __restrict__ y, int C, int T,
                                long long bstride, long long pstride, float inv_slope) {
    constexpr int CW = Prec<P>::CW;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.y, b = blockIdx.z;
    if (t >= T) return;
    const uint8_t* p = in + (long long)b * bstride + (long long)chunk * pstride + (long long)(kPadL + t) * 16;
    float v[CW];
    cell_to_floats<P>(*reinterpret_cast<const uint4*>(p), v);
#pragma unroll
    for (int i = 0; i < CW; ++i)
        y[((size_t)b * C + chunk * CW + i) * T + t] = lrelu_inv(v[i], inv_slope);
}

// Zero the padding rows a VALID output can depend on: the kPadL rows in front of the data and the first
// kZeroTail rows after it (receptive halo of any layer <= 32 rows; polyphase / conv_post overhang <= 4).
// Rows further out are only ever read into accumulator rows that are never stored (a UMMA output row
// depends on its own operand rows only), so they may hold anything.  Up to 40 buffers per launch.
constexpr int kZeroTail = 96;
struct PadJob { uint8_t* base; long long planes; int TP; int T; };
struct PadJobs { PadJob job[40]; int n; };
__global__ void tc_zero_pads(const PadJobs jobs) {
    const PadJob j = jobs.job[blockIdx.y];
    const int tail = (j.TP - kPadL - j.T) < kZeroTail ? (j.TP - kPadL - j.T) : kZeroTail;
    const int pad_rows = kPadL + tail;                     // front rows + tail rows
    const long long total = j.planes * pad_rows;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long plane = e / pad_rows;
        int r = (int)(e - plane * pad_rows);
        if (r >= kPadL) r += j.T;                          // skip the valid rows
        *reinterpret_cast<uint4*>(j.base + (plane * j.TP + r) * 16) = make_uint4(0, 0, 0, 0);
    }
}

// Variable-length batches: per-utterance row counts of every stage, from the valid frame counts.
//   tab[0][b]      = min(T, len_b + halo)                    frames read by conv_pre
//   tab[1+i][b]    = rows of stage i (after ups[i]) for those frames
//   tab[1+n][b]    = waveform samples of the len_b valid frames (no halo): everything beyond is returned as 0
struct LenGeom { int n_stages; int u[HFG_MAX_STAGES], k[HFG_MAX_STAGES], p[HFG_MAX_STAGES]; };
__global__ void tc_len_table(const int* __restrict__ lengths, int B, int T, int halo, const LenGeom g, int* __restrict__ tab) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int len = lengths[b];
    len = len < 0 ? 0 : (len > T ? T : len);
    long long eff = len + halo < T ? len + halo : T, val = len;
    tab[b] = (int)eff;
    for (int i = 0; i < g.n_stages; ++i) {
        eff = eff > 0 ? (eff - 1) * g.u[i] - 2 * g.p[i] + g.k[i] : 0;
        val = val > 0 ? (val - 1) * g.u[i] - 2 * g.p[i] + g.k[i] : 0;
        tab[(size_t)(1 + i) * B + b] = (int)eff;
    }
    tab[(size_t)(1 + g.n_stages) * B + b] = (int)val;
}
// fp32 mode: the full batch is generated; samples beyond the valid length are zeroed afterwards
__global__ void tc_mask_tail(float* __restrict__ y, const int* __restrict__ valid_len, int T) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T && t >= valid_len[b]) y[(size_t)b * T + t] = 0.f;
}

// conv_post (C_out = 1, K taps) + tanh from chunk planes that already hold leaky_relu(x)
// (reference models/hifigan.py:254-256).  HBM-bound: reads C channels per step, writes one sample.
//
// A block stages (kPostTile + K - 1) rows of up to kPostGC chunk columns in shared memory with 16-byte cp.async
// copies (no register round trip, every copy of the block in flight at once: the first version of this kernel
// spent its time in long-scoreboard stalls of a load -> store staging loop), then every thread produces
// kPostR CONSECUTIVE samples from a sliding window: a cell is converted once and feeds up to kPostR * CW FMAs,
// a chunk's K * CW weights are loaded once per thread.  Threads of a warp start kPostR cells apart, which would
// put 8 lanes of a 16-byte shared load on 2 bank groups; one pad cell after every 4 cells (position
// c + c / 4) spreads them over all 8 (5 t mod 8 is a permutation).  Accumulation order per sample is
// chunk -> tap -> channel, as in the scalar version.
constexpr int kPostThreads = 128, kPostR = 4, kPostTile = kPostThreads * kPostR, kPostGC = 4;
__host__ __device__ constexpr int post_padded(int c) { return c + (c >> 2); }
template <int K>
__host__ __device__ constexpr int post_col_cells() { return post_padded(kPostTile + K - 1) + 1; }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
template <int P, int K>
__global__ void __launch_bounds__(kPostThreads)
tc_conv_post_tanh(const uint8_t* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ y, int C, int T, int pad, long long bstride, long long pstride,
                  const int* __restrict__ len_rows, const int* __restrict__ valid_len) {
    constexpr int CW = Prec<P>::CW;
    constexpr int ROWS = kPostTile + K - 1, COL = post_col_cells<K>();
    extern __shared__ __align__(16) uint8_t post_smem[];
    const int n_chunks = C / CW;
    uint4* cells = reinterpret_cast<uint4*>(post_smem);     // [kPostGC][COL]: one group of chunk columns at a time
    float* wsm = reinterpret_cast<float*>(cells + (size_t)kPostGC * COL);     // [chunk][K][CW]
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * kPostTile;
    // variable-length batches: samples at or beyond valid_len[b] are zeros; tiles beyond the rows that were
    // generated for utterance b (len_rows[b] = valid + halo) are not even read
    const int Tv = valid_len ? (valid_len[b] < T ? valid_len[b] : T) : T;
    if (len_rows && t0 >= len_rows[b]) {
        for (int e = threadIdx.x; e < kPostTile && t0 + e < T; e += blockDim.x) y[(size_t)b * T + t0 + e] = 0.f;
        return;
    }
    for (int e = threadIdx.x; e < C * K; e += blockDim.x) {
        const int ci = e / K, j = e - ci * K;               // w is [C][K]
        wsm[((ci / CW) * K + j) * CW + (ci % CW)] = w[e];
    }
    // rows t0 - pad .. t0 - pad + ROWS: the planes carry kPadL zero rows in front of the data and the tile
    // overhang behind it (tc_tp); rows past T + kZeroTail may hold anything but only feed samples >= T
    const uint8_t* base = in + (long long)b * bstride + (long long)(kPadL + t0 - pad) * 16;
    const int tl = threadIdx.x * kPostR;                    // first local sample of this thread
    float acc[kPostR];
#pragma unroll
    for (int i = 0; i < kPostR; ++i) acc[i] = 0.f;
    for (int c0 = 0; c0 < n_chunks; c0 += kPostGC) {
        const int gc = (n_chunks - c0) < kPostGC ? (n_chunks - c0) : kPostGC;
        __syncthreads();                                    // previous group consumed
        for (int e = threadIdx.x; e < gc * ROWS; e += blockDim.x) {
            const int chunk = e / ROWS, r = e - chunk * ROWS;
            cp_async16(cells + chunk * COL + post_padded(r), base + (long long)(c0 + chunk) * pstride + (long long)r * 16);
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                    // staged cells (and, first time round, wsm) visible
        for (int chunk = 0; chunk < gc; ++chunk) {
            float wv[K][CW];
#pragma unroll
            for (int j = 0; j < K; ++j)
#pragma unroll
                for (int i = 0; i < CW; i += 4)
                    *reinterpret_cast<float4*>(&wv[j][i]) =
                        *reinterpret_cast<const float4*>(wsm + ((c0 + chunk) * K + j) * CW + i);
            // cell tl + m sits at post_padded(tl + m) = 5 * tid + m + m / 4   (tl = 4 * tid)
            const uint4* col = cells + chunk * COL + 5 * threadIdx.x;
#pragma unroll
            for (int m = 0; m < kPostR + K - 1; ++m) {      // window row m feeds sample r through tap j = m - r
                float v[CW];
                cell_to_floats<P>(col[m + (m >> 2)], v);
#pragma unroll
                for (int r = 0; r < kPostR; ++r) {
                    const int j = m - r;
                    if (j >= 0 && j < K) {
#pragma unroll
                        for (int i = 0; i < CW; ++i) acc[r] = fmaf(wv[j][i], v[i], acc[r]);
                    }
                }
            }
        }
    }
    const float bv = bias[0];
    const size_t o = (size_t)b * T + t0 + tl;
    float out[kPostR];
#pragma unroll
    for (int r = 0; r < kPostR; ++r) out[r] = (t0 + tl + r < Tv) ? tanhf(acc[r] + bv) : 0.f;
    if (t0 + tl + kPostR <= T && (o & 3) == 0) {
        *reinterpret_cast<float4*>(y + o) = make_float4(out[0], out[1], out[2], out[3]);
    } else {
#pragma unroll
        for (int r = 0; r < kPostR; ++r)
            if (t0 + tl + r < T) y[o + r] = out[r];
    }
}

}  // namespace hfg
