// Fused ResBlock pair on tcgen05 (sm_100a):
//
//     x_new = x + conv2( leaky_relu( conv1_d( leaky_relu(x) ) + b1 ) ) + b2
//                                             (reference models/hifigan.py:80-85)
//
// in ONE persistent kernel: the intermediate never leaves the SM, the residual is
// never re-read from global memory, and both leaky_relus, both biases, the
// residual add, the MRF running sum / division (reference :126-131) and the
// conversion to the operand dtype are fused.
//
// Per tile (one utterance, TO = MT*128 - (k-1) output time steps):
//
//   producer warp   bulk-copies the leaky_relu(x) tile (rows t0-p2-p1 .. , 8-chunk K blocks)
//                   and streams W1 then W2 (one (K block, tap) stage at a time)
//   MMA warp        conv1: D1[mt] += A(shifted by tap*d rows) * W1[tap]      (acc1, TMEM)
//                   conv2: D2[mt] += H(shifted by tap rows)   * W2[tap]      (acc2, TMEM)
//   8 epilogue warps
//     pre2  (overlaps conv1 MMAs)  acc2 <- x + b2 (+ MRF partial sum): the residual is the
//           centre rows of the A tile already in smem (leaky_relu inverted exactly), written
//           to TMEM with tcgen05.st, so conv2 simply accumulates on top of it
//     epi1  acc1 -> +b1 -> leaky_relu -> zero outside [0,T) (conv2's own zero padding)
//           -> operand dtype -> smem H tile in the same chunk-plane layout conv2's A
//           descriptors address with row shifts
//     epi2  acc2 -> (MRF sum bookkeeping) -> leaky_relu for the next consumer -> global
//
// TMEM: acc1 = MT*N columns, acc2 = MT*N columns (MT*N <= 256).  While the epilogue
// warps drain acc2 of tile i the MMA warp already runs conv1 of tile i+1 into acc1.
#pragma once
#include "tc_kernels.cuh"

namespace hfg {

constexpr int kPairEpiWarps = 8;                  // default; the one-CTA-per-SM variants can run 12 (template parameter EW)
constexpr int pair_threads(int ew) { return 64 + 32 * ew; }
constexpr int kPairThreads = pair_threads(kPairEpiWarps);
constexpr int kPairMaxSA = 6;                     // activation ring slots (K blocks, possibly of several tiles ahead)

struct TcPairArgs {
    const uint8_t* a; long long a_bstride, a_pstride;     // leaky_relu(x) planes
    const uint8_t* w1; const uint8_t* w2;                 // packed [kb][tap][chunk][N][16 B]; CTA pair: [half][..][N/2][16 B]
    long long w_half_stride, w2_half_stride;              // bytes between the two N-halves (CTA pair only)
    const float* b1; const float* b2;
    uint8_t* out; long long o_bstride, o_pstride;         // leaky_relu(x_new) planes (may be null)
    float* acc; long long acc_bstride, acc_pstride;       // MRF fp32 partial sums (bytes): unfused-neighbour fallback only
    int acc_mode; float div;
    // MRF sum without an fp32 round trip (reference :126-131): the last pair of the LAST resblock (TC_ACC_FINAL)
    // adds the finished outputs of the other resblocks -- ordinary leaky_relu(x_j) planes with the geometry of
    // `out`, inverted exactly on load -- to its accumulator before conv2, then scales by 1 / n_resblocks.
    const uint8_t* sum_in[HFG_MAX_STAGES - 1]; int n_sum;
    // LO kernels (tf32 mode on fp16 operand planes, tc_path.cuh "split plan"): the residual stream is stored as the fp16
    // pair hi + lo (22 mantissa bits).  `a` / `out` are the hi planes -- what the MMAs read --, a_lo / out_lo their lo
    // twins (same geometry).  The producer copies the lo cells of the tile's own MT * 128 rows into the H-TILE BUFFER,
    // which is idle between conv2 of tile i (tcgen05.commit -> ACC2_FULL) and epi1 of tile i+1: no extra shared
    // memory, nothing loaded synchronously; pre2 forms the residual hi + lo from shared memory, and a named barrier
    // of the epilogue warps separates those reads from epi1's writes of H.  out_lo may be null (the result only
    // feeds MMAs); out32 (fp32 cells, geometry o32_*) replaces out / out_lo where the result only feeds the MRF sum
    // or conv_post; sum_in then points at such fp32 planes.
    const uint8_t* a_lo; uint8_t* out_lo; uint8_t* out32; long long o32_bstride, o32_pstride;
    int N;            // channels (C_in = C_out = N)
    int n_chunks;     // N / CW
    int MT;           // 128-row sub-tiles per tile
    int T;            // valid time steps
    int k, dil, p1, p2;
    int R1;           // rows per A stage   = MT*128 + 2*p1
    int RH;           // rows per H plane  >= MT*128 + 2*p2
    int TO;           // output rows per tile = MT*128 - 2*p2
    int sa, sw;
    int tap_group;    // taps per W stage (conv1)
    // conv2 in space-to-depth form (s2d = 1, narrow layers): GEMM row m of conv2 holds the two time steps 2m, 2m+1
    // of the tile -- H' has 2N channels (parity, channel), D2 has 2N columns (parity, channel) and MT / 2 sub-tiles,
    // the k taps collapse into k2 = (k + 1) / 2 taps of a 2N x 2N matrix of which k / (k + 1) is non-zero.  One
    // A-operand read (the 64 B/clk shared-memory read that bounds the narrow layers) then feeds twice the output
    // columns: conv2 needs (k + 1) / 2 * 2 MMAs per 256 time steps instead of 2 * k.  s2d = 0: k2 = k, tap_group2 = tap_group.
    int s2d, k2, tap_group2;
    // groups = 2: the tile runs as TWO INDEPENDENT HALVES (sub-tiles [0, MT/2) and [MT/2, MT)), each with its own
    // accumulator / H-tile barriers and its own four epilogue warps.  The MMA warp runs conv1 of the upper half, conv1
    // of the lower half, conv2 of the upper half (it needs no H row of the lower one), conv2 of the lower half; so one
    // half's epi1 overlaps the other half's conv1, and its epi2 + the next tile's pre2 overlap the other half's
    // conv2 -- the k = 3 / 7 layers, whose tile time is the serial chain pre2 -> epi1 -> conv2 -> epi2 of ALL eight
    // warps, lose most of the time the epilogue warps spent waiting for whole-tile accumulators.  Costs: the
    // weight stages of each convolution are streamed once per half, and every activation K block of the tile must
    // be resident at once (sa == number of K blocks).  groups = 1: one accumulator set per tile, as before.
    int groups;
    unsigned w_stage_bytes;   // bytes of one W ring slot (the larger of the conv1 / conv2 stage)
    int kbc;          // 16-byte cells per K block
    int poll_ns;      // producer back-off when both rings are full
    int epi_sleep_ns; // epilogue warps: longest sleep between polls of a barrier (0 = spin)
    int pdl;          // launched with programmatic stream serialisation: the prologue above griddepcontrol.wait (barrier
                      // init, TMEM allocation, bias staging -- nothing an earlier kernel writes) overlaps the predecessor's tail
    int dbg;          // HFG_TUNING builds only -- timing experiments (results are wrong): 1 = no weight copies, 2 = no activation copies, 4 = tap shifts of 8 rows (128-byte aligned operand reads), 8 = epilogue warps only keep the barrier protocol, 16 = no MMAs issued
    int tiles_per_batch, n_tiles;
    // variable-length batches: rows of this stage that utterance b needs (valid frames + receptive halo, scaled to
    // this stage), or null.  Tiles that start at or beyond len_rows[b] are never loaded, multiplied or stored.
    const int* len_rows;
    float slope;
    unsigned long long* timeline;   // tuning only (tools/pair_timeline.py): clock64 stamps, [cta < 4][tile < 16][event < 16]
};

// per-tile phase stamps of the first CTAs (HFG_TUNING builds only)
#ifndef HFG_TUNING
#define HFG_TL(ev, tile_ord) do { } while (0)
#else
#define HFG_TL(ev, tile_ord)                                                                                  \
    do {                                                                                                      \
        if (a.timeline && blockIdx.x < 4 && (tile_ord) < 16 && lane == 0)                                     \
            a.timeline[((size_t)blockIdx.x * 16 + (tile_ord)) * 16 + (ev)] = (unsigned long long)clock64();   \
    } while (0)
#endif

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// P2 = operand precision of conv2, i.e. of the on-chip intermediate H and of W2.  P2 == P except in tf32 mode,
// where P2 = PREC_FP16: the MMA would round the fp32 H tile to tf32's 10-bit mantissa anyway, so the epilogue
// stores it as fp16 (same mantissa, round-to-nearest, saturating) -- conv2 then runs at the fp16 MMA rate on
// half the shared-memory operand bytes, and the H tile of C = 256 fits next to the operand rings.  The
// residual stream (A tile -> acc2 -> output planes) stays fp32.
// MINB = CTAs per SM the register allocation must allow (2 for the narrow layers, whose tiles are
// latency-bound and want a second CTA to fill the tensor pipe while the first one is in an epilogue)
// CTAS = 2: a cluster of two CTAs runs two adjacent tiles in lockstep; the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256) for both, each CTA stages only its half of every weight tile
// (half the L2->smem weight traffic and half the B-operand smem reads per SM).
// LO = true (P = P2 = PREC_FP16): operands are the fp16 hi planes, the residual stream is the pair hi + lo -- the
// arithmetic of the tf32 mode (10-bit-mantissa operands, fp32 accumulate, >= 22-bit residual stream) at the fp16
// mode's MMA rate and shared-memory operand bytes, with the fp32 planes' HBM bytes.
// EW = epilogue warps (8 or 12): EW / 4 warps share each TMEM lane quarter.  With three, the (sub-tile, 32-column
// group) units of an accumulator go round-robin to the warps of a quarter -- the same map in pre2 and epi2, so a
// warp only ever writes accumulator columns it has itself drained.
template <int P, int P2, int MINB, int CTAS, bool LO = false, int EW = kPairEpiWarps>
__global__ void __launch_bounds__(pair_threads(EW), MINB)
tc_pair_kernel(const TcPairArgs a) {
    static_assert(EW == 8 || EW == 12, "two or three epilogue warps per TMEM lane quarter");
    extern __shared__ __align__(128) uint8_t tc_pair_smem[];
    uint8_t* smem = tc_pair_smem;
    constexpr int CW = Prec<P>::CW, CW2 = Prec<P2>::CW;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = a.N, MT = a.MT, R1 = a.R1, RH = a.RH, k = a.k;
    const int NB = N / CTAS;                          // rows of the weight tile staged by this CTA
    const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
    const int n_chunks = a.n_chunks;
    const int KBC = a.kbc;                            // 16-byte cells per K block (8, or 4 when smem is tight)
    const int nck_max = n_chunks < KBC ? n_chunks : KBC;
    const int n_kb = (n_chunks + KBC - 1) / KBC;
    const int PP = a.s2d ? 2 : 1;                     // time steps per GEMM row of conv2
    const int N2 = N * PP, NB2 = N2 / CTAS, MT2 = MT / PP, k2 = a.k2;
    const int n_chunks2 = (N / CW2) * PP;             // cells per row of the H tile (conv2's K extent)
    const int n_kb2 = (n_chunks2 + KBC - 1) / KBC;
    const int RL = MT * 128;                          // LO kernels: lo rows per chunk, parked in the H-tile buffer
    const uint32_t a_stage_bytes = (uint32_t)R1 * nck_max * 16;
    const int G = a.tap_group, G2 = a.tap_group2;
    const uint32_t w_stage_bytes = a.w_stage_bytes;
    uint8_t* sA = smem;
    uint8_t* sW = sA + (size_t)a.sa * a_stage_bytes;
    uint8_t* sH = sW + (size_t)a.sw * w_stage_bytes;
    float* sB1 = reinterpret_cast<float*>(sH + (size_t)n_chunks2 * RH * 16);
    float* sB2 = sB1 + N;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB2 + N);
    const uint32_t bar0 = smem_u32(bars);
    auto A_FULL = [&](int i) { return bar0 + 8u * i; };
    auto A_EMPTY = [&](int i) { return bar0 + 8u * (kPairMaxSA + i); };
    auto W_FULL = [&](int i) { return bar0 + 8u * (2 * kPairMaxSA + i); };
    auto W_EMPTY = [&](int i) { return bar0 + 8u * (2 * kPairMaxSA + kMaxSW + i); };
    const uint32_t ACC1_FULL = bar0 + 8u * (2 * kPairMaxSA + 2 * kMaxSW);
    const uint32_t H_READY = ACC1_FULL + 8, ACC2_FULL = ACC1_FULL + 16, LO_FULL = ACC1_FULL + 24;
    // second set for the upper half of a tile run as two halves (a.groups == 2); the lower half -- the one whose
    // conv2 completes last -- uses the first set
    const uint32_t ACC1_FULL_B = ACC1_FULL + 32, H_READY_B = ACC1_FULL + 40, ACC2_FULL_B = ACC1_FULL + 48;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPairMaxSA + 2 * kMaxSW + 7);
    const int GR = a.groups;
    const int MTg = MT / GR, MT2g = MT2 / GR;         // sub-tiles per half

    uint32_t ncols = 32;
    while ((int)ncols < 2 * MT * N) ncols <<= 1;

    if (warp == 0 && lane == 0) {
        // pair mode: the LEADER's full barriers also collect the peer's "my copy of this stage has landed"
        // (one remote arrive), so its MMA warp waits on a single barrier per stage
        const uint32_t full_count = (CTAS == 2 && rank == 0) ? 2u : 1u;
        for (int i = 0; i < a.sa; ++i) { mbar_init(A_FULL(i), full_count); mbar_init(A_EMPTY(i), 1 + EW); }
        for (int i = 0; i < a.sw; ++i) { mbar_init(W_FULL(i), full_count); mbar_init(W_EMPTY(i), 1); }
        mbar_init(ACC1_FULL, 1);
        mbar_init(H_READY, (EW / GR) * CTAS);
        mbar_init(ACC2_FULL, 1);
        mbar_init(LO_FULL, 1);
        mbar_init(ACC1_FULL_B, 1);
        mbar_init(H_READY_B, (EW / GR) * CTAS);
        mbar_init(ACC2_FULL_B, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (CTAS == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(smem_u32(tmem_slot)), "r"(ncols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(smem_u32(tmem_slot)), "r"(ncols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < N; i += 32 * EW) { sB1[i] = a.b1[i]; sB2[i] = a.b2[i]; }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CTAS == 2) cluster_sync_all();      // peer barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc1 = tmem_base, acc2 = tmem_base + (uint32_t)(MT * N);
    if (a.pdl) {
        // let the next launch of this stream place its CTAs as ours retire, then wait until everything this
        // kernel depends on (the previous launch of the stream and all it waited for) has completed and is visible
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    // tile schedule: cluster c takes tile pairs c, c + n_clusters, ...; CTA `rank` runs tile 2*pair + rank.
    // An odd tail tile is duplicated on the peer (same coordinates, stores suppressed) so both CTAs
    // stay in lockstep on the shared weight ring.
    const int n_sched = (a.n_tiles + CTAS - 1) / CTAS;
    const int sched0 = (int)blockIdx.x / CTAS, sched_step = (int)gridDim.x / CTAS;
    auto tile_of = [&](int sc) { const int t = sc * CTAS + (int)rank; return t < a.n_tiles ? t : a.n_tiles - 1; };
    auto tile_live = [&](int t) {                      // inside the batch and not beyond its utterance's length
        if (t >= a.n_tiles) return false;
        if (!a.len_rows) return true;
        return (t % a.tiles_per_batch) * a.TO < a.len_rows[t / a.tiles_per_batch];
    };
    auto tile_real = [&](int sc) { return tile_live(sc * CTAS + (int)rank); };
    // a schedule slot runs when any of its CTAS tiles is live; the other CTA of a pair then runs its (dead)
    // tile with stores suppressed, exactly like the duplicated odd tail tile.  Every role of both CTAs
    // evaluates this predicate on the same data, so they skip the same slots.
    auto sched_live = [&](int sc) {
        if (!a.len_rows) return true;
        for (int r = 0; r < CTAS; ++r)
            if (tile_live(sc * CTAS + r)) return true;
        return false;
    };
    auto next_live = [&](int sc) { while (sc < n_sched && !sched_live(sc)) sc += sched_step; return sc; };

    if (warp == 0) {
        // ===================== producer =====================
        const bool leader = elect_one();
        int sa_i = 0, sa_ph = 0, sw_i = 0, sw_ph = 0;
        auto tile_src = [&](int tile) {
            const int b = tile / a.tiles_per_batch;
            const int t0 = (tile % a.tiles_per_batch) * a.TO;
            return a.a + (long long)b * a.a_bstride + (long long)(kPadL + t0 - a.p2 - a.p1) * 16;
        };
        // Two independent streams -- activation K blocks (A ring) and weight stages (W ring) -- issued by
        // one polling thread: whichever ring has a free slot gets its next copy.  (Issuing them in one
        // blocking program order let a full A ring stall the weight stream and starve the MMAs:
        // profiles/r1_tuning.md section 6.)
        const int groups = (k + G - 1) / G, groups2 = (k2 + G2 - 1) / G2;
        int a_sc = next_live(sched0), a_kb = 0;                  // next A block: schedule slot, K block
        int w_sc = a_sc, w_conv = 0, w_kb = 0, w_g = 0, w_pass = 0;   // next W stage (each convolution's stages once per half)
        int a_t = 0, w_t = 0;                                    // ordinals of those slots among this CTA's live ones
        int l_sc = LO ? a_sc : n_sched, l_t = 0;                 // LO kernels: next tile whose lo rows go into the H-tile buffer
        uint32_t idle = 0;
        long long t_idle0 = 0;
        while (a_sc < n_sched || w_sc < n_sched || l_sc < n_sched) {
            bool did = false;
            if constexpr (LO) {
                // the H-tile buffer is free once conv2 of the previous tile has completed (its tcgen05.commit on ACC2_FULL)
                if (l_sc < n_sched && (l_t == 0 || mbar_test(ACC2_FULL, (uint32_t)((l_t - 1) & 1)))) {
                    if (leader) {
                        const int tile = tile_of(l_sc);
                        const uint8_t* lb = a.a_lo + (long long)(tile / a.tiles_per_batch) * a.a_bstride +
                                            (long long)(kPadL + (tile % a.tiles_per_batch) * a.TO) * 16;
                        mbar_expect_tx(LO_FULL, (uint32_t)n_chunks * RL * 16);
                        const uint32_t dst = smem_u32(sH);
                        for (int c = 0; c < n_chunks; ++c)
                            bulk_g2s(dst + (uint32_t)c * RL * 16, lb + (long long)c * a.a_pstride, (uint32_t)RL * 16, LO_FULL);
                    }
                    __syncwarp();
                    ++l_t;
                    l_sc = next_live(l_sc + sched_step);
                    did = true;
                }
            }
            if (a_sc < n_sched && mbar_test(A_EMPTY(sa_i), sa_ph ^ 1)) {
                const int nck = (n_chunks - KBC * a_kb) < KBC ? (n_chunks - KBC * a_kb) : KBC;
                if (a_kb == 0) HFG_TL(0, a_t);
                if (leader && HFG_DBG(a, 2)) mbar_arrive(A_FULL(sa_i));
                if (leader && !HFG_DBG(a, 2)) {
                    const uint8_t* ab = tile_src(tile_of(a_sc));
                    mbar_expect_tx(A_FULL(sa_i), (uint32_t)nck * R1 * 16);
                    const uint32_t dst = smem_u32(sA + (size_t)sa_i * a_stage_bytes);
                    for (int c = 0; c < nck; ++c)
                        bulk_g2s(dst + (uint32_t)c * R1 * 16, ab + (long long)(KBC * a_kb + c) * a.a_pstride,
                                 (uint32_t)R1 * 16, A_FULL(sa_i));
                }
                __syncwarp();
                if (++sa_i == a.sa) { sa_i = 0; sa_ph ^= 1; }
                if (++a_kb == n_kb) { a_kb = 0; ++a_t; a_sc = next_live(a_sc + sched_step); }
                did = true;
            }
            if (w_sc < n_sched && mbar_test(W_EMPTY(sw_i), sw_ph ^ 1)) {
                const int w_chunks = w_conv ? n_chunks2 : n_chunks, w_nkb = w_conv ? n_kb2 : n_kb;
                const int w_k = w_conv ? k2 : k, w_G = w_conv ? G2 : G, w_groups = w_conv ? groups2 : groups;
                const int w_NB = w_conv ? NB2 : NB;
                const int nck = (w_chunks - KBC * w_kb) < KBC ? (w_chunks - KBC * w_kb) : KBC;
                const int tap0 = w_g * w_G;
                const int g = (w_k - tap0) < w_G ? (w_k - tap0) : w_G;
                if (w_conv == 0 && w_kb == 0 && w_g == 0) HFG_TL(11, w_t);
                if (leader && HFG_DBG(a, 1)) mbar_arrive(W_FULL(sw_i));
                if (leader && !HFG_DBG(a, 1)) {
                    const uint8_t* w = w_conv ? a.w2 : a.w1;
                    mbar_expect_tx(W_FULL(sw_i), (uint32_t)g * nck * w_NB * 16);
                    bulk_g2s(smem_u32(sW + (size_t)sw_i * w_stage_bytes),
                             w + (long long)rank * (w_conv ? a.w2_half_stride : a.w_half_stride) +
                                 ((long long)w_kb * w_k * KBC + (long long)tap0 * nck) * w_NB * 16,
                             (uint32_t)g * nck * w_NB * 16, W_FULL(sw_i));
                }
                __syncwarp();
                if (++sw_i == a.sw) { sw_i = 0; sw_ph ^= 1; }
                if (++w_g == w_groups) { w_g = 0; if (++w_kb == w_nkb) { w_kb = 0; if (++w_pass == GR) { w_pass = 0; if (++w_conv == 2) { w_conv = 0; ++w_t; w_sc = next_live(w_sc + sched_step); } } } }
                did = true;
            }
            if (did) { idle = 0; t_idle0 = 0; continue; }
            if (a.poll_ns > 0) __nanosleep(a.poll_ns);
            if ((++idle & 4095u) == 0) {                         // bounded: trap instead of hanging the GPU
                const long long now = clock64();
                if (t_idle0 == 0) t_idle0 = now;
                else if (now - t_idle0 > 4000000000ll) __trap();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc<P>(N, 128 * CTAS), idesc2 = umma_idesc<P2>(N2, 128 * CTAS);
        const uint32_t d_hi = (128u >> 4) | (1u << 14);                     // SBO = 128 B, version 1
        const uint32_t a_lbo = ((uint32_t)R1) << 16, h_lbo = ((uint32_t)RH) << 16, b_lbo = ((uint32_t)NB) << 16;
        const uint32_t b2_lbo = ((uint32_t)NB2) << 16;
        auto commit = [&](uint32_t bar) { if constexpr (CTAS == 2) tc_commit2(bar); else tc_commit(bar); };
        if (CTAS == 2 && rank == 1) {
            // ---- peer CTA: no MMA issue.  Forward "my stage has landed" to the leader: one lane per ring
            // slot, so slots are forwarded independently instead of through one serial wait chain ----
            int n_my = 0;
            for (int sc = next_live(sched0); sc < n_sched; sc = next_live(sc + sched_step)) ++n_my;
            const int stages_per_tile = GR * (n_kb * ((k + G - 1) / G) + n_kb2 * ((k2 + G2 - 1) / G2));
            const int total_w = n_my * stages_per_tile, total_a = n_my * n_kb;
            if (lane < a.sw) {
                const int uses = total_w / a.sw + (lane < total_w % a.sw ? 1 : 0);
                for (int u = 0; u < uses; ++u) {
                    mbar_wait(W_FULL(lane), (uint32_t)(u & 1));
                    mbar_arrive_remote(W_FULL(lane), 0);
                }
            } else if (lane < a.sw + a.sa) {
                const int sl = lane - a.sw;
                const int uses = total_a / a.sa + (sl < total_a % a.sa ? 1 : 0);
                for (int u = 0; u < uses; ++u) {
                    mbar_wait(A_FULL(sl), (uint32_t)(u & 1));
                    mbar_arrive_remote(A_FULL(sl), 0);
                }
            }
            __syncwarp();
        } else {
        const uint32_t h_lo_base = ((smem_u32(sH) & 0x3FFFFu) >> 4) | h_lbo;
        int sa_i = 0, sa_ph = 0, sw_i = 0, sw_ph = 0;
        uint32_t it = 0;
        for (int sc = next_live(sched0); sc < n_sched; sc = next_live(sc + sched_step), ++it) {
            // ---- conv1: acc1 = sum_{kb,tap} A(+tap*d rows) * W1, one pass per half (upper half first) ----
            const int sa_i0 = sa_i, sa_ph0 = sa_ph;
            for (int gp = 0; gp < GR; ++gp) {
                const int grp = GR - 1 - gp;
                sa_i = sa_i0; sa_ph = sa_ph0;                  // every pass walks the tile's activation stages
                uint32_t acc_on = 0;
                for (int kb = 0; kb < n_kb; ++kb) {
                    const int nck = (n_chunks - KBC * kb) < KBC ? (n_chunks - KBC * kb) : KBC;
                    const int ksteps = nck >> 1;
                    if (gp == 0) {                             // (still resident in the later pass)
                        mbar_wait(A_FULL(sa_i), sa_ph);
                        tc_fence_after();
                    }
                    if (kb == 0 && gp == 0) HFG_TL(1, it);
                    const uint32_t a_lo0 = ((smem_u32(sA + (size_t)sa_i * a_stage_bytes) & 0x3FFFFu) >> 4) | a_lbo;
                    for (int tap0 = 0; tap0 < k; tap0 += G) {
                        const int g = (k - tap0) < G ? (k - tap0) : G;
                        mbar_wait(W_FULL(sw_i), sw_ph);
                        tc_fence_after();
                        const uint32_t b_stage = ((smem_u32(sW + (size_t)sw_i * w_stage_bytes) & 0x3FFFFu) >> 4) | b_lbo;
                        if (leader) {
                            for (int tt = 0; tt < g && !HFG_DBG(a, 16); ++tt) {
                                const uint32_t b_lo = b_stage + (uint32_t)(tt * nck * NB);
                                const uint32_t a_lo1 = a_lo0 + (uint32_t)((tap0 + tt) * (HFG_DBG(a, 4) ? 8 : a.dil));
                                for (int mt = grp * MTg; mt < (grp + 1) * MTg; ++mt)
                                    umma_ksteps<P, CTAS>(acc1 + (uint32_t)(mt * N), d_hi, a_lo1 + (uint32_t)(mt * 128), b_lo,
                                                            2u * (uint32_t)R1, 2u * (uint32_t)NB, idesc, ksteps,
                                                            acc_on | (uint32_t)tt);
                            }
                            commit(W_EMPTY(sw_i));
                        }
                        __syncwarp();
                        acc_on = 1;
                        if (++sw_i == a.sw) { sw_i = 0; sw_ph ^= 1; }
                    }
                    if (leader && gp == GR - 1) commit(A_EMPTY(sa_i));
                    __syncwarp();
                    if (++sa_i == a.sa) { sa_i = 0; sa_ph ^= 1; }
                }
                if (leader) commit(grp ? ACC1_FULL_B : ACC1_FULL);
                __syncwarp();
            }
            HFG_TL(2, it);
            // ---- conv2: acc2 (pre-loaded with x + b2) += sum_{kb,tap} H(+tap rows) * W2; the upper half reads H rows of
            // its own half only, the lower half also the first 2 p2 rows of the upper one (complete by then) ----
            for (int gp = 0; gp < GR; ++gp) {
                const int grp = GR - 1 - gp;
                const uint32_t h_ready = grp ? H_READY_B : H_READY;
                if constexpr (CTAS == 2) mbar_wait_cluster(h_ready, it & 1); else mbar_wait(h_ready, it & 1);
                tc_fence_after();
                if (gp == 0) HFG_TL(3, it);
                for (int kb = 0; kb < n_kb2; ++kb) {
                    const int nck = (n_chunks2 - KBC * kb) < KBC ? (n_chunks2 - KBC * kb) : KBC;
                    const int ksteps = nck >> 1;
                    const uint32_t h_lo0 = h_lo_base + (uint32_t)(KBC * kb * RH);
                    for (int tap0 = 0; tap0 < k2; tap0 += G2) {
                        const int g = (k2 - tap0) < G2 ? (k2 - tap0) : G2;
                        mbar_wait(W_FULL(sw_i), sw_ph);
                        tc_fence_after();
                        const uint32_t b_stage = ((smem_u32(sW + (size_t)sw_i * w_stage_bytes) & 0x3FFFFu) >> 4) | b2_lbo;
                        if (leader) {
                            for (int tt = 0; tt < g && !HFG_DBG(a, 16); ++tt) {
                                const uint32_t b_lo = b_stage + (uint32_t)(tt * nck * NB2);
                                const uint32_t h_lo1 = h_lo0 + (uint32_t)((tap0 + tt) * (HFG_DBG(a, 4) ? 8 : 1));
                                for (int mt = grp * MT2g; mt < (grp + 1) * MT2g; ++mt)
                                    umma_ksteps<P2, CTAS>(acc2 + (uint32_t)(mt * N2), d_hi, h_lo1 + (uint32_t)(mt * 128), b_lo,
                                                            2u * (uint32_t)RH, 2u * (uint32_t)NB2, idesc2, ksteps, 1u);
                            }
                            commit(W_EMPTY(sw_i));
                        }
                        __syncwarp();
                        if (++sw_i == a.sw) { sw_i = 0; sw_ph ^= 1; }
                    }
                }
                if (leader) commit(grp ? ACC2_FULL_B : ACC2_FULL);
                __syncwarp();
            }
            HFG_TL(4, it);
        }
        }   // leader / single-CTA issue path
    } else {
        // ===================== epilogue warps =====================
        const int e = warp - 2;
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch
        const int half = e >> 2;                      // the two warps of a quarter split the sub-tiles ...
        // ... or, with a single sub-tile, alternate 32-column steps.  acc1 has MT sub-tiles of N columns, acc2 has
        // MT2 sub-tiles of N2 columns (MT / 2 and 2 N in space-to-depth form): each accumulator has its own split,
        // and every phase that touches acc2 (pre2, epi2) uses the same one
        // With the tile run as two halves (GR == 2) the warps of `half` own ALL sub-tiles and columns of that half.
        // Three warps per quarter (EW == 12, GR == 1): every warp walks all sub-tiles and 32-column groups and keeps the
        // units u = sub-tile * groups + group with u % 3 == its index (`own1` / `own2` below).
        constexpr bool W3 = EW == 12;
        const bool split1 = !W3 && GR == 1 && MT == 1, split2 = !W3 && GR == 1 && MT2 == 1;
        const int mt_first = W3 ? 0 : (GR == 2 ? half * MTg : (split1 ? 0 : half)), mt_end = (!W3 && GR == 2) ? (half + 1) * MTg : MT;
        const int mt_step = (W3 || GR == 2 || split1) ? 1 : 2;
        const int cs = split1 ? 2 : 1, ch = split1 ? half : 0;           // acc1: 32-column-step stride / phase
        const int mt2_first = W3 ? 0 : (GR == 2 ? half * MT2g : (split2 ? 0 : half)), mt2_end = (!W3 && GR == 2) ? (half + 1) * MT2g : MT2;
        const int mt2_step = (W3 || GR == 2 || split2) ? 1 : 2;
        const int cs2 = split2 ? 2 : 1, ch2 = split2 ? half : 0;         // acc2
        auto own1 = [&](int mt, int col) { return !W3 || (mt * (N >> 5) + (col >> 5)) % 3 == half; };     // acc1 columns
        auto own2 = [&](int mt, int col) { return !W3 || (mt * (N2 >> 5) + (col >> 5)) % 3 == half; };   // acc2 columns
        const bool upper = GR == 2 && half == 1;                         // this warp's half uses the second barrier set
        const uint32_t acc1_full = upper ? ACC1_FULL_B : ACC1_FULL, h_ready = upper ? H_READY_B : H_READY;
        const uint32_t acc2_full = upper ? ACC2_FULL_B : ACC2_FULL;
        const int row = quarter * 32 + lane;          // row inside a 128-row sub-tile
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        const float slope = a.slope, inv_slope = 1.0f / a.slope;
        const float inv_div = 1.0f / a.div;           // MRF mean as a multiplication (tensor-core modes are not bit-exact anyway)
        const long long a_plane = (long long)R1 * 16, h_plane = (long long)RH * 16;
        const bool add_prev_mode = (a.acc_mode == TC_ACC_ADD || a.acc_mode == TC_ACC_FINAL);
        const bool acc_store_mode = (a.acc_mode == TC_ACC_WRITE || a.acc_mode == TC_ACC_ADD);
        int sa_i = 0, sa_ph = 0;
        uint32_t it = 0;
        for (int sc = next_live(sched0); sc < n_sched; sc = next_live(sc + sched_step), ++it) {
            const int tile = tile_of(sc);
            const bool real = tile_real(sc);          // false: duplicated tail tile, no global side effects
            const int b = tile / a.tiles_per_batch;
            const int t0 = (tile % a.tiles_per_batch) * a.TO;
            // ---------- pre2: acc2 <- x + b2 (+ partial MRF sum), per K block as it lands ----------
            if constexpr (LO) mbar_wait_sleep(LO_FULL, it & 1, (uint32_t)a.epi_sleep_ns);
            for (int kb = 0; kb < n_kb; ++kb) {
                const int nck = (n_chunks - KBC * kb) < KBC ? (n_chunks - KBC * kb) : KBC;
                mbar_wait_sleep(A_FULL(sa_i), sa_ph, (uint32_t)a.epi_sleep_ns);
                if (kb == 0 && e == 0) HFG_TL(5, it);
                const uint8_t* sa_p = sA + (size_t)sa_i * a_stage_bytes;
                for (int mt = mt2_first; mt < mt2_end && !HFG_DBG(a, 8); mt += mt2_step) {
                  for (int pp = 0; pp < PP; ++pp) {                     // the time steps of this GEMM row
                    const int lr = (mt * 128 + row) * PP + pp;          // output row inside the tile
                    const int t = t0 + lr;
                    const bool add_prev = add_prev_mode && real && lr < a.TO && t < a.T;
                    const uint8_t* rp = sa_p + (size_t)(lr + a.p2 + a.p1) * 16;
                    const uint8_t* accp = reinterpret_cast<const uint8_t*>(a.acc) + (long long)b * a.acc_bstride +
                                          (long long)(kPadL + t) * 16;
                    const uint32_t tbase = acc2 + lane_sel + (uint32_t)(mt * N2 + pp * N + kb * KBC * CW);
                    const long long sum_b = LO ? a.o32_bstride : a.o_bstride, sum_p = LO ? a.o32_pstride : a.o_pstride;
                    const uint8_t* sump = a.n_sum ? a.sum_in[0] + (long long)b * sum_b + (long long)(kPadL + t) * 16 : nullptr;
                    const uint8_t* lp = sH + (size_t)(kb * KBC) * RL * 16 + (size_t)lr * 16;   // lo twin of row lr (LO kernels)
                    for (int c16 = 0; c16 < nck * CW; c16 += 16) {              // 16 columns at a time
                        const int col = kb * KBC * CW + c16;                    // channel; acc2 column = pp * N + col
                        // column ownership must match epi2 (32-column groups of acc2 alternate between the two
                        // warps of a lane quarter): this warp's tcgen05.st for tile i+1 may only touch columns
                        // whose tile-i values it has itself already read in epi2
                        if (split2 && (((pp * N + col) >> 5) & 1) != half) continue;
                        if (!own2(mt, pp * N + col)) continue;
                        float v[16];
                        if constexpr (LO)
                            load_split16(rp + (long long)(c16 / CW) * a_plane, a_plane,
                                         lp + (long long)(c16 / CW) * ((long long)RL * 16), (long long)RL * 16, v);
                        else
                            load_cells16<P>(rp + (long long)(c16 / CW) * a_plane, a_plane, v);
                        float prev[16];
                        if (add_prev && a.n_sum == 0) load_f32x16(accp + (long long)(col / 4) * a.acc_pstride, a.acc_pstride, prev);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = lrelu_inv(v[i], inv_slope);
                        add_bias16(v, sB2 + col);
                        if (add_prev && a.n_sum == 0) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] += prev[i];
                        }
                        if (add_prev && a.n_sum > 0) {                          // + rb_j(x) for the other resblocks
                            for (int s = 0; s < a.n_sum; ++s) {
                                if constexpr (LO) load_f32x16(sump + (a.sum_in[s] - a.sum_in[0]) + (long long)(col / 4) * sum_p, sum_p, prev);
                                else load_cells16<P>(sump + (a.sum_in[s] - a.sum_in[0]) + (long long)(col / CW) * sum_p, sum_p, prev);
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] += lrelu_inv(prev[i], inv_slope);
                            }
                        }
                        tmem_st16(tbase + (uint32_t)c16, v);
                    }
                  }
                }
                tmem_st_wait();
                __syncwarp();
                if (lane == 0) mbar_arrive(A_EMPTY(sa_i));
                if (++sa_i == a.sa) { sa_i = 0; sa_ph ^= 1; }
            }
            // LO kernels: every epilogue warp has read its lo cells before any of them overwrites the buffer with H
            if constexpr (LO) asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
            // ---------- epi1: acc1 -> leaky_relu(. + b1) -> H tile in smem ----------
            if (e == 0) HFG_TL(6, it);
            mbar_wait_sleep(acc1_full, it & 1, (uint32_t)a.epi_sleep_ns);
            tc_fence_after();
            if (e == 0) HFG_TL(7, it);
            // only tiles that touch an utterance edge have H rows outside [0, T) to zero
            const bool edge_tile = (t0 - a.p2 < 0) || (t0 - a.p2 + MT * 128 > a.T);
            for (int mt = mt_first; mt < mt_end && !HFG_DBG(a, 8); mt += mt_step) {
                const int hr = mt * 128 + row;                          // H row inside the tile
                const int th = t0 - a.p2 + hr;                          // its time step
                const bool drop = edge_tile && !(th >= 0 && th < a.T);  // conv2 zero-pads ITS input
                // H row hr of the tile: row hr of plane chunk, or -- space-to-depth -- row hr / 2 of plane
                // (hr % 2) * (N / CW2) + chunk
                uint8_t* hp = a.s2d ? sH + (size_t)(hr >> 1) * 16 + (size_t)(hr & 1) * (N / CW2) * h_plane : sH + (size_t)hr * 16;
                const uint32_t tbase = acc1 + lane_sel + (uint32_t)(mt * N);
                for (int c0 = 32 * ch; c0 < N; c0 += 32 * cs) {
                    if (!own1(mt, c0)) continue;
                    uint32_t r0[16], r1[16];
                    const bool two = c0 + 16 < N;
                    tmem_ld16(tbase + (uint32_t)c0, r0);
                    if (two) tmem_ld16(tbase + (uint32_t)(c0 + 16), r1);
                    tmem_ld_wait();
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r0[i]);
                    add_bias16(v, sB1 + c0);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = drop ? 0.f : lrelu(v[i], slope);
                    store_cells16<P2>(hp + (long long)(c0 / CW2) * h_plane, h_plane, v);
                    if (two) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r1[i]);
                        add_bias16(v, sB1 + c0 + 16);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = drop ? 0.f : lrelu(v[i], slope);
                        store_cells16<P2>(hp + (long long)((c0 + 16) / CW2) * h_plane, h_plane, v);
                    }
                }
            }
            fence_async_smem();          // H (generic-proxy writes) must be visible to the MMA (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTAS == 2 && rank == 1) mbar_arrive_remote(h_ready, 0);
                else mbar_arrive(h_ready);
            }
            // ---------- epi2: acc2 -> global ----------
            if (e == 0) HFG_TL(8, it);
            mbar_wait_sleep(acc2_full, it & 1, (uint32_t)a.epi_sleep_ns);
            tc_fence_after();
            if (e == 0) HFG_TL(9, it);
            // space-to-depth form with both time steps of a GEMM row in this thread: rows 2m and 2m + 1 of a chunk plane
            // are adjacent, so every cell pair leaves as ONE 32-byte store (lanes 32 bytes apart: full sectors) instead
            // of two 16-byte stores with a 32-byte lane stride
            // (compiled only into the single-CTA 2-byte variants, the ones the space-to-depth rule selects: the CTA-pair
            // and tf32 variants run at their register cap and would pay for the extra path with spills)
            const bool rows2 = CTAS == 1 && P != PREC_TF32 && !W3 && a.s2d && !split2 && !acc_store_mode;
            if constexpr (CTAS == 1 && P != PREC_TF32)
            for (int mt = mt2_first; mt < mt2_end && !HFG_DBG(a, 8) && rows2; mt += mt2_step) {
                const uint32_t tbase = acc2 + lane_sel + (uint32_t)(mt * N2);
                const int lr = (mt * 128 + row) * 2;
                const int t = t0 + lr;
                const bool v0 = real && lr < a.TO && t < a.T, v1 = real && lr + 1 < a.TO && t + 1 < a.T;
                const long long row_bytes = (long long)(kPadL + t) * 16;
                uint8_t* op = a.out + (long long)b * a.o_bstride + row_bytes;
                for (int cb = 0; cb < N; cb += 8) {          // 8 channels of both rows at a time (register budget of the MINB = 2 variants)
                    uint32_t r0[8], r1[8];
                    tmem_ld8(tbase + (uint32_t)cb, r0);
                    tmem_ld8(tbase + (uint32_t)(N + cb), r1);
                    tmem_ld_wait();
                    if (!v0 && !v1) continue;
                    float x0[8], x1[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { x0[i] = __uint_as_float(r0[i]); x1[i] = __uint_as_float(r1[i]); }
                    if (a.acc_mode == TC_ACC_FINAL) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { x0[i] *= inv_div; x1[i] *= inv_div; }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { x0[i] = lrelu(x0[i], slope); x1[i] = lrelu(x1[i], slope); }
                    if (LO && a.out32) {
                        uint8_t* p32 = a.out32 + (long long)b * a.o32_bstride + row_bytes + (long long)(cb / 4) * a.o32_pstride;
#pragma unroll
                        for (int g = 0; g < 2; ++g)
                            store_cell_rows2(p32 + g * a.o32_pstride, floats_to_cell<PREC_TF32>(x0 + 4 * g),
                                             floats_to_cell<PREC_TF32>(x1 + 4 * g), v0, v1);
                    } else if (LO && a.out_lo) {
                        uint4 h0, l0, h1, l1;
                        split16(x0, h0, l0);
                        split16(x1, h1, l1);
                        store_cell_rows2(op + (long long)(cb / CW) * a.o_pstride, h0, h1, v0, v1);
                        store_cell_rows2(a.out_lo + (long long)b * a.o_bstride + row_bytes + (long long)(cb / CW) * a.o_pstride, l0, l1, v0, v1);
                    } else {
                        uint8_t* ph = op + (long long)(cb / CW) * a.o_pstride;
#pragma unroll
                        for (int g = 0; g < 8 / CW; ++g)
                            store_cell_rows2(ph + g * a.o_pstride, floats_to_cell<P>(x0 + CW * g), floats_to_cell<P>(x1 + CW * g), v0, v1);
                    }
                }
            }
            for (int mt = mt2_first; mt < mt2_end && !HFG_DBG(a, 8) && !rows2; mt += mt2_step) {
                const uint32_t tbase = acc2 + lane_sel + (uint32_t)(mt * N2);
                for (int c0 = 32 * ch2; c0 < N2; c0 += 32 * cs2) {
                    if (!own2(mt, c0)) continue;
                    const int pp = c0 >= N ? 1 : 0, cb = c0 - pp * N;   // time step of this GEMM row, channel base (N % 32 == 0 when PP = 2)
                    const int lr = (mt * 128 + row) * PP + pp;
                    const int t = t0 + lr;
                    const bool valid = real && lr < a.TO && t < a.T;
                    const long long row_bytes = (long long)(kPadL + t) * 16;
                    uint8_t* ap = reinterpret_cast<uint8_t*>(a.acc) + (long long)b * a.acc_bstride + row_bytes;
                    uint8_t* op = a.out + (long long)b * a.o_bstride + row_bytes;
                    uint32_t r0[16], r1[16];
                    const bool two = cb + 16 < N;
                    tmem_ld16(tbase + (uint32_t)c0, r0);
                    if (two) tmem_ld16(tbase + (uint32_t)(c0 + 16), r1);
                    tmem_ld_wait();
                    if (!valid) continue;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        if (hh == 1 && !two) break;
                        const int cc = cb + 16 * hh;
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(hh ? r1[i] : r0[i]);
                        if (acc_store_mode) {
                            store_f32x16(ap + (long long)(cc / 4) * a.acc_pstride, a.acc_pstride, v);
                        } else {
                            if (a.acc_mode == TC_ACC_FINAL) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] *= inv_div;
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = lrelu(v[i], slope);
                            if constexpr (LO) {
                                if (a.out32)
                                    store_f32x16(a.out32 + (long long)b * a.o32_bstride + row_bytes + (long long)(cc / 4) * a.o32_pstride,
                                                 a.o32_pstride, v);
                                else if (a.out_lo)
                                    store_split16(op + (long long)(cc / CW) * a.o_pstride,
                                                  a.out_lo + (long long)b * a.o_bstride + row_bytes + (long long)(cc / CW) * a.o_pstride,
                                                  a.o_pstride, v);
                                else store_cells16<P>(op + (long long)(cc / CW) * a.o_pstride, a.o_pstride, v);
                            } else {
                                store_cells16<P>(op + (long long)(cc / CW) * a.o_pstride, a.o_pstride, v);
                            }
                        }
                    }
                }
            }
            tc_fence_before();           // order these TMEM reads before the next tile's tcgen05.st / MMA
            if (e == 0) HFG_TL(10, it);
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CTAS == 2) cluster_sync_all();      // nobody exits while the peer may still signal / read it
    if (warp == 1) {
        tc_fence_after();
        if constexpr (CTAS == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

}  // namespace hfg
