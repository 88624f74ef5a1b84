// Persistent implicit-GEMM convolution: conv_pre, the polyphase ConvTranspose1d upsamplers and the
// ResBlock convolutions that do not fit the fused pair kernel (reference models/hifigan.py:80-85,238,245).
//
// Same operand layout and tap-shift trick as tc_conv_kernel (tc_kernels.cuh), but:
//   * one CTA per SM loops over work items (tile, phase, channel tile) instead of one CTA per item, so the
//     barrier / TMEM set-up and the first activation load are paid once per SM, not once per tile;
//   * the accumulator is double-buffered in TMEM: while the 8 epilogue warps drain item i (TMEM -> bias ->
//     leaky_relu -> polyphase-interleaved 16-byte cells), the MMA warp already runs item i+1 and the
//     producer streams the operands of item i+2.  The per-CTA timeline of the one-shot kernel showed
//     load -> MMA -> epilogue strictly serial with the epilogue as long as the MMAs
//     (profiles/r1_tuning.md section 11).
#pragma once
#include "tc_kernels.cuh"
#include "tc_pair_kernel.cuh"   // mbar_arrive

namespace hfg {

constexpr int kUpMaxSA = 4, kUpMaxSW = 8;

struct TcUpArgs {
    TcConvArgs c;        // geometry, operands, epilogue parameters (res / acc unused)
    int n_items;         // tiles * phases * channel tiles
    int pn_per_tile;     // phases * channel tiles: items of one tile are adjacent (they share the activation tile)
    int n_ctile;         // channel tiles per phase
    int cout_total;      // bias entries staged in shared memory (virtual channels when phases are stacked)
    int epi_sleep_ns;    // epilogue warps: longest sleep between polls of the accumulator barrier (0 = spin)
    int contig;          // > 0: CTA c runs the contiguous items [c * contig, (c + 1) * contig); 0: items c, c + grid, ...
};

template <int P>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_up_kernel(const TcUpArgs ua) {
    const TcConvArgs& a = ua.c;
    extern __shared__ __align__(128) uint8_t tc_up_smem[];
    uint8_t* smem = tc_up_smem;
    constexpr int CW = Prec<P>::CW;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = a.N, MT = a.MT, R = a.R;
    const int n_chunks = a.a_nchunks;
    const int nck_max = n_chunks < 8 ? n_chunks : 8;
    const int n_kb = (n_chunks + 7) / 8;
    const uint32_t a_stage_bytes = (uint32_t)R * nck_max * 16;
    const int G = a.tap_group;
    const uint32_t w_stage_bytes = (uint32_t)G * N * nck_max * 16;
    uint8_t* sA = smem;
    uint8_t* sW = sA + (size_t)a.sa * a_stage_bytes;
    float* sBias = reinterpret_cast<float*>(sW + (size_t)a.sw * w_stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + ((ua.cout_total + 3) & ~3));
    const uint32_t bar0 = smem_u32(bars);
    auto A_FULL = [&](int i) { return bar0 + 8u * i; };
    auto A_EMPTY = [&](int i) { return bar0 + 8u * (kUpMaxSA + i); };
    auto W_FULL = [&](int i) { return bar0 + 8u * (2 * kUpMaxSA + i); };
    auto W_EMPTY = [&](int i) { return bar0 + 8u * (2 * kUpMaxSA + kUpMaxSW + i); };
    auto ACC_FULL = [&](int i) { return bar0 + 8u * (2 * kUpMaxSA + 2 * kUpMaxSW + i); };
    auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (2 * kUpMaxSA + 2 * kUpMaxSW + 2 + i); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kUpMaxSA + 2 * kUpMaxSW + 4);
    const int n_epi_warps = ((int)blockDim.x - 64) / 32;

    uint32_t ncols = 32;
    while ((int)ncols < 2 * MT * N) ncols <<= 1;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < a.sa; ++i) { mbar_init(A_FULL(i), 1); mbar_init(A_EMPTY(i), 1); }
        for (int i = 0; i < a.sw; ++i) { mbar_init(W_FULL(i), 1); mbar_init(W_EMPTY(i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(ACC_FULL(i), 1); mbar_init(ACC_EMPTY(i), (uint32_t)n_epi_warps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < ua.cout_total; i += (int)blockDim.x - 64)
            sBias[i] = a.bias[a.stack_cout ? i % a.stack_cout : i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work items of this CTA: item, item + gridDim.x, ...
    const int item0 = ua.contig ? (int)blockIdx.x * ua.contig : (int)blockIdx.x, item_step = ua.contig ? 1 : (int)gridDim.x;
    const int item_end = ua.contig ? (item0 + ua.contig < ua.n_items ? item0 + ua.contig : ua.n_items) : ua.n_items;
    auto taps_of = [&](int phase) { return a.phases > 1 ? (a.k - phase + a.u - 1) / a.u : a.taps_max; };
    // variable-length batches (a.len_rows: OUTPUT rows utterance b needs, or null): an item whose tile starts at
    // or beyond the last q position that utterance needs is skipped by every role alike
    auto item_live = [&](int item) {
        if (!a.len_rows) return true;
        const int tile = item / ua.pn_per_tile;
        const int q0 = (tile % a.tiles_per_batch) * MT * 128;
        return q0 < tc_len_nq(a, a.len_rows[tile / a.tiles_per_batch]);
    };
    auto next_live = [&](int item) { while (item < item_end && !item_live(item)) item += item_step; return item; };

    if (warp == 0) {
        // ===================== producer: two independent streams (activation K blocks, weight stages) =====================
        const bool leader = elect_one();
        int sa_i = 0, sa_ph = 0, sw_i = 0, sw_ph = 0;
        int a_item = next_live(item0), a_kb = 0;            // next activation block
        int w_item = a_item, w_kb = 0, w_tap0 = 0;          // next weight stage
        uint32_t idle = 0;
        long long t_idle0 = 0;
        while (a_item < item_end || w_item < item_end) {
            bool did = false;
            if (a_item < item_end && mbar_test(A_EMPTY(sa_i), sa_ph ^ 1)) {
                const int tile = a_item / ua.pn_per_tile;
                const int b = tile / a.tiles_per_batch;
                const int q0 = (tile % a.tiles_per_batch) * MT * 128;
                const int nck = (n_chunks - 8 * a_kb) < 8 ? (n_chunks - 8 * a_kb) : 8;
                if (leader) {
                    const uint8_t* ab = a.a + (long long)b * a.a_bstride + (long long)(kPadL + q0 + a.min_off) * 16;
                    mbar_expect_tx(A_FULL(sa_i), (uint32_t)nck * R * 16);
                    const uint32_t dst = smem_u32(sA + (size_t)sa_i * a_stage_bytes);
                    for (int c = 0; c < nck; ++c)
                        bulk_g2s(dst + (uint32_t)c * R * 16, ab + (long long)(8 * a_kb + c) * a.a_pstride,
                                 (uint32_t)R * 16, A_FULL(sa_i));
                }
                __syncwarp();
                if (++sa_i == a.sa) { sa_i = 0; sa_ph ^= 1; }
                if (++a_kb == n_kb) { a_kb = 0; a_item = next_live(a_item + item_step); }
                did = true;
            }
            if (w_item < item_end && mbar_test(W_EMPTY(sw_i), sw_ph ^ 1)) {
                const int pn = w_item % ua.pn_per_tile;
                const int phase = pn / ua.n_ctile, ntile = pn % ua.n_ctile;
                const int taps = taps_of(phase);
                const int nck = (n_chunks - 8 * w_kb) < 8 ? (n_chunks - 8 * w_kb) : 8;
                const int g = (taps - w_tap0) < G ? (taps - w_tap0) : G;
                if (leader) {
                    const uint8_t* wb = a.w + (long long)phase * a.w_phase_stride + (long long)ntile * a.w_ntile_stride;
                    mbar_expect_tx(W_FULL(sw_i), (uint32_t)g * nck * N * 16);
                    bulk_g2s(smem_u32(sW + (size_t)sw_i * w_stage_bytes),
                             wb + ((long long)w_kb * a.taps_max * 8 + (long long)w_tap0 * nck) * N * 16,
                             (uint32_t)g * nck * N * 16, W_FULL(sw_i));
                }
                __syncwarp();
                if (++sw_i == a.sw) { sw_i = 0; sw_ph ^= 1; }
                w_tap0 += G;
                if (w_tap0 >= taps) { w_tap0 = 0; if (++w_kb == n_kb) { w_kb = 0; w_item = next_live(w_item + item_step); } }
                did = true;
            }
            if (did) { idle = 0; t_idle0 = 0; continue; }
            __nanosleep(40);
            if ((++idle & 4095u) == 0) {                         // bounded: trap instead of hanging the GPU
                const long long now = clock64();
                if (t_idle0 == 0) t_idle0 = now;
                else if (now - t_idle0 > 4000000000ll) __trap();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc<P>(N);
        const uint32_t a_hi = (128u >> 4) | (1u << 14);                     // SBO = 128 B, version 1
        const uint32_t a_lbo = ((uint32_t)R) << 16, b_lbo = ((uint32_t)N) << 16;
        int sa_i = 0, sa_ph = 0, sw_i = 0, sw_ph = 0;
        uint32_t it = 0;
        for (int item = next_live(item0); item < item_end; item = next_live(item + item_step), ++it) {
            const int pn = item % ua.pn_per_tile;
            const int taps = taps_of(pn / ua.n_ctile);
            const uint32_t buf = it & 1u;
            const uint32_t acc = tmem_base + buf * (uint32_t)(MT * N);
            mbar_wait(ACC_EMPTY(buf), ((it >> 1) & 1u) ^ 1u);              // epilogue of item it-2 has drained this buffer
            tc_fence_after();
            uint32_t acc_on = 0;
            for (int kb = 0; kb < n_kb; ++kb) {
                const int nck = (n_chunks - 8 * kb) < 8 ? (n_chunks - 8 * kb) : 8;
                const int ksteps = nck >> 1;
                mbar_wait(A_FULL(sa_i), sa_ph);
                tc_fence_after();
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
                                umma_ksteps<P>(acc + (uint32_t)(mt * N), a_hi, a_lo1 + (uint32_t)(mt * 128), b_lo,
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
            if (leader) tc_commit(ACC_FULL(buf));
            __syncwarp();
        }
    } else {
        // ===================== epilogue: TMEM -> bias (+ residual, MRF sum) -> leaky_relu -> operand dtype -> global =====================
        // same arithmetic, in the same order, as tc_conv_kernel's epilogue (bit-identical results)
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;                   // the two warps of a quarter alternate 32-column steps
        const int qlane = quarter * 32 + lane;
        const float slope = a.slope, inv_slope = 1.0f / a.slope;
        const float inv_div = 1.0f / a.div;
        const bool add_prev = a.acc_mode == TC_ACC_ADD || a.acc_mode == TC_ACC_FINAL;
        const bool acc_store = a.acc_mode == TC_ACC_WRITE || a.acc_mode == TC_ACC_ADD;
        const int col_step = 32 * (n_epi_warps / 4);
        uint32_t it = 0;
        for (int item = next_live(item0); item < item_end; item = next_live(item + item_step), ++it) {
            const int tile = item / ua.pn_per_tile, pn = item % ua.pn_per_tile;
            const int phase = pn / ua.n_ctile, ntile = pn % ua.n_ctile;
            const int b = tile / a.tiles_per_batch;
            const int q0 = (tile % a.tiles_per_batch) * MT * 128;
            const uint32_t buf = it & 1u;
            mbar_wait_sleep(ACC_FULL(buf), (it >> 1) & 1u, (uint32_t)ua.epi_sleep_ns);
            tc_fence_after();
            for (int mt = 0; mt < MT; ++mt) {
                const int q = q0 + mt * 128 + qlane;
                const uint32_t tbase = tmem_base + buf * (uint32_t)(MT * N) + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * N);
                for (int c0 = 32 * half; c0 < N; c0 += col_step) {
                    const bool two = c0 + 16 < N;
                    uint32_t r0[16], r1[16];
                    tmem_ld16(tbase + (uint32_t)c0, r0);
                    if (two) tmem_ld16(tbase + (uint32_t)(c0 + 16), r1);
                    float x0[16], x1[16], p0[16], p1[16];
                    const int vc0 = ntile * N + c0;             // first (virtual) output channel of this step
                    int ch0 = vc0, ph = phase;
                    if (a.stack_cout) { ph = vc0 / a.stack_cout; ch0 = vc0 - ph * a.stack_cout; }   // a step never straddles phases
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
                        const int ch = ch0 + 16 * hh;
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(hh ? r1[i] : r0[i]);
                        add_bias16(v, sBias + vc0 + 16 * hh);
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
                                if (a.out_lo) {                  // fp16 hi + lo pair (tf32 mode on fp16 operand planes)
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
            tc_fence_before();                                  // TMEM reads before the MMA warp reuses this buffer
            __syncwarp();
            if (lane == 0) mbar_arrive(ACC_EMPTY(buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

}  // namespace hfg
