/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Plain-C CPU restatement of the reference HiFi-GAN generator forward pass
 * (reference: models/hifigan.py:224-261).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this library; the product
 * (tts-sambert_hifigan_b200/) never links, imports or calls it.
 *
 * Parity status: the reference's own tests pin no output VALUE for this path
 *  and ship no golden vectors, so this file is pinned
 * against the live reference module instead: tests/golden/make_golden.py runs
 * models.hifigan.HiFiGANGenerator from /root/reference on seeded inputs and
 * commits its outputs under tests/golden/; tests/test_oracle.py checks this
 * restatement against those vectors.
 *
 * Arithmetic: every layer accumulates in double and stores float, layer by
 * layer, exactly where the reference materialises an fp32 tensor.  The
 * arithmetic itself lives in PyTorch ATen (conv1d / conv_transpose1d /
 * leaky_relu / tanh; torch 2.11.0 is the effective pin, requirements.txt:2 says
 * torch>=2.0.0); what is restated here is the published definition of those
 * operators.
 *
 * Layout: activations [B, C, T] row-major (time fastest), as the reference.
 * Weights arrive as an array of pointers in state_dict registration order
 * (see weight_shapes() in tts-sambert_hifigan_b200/synth.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HFGO_MAX 8

typedef struct {
    int32_t n_mels;
    int32_t n_up;                 /* number of upsample stages            */
    int32_t c0;                   /* upsample_initial_channel             */
    int32_t n_rk;                 /* number of resblocks per MRF          */
    int32_t up_rates[HFGO_MAX];
    int32_t up_ks[HFGO_MAX];
    int32_t rk[HFGO_MAX];         /* resblock kernel sizes                */
    int32_t n_dil[HFGO_MAX];      /* dilations per resblock               */
    int32_t dil[HFGO_MAX][HFGO_MAX];
} hfgo_config;

/* reference models/hifigan.py:21-23 */
static int same_padding(int k, int d) { return (k * d - d) / 2; }

/* reference models/hifigan.py:81,83,244,254: F.leaky_relu(x, 0.1) */
static void leaky_relu(const float* x, float* y, size_t n) {
    for (size_t i = 0; i < n; ++i) y[i] = x[i] > 0.0f ? x[i] : 0.1f * x[i];
}

/* nn.Conv1d(cin, cout, k, stride=1, dilation=d, padding=p), weight [cout,cin,k]
 * (call sites: reference models/hifigan.py:52-69,177-183,216-222) */
static void conv1d(const float* x, int B, int cin, int T, const float* w, const float* bias,
                   int cout, int k, int d, int p, float* y) {
    #pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int co = 0; co < cout; ++co) {
            double* acc = (double*)malloc(sizeof(double) * (size_t)T);
            for (int t = 0; t < T; ++t) acc[t] = (double)bias[co];
            for (int ci = 0; ci < cin; ++ci) {
                const float* xr = x + ((size_t)b * cin + ci) * T;
                const float* wr = w + ((size_t)co * cin + ci) * k;
                for (int j = 0; j < k; ++j) {
                    const double wv = (double)wr[j];
                    const int off = j * d - p;
                    int lo = off < 0 ? -off : 0;
                    int hi = T - off < T ? T - off : T;
                    for (int t = lo; t < hi; ++t) acc[t] += wv * (double)xr[t + off];
                }
            }
            float* yr = y + ((size_t)b * cout + co) * T;
            for (int t = 0; t < T; ++t) yr[t] = (float)acc[t];
            free(acc);
        }
}

/* nn.ConvTranspose1d(cin, cout, k, stride=u, padding=p), weight [cin,cout,k]
 * (reference models/hifigan.py:196-202):  y[co, s*u - p + j] += x[ci, s] * w[ci, co, j] */
static int convt_out_len(int T, int k, int u, int p) { return (T - 1) * u - 2 * p + k; }

static void conv_transpose1d(const float* x, int B, int cin, int T, const float* w,
                             const float* bias, int cout, int k, int u, int p, float* y) {
    const int To = convt_out_len(T, k, u, p);
    #pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int co = 0; co < cout; ++co) {
            double* acc = (double*)malloc(sizeof(double) * (size_t)To);
            for (int t = 0; t < To; ++t) acc[t] = (double)bias[co];
            for (int ci = 0; ci < cin; ++ci) {
                const float* xr = x + ((size_t)b * cin + ci) * T;
                const float* wr = w + ((size_t)ci * cout + co) * k;
                for (int s = 0; s < T; ++s) {
                    const double xv = (double)xr[s];
                    for (int j = 0; j < k; ++j) {
                        const int t = s * u - p + j;
                        if (t >= 0 && t < To) acc[t] += xv * (double)wr[j];
                    }
                }
            }
            float* yr = y + ((size_t)b * cout + co) * To;
            for (int t = 0; t < To; ++t) yr[t] = (float)acc[t];
            free(acc);
        }
}

int64_t hfgo_out_len(const hfgo_config* c, int T) {
    int64_t t = T;
    for (int i = 0; i < c->n_up; ++i)
        t = convt_out_len((int)t, c->up_ks[i], c->up_rates[i], (c->up_ks[i] - c->up_rates[i]) / 2);
    return t;
}

/* Number of weight tensors the caller must pass, in state_dict order. */
int hfgo_num_weights(const hfgo_config* c) {
    int n = 2 + 2;
    for (int i = 0; i < c->n_up; ++i) {
        n += 2;
        for (int j = 0; j < c->n_rk; ++j) n += 4 * c->n_dil[j];
    }
    return n;
}

/*
 * stage_out (may be NULL): 2*n_up+1 optional float buffers receiving
 *   [0] conv_pre output, [1+2i] ups[i] output, [2+2i] mrfs[i] output.
 * Returns 0, or -1 on allocation failure.
 */
int hfgo_forward(const hfgo_config* c, const float* const* wts, const float* mel, int B, int T,
                 float* wav, float* const* stage_out) {
    int wi = 0;
    const float* pre_w = wts[wi++];
    const float* pre_b = wts[wi++];
    const int ups_base = wi;
    wi += 2 * c->n_up;
    const int mrf_base = wi;

    int C = c->c0;
    size_t n = (size_t)B * C * T;
    float* x = (float*)malloc(sizeof(float) * n);
    if (!x) return -1;
    /* reference models/hifigan.py:238 */
    conv1d(mel, B, c->n_mels, T, pre_w, pre_b, C, 7, 1, 3, x);
    if (stage_out && stage_out[0]) memcpy(stage_out[0], x, sizeof(float) * n);

    int Tc = T;
    int mw = mrf_base;
    for (int i = 0; i < c->n_up; ++i) {
        const int u = c->up_rates[i], k = c->up_ks[i], p = (k - u) / 2;
        const int Co = C / 2, To = convt_out_len(Tc, k, u, p);
        /* reference models/hifigan.py:244-245 */
        leaky_relu(x, x, (size_t)B * C * Tc);
        size_t no = (size_t)B * Co * To;
        float* y = (float*)malloc(sizeof(float) * no);
        if (!y) return -1;
        conv_transpose1d(x, B, C, Tc, wts[ups_base + 2 * i], wts[ups_base + 2 * i + 1], Co, k, u, p, y);
        free(x);
        x = y; C = Co; Tc = To;
        if (stage_out && stage_out[1 + 2 * i]) memcpy(stage_out[1 + 2 * i], x, sizeof(float) * no);

        /* MRF: reference models/hifigan.py:116-131 */
        float* sum = (float*)calloc(no, sizeof(float));
        float* r = (float*)malloc(sizeof(float) * no);
        float* a = (float*)malloc(sizeof(float) * no);
        float* h = (float*)malloc(sizeof(float) * no);
        if (!sum || !r || !a || !h) return -1;
        for (int j = 0; j < c->n_rk; ++j) {
            const int rk = c->rk[j], nd = c->n_dil[j];
            /* state_dict order inside a ResBlock: convs1.0..nd-1 then convs2.0..nd-1 */
            const float* const* w1 = wts + mw;
            const float* const* w2 = wts + mw + 2 * nd;
            mw += 4 * nd;
            memcpy(r, x, sizeof(float) * no);
            /* ResBlock: reference models/hifigan.py:80-85 */
            for (int l = 0; l < nd; ++l) {
                const int d = c->dil[j][l];
                leaky_relu(r, a, no);
                conv1d(a, B, C, Tc, w1[2 * l], w1[2 * l + 1], C, rk, d, same_padding(rk, d), h);
                leaky_relu(h, h, no);
                conv1d(h, B, C, Tc, w2[2 * l], w2[2 * l + 1], C, rk, 1, same_padding(rk, 1), a);
                for (size_t e = 0; e < no; ++e) r[e] = r[e] + a[e];
            }
            /* reference :126-129: output = rb0(x); output = output + rb_j(x) */
            if (j == 0) memcpy(sum, r, sizeof(float) * no);
            else for (size_t e = 0; e < no; ++e) sum[e] = sum[e] + r[e];
        }
        /* reference :131: output / len(resblocks) */
        const float div = (float)c->n_rk;
        for (size_t e = 0; e < no; ++e) x[e] = sum[e] / div;
        free(sum); free(r); free(a); free(h);
        if (stage_out && stage_out[2 + 2 * i]) memcpy(stage_out[2 + 2 * i], x, sizeof(float) * no);
    }
    /* reference models/hifigan.py:254-256 */
    leaky_relu(x, x, (size_t)B * C * Tc);
    conv1d(x, B, C, Tc, wts[mw], wts[mw + 1], 1, 7, 1, 3, wav);
    for (size_t e = 0; e < (size_t)B * Tc; ++e) wav[e] = tanhf(wav[e]);
    free(x);
    return 0;
}
