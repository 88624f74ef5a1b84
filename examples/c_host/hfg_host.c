/*
 * hfg_host.c -- the generator driven from a plain-C host through include/hfg.h alone
 * (no Python, no torch, no CUDA headers): what a compiled-language maintainer binds.
 *
 * The reference's boundary is a Python class (models/hifigan.py:149-261); this program is the same
 * life cycle spelt through the C ABI:
 *   HiFiGANGenerator(**kwargs)   -> hfg_create            (reference models/hifigan.py:149-222)
 *   load_state_dict(sd)          -> hfg_set_weight x N    (156- or 232-key schema)
 *   remove_weight_norm()         -> hfg_commit_weights    (reference :263-272)
 *   wav = generator(mel)         -> hfg_forward_host      (reference :224-261; host buffers in and out)
 *
 * usage: hfg_host <dir> [mode]        mode: 0 fp32, 1 tf32 (default), 2 bf16, 3 fp16
 * <dir>/manifest.txt (written by tests/test_zz_c_host.py or by any exporter):
 *   cfg <n_mels> <initial_channel> <n_ups> <rate>*n_ups <kernel>*n_ups <n_rb> { <k> <n_dil> <dil>*n_dil }*n_rb
 *   mel <B> <T> <file>                                  raw little-endian fp32 [B, n_mels, T]
 *   w <state_dict key> <ndim> <d0> [<d1> [<d2>]] <file>  raw fp32 in the reference's own layout
 * writes <dir>/wav.bin (fp32 [B, 1, T_out]) and prints one line "ok B T_out launches n_weights sum".
 *
 * Exit codes: 0 ok, 2 bad input files, 3 the library refused (message printed) -- in particular
 * without a CUDA device hfg_create fails with HFG_ERR_CUDA: there is no CPU fallback.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hfg.h"

static float* read_floats(const char* dir, const char* file, size_t n) {
    char path[4096];
    snprintf(path, sizeof path, "%s/%s", dir, file);
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); return NULL; }
    float* p = (float*)malloc(n ? n * sizeof(float) : 1);
    size_t got = p ? fread(p, sizeof(float), n, f) : 0;
    fclose(f);
    if (got != n) { fprintf(stderr, "%s: wanted %zu floats, got %zu\n", path, n, got); free(p); return NULL; }
    return p;
}

static const char* status_name(int rc) {
    switch (rc) {
        case HFG_OK: return "HFG_OK";
        case HFG_ERR_INVALID: return "HFG_ERR_INVALID";
        case HFG_ERR_CUDA: return "HFG_ERR_CUDA";
        case HFG_ERR_STATE: return "HFG_ERR_STATE";
        case HFG_ERR_WORKSPACE: return "HFG_ERR_WORKSPACE";
        case HFG_ERR_UNSUPPORTED: return "HFG_ERR_UNSUPPORTED";
        default: return "unknown status";
    }
}

static int fail(hfg_handle* h, const char* what, int rc) {
    fprintf(stderr, "%s failed: %d %s (%s)\n", what, rc, status_name(rc),
            h ? hfg_last_error(h) : "no handle was created: no CUDA device, and there is no CPU fallback");
    if (h) hfg_destroy(h);
    return 3;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <dir> [mode]\n", argv[0]); return 2; }
    const char* dir = argv[1];
    const int mode = argc > 2 ? atoi(argv[2]) : HFG_MODE_TF32;

    char path[4096];
    snprintf(path, sizeof path, "%s/manifest.txt", dir);
    FILE* mf = fopen(path, "r");
    if (!mf) { fprintf(stderr, "cannot open %s\n", path); return 2; }

    /* --- HiFiGANGenerator.__init__ ------------------------------------------------------------ */
    hfg_config cfg;
    memset(&cfg, 0, sizeof cfg);
    char tag[16];
    if (fscanf(mf, "%15s %d %d %d", tag, &cfg.n_mels, &cfg.upsample_initial_channel, &cfg.num_upsamples) != 4 ||
        strcmp(tag, "cfg") != 0 || cfg.num_upsamples < 1 || cfg.num_upsamples > HFG_MAX_STAGES) {
        fprintf(stderr, "manifest: bad cfg line\n"); return 2;
    }
    for (int i = 0; i < cfg.num_upsamples; ++i) if (fscanf(mf, "%d", &cfg.upsample_rates[i]) != 1) return 2;
    for (int i = 0; i < cfg.num_upsamples; ++i) if (fscanf(mf, "%d", &cfg.upsample_kernel_sizes[i]) != 1) return 2;
    if (fscanf(mf, "%d", &cfg.num_resblocks) != 1 || cfg.num_resblocks < 1 || cfg.num_resblocks > HFG_MAX_STAGES) return 2;
    for (int j = 0; j < cfg.num_resblocks; ++j) {
        if (fscanf(mf, "%d %d", &cfg.resblock_kernel_sizes[j], &cfg.num_dilations[j]) != 2 ||
            cfg.num_dilations[j] < 1 || cfg.num_dilations[j] > HFG_MAX_STAGES) return 2;
        for (int l = 0; l < cfg.num_dilations[j]; ++l) if (fscanf(mf, "%d", &cfg.resblock_dilations[j][l]) != 1) return 2;
    }

    /* host-only entry points: these answer on a box without a GPU too */
    int32_t radius = -1;
    int rc = hfg_receptive_radius(&cfg, &radius);
    printf("abi %d (header %d) receptive_radius %d rc %d\n", hfg_abi_version(), HFG_ABI_VERSION, (int)radius, rc);
    if (hfg_abi_version() != HFG_ABI_VERSION) { fprintf(stderr, "header / library ABI mismatch\n"); return 3; }

    hfg_handle* h = NULL;
    rc = hfg_create(&cfg, &h);
    if (rc != HFG_OK) return fail(h, "hfg_create", rc);

    /* --- mel + load_state_dict ----------------------------------------------------------------- */
    int B = 0, T = 0, n_weights = 0;
    float* mel = NULL;
    char name[512], file[512];
    while (fscanf(mf, "%15s", tag) == 1) {
        if (strcmp(tag, "mel") == 0) {
            if (fscanf(mf, "%d %d %511s", &B, &T, file) != 3 || B < 1 || T < 1) return 2;
            mel = read_floats(dir, file, (size_t)B * cfg.n_mels * T);
            if (!mel) return 2;
        } else if (strcmp(tag, "w") == 0) {
            int ndim = 0;
            int64_t shape[3] = {1, 1, 1};
            if (fscanf(mf, "%511s %d", name, &ndim) != 2 || ndim < 1 || ndim > 3) return 2;
            size_t n = 1;
            for (int i = 0; i < ndim; ++i) {
                long long d;
                if (fscanf(mf, "%lld", &d) != 1 || d < 1) return 2;
                shape[i] = d; n *= (size_t)d;
            }
            if (fscanf(mf, "%511s", file) != 1) return 2;
            float* w = read_floats(dir, file, n);
            if (!w) return 2;
            rc = hfg_set_weight(h, name, w, shape, ndim);     /* copies */
            free(w);
            if (rc != HFG_OK) return fail(h, name, rc);
            ++n_weights;
        } else {
            fprintf(stderr, "manifest: unknown tag %s\n", tag); return 2;
        }
    }
    fclose(mf);
    if (!mel) { fprintf(stderr, "manifest: no mel line\n"); return 2; }

    /* --- remove_weight_norm + repack ----------------------------------------------------------- */
    rc = hfg_commit_weights(h);
    if (rc != HFG_OK) return fail(h, "hfg_commit_weights", rc);

    /* --- forward ------------------------------------------------------------------------------- */
    int64_t out_len = 0;
    rc = hfg_out_len(h, T, &out_len);
    if (rc != HFG_OK) return fail(h, "hfg_out_len", rc);
    float* wav = (float*)malloc((size_t)B * out_len * sizeof(float));
    rc = hfg_forward_host(h, mel, B, T, wav, mode);           /* first call: plain launches   */
    if (rc == HFG_OK) rc = hfg_forward_host(h, mel, B, T, wav, mode);   /* second call: CUDA-graph replay */
    if (rc != HFG_OK) return fail(h, "hfg_forward_host", rc);
    int64_t launches = 0;
    hfg_last_launch_count(h, &launches);

    snprintf(path, sizeof path, "%s/wav.bin", dir);
    FILE* of = fopen(path, "wb");
    if (!of || fwrite(wav, sizeof(float), (size_t)B * out_len, of) != (size_t)B * out_len) {
        fprintf(stderr, "cannot write %s\n", path); return 2;
    }
    fclose(of);
    double sum = 0.0;
    for (size_t i = 0; i < (size_t)B * out_len; ++i) sum += wav[i];
    printf("ok %d %lld %lld %d %.9g\n", B, (long long)out_len, (long long)launches, n_weights, sum);
    free(wav); free(mel);
    hfg_destroy(h);
    return 0;
}
