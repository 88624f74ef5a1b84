/*
 * hfg_ard.h -- C ABI of the KV-cached autoregressive mel decoder.
 *
 * Stands behind the reference's PNCAARDecoder in inference mode
 * (reference models/ar_decoder.py:167-238, called from models/acoustic_model.py:258): prenet ->
 * positional encoding -> 6-layer post-norm nn.TransformerDecoder over the encoder memory -> mel_proj,
 * one frame per step.  The reference re-runs the whole decoder on the growing prefix for every frame
 * (O(T^2) layer evaluations, no cache); the self-attention is causal, so position t's keys / values never
 * change: this library keeps them per layer (K/V cache), projects the encoder memory once per layer, and
 * evaluates one position per step -- the same frames to fp32 round-off, O(T) layer evaluations.
 *
 * Same conventions as hfg.h: plain C, 0 / negative hfg_status, device pointers owned by the caller, `stream`
 * a cudaStream_t, no CPU fallback.  fp32 arithmetic (FFMA kernels; the work per step is a few MFLOP per
 * utterance, the path is latency-bound): one CUDA graph of the ~70 launches of a step, replayed max_len times.
 */
#ifndef HFG_ARD_H_
#define HFG_ARD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hfg_ard_handle hfg_ard_handle;

/* PNCAARDecoder.__init__ arguments (reference models/ar_decoder.py:38-39). */
typedef struct hfg_ard_config {
    int32_t d_model;   /* 256  */
    int32_t n_mels;    /* 80   */
    int32_t n_layers;  /* 6    */
    int32_t n_heads;   /* 8    (head_dim = d_model / n_heads must be 16, 32, 64 or 128) */
    int32_t d_ff;      /* 2048 */
    int32_t max_pos;   /* rows of the positional-encoding table (reference: 5000, models/ar_decoder.py:69) */
} hfg_ard_config;

int hfg_ard_create(const hfg_ard_config* cfg, hfg_ard_handle** out);
void hfg_ard_destroy(hfg_ard_handle* h);
const char* hfg_ard_last_error(const hfg_ard_handle* h);

/* One tensor of the reference module's state_dict by its key ("prenet.0.weight", "pos_encoding.pe",
 * "decoder.layers.3.multihead_attn.in_proj_weight", "mel_proj.bias", ...): HOST fp32, copied. */
int hfg_ard_set_weight(hfg_ard_handle* h, const char* name, const float* data, const int64_t* shape, int32_t ndim);
int hfg_ard_commit_weights(hfg_ard_handle* h);

int hfg_ard_workspace_bytes(const hfg_ard_handle* h, int32_t batch, int32_t frames, int32_t max_len, size_t* bytes);

/* PNCAARDecoder._forward_autoregressive (reference models/ar_decoder.py:167-238):
 * hvar_dev fp32 [B, frames, d_model] (encoder memory, every frame attended, no mask -- as the reference)
 * -> mel_dev fp32 [B, max_len, n_mels].  max_len <= max_pos.  Asynchronous on `stream`. */
int hfg_ard_decode(hfg_ard_handle* h, const float* hvar_dev, int32_t batch, int32_t frames, int32_t max_len,
                   float* mel_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Kernels launched by the last hfg_ard_decode call (graph replays counted per kernel node). */
int hfg_ard_last_launch_count(const hfg_ard_handle* h, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* HFG_ARD_H_ */
