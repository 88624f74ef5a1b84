/*
 * hfg.h -- C ABI of the B200-native HiFi-GAN generator inference path.
 *
 * This is the whole drop-in boundary: a plain-C shared library
 * (libhfg_b200.so; no torch, no C++ types in any signature) that the host-side
 * mirror of the reference's nn.Module binds with ctypes
 * (tts-sambert_hifigan_b200/_capi.py).  The reference has no FFI of its own --
 * its boundary is the Python class models/hifigan.py::HiFiGANGenerator used by
 * composition (reference models/hifigan.py:681-689, called at :719) -- so each
 * entry point below cites the reference method or attribute it stands behind.
 *
 * Conventions
 *   - every function returns 0 (HFG_OK) or a negative hfg_status; nothing
 *     throws across the ABI; hfg_last_error() gives the message of the last
 *     failure on that handle;
 *   - a handle belongs to the CUDA device that was current in hfg_create and is
 *     not thread-safe (one handle per GPU, as one process drives one GPU); it owns one
 *     set of internal side streams and events, so calls on one handle must be issued on
 *     one caller stream at a time (a second handle overlaps independent batches);
 *   - device pointers are plain CUDA device pointers owned by the caller
 *     (PyTorch allocates them); `stream` is a cudaStream_t passed as void*;
 *   - hfg_forward is asynchronous on `stream`.
 *   - there is NO CPU fallback: without a CUDA device hfg_create fails with
 *     HFG_ERR_CUDA.
 */
#ifndef HFG_H_
#define HFG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HFG_MAX_STAGES 8
#define HFG_ABI_VERSION 3   /* 3: + hfg_tf32_plan (additive) */

typedef struct hfg_handle hfg_handle;

/* Constructor arguments of HiFiGANGenerator.__init__ (reference
 * models/hifigan.py:149-158), flattened. */
typedef struct hfg_config {
    int32_t n_mels;                                   /* :151 */
    int32_t num_upsamples;                            /* len(upsample_rates), :173 */
    int32_t upsample_initial_channel;                 /* :154 */
    int32_t num_resblocks;                            /* len(resblock_kernel_sizes), :172 */
    int32_t upsample_rates[HFG_MAX_STAGES];           /* :152 */
    int32_t upsample_kernel_sizes[HFG_MAX_STAGES];    /* :153 */
    int32_t resblock_kernel_sizes[HFG_MAX_STAGES];    /* :155 */
    int32_t num_dilations[HFG_MAX_STAGES];            /* len(resblock_dilation_sizes[j]) */
    int32_t resblock_dilations[HFG_MAX_STAGES][HFG_MAX_STAGES]; /* :156 */
} hfg_config;

typedef enum hfg_status {
    HFG_OK = 0,
    HFG_ERR_INVALID = -1,     /* bad argument / shape mismatch (reference: ATen shape error) */
    HFG_ERR_CUDA = -2,        /* CUDA runtime failure, or no CUDA device                     */
    HFG_ERR_STATE = -3,       /* weights missing / not committed                             */
    HFG_ERR_WORKSPACE = -4,   /* workspace too small or misaligned                           */
    HFG_ERR_UNSUPPORTED = -5  /* mode not available for this configuration                   */
} hfg_status;

/* Arithmetic modes.  The output is always fp32 [B,1,T_out]
 * (reference tests/test_hifigan_integration.py:50). */
typedef enum hfg_mode {
    HFG_MODE_FP32 = 0,  /* fp32 FFMA kernels, fp32 activations: strict parity mode      */
    HFG_MODE_TF32 = 1,  /* TF32-class arithmetic on tensor cores: 10-bit-mantissa operands, fp32 accumulate,
                           residual stream kept to >= 22 mantissa bits (parity <= 1e-3).  Configurations whose
                           every ResBlock pair fits the fused kernel (the default one does) run it on fp16
                           operand planes -- the same 10-bit mantissa, rounded to nearest -- with the residual
                           stream stored as an fp16 pair hi + lo; others use tcgen05 kind::tf32 on fp32 planes */
    HFG_MODE_BF16 = 2,  /* tcgen05 kind::f16 (bf16 operands), bf16 activations, fp32 acc  */
    HFG_MODE_FP16 = 3   /* tcgen05 kind::f16 (fp16 operands), fp16 activations, fp32 acc: tf32's 10-bit
                           mantissa at bf16's MMA rate; conversions saturate at +-65504 (parity <= 1e-3) */
} hfg_mode;

int hfg_abi_version(void);

/* HiFiGANGenerator.__init__ (reference models/hifigan.py:149-222). */
int hfg_create(const hfg_config* cfg, hfg_handle** out);
void hfg_destroy(hfg_handle* h);
const char* hfg_last_error(const hfg_handle* h);

/* One tensor of the module's state_dict, by its reference key
 * ("conv_pre.weight", "ups.0.bias", "mrfs.1.resblocks.2.convs1.0.weight_g", ...;
 * schema: SURVEY.md section 8b).  `data` is a HOST fp32 pointer in the
 * reference's own layout ([C_out,C_in,k] for Conv1d, [C_in,C_out,k] for
 * ConvTranspose1d); it is copied.  Replaces nn.Module.load_state_dict for this
 * module.  Both the plain (.weight) and the weight-normed (.weight_g/.weight_v,
 * reference models/hifigan.py:274-283) schema are accepted. */
int hfg_set_weight(hfg_handle* h, const char* name, const float* data,
                   const int64_t* shape, int32_t ndim);

/* Fold weight-norm (w = g*v/||v||, norm over all dims but 0), repack every
 * layer into kernel layout and upload.  Replaces remove_weight_norm()
 * (reference models/hifigan.py:263-272).  Fails with HFG_ERR_STATE and names
 * the first missing key if the state is incomplete. */
int hfg_commit_weights(hfg_handle* h);

/* T_out of forward() for Tfrm frames: each ConvTranspose1d maps
 * T -> (T-1)*u - 2*((k-u)//2) + k (reference models/hifigan.py:196-202). */
int hfg_out_len(const hfg_handle* h, int32_t frames, int64_t* out_len);

int hfg_workspace_bytes(const hfg_handle* h, int32_t batch, int32_t frames, int32_t mode,
                        size_t* bytes);

/* HiFiGANGenerator.forward (reference models/hifigan.py:224-261):
 * mel_dev fp32 [B, n_mels, Tfrm] contiguous  ->  wav_dev fp32 [B, 1, T_out].
 * Asynchronous: everything is ordered after the work already in `stream` and
 * before whatever is enqueued on it next.  In the tensor-core modes the three
 * resblocks of each MRF run on `stream` plus two internal streams of the
 * handle (fork / join with events, no host synchronisation); the workspace may
 * be dirty on entry (the pad rows a valid output depends on are re-zeroed). */
int hfg_forward(hfg_handle* h, const float* mel_dev, int32_t batch, int32_t frames,
                float* wav_dev, void* workspace_dev, size_t workspace_bytes, int32_t mode,
                void* stream);

/* The acoustic model upstream emits mel_pred as [B, Tfrm, n_mels] (reference
 * models/acoustic_model.py:181-265; the reference design transposes it first,
 * .kiro/specs/tts-sam-bert-hifigan/design.md:905-906).  layout = HFG_MEL_FRAMES_LAST makes every
 * following hfg_forward* call on this handle read that layout directly (the transpose is folded
 * into the first kernel's load); HFG_MEL_CHANNELS_FIRST (default) is the reference's [B, n_mels, Tfrm]. */
#define HFG_MEL_CHANNELS_FIRST 0
#define HFG_MEL_FRAMES_LAST 1
int hfg_set_mel_layout(hfg_handle* h, int32_t layout);

/* Same, and also copies the 2*num_upsamples+1 stage-boundary activations
 * (conv_pre, then ups[i], mrfs[i] outputs; fp32 [B,C,T] device buffers; NULL
 * entries are skipped) -- what forward hooks on the reference module observe.
 * Test/debug entry point. */
int hfg_forward_stages(hfg_handle* h, const float* mel_dev, int32_t batch, int32_t frames,
                       float* wav_dev, void* workspace_dev, size_t workspace_bytes, int32_t mode,
                       void* stream, float* const* stage_out_dev);

/* Variable-length batch.  The reference has no masks: LengthRegulator pads every
 * utterance to the batch maximum (reference models/variance_adaptor.py:240-264) and the generator synthesises
 * the padding.  Here lengths_dev (device, int32 [batch]) gives the valid frame count of each utterance; the
 * tile schedulers of every kernel skip tiles beyond (length + halo_frames) frames, so padded frames cost
 * nothing.  Samples [0, T_out(length)) of every utterance are exactly what hfg_forward returns for the same
 * padded mel (halo_frames must be >= hfg_receptive_radius, else HFG_ERR_INVALID); samples beyond are 0.
 * Everything, including the per-utterance row counts, is computed on the device: the call stays
 * asynchronous and graph-capturable. */
int hfg_forward_lengths(hfg_handle* h, const float* mel_dev, const int32_t* lengths_dev, int32_t halo_frames,
                        int32_t batch, int32_t frames, float* wav_dev, void* workspace_dev,
                        size_t workspace_bytes, int32_t mode, void* stream);

/* Receptive radius of one output frame, in mel frames, derived from the configuration (13 for the default
 * one): time chunking (config 4) and hfg_forward_lengths are exact with a halo of at least this many frames.
 * Host-only; needs no device. */
int hfg_receptive_radius(const hfg_config* cfg, int32_t* frames);

/* End-to-end call with HOST buffers: stages mel through pinned memory, copies
 * host->device, runs forward, copies the waveform back and synchronises.
 * Pinned staging, workspace and stream are owned by the handle and grown on
 * demand.  This is what `module(mel_cpu_tensor)` costs a caller whose data
 * lives on the host.  From the second call with the same (batch, frames,
 * mode, layout) the launch sequence is replayed from a CUDA graph
 * (environment HFG_HOST_GRAPH=0 turns that off). */
int hfg_forward_host(hfg_handle* h, const float* mel_host, int32_t batch, int32_t frames,
                     float* wav_host, int32_t mode);

/* Same with hints: a buffer flagged as page-locked (cudaHostAlloc / torch pin_memory) is used
 * for the DMA directly instead of being staged through the handle's own pinned buffer. */
#define HFG_HOST_MEL_PINNED 1u
#define HFG_HOST_WAV_PINNED 2u
int hfg_forward_host_ex(hfg_handle* h, const float* mel_host, int32_t batch, int32_t frames,
                        float* wav_host, int32_t mode, uint32_t flags);

/* Streaming form of the host-buffer call, for callers that generate batch after batch: up to two submissions
 * (slot 0 / 1) are in flight.  Submit queues the H2D copy of `mel_host`, the forward (CUDA-graph replay) and the
 * D2H copy into `wav_host` on three streams of the handle and returns at once; wait(slot) blocks until that
 * slot's waveform has landed.  The copies of one submission overlap the kernels of its neighbours, so a steady
 * stream of batches runs at the device rate.  Both host buffers MUST be page-locked and must stay untouched until
 * wait(slot) returns; a slot must be waited for before it is submitted again (HFG_ERR_STATE otherwise).  Results are
 * identical to hfg_forward_host.  The blocking and the streaming calls share the handle's compute stream and workspace:
 * drain the slots (wait) before mixing in hfg_forward_host* calls or re-committing weights. */
int hfg_forward_host_submit(hfg_handle* h, int32_t slot, const float* mel_host, int32_t batch, int32_t frames,
                            float* wav_host, int32_t mode);
int hfg_forward_host_wait(hfg_handle* h, int32_t slot);

/* Per-launch device timing.  When enabled, every kernel launch of hfg_forward*
 * is bracketed by a CUDA event pair on the launching stream; hfg_get_profile
 * synchronises and returns a JSON array, one entry per kernel label:
 *   [{"kernel":"mrf0","launches":18,"ms":..,"flops":..,"bytes":..}, ...]
 * (flops / bytes are the ALGORITHMIC work of those launches).  Call with
 * buf == NULL to query the size.  Used by bench.py for the roofline figures.
 * enable = 1: per launch; the launches of a forward are serialised on the
 *   caller's stream so that every kernel is timed alone.
 * enable = 2: per stage ("head", "ups<i>", "mrf<i>", "tail"), with the
 *   resblocks of each MRF running concurrently as in a normal forward. */
int hfg_set_profiling(hfg_handle* h, int32_t enable);
int hfg_get_profile(hfg_handle* h, char* buf, size_t buf_bytes, size_t* needed);

/* Tuning / profiling aid: time `iters` launches of ONE MRF convolution
 * (mrfs[stage].resblocks[resblock].convs{1,2}[pair]; which = 0 / 1, or 2 for the
 * fused conv1+conv2 pair kernel) of the
 * tensor-core path on scratch buffers of `batch` x `rows` time steps.
 * Returns the average milliseconds per launch. */
int hfg_bench_layer(hfg_handle* h, int32_t stage, int32_t resblock, int32_t pair, int32_t which,
                    int32_t batch, int32_t rows, int32_t mode, int32_t iters, float* ms);

/* ---- upstream glue: integer frame indexing of the acoustic model ----
 * Stand-alone (no handle); asynchronous on `stream`; all pointers are device pointers.
 *
 * hfg_durations_from_log: dur = clamp(round_half_even(exp(log_dur)), min=1) as int64
 *   (VarianceAdaptor inference branch, reference models/variance_adaptor.py:746-748).
 * hfg_length_regulate: LengthRegulator.forward (reference models/variance_adaptor.py:171-269):
 *   out[b, t, :] = henc[b, p(t), :] (repeat_interleave by clamp(dur, min=0)), zeros for
 *   t >= sum(dur[b]); out is [B, Tfrm, D] with Tfrm chosen by the caller (the reference uses the
 *   batch maximum, which hfg_length_regulate_frames computes -- that call synchronises `stream`). */
int hfg_durations_from_log(const float* log_dur, int64_t n, int64_t* dur, void* stream);
int hfg_length_regulate_frames(const int64_t* dur, int32_t batch, int32_t n_phonemes, int64_t* max_frames,
                               void* stream);
int hfg_length_regulate(const float* henc, const int64_t* dur, int32_t batch, int32_t n_phonemes,
                        int32_t d_model, int32_t frames, float* out, void* stream);

/* Number of kernels the last hfg_forward* call on this handle launched. */
int hfg_last_launch_count(const hfg_handle* h, int64_t* launches);

/* How HFG_MODE_TF32 runs on this (committed) handle: *split = 1 when the MMAs read fp16 planes and the residual
 * stream is stored as an fp16 pair hi + lo (every ResBlock pair fits the fused kernel), 0 when the fp32-plane
 * tcgen05 kind::tf32 kernels are used.  Both compute 10-bit-mantissa products accumulated in fp32. */
int hfg_tf32_plan(const hfg_handle* h, int32_t* split);

#ifdef __cplusplus
}
#endif
#endif /* HFG_H_ */
