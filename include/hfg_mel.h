/*
 * hfg_mel.h -- C ABI of the on-device log-mel front/back end.
 *
 * Stands behind the reference's mel definition -- torchaudio MelSpectrogram(power = 2) + log10(. + 1e-10)
 * (reference data/audio_processing.py:99-127, parameters configs/config.yaml:4-14) -- and its log-mel L1
 * (reference models/losses.py:708-797, VocoderLoss.mel_reconstruction_loss), which is the figure the bf16 mode of
 * the generator is reported with.  One kernel per call: reflect-padded framing (center = True), periodic Hann
 * window, radix-2 FFT in shared memory, power spectrum, slaney-scale / slaney-normalised triangular filterbank,
 * log10; the L1 variant does both waveforms in the same block and reduces |a - b| deterministically.
 * HBM-bound by design: reads each sample n_fft / hop times from L2, writes n_mels values per frame.
 *
 * Conventions as hfg.h (plain C, 0 / negative hfg_status, caller-owned device pointers, cudaStream_t as void*).
 */
#ifndef HFG_MEL_H_
#define HFG_MEL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hfg_mel_handle hfg_mel_handle;

/* configs/config.yaml `audio:` section of the reference (defaults in brackets). */
typedef struct hfg_mel_config {
    int32_t sample_rate;  /* [22050] */
    int32_t n_fft;        /* [1024]  power of two, 64 .. 4096 */
    int32_t hop_length;   /* [256]   */
    int32_t win_length;   /* [1024]  must equal n_fft (the reference's setting) */
    int32_t n_mels;       /* [80]    <= 256 */
    float fmin;           /* [0]     */
    float fmax;           /* [8000]  */
} hfg_mel_config;

int hfg_mel_create(const hfg_mel_config* cfg, hfg_mel_handle** out);
void hfg_mel_destroy(hfg_mel_handle* h);
const char* hfg_mel_last_error(const hfg_mel_handle* h);

/* Frames of a waveform of `samples` samples: samples / hop + 1 (center = True). */
int hfg_mel_frames(const hfg_mel_handle* h, int64_t samples, int64_t* frames);

/* extract_mel / the mel_transform + log of mel_reconstruction_loss:
 * wav_dev fp32 [B, samples] -> out_dev fp32 [B, n_mels, frames].  samples must exceed n_fft / 2 (reflect padding). */
int hfg_log_mel(hfg_mel_handle* h, const float* wav_dev, int32_t batch, int64_t samples, float* out_dev, void* stream);

/* VocoderLoss.mel_reconstruction_loss: mean |log_mel(new) - log_mel(ref)| over [B, n_mels, frames] into *loss_dev
 * (one float on the device).  scratch_dev: batch * frames floats (per-frame partial sums). */
int hfg_log_mel_l1(hfg_mel_handle* h, const float* wav_ref_dev, const float* wav_new_dev, int32_t batch, int64_t samples,
                   float* loss_dev, float* scratch_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HFG_MEL_H_ */
