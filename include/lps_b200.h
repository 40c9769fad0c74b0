/*
 * lps_b200.h -- C ABI of the B200-native log-power-spectrum (LPS) front end, the drop-in for the
 * arithmetic of the reference's Wav2LPS_be at 16 kHz
 * (Feature_prepare/SourceCode_Wav2LogSpec_be/Wav2LogSpec_be.c:395-404, 413-576; FEfunc.c:80-118, 146-293;
 *  fileio.c:231-243, 268-282): int16 PCM -> 512-sample frames every 256 samples -> Hamming window ->
 * 512-point real split-radix FFT -> power -> floored natural log -> 257 bins per frame.
 *
 * Two kernels: the default computes the same transform with a register-resident radix-8 FFT (throughput path: fp32,
 * exactly rounded twiddles); with LPS_FLAG_EXACT the kernel executes the reference's butterfly network itself (same
 * operations, same order within each butterfly, no FMA contraction), so features are bit-identical to the reference
 * except where the double-precision log of the GPU differs from glibc's in the last place.
 *
 * Entry point               replaces
 * ------------------------  ---------------------------------------------------------------------
 * lps_nframes               frame count implied by the main loop, Wav2LogSpec_be.c:401-416
 * lps_extract               main frame loop for one utterance (host buffers)
 * lps_extract_batch         the same for many utterances in one launch (host buffers)
 * lps_extract_batch_device  the same with device-resident PCM / features
 * lps_norm_reset/finalize     qnnorm (tools_pfile/get_norm.pl:4): mean / reciprocal std per bin, accumulated on the device
 * Output options (flags) cover the consumers of the features: the HTK writer (big-endian floats,
 * fileio.c:231-243), the trainer's z-score with the reciprocal-std .norm file (Interface.cc:760-766) and
 * the pfile records feacat builds from the HTK files (tools_pfile/pfile_noisy.pl:33).
 */
#ifndef LPS_B200_H_
#define LPS_B200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LPS_BINS 257
#define LPS_FRAME_LEN 512
#define LPS_FRAME_SHIFT 256

enum {
    LPS_FLAG_BIG_ENDIAN = 1,  /* byte-swap each float (ready for fwrite into an HTK file) */
    LPS_FLAG_ZSCORE = 2,      /* (x - mean[k]) * dvar[k] with the .norm constants; not combinable with BIG_ENDIAN */
    LPS_FLAG_PFILE = 8,       /* write QuickNet pfile RECORDS instead of bare features: per frame 259 big-endian words {sentence index,
                                 frame index in the sentence, 257 float32} -- the payload feacat produces (tools_pfile/pfile_noisy.pl:33),
                                 ready to be written after the 32 768-byte header.  `out` must hold frames * 259 words. */
    LPS_FLAG_ACCUM_NORM = 16, /* also accumulate the per-bin sum and sum of squares of the produced frames on the device: the input of
                                 lps_norm_finalize (qnnorm, tools_pfile/get_norm.pl:4) */
    LPS_FLAG_EXACT = 4        /* execute the reference's split-radix butterfly network itself (bit-identical spectrum, HTK-identical
                                 files) instead of the default register-resident radix-8 FFT (fp32; agrees with the reference to
                                 ~1e-6 of the feature range, see tests/test_lps_gpu.py) */
};

typedef struct lps_handle lps_handle;

long lps_nframes(long n_samples);
int lps_create(int gpu, lps_handle **out);
int lps_destroy(lps_handle *h);
/* mean / dvar: 257 floats each (host), needed for LPS_FLAG_ZSCORE */
int lps_set_norm(lps_handle *h, const float *mean, const float *dvar);

/* one utterance, host in / host out; out must hold lps_nframes(n_samples) * 257 floats */
int lps_extract(lps_handle *h, const int16_t *pcm, long n_samples, float *out, int flags);
/* utt_off: n_utts + 1 sample offsets into pcm (host). Frames of utterance u follow those of u-1 in out.
 * total_frames (optional) receives the frame count. */
int lps_extract_batch(lps_handle *h, const int16_t *pcm, const long *utt_off, int n_utts, float *out, int flags, long *total_frames);
/* device-resident variant; utt_off stays a host array. d_out must hold sum_u lps_nframes(len_u) * 257 floats */
int lps_extract_batch_device(lps_handle *h, const int16_t *d_pcm, const long *utt_off, int n_utts, float *d_out, int flags, long *total_frames);
/* .norm production (qnnorm): mean and RECIPROCAL standard deviation (ddof 0) per bin over every frame produced with
 * LPS_FLAG_ACCUM_NORM since the last lps_norm_reset.  mean / dvar: 257 floats each (host). */
int lps_norm_reset(lps_handle *h);
/* the same accumulation for features already on the device: n_frames rows of `pitch` 32-bit words, the 257 values `skip`
 * words into a row, big-endian when big_endian != 0 (pitch 259 / skip 2 / big_endian 1 = pfile records as they lie in the file) */
int lps_norm_accumulate_device(lps_handle *h, const float *d_feats, long n_frames, int pitch, int skip, int big_endian);
int lps_norm_finalize(lps_handle *h, float *mean, float *dvar, long *n_frames);
/* CUDA-event time (ms) of the kernel(s) of the last call */
double lps_last_kernel_ms(lps_handle *h);
const char *lps_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
