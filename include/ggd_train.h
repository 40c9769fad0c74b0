/*
 * ggd_train.h -- C ABI of the B200-native drop-in for the reference's BP_GPU class
 * (Train_code_ML_GGD/BP_GPU.h:45-70), i.e. the device path of BPtrain_Sigmoid:
 * forward -> GGD / beta-norm loss gradient -> backward -> momentum SGD, plus the three
 * cross-validation metrics.  Plain pointers and sizes only; every entry point returns 0 on
 * success or a negative GGD_E* code (ggd_last_error() gives the text).  There is no CPU
 * fallback: every call fails with GGD_ECUDA when no sm_100 device is usable.
 *
 * Entry point                  replaces (reference file:line)
 * ---------------------------  -------------------------------------------------------------
 * ggd_create                   BP_GPU::BP_GPU               BP_GPU.cu:9-113
 * ggd_destroy                  BP_GPU::~BP_GPU              BP_GPU.cu:115-150
 * ggd_train                    BP_GPU::train                BP_GPU.cu:152-185  (+ train_bunch_single :308-440)
 * ggd_cv_sqerr                 BP_GPU::CrossValid           BP_GPU.cu:187-221
 * ggd_cv_abserr                BP_GPU::CrossValiddB         BP_GPU.cu:222-255
 * ggd_cv_loglik                BP_GPU::CrossValid2 + Gamma  BP_GPU.cu:256-306, 593-640
 * ggd_get_weights              BP_GPU::returnWeights        BP_GPU.cu:514-525
 *
 * Layouts are the reference's: `in` is n_frames x layersizes[0] row-major (already z-scored,
 * context-expanded, shuffled: Interface.cc:719-838), `targ` is n_frames x layersizes[last],
 * weights[l] (l = 1..numlayers-1) is float[layersizes[l]*layersizes[l-1]] with
 * index = out + in*layersizes[l] (the .wts / MATLAB order, DevFunc.h:65-75), bias[l] is
 * float[layersizes[l]].  Index 0 of the weights/bias pointer arrays is unused, as in the reference.
 * The caller owns every host array; the library copies in and out.
 */
#ifndef GGD_TRAIN_H_
#define GGD_TRAIN_H_
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGD_MAXLAYER 10          /* BP_GPU.h:6 */
#define GGD_MAXCACHEFRAME 200000 /* BP_GPU.h:7: upper bound on n_frames per call */

enum {
    GGD_OK = 0,
    GGD_EINVAL = -1,   /* bad argument / unsupported configuration */
    GGD_ECUDA = -2,    /* CUDA runtime or driver error, or no sm_100 device */
    GGD_ENOMEM = -3,
    GGD_ENCCL = -4,
    GGD_EUNSUPPORTED = -5
};

/* arithmetic used for the dense contractions */
enum {
    GGD_PREC_BF16X3 = 0,  /* default: tcgen05 kind::f16, every fp32 operand split into bf16 hi+lo,
                             3 MMAs (hi*hi + hi*lo + lo*hi), fp32 accumulation in TMEM */
    GGD_PREC_FP32_SIMT = 1 /* validation path: plain fp32 FMA GEMMs on CUDA cores (slow) */
};

typedef struct ggd_config {
    int   numlayers;                  /* <= GGD_MAXLAYER, counts the input layer */
    int   layersizes[GGD_MAXLAYER];
    int   bunchsize;                  /* frames per minibatch ON THIS RANK */
    float lrate, momentum, weightcost;
    float shapefactor;                /* beta */
    int   MLflag;                     /* 1: GGD maximum-likelihood gradient; else beta-norm */
    int   dropoutflag;                /* must be 0 (dropout is outside the named path) */
    float visible_omit, hid_omit;     /* accepted, unused while dropoutflag == 0 */
    int   gpu;                        /* CUDA device ordinal (gpu_used=) */
    int   seed;                       /* accepted for signature parity (only seeds dropout in the reference) */
    int   precision;                  /* GGD_PREC_* */
    /* frame-sharded data parallelism; world_size <= 1 means single GPU */
    int   world_size, rank;
    const void *nccl_unique_id;       /* 128-byte ncclUniqueId shared by all ranks, or NULL */
    int   flags;                      /* GGD_FLAG_* */
} ggd_config;

enum {
    GGD_FLAG_UNFUSED_UPDATE = 1,  /* materialise the weight gradient and run the stand-alone update kernel instead of the
                                     fused gradient+update kernel (validation / debugging: the gradient stays readable) */
    GGD_FLAG_NO_GRAPH = 2,        /* launch kernels directly instead of replaying a CUDA graph */
    GGD_FLAG_KEEP_DEBUG = 4,      /* keep per-step tensors readable through ggd_debug_read */
    GGD_FLAG_PIN_HOST = 8         /* cudaHostRegister the caller's (long-lived, reused) chunk buffers on first use;
                                     the reference's Interface allocates them once (Interface.cc:476-480).  A buffer
                                     registered this way must stay allocated until ggd_release_host() or ggd_destroy() */
};

typedef struct ggd_handle ggd_handle;

int ggd_create(const ggd_config *cfg, const float *const *weights, const float *const *bias, ggd_handle **out);
int ggd_destroy(ggd_handle *h);

/* Optional: allocate the chunk staging for up to n_frames and capture the step graphs now instead of inside
 * the first ggd_train* call (the reference allocates for MAXCACHEFRAME in its constructor, BP_GPU.cu:72-75). */
int ggd_reserve(ggd_handle *h, int n_frames);

/* One call = one chunk: H2D copy, then one training step per full bunch; a trailing partial bunch is
 * dropped (BP_GPU.cu:173-180).  Blocking, like the reference. */
int ggd_train(ggd_handle *h, int n_frames, const float *in, const float *targ);
/* Device-side loader: replaces the ARITHMETIC of Interface::Readchunk (Interface.cc:735-838).  The caller
 * hands over the raw pfile records of one chunk exactly as they lie in the files (big-endian 32-bit words, 2 + dim per
 * frame: sentence id, frame id, features) and, for every net-input row (already in the shuffled order of :750-754), the
 * index of its first context frame inside the chunk.  Byte swap, z-score with mean / reciprocal std (noisy statistics on
 * BOTH streams, indexed j % fea_dim for the targets), context expansion, target-frame selection and the operand split
 * run on the GPU: 2 KB per frame cross PCIe instead of 8.2 KB per sample, and the host does no per-element work.
 * Then trains like ggd_train (trailing partial bunch dropped).  Results are bit-identical to ggd_train on the host-expanded
 * arrays. */
typedef struct ggd_raw_chunk {
    const unsigned int *fea_records;     /* n_frames * (2 + fea_dim) words, noisy-speech pfile */
    const unsigned int *targ_records;    /* n_frames * (2 + layersizes[last]) words, clean-speech pfile */
    int n_frames;                        /* raw frames in the chunk ("need", Interface.cc:729-733) */
    int n_samples;                       /* net-input rows */
    const int *sample_first_frame;       /* [n_samples], 0 <= f and f + fea_context <= n_frames */
    int fea_dim, fea_context, targ_offset;
    const float *mean, *dvar;            /* [fea_dim] */
    /* Data parallelism: every rank may supply only ITS slice of the chunk's records and the library all-gathers them over
     * NVLink (ncclAllGather) -- each pfile byte is then read from disk and crosses PCIe once per node instead of once per
     * GPU.  rec_frames == 0: the arrays above hold all n_frames records.  Otherwise they hold the records of frames
     * [rec_frame0, rec_frame0 + rec_frames) with rec_frame0 = rank * S, S = ceil(n_frames / world_size),
     * rec_frames = min(S, n_frames - rec_frame0) (clamped at 0). */
    int rec_frame0, rec_frames;
} ggd_raw_chunk;
int ggd_train_raw(ggd_handle *h, const ggd_raw_chunk *chunk);
/* Page-locked host memory for chunk buffers (the record arrays of ggd_raw_chunk, the arrays of ggd_train): copies from it
 * run at PCIe speed and overlap the training steps; pageable memory is staged by the driver at a fraction of that. */
/* makes the handle's GPU the calling thread's current device (call it once in a loader thread before ggd_host_alloc) */
int ggd_bind_thread(ggd_handle *h);
void *ggd_host_alloc(size_t bytes);
void ggd_host_free(void *p);

/* Drops the GGD_FLAG_PIN_HOST registration of a host buffer previously passed to ggd_train (call it before freeing or
 * reallocating such a buffer while the handle lives; unknown pointers are ignored). */
int ggd_release_host(ggd_handle *h, const void *host_ptr);

/* Same, for a chunk that is already resident in device memory (fp32, same layouts). */
int ggd_train_device(ggd_handle *h, int n_frames, const float *d_in, const float *d_targ);

int ggd_cv_sqerr(ggd_handle *h, int n_frames, const float *in, const float *targ, float *result);
int ggd_cv_abserr(ggd_handle *h, int n_frames, const float *in, const float *targ, float *result);
int ggd_cv_loglik(ggd_handle *h, int n_frames, const float *in, const float *targ, float *result);
/* The three CV metrics of BPtrain.cc:124-128 from ONE forward pass: result3 = {CrossValid, CrossValiddB, CrossValid2}
 * (CrossValid2 is 0 unless MLflag == 1); identical values to the three separate calls. */
int ggd_cv_all(ggd_handle *h, int n_frames, const float *in, const float *targ, float *result3);
/* Inference path of Test_code/decode.m:28-62 + frame_expand.m:6-25 for ONE utterance: z-score of the raw LPS frames
 * lps[n_frames][fea_dim] with the .norm constants, fea_context (odd) frames of context with the first / last frame
 * replicated at the utterance edges, forward pass, de-normalisation out/dvar + mean.  out: n_frames x layersizes[last]. */
int ggd_enhance(ggd_handle *h, int n_frames, const float *lps, int fea_dim, int fea_context, const float *mean, const float *dvar,
                float *out);
/* forward only: out is n_frames x layersizes[last] (cv_bunch_single, BP_GPU.cu:442-512) */
int ggd_forward(ggd_handle *h, int n_frames, const float *in, float *out);

int ggd_get_weights(ggd_handle *h, float *const *weights, float *const *bias);
/* GGD scale factors alpha_d left by the last training bunch (dev.scalefactor) */
int ggd_get_alpha(ggd_handle *h, float *alpha);
/* per-bunch loss trace of the last ggd_train* call (SURVEY.md 8c definition); *n = bunches run */
int ggd_get_losses(ggd_handle *h, float *losses, int max, int *n);

/* timing / accounting of the last ggd_train* call */
typedef struct ggd_stats {
    double device_ms;        /* CUDA-event time of the training steps on the compute stream */
    double h2d_ms;           /* CUDA-event time of the chunk upload (0 for ggd_train_device) */
    long long launches;      /* kernels launched (graph nodes count individually) */
    long long steps;         /* bunches trained */
    long long h2d_bytes, d2h_bytes;
} ggd_stats;
int ggd_get_stats(ggd_handle *h, ggd_stats *s);

/* Per-kernel timing: runs n_frames / bunchsize training steps WITHOUT the graph, with a CUDA-event pair
 * around every launch on the compute stream, and returns the summed milliseconds and launch counts per
 * kernel class (this is what bench.py divides the algorithmic bytes / FLOPs by). */
enum { GGD_KC_FWD = 0, GGD_KC_LOSS, GGD_KC_DX, GGD_KC_DW, GGD_KC_BIAS, GGD_KC_ALLREDUCE, GGD_KC_UPDATE, GGD_KC_ADVANCE,
       GGD_KC_SPLIT, GGD_KC_DWUPD, GGD_KC_PUSH, GGD_KC_COUNT };
typedef struct ggd_kernel_times {
    double ms[16];
    long long launches[16];
    long long steps;
    long long param_elems;   /* elements of the (padded) parameter arena the update kernel streams */
} ggd_kernel_times;
int ggd_profile_kernels(ggd_handle *h, int n_frames, const float *d_in, const float *d_targ, ggd_kernel_times *out);

/* Parity hooks (tests only): run ONE bunch with or without applying the update, then read tensors.
 * what: 0 = out [M][D], 1 = dedx of layer l [M][units], 2 = y of layer l, 3 = weight gradient of
 * layer l (reference order out + in*cur), 4 = bias gradient of layer l. */
int ggd_debug_step(ggd_handle *h, int n_frames, const float *in, const float *targ, int apply_update);
int ggd_debug_read(ggd_handle *h, int what, int layer, float *dst);

/* Raw tcgen05 GEMM probe (tests only): D[i][j] = sum_r A(i,r) B(j,r), fp32 host arrays.
 * a_mn / b_mn = 0: operand is [rows][R] (reduction contiguous); 1: operand is [R][rows]. */
int ggd_debug_gemm(int a_mn, int b_mn, int I, int J, int R, int bn, int splits, const float *A, const float *B, float *D);
/* same, timed: `reps` back-to-back launches between CUDA events (avg_ms), plus an optional per-CTA trace of
 * globaltimer stamps (trace_host[ctas][16]; slots: 0 entry, 1 setup done, 2 first stage issued, 3 first stage landed,
 * 4 last stage landed, 5 last MMA issued, 6 accumulator ready, 7 cluster exchange done, 8 epilogue done, 9 exit) */
int ggd_debug_gemm_timed(int a_mn, int b_mn, int I, int J, int R, int bn, int splits, const float *A, const float *B, float *D,
                         int reps, float *avg_ms, unsigned long long *trace_host, int trace_ctas);
/* one traced training step (tuning aid): see csrc/ggd_train.cu */
int ggd_debug_trace_step(ggd_handle *h, const float *in, const float *targ, int fused, float *out, int max_launches, int *n_launches);
/* writes a 128-byte ncclUniqueId (rank 0 creates it, every rank passes it in ggd_config) */
int ggd_nccl_unique_id(void *out128);

const char *ggd_last_error(void);
const char *ggd_version(void);

#ifdef __cplusplus
}
#endif
#endif
