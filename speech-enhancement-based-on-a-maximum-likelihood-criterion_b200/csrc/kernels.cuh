// kernels.cuh -- launch wrappers of the HBM-bound kernels of the training step
// (input split, fused loss gradient, bias gradient, fused momentum-SGD update) and of the
// fp32 CUDA-core validation GEMM.
#pragma once
#include "common.cuh"

namespace ggd {

typedef __nv_bfloat16 bf16;

// ---- flag layout of the data-parallel exchanges over NVLink peer memory (dp_factor.cuh) ----------------------------
constexpr int FX_MAX = 8;                 // ranks
constexpr int LOSS_FLAGS_PER_RANK = 32;   // alpha exchange: one flag per source rank and 16-column chunk (32-column block in loss_kernel)
// Flag block of one rank: [source rank][FX_STRIDE] 32-bit words, written by the peers with st.release.sys, values = step.
//   [FX_EV_Y + l]    source's activations y_l of this step have landed in my factor arena (l = 0: the net-input rows)
//   [FX_EV_DX + l]   source's dE/dx_l of this step have landed
//   [FX_EV_DONE]     source has finished step `value`: it no longer reads its factor arena (my next pushes may overwrite it)
//   [FX_EV_LOSS + c] source's partial sum_m|e|^beta of column chunk c
constexpr int FX_EV_Y = 0, FX_EV_DX = 10, FX_EV_DONE = 20, FX_EV_LOSS = 32, FX_STRIDE = 64;
static_assert(FX_EV_LOSS + LOSS_FLAGS_PER_RANK <= FX_STRIDE, "flag block layout");

// Device-side step control block: read by every kernel of a step so that one captured CUDA graph
// can be replayed for every bunch of a chunk (the bunch index advances on the device).
struct StepCtl {
    int bunch_idx;        // index of the current bunch inside the chunk
    int pad;
    const float *in32;    // chunk input, fp32 [frames][units0]   (reference layout)
    const float *targ;    // chunk targets, fp32 [frames][D]
};

// fp32 [rows][cols] -> bf16 hi / lo [rows][ld] (columns >= cols are zeroed)
void launch_split_rows(const float *src, int rows, int cols, bf16 *hi, bf16 *lo, int ld, cudaStream_t s);

// Device-side restatement of the arithmetic of Interface::Readchunk (Interface.cc:735-838): raw big-endian pfile records
// -> byte swap, z-score with the noisy-speech mean / reciprocal std on BOTH streams (:760-766, :804-810), context expansion
// (:778-785), target frame selection (:822-827) and the bf16 hi/lo operand split of the net input, in one pass.
struct ExpandArgs {
    const unsigned int *fea_rec, *targ_rec;   // [need][2 + fea_dim] / [need][2 + D] big-endian words
    const int *first;                         // [samples] first context frame of every (already shuffled) row
    const float *mean, *dvar;                 // [fea_dim]
    int samples, fea_dim, ctx, targ_offset, D;
    float *in32;                              // [samples][fea_dim*ctx] fp32 (validation path) or NULL
    bf16 *in_hi, *in_lo;                      // [samples][ld] (tensor path) or NULL
    int ld;
    float *targ;                              // [samples][D]
};
void launch_expand_chunk(const ExpandArgs &a, cudaStream_t s);

// Inference-side input builder (Test_code/decode.m:28-34 + frame_expand.m:6-25): z-score of the raw LPS frames with the .norm
// constants, then per frame t the context frames t-c..t+c with the utterance's first / last frame replicated at the edges.
struct EdgeExpandArgs {
    const float *lps;            // [frames][fea_dim] raw (un-normalised) features of ONE utterance
    const float *mean, *dvar;    // [fea_dim]
    int frames, fea_dim, ctx;
    float *in32;                 // [frames][fea_dim*ctx] fp32 (validation path) or NULL
    bf16 *in_hi, *in_lo;         // [frames][ld] (tensor path) or NULL
    int ld;
};
void launch_expand_edges(const EdgeExpandArgs &a, cudaStream_t s);
// out[f][d] = out[f][d] / dvar[d % fea_dim] + mean[d % fea_dim]   (decode.m:60-62)
void launch_denorm(float *out, long long frames, int D, const float *mean, const float *dvar, int fea_dim, cudaStream_t s);

struct LossArgs {
    const StepCtl *ctl;
    const float *out;   // [M][ldo] network output of this bunch (fp32)
    int ldo;
    int M;              // frames of this bunch on this rank
    int Mg;             // frames of the GLOBAL minibatch (== M on one GPU)
    int D;              // output units
    float beta;         // shapefactor
    int ml;             // MLflag == 1
    float *dx32;        // top-layer dE/dx, fp32 [M][ldx] (validation path) or NULL
    bf16 *dx_hi, *dx_lo;  // top-layer dE/dx as bf16 hi/lo [M][ldx] or NULL
    int ldx;
    float *alpha;       // [D] GGD scale factors (kept for CrossValid2)
    float *colsum;      // [D] sum_m |e|^beta  (partial in mode 1, global in mode 2)
    double *trace;      // per-bunch loss trace, indexed by ctl->bunch_idx (may be NULL)
    int mode;           // 0 = fused single pass-pair, 1 = column sums only, 2 = gradient from `colsum`,
                        // 3 = data-parallel in ONE kernel: partial column sums are stored into every rank's slot over
                        //     NVLink peer memory, flagged, and summed in rank order (dp_factor.cuh)
    float *asum_slot[8];            // mode 3: every rank's receive area [world][D]
    unsigned int *lflags[8];        // mode 3: every rank's flag block [world][FX_STRIDE]; loss flags at FX_EV_LOSS + block
    const unsigned int *step_counter;   // mode 3: completed data-parallel steps (flag value = *step_counter + 1)
    unsigned int *error_flag;
    int world, rank;
};
void launch_loss(const LossArgs &a, cudaStream_t s);

// bias gradients of ALL layers in one launch: dst[n] = sum_{m<M} dedx[m][n]   (kernAccSumrow, DevFunc.cu:267-285)
struct BiasGradLayer {
    const float *dx32;          // fp32 deltas (validation path) or NULL
    const bf16 *hi, *lo;        // bf16 hi/lo deltas
    int ld, N;
    float *dst;                 // gradient out (unfused path)
    float *b, *db;              // bias and its momentum (fused path: updated in place)
};
struct BiasGradArgs {
    BiasGradLayer layer[10];
    int nlayers, M;
    int apply;                  // 1: apply the momentum-SGD update to the biases here (no weight cost, BP_GPU.cu:435)
    float mom, lr, Mg;
    StepCtl *ctl;               // when set (last kernel of a fused step) the kernel advances ctl->bunch_idx
};
void launch_bias_grad(const BiasGradArgs &a, cudaStream_t s);

struct UpdSeg {
    long long off;   // element offset in the parameter arena (multiple of 4)
    long long goff;  // element offset of the matching gradients in G
    long long n;     // elements (multiple of 4)
    float wc;        // weight cost (0 for biases)
    int shadow;      // write bf16 hi/lo shadows
};
struct UpdArgs {
    UpdSeg seg[2 * 10];
    int nseg;
    float *P, *Dl;       // parameters, momentum
    const float *G;      // gradients (same arena layout)
    bf16 *Phi, *Plo;     // bf16 split shadows of P (weights only)
    float mom, lr, Mg;
    StepCtl *ctl;        // when set, the kernel also advances ctl->bunch_idx (it is the last kernel of a step)
};
// kernUpdatedelta + kernAccSum fused (DevFunc.cu:490-507, 427-443)
void launch_update(const UpdArgs &a, int sm_count, cudaStream_t s);

void launch_advance(StepCtl *ctl, cudaStream_t s);

// fp32 validation GEMM: C[i][j] = sum_r A[i*sAi + r*sAr] * B[j*sBj + r*sBr]
void launch_simt_gemm(const float *A, long sAi, long sAr, const float *B, long sBj, long sBr, float *C, long ldc,
                      int I, int J, int R, cudaStream_t s);
// x[m][n] + bias[n] -> sigmoid -> y  (or plain copy to y when linear)
void launch_simt_bias_act(const float *x, int ld, const float *bias, float *y, int M, int N, int linear, cudaStream_t s);
void launch_simt_dsigmoid(const float *y, const float *dedy, float *dedx, int ld, int M, int N, cudaStream_t s);
// A-operand view of the current bunch of the fp32 chunk input (validation path)
void launch_simt_gather_in(const StepCtl *ctl, int M, int K, float *dst, int ld, cudaStream_t s);

}  // namespace ggd
