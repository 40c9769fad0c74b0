// gemm_tc.cuh -- the tcgen05 / TMEM / TMA GEMM of the training step.
//
// One kernel template serves the three dense contractions of BP_GPU::train_bunch_single
// (BP_GPU.cu:361 SgemmNN, :430 SgemmTN, :432 SgemmNT), written as
//        D[i][j] = sum_r A(i,r) * B(j,r)            (tile: 128 rows i  x  BN columns j)
// with each operand either K-major (r contiguous in memory) or MN-major (i / j contiguous):
//   forward   x[m][n]   = sum_k y[m][k] W[k][n]      A = y    K-major   B = W     MN-major
//   backward  dy[m][k]  = sum_n dx[m][n] W[k][n]     A = dx   K-major   B = W     K-major
//   gradient  g[k][n]   = sum_m y[m][k] dx[m][n]     A = y    MN-major  B = dx    MN-major
// so the weight matrix is never transposed in memory (it keeps the reference's .wts order).
//
// Arithmetic: every fp32 operand is stored as a bf16 pair (hi, lo); a product is formed as
// hi*hi + hi*lo + lo*hi by three kind::f16 MMAs into one fp32 TMEM accumulator (bf16x3).
#pragma once
#include "kernels.cuh"

namespace ggd {

enum GemmEpilogue {
    EPI_FWD_SIGMOID = 0,  // + bias, sigmoid, write bf16 hi/lo activations (kernMultiCopy + kernSigmoid fused)
    EPI_FWD_LINEAR = 1,   // + bias, write fp32 network output
    EPI_DX_DSIGMOID = 2,  // * y(1-y), write bf16 hi/lo dE/dx of the previous layer (kernDsigmoid fused)
    EPI_STORE_F32 = 3,    // plain fp32 store (weight gradient)
    EPI_FWD_LOSS = 4      // output layer of a TRAINING step: + bias, fp32 network output AND the whole loss-gradient chain
                          // (BP_GPU.cu:408-424) in the epilogue: e = out - targ, s_d = sum_m |e|^beta, alpha_d, dE/dx as bf16 hi/lo
};


struct GemmArgs {
    const StepCtl *ctl;
    int a_rows_from_ctl;   // add ctl->bunch_idx * rows_per_bunch to the ROW coordinate of A's tensor map
    int rows_per_bunch;
    int I, J;              // valid output extent; rows >= I or columns >= J are written as zeros
    int kblocks;           // reduction length / 64
    const float *bias;     // EPI_FWD_*
    bf16 *o_hi, *o_lo;     // bf16 outputs, pitch ldo
    int ldo;
    float *o32;            // fp32 output, pitch ld32
    int ld32;
    const bf16 *y_hi, *y_lo;  // EPI_DX_DSIGMOID: activations of the layer whose dE/dx is produced, pitch ldy
    int ldy;
    unsigned long long *trace;   // optional [ctas][16] globaltimer stamps of the pipeline phases (ggd_debug_gemm_timed)
    unsigned int *hang;          // host-mapped watchdog record (pipe.cuh: mbar_wait_bounded); may be NULL
    int b_early;                 // B_F32: the weight tiles may be loaded before griddepcontrol.wait (they were not written by the
                                 // kernel launched just before this one)
    // ---- EPI_FWD_LOSS only
    int D;                 // real output units (== J)
    int Mg;                // frames of the GLOBAL minibatch
    float beta;            // shapefactor
    int ml;                // MLflag == 1
    float *alpha;          // [D]
    double *loss_trace;    // per-bunch loss, indexed by ctl->bunch_idx (may be NULL)
    int world, rank;       // world > 1: partial column sums are exchanged over peer memory (dp_factor.cuh)
    float *asum_slot[8];   // every rank's receive area [world][D]
    unsigned int *lflags[8];   // every rank's flag block [world][FX_STRIDE] (kernels.cuh); the loss flags sit at FX_EV_LOSS
    const unsigned int *step_counter;
    unsigned int *error_flag;
};

struct GemmPlan {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    GemmArgs args;
    int bn;          // 64 or 128
    int a_mn, b_mn;  // operand majors
    int b_f32;       // B operand is the fp32 master weight matrix (b_hi = fp32 tensor map, box {64, 64 | bn}, no swizzle): split in-kernel
    int epi;
    int splits;      // cluster size along the reduction (1, 2, 4 or 8)
    int no_pdl;      // launch WITHOUT programmatic stream serialization (first kernel after a cross-stream join)
    int tiles_i, tiles_j;
};

// 2-D bf16 row-major tensor [rows][cols] (pitch ld elements) with a {64, box_rows} box, 128-byte swizzle
int make_tmap_bf16(CUtensorMap *m, const bf16 *base, long long rows, long long cols, long long ld, int box_rows);

// generic 2-D row-major tensor map: fp32 (dtype_f32 = 1) or bf16 elements, {box_cols, box_rows} box, optional 128-byte swizzle
int make_tmap_2d(CUtensorMap *m, const void *base, int dtype_f32, long long rows, long long cols, long long ld, int box_cols, int box_rows,
                 int swizzle128);

int launch_gemm_tc(const GemmPlan &p, cudaStream_t s);

// ---- persistent gradient + update kernel for ALL layers and biases in one launch (dw_persist.cu; Mp == 128) ---------
// The argument block lives in device memory (its tensor maps are read by TMA from there).
struct DwpLayer {
    CUtensorMap a_hi, a_lo;   // dE/dx of this layer, bf16 [Mp][Np], box {64 units, 64 frames}
    CUtensorMap b_hi, b_lo;   // activations of the layer below (chunk input for layer 1), bf16 [rows][Kp], box {64, 64}
    CUtensorMap w_map, d_map; // fp32 weights / momentum [Kp][Np], box {128 n, 16 k}, no swizzle (TMA load AND store)
    CUtensorMap hi_map, lo_map;  // bf16 shadows [Kp][Np], box {128 n, 16 k}, no swizzle (TMA store)
    float *W, *D;             // fp32 weights / momentum [Kp][Np]
    bf16 *w_hi, *w_lo;        // shadows [Kp][Np]
    float *b, *db;            // bias and its momentum
    const bf16 *dx_hi, *dx_lo;  // same arrays as a_hi / a_lo, for the bias gradient
    int Kp, Np, N;            // padded dims, real output units
    int k_tiles;              // Kp / 64
    int tile_base;            // index of this layer's first tile in the global list
    int b_rows_from_ctl;      // add ctl->bunch_idx * rows_per_bunch to the frame coordinate of b_hi / b_lo
    float wc;
    int pad;
};
struct DwpArgs {
    DwpLayer layer[10];
    int nlayers, total_tiles;
    StepCtl *ctl;
    int rows_per_bunch, M;
    float mom, lr, Mg;
    int advance;                  // last CTA out increments ctl->bunch_idx
    unsigned int *done_counter;
    int shadows;                  // 1: also write the bf16 hi/lo shadows of the weights (GEMMs in shadow mode, GGD_W_F32=0)
    int l2_hints;                 // evict-first fp32 streams, evict-last shadows (GGD_L2_HINTS, default 1)
    unsigned int *hang;           // host-mapped [8]: filled by a waiter that gave up (see mbar_wait_bounded)
};
int launch_dw_persist(const DwpArgs *dev_args, int grid, int shadows, cudaStream_t s);
int dw_persist_init();
int gemm_tc_init();   // resolves the driver entry point, sets the shared-memory attributes

}  // namespace ggd
