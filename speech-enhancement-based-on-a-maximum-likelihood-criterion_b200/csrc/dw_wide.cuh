// dw_wide.cuh -- argument blocks of the wide gradient + update kernel (dw_wide.cu) and the flag layout of the
// factor-exchange data parallelism (dp_factor.cuh) it takes part in.
#pragma once
#include "gemm_tc.cuh"

namespace ggd {


struct DwwLayer {
    CUtensorMap a_hi, a_lo;   // dE/dx of this layer over the WHOLE minibatch, bf16 [rows][Np], box {64 units, 32 frames}, 128-byte swizzle
    CUtensorMap b_hi, b_lo;   // activations of the layer below, bf16 [rows][Kp], box {64 units, 32 frames}, 128-byte swizzle
    CUtensorMap w_map, d_map; // fp32 weights / momentum [Kp][Np], box {128 n, 16 k}, no swizzle (TMA load AND store)
    int Kp, Np;
    int k_slabs;              // Kp / 64
    int slab_base;            // index of this layer's first slab in the global list
    int n_slabs;              // slabs of this layer: ceil(Np / 128) * k_slabs
    int b_rows_from_ctl;      // add ctl->bunch_idx * rows_per_bunch to the frame coordinate of b_hi / b_lo (one GPU, layer 1)
    int ev_dx, ev_y;          // data parallel: flags that every peer must have raised before the operands are read (-1: none)
    const bf16 *dx_hi, *dx_lo; // the same dE/dx arrays as a_hi / a_lo (pitch Np), for the bias gradient
    float *b, *db;            // bias and its momentum
    int N;                    // real output units
    float wc;
};
struct DwwArgs {
    DwwLayer layer[10];       // in the order the kernel walks them: TOP layer first (its factors arrive first)
    int nlayers, total_slabs;
    int ngroups;              // slab groups walked one after the other by every CTA (dw_wide.cu: SegIter)
    int group_base[11];       // first slab of each group, then total_slabs
    StepCtl *ctl;
    int rows_per_bunch;
    int fblocks;              // frames of the whole (padded) minibatch / 32
    int rows;                 // = 32 * fblocks
    int op_stages, wd_stages; // ring depths (dw_wide_smem)
    float mom, lr, Mg;
    int advance;              // last CTA out increments ctl->bunch_idx (and the data-parallel step counter)
    unsigned int *done_counter;
    int l2_hints;
    unsigned int *hang;
    // data parallelism (world > 1)
    int world, rank;
    const unsigned int *flags;            // my flag block [world][FX_STRIDE]
    unsigned int *peer_flags[FX_MAX];     // every rank's flag block
    unsigned int *step_counter;           // completed data-parallel steps
    unsigned int *error_flag;
    unsigned long long *trace;            // optional: [FX_TRACE_WIDE + {start, flags of list layer 0..9 seen, end}] (block 0 / last block)
};
int dw_wide_smem(int fblocks, int *op_stages, int *wd_stages);
int launch_dw_wide(const DwwArgs *dev_args, int grid, int smem_bytes, cudaStream_t s);
int dw_wide_init();

}  // namespace ggd
