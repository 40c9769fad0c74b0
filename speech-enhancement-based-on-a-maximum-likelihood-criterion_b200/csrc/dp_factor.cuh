// dp_factor.cuh -- frame-sharded data parallelism by FACTOR exchange over NVLink peer memory.
//
// The weight gradient of a layer, dW_l = sum_m dx_l[m]^T y_{l-1}[m], has rank <= frames of the global minibatch while it
// is Np x Kp large: its factors (dE/dx_l and y_{l-1} of one rank's 128 frames, bf16 hi+lo: 7.4 MB for the named net) are
// 7x smaller than the gradient (51 MB).  So instead of reduce-scattering gradients and all-gathering weights
// (2 (N-1)/N 51 MB per rank and step, round 1), every rank
//   1. writes its own slice of every factor array straight from the GEMM epilogues into its FACTOR ARENA
//      (rows [rank*Mp, (rank+1)*Mp) of [world*Mp][units] arrays),
//   2. pushes that slice into every peer's arena with `factor_push_kernel` on a second stream the moment the producing
//      GEMM has finished -- the transfer runs under the rest of the forward / backward chain (the push CTAs use no
//      shared memory and co-reside with the GEMM CTAs) -- and raises a per-array flag at every peer,
//   3. runs the gradient + momentum update (weights and biases) REPLICATED over the whole minibatch (dw_wide.cu waits for
//      the flags layer by layer, top layer first).
// Every rank computes the same sums in the same order: weights stay bit-identical without ever being exchanged, and
// ggd_get_weights needs no gather.  The per-dimension sum_m|e|^beta of the GGD scale is exchanged the same way inside
// the loss epilogue (gemm_tc.cu) / loss_kernel mode 3, so alpha equals the unsharded minibatch's on every rank.
// NVLink bytes per rank and step: (N-1) x 7.4 MB out and in; no NCCL kernel is on the step (NCCL only carries the
// CUDA IPC handles at set-up and the host-side barriers).
// Reuse of the arenas across steps: the last CTA of dw_wide raises FX_EV_DONE at every peer; the first push of the next
// step waits for every peer's DONE before it overwrites their arenas.
#pragma once
#include "dw_wide.cuh"

namespace ggd {

constexpr int FX_TRACE_WIDE = 4 * FX_STRIDE, FX_TRACE_WORDS = FX_TRACE_WIDE + 16;

struct FxSeg {
    const uint8_t *src;          // local source (my slice)
    long long src_bunch_stride;  // bytes added per ctl->bunch_idx (the net-input rows live in the chunk arrays), else 0
    long long dst_off;           // byte offset of my slice in every rank's factor arena
    long long bytes;             // multiple of 16
};
struct FxPushArgs {
    FxSeg seg[2];
    int nseg;
    uint8_t *peer_arena[FX_MAX];          // every rank's factor arena
    unsigned int *peer_flags[FX_MAX];     // every rank's flag block [world][FX_STRIDE]
    const unsigned int *my_flags;
    const StepCtl *ctl;
    const unsigned int *step_counter;     // completed steps; this step's flag value = *step_counter + 1
    unsigned int *block_counter;          // one word per event
    unsigned int *error_flag;
    unsigned int *hang;
    int world, rank;
    int event;                            // FX_EV_Y + l or FX_EV_DX + l
    int wait_done;                        // first push of a step: wait until every peer has finished the previous step
    int include_self;                     // also copy into my own arena (net-input rows)
    unsigned long long *trace;            // optional globaltimer stamps [event*4 + {start, waited, copied, flagged}] (GGD_FX_TRACE=1)
};
void launch_factor_push(const FxPushArgs &a, int grid, cudaStream_t s);

}  // namespace ggd
