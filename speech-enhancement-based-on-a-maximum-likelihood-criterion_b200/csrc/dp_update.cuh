// dp_update.cuh -- fused cross-GPU gradient reduce-scatter + sharded momentum-SGD update + all-gather of the operand
// shadows, in ONE kernel per step over NVLink peer memory (frame-sharded data parallelism).
//
// Rank r owns the r-th slice of the parameter arena.  After its backward pass every rank
//   1. tells all peers "my gradient arena of step s is complete" (flag store into the peers' memory) and waits for theirs,
//   2. for its slice: g = sum_p G_p[i] read straight from the peers' HBM over NVLink (fixed rank order: deterministic),
//      delta <- mom*delta - lr*(g/Mg + wc*W), W <- W + delta on ITS master copy only (update cost / world),
//      and stores the new bf16 hi/lo operand shadows (and fp32 biases) into EVERY rank's memory,
//   3. after a system-scope fence signals "my slice is published" and waits until every peer has published.
// NVLink traffic per rank and step: (N-1)/N * 4 B/param in (gradients) and out (shadows): the same bytes as a ring
// allreduce, but without the replicated update and without a separate NCCL kernel competing for SMs.
#pragma once
#include "kernels.cuh"

namespace ggd {

constexpr int DP_MAX_RANKS = 8;

struct DpPiece {          // intersection of this rank's slice with one weight or bias segment
    long long off;        // element offset in the arena (P, Dl, shadows)
    long long goff;       // element offset of its gradients in G
    long long n;          // elements (multiple of 4)
    float wc;
    int shadow;           // 1: weights (publish bf16 hi/lo);  0: biases (publish fp32 into the peers' P)
};

struct DpArgs {
    float *G[DP_MAX_RANKS];           // every rank's gradient arena (index = rank; own entry = local pointer)
    bf16 *hi[DP_MAX_RANKS], *lo[DP_MAX_RANKS];
    float *P[DP_MAX_RANKS];
    unsigned int *flags[DP_MAX_RANKS];   // [2][DP_MAX_RANKS] per rank: phase A / phase B arrival counters
    float *Dl;                        // local momentum
    DpPiece piece[24];
    int npieces;
    int world, rank;
    float mom, lr, Mg;
    unsigned int *step_counter;       // local: number of completed dp_update launches (the flag value to publish)
    unsigned int *block_counter;      // local: blocks that finished their stores in this launch
    unsigned int *error_flag;         // local: set when a peer did not arrive in time
    StepCtl *ctl;
};

void launch_dp_update(const DpArgs &a, int blocks, cudaStream_t s);

}  // namespace ggd
