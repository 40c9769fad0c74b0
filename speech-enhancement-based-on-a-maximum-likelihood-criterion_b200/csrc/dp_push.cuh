// dp_push.cuh -- frame-sharded data parallelism over NVLink peer memory, push model.
//
// The global tile list of dw_persist.cu (layer, 128-unit n-tile, 64-unit k-tile) is cut into `world` contiguous
// ranges; rank o OWNS range o: it holds the fp32 master weights + momentum of those tiles.  Per step and rank:
//   K1  dw_push_kernel        persistent tcgen05 kernel: the gradient tile of EVERY tile from this rank's 128 frames,
//                             TMEM -> shared memory -> TMA store straight into the owner's receive slot
//                             [source rank][tile] (peer memory for foreign tiles).  Bias-gradient partial sums are
//                             stored into every rank's bias slot.  When all stores of the grid have completed, the
//                             rank raises flag A at every peer.
//   K2  reduce_update_kernel  persistent TMA-pipelined kernel over the OWNED tiles: waits for flag A of every peer,
//                             sums the `world` partial tiles in rank order (deterministic), applies the momentum-SGD
//                             update to its master copy and TMA-stores the new bf16 hi/lo operand shadows into EVERY
//                             rank's shadow arrays.  Biases: every rank sums the bias slots in rank order and updates
//                             its own copy (identical everywhere).  Flag B (raise + wait) closes the step: no rank
//                             starts the next forward before every shadow slice has landed.
//   loss_kernel mode 3        the per-dimension sum_m |e|^beta of the GGD scale (257 floats) is exchanged the same way
//                             (store partials into every rank's slot, flag, sum in rank order) inside ONE kernel, so
//                             alpha equals the unsharded minibatch's on every rank.
// NVLink traffic per rank and step: (N-1)/N * 4 B/param out (gradient tiles) + (N-1)/N * 4 B/param out (shadows of the
// owned tiles to N-1 peers), the same bytes as reduce-scatter + all-gather; no NCCL kernel is on the step.
#pragma once
#include "gemm_tc.cuh"

namespace ggd {

constexpr int DPX_MAX = 8;

struct DpxLayer {
    CUtensorMap a_hi, a_lo;   // dE/dx of this layer, bf16 [Mp][Np], box {64 units, 64 frames}, 128-byte swizzle
    CUtensorMap b_hi, b_lo;   // activations of the layer below, bf16 [rows][Kp], box {64, 64}, 128-byte swizzle
    CUtensorMap w_map, d_map; // local fp32 weights / momentum [Kp][Np], box {128 n, 8 k}
    CUtensorMap hi_map[DPX_MAX], lo_map[DPX_MAX];   // bf16 shadows of EVERY rank [Kp][Np], box {128 n, 8 k}
    CUtensorMap wp_map[DPX_MAX];                    // fp32 weights of EVERY rank (w_f32 mode: the updated weights themselves are broadcast)
    const bf16 *dx_hi, *dx_lo;
    float *b, *db;
    int Kp, Np, N, k_tiles, tile_base, b_rows_from_ctl, bias_off, pad;
    float wc, pad2;
};

struct DpxArgs {
    DpxLayer layer[10];
    CUtensorMap push_map[DPX_MAX];   // rank o's receive slot for MY partial tiles: fp32 [slot_tiles*64][128], box {128, 16}
    CUtensorMap part_map[DPX_MAX];   // MY receive slot holding source p's partial tiles: same geometry, box {128, 8}
    float *bias_slot[DPX_MAX];       // every rank's bias receive area [world][nbias]; this rank writes row `rank`
    float *asum_slot[DPX_MAX];       // every rank's sum|e|^beta receive area [world][D]
    unsigned int *flags[DPX_MAX];    // every rank's flag block: A [DPX_MAX], B [DPX_MAX], loss [DPX_MAX][16]
    unsigned int *counters;          // local: {step, k1_done, k2_done}
    unsigned int *error_flag;        // local: 1 + rank that did not arrive in time
    unsigned int *hang;
    unsigned long long *trace;       // optional (GGD_DPX_TRACE=1): [2 kernels][grid][8] globaltimer stamps of the last step
    StepCtl *ctl;
    int own_begin[DPX_MAX + 1];      // owner o holds tiles [own_begin[o], own_begin[o+1])
    int nlayers, total_tiles, world, rank, nbias, rows_per_bunch, M;
    int k2_stages, k2_stage_bytes;
    int w_f32, pad4;                 // 1: broadcast fp32 weights into every rank's master array instead of bf16 shadows
    float mom, lr, Mg;
};

constexpr int DPX_FLAG_A = 0, DPX_FLAG_B = DPX_MAX, DPX_FLAG_LOSS = 2 * DPX_MAX, DPX_FLAG_WORDS = 2 * DPX_MAX + LOSS_FLAGS_PER_RANK * DPX_MAX;

int dp_push_init();
int dp_push_k2_smem(int world, int *stages, int *stage_bytes);
int launch_dw_push(const DpxArgs *dev_args, int grid, cudaStream_t s);
int launch_reduce_update(const DpxArgs *dev_args, int grid, int smem_bytes, cudaStream_t s);
// copies the fp32 master weights of the tiles owned by other ranks from their owners (before exporting weights)
void launch_gather_master(const DpxArgs *dev_args, float *const *peerP, const long long *w_off, int grid, cudaStream_t s);

}  // namespace ggd
