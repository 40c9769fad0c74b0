// dw_persist.cu -- ONE persistent launch for the weight gradients, the momentum-SGD update of every layer and the bias
// gradients + bias update (BP_GPU.cu:432-437: SgemmNT, updatedelta, DevAccSum, DevAccSumrow, updatedelta, DevAccSum for
// all layers), used when the bunch fits one reduction tile (Mp == 128 frames).
//
// Grid = one CTA per SM; every CTA owns a CONTIGUOUS range of the global tile list (layer, n-tile, k-tile; k fastest).
// A tile is 128 output units n  x  64 input units k of one weight matrix:
//     g[n][k] = sum_m dx[m][n] * y[m][k]        UMMA: M = 128 (n, TMEM lanes), N = 64 (k, TMEM columns), K = 128 frames
// A TMEM lane is an OUTPUT unit n, and n is the contiguous index of the reference weight layout (W[k][n], index =
// out + in*cur), so a warp's 32 lanes work on 32 consecutive floats of one W row (conflict-free shared-memory rows).
//
// ALL global traffic is TMA; the SM's load/store units only touch shared memory:
//   warp 0      producer: dx^T of the current n-tile (128 n x 128 frames bf16 hi/lo, 64 KB, resident while the CTA walks
//               the k-tiles), y^T of the tile (64 k x 128 frames, 32 KB), and the fp32 weights + momentum in quarter
//               tiles of 16 k rows (W and delta, 16 KB per stage, 5-stage ring) -- ~60 KB of HBM reads in flight per SM
//               without holding a register;
//   warp 1      TMEM allocator + single-thread tcgen05 MMA issuer (bf16x3), accumulator double-buffered in TMEM so the
//               MMAs of tile t+1 run under the update of tile t;
//   warps 2..9  update warps: every warp takes part in every ring stage (8 rows x 32 n each): gradient from TMEM,
//               W / delta from shared memory, delta <- mom*delta - lr*(g/Mg + wc*W), W <- W + delta written back IN
//               PLACE (and, in shadow mode only, the bf16 hi/lo shadows into the stage's shadow buffers);
//   warp 10     store warp: one TMA store per array and stage (W, delta [, hi, lo]), and it releases the stage to the
//               producer once the store engine has read it;
//   warp 11     bias warp: column sums of dx over the frames and the bias update for the n-tiles whose k-tile 0 belongs
//               to this CTA (latency-bound L2 reads, hidden beside the tile pipeline).
// Every consumer follows every ring stage in order: a parity wait cannot tell phase f from phase f+2, so no warp may
// ever skip a stage (TMA loads land out of order).
// HBM traffic: 16 B/param (shadow mode: +4 B/param of bf16 shadows; by default the GEMMs read the fp32 weights and split them
// in-kernel, gemm_tc.cu B_F32).  The last CTA to finish advances the device-side bunch counter.
#include "gemm_tc.cuh"
#include "pipe.cuh"
#include "../../include/ggd_train.h"

namespace ggd {

namespace dwp {
constexpr int TN = 128, TK = 64, BK = 64;
constexpr int KB = 2;                              // 128 frames = 2 reduction blocks of 64
constexpr int A_HALF = 64 * BK * 2;                // 8 KB: 64 units x 64 frames bf16
constexpr int A_PART = 2 * A_HALF;                 // 16 KB: 128 n x 64 frames (hi or lo)
constexpr int A_SLOT = KB * 2 * A_PART;            // 64 KB
constexpr int B_PART = TK * BK * 2;                // 8 KB
constexpr int B_STAGE = KB * 2 * B_PART;           // 32 KB
constexpr int WD_ROWS = 16;                        // k rows per ring stage (a quarter tile)
constexpr int QUARTERS = TK / WD_ROWS;             // 4
constexpr int WD_F32 = WD_ROWS * TN * 4;           // 8 KB: 16 k x 128 n fp32
constexpr int WD_B16 = WD_ROWS * TN * 2;           // 4 KB
constexpr int WD_LOAD = 2 * WD_F32;                // bytes that arrive by TMA per stage
// shadow mode: a stage also stages the bf16 hi/lo tiles (24 KB, 5 stages); without shadows 16 KB, 7 stages
template <bool SHADOWS> struct Ring {
    static constexpr int STAGE = 2 * WD_F32 + (SHADOWS ? 2 * WD_B16 : 0);
    static constexpr int STAGES = SHADOWS ? 5 : 8;
    static constexpr int SMEM = A_SLOT + B_STAGE + STAGES * STAGE + 1024;
};
constexpr int NTHREADS = 384;
constexpr int TMEM_COLS = 2 * TK;                  // double-buffered accumulator
}  // namespace dwp

struct TileRef {
    const DwpLayer *L;
    int nt, kt;
    int key;   // identifies the (layer, n-tile) pair, i.e. the resident dx^T operand
};

__device__ __forceinline__ TileRef decode_tile(const DwpArgs *gp, int t)
{
    int l = 0;
#pragma unroll 1
    while (l + 1 < gp->nlayers && t >= gp->layer[l + 1].tile_base) l++;
    const DwpLayer *L = &gp->layer[l];
    const int r = t - L->tile_base;
    TileRef tr;
    tr.L = L; tr.nt = r / L->k_tiles; tr.kt = r - tr.nt * L->k_tiles;
    tr.key = (l << 16) | tr.nt;
    return tr;
}

template <bool SHADOWS>
__global__ void __launch_bounds__(dwp::NTHREADS, 1) dw_persist_kernel(const DwpArgs *__restrict__ gp)
{
    using namespace dwp;
    constexpr int WD_STAGE = Ring<SHADOWS>::STAGE, WD_STAGES = Ring<SHADOWS>::STAGES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *a_slot = smem, *b_stage = smem + A_SLOT, *wd_ring = b_stage + B_STAGE;
    __shared__ __align__(8) uint64_t a_full, a_empty, b_full, b_empty, t_full[2], t_empty[2];
    __shared__ __align__(8) uint64_t wd_full[WD_STAGES], wd_done[WD_STAGES], wd_empty[WD_STAGES];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = gp->total_tiles;
    const int t0 = (int)((long long)T * blockIdx.x / gridDim.x), t1 = (int)((long long)T * (blockIdx.x + 1) / gridDim.x);
    unsigned int *const hang = gp->hang;

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(&a_full, 1); mbar_init(&a_empty, 1); mbar_init(&b_full, 1); mbar_init(&b_empty, 1);
            for (int s = 0; s < 2; s++) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 8); }
            for (int s = 0; s < WD_STAGES; s++) { mbar_init(&wd_full[s], 1); mbar_init(&wd_done[s], 8); mbar_init(&wd_empty[s], 1); }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<TMEM_COLS>(&tmem_base_s);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    pdl_wait();   // the deltas / activations of this step are complete and visible from here on
    const int bunch_row0 = gp->ctl->bunch_idx * gp->rows_per_bunch;

    if (warp == 0) {
        if (lane == 0 && t0 < t1) {
            // ===== TMA producer =====
            int key = -1, a_cnt = 0;
            const uint64_t pol_stream = l2_policy_evict_first();
            for (int t = t0, it = 0; t < t1; t++, it++) {
                const TileRef tr = decode_tile(gp, t);
                const DwpLayer *L = tr.L;
                if (tr.key != key) {
                    key = tr.key;
                    mbar_wait_bounded(&a_empty, (a_cnt & 1) ^ 1, hang, 1, it);
                    a_cnt++;
                    mbar_expect_tx(&a_full, A_SLOT);
#pragma unroll
                    for (int kb = 0; kb < KB; kb++)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            tma_load_2d(a_slot + kb * 2 * A_PART + h * A_HALF, &L->a_hi, &a_full, tr.nt * TN + 64 * h, kb * BK);
                            tma_load_2d(a_slot + kb * 2 * A_PART + A_PART + h * A_HALF, &L->a_lo, &a_full, tr.nt * TN + 64 * h, kb * BK);
                        }
                }
                mbar_wait_bounded(&b_empty, (it & 1) ^ 1, hang, 2, it);
                mbar_expect_tx(&b_full, B_STAGE);
                const int r0 = L->b_rows_from_ctl ? bunch_row0 : 0;
#pragma unroll
                for (int kb = 0; kb < KB; kb++) {
                    tma_load_2d(b_stage + kb * 2 * B_PART, &L->b_hi, &b_full, tr.kt * TK, r0 + kb * BK);
                    tma_load_2d(b_stage + kb * 2 * B_PART + B_PART, &L->b_lo, &b_full, tr.kt * TK, r0 + kb * BK);
                }
                // fp32 weights / momentum of this tile: four quarter tiles of 16 k rows
#pragma unroll
                for (int qt = 0; qt < QUARTERS; qt++) {
                    const int seq = QUARTERS * it + qt, ws = seq % WD_STAGES, wph = (seq / WD_STAGES) & 1;
                    mbar_wait_bounded(&wd_empty[ws], wph ^ 1, hang, 3, it);
                    mbar_expect_tx(&wd_full[ws], WD_LOAD);
                    uint8_t *wdst = wd_ring + ws * WD_STAGE;
                    if (gp->l2_hints) {
                        tma_load_2d_hint(wdst, &L->w_map, &wd_full[ws], tr.nt * TN, tr.kt * TK + qt * WD_ROWS, pol_stream);
                        tma_load_2d_hint(wdst + WD_F32, &L->d_map, &wd_full[ws], tr.nt * TN, tr.kt * TK + qt * WD_ROWS, pol_stream);
                    } else {
                        tma_load_2d(wdst, &L->w_map, &wd_full[ws], tr.nt * TN, tr.kt * TK + qt * WD_ROWS);
                        tma_load_2d(wdst + WD_F32, &L->d_map, &wd_full[ws], tr.nt * TN, tr.kt * TK + qt * WD_ROWS);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && t0 < t1) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = make_idesc_bf16(TN, TK, true, true);
            int key = -1, a_cnt = 0;
            TileRef tr = decode_tile(gp, t0);
            for (int t = t0, it = 0; t < t1; t++, it++) {
                if (tr.key != key) {
                    key = tr.key;
                    mbar_wait_bounded(&a_full, a_cnt & 1, hang, 4, it);
                    a_cnt++;
                }
                mbar_wait_bounded(&b_full, it & 1, hang, 5, it);
                const int acc = it & 1;
                mbar_wait_bounded(&t_empty[acc], ((it >> 1) & 1) ^ 1, hang, 6, it);
                tc_fence_after();
                const uint32_t a0 = smem_u32(a_slot), b0 = smem_u32(b_stage);
                const uint32_t d = tmem + acc * TK;
#pragma unroll
                for (int kb = 0; kb < KB; kb++) {
                    const uint32_t a_hi = a0 + kb * 2 * A_PART, a_lo = a_hi + A_PART;
                    const uint32_t b_hi = b0 + kb * 2 * B_PART, b_lo = b_hi + B_PART;
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) {
                        const uint64_t dah = make_smem_desc(a_hi + k * 2048, 8192, 1024), dal = make_smem_desc(a_lo + k * 2048, 8192, 1024);
                        const uint64_t dbh = make_smem_desc(b_hi + k * 2048, 8192, 1024), dbl = make_smem_desc(b_lo + k * 2048, 8192, 1024);
                        umma_bf16(d, dal, dbh, idesc, (kb | k) != 0);   // small terms first
                        umma_bf16(d, dah, dbl, idesc, 1);
                        umma_bf16(d, dah, dbh, idesc, 1);
                    }
                }
                umma_commit(&b_empty);
                umma_commit(&t_full[acc]);
                TileRef nx = tr;
                if (t + 1 < t1) nx = decode_tile(gp, t + 1);
                if (t + 1 >= t1 || nx.key != key) umma_commit(&a_empty);   // last tile that reads this dx^T operand
                tr = nx;
            }
        }
        __syncwarp();
    } else if (warp == 10) {
        if (lane == 0 && t0 < t1) {
            // ===== store warp: TMA stores of finished stages; a stage goes back to the producer once it has been read =====
            int prev_ws = -1;
            const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
            for (int t = t0, it = 0; t < t1; t++, it++) {
                const TileRef tr = decode_tile(gp, t);
                const DwpLayer *L = tr.L;
#pragma unroll
                for (int qt = 0; qt < QUARTERS; qt++) {
                    const int seq = QUARTERS * it + qt, ws = seq % WD_STAGES;
                    mbar_wait_bounded(&wd_done[ws], (seq / WD_STAGES) & 1, hang, 10, it);
                    const uint8_t *src = wd_ring + ws * WD_STAGE;
                    const int c0 = tr.nt * TN, c1 = tr.kt * TK + qt * WD_ROWS;
                    if (gp->l2_hints) {
                        // without shadows the fp32 weights themselves are what the next step's GEMMs read: keep them in L2
                        tma_store_2d_hint(&L->w_map, src, c0, c1, SHADOWS ? pol_stream : pol_keep);
                        tma_store_2d_hint(&L->d_map, src + WD_F32, c0, c1, pol_stream);
                        if (SHADOWS) {
                            tma_store_2d_hint(&L->hi_map, src + 2 * WD_F32, c0, c1, pol_keep);
                            tma_store_2d_hint(&L->lo_map, src + 2 * WD_F32 + WD_B16, c0, c1, pol_keep);
                        }
                    } else {
                        tma_store_2d(&L->w_map, src, c0, c1);
                        tma_store_2d(&L->d_map, src + WD_F32, c0, c1);
                        if (SHADOWS) {
                            tma_store_2d(&L->hi_map, src + 2 * WD_F32, c0, c1);
                            tma_store_2d(&L->lo_map, src + 2 * WD_F32 + WD_B16, c0, c1);
                        }
                    }
                    tma_store_commit();
                    if (prev_ws >= 0) {
                        tma_store_wait_read<1>();            // everything but the newest group has left shared memory
                        mbar_arrive(&wd_empty[prev_ws]);
                    }
                    prev_ws = ws;
                }
            }
            tma_store_wait_all<0>();   // all writes performed before the CTA (and with it the grid) completes
        }
        __syncwarp();
    } else if (warp == 11) {
        // ===== bias warp: bias gradient + bias update of every n-tile whose k-tile 0 belongs to this CTA (kernAccSumrow,
        // DevFunc.cu:267-285; BP_GPU.cu:434-437).  It only needs dx, so it runs beside the tile pipeline from the start;
        // lane = 4 consecutive units, frames summed in ascending order like the reference.
        const float mom = gp->mom, lr = gp->lr;
        const int M = gp->M;
        for (int t = t0; t < t1; t++) {
            const TileRef tr = decode_tile(gp, t);
            if (tr.kt != 0) continue;
            const DwpLayer *L = tr.L;
            const int Np = L->Np, n = tr.nt * TN + 4 * lane;
            if (n >= Np) continue;
            const bf16 *xh = L->dx_hi + n, *xl = L->dx_lo + n;
            float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            int m = 0;
            for (; m + 8 <= M; m += 8) {
                uint2 vh[8], vl[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    vh[u] = *reinterpret_cast<const uint2 *>(xh + (size_t)(m + u) * Np);
                    vl[u] = *reinterpret_cast<const uint2 *>(xl + (size_t)(m + u) * Np);
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    s[0] += __uint_as_float(vh[u].x << 16) + __uint_as_float(vl[u].x << 16);
                    s[1] += __uint_as_float(vh[u].x & 0xFFFF0000u) + __uint_as_float(vl[u].x & 0xFFFF0000u);
                    s[2] += __uint_as_float(vh[u].y << 16) + __uint_as_float(vl[u].y << 16);
                    s[3] += __uint_as_float(vh[u].y & 0xFFFF0000u) + __uint_as_float(vl[u].y & 0xFFFF0000u);
                }
            }
            for (; m < M; m++) {
                const uint2 vh = *reinterpret_cast<const uint2 *>(xh + (size_t)m * Np), vl = *reinterpret_cast<const uint2 *>(xl + (size_t)m * Np);
                s[0] += __uint_as_float(vh.x << 16) + __uint_as_float(vl.x << 16);
                s[1] += __uint_as_float(vh.x & 0xFFFF0000u) + __uint_as_float(vl.x & 0xFFFF0000u);
                s[2] += __uint_as_float(vh.y << 16) + __uint_as_float(vl.y << 16);
                s[3] += __uint_as_float(vh.y & 0xFFFF0000u) + __uint_as_float(vl.y & 0xFFFF0000u);
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if (n + c < L->N) {
                    const float db = mom * L->db[n + c] - lr * (s[c] / gp->Mg);   // no weight cost on biases (BP_GPU.cu:435)
                    L->db[n + c] = db;
                    L->b[n + c] = db + L->b[n + c];
                }
            }
        }
    } else if (t0 < t1) {
        // ===== update warps (8): quadrant q = accumulator lanes [32q, 32q+32); `h` = which 8 of a stage's 16 rows =====
        const int e = warp - 2, q = warp & 3, h = e >> 2;
        const float mom = gp->mom, lr = gp->lr, inv_mg = 1.0f / gp->Mg;
        constexpr bool shadows = SHADOWS;
        TileRef tr = decode_tile(gp, t0);
        for (int t = t0, it = 0; t < t1; t++, it++) {
            const DwpLayer *L = tr.L;
            const float wc = L->wc;
            const int acc = it & 1;
            mbar_wait_bounded(&t_full[acc], (it >> 1) & 1, hang, 8, it);
            __syncwarp();
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + acc * TK + h * 8;
#pragma unroll
            for (int qt = 0; qt < QUARTERS; qt++) {
                const int seq = QUARTERS * it + qt, ws = seq % WD_STAGES;
                uint8_t *st = wd_ring + ws * WD_STAGE;
                float *sw = reinterpret_cast<float *>(st) + (h * 8) * TN + q * 32 + lane;
                float *sd = sw + WD_F32 / 4;
                bf16 *shi = reinterpret_cast<bf16 *>(st + 2 * WD_F32) + (h * 8) * TN + q * 32 + lane;
                bf16 *slo = shi + WD_B16 / 2;
                float g[8];
                tmem_ld8(taddr + qt * WD_ROWS, g);
                if (qt == QUARTERS - 1) {   // the accumulator has been drained by this warp: hand it back to the MMA issuer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[acc]);
                }
                mbar_wait_bounded(&wd_full[ws], (seq / WD_STAGES) & 1, hang, 7, it);
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    // kernUpdatedelta + kernAccSum (DevFunc.cu:490-507, 427-443); g/Mg as g*(1/Mg) (<= 1 ulp)
                    const float ww = sw[x * TN];
                    const float dd = mom * sd[x * TN] - lr * (g[x] * inv_mg + wc * ww);
                    const float wn = dd + ww;
                    sw[x * TN] = wn;
                    sd[x * TN] = dd;
                    if (shadows) {
                        bf16 hv, lv;
                        split_bf16(wn, hv, lv);
                        shi[x * TN] = hv;
                        slo[x * TN] = lv;
                    }
                }
                fence_async_proxy();      // the stage is read by the TMA store engine next
                __syncwarp();
                if (lane == 0) mbar_arrive(&wd_done[ws]);
            }
            if (t + 1 < t1) tr = decode_tile(gp, t + 1);
        }
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem);
    // last CTA out moves the device-side bunch counter on (every CTA has read it above)
    if (threadIdx.x == 0 && gp->advance) {
        __threadfence();
        const unsigned int prev = atomicAdd(gp->done_counter, 1u);
        if (prev == gridDim.x - 1) {
            *gp->done_counter = 0;
            gp->ctl->bunch_idx += 1;
        }
    }
}

int launch_dw_persist(const DwpArgs *dev_args, int grid, int shadows, cudaStream_t s)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(dwp::NTHREADS);
    cfg.dynamicSmemBytes = shadows ? dwp::Ring<true>::SMEM : dwp::Ring<false>::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (shadows) GGD_CUDA(cudaLaunchKernelEx(&cfg, dw_persist_kernel<true>, dev_args));
    else GGD_CUDA(cudaLaunchKernelEx(&cfg, dw_persist_kernel<false>, dev_args));
    return GGD_OK;
}

int dw_persist_init()
{
    GGD_CUDA(cudaFuncSetAttribute(dw_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dwp::Ring<true>::SMEM));
    GGD_CUDA(cudaFuncSetAttribute(dw_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dwp::Ring<false>::SMEM));
    return GGD_OK;
}

}  // namespace ggd
