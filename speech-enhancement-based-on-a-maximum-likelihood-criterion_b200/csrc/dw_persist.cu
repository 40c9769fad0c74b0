// dw_persist.cu -- ONE persistent launch for the weight gradients, the momentum-SGD update of every layer and the bias
// gradients + bias update (BP_GPU.cu:432-437: SgemmNT, updatedelta, DevAccSum, DevAccSumrow, updatedelta, DevAccSum for
// all layers), used when the bunch fits one reduction tile (Mp == 128 frames).
//
// Grid = one CTA per SM; every CTA owns a CONTIGUOUS range of the global tile list (layer, n-tile, k-tile; k fastest).
// A tile is 128 output units n  x  64 input units k of one weight matrix:
//     g[n][k] = sum_m dx[m][n] * y[m][k]        UMMA: M = 128 (n, TMEM lanes), N = 64 (k, TMEM columns), K = 128 frames
// The operand roles are swapped with respect to dw_update.cu on purpose: a TMEM lane is an OUTPUT unit n, and n is the
// contiguous index of the reference weight layout (W[k][n], index = out + in*cur).  The 32 lanes of an update warp
// therefore address 32 consecutive floats of one W row: the fp32 weight / momentum / bf16-shadow traffic is fully
// coalesced straight from registers, with no shared-memory transposition and no staging buffer.
//
// CTA = 320 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = update warps.
//   * dx^T tile of the current n-tile (128 n x 128 frames, bf16 hi/lo, 64 KB) stays RESIDENT while the CTA walks the
//     k-tiles; only the y^T tile (64 k x 128 frames, 32 KB) streams through a 2-stage ring.
//   * the accumulator is double-buffered in TMEM (2 x 64 columns): the MMAs of tile t+1 run under the update of tile t.
//   * the fp32 weights and momentum stream in by TMA as well (half tiles: 32 k x 128 n of W and of delta, 32 KB per
//     stage, 3 stages = 1.5 tiles ahead of the update warps), so ~96 KB of HBM reads per SM are always in flight
//     without holding a single register; the update warps read them conflict-free from shared memory (lane = n) and
//     store W, delta and the bf16 shadows straight from registers: 16 B/param (+4 B/param of bf16 shadows).
// The CTA that owns k-tile 0 of an n-tile also forms the bias gradient of those 128 units (column sums of dx over the
// frames) and applies the bias update; the last CTA to finish advances the device-side bunch counter.
#include "gemm_tc.cuh"
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
constexpr int A_SLOTS = 1, B_STAGES = 2;
constexpr int WD_ROWS = 32;                        // k rows per W/delta half-tile stage
constexpr int WD_PART = WD_ROWS * TN * 4;          // 16 KB: 32 k x 128 n fp32
constexpr int WD_STAGE = 2 * WD_PART;              // W then delta
constexpr int WD_STAGES = 3;
constexpr int SMEM = A_SLOTS * A_SLOT + B_STAGES * B_STAGE + WD_STAGES * WD_STAGE + 1024;
constexpr int NTHREADS = 320;
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

// Bounded mbarrier wait: a pipeline bug must surface as an error, never as a hung GPU.  After ~2^22 failed probes
// (seconds) the waiter records {code, block, iteration, parity} in host-mapped memory and traps.
__device__ __noinline__ void hang_report(unsigned int *rec, int code, int it, uint32_t parity)
{
    if (rec) {
        rec[1] = (unsigned int)code; rec[2] = blockIdx.x; rec[3] = (unsigned int)it; rec[4] = parity; rec[5] = threadIdx.x;
        __threadfence_system();
        rec[0] = 0xDEADu;
        __threadfence_system();
    }
    __trap();
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity, unsigned int *rec, int code, int it)
{
    unsigned int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        ++spins;
        if (spins == (1u << 16) && rec && (threadIdx.x & 31) == 0) {   // note who is waiting on what, per warp
            unsigned int *w = rec + 8 + (blockIdx.x * 10 + (threadIdx.x >> 5)) * 4;
            w[0] = (unsigned int)code; w[1] = (unsigned int)it; w[2] = parity; w[3] |= 0x80000000u;
            __threadfence_system();
        }
        if (spins > (1u << 22)) hang_report(rec, code, it, parity);
    }
}

__device__ __forceinline__ void st_stream_f32(float *p, float v)
{
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__global__ void __launch_bounds__(dwp::NTHREADS, 1) dw_persist_kernel(const DwpArgs *__restrict__ gp)
{
    using namespace dwp;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *a_slots = smem, *b_ring = smem + A_SLOTS * A_SLOT, *wd_ring = b_ring + B_STAGES * B_STAGE;
    __shared__ __align__(8) uint64_t a_full[A_SLOTS], a_empty[A_SLOTS], b_full[B_STAGES], b_empty[B_STAGES], t_full[2], t_empty[2];
    __shared__ __align__(8) uint64_t wd_full[WD_STAGES], wd_empty[WD_STAGES];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = gp->total_tiles;
    const int t0 = (int)((long long)T * blockIdx.x / gridDim.x), t1 = (int)((long long)T * (blockIdx.x + 1) / gridDim.x);

    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < A_SLOTS; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
            for (int s = 0; s < B_STAGES; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
            for (int s = 0; s < 2; s++) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 8); }
            for (int s = 0; s < WD_STAGES; s++) { mbar_init(&wd_full[s], 1); mbar_init(&wd_empty[s], 8); }
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
            for (int t = t0, it = 0; t < t1; t++, it++) {
                const TileRef tr = decode_tile(gp, t);
                const DwpLayer *L = tr.L;
                if (tr.key != key) {
                    key = tr.key;
                    const int sl = a_cnt % A_SLOTS, ph = (a_cnt / A_SLOTS) & 1;
                    a_cnt++;
                    mbar_wait_bounded(&a_empty[sl], ph ^ 1, gp->hang, 1, it);
                    mbar_expect_tx(&a_full[sl], A_SLOT);
                    uint8_t *dst = a_slots + sl * A_SLOT;
#pragma unroll
                    for (int kb = 0; kb < KB; kb++)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            tma_load_2d(dst + kb * 2 * A_PART + h * A_HALF, &L->a_hi, &a_full[sl], tr.nt * TN + 64 * h, kb * BK);
                            tma_load_2d(dst + kb * 2 * A_PART + A_PART + h * A_HALF, &L->a_lo, &a_full[sl], tr.nt * TN + 64 * h, kb * BK);
                        }
                }
                const int s = it % B_STAGES, ph = (it / B_STAGES) & 1;
                mbar_wait_bounded(&b_empty[s], ph ^ 1, gp->hang, 2, it);
                mbar_expect_tx(&b_full[s], B_STAGE);
                uint8_t *dst = b_ring + s * B_STAGE;
                const int r0 = L->b_rows_from_ctl ? bunch_row0 : 0;
#pragma unroll
                for (int kb = 0; kb < KB; kb++) {
                    tma_load_2d(dst + kb * 2 * B_PART, &L->b_hi, &b_full[s], tr.kt * TK, r0 + kb * BK);
                    tma_load_2d(dst + kb * 2 * B_PART + B_PART, &L->b_lo, &b_full[s], tr.kt * TK, r0 + kb * BK);
                }
                // fp32 weights / momentum of this tile: two half tiles of 32 k rows
#pragma unroll
                for (int hs = 0; hs < 2; hs++) {
                    const int seq = 2 * it + hs, ws = seq % WD_STAGES, wph = (seq / WD_STAGES) & 1;
                    mbar_wait_bounded(&wd_empty[ws], wph ^ 1, gp->hang, 3, it);
                    mbar_expect_tx(&wd_full[ws], WD_STAGE);
                    uint8_t *wdst = wd_ring + ws * WD_STAGE;
                    tma_load_2d(wdst, &L->w_map, &wd_full[ws], tr.nt * TN, tr.kt * TK + hs * WD_ROWS);
                    tma_load_2d(wdst + WD_PART, &L->d_map, &wd_full[ws], tr.nt * TN, tr.kt * TK + hs * WD_ROWS);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && t0 < t1) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = make_idesc_bf16(TN, TK, true, true);
            int key = -1, a_cnt = 0, sl = 0;
            TileRef tr = decode_tile(gp, t0);
            for (int t = t0, it = 0; t < t1; t++, it++) {
                if (tr.key != key) {
                    key = tr.key;
                    sl = a_cnt % A_SLOTS;
                    mbar_wait_bounded(&a_full[sl], (a_cnt / A_SLOTS) & 1, gp->hang, 4, it);
                    a_cnt++;
                }
                const int s = it % B_STAGES;
                mbar_wait_bounded(&b_full[s], (it / B_STAGES) & 1, gp->hang, 5, it);
                const int acc = it & 1;
                mbar_wait_bounded(&t_empty[acc], ((it >> 1) & 1) ^ 1, gp->hang, 6, it);
                tc_fence_after();
                const uint32_t a0 = smem_u32(a_slots + sl * A_SLOT), b0 = smem_u32(b_ring + s * B_STAGE);
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
                umma_commit(&b_empty[s]);
                umma_commit(&t_full[acc]);
                TileRef nx = tr;
                if (t + 1 < t1) nx = decode_tile(gp, t + 1);
                if (t + 1 >= t1 || nx.key != key) umma_commit(&a_empty[sl]);   // last tile that reads this dx^T slot
                tr = nx;
            }
        }
        __syncwarp();
    } else if (t0 < t1) {
        // ===== update warps (8): quadrant q = lanes [32q, 32q+32) of the accumulator, `half` = 32 of its 64 columns =====
        const int e = warp - 2, q = warp & 3, half = e >> 2;
        const float mom = gp->mom, lr = gp->lr, inv_mg = 1.0f / gp->Mg;
        TileRef tr = decode_tile(gp, t0);
        for (int t = t0, it = 0; t < t1; t++, it++) {
            const DwpLayer *L = tr.L;
            float *W = L->W, *D = L->D;
            bf16 *Hi = L->w_hi, *Lo = L->w_lo;
            const int Np = L->Np;
            const float wc = L->wc;
            const int n = tr.nt * TN + q * 32 + lane;
            const bool ok = n < Np;
            const size_t off0 = (size_t)(tr.kt * TK + half * 32) * Np + n;
            TileRef nx = tr;
            if (t + 1 < t1) nx = decode_tile(gp, t + 1);

            if (gp->dbg_progress && lane == 0) { gp->hang[8 + (blockIdx.x * 10 + warp) * 4 + 3] = 1000u + (unsigned int)it; }
            const int acc = it & 1;
            // Every update warp follows EVERY stage of the W/delta ring in order (full -> empty), although it only reads
            // the half tile of its own `half`: a warp that skipped the other half's stages could get two phases away from
            // a barrier (TMA loads land out of order) and a parity wait cannot tell phase f from phase f+2.
            const int seq = 2 * it + half, ws = seq % WD_STAGES;
            const int oseq = 2 * it + (half ^ 1), ows = oseq % WD_STAGES;
            const float *ws_w = reinterpret_cast<const float *>(wd_ring + ws * WD_STAGE) + q * 32 + lane;
            const float *ws_d = ws_w + WD_PART / 4;
            if (half == 0) {
                mbar_wait_bounded(&wd_full[ws], (seq / WD_STAGES) & 1, gp->hang, 7, it);
                mbar_wait_bounded(&wd_full[ows], (oseq / WD_STAGES) & 1, gp->hang, 9, it);
            } else {
                mbar_wait_bounded(&wd_full[ows], (oseq / WD_STAGES) & 1, gp->hang, 9, it);
                mbar_wait_bounded(&wd_full[ws], (seq / WD_STAGES) & 1, gp->hang, 7, it);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&wd_empty[ows]);   // not read by this warp
            mbar_wait_bounded(&t_full[acc], (it >> 1) & 1, gp->hang, 8, it);
            __syncwarp();
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + acc * TK + half * 32;
#pragma unroll
            for (int c = 0; c < 32; c += 16) {
                float g[16], w[16], d[16];
#pragma unroll
                for (int x = 0; x < 16; x++) { w[x] = ws_w[(c + x) * TN]; d[x] = ws_d[(c + x) * TN]; }
                tmem_ld16(taddr + c, g);
                if (c == 16) {   // accumulator and W/delta stage have been drained by this warp: hand them back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(&t_empty[acc]); mbar_arrive(&wd_empty[ws]); }
                }
                if (ok) {
#pragma unroll
                    for (int x = 0; x < 16; x++) {
                        // kernUpdatedelta + kernAccSum (DevFunc.cu:490-507, 427-443); g/Mg as g*(1/Mg) (<= 1 ulp)
                        const float ww = w[x];
                        const float dd = mom * d[x] - lr * (g[x] * inv_mg + wc * ww);
                        const float wn = dd + ww;
                        const size_t o = off0 + (size_t)(c + x) * Np;
                        st_stream_f32(W + o, wn);
                        st_stream_f32(D + o, dd);
                        bf16 h, l;
                        split_bf16(wn, h, l);
                        Hi[o] = h;
                        Lo[o] = l;
                    }
                }
            }
            // bias gradient + bias update of the 128 units of this n-tile (kernAccSumrow, DevFunc.cu:267-285; BP_GPU.cu:434-437)
            if (tr.kt == 0 && half == 0 && n < L->N) {
                const bf16 *xh = L->dx_hi + n, *xl = L->dx_lo + n;
                float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
                const int M = gp->M;
                int m = 0;
                for (; m + 4 <= M; m += 4) {
                    s0 += join_bf16(xh[(size_t)m * Np], xl[(size_t)m * Np]);
                    s1 += join_bf16(xh[(size_t)(m + 1) * Np], xl[(size_t)(m + 1) * Np]);
                    s2 += join_bf16(xh[(size_t)(m + 2) * Np], xl[(size_t)(m + 2) * Np]);
                    s3 += join_bf16(xh[(size_t)(m + 3) * Np], xl[(size_t)(m + 3) * Np]);
                }
                for (; m < M; m++) s0 += join_bf16(xh[(size_t)m * Np], xl[(size_t)m * Np]);
                const float sum = (s0 + s1) + (s2 + s3);
                const float db = mom * L->db[n] - lr * (sum / gp->Mg);   // no weight cost on biases (BP_GPU.cu:435)
                L->db[n] = db;
                L->b[n] = db + L->b[n];
            }
            tr = nx;
        }
        if (gp->dbg_progress && lane == 0) { gp->hang[8 + (blockIdx.x * 10 + warp) * 4 + 3] = 5000u; }
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

int launch_dw_persist(const DwpArgs *dev_args, int grid, cudaStream_t s)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(dwp::NTHREADS);
    cfg.dynamicSmemBytes = dwp::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    GGD_CUDA(cudaLaunchKernelEx(&cfg, dw_persist_kernel, dev_args));
    return GGD_OK;
}

int dw_persist_init()
{
    GGD_CUDA(cudaFuncSetAttribute(dw_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dwp::SMEM));
    return GGD_OK;
}

}  // namespace ggd
