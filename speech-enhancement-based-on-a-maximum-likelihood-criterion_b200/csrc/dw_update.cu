// dw_update.cu -- fused weight-gradient GEMM + momentum-SGD update (see gemm_tc.cuh: DwUpdPlan).
//
// CTA = 320 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue/update.
// Per 128 (input units k) x 64 (output units n) tile of one weight matrix:
//   1. TMA prefetches the fp32 W and delta tiles (2 x 32 KB) while
//   2. the operand k-blocks (64 frames each: y^T 128x64 and dx^T 64x64 as bf16 hi/lo, both MN-major) stream in and
//      one thread issues the bf16x3 tcgen05 MMAs into a 64-column TMEM accumulator;
//   3. the epilogue warps move the gradient tile TMEM -> registers -> padded shared memory (row-per-thread), then
//      re-read it row-wise (a warp owns a whole 256-byte row: conflict-free, and coalesced for the shadow stores),
//      apply the update in place on the shared-memory W / delta tiles and
//   4. TMA stores W and delta back; the bf16 hi/lo shadows leave through coalesced 128-byte row stores.
#include "gemm_tc.cuh"
#include "../../include/ggd_train.h"

namespace ggd {

namespace dwu {
constexpr int BK = 64, TILE_I = 128, BN = 64;
constexpr int A_TILE = TILE_I * BK * 2;        // 16 KB  (128 k x 64 frames, bf16)
constexpr int B_TILE = BN * BK * 2;            //  8 KB
constexpr int STAGE = 2 * A_TILE + 2 * B_TILE; // 48 KB
constexpr int STAGES = 2;
constexpr int WD_TILE = TILE_I * BN * 4;       // 32 KB
constexpr int G_PITCH = BN + 4;                // floats; 272-byte rows keep the row-per-thread float4 stores conflict-free
constexpr int SMEM = STAGES * STAGE + 2 * WD_TILE + 1024;
constexpr int NTHREADS = 320;
static_assert(TILE_I * G_PITCH * 4 <= STAGES * STAGE, "gradient staging must fit in the drained operand ring");
}  // namespace dwu

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *m, const void *smem_src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(dwu::NTHREADS, 1)
dw_update_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                 const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_d, const DwUpdArgs g)
{
    using namespace dwu;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *ring = smem;
    float *Wt = reinterpret_cast<float *>(smem + STAGES * STAGE);
    float *Dt = reinterpret_cast<float *>(smem + STAGES * STAGE + WD_TILE);
    float *Gs = reinterpret_cast<float *>(ring);          // aliases the operand ring once the MMAs have retired
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar, wd_bar;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j0 = blockIdx.x * BN, i0 = blockIdx.y * TILE_I;
    const int a_row_off = g.a_rows_from_ctl ? g.ctl->bunch_idx * g.rows_per_bunch : 0;
    const int nkb = g.kblocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo); tma_prefetch_desc(&tm_b_hi); tma_prefetch_desc(&tm_b_lo);
        tma_prefetch_desc(&tm_w); tma_prefetch_desc(&tm_d);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            mbar_init(&tmem_full_bar, 1);
            mbar_init(&wd_bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<BN>(&tmem_base_s);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            // the fp32 W / delta tiles first: they are the HBM traffic and have the whole GEMM to arrive
            mbar_expect_tx(&wd_bar, 2 * WD_TILE);
            tma_load_2d(Wt, &tm_w, &wd_bar, j0, i0);
            tma_load_2d(Dt, &tm_d, &wd_bar, j0, i0);
            for (int it = 0; it < nkb; it++) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], STAGE);
                uint8_t *st = ring + s * STAGE;
                const int r0 = it * BK + a_row_off;     // frame rows of this k-block
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    tma_load_2d(st + h * 8192, &tm_a_hi, &full_bar[s], i0 + 64 * h, r0);
                    tma_load_2d(st + A_TILE + h * 8192, &tm_a_lo, &full_bar[s], i0 + 64 * h, r0);
                }
                tma_load_2d(st + 2 * A_TILE, &tm_b_hi, &full_bar[s], j0, it * BK);
                tma_load_2d(st + 2 * A_TILE + B_TILE, &tm_b_lo, &full_bar[s], j0, it * BK);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(TILE_I, BN, true, true);
            for (int it = 0; it < nkb; it++) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(ring + s * STAGE), a_lo = a_hi + A_TILE;
                const uint32_t b_hi = a_hi + 2 * A_TILE, b_lo = b_hi + B_TILE;
#pragma unroll
                for (int k = 0; k < BK / 16; k++) {
                    const uint64_t dah = make_smem_desc(a_hi + k * 2048, 8192, 1024), dal = make_smem_desc(a_lo + k * 2048, 8192, 1024);
                    const uint64_t dbh = make_smem_desc(b_hi + k * 2048, 8192, 1024), dbl = make_smem_desc(b_lo + k * 2048, 8192, 1024);
                    umma_bf16(tmem, dal, dbh, idesc, (it | k) != 0);
                    umma_bf16(tmem, dah, dbl, idesc, 1);
                    umma_bf16(tmem, dah, dbh, idesc, 1);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&tmem_full_bar);
        }
        __syncwarp();
    } else {
        // ===== epilogue / update warps (8) =====
        const int e = warp - 2;                 // 0..7
        const int q = warp & 3;                 // TMEM lane quadrant
        const int half = e >> 2;                // column half [32*half, 32*half+32)
        const int row = q * 32 + lane;
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        // gradient tile -> padded shared memory (all MMAs have retired: the operand ring is free)
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + half * 32;
        float *grow = Gs + (size_t)row * G_PITCH + half * 32;
#pragma unroll
        for (int c = 0; c < 32; c += 16) {
            float v[16];
            tmem_ld16(taddr + c, v);
#pragma unroll
            for (int x = 0; x < 4; x++) *reinterpret_cast<float4 *>(grow + c + 4 * x) = make_float4(v[4 * x], v[4 * x + 1], v[4 * x + 2], v[4 * x + 3]);
        }
        epi_bar_sync();
        mbar_wait(&wd_bar, 0);
        // update: warp e owns rows e, e+8, ...; a lane owns two adjacent columns
        const float mom = g.mom, lr = g.lr, Mg = g.Mg, wc = g.wc;
        for (int r = e; r < TILE_I; r += 8) {
            const int k = i0 + r;
            float2 w = *reinterpret_cast<float2 *>(Wt + r * BN + 2 * lane);
            float2 d = *reinterpret_cast<float2 *>(Dt + r * BN + 2 * lane);
            const float2 gr = *reinterpret_cast<const float2 *>(Gs + (size_t)r * G_PITCH + 2 * lane);
            d.x = mom * d.x - lr * (gr.x / Mg + wc * w.x);
            d.y = mom * d.y - lr * (gr.y / Mg + wc * w.y);
            w.x = d.x + w.x;
            w.y = d.y + w.y;
            *reinterpret_cast<float2 *>(Wt + r * BN + 2 * lane) = w;
            *reinterpret_cast<float2 *>(Dt + r * BN + 2 * lane) = d;
            if (k < g.Kp) {
                bf16 h0, l0, h1, l1;
                split_bf16(w.x, h0, l0);
                split_bf16(w.y, h1, l1);
                const size_t o = (size_t)k * g.Np + j0 + 2 * lane;
                *reinterpret_cast<uint32_t *>(g.w_hi + o) = pack_bf16x2(h0, h1);   // 32 lanes x 4 B = one 128-byte line
                *reinterpret_cast<uint32_t *>(g.w_lo + o) = pack_bf16x2(l0, l1);
            }
        }
        fence_async_proxy();       // generic-proxy writes of Wt / Dt -> visible to the TMA store
        epi_bar_sync();
        if (e == 0 && lane == 0) {
            tma_store_2d(&tm_w, Wt, j0, i0);
            tma_store_2d(&tm_d, Dt, j0, i0);
            tma_store_commit_wait();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<BN>(tmem);
}

int launch_dw_update(const DwUpdPlan &p, cudaStream_t s)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.tiles_j, p.tiles_i, 1);
    cfg.blockDim = dim3(dwu::NTHREADS);
    cfg.dynamicSmemBytes = dwu::SMEM;
    cfg.stream = s;
    GGD_CUDA(cudaLaunchKernelEx(&cfg, dw_update_kernel, p.a_hi, p.a_lo, p.b_hi, p.b_lo, p.w, p.d, p.args));
    return GGD_OK;
}

int dw_update_init()
{
    GGD_CUDA(cudaFuncSetAttribute(dw_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dwu::SMEM));
    return GGD_OK;
}

}  // namespace ggd
