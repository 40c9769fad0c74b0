// dw_update.cu -- fused weight-gradient GEMM + momentum-SGD update (see gemm_tc.cuh: DwUpdPlan).
//
// CTA = 320 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue/update.
// Per 128 (input units k) x 64 (output units n) tile of one weight matrix:
//   1. the operand k-blocks (64 frames each: y^T 128x64 and dx^T 64x64 as bf16 hi/lo, both MN-major) stream in by TMA
//      and one thread issues the bf16x3 tcgen05 MMAs into a 64-column TMEM accumulator; meanwhile every update warp
//      has already issued the global loads of its first W / delta rows (they do not depend on the GEMM);
//   2. the gradient tile goes TMEM -> registers -> padded shared memory (row-per-thread, conflict-free) and is
//      re-read row-wise: a warp owns a whole 256-byte row, so W / delta / bf16-shadow traffic is fully coalesced;
//   3. delta <- mom*delta - lr*(g/Mg + wc*W);  W <- W + delta;  shadows <- bf16 split(W).
// Shared memory is only the operand ring (48 KB per stage), so 2-4 CTAs share an SM and one tile's loads overlap
// another tile's stores: the kernel is meant to run at HBM speed (16 B/param + 4 B/param of shadows).
#include "gemm_tc.cuh"
#include "../../include/ggd_train.h"

namespace ggd {

namespace dwu {
constexpr int BK = 64, TILE_I = 128, BN = 64;
constexpr int A_TILE = TILE_I * BK * 2;        // 16 KB  (128 k x 64 frames, bf16)
constexpr int B_TILE = BN * BK * 2;            //  8 KB
constexpr int STAGE = 2 * A_TILE + 2 * B_TILE; // 48 KB
constexpr int G_PITCH = BN + 4;                // floats; 272-byte rows keep the row-per-thread float4 stores conflict-free
constexpr int NTHREADS = 320;
constexpr int ROWS_PER_WARP = TILE_I / 8;      // 16
constexpr int RB = ROWS_PER_WARP;              // ALL rows of a warp have their W / delta loads in flight during the GEMM
static_assert(TILE_I * G_PITCH * 4 <= STAGE, "gradient staging must fit in one drained operand stage");
template <int STAGES> constexpr int smem_bytes() { return STAGES * STAGE + 1024; }
}  // namespace dwu

__device__ __forceinline__ void dstamp(const DwUpdArgs &g, int slot)
{
    if (g.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g.trace[(size_t)(blockIdx.x + gridDim.x * blockIdx.y) * 16 + slot] = t;
    }
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float2 ld_stream_f2(const float *p)
{
    float2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

template <int STAGES>
__global__ void __launch_bounds__(dwu::NTHREADS, 2)
dw_update_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, const DwUpdArgs g)
{
    using namespace dwu;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *ring = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float *Gs = reinterpret_cast<float *>(ring);          // aliases operand stage 0 once the MMAs have retired
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j0 = blockIdx.x * BN, i0 = blockIdx.y * TILE_I;
    const int nkb = g.kblocks;
    if (threadIdx.x == 0) dstamp(g, 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo); tma_prefetch_desc(&tm_b_hi); tma_prefetch_desc(&tm_b_lo);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            mbar_init(&tmem_full_bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<BN>(&tmem_base_s);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    pdl_wait();      // prologue overlapped the previous kernel's tail; its outputs are visible from here on
    const int a_row_off = g.a_rows_from_ctl ? g.ctl->bunch_idx * g.rows_per_bunch : 0;
    if (threadIdx.x == 0) dstamp(g, 1);

    if (warp == 0) {
        if (lane == 0) {
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
                if (it == 0) dstamp(g, 3);
                if (it == nkb - 1) dstamp(g, 4);
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
        const int e = warp - 2;                 // 0..7: owns rows e*16 .. e*16+15 of the tile in the update phase
        const int q = warp & 3;                 // TMEM lane quadrant
        const int half = e >> 2;                // column half [32*half, 32*half+32) in the TMEM -> smem phase
        const int row = q * 32 + lane;
        const int rbase = e * ROWS_PER_WARP;
        const size_t colo = (size_t)j0 + 2 * lane;
        // W / delta of the first RB rows: independent of the GEMM, so they are requested before waiting for it
        float2 w[RB], d[RB];
#pragma unroll
        for (int x = 0; x < RB; x++) {
            const int k = i0 + rbase + x;
            const size_t o = (size_t)(k < g.Kp ? k : 0) * g.Np + colo;
            w[x] = ld_stream_f2(g.W + o);
            d[x] = ld_stream_f2(g.D + o);
        }
        if (threadIdx.x == 64) dstamp(g, 2);
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        if (threadIdx.x == 64) dstamp(g, 6);
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
        pdl_trigger();
        if (threadIdx.x == 64) dstamp(g, 7);
        // the update phase is issue-bound (16 update warps share 4 schedulers), so the arithmetic is kept minimal:
        // g/Mg becomes g*(1/Mg) (<= 1 ulp from the reference's division) and the bf16 splits use the paired converts
        const float mom = g.mom, lr = g.lr, inv_mg = 1.0f / g.Mg, wc = g.wc;
#pragma unroll
        for (int x = 0; x < RB; x++) {
            const int r = rbase + x, k = i0 + r;
            if (k < g.Kp) {
                const float2 gr = *reinterpret_cast<const float2 *>(Gs + (size_t)r * G_PITCH + 2 * lane);
                float2 dd = d[x], ww = w[x];
                dd.x = mom * dd.x - lr * (gr.x * inv_mg + wc * ww.x);
                dd.y = mom * dd.y - lr * (gr.y * inv_mg + wc * ww.y);
                ww.x = dd.x + ww.x;
                ww.y = dd.y + ww.y;
                const size_t o = (size_t)k * g.Np + colo;
                *reinterpret_cast<float2 *>(g.W + o) = ww;        // 32 lanes x 8 B = one 256-byte row segment
                *reinterpret_cast<float2 *>(g.D + o) = dd;
                const __nv_bfloat162 hi2 = __floats2bfloat162_rn(ww.x, ww.y);
                const float2 hf = __bfloat1622float2(hi2);
                const __nv_bfloat162 lo2 = __floats2bfloat162_rn(ww.x - hf.x, ww.y - hf.y);
                *reinterpret_cast<__nv_bfloat162 *>(g.w_hi + o) = hi2;
                *reinterpret_cast<__nv_bfloat162 *>(g.w_lo + o) = lo2;
            }
        }
    }
    if (threadIdx.x == 64) dstamp(g, 8);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<BN>(tmem);
    if (threadIdx.x == 0) dstamp(g, 9);
}

int launch_dw_update(const DwUpdPlan &p, cudaStream_t s)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.tiles_j, p.tiles_i, 1);
    cfg.blockDim = dim3(dwu::NTHREADS);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (p.stages == 1) {
        cfg.dynamicSmemBytes = dwu::smem_bytes<1>();
        GGD_CUDA(cudaLaunchKernelEx(&cfg, dw_update_kernel<1>, p.a_hi, p.a_lo, p.b_hi, p.b_lo, p.args));
    } else {
        cfg.dynamicSmemBytes = dwu::smem_bytes<2>();
        GGD_CUDA(cudaLaunchKernelEx(&cfg, dw_update_kernel<2>, p.a_hi, p.a_lo, p.b_hi, p.b_lo, p.args));
    }
    return GGD_OK;
}

int dw_update_init()
{
    GGD_CUDA(cudaFuncSetAttribute(dw_update_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dwu::smem_bytes<1>()));
    GGD_CUDA(cudaFuncSetAttribute(dw_update_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dwu::smem_bytes<2>()));
    return GGD_OK;
}

}  // namespace ggd
