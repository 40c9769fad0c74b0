// dp_push.cu -- see dp_push.cuh: gradient-tile push (K1), owner-side reduce + update + shadow broadcast (K2).
#include "dp_push.cuh"
#include "pipe.cuh"
#include "../../include/ggd_train.h"

namespace ggd {

struct XTile {
    const DpxLayer *L;
    int nt, kt, key;
};
__device__ __forceinline__ XTile decode_xtile(const DpxArgs *gp, int t)
{
    int l = 0;
#pragma unroll 1
    while (l + 1 < gp->nlayers && t >= gp->layer[l + 1].tile_base) l++;
    const DpxLayer *L = &gp->layer[l];
    const int r = t - L->tile_base;
    XTile x;
    x.L = L; x.nt = r / L->k_tiles; x.kt = r - x.nt * L->k_tiles; x.key = (l << 16) | x.nt;
    return x;
}
__device__ __forceinline__ int owner_of(const DpxArgs *gp, int t)
{
    int o = 0;
#pragma unroll 1
    while (o + 1 < gp->world && t >= gp->own_begin[o + 1]) o++;
    return o;
}

__device__ __forceinline__ void xstamp(const DpxArgs *gp, int kernel, int slot)
{
    if (gp->trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        gp->trace[((size_t)kernel * gridDim.x + blockIdx.x) * 8 + slot] = t;
    }
}

// wait until flags[base + p] >= step for every peer p; bounded (a lost peer must not hang the GPU)
__device__ bool wait_peers(const DpxArgs *gp, int base, unsigned int step)
{
    const unsigned int *f = gp->flags[gp->rank] + base;
    const long long t0 = clock64();
    for (int p = 0; p < gp->world; p++) {
        if (p == gp->rank) continue;
        while ((int)(ld_acquire_sys_u32(f + p) - step) < 0) {
            if (clock64() - t0 > (1ll << 32)) { *gp->error_flag = 1u + p; __threadfence_system(); return false; }   // ~2 s
            __nanosleep(64);
        }
    }
    return true;
}

// =====================================================================================================================
// K1: gradient tiles of every tile -> owner's receive slot
// =====================================================================================================================
namespace k1 {
constexpr int TN = 128, TK = 64, BK = 64, KB = 2;
constexpr int A_HALF = 64 * BK * 2, A_PART = 2 * A_HALF, A_SLOT = KB * 2 * A_PART;   // 64 KB
constexpr int B_PART = TK * BK * 2, B_STAGE = KB * 2 * B_PART;                       // 32 KB
constexpr int ROWS = 16, QUARTERS = TK / ROWS;
constexpr int STAGE = ROWS * TN * 4;        // 8 KB
constexpr int STAGES = 8;
constexpr int B_STAGES = 3;                 // y^T ring: the MMA chain must not wait for an L2 round trip per tile
constexpr int SMEM = A_SLOT + B_STAGES * B_STAGE + STAGES * STAGE + 1024;
constexpr int NTHREADS = 384;
constexpr int TMEM_COLS = 2 * TK;
}  // namespace k1

__global__ void __launch_bounds__(k1::NTHREADS, 1) dw_push_kernel(const DpxArgs *__restrict__ gp)
{
    using namespace k1;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *a_slot = smem, *b_ring = smem + A_SLOT, *ring = b_ring + B_STAGES * B_STAGE;
    __shared__ __align__(8) uint64_t a_full, a_empty, b_full[B_STAGES], b_empty[B_STAGES], t_full[2], t_empty[2], st_done[STAGES], st_empty[STAGES];
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = gp->total_tiles;
    const int t0 = (int)((long long)T * blockIdx.x / gridDim.x), t1 = (int)((long long)T * (blockIdx.x + 1) / gridDim.x);
    unsigned int *const hang = gp->hang;

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(&a_full, 1); mbar_init(&a_empty, 1);
            for (int s = 0; s < B_STAGES; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
            for (int s = 0; s < 2; s++) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 8); }
            for (int s = 0; s < STAGES; s++) { mbar_init(&st_done[s], 8); mbar_init(&st_empty[s], 1); }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<TMEM_COLS>(&tmem_base_s);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) xstamp(gp, 0, 0);
    pdl_wait();
    if (threadIdx.x == 0) xstamp(gp, 0, 1);
    const int bunch_row0 = gp->ctl->bunch_idx * gp->rows_per_bunch;
    const unsigned int step = gp->counters[0] + 1u;

    if (warp == 0) {
        if (lane == 0 && t0 < t1) {
            // ===== TMA producer: operands only =====
            int key = -1, a_cnt = 0;
            for (int t = t0, it = 0; t < t1; t++, it++) {
                const XTile tr = decode_xtile(gp, t);
                const DpxLayer *L = tr.L;
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
                const int bs = it % B_STAGES;
                mbar_wait_bounded(&b_empty[bs], ((it / B_STAGES) & 1) ^ 1, hang, 2, it);
                mbar_expect_tx(&b_full[bs], B_STAGE);
                uint8_t *b_stage = b_ring + bs * B_STAGE;
                const int r0 = L->b_rows_from_ctl ? bunch_row0 : 0;
#pragma unroll
                for (int kb = 0; kb < KB; kb++) {
                    tma_load_2d(b_stage + kb * 2 * B_PART, &L->b_hi, &b_full[bs], tr.kt * TK, r0 + kb * BK);
                    tma_load_2d(b_stage + kb * 2 * B_PART + B_PART, &L->b_lo, &b_full[bs], tr.kt * TK, r0 + kb * BK);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && t0 < t1) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = make_idesc_bf16(TN, TK, true, true);
            int key = -1, a_cnt = 0;
            XTile tr = decode_xtile(gp, t0);
            for (int t = t0, it = 0; t < t1; t++, it++) {
                if (tr.key != key) {
                    key = tr.key;
                    mbar_wait_bounded(&a_full, a_cnt & 1, hang, 4, it);
                    a_cnt++;
                }
                const int bs = it % B_STAGES;
                mbar_wait_bounded(&b_full[bs], (it / B_STAGES) & 1, hang, 5, it);
                const int acc = it & 1;
                mbar_wait_bounded(&t_empty[acc], ((it >> 1) & 1) ^ 1, hang, 6, it);
                tc_fence_after();
                const uint32_t a0 = smem_u32(a_slot), b0 = smem_u32(b_ring + bs * B_STAGE);
                const uint32_t d = tmem + acc * TK;
#pragma unroll
                for (int kb = 0; kb < KB; kb++) {
                    const uint32_t a_hi = a0 + kb * 2 * A_PART, a_lo = a_hi + A_PART;
                    const uint32_t b_hi = b0 + kb * 2 * B_PART, b_lo = b_hi + B_PART;
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) {
                        const uint64_t dah = make_smem_desc(a_hi + k * 2048, 8192, 1024), dal = make_smem_desc(a_lo + k * 2048, 8192, 1024);
                        const uint64_t dbh = make_smem_desc(b_hi + k * 2048, 8192, 1024), dbl = make_smem_desc(b_lo + k * 2048, 8192, 1024);
                        umma_bf16(d, dal, dbh, idesc, (kb | k) != 0);
                        umma_bf16(d, dah, dbl, idesc, 1);
                        umma_bf16(d, dah, dbh, idesc, 1);
                    }
                }
                umma_commit(&b_empty[bs]);
                umma_commit(&t_full[acc]);
                XTile nx = tr;
                if (t + 1 < t1) nx = decode_xtile(gp, t + 1);
                if (t + 1 >= t1 || nx.key != key) umma_commit(&a_empty);
                tr = nx;
            }
        }
        __syncwarp();
    } else if (warp == 10) {
        if (lane == 0 && t0 < t1) {
            // ===== store warp: gradient quarter tiles -> owner's receive slot =====
            int prev_ws = -1;
            for (int t = t0, it = 0; t < t1; t++, it++) {
                const int o = owner_of(gp, t), tl = t - gp->own_begin[o];
#pragma unroll
                for (int qt = 0; qt < QUARTERS; qt++) {
                    const int seq = QUARTERS * it + qt, ws = seq % STAGES;
                    mbar_wait_bounded(&st_done[ws], (seq / STAGES) & 1, hang, 10, it);
                    tma_store_2d(&gp->push_map[o], ring + ws * STAGE, 0, tl * TK + qt * ROWS);
                    tma_store_commit();
                    if (prev_ws >= 0) {
                        tma_store_wait_read<1>();
                        mbar_arrive(&st_empty[prev_ws]);
                    }
                    prev_ws = ws;
                }
            }
            xstamp(gp, 0, 2);
            tma_store_wait_all<0>();      // every gradient tile of this CTA has been written at its owner
            xstamp(gp, 0, 3);
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        __syncwarp();
    } else if (warp == 11) {
        // ===== bias warp: partial bias gradients (column sums of dx over this rank's frames) into every rank's slot =====
        const int M = gp->M;
        int item = 0;
        for (int l = 0; l < gp->nlayers; l++) {
            const DpxLayer *L = &gp->layer[l];
            const int ntiles = (L->Np + TN - 1) / TN;
            for (int nt = 0; nt < ntiles; nt++, item++) {
                if (item % gridDim.x != blockIdx.x) continue;
                const int Np = L->Np, n = nt * TN + 4 * lane;
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
                const size_t o = (size_t)gp->rank * gp->nbias + L->bias_off + n;
                for (int p = 0; p < gp->world; p++)
                    *reinterpret_cast<float4 *>(gp->bias_slot[p] + o) = make_float4(s[0], s[1], s[2], s[3]);
            }
        }
    } else if (t0 < t1) {
        // ===== copy warps (8): accumulator -> shared-memory stage, [k row][n] like the weights =====
        const int e = warp - 2, q = warp & 3, h = e >> 2;
        for (int t = t0, it = 0; t < t1; t++, it++) {
            const int acc = it & 1;
            mbar_wait_bounded(&t_full[acc], (it >> 1) & 1, hang, 8, it);
            __syncwarp();
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + acc * TK + h * 8;
#pragma unroll
            for (int qt = 0; qt < QUARTERS; qt++) {
                const int seq = QUARTERS * it + qt, ws = seq % STAGES;
                float *sg = reinterpret_cast<float *>(ring + ws * STAGE) + (h * 8) * TN + q * 32 + lane;
                float g[8];
                tmem_ld8(taddr + qt * ROWS, g);
                if (qt == QUARTERS - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[acc]);
                }
                mbar_wait_bounded(&st_empty[ws], ((seq / STAGES) & 1) ^ 1, hang, 7, it);
#pragma unroll
                for (int x = 0; x < 8; x++) sg[x * TN] = g[x];
                fence_async_proxy();
                __syncwarp();
                if (lane == 0) mbar_arrive(&st_done[ws]);
            }
        }
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem);
    // flag A: once EVERY CTA's stores (tiles and bias slots) are performed system-wide, tell all peers
    if (threadIdx.x == 0) {
        xstamp(gp, 0, 4);
        __threadfence_system();
        const unsigned int prev = atomicAdd(gp->counters + 1, 1u);
        s_last = (prev == gridDim.x - 1);
        if (s_last) {
            gp->counters[1] = 0;
            __threadfence_system();
            for (int p = 0; p < gp->world; p++)
                if (p != gp->rank) st_release_sys_u32(gp->flags[p] + DPX_FLAG_A + gp->rank, step);
        }
    }
}

// =====================================================================================================================
// K2: owned tiles: sum the partial tiles in rank order, momentum-SGD update, broadcast the bf16 shadows
// =====================================================================================================================
namespace k2 {
constexpr int TN = 128, ROWS = 8, EIGHTHS = 64 / ROWS;
constexpr int F32 = ROWS * TN * 4;    // 4 KB
constexpr int B16 = ROWS * TN * 2;    // 2 KB
constexpr int MAX_STAGES = 10;
constexpr int NTHREADS = 352;         // producer, 8 update warps, store warp, bias warp
}  // namespace k2

int dp_push_k2_smem(int world, int *stages, int *stage_bytes)
{
    const int sb = (2 + world) * k2::F32 + 2 * k2::B16;
    int st = (200 * 1024) / sb;
    if (st > k2::MAX_STAGES) st = k2::MAX_STAGES;
    *stages = st; *stage_bytes = sb;
    return st * sb + 1024;
}

__global__ void __launch_bounds__(k2::NTHREADS, 1) reduce_update_kernel(const DpxArgs *__restrict__ gp)
{
    using namespace k2;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *ring = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full[MAX_STAGES], done[MAX_STAGES], empty[MAX_STAGES];
    __shared__ int s_ok;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int world = gp->world, rank = gp->rank;
    const int STG = gp->k2_stages, SB = gp->k2_stage_bytes;
    const int tb = gp->own_begin[rank], own = gp->own_begin[rank + 1] - tb;
    const int items = own * EIGHTHS;
    const int i0 = (int)((long long)items * blockIdx.x / gridDim.x), i1 = (int)((long long)items * (blockIdx.x + 1) / gridDim.x);
    unsigned int *const hang = gp->hang;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STG; s++) { mbar_init(&full[s], 1); mbar_init(&done[s], 8); mbar_init(&empty[s], 1); }
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) xstamp(gp, 1, 0);
    pdl_wait();
    if (threadIdx.x == 0) xstamp(gp, 1, 1);
    const unsigned int step = gp->counters[0] + 1u;
    // every peer's gradient tiles and bias partials of this step have landed in my receive slots
    if (threadIdx.x == 0) { s_ok = wait_peers(gp, DPX_FLAG_A, step) ? 1 : 0; xstamp(gp, 1, 2); }
    __syncthreads();
    const bool ok = s_ok != 0;

    if (warp == 0) {
        if (lane == 0 && ok) {
            // ===== TMA producer =====
            const uint64_t pol_stream = l2_policy_evict_first();
            for (int i = i0, it = 0; i < i1; i++, it++) {
                const int tl = i / EIGHTHS, e8 = i - tl * EIGHTHS;
                const XTile tr = decode_xtile(gp, tb + tl);
                const DpxLayer *L = tr.L;
                const int s = it % STG;
                mbar_wait_bounded(&empty[s], ((it / STG) & 1) ^ 1, hang, 3, it);
                mbar_expect_tx(&full[s], (2 + world) * F32);
                uint8_t *st = ring + s * SB;
                const int c0 = tr.nt * TN, c1 = tr.kt * 64 + e8 * ROWS;
                tma_load_2d_hint(st, &L->w_map, &full[s], c0, c1, pol_stream);
                tma_load_2d_hint(st + F32, &L->d_map, &full[s], c0, c1, pol_stream);
                for (int p = 0; p < world; p++)
                    tma_load_2d_hint(st + (2 + p) * F32, &gp->part_map[p], &full[s], 0, tl * 64 + e8 * ROWS, pol_stream);
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        if (lane == 0 && ok) {
            // ===== store warp =====
            const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
            int prev_s = -1;
            for (int i = i0, it = 0; i < i1; i++, it++) {
                const int tl = i / EIGHTHS, e8 = i - tl * EIGHTHS;
                const XTile tr = decode_xtile(gp, tb + tl);
                const DpxLayer *L = tr.L;
                const int s = it % STG;
                mbar_wait_bounded(&done[s], (it / STG) & 1, hang, 10, it);
                const uint8_t *st = ring + s * SB;
                const int c0 = tr.nt * TN, c1 = tr.kt * 64 + e8 * ROWS;
                tma_store_2d_hint(&L->w_map, st, c0, c1, gp->w_f32 ? pol_keep : pol_stream);
                tma_store_2d_hint(&L->d_map, st + F32, c0, c1, pol_stream);
                if (gp->w_f32) {
                    for (int p = 0; p < world; p++)
                        if (p != rank) tma_store_2d_hint(&L->wp_map[p], st, c0, c1, pol_keep);   // the weights themselves, to every peer
                } else {
                    const uint8_t *sh = st + (2 + world) * F32;
                    for (int p = 0; p < world; p++) {
                        tma_store_2d_hint(&L->hi_map[p], sh, c0, c1, pol_keep);
                        tma_store_2d_hint(&L->lo_map[p], sh + B16, c0, c1, pol_keep);
                    }
                }
                tma_store_commit();
                if (prev_s >= 0) {
                    tma_store_wait_read<1>();
                    mbar_arrive(&empty[prev_s]);
                }
                prev_s = s;
            }
            xstamp(gp, 1, 3);
            tma_store_wait_all<0>();      // my shadows are written in every rank's memory
            xstamp(gp, 1, 4);
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        __syncwarp();
    } else if (warp == 10) {
        // ===== bias warp: g = sum over ranks (rank order) of the bias partials; identical update on every rank =====
        if (ok) {
            const float mom = gp->mom, lr = gp->lr;
            const float *slot = gp->bias_slot[rank];
            for (int l = 0; l < gp->nlayers; l++) {
                const DpxLayer *L = &gp->layer[l];
                for (int n = blockIdx.x * 32 + lane; n < L->N; n += gridDim.x * 32) {
                    float g = 0.0f;
                    for (int p = 0; p < world; p++) g += ld_relaxed_sys_f32(slot + (size_t)p * gp->nbias + L->bias_off + n);
                    const float db = mom * L->db[n] - lr * (g / gp->Mg);   // no weight cost on biases (BP_GPU.cu:435)
                    L->db[n] = db;
                    L->b[n] = db + L->b[n];
                }
            }
        }
    } else if (ok) {
        // ===== update warps (8): warp = one k row of the stage, lane = 4 consecutive units =====
        const int row = warp - 1;
        const float mom = gp->mom, lr = gp->lr, inv_mg = 1.0f / gp->Mg;
        for (int i = i0, it = 0; i < i1; i++, it++) {
            const int tl = i / EIGHTHS;
            const XTile tr = decode_xtile(gp, tb + tl);
            const float wc = tr.L->wc;
            const int s = it % STG;
            uint8_t *st = ring + s * SB;
            float4 *sw = reinterpret_cast<float4 *>(st + (size_t)row * TN * 4) + lane;
            float4 *sd = reinterpret_cast<float4 *>(st + F32 + (size_t)row * TN * 4) + lane;
            mbar_wait_bounded(&full[s], (it / STG) & 1, hang, 7, it);
            float4 g = *(reinterpret_cast<const float4 *>(st + 2 * F32 + (size_t)row * TN * 4) + lane);
            for (int p = 1; p < world; p++) {
                const float4 x = *(reinterpret_cast<const float4 *>(st + (2 + p) * F32 + (size_t)row * TN * 4) + lane);
                g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
            }
            float4 w = *sw, d = *sd;
            // kernUpdatedelta + kernAccSum (DevFunc.cu:490-507, 427-443) on the gradient of the GLOBAL minibatch
            d.x = mom * d.x - lr * (g.x * inv_mg + wc * w.x);
            d.y = mom * d.y - lr * (g.y * inv_mg + wc * w.y);
            d.z = mom * d.z - lr * (g.z * inv_mg + wc * w.z);
            d.w = mom * d.w - lr * (g.w * inv_mg + wc * w.w);
            w.x += d.x; w.y += d.y; w.z += d.z; w.w += d.w;
            *sw = w;
            *sd = d;
            if (!gp->w_f32) {
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(w.x, w.y), h23 = __floats2bfloat162_rn(w.z, w.w);
            const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(w.x - f01.x, w.y - f01.y), l23 = __floats2bfloat162_rn(w.z - f23.x, w.w - f23.y);
            uint2 hv, lv;
            hv.x = *reinterpret_cast<const uint32_t *>(&h01); hv.y = *reinterpret_cast<const uint32_t *>(&h23);
            lv.x = *reinterpret_cast<const uint32_t *>(&l01); lv.y = *reinterpret_cast<const uint32_t *>(&l23);
            uint8_t *sh = st + (2 + world) * F32;
            *(reinterpret_cast<uint2 *>(sh + (size_t)row * TN * 2) + lane) = hv;
            *(reinterpret_cast<uint2 *>(sh + B16 + (size_t)row * TN * 2) + lane) = lv;
            }
            fence_async_proxy();
            __syncwarp();
            if (lane == 0) mbar_arrive(&done[s]);
        }
    }
    pdl_trigger();
    __syncthreads();
    // flag B: all my shadow slices are written everywhere; nobody leaves the step before every slice has landed
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int prev = atomicAdd(gp->counters + 2, 1u);
        if (prev == gridDim.x - 1) {
            gp->counters[2] = 0;
            __threadfence_system();
            for (int p = 0; p < world; p++)
                if (p != rank) st_release_sys_u32(gp->flags[p] + DPX_FLAG_B + rank, step);
            xstamp(gp, 1, 5);
            wait_peers(gp, DPX_FLAG_B, step);
            xstamp(gp, 1, 6);
            gp->counters[0] = step;
            gp->ctl->bunch_idx += 1;
            __threadfence();
        }
    }
}

// fp32 master weights of foreign tiles <- their owners (weight export only)
__global__ void gather_master_kernel(const DpxArgs *__restrict__ gp, float *const *peerP, const long long *w_off)
{
    for (int t = blockIdx.x; t < gp->total_tiles; t += gridDim.x) {
        const int o = owner_of(gp, t);
        if (o == gp->rank) continue;
        const XTile tr = decode_xtile(gp, t);
        const DpxLayer *L = tr.L;
        const int l = (int)(L - gp->layer);
        const float *src = peerP[o] + w_off[l];
        float *dst = peerP[gp->rank] + w_off[l];
        for (int e = threadIdx.x; e < 64 * 128; e += blockDim.x) {
            const int k = tr.kt * 64 + e / 128, n = tr.nt * 128 + e % 128;
            if (k < L->Kp && n < L->Np) dst[(size_t)k * L->Np + n] = src[(size_t)k * L->Np + n];
        }
    }
}

int launch_dw_push(const DpxArgs *dev_args, int grid, cudaStream_t s)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(k1::NTHREADS);
    cfg.dynamicSmemBytes = k1::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    GGD_CUDA(cudaLaunchKernelEx(&cfg, dw_push_kernel, dev_args));
    return GGD_OK;
}

int launch_reduce_update(const DpxArgs *dev_args, int grid, int smem_bytes, cudaStream_t s)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(k2::NTHREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    GGD_CUDA(cudaLaunchKernelEx(&cfg, reduce_update_kernel, dev_args));
    return GGD_OK;
}

void launch_gather_master(const DpxArgs *dev_args, float *const *peerP, const long long *w_off, int grid, cudaStream_t s)
{
    gather_master_kernel<<<grid, 256, 0, s>>>(dev_args, peerP, w_off);
}

int dp_push_init()
{
    GGD_CUDA(cudaFuncSetAttribute(dw_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k1::SMEM));
    GGD_CUDA(cudaFuncSetAttribute(reduce_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
    return GGD_OK;
}

}  // namespace ggd
