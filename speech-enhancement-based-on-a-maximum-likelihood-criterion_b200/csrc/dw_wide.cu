// dw_wide.cu -- persistent weight-gradient + momentum-SGD update kernel for minibatches of ANY number of frames
// (BP_GPU.cu:432-437 for all layers: SgemmNT, updatedelta, DevAccSum), the generalisation of dw_persist.cu:
//   * one GPU with a bunch above 128 frames (BASELINE config 4 at N = 1: 1024 frames), and
//   * frame-sharded data parallelism by FACTOR exchange (dp_factor.cuh): every rank holds the dE/dx and activation
//     factors of the WHOLE global minibatch (its own slice written by its GEMM epilogues, the other slices pushed by the
//     peers over NVLink) and runs this kernel REPLICATED -- same operands, same order, bit-identical weights on every
//     rank, and no gradient or weight ever crosses NVLink (7.4 MB of factors per rank and step instead of 2 x 51 MB).
//
//     g[n][k] = sum_m dx[m][n] * y[m][k]      UMMA: M = 128 (n, TMEM lanes), N = 64..256 (k, TMEM columns), K = frames
//
// Work unit: a SLAB = 128 output units n x 64 input units k of one weight matrix.  Every layer's slab list (n-tile, k-slab;
// k fastest) is cut into gridDim.x contiguous ranges (balanced to one slab); inside its range a CTA forms
// SEGMENTS of up to 4 consecutive slabs of one n-tile (UMMA N = 64 w <= 256), so that the dx^T operand is re-read once
// per 256 input units instead of once per 64: L2 -> SM operand traffic per weight drops 2.4x against 128 x 64 tiles.
// The frames are streamed in blocks of 32 through a TMA ring (48 KB per stage: dx^T 128 n x 32 frames hi+lo, y^T
// 64 w k x 32 frames hi+lo); two 256-column fp32 accumulators alternate in TMEM (all 512 columns), so the MMAs of
// segment t+1 run under the update of segment t.
//
// Warp roles (416 threads, one CTA per SM):
//   warp 0      operand producer (TMA).  In data-parallel mode it first waits (bounded) for the peers' flags of the layer.
//   warp 1      TMEM allocator + single-thread tcgen05 MMA issuer (bf16x3: lo*hi + hi*lo + hi*hi, small terms first)
//   warps 2..9  update warps: gradient quarter (16 k rows x 128 n) from TMEM, W / delta from the ring stage,
//               delta <- mom*delta - lr*(g/Mg + wc*W), W <- W + delta, written back in place
//   warp 10     store warp: TMA stores of W and delta, releases the stage once the store engine has read it
//   warp 11     weight producer: fp32 W and delta quarter tiles by TMA, running ahead of the update by the ring depth
//   warp 12     bias warp: column sums of dE/dx over the whole minibatch + bias update for this CTA's share of the columns
// All weight traffic is TMA.  HBM bytes: 16 B/param (read W, delta; write W, delta).
#include "dp_factor.cuh"
#include "pipe.cuh"
#include "../../include/ggd_train.h"
#include <stdlib.h>

namespace ggd {

namespace dww {
constexpr int TN = 128, SLAB = 64, MAXW = 4, FBK = 32;
constexpr int BOX = 64 * FBK * 2;                  // 4 KB: 64 units x 32 frames bf16 (one TMA box, 128-byte swizzle)
constexpr int A_PART = 2 * BOX;                    // 8 KB: 128 n x 32 frames, hi or lo
constexpr int B_PART = MAXW * BOX;                 // 16 KB: up to 256 k x 32 frames, hi or lo
constexpr int OP_STAGE = 2 * A_PART + 2 * B_PART;  // 48 KB
constexpr int WD_ROWS = 16;
constexpr int WD_F32 = WD_ROWS * TN * 4;           // 8 KB: 16 k x 128 n fp32
constexpr int WD_STAGE = 2 * WD_F32;               // W + delta
constexpr int MAX_OPS = 4, MAX_WDS = 8;
constexpr int NTHREADS = 416;                     // 13 warps: see the role list
constexpr int TMEM_COLS = 512;
constexpr int ACC_COLS = 256;
}  // namespace dww

struct WSeg {
    const DwwLayer *L;
    int li, nt, ks, w;
};

// Slab GROUPS: the layers (in list order, top layer first) are joined into groups; every group's slab list is cut into
// gridDim.x contiguous ranges and a CTA walks its range of group 0, then of group 1, ...  One GPU: one group (the fewest, widest
// segments).  Data parallel: the bottom layer is a group of its own -- its dE/dx factors arrive last, ~10-20 us after the
// backward chain, so every CTA first works through its share of the upper layers and nobody idles waiting for the peers.
struct SegIter {
    const DwwArgs *gp;
    int grp, s, s1;
    __device__ __forceinline__ explicit SegIter(const DwwArgs *g) : gp(g), grp(-1), s(0), s1(0) {}
    __device__ __forceinline__ bool next(WSeg &g)
    {
        while (s >= s1) {
            if (++grp >= gp->ngroups) return false;
            const int b = gp->group_base[grp], n = gp->group_base[grp + 1] - b;
            s = b + (int)((long long)n * blockIdx.x / gridDim.x);
            s1 = b + (int)((long long)n * (blockIdx.x + 1) / gridDim.x);
        }
        int li = 0;
#pragma unroll 1
        while (li + 1 < gp->nlayers && s >= gp->layer[li + 1].slab_base) li++;
        const DwwLayer *L = &gp->layer[li];
        const int r = s - L->slab_base;
        g.L = L; g.li = li; g.nt = r / L->k_slabs; g.ks = r - g.nt * L->k_slabs;
        int w = L->k_slabs - g.ks;
        if (w > dww::MAXW) w = dww::MAXW;
        if (w > s1 - s) w = s1 - s;
        g.w = w;
        s += w;
        return true;
    }
};

// bounded wait for the factor flags `ev` of every peer (value >= step); a lost peer must not hang the GPU
__device__ __noinline__ void wide_wait_flags(const DwwArgs *gp, int ev, unsigned int step)
{
    if (ev < 0) return;
    const long long t0 = clock64();
    for (int p = 0; p < gp->world; p++) {
        if (p == gp->rank) continue;
        const unsigned int *f = gp->flags + p * FX_STRIDE + ev;
        while ((int)(ld_acquire_sys_u32(f) - step) < 0) {
            if (clock64() - t0 > (1ll << 33)) {   // ~4 s
                *gp->error_flag = 1u + p;
                __threadfence_system();
                hang_report(gp->hang, 100 + ev, (int)step, (uint32_t)p);
            }
            __nanosleep(64);
        }
    }
}

__global__ void __launch_bounds__(dww::NTHREADS, 1) dw_wide_kernel(const DwwArgs *__restrict__ gp)
{
    using namespace dww;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int OPS = gp->op_stages, WDS = gp->wd_stages;
    uint8_t *op_ring = smem, *wd_ring = smem + OPS * OP_STAGE;
    __shared__ __align__(8) uint64_t op_full[MAX_OPS], op_empty[MAX_OPS], t_full[2], t_empty[2];
    __shared__ __align__(8) uint64_t wd_full[MAX_WDS], wd_done[MAX_WDS], wd_empty[MAX_WDS];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int FB = gp->fblocks;
    unsigned int *const hang = gp->hang;

    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < MAX_OPS; s++) { mbar_init(&op_full[s], 1); mbar_init(&op_empty[s], 1); }
            for (int s = 0; s < 2; s++) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 8); }
            for (int s = 0; s < MAX_WDS; s++) { mbar_init(&wd_full[s], 1); mbar_init(&wd_done[s], 8); mbar_init(&wd_empty[s], 1); }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<TMEM_COLS>(&tmem_base_s);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (gp->trace && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); gp->trace[FX_TRACE_WIDE] = t; }
    // The weight producer does NOT wait for the previous kernel: W and delta are only read by the chain, and nothing is stored
    // before the MMAs (which need the factors, i.e. griddepcontrol.wait) have finished -- the HBM stream is primed while the
    // last GEMM of the backward chain drains.
    if (warp != 11) pdl_wait();   // the factors of this step are complete and visible from here on
    if (gp->trace && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); gp->trace[FX_TRACE_WIDE + 1] = t; }
    const int bunch_row0 = (warp == 11) ? 0 : gp->ctl->bunch_idx * gp->rows_per_bunch;

    if (warp == 0) {
        if (lane == 0) {
            // ===== operand producer =====
            const unsigned int step = gp->world > 1 ? *gp->step_counter + 1u : 0u;
            int st = 0, ph = 0, ready = -1, it = 0;
            SegIter si(gp);
            WSeg g;
            while (si.next(g)) {
                const DwwLayer *L = g.L;
                if (gp->world > 1 && g.li != ready) {
                    wide_wait_flags(gp, L->ev_dx, step);
                    wide_wait_flags(gp, L->ev_y, step);
                    asm volatile("fence.proxy.async;" ::: "memory");   // peer (generic-proxy) writes -> my TMA reads
                    ready = g.li;
                    if (gp->trace) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMax(gp->trace + FX_TRACE_WIDE + 2 + g.li, t); }
                }
                const int r0 = L->b_rows_from_ctl ? bunch_row0 : 0;
                const uint32_t bytes = 2 * A_PART + 2 * g.w * BOX;
                for (int fb = 0; fb < FB; fb++, it++) {
                    mbar_wait_bounded(&op_empty[st], ph ^ 1, hang, 1, it);
                    mbar_expect_tx(&op_full[st], bytes);
                    uint8_t *dst = op_ring + st * OP_STAGE;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        tma_load_2d(dst + h * BOX, &L->a_hi, &op_full[st], g.nt * TN + 64 * h, fb * FBK);
                        tma_load_2d(dst + A_PART + h * BOX, &L->a_lo, &op_full[st], g.nt * TN + 64 * h, fb * FBK);
                    }
                    for (int j = 0; j < g.w; j++) {
                        tma_load_2d(dst + 2 * A_PART + j * BOX, &L->b_hi, &op_full[st], (g.ks + j) * SLAB, r0 + fb * FBK);
                        tma_load_2d(dst + 2 * A_PART + B_PART + j * BOX, &L->b_lo, &op_full[st], (g.ks + j) * SLAB, r0 + fb * FBK);
                    }
                    if (++st == OPS) { st = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            int st = 0, ph = 0, it = 0, seg_it = 0;
            SegIter si(gp);
            WSeg g;
            for (; si.next(g); seg_it++) {
                const int acc = seg_it & 1;
                mbar_wait_bounded(&t_empty[acc], ((seg_it >> 1) & 1) ^ 1, hang, 6, seg_it);
                tc_fence_after();
                const uint32_t idesc = make_idesc_bf16(TN, SLAB * g.w, true, true);
                const uint32_t d = tmem + acc * ACC_COLS;
                for (int fb = 0; fb < FB; fb++, it++) {
                    mbar_wait_bounded(&op_full[st], ph, hang, 5, it);
                    tc_fence_after();
                    const uint32_t base = smem_u32(op_ring + st * OP_STAGE);
#pragma unroll
                    for (int k = 0; k < FBK / 16; k++) {
                        const uint64_t dah = make_smem_desc(base + k * 2048, BOX, 1024);
                        const uint64_t dal = make_smem_desc(base + A_PART + k * 2048, BOX, 1024);
                        const uint64_t dbh = make_smem_desc(base + 2 * A_PART + k * 2048, BOX, 1024);
                        const uint64_t dbl = make_smem_desc(base + 2 * A_PART + B_PART + k * 2048, BOX, 1024);
                        umma_bf16(d, dal, dbh, idesc, (fb | k) != 0);   // small terms first
                        umma_bf16(d, dah, dbl, idesc, 1);
                        umma_bf16(d, dah, dbh, idesc, 1);
                    }
                    umma_commit(&op_empty[st]);
                    if (++st == OPS) { st = 0; ph ^= 1; }
                }
                umma_commit(&t_full[acc]);
            }
        }
        __syncwarp();
    } else if (warp == 10) {
        if (lane == 0) {
            // ===== store warp =====
            const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
            int ws = 0, wph = 0, prev_ws = -1, it = 0;
            SegIter si(gp);
            WSeg g;
            while (si.next(g)) {
                const DwwLayer *L = g.L;
                const int nq = 4 * g.w;
                for (int qt = 0; qt < nq; qt++, it++) {
                    mbar_wait_bounded(&wd_done[ws], wph, hang, 10, it);
                    const uint8_t *src = wd_ring + ws * WD_STAGE;
                    const int c0 = g.nt * TN, c1 = g.ks * SLAB + qt * WD_ROWS;
                    if (gp->l2_hints) {
                        // the fp32 weights are what the next step's GEMMs read: keep them in L2; the momentum is streamed
                        tma_store_2d_hint(&L->w_map, src, c0, c1, pol_keep);
                        tma_store_2d_hint(&L->d_map, src + WD_F32, c0, c1, pol_stream);
                    } else {
                        tma_store_2d(&L->w_map, src, c0, c1);
                        tma_store_2d(&L->d_map, src + WD_F32, c0, c1);
                    }
                    tma_store_commit();
                    if (prev_ws >= 0) {
                        tma_store_wait_read<1>();
                        mbar_arrive(&wd_empty[prev_ws]);
                    }
                    prev_ws = ws;
                    if (++ws == WDS) { ws = 0; wph ^= 1; }
                }
            }
            tma_store_wait_all<0>();   // all writes performed before the CTA (and with it the grid) completes
        }
        __syncwarp();
    } else if (warp == 12) {
        // ===== bias warp (kernAccSumrow + updatedelta + DevAccSum, DevFunc.cu:267-285; BP_GPU.cu:434-437): column sums of dE/dx
        // over the WHOLE minibatch in a fixed order (identical on every rank) and the bias update, 32 columns per item, items dealt
        // round-robin to the CTAs.  It runs BESIDE the slab pipeline (latency-bound L2 reads), not as a kernel of its own: a second
        // kernel spinning on the peers' flags can fill the SMs and starve the very push kernels the peers wait for.
        const unsigned int step = gp->world > 1 ? *gp->step_counter + 1u : 0u;
        const int cpair = lane & 15, half = lane >> 4;      // 16 column pairs x 2 row halves (even / odd frames)
        int item = 0;
        for (int li = 0; li < gp->nlayers; li++) {
            const DwwLayer *L = &gp->layer[li];
            const int nblk = (L->N + 31) / 32;
            bool waited = false;
            for (int bk = 0; bk < nblk; bk++, item++) {
                if (item % (int)gridDim.x != (int)blockIdx.x) continue;
                if (!waited && gp->world > 1) {
                    if (lane == 0) wide_wait_flags(gp, L->ev_dx, step);
                    __syncwarp();
                    waited = true;
                }
                const int n = bk * 32 + 2 * cpair;
                float s0 = 0.0f, s1 = 0.0f;
                if (n < L->Np) {
                    const unsigned int *ph = reinterpret_cast<const unsigned int *>(L->dx_hi + n), *pl = reinterpret_cast<const unsigned int *>(L->dx_lo + n);
                    const size_t pitch = (size_t)L->Np / 2;            // row pitch in 32-bit words (Np is a multiple of 64)
                    int m = half;
                    for (; m + 30 < gp->rows; m += 32) {               // 16 rows of this half per round, loads first
                        unsigned int vh[16], vl[16];
#pragma unroll
                        for (int u = 0; u < 16; u++) { vh[u] = __ldcg(ph + (size_t)(m + 2 * u) * pitch); vl[u] = __ldcg(pl + (size_t)(m + 2 * u) * pitch); }
#pragma unroll
                        for (int u = 0; u < 16; u++) {
                            s0 += __uint_as_float(vh[u] << 16) + __uint_as_float(vl[u] << 16);
                            s1 += __uint_as_float(vh[u] & 0xFFFF0000u) + __uint_as_float(vl[u] & 0xFFFF0000u);
                        }
                    }
                    for (; m < gp->rows; m += 2) {
                        const unsigned int vh = __ldcg(ph + (size_t)m * pitch), vl = __ldcg(pl + (size_t)m * pitch);
                        s0 += __uint_as_float(vh << 16) + __uint_as_float(vl << 16);
                        s1 += __uint_as_float(vh & 0xFFFF0000u) + __uint_as_float(vl & 0xFFFF0000u);
                    }
                }
                // even-frame half + odd-frame half, fixed order
                const float o0 = __shfl_xor_sync(0xffffffffu, s0, 16), o1 = __shfl_xor_sync(0xffffffffu, s1, 16);
                if (half == 0) {
                    const float g0 = s0 + o0, g1 = s1 + o1;
                    if (n < L->N) { const float db = gp->mom * L->db[n] - gp->lr * (g0 / gp->Mg); L->db[n] = db; L->b[n] = db + L->b[n]; }      // no weight cost on biases (BP_GPU.cu:435)
                    if (n + 1 < L->N) { const float db = gp->mom * L->db[n + 1] - gp->lr * (g1 / gp->Mg); L->db[n + 1] = db; L->b[n + 1] = db + L->b[n + 1]; }
                }
            }
        }
    } else if (warp == 11) {
        if (lane == 0) {
            // ===== weight producer: W and delta quarter tiles, ahead of the update by the ring depth =====
            const uint64_t pol_stream = l2_policy_evict_first();
            int ws = 0, wph = 0, it = 0;
            SegIter si(gp);
            WSeg g;
            while (si.next(g)) {
                const DwwLayer *L = g.L;
                const int nq = 4 * g.w;
                for (int qt = 0; qt < nq; qt++, it++) {
                    mbar_wait_bounded(&wd_empty[ws], wph ^ 1, hang, 3, it);
                    mbar_expect_tx(&wd_full[ws], WD_STAGE);
                    uint8_t *dst = wd_ring + ws * WD_STAGE;
                    const int c0 = g.nt * TN, c1 = g.ks * SLAB + qt * WD_ROWS;
                    if (gp->l2_hints) {
                        tma_load_2d_hint(dst, &L->w_map, &wd_full[ws], c0, c1, pol_stream);
                        tma_load_2d_hint(dst + WD_F32, &L->d_map, &wd_full[ws], c0, c1, pol_stream);
                    } else {
                        tma_load_2d(dst, &L->w_map, &wd_full[ws], c0, c1);
                        tma_load_2d(dst + WD_F32, &L->d_map, &wd_full[ws], c0, c1);
                    }
                    if (++ws == WDS) { ws = 0; wph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 2 && warp <= 9) {
        // ===== update warps (8): quadrant q = accumulator lanes [32q, 32q+32); `h` = which 8 of a stage's 16 rows =====
        const int e = warp - 2, q = warp & 3, h = e >> 2;
        const float mom = gp->mom, lr = gp->lr, inv_mg = 1.0f / gp->Mg;
        int ws = 0, wph = 0, it = 0, seg_it = 0;
        SegIter si(gp);
        WSeg g;
        for (; si.next(g); seg_it++) {
            const float wc = g.L->wc;
            const int acc = seg_it & 1;
            mbar_wait_bounded(&t_full[acc], (seg_it >> 1) & 1, hang, 8, seg_it);
            __syncwarp();
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + h * 8;
            const int nq = 4 * g.w;
#pragma unroll 1
            for (int qt = 0; qt < nq; qt++, it++) {
                uint8_t *st = wd_ring + ws * WD_STAGE;
                float *sw = reinterpret_cast<float *>(st) + (h * 8) * TN + q * 32 + lane;
                float *sd = sw + WD_F32 / 4;
                float gr[8];
                tmem_ld8(taddr + qt * WD_ROWS, gr);
                if (qt == nq - 1) {   // the accumulator has been drained by this warp: hand it back to the MMA issuer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[acc]);
                }
                mbar_wait_bounded(&wd_full[ws], wph, hang, 7, it);
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    // kernUpdatedelta + kernAccSum (DevFunc.cu:490-507, 427-443); g/Mg as g*(1/Mg) (<= 1 ulp)
                    const float ww = sw[x * TN];
                    const float dd = mom * sd[x * TN] - lr * (gr[x] * inv_mg + wc * ww);
                    sw[x * TN] = dd + ww;
                    sd[x * TN] = dd;
                }
                fence_async_proxy();      // the stage is read by the TMA store engine next
                __syncwarp();
                if (lane == 0) mbar_arrive(&wd_done[ws]);
                if (++ws == WDS) { ws = 0; wph ^= 1; }
            }
        }
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem);
    // last CTA out: the device-side bunch counter moves on; in data-parallel mode the step counter too, and every peer
    // learns that this rank no longer reads its factor arena (the peers' next pushes wait for that)
    if (threadIdx.x == 0 && gp->advance) {
        __threadfence();
        const unsigned int prev = atomicAdd(gp->done_counter, 1u);
        if (prev == gridDim.x - 1) {
            *gp->done_counter = 0;
            gp->ctl->bunch_idx += 1;
            if (gp->trace) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); gp->trace[FX_TRACE_WIDE + 14] = t; }
            if (gp->world > 1) {
                const unsigned int step = *gp->step_counter + 1u;
                *gp->step_counter = step;
                __threadfence_system();
                for (int p = 0; p < gp->world; p++)
                    if (p != gp->rank) st_relaxed_sys_u32(gp->peer_flags[p] + gp->rank * FX_STRIDE + FX_EV_DONE, step);
            }
        }
    }
}

int dw_wide_smem(int fblocks, int *op_stages, int *wd_stages)
{
    // Both streams are latency bound (bytes in flight per SM / round trip): the operand stream needs FB stages per segment at
    // ~1.9 us per ring round, the weight stream 204 MB through (stages x 16 KB) per ~2 us.  Measured on B200: 128-256 frames
    // are HBM bound (deep weight ring), 1024 frames operand bound (deep operand ring).
    // (2 weight stages at 1024 frames were a disaster: 260 us -- the weight stage cycle load -> update -> store is ~5 us)
    int ops = fblocks >= 16 ? 3 : 2;
    int wds = (int)((224 * 1024 - ops * dww::OP_STAGE) / dww::WD_STAGE);
    if (wds > dww::MAX_WDS) wds = dww::MAX_WDS;
    { const char *ev = getenv("GGD_WIDE_OPS"); if (ev && atoi(ev) >= 2 && atoi(ev) <= dww::MAX_OPS) { ops = atoi(ev); wds = (int)((224 * 1024 - ops * dww::OP_STAGE) / dww::WD_STAGE); if (wds > dww::MAX_WDS) wds = dww::MAX_WDS; } }
    if (wds < 2) wds = 2;
    *op_stages = ops; *wd_stages = wds;
    return ops * dww::OP_STAGE + wds * dww::WD_STAGE + 1024;
}

int launch_dw_wide(const DwwArgs *dev_args, int grid, int smem_bytes, cudaStream_t s)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(dww::NTHREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    GGD_CUDA(cudaLaunchKernelEx(&cfg, dw_wide_kernel, dev_args));
    return GGD_OK;
}

int dw_wide_init()
{
    GGD_CUDA(cudaFuncSetAttribute(dw_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    return GGD_OK;
}

}  // namespace ggd
