// gemm_tc.cu -- tcgen05 / TMEM / TMA GEMM kernel (see gemm_tc.cuh for the operand conventions).
//
// CTA = 192 threads: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (one TMEM lane quadrant each).  Output tile 128 x BN, reduction in 64-element
// blocks through a 3/4-stage mbarrier ring.  The reduction can be split over a thread-block cluster of
// S CTAs (gridDim.x = S): every CTA accumulates its share of the k-blocks in its own TMEM, then the
// partial tiles are reduce-scattered through distributed shared memory (CTA r owns columns
// [r*BN/S, (r+1)*BN/S)), summed in a fixed order, and only then the epilogue runs -- no partial sums
// ever touch L2/HBM and the result is deterministic.
#include "gemm_tc.cuh"
#include "pipe.cuh"
#include "../../include/ggd_train.h"
#include <cudaTypedefs.h>

namespace ggd {

constexpr int BK = 64;                 // reduction elements per stage (128 bytes of bf16: one swizzle row)
constexpr int TILE_I = 128;            // UMMA M
constexpr int A_TILE = TILE_I * BK * 2;
constexpr int NTHREADS = 192;           // shadow mode
constexpr int CONV_WARPS = 4;            // B_F32: warps that convert every weight tile together (4 = the epilogue warps).  Measured alternatives,
                                        // all slower or equal (118.2 us per step with this setting): 8 warps on every tile 118.3; two groups of 4
                                        // alternating over the k-blocks with separate A(3)/B(4)/raw(2) rings 120.5 -- the operand stream is bound by
                                        // the bytes TMA keeps in flight (3 operand stages + 4 raw stages here), not by the conversion arithmetic
constexpr int NTHREADS_F32 = 64 + 32 * CONV_WARPS;

// B_F32: the B operand (the weight matrix) is read as fp32 straight from the MASTER weights and split into bf16 hi/lo
// inside the kernel by the epilogue warps (idle during the main loop): no bf16 shadow copy of the weights has to be
// maintained by the update kernel (4 B/param less HBM traffic per step), same operand bytes through TMA.
template <int BN, bool B_F32 = false> struct TileCfg {
    static constexpr int B_TILE = BN * BK * 2;
    static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
    static constexpr int STAGES = (BN == 128) ? 3 : 4;             // shadow mode: A and B of a k-block share one stage
    // B_F32: three rings -- A (TMA, hi+lo), converted B (bf16 hi+lo written by the converter warps), raw fp32 B (TMA).
    // The activation operand is 2/3 of the bytes a CTA streams, so its ring stays as deep as in shadow mode.
    // B_F32: the operand stages keep the interleaved A|B layout (B = bf16 hi+lo written by the converter warps) with one
    // stage less, and a separate raw ring receives the fp32 weight tiles by TMA.
    static constexpr int A_ST = B_F32 ? ((BN == 128) ? 2 : 3) : STAGES;
    static constexpr int B_ST = A_ST;
    static constexpr int RAW_TILE = BN * BK * 4;                    // fp32 [BN rows][64], unswizzled
    static constexpr int RAW_STAGES = B_F32 ? ((BN == 128) ? 2 : 4) : 0;
    static constexpr int SMEM = A_ST * STAGE + RAW_STAGES * RAW_TILE + 1024;
    static_assert(SMEM + 12 * 1024 <= 227 * 1024, "dynamic + static shared memory (barriers, loss tile) must fit one SM");
};

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_saddr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t raddr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void stamp(const GemmArgs &g, int slot)
{
    if (g.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        g.trace[(size_t)cta * 16 + slot] = t;
    }
}

// ---- epilogues: 16 consecutive output columns [j, j+16) of output row i ----------------------------
template <int EPI>
__device__ __forceinline__ void epilogue16(const GemmArgs &g, int i, int j, float *v)
{
    const bool row_ok = i < g.I;
    if constexpr (EPI == EPI_STORE_F32) {
        if (!row_ok) return;
        float4 *dst = reinterpret_cast<float4 *>(g.o32 + (size_t)i * g.ld32 + j);
#pragma unroll
        for (int q = 0; q < 4; q++) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else if constexpr (EPI == EPI_FWD_LINEAR) {
        float4 *dst = reinterpret_cast<float4 *>(g.o32 + (size_t)i * g.ld32 + j);
        const float4 *b4 = reinterpret_cast<const float4 *>(g.bias + j);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float4 b = __ldg(b4 + q);
            float4 o;
            o.x = (row_ok && j + 4 * q + 0 < g.J) ? v[4 * q + 0] + b.x : 0.0f;
            o.y = (row_ok && j + 4 * q + 1 < g.J) ? v[4 * q + 1] + b.y : 0.0f;
            o.z = (row_ok && j + 4 * q + 2 < g.J) ? v[4 * q + 2] + b.z : 0.0f;
            o.w = (row_ok && j + 4 * q + 3 < g.J) ? v[4 * q + 3] + b.w : 0.0f;
            dst[q] = o;
        }
    } else {
        float r[16];
        if constexpr (EPI == EPI_FWD_SIGMOID) {
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const float x = v[e] + __ldg(g.bias + j + e);
                r[e] = (row_ok && j + e < g.J) ? __fdividef(1.0f, 1.0f + __expf(-x)) : 0.0f;   // kernSigmoid, DevFunc.cu:36-51 (rel. error ~2e-7)
            }
        } else {  // EPI_DX_DSIGMOID
            const uint4 *yh = reinterpret_cast<const uint4 *>(g.y_hi + (size_t)i * g.ldy + j);
            const uint4 *yl = reinterpret_cast<const uint4 *>(g.y_lo + (size_t)i * g.ldy + j);
            uint32_t h[8], l[8];
            *reinterpret_cast<uint4 *>(h) = __ldg(yh);
            *reinterpret_cast<uint4 *>(h + 4) = __ldg(yh + 1);
            *reinterpret_cast<uint4 *>(l) = __ldg(yl);
            *reinterpret_cast<uint4 *>(l + 4) = __ldg(yl + 1);
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const uint32_t hw = h[e >> 1], lw = l[e >> 1];
                const float yhi = __uint_as_float((e & 1) ? (hw & 0xFFFF0000u) : (hw << 16));
                const float ylo = __uint_as_float((e & 1) ? (lw & 0xFFFF0000u) : (lw << 16));
                const float y = yhi + ylo;
                r[e] = (row_ok && j + e < g.J) ? (1.0f - y) * y * v[e] : 0.0f;      // kernDsigmoid, DevFunc.cu:53-71
            }
        }
        uint32_t ph[8], pl[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            bf16 h0, l0, h1, l1;
            split_bf16(r[2 * e], h0, l0);
            split_bf16(r[2 * e + 1], h1, l1);
            ph[e] = pack_bf16x2(h0, h1);
            pl[e] = pack_bf16x2(l0, l1);
        }
        uint4 *oh = reinterpret_cast<uint4 *>(g.o_hi + (size_t)i * g.ldo + j);
        uint4 *ol = reinterpret_cast<uint4 *>(g.o_lo + (size_t)i * g.ldo + j);
        oh[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        oh[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
        ol[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        ol[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
    }
}

// |e|^p as in loss_kernel (kernels.cu): exact for beta = 2 and 1, otherwise exp2(p*log2 a)
__device__ __forceinline__ float pow_abs_g(float a, float p)
{
    if (p == 2.0f) return a * a;
    if (p == 1.0f) return a;
    if (p == 0.0f) return 1.0f;
    return (a > 0.0f) ? exp2f(p * __log2f(a)) : 0.0f;
}

// EPI_FWD_LOSS: the 128 epilogue threads hold rows (frames) i0..i0+127 and 16 columns [j, j+16) of the network output.
// The tile (out = acc + bias, also stored as fp32) is staged in shared memory and the loss chain then runs COLUMN-parallel
// in rolled loops (thread = one column x 16 rows): the epilogue is executed once per CTA, so straight-line unrolled code
// (3 000 instructions in the first version) was bound by instruction fetch -- 11 us instead of 1.
//   e = out - targ, s_d = sum over the 128 rows (8 row groups, fixed order), alpha_d = (beta*s_d/Mg)^(1/beta),
//   dE/dx = (1/Mg) sgn(e)|e|^(beta-1) beta [/ alpha_d^beta] as bf16 hi/lo, exactly 0 at e == 0 (DevFunc.cu:388-391, 479-482).
// With world > 1 the partial s_d are exchanged over peer memory (dp_factor.cuh) so that alpha is the unsharded minibatch's.
// `tg` / `bs`: targets and biases of this thread's row and 16 columns, loaded by the caller BEFORE it waits for the accumulator
// (they do not depend on the GEMM; the epilogue warps are idle during the main loop anyway).
__device__ __forceinline__ void epilogue_loss16(const GemmArgs &g, int i, int j, float *v, int t, const float *tg, const float *bs, int bunch)
{
    __shared__ float tile[128][17];
    __shared__ float part[8][16], colscale[16];
    const int row = i & 127;
    const float beta = g.beta;
    {
        float4 *dst = reinterpret_cast<float4 *>(g.o32 + (size_t)i * g.ld32 + j);
        const bool row_ok = i < g.I;
#pragma unroll
        for (int x = 0; x < 16; x++) {
            const float o = (row_ok && j + x < g.D) ? v[x] + bs[x] : 0.0f;
            v[x] = o;
            tile[row][x] = o - tg[x];     // e (0 for rows / columns outside the net output: tg is 0 there)
        }
#pragma unroll
        for (int x = 0; x < 4; x++) dst[x] = make_float4(v[4 * x], v[4 * x + 1], v[4 * x + 2], v[4 * x + 3]);
    }
    if (t == 64) stamp(g, 10);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (t == 64) stamp(g, 11);
    const int col = t & 15, grp = t >> 4, d = j + col;
    const int i0 = i - row;
    const bool live = d < g.D;
    {
        float s = 0.0f;
#pragma unroll
        for (int m = 0; m < 16; m++) s += pow_abs_g(fabsf(tile[grp * 16 + m][col]), beta);
        part[grp][col] = s;
    }
    if (t == 64) stamp(g, 12);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (t < 16) {
        float s = 0.0f;
#pragma unroll 1
        for (int q = 0; q < 8; q++) s += part[q][t];
        if (g.world > 1) {
            // frame-sharded data parallelism: alpha (and the loss trace) need the sum over the GLOBAL minibatch
            const unsigned int step = *g.step_counter + 1u;
            const int chunk = j >> 4;
            if (live)
                for (int pr = 0; pr < g.world; pr++) g.asum_slot[pr][(size_t)g.rank * g.D + d] = s;
            __threadfence_system();
            __syncwarp(0x0000ffffu);
            if (t == 0)
                for (int pr = 0; pr < g.world; pr++)
                    if (pr != g.rank) st_relaxed_sys_u32(g.lflags[pr] + g.rank * FX_STRIDE + FX_EV_LOSS + chunk, step);   // (fenced above)
            if (t < g.world && t != g.rank) {
                const unsigned int *f = g.lflags[g.rank] + t * FX_STRIDE + FX_EV_LOSS + chunk;
                const long long t0 = clock64();
                while ((int)(ld_acquire_sys_u32(f) - step) < 0) {
                    if (clock64() - t0 > (1ll << 32)) { *g.error_flag = 1u + t; break; }
                    __nanosleep(32);
                }
            }
            __syncwarp(0x0000ffffu);
            if (live) {
                s = 0.0f;
                for (int pr = 0; pr < g.world; pr++) s += ld_relaxed_sys_f32(g.asum_slot[g.rank] + (size_t)pr * g.D + d);   // rank order
            }
        }
        float pa = 1.0f, contrib = 0.0f;
        if (live) {
            if (g.ml) {
                const float v1 = s / (float)g.Mg;
                const float v2 = v1 * beta;
                // (the 16 threads of this block are alone on their scheduler: hardware exp2/log2, rel. error ~1e-6)
                const float al = (v2 > 0.0f) ? exp2f(__log2f(v2) / beta) : 0.0f;
                g.alpha[d] = al;
                pa = (beta == 2.0f) ? al * al : ((beta == 1.0f) ? al : v2);     // alpha^beta == beta*s/Mg by definition
                contrib = 0.69314718056f * __log2f(al) + s / ((float)g.Mg * pa);     // ln alpha_d + sum_m (|e|/alpha_d)^beta / M
            } else {
                contrib = s / (float)g.Mg;                        // E_beta = sum |e|^beta / M
            }
        }
        colscale[t] = (g.ml ? beta / pa : beta) / (float)g.Mg;
        if (g.loss_trace) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) contrib += __shfl_xor_sync(0x0000ffffu, contrib, o);
            if (t == 0) atomicAdd(g.loss_trace + bunch, (double)contrib);
        }
    }
    if (t == 0) stamp(g, 13);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    {
        const float scale = colscale[col];
        bf16 *oh = g.o_hi + (size_t)i0 * g.ldo + d, *ol = g.o_lo + (size_t)i0 * g.ldo + d;
#pragma unroll 8
        for (int m = 0; m < 16; m++) {
            const int r = grp * 16 + m;
            const float ee = tile[r][col];
            float rr = 0.0f;
            if (ee != 0.0f) {
                rr = pow_abs_g(fabsf(ee), beta - 1.0f) * scale;
                rr = (ee > 0.0f) ? rr : -rr;
            }
            bf16 hv, lv;
            split_bf16(rr, hv, lv);
            oh[(size_t)r * g.ldo] = hv;
            ol[(size_t)r * g.ldo] = lv;
        }
    }
    if (t == 64) stamp(g, 14);
    asm volatile("bar.sync 1, 128;" ::: "memory");   // tile / part / colscale are reused by the next 16-column chunk
}

template <int BN, bool A_MN, bool B_MN, int EPI, bool B_F32>
__global__ void __launch_bounds__(B_F32 ? NTHREADS_F32 : NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, const GemmArgs g)
{
    using Cfg = TileCfg<BN, B_F32>;
    constexpr int B_TILE = Cfg::B_TILE, STAGE = Cfg::STAGE, A_ST = Cfg::A_ST, B_ST = Cfg::B_ST;
    constexpr int RAW_TILE = Cfg::RAW_TILE, RAW_STAGES = B_F32 ? Cfg::RAW_STAGES : 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // operand stage addresses: shadow mode interleaves A and B per stage; B_F32 keeps an A ring, a converted-B ring, a raw ring
    auto a_stage = [&](int it) -> uint8_t * { return smem + (it % A_ST) * STAGE; };
    auto b_stage = [&](int it) -> uint8_t * { return smem + (it % B_ST) * STAGE + 2 * A_TILE; };
    uint8_t *raw_ring = smem + A_ST * STAGE;   // B_F32 only
    __shared__ __align__(8) uint64_t full_bar[A_ST], empty_bar[A_ST], tmem_full_bar;          // A (and B in shadow mode)
    __shared__ __align__(8) uint64_t bconv_bar[B_ST], bempty_bar[B_ST], raw_full_bar[RAW_STAGES], raw_empty_bar[RAW_STAGES];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = gridDim.x, rank = blockIdx.x;          // cluster = the gridDim.x CTAs that share one output tile
    const int j0 = blockIdx.y * BN, i0 = blockIdx.z * TILE_I;
    const int kb0 = (g.kblocks * rank) / S, kb1 = (g.kblocks * (rank + 1)) / S;
    const int nkb = kb1 - kb0;
    if (threadIdx.x == 0) stamp(g, 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo);
        tma_prefetch_desc(&tm_b_hi); tma_prefetch_desc(&tm_b_lo);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < A_ST; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            for (int s = 0; s < B_ST; s++) { mbar_init(&bconv_bar[s], B_F32 ? CONV_WARPS : 4); mbar_init(&bempty_bar[s], 1); }
            for (int s = 0; s < RAW_STAGES; s++) { mbar_init(&raw_full_bar[s], 1); mbar_init(&raw_empty_bar[s], B_F32 ? CONV_WARPS : 4); }
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
    float loss_tg[16], loss_bs[16];   // EPI_FWD_LOSS: prefetched targets / biases (dead code otherwise)
    int loss_bunch = 0;
    // Everything above overlapped the tail of the previous kernel (PDL).  Each role waits for the previous grid
    // (griddepcontrol.wait) right before it first touches what that grid produced -- and not earlier: with B_F32 the weight
    // tiles do not depend on the previous kernel of the chain (g.b_early), so their TMA loads and their conversion start
    // while the previous kernel is still draining.

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int early = 0;
            if constexpr (B_F32) {
                if (g.b_early) {
                    early = nkb < RAW_STAGES ? nkb : RAW_STAGES;
                    for (int it = 0; it < early; it++) {
                        const int r0 = (kb0 + it) * BK;
                        mbar_expect_tx(&raw_full_bar[it], RAW_TILE);
                        uint8_t *rw = raw_ring + it * RAW_TILE;
                        if constexpr (!B_MN) {
                            tma_load_2d(rw, &tm_b_hi, &raw_full_bar[it], r0, j0);
                        } else {
#pragma unroll
                            for (int h = 0; h < BN / 64; h++) tma_load_2d(rw + h * 16384, &tm_b_hi, &raw_full_bar[it], j0 + 64 * h, r0);
                        }
                    }
                }
            }
            if constexpr (!B_F32) {
                // shadow mode: the bf16 weight tiles go straight into the operand stages; their bytes are announced without
                // the arrival, which follows with the activation tiles after griddepcontrol.wait
                if (g.b_early) {
                    early = nkb < A_ST ? nkb : A_ST;
                    for (int it = 0; it < early; it++) {
                        const int r0 = (kb0 + it) * BK;
                        mbar_expect_tx_only(&full_bar[it], 2 * B_TILE);
                        uint8_t *sb = b_stage(it);
                        if constexpr (!B_MN) {
                            tma_load_2d(sb, &tm_b_hi, &full_bar[it], r0, j0);
                            tma_load_2d(sb + B_TILE, &tm_b_lo, &full_bar[it], r0, j0);
                        } else {
#pragma unroll
                            for (int h = 0; h < BN / 64; h++) {
                                tma_load_2d(sb + h * 8192, &tm_b_hi, &full_bar[it], j0 + 64 * h, r0);
                                tma_load_2d(sb + B_TILE + h * 8192, &tm_b_lo, &full_bar[it], j0 + 64 * h, r0);
                            }
                        }
                    }
                }
            }
            pdl_wait();
            const int a_row_off = g.a_rows_from_ctl ? g.ctl->bunch_idx * g.rows_per_bunch : 0;
            stamp(g, 1);
            for (int it = 0; it < nkb; it++) {
                const int s = it % A_ST, ph = (it / A_ST) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], (B_F32 || it < early) ? 2 * A_TILE : STAGE);
                uint8_t *st = a_stage(it);
                const int r0 = (kb0 + it) * BK;
                if constexpr (!A_MN) {
                    tma_load_2d(st, &tm_a_hi, &full_bar[s], r0, i0 + a_row_off);
                    tma_load_2d(st + A_TILE, &tm_a_lo, &full_bar[s], r0, i0 + a_row_off);
                } else {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        tma_load_2d(st + h * 8192, &tm_a_hi, &full_bar[s], i0 + 64 * h, r0 + a_row_off);
                        tma_load_2d(st + A_TILE + h * 8192, &tm_a_lo, &full_bar[s], i0 + 64 * h, r0 + a_row_off);
                    }
                }
                uint8_t *sb = b_stage(it);
                if constexpr (B_F32) {
                    // fp32 weight tile [BN rows][64 floats] (tm_b_hi is the fp32 map of the master weights)
                    if (it < early) { if (it == 0) stamp(g, 2); continue; }
                    const int rs = it % RAW_STAGES, rph = (it / RAW_STAGES) & 1;
                    mbar_wait(&raw_empty_bar[rs], rph ^ 1);
                    mbar_expect_tx(&raw_full_bar[rs], RAW_TILE);
                    uint8_t *rw = raw_ring + rs * RAW_TILE;
                    if constexpr (!B_MN) {
                        tma_load_2d(rw, &tm_b_hi, &raw_full_bar[rs], r0, j0);
                    } else {
#pragma unroll
                        for (int h = 0; h < BN / 64; h++) tma_load_2d(rw + h * 16384, &tm_b_hi, &raw_full_bar[rs], j0 + 64 * h, r0);
                    }
                } else if (it < early) {
                    // already requested before griddepcontrol.wait
                } else if constexpr (!B_MN) {
                    tma_load_2d(sb, &tm_b_hi, &full_bar[s], r0, j0);
                    tma_load_2d(sb + B_TILE, &tm_b_lo, &full_bar[s], r0, j0);
                } else {
#pragma unroll
                    for (int h = 0; h < BN / 64; h++) {
                        tma_load_2d(sb + h * 8192, &tm_b_hi, &full_bar[s], j0 + 64 * h, r0);
                        tma_load_2d(sb + B_TILE + h * 8192, &tm_b_lo, &full_bar[s], j0 + 64 * h, r0);
                    }
                }
                if (it == 0) stamp(g, 2);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = make_idesc_bf16(TILE_I, BN, A_MN, B_MN);
            constexpr uint32_t A_LBO = A_MN ? 8192u : 16u, B_LBO = B_MN ? 8192u : 16u;
            constexpr uint32_t A_KSTEP = A_MN ? 2048u : 32u, B_KSTEP = B_MN ? 2048u : 32u;
            for (int it = 0; it < nkb; it++) {
                const int s = it % A_ST, ph = (it / A_ST) & 1;
                const int sb = it % B_ST;
                mbar_wait(&full_bar[s], ph);
                if constexpr (B_F32) mbar_wait(&bconv_bar[sb], (it / B_ST) & 1);
                tc_fence_after();
                if (it == 0) stamp(g, 3);
                if (it == nkb - 1) stamp(g, 4);
                const uint32_t a_hi = smem_u32(a_stage(it)), a_lo = a_hi + A_TILE;
                const uint32_t b_hi = smem_u32(b_stage(it)), b_lo = b_hi + B_TILE;
#pragma unroll
                for (int k = 0; k < BK / 16; k++) {
                    const uint64_t dah = make_smem_desc(a_hi + k * A_KSTEP, A_LBO, 1024);
                    const uint64_t dal = make_smem_desc(a_lo + k * A_KSTEP, A_LBO, 1024);
                    const uint64_t dbh = make_smem_desc(b_hi + k * B_KSTEP, B_LBO, 1024);
                    const uint64_t dbl = make_smem_desc(b_lo + k * B_KSTEP, B_LBO, 1024);
                    umma_bf16(tmem, dal, dbh, idesc, (it | k) != 0);   // small terms first
                    umma_bf16(tmem, dah, dbl, idesc, 1);
                    umma_bf16(tmem, dah, dbh, idesc, 1);
                }
                umma_commit(&empty_bar[s]);   // frees the smem stage when these MMAs retire
                if constexpr (B_F32) umma_commit(&bempty_bar[sb]);
            }
            umma_commit(&tmem_full_bar);
            stamp(g, 5);
        }
        __syncwarp();
    } else {
        // ===== epilogue warps (2..5; with B_F32 also converter group 0) and converter group 1 (warps 6..9, B_F32 only) =====
        const bool epi_warp = warp < 6;
        if constexpr (!B_F32) pdl_wait();
        if constexpr (EPI == EPI_FWD_LOSS) if (epi_warp) {
            pdl_wait();
            // targets and biases of this thread's row / first 16-column chunk: requested now, used after the main loop
            const int qq = warp & 3, ii = i0 + qq * 32 + lane, jj = j0 + rank * (BN / S);
            loss_bunch = g.ctl->bunch_idx;
            const bool row_ok = ii < g.I;
            const float *trow = g.ctl->targ + ((size_t)loss_bunch * g.I + (row_ok ? ii : 0)) * g.D + jj;
#pragma unroll
            for (int x = 0; x < 16; x++) {
                loss_tg[x] = (row_ok && jj + x < g.D) ? __ldg(trow + x) : 0.0f;
                loss_bs[x] = (jj + x < g.D) ? __ldg(g.bias + jj + x) : 0.0f;
            }
        }
        if constexpr (B_F32) {
            // ===== weight converter: fp32 tile -> bf16 hi / lo tiles in the SWIZZLE_128B layout the MMA descriptors expect
            // (row r of 128 bytes, 16-byte chunk c stored at chunk position c ^ (r & 7)) =====
            constexpr int CT = 32 * CONV_WARPS;
            const int ct = threadIdx.x - 64;                       // 0..CT-1
            for (int it = 0; it < nkb; it++) {
                const int sb = it % B_ST, bph = (it / B_ST) & 1;
                const int rs = it % RAW_STAGES, rph = (it / RAW_STAGES) & 1;
                mbar_wait(&raw_full_bar[rs], rph);
                const uint8_t *rw = raw_ring + rs * RAW_TILE;
                uint8_t *bh = b_stage(it), *bl = bh + B_TILE;
                constexpr int GROUPS = BN * 8 / CT;                 // (row, 16-byte chunk) groups per thread
                uint4 hv[GROUPS], lv[GROUPS];
#pragma unroll
                for (int u = 0; u < GROUPS; u++) {
                    const int gi = u * CT + ct, r = gi >> 3, c = gi & 7;
                    const float4 f0 = *reinterpret_cast<const float4 *>(rw + (size_t)r * 256 + c * 32);
                    const float4 f1 = *reinterpret_cast<const float4 *>(rw + (size_t)r * 256 + c * 32 + 16);
                    const float x[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
                    uint32_t hh[4], ll[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        // hi = RN(x), lo = RN(x - hi), two elements per (packed) convert
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
                        const float2 hf = __bfloat1622float2(h2);
                        const __nv_bfloat162 l2 = __floats2bfloat162_rn(x[2 * e] - hf.x, x[2 * e + 1] - hf.y);
                        hh[e] = *reinterpret_cast<const uint32_t *>(&h2);
                        ll[e] = *reinterpret_cast<const uint32_t *>(&l2);
                    }
                    hv[u] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                    lv[u] = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&raw_empty_bar[rs]);     // the fp32 tile is in registers
                mbar_wait(&bempty_bar[sb], bph ^ 1);                // the MMAs that read this converted-B stage have retired
#pragma unroll
                for (int u = 0; u < GROUPS; u++) {
                    const int gi = u * CT + ct, r = gi >> 3, c = gi & 7;
                    const size_t o = (size_t)r * 128 + (size_t)((c ^ (r & 7)) << 4);
                    *reinterpret_cast<uint4 *>(bh + o) = hv[u];
                    *reinterpret_cast<uint4 *>(bl + o) = lv[u];
                }
                fence_async_proxy();                                // generic-proxy writes -> tensor-core (async proxy) reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&bconv_bar[sb]);
            }
        }
        if constexpr (B_F32) pdl_wait();   // (returns at once by now) before the epilogue reads / writes global memory
        if (epi_warp) {
            mbar_wait(&tmem_full_bar, 0);
            tc_fence_after();
        }
        if (threadIdx.x == 64) stamp(g, 6);
    }
    pdl_trigger();   // mainloop done on this CTA: the next kernel may be scheduled as SMs drain

    const int q = warp & 3;                 // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;          // output row inside the tile
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int W = BN / S;                   // columns owned by each CTA of the cluster
    // receive buffer [S][W/4][128] float4, aliases the (drained) pipeline stages.  The row index is the fastest
    // dimension so that the 32 lanes of a warp write 512 contiguous bytes per remote store instruction.
    float4 *recv = reinterpret_cast<float4 *>(smem);
    const int W4 = W >> 2;

    if (S > 1) {
        cluster_sync_all();   // every CTA of the cluster has retired its MMAs: stage memory is free everywhere
        if (warp >= 2 && warp < 6) {
            for (int p = 0; p < S; p++) {
                if (p == rank) continue;
                const uint32_t dst0 = map_to_cta(smem_u32(recv + ((size_t)rank * W4) * TILE_I + row), p);
                for (int c = 0; c < W; c += 16) {
                    float v[16];
                    tmem_ld16(trow + p * W + c, v);
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        st_cluster_f4(dst0 + (uint32_t)(((c >> 2) + e) * TILE_I) * 16u, v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                }
            }
        }
        cluster_sync_all();   // all partial slabs have landed
    }
    if (threadIdx.x == 64) stamp(g, 7);
    if (warp >= 2 && warp < 6) {
        for (int c = 0; c < W; c += 16) {
            float v[16];
            tmem_ld16(trow + rank * W + c, v);
            for (int p = 0; p < S; p++) {
                if (p == rank) continue;
                const float4 *src = recv + ((size_t)p * W4 + (c >> 2)) * TILE_I + row;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float4 t = src[e * TILE_I];
                    v[4 * e] += t.x; v[4 * e + 1] += t.y; v[4 * e + 2] += t.z; v[4 * e + 3] += t.w;
                }
            }
            if constexpr (EPI == EPI_FWD_LOSS) {
                if (c > 0) {   // further chunks of a wide tile (S < 4): load inline
                    const int ii = i0 + row, jj = j0 + rank * W + c;
                    const bool row_ok = ii < g.I;
                    const float *trow = g.ctl->targ + ((size_t)loss_bunch * g.I + (row_ok ? ii : 0)) * g.D + jj;
#pragma unroll
                    for (int x = 0; x < 16; x++) {
                        loss_tg[x] = (row_ok && jj + x < g.D) ? __ldg(trow + x) : 0.0f;
                        loss_bs[x] = (jj + x < g.D) ? __ldg(g.bias + jj + x) : 0.0f;
                    }
                }
                epilogue_loss16(g, i0 + row, j0 + rank * W + c, v, row, loss_tg, loss_bs, loss_bunch);
            }
            else epilogue16<EPI>(g, i0 + row, j0 + rank * W + c, v);
        }
    }
    if (threadIdx.x == 64) stamp(g, 8);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<BN>(tmem);
    if (threadIdx.x == 0) stamp(g, 9);
}

// ---- host side -----------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int make_tmap_bf16(CUtensorMap *m, const bf16 *base, long long rows, long long cols, long long ld, int box_rows)
{
    if (!g_encode) { set_error("tensor-map encoder not initialised"); return GGD_ECUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(bf16)};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16 *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld); return GGD_ECUDA; }
    return GGD_OK;
}

int make_tmap_2d(CUtensorMap *m, const void *base, int dtype_f32, long long rows, long long cols, long long ld, int box_cols, int box_rows,
                 int swizzle128)
{
    if (!g_encode) { set_error("tensor-map encoder not initialised"); return GGD_ECUDA; }
    const size_t es = dtype_f32 ? 4 : 2;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = g_encode(m, dtype_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r, rows, cols, ld, box_cols, box_rows); return GGD_ECUDA; }
    return GGD_OK;
}

template <int BN, bool A_MN, bool B_MN, int EPI, bool B_F32 = false>
static int launch_inst(const GemmPlan &p, cudaStream_t s)
{
    auto kern = gemm_tc_kernel<BN, A_MN, B_MN, EPI, B_F32>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.splits, p.tiles_j, p.tiles_i);
    cfg.blockDim = dim3(B_F32 ? NTHREADS_F32 : NTHREADS);
    cfg.dynamicSmemBytes = TileCfg<BN, B_F32>::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = p.splits; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = p.no_pdl ? 1 : 2;
    GGD_CUDA(cudaLaunchKernelEx(&cfg, kern, p.a_hi, p.a_lo, p.b_hi, p.b_lo, p.args));
    return GGD_OK;
}

template <int BN>
static int launch_bn(const GemmPlan &p, cudaStream_t s)
{
    const int key = p.a_mn * 100 + p.b_mn * 10 + p.epi;
    if (p.b_f32) {   // B operand = fp32 master weights, split in the kernel
        switch (key) {
        case 10 + EPI_FWD_SIGMOID: return launch_inst<BN, false, true, EPI_FWD_SIGMOID, true>(p, s);
        case 10 + EPI_FWD_LINEAR:  return launch_inst<BN, false, true, EPI_FWD_LINEAR, true>(p, s);
        case 10 + EPI_FWD_LOSS:    return launch_inst<BN, false, true, EPI_FWD_LOSS, true>(p, s);
        case 0 + EPI_DX_DSIGMOID:  return launch_inst<BN, false, false, EPI_DX_DSIGMOID, true>(p, s);
        default: set_error("gemm_tc: no fp32-B variant for a_mn=%d b_mn=%d epi=%d", p.a_mn, p.b_mn, p.epi); return GGD_EINVAL;
        }
    }
    switch (key) {
    case 10 + EPI_FWD_SIGMOID: return launch_inst<BN, false, true, EPI_FWD_SIGMOID>(p, s);
    case 10 + EPI_FWD_LINEAR:  return launch_inst<BN, false, true, EPI_FWD_LINEAR>(p, s);
    case 10 + EPI_FWD_LOSS:    return launch_inst<BN, false, true, EPI_FWD_LOSS>(p, s);
    case 10 + EPI_STORE_F32:   return launch_inst<BN, false, true, EPI_STORE_F32>(p, s);
    case 0 + EPI_DX_DSIGMOID:  return launch_inst<BN, false, false, EPI_DX_DSIGMOID>(p, s);
    case 0 + EPI_STORE_F32:    return launch_inst<BN, false, false, EPI_STORE_F32>(p, s);
    case 110 + EPI_STORE_F32:  return launch_inst<BN, true, true, EPI_STORE_F32>(p, s);
    default: set_error("gemm_tc: unsupported variant a_mn=%d b_mn=%d epi=%d", p.a_mn, p.b_mn, p.epi); return GGD_EINVAL;
    }
}

int launch_gemm_tc(const GemmPlan &p, cudaStream_t s)
{
    if (p.splits < 1 || p.splits > 8 || (p.splits & (p.splits - 1)) || (p.bn / p.splits) % 16 != 0 || p.splits > p.args.kblocks) {
        set_error("gemm_tc: bad split %d for bn=%d kblocks=%d", p.splits, p.bn, p.args.kblocks);
        return GGD_EINVAL;
    }
    if (p.bn == 128) return launch_bn<128>(p, s);
    if (p.bn == 64) return launch_bn<64>(p, s);
    set_error("gemm_tc: bn must be 64 or 128");
    return GGD_EINVAL;
}

template <int BN, bool A_MN, bool B_MN, int EPI, bool B_F32 = false>
static int prep_inst()
{
    GGD_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN, EPI, B_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<BN, B_F32>::SMEM));
    return GGD_OK;
}
template <int BN>
static int prep_bn()
{
    int rc;
    if ((rc = prep_inst<BN, false, true, EPI_FWD_SIGMOID>())) return rc;
    if ((rc = prep_inst<BN, false, true, EPI_FWD_LINEAR>())) return rc;
    if ((rc = prep_inst<BN, false, true, EPI_FWD_LOSS>())) return rc;
    if ((rc = prep_inst<BN, false, true, EPI_STORE_F32>())) return rc;
    if ((rc = prep_inst<BN, false, false, EPI_DX_DSIGMOID>())) return rc;
    if ((rc = prep_inst<BN, false, false, EPI_STORE_F32>())) return rc;
    if ((rc = prep_inst<BN, true, true, EPI_STORE_F32>())) return rc;
    if ((rc = prep_inst<BN, false, true, EPI_FWD_SIGMOID, true>())) return rc;
    if ((rc = prep_inst<BN, false, true, EPI_FWD_LINEAR, true>())) return rc;
    if ((rc = prep_inst<BN, false, true, EPI_FWD_LOSS, true>())) return rc;
    if ((rc = prep_inst<BN, false, false, EPI_DX_DSIGMOID, true>())) return rc;
    return GGD_OK;
}

int gemm_tc_init()
{
    // per-device function attributes (cheap; repeated calls are harmless)
    int rc;
    if ((rc = prep_bn<64>())) return rc;
    if ((rc = prep_bn<128>())) return rc;
    if (g_encode) return GGD_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GGD_CUDA(cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) { set_error("cuTensorMapEncodeTiled not available"); return GGD_ECUDA; }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return GGD_OK;
}

}  // namespace ggd
