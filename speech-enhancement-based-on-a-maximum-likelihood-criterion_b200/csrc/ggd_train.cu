// ggd_train.cu -- the C ABI of libggd_b200 (include/ggd_train.h): device workspace, chunk upload,
// the per-bunch training step (captured once as a CUDA graph and replayed for every bunch of a chunk),
// cross-validation forward passes and weight export.
//
// Reference behaviour restated (not ported): Train_code_ML_GGD/BP_GPU.cu.  Differences by design:
//   * 53 launches + 5 host syncs per bunch become one graph replay per 16 bunches;
//   * cuBLAS fp32 SGEMM becomes the tcgen05 bf16x3 GEMM of gemm_tc.cu (no cuBLAS anywhere);
//   * the 11-kernel loss chain runs in the output-layer GEMM's epilogue; the 4-launch-per-layer update is one launch per step;
//   * all weights live in one padded arena so that update / allreduce are single flat passes;
//   * data parallelism exchanges the low-rank FACTORS of the gradient over NVLink peer memory (dp_factor.cuh).
#include "../../include/ggd_train.h"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "dw_wide.cuh"
#include "dp_factor.cuh"
#include <nccl.h>
#include <stdarg.h>
#include <unistd.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <string>
#include <map>
#include <algorithm>

namespace ggd {
static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
}  // namespace ggd

using namespace ggd;

#define GGD_NCCL(expr)                                                                            \
    do {                                                                                          \
        ncclResult_t _r = (expr);                                                                 \
        if (_r != ncclSuccess) {                                                                  \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, ncclGetErrorString(_r));      \
            return GGD_ENCCL;                                                                     \
        }                                                                                         \
    } while (0)
#define GGD_TRY(expr)              \
    do {                           \
        int _rc = (expr);          \
        if (_rc != GGD_OK) return _rc; \
    } while (0)

constexpr int HANG_WORDS = 8 + 160 * 12 * 4;
constexpr int FX_NEVENTS = 64;      // fork / join events of one step
constexpr int FX_NSIDE = 12;        // side streams: every factor push (and the bias update) is its own branch of the step graph
// fx_counters words
constexpr int FXC_STEP = 0, FXC_ERROR = 3, FXC_PUSH_BLOCKS = 8, FXC_WORDS = 8 + FX_STRIDE;

struct LayerInfo {
    int prev, cur;      // real units
    int Kp, Np;         // padded units (multiples of 64)
    size_t w_off, b_off;  // element offsets in the parameter arenas
    size_t gb_off;        // offset of the bias gradient in G (packed after the arena so that one allreduce covers all biases)
};

struct ggd_handle {
    ggd_config cfg;
    int L, M, Mp, Mg, sm_count;
    int units[GGD_MAXLAYER], upad[GGD_MAXLAYER];
    LayerInfo lay[GGD_MAXLAYER];
    bool tensor;        // GGD_PREC_BF16X3
    // parameter arenas (same element offsets in each)
    size_t arena;
    float *P, *Dl, *G;
    bf16 *Phi, *Plo;
    // FACTOR ARENA (one allocation, tensor path): per layer the activations and dE/dx of one minibatch as bf16 hi/lo
    // [fx_rows][upad].  fx_rows = Mp on one GPU; world*Mp with factor-exchange data parallelism, where this rank's GEMMs
    // read and write the slice of rows [loc_row, loc_row + Mp) and the peers push the other slices (dp_factor.cuh).
    uint8_t *fx_arena;
    size_t fx_bytes;
    int fx_rows, loc_row;
    bf16 *act_hi[GGD_MAXLAYER], *act_lo[GGD_MAXLAYER], *dx_hi[GGD_MAXLAYER], *dx_lo[GGD_MAXLAYER];   // whole arrays (row 0)
    float *x32[GGD_MAXLAYER], *y32[GGD_MAXLAYER], *dy32[GGD_MAXLAYER], *dx32[GGD_MAXLAYER];
    float *out32;       // [Mp][upad[L-1]]
    float *alpha, *colsum;
    double *trace;
    size_t trace_cap;
    StepCtl *ctl;
    // raw-record staging of the device-side loader (grow-only)
    unsigned int *r_fea, *r_targ; int *r_first; float *r_norm; size_t r_cap_fea, r_cap_targ, r_cap_first, r_cap_norm;
    // chunk staging
    float *c_in, *c_targ, *c_out;
    bf16 *c_hi, *c_lo;
    size_t cap;         // frames
    std::map<const void *, size_t> pinned;
    cudaStream_t s_main, s_side[FX_NSIDE], s_copy;
    std::vector<cudaEvent_t> ev_piece;   // upload pipeline: piece p of the chunk has landed
    cudaEvent_t ev_c0, ev_c1;
    cudaEvent_t ev0, ev1, ev2;
    cudaEvent_t ev_fx[FX_NEVENTS];       // fork / join between the compute stream and the side stream (pushes, bias update)
    size_t nbias;       // packed bias-gradient elements
    // plans + graphs
    GemmPlan fwd[GGD_MAXLAYER], dxp[GGD_MAXLAYER], dwp[GGD_MAXLAYER];
    GemmPlan fwd_loss;  // output layer of a training step: the loss-gradient chain runs in its epilogue (EPI_FWD_LOSS)
    bool fuse_loss;
    bool w_f32;         // GEMMs read the fp32 master weights and split them in-kernel: no bf16 weight shadows are maintained
    bool fused;         // gradient GEMM + update fused (tensor path; one GPU or factor-exchange data parallelism)
    bool persist;       // fused, one GPU, the bunch is one reduction tile: dw_persist.cu
    bool wide;          // fused, any other case: dw_wide.cu (weights and biases)
    DwpArgs *dwp_dev;   // argument block of dw_persist (device memory)
    DwwArgs *dww_dev;   // argument block of dw_wide (device memory)
    int wide_smem;
    unsigned int *dwp_counter;
    unsigned int *hang_host, *hang_dev;   // host-mapped record written by a device-side watchdog before it traps
    cudaGraphExec_t g1, g4, gN;     // step graphs of 1, 4 and 16 bunches (a chunk is 16a + 4b + c bunches)
    int gN_steps;
    int launches_per_step;
    // data parallelism
    ncclComm_t comm;
    bool has_comm;
    bool dp_fx;         // factor exchange over NVLink peer memory (dp_factor.cuh); otherwise NCCL allreduce of the gradient arena
    bool fx_loss;       // sum|e|^beta exchanged over peer memory inside the loss epilogue / loss_kernel
    float *fx_asum;                 // receive slots of the sum|e|^beta partials [world][D]
    unsigned int *fx_flags;         // my flag block [world][FX_STRIDE]
    unsigned int *fx_counters;      // FXC_*
    void *fx_peer[3][FX_MAX];       // IPC-mapped: factor arena, asum, flags of every rank
    FxPushArgs fx_push[FX_STRIDE];  // by event (FX_EV_Y + l, FX_EV_DX + l)
    int fx_push_ctas;
    unsigned long long *fx_trace;   // GGD_FX_TRACE=1: globaltimer stamps of the last step's pushes / wide / bias kernels
    // host mirrors / stats
    std::vector<float> losses;
    std::vector<float> h_out;
    ggd_stats stats;
    // optional per-kernel timing (ggd_profile_kernels): events around every launch, by kernel class
    bool prof_on;
    std::vector<cudaEvent_t> prof_ev;     // pairs
    std::vector<int> prof_cls;
};

enum { KC_FWD = 0, KC_LOSS, KC_DX, KC_DW, KC_BIAS, KC_ALLREDUCE, KC_UPDATE, KC_ADVANCE, KC_SPLIT, KC_DWUPD, KC_PUSH, KC_COUNT };
static_assert((int)KC_COUNT == (int)GGD_KC_COUNT, "kernel classes out of sync with include/ggd_train.h");

struct ProfScope {
    ggd_handle *h; cudaStream_t s;
    ProfScope(ggd_handle *h_, int cls, cudaStream_t s_) : h(h_), s(s_) {
        if (!h->prof_on) return;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        h->prof_ev.push_back(a); h->prof_ev.push_back(b); h->prof_cls.push_back(cls);
        cudaEventRecord(a, s);
    }
    ~ProfScope() { if (h->prof_on) cudaEventRecord(h->prof_ev.back(), s); }
};

// this rank's slice of a factor array
static inline bf16 *loc(const ggd_handle *h, bf16 *base, int l) { return base + (size_t)h->loc_row * h->upad[l]; }

static void free_chunk(ggd_handle *h)
{
    cudaFree(h->c_in); cudaFree(h->c_targ); cudaFree(h->c_out); cudaFree(h->c_hi); cudaFree(h->c_lo); cudaFree(h->trace);
    h->c_in = h->c_targ = h->c_out = nullptr; h->c_hi = h->c_lo = nullptr; h->trace = nullptr;
    h->cap = 0;
    if (h->g1) { cudaGraphExecDestroy(h->g1); h->g1 = nullptr; }
    if (h->gN) { cudaGraphExecDestroy(h->gN); h->gN = nullptr; }
    if (h->g4) { cudaGraphExecDestroy(h->g4); h->g4 = nullptr; }
}

// Tile width and cluster split of one GEMM.  Measured on B200 (tools/gemm_probe.py): an SM pulls ~100 GB/s through
// TMA, clusters of 8 CTAs with ~200 KB of shared memory each only co-schedule 64 CTAs (two waves), and the
// DSMEM reduce-scatter costs ~1 us per 8 KB received.  The model below picks (bn, S) with S <= 4 accordingly.
static void pick_tile(int tiles_i, int Np, int kblocks, int sm, int *bn_out, int *splits_out)
{
    double best = 1e30;
    *bn_out = 64; *splits_out = 1;
    for (int bn = 64; bn <= 128; bn += 64) {
        if (Np % bn) continue;
        for (int s = 1; s <= 4; s *= 2) {
            if (s > kblocks || (bn / s) % 16) continue;
            const int ctas = tiles_i * (Np / bn) * s;
            const int waves = (ctas + sm - 1) / sm;
            const double stage_kb = 32.0 + bn * 0.25;                       // A hi/lo + B hi/lo per k-block
            const double kb_per_cta = (double)((kblocks + s - 1) / s);
            const double t_main = 1.3 + kb_per_cta * stage_kb / 100.0;       // us: latency + bytes at ~100 KB/us per SM
            const double t_xchg = (s > 1) ? 0.8 + (s - 1) * (128.0 * (bn / s) * 4.0 / 1024.0) / 16.0 : 0.0;
            const double t = waves * (t_main + t_xchg + 1.0 + bn / 128.0);
            if (t < best) { best = t; *bn_out = bn; *splits_out = s; }
        }
    }
}

// (re)build tensor maps and GEMM plans; needs the chunk buffers for the layer-1 operands
static int build_plans(ggd_handle *h)
{
    const int L = h->L;
    const int world = h->dp_fx ? h->cfg.world_size : 1, rank = h->dp_fx ? h->cfg.rank : 0;
    for (int l = 1; l < L; l++) {
        const LayerInfo &ly = h->lay[l];
        const bool first = (l == 1), last = (l == L - 1);
        const bf16 *ah = first ? h->c_hi : loc(h, h->act_hi[l - 1], l - 1), *al = first ? h->c_lo : loc(h, h->act_lo[l - 1], l - 1);
        const long long arows = first ? (long long)h->cap : h->Mp;
        bf16 *dxh = loc(h, h->dx_hi[l], l), *dxl = loc(h, h->dx_lo[l], l);
        // ---- forward: x[m][n] = sum_k y[m][k] W[k][n]
        {
            GemmPlan &p = h->fwd[l];
            memset(&p, 0, sizeof p);
            pick_tile(h->Mp / 128, ly.Np, ly.Kp / 64, h->sm_count, &p.bn, &p.splits);
            p.a_mn = 0; p.b_mn = 1;
            p.epi = last ? EPI_FWD_LINEAR : EPI_FWD_SIGMOID;
            p.tiles_i = h->Mp / 128; p.tiles_j = ly.Np / p.bn;
            GGD_TRY(make_tmap_bf16(&p.a_hi, ah, arows, ly.Kp, ly.Kp, 128));
            GGD_TRY(make_tmap_bf16(&p.a_lo, al, arows, ly.Kp, ly.Kp, 128));
            if (h->w_f32) {
                p.b_f32 = 1;
                GGD_TRY(make_tmap_2d(&p.b_hi, h->P + ly.w_off, 1, ly.Kp, ly.Np, ly.Np, 64, 64, 0));
                p.b_lo = p.b_hi;
            } else {
                GGD_TRY(make_tmap_bf16(&p.b_hi, h->Phi + ly.w_off, ly.Kp, ly.Np, ly.Np, 64));
                GGD_TRY(make_tmap_bf16(&p.b_lo, h->Plo + ly.w_off, ly.Kp, ly.Np, ly.Np, 64));
            }
            GemmArgs &a = p.args;
            a.ctl = h->ctl; a.a_rows_from_ctl = first; a.rows_per_bunch = h->M;
            a.I = h->M; a.J = ly.cur; a.kblocks = ly.Kp / 64;
            a.bias = h->P + ly.b_off;
            a.o_hi = loc(h, h->act_hi[l], l); a.o_lo = loc(h, h->act_lo[l], l); a.ldo = ly.Np;
            a.o32 = h->out32; a.ld32 = ly.Np;
        }
        // ---- backward: dE/dy[m][k] = sum_n dE/dx[m][n] W[k][n], times y(1-y) of layer l-1
        if (!first) {
            GemmPlan &p = h->dxp[l];
            memset(&p, 0, sizeof p);
            pick_tile(h->Mp / 128, ly.Kp, ly.Np / 64, h->sm_count, &p.bn, &p.splits);
            p.a_mn = 0; p.b_mn = 0;
            p.epi = EPI_DX_DSIGMOID;
            p.tiles_i = h->Mp / 128; p.tiles_j = ly.Kp / p.bn;
            GGD_TRY(make_tmap_bf16(&p.a_hi, dxh, h->Mp, ly.Np, ly.Np, 128));
            GGD_TRY(make_tmap_bf16(&p.a_lo, dxl, h->Mp, ly.Np, ly.Np, 128));
            if (h->w_f32) {
                p.b_f32 = 1;
                GGD_TRY(make_tmap_2d(&p.b_hi, h->P + ly.w_off, 1, ly.Kp, ly.Np, ly.Np, 64, p.bn, 0));
                p.b_lo = p.b_hi;
            } else {
                GGD_TRY(make_tmap_bf16(&p.b_hi, h->Phi + ly.w_off, ly.Kp, ly.Np, ly.Np, p.bn));
                GGD_TRY(make_tmap_bf16(&p.b_lo, h->Plo + ly.w_off, ly.Kp, ly.Np, ly.Np, p.bn));
            }
            GemmArgs &a = p.args;
            a.ctl = h->ctl; a.a_rows_from_ctl = 0; a.rows_per_bunch = h->M;
            a.I = h->M; a.J = ly.prev; a.kblocks = ly.Np / 64;
            a.o_hi = loc(h, h->dx_hi[l - 1], l - 1); a.o_lo = loc(h, h->dx_lo[l - 1], l - 1); a.ldo = ly.Kp;
            a.y_hi = loc(h, h->act_hi[l - 1], l - 1); a.y_lo = loc(h, h->act_lo[l - 1], l - 1); a.ldy = ly.Kp;
        }
        // ---- gradient of this rank's frames, materialised: g[k][n] = sum_m y[m][k] dE/dx[m][n]  (validation / NCCL mode)
        {
            GemmPlan &p = h->dwp[l];
            memset(&p, 0, sizeof p);
            p.bn = (ly.Np % 128 == 0) ? 128 : 64;
            p.a_mn = 1; p.b_mn = 1;
            p.epi = EPI_STORE_F32;
            p.tiles_i = ceil_div(ly.Kp, 128); p.tiles_j = ly.Np / p.bn;
            GGD_TRY(make_tmap_bf16(&p.a_hi, ah, arows, ly.Kp, ly.Kp, 64));
            GGD_TRY(make_tmap_bf16(&p.a_lo, al, arows, ly.Kp, ly.Kp, 64));
            GGD_TRY(make_tmap_bf16(&p.b_hi, dxh, h->Mp, ly.Np, ly.Np, 64));
            GGD_TRY(make_tmap_bf16(&p.b_lo, dxl, h->Mp, ly.Np, ly.Np, 64));
            GemmArgs &a = p.args;
            a.ctl = h->ctl; a.a_rows_from_ctl = first; a.rows_per_bunch = h->M;
            a.I = ly.Kp; a.J = ly.Np; a.kblocks = h->Mp / 64;
            a.o32 = h->G + ly.w_off; a.ld32 = ly.Np;
            p.splits = 1;
        }
    }
    for (int l = 1; l < L; l++) {
        h->fwd[l].args.hang = h->hang_dev; h->dxp[l].args.hang = h->hang_dev; h->dwp[l].args.hang = h->hang_dev;
        // the weights are written by the update kernel(s) at the END of a step; only the first forward launch follows them directly
        h->fwd[l].args.b_early = (l != 1); h->dxp[l].args.b_early = 1;
    }
    { const char *ev = getenv("GGD_WIDE_PDL"); h->fwd[1].no_pdl = h->wide && ev && atoi(ev) == 0; }   // fwd[1] follows the join with the side streams; PDL on that edge still captures (-9 us per step)
    if (h->fuse_loss) {
        const LayerInfo &top = h->lay[L - 1];
        GemmPlan &p = h->fwd_loss;
        p = h->fwd[L - 1];
        p.epi = EPI_FWD_LOSS;
        GemmArgs &a = p.args;
        a.o_hi = loc(h, h->dx_hi[L - 1], L - 1); a.o_lo = loc(h, h->dx_lo[L - 1], L - 1); a.ldo = top.Np;
        a.D = top.cur; a.Mg = h->Mg; a.beta = h->cfg.shapefactor; a.ml = (h->cfg.MLflag == 1);
        a.alpha = h->alpha; a.loss_trace = h->trace;
        a.world = h->fx_loss ? h->cfg.world_size : 1; a.rank = h->cfg.rank;
        if (h->fx_loss) {
            a.step_counter = h->fx_counters + FXC_STEP; a.error_flag = h->fx_counters + FXC_ERROR;
            for (int q = 0; q < a.world; q++) { a.asum_slot[q] = (float *)h->fx_peer[1][q]; a.lflags[q] = (unsigned int *)h->fx_peer[2][q]; }
        }
    }
    if (h->persist) {
        // one tile list over all layers: (layer, 128-unit n-tile, 64-unit k-tile), k fastest
        DwpArgs a;
        memset(&a, 0, sizeof a);
        int base = 0;
        for (int l = 1; l < L; l++) {
            const LayerInfo &ly = h->lay[l];
            const bool first = (l == 1);
            DwpLayer &d = a.layer[a.nlayers++];
            GGD_TRY(make_tmap_bf16(&d.a_hi, h->dx_hi[l], h->Mp, ly.Np, ly.Np, 64));      // dE/dx, box {64, 64}
            GGD_TRY(make_tmap_bf16(&d.a_lo, h->dx_lo[l], h->Mp, ly.Np, ly.Np, 64));
            GGD_TRY(make_tmap_bf16(&d.b_hi, first ? h->c_hi : h->act_hi[l - 1], first ? (long long)h->cap : h->Mp, ly.Kp, ly.Kp, 64));   // activations below
            GGD_TRY(make_tmap_bf16(&d.b_lo, first ? h->c_lo : h->act_lo[l - 1], first ? (long long)h->cap : h->Mp, ly.Kp, ly.Kp, 64));
            d.W = h->P + ly.w_off; d.D = h->Dl + ly.w_off; d.w_hi = h->Phi + ly.w_off; d.w_lo = h->Plo + ly.w_off;
            d.b = h->P + ly.b_off; d.db = h->Dl + ly.b_off;
            GGD_TRY(make_tmap_2d(&d.w_map, d.W, 1, ly.Kp, ly.Np, ly.Np, 128, 16, 0));
            GGD_TRY(make_tmap_2d(&d.d_map, d.D, 1, ly.Kp, ly.Np, ly.Np, 128, 16, 0));
            GGD_TRY(make_tmap_2d(&d.hi_map, d.w_hi, 0, ly.Kp, ly.Np, ly.Np, 128, 16, 0));
            GGD_TRY(make_tmap_2d(&d.lo_map, d.w_lo, 0, ly.Kp, ly.Np, ly.Np, 128, 16, 0));
            d.dx_hi = h->dx_hi[l]; d.dx_lo = h->dx_lo[l];
            d.Kp = ly.Kp; d.Np = ly.Np; d.N = ly.cur;
            d.k_tiles = ly.Kp / 64;
            d.tile_base = base;
            d.b_rows_from_ctl = first;
            d.wc = h->cfg.weightcost;
            base += ceil_div(ly.Np, 128) * d.k_tiles;
        }
        a.total_tiles = base;
        a.ctl = h->ctl; a.rows_per_bunch = h->M; a.M = h->M;
        a.mom = h->cfg.momentum; a.lr = h->cfg.lrate; a.Mg = (float)h->Mg;
        a.advance = 1; a.done_counter = h->dwp_counter; a.shadows = h->w_f32 ? 0 : 1; a.hang = h->hang_dev;
        { const char *ev = getenv("GGD_L2_HINTS"); a.l2_hints = !(ev && atoi(ev) == 0); }
        GGD_CUDA(cudaMemcpy(h->dwp_dev, &a, sizeof a, cudaMemcpyHostToDevice));
    }
    if (h->wide) {
        // slab list over all layers, TOP layer first (in data-parallel mode its factors arrive first)
        DwwArgs *a = new DwwArgs();
        memset(a, 0, sizeof *a);
        int base = 0, rc = GGD_OK;
        for (int l = L - 1; l >= 1 && rc == GGD_OK; l--) {
            const LayerInfo &ly = h->lay[l];
            const bool in_chunk = (l == 1) && !h->dp_fx;       // layer 1 reads the net input from the chunk arrays on one GPU
            DwwLayer &d = a->layer[a->nlayers++];
            rc = make_tmap_bf16(&d.a_hi, h->dx_hi[l], h->fx_rows, ly.Np, ly.Np, 32);
            if (!rc) rc = make_tmap_bf16(&d.a_lo, h->dx_lo[l], h->fx_rows, ly.Np, ly.Np, 32);
            if (!rc) rc = make_tmap_bf16(&d.b_hi, in_chunk ? h->c_hi : h->act_hi[l - 1], in_chunk ? (long long)h->cap : h->fx_rows, ly.Kp, ly.Kp, 32);
            if (!rc) rc = make_tmap_bf16(&d.b_lo, in_chunk ? h->c_lo : h->act_lo[l - 1], in_chunk ? (long long)h->cap : h->fx_rows, ly.Kp, ly.Kp, 32);
            if (!rc) rc = make_tmap_2d(&d.w_map, h->P + ly.w_off, 1, ly.Kp, ly.Np, ly.Np, 128, 16, 0);
            if (!rc) rc = make_tmap_2d(&d.d_map, h->Dl + ly.w_off, 1, ly.Kp, ly.Np, ly.Np, 128, 16, 0);
            d.Kp = ly.Kp; d.Np = ly.Np; d.k_slabs = ly.Kp / 64; d.slab_base = base; d.n_slabs = ceil_div(ly.Np, 128) * d.k_slabs;
            d.b_rows_from_ctl = in_chunk;
            d.ev_dx = h->dp_fx ? FX_EV_DX + l : -1; d.ev_y = h->dp_fx ? FX_EV_Y + (l - 1) : -1;
            d.wc = h->cfg.weightcost;
            d.dx_hi = h->dx_hi[l]; d.dx_lo = h->dx_lo[l]; d.b = h->P + ly.b_off; d.db = h->Dl + ly.b_off; d.N = ly.cur;
            base += ceil_div(ly.Np, 128) * d.k_slabs;
        }
        a->total_slabs = base;
        // slab groups: one GPU = one group; data parallel = {all layers above the bottom one}, {bottom layer} (its factors arrive last)
        { const char *ev = getenv("GGD_WIDE_GROUPS");     // tuning: 0 = one group, 1 = one group per layer
          const int mode = ev ? atoi(ev) : -1;
          a->ngroups = 0;
          a->group_base[a->ngroups++] = 0;
          for (int i = 1; i < a->nlayers; i++) {
              // (at 1024+ frames one contiguous range is better: fewer, wider segments outweigh the wait for the bottom layer)
              const bool cut = mode == 1 || (mode == -1 && h->dp_fx && h->fx_rows < 1024 && i == a->nlayers - 1);
              if (cut) a->group_base[a->ngroups++] = a->layer[i].slab_base;
          }
          a->group_base[a->ngroups] = base; }
        a->ctl = h->ctl; a->rows_per_bunch = h->M; a->fblocks = h->fx_rows / 32; a->rows = h->fx_rows;
        h->wide_smem = dw_wide_smem(a->fblocks, &a->op_stages, &a->wd_stages);
        a->mom = h->cfg.momentum; a->lr = h->cfg.lrate; a->Mg = (float)h->Mg;
        a->advance = 1; a->done_counter = h->dwp_counter; a->hang = h->hang_dev;
        { const char *ev = getenv("GGD_L2_HINTS"); a->l2_hints = !(ev && atoi(ev) == 0); }
        a->world = world; a->rank = rank;
        if (h->dp_fx) {
            a->flags = h->fx_flags; a->step_counter = h->fx_counters + FXC_STEP; a->error_flag = h->fx_counters + FXC_ERROR;
            for (int p = 0; p < world; p++) a->peer_flags[p] = (unsigned int *)h->fx_peer[2][p];
        }
        a->trace = h->fx_trace;
        cudaError_t e = (rc == GGD_OK) ? cudaMemcpy(h->dww_dev, a, sizeof *a, cudaMemcpyHostToDevice) : cudaSuccess;
        delete a;
        GGD_TRY(rc);
        GGD_CUDA(e);
    }
    if (h->dp_fx) {
        // one push per factor array: my slice -> the same rows of every peer's arena
        auto make_push = [&](int event, bf16 *hi, bf16 *lo, int l, const bf16 *src_hi, const bf16 *src_lo, long long bunch_stride, int first_of_step) {
            FxPushArgs &p = h->fx_push[event];
            memset(&p, 0, sizeof p);
            const long long bytes = (long long)h->Mp * h->upad[l] * sizeof(bf16);
            p.seg[0] = {(const uint8_t *)src_hi, bunch_stride, (long long)((uint8_t *)loc(h, hi, l) - h->fx_arena), bytes};
            p.seg[1] = {(const uint8_t *)src_lo, bunch_stride, (long long)((uint8_t *)loc(h, lo, l) - h->fx_arena), bytes};
            p.nseg = 2;
            for (int q = 0; q < world; q++) { p.peer_arena[q] = (uint8_t *)h->fx_peer[0][q]; p.peer_flags[q] = (unsigned int *)h->fx_peer[2][q]; }
            p.my_flags = h->fx_flags; p.ctl = h->ctl; p.step_counter = h->fx_counters + FXC_STEP;
            p.block_counter = h->fx_counters + FXC_PUSH_BLOCKS + event; p.error_flag = h->fx_counters + FXC_ERROR; p.hang = h->hang_dev;
            p.world = world; p.rank = rank; p.event = event; p.wait_done = 1; p.include_self = first_of_step;   // (the pushes of a step run in parallel: each checks DONE)
            p.trace = h->fx_trace;
        };
        // the net-input rows of the current bunch live in the chunk arrays: they are copied into EVERY arena (mine included)
        make_push(FX_EV_Y + 0, h->act_hi[0], h->act_lo[0], 0, h->c_hi, h->c_lo, (long long)h->M * h->upad[0] * sizeof(bf16), 1);
        for (int l = 1; l < L - 1; l++) make_push(FX_EV_Y + l, h->act_hi[l], h->act_lo[l], l, loc(h, h->act_hi[l], l), loc(h, h->act_lo[l], l), 0, 0);
        for (int l = 1; l < L; l++) make_push(FX_EV_DX + l, h->dx_hi[l], h->dx_lo[l], l, loc(h, h->dx_hi[l], l), loc(h, h->dx_lo[l], l), 0, 0);
    }
    return GGD_OK;
}

static int ensure_chunk(ggd_handle *h, size_t frames)
{
    if (frames <= h->cap) return GGD_OK;
    free_chunk(h);
    size_t cap = frames + h->Mp;   // slack so that a padded last tile never leaves the allocation
    const int D = h->units[h->L - 1];
    GGD_CUDA(cudaMalloc(&h->c_in, cap * h->units[0] * sizeof(float)));
    GGD_CUDA(cudaMalloc(&h->c_targ, cap * D * sizeof(float)));
    GGD_CUDA(cudaMalloc(&h->c_out, cap * D * sizeof(float)));
    GGD_CUDA(cudaMemset(h->c_in, 0, cap * h->units[0] * sizeof(float)));
    GGD_CUDA(cudaMemset(h->c_targ, 0, cap * D * sizeof(float)));
    if (h->tensor) {
        GGD_CUDA(cudaMalloc(&h->c_hi, cap * h->upad[0] * sizeof(bf16)));
        GGD_CUDA(cudaMalloc(&h->c_lo, cap * h->upad[0] * sizeof(bf16)));
        GGD_CUDA(cudaMemset(h->c_hi, 0, cap * h->upad[0] * sizeof(bf16)));
        GGD_CUDA(cudaMemset(h->c_lo, 0, cap * h->upad[0] * sizeof(bf16)));
    }
    h->trace_cap = cap / h->M + 2;
    GGD_CUDA(cudaMalloc(&h->trace, h->trace_cap * sizeof(double)));
    h->cap = cap;
    if (h->tensor) GGD_TRY(build_plans(h));
    return GGD_OK;
}

// ---- peer memory: exchange CUDA IPC handles through NCCL, map every peer's buffers ---------------------------------
static int ipc_exchange(ggd_handle *h, void *const *local, int nbuf, void *(*peer)[FX_MAX])
{
    const int world = h->cfg.world_size, rank = h->cfg.rank;
    std::vector<cudaIpcMemHandle_t> mine(nbuf), all((size_t)nbuf * world);
    for (int k = 0; k < nbuf; k++) GGD_CUDA(cudaIpcGetMemHandle(&mine[k], local[k]));
    cudaIpcMemHandle_t *d_all = nullptr, *d_mine = nullptr;
    const size_t bytes = sizeof(cudaIpcMemHandle_t) * nbuf;
    GGD_CUDA(cudaMalloc(&d_all, bytes * world));
    GGD_CUDA(cudaMalloc(&d_mine, bytes));
    GGD_CUDA(cudaMemcpy(d_mine, mine.data(), bytes, cudaMemcpyHostToDevice));
    GGD_NCCL(ncclAllGather(d_mine, d_all, bytes, ncclChar, h->comm, h->s_main));
    GGD_CUDA(cudaStreamSynchronize(h->s_main));
    GGD_CUDA(cudaMemcpy(all.data(), d_all, bytes * world, cudaMemcpyDeviceToHost));
    cudaFree(d_all); cudaFree(d_mine);
    for (int p = 0; p < world; p++)
        for (int k = 0; k < nbuf; k++) {
            if (p == rank) { peer[k][p] = local[k]; continue; }
            cudaError_t e = cudaIpcOpenMemHandle(&peer[k][p], all[(size_t)p * nbuf + k], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle(rank %d, buffer %d): %s", p, k, cudaGetErrorString(e)); return GGD_ECUDA; }
        }
    return GGD_OK;
}

// host-side cross-rank barrier (NCCL allreduce of one float + stream sync)
static int dp_barrier(ggd_handle *h)
{
    GGD_NCCL(ncclAllReduce(h->colsum, h->colsum, 1, ncclFloat, ncclSum, h->comm, h->s_main));
    GGD_CUDA(cudaStreamSynchronize(h->s_main));
    return GGD_OK;
}

// factor-exchange data parallelism: receive slots + flags, IPC exchange of the arenas
static int dp_fx_setup(ggd_handle *h)
{
    const int world = h->cfg.world_size, D = h->units[h->L - 1];
    const size_t ab = (size_t)world * D;
    GGD_CUDA(cudaMalloc(&h->fx_asum, ab * sizeof(float)));   GGD_CUDA(cudaMemset(h->fx_asum, 0, ab * sizeof(float)));
    GGD_CUDA(cudaMalloc(&h->fx_flags, (size_t)FX_MAX * FX_STRIDE * sizeof(unsigned int)));
    GGD_CUDA(cudaMemset(h->fx_flags, 0, (size_t)FX_MAX * FX_STRIDE * sizeof(unsigned int)));
    void *local[3] = {h->fx_arena, h->fx_asum, h->fx_flags};
    GGD_TRY(ipc_exchange(h, local, 3, h->fx_peer));
    { const char *ev = getenv("GGD_FX_PUSH_CTAS"); h->fx_push_ctas = (ev && atoi(ev) > 0) ? atoi(ev) : 32; }
    // (32 CTAs at every N: the push CTAs' peer stores share the SMs' store paths with the GEMM epilogues -- at N = 8, 56 CTAs per
    //  push slowed every GEMM of the chain from ~15 to ~21 us: 3.79 -> 4.17 M frames/s with 32)
    GGD_TRY(dp_barrier(h));    // nobody may enter the first step before every rank has mapped everyone
    return GGD_OK;
}

// ---- one training step (forward, loss gradient, backward, update); stream-ordered, no host sync ----
// Side branches of a step.  Every factor push (and the bias update) runs on its OWN side stream, forked from the compute
// stream right after its producer: in the captured graph these are parallel branches, so neither the launch gaps nor the
// NVLink round trips of one push delay the next one (8 pushes in a row on one stream took 125 us against a 75 us chain).
struct SideCtx { int nfork, nside; };
static int fx_fork(ggd_handle *h, cudaStream_t s, SideCtx *sc, cudaStream_t *side)
{
    cudaEvent_t e = h->ev_fx[sc->nfork++ % FX_NEVENTS];
    *side = h->s_side[sc->nside++ % FX_NSIDE];
    GGD_CUDA(cudaEventRecord(e, s));
    GGD_CUDA(cudaStreamWaitEvent(*side, e, 0));
    return GGD_OK;
}
// all side branches of this step join the compute stream
static int fx_join(ggd_handle *h, cudaStream_t s, SideCtx *sc)
{
    const int used = sc->nside < FX_NSIDE ? sc->nside : FX_NSIDE;
    for (int i = 0; i < used; i++) {
        cudaEvent_t e = h->ev_fx[sc->nfork++ % FX_NEVENTS];
        GGD_CUDA(cudaEventRecord(e, h->s_side[i]));
        GGD_CUDA(cudaStreamWaitEvent(s, e, 0));
    }
    sc->nside = 0;
    return GGD_OK;
}
// factor push of one array on a side stream, after everything queued on `s` so far
static int fx_push(ggd_handle *h, cudaStream_t s, int event, SideCtx *sc, int *launches)
{
    cudaStream_t side;
    GGD_TRY(fx_fork(h, s, sc, &side));
    ProfScope ps(h, KC_PUSH, side);
    // the LAST push of a step (dE/dx of the bottom layer) runs alone after the backward chain and is on the critical path: more CTAs
    const int grid = (event == FX_EV_DX + 1) ? std::min(h->sm_count, 2 * h->fx_push_ctas) : h->fx_push_ctas;
    launch_factor_push(h->fx_push[event], grid, side); (*launches)++;
    return GGD_OK;
}

static int enqueue_forward(ggd_handle *h, cudaStream_t s, int *launches, bool train = false, bool fx = false, SideCtx *sc = nullptr)
{
    const int L = h->L;
    if (h->tensor) {
        for (int l = 1; l < L; l++) {
            { ProfScope ps(h, KC_FWD, s);
              GGD_TRY(launch_gemm_tc((train && h->fuse_loss && l == L - 1) ? h->fwd_loss : h->fwd[l], s)); (*launches)++; }
            if (fx && l < L - 1) GGD_TRY(fx_push(h, s, FX_EV_Y + l, sc, launches));
        }
    } else {
        launch_simt_gather_in(h->ctl, h->M, h->units[0], h->y32[0], h->upad[0], s); (*launches)++;
        for (int l = 1; l < L; l++) {
            const LayerInfo &ly = h->lay[l];
            launch_simt_gemm(h->y32[l - 1], ly.Kp, 1, h->P + ly.w_off, 1, ly.Np, h->x32[l], ly.Np, h->M, ly.cur, ly.prev, s);
            launch_simt_bias_act(h->x32[l], ly.Np, h->P + ly.b_off, (l == L - 1) ? h->out32 : h->y32[l], h->M, ly.cur, l == L - 1, s);
            (*launches) += 2;
        }
    }
    return GGD_OK;
}

static int enqueue_step(ggd_handle *h, cudaStream_t s, bool apply_update, int *launches, bool allow_fused = true)
{
    const bool fused = h->fused && allow_fused && apply_update;
    const bool fx = h->dp_fx && fused;      // factor exchange: pushes on the side stream, replicated update
    const int L = h->L;
    const LayerInfo &top = h->lay[L - 1];
    SideCtx sc = {0, 0};
    if (h->dp_fx && !fused) { set_error("data-parallel steps need the fused update path"); return GGD_EUNSUPPORTED; }
    if (fx) GGD_TRY(fx_push(h, s, FX_EV_Y + 0, &sc, launches));    // net-input rows of this bunch -> every arena
    GGD_TRY(enqueue_forward(h, s, launches, true, fx, &sc));
    // ---- fused loss gradient (BP_GPU.cu:408-424); with fuse_loss it already ran in the output layer's epilogue
    LossArgs la;
    memset(&la, 0, sizeof la);
    la.ctl = h->ctl; la.out = h->out32; la.ldo = top.Np; la.M = h->M; la.Mg = h->Mg; la.D = top.cur;
    la.beta = h->cfg.shapefactor; la.ml = (h->cfg.MLflag == 1);
    la.dx32 = h->tensor ? nullptr : h->dx32[L - 1];
    la.dx_hi = h->tensor ? loc(h, h->dx_hi[L - 1], L - 1) : nullptr; la.dx_lo = h->tensor ? loc(h, h->dx_lo[L - 1], L - 1) : nullptr;
    la.ldx = top.Np; la.alpha = h->alpha; la.colsum = h->colsum; la.trace = h->trace;
    if (h->fuse_loss) {
    } else if (h->fx_loss) {
        // partial sum|e|^beta exchanged over peer memory inside the loss kernel (no NCCL on the step)
        ProfScope ps(h, KC_LOSS, s);
        la.mode = 3; la.world = h->cfg.world_size; la.rank = h->cfg.rank;
        la.step_counter = h->fx_counters + FXC_STEP; la.error_flag = h->fx_counters + FXC_ERROR;
        for (int p = 0; p < la.world; p++) { la.asum_slot[p] = (float *)h->fx_peer[1][p]; la.lflags[p] = (unsigned int *)h->fx_peer[2][p]; }
        launch_loss(la, s); (*launches)++;
    } else if (h->has_comm) {
        { ProfScope ps(h, KC_LOSS, s); la.mode = 1; launch_loss(la, s); }
        { ProfScope ps(h, KC_ALLREDUCE, s); GGD_NCCL(ncclAllReduce(h->colsum, h->colsum, top.cur, ncclFloat, ncclSum, h->comm, s)); }
        { ProfScope ps(h, KC_LOSS, s); la.mode = 2; launch_loss(la, s); }
        (*launches) += 3;
    } else {
        ProfScope ps(h, KC_LOSS, s);
        la.mode = 0; launch_loss(la, s); (*launches)++;
    }
    if (fx) GGD_TRY(fx_push(h, s, FX_EV_DX + (L - 1), &sc, launches));
    // ---- backward (BP_GPU.cu:371-438); every GEMM of the step sees the pre-update weights
    for (int l = L - 1; l > 0; l--) {
        const LayerInfo &ly = h->lay[l];
        if (h->tensor) {
            if (l != 1) {
                { ProfScope ps(h, KC_DX, s); GGD_TRY(launch_gemm_tc(h->dxp[l], s)); (*launches)++; }
                if (fx) GGD_TRY(fx_push(h, s, FX_EV_DX + (l - 1), &sc, launches));
            }
            if (!fused) { ProfScope ps(h, KC_DW, s); GGD_TRY(launch_gemm_tc(h->dwp[l], s)); (*launches)++; }
            // fused: all layers in one persistent launch after the backward chain
        } else {
            if (l != L - 1) { launch_simt_dsigmoid(h->y32[l], h->dy32[l], h->dx32[l], ly.Np, h->M, ly.cur, s); (*launches)++; }
            if (l != 1) {
                launch_simt_gemm(h->dx32[l], ly.Np, 1, h->P + ly.w_off, ly.Np, 1, h->dy32[l - 1], ly.Kp, h->M, ly.prev, ly.cur, s);
                (*launches)++;
            }
            launch_simt_gemm(h->y32[l - 1], 1, ly.Kp, h->dx32[l], 1, ly.Np, h->G + ly.w_off, ly.Np, ly.prev, ly.cur, h->M, s);
            (*launches)++;
        }
    }
    if (fused && h->persist) {
        // weight gradients + updates of all layers, bias gradients + updates and the bunch counter: one launch
        ProfScope ps(h, KC_DWUPD, s);
        GGD_TRY(launch_dw_persist(h->dwp_dev, h->sm_count, h->w_f32 ? 0 : 1, s)); (*launches)++;
        GGD_CUDA(cudaGetLastError());
        return GGD_OK;
    }
    if (fused && h->wide) {
        // weights, biases and the bunch counter of the whole (global) minibatch: one persistent launch
        { ProfScope ps(h, KC_DWUPD, s); GGD_TRY(launch_dw_wide(h->dww_dev, h->sm_count, h->wide_smem, s)); (*launches)++; }
        GGD_TRY(fx_join(h, s, &sc));
        GGD_CUDA(cudaGetLastError());
        return GGD_OK;
    }
    {
        ProfScope ps(h, KC_BIAS, s);
        BiasGradArgs ba;
        memset(&ba, 0, sizeof ba);
        ba.M = h->M;
        for (int l = 1; l < L; l++) {
            const LayerInfo &ly = h->lay[l];
            BiasGradLayer &b = ba.layer[ba.nlayers++];
            b.dx32 = h->tensor ? nullptr : h->dx32[l];
            b.hi = h->tensor ? loc(h, h->dx_hi[l], l) : nullptr; b.lo = h->tensor ? loc(h, h->dx_lo[l], l) : nullptr;
            b.ld = ly.Np; b.N = ly.cur; b.dst = h->G + ly.gb_off;
            b.b = h->P + ly.b_off; b.db = h->Dl + ly.b_off;
        }
        launch_bias_grad(ba, s); (*launches)++;
    }
    if (h->has_comm) {
        // NCCL mode (fp32 validation path, GGD_DP_MODE=nccl): one allreduce of the whole gradient arena
        ProfScope ps(h, KC_ALLREDUCE, s);
        GGD_NCCL(ncclAllReduce(h->G, h->G, h->arena + h->nbias, ncclFloat, ncclSum, h->comm, s)); (*launches)++;
    }
    if (apply_update) {
        UpdArgs ua;
        memset(&ua, 0, sizeof ua);
        ua.P = h->P; ua.Dl = h->Dl; ua.G = h->G; ua.Phi = h->Phi; ua.Plo = h->Plo;
        ua.mom = h->cfg.momentum; ua.lr = h->cfg.lrate; ua.Mg = (float)h->Mg;
        for (int l = 1; l < L; l++) {
            const LayerInfo &ly = h->lay[l];
            ua.seg[ua.nseg++] = {(long long)ly.w_off, (long long)ly.w_off, (long long)ly.Kp * ly.Np, h->cfg.weightcost, 1};
            ua.seg[ua.nseg++] = {(long long)ly.b_off, (long long)ly.gb_off, (long long)ly.Np, 0.0f, 0};   // no decay on biases (BP_GPU.cu:435)
        }
        ua.ctl = h->ctl;   // the update kernel also advances the bunch counter
        ProfScope ps(h, KC_UPDATE, s);
        launch_update(ua, h->sm_count, s); (*launches)++;
    } else {
        ProfScope ps(h, KC_ADVANCE, s); launch_advance(h->ctl, s); (*launches)++;
    }
    GGD_CUDA(cudaGetLastError());
    return GGD_OK;
}

static int capture_graph(ggd_handle *h, int steps, cudaGraphExec_t *out)
{
    cudaGraph_t g;
    int launches = 0;
    GGD_CUDA(cudaStreamBeginCapture(h->s_main, cudaStreamCaptureModeThreadLocal));
    int rc = GGD_OK;
    for (int i = 0; i < steps && rc == GGD_OK; i++) rc = enqueue_step(h, h->s_main, true, &launches);
    cudaError_t e = cudaStreamEndCapture(h->s_main, &g);
    if (rc != GGD_OK) { if (e == cudaSuccess) cudaGraphDestroy(g); return rc; }
    GGD_CUDA(e);
    GGD_CUDA(cudaGraphInstantiate(out, g, 0));
    cudaGraphDestroy(g);
    h->launches_per_step = launches / steps;
    return GGD_OK;
}

static int set_ctl(ggd_handle *h, const float *d_in, const float *d_targ)
{
    StepCtl c;
    c.bunch_idx = 0; c.pad = 0; c.in32 = d_in; c.targ = d_targ;
    GGD_CUDA(cudaMemcpyAsync(h->ctl, &c, sizeof c, cudaMemcpyHostToDevice, h->s_main));
    return GGD_OK;
}

static int pin_host(ggd_handle *h, const float *src, size_t bytes)
{
    // pin the caller's (reused) buffer once so that the copy is a real DMA (the reference hands us pageable memory)
    if (!(h->cfg.flags & GGD_FLAG_PIN_HOST) || bytes < (1u << 20)) return GGD_OK;
    auto it = h->pinned.find(src);
    if (it != h->pinned.end() && (it->second == 0 || it->second >= bytes)) return GGD_OK;   // pinned large enough (or unpinnable)
    // a reused buffer may carry a LARGER chunk than the one it was first seen with (the chunks of an epoch differ in size
    // and come in shuffled order): a copy that spans registered and unregistered pages is an error, so re-register
    if (it != h->pinned.end()) cudaHostUnregister(const_cast<float *>(src));
    cudaError_t e = cudaHostRegister(const_cast<float *>(src), bytes, cudaHostRegisterDefault);
    if (e == cudaSuccess) h->pinned[src] = bytes;
    else { cudaGetLastError(); h->pinned[src] = 0; }
    return GGD_OK;
}

static int upload(ggd_handle *h, const float *src, float *dst, size_t bytes)
{
    GGD_TRY(pin_host(h, src, bytes));
    GGD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->s_main));
    return GGD_OK;
}

// synchronise the main stream; a device-side watchdog trap is reported with its record (which barrier, block, iteration)
static int sync_main(ggd_handle *h)
{
    static_assert(HANG_WORDS >= 8 + 160 * 12 * 4, "hang record");
    cudaError_t e = cudaStreamSynchronize(h->s_main);
    if (e == cudaSuccess) return GGD_OK;
    const unsigned int *r = h->hang_host;
    if (r && r[0] == 0xDEADu)
    {
        char buf[700]; int n = 0;
        n += snprintf(buf + n, sizeof buf - n, "device pipeline watchdog: wait code %u gave up in block %u at iteration %u (parity %u, thread %u); waiters of that block [warp:code@it/parity]:", r[1], r[2], r[3], r[4], r[5]);
        for (int w = 0; w < 12 && n < 600; w++) {
            const unsigned int *q = r + 8 + (r[2] * 12 + w) * 4;
            if (q[3]) n += snprintf(buf + n, sizeof buf - n, " %d:%u@%u/%u[%u]", w, q[0], q[1], q[2], q[3]);
        }
        int stuck = 0;
        for (int b = 0; b < 160; b++) { bool any = false; for (int w = 0; w < 12; w++) any |= r[8 + (b * 12 + w) * 4 + 3] != 0; stuck += any; }
        n += snprintf(buf + n, sizeof buf - n, "; blocks with waiters: %d", stuck);
        set_error("%s: %s", buf, cudaGetErrorString(e));
    }
    else
        set_error("cudaStreamSynchronize -> %s", cudaGetErrorString(e));
    return GGD_ECUDA;
}

// Steps over one chunk.  With host_in/host_targ the chunk is uploaded in pieces of PIECE bunches on the copy stream
// while the compute stream already trains on the pieces that have landed (H2D hidden behind the steps).
static int run_chunk(ggd_handle *h, int n_frames, const float *d_in, const float *d_targ, const float *host_in = nullptr,
                     const float *host_targ = nullptr, bool presplit = false)
{
    const int nb = n_frames / h->M;   // trailing partial bunch dropped (BP_GPU.cu:173-180)
    const int PIECE = 16;             // bunches per upload piece (= the 16-step graph): 17 MB, 0.7 ms of PCIe under 1.8 ms of steps
    // The first pieces are smaller (1, 1, 2, 4, 8 bunches) so that the first step waits for 1 MB, not 17 MB: a short chunk
    // (bench --steps 20) otherwise spends a quarter of its time waiting for the first piece.
    std::vector<int> pb;              // piece boundaries in bunches: piece p = bunches [pb[p], pb[p+1])
    pb.push_back(0);
    if (host_in != nullptr) {
        for (int len = 1, first = 1; pb.back() < nb;) {
            pb.push_back(std::min(nb, pb.back() + len));
            if (first) first = 0; else if (len < PIECE) len *= 2;
        }
    } else pb.push_back(nb);
    const bool piped = host_in != nullptr;
    const int D_ = h->units[h->L - 1];
    h->stats.steps = nb; h->stats.launches = 0;
    h->losses.assign(nb, 0.0f);
    if (nb == 0) { h->stats.device_ms = 0; return GGD_OK; }
    // data parallel: every rank enters the chunk together (each rank feeds its own loader / disk; the in-kernel peer waits
    // are bounded, so skew between ranks must be absorbed HERE, on the host, where waiting is free)
    if (h->has_comm) GGD_TRY(dp_barrier(h));
    GGD_TRY(set_ctl(h, d_in, d_targ));
    GGD_CUDA(cudaMemsetAsync(h->trace, 0, h->trace_cap * sizeof(double), h->s_main));
    GGD_CUDA(cudaEventRecord(h->ev1, h->s_main));
    const bool graphs = !(h->cfg.flags & GGD_FLAG_NO_GRAPH);
    if (graphs && !h->g1) {
        GGD_TRY(capture_graph(h, 1, &h->g1));
        GGD_TRY(capture_graph(h, 4, &h->g4));
        h->gN_steps = 16;
        GGD_TRY(capture_graph(h, h->gN_steps, &h->gN));
    }
    const int npieces = (int)pb.size() - 1;
    if (piped) {
        while ((int)h->ev_piece.size() < npieces) { cudaEvent_t e; GGD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->ev_piece.push_back(e); }
        // the copy stream starts after everything queued so far on the compute stream (the previous chunk is done: calls are blocking)
        GGD_CUDA(cudaEventRecord(h->ev_c0, h->s_copy));
        for (int p = 0; p < npieces; p++) {
            const size_t f0 = (size_t)pb[p] * h->M, f1 = (p == npieces - 1) ? (size_t)n_frames : (size_t)pb[p + 1] * h->M;
            GGD_CUDA(cudaMemcpyAsync(const_cast<float *>(d_in) + f0 * h->units[0], host_in + f0 * h->units[0], (f1 - f0) * h->units[0] * sizeof(float), cudaMemcpyHostToDevice, h->s_copy));
            GGD_CUDA(cudaMemcpyAsync(const_cast<float *>(d_targ) + f0 * D_, host_targ + f0 * D_, (f1 - f0) * D_ * sizeof(float), cudaMemcpyHostToDevice, h->s_copy));
            GGD_CUDA(cudaEventRecord(h->ev_piece[p], h->s_copy));
        }
        GGD_CUDA(cudaEventRecord(h->ev_c1, h->s_copy));
    }
    int launches = 0;
    for (int p = 0; p < npieces; p++) {
        const int b0 = pb[p], b1 = pb[p + 1];
        if (piped) GGD_CUDA(cudaStreamWaitEvent(h->s_main, h->ev_piece[p], 0));
        if (h->tensor && !presplit) {
            launch_split_rows(d_in + (size_t)b0 * h->M * h->units[0], (b1 - b0) * h->M, h->units[0], h->c_hi + (size_t)b0 * h->M * h->upad[0],
                              h->c_lo + (size_t)b0 * h->M * h->upad[0], h->upad[0], h->s_main);
            h->stats.launches++;
        }
        if (!graphs) {
            for (int b = b0; b < b1; b++) GGD_TRY(enqueue_step(h, h->s_main, true, &launches));
        } else {
            int b = b0;
            for (; b + h->gN_steps <= b1; b += h->gN_steps) GGD_CUDA(cudaGraphLaunch(h->gN, h->s_main));
            for (; b + 4 <= b1; b += 4) GGD_CUDA(cudaGraphLaunch(h->g4, h->s_main));
            for (; b < b1; b++) GGD_CUDA(cudaGraphLaunch(h->g1, h->s_main));
        }
    }
    h->stats.launches += graphs ? (long long)nb * h->launches_per_step : launches;
    GGD_CUDA(cudaEventRecord(h->ev2, h->s_main));
    std::vector<double> tr(nb);
    if (cudaMemcpyAsync(tr.data(), h->trace, nb * sizeof(double), cudaMemcpyDeviceToHost, h->s_main) != cudaSuccess) cudaGetLastError();
    GGD_TRY(sync_main(h));
    float ms = 0;
    GGD_CUDA(cudaEventElapsedTime(&ms, h->ev1, h->ev2));
    h->stats.device_ms = ms;
    for (int b = 0; b < nb; b++) h->losses[b] = (float)tr[b];
    h->stats.d2h_bytes = nb * sizeof(double);
    if (h->fx_trace) {
        // per-kernel globaltimer stamps of the LAST step (tuning aid, GGD_FX_TRACE=1), in us relative to the earliest stamp
        std::vector<unsigned long long> tr(FX_TRACE_WORDS);
        GGD_CUDA(cudaMemcpy(tr.data(), h->fx_trace, tr.size() * 8, cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull;
        for (unsigned long long x : tr) if (x && x < t0) t0 = x;
        auto us = [&](int i) { return tr[i] ? (double)(tr[i] - t0) * 1e-3 : -1.0; };
        std::string line = "[rank " + std::to_string(h->cfg.rank) + "] fx trace (us): ";
        char buf[160];
        for (int ev = 0; ev < FX_STRIDE; ev++)
            if (tr[ev * 4]) { snprintf(buf, sizeof buf, "push%d[start %.1f waited %.1f copied %.1f flagged %.1f] ", ev, us(ev * 4), us(ev * 4 + 1), us(ev * 4 + 2), us(ev * 4 + 3)); line += buf; }
        snprintf(buf, sizeof buf, "wide[start %.1f pdl %.1f", us(FX_TRACE_WIDE), us(FX_TRACE_WIDE + 1)); line += buf;
        for (int l = 0; l < 10; l++) if (tr[FX_TRACE_WIDE + 2 + l]) { snprintf(buf, sizeof buf, " ready%d %.1f", l, us(FX_TRACE_WIDE + 2 + l)); line += buf; }
        snprintf(buf, sizeof buf, " end %.1f]", us(FX_TRACE_WIDE + 14)); line += buf;
        fprintf(stderr, "%s\n", line.c_str());
        GGD_CUDA(cudaMemset(h->fx_trace, 0, FX_TRACE_WORDS * 8));
    }
    if (h->dp_fx) {
        unsigned int err = 0;
        GGD_CUDA(cudaMemcpy(&err, h->fx_counters + FXC_ERROR, sizeof err, cudaMemcpyDeviceToHost));
        if (err) { set_error("data-parallel step: rank %u did not arrive within the timeout (ranks must train the same number of bunches)", err - 1); return GGD_ENCCL; }
    }
    return GGD_OK;
}

template <typename T>
static int grow(T **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return GGD_OK;
    cudaFree(*buf); *buf = nullptr; *cap = 0;
    GGD_CUDA(cudaMalloc(buf, need * sizeof(T)));
    *cap = need;
    return GGD_OK;
}

// ---------------------------------------------------------------------------------------------------
extern "C" {

const char *ggd_last_error(void) { return g_err; }
const char *ggd_version(void) { return "ggd_b200 0.1 (sm_100a; tcgen05 bf16x3)"; }

int ggd_create(const ggd_config *cfg, const float *const *weights, const float *const *bias, ggd_handle **out)
{
    if (!cfg || !weights || !bias || !out) { set_error("ggd_create: null argument"); return GGD_EINVAL; }
    if (cfg->numlayers < 2 || cfg->numlayers > GGD_MAXLAYER) { set_error("numlayers %d out of range", cfg->numlayers); return GGD_EINVAL; }
    if (cfg->bunchsize < 1) { set_error("bunchsize must be >= 1"); return GGD_EINVAL; }
    if (cfg->dropoutflag == 1) { set_error("dropoutflag=1 is outside the named path and not supported"); return GGD_EUNSUPPORTED; }
    if (!(cfg->shapefactor > 0)) { set_error("shapefactor must be > 0"); return GGD_EINVAL; }
    for (int i = 0; i < cfg->numlayers; i++)
        if (cfg->layersizes[i] < 1) { set_error("layersizes[%d] invalid", i); return GGD_EINVAL; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("no CUDA device: libggd_b200 has no CPU fallback"); return GGD_ECUDA; }
    if (cfg->gpu < 0 || cfg->gpu >= ndev) { set_error("GPU Num %d Not In Range 0-%d", cfg->gpu, ndev - 1); return GGD_EINVAL; }
    GGD_CUDA(cudaSetDevice(cfg->gpu));
    cudaDeviceProp prop;
    GGD_CUDA(cudaGetDeviceProperties(&prop, cfg->gpu));
    if (prop.major != 10) { set_error("device %d is sm_%d%d; this library is built for sm_100a only", cfg->gpu, prop.major, prop.minor); return GGD_ECUDA; }

    ggd_handle *h = new ggd_handle();
    h->cfg = *cfg;
    h->L = cfg->numlayers; h->M = cfg->bunchsize; h->Mp = round_up(h->M, 128);
    h->sm_count = prop.multiProcessorCount;
    h->tensor = (cfg->precision == GGD_PREC_BF16X3);
    const int world = cfg->world_size > 1 ? cfg->world_size : 1;
    {
        // data parallelism: factor exchange over NVLink peer memory (default, tensor path) or NCCL allreduce of the gradient
        // arena (GGD_DP_MODE=nccl, the fp32 validation path, GGD_FLAG_UNFUSED_UPDATE)
        const char *dm = getenv("GGD_DP_MODE");
        const bool want_nccl = dm && !strcmp(dm, "nccl");
        h->dp_fx = world > 1 && world <= FX_MAX && h->tensor && !want_nccl && !(cfg->flags & GGD_FLAG_UNFUSED_UPDATE);
        h->fx_loss = h->dp_fx && cfg->layersizes[cfg->numlayers - 1] <= 16 * LOSS_FLAGS_PER_RANK;   // one flag per 16-column chunk
    }
    h->fused = h->tensor && !(cfg->flags & GGD_FLAG_UNFUSED_UPDATE) && (world == 1 || h->dp_fx);
    {
        const char *ev = getenv("GGD_DW_PERSIST");   // 0: the wide kernel also at 128 frames (tuning / A-B only)
        h->persist = h->fused && world == 1 && h->Mp == 128 && !(ev && atoi(ev) == 0);
        h->wide = h->fused && !h->persist;
    }
    {
        // GGD_W_F32: 1 (default) = GEMMs split the fp32 master weights in-kernel (no shadows; update kernel 57 -> 44 us, GEMMs
        // +1.3 us each: 124.7 -> 121.9 us per step on one GPU), 0 = bf16 hi/lo shadows maintained by the update kernels
        // (dw_persist / update_kernel only: the wide kernel keeps no shadows)
        const char *ev = getenv("GGD_W_F32");
        h->w_f32 = h->tensor && ((ev ? atoi(ev) != 0 : true) || h->wide);
    }
    h->Mg = h->M * world;
    h->fx_rows = h->dp_fx ? world * h->Mp : h->Mp;
    h->loc_row = h->dp_fx ? cfg->rank * h->Mp : 0;
    size_t off = 0;
    for (int i = 0; i < h->L; i++) { h->units[i] = cfg->layersizes[i]; h->upad[i] = round_up(h->units[i], 64); }
    for (int l = 1; l < h->L; l++) {
        LayerInfo &ly = h->lay[l];
        ly.prev = h->units[l - 1]; ly.cur = h->units[l]; ly.Kp = h->upad[l - 1]; ly.Np = h->upad[l];
        ly.w_off = off; off += (size_t)ly.Kp * ly.Np;
        ly.b_off = off; off += ly.Np;
    }
    h->arena = off;
    h->nbias = 0;
    for (int l = 1; l < h->L; l++) { h->lay[l].gb_off = off + h->nbias; h->nbias += h->lay[l].Np; }
    auto fail = [&](int rc) { ggd_destroy(h); return rc; };
#define CK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("%s -> %s", #expr, cudaGetErrorString(_e)); return fail(_e == cudaErrorMemoryAllocation ? GGD_ENOMEM : GGD_ECUDA); } } while (0)
    CK(cudaStreamCreateWithFlags(&h->s_main, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h->ev_c0)); CK(cudaEventCreate(&h->ev_c1));
    { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi); for (int k = 0; k < FX_NSIDE; k++) CK(cudaStreamCreateWithPriority(&h->s_side[k], cudaStreamNonBlocking, hi)); }
    for (int k = 0; k < FX_NEVENTS; k++) CK(cudaEventCreateWithFlags(&h->ev_fx[k], cudaEventDisableTiming));
    CK(cudaEventCreate(&h->ev0)); CK(cudaEventCreate(&h->ev1)); CK(cudaEventCreate(&h->ev2));
    CK(cudaMalloc(&h->P, off * sizeof(float))); CK(cudaMalloc(&h->Dl, off * sizeof(float))); CK(cudaMalloc(&h->G, (off + h->nbias) * sizeof(float)));
    CK(cudaMalloc(&h->Phi, off * sizeof(bf16))); CK(cudaMalloc(&h->Plo, off * sizeof(bf16)));
    CK(cudaMemset(h->P, 0, off * sizeof(float))); CK(cudaMemset(h->Dl, 0, off * sizeof(float))); CK(cudaMemset(h->G, 0, (off + h->nbias) * sizeof(float)));
    CK(cudaMemset(h->Phi, 0, off * sizeof(bf16))); CK(cudaMemset(h->Plo, 0, off * sizeof(bf16)));
    if (h->tensor) {
        // factor arena: act_hi, act_lo, dx_hi, dx_lo of every layer, [fx_rows][upad] bf16 each (1 KB aligned)
        size_t o = 0;
        auto carve = [&](int l) { const size_t at = o; o += (((size_t)h->fx_rows * h->upad[l] * sizeof(bf16)) + 1023) & ~(size_t)1023; return at; };
        size_t at[GGD_MAXLAYER][4];
        for (int l = 0; l < h->L; l++) for (int k = 0; k < 4; k++) at[l][k] = carve(l);
        h->fx_bytes = o;
        CK(cudaMalloc(&h->fx_arena, o));
        CK(cudaMemset(h->fx_arena, 0, o));
        for (int l = 0; l < h->L; l++) {
            h->act_hi[l] = (bf16 *)(h->fx_arena + at[l][0]); h->act_lo[l] = (bf16 *)(h->fx_arena + at[l][1]);
            h->dx_hi[l] = (bf16 *)(h->fx_arena + at[l][2]); h->dx_lo[l] = (bf16 *)(h->fx_arena + at[l][3]);
        }
    }
    { const char *ev = getenv("GGD_FX_TRACE"); if (ev && atoi(ev) == 1) { CK(cudaMalloc(&h->fx_trace, FX_TRACE_WORDS * 8)); CK(cudaMemset(h->fx_trace, 0, FX_TRACE_WORDS * 8)); } }
    CK(cudaMalloc(&h->fx_counters, FXC_WORDS * sizeof(unsigned int))); CK(cudaMemset(h->fx_counters, 0, FXC_WORDS * sizeof(unsigned int)));
    for (int l = 0; l < h->L; l++) {
        const size_t n = (size_t)h->Mp * h->upad[l];
        if (h->tensor) {
        } else {
            CK(cudaMalloc(&h->x32[l], n * sizeof(float))); CK(cudaMalloc(&h->y32[l], n * sizeof(float)));
            CK(cudaMalloc(&h->dy32[l], n * sizeof(float))); CK(cudaMalloc(&h->dx32[l], n * sizeof(float)));
            CK(cudaMemset(h->x32[l], 0, n * sizeof(float))); CK(cudaMemset(h->y32[l], 0, n * sizeof(float)));
            CK(cudaMemset(h->dy32[l], 0, n * sizeof(float))); CK(cudaMemset(h->dx32[l], 0, n * sizeof(float)));
        }
    }
    const int D = h->units[h->L - 1];
    CK(cudaMalloc(&h->out32, (size_t)h->Mp * h->upad[h->L - 1] * sizeof(float)));
    CK(cudaMemset(h->out32, 0, (size_t)h->Mp * h->upad[h->L - 1] * sizeof(float)));
    CK(cudaMalloc(&h->alpha, D * sizeof(float))); CK(cudaMalloc(&h->colsum, D * sizeof(float)));
    CK(cudaMemset(h->alpha, 0, D * sizeof(float))); CK(cudaMemset(h->colsum, 0, D * sizeof(float)));
    CK(cudaMalloc(&h->ctl, sizeof(StepCtl))); CK(cudaMemset(h->ctl, 0, sizeof(StepCtl)));
    CK(cudaMalloc(&h->dwp_dev, sizeof(DwpArgs))); CK(cudaMalloc(&h->dww_dev, sizeof(DwwArgs))); CK(cudaMalloc(&h->dwp_counter, sizeof(unsigned int)));
    CK(cudaMemset(h->dwp_counter, 0, sizeof(unsigned int)));
    CK(cudaHostAlloc(&h->hang_host, HANG_WORDS * sizeof(unsigned int), cudaHostAllocMapped)); memset(h->hang_host, 0, HANG_WORDS * sizeof(unsigned int));
    CK(cudaHostGetDevicePointer(&h->hang_dev, h->hang_host, 0));
    // weights in: reference order (out + in*cur) -> padded pitch Np; then build the bf16 shadows with a zero-gradient-free pass
    for (int l = 1; l < h->L; l++) {
        const LayerInfo &ly = h->lay[l];
        if (!weights[l] || !bias[l]) { set_error("weights[%d]/bias[%d] is null", l, l); return fail(GGD_EINVAL); }
        CK(cudaMemcpy2D(h->P + ly.w_off, ly.Np * sizeof(float), weights[l], ly.cur * sizeof(float), ly.cur * sizeof(float), ly.prev, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->P + ly.b_off, bias[l], ly.cur * sizeof(float), cudaMemcpyHostToDevice));
        launch_split_rows(h->P + ly.w_off, ly.Kp, ly.Np, h->Phi + ly.w_off, h->Plo + ly.w_off, ly.Np, 0);
    }
    CK(cudaDeviceSynchronize());
    if (h->tensor) { int rc = gemm_tc_init(); if (rc == GGD_OK) rc = dw_persist_init(); if (rc == GGD_OK) rc = dw_wide_init(); if (rc != GGD_OK) return fail(rc); }
    if (world > 1) {
        if (!cfg->nccl_unique_id) { set_error("world_size > 1 needs nccl_unique_id"); return fail(GGD_EINVAL); }
        ncclUniqueId id;
        memcpy(&id, cfg->nccl_unique_id, sizeof id);
        ncclConfig_t ncfg = NCCL_CONFIG_INITIALIZER;
        const char *mc = getenv("GGD_NCCL_MAX_CTAS");
        if (mc) { ncfg.maxCTAs = atoi(mc); ncfg.minCTAs = ncfg.maxCTAs < 4 ? ncfg.maxCTAs : 4; }
        ncclResult_t r = ncclCommInitRankConfig(&h->comm, world, id, cfg->rank, &ncfg);
        if (r != ncclSuccess) { set_error("ncclCommInitRankConfig: %s", ncclGetErrorString(r)); return fail(GGD_ENCCL); }
        h->has_comm = true;
        if (h->dp_fx) { int rc = dp_fx_setup(h); if (rc != GGD_OK) return fail(rc); }
    }
    {
        const char *ev = getenv("GGD_FUSE_LOSS");   // 0: separate loss kernel (A-B / tuning)
        h->fuse_loss = h->tensor && h->Mp == 128 && (world == 1 || h->fx_loss) && !(ev && atoi(ev) == 0);
    }
#undef CK
    *out = h;
    return GGD_OK;
}

int ggd_destroy(ggd_handle *h)
{
    if (!h) return GGD_OK;
    cudaSetDevice(h->cfg.gpu);
    if (h->s_main) cudaStreamSynchronize(h->s_main);
    if (h->s_copy) { cudaStreamSynchronize(h->s_copy); cudaStreamDestroy(h->s_copy); }
    if (h->dp_fx && h->fx_flags && h->fx_counters && h->fx_peer[0][h->cfg.rank]) {
        // Quiesce before my buffers go away: the LAST write a peer makes into my memory is its DONE flag of my last step (raised
        // by the last CTA of its update kernel, possibly a little after my own step has finished).  Bounded: a lost peer must
        // not turn destruction into a hang.
        unsigned int step = 0;
        const int world = h->cfg.world_size;
        if (cudaMemcpy(&step, h->fx_counters + FXC_STEP, sizeof step, cudaMemcpyDeviceToHost) == cudaSuccess) {
            std::vector<unsigned int> fl((size_t)world * FX_STRIDE);
            for (int tries = 0; tries < 2000; tries++) {
                if (cudaMemcpy(fl.data(), h->fx_flags, fl.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost) != cudaSuccess) break;
                bool all = true;
                for (int p = 0; p < world; p++)
                    if (p != h->cfg.rank && (int)(fl[(size_t)p * FX_STRIDE + FX_EV_DONE] - step) < 0) all = false;
                if (all) break;
                usleep(1000);
            }
        }
        cudaGetLastError();
    }
    for (cudaEvent_t e : h->ev_piece) cudaEventDestroy(e);
    if (h->ev_c0) cudaEventDestroy(h->ev_c0);
    if (h->ev_c1) cudaEventDestroy(h->ev_c1);
    free_chunk(h);
    for (auto &kv : h->pinned) if (kv.second) cudaHostUnregister(const_cast<void *>(kv.first));
    if (h->dp_fx && h->fx_peer[0][h->cfg.rank])
        for (int p = 0; p < h->cfg.world_size; p++)
            for (int k = 0; k < 3; k++) if (p != h->cfg.rank && h->fx_peer[k][p]) cudaIpcCloseMemHandle(h->fx_peer[k][p]);
    cudaFree(h->fx_trace); cudaFree(h->fx_asum); cudaFree(h->fx_flags); cudaFree(h->fx_counters); cudaFree(h->fx_arena); cudaFree(h->dww_dev);
    cudaFree(h->r_fea); cudaFree(h->r_targ); cudaFree(h->r_first); cudaFree(h->r_norm);
    if (h->has_comm) ncclCommDestroy(h->comm);
    cudaFree(h->P); cudaFree(h->Dl); cudaFree(h->G); cudaFree(h->Phi); cudaFree(h->Plo);
    for (int l = 0; l < GGD_MAXLAYER; l++) { cudaFree(h->x32[l]); cudaFree(h->y32[l]); cudaFree(h->dy32[l]); cudaFree(h->dx32[l]); }
    cudaFree(h->out32); cudaFree(h->alpha); cudaFree(h->colsum); cudaFree(h->ctl); cudaFree(h->dwp_dev); cudaFree(h->dwp_counter); if (h->hang_host) cudaFreeHost(h->hang_host);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev2) cudaEventDestroy(h->ev2);
    for (int k = 0; k < FX_NEVENTS; k++) if (h->ev_fx[k]) cudaEventDestroy(h->ev_fx[k]);
    for (int k = 0; k < FX_NSIDE; k++) if (h->s_side[k]) cudaStreamDestroy(h->s_side[k]);
    if (h->s_main) cudaStreamDestroy(h->s_main);
    delete h;
    return GGD_OK;
}

int ggd_reserve(ggd_handle *h, int n_frames)
{
    if (!h || n_frames < 1 || n_frames > GGD_MAXCACHEFRAME) { set_error("ggd_reserve: bad argument"); return GGD_EINVAL; }
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    GGD_TRY(ensure_chunk(h, n_frames));
    if (!(h->cfg.flags & GGD_FLAG_NO_GRAPH) && !h->g1) {
        GGD_TRY(set_ctl(h, h->c_in, h->c_targ));
        GGD_TRY(capture_graph(h, 1, &h->g1));
        GGD_TRY(capture_graph(h, 4, &h->g4));
        h->gN_steps = 16;
        GGD_TRY(capture_graph(h, h->gN_steps, &h->gN));
    }
    GGD_CUDA(cudaStreamSynchronize(h->s_main));
    return GGD_OK;
}

int ggd_train(ggd_handle *h, int n_frames, const float *in, const float *targ)
{
    if (!h || !in || !targ || n_frames < 0) { set_error("ggd_train: bad argument"); return GGD_EINVAL; }
    if (n_frames > GGD_MAXCACHEFRAME) { set_error("n_frames %d exceeds MAXCACHEFRAME %d", n_frames, GGD_MAXCACHEFRAME); return GGD_EINVAL; }
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    GGD_TRY(ensure_chunk(h, n_frames));
    const int D = h->units[h->L - 1];
    const size_t bi = (size_t)n_frames * h->units[0] * sizeof(float), bt = (size_t)n_frames * D * sizeof(float);
    GGD_TRY(pin_host(h, in, bi));
    GGD_TRY(pin_host(h, targ, bt));
    GGD_TRY(run_chunk(h, n_frames, h->c_in, h->c_targ, in, targ));
    float ms = 0;
    if (n_frames / h->M > 0) cudaEventElapsedTime(&ms, h->ev_c0, h->ev_c1);
    h->stats.h2d_ms = ms; h->stats.h2d_bytes = bi + bt;   // copy-stream time; it overlaps the steps of the earlier pieces
    return GGD_OK;
}

int ggd_release_host(ggd_handle *h, const void *host_ptr)
{
    if (!h) { set_error("ggd_release_host: null handle"); return GGD_EINVAL; }
    auto it = h->pinned.find(host_ptr);
    if (it == h->pinned.end()) return GGD_OK;
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    if (h->s_copy) cudaStreamSynchronize(h->s_copy);
    if (it->second) cudaHostUnregister(const_cast<void *>(host_ptr));
    h->pinned.erase(it);
    return GGD_OK;
}

int ggd_train_raw(ggd_handle *h, const ggd_raw_chunk *c)
{
    if (!h || !c || !c->fea_records || !c->targ_records || !c->sample_first_frame || !c->mean || !c->dvar || c->n_samples < 0 || c->n_frames < 0) {
        set_error("ggd_train_raw: bad argument"); return GGD_EINVAL;
    }
    const int D = h->units[h->L - 1];
    if (c->fea_dim < 1 || c->fea_context < 1 || c->fea_dim * c->fea_context != h->units[0]) {
        set_error("ggd_train_raw: fea_dim %d x fea_context %d does not match layersizes[0] %d", c->fea_dim, c->fea_context, h->units[0]); return GGD_EINVAL;
    }
    if (c->targ_offset < 0 || c->targ_offset >= c->fea_context) { set_error("ggd_train_raw: targ_offset %d outside the context window", c->targ_offset); return GGD_EINVAL; }
    if (c->n_samples > GGD_MAXCACHEFRAME) { set_error("n_samples %d exceeds MAXCACHEFRAME %d", c->n_samples, GGD_MAXCACHEFRAME); return GGD_EINVAL; }
    for (int i = 0; i < c->n_samples; i++)
        if (c->sample_first_frame[i] < 0 || c->sample_first_frame[i] + c->fea_context > c->n_frames) {
            set_error("ggd_train_raw: sample %d starts at frame %d, outside the %d frames of the chunk", i, c->sample_first_frame[i], c->n_frames); return GGD_EINVAL;
        }
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    GGD_TRY(ensure_chunk(h, c->n_samples));
    const int world = h->has_comm ? h->cfg.world_size : 1;
    const bool sliced = c->rec_frames > 0 || c->rec_frame0 > 0;
    const int S = ceil_div(c->n_frames, world);                 // frames per rank slice
    if (sliced) {
        const int want0 = h->cfg.rank * S, wantn = std::max(0, std::min(S, c->n_frames - want0));
        if (world < 2 || c->rec_frame0 != want0 || c->rec_frames != wantn) {
            set_error("ggd_train_raw: record slice [%d, +%d) given, rank %d of %d must supply [%d, +%d)", c->rec_frame0, c->rec_frames, h->cfg.rank, world, want0, wantn);
            return GGD_EINVAL;
        }
    }
    const size_t nf = (size_t)(sliced ? S * world : c->n_frames) * (2 + c->fea_dim), nt = (size_t)(sliced ? S * world : c->n_frames) * (2 + D);
    GGD_TRY(grow(&h->r_fea, &h->r_cap_fea, nf));
    GGD_TRY(grow(&h->r_targ, &h->r_cap_targ, nt));
    GGD_TRY(grow(&h->r_first, &h->r_cap_first, (size_t)c->n_samples));
    GGD_TRY(grow(&h->r_norm, &h->r_cap_norm, (size_t)2 * c->fea_dim));
    // (the record buffers are NOT pinned: their size varies from chunk to chunk, so a caller is free to reallocate them,
    // and a registration that outlives its allocation corrupts the address space; the raw copy is 4x smaller anyway)
    GGD_CUDA(cudaEventRecord(h->ev_c0, h->s_main));
    if (sliced) {
        // my slice lands at its place in the chunk-sized buffers; the other slices arrive over NVLink (in-place all-gather)
        const size_t wf = (size_t)S * (2 + c->fea_dim), wt = (size_t)S * (2 + D);
        if (c->rec_frames > 0) {
            GGD_CUDA(cudaMemcpyAsync(h->r_fea + h->cfg.rank * wf, c->fea_records, (size_t)c->rec_frames * (2 + c->fea_dim) * 4, cudaMemcpyHostToDevice, h->s_main));
            GGD_CUDA(cudaMemcpyAsync(h->r_targ + h->cfg.rank * wt, c->targ_records, (size_t)c->rec_frames * (2 + D) * 4, cudaMemcpyHostToDevice, h->s_main));
        }
        GGD_NCCL(ncclAllGather(h->r_fea + h->cfg.rank * wf, h->r_fea, wf, ncclUint32, h->comm, h->s_main));
        GGD_NCCL(ncclAllGather(h->r_targ + h->cfg.rank * wt, h->r_targ, wt, ncclUint32, h->comm, h->s_main));
    } else {
        GGD_CUDA(cudaMemcpyAsync(h->r_fea, c->fea_records, nf * 4, cudaMemcpyHostToDevice, h->s_main));
        GGD_CUDA(cudaMemcpyAsync(h->r_targ, c->targ_records, nt * 4, cudaMemcpyHostToDevice, h->s_main));
    }
    GGD_CUDA(cudaMemcpyAsync(h->r_first, c->sample_first_frame, (size_t)c->n_samples * sizeof(int), cudaMemcpyHostToDevice, h->s_main));
    GGD_CUDA(cudaMemcpyAsync(h->r_norm, c->mean, (size_t)c->fea_dim * sizeof(float), cudaMemcpyHostToDevice, h->s_main));
    GGD_CUDA(cudaMemcpyAsync(h->r_norm + c->fea_dim, c->dvar, (size_t)c->fea_dim * sizeof(float), cudaMemcpyHostToDevice, h->s_main));
    GGD_CUDA(cudaEventRecord(h->ev_c1, h->s_main));
    ExpandArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.fea_rec = h->r_fea; ea.targ_rec = h->r_targ; ea.first = h->r_first; ea.mean = h->r_norm; ea.dvar = h->r_norm + c->fea_dim;
    ea.samples = c->n_samples; ea.fea_dim = c->fea_dim; ea.ctx = c->fea_context; ea.targ_offset = c->targ_offset; ea.D = D;
    ea.in32 = h->tensor ? nullptr : h->c_in;
    ea.in_hi = h->tensor ? h->c_hi : nullptr; ea.in_lo = h->tensor ? h->c_lo : nullptr; ea.ld = h->upad[0];
    ea.targ = h->c_targ;
    launch_expand_chunk(ea, h->s_main);
    GGD_CUDA(cudaGetLastError());
    GGD_TRY(run_chunk(h, c->n_samples, h->c_in, h->c_targ, nullptr, nullptr, true));
    h->stats.launches += 1;
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev_c0, h->ev_c1);
    h->stats.h2d_ms = ms;
    h->stats.h2d_bytes = (sliced ? (size_t)c->rec_frames * (4 + c->fea_dim + D) : nf + nt) * 4 + (size_t)c->n_samples * sizeof(int);
    return GGD_OK;
}

int ggd_bind_thread(ggd_handle *h)
{
    if (!h) { set_error("ggd_bind_thread: null handle"); return GGD_EINVAL; }
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    return GGD_OK;
}

void *ggd_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); set_error("ggd_host_alloc(%zu) failed", bytes); return nullptr; }
    return p;
}
void ggd_host_free(void *p) { if (p) cudaFreeHost(p); }

int ggd_train_device(ggd_handle *h, int n_frames, const float *d_in, const float *d_targ)
{
    if (!h || !d_in || !d_targ || n_frames < 0) { set_error("ggd_train_device: bad argument"); return GGD_EINVAL; }
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    GGD_TRY(ensure_chunk(h, n_frames));
    h->stats.h2d_ms = 0; h->stats.h2d_bytes = 0;
    return run_chunk(h, n_frames, d_in, d_targ);
}

// forward-only over the n frames whose net input is already in the chunk arrays (c_in, or c_hi / c_lo on the tensor path),
// in bunches of M (a partial last bunch IS processed: BP_GPU.cu:203-218); the outputs land in c_out and in h_out
static int forward_resident(ggd_handle *h, int n_frames, bool denorm = false, const float *d_mean = nullptr, const float *d_dvar = nullptr, int fea_dim = 0)
{
    const int D = h->units[h->L - 1], ldo = h->upad[h->L - 1];
    GGD_TRY(set_ctl(h, h->c_in, h->c_targ));
    const int nb = ceil_div(n_frames, h->M);
    int launches = 0;
    for (int b = 0; b < nb; b++) {
        const int f = (n_frames - b * h->M < h->M) ? n_frames - b * h->M : h->M;
        GGD_TRY(enqueue_forward(h, h->s_main, &launches));
        GGD_CUDA(cudaMemcpy2DAsync(h->c_out + (size_t)b * h->M * D, D * sizeof(float), h->out32, ldo * sizeof(float), D * sizeof(float), f,
                                   cudaMemcpyDeviceToDevice, h->s_main));
        launch_advance(h->ctl, h->s_main);
    }
    if (denorm) launch_denorm(h->c_out, n_frames, D, d_mean, d_dvar, fea_dim, h->s_main);
    h->h_out.resize((size_t)n_frames * D);
    GGD_CUDA(cudaMemcpyAsync(h->h_out.data(), h->c_out, (size_t)n_frames * D * sizeof(float), cudaMemcpyDeviceToHost, h->s_main));
    GGD_CUDA(cudaStreamSynchronize(h->s_main));
    return GGD_OK;
}

static int forward_chunk(ggd_handle *h, int n_frames, const float *in)
{
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    GGD_TRY(ensure_chunk(h, n_frames));
    GGD_TRY(upload(h, in, h->c_in, (size_t)n_frames * h->units[0] * sizeof(float)));
    if (h->tensor) launch_split_rows(h->c_in, n_frames, h->units[0], h->c_hi, h->c_lo, h->upad[0], h->s_main);
    return forward_resident(h, n_frames);
}

int ggd_forward(ggd_handle *h, int n_frames, const float *in, float *out)
{
    if (!h || !in || !out || n_frames < 0) { set_error("ggd_forward: bad argument"); return GGD_EINVAL; }
    GGD_TRY(forward_chunk(h, n_frames, in));
    memcpy(out, h->h_out.data(), h->h_out.size() * sizeof(float));
    return GGD_OK;
}

// The three CV metrics accumulate on the host in float, frame-major order, exactly like the reference.
static float metric_sqerr(const ggd_handle *h, int n_frames, const float *targ)
{
    const size_t n = (size_t)n_frames * h->units[h->L - 1];
    const float *o = h->h_out.data();
    float s = 0.0f;
    for (size_t i = 0; i < n; i++) s = s + (o[i] - targ[i]) * (o[i] - targ[i]);   // BP_GPU.cu:211
    return s;
}
static float metric_abserr(const ggd_handle *h, int n_frames, const float *targ)
{
    const int D = h->units[h->L - 1];
    const size_t n = (size_t)n_frames * D;
    const float *o = h->h_out.data();
    float s = 0.0f;
    for (size_t i = 0; i < n; i++) s = s + fabsf(o[i] - targ[i]);                  // BP_GPU.cu:244
    return s / D;                                                                  // BP_GPU.cu:250
}

int ggd_cv_sqerr(ggd_handle *h, int n_frames, const float *in, const float *targ, float *result)
{
    if (!h || !in || !targ || !result) { set_error("ggd_cv_sqerr: bad argument"); return GGD_EINVAL; }
    GGD_TRY(forward_chunk(h, n_frames, in));
    *result = metric_sqerr(h, n_frames, targ);
    return GGD_OK;
}

int ggd_cv_abserr(ggd_handle *h, int n_frames, const float *in, const float *targ, float *result)
{
    if (!h || !in || !targ || !result) { set_error("ggd_cv_abserr: bad argument"); return GGD_EINVAL; }
    GGD_TRY(forward_chunk(h, n_frames, in));
    *result = metric_abserr(h, n_frames, targ);
    return GGD_OK;
}

static float gamma_ref(float x)   // BP_GPU::Gamma, BP_GPU.cu:593-640
{
    if (x > 2 && x <= 3) {
        static const float c[11] = {0.0000677106f, -0.0003442342f, 0.0015397681f, -0.0024467480f, 0.0109736958f, -0.0002109075f,
                                    0.0742379071f, 0.0815782188f,  0.4118402518f, 0.4227843370f,  1.0000000000f};
        const double t = x - 2.0;
        float temp = 0;
        temp = temp + c[0] * pow(t, 10.0) + c[1] * pow(t, 9.0);
        temp = temp + c[2] * pow(t, 8.0) + c[3] * pow(t, 7.0);
        temp = temp + c[4] * pow(t, 6.0) + c[5] * pow(t, 5.0);
        temp = temp + c[6] * pow(t, 4.0) + c[7] * pow(t, 3.0);
        temp = temp + c[8] * pow(t, 2.0) + c[9] * t + c[10];
        return temp;
    }
    if (x > 0 && x <= 1) return gamma_ref(x + 2) / (x * (x + 1));
    if (x > 1 && x <= 2) return gamma_ref(x + 1) / x;
    if (x > 3) {
        int i = 1;
        float temp = 1;
        while (!((x - i) > 2 && (x - i) <= 3)) { temp = (x - i) * temp; i++; }
        temp = temp * (x - i);
        return temp * gamma_ref(x - i);
    }
    return 0;
}

static int metric_loglik(ggd_handle *h, int n_frames, const float *targ, float *result);

int ggd_cv_loglik(ggd_handle *h, int n_frames, const float *in, const float *targ, float *result)
{
    if (!h || !in || !targ || !result) { set_error("ggd_cv_loglik: bad argument"); return GGD_EINVAL; }
    GGD_TRY(forward_chunk(h, n_frames, in));
    return metric_loglik(h, n_frames, targ, result);
}

int ggd_cv_all(ggd_handle *h, int n_frames, const float *in, const float *targ, float *result3)
{
    if (!h || !in || !targ || !result3) { set_error("ggd_cv_all: bad argument"); return GGD_EINVAL; }
    GGD_TRY(forward_chunk(h, n_frames, in));          // ONE forward pass feeds the three metrics (the reference runs it three times)
    result3[0] = metric_sqerr(h, n_frames, targ);
    result3[1] = metric_abserr(h, n_frames, targ);
    result3[2] = 0.0f;
    if (h->cfg.MLflag == 1) GGD_TRY(metric_loglik(h, n_frames, targ, &result3[2]));
    return GGD_OK;
}

int ggd_enhance(ggd_handle *h, int n_frames, const float *lps, int fea_dim, int fea_context, const float *mean, const float *dvar, float *out)
{
    if (!h || !lps || !mean || !dvar || !out || n_frames < 0) { set_error("ggd_enhance: bad argument"); return GGD_EINVAL; }
    const int D = h->units[h->L - 1];
    if (fea_dim < 1 || fea_context < 1 || (fea_context & 1) == 0 || fea_dim * fea_context != h->units[0]) {
        set_error("ggd_enhance: fea_dim %d x fea_context %d (odd) must equal layersizes[0] %d", fea_dim, fea_context, h->units[0]); return GGD_EINVAL;
    }
    if (n_frames > GGD_MAXCACHEFRAME) { set_error("n_frames %d exceeds MAXCACHEFRAME %d", n_frames, GGD_MAXCACHEFRAME); return GGD_EINVAL; }
    if (n_frames == 0) return GGD_OK;
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    GGD_TRY(ensure_chunk(h, n_frames));
    // raw features + norm constants ride in the loader's staging buffers
    GGD_TRY(grow(&h->r_fea, &h->r_cap_fea, (size_t)n_frames * fea_dim));
    GGD_TRY(grow(&h->r_norm, &h->r_cap_norm, (size_t)2 * fea_dim));
    float *d_lps = reinterpret_cast<float *>(h->r_fea);
    GGD_CUDA(cudaMemcpyAsync(d_lps, lps, (size_t)n_frames * fea_dim * sizeof(float), cudaMemcpyHostToDevice, h->s_main));
    GGD_CUDA(cudaMemcpyAsync(h->r_norm, mean, (size_t)fea_dim * sizeof(float), cudaMemcpyHostToDevice, h->s_main));
    GGD_CUDA(cudaMemcpyAsync(h->r_norm + fea_dim, dvar, (size_t)fea_dim * sizeof(float), cudaMemcpyHostToDevice, h->s_main));
    EdgeExpandArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.lps = d_lps; ea.mean = h->r_norm; ea.dvar = h->r_norm + fea_dim; ea.frames = n_frames; ea.fea_dim = fea_dim; ea.ctx = fea_context;
    ea.in32 = h->tensor ? nullptr : h->c_in; ea.in_hi = h->tensor ? h->c_hi : nullptr; ea.in_lo = h->tensor ? h->c_lo : nullptr; ea.ld = h->upad[0];
    launch_expand_edges(ea, h->s_main);
    GGD_CUDA(cudaGetLastError());
    GGD_TRY(forward_resident(h, n_frames, true, h->r_norm, h->r_norm + fea_dim, fea_dim));
    memcpy(out, h->h_out.data(), (size_t)n_frames * D * sizeof(float));
    return GGD_OK;
}

static int metric_loglik(ggd_handle *h, int n_frames, const float *targ, float *result)
{
    const int D = h->units[h->L - 1];
    std::vector<float> al(D);
    GGD_CUDA(cudaMemcpy(al.data(), h->alpha, D * sizeof(float), cudaMemcpyDeviceToHost));
    const float sf = h->cfg.shapefactor;
    const float *o = h->h_out.data();
    float d1 = n_frames * D * logf(sf / (2 * gamma_ref((float)(1.0 / sf))));
    float d2 = 0, d3 = 0;
    for (int u = 0; u < D; u++) d2 += logf(al[u]);
    d2 = d2 * n_frames;
    for (int f = 0; f < n_frames; f++)
        for (int d = 0; d < D; d++) d3 += powf(fabsf(targ[(size_t)f * D + d] - o[(size_t)f * D + d]) / al[d], sf);
    *result = d1 - d2 - d3;                                                        // BP_GPU.cu:287-301
    return GGD_OK;
}

int ggd_get_weights(ggd_handle *h, float *const *weights, float *const *bias)
{
    if (!h || !weights || !bias) { set_error("ggd_get_weights: bad argument"); return GGD_EINVAL; }
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    GGD_CUDA(cudaStreamSynchronize(h->s_main));
    // (data parallel: every rank holds the full, bit-identical fp32 master weights -- nothing to gather)
    for (int l = 1; l < h->L; l++) {
        const LayerInfo &ly = h->lay[l];
        GGD_CUDA(cudaMemcpy2D(weights[l], ly.cur * sizeof(float), h->P + ly.w_off, ly.Np * sizeof(float), ly.cur * sizeof(float), ly.prev, cudaMemcpyDeviceToHost));
        GGD_CUDA(cudaMemcpy(bias[l], h->P + ly.b_off, ly.cur * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return GGD_OK;
}

int ggd_get_alpha(ggd_handle *h, float *alpha)
{
    if (!h || !alpha) { set_error("ggd_get_alpha: bad argument"); return GGD_EINVAL; }
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    GGD_CUDA(cudaStreamSynchronize(h->s_main));
    GGD_CUDA(cudaMemcpy(alpha, h->alpha, h->units[h->L - 1] * sizeof(float), cudaMemcpyDeviceToHost));
    return GGD_OK;
}

int ggd_get_losses(ggd_handle *h, float *losses, int max, int *n)
{
    if (!h || !n) { set_error("ggd_get_losses: bad argument"); return GGD_EINVAL; }
    const int k = (int)h->losses.size() < max ? (int)h->losses.size() : max;
    if (losses) memcpy(losses, h->losses.data(), k * sizeof(float));
    *n = (int)h->losses.size();
    return GGD_OK;
}

int ggd_get_stats(ggd_handle *h, ggd_stats *s)
{
    if (!h || !s) { set_error("ggd_get_stats: bad argument"); return GGD_EINVAL; }
    *s = h->stats;
    return GGD_OK;
}

int ggd_debug_step(ggd_handle *h, int n_frames, const float *in, const float *targ, int apply_update)
{
    if (!h || !in || !targ || n_frames != h->M) { set_error("ggd_debug_step: n_frames must equal bunchsize"); return GGD_EINVAL; }
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    GGD_TRY(ensure_chunk(h, n_frames));
    const int D = h->units[h->L - 1];
    GGD_CUDA(cudaMemcpyAsync(h->c_in, in, (size_t)n_frames * h->units[0] * sizeof(float), cudaMemcpyHostToDevice, h->s_main));
    GGD_CUDA(cudaMemcpyAsync(h->c_targ, targ, (size_t)n_frames * D * sizeof(float), cudaMemcpyHostToDevice, h->s_main));
    GGD_TRY(set_ctl(h, h->c_in, h->c_targ));
    GGD_CUDA(cudaMemsetAsync(h->trace, 0, h->trace_cap * sizeof(double), h->s_main));
    if (h->tensor) launch_split_rows(h->c_in, n_frames, h->units[0], h->c_hi, h->c_lo, h->upad[0], h->s_main);
    int launches = 0;
    GGD_TRY(enqueue_step(h, h->s_main, apply_update != 0, &launches, false));   // unfused: the gradient stays readable
    double tr = 0;
    GGD_CUDA(cudaMemcpyAsync(&tr, h->trace, sizeof(double), cudaMemcpyDeviceToHost, h->s_main));
    GGD_CUDA(cudaStreamSynchronize(h->s_main));
    h->losses.assign(1, (float)tr);
    return GGD_OK;
}

static int read_pair(ggd_handle *h, const bf16 *hi, const bf16 *lo, int ld, int rows, int cols, float *dst)
{
    std::vector<uint16_t> a((size_t)rows * ld), b((size_t)rows * ld);
    GGD_CUDA(cudaMemcpy(a.data(), hi, a.size() * 2, cudaMemcpyDeviceToHost));
    GGD_CUDA(cudaMemcpy(b.data(), lo, b.size() * 2, cudaMemcpyDeviceToHost));
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++) {
            uint32_t x = (uint32_t)a[(size_t)r * ld + c] << 16, y = (uint32_t)b[(size_t)r * ld + c] << 16;
            float fx, fy;
            memcpy(&fx, &x, 4); memcpy(&fy, &y, 4);
            dst[(size_t)r * cols + c] = fx + fy;
        }
    return GGD_OK;
}

int ggd_debug_read(ggd_handle *h, int what, int layer, float *dst)
{
    if (!h || !dst) { set_error("ggd_debug_read: bad argument"); return GGD_EINVAL; }
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    GGD_CUDA(cudaStreamSynchronize(h->s_main));
    const int L = h->L;
    if (what == 0) {
        const int D = h->units[L - 1];
        GGD_CUDA(cudaMemcpy2D(dst, D * sizeof(float), h->out32, h->upad[L - 1] * sizeof(float), D * sizeof(float), h->M, cudaMemcpyDeviceToHost));
        return GGD_OK;
    }
    if (layer < 1 || layer >= L) { set_error("ggd_debug_read: layer %d out of range", layer); return GGD_EINVAL; }
    const LayerInfo &ly = h->lay[layer];
    switch (what) {
    case 1:
        if (h->tensor) return read_pair(h, loc(h, h->dx_hi[layer], layer), loc(h, h->dx_lo[layer], layer), ly.Np, h->M, ly.cur, dst);
        GGD_CUDA(cudaMemcpy2D(dst, ly.cur * sizeof(float), h->dx32[layer], ly.Np * sizeof(float), ly.cur * sizeof(float), h->M, cudaMemcpyDeviceToHost));
        return GGD_OK;
    case 2:
        if (layer == L - 1) { set_error("layer %d is linear: read what=0", layer); return GGD_EINVAL; }
        if (h->tensor) return read_pair(h, loc(h, h->act_hi[layer], layer), loc(h, h->act_lo[layer], layer), ly.Np, h->M, ly.cur, dst);
        GGD_CUDA(cudaMemcpy2D(dst, ly.cur * sizeof(float), h->y32[layer], ly.Np * sizeof(float), ly.cur * sizeof(float), h->M, cudaMemcpyDeviceToHost));
        return GGD_OK;
    case 3:
        GGD_CUDA(cudaMemcpy2D(dst, ly.cur * sizeof(float), h->G + ly.w_off, ly.Np * sizeof(float), ly.cur * sizeof(float), ly.prev, cudaMemcpyDeviceToHost));
        return GGD_OK;
    case 4:
        GGD_CUDA(cudaMemcpy(dst, h->G + ly.gb_off, ly.cur * sizeof(float), cudaMemcpyDeviceToHost));
        return GGD_OK;
    }
    set_error("ggd_debug_read: unknown selector %d", what);
    return GGD_EINVAL;
}

int ggd_profile_kernels(ggd_handle *h, int n_frames, const float *d_in, const float *d_targ, ggd_kernel_times *out)
{
    if (!h || !d_in || !d_targ || !out) { set_error("ggd_profile_kernels: bad argument"); return GGD_EINVAL; }
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    GGD_TRY(ensure_chunk(h, n_frames));
    const int nb = n_frames / h->M;
    memset(out, 0, sizeof *out);
    if (nb == 0) return GGD_OK;
    if (h->has_comm) GGD_TRY(dp_barrier(h));
    GGD_TRY(set_ctl(h, d_in, d_targ));
    GGD_CUDA(cudaMemsetAsync(h->trace, 0, h->trace_cap * sizeof(double), h->s_main));
    h->prof_on = true;
    int rc = GGD_OK, launches = 0;
    if (h->tensor) { ProfScope ps(h, KC_SPLIT, h->s_main); launch_split_rows(d_in, nb * h->M, h->units[0], h->c_hi, h->c_lo, h->upad[0], h->s_main); }
    for (int b = 0; b < nb && rc == GGD_OK; b++) rc = enqueue_step(h, h->s_main, true, &launches);
    h->prof_on = false;
    if (rc == GGD_OK) rc = sync_main(h);
    else if (h->hang_host && h->hang_host[0] == 0xDEADu) {
        // a device-side watchdog fired while launches were still being queued: say which one
        const unsigned int *r = h->hang_host;
        std::string prev = ggd_last_error();
        set_error("%s [device watchdog: wait code %u gave up in block %u at iteration %u (parity/peer %u, thread %u)]", prev.c_str(), r[1], r[2], r[3], r[4], r[5]);
    }
    for (size_t i = 0; i < h->prof_cls.size(); i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]);
        out->ms[h->prof_cls[i]] += ms;
        out->launches[h->prof_cls[i]] += 1;
    }
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    h->prof_ev.clear(); h->prof_cls.clear();
    out->steps = nb;
    out->param_elems = (long long)h->arena;
    return rc;
}

// One training step with per-CTA phase stamps in every tensor-core kernel (tests / tuning only).
// out[launch][12]: kind (0 fwd, 2 dx, 3 dw), ctas, first entry (us since the first kernel), last exit,
// then medians of slots 1,2,3,4,6,7,8,9 relative to the CTA's own entry.
int ggd_debug_trace_step(ggd_handle *h, const float *in, const float *targ, int fused, float *out, int max_launches, int *n_launches)
{
    if (!h || !in || !targ || !out || !n_launches || !h->tensor) { set_error("ggd_debug_trace_step: bad argument"); return GGD_EINVAL; }
    GGD_CUDA(cudaSetDevice(h->cfg.gpu));
    GGD_TRY(ensure_chunk(h, h->M));
    const int D = h->units[h->L - 1], L = h->L;
    struct Item { int kind; int ctas; unsigned long long **slot; };
    std::vector<Item> items;
    for (int l = 1; l < L; l++) {
        GemmPlan &fp = (h->fuse_loss && l == L - 1) ? h->fwd_loss : h->fwd[l];
        items.push_back({0, fp.splits * fp.tiles_i * fp.tiles_j, &fp.args.trace});
    }
    for (int l = L - 1; l > 0; l--) {
        if (l != 1) items.push_back({2, h->dxp[l].splits * h->dxp[l].tiles_i * h->dxp[l].tiles_j, &h->dxp[l].args.trace});
        if (fused && h->fused) continue;     // the persistent gradient+update kernels carry no stamps
        items.push_back({3, h->dwp[l].tiles_i * h->dwp[l].tiles_j, &h->dwp[l].args.trace});
    }
    size_t total = 0;
    for (auto &it : items) total += (size_t)it.ctas * 16;
    unsigned long long *dbuf = nullptr;
    GGD_CUDA(cudaMalloc(&dbuf, total * 8));
    GGD_CUDA(cudaMemset(dbuf, 0, total * 8));
    GGD_CUDA(cudaMemcpyAsync(h->c_in, in, (size_t)h->M * h->units[0] * sizeof(float), cudaMemcpyHostToDevice, h->s_main));
    GGD_CUDA(cudaMemcpyAsync(h->c_targ, targ, (size_t)h->M * D * sizeof(float), cudaMemcpyHostToDevice, h->s_main));
    launch_split_rows(h->c_in, h->M, h->units[0], h->c_hi, h->c_lo, h->upad[0], h->s_main);
    int launches = 0, rc = GGD_OK;
    for (int rep = 0; rep < 3 && rc == GGD_OK; rep++) {     // two warm steps, then the traced one
        GGD_TRY(set_ctl(h, h->c_in, h->c_targ));
        if (rep == 2) { size_t off = 0; for (auto &it : items) { *it.slot = dbuf + off; off += (size_t)it.ctas * 16; } }
        rc = enqueue_step(h, h->s_main, true, &launches, fused != 0);
    }
    cudaStreamSynchronize(h->s_main);
    for (auto &it : items) *it.slot = nullptr;
    std::vector<unsigned long long> hb(total);
    cudaMemcpy(hb.data(), dbuf, total * 8, cudaMemcpyDeviceToHost);
    cudaFree(dbuf);
    if (rc != GGD_OK) return rc;
    unsigned long long t0 = ~0ull;
    for (size_t i = 0; i < total; i += 16) if (hb[i] && hb[i] < t0) t0 = hb[i];
    size_t off = 0;
    int n = 0;
    for (auto &it : items) {
        if (n >= max_launches) break;
        float *o = out + (size_t)n * 12;
        unsigned long long first = ~0ull, last = 0;
        int slots[8] = {1, 2, 3, 4, 6, 7, 8, 9};
        if (getenv("GGD_TRACE_LOSS")) { const int alt[8] = {7, 10, 11, 12, 13, 14, 8, 9}; for (int k = 0; k < 8; k++) slots[k] = alt[k]; }
        std::vector<double> med[8];
        for (int c = 0; c < it.ctas; c++) {
            const unsigned long long *t = &hb[off + (size_t)c * 16];
            if (!t[0]) continue;
            if (t[0] < first) first = t[0];
            if (t[9] > last) last = t[9];
            for (int k = 0; k < 8; k++) if (t[slots[k]]) med[k].push_back((double)(t[slots[k]] - t[0]) / 1000.0);
        }
        o[0] = (float)it.kind; o[1] = (float)it.ctas; o[2] = (float)((double)(first - t0) / 1000.0); o[3] = (float)((double)(last - t0) / 1000.0);
        for (int k = 0; k < 8; k++) {
            if (med[k].empty()) { o[4 + k] = -1; continue; }
            std::sort(med[k].begin(), med[k].end());
            o[4 + k] = (float)med[k][med[k].size() / 2];
        }
        off += (size_t)it.ctas * 16;
        n++;
    }
    *n_launches = n;
    return GGD_OK;
}

int ggd_nccl_unique_id(void *out128)
{
    if (!out128) { set_error("ggd_nccl_unique_id: null argument"); return GGD_EINVAL; }
    ncclUniqueId id;
    GGD_NCCL(ncclGetUniqueId(&id));
    memcpy(out128, &id, sizeof id);
    return GGD_OK;
}

static unsigned long long *g_gemm_trace = nullptr;   // set by ggd_debug_gemm_timed around ggd_debug_gemm
static int g_gemm_reps = 1;
static float g_gemm_ms = 0;

int ggd_debug_gemm_timed(int a_mn, int b_mn, int I, int J, int R, int bn, int splits, const float *A, const float *B, float *D,
                         int reps, float *avg_ms, unsigned long long *trace_host, int trace_ctas)
{
    unsigned long long *dtrace = nullptr;
    if (trace_host && trace_ctas > 0) {
        if (cudaMalloc(&dtrace, (size_t)trace_ctas * 16 * 8) != cudaSuccess) { set_error("trace alloc failed"); return GGD_ENOMEM; }
        cudaMemset(dtrace, 0, (size_t)trace_ctas * 16 * 8);
    }
    g_gemm_trace = dtrace; g_gemm_reps = reps > 0 ? reps : 1;
    int rc = ggd_debug_gemm(a_mn, b_mn, I, J, R, bn, splits, A, B, D);
    g_gemm_trace = nullptr; g_gemm_reps = 1;
    if (avg_ms) *avg_ms = g_gemm_ms;
    if (dtrace) {
        if (rc == GGD_OK) cudaMemcpy(trace_host, dtrace, (size_t)trace_ctas * 16 * 8, cudaMemcpyDeviceToHost);
        cudaFree(dtrace);
    }
    return rc;
}

int ggd_debug_gemm(int a_mn, int b_mn, int I, int J, int R, int bn, int splits, const float *A, const float *B, float *D)
{
    if (!A || !B || !D || I < 1 || J < 1 || R < 1) { set_error("ggd_debug_gemm: bad argument"); return GGD_EINVAL; }
    GGD_TRY(gemm_tc_init());
    const int Ip = round_up(I, 128), Jp = round_up(J, bn), Rp = round_up(R, 64);
    const int Ic = round_up(I, 64), Jc = round_up(J, 64);
    // operand shapes in memory: K-major [rows_p][Rp]; MN-major [Rp][rows rounded to 64]
    const int a_rows = a_mn ? R : I, a_cols = a_mn ? I : R, a_ld = a_mn ? Ic : Rp, a_rp = a_mn ? Rp : Ip;
    const int b_rows = b_mn ? R : J, b_cols = b_mn ? J : R, b_ld = b_mn ? Jc : Rp, b_rp = b_mn ? Rp : Jp;
    float *dA = nullptr, *dB = nullptr, *dD = nullptr;
    bf16 *ah = nullptr, *al = nullptr, *bh = nullptr, *bl = nullptr;
    int rc = GGD_OK;
    auto cleanup = [&]() { cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(ah); cudaFree(al); cudaFree(bh); cudaFree(bl); };
#define DG(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("%s -> %s", #expr, cudaGetErrorString(_e)); cleanup(); return GGD_ECUDA; } } while (0)
    DG(cudaMalloc(&dA, (size_t)a_rows * a_cols * 4)); DG(cudaMalloc(&dB, (size_t)b_rows * b_cols * 4));
    DG(cudaMalloc(&dD, (size_t)Ip * Jp * 4));
    DG(cudaMalloc(&ah, (size_t)a_rp * a_ld * 2)); DG(cudaMalloc(&al, (size_t)a_rp * a_ld * 2));
    DG(cudaMalloc(&bh, (size_t)b_rp * b_ld * 2)); DG(cudaMalloc(&bl, (size_t)b_rp * b_ld * 2));
    DG(cudaMemset(ah, 0, (size_t)a_rp * a_ld * 2)); DG(cudaMemset(al, 0, (size_t)a_rp * a_ld * 2));
    DG(cudaMemset(bh, 0, (size_t)b_rp * b_ld * 2)); DG(cudaMemset(bl, 0, (size_t)b_rp * b_ld * 2));
    DG(cudaMemset(dD, 0xFF, (size_t)Ip * Jp * 4));
    DG(cudaMemcpy(dA, A, (size_t)a_rows * a_cols * 4, cudaMemcpyHostToDevice));
    DG(cudaMemcpy(dB, B, (size_t)b_rows * b_cols * 4, cudaMemcpyHostToDevice));
    launch_split_rows(dA, a_rows, a_cols, ah, al, a_ld, 0);
    launch_split_rows(dB, b_rows, b_cols, bh, bl, b_ld, 0);
    GemmPlan p;
    memset(&p, 0, sizeof p);
    p.bn = bn; p.a_mn = a_mn; p.b_mn = b_mn; p.epi = EPI_STORE_F32; p.splits = splits;
    p.tiles_i = Ip / 128; p.tiles_j = Jp / bn;
    rc = make_tmap_bf16(&p.a_hi, ah, a_rp, a_ld, a_ld, a_mn ? 64 : 128);
    if (!rc) rc = make_tmap_bf16(&p.a_lo, al, a_rp, a_ld, a_ld, a_mn ? 64 : 128);
    if (!rc) rc = make_tmap_bf16(&p.b_hi, bh, b_rp, b_ld, b_ld, b_mn ? 64 : bn);
    if (!rc) rc = make_tmap_bf16(&p.b_lo, bl, b_rp, b_ld, b_ld, b_mn ? 64 : bn);
    p.args.I = Ip; p.args.J = Jp; p.args.kblocks = Rp / 64; p.args.o32 = dD; p.args.ld32 = Jp;
    if (!rc) rc = launch_gemm_tc(p, 0);
    if (rc) { cleanup(); return rc; }
    DG(cudaDeviceSynchronize());
    if (g_gemm_reps > 1) {   // warm timing: back-to-back launches between two events
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, 0);
        for (int i = 0; i < g_gemm_reps && !rc; i++) rc = launch_gemm_tc(p, 0);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&g_gemm_ms, e0, e1);
        g_gemm_ms /= g_gemm_reps;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (g_gemm_trace) { p.args.trace = g_gemm_trace; if (!rc) rc = launch_gemm_tc(p, 0); cudaDeviceSynchronize(); }   // traced launch, warm
        if (rc) { cleanup(); return rc; }
    }
    DG(cudaMemcpy2D(D, (size_t)J * 4, dD, (size_t)Jp * 4, (size_t)J * 4, I, cudaMemcpyDeviceToHost));
#undef DG
    cleanup();
    return GGD_OK;
}

}  // extern "C"
