/*
 * ggd_oracle.c -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * Plain-C fp32 restatement of the reference's BPtrain_Sigmoid device path:
 *   BP_GPU::train               Train_code_ML_GGD/BP_GPU.cu:152-185
 *   BP_GPU::train_bunch_single  Train_code_ML_GGD/BP_GPU.cu:308-440
 *   BP_GPU::cv_bunch_single     Train_code_ML_GGD/BP_GPU.cu:442-512
 *   CrossValid / CrossValiddB / CrossValid2 / Gamma   BP_GPU.cu:187-306, 593-640
 * and of the DevFunc.cu kernels those call (cited at each function).
 *
 * Pinning: the reference ships NO golden vectors for the training step
 * (SURVEY.md section 4), so this restatement is pinned against the reference's own
 * CUDA binary (oracle/_ref/BPtrain_ref, built from /root/reference by
 * oracle/build_ref.sh, and libref_bpgpu.so = BP_GPU.cu + DevFunc.cu behind a C shim) run on
 * a B200: tests/test_vs_reference_gpu.py; an independent float64 derivation pins the
 * arithmetic on the CPU (tests/test_oracle_cpu.py).
 *
 * Layout conventions (identical to the reference):
 *   activations  row-major [frame][unit]
 *   weights W_l  index = out + in * cur_units   (l = 1..numlayers-1)
 * All arithmetic is float; summation orders follow the reference kernels where
 * the reference fixes one (column sums, bias sums, CV accumulators); the GEMM
 * summation order (cuBLAS, unspecified) is sequential-k here.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define GGD_MAXLAYER 10

typedef struct {
    int   numlayers;                 /* BP_GPU.h:58 */
    int   layersizes[GGD_MAXLAYER];  /* BP_GPU.h:59 */
    int   bunchsize;
    float lrate, momentum, weightcost, shapefactor;
    int   MLflag;
    /* state */
    float *W[GGD_MAXLAYER], *b[GGD_MAXLAYER];     /* 1..numlayers-1 */
    float *dW[GGD_MAXLAYER], *db[GGD_MAXLAYER];   /* momentum buffers, zero at creation (BP_GPU.cu:91-92,536-541) */
    float *x[GGD_MAXLAYER], *y[GGD_MAXLAYER], *dedy[GGD_MAXLAYER], *dedx[GGD_MAXLAYER];
    float *ydedx[GGD_MAXLAYER], *sumdedx[GGD_MAXLAYER];
    float *out, *scalefactor, *vec1, *vec2, *realerror, *errabs, *errabs2, *newobj;
    /* trace of the last bunch (oracle-defined "loss curve") */
    float last_loss;
    int   world;        /* >1: emulate frame-sharded data parallelism (see oracle_train_bunch_sharded) */
} ggd_oracle;

static float *zalloc(size_t n) { float *p = (float *)calloc(n ? n : 1, sizeof(float)); return p; }

ggd_oracle *ggd_oracle_create(int numlayers, const int *layersizes, int bunchsize, float lrate,
                              float momentum, float weightcost, float shapefactor, int MLflag,
                              const float *const *W, const float *const *b)
{
    ggd_oracle *o = (ggd_oracle *)calloc(1, sizeof(ggd_oracle));
    o->numlayers = numlayers;
    for (int i = 0; i < numlayers; i++) o->layersizes[i] = layersizes[i];
    o->bunchsize = bunchsize; o->lrate = lrate; o->momentum = momentum; o->weightcost = weightcost;
    o->shapefactor = shapefactor; o->MLflag = MLflag; o->world = 1;
    int od = layersizes[numlayers - 1];
    for (int l = 1; l < numlayers; l++) {
        size_t nw = (size_t)layersizes[l] * layersizes[l - 1];
        o->W[l] = zalloc(nw); o->dW[l] = zalloc(nw); o->ydedx[l] = zalloc(nw);
        o->b[l] = zalloc(layersizes[l]); o->db[l] = zalloc(layersizes[l]); o->sumdedx[l] = zalloc(layersizes[l]);
        memcpy(o->W[l], W[l], nw * sizeof(float));
        memcpy(o->b[l], b[l], layersizes[l] * sizeof(float));
        size_t na = (size_t)bunchsize * layersizes[l];
        o->x[l] = zalloc(na); o->y[l] = zalloc(na); o->dedy[l] = zalloc(na); o->dedx[l] = zalloc(na);
    }
    size_t no = (size_t)bunchsize * od;
    o->out = zalloc(no); o->realerror = zalloc(no); o->errabs = zalloc(no); o->errabs2 = zalloc(no); o->newobj = zalloc(no);
    o->scalefactor = zalloc(od); o->vec1 = zalloc(od); o->vec2 = zalloc(od);
    return o;
}

void ggd_oracle_destroy(ggd_oracle *o)
{
    if (!o) return;
    for (int l = 1; l < o->numlayers; l++) {
        free(o->W[l]); free(o->dW[l]); free(o->ydedx[l]); free(o->b[l]); free(o->db[l]); free(o->sumdedx[l]);
        free(o->x[l]); free(o->y[l]); free(o->dedy[l]); free(o->dedx[l]);
    }
    free(o->out); free(o->realerror); free(o->errabs); free(o->errabs2); free(o->newobj);
    free(o->scalefactor); free(o->vec1); free(o->vec2);
    free(o);
}

/* ---- GEMMs: restatement of the three cublasSgemm wrappers, DevFunc.h:49-87 ---------------- */

/* SgemmNN (DevFunc.h:65-75) as called at BP_GPU.cu:361,494: x[m][o] += sum_i W[o + i*cur] * yprev[m][i]
 * (cuBLAS alpha=1, beta=1 on top of the bias broadcast done by kernMultiCopy, DevFunc.cu:134-165). */
__attribute__((target_clones("avx2", "default")))
static void fwd_gemm(int M, int prev, int cur, const float *W, const float *yprev, const float *bias, float *x)
{
#pragma omp parallel for schedule(static)
    for (int m = 0; m < M; m++) {
        float *xr = x + (size_t)m * cur;
        for (int o = 0; o < cur; o++) xr[o] = bias[o];
        const float *yr = yprev + (size_t)m * prev;
        for (int i = 0; i < prev; i++) {
            const float a = yr[i];
            const float *wr = W + (size_t)i * cur;
            for (int o = 0; o < cur; o++) xr[o] += wr[o] * a;
        }
    }
}

/* SgemmTN (DevFunc.h:49-63) as called at BP_GPU.cu:430: dedy_prev[m][i] = sum_o W[o + i*cur] * dedx[m][o] */
__attribute__((target_clones("avx2", "default")))
static void dx_gemm(int M, int prev, int cur, const float *W, const float *dedx, float *dedy_prev)
{
#pragma omp parallel for schedule(static)
    for (int m = 0; m < M; m++) {
        const float *dr = dedx + (size_t)m * cur;
        float *pr = dedy_prev + (size_t)m * prev;
        for (int i = 0; i < prev; i++) {
            const float *wr = W + (size_t)i * cur;
            float s = 0.0f;
            for (int o = 0; o < cur; o++) s += wr[o] * dr[o];
            pr[i] = s;
        }
    }
}

/* SgemmNT (DevFunc.h:77-87) as called at BP_GPU.cu:432: ydedx[o + i*cur] = sum_m dedx[m][o] * yprev[m][i] */
__attribute__((target_clones("avx2", "default")))
static void dw_gemm(int M, int prev, int cur, const float *dedx, const float *yprev, float *g)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < prev; i++) {
        float *gr = g + (size_t)i * cur;
        for (int o = 0; o < cur; o++) gr[o] = 0.0f;
        for (int m = 0; m < M; m++) {
            const float a = yprev[(size_t)m * prev + i];
            const float *dr = dedx + (size_t)m * cur;
            for (int o = 0; o < cur; o++) gr[o] += dr[o] * a;
        }
    }
}

/* kernSigmoid, DevFunc.cu:36-51 (non-RELU branch: the Makefile typo means -DRELU is never applied) */
static void sigmoid_vec(size_t n, const float *x, float *y)
{
    for (size_t i = 0; i < n; i++) y[i] = 1.0f / (1.0f + expf(-x[i]));
}

/* forward pass shared by train and CV: BP_GPU.cu:334-369 and :472-508 */
static void forward(ggd_oracle *o, int M, const float *in, float *out)
{
    const int L = o->numlayers;
    for (int l = 1; l < L; l++) {
        const float *yprev = (l == 1) ? in : o->y[l - 1];
        fwd_gemm(M, o->layersizes[l - 1], o->layersizes[l], o->W[l], yprev, o->b[l], o->x[l]);
        if (l != L - 1) sigmoid_vec((size_t)M * o->layersizes[l], o->x[l], o->y[l]);
        else memcpy(out, o->x[l], (size_t)M * o->layersizes[l] * sizeof(float));  /* cudaMemcpy D2D, :367 */
    }
}

/* The loss-gradient chain at the top layer, BP_GPU.cu:408-424.
 * Mg = frames in the GLOBAL minibatch (== M on one GPU); partial column sums of the other
 * shards (if any) are passed in `extra_colsum` (NULL on one GPU). */
static void loss_gradient(ggd_oracle *o, int M, int Mg, const float *targ, const float *extra_colsum, float *colsum_out)
{
    const int D = o->layersizes[o->numlayers - 1];
    const float beta = o->shapefactor;
    float *dedx = o->dedx[o->numlayers - 1];
    /* DevSubClean2 (DevFunc.cu:376-398) then DevVecMulNum by 1.0f/n_frames (DevFunc.cu:287-293) */
    const float invM = 1.0f / Mg;
    double loss = 0.0;
    for (int m = 0; m < M; m++)
        for (int d = 0; d < D; d++) {
            const float a = o->out[(size_t)m * D + d], t = targ[(size_t)m * D + d];
            float r;
            if (a > t) r = beta * powf(a - t, beta - 1);
            else if (a == t) r = 0;
            else r = -beta * powf(t - a, beta - 1);
            dedx[(size_t)m * D + d] = r * invM;
            loss += pow(fabs((double)a - (double)t), (double)beta);
        }
    o->last_loss = (float)(loss / Mg);   /* E_beta = sum |e|^beta / M    */
    if (o->MLflag != 1) return;
    /* Deverror (DevFunc.cu:399-409), Devabsolutevalus (:186-191), Devindex2 (:219-227) */
    for (size_t i = 0; i < (size_t)M * D; i++) {
        o->realerror[i] = o->out[i] - targ[i];
        o->errabs[i] = fabsf(o->realerror[i]);
        o->errabs2[i] = powf(o->errabs[i], beta);
    }
    /* DevSumcol (DevFunc.cu:167-185): top = row0; then += rows 1.. sequentially */
    for (int d = 0; d < D; d++) {
        float s = o->errabs2[d];
        for (int m = 1; m < M; m++) s += o->errabs2[(size_t)m * D + d];
        o->vec1[d] = s;
    }
    if (colsum_out) memcpy(colsum_out, o->vec1, D * sizeof(float));
    if (extra_colsum) for (int d = 0; d < D; d++) o->vec1[d] += extra_colsum[d];
    /* DevDivide by n_frames (:445-450), DevVecMulNum by beta, Devindex2 with 1.0f/beta  (BP_GPU.cu:417-420) */
    const float ppp = 1.0f / beta;
    for (int d = 0; d < D; d++) {
        o->vec1[d] = o->vec1[d] / (float)Mg;
        o->vec2[d] = o->vec1[d] * beta;
        o->scalefactor[d] = powf(o->vec2[d], ppp);
    }
    /* Devfunc2 (DevFunc.cu:468-489) then DevVecMulNum by 1.0f/n_frames into dedx */
    for (int m = 0; m < M; m++)
        for (int d = 0; d < D; d++) {
            const float e = o->realerror[(size_t)m * D + d];
            float r;
            if (e > 0) r = powf(e, beta - 1.0f) * beta / powf(o->scalefactor[d], beta);
            else if (e == 0) r = 0;
            else r = -powf(-e, beta - 1.0f) * beta / powf(o->scalefactor[d], beta);
            o->newobj[(size_t)m * D + d] = r;
            dedx[(size_t)m * D + d] = r * invM;
        }
}

/* oracle-defined GGD loss  E = sum_d ln alpha_d + sum |e/alpha|^beta / M  (README.md:97) */
static float ggd_loss(ggd_oracle *o, int M, int Mg)
{
    const int D = o->layersizes[o->numlayers - 1];
    double s = 0.0;
    for (int d = 0; d < D; d++) s += log((double)o->scalefactor[d]);
    double q = 0.0;
    for (int m = 0; m < M; m++)
        for (int d = 0; d < D; d++)
            q += pow(fabs((double)o->realerror[(size_t)m * D + d]) / (double)o->scalefactor[d], (double)o->shapefactor);
    return (float)(s + q / Mg);
}

/* backward + momentum-SGD, BP_GPU.cu:371-438. Mg = global frames (divisor in kernUpdatedelta),
 * extra_g / extra_sum: optional gradient contributions of other shards (already summed), per layer. */
static void backward_update(ggd_oracle *o, int M, int Mg, const float *in,
                            float *const *extra_g, float *const *extra_sum, int apply_update)
{
    const int L = o->numlayers;
    for (int l = L - 1; l > 0; l--) {
        const int cur = o->layersizes[l], prev = o->layersizes[l - 1];
        const float *yprev = (l == 1) ? in : o->y[l - 1];
        if (l != L - 1) {   /* DevDsigmoid, DevFunc.cu:53-71 */
            const float *y = o->y[l], *dy = o->dedy[l];
            float *dx = o->dedx[l];
            for (size_t i = 0; i < (size_t)M * cur; i++) dx[i] = (1.0f - y[i]) * y[i] * dy[i];
        }
        if (l != 1) dx_gemm(M, prev, cur, o->W[l], o->dedx[l], o->dedy[l - 1]);   /* pre-update W */
        dw_gemm(M, prev, cur, o->dedx[l], yprev, o->ydedx[l]);
        /* DevAccSumrow (DevFunc.cu:267-285): alpha=0, beta=1, sequential over frames */
        for (int u = 0; u < cur; u++) {
            float s = o->sumdedx[l][u] * 0.0f + 1.0f * o->dedx[l][u];
            for (int m = 1; m < M; m++) s += 1.0f * o->dedx[l][(size_t)m * cur + u];
            o->sumdedx[l][u] = s;
        }
        if (extra_g && extra_g[l]) {
            const size_t nw = (size_t)cur * prev;
            for (size_t i = 0; i < nw; i++) o->ydedx[l][i] += extra_g[l][i];
            for (int u = 0; u < cur; u++) o->sumdedx[l][u] += extra_sum[l][u];
        }
        if (!apply_update) continue;
        /* kernUpdatedelta (DevFunc.cu:490-507) + kernAccSum (:427-443) */
        const size_t nw = (size_t)cur * prev;
        float *W = o->W[l], *dW = o->dW[l];
        const float *g = o->ydedx[l];
        const float mom = o->momentum, lr = o->lrate, wc = o->weightcost;
        for (size_t i = 0; i < nw; i++) {
            dW[i] = mom * dW[i] - lr * (g[i] / Mg + wc * W[i]);
            W[i] = dW[i] + 1.0f * W[i];
        }
        for (int u = 0; u < cur; u++) {
            o->db[l][u] = mom * o->db[l][u] - lr * (o->sumdedx[l][u] / Mg + 0.0f * o->b[l][u]);
            o->b[l][u] = o->db[l][u] + 1.0f * o->b[l][u];
        }
    }
}

/* BP_GPU::train_bunch_single, BP_GPU.cu:308-440 */
void ggd_oracle_train_bunch(ggd_oracle *o, int M, const float *in, const float *targ)
{
    forward(o, M, in, o->out);
    loss_gradient(o, M, M, targ, NULL, NULL);
    if (o->MLflag == 1) o->last_loss = ggd_loss(o, M, M);
    backward_update(o, M, M, in, NULL, NULL, 1);
}

/* BP_GPU::train, BP_GPU.cu:152-185: a trailing bunch smaller than bunchsize is dropped.
 * losses (optional) receives one value per processed bunch; returns the number of bunches processed. */
int ggd_oracle_train(ggd_oracle *o, int n_frames, const float *in, const float *targ, float *losses, float *alphas)
{
    const int n_in = o->layersizes[0], D = o->layersizes[o->numlayers - 1];
    int nb = 0;
    for (int i = 0; i < n_frames; i += o->bunchsize) {
        int f = (o->bunchsize > n_frames - i) ? (n_frames - i) : o->bunchsize;
        if (f == o->bunchsize) {
            ggd_oracle_train_bunch(o, f, in + (size_t)i * n_in, targ + (size_t)i * D);
            if (losses) losses[nb] = o->last_loss;
            if (alphas) memcpy(alphas + (size_t)nb * D, o->scalefactor, D * sizeof(float));
            nb++;
        }
    }
    return nb;
}

/* Frame-sharded data-parallel bunch emulated on one host: the global minibatch of
 * Mg = world*Ms frames is split into `world` contiguous shards, the per-dimension sum |e|^beta and
 * the weight/bias gradients are summed over shards (what NCCL allreduce does), alpha and the
 * update use Mg.  Used to check the N>1 path against the unsharded bunch. */
void ggd_oracle_train_bunch_sharded(ggd_oracle *o, int world, int Ms, const float *in, const float *targ)
{
    const int L = o->numlayers, D = o->layersizes[L - 1], n_in = o->layersizes[0];
    const int Mg = world * Ms;
    float *colsum = zalloc((size_t)world * D), *tot = zalloc(D);
    float *outs = zalloc((size_t)Mg * D);
    /* pass 1: per-shard forward and partial column sums */
    for (int r = 0; r < world; r++) {
        forward(o, Ms, in + (size_t)r * Ms * n_in, o->out);
        memcpy(outs + (size_t)r * Ms * D, o->out, (size_t)Ms * D * sizeof(float));
        if (o->MLflag == 1) loss_gradient(o, Ms, Mg, targ + (size_t)r * Ms * D, NULL, colsum + (size_t)r * D);
    }
    for (int r = 0; r < world; r++) for (int d = 0; d < D; d++) tot[d] += colsum[(size_t)r * D + d];
    /* pass 2: per-shard gradient with the global alpha; accumulate gradients over shards */
    float *accg[GGD_MAXLAYER] = {0}, *accs[GGD_MAXLAYER] = {0};
    for (int l = 1; l < L; l++) { accg[l] = zalloc((size_t)o->layersizes[l] * o->layersizes[l - 1]); accs[l] = zalloc(o->layersizes[l]); }
    float *others = zalloc(D);
    for (int r = 0; r < world; r++) {
        const float *inr = in + (size_t)r * Ms * n_in, *tr = targ + (size_t)r * Ms * D;
        forward(o, Ms, inr, o->out);
        for (int d = 0; d < D; d++) others[d] = tot[d] - colsum[(size_t)r * D + d];
        loss_gradient(o, Ms, Mg, tr, o->MLflag == 1 ? others : NULL, NULL);
        if (r < world - 1) {
            backward_update(o, Ms, Mg, inr, NULL, NULL, 0);
            for (int l = 1; l < L; l++) {
                size_t nw = (size_t)o->layersizes[l] * o->layersizes[l - 1];
                for (size_t i = 0; i < nw; i++) accg[l][i] += o->ydedx[l][i];
                for (int u = 0; u < o->layersizes[l]; u++) accs[l][u] += o->sumdedx[l][u];
            }
        } else {
            backward_update(o, Ms, Mg, inr, accg, accs, 1);
        }
    }
    for (int l = 1; l < L; l++) { free(accg[l]); free(accs[l]); }
    free(colsum); free(tot); free(outs); free(others);
}

/* ---- phase-wise entry points used by the multi-process (gloo, world_size 2) CPU tests: they mirror what each
 * rank of the frame-sharded CUDA path does between its two collectives. ------------------------ */
/* phase 1: forward on the local shard, local sum_m |e|^beta per output dimension */
void ggd_oracle_dp_colsum(ggd_oracle *o, int Ms, int Mg, const float *in, const float *targ, float *colsum)
{
    const int D = o->layersizes[o->numlayers - 1];
    forward(o, Ms, in, o->out);
    const int ml = o->MLflag;
    o->MLflag = 1;
    loss_gradient(o, Ms, Mg, targ, NULL, colsum);
    o->MLflag = ml;
    (void)D;
}
/* phase 2: local gradients given the GLOBAL column sums (after allreduce); no update */
void ggd_oracle_dp_backward(ggd_oracle *o, int Ms, int Mg, const float *in, const float *targ, const float *colsum_global, const float *colsum_local)
{
    const int D = o->layersizes[o->numlayers - 1];
    float *others = zalloc(D);
    for (int d = 0; d < D; d++) others[d] = colsum_global[d] - colsum_local[d];
    loss_gradient(o, Ms, Mg, targ, o->MLflag == 1 ? others : NULL, NULL);
    backward_update(o, Ms, Mg, in, NULL, NULL, 0);
    free(others);
}
/* phase 3: momentum-SGD update from the (allreduced) gradients stored back into ydedx / sumdedx */
void ggd_oracle_dp_update(ggd_oracle *o, int Mg)
{
    for (int l = o->numlayers - 1; l > 0; l--) {
        const size_t nw = (size_t)o->layersizes[l] * o->layersizes[l - 1];
        for (size_t i = 0; i < nw; i++) {
            o->dW[l][i] = o->momentum * o->dW[l][i] - o->lrate * (o->ydedx[l][i] / Mg + o->weightcost * o->W[l][i]);
            o->W[l][i] = o->dW[l][i] + 1.0f * o->W[l][i];
        }
        for (int u = 0; u < o->layersizes[l]; u++) {
            o->db[l][u] = o->momentum * o->db[l][u] - o->lrate * (o->sumdedx[l][u] / Mg + 0.0f * o->b[l][u]);
            o->b[l][u] = o->db[l][u] + 1.0f * o->b[l][u];
        }
    }
}
float *ggd_oracle_bias_grad(ggd_oracle *o, int l) { return o->sumdedx[l]; }

/* cv forward for n frames in bunches (partial last bunch IS processed, BP_GPU.cu:203-218) */
void ggd_oracle_forward(ggd_oracle *o, int n_frames, const float *in, float *out)
{
    const int n_in = o->layersizes[0], D = o->layersizes[o->numlayers - 1];
    for (int i = 0; i < n_frames; i += o->bunchsize) {
        int f = (o->bunchsize > n_frames - i) ? (n_frames - i) : o->bunchsize;
        forward(o, f, in + (size_t)i * n_in, out + (size_t)i * D);
    }
}

/* BP_GPU::CrossValid, BP_GPU.cu:187-221: float accumulator, frame-major then dim order */
float ggd_oracle_cv_sqerr(ggd_oracle *o, int n_frames, const float *in, const float *targ)
{
    const int D = o->layersizes[o->numlayers - 1];
    float *out = zalloc((size_t)n_frames * D);
    ggd_oracle_forward(o, n_frames, in, out);
    float s = 0.0f;
    for (size_t i = 0; i < (size_t)n_frames * D; i++) s = s + (out[i] - targ[i]) * (out[i] - targ[i]);
    free(out);
    return s;
}

/* BP_GPU::CrossValiddB, BP_GPU.cu:222-255 (abs() resolves to the float overload under g++/nvcc) */
float ggd_oracle_cv_abserr(ggd_oracle *o, int n_frames, const float *in, const float *targ)
{
    const int D = o->layersizes[o->numlayers - 1];
    float *out = zalloc((size_t)n_frames * D);
    ggd_oracle_forward(o, n_frames, in, out);
    float s = 0.0f;
    for (size_t i = 0; i < (size_t)n_frames * D; i++) s = s + fabsf(out[i] - targ[i]);
    s = s / D;
    free(out);
    return s;
}

/* BP_GPU::Gamma, BP_GPU.cu:593-640 (polynomial on (2,3], recursion elsewhere; double pow, float store) */
float ggd_oracle_gamma(float x)
{
    if (x > 2 && x <= 3) {
        const float c0 = 0.0000677106, c1 = -0.0003442342, c2 = 0.0015397681, c3 = -0.0024467480,
                    c4 = 0.0109736958, c5 = -0.0002109075, c6 = 0.0742379071, c7 = 0.0815782188,
                    c8 = 0.4118402518, c9 = 0.4227843370, c10 = 1.0000000000;
        float temp = 0;
        temp = temp + c0 * pow(x - 2.0, 10.0) + c1 * pow(x - 2.0, 9.0);
        temp = temp + c2 * pow(x - 2.0, 8.0) + c3 * pow(x - 2.0, 7.0);
        temp = temp + c4 * pow(x - 2.0, 6.0) + c5 * pow(x - 2.0, 5.0);
        temp = temp + c6 * pow(x - 2.0, 4.0) + c7 * pow(x - 2.0, 3.0);
        temp = temp + c8 * pow(x - 2.0, 2.0) + c9 * (x - 2.0) + c10;
        return temp;
    } else if (x > 0 && x <= 1) {
        return ggd_oracle_gamma(x + 2) / (x * (x + 1));
    } else if (x > 1 && x <= 2) {
        return ggd_oracle_gamma(x + 1) / x;
    } else if (x > 3) {
        int i = 1;
        float temp = 1;
        while (((x - i) > 2 && (x - i) <= 3) == 0) { temp = (x - i) * temp; i++; }
        temp = temp * (x - i);
        return temp * ggd_oracle_gamma(x - i);
    }
    return 0;
}

/* BP_GPU::CrossValid2, BP_GPU.cu:256-306: GGD log-likelihood with alpha left by the last training bunch */
float ggd_oracle_cv_loglik(ggd_oracle *o, int n_frames, const float *in, const float *targ)
{
    const int D = o->layersizes[o->numlayers - 1];
    float *out = zalloc((size_t)n_frames * D), *err = zalloc((size_t)n_frames * D);
    ggd_oracle_forward(o, n_frames, in, out);
    for (size_t i = 0; i < (size_t)n_frames * D; i++) err[i] = targ[i] - out[i];
    const float sf = o->shapefactor;
    float density1, density2 = 0, density3 = 0;
    density1 = n_frames * D * logf(sf / (2 * ggd_oracle_gamma((float)(1.0 / sf))));
    for (int u = 0; u < D; u++) density2 += logf(o->scalefactor[u]);
    density2 = density2 * n_frames;
    for (int uu = 0; uu < n_frames; uu++)
        for (int d = 0; d < D; d++)
            density3 += powf(fabsf(err[(size_t)uu * D + d]) / o->scalefactor[d], sf);
    free(out); free(err);
    return density1 - density2 - density3;
}

/* accessors for the ctypes wrapper */
float *ggd_oracle_W(ggd_oracle *o, int l) { return o->W[l]; }
float *ggd_oracle_b(ggd_oracle *o, int l) { return o->b[l]; }
float *ggd_oracle_dW(ggd_oracle *o, int l) { return o->dW[l]; }
float *ggd_oracle_alpha(ggd_oracle *o) { return o->scalefactor; }
float *ggd_oracle_out(ggd_oracle *o) { return o->out; }
float *ggd_oracle_dedx(ggd_oracle *o, int l) { return o->dedx[l]; }
float *ggd_oracle_grad(ggd_oracle *o, int l) { return o->ydedx[l]; }
float *ggd_oracle_y(ggd_oracle *o, int l) { return o->y[l]; }
float  ggd_oracle_last_loss(ggd_oracle *o) { return o->last_loss; }
