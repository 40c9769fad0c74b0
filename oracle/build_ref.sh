#!/bin/bash
# Builds the UNMODIFIED reference sources (where they lie under /root/reference) into oracle/_ref/.
# Outputs only (git-ignored): oracle/_ref/Wav2LPS_be_ref, oracle/_ref/libref_interface.so,
# oracle/_ref/BPtrain_ref.  Nothing is copied into the repository; the only edit is an
# in-flight sed of four hard-coded "/usr/local/cuda-9.0/include" include paths
# (BP_GPU.h:3-5, DevFunc.h:3) applied to a temporary directory that is deleted afterwards.
set -e
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "no reference at $REF: skipping oracle/_ref build"; exit 0; }
mkdir -p "$OUT"
# 1. LPS front end: three C files, libm only (reference makefile flags: -std=c99, no -O)
L="$REF/Feature_prepare/SourceCode_Wav2LogSpec_be"
/usr/bin/gcc -std=c99 -O2 -w -o "$OUT/Wav2LPS_be_ref" "$L/Wav2LogSpec_be.c" "$L/FEfunc.c" "$L/fileio.c" -lm
/usr/bin/gcc -std=c99 -w -o "$OUT/Wav2LPS_be_ref_O0" "$L/Wav2LogSpec_be.c" "$L/FEfunc.c" "$L/fileio.c" -lm
# 2/3. trainer: needs the include-path patch -> temp copy, removed on exit
T="$REF/Train_code_ML_GGD"
TMP="$(mktemp -d "$OUT/.build.XXXXXX")"
trap 'rm -rf "$TMP"' EXIT
for f in BP_GPU.h DevFunc.h Interface.h Interface.cc BPtrain.cc BP_GPU.cu DevFunc.cu; do
  sed 's#/usr/local/cuda-9.0/include/##' "$T/$f" > "$TMP/$f"
done
CUDA=${CUDA_HOME:-/usr/local/cuda}
# 2. host loader only (Interface.cc links without CUDA): used to pin the product loader
cat > "$TMP/shim.cc" <<'EOS'
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Interface.h"
extern "C" {
void *refif_create(int argc, char **argv) { Interface *o = new Interface; o->Initial(argc, argv); o->get_pfile_info(); return o; }
int refif_numlayers(void *p) { return ((Interface *)p)->numlayers; }
void refif_train_info(void *p, const char *range, int *chunks, int *samples) {
  Interface *o = (Interface *)p; char buf[256]; snprintf(buf, sizeof buf, "%s", range); o->get_chunk_info(buf);
  *chunks = o->total_chunks; *samples = o->total_samples; }
void refif_cv_info(void *p, const char *range, int *chunks, int *samples) {
  Interface *o = (Interface *)p; char buf[256]; snprintf(buf, sizeof buf, "%s", range); o->get_chunk_info_cv(buf);
  *chunks = o->cv_total_chunks; *samples = o->cv_total_samples; }
void refif_shuffle_chunks(void *p, int *idx, int n) { ((Interface *)p)->GetRandIndex(idx, n); }
int refif_readchunk(void *p, int idx) { return ((Interface *)p)->Readchunk(idx); }
int refif_readchunk_cv(void *p, int idx) { return ((Interface *)p)->Readchunk_cv(idx); }
float *refif_in(void *p) { return ((Interface *)p)->para->indata[0]; }
float *refif_targ(void *p) { return ((Interface *)p)->para->targ[0]; }
float *refif_W(void *p, int l) { return ((Interface *)p)->para->weights[l]; }
float *refif_b(void *p, int l) { return ((Interface *)p)->para->bias[l]; }
void refif_writeweights(void *p) { ((Interface *)p)->Writeweights(); }
}
EOS
/usr/bin/g++ -O1 -w -fPIC -shared -fpermissive -I"$CUDA/include" -o "$OUT/libref_interface.so" "$TMP/Interface.cc" "$TMP/shim.cc" -lpthread
# 2b. the reference's device path (BP_GPU.cu + DevFunc.cu, cuBLAS/cuRAND) as a shared library with a C shim,
#     so that tests and bench.py can drive the UNMODIFIED reference CUDA code on identical in-memory inputs.
cat > "$TMP/shim_bp.cu" <<'EOS'
#include <stdio.h>
#include <stdlib.h>
#include "BP_GPU.h"
extern "C" {
void *refbp_create(int seed, int gpu, int numlayers, int *layersizes, int bunchsize, float lrate, float momentum, float weightcost,
                   float **weights, float **bias, float shapefactor, int MLflag) {
  BP_GPU *o = new BP_GPU(seed, gpu, numlayers, layersizes, bunchsize, lrate, momentum, weightcost, weights, bias, shapefactor, MLflag, 0, 0.0f, 0.0f);
  cudaDeviceSynchronize();
  return o; }
void refbp_train(void *p, int n, float *in, const float *targ) { ((BP_GPU *)p)->train(n, in, targ); cudaDeviceSynchronize(); }
float refbp_cv(void *p, int which, int n, const float *in, const float *targ) {
  BP_GPU *o = (BP_GPU *)p; float r = which == 0 ? o->CrossValid(n, in, targ) : which == 1 ? o->CrossValiddB(n, in, targ) : o->CrossValid2(n, in, targ);
  return r; }
void refbp_weights(void *p, float **weights, float **bias) { ((BP_GPU *)p)->returnWeights(weights, bias); cudaDeviceSynchronize(); }
void refbp_destroy(void *p) { delete (BP_GPU *)p; }
}
EOS
if command -v nvcc >/dev/null; then
  nvcc -w -O2 -gencode arch=compute_100a,code=sm_100a -I"$CUDA/include" -Xcompiler -fPIC -Xcompiler -fpermissive -shared -cudart shared \
    -o "$OUT/libref_bpgpu.so" "$TMP/shim_bp.cu" "$TMP/BP_GPU.cu" "$TMP/DevFunc.cu" -lcublas -lcurand || echo "reference device library did not build"
fi
# 3. the reference CUDA trainer for sm_100a (second baseline + strongest oracle; runs on the GPU box)
if command -v nvcc >/dev/null; then
  nvcc -w -gencode arch=compute_100a,code=sm_100a -I"$CUDA/include" -Xcompiler -fpermissive \
    -o "$OUT/BPtrain_ref" "$TMP/BPtrain.cc" "$TMP/Interface.cc" "$TMP/BP_GPU.cu" "$TMP/DevFunc.cu" \
    -lcublas -lcurand -lpthread || echo "reference CUDA trainer did not build"
fi
ls -la "$OUT"
