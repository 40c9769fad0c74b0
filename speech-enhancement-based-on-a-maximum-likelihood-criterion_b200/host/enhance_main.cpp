// Enhance_LPS: the network part of Test_code/decode.m as an executable -- noisy LPS (HTK file, as written by Wav2LPS_be) ->
// z-score with the .norm constants (decode.m:31-33) -> context expansion with the first / last frame replicated at the
// utterance edges (frame_expand.m:6-25) -> sigmoid layers + linear output (decode.m:37-58, weights from the MAT-v4 file
// BPtrain_Sigmoid writes) -> de-normalisation (decode.m:60-62) -> enhanced LPS as an HTK file (writeHTK_new, decode.m:63).
// The arithmetic runs on the GPU through ggd_enhance (include/ggd_train.h); resynthesis (LPS2Wav_be) is outside this path.
//
//   Enhance_LPS [-gpu n] [-ctx 7] -wts mlp.wts -norm train_noisy.norm -layers 1799,2048,2048,2048,257 in.lps out.lps [in2 out2 ...]
#include "../../include/ggd_train.h"
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static bool read_norm(const char *path, int dim, std::vector<float> &mean, std::vector<float> &dvar)
{
    FILE *f = fopen(path, "rt");                       // Interface.cc:385-396 / decode.m:6-8: "vec N", N means, "vec N", N reciprocal stds
    if (!f) return false;
    char line[1024];
    mean.assign(dim, 0.f); dvar.assign(dim, 0.f);
    bool ok = fgets(line, sizeof line, f) != nullptr;
    for (int j = 0; j < dim && ok; j++) { ok = fgets(line, sizeof line, f) != nullptr; mean[j] = (float)atof(line); }
    ok = ok && fgets(line, sizeof line, f) != nullptr;
    for (int j = 0; j < dim && ok; j++) { ok = fgets(line, sizeof line, f) != nullptr; dvar[j] = (float)atof(line); }
    fclose(f);
    return ok;
}

// MAT level-4 matrices weights<i><i+1> (rows = L[i], cols = L[i-1], column-major = index out + in*rows) and bias<i+1>,
// as written by Interface::Writeweights (Interface.cc:489-514)
static bool read_wts(const char *path, const std::vector<int> &ls, std::vector<std::vector<float>> &W, std::vector<std::vector<float>> &b)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    const int L = (int)ls.size();
    W.assign(L, {}); b.assign(L, {});
    bool ok = true;
    for (int l = 1; l < L && ok; l++) {
        for (int part = 0; part < 2 && ok; part++) {
            int32_t st[5];
            ok = fread(st, 4, 5, f) == 5 && st[4] > 0 && st[4] < 256;
            if (!ok) break;
            char name[256];
            ok = fread(name, 1, st[4], f) == (size_t)st[4];
            const long rows = st[1], cols = st[2];
            const long want_r = part == 0 ? ls[l] : 1, want_c = part == 0 ? ls[l - 1] : ls[l];
            if (!ok || rows != want_r || cols != want_c) { fprintf(stderr, "ERROR:   matrix %d of layer %d is %ld x %ld, expected %ld x %ld\n", part, l, rows, cols, want_r, want_c); ok = false; break; }
            std::vector<float> &dst = part == 0 ? W[l] : b[l];
            dst.resize((size_t)rows * cols);
            ok = fread(dst.data(), 4, dst.size(), f) == dst.size();
        }
    }
    fclose(f);
    return ok;
}

static bool read_htk(const char *path, int dim, std::vector<float> &x, long *frames, uint32_t *period)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    uint32_t h32[2]; uint16_t h16[2];
    bool ok = fread(h32, 4, 2, f) == 2 && fread(h16, 2, 2, f) == 2;
    const long n = ok ? (long)__builtin_bswap32(h32[0]) : 0;
    ok = ok && __builtin_bswap16(h16[0]) == dim * 4;
    if (ok) {
        std::vector<uint32_t> raw((size_t)n * dim);
        ok = fread(raw.data(), 4, raw.size(), f) == raw.size();
        x.resize(raw.size());
        for (size_t i = 0; i < raw.size() && ok; i++) { const uint32_t w = __builtin_bswap32(raw[i]); memcpy(&x[i], &w, 4); }
        *frames = n; *period = __builtin_bswap32(h32[1]);
    }
    fclose(f);
    return ok;
}

static bool write_htk(const char *path, int dim, const std::vector<float> &x, long frames, uint32_t period)
{
    FILE *f = fopen(path, "wb");
    if (!f) return false;
    const uint32_t h32[2] = {__builtin_bswap32((uint32_t)frames), __builtin_bswap32(period)};
    const uint16_t h16[2] = {__builtin_bswap16((uint16_t)(dim * 4)), __builtin_bswap16(9)};
    std::vector<uint32_t> raw(x.size());
    for (size_t i = 0; i < x.size(); i++) { uint32_t w; memcpy(&w, &x[i], 4); raw[i] = __builtin_bswap32(w); }
    bool ok = fwrite(h32, 4, 2, f) == 2 && fwrite(h16, 2, 2, f) == 2 && fwrite(raw.data(), 4, raw.size(), f) == raw.size();
    return (fclose(f) == 0) && ok;
}

int main(int argc, char **argv)
{
    const char *wts = nullptr, *norm = nullptr;
    int gpu = 0, ctx = 7;
    std::vector<int> ls;
    std::vector<std::string> files;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-gpu") && i + 1 < argc) gpu = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-ctx") && i + 1 < argc) ctx = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-wts") && i + 1 < argc) wts = argv[++i];
        else if (!strcmp(argv[i], "-norm") && i + 1 < argc) norm = argv[++i];
        else if (!strcmp(argv[i], "-layers") && i + 1 < argc) {
            for (const char *q = argv[++i]; *q;) { ls.push_back(atoi(q)); q = strchr(q, ','); if (!q) break; q++; }
        }
        else if (argv[i][0] == '-') fprintf(stderr, "WARNING:  Un-recognized flag '%s' !\n", argv[i]);
        else files.push_back(argv[i]);
    }
    if (!wts || !norm || ls.size() < 2 || ls.size() > GGD_MAXLAYER || files.empty() || files.size() % 2) {
        fprintf(stderr, "USAGE:   %s [-gpu n] [-ctx 7] -wts mlp.wts -norm x.norm -layers 1799,2048,2048,2048,257 in.lps out.lps [in2 out2 ...]\n", argv[0]);
        return 1;
    }
    const int dim = ls.back();
    if (ctx < 1 || (ctx & 1) == 0 || dim * ctx != ls[0]) { fprintf(stderr, "ERROR:   %d context frames x %d bins do not make the %d net inputs\n", ctx, dim, ls[0]); return 1; }
    std::vector<float> mean, dvar;
    if (!read_norm(norm, dim, mean, dvar)) { fprintf(stderr, "ERROR:   Could not read norm file '%s' !\n", norm); return 1; }
    std::vector<std::vector<float>> W, b;
    if (!read_wts(wts, ls, W, b)) { fprintf(stderr, "ERROR:   Could not read weights file '%s' !\n", wts); return 1; }
    ggd_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.numlayers = (int)ls.size();
    for (size_t i = 0; i < ls.size(); i++) cfg.layersizes[i] = ls[i];
    cfg.bunchsize = 128; cfg.shapefactor = 2.0f; cfg.gpu = gpu; cfg.world_size = 1;     // forward only: the training settings are unused
    const float *Wp[GGD_MAXLAYER] = {nullptr}, *bp[GGD_MAXLAYER] = {nullptr};
    for (size_t l = 1; l < ls.size(); l++) { Wp[l] = W[l].data(); bp[l] = b[l].data(); }
    ggd_handle *net = nullptr;
    if (ggd_create(&cfg, Wp, bp, &net) != GGD_OK) { fprintf(stderr, "ERROR:   %s\n", ggd_last_error()); return 1; }
    for (size_t k = 0; k < files.size(); k += 2) {
        std::vector<float> x;
        long frames = 0; uint32_t period = 160000;
        if (!read_htk(files[k].c_str(), dim, x, &frames, &period)) { fprintf(stderr, "ERROR:   Could not read HTK file '%s' (%d-bin float features expected) !\n", files[k].c_str(), dim); return 1; }
        std::vector<float> y((size_t)frames * dim);
        if (frames > 0 && ggd_enhance(net, (int)frames, x.data(), dim, ctx, mean.data(), dvar.data(), y.data()) != GGD_OK) { fprintf(stderr, "ERROR:   %s\n", ggd_last_error()); return 1; }
        if (!write_htk(files[k + 1].c_str(), dim, y, frames, period)) { fprintf(stderr, "ERROR:   Could not write '%s' !\n", files[k + 1].c_str()); return 1; }
        fprintf(stderr, "Enhanced: %ld Frames  %s -> %s\n", frames, files[k].c_str(), files[k + 1].c_str());
    }
    ggd_destroy(net);
    return 0;
}
