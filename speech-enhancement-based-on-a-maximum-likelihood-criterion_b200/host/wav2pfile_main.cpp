// Wav2Pfile: raw 16 kHz PCM utterances -> QuickNet pfile (+ .norm) in one pass on the GPU.
// Replaces the chain  Wav2LPS_be (per utterance) -> feacat -ipformat htk + pfile_concat (tools_pfile/pfile_noisy.pl:33,45)
// -> qnnorm (tools_pfile/get_norm.pl:4): the LPS kernel writes pfile records directly (LPS_FLAG_PFILE) and accumulates the
// per-bin statistics of the .norm file on the device (LPS_FLAG_ACCUM_NORM); the host only reads PCM and writes the container.
//
//   Wav2Pfile [-gpu n] [-exact] [-swap] [-norm out.norm] -o out.pfile (-S list.scp | in1.raw in2.raw ...)
#include "../../include/lps_b200.h"
#include "pfile_writer.h"
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static long file_samples(const std::string &p)
{
    FILE *f = fopen(p.c_str(), "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    const long b = ftell(f);
    fclose(f);
    return b / 2;
}

int main(int argc, char **argv)
{
    std::vector<std::string> in;
    const char *out = nullptr, *norm = nullptr, *scp = nullptr;
    int gpu = 0;
    bool exact = false, swap = false;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-gpu") && i + 1 < argc) gpu = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-exact")) exact = true;
        else if (!strcmp(argv[i], "-swap")) swap = true;
        else if (!strcmp(argv[i], "-norm") && i + 1 < argc) norm = argv[++i];
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "-S") && i + 1 < argc) scp = argv[++i];
        else if (argv[i][0] == '-') { fprintf(stderr, "WARNING:  Un-recognized flag '%s' !\n", argv[i]); }
        else in.push_back(argv[i]);
    }
    if (scp) {
        FILE *f = fopen(scp, "rt");
        if (!f) { fprintf(stderr, "ERROR:   Could not open list '%s' !\n", scp); return 1; }
        char line[4096];
        while (fgets(line, sizeof line, f)) {
            std::string s(line);
            while (!s.empty() && (s.back() == '\n' || s.back() == '\r' || s.back() == ' ')) s.pop_back();
            if (!s.empty()) in.push_back(s);
        }
        fclose(f);
    }
    if (!out || in.empty()) { fprintf(stderr, "USAGE:   %s [-gpu n] [-exact] [-swap] [-norm out.norm] -o out.pfile (-S list.scp | in.raw ...)\n", argv[0]); return 1; }
    // sentences = utterances with at least one frame, in list order
    std::vector<std::string> kept;
    std::vector<long> samples, frames;
    for (const auto &p : in) {
        const long n = file_samples(p);
        if (n < 0) { fprintf(stderr, "ERROR:   Could not open file '%s' !\n", p.c_str()); return 1; }
        const long nf = lps_nframes(n);
        if (nf == 0) { fprintf(stderr, "WARNING:  '%s' is shorter than one frame and is skipped\n", p.c_str()); continue; }
        kept.push_back(p); samples.push_back(n); frames.push_back(nf);
    }
    if (kept.empty()) { fprintf(stderr, "ERROR:   no utterance holds a frame\n"); return 1; }
    bphost::PfileWriter w;
    if (!w.open(out, frames, LPS_BINS)) { fprintf(stderr, "ERROR:   Could not open file '%s' !\n", out); return 1; }
    lps_handle *h = nullptr;
    if (lps_create(gpu, &h) != 0) { fprintf(stderr, "ERROR:   %s\n", lps_last_error()); return 1; }
    lps_norm_reset(h);
    const int flags = LPS_FLAG_PFILE | (norm ? LPS_FLAG_ACCUM_NORM : 0) | (exact ? LPS_FLAG_EXACT : 0);
    // groups of utterances of ~2 M frames (2 GB of records) per library call
    const long GROUP_FRAMES = 2000000;
    std::vector<int16_t> pcm;
    std::vector<uint32_t> rec;
    std::vector<long> off;
    size_t u = 0;
    while (u < kept.size()) {
        size_t v = u;
        long gf = 0, gs = 0;
        while (v < kept.size() && (v == u || gf + frames[v] <= GROUP_FRAMES)) { gf += frames[v]; gs += samples[v]; v++; }
        pcm.resize((size_t)gs);
        off.assign(1, 0);
        long pos = 0;
        for (size_t k = u; k < v; k++) {
            FILE *f = fopen(kept[k].c_str(), "rb");
            if (!f || fread(pcm.data() + pos, 2, (size_t)samples[k], f) != (size_t)samples[k]) { fprintf(stderr, "ERROR:   short read on '%s'\n", kept[k].c_str()); return 1; }
            fclose(f);
            pos += samples[k];
            off.push_back(pos);
        }
        if (swap) for (auto &s : pcm) s = (int16_t)(((uint16_t)s << 8) | ((uint16_t)s >> 8));
        rec.resize((size_t)gf * (LPS_BINS + 2));
        long total = 0;
        if (lps_extract_batch(h, pcm.data(), off.data(), (int)(v - u), reinterpret_cast<float *>(rec.data()), flags, &total) != 0 || total != gf) {
            fprintf(stderr, "ERROR:   %s\n", lps_last_error()); return 1;
        }
        if (u > 0)      // the kernel numbers the sentences of one call from 0: shift to the global sentence index
            for (long fr = 0; fr < gf; fr++) {
                uint32_t &sw = rec[(size_t)fr * (LPS_BINS + 2)];
                sw = __builtin_bswap32(__builtin_bswap32(sw) + (uint32_t)u);
            }
        if (!w.append(rec.data(), gf)) { fprintf(stderr, "ERROR:   write to '%s' failed\n", out); return 1; }
        u = v;
    }
    if (!w.close()) { fprintf(stderr, "ERROR:   write to '%s' failed\n", out); return 1; }
    if (norm) {
        float mean[LPS_BINS], dvar[LPS_BINS];
        long n = 0;
        if (lps_norm_finalize(h, mean, dvar, &n) != 0 || !bphost::write_norm_file(norm, mean, dvar, LPS_BINS)) { fprintf(stderr, "ERROR:   %s\n", lps_last_error()); return 1; }
    }
    lps_destroy(h);
    fprintf(stderr, "Processed: %ld Frames of %zu utterances.\n", w.frames_written(), kept.size());
    return 0;
}
