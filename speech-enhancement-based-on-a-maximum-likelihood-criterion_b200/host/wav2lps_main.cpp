// Wav2LPS_be drop-in: `Wav2LPS_be -F RAW -fs 16 in.raw out.lps` (Feature_prepare/LPS_extract.m:13).
// Writes the reference's HTK big-endian file: 12-byte header {nSamples, 160000, 1028, 9} then
// frames x 257 float32 (Wav2LogSpec_be.c:371-377, 575-576; fileio.c:187-243).  The arithmetic runs in
// the LPS kernel of libggd_b200 (include/lps_b200.h), which returns the payload already byte-swapped.
#include "../../include/lps_b200.h"
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static void usage(const char *a0)
{
    fprintf(stderr, "\r\nUSAGE:   %s infile HTK_outfile [options]\r\n\r\nOPTIONS:\r\n"
                    "     -q            Quiet Mode\r\n     -F    format  Input file format (RAW)\r\n"
                    "     -fs   freq    Sampling frequency in kHz (16)\r\n     -swap         Change input byte ordering\r\n"
                    "     -gpu  n       CUDA device (default 0)\r\n"
                    "     -fast         register-resident FFT kernel (features within 1e-4 relative of the reference) instead of\r\n"
                    "                   the reference's own butterfly network (default: byte-identical HTK files)\r\n", a0);
}

int main(int argc, char **argv)
{
    const char *in = nullptr, *out = nullptr;
    bool quiet = false, swap = false, fast = false;
    int gpu = 0, fs = 16;
    std::string fmt = "RAW";
    for (int i = 1; i < argc; i++) {          // ParseCommLine, Wav2LogSpec_be.c:127-259
        if (!strcmp(argv[i], "-q")) quiet = true;
        else if (!strcmp(argv[i], "-swap")) swap = true;
        else if (!strcmp(argv[i], "-fast")) fast = true;
        else if (!strcmp(argv[i], "-F") && i + 1 < argc) fmt = argv[++i];
        else if (!strcmp(argv[i], "-fs") && i + 1 < argc) fs = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-gpu") && i + 1 < argc) gpu = atoi(argv[++i]);
        else if (argv[i][0] == '-') { fprintf(stderr, "WARNING:  Un-recognized flag '%s' !\r\n", argv[i]); }
        else if (!in) in = argv[i];
        else if (!out) out = argv[i];
    }
    if (!in || !out) { usage(argv[0]); return 1; }
    if (fmt != "RAW") { fprintf(stderr, "ERROR:   only -F RAW is supported by this build (NIST/HTK inputs are outside the named path)\r\n"); return 1; }
    if (fs != 16) { fprintf(stderr, "ERROR:   Invalid sampling frequency '%d'! (this build covers the 16 kHz configuration)\r\n", fs * 1000); return 1; }
    if (!quiet) fprintf(stderr, "\r\nDSR Front-End v2.0 (B200)\r\n");
    FILE *fi = fopen(in, "rb");
    if (!fi) { fprintf(stderr, "ERROR:   Could not open file '%s' !\r\n", in); return 1; }
    fseek(fi, 0, SEEK_END);
    const long bytes = ftell(fi);
    fseek(fi, 0, SEEK_SET);
    std::vector<int16_t> pcm(bytes / 2);
    if (fread(pcm.data(), 2, pcm.size(), fi) != pcm.size()) { fprintf(stderr, "ERROR:   short read on '%s'\r\n", in); return 1; }
    fclose(fi);
    if (swap) for (auto &s : pcm) s = (int16_t)(((uint16_t)s << 8) | ((uint16_t)s >> 8));
    const long nf = lps_nframes((long)pcm.size());
    std::vector<float> feat((size_t)nf * LPS_BINS);
    lps_handle *h = nullptr;
    if (lps_create(gpu, &h) != 0) { fprintf(stderr, "ERROR:   %s\r\n", lps_last_error()); return 1; }
    if (lps_extract(h, pcm.data(), (long)pcm.size(), feat.data(), LPS_FLAG_BIG_ENDIAN | (fast ? 0 : LPS_FLAG_EXACT)) != 0) { fprintf(stderr, "ERROR:   %s\r\n", lps_last_error()); return 1; }
    lps_destroy(h);
    FILE *fo = fopen(out, "wb");
    if (!fo) { fprintf(stderr, "ERROR:   Could not open file '%s' !\r\n", out); return 1; }
    const uint32_t hdr32[2] = {__builtin_bswap32((uint32_t)nf), __builtin_bswap32(160000u)};
    const uint16_t hdr16[2] = {__builtin_bswap16((uint16_t)(LPS_BINS * 4)), __builtin_bswap16(9)};
    fwrite(hdr32, 4, 2, fo); fwrite(hdr16, 2, 2, fo);
    fwrite(feat.data(), 4, feat.size(), fo);
    fclose(fo);
    if (!quiet) fprintf(stderr, "\rProcessed: %ld Frames.                      \r\n", nf);
    return 0;
}
