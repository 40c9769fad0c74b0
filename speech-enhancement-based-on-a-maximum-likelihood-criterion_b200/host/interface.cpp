// interface.cpp -- see interface.h.  Reference behaviour cited per function (Train_code_ML_GGD/Interface.cc).
#include "interface.h"
#include <cstdarg>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <thread>
#include <unistd.h>

namespace bphost {

// bulk read of [off, off + bytes) with `threads` concurrent pread()s (page cache / NVMe queues are not saturated by one reader)
static bool pread_parallel(int fd, void *buf, size_t bytes, off_t off, int threads)
{
    if (threads < 1) threads = 1;
    const size_t piece = (bytes / threads + 4095) & ~(size_t)4095;
    std::vector<std::thread> ts;
    std::vector<char> ok(threads, 1);
    for (int t = 0; t < threads; t++) {
        const size_t b0 = (size_t)t * piece, b1 = b0 + piece < bytes ? b0 + piece : bytes;
        if (b0 >= bytes) break;
        ts.emplace_back([=, &ok] {
            size_t done = b0;
            while (done < b1) {
                const ssize_t r = pread(fd, (char *)buf + done, b1 - done, off + (off_t)done);
                if (r <= 0) { ok[t] = 0; return; }
                done += (size_t)r;
            }
        });
    }
    for (auto &th : ts) th.join();
    for (char c : ok) if (!c) return false;
    return true;
}

static inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }

Host::~Host()
{
    if (fp_data) fclose(fp_data);
    if (fp_targ) fclose(fp_targ);
    if (fp_out) fclose(fp_out);
    if (fp_log) fclose(fp_log);
}

void Host::logf(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    if (fp_log) { vfprintf(fp_log, fmt, ap); fflush(fp_log); }
    else vfprintf(stderr, fmt, ap);
    va_end(ap);
}

// Interface::Initial, Interface.cc:133-482
bool Host::init(int argc, char **argv)
{
    for (int i = 1; i < argc; i++) {
        std::string a(argv[i]);
        const size_t eq = a.find('=');
        if (eq == std::string::npos) { fprintf(stderr, "Arg: %s  Format Error\n", argv[i]); return false; }   // :152-156
        const std::string k = a.substr(0, eq), v = a.substr(eq + 1);
        if (k == "fea_file") p.fea_file = v;
        else if (k == "norm_file") p.norm_file = v;
        else if (k == "targ_file") p.targ_file = v;
        else if (k == "outwts_file") p.outwts_file = v;
        else if (k == "log_file") p.log_file = v;
        else if (k == "initwts_file") p.initwts_file = v;
        else if (k == "train_sent_range") p.train_sent_range = v;
        else if (k == "cv_sent_range") p.cv_sent_range = v;
        else if (k == "fea_dim") p.fea_dim = atoi(v.c_str());
        else if (k == "fea_context") p.fea_context = atoi(v.c_str());
        else if (k == "targ_offset") p.targ_offset = atoi(v.c_str());
        else if (k == "dropoutflag") p.dropoutflag = atoi(v.c_str());
        else if (k == "MLflag") p.MLflag = atoi(v.c_str());
        else if (k == "traincache") p.traincache = atoi(v.c_str());
        else if (k == "bunchsize") p.bunchsize = atoi(v.c_str());
        else if (k == "gpu_used") {                                             // a comma list selects data parallelism (extension)
            p.gpus.clear();
            size_t pos = 0;
            while (pos <= v.size()) {
                const size_t c = v.find(',', pos);
                p.gpus.push_back(atoi(v.substr(pos, c == std::string::npos ? std::string::npos : c - pos).c_str()));
                if (c == std::string::npos) break;
                pos = c + 1;
            }
            p.gpu_used = p.gpus.empty() ? 0 : p.gpus[0];
        }
        else if (k == "init_randem_seed") p.init_randem_seed = atoi(v.c_str());
        else if (k == "momentum") p.momentum = (float)atof(v.c_str());
        else if (k == "shapefactor") p.shapefactor = (float)atof(v.c_str());
        else if (k == "weightcost") p.weightcost = (float)atof(v.c_str());
        else if (k == "lrate") p.lrate = (float)atof(v.c_str());
        else if (k == "visible_omit") p.visible_omit = (float)atof(v.c_str());
        else if (k == "hid_omit") p.hid_omit = (float)atof(v.c_str());
        else if (k == "init_randem_weight_max") p.init_randem_weight_max = (float)atof(v.c_str());
        else if (k == "init_randem_weight_min") p.init_randem_weight_min = (float)atof(v.c_str());
        else if (k == "init_randem_bias_max") p.init_randem_bias_max = (float)atof(v.c_str());
        else if (k == "init_randem_bias_min") p.init_randem_bias_min = (float)atof(v.c_str());
        else if (k == "layersizes") {                                           // :298-313: the list defines numlayers
            p.numlayers = 0;
            size_t pos = 0;
            while (pos <= v.size() && p.numlayers < kMaxLayer) {
                const size_t c = v.find(',', pos);
                p.layersizes[p.numlayers++] = atoi(v.substr(pos, c == std::string::npos ? std::string::npos : c - pos).c_str());
                if (c == std::string::npos) break;
                pos = c + 1;
            }
        }
        else if (k == "precision") p.precision = (v == "fp32" || v == "1") ? 1 : 0;
        else if (k == "no_graph") p.no_graph = atoi(v.c_str());
        else if (k == "host_loader") p.host_loader = atoi(v.c_str());
        else if (k == "read_threads") p.read_threads = atoi(v.c_str());
        // anything else (e.g. numlayers=) is silently ignored, as in the reference
    }
    if (!(fp_log = fopen(p.log_file.c_str(), "wt"))) { printf("can not open output log file: %s\n", p.log_file.c_str()); return false; }
    if (!(fp_data = fopen(p.fea_file.c_str(), "rb"))) { logf("can not open feature file: %s\n", p.fea_file.c_str()); return false; }
    if (!(fp_targ = fopen(p.targ_file.c_str(), "rb"))) { logf("can not open target file: %s\n", p.targ_file.c_str()); return false; }
    if (!(fp_out = fopen(p.outwts_file.c_str(), "wb"))) { logf("can not open output weights file: %s\n", p.outwts_file.c_str()); return false; }
    // parameter echo, same lines as Interface.cc:338-371
    logf("parameters input:\n");
    logf("fea_file:             %s\n", p.fea_file.c_str());
    logf("norm_file:            %s\n", p.norm_file.c_str());
    logf("targ_file:            %s\n", p.targ_file.c_str());
    logf("outwts_file:          %s\n", p.outwts_file.c_str());
    logf("log_file:\t\t          %s\n", p.log_file.c_str());
    logf("initwts_file:         %s\n", p.initwts_file.c_str());
    logf("train_sent_range:     %s\n", p.train_sent_range.c_str());
    logf("cv_sent_range:        %s\n", p.cv_sent_range.c_str());
    logf("fea_dim:\t\t          %d\n", p.fea_dim);
    logf("fea_context:\t\t      %d\n", p.fea_context);
    logf("bunchsize:\t\t        %d\n", p.bunchsize);
    logf("gpu_used:\t\t          %d\n", p.gpu_used);
    logf("train_cache:\t\t      %d\n", p.traincache);
    logf("init_randem_seed:\t\t  %d\n", p.init_randem_seed);
    logf("targ_offset:\t\t      %d\n", p.targ_offset);
    logf("dropoutflag:\t\t      %d\n", p.dropoutflag);
    logf("MLflag:\t\t      %d\n", p.MLflag);
    logf("init_randem_weight_max:\t\t  %f\n", p.init_randem_weight_max);
    logf("init_randem_weight_min:\t\t  %f\n", p.init_randem_weight_min);
    logf("init_randem_bias_max:\t\t    %f\n", p.init_randem_bias_max);
    logf("init_randem_bias_min:\t\t    %f\n", p.init_randem_bias_min);
    logf("momentum:\t\t                %f\n", p.momentum);
    logf("shapefactor:\t\t                %f\n", p.shapefactor);
    logf("weightcost:\t\t              %f\n", p.weightcost);
    logf("learnrate:\t\t              %f\n", p.lrate);
    logf("visible_omit:\t\t      %f\n", p.visible_omit);
    logf("hid_omit:\t\t      %f\n", p.hid_omit);
    logf("layersizes:\t\t              ");
    for (int j = 0; j < p.numlayers; j++) logf("%d,", p.layersizes[j]);
    logf("\n");
    logf("Please check...\n");
    if (p.numlayers < 2) { logf("layersizes must name at least two layers\n"); return false; }

    // norm file, Interface.cc:374-399: skip a line, fea_dim means, skip a line, fea_dim reciprocal stds
    FILE *fn = fopen(p.norm_file.c_str(), "rt");
    if (!fn) { logf("can not open normalization file: %s\n", p.norm_file.c_str()); return false; }
    logf("Loading Norm file...\n");
    char line[1024];
    mean.assign(p.fea_dim, 0.f); dvar.assign(p.fea_dim, 0.f);
    if (!fgets(line, sizeof line, fn)) line[0] = 0;
    for (int j = 0; j < p.fea_dim; j++) { if (!fgets(line, sizeof line, fn)) line[0] = 0; mean[j] = (float)atof(line); }
    if (!fgets(line, sizeof line, fn)) line[0] = 0;
    for (int j = 0; j < p.fea_dim; j++) { if (!fgets(line, sizeof line, fn)) line[0] = 0; dvar[j] = (float)atof(line); }
    fclose(fn);
    logf("Norm file loaded.\n");

    for (int l = 1; l < p.numlayers; l++) {
        W[l].assign((size_t)p.layersizes[l] * p.layersizes[l - 1], 0.f);
        b[l].assign(p.layersizes[l], 0.f);
    }
    // :411 srand48(seed) -- one stream for chunk order and sample order.  The POSIX 48-bit LCG is restated here
    // (X' = 0x5DEECE66D * X + 0xB mod 2^48, seed -> (seed << 16) | 0x330E) so that the loader owns its state.
    rng_state = (((uint64_t)(uint32_t)p.init_randem_seed) << 16) | 0x330E;
    if (p.initwts_file.empty()) { logf("fatal_error, please set initial weights file\n"); return false; }   // :425-427
    FILE *fw = fopen(p.initwts_file.c_str(), "rb");
    if (!fw) { logf("can not open initial weights file: %s\n", p.initwts_file.c_str()); return false; }
    logf("Loading Init weight file...\n");
    for (int l = 1; l < p.numlayers; l++) {   // MAT-v4 records, :442-464
        int32_t st[5];
        char name[256];
        if (fread(st, 4, 5, fw) != 5 || st[4] < 0 || st[4] > 255 || fread(name, 1, st[4], fw) != (size_t)st[4]) { logf("init weights file truncated\n"); fclose(fw); return false; }
        if (st[1] != p.layersizes[l] || st[2] != p.layersizes[l - 1]) {
            logf("%d,%d,%d,%d\n", st[1], st[2], p.layersizes[l], p.layersizes[l - 1]);
            logf("init weights node nums do not match\n");
            fclose(fw);
            return false;
        }
        if (fread(W[l].data(), 4, W[l].size(), fw) != W[l].size()) { logf("init weights file truncated\n"); fclose(fw); return false; }
        if (fread(st, 4, 5, fw) != 5 || st[4] < 0 || st[4] > 255 || fread(name, 1, st[4], fw) != (size_t)st[4]) { logf("init weights file truncated\n"); fclose(fw); return false; }
        if (st[2] != p.layersizes[l] || st[1] != 1) { logf("init bias node nums do not match\n"); fclose(fw); return false; }
        if (fread(b[l].data(), 4, b[l].size(), fw) != b[l].size()) { logf("init weights file truncated\n"); fclose(fw); return false; }
    }
    fclose(fw);
    logf("Init weight file loaded.\n");
    if (p.fea_dim * p.fea_context != p.layersizes[0]) { logf("feadim times context must be equal to layersizes[0]\n"); return false; }   // :471-475
    if (p.traincache < 1 || p.bunchsize < 1) { logf("traincache and bunchsize must be positive\n"); return false; }
    return true;
}

bool Host::header_uint(const std::vector<char> &hdr, const char *name, unsigned *val)   // get_uint, :988-1009
{
    const char *q = strstr(hdr.data(), name);
    if (!q) { logf("pfile header format is Not correct.\n"); return false; }
    q += strlen(name);
    int count = 0;
    if (sscanf(q, " %u%n", val, &count) < 1 || count <= 1) { logf("%s num in pfile header is Not correct.\n", name); return false; }
    return true;
}

bool Host::read_tail(FILE *fp, long off, unsigned n, std::vector<int> &out)   // read_tail, :1011-1024 (first offset, 0, skipped)
{
    out.assign(n, 0);
    fseek(fp, off + 4, SEEK_SET);
    if (fread(out.data(), 4, n, fp) != n) { logf("pfile tail is Not correct.\n"); return false; }
    for (unsigned i = 0; i < n; i++) out[i] = (int)bswap32((uint32_t)out[i]);
    return true;
}

bool Host::pfile_info()
{
    std::vector<char> hdr(kPfileHeader + 1, 0);
    logf("begin to read in_pfile\n");
    fseek(fp_data, 0, SEEK_SET);
    if (fread(hdr.data(), kPfileHeader, 1, fp_data) != 1) { logf("Failed to read data pfile header.\n"); return false; }
    if (!header_uint(hdr, "-num_sentences", &total_sents) || !header_uint(hdr, "-num_frames", &total_frames)) return false;
    if (!read_tail(fp_data, (long)total_frames * 4L * (2 + p.fea_dim) + kPfileHeader, total_sents, frames_before_sent)) return false;
    logf("begin to read target_pfile\n");
    fseek(fp_targ, 0, SEEK_SET);
    std::fill(hdr.begin(), hdr.end(), 0);
    if (fread(hdr.data(), kPfileHeader, 1, fp_targ) != 1) { logf("Failed to read target pfile header.\n"); return false; }
    unsigned ts = 0, tf = 0;
    if (!header_uint(hdr, "-num_sentences", &ts) || !header_uint(hdr, "-num_frames", &tf)) return false;
    std::vector<int> tt;
    const int D = p.layersizes[p.numlayers - 1];
    if (!read_tail(fp_targ, (long)tf * 4L * (2 + D) + kPfileHeader, ts, tt)) return false;
    logf("tmpsentnum=%d,tmpframenum=%d,total_frames=%d\n", ts, tf, total_frames);
    if (ts != total_sents || tf != total_frames) { logf("frames or sentence num in target pfile and data pfile is not consistent.\n"); return false; }
    logf("frames or sentence num in target pfile and data pfile is consistent.\n");
    for (unsigned i = 0; i < total_sents; i++)
        if (tt[i] != frames_before_sent[i]) { logf("tails in target pfile and data pfile is not consistent---%d.\n", i); return false; }
    logf("Get pfile info over: Training data has %u frames, %u sentences.\n", total_frames, total_sents);
    return true;
}

bool Host::chunk_info(const std::string &range, bool cv)
{
    const size_t dash = range.find('-');
    if (dash == std::string::npos) { logf("%ssent range: %s format error.\n", cv ? "cv " : "", range.c_str()); return false; }
    const int st = atoi(range.substr(0, dash).c_str()), en = atoi(range.substr(dash + 1).c_str());
    if (en < st || st < 0 || en >= (int)total_sents) { logf("%ssent range: %d to %d number error.\n", cv ? "cv " : "", st, en); return false; }
    std::vector<int> &starts = cv ? cv_chunk_st : chunk_st;
    starts.clear();
    int cur_frame = st == 0 ? 0 : frames_before_sent[st - 1];
    starts.push_back(cur_frame);
    int in_chunk = 0;
    for (int s = st; s <= en; s++) {
        const int inc = frames_before_sent[s] - cur_frame;
        cur_frame = frames_before_sent[s];
        const int lost = inc >= p.fea_context ? p.fea_context - 1 : inc;   // a sentence shorter than the context yields nothing
        in_chunk += inc - lost;
        while (in_chunk >= p.traincache) {                                 // chunk boundary inside this sentence
            const int next_st = cur_frame - (in_chunk - p.traincache);
            starts.push_back(next_st);
            in_chunk = (cur_frame - next_st > p.fea_context - 1) ? (cur_frame - next_st - p.fea_context + 1) : 0;
        }
    }
    const int chunks = (int)starts.size(), samples = (chunks - 1) * p.traincache + in_chunk;
    if (cv) { cv_sent_st = st; cv_sent_en = en; cv_total_chunks = chunks; cv_total_samples = samples;
              logf("Get cv chunk info over: CV sentences have %d chunks, %d samples.\n", chunks, samples); }
    else { sent_st = st; sent_en = en; total_chunks = chunks; total_samples = samples;
           logf("Get chunk info over: Training sentences have %d chunks, %d samples.\n", chunks, samples); }
    return true;
}

void Host::shuffle(std::vector<int> &v)
{
    const int n = (int)v.size();
    for (int i = 0; i < n - 1; i++) {
        rng_state = (0x5DEECE66DULL * rng_state + 0xBULL) & ((1ULL << 48) - 1);
        const int idx = (int)((long)(rng_state >> 17) % (n - i));   // lrand48() % (len - i), :982
        std::swap(v[idx], v[n - 1 - i]);
    }
}

// Readchunk / Readchunk_cv, Interface.cc:719-972
int Host::read_chunk(int idx, bool cv, std::vector<float> &in, std::vector<float> &targ)
{
    const std::vector<int> &starts = cv ? cv_chunk_st : chunk_st;
    const int nchunks = cv ? cv_total_chunks : total_chunks, tot = cv ? cv_total_samples : total_samples, last_sent = cv ? cv_sent_en : sent_en;
    const int D = p.layersizes[p.numlayers - 1], in_dim = p.layersizes[0], fd = p.fea_dim, ctx = p.fea_context;
    int need, samples;
    if (idx == nchunks - 1) { need = frames_before_sent[last_sent] - starts[idx]; samples = tot - p.traincache * idx; }
    else { samples = p.traincache; need = starts[idx + 1] - starts[idx]; }
    std::vector<int> order(samples);
    for (int i = 0; i < samples; i++) order[i] = i;
    if (!cv) shuffle(order);                                               // per-sample shuffle, :750-754
    in.resize((size_t)samples * in_dim);
    targ.resize((size_t)samples * D);

    for (int stream = 0; stream < 2; stream++) {
        const int dim = stream == 0 ? fd : D;
        FILE *fp = stream == 0 ? fp_data : fp_targ;
        const long rec = 4L * (dim + 2);
        if (fseek(fp, kPfileHeader + (long)starts[idx] * rec, SEEK_SET) != 0) { logf("%s pfile cannot fseek to chunk %d.\n", stream ? "targ" : "data", idx); return -1; }
        std::vector<uint32_t> raw((size_t)need * (dim + 2));
        if (fread(raw.data(), rec, need, fp) != (size_t)need) { logf("%s pfile short read in chunk %d.\n", stream ? "targ" : "data", idx); return -1; }
        int cur_sent = (int)bswap32(raw[0]);                               // sentence id of the first record
        std::vector<float> x((size_t)need * dim);
        for (int f = 0; f < need; f++)
            for (int j = 0; j < dim; j++) {
                uint32_t u = bswap32(raw[(size_t)f * (dim + 2) + 2 + j]);
                float v;
                memcpy(&v, &u, 4);
                v -= mean[j % fd];                                         // targets use the NOISY mean/dVar too, :807-808
                v *= dvar[j % fd];
                x[(size_t)f * dim + j] = v;
            }
        int processed = 0, cur_frame = starts[idx], cur_sample = 0;
        while (processed != need) {
            int n;
            if (frames_before_sent[cur_sent] > need + starts[idx]) n = need - processed;
            else n = frames_before_sent[cur_sent] - cur_frame;
            for (int j = 0; j <= n - ctx; j++) {
                if (cur_sample >= samples) break;                          // (the reference would overrun its index array here)
                const size_t row = (size_t)order[cur_sample];
                if (stream == 0)
                    memcpy(&in[row * in_dim], &x[(size_t)(processed + j) * fd], sizeof(float) * fd * ctx);   // frames j..j+ctx-1, :778-785
                else
                    memcpy(&targ[row * D], &x[(size_t)(processed + j + p.targ_offset) * D], sizeof(float) * D);   // :822-827
                cur_sample++;
            }
            cur_frame = frames_before_sent[cur_sent];
            cur_sent++;
            processed += n;
        }
    }
    return samples;
}

// Readchunk (Interface.cc:719-838) without its arithmetic: raw records + row -> first-frame map (see interface.h).
// The records of frames [rec0, rec0 + nrec) of the chunk are read into caller-owned buffers (page-locked in the trainer);
// the row map depends only on the sentence table, so every data-parallel rank builds the same map from its own slice.
int Host::read_chunk_raw_slice(int idx, int rank, int world, unsigned **fea_buf, size_t *fea_cap, unsigned **targ_buf, size_t *targ_cap,
                               void *(*alloc)(size_t), void (*release)(void *), std::vector<int> &first, int *need_out, int *rec0_out, int *nrec_out)
{
    const int D = p.layersizes[p.numlayers - 1], fd = p.fea_dim, ctx = p.fea_context;
    int need, samples;
    if (idx == total_chunks - 1) { need = frames_before_sent[sent_en] - chunk_st[idx]; samples = total_samples - p.traincache * idx; }
    else { samples = p.traincache; need = chunk_st[idx + 1] - chunk_st[idx]; }
    std::vector<int> order(samples);
    for (int i = 0; i < samples; i++) order[i] = i;
    shuffle(order);                                                        // per-sample shuffle, :750-754
    first.assign(samples, 0);
    const int S = (need + world - 1) / world;
    const int rec0 = world > 1 ? rank * S : 0;
    const int nrec = world > 1 ? std::max(0, std::min(S, need - rec0)) : need;
    // both streams at once, each with read_threads concurrent pread()s (the records of a chunk are contiguous in the pfile)
    bool ok_stream[2] = {true, true};
    std::thread readers[2];
    for (int stream = 0; stream < 2; stream++) {
        const int dim = stream == 0 ? fd : D;
        FILE *fp = stream == 0 ? fp_data : fp_targ;
        unsigned **buf = stream == 0 ? fea_buf : targ_buf;
        size_t *cap = stream == 0 ? fea_cap : targ_cap;
        const long rec = 4L * (dim + 2);
        const size_t bytes = (size_t)std::max(nrec, 1) * rec;
        if (bytes > *cap) {
            if (*buf) release(*buf);
            *cap = bytes + bytes / 8;
            *buf = (unsigned *)alloc(*cap);
            if (!*buf) { *cap = 0; logf("cannot allocate %zu bytes for the records of chunk %d.\n", bytes, idx); return -1; }
        }
        unsigned *dst = *buf;
        readers[stream] = std::thread([=, &ok_stream] {
            ok_stream[stream] = nrec == 0 || pread_parallel(fileno(fp), dst, (size_t)nrec * rec, (off_t)kPfileHeader + (off_t)(chunk_st[idx] + rec0) * rec, p.read_threads);
        });
    }
    for (auto &r : readers) r.join();
    for (int stream = 0; stream < 2; stream++)
        if (!ok_stream[stream]) { logf("%s pfile short read in chunk %d.\n", stream ? "targ" : "data", idx); return -1; }
    // sentence that contains the first frame of the chunk (the reference reads it from the first record, Interface.cc:744)
    int cur_sent = (int)(std::upper_bound(frames_before_sent.begin(), frames_before_sent.end(), chunk_st[idx]) - frames_before_sent.begin());
    if (rec0 == 0 && nrec > 0 && (int)bswap32((*fea_buf)[0]) != cur_sent) { logf("chunk %d: sentence id %d of the first record disagrees with the sentence table (%d).\n", idx, (int)bswap32((*fea_buf)[0]), cur_sent); return -1; }
    int processed = 0, cur_frame = chunk_st[idx], cur_sample = 0;
    while (processed != need) {
        int n;
        if (frames_before_sent[cur_sent] > need + chunk_st[idx]) n = need - processed;
        else n = frames_before_sent[cur_sent] - cur_frame;
        for (int j = 0; j <= n - ctx; j++) {
            if (cur_sample >= samples) break;
            first[order[cur_sample]] = processed + j;                      // rows = frames j..j+ctx-1, target = frame j+targ_offset
            cur_sample++;
        }
        cur_frame = frames_before_sent[cur_sent];
        cur_sent++;
        processed += n;
    }
    *need_out = need;
    if (rec0_out) *rec0_out = rec0;
    if (nrec_out) *nrec_out = nrec;
    return samples;
}

int Host::read_chunk_raw(int idx, std::vector<unsigned> &fea_rec, std::vector<unsigned> &targ_rec, std::vector<int> &first, int *need_out)
{
    unsigned *fb = nullptr, *tb = nullptr;
    size_t fc = 0, tc = 0;
    const int n = read_chunk_raw_slice(idx, 0, 1, &fb, &fc, &tb, &tc, malloc, free, first, need_out, nullptr, nullptr);
    if (n >= 0) {
        fea_rec.assign(fb, fb + (size_t)*need_out * (2 + p.fea_dim));
        targ_rec.assign(tb, tb + (size_t)*need_out * (2 + p.layersizes[p.numlayers - 1]));
    }
    free(fb); free(tb);
    return n;
}

bool Host::write_weights()
{
    logf("Saving weights to file...\n");
    for (int l = 1; l < p.numlayers; l++) {
        char name[64];
        snprintf(name, sizeof name, "weights%d%d", l, l + 1);
        int32_t st[5] = {10, p.layersizes[l], p.layersizes[l - 1], 0, (int32_t)strlen(name) + 1};
        fwrite(st, 4, 5, fp_out); fwrite(name, 1, st[4], fp_out); fwrite(W[l].data(), 4, W[l].size(), fp_out);
        snprintf(name, sizeof name, "bias%d", l + 1);
        int32_t sb[5] = {10, 1, p.layersizes[l], 0, (int32_t)strlen(name) + 1};
        fwrite(sb, 4, 5, fp_out); fwrite(name, 1, sb[4], fp_out); fwrite(b[l].data(), 4, b[l].size(), fp_out);
    }
    fflush(fp_out);
    logf("Saving over.\n");
    return true;
}

}  // namespace bphost

// ---- plain-C view of the loader (used by the CPU tests and by bindings written in other languages) ----------
extern "C" {
void *bph_create(int argc, char **argv)
{
    bphost::Host *h = new bphost::Host();
    if (!h->init(argc, argv) || !h->pfile_info()) { delete h; return nullptr; }
    return h;
}
void bph_destroy(void *p) { delete (bphost::Host *)p; }
int bph_numlayers(void *p) { return ((bphost::Host *)p)->p.numlayers; }
int bph_chunk_info(void *p, const char *range, int cv, int *chunks, int *samples)
{
    bphost::Host *h = (bphost::Host *)p;
    if (!h->chunk_info(range, cv != 0)) return -1;
    *chunks = cv ? h->cv_total_chunks : h->total_chunks;
    *samples = cv ? h->cv_total_samples : h->total_samples;
    return 0;
}
void bph_shuffle(void *p, int *idx, int n)
{
    std::vector<int> v(idx, idx + n);
    ((bphost::Host *)p)->shuffle(v);
    for (int i = 0; i < n; i++) idx[i] = v[i];
}
// two-call protocol: in == NULL returns the sample count of the NEXT read without consuming anything is not possible
// (the shuffle consumes random numbers), so the caller passes buffers sized traincache * dim.
int bph_read_chunk(void *p, int idx, int cv, float *in, float *targ)
{
    bphost::Host *h = (bphost::Host *)p;
    std::vector<float> a, b;
    const int n = h->read_chunk(idx, cv != 0, a, b);
    if (n < 0) return n;
    memcpy(in, a.data(), a.size() * sizeof(float));
    memcpy(targ, b.data(), b.size() * sizeof(float));
    return n;
}
// raw chunk for the device-side loader: fea / targ must hold need*(2+dim) words (need <= max_frames), first >= samples ints
int bph_read_chunk_raw(void *p, int idx, unsigned *fea, unsigned *targ, int *first, int max_frames, int *need)
{
    bphost::Host *h = (bphost::Host *)p;
    std::vector<unsigned> a, b;
    std::vector<int> f;
    const int n = h->read_chunk_raw(idx, a, b, f, need);
    if (n < 0 || *need > max_frames) return -1;
    memcpy(fea, a.data(), a.size() * sizeof(unsigned));
    memcpy(targ, b.data(), b.size() * sizeof(unsigned));
    memcpy(first, f.data(), f.size() * sizeof(int));
    return n;
}
// data-parallel view: the records of rank `rank`'s slice of chunk idx and the (rank-independent) row map
int bph_read_chunk_raw_slice(void *p, int idx, int rank, int world, unsigned *fea, unsigned *targ, int *first, int max_frames, int *need, int *rec0, int *nrec)
{
    bphost::Host *h = (bphost::Host *)p;
    unsigned *fb = nullptr, *tb = nullptr;
    size_t fc = 0, tc = 0;
    std::vector<int> f;
    const int n = h->read_chunk_raw_slice(idx, rank, world, &fb, &fc, &tb, &tc, malloc, free, f, need, rec0, nrec);
    if (n >= 0 && *nrec <= max_frames) {
        memcpy(fea, fb, (size_t)*nrec * (2 + h->p.fea_dim) * sizeof(unsigned));
        memcpy(targ, tb, (size_t)*nrec * (2 + h->p.layersizes[h->p.numlayers - 1]) * sizeof(unsigned));
        memcpy(first, f.data(), f.size() * sizeof(int));
    }
    free(fb); free(tb);
    return (n >= 0 && *nrec <= max_frames) ? n : -1;
}
const float *bph_weights(void *p, int l) { return ((bphost::Host *)p)->W[l].data(); }
const float *bph_bias(void *p, int l) { return ((bphost::Host *)p)->b[l].data(); }
int bph_write_weights(void *p) { return ((bphost::Host *)p)->write_weights() ? 0 : -1; }
}
