// BPtrain_Sigmoid drop-in: the process finetune.pl launches once per epoch (finetune.pl:50-76).
// Same flags, same input / output files, same log lines as the reference's BPtrain.cc:55-146; the device
// work goes through the C ABI of libggd_b200 (include/ggd_train.h).  A loader thread prefetches the next
// chunk while the GPU trains on the current one (the reference's pthread double buffer, BPtrain.cc:15-54).
//
// Extension: gpu_used=0,1,...,N-1 trains frame-sharded on N GPUs.  The process forks one worker per
// extra GPU before CUDA is touched; `bunchsize` stays the GLOBAL minibatch, every rank trains rows
// [rank*bunchsize/N, (rank+1)*bunchsize/N) of every bunch of the same shuffled chunk, so weights, log and CV lines
// are those of the one-GPU run with the same flags (within the arithmetic tolerance).  Rank 0 alone writes files.
#include "../../include/ggd_train.h"
#include "interface.h"
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <ctime>
#include <mutex>
#include <string>
#include <thread>
#include <sys/wait.h>
#include <unistd.h>

using namespace bphost;

static std::vector<int> scan_gpus(int argc, char **argv)
{
    std::vector<int> g;
    for (int i = 1; i < argc; i++)
        if (!strncmp(argv[i], "gpu_used=", 9)) {
            g.clear();
            for (const char *q = argv[i] + 9; *q;) { g.push_back(atoi(q)); q = strchr(q, ','); if (!q) break; q++; }
        }
    if (g.empty()) g.push_back(0);
    return g;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
    const time_t t0 = time(nullptr);
    const double T0 = now_s();
    const bool timing = getenv("GGD_CLI_TIMING") != nullptr;     // per-phase wall times on stderr (tuning aid)
    // ---- data parallelism: one process per listed GPU, forked BEFORE anything touches CUDA
    const std::vector<int> gpus = scan_gpus(argc, argv);
    const int world = (int)gpus.size();
    int rank = 0;
    std::vector<int> to_child(world, -1);    // parent's write ends
    int from_parent = -1;
    std::vector<pid_t> kids;
    for (int r = 1; r < world; r++) {
        int fd[2];
        if (pipe(fd) != 0) { perror("pipe"); return 1; }
        const pid_t pid = fork();
        if (pid < 0) { perror("fork"); return 1; }
        if (pid == 0) {
            rank = r; from_parent = fd[0]; close(fd[1]);
            for (int q = 1; q < r; q++) if (to_child[q] >= 0) close(to_child[q]);
            break;
        }
        close(fd[0]); to_child[r] = fd[1]; kids.push_back(pid);
    }
    std::vector<std::string> extra;
    std::vector<char *> av(argv, argv + argc);
    if (rank > 0) {   // workers write no files and stay quiet
        extra = {"outwts_file=/dev/null", "log_file=/dev/null"};
        for (auto &e : extra) av.push_back(const_cast<char *>(e.c_str()));
        if (!freopen("/dev/null", "w", stdout)) return 1;
    }
    printf("--------activation functin is sigmoid--------\n");
    Host H;
    if (!H.init((int)av.size(), av.data())) return 1;
    Params &p = H.p;
    if (world > 1 && (p.bunchsize % world != 0 || p.host_loader)) {
        H.logf("gpu_used lists %d GPUs: bunchsize %d must be divisible by it and host_loader must be 0\n", world, p.bunchsize);
        return 1;
    }
    unsigned char uid[128] = {0};
    if (world > 1) {
        if (rank == 0) {
            if (ggd_nccl_unique_id(uid) != GGD_OK) { H.logf("%s\n", ggd_last_error()); return 1; }
            for (int r = 1; r < world; r++) if (write(to_child[r], uid, sizeof uid) != (ssize_t)sizeof uid) { perror("write"); return 1; }
        } else if (read(from_parent, uid, sizeof uid) != (ssize_t)sizeof uid) { perror("read"); return 1; }
    }
    ggd_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.numlayers = p.numlayers;
    for (int i = 0; i < p.numlayers; i++) cfg.layersizes[i] = p.layersizes[i];
    cfg.bunchsize = p.bunchsize / world; cfg.lrate = p.lrate; cfg.momentum = p.momentum; cfg.weightcost = p.weightcost;
    cfg.shapefactor = p.shapefactor; cfg.MLflag = p.MLflag; cfg.dropoutflag = p.dropoutflag;
    cfg.visible_omit = p.visible_omit; cfg.hid_omit = p.hid_omit; cfg.gpu = gpus[rank]; cfg.seed = p.init_randem_seed;
    cfg.precision = p.precision; cfg.world_size = world; cfg.rank = rank; cfg.nccl_unique_id = world > 1 ? uid : nullptr; cfg.flags = GGD_FLAG_PIN_HOST | (p.no_graph ? GGD_FLAG_NO_GRAPH : 0);
    const float *Wp[GGD_MAXLAYER] = {nullptr}, *bp[GGD_MAXLAYER] = {nullptr};
    for (int l = 1; l < p.numlayers; l++) { Wp[l] = H.W[l].data(); bp[l] = H.b[l].data(); }
    ggd_handle *net = nullptr;
    if (timing) fprintf(stderr, "[rank %d] %.3f s: host init done (norm, weights, files)\n", rank, now_s() - T0);
    if (ggd_create(&cfg, Wp, bp, &net) != GGD_OK) { H.logf("%s\n", ggd_last_error()); printf("%s\n", ggd_last_error()); return 1; }
    printf("Created net with %d layers, bunchsize %d.\n", p.numlayers, p.bunchsize);
    if (timing) fprintf(stderr, "[rank %d] %.3f s: ggd_create done (CUDA context, NCCL, peer mapping)\n", rank, now_s() - T0);
    if (!H.pfile_info()) return 1;

    // ---- train
    if (!H.chunk_info(p.train_sent_range, false)) return 1;
    std::vector<int> order(H.total_chunks);
    for (int i = 0; i < H.total_chunks; i++) order[i] = i;
    H.shuffle(order);
    // double buffer; the buffers are allocated once at full chunk size so that the library can pin them
    std::vector<float> in[2], tg[2];
    // device-side loader (default): the loader thread only reads the raw records and builds the row -> first-frame map;
    // byte swap, z-score, context expansion and target selection run on the GPU (ggd_train_raw).  host_loader=1 restores
    // the reference's division of labour (everything on the CPU, ggd_train).
    const bool raw = !p.host_loader;
    // page-locked record buffers (ggd_host_alloc): the pread()s land where the DMA engine reads; with several GPUs each rank
    // reads only its slice of a chunk's records and the library all-gathers the slices over NVLink
    unsigned *rfea[2] = {nullptr, nullptr}, *rtg[2] = {nullptr, nullptr};
    size_t rfea_cap[2] = {0, 0}, rtg_cap[2] = {0, 0};
    std::vector<int> first[2];
    int need[2] = {0, 0}, rec0[2] = {0, 0}, nrec[2] = {0, 0};
    if (!raw) for (int k = 0; k < 2; k++) { in[k].reserve((size_t)p.traincache * p.layersizes[0]); tg[k].reserve((size_t)p.traincache * p.layersizes[p.numlayers - 1]); }
    int samples[2] = {0, 0}, local_samples[2] = {0, 0};
    std::mutex mu;
    std::condition_variable cv;
    int filled = 0, consumed = 0;   // chunks produced / released
    bool load_failed = false;
    std::thread loader([&] {
        ggd_bind_thread(net);     // page-locked allocations of this thread belong to this rank's GPU
        for (int i = 0; i < H.total_chunks; i++) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return i - consumed < 2; });
            }
            const double tl0 = now_s();
            const int n = raw ? H.read_chunk_raw_slice(order[i], rank, world, &rfea[i & 1], &rfea_cap[i & 1], &rtg[i & 1], &rtg_cap[i & 1], ggd_host_alloc, ggd_host_free,
                                                       first[i & 1], &need[i & 1], &rec0[i & 1], &nrec[i & 1])
                              : H.read_chunk(order[i], false, in[i & 1], tg[i & 1]);
            if (timing) fprintf(stderr, "[rank %d] %.3f s: chunk %d loaded in %.1f ms\n", rank, now_s() - T0, i, (now_s() - tl0) * 1e3);
            int nl = n;
            if (world > 1 && n > 0) {
                // this rank's rows of every GLOBAL bunch (the trailing partial bunch is dropped by the trainer anyway)
                const int M = p.bunchsize, Ml = M / world, nb = n / M;
                std::vector<int> mine((size_t)nb * Ml);
                for (int b = 0; b < nb; b++)
                    for (int j = 0; j < Ml; j++) mine[(size_t)b * Ml + j] = first[i & 1][(size_t)b * M + rank * Ml + j];
                first[i & 1].swap(mine);
                nl = nb * Ml;
            }
            std::lock_guard<std::mutex> lk(mu);
            samples[i & 1] = n; local_samples[i & 1] = nl;
            if (n < 0) load_failed = true;
            filled = i + 1;
            cv.notify_all();
            if (n < 0) return;
        }
    });
    int rc = 0;
    for (int i = 0; i < H.total_chunks && rc == 0; i++) {
        const double tw0 = now_s();
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return filled > i || load_failed; });
            if (load_failed) { rc = 1; break; }
        }
        H.logf("Starting chunk %d of %d containing %d samples.\n", i + 1, H.total_chunks, samples[i & 1]);
        const double tt0 = now_s();
        if (samples[i & 1] % p.bunchsize) printf("this bunch has only %d samples and is ignored.\n", samples[i & 1] % p.bunchsize);
        if (raw) {
            ggd_raw_chunk c;
            memset(&c, 0, sizeof c);
            c.fea_records = rfea[i & 1]; c.targ_records = rtg[i & 1];
            if (world > 1) { c.rec_frame0 = rec0[i & 1]; c.rec_frames = nrec[i & 1]; }
            c.n_frames = need[i & 1]; c.n_samples = local_samples[i & 1]; c.sample_first_frame = first[i & 1].data();
            c.fea_dim = p.fea_dim; c.fea_context = p.fea_context; c.targ_offset = p.targ_offset;
            c.mean = H.mean_ptr(); c.dvar = H.dvar_ptr();
            if (ggd_train_raw(net, &c) != GGD_OK) { H.logf("%s\n", ggd_last_error()); rc = 1; }
        } else if (ggd_train(net, samples[i & 1], in[i & 1].data(), tg[i & 1].data()) != GGD_OK) { H.logf("%s\n", ggd_last_error()); rc = 1; }
        if (timing) fprintf(stderr, "[rank %d] %.3f s: chunk %d trained in %.1f ms (waited for it %.1f ms)\n", rank, now_s() - T0, i, (now_s() - tt0) * 1e3, (tt0 - tw0) * 1e3);
        std::lock_guard<std::mutex> lk(mu);
        consumed = i + 1;
        cv.notify_all();
    }
    { std::lock_guard<std::mutex> lk(mu); consumed = H.total_chunks + 2; cv.notify_all(); }
    loader.join();
    for (int k = 0; k < 2; k++) { ggd_host_free(rfea[k]); ggd_host_free(rtg[k]); }
    if (rc) return rc;
    H.logf("Total cost time: %.1f s.\n", (double)(time(nullptr) - t0));
    if (rank > 0) {
        // workers hold their (bit-identical) weights until rank 0 has finished writing and cross-validating
        char done;
        if (read(from_parent, &done, 1) < 0) perror("read");
        ggd_destroy(net);
        return 0;
    }

    printf("begin to write weights\n");
    float *Wo[GGD_MAXLAYER] = {nullptr}, *bo[GGD_MAXLAYER] = {nullptr};
    for (int l = 1; l < p.numlayers; l++) { Wo[l] = H.W[l].data(); bo[l] = H.b[l].data(); }
    if (ggd_get_weights(net, Wo, bo) != GGD_OK) { H.logf("%s\n", ggd_last_error()); return 1; }
    H.write_weights();
    printf("finish to write weights\n\n");

    // ---- cross validation (BPtrain.cc:114-139)
    printf("begin to CV\n");
    H.logf("Starting CV.\n");
    if (!H.chunk_info(p.cv_sent_range, true)) return 1;
    float squared_err = 0.f, db_err = 0.f, likelihood = 0.f;
    for (int i = 0; i < H.cv_total_chunks; i++) {
        const int n = H.read_chunk(i, true, in[0], tg[0]);
        if (n < 0) return 1;
        printf("cur_chunk_samples=%d\n", n);
        // one forward pass feeds CrossValid, CrossValiddB and CrossValid2 (BPtrain.cc:124-128 runs it three times)
        float r[3] = {0, 0, 0};
        if (ggd_cv_all(net, n, in[0].data(), tg[0].data(), r) != GGD_OK) { H.logf("%s\n", ggd_last_error()); return 1; }
        squared_err += r[0];
        db_err += r[1];
        if (p.MLflag == 1) likelihood += r[2];
    }
    H.logf("CV over. squared error: %f\n", squared_err / H.cv_total_samples);
    H.logf("CV over. square root squared error: %f\n", db_err / H.cv_total_samples);
    if (p.MLflag == 1) H.logf("CV2 over. CV log likelihood: %f\n", likelihood / H.cv_total_samples);
    printf("all finish!\n");
    for (int r = 1; r < world; r++) { const char done = 1; if (write(to_child[r], &done, 1) < 0) perror("write"); close(to_child[r]); }
    ggd_destroy(net);
    for (pid_t k : kids) { int st = 0; waitpid(k, &st, 0); }
    return 0;
}
