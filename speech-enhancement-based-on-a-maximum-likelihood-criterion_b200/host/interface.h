// interface.h -- host side of the BPtrain_Sigmoid drop-in: key=value command line, pfile / norm / MAT-v4
// weight files, chunking, z-score, context expansion and the lrand48 shuffles.  A from-scratch restatement
// of the behaviour of the reference's Interface class (Train_code_ML_GGD/Interface.{h,cc}); file formats,
// flag names, random sequence and log lines are the reference's, the code is not.
#pragma once
#include <cstdio>
#include <string>
#include <vector>

namespace bphost {

constexpr int kMaxLayer = 10;            // Interface.h:6
constexpr long kPfileHeader = 32768;     // Interface.cc:13

struct Params {                          // struct WorkPara, Interface.h:31-69
    std::string fea_file, norm_file, targ_file, outwts_file, log_file, initwts_file, train_sent_range, cv_sent_range;
    int fea_dim = 0, fea_context = 0, targ_offset = 0, dropoutflag = 0, MLflag = 0, traincache = 0, bunchsize = 0;
    int gpu_used = 0, init_randem_seed = 0;
    std::vector<int> gpus;      // gpu_used=0,1,2,3 (extension): frame-sharded data parallelism, one process per listed GPU
    float momentum = 0, shapefactor = 0, weightcost = 0, lrate = 0, visible_omit = 0, hid_omit = 0;
    float init_randem_weight_min = -0.1f, init_randem_weight_max = 0.1f, init_randem_bias_min = -0.1f, init_randem_bias_max = 0.1f;
    int numlayers = 0;
    int layersizes[kMaxLayer] = {0};
    // extensions of this implementation (ignored by the reference, which skips unknown names)
    int precision = 0;      // precision=fp32 selects the CUDA-core validation path
    int no_graph = 0;
    int host_loader = 0;    // host_loader=1: z-score / context expansion on the CPU (ggd_train) instead of the device-side loader
    int read_threads = 4;   // read_threads=N: threads per pfile for the bulk record reads of the device-side loader
};

class Host {
public:
    ~Host();
    // parse argv, open log/pfiles/output, load norm + initial weights. Returns false after logging the reason.
    bool init(int argc, char **argv);
    bool pfile_info();                                        // Interface::get_pfile_info, Interface.cc:519-586
    bool chunk_info(const std::string &range, bool cv);      // get_chunk_info{,_cv}, Interface.cc:588-716
    void shuffle(std::vector<int> &v);                        // GetRandIndex, Interface.cc:975-986
    // fills in/targ (resized to samples*dim) with chunk `idx`; returns the number of samples, <0 on error
    int read_chunk(int idx, bool cv, std::vector<float> &in, std::vector<float> &targ);
    // the same chunk for the device-side loader (ggd_train_raw): the raw big-endian records as they lie in the pfiles and,
    // per shuffled net-input row, its first context frame inside the chunk; consumes the same random numbers as read_chunk
    int read_chunk_raw(int idx, std::vector<unsigned> &fea_rec, std::vector<unsigned> &targ_rec, std::vector<int> &first, int *need_out);
    // the same with caller-owned (e.g. page-locked) record buffers, grown through alloc / release; with world > 1 only the
    // records of this rank's slice [rank*S, rank*S + S) of the chunk, S = ceil(need / world), are read (ggd_raw_chunk::rec_frame0)
    int read_chunk_raw_slice(int idx, int rank, int world, unsigned **fea_buf, size_t *fea_cap, unsigned **targ_buf, size_t *targ_cap,
                             void *(*alloc)(size_t), void (*release)(void *), std::vector<int> &first, int *need_out, int *rec0_out, int *nrec_out);
    const float *mean_ptr() const { return mean.data(); }
    const float *dvar_ptr() const { return dvar.data(); }
    bool write_weights();                                     // Interface::Writeweights, Interface.cc:484-516
    void logf(const char *fmt, ...);

    Params p;
    std::vector<float> W[kMaxLayer], b[kMaxLayer];           // 1..numlayers-1, reference order out + in*cur
    unsigned total_frames = 0, total_sents = 0;
    int total_chunks = 0, total_samples = 0, cv_total_chunks = 0, cv_total_samples = 0;
    FILE *fp_log = nullptr;

private:
    bool read_tail(FILE *fp, long off, unsigned n, std::vector<int> &out);
    bool header_uint(const std::vector<char> &hdr, const char *name, unsigned *val);
    FILE *fp_data = nullptr, *fp_targ = nullptr, *fp_out = nullptr;
    std::vector<float> mean, dvar;
    std::vector<int> frames_before_sent, chunk_st, cv_chunk_st;
    int sent_st = 0, sent_en = 0, cv_sent_st = 0, cv_sent_en = 0;
    unsigned long long rng_state = 0;     // drand48-family state (srand48 / lrand48, Interface.cc:411, 982)
};

}  // namespace bphost
