// pfile_writer.h -- the QuickNet pfile container and the .norm text file, as the reference's tools produce them
// (tools_pfile/pfile_noisy.pl:33,45: feacat + pfile_concat; get_norm.pl:4: qnnorm) and as Interface.cc reads them
// (header keys :531-537, 988-1009; records :735-766; sentence table :1011-1024; norm :385-396).
// The records themselves (259 big-endian words per frame) come from the LPS kernel (LPS_FLAG_PFILE, include/lps_b200.h).
#pragma once
#include <cstdint>
#include <cstdio>
#include <vector>

namespace bphost {

class PfileWriter {
public:
    ~PfileWriter();
    // sent_frames: frames of every sentence, known up front (lps_nframes of each utterance); dim = features per frame
    bool open(const char *path, const std::vector<long> &sent_frames, int dim);
    // appends `frames` records of (2 + dim) big-endian words; the sentence / frame words are taken as they are
    bool append(const uint32_t *records, long frames);
    bool close();                      // writes the sentence table; false when the appended frames do not add up
    long frames_written() const { return written_; }
private:
    FILE *fp_ = nullptr;
    std::vector<long> sent_frames_;
    long total_ = 0, written_ = 0;
    int dim_ = 0;
};

// "vec N" + N means + "vec N" + N reciprocal standard deviations, %g like qnnorm's output (train_noisy.norm)
bool write_norm_file(const char *path, const float *mean, const float *dvar, int dim);

}  // namespace bphost
