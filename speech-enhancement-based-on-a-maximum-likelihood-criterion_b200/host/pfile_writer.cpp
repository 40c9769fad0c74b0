// pfile_writer.cpp -- see pfile_writer.h.
#include "pfile_writer.h"
#include <cstring>
#include <string>

namespace bphost {

static const long kHeaderBytes = 32768;      // Interface.cc:13

PfileWriter::~PfileWriter() { if (fp_) fclose(fp_); }

bool PfileWriter::open(const char *path, const std::vector<long> &sent_frames, int dim)
{
    sent_frames_ = sent_frames; dim_ = dim; total_ = 0; written_ = 0;
    for (long n : sent_frames) { if (n < 0) return false; total_ += n; }
    if (!(fp_ = fopen(path, "wb"))) return false;
    const long words = total_ * (2 + dim);
    std::string h = "-pfile_header version 0 size 32768\n";
    h += "-num_sentences " + std::to_string(sent_frames.size()) + "\n";
    h += "-num_frames " + std::to_string(total_) + "\n";
    h += "-first_feature_column 2\n-num_features " + std::to_string(dim) + "\n";
    h += "-first_label_column " + std::to_string(2 + dim) + "\n-num_labels 0\n";
    h += "-format dd" + std::string((size_t)dim, 'f') + "\n";
    h += "-data size " + std::to_string(words) + " offset 0 ndim 2 nrow " + std::to_string(total_) + " ncol " + std::to_string(2 + dim) + "\n";
    h += "-sent_table_data size " + std::to_string(sent_frames.size() + 1) + " offset " + std::to_string(words) + " ndim 1\n-end\n";
    if ((long)h.size() > kHeaderBytes) return false;
    std::vector<char> block(kHeaderBytes, 0);
    memcpy(block.data(), h.data(), h.size());
    return fwrite(block.data(), 1, block.size(), fp_) == block.size();
}

bool PfileWriter::append(const uint32_t *records, long frames)
{
    if (!fp_ || frames < 0 || written_ + frames > total_) return false;
    const size_t n = (size_t)frames * (2 + dim_);
    if (fwrite(records, sizeof(uint32_t), n, fp_) != n) return false;
    written_ += frames;
    return true;
}

bool PfileWriter::close()
{
    if (!fp_) return false;
    bool ok = written_ == total_;
    // sentence table: num_sentences + 1 big-endian cumulative frame offsets (the reader skips the first, Interface.cc:1011-1024)
    uint32_t acc = 0;
    std::vector<uint32_t> tab;
    tab.push_back(__builtin_bswap32(0u));
    for (long n : sent_frames_) { acc += (uint32_t)n; tab.push_back(__builtin_bswap32(acc)); }
    ok = ok && fwrite(tab.data(), sizeof(uint32_t), tab.size(), fp_) == tab.size();
    ok = (fclose(fp_) == 0) && ok;
    fp_ = nullptr;
    return ok;
}

bool write_norm_file(const char *path, const float *mean, const float *dvar, int dim)
{
    FILE *f = fopen(path, "wt");
    if (!f) return false;
    fprintf(f, "vec %d\n", dim);
    for (int i = 0; i < dim; i++) fprintf(f, "%g\n", mean[i]);
    fprintf(f, "vec %d\n", dim);
    for (int i = 0; i < dim; i++) fprintf(f, "%g\n", dvar[i]);
    return fclose(f) == 0;
}

}  // namespace bphost

// ---- plain-C view (CPU tests, bindings) ----
extern "C" {
int bph_write_pfile(const char *path, const unsigned *records, long frames, int dim, const long *sent_frames, int n_sents)
{
    bphost::PfileWriter w;
    std::vector<long> sf(sent_frames, sent_frames + n_sents);
    if (!w.open(path, sf, dim)) return -1;
    // two appends on purpose when possible: the streaming path is the one the tool uses
    const long half = frames / 2;
    if (!w.append(records, half) || !w.append(records + (size_t)half * (2 + dim), frames - half)) return -1;
    return w.close() ? 0 : -1;
}
int bph_write_norm(const char *path, const float *mean, const float *dvar, int dim) { return bphost::write_norm_file(path, mean, dvar, dim) ? 0 : -1; }
}
