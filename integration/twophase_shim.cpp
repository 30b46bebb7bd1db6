// C entry for tests and the bench: nqr::NyquistIO::Load (the reference's own Common.cpp, unmodified)
// on top of the two-phase OpusDecoder.  ctypes cannot call C++ directly.  The samples are handed out
// in place (the AudioData lives until nq_twophase_free): copying 86 MB out of the vector would be a
// cost of this wrapper, not of Load.
#include "Decoders.h"
#include "OpusBatchLoader.h"

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace {
std::mutex g_mu;
std::map<const float *, std::unique_ptr<nqr::AudioData>> g_live;
}

extern "C" {

extern void nq_twophase_last_timing(double out[8]);

// Returns 0 and the interleaved float samples (valid until nq_twophase_free); -1 and a message on
// stderr if Load throws.  timing: 8 doubles, see OpusDecoderTwoPhase.cpp.
__attribute__((visibility("default"))) int nq_twophase_load(const char *path, float **samples, size_t *count,
                                                             int *channels, int *sample_rate, double timing[8])
{
    try {
        nqr::NyquistIO loader;
        std::unique_ptr<nqr::AudioData> data(new nqr::AudioData());
        loader.Load(data.get(), std::string(path));
        *count = data->samples.size();
        *channels = data->channelCount;
        *sample_rate = data->sampleRate;
        *samples = data->samples.data();
        if (timing) nq_twophase_last_timing(timing);
        std::lock_guard<std::mutex> lk(g_mu);
        g_live[data->samples.data()] = std::move(data);
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "nq_twophase_load: " << e.what() << std::endl;
        return -1;
    }
}

// nqr::LoadOpusBatch: n paths -> samples[i] / counts[i] / channels[i] (each freed with nq_twophase_free).
// stats: phase 1 s, phase 2 s, total s, files batched, files single, CELT frames, kernel launches.
__attribute__((visibility("default"))) int nq_twophase_load_batch(const char *const *paths, int n, int threads, float **samples,
                                                                   size_t *counts, int *channels, double stats[7])
{
    try {
        std::vector<std::string> p(paths, paths + n);
        std::vector<std::shared_ptr<nqr::AudioData>> out;
        const nqr::OpusBatchStats st = nqr::LoadOpusBatch(p, out, threads);
        std::lock_guard<std::mutex> lk(g_mu);
        for (int i = 0; i < n; i++) {
            std::unique_ptr<nqr::AudioData> d(new nqr::AudioData(std::move(*out[i])));
            counts[i] = d->samples.size();
            channels[i] = d->channelCount;
            samples[i] = d->samples.data();
            g_live[d->samples.data()] = std::move(d);
        }
        if (stats) {
            stats[0] = st.phase1Seconds; stats[1] = st.phase2Seconds; stats[2] = st.totalSeconds;
            stats[3] = st.filesBatched; stats[4] = st.filesSingle; stats[5] = (double)st.frames; stats[6] = st.launches;
        }
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "nq_twophase_load_batch: " << e.what() << std::endl;
        return -1;
    }
}

__attribute__((visibility("default"))) void nq_twophase_free(float *p)
{
    std::lock_guard<std::mutex> lk(g_mu);
    g_live.erase(p);
}

}  // extern "C"
