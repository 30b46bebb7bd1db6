// TEST INFRASTRUCTURE.  C entry to the UNMODIFIED reference's nqr::NyquistIO::Load (src/Common.cpp,
// src/OpusDecoder.cpp, src/OpusDependencies.c compiled in place by oracle/Makefile into
// oracle/_ref/libnyquist_ref.so), so tests and bench.py can time and compare the reference's own
// end-to-end file decode next to the two-phase GPU build (integration/).  Same signature as
// integration/twophase_shim.cpp.
#include "Decoders.h"

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>

namespace {
std::mutex g_mu;
std::map<const float *, std::unique_ptr<nqr::AudioData>> g_live;   // samples handed out in place: no copy inside the timed call
}

extern "C" {

__attribute__((visibility("default"))) int nqref_load(const char *path, float **samples, size_t *count, int *channels,
                                                       int *sample_rate)
{
    try {
        nqr::NyquistIO loader;
        std::unique_ptr<nqr::AudioData> data(new nqr::AudioData());
        loader.Load(data.get(), std::string(path));
        *count = data->samples.size();
        *channels = data->channelCount;
        *sample_rate = data->sampleRate;
        *samples = data->samples.data();
        std::lock_guard<std::mutex> lk(g_mu);
        g_live[data->samples.data()] = std::move(data);
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "nqref_load: " << e.what() << std::endl;
        return -1;
    }
}

__attribute__((visibility("default"))) void nqref_free(float *p)
{
    std::lock_guard<std::mutex> lk(g_mu);
    g_live.erase(p);
}

}  // extern "C"
