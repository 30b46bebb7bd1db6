// C entry for tests and the bench: nqr::NyquistIO::Load (the reference's own Common.cpp, unmodified)
// on top of the two-phase OpusDecoder.  ctypes cannot call C++ directly.
#include "Decoders.h"

#include <cstdlib>
#include <cstring>
#include <iostream>

extern "C" {

extern void nq_twophase_last_timing(double out[3]);

// Returns 0 and a malloc'ed interleaved float buffer the caller frees with nq_twophase_free;
// -1 and a message on stderr if Load throws.
__attribute__((visibility("default"))) int nq_twophase_load(const char *path, float **samples, size_t *count,
                                                             int *channels, int *sample_rate, double timing[3])
{
    try {
        nqr::NyquistIO loader;
        nqr::AudioData data;
        loader.Load(&data, std::string(path));
        *count = data.samples.size();
        *channels = data.channelCount;
        *sample_rate = data.sampleRate;
        *samples = (float *)malloc(sizeof(float) * data.samples.size());
        memcpy(*samples, data.samples.data(), sizeof(float) * data.samples.size());
        if (timing) nq_twophase_last_timing(timing);
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "nq_twophase_load: " << e.what() << std::endl;
        return -1;
    }
}

__attribute__((visibility("default"))) void nq_twophase_free(float *p) { free(p); }

}  // extern "C"
