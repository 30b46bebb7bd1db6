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
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

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

// K loads of the unmodified reference at the same time, one thread each (`threads` at a time): what a
// program that needs K files does with the reference on the same cores.  Returns the wall seconds,
// or a negative number if a Load threw.  The samples are dropped.
__attribute__((visibility("default"))) double nqref_load_many(const char *const *paths, int n, int threads)
{
    std::atomic<int> next{0}, failed{0};
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    if (threads < 1) threads = 1;
    if (threads > n) threads = n;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&] {
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= n) return;
                try {
                    nqr::NyquistIO loader;
                    nqr::AudioData data;
                    loader.Load(&data, std::string(paths[i]));
                } catch (const std::exception &) {
                    failed.fetch_add(1);
                }
            }
        });
    for (std::thread &t : pool) t.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return failed.load() ? -1.0 : dt;
}

__attribute__((visibility("default"))) void nqref_free(float *p)
{
    std::lock_guard<std::mutex> lk(g_mu);
    g_live.erase(p);
}

}  // extern "C"
