// Many Opus files in one call: the producer of the many-streams phase 2 (SURVEY.md section 8(f)
// rows 1-2).  The reference decodes a file with the loop src/OpusDecoder.cpp:101-119 and a program
// that loads K files runs that loop K times; here phase 1 (entropy decode, CPU) of the K files runs
// on up to `threads` host threads at once, and phase 2 of ALL of them is one synthesis launch and
// one post-filter launch on the B200 (nq_celt_sink_flush_many: every file is a segment that starts
// from a reset decoder).  Same AudioData contents as nqr::NyquistIO::Load file by file.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "Common.h"

namespace nqr
{

struct OpusBatchStats {
    double phase1Seconds = 0;   // wall time of the parallel entropy decode of all files
    double phase2Seconds = 0;   // gather + the two launches + copy back, all files
    double totalSeconds = 0;
    int filesBatched = 0;       // CELT-only, single-link files: decoded by the batched phase 2
    int filesSingle = 0;        // everything else (SILK / hybrid / mode switches): the ordinary two-phase Load, one by one
    long long frames = 0;       // CELT frames in the batched phase 2
    int launches = 0;           // kernel launches of the batched phase 2
};

// out[i] receives paths[i]; throws (like Load) if any file cannot be decoded.  threads <= 0: one per
// hardware thread, at most one per file.
OpusBatchStats LoadOpusBatch(const std::vector<std::string> &paths, std::vector<std::shared_ptr<AudioData>> &out, int threads = 0);

}   // namespace nqr
