// TEST INFRASTRUCTURE ONLY: C entry points over the UNMODIFIED reference C++ classes, compiled into oracle/_ref/ by
// oracle/Makefile from the sources under /root/reference/runtime/core.  They expose
//   * wespeaker::SpeakerEngine::ExtractFeature (speaker_engine.cc:77-139): the reference's own fbank (frontend/fbank.h,
//     fft.cc) + the chunk-and-pad logic of the long-audio mode, and ApplyMean (:62-75);
// so that oracle/speaker_engine_py.py (the Python restatement used on the GPU box) and oracle/fbank_np.py are pinned
// against outputs of the reference itself (tests/golden/make_golden_ref_engine.py stores them as fixtures).
// every standard header the reference headers pull in is included BEFORE the access hack below (include guards then keep
// them out of its reach)
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <queue>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include "glog/logging.h"

#define private public   // ApplyMean is private in the reference header; the wrapper only calls it
#include "speaker/speaker_engine.h"
#undef private

extern "C" {

// Returns the number of chunks; fills shape[0..2] = {chunks, frames per chunk, feat_dim}.  *out is malloc'ed
// [chunks][frames][dim] floats (caller frees with ref_free).  apply_mean != 0 runs SpeakerEngine::ApplyMean per chunk.
int ref_extract_feature(const short* pcm, int nsamples, int samples_per_chunk, int apply_mean, float** out, int* shape) {
    wespeaker::SpeakerEngine eng("", 80, 16000, 256, samples_per_chunk);
    std::vector<std::vector<std::vector<float>>> chunks;
    eng.ExtractFeature(pcm, nsamples, &chunks);
    const int nc = (int)chunks.size();
    shape[0] = nc; shape[1] = nc ? (int)chunks[0].size() : 0; shape[2] = (nc && shape[1]) ? (int)chunks[0][0].size() : 0;
    for (auto& c : chunks)
        if ((int)c.size() != shape[1]) return -1;          // ragged chunk lists (full mode has exactly one)
    *out = (float*)malloc(sizeof(float) * (size_t)(nc ? nc : 1) * (shape[1] ? shape[1] : 1) * (shape[2] ? shape[2] : 1));
    float* p = *out;
    for (auto& c : chunks) {
        if (apply_mean) eng.ApplyMean(&c, (unsigned)shape[2]);
        for (auto& row : c) { memcpy(p, row.data(), sizeof(float) * row.size()); p += row.size(); }
    }
    return nc;
}

float ref_cosine_similarity(const float* a, const float* b, int n) {
    wespeaker::SpeakerEngine eng("", 80, 16000, n, 0);
    return eng.CosineSimilarity(std::vector<float>(a, a + n), std::vector<float>(b, b + n));
}

void ref_free(float* p) { free(p); }

}  // extern "C"
