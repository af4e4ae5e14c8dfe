// TEST INFRASTRUCTURE ONLY: drives the reference's own wespeaker::SpeakerEngine (runtime/core/speaker/speaker_engine.cc,
// compiled unmodified) with wespeaker::B200SpeakerModel plugged into its `model_` slot, exactly where a maintainer's
// `#elif defined(USE_B200)` branch would put it: reference fbank -> reference chunking -> reference ApplyMean ->
// SpeakerModel::ExtractEmbedding (virtual) -> B200 engine -> reference averaging.
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

#define private public   // the model slot is private; the reference fills it inside its constructor from compile-time macros
#include "speaker/speaker_engine.h"
#undef private

#include "b200_speaker_model.h"

extern "C" int ref_engine_extract_with_b200(const char* flat_model_path, int device, const short* pcm, int nsamples,
                                            int samples_per_chunk, float* emb_out, int emb_size) {
    auto model = std::make_shared<wespeaker::B200SpeakerModel>(flat_model_path, device);
    if (model->EmbeddingSize() != emb_size) return -1;
    wespeaker::SpeakerEngine eng("", 80, 16000, emb_size, samples_per_chunk);
    eng.model_ = model;
    std::vector<float> emb;
    eng.ExtractEmbedding(pcm, nsamples, &emb);
    if ((int)emb.size() != emb_size) return -2;
    std::memcpy(emb_out, emb.data(), sizeof(float) * emb.size());
    return 0;
}
