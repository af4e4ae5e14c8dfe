// B200 back-end for the reference's C++ runtime seam:
//
//     class SpeakerModel { virtual void ExtractEmbedding(const std::vector<std::vector<float>>& feats,
//                                                        std::vector<float>* embed); };
//                                                  -- /root/reference/runtime/core/speaker/speaker_model.h:25-32
//
// `B200SpeakerModel` is what a maintainer selects next to OnnxSpeakerModel / MnnSpeakerModel in
// runtime/core/speaker/speaker_engine.cc:47-58 (`#elif defined(USE_B200)  model_ = std::make_shared<B200SpeakerModel>(path);`).
// It owns one ws_engine (include/wespeaker_b200.h) and feeds it the reference's [T][F] feature rows; batch = 1, caller owns
// both vectors, fatal errors abort with a message (the reference's glog CHECK / LOG(FATAL) behaviour).
//
// Weights come from a flat file written by `wespeaker_b200.models.B200SpeakerModel.export_flat(path)`:
//   "WSPKB200" | u32 version=1 | str model | str precision | i32 feat_dim | i32 embed_dim | u32 n_options {str key, i64 value}
//   | u32 n_tensors { str name | u32 ndim | i64 dims[ndim] | f32 data[prod(dims)] }        (str = u32 length + bytes)
#ifndef WESPEAKER_B200_RUNTIME_B200_SPEAKER_MODEL_H_
#define WESPEAKER_B200_RUNTIME_B200_SPEAKER_MODEL_H_

#include <string>
#include <vector>

#include "speaker/speaker_model.h"   // the reference's abstract class (runtime/core on the include path)

struct ws_engine;

namespace wespeaker {

class B200SpeakerModel : public SpeakerModel {
 public:
  explicit B200SpeakerModel(const std::string& model_path, int device = 0);
  ~B200SpeakerModel() override;
  B200SpeakerModel(const B200SpeakerModel&) = delete;
  B200SpeakerModel& operator=(const B200SpeakerModel&) = delete;

  void ExtractEmbedding(const std::vector<std::vector<float>>& feats, std::vector<float>* embed) override;
  int EmbeddingSize() const { return embed_dim_; }

 private:
  ws_engine* engine_ = nullptr;
  int feat_dim_ = 0, embed_dim_ = 0;
  std::vector<float> flat_;   // [T][F] staging of one utterance
};

}  // namespace wespeaker

#endif  // WESPEAKER_B200_RUNTIME_B200_SPEAKER_MODEL_H_
