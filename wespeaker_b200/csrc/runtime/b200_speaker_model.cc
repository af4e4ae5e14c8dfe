// See b200_speaker_model.h.  Host-only C++ over the C ABI of libwespeaker_b200.so (no CUDA headers needed here).
#include "b200_speaker_model.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "../../../include/wespeaker_b200.h"

namespace wespeaker {

namespace {
[[noreturn]] void Fatal(const std::string& what) {
  const char* e = ws_last_error();
  std::fprintf(stderr, "B200SpeakerModel: %s%s%s\n", what.c_str(), (e && *e) ? ": " : "", (e && *e) ? e : "");
  std::abort();   // the reference runtime reports fatal errors through glog CHECK / LOG(FATAL), which abort
}
template <class T>
T ReadPod(std::ifstream& f) {
  T v{};
  f.read(reinterpret_cast<char*>(&v), sizeof(T));
  if (!f) Fatal("truncated model file");
  return v;
}
std::string ReadStr(std::ifstream& f) {
  const uint32_t n = ReadPod<uint32_t>(f);
  if (n > (1u << 20)) Fatal("corrupt model file (string length)");
  std::string s(n, '\0');
  f.read(&s[0], n);
  if (!f) Fatal("truncated model file");
  return s;
}
}  // namespace

B200SpeakerModel::B200SpeakerModel(const std::string& model_path, int device) {
  std::ifstream f(model_path, std::ios::binary);
  if (!f) Fatal("cannot open " + model_path);
  char magic[8];
  f.read(magic, 8);
  if (!f || std::string(magic, 8) != "WSPKB200") Fatal(model_path + " is not a wespeaker_b200 flat model file");
  if (ReadPod<uint32_t>(f) != 1u) Fatal("unsupported flat model version");
  const std::string model = ReadStr(f), precision = ReadStr(f);
  feat_dim_ = ReadPod<int32_t>(f);
  embed_dim_ = ReadPod<int32_t>(f);
  if (ws_engine_create(model.c_str(), precision.c_str(), feat_dim_, embed_dim_, device, &engine_) != 0) Fatal("ws_engine_create");
  const uint32_t nopt = ReadPod<uint32_t>(f);
  for (uint32_t i = 0; i < nopt; ++i) {
    const std::string key = ReadStr(f);
    const long long value = ReadPod<int64_t>(f);
    if (ws_engine_set_option(engine_, key.c_str(), value) != 0) Fatal("ws_engine_set_option " + key);
  }
  const uint32_t nt = ReadPod<uint32_t>(f);
  std::vector<float> data;
  for (uint32_t i = 0; i < nt; ++i) {
    const std::string name = ReadStr(f);
    const uint32_t ndim = ReadPod<uint32_t>(f);
    if (ndim > 8) Fatal("corrupt model file (ndim)");
    long long dims[8] = {0}, n = 1;
    for (uint32_t d = 0; d < ndim; ++d) { dims[d] = ReadPod<int64_t>(f); n *= dims[d]; }
    data.resize((size_t)n);
    f.read(reinterpret_cast<char*>(data.data()), (std::streamsize)(n * 4));
    if (!f) Fatal("truncated model file");
    if (ws_engine_set_tensor(engine_, name.c_str(), data.data(), dims, (int)ndim) != 0) Fatal("ws_engine_set_tensor " + name);
  }
  if (ws_engine_finalize(engine_) != 0) Fatal("ws_engine_finalize");
}

B200SpeakerModel::~B200SpeakerModel() {
  if (engine_) ws_engine_destroy(engine_);
}

void B200SpeakerModel::ExtractEmbedding(const std::vector<std::vector<float>>& feats, std::vector<float>* embed) {
  const int T = static_cast<int>(feats.size());
  if (T == 0 || embed == nullptr) Fatal("ExtractEmbedding: empty features");
  flat_.resize((size_t)T * feat_dim_);
  for (int t = 0; t < T; ++t) {
    if (static_cast<int>(feats[t].size()) != feat_dim_) Fatal("ExtractEmbedding: feature dimension mismatch");
    std::copy(feats[t].begin(), feats[t].end(), flat_.begin() + (size_t)t * feat_dim_);
  }
  embed->resize(embed_dim_);
  if (ws_engine_forward_host(engine_, flat_.data(), 1, T, embed->data()) != 0) Fatal("ws_engine_forward_host");
}

}  // namespace wespeaker
