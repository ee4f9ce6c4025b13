// Asset containers around the ONNX graphs: .sbv2 (zstd(tar)), style_vectors.json, .aivmx
// metadata (base64 .npy) and the float32 WAV writer.  Host-only code.
// Reference: crates/sbv2_core/src/sbv2file.rs:15-37, style.rs:5-28, tts.rs:95-108,
// tts_util.rs:163-180.
#pragma once
#include "common.h"

namespace sbv2 {

std::vector<uint8_t> zstd_decompress(const uint8_t* p, size_t n);
std::vector<uint8_t> zstd_compress(const uint8_t* p, size_t n, int level);

struct TarEntry {
  std::string name;
  const uint8_t* data;
  size_t size;
};
std::vector<TarEntry> tar_entries(const uint8_t* p, size_t n);

struct Sbv2File {
  std::vector<uint8_t> tar;  // decompressed archive; entries below point into it
  const uint8_t* style_json = nullptr;
  size_t style_n = 0;
  const uint8_t* onnx = nullptr;
  size_t onnx_n = 0;
};
Sbv2File parse_sbv2file(const uint8_t* p, size_t n);

struct StyleVectors {
  int64_t rows = 0, cols = 0;
  std::vector<float> data;  // row-major
};
StyleVectors load_style_json(const uint8_t* p, size_t n);
StyleVectors load_style_npy_base64(const char* b64, size_t n);
std::vector<float> get_style_vector(const StyleVectors& s, int32_t style_id, float weight);

std::vector<uint8_t> base64_decode(const char* p, size_t n);
std::vector<uint8_t> wav_from_f32(const float* samples, int64_t n);
std::vector<uint8_t> wav_pcm16_from_f32(const float* samples, int64_t n);

}  // namespace sbv2
