// Minimal ONNX ModelProto reader (protobuf wire format, no dependency on the onnx package).
// Reads what the loader needs: initializers (name, dims, dtype, payload), graph nodes (op type,
// inputs/outputs, integer attributes — for structural binding of anonymous weights) and
// metadata_props (for .aivmx style vectors, reference crates/sbv2_core/src/tts.rs:93-94).
// Field numbers: SURVEY.md §A.7 (onnx.proto3).
#pragma once
#include "common.h"

namespace sbv2 {

enum OnnxDType { ONNX_FLOAT = 1, ONNX_INT32 = 6, ONNX_INT64 = 7, ONNX_FLOAT16 = 10, ONNX_DOUBLE = 11, ONNX_BFLOAT16 = 16 };

struct OnnxTensor {
  std::string name;
  std::vector<int64_t> dims;
  int dtype = 0;
  // Payload view into the model bytes (raw_data or packed float_data) or owned storage.
  const uint8_t* data = nullptr;
  size_t nbytes = 0;
  std::vector<uint8_t> owned;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : dims) n *= d;
    return n;
  }
};

struct OnnxNode {
  std::string op_type, name;
  std::vector<std::string> inputs, outputs;
  std::map<std::string, std::vector<int64_t>> int_attrs;  // i and ints
};

struct OnnxModel {
  std::vector<OnnxTensor> initializers;
  std::map<std::string, size_t> by_name;
  std::vector<OnnxNode> nodes;
  std::map<std::string, std::string> metadata;
  std::vector<std::string> graph_inputs, graph_outputs;
  std::string producer;
  int64_t ir_version = 0;

  const OnnxTensor* find(const std::string& n) const {
    auto it = by_name.find(n);
    return it == by_name.end() ? nullptr : &initializers[it->second];
  }
  // Converts any supported float payload to float32.
  std::vector<float> as_f32(const OnnxTensor& t) const;
};

// `bytes` must outlive the returned model (tensor payloads are views).
OnnxModel parse_onnx(const uint8_t* bytes, size_t n);

}  // namespace sbv2
