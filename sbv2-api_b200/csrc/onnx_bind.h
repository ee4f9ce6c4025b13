// Structural binding of anonymous initializers.
//
// The reference's export scripts trace the PyTorch modules with TorchScript and run onnxsim
// (scripts/convert/convert_model.py:115-156, convert_deberta.py:36-52).  Constant folding then renames
//   * every weight-normed Conv1d / ConvTranspose1d weight (decoder ups / resblocks, WN flow layers) to onnx::Conv_N /
//     onnx::ConvTranspose_N, and
//   * every nn.Linear applied to a 3-D input (all of DeBERTa's q/k/v/o/FFN projections, the encoders' spk_emb_linear)
//     to a TRANSPOSED [in, out] onnx::MatMul_N initializer feeding MatMul -> Add,
// while biases, embeddings, LayerNorm parameters and plain Conv weights keep their dotted PyTorch names
// (SURVEY.md §A.7).  The loaders ask for weights by their PyTorch names; this class answers those requests by walking the
// graph: an anonymous weight is identified by the still-named bias it is used with
//   Conv / ConvTranspose (x, W, "<p>.bias")            ->  "<p>.weight" = W
//   Gemm (x, W, "<p>.bias"), transB = 1 / 0             ->  "<p>.weight" = W / W^T
//   MatMul (x, W) -> Add ("<p>.bias", .) (either order)  ->  "<p>.weight" = W^T
// Transposed bindings are materialised once in PyTorch layout ([out, in]) so that callers see dims and data exactly as a
// state_dict has them.  A "<p>.bias" whose "<p>.weight" cannot be found either way is a hard error (a silently skipped
// Linear — e.g. the speaker conditioning of the encoders — would change the output instead of failing the load).
#pragma once
#include <map>
#include <set>
#include <string>

#include "onnx_reader.h"

namespace sbv2 {

class WeightBinder {
 public:
  explicit WeightBinder(const OnnxModel& m);
  // canonical (PyTorch) name -> tensor, or null
  const OnnxTensor* find(const std::string& name) const;
  bool has(const std::string& name) const { return find(name) != nullptr; }
  // manual alias (the decoder's sequence walker for convs without a named bias)
  void alias_to(const std::string& canonical, const std::string& initializer) { alias_[canonical] = initializer; }
  bool any_structural() const { return !alias_.empty() || !materialized_.empty(); }
  // every "<p>.bias" under `prefix` must have a "<p>.weight" (named or bound); throws SBV2_ERR_UNSUPPORTED otherwise
  void require_weights_for_biases(const std::string& prefix) const;
  // {"bound": {"<canonical>": {"initializer": "...", "transposed": bool, "via": "Conv|Gemm|MatMul+Add|sequence"}}, ...}
  std::string report_json() const;
  const OnnxModel& model() const { return m_; }

 private:
  void bind(const std::string& bias_name, const std::string& init, bool transposed, const char* via);
  const OnnxModel& m_;
  std::map<std::string, std::string> alias_;          // canonical -> initializer (same layout)
  std::map<std::string, OnnxTensor> materialized_;    // canonical -> transposed copy in PyTorch layout
  std::map<std::string, std::pair<std::string, std::string>> how_;  // canonical -> (initializer, via)
};

}  // namespace sbv2
