#include "onnx_bind.h"

#include <sstream>

namespace sbv2 {
namespace {
bool ends_with(const std::string& s, const char* suf) {
  const size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}
int64_t attr_i(const OnnxNode& n, const char* name, int64_t dflt) {
  auto it = n.int_attrs.find(name);
  return it == n.int_attrs.end() || it->second.empty() ? dflt : it->second[0];
}
}  // namespace

WeightBinder::WeightBinder(const OnnxModel& m) : m_(m) {
  std::map<std::string, size_t> producer;  // value name -> node index
  for (size_t i = 0; i < m.nodes.size(); ++i)
    for (const auto& o : m.nodes[i].outputs) producer[o] = i;
  auto named_bias = [&](const std::string& v) -> const OnnxTensor* {
    const OnnxTensor* t = m.find(v);
    return t && ends_with(v, ".bias") && t->dims.size() == 1 ? t : nullptr;
  };
  for (const OnnxNode& n : m.nodes) {
    if ((n.op_type == "Conv" || n.op_type == "ConvTranspose") && n.inputs.size() >= 3) {
      const OnnxTensor* w = m.find(n.inputs[1]);
      if (w && named_bias(n.inputs[2])) bind(n.inputs[2], n.inputs[1], false, n.op_type.c_str());
    } else if (n.op_type == "Gemm" && n.inputs.size() >= 3) {
      const OnnxTensor* w = m.find(n.inputs[1]);
      const OnnxTensor* b = named_bias(n.inputs[2]);
      if (w && b && w->dims.size() == 2 && attr_i(n, "transA", 0) == 0) {
        const bool trans_b = attr_i(n, "transB", 0) != 0;  // transB = 1: W is [out, in] as PyTorch stores it
        const int64_t out_dim = trans_b ? w->dims[0] : w->dims[1];
        if (out_dim == b->dims[0]) bind(n.inputs[2], n.inputs[1], !trans_b, "Gemm");
      }
    } else if (n.op_type == "Add" && n.inputs.size() == 2) {
      for (int side = 0; side < 2; ++side) {
        const OnnxTensor* b = named_bias(n.inputs[side]);
        if (!b) continue;
        auto p = producer.find(n.inputs[1 - side]);
        if (p == producer.end()) continue;
        const OnnxNode& mm = m.nodes[p->second];
        if (mm.op_type != "MatMul" || mm.inputs.size() != 2) continue;
        const OnnxTensor* w = m.find(mm.inputs[1]);
        if (w && w->dims.size() == 2 && w->dims[1] == b->dims[0]) bind(n.inputs[side], mm.inputs[1], true, "MatMul+Add");
      }
    }
  }
}

void WeightBinder::bind(const std::string& bias_name, const std::string& init, bool transposed, const char* via) {
  const std::string canonical = bias_name.substr(0, bias_name.size() - 5) + ".weight";
  if (init == canonical && !transposed) return;       // already named as PyTorch names it
  if (m_.find(canonical) && init != canonical) return;  // a properly named weight exists: the name wins
  if (alias_.count(canonical) || materialized_.count(canonical)) return;  // first use wins (shared weights)
  how_[canonical] = std::make_pair(init, std::string(via));
  if (!transposed) {
    alias_[canonical] = init;
    return;
  }
  const OnnxTensor& w = *m_.find(init);
  const int64_t rows = w.dims[0], cols = w.dims[1];  // stored [in, out]
  std::vector<float> src = m_.as_f32(w);
  OnnxTensor t;
  t.name = canonical;
  t.dims = {cols, rows};  // PyTorch layout [out, in]
  t.dtype = ONNX_FLOAT;
  t.owned.resize(size_t(rows) * size_t(cols) * 4);
  float* dst = reinterpret_cast<float*>(t.owned.data());
  for (int64_t i = 0; i < rows; ++i)
    for (int64_t o = 0; o < cols; ++o) dst[size_t(o) * size_t(rows) + size_t(i)] = src[size_t(i) * size_t(cols) + size_t(o)];
  t.nbytes = t.owned.size();
  auto it = materialized_.emplace(canonical, std::move(t)).first;
  it->second.data = it->second.owned.data();  // the vector's buffer survives the move; re-point for clarity
}

const OnnxTensor* WeightBinder::find(const std::string& name) const {
  auto mt = materialized_.find(name);
  if (mt != materialized_.end()) return &mt->second;
  auto a = alias_.find(name);
  return m_.find(a == alias_.end() ? name : a->second);
}

void WeightBinder::require_weights_for_biases(const std::string& prefix) const {
  for (const OnnxTensor& t : m_.initializers) {
    if (t.name.compare(0, prefix.size(), prefix) != 0 || !ends_with(t.name, ".bias") || t.dims.size() != 1) continue;
    const std::string w = t.name.substr(0, t.name.size() - 5) + ".weight";
    if (!has(w))
      fail(SBV2_ERR_UNSUPPORTED, "initializer '" + t.name + "' has no matching '" + w +
                                     "': its weight is anonymous in this export and no Conv / Gemm / MatMul+Add node binds it");
  }
}

std::string WeightBinder::report_json() const {
  std::ostringstream js;
  js << "{\"bound\":{";
  bool first = true;
  for (const auto& kv : how_) {
    js << (first ? "" : ",") << "\"" << kv.first << "\":{\"initializer\":\"" << kv.second.first << "\",\"transposed\":"
       << (materialized_.count(kv.first) ? "true" : "false") << ",\"via\":\"" << kv.second.second << "\"}";
    first = false;
  }
  for (const auto& kv : alias_)
    if (!how_.count(kv.first)) {
      js << (first ? "" : ",") << "\"" << kv.first << "\":{\"initializer\":\"" << kv.second << "\",\"transposed\":false,\"via\":\"sequence\"}";
      first = false;
    }
  js << "},\"unbound_biases\":[";
  first = true;
  for (const OnnxTensor& t : m_.initializers) {
    if (!ends_with(t.name, ".bias") || t.dims.size() != 1) continue;
    if (has(t.name.substr(0, t.name.size() - 5) + ".weight")) continue;
    js << (first ? "" : ",") << "\"" << t.name << "\"";
    first = false;
  }
  js << "],\"n_initializers\":" << m_.initializers.size() << ",\"n_nodes\":" << m_.nodes.size() << "}";
  return js.str();
}

}  // namespace sbv2
