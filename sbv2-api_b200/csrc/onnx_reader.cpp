#include "onnx_reader.h"

#include <cmath>

namespace sbv2 {
namespace {

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    while (true) {
      if (p >= end) fail(SBV2_ERR_PARSE, "onnx: truncated varint");
      uint8_t b = *p++;
      v |= uint64_t(b & 0x7f) << shift;
      if (!(b & 0x80)) break;
      shift += 7;
      if (shift > 63) fail(SBV2_ERR_PARSE, "onnx: varint too long");
    }
    return v;
  }
  Reader sub() {
    uint64_t n = varint();
    if (n > uint64_t(end - p)) fail(SBV2_ERR_PARSE, "onnx: length-delimited field exceeds buffer");
    Reader r{p, p + n};
    p += n;
    return r;
  }
  void skip(int wt) {
    switch (wt) {
      case 0: varint(); break;
      case 1:
        if (end - p < 8) fail(SBV2_ERR_PARSE, "onnx: truncated fixed64");
        p += 8;
        break;
      case 2: sub(); break;
      case 5:
        if (end - p < 4) fail(SBV2_ERR_PARSE, "onnx: truncated fixed32");
        p += 4;
        break;
      default: fail(SBV2_ERR_PARSE, "onnx: unsupported wire type " + std::to_string(wt));
    }
  }
  std::string str() {
    Reader r = sub();
    return std::string(reinterpret_cast<const char*>(r.p), r.end - r.p);
  }
};

void parse_tensor(Reader r, OnnxTensor& t) {
  std::vector<int32_t> i32;
  std::vector<int64_t> i64;
  const uint8_t* fdata = nullptr;
  size_t fbytes = 0;
  std::vector<float> fsingle;
  while (!r.done()) {
    uint64_t key = r.varint();
    int field = int(key >> 3), wt = int(key & 7);
    if (field == 1) {  // dims
      if (wt == 2) {
        Reader s = r.sub();
        while (!s.done()) t.dims.push_back(int64_t(s.varint()));
      } else {
        t.dims.push_back(int64_t(r.varint()));
      }
    } else if (field == 2 && wt == 0) {
      t.dtype = int(r.varint());
    } else if (field == 4) {  // float_data
      if (wt == 2) {
        Reader s = r.sub();
        fdata = s.p;
        fbytes = size_t(s.end - s.p);
      } else if (wt == 5) {
        float f;
        if (r.end - r.p < 4) fail(SBV2_ERR_PARSE, "onnx: truncated float");
        memcpy(&f, r.p, 4);
        r.p += 4;
        fsingle.push_back(f);
      } else {
        r.skip(wt);
      }
    } else if (field == 5) {  // int32_data
      if (wt == 2) {
        Reader s = r.sub();
        while (!s.done()) i32.push_back(int32_t(s.varint()));
      } else {
        i32.push_back(int32_t(r.varint()));
      }
    } else if (field == 7) {  // int64_data
      if (wt == 2) {
        Reader s = r.sub();
        while (!s.done()) i64.push_back(int64_t(s.varint()));
      } else {
        i64.push_back(int64_t(r.varint()));
      }
    } else if (field == 8 && wt == 2) {
      t.name = r.str();
    } else if (field == 9 && wt == 2) {  // raw_data
      Reader s = r.sub();
      t.data = s.p;
      t.nbytes = size_t(s.end - s.p);
    } else if (field == 14 && wt == 0) {  // data_location
      if (r.varint() != 0) fail(SBV2_ERR_UNSUPPORTED, "onnx: external tensor data is not supported (" + t.name + ")");
    } else {
      r.skip(wt);
    }
  }
  if (!t.data) {
    if (fdata) {
      t.data = fdata;
      t.nbytes = fbytes;
    } else if (!fsingle.empty()) {
      t.owned.resize(fsingle.size() * 4);
      memcpy(t.owned.data(), fsingle.data(), t.owned.size());
    } else if (!i64.empty()) {
      t.owned.resize(i64.size() * 8);
      memcpy(t.owned.data(), i64.data(), t.owned.size());
    } else if (!i32.empty()) {
      if (t.dtype == ONNX_FLOAT16 || t.dtype == ONNX_BFLOAT16) {
        t.owned.resize(i32.size() * 2);
        for (size_t i = 0; i < i32.size(); ++i) {
          uint16_t h = uint16_t(i32[i]);
          memcpy(t.owned.data() + 2 * i, &h, 2);
        }
      } else {
        t.owned.resize(i32.size() * 4);
        memcpy(t.owned.data(), i32.data(), t.owned.size());
      }
    }
    if (!t.owned.empty()) {
      t.nbytes = t.owned.size();
    }
  }
}

void parse_attribute(Reader r, OnnxNode& n) {
  std::string name;
  std::vector<int64_t> ints;
  bool has = false;
  while (!r.done()) {
    uint64_t key = r.varint();
    int field = int(key >> 3), wt = int(key & 7);
    if (field == 1 && wt == 2) {
      name = r.str();
    } else if (field == 3 && wt == 0) {
      ints.push_back(int64_t(r.varint()));
      has = true;
    } else if (field == 8) {
      has = true;
      if (wt == 2) {
        Reader s = r.sub();
        while (!s.done()) ints.push_back(int64_t(s.varint()));
      } else {
        ints.push_back(int64_t(r.varint()));
      }
    } else {
      r.skip(wt);
    }
  }
  if (has) n.int_attrs[name] = ints;
}

void parse_node(Reader r, OnnxNode& n) {
  while (!r.done()) {
    uint64_t key = r.varint();
    int field = int(key >> 3), wt = int(key & 7);
    if (field == 1 && wt == 2) n.inputs.push_back(r.str());
    else if (field == 2 && wt == 2) n.outputs.push_back(r.str());
    else if (field == 3 && wt == 2) n.name = r.str();
    else if (field == 4 && wt == 2) n.op_type = r.str();
    else if (field == 5 && wt == 2) parse_attribute(r.sub(), n);
    else r.skip(wt);
  }
}

std::string parse_value_info_name(Reader r) {
  std::string name;
  while (!r.done()) {
    uint64_t key = r.varint();
    int field = int(key >> 3), wt = int(key & 7);
    if (field == 1 && wt == 2) name = r.str();
    else r.skip(wt);
  }
  return name;
}

void parse_graph(Reader r, OnnxModel& m) {
  while (!r.done()) {
    uint64_t key = r.varint();
    int field = int(key >> 3), wt = int(key & 7);
    if (field == 1 && wt == 2) {
      m.nodes.emplace_back();
      parse_node(r.sub(), m.nodes.back());
    } else if (field == 5 && wt == 2) {
      m.initializers.emplace_back();
      parse_tensor(r.sub(), m.initializers.back());
    } else if (field == 11 && wt == 2) {
      m.graph_inputs.push_back(parse_value_info_name(r.sub()));
    } else if (field == 12 && wt == 2) {
      m.graph_outputs.push_back(parse_value_info_name(r.sub()));
    } else {
      r.skip(wt);
    }
  }
}

float half_to_float(uint16_t h) {
  uint32_t sign = uint32_t(h & 0x8000) << 16;
  uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ff;
  uint32_t f;
  if (exp == 0) {
    if (man == 0) {
      f = sign;
    } else {
      int e = -1;
      do {
        man <<= 1;
        ++e;
      } while (!(man & 0x400));
      f = sign | uint32_t(127 - 15 - e) << 23 | (man & 0x3ff) << 13;
    }
  } else if (exp == 31) {
    f = sign | 0x7f800000u | man << 13;
  } else {
    f = sign | (exp + 112) << 23 | man << 13;
  }
  float out;
  memcpy(&out, &f, 4);
  return out;
}

}  // namespace

OnnxModel parse_onnx(const uint8_t* bytes, size_t n) {
  if (!bytes || n == 0) fail(SBV2_ERR_PARSE, "onnx: empty model bytes");
  OnnxModel m;
  Reader r{bytes, bytes + n};
  bool saw_graph = false;
  while (!r.done()) {
    uint64_t key = r.varint();
    int field = int(key >> 3), wt = int(key & 7);
    if (field == 0) fail(SBV2_ERR_PARSE, "onnx: invalid field number 0 (not a ModelProto)");
    if (field == 1 && wt == 0) {
      m.ir_version = int64_t(r.varint());
    } else if (field == 2 && wt == 2) {
      m.producer = r.str();
    } else if (field == 7 && wt == 2) {
      parse_graph(r.sub(), m);
      saw_graph = true;
    } else if (field == 14 && wt == 2) {
      Reader s = r.sub();
      std::string k, v;
      while (!s.done()) {
        uint64_t key2 = s.varint();
        int f2 = int(key2 >> 3), w2 = int(key2 & 7);
        if (f2 == 1 && w2 == 2) k = s.str();
        else if (f2 == 2 && w2 == 2) v = s.str();
        else s.skip(w2);
      }
      m.metadata[k] = v;
    } else {
      r.skip(wt);
    }
  }
  if (!saw_graph) fail(SBV2_ERR_PARSE, "onnx: ModelProto has no graph");
  for (size_t i = 0; i < m.initializers.size(); ++i) {
    auto& t = m.initializers[i];
    if (!t.owned.empty()) t.data = t.owned.data();
    m.by_name[t.name] = i;
  }
  return m;
}

std::vector<float> OnnxModel::as_f32(const OnnxTensor& t) const {
  int64_t n = t.numel();
  std::vector<float> out(size_t(n > 0 ? n : 0));
  auto need = [&](size_t elt) {
    if (t.nbytes != size_t(n) * elt)
      fail(SBV2_ERR_PARSE, "onnx: initializer '" + t.name + "' payload size " + std::to_string(t.nbytes) +
                               " does not match dims (" + std::to_string(n) + " elements)");
  };
  switch (t.dtype) {
    case ONNX_FLOAT:
      need(4);
      memcpy(out.data(), t.data, t.nbytes);
      break;
    case ONNX_DOUBLE:
      need(8);
      for (int64_t i = 0; i < n; ++i) {
        double d;
        memcpy(&d, t.data + 8 * i, 8);
        out[i] = float(d);
      }
      break;
    case ONNX_FLOAT16:
      need(2);
      for (int64_t i = 0; i < n; ++i) {
        uint16_t h;
        memcpy(&h, t.data + 2 * i, 2);
        out[i] = half_to_float(h);
      }
      break;
    case ONNX_BFLOAT16:
      need(2);
      for (int64_t i = 0; i < n; ++i) {
        uint16_t h;
        memcpy(&h, t.data + 2 * i, 2);
        uint32_t u = uint32_t(h) << 16;
        memcpy(&out[i], &u, 4);
      }
      break;
    default:
      fail(SBV2_ERR_UNSUPPORTED, "onnx: initializer '" + t.name + "' has non-float dtype " + std::to_string(t.dtype));
  }
  return out;
}

}  // namespace sbv2
