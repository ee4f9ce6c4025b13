#include "container.h"

#include <dlfcn.h>

#include <cerrno>
#include <cmath>
#include <mutex>

namespace sbv2 {
namespace {

// libzstd.so.1 ships in the image without headers; bind the four stable entry points by hand.
struct Zstd {
  void* handle = nullptr;
  unsigned long long (*getFrameContentSize)(const void*, size_t) = nullptr;
  size_t (*decompress)(void*, size_t, const void*, size_t) = nullptr;
  size_t (*compress)(void*, size_t, const void*, size_t, int) = nullptr;
  size_t (*compressBound)(size_t) = nullptr;
  unsigned (*isError)(size_t) = nullptr;
  const char* (*getErrorName)(size_t) = nullptr;
  // streaming (frames without a content size)
  void* (*createDStream)() = nullptr;
  size_t (*freeDStream)(void*) = nullptr;
  size_t (*decompressStream)(void*, void*, void*) = nullptr;
};

Zstd& zstd() {
  static Zstd z;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libzstd.so.1", "libzstd.so"}) {
      z.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (z.handle) break;
    }
    if (!z.handle) return;
    auto sym = [&](const char* s) { return dlsym(z.handle, s); };
    z.getFrameContentSize = reinterpret_cast<decltype(z.getFrameContentSize)>(sym("ZSTD_getFrameContentSize"));
    z.decompress = reinterpret_cast<decltype(z.decompress)>(sym("ZSTD_decompress"));
    z.compress = reinterpret_cast<decltype(z.compress)>(sym("ZSTD_compress"));
    z.compressBound = reinterpret_cast<decltype(z.compressBound)>(sym("ZSTD_compressBound"));
    z.isError = reinterpret_cast<decltype(z.isError)>(sym("ZSTD_isError"));
    z.getErrorName = reinterpret_cast<decltype(z.getErrorName)>(sym("ZSTD_getErrorName"));
    z.createDStream = reinterpret_cast<decltype(z.createDStream)>(sym("ZSTD_createDStream"));
    z.freeDStream = reinterpret_cast<decltype(z.freeDStream)>(sym("ZSTD_freeDStream"));
    z.decompressStream = reinterpret_cast<decltype(z.decompressStream)>(sym("ZSTD_decompressStream"));
  });
  if (!z.handle || !z.decompress || !z.isError || !z.getFrameContentSize)
    fail(SBV2_ERR_INTERNAL, "libzstd.so.1 could not be loaded (needed for .sbv2 containers)");
  return z;
}

struct ZBuf {
  void* ptr;
  size_t size, pos;
};

}  // namespace

std::vector<uint8_t> zstd_decompress(const uint8_t* p, size_t n) {
  Zstd& z = zstd();
  if (!p || n < 4) fail(SBV2_ERR_PARSE, ".sbv2: not a zstd frame (too short)");
  const unsigned long long kUnknown = 0ULL - 1, kError = 0ULL - 2;
  unsigned long long sz = z.getFrameContentSize(p, n);
  if (sz == kError) fail(SBV2_ERR_PARSE, ".sbv2: not a zstd frame");
  if (sz != kUnknown) {
    std::vector<uint8_t> out(sz);
    size_t r = z.decompress(out.data(), out.size(), p, n);
    if (z.isError(r)) fail(SBV2_ERR_PARSE, std::string(".sbv2: zstd: ") + z.getErrorName(r));
    out.resize(r);
    return out;
  }
  // streaming path (ZstdCompressor(threads=-1) may omit the content size)
  if (!z.createDStream || !z.decompressStream) fail(SBV2_ERR_INTERNAL, "libzstd streaming API missing");
  void* ds = z.createDStream();
  std::vector<uint8_t> out;
  std::vector<uint8_t> chunk(1 << 20);
  ZBuf in{const_cast<uint8_t*>(p), n, 0};
  size_t last = 1;
  // Keep calling while there is input left OR the last call filled the whole output chunk (the decoder may still hold
  // data to flush even though every input byte has been consumed); stop when a frame ends exactly at the end of input.
  bool out_was_full = false;
  while (in.pos < in.size || out_was_full) {
    ZBuf o{chunk.data(), chunk.size(), 0};
    size_t r = z.decompressStream(ds, &o, &in);
    if (z.isError(r)) {
      z.freeDStream(ds);
      fail(SBV2_ERR_PARSE, std::string(".sbv2: zstd: ") + z.getErrorName(r));
    }
    out.insert(out.end(), chunk.begin(), chunk.begin() + o.pos);
    last = r;
    out_was_full = o.pos == o.size;
    if (r == 0 && in.pos >= in.size) break;
    if (o.pos == 0 && in.pos >= in.size) break;  // no progress and nothing left to feed: truncated
  }
  z.freeDStream(ds);
  if (last != 0) fail(SBV2_ERR_PARSE, ".sbv2: truncated zstd frame");
  return out;
}

std::vector<uint8_t> zstd_compress(const uint8_t* p, size_t n, int level) {
  Zstd& z = zstd();
  if (!z.compress || !z.compressBound) fail(SBV2_ERR_INTERNAL, "libzstd compress API missing");
  std::vector<uint8_t> out(z.compressBound(n));
  size_t r = z.compress(out.data(), out.size(), p, n, level);
  if (z.isError(r)) fail(SBV2_ERR_INTERNAL, std::string("zstd: ") + z.getErrorName(r));
  out.resize(r);
  return out;
}

std::vector<TarEntry> tar_entries(const uint8_t* p, size_t n) {
  std::vector<TarEntry> out;
  size_t off = 0;
  while (off + 512 <= n) {
    const uint8_t* h = p + off;
    bool all_zero = true;
    for (int i = 0; i < 512; ++i)
      if (h[i]) {
        all_zero = false;
        break;
      }
    if (all_zero) break;
    std::string name(reinterpret_cast<const char*>(h), strnlen(reinterpret_cast<const char*>(h), 100));
    // ustar prefix
    if (memcmp(h + 257, "ustar", 5) == 0 && h[345]) {
      std::string prefix(reinterpret_cast<const char*>(h + 345), strnlen(reinterpret_cast<const char*>(h + 345), 155));
      name = prefix + "/" + name;
    }
    uint64_t size = 0;
    if (h[124] & 0x80) {  // GNU base-256
      for (int i = 125; i < 136; ++i) size = (size << 8) | h[i];
    } else {
      for (int i = 124; i < 136; ++i) {
        uint8_t c = h[i];
        if (c == 0 || c == ' ') {
          if (size == 0 && c == ' ') continue;
          break;
        }
        if (c < '0' || c > '7') fail(SBV2_ERR_PARSE, ".sbv2: bad tar size field");
        size = size * 8 + (c - '0');
      }
    }
    char type = char(h[156]);
    off += 512;
    if (size > n - off) fail(SBV2_ERR_PARSE, ".sbv2: tar entry '" + name + "' exceeds archive");  // off <= n here; no wrap-around
    if (type == '0' || type == 0) {
      if (name.rfind("./", 0) == 0) name = name.substr(2);
      out.push_back({name, p + off, size_t(size)});
    }
    const uint64_t padded = (size + 511) / 512 * 512;  // size <= n - off, so this cannot wrap
    off = padded > n - off ? n : off + size_t(padded);
  }
  return out;
}

Sbv2File parse_sbv2file(const uint8_t* p, size_t n) {
  Sbv2File f;
  f.tar = zstd_decompress(p, n);
  for (auto& e : tar_entries(f.tar.data(), f.tar.size())) {
    if (e.name == "model.onnx") {
      f.onnx = e.data;
      f.onnx_n = e.size;
    } else if (e.name == "style_vectors.json") {
      f.style_json = e.data;
      f.style_n = e.size;
    }
  }
  // same order of checks as the reference (sbv2file.rs:30-35)
  if (!f.style_json) fail(SBV2_ERR_MODEL_NOT_FOUND, "model not found error: style_vectors");
  if (!f.onnx) fail(SBV2_ERR_MODEL_NOT_FOUND, "model not found error: vits2");
  return f;
}

// ---- style_vectors.json -------------------------------------------------------------------

namespace {
struct Json {
  const char* p;
  const char* end;
  void ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
  }
  bool eat(char c) {
    ws();
    if (p < end && *p == c) {
      ++p;
      return true;
    }
    return false;
  }
  void expect(char c) {
    if (!eat(c)) fail(SBV2_ERR_PARSE, std::string("style json: expected '") + c + "'");
  }
  std::string string() {
    expect('"');
    std::string s;
    while (p < end && *p != '"') {
      if (*p == '\\' && p + 1 < end) ++p;
      s.push_back(*p++);
    }
    if (p >= end) fail(SBV2_ERR_PARSE, "style json: unterminated string");
    ++p;
    return s;
  }
  double number() {
    ws();
    // JSON number grammar only ([-]digits[.digits][(e|E)[+-]digits]); the token is copied into a bounded NUL-terminated
    // buffer because the input is caller memory of `end - p` bytes, not a C string (strtod would also take nan/inf/hex).
    const char* q = p;
    if (q < end && *q == '-') ++q;
    const char* d0 = q;
    while (q < end && *q >= '0' && *q <= '9') ++q;
    if (q == d0) fail(SBV2_ERR_PARSE, "style json: expected number");
    if (q < end && *q == '.') {
      ++q;
      const char* f0 = q;
      while (q < end && *q >= '0' && *q <= '9') ++q;
      if (q == f0) fail(SBV2_ERR_PARSE, "style json: malformed number");
    }
    if (q < end && (*q == 'e' || *q == 'E')) {
      ++q;
      if (q < end && (*q == '+' || *q == '-')) ++q;
      const char* e0 = q;
      while (q < end && *q >= '0' && *q <= '9') ++q;
      if (q == e0) fail(SBV2_ERR_PARSE, "style json: malformed number");
    }
    char buf[64];
    const size_t len = size_t(q - p);
    if (len >= sizeof(buf)) fail(SBV2_ERR_PARSE, "style json: number token too long");
    memcpy(buf, p, len);
    buf[len] = 0;
    p = q;
    return strtod(buf, nullptr);
  }
  void skip_value() {
    ws();
    if (p >= end) fail(SBV2_ERR_PARSE, "style json: truncated");
    if (*p == '"') {
      string();
    } else if (*p == '[' || *p == '{') {
      char open = *p, close = (open == '[') ? ']' : '}';
      int depth = 0;
      bool in_str = false;
      for (; p < end; ++p) {
        if (in_str) {
          if (*p == '\\') ++p;
          else if (*p == '"') in_str = false;
        } else if (*p == '"') in_str = true;
        else if (*p == open) ++depth;
        else if (*p == close && --depth == 0) {
          ++p;
          return;
        }
      }
      fail(SBV2_ERR_PARSE, "style json: unbalanced brackets");
    } else {
      while (p < end && *p != ',' && *p != '}' && *p != ']') ++p;
    }
  }
};
}  // namespace

StyleVectors load_style_json(const uint8_t* bytes, size_t n) {
  Json j{reinterpret_cast<const char*>(bytes), reinterpret_cast<const char*>(bytes) + n};
  StyleVectors s;
  std::vector<int64_t> shape;
  std::vector<float> data;
  bool has_shape = false, has_data = false;
  j.expect('{');
  if (!j.eat('}')) {
    do {
      std::string key = j.string();
      j.expect(':');
      if (key == "shape") {
        j.expect('[');
        if (!j.eat(']')) {
          do shape.push_back(int64_t(j.number()));
          while (j.eat(','));
          j.expect(']');
        }
        has_shape = true;
      } else if (key == "data") {
        j.expect('[');
        if (!j.eat(']')) {
          do {
            j.expect('[');
            if (!j.eat(']')) {
              do data.push_back(float(j.number()));
              while (j.eat(','));
              j.expect(']');
            }
          } while (j.eat(','));
          j.expect(']');
        }
        has_data = true;
      } else {
        j.skip_value();
      }
    } while (j.eat(','));
    j.expect('}');
  }
  if (!has_shape || !has_data) fail(SBV2_ERR_PARSE, "style json: missing field 'shape' or 'data'");
  if (shape.size() != 2) fail(SBV2_ERR_PARSE, "style json: shape must have 2 entries");
  const bool shape_ok = shape[0] >= 0 && shape[1] >= 0 &&
                        (shape[1] == 0 ? data.empty() : (data.size() % size_t(shape[1]) == 0 && data.size() / size_t(shape[1]) == size_t(shape[0])));
  if (!shape_ok)
    fail(SBV2_ERR_INVALID_ARGUMENT, "NDArray error: style data does not match shape");  // ShapeError in the reference
  s.rows = shape[0];
  s.cols = shape[1];
  s.data = std::move(data);
  return s;
}

std::vector<float> get_style_vector(const StyleVectors& s, int32_t style_id, float weight) {
  if (s.rows < 1) fail(SBV2_ERR_INVALID_ARGUMENT, "style matrix is empty");
  if (style_id < 0 || style_id >= s.rows) fail(SBV2_ERR_INVALID_ARGUMENT, "style_id out of range");
  std::vector<float> out(size_t(s.cols));
  const float* mean = s.data.data();
  const float* v = s.data.data() + size_t(style_id) * s.cols;
  for (int64_t i = 0; i < s.cols; ++i) {
    float diff = (v[i] - mean[i]) * weight;  // same operation order as style.rs:24-27
    out[i] = mean[i] + diff;
  }
  return out;
}

std::vector<uint8_t> base64_decode(const char* p, size_t n) {
  static int8_t table[256];
  static std::once_flag once;
  std::call_once(once, [] {
    memset(table, -1, sizeof(table));
    const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for (int i = 0; i < 64; ++i) table[uint8_t(a[i])] = int8_t(i);
  });
  std::vector<uint8_t> out;
  out.reserve(n * 3 / 4);
  uint32_t acc = 0;
  int bits = 0;
  for (size_t i = 0; i < n; ++i) {
    uint8_t c = uint8_t(p[i]);
    if (c == '=') break;
    if (c == '\n' || c == '\r') continue;
    int8_t v = table[c];
    if (v < 0) fail(SBV2_ERR_PARSE, "base64 error");
    acc = (acc << 6) | uint32_t(v);
    bits += 6;
    if (bits >= 8) {
      bits -= 8;
      out.push_back(uint8_t(acc >> bits));
    }
  }
  return out;
}

StyleVectors load_style_npy_base64(const char* b64, size_t n) {
  std::vector<uint8_t> npy = base64_decode(b64, n);
  if (npy.size() < 10 || memcmp(npy.data(), "\x93NUMPY", 6) != 0) fail(SBV2_ERR_PARSE, "aivmx: style vectors are not a .npy file");
  int major = npy[6];
  size_t hlen, hoff;
  if (major == 1) {
    hlen = npy[8] | (npy[9] << 8);
    hoff = 10;
  } else {
    if (npy.size() < 12) fail(SBV2_ERR_PARSE, "aivmx: truncated .npy");
    hlen = npy[8] | (npy[9] << 8) | (npy[10] << 16) | (size_t(npy[11]) << 24);
    hoff = 12;
  }
  if (hoff + hlen > npy.size()) fail(SBV2_ERR_PARSE, "aivmx: truncated .npy header");
  std::string hdr(reinterpret_cast<const char*>(npy.data() + hoff), hlen);
  auto find_after = [&](const std::string& key) -> size_t {
    size_t k = hdr.find(key);
    if (k == std::string::npos) fail(SBV2_ERR_PARSE, "aivmx: .npy header lacks " + key);
    size_t c = hdr.find(':', k);
    if (c == std::string::npos) fail(SBV2_ERR_PARSE, "aivmx: malformed .npy header near " + key);
    return c + 1;
  };
  const size_t npos = std::string::npos;
  size_t d = find_after("'descr'");
  size_t q1 = hdr.find('\'', d), q2 = q1 == npos ? npos : hdr.find('\'', q1 + 1);
  if (q1 == npos || q2 == npos) fail(SBV2_ERR_PARSE, "aivmx: malformed .npy header (descr)");
  std::string descr = hdr.substr(q1 + 1, q2 - q1 - 1);
  if (descr != "<f4" && descr != "|f4" && descr != "=f4") fail(SBV2_ERR_UNSUPPORTED, "aivmx: style vectors must be float32, got " + descr);
  size_t f = find_after("'fortran_order'");
  const size_t fv = hdr.find_first_not_of(' ', f);
  if (fv == npos) fail(SBV2_ERR_PARSE, "aivmx: malformed .npy header (fortran_order)");
  bool fortran = hdr.compare(fv, 4, "True") == 0;
  size_t s = find_after("'shape'");
  size_t p1 = hdr.find('(', s), p2 = p1 == npos ? npos : hdr.find(')', p1);
  if (p1 == npos || p2 == npos) fail(SBV2_ERR_PARSE, "aivmx: malformed .npy header (shape)");
  std::vector<int64_t> shape;
  {
    std::string inner = hdr.substr(p1 + 1, p2 - p1 - 1);
    const char* c = inner.c_str();
    while (*c) {
      while (*c == ' ' || *c == ',') ++c;
      if (!*c) break;
      char* e;
      shape.push_back(strtoll(c, &e, 10));
      if (e == c) break;
      c = e;
    }
  }
  if (shape.size() != 2) fail(SBV2_ERR_INVALID_ARGUMENT, "aivmx: expected 2D array");
  StyleVectors out;
  out.rows = shape[0];
  out.cols = shape[1];
  if (out.rows < 0 || out.cols < 0) fail(SBV2_ERR_PARSE, "aivmx: negative .npy dimension");
  // checked arithmetic: the payload bound decides, so neither rows*cols nor count*4 may wrap
  const size_t avail = (npy.size() - hoff - hlen) / 4;
  if (out.cols != 0 && size_t(out.rows) > avail / size_t(out.cols)) fail(SBV2_ERR_PARSE, "aivmx: .npy payload shorter than shape");
  size_t count = size_t(out.rows) * size_t(out.cols);
  const uint8_t* payload = npy.data() + hoff + hlen;
  out.data.resize(count);
  if (!fortran) {
    memcpy(out.data.data(), payload, count * 4);
  } else {
    for (int64_t r = 0; r < out.rows; ++r)
      for (int64_t c = 0; c < out.cols; ++c) memcpy(&out.data[size_t(r * out.cols + c)], payload + 4 * size_t(c * out.rows + r), 4);
  }
  return out;
}

// ---- WAV ------------------------------------------------------------------------------------

std::vector<uint8_t> wav_from_f32(const float* samples, int64_t n) {
  // Layout hound 3.5 produces for WavSpec{channels:1, sample_rate:44100, bits_per_sample:32,
  // sample_format:Float}: RIFF / fmt (WAVE_FORMAT_EXTENSIBLE, 40 bytes) / data.
  if (n < 0) fail(SBV2_ERR_INVALID_ARGUMENT, "negative sample count");
  if (n > (int64_t(0xFFFFFFFF) - 60) / 4) fail(SBV2_ERR_INVALID_ARGUMENT, "too many samples for a RIFF container (4 GiB)");
  const uint32_t data_bytes = uint32_t(n * 4);
  std::vector<uint8_t> w(68 + size_t(data_bytes));
  uint8_t* p = w.data();
  auto u16 = [&](uint16_t v) { memcpy(p, &v, 2); p += 2; };
  auto u32 = [&](uint32_t v) { memcpy(p, &v, 4); p += 4; };
  auto tag = [&](const char* t) { memcpy(p, t, 4); p += 4; };
  tag("RIFF"); u32(60 + data_bytes); tag("WAVE");
  tag("fmt "); u32(40);
  u16(0xFFFE); u16(1); u32(44100); u32(44100 * 4); u16(4); u16(32);
  u16(22); u16(32); u32(0x1);
  const uint8_t guid_float[16] = {0x03, 0x00, 0x00, 0x00, 0x00, 0x00, 0x10, 0x00, 0x80, 0x00, 0x00, 0xaa, 0x00, 0x38, 0x9b, 0x71};
  memcpy(p, guid_float, 16); p += 16;
  tag("data"); u32(data_bytes);
  if (n) memcpy(p, samples, size_t(data_bytes));
  return w;
}

// SURVEY.md 8f row 3, optional: 16-bit PCM (plain 44-byte WAVE_FORMAT_PCM header), half the bytes of the
// float container.  Samples are clamped to [-1, 1] and scaled by 32767 with round-to-nearest-even.
std::vector<uint8_t> wav_pcm16_from_f32(const float* samples, int64_t n) {
  if (n < 0) fail(SBV2_ERR_INVALID_ARGUMENT, "negative sample count");
  if (n > (int64_t(0xFFFFFFFF) - 36) / 2) fail(SBV2_ERR_INVALID_ARGUMENT, "too many samples for a RIFF container (4 GiB)");
  const uint32_t data_bytes = uint32_t(n * 2);
  std::vector<uint8_t> w(44 + size_t(data_bytes));
  uint8_t* p = w.data();
  auto u16 = [&](uint16_t v) { memcpy(p, &v, 2); p += 2; };
  auto u32 = [&](uint32_t v) { memcpy(p, &v, 4); p += 4; };
  auto tag = [&](const char* t) { memcpy(p, t, 4); p += 4; };
  tag("RIFF"); u32(36 + data_bytes); tag("WAVE");
  tag("fmt "); u32(16);
  u16(1); u16(1); u32(44100); u32(44100 * 2); u16(2); u16(16);
  tag("data"); u32(data_bytes);
  int16_t* o = reinterpret_cast<int16_t*>(p);
  for (int64_t i = 0; i < n; ++i) {
    float x = samples[i];
    x = x != x ? 0.f : (x < -1.f ? -1.f : (x > 1.f ? 1.f : x));  // NaN -> silence
    o[i] = int16_t(std::nearbyint(x * 32767.f));
  }
  return w;
}

}  // namespace sbv2
