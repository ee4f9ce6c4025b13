// sbv2_model base: device binding, weight uploads, out-pointer registry.
#include "model.h"

#include <nvtx3/nvToolsExt.h>

#include <mutex>
#include <unordered_map>

namespace sbv2 {
namespace {
std::mutex g_out_mu;
std::unordered_map<void*, bool> g_out;  // ptr -> pinned?
thread_local std::string g_last_error;
}  // namespace

void set_last_error(const std::string& m) { g_last_error = m; }
const char* last_error_cstr() { return g_last_error.c_str(); }

// Pinned result blocks are recycled: cudaMallocHost of a 45 MB waveform buffer costs ~10 ms, which
// would otherwise sit inside every synthesize call.
namespace {
struct PinnedBlock {
  void* p;
  size_t cap;
};
std::vector<PinnedBlock> g_pinned_free;
std::unordered_map<void*, size_t> g_pinned_cap;
size_t g_pinned_cached = 0;
constexpr size_t kPinnedCacheLimit = size_t(1) << 30;
}  // namespace

void* alloc_out(size_t bytes, bool pinned) {
  void* p = nullptr;
  if (bytes == 0) bytes = 1;
  if (pinned) {
    const size_t want = (bytes + ((size_t(1) << 20) - 1)) & ~((size_t(1) << 20) - 1);
    {
      std::lock_guard<std::mutex> lk(g_out_mu);
      size_t best = g_pinned_free.size();
      for (size_t i = 0; i < g_pinned_free.size(); ++i)
        if (g_pinned_free[i].cap >= want && g_pinned_free[i].cap <= 2 * want &&
            (best == g_pinned_free.size() || g_pinned_free[i].cap < g_pinned_free[best].cap))
          best = i;
      if (best != g_pinned_free.size()) {
        p = g_pinned_free[best].p;
        g_pinned_cached -= g_pinned_free[best].cap;
        g_pinned_free.erase(g_pinned_free.begin() + long(best));
        g_out[p] = true;
        return p;
      }
    }
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      pinned = false;
    } else {
      std::lock_guard<std::mutex> lk(g_out_mu);
      g_pinned_cap[p] = want;
      g_out[p] = true;
      return p;
    }
  }
  p = malloc(bytes);
  if (!p) throw std::bad_alloc();
  std::lock_guard<std::mutex> lk(g_out_mu);
  g_out[p] = false;
  return p;
}

void free_out(void* p) {
  if (!p) return;
  bool pinned = false, found = false, release = false;
  {
    std::lock_guard<std::mutex> lk(g_out_mu);
    auto it = g_out.find(p);
    if (it != g_out.end()) {
      pinned = it->second;
      found = true;
      g_out.erase(it);
      if (pinned) {
        const size_t cap = g_pinned_cap[p];
        if (g_pinned_cached + cap <= kPinnedCacheLimit) {
          g_pinned_free.push_back({p, cap});
          g_pinned_cached += cap;
        } else {
          g_pinned_cap.erase(p);
          release = true;
        }
      }
    }
  }
  if (!found) return;  // not ours: ignore rather than corrupt the heap
  if (pinned) {
    if (release) cudaFreeHost(p);
  } else {
    free(p);
  }
}

}  // namespace sbv2

sbv2_model::~sbv2_model() {
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  for (auto& kv : regions) {
    cudaEventDestroy(kv.second.first);
    cudaEventDestroy(kv.second.second);
  }
  if (wait_event) cudaEventDestroy(wait_event);
  for (void* p : owned_device) cudaFree(p);
  if (stream) cudaStreamDestroy(stream);
}

void sbv2_model::bind_device() const { CUDA_CHECK(cudaSetDevice(device)); }

void sbv2_model::wait_stream(bool yield) {
  if (!yield) {
    CUDA_CHECK(cudaStreamSynchronize(stream));
    return;
  }
  if (!wait_event) CUDA_CHECK(cudaEventCreateWithFlags(&wait_event, cudaEventBlockingSync | cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(wait_event, stream));
  CUDA_CHECK(cudaEventSynchronize(wait_event));
}

void* sbv2_model::upload_bytes(const void* host, size_t bytes) {
  void* d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, bytes ? bytes : 16));
  owned_device.push_back(d);
  if (bytes) CUDA_CHECK(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, stream));
  // `host` is usually a temporary vector: make the copy complete before returning
  CUDA_CHECK(cudaStreamSynchronize(stream));
  return d;
}

// Regions ("text", "flow", "decoder", "bert") are always NVTX ranges (header-only NVTX3: a no-op unless a tool is
// attached) and, with sbv2_model_enable_timing, CUDA-event pairs on the model's stream.
void sbv2_model::region_begin(const std::string& name) {
  nvtxRangePushA(name.c_str());
  if (!timing) return;
  auto it = regions.find(name);
  if (it == regions.end()) {
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a));
    CUDA_CHECK(cudaEventCreate(&b));
    it = regions.emplace(name, std::make_pair(a, b)).first;
  }
  CUDA_CHECK(cudaEventRecord(it->second.first, stream));
}

void sbv2_model::region_end(const std::string& name) {
  nvtxRangePop();
  if (!timing) return;
  auto it = regions.find(name);
  if (it == regions.end()) return;
  CUDA_CHECK(cudaEventRecord(it->second.second, stream));
}

float sbv2_model::region_ms(const std::string& name) {
  auto it = regions.find(name);
  if (it == regions.end()) sbv2::fail(SBV2_ERR_INVALID_ARGUMENT, "no timed region named " + name);
  CUDA_CHECK(cudaEventSynchronize(it->second.second));
  float ms = 0.f;
  CUDA_CHECK(cudaEventElapsedTime(&ms, it->second.first, it->second.second));
  return ms;
}
