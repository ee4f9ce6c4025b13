// Micro-benchmark (debug hook): issue rate of tcgen05.mma.cta_group::1.kind::f16 M=128 for
// different N and shared-memory layouts, no loads — isolates the tensor-pipe / smem-operand cost
// from the rest of the conv kernel.
#include <cuda_runtime.h>

#include "common.h"
#include "umma_device.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n, int layout, int shift_rows, int iters, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && lane == 0) {
    const uint32_t sA = smem_u32(smem), sB = sA + 96 * 1024;
    const uint32_t idesc = (1u << 4) | ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    uint64_t ad, bd;
    const int RA = 306;
    if (layout == 0) {  // no swizzle, plane-major: LBO = RA*16, SBO = 128
      const uint64_t hi = (uint64_t)(0x4000u | 8u) << 32;
      ad = hi | ((uint64_t)RA << 16) | ((sA >> 4) + shift_rows);
      bd = hi | ((uint64_t)n << 16) | (sB >> 4);
    } else {  // 128B swizzle, K-major: SBO = 1024, layout type 2
      const uint64_t hi = ((uint64_t)(0x4000u | (1024u >> 4)) << 32) | ((uint64_t)2 << 61);
      ad = hi | ((sA >> 4) + shift_rows * 8);  // shift by rows*128B
      bd = hi | (sB >> 4);
    }
    // descriptors for 4 K-steps precomputed; the loop body is 8 back-to-back MMAs with no address math
    uint64_t a4[4], b4[4];
    for (int q = 0; q < 4; ++q) {
      a4[q] = ad + (layout == 0 ? (uint64_t)(q * 2 * RA) : (uint64_t)(q * 2));
      b4[q] = bd + (layout == 0 ? (uint64_t)(q * 2 * n) : (uint64_t)(q * 2));
    }
    const uint32_t acc0 = tmem, acc1 = tmem + (nacc > 1 ? n : 0);
#define MMA(ACC, A, B, EN)                                                                                          \
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(ACC), \
               "l"(A), "l"(B), "r"(idesc), "r"(EN)                                                                   \
               : "memory")
    long long t0 = clock64();
    MMA(acc0, a4[0], b4[0], 0u);
    MMA(acc1, a4[0], b4[0], 0u);
    for (int i = 0; i < iters; i += 8) {
      MMA(acc0, a4[0], b4[0], 1u);
      MMA(acc0, a4[1], b4[1], 1u);
      MMA(acc0, a4[2], b4[2], 1u);
      MMA(acc0, a4[3], b4[3], 1u);
      MMA(acc1, a4[0], b4[0], 1u);
      MMA(acc1, a4[1], b4[1], 1u);
      MMA(acc1, a4[2], b4[2], 1u);
      MMA(acc1, a4[3], b4[3], 1u);
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
            smem_u32(&bar))
        : "memory");
    long long t2 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// Variant with the other warps of the CTA generating the traffic of an epilogue while one thread issues MMAs:
// noise bit 0: tcgen05.ld of other TMEM columns, bit 1: 16-byte st.shared, bit 2: 16-byte global loads.
__global__ void __launch_bounds__(544, 1) mma_rate2_kernel(int n, int ra, int iters, int nacc, int noise, const uint4* gsrc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) {
    uint32_t v = 0;
    if (noise & 16) {  // pseudo-random fp16 pairs in [-2, 2): exercises the multipliers (power) unlike all-zero operands
      uint32_t h = (uint32_t)i * 2654435761u + 12345u;
      h ^= h >> 13;
      h *= 1274126177u;
      v = (h & 0x83FF83FFu) | 0x3C003C00u;
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) {
    done = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && (noise & 32)) {
    // the conv kernels' issue shape: whole warp in the loop, elect.sync + unrolled dispatch + __syncwarp per tap
    const uint32_t sA = smem_u32(smem), sB = sA + 150 * 1024;
    const uint32_t idesc = (1u << 4) | ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    const uint64_t hi = (uint64_t)(0x4000u | 8u) << 32;
    const uint64_t ad = hi | ((uint64_t)ra << 16) | ((sA >> 4) + 3);
    const uint64_t bd = hi | ((uint64_t)n << 16) | (sB >> 4);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 88) {
      for (int tap = 0; tap < 11; ++tap) {
        const uint64_t at = ad + (uint64_t)tap, bt = bd + (uint64_t)(tap * 2 * n);
        if (sbv2::elect_one_sync()) sbv2::issue_mmas_dyn(nacc, 1, tmem, at, bt, 2u * ra, 2u * n, (uint32_t)n, idesc, tap > 0 ? 1u : 0u);
        __syncwarp();
      }
    }
    long long t1 = clock64();
    if (lane == 0) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      asm volatile(
          "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
              smem_u32(&bar))
          : "memory");
      long long t2 = clock64();
      done = 1;
      if (blockIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
      }
    }
  } else if (warp == 0) {
    if (lane == 0) {
      const uint32_t sA = smem_u32(smem), sB = sA + 150 * 1024;
      const uint32_t idesc = (1u << 4) | ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
      const uint64_t hi = (uint64_t)(0x4000u | 8u) << 32;
      const uint64_t ad = hi | ((uint64_t)ra << 16) | ((sA >> 4) + 3);
      const uint64_t bd = hi | ((uint64_t)n << 16) | (sB >> 4);
      long long t0 = clock64();
      if (noise & 8) {
        // conv-like pattern: 11 taps x 8 row tiles, the tap shifts the A start by one row and selects another B block
        for (int i = 0; i < iters; i += 88) {
          for (int tap = 0; tap < 11; ++tap) {
            const uint64_t at = ad + (uint64_t)tap, bt = bd + (uint64_t)(tap * 2 * n);
#pragma unroll
            for (int q = 0; q < 8; ++q) MMA(tmem + (uint32_t)(q * n), at + (uint64_t)(q * 128), bt, tap > 0 ? 1u : 0u);
          }
        }
      } else {
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t acc = tmem + (uint32_t)((q % nacc) * n);
            MMA(acc, ad + (uint64_t)((q & 3) * 128), bd, i > 0 ? 1u : 0u);
          }
        }
      }
      long long t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      asm volatile(
          "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
              smem_u32(&bar))
          : "memory");
      long long t2 = clock64();
      done = 1;
      if (blockIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
      }
    }
  } else if (noise) {
    uint32_t acc = 0;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (uint32_t)((warp >> 2) * 16);
    uint4* sdst = reinterpret_cast<uint4*>(smem + 160 * 1024) + threadIdx.x;
    const uint4* g = gsrc + (size_t)blockIdx.x * 65536 + threadIdx.x;
    int it = 0;
    while (!done) {
      if (noise & 1) {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += v[0] + v[15];
      }
      if (noise & 4) {
        const uint4 q = g[(it & 63) * 1024];
        acc += q.x;
      }
      if (noise & 2) *sdst = make_uint4(acc, acc, acc, acc);
      ++it;
    }
    if (acc == 0x12345678u) out[1] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace

extern "C" int sbv2_debug_mma_rate2(int n, int ra, int iters, int nacc, int blocks, int noise, long long* out2) {
  return sbv2::guarded([&] {
    CUDA_CHECK(cudaSetDevice(0));
    CUDA_CHECK(cudaFuncSetAttribute(mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long* d = nullptr;
    uint4* g = nullptr;
    CUDA_CHECK(cudaMalloc(&d, 16));
    CUDA_CHECK(cudaMalloc(&g, size_t(blocks + 1) * 65536 * 16));
    CUDA_CHECK(cudaMemset(g, 0, size_t(blocks + 1) * 65536 * 16));
    mma_rate2_kernel<<<blocks, 544, 200 * 1024>>>(n, ra, iters, nacc, noise, g, d);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(out2, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
    cudaFree(g);
  });
}

extern "C" int sbv2_debug_mma_rate(int n, int layout, int shift_rows, int iters, int nacc, int blocks, long long* out2) {
  return sbv2::guarded([&] {
    CUDA_CHECK(cudaSetDevice(0));
    CUDA_CHECK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long* d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, 16));
    mma_rate_kernel<<<blocks, 128, 200 * 1024>>>(n, layout, shift_rows, iters, nacc, d);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(out2, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
  });
}

// ---- cluster multicast probe: every CTA of a cluster of NC fetches 1/NC of a buffer and multicasts it to all CTAs; each
// CTA then dumps what landed in its shared memory.  Also reports %cluster_ctarank and the shared-window addresses.
namespace {
__global__ void __launch_bounds__(128) multicast_probe_kernel(const uint32_t* src, uint32_t* out, int words, int nc, uint32_t* info) {
  extern __shared__ __align__(128) uint8_t psm[];
  const uint32_t sbuf = sbv2::smem_u32(psm);
  const uint32_t bar = sbuf + (uint32_t)words * 4;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    sbv2::mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < words; i += blockDim.x) reinterpret_cast<uint32_t*>(psm)[i] = 0xdeadbeefu;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)words * 4, slice = bytes / (uint32_t)nc;
    sbv2::mbar_expect_tx(bar, bytes);
    sbv2::bulk_g2s_multicast(sbuf + rank * slice, reinterpret_cast<const uint8_t*>(src) + rank * slice, slice, bar,
                             (uint16_t)((1u << nc) - 1u));
    info[blockIdx.x * 4 + 0] = rank;
    info[blockIdx.x * 4 + 1] = sbuf;
    info[blockIdx.x * 4 + 2] = bar;
  }
  sbv2::mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < words; i += blockDim.x) out[(size_t)blockIdx.x * words + i] = reinterpret_cast<volatile uint32_t*>(psm)[i];
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
}  // namespace

extern "C" int sbv2_debug_multicast_probe(int nc, int words, int blocks, uint32_t* out /*[blocks*words]*/, uint32_t* info /*[blocks*4]*/) {
  return sbv2::guarded([&] {
    CUDA_CHECK(cudaSetDevice(0));
    uint32_t *d_src = nullptr, *d_out = nullptr, *d_info = nullptr;
    CUDA_CHECK(cudaMalloc(&d_src, size_t(words) * 4));
    CUDA_CHECK(cudaMalloc(&d_out, size_t(blocks) * words * 4));
    CUDA_CHECK(cudaMalloc(&d_info, size_t(blocks) * 16));
    std::vector<uint32_t> h;
    h.resize(size_t(words));
    for (int i = 0; i < words; ++i) h[size_t(i)] = uint32_t(i);
    CUDA_CHECK(cudaMemcpy(d_src, h.data(), size_t(words) * 4, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemset(d_out, 0xff, size_t(blocks) * words * 4));
    sbv2::launch_pdl_cluster(false, nc, multicast_probe_kernel, dim3(blocks), dim3(128), size_t(words) * 4 + 64, nullptr, d_src, d_out, words, nc,
                             d_info);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(out, d_out, size_t(blocks) * words * 4, cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(info, d_info, size_t(blocks) * 16, cudaMemcpyDeviceToHost));
    cudaFree(d_src);
    cudaFree(d_out);
    cudaFree(d_info);
  });
}
