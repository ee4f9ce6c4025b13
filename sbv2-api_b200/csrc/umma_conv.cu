// tcgen05 / TMEM implicit-GEMM convolution for the HiFi-GAN decoder (sm_100a only).
//
// Data layout ("planar fp16"): a [rows, C] activation is stored as C/8 planes, plane p holding
// channels [8p, 8p+8) of every row as one 16-byte item: half buf[C/8][rows][8].  Utterances of a
// batch are packed along rows with GAP zero rows around each one, so "same" zero padding is
// obtained by reading neighbouring rows and never written rows stay zero.
//
// Why this layout: tcgen05.mma reads K-major operands as 8-row x 16-byte core matrices.  With
// SWIZZLE_NONE the descriptor (start, LBO = byte distance between the two 16-byte K chunks,
// SBO = byte distance between 8-row groups) addresses exactly plane-major shared memory
// [plane][row][16 B]: LBO = rows*16, SBO = 128.  A filter tap shifted by s rows is the same tile
// with start += 16*s — any s, no swizzle phase to respect — so one halo'd activation tile, loaded
// once with plain 1-D bulk copies (cp.async.bulk, no tensor map), feeds every tap of the filter.
//
// Per CTA: rows [t0, t0 + 128*MT) of one utterance x NB output channels.
//   warp 0  : producer — bulk copies of the activation tile (per 64-channel chunk) and of the
//             packed weights (ring of stages), completion on mbarriers
//   warp 1  : TMEM alloc + single-thread tcgen05.mma issue, MT accumulators of 128 x NB fp32
//   warps 2-5: epilogue — tcgen05.ld, bias, residual, MRF accumulation, activation, fp16 store
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>

#include "kernels.h"
#include "model.h"
#include "umma_conv.h"

namespace sbv2 {
namespace {

constexpr int GAP = 32;        // zero rows around each utterance (>= largest halo: 5*(11-1)/2 = 25)
constexpr int TAIL_ROWS = 640; // slack rows after the last utterance (a tile may over-read)
constexpr int MAX_TAPS = 16;
constexpr int MAX_KCHUNKS = 8;
constexpr int MAX_STAGES = 4;
constexpr int SMEM_LIMIT = 227 * 1024;

enum ResMode { RES_NONE = 0, RES_LRELU_INV = 1 };
enum UAccum { UACC_NONE = 0, UACC_SET = 1, UACC_ADD = 2, UACC_FINAL = 3 };

struct UmmaConvArgs {
  const __half* in;
  long long in_plane_stride;   // elements between planes of `in`
  __half* out;
  long long out_plane_stride;
  const __half* residual;      // geometry of `out`; stored post-lrelu(0.1) when res_mode == RES_LRELU_INV
  float* accum;                // fp32 planar, geometry of `out`
  const __half* w;             // packed [nblk][step][KC/8][NB][8]
  const float* bias;           // [n_nblk*NB]
  const float* bias_utt;       // [n_utt][n_nblk*NB] or null
  const int* tile_prefix;      // [n_utt+1]
  const int* pstart_in;        // [n_utt] first planar row of each utterance in `in`
  const int* pstart_out;
  const int* len;              // [n_utt] rows (input resolution)
  int n_utt;
  int cin, nb, taps, kc, nkc, mt, sps, nstages, nloads, total_steps;
  int tap_shift[MAX_TAPS];
  int halo_lo, halo_hi;
  int out_mul, out_off;
  int act_out;                 // ACT_NONE / ACT_LRELU / ACT_LRELU01
  int res_mode, accum_mode;
  float accum_div;
  int tmem_cols;
  unsigned idesc;
  int dbg_swap;                // debugging: swap LBO/SBO roles
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// SWIZZLE_NONE, K-major shared-memory matrix descriptor (sm_100 "version 1")
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ACT_LRELU) return v > 0.f ? v : v * 0.1f;
  if (act == ACT_LRELU01) return v > 0.f ? v : v * 0.01f;
  return v;
}

// ---- the kernel ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(192, 1) umma_conv_kernel(const UmmaConvArgs p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int TM = 128 * p.mt;
  // tile -> (utterance, first row)
  int b;
  {
    int lo = 0, hi = p.n_utt;  // largest b with tile_prefix[b] <= tile
    const int tile = blockIdx.x;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (p.tile_prefix[mid] <= tile) lo = mid;
      else hi = mid;
    }
    b = lo;
  }
  const int t0 = (blockIdx.x - p.tile_prefix[b]) * TM;
  const int len = p.len[b];
  const int nblk = blockIdx.y;
  const int RA = TM + p.halo_lo + p.halo_hi;
  const int planes_per_chunk = p.kc / 8;
  const uint32_t a_bytes = (uint32_t)(p.cin / 8) * RA * 16;
  const uint32_t step_bytes = (uint32_t)p.nb * p.kc * 2;
  const uint32_t stage_bytes = step_bytes * p.sps;
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + ((a_bytes + 127u) & ~127u);
  const uint32_t sBar = sB + stage_bytes * p.nstages;
  // barriers: a_full[nkc], b_full[nstages], b_empty[nstages], acc_full ; then tmem slot
  const uint32_t bar_a = sBar, bar_bf = bar_a + 8 * MAX_KCHUNKS, bar_be = bar_bf + 8 * MAX_STAGES, bar_acc = bar_be + 8 * MAX_STAGES;
  const uint32_t tmem_slot = bar_acc + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - sA));

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nkc; ++i) mbar_init(bar_a + 8 * i, 1);
    for (int i = 0; i < p.nstages; ++i) {
      mbar_init(bar_bf + 8 * i, 1);
      mbar_init(bar_be + 8 * i, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- producer ----------------
      const __half* wbase = p.w + (size_t)nblk * p.total_steps * (step_bytes / 2);
      const long long in_row0 = (long long)p.pstart_in[b] + t0 - p.halo_lo;
      auto load_a_chunk = [&](int kc) {
        mbar_expect_tx(bar_a + 8 * kc, (uint32_t)planes_per_chunk * RA * 16);
        for (int q = 0; q < planes_per_chunk; ++q) {
          int plane = kc * planes_per_chunk + q;
          bulk_g2s(sA + (uint32_t)plane * RA * 16, p.in + (size_t)plane * p.in_plane_stride + in_row0 * 8, (uint32_t)RA * 16,
                   bar_a + 8 * kc);
        }
      };
      auto load_b = [&](int i) {  // i-th stage load
        int st = i % p.nstages;
        uint32_t par = ((i / p.nstages) & 1) ^ 1;
        mbar_wait(bar_be + 8 * st, par);
        int first = i * p.sps;
        int nsteps = min(p.sps, p.total_steps - first);
        uint32_t bytes = step_bytes * nsteps;
        mbar_expect_tx(bar_bf + 8 * st, bytes);
        bulk_g2s(sB + stage_bytes * st, wbase + (size_t)first * (step_bytes / 2), bytes, bar_bf + 8 * st);
      };
      load_a_chunk(0);
      load_b(0);
      for (int kc = 1; kc < p.nkc; ++kc) load_a_chunk(kc);
      for (int i = 1; i < p.nloads; ++i) load_b(i);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      const uint32_t a_lbo = p.dbg_swap ? 128u : (uint32_t)RA * 16, a_sbo = p.dbg_swap ? (uint32_t)RA * 16 : 128u;
      const uint32_t b_lbo = p.dbg_swap ? 128u : (uint32_t)p.nb * 16, b_sbo = p.dbg_swap ? (uint32_t)p.nb * 16 : 128u;
      const int k16_per_chunk = p.kc / 16;
      int step = 0;
      for (int tap = 0; tap < p.taps; ++tap) {
        const int row_off = p.halo_lo + p.tap_shift[tap];
        for (int kc = 0; kc < p.nkc; ++kc, ++step) {
          const int load = step / p.sps, si = step - load * p.sps;
          const int st = load % p.nstages;
          if (si == 0) mbar_wait(bar_bf + 8 * st, (load / p.nstages) & 1);
          if (tap == 0) mbar_wait(bar_a + 8 * kc, 0);
          tc_fence_after();
          const uint32_t b_stage = sB + stage_bytes * st + step_bytes * si;
          for (int a = 0; a < p.mt; ++a) {
            for (int k = 0; k < k16_per_chunk; ++k) {
              const int plane = kc * planes_per_chunk + 2 * k;
              uint64_t ad = make_desc(sA + (uint32_t)plane * RA * 16 + (uint32_t)(a * 128 + row_off) * 16, a_lbo, a_sbo);
              uint64_t bd = make_desc(b_stage + (uint32_t)(2 * k) * p.nb * 16, b_lbo, b_sbo);
              tc_mma_f16(tmem_base + (uint32_t)(a * p.nb), ad, bd, p.idesc, (step > 0 || k > 0) ? 1u : 0u);
            }
          }
          if (si == p.sps - 1 || step == p.total_steps - 1) tc_commit(bar_be + 8 * st);
        }
      }
      tc_commit(bar_acc);
    }
  } else {
    // ---------------- epilogue ----------------
    const int wq = warp & 3;
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const float* bias = p.bias + (size_t)nblk * p.nb;
    const float* bias_u = p.bias_utt ? p.bias_utt + (size_t)b * gridDim.y * p.nb + (size_t)nblk * p.nb : nullptr;
    for (int a = 0; a < p.mt; ++a) {
      const int t = t0 + a * 128 + wq * 32 + lane;
      const bool valid = t < len;
      const long long orow = (long long)p.pstart_out[b] + (long long)t * p.out_mul + p.out_off;
      for (int c0 = 0; c0 < p.nb; c0 += 16) {
        uint32_t v[16];
        tc_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(a * p.nb + c0), v);
        tc_wait_ld();
        if (!valid) continue;
#pragma unroll
        for (int hp = 0; hp < 2; ++hp) {
          const int co = c0 + 8 * hp;  // within this N block
          const long long plane = ((long long)nblk * p.nb + co) >> 3;
          const long long eoff = plane * p.out_plane_stride + orow * 8;
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            f[e] = __uint_as_float(v[8 * hp + e]) + bias[co + e];
            if (bias_u) f[e] += bias_u[co + e];
          }
          if (p.res_mode == RES_LRELU_INV) {
            uint4 r = *reinterpret_cast<const uint4*>(p.residual + eoff);
            const __half2* rh = reinterpret_cast<const __half2*>(&r);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 y = __half22float2(rh[e]);
              f[2 * e] += y.x >= 0.f ? y.x : y.x * 10.f;
              f[2 * e + 1] += y.y >= 0.f ? y.y : y.y * 10.f;
            }
          }
          if (p.accum_mode != UACC_NONE) {
            float4* s = reinterpret_cast<float4*>(p.accum + eoff);
            if (p.accum_mode == UACC_SET) {
              s[0] = make_float4(f[0], f[1], f[2], f[3]);
              s[1] = make_float4(f[4], f[5], f[6], f[7]);
            } else {
              float4 s0 = s[0], s1 = s[1];
              f[0] += s0.x; f[1] += s0.y; f[2] += s0.z; f[3] += s0.w;
              f[4] += s1.x; f[5] += s1.y; f[6] += s1.z; f[7] += s1.w;
              if (p.accum_mode == UACC_ADD) {
                s[0] = make_float4(f[0], f[1], f[2], f[3]);
                s[1] = make_float4(f[4], f[5], f[6], f[7]);
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = f[e] / p.accum_div;
              }
            }
          }
          if (p.out && (p.accum_mode == UACC_NONE || p.accum_mode == UACC_FINAL)) {
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(act_apply(f[2 * e], p.act_out), act_apply(f[2 * e + 1], p.act_out));
            *reinterpret_cast<uint4*>(p.out + eoff) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ---- glue kernels ----------------------------------------------------------------------------------
// packed fp32 [rows, C] (utterance b at rows start[b]..) -> planar fp16 with gaps
__global__ void to_planar_kernel(__half* out, long long plane_stride, const float* in, int C, const int* start, const int* pstart,
                                 const int* len, int act) {
  int b = blockIdx.y;
  int t = blockIdx.x * blockDim.y + threadIdx.y;
  if (t >= len[b]) return;
  const float* row = in + (size_t)(start[b] + t) * C;
  for (int pl = threadIdx.x; pl < C / 8; pl += blockDim.x) {
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(act_apply(row[pl * 8 + 2 * e], act), act_apply(row[pl * 8 + 2 * e + 1], act));
    *reinterpret_cast<uint4*>(out + (size_t)pl * plane_stride + (size_t)(pstart[b] + t) * 8) = o;
  }
}

// planar fp16 -> packed fp32 [rows, C] (debug / tests)
__global__ void from_planar_kernel(float* out, const __half* in, long long plane_stride, int C, const int* start, const int* pstart,
                                   const int* len) {
  int b = blockIdx.y;
  int t = blockIdx.x * blockDim.y + threadIdx.y;
  if (t >= len[b]) return;
  float* row = out + (size_t)(start[b] + t) * C;
  for (int pl = threadIdx.x; pl < C / 8; pl += blockDim.x) {
    uint4 r = *reinterpret_cast<const uint4*>(in + (size_t)pl * plane_stride + (size_t)(pstart[b] + t) * 8);
    const __half2* rh = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 y = __half22float2(rh[e]);
      row[pl * 8 + 2 * e] = y.x;
      row[pl * 8 + 2 * e + 1] = y.y;
    }
  }
}

// zero the GAP rows before each utterance, after the last one, plus the tail
__global__ void zero_gaps_kernel(__half* buf, long long plane_stride, int planes, const int* pstart, const int* len, int n_utt,
                                 int mul, int tail) {
  // block (x: gap index 0..n_utt, y: plane)
  int gi = blockIdx.x, pl = blockIdx.y;
  long long r0, r1;
  if (gi == 0) {
    r0 = 0;
    r1 = pstart[0];
  } else {
    r0 = (long long)pstart[gi - 1] + (long long)len[gi - 1] * mul;
    r1 = (gi < n_utt) ? pstart[gi] : r0 + tail;
  }
  uint4 z = make_uint4(0, 0, 0, 0);
  for (long long r = r0 + threadIdx.x; r < r1; r += blockDim.x)
    *reinterpret_cast<uint4*>(buf + (size_t)pl * plane_stride + r * 8) = z;
}

// decoder post: out[t] = tanh(sum_j sum_c w[c][j] * x[t+j-pad][c]); x planar fp16 already activated (lrelu 0.01)
__global__ void post_planar_kernel(float* wave, const __half* x, long long plane_stride, const float* w, int C, int k, const int* pstart,
                                   const int* wstart, const int* len) {
  extern __shared__ float ws[];  // [k][C]
  for (int i = threadIdx.x; i < C * k; i += blockDim.x) {
    int c = i / k, j = i % k;
    ws[j * C + c] = w[i];
  }
  __syncthreads();
  int b = blockIdx.y;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= len[b]) return;
  const int pad = (k - 1) / 2;
  float acc = 0.f;
  for (int j = 0; j < k; ++j) {
    long long r = (long long)pstart[b] + t + j - pad;  // gap rows are zero: no bounds test needed
    for (int pl = 0; pl < C / 8; ++pl) {
      uint4 q = *reinterpret_cast<const uint4*>(x + (size_t)pl * plane_stride + r * 8);
      const __half2* qh = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 y = __half22float2(qh[e]);
        acc = fmaf(ws[j * C + pl * 8 + 2 * e], y.x, acc);
        acc = fmaf(ws[j * C + pl * 8 + 2 * e + 1], y.y, acc);
      }
    }
  }
  wave[(size_t)wstart[b] + t] = tanhf(acc);
}

// ---- host side ---------------------------------------------------------------------------------------
struct ConvLayer {
  __half* w = nullptr;
  float* bias = nullptr;
  int cin = 0, cout = 0, nb = 0, n_nblk = 1, taps = 1, kc = 64, nkc = 1, mt = 1, sps = 1, nstages = 2, nloads = 1, total_steps = 1;
  int tap_shift[MAX_TAPS] = {0};
  int halo_lo = 0, halo_hi = 0;
  size_t smem = 0;
  int tmem_cols = 32;
  unsigned idesc = 0;
};

int pow2_at_least(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

uint16_t f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }

// wsel(co, ci, tap) returns the weight of output channel co, input channel ci, tap index `tap`
template <class WSel>
ConvLayer make_layer(sbv2_model* owner, int cin, int cout, int taps, const int* shifts, const float* bias, WSel wsel, int mt_pref) {
  ConvLayer L;
  L.cin = cin;
  L.cout = cout;
  L.taps = taps;
  if (taps > MAX_TAPS) fail(SBV2_ERR_UNSUPPORTED, "conv has too many taps for the tensor-core plan");
  if (cin % 16 != 0 || cout % 16 != 0) fail(SBV2_ERR_UNSUPPORTED, "tensor-core conv needs channel counts divisible by 16");
  L.nb = std::min(cout, 256);
  if (cout % L.nb != 0) fail(SBV2_ERR_UNSUPPORTED, "cout not divisible by the N block");
  L.n_nblk = cout / L.nb;
  L.kc = std::min(cin, 64);
  if (cin % L.kc != 0) fail(SBV2_ERR_UNSUPPORTED, "cin not divisible by the K chunk");
  L.nkc = cin / L.kc;
  if (L.nkc > MAX_KCHUNKS) fail(SBV2_ERR_UNSUPPORTED, "too many K chunks");
  int lo = 0, hi = 0;
  for (int i = 0; i < taps; ++i) {
    L.tap_shift[i] = shifts[i];
    lo = std::min(lo, shifts[i]);
    hi = std::max(hi, shifts[i]);
  }
  L.halo_lo = -lo;
  L.halo_hi = hi;
  if (L.halo_lo > GAP || L.halo_hi > GAP) fail(SBV2_ERR_UNSUPPORTED, "conv halo exceeds the packing gap");
  L.total_steps = taps * L.nkc;
  const size_t step_bytes = size_t(L.nb) * L.kc * 2;
  L.sps = int(std::max<size_t>(1, (32 * 1024) / step_bytes));
  L.sps = std::min(L.sps, L.total_steps);
  L.nloads = (L.total_steps + L.sps - 1) / L.sps;
  // choose MT (accumulators per CTA) and stage count under the smem / TMEM limits
  int mt = mt_pref;
  while (mt > 1 && mt * L.nb > 512) mt >>= 1;
  for (;; mt >>= 1) {
    size_t a_bytes = (size_t(cin / 8) * (128 * mt + L.halo_lo + L.halo_hi) * 16 + 127) & ~size_t(127);
    int ns = std::min(MAX_STAGES, L.nloads);
    while (ns > 1 && a_bytes + ns * step_bytes * L.sps + 256 > size_t(SMEM_LIMIT)) --ns;
    if (a_bytes + ns * step_bytes * L.sps + 256 <= size_t(SMEM_LIMIT) && (ns >= 2 || L.nloads == 1)) {
      L.mt = mt;
      L.nstages = ns;
      L.smem = a_bytes + ns * step_bytes * L.sps + 256;
      break;
    }
    if (mt == 1) fail(SBV2_ERR_UNSUPPORTED, "conv tile does not fit in shared memory");
  }
  L.tmem_cols = pow2_at_least(L.mt * L.nb);
  L.idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((unsigned)(L.nb >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
  // pack weights: [nblk][tap][kc][KC/8][NB][8]
  std::vector<uint16_t> pk(size_t(L.n_nblk) * L.total_steps * L.nb * L.kc);
  size_t o = 0;
  for (int nbk = 0; nbk < L.n_nblk; ++nbk)
    for (int tap = 0; tap < taps; ++tap)
      for (int kc = 0; kc < L.nkc; ++kc)
        for (int pl = 0; pl < L.kc / 8; ++pl)
          for (int n = 0; n < L.nb; ++n)
            for (int e = 0; e < 8; ++e) pk[o++] = f2h(wsel(nbk * L.nb + n, kc * L.kc + pl * 8 + e, tap));
  L.w = static_cast<__half*>(owner->upload_bytes(pk.data(), pk.size() * 2));
  std::vector<float> bz(size_t(cout), 0.f);
  if (bias) std::copy(bias, bias + cout, bz.begin());
  L.bias = static_cast<float*>(owner->upload_bytes(bz.data(), bz.size() * 4));
  return L;
}

ConvLayer make_conv1d(sbv2_model* owner, const HostConv& c, int dil, int mt_pref) {
  // Conv1d weight [Cout][Cin][k]
  int shifts[MAX_TAPS];
  if (c.k > MAX_TAPS) fail(SBV2_ERR_UNSUPPORTED, "kernel size too large");
  for (int j = 0; j < c.k; ++j) shifts[j] = (j - (c.k - 1) / 2) * dil;
  const int cin = c.d1, k = c.k;
  const float* w = c.w.data();
  return make_layer(owner, c.d1, c.d0, c.k, shifts, c.b.empty() ? nullptr : c.b.data(),
                    [=](int co, int ci, int tap) { return w[(size_t(co) * cin + ci) * k + tap]; }, mt_pref);
}

// phase r (= output index mod u) of ConvTranspose1d weight [Cin][Cout][k], stride u, pad (k-u)/2
ConvLayer make_up_phase(sbv2_model* owner, const HostConv& c, int u, int r, int mt_pref) {
  const int k = c.k, pad = (k - u) / 2, taps = k / u;
  const int s = r + pad, rr = s % u, cc = s / u;
  int shifts[MAX_TAPS];
  for (int m = 0; m < taps; ++m) shifts[m] = cc - m;
  const int cout = c.d1;
  const float* w = c.w.data();
  return make_layer(owner, c.d0, c.d1, taps, shifts, c.b.empty() ? nullptr : c.b.data(),
                    [=](int co, int ci, int tap) { return w[(size_t(ci) * cout + co) * k + rr + u * tap]; }, mt_pref);
}

}  // namespace

// ---- geometry -----------------------------------------------------------------------------------------
namespace {

struct Geom {            // one time resolution of one batch
  int mul = 1;
  long long rows_tot = 0;
  std::vector<int> pstart, len;
  const int* d_pstart = nullptr;
  const int* d_len = nullptr;
  const int* d_prefix[3] = {nullptr, nullptr, nullptr};  // mt = 1, 2, 4
  int n_tiles[3] = {0, 0, 0};
  int max_len = 0;
};

int mt_slot(int mt) { return mt == 1 ? 0 : (mt == 2 ? 1 : 2); }

void set_smem_attr() {
  static bool done = false;
  if (!done) {
    CUDA_CHECK(cudaFuncSetAttribute(umma_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    done = true;
  }
}

struct ConvCall {
  const __half* in = nullptr;
  __half* out = nullptr;
  const __half* residual = nullptr;
  float* accum = nullptr;
  int accum_mode = UACC_NONE;
  float accum_div = 1.f;
  int act_out = ACT_NONE;
  const float* bias_utt = nullptr;
  int out_mul = 1, out_off = 0;
};

int g_dbg_swap = 0;

void launch_umma(const LaunchCtx& ctx, const ConvLayer& L, const Geom& gi, const Geom& go, const ConvCall& c, int n_utt) {
  set_smem_attr();
  UmmaConvArgs a;
  a.in = c.in;
  a.in_plane_stride = gi.rows_tot * 8;
  a.out = c.out;
  a.out_plane_stride = go.rows_tot * 8;
  a.residual = c.residual;
  a.accum = c.accum;
  a.w = L.w;
  a.bias = L.bias;
  a.bias_utt = c.bias_utt;
  const int slot = mt_slot(L.mt);
  a.tile_prefix = gi.d_prefix[slot];
  a.pstart_in = gi.d_pstart;
  a.pstart_out = go.d_pstart;
  a.len = gi.d_len;
  a.n_utt = n_utt;
  a.cin = L.cin;
  a.nb = L.nb;
  a.taps = L.taps;
  a.kc = L.kc;
  a.nkc = L.nkc;
  a.mt = L.mt;
  a.sps = L.sps;
  a.nstages = L.nstages;
  a.nloads = L.nloads;
  a.total_steps = L.total_steps;
  for (int i = 0; i < MAX_TAPS; ++i) a.tap_shift[i] = L.tap_shift[i];
  a.halo_lo = L.halo_lo;
  a.halo_hi = L.halo_hi;
  a.out_mul = c.out_mul;
  a.out_off = c.out_off;
  a.act_out = c.act_out;
  a.res_mode = c.residual ? RES_LRELU_INV : RES_NONE;
  a.accum_mode = c.accum_mode;
  a.accum_div = c.accum_div;
  a.tmem_cols = L.tmem_cols;
  a.idesc = L.idesc;
  a.dbg_swap = g_dbg_swap;
  if (gi.n_tiles[slot] <= 0) return;
  dim3 grid(gi.n_tiles[slot], L.n_nblk);
  umma_conv_kernel<<<grid, 192, L.smem, ctx.stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  ctx.count();
}

void launch_zero_gaps(const LaunchCtx& ctx, __half* buf, int C, const Geom& g, int n_utt) {
  dim3 grid(n_utt + 1, C / 8);
  // len passed at this geometry's resolution (mul = 1 because g.len is already scaled)
  zero_gaps_kernel<<<grid, 64, 0, ctx.stream>>>(buf, g.rows_tot * 8, C / 8, g.d_pstart, g.d_len, n_utt, 1, GAP);
  CUDA_CHECK(cudaGetLastError());
  ctx.count();
}

// Builds the geometries of every resolution of a batch and uploads them in one blob.
// frame-level lengths ylen, multipliers mul[0..n]; returns geoms; also uploads start arrays
// (fp32-packed row starts) and wave starts.
struct BatchGeom {
  std::vector<Geom> g;
  const int* d_ystart = nullptr;  // packed fp32 rows (frame level)
  const int* d_wstart = nullptr;  // sample offsets in the wave buffer
};

BatchGeom build_geoms(sbv2_model* owner, DBuf& dev, PinnedBuf& pin, const std::vector<int>& ystart, const std::vector<int>& ylen,
                      const std::vector<int>& muls) {
  const int B = int(ylen.size());
  BatchGeom bg;
  bg.g.resize(muls.size());
  // ints: per geom: pstart[B], len[B], prefix x3 [(B+1)*3]; then ystart[B], wstart[B]
  const size_t per = size_t(2) * B + size_t(3) * (B + 1);
  const size_t total = per * muls.size() + size_t(2) * B;
  pin.ensure(total * 4);
  int* h = pin.as<int>();
  for (size_t s = 0; s < muls.size(); ++s) {
    Geom& G = bg.g[s];
    G.mul = muls[s];
    G.pstart.resize(B);
    G.len.resize(B);
    long long r = GAP;
    int* hp = h + per * s;
    for (int b = 0; b < B; ++b) {
      long long l = (long long)ylen[b] * muls[s];
      if (r + l + GAP + TAIL_ROWS > 2000000000LL) fail(SBV2_ERR_INVALID_ARGUMENT, "batch too long for the decoder's 32-bit row index");
      G.pstart[b] = int(r);
      G.len[b] = int(l);
      G.max_len = std::max(G.max_len, int(l));
      hp[b] = int(r);
      hp[B + b] = int(l);
      r += l + GAP;
    }
    G.rows_tot = r + TAIL_ROWS;
    for (int slot = 0; slot < 3; ++slot) {
      int tm = 128 << slot;
      int* pf = hp + 2 * B + slot * (B + 1);
      int acc = 0;
      for (int b = 0; b < B; ++b) {
        pf[b] = acc;
        acc += (G.len[b] + tm - 1) / tm;
      }
      pf[B] = acc;
      G.n_tiles[slot] = acc;
    }
  }
  int* tail = h + per * muls.size();
  long long hop = muls.back();
  for (int b = 0; b < B; ++b) {
    tail[b] = ystart[b];
    tail[B + b] = int((long long)ystart[b] * hop);
  }
  dev.stream = owner->stream;
  dev.ensure(total * 4);
  CUDA_CHECK(cudaMemcpyAsync(dev.p, h, total * 4, cudaMemcpyHostToDevice, owner->stream));
  const int* d = dev.as<int>();
  for (size_t s = 0; s < muls.size(); ++s) {
    Geom& G = bg.g[s];
    const int* dp = d + per * s;
    G.d_pstart = dp;
    G.d_len = dp + B;
    for (int slot = 0; slot < 3; ++slot) G.d_prefix[slot] = dp + 2 * B + slot * (B + 1);
  }
  bg.d_ystart = d + per * muls.size();
  bg.d_wstart = bg.d_ystart + B;
  return bg;
}

void launch_to_planar(const LaunchCtx& ctx, __half* out, const float* in, int C, const int* d_start, const Geom& g, int n_utt, int act) {
  dim3 block(std::min(32, C / 8), 8);
  dim3 grid((g.max_len + 7) / 8, n_utt);
  to_planar_kernel<<<grid, block, 0, ctx.stream>>>(out, g.rows_tot * 8, in, C, d_start, g.d_pstart, g.d_len, act);
  CUDA_CHECK(cudaGetLastError());
  ctx.count();
}

void launch_from_planar(const LaunchCtx& ctx, float* out, const __half* in, int C, const int* d_start, const Geom& g, int n_utt) {
  dim3 block(std::min(32, C / 8), 8);
  dim3 grid((g.max_len + 7) / 8, n_utt);
  from_planar_kernel<<<grid, block, 0, ctx.stream>>>(out, in, g.rows_tot * 8, C, d_start, g.d_pstart, g.d_len);
  CUDA_CHECK(cudaGetLastError());
  ctx.count();
}

}  // namespace

// ---- decoder plan -----------------------------------------------------------------------------------------
struct UmmaDecoder {
  int gin = 0, cin0 = 0, c0 = 0, n_stages = 0, per = 0, post_c = 0, post_k = 0;
  ConvLayer pre;
  float* cond_w = nullptr;  // fp32 [1][gin][c0]
  float* cond_b = nullptr;
  std::vector<std::vector<ConvLayer>> ups;  // [stage][phase]
  std::vector<int> up_u, stage_c;
  std::vector<std::vector<ConvLayer>> c1, c2;  // [resblock][layer]
  float* post_w = nullptr;
  DBuf zp, xs, xu, t1, r, sum, meta, gcond;
  PinnedBuf pin_meta;
};

UmmaDecoder* umma_decoder_create(const DecoderHostWeights& w, sbv2_model* owner) {
  std::unique_ptr<UmmaDecoder> D(new UmmaDecoder());
  D->gin = w.gin;
  D->cin0 = w.pre.d1;
  D->c0 = w.pre.d0;
  D->n_stages = int(w.ups.size());
  D->per = w.per;
  D->pre = make_conv1d(owner, w.pre, 1, 2);
  {
    // cond: Conv1d(gin -> c0, 1) evaluated on [B, gin] rows with the fp32 kernel
    std::vector<float> t(size_t(w.cond.d0) * w.cond.d1);
    for (int co = 0; co < w.cond.d0; ++co)
      for (int ci = 0; ci < w.cond.d1; ++ci) t[size_t(ci) * w.cond.d0 + co] = w.cond.w[size_t(co) * w.cond.d1 + ci];
    D->cond_w = owner->upload_f32(t);
    D->cond_b = owner->upload_f32(w.cond.b);
  }
  int C = D->c0;
  for (int s = 0; s < D->n_stages; ++s) {
    const HostConv& U = w.ups[s];
    if (U.d0 != C) fail(SBV2_ERR_UNSUPPORTED, "decoder upsample channel mismatch");
    std::vector<ConvLayer> phases;
    for (int r = 0; r < w.up_u[s]; ++r) phases.push_back(make_up_phase(owner, U, w.up_u[s], r, 4));
    D->ups.push_back(phases);
    D->up_u.push_back(w.up_u[s]);
    C = U.d1;
    D->stage_c.push_back(C);
    for (int j = 0; j < w.per; ++j) {
      size_t rb = size_t(s) * w.per + j;
      std::vector<ConvLayer> l1, l2;
      for (size_t l = 0; l < w.res_c1[rb].size(); ++l) {
        l1.push_back(make_conv1d(owner, w.res_c1[rb][l], w.res_dil[rb][l], 4));
        l2.push_back(make_conv1d(owner, w.res_c2[rb][l], 1, 4));
      }
      D->c1.push_back(l1);
      D->c2.push_back(l2);
    }
  }
  D->post_c = w.post.d1;
  D->post_k = w.post.k;
  if (D->post_c != C || D->post_c % 8 != 0) fail(SBV2_ERR_UNSUPPORTED, "decoder conv_post channel mismatch");
  D->post_w = owner->upload_f32(w.post.w);
  for (DBuf* b : {&D->zp, &D->xs, &D->xu, &D->t1, &D->r, &D->sum, &D->meta, &D->gcond}) b->stream = owner->stream;
  return D.release();
}

void umma_decoder_free(UmmaDecoder* d) { delete d; }

void umma_decoder_run(UmmaDecoder* D, sbv2_model* owner, const float* z, const float* g, int B, const std::vector<int>& ystart,
                      const std::vector<int>& ylen, float* wave) {
  LaunchCtx ctx = owner->ctx();
  std::vector<int> muls(1, 1);
  for (int s = 0; s < D->n_stages; ++s) muls.push_back(muls.back() * D->up_u[s]);
  BatchGeom bg = build_geoms(owner, D->meta, D->pin_meta, ystart, ylen, muls);
  // buffer sizes
  size_t max_half = size_t(bg.g[0].rows_tot) * D->c0;
  for (int s = 0; s < D->n_stages; ++s) max_half = std::max(max_half, size_t(bg.g[s + 1].rows_tot) * D->stage_c[s]);
  D->zp.ensure(size_t(bg.g[0].rows_tot) * D->cin0 * 2);
  D->xs.ensure(max_half * 2);
  D->xu.ensure(max_half * 2);
  D->t1.ensure(max_half * 2);
  D->r.ensure(max_half * 2);
  D->sum.ensure(max_half * 4);
  D->gcond.ensure(size_t(B) * D->c0 * 4);
  __half* zp = D->zp.as<__half>();
  __half* xs = D->xs.as<__half>();
  __half* xu = D->xu.as<__half>();
  __half* t1 = D->t1.as<__half>();
  __half* r = D->r.as<__half>();
  float* sum = D->sum.as<float>();
  float* gcond = D->gcond.as<float>();

  // cond(g) -> per-utterance bias of conv_pre
  {
    // one segment of B rows: reuse the first geometry's arrays is not possible; a tiny dedicated pair lives after wstart
    // (start = 0, len = B) — build it on the fly in the gcond buffer's tail is overkill: launch with an explicit Segs
    // whose arrays are the (ystart-independent) prefix of geometry 0: prefix[0] == 0 and we need len == B.
    static_assert(sizeof(int) == 4, "");
  }
  {
    // Segs {start=[0], len=[B]}: store in pinned+device meta tail
    // (appended by build_geoms would complicate its layout; use a small separate upload)
    int two[2] = {0, B};
    D->gcond.ensure(size_t(B) * D->c0 * 4 + 64);
    gcond = D->gcond.as<float>();
    int* d_two = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(gcond) + size_t(B) * D->c0 * 4);
    CUDA_CHECK(cudaMemcpyAsync(d_two, two, 8, cudaMemcpyHostToDevice, owner->stream));
    ConvArgs a;
    a.in = g;
    a.in_ld = D->gin;
    a.w = D->cond_w;
    a.bias = D->cond_b;
    a.out = gcond;
    a.out_ld = D->c0;
    a.cin = D->gin;
    a.cout = D->c0;
    a.seg.start = d_two;
    a.seg.len = d_two + 1;
    a.seg.n = 1;
    a.seg.max_len = B;
    launch_conv(ctx, a);
  }

  const Geom& G0 = bg.g[0];
  launch_zero_gaps(ctx, zp, D->cin0, G0, B);
  launch_to_planar(ctx, zp, z, D->cin0, bg.d_ystart, G0, B, ACT_NONE);
  launch_zero_gaps(ctx, xs, D->c0, G0, B);
  {
    ConvCall c;
    c.in = zp;
    c.out = xs;
    c.act_out = ACT_LRELU;
    c.bias_utt = gcond;
    launch_umma(ctx, D->pre, G0, G0, c, B);
  }
  for (int s = 0; s < D->n_stages; ++s) {
    const Geom& Gi = bg.g[s];
    const Geom& Go = bg.g[s + 1];
    const int C = D->stage_c[s];
    const int u = D->up_u[s];
    launch_zero_gaps(ctx, xu, C, Go, B);
    for (int ph = 0; ph < u; ++ph) {
      ConvCall c;
      c.in = xs;
      c.out = xu;
      c.act_out = ACT_LRELU;
      c.out_mul = u;
      c.out_off = ph;
      launch_umma(ctx, D->ups[s][ph], Gi, Go, c, B);
    }
    launch_zero_gaps(ctx, t1, C, Go, B);
    launch_zero_gaps(ctx, r, C, Go, B);
    const bool last_stage = s + 1 == D->n_stages;
    for (int j = 0; j < D->per; ++j) {
      size_t rb = size_t(s) * D->per + j;
      const __half* cur = xu;
      const size_t nl = D->c1[rb].size();
      for (size_t l = 0; l < nl; ++l) {
        {
          ConvCall c;
          c.in = cur;
          c.out = t1;
          c.act_out = ACT_LRELU;
          launch_umma(ctx, D->c1[rb][l], Go, Go, c, B);
        }
        ConvCall c;
        c.in = t1;
        c.residual = cur;
        if (l + 1 < nl) {
          c.out = r;
          c.act_out = ACT_LRELU;
        } else {
          c.accum = sum;
          c.accum_div = float(D->per);
          if (D->per == 1) {
            c.accum_mode = UACC_FINAL;  // (0 + v)/1 — needs sum zero: use SET semantics via out only
          }
          if (j == 0 && D->per > 1) c.accum_mode = UACC_SET;
          else if (j + 1 < D->per) c.accum_mode = UACC_ADD;
          else c.accum_mode = UACC_FINAL;
          if (c.accum_mode == UACC_FINAL) {
            launch_zero_gaps(ctx, xs, C, Go, B);
            c.out = xs;
            c.act_out = last_stage ? ACT_LRELU01 : ACT_LRELU;
          }
        }
        launch_umma(ctx, D->c2[rb][l], Go, Go, c, B);
        cur = r;
      }
    }
  }
  const Geom& GL = bg.g.back();
  {
    dim3 grid((GL.max_len + 255) / 256, B);
    post_planar_kernel<<<grid, 256, sizeof(float) * D->post_c * D->post_k, ctx.stream>>>(wave, xs, GL.rows_tot * 8, D->post_w, D->post_c,
                                                                                        D->post_k, GL.d_pstart, bg.d_wstart, GL.d_len);
    CUDA_CHECK(cudaGetLastError());
    ctx.count();
  }
  // the pinned geometry blob is rewritten by the next run: make sure its upload finished
  // (it did: every kernel above depends on it and the caller synchronises before returning results)
}

}  // namespace sbv2

// ---- test hook: one convolution through both the fp32 kernel and the tensor-core kernel -------------------
extern "C" int sbv2_debug_conv_compare(const float* x, int64_t T, int cin, const float* w, const float* bias, int cout, int k,
                                       int dil, int mt_pref, int with_residual, int dbg_swap, float* out_umma, float* out_ref) {
  using namespace sbv2;
  return guarded([&] {
    SBV2_REQUIRE(x && w && out_umma && out_ref && T > 0, "bad arguments");
    sbv2_model owner;
    owner.device = 0;
    CUDA_CHECK(cudaSetDevice(0));
    CUDA_CHECK(cudaStreamCreateWithFlags(&owner.stream, cudaStreamNonBlocking));
    LaunchCtx ctx = owner.ctx();
    HostConv hc;
    hc.d0 = cout;
    hc.d1 = cin;
    hc.k = k;
    hc.w.assign(w, w + size_t(cout) * cin * k);
    if (bias) hc.b.assign(bias, bias + cout);
    ConvLayer L = make_conv1d(&owner, hc, dil, mt_pref);
    // fp32 reference weights [k][cin][cout]
    std::vector<float> wr(size_t(k) * cin * cout);
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int j = 0; j < k; ++j) wr[(size_t(j) * cin + ci) * cout + co] = w[(size_t(co) * cin + ci) * k + j];
    float* d_wr = owner.upload_f32(wr);
    float* d_b = bias ? static_cast<float*>(owner.upload_bytes(bias, size_t(cout) * 4)) : nullptr;
    float* d_x = static_cast<float*>(owner.upload_bytes(x, size_t(T) * cin * 4));
    DBuf meta, xin, xout, xres, ref, back;
    PinnedBuf pin;
    for (DBuf* b : {&meta, &xin, &xout, &xres, &ref, &back}) b->stream = owner.stream;
    std::vector<int> ystart{0}, ylen{int(T)}, muls{1};
    BatchGeom bg = build_geoms(&owner, meta, pin, ystart, ylen, muls);
    const Geom& G = bg.g[0];
    xin.ensure(size_t(G.rows_tot) * cin * 2);
    xout.ensure(size_t(G.rows_tot) * cout * 2);
    ref.ensure(size_t(T) * cout * 4);
    back.ensure(size_t(T) * cout * 4);
    launch_zero_gaps(ctx, xin.as<__half>(), cin, G, 1);
    launch_to_planar(ctx, xin.as<__half>(), d_x, cin, bg.d_ystart, G, 1, ACT_NONE);
    ConvCall c;
    c.in = xin.as<__half>();
    c.out = xout.as<__half>();
    if (with_residual) {
      SBV2_REQUIRE(cin == cout, "residual test needs cin == cout");
      c.residual = xin.as<__half>();  // interpreted as lrelu-stored values
    }
    g_dbg_swap = dbg_swap;
    launch_umma(ctx, L, G, G, c, 1);
    g_dbg_swap = 0;
    launch_from_planar(ctx, back.as<float>(), xout.as<__half>(), cout, bg.d_ystart, G, 1);
    // reference
    ConvArgs a;
    a.in = d_x;
    a.in_ld = cin;
    a.w = d_wr;
    a.bias = d_b;
    a.out = ref.as<float>();
    a.out_ld = cout;
    a.cin = cin;
    a.cout = cout;
    a.taps = k;
    a.dil = dil;
    a.off = -dil * ((k - 1) / 2);
    a.seg.start = bg.d_ystart;      // [0]
    a.seg.len = G.d_len;            // [T]
    a.seg.n = 1;
    a.seg.max_len = int(T);
    launch_conv(ctx, a);
    CUDA_CHECK(cudaMemcpyAsync(out_umma, back.p, size_t(T) * cout * 4, cudaMemcpyDeviceToHost, owner.stream));
    CUDA_CHECK(cudaMemcpyAsync(out_ref, ref.p, size_t(T) * cout * 4, cudaMemcpyDeviceToHost, owner.stream));
    CUDA_CHECK(cudaStreamSynchronize(owner.stream));
  });
}
