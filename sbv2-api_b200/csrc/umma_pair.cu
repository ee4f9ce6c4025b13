// Fused ResBlock1 pair on tcgen05:  out = conv2(lrelu(conv1(lrelu(x)) + b1)) + b2 + x   (HiFi-GAN MRF,
// oracle/vits.py ResBlock1.forward) with the intermediate activation kept in shared memory.
//
// Activations are the post-LeakyReLU planar fp16 tensors of umma_conv.cu.  Per work item the CTA
//   phase 1: computes t1 = lrelu(conv1(y) + b1) for 128*MT rows (dilated taps = shifted descriptors on
//            the halo'd input chunk ring), epilogue-1 writes it as fp16 into a shared-memory tile in
//            the same plane-major layout, zeroing rows outside the utterance (conv2's zero padding);
//   phase 2: runs conv2 over that tile (taps = shifted descriptors again) into a second TMEM
//            accumulator set; epilogue-2 adds b2 and the residual (and, for the last pair of the last
//            ResBlock of a stage, the other ResBlocks' outputs and the 1/3 of the MRF mean) and stores.
// Each item yields 128*MT - 2*h2 output rows (h2 = conv2 halo).  The MMA warp is software-pipelined
// P1(i+1) before P2(i), so the tensor pipe runs phase 1 of the next item while the epilogue warps
// convert the current one; both accumulator sets and the t1 tile are double-buffered.
// Compared with two umma_conv launches this removes the write + read of t1 (40 % of the pair's HBM
// traffic) and one of the two global epilogues.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.h"
#include "model.h"
#include "umma_conv.h"
#include "umma_device.cuh"

namespace sbv2 {
namespace {

constexpr int P_MAX_ASLOTS = 8, P_MAX_STAGES = 4, P_EPI_WARPS = 16, P_THREADS = 64 + 32 * P_EPI_WARPS + 32;
constexpr int P_B_PRODUCER_WARP = 2 + P_EPI_WARPS;  // weight stages have their own producer thread (see umma_conv.cu)
constexpr int P_SMEM_LIMIT = 227 * 1024;

struct PairArgs {
  // epilogue-2 fields (names shared with UmmaConvArgs: epilogue_item is a template on the args type)
  __half* out;
  long long out_plane_stride;
  const __half* residual;
  const __half* residual2;
  const __half* residual3;
  float out_div;
  float* accum;
  int accum_mode, act_on_accum;
  float accum_div;
  int act_out;
  float act_slope;
  // the pair
  const __half* in;            // y = lrelu(x), planar fp16 (same geometry as out)
  const __half* w1;            // packed [kc][tap][KC/8][NB][8]
  const __half* w2;
  const float* bias1;
  const float* bias2;
  const int* tile_prefix;      // tiles of out_rows rows
  const int* pstart;
  const int* len;
  int n_utt, n_items;
  int c, taps, kc, nkc, mt, t1rows, t1pitch, out_rows, h1, h2;
  int shift1[UMMA_MAX_TAPS];   // conv1 row shifts ((j - (k-1)/2) * dil)
  int sps, nstages, nloads, total_steps, a_slots, b_resident;
  int has_res;
  int tmem_cols;
  unsigned idesc;
  long long* trace;  // debug timeline of CTA 0: [item][8] clock64 values (sbv2_debug_pair_compare)
};

#define PTRACE(ev, itv)                                                                          \
  do {                                                                                            \
    if (p.trace != nullptr && blockIdx.x == 0 && (itv) < 64) p.trace[(itv) * 16 + (ev)] = clock64(); \
  } while (0)

struct PTile {
  int b, t0, len, pstart;
};
constexpr int P_TABLE_UTTS = 128;  // batches up to this size keep the geometry tables in shared memory
// geometry tables (tile prefix [n+1], len [n], pstart [n]): shared-memory copies when the batch is small enough — the
// binary search is five dependent loads, ~2000 cycles from L2 per epilogue call, ~30 from shared memory
struct PTables {
  const int* prefix;
  const int* len;
  const int* pstart;
};
__device__ __forceinline__ PTile locate_pair_item(const PairArgs& p, const PTables& tb, int item) {
  PTile ti;
  int lo = 0, hi = p.n_utt;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tb.prefix[mid] <= item) lo = mid;
    else hi = mid;
  }
  ti.b = lo;
  ti.t0 = (item - tb.prefix[lo]) * p.out_rows;
  ti.len = tb.len[lo];
  ti.pstart = tb.pstart[lo];
  return ti;
}

template <int C>
struct PairShape {
  static constexpr int c = C, mt = 128 / C, kc = C < 64 ? C : 64, nkc = C / kc, t1rows = 128 * mt, t1pitch = t1rows + 16;
};

// Channel count as a template parameter: the row-tile count, K chunking and tile pitches become constants, so the
// epilogues' index arithmetic (divisions by C / 16, plane and pitch multiplies) folds away — it was a third of the
// kernel's instructions (profiles/r1_ncu_full_summary.txt).
template <int C>
__global__ void __launch_bounds__(P_THREADS, 1) umma_pair_kernel(const __grid_constant__ PairArgs p) {
  using K = PairShape<C>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int RA = K::t1rows + 2 * p.h1;           // input rows per chunk: t1 rows + conv1 halo
  const int ppc = K::kc / 8;
  const uint32_t slot_bytes = (uint32_t)ppc * RA * 16;
  const uint32_t step_bytes = (uint32_t)K::c * K::kc * 2;
  const uint32_t stage_bytes = step_bytes * p.sps;
  const uint32_t t1_bytes = (uint32_t)(K::c / 8) * K::t1pitch * 16;
  const uint32_t sA = smem_u32(smem);
  const uint32_t sT1 = sA + ((slot_bytes * p.a_slots + 127u) & ~127u);
  const uint32_t sB = sT1 + 2 * t1_bytes;
  const uint32_t sBar = sB + (p.b_resident ? 2 * step_bytes * p.total_steps : stage_bytes * p.nstages);
  const uint32_t bar_af = sBar, bar_ae = bar_af + 8 * P_MAX_ASLOTS, bar_bf = bar_ae + 8 * P_MAX_ASLOTS, bar_be = bar_bf + 8 * P_MAX_STAGES,
                 bar_a1f = bar_be + 8 * P_MAX_STAGES, bar_a1e = bar_a1f + 16, bar_a2f = bar_a1e + 16, bar_a2e = bar_a2f + 16,
                 bar_t1r = bar_a2e + 16, bar_t1f = bar_t1r + 16;
  const uint32_t tmem_slot = bar_t1f + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - sA));
  float* bias_s = reinterpret_cast<float*>(smem + (sBar - sA) + 384);  // [2][C]: b1 | b2
  const int acc_cols = K::mt * K::c;  // one accumulator set; TMEM: acc1[0], acc1[1], acc2[0], acc2[1]

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_slots; ++i) {
      mbar_init(bar_af + 8 * i, 1);
      mbar_init(bar_ae + 8 * i, 1);
    }
    for (int i = 0; i < P_MAX_STAGES; ++i) {
      mbar_init(bar_bf + 8 * i, 1);
      mbar_init(bar_be + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_a1f + 8 * i, 1);
      mbar_init(bar_a1e + 8 * i, P_EPI_WARPS);
      mbar_init(bar_a2f + 8 * i, 1);
      mbar_init(bar_a2e + 8 * i, P_EPI_WARPS);
      mbar_init(bar_t1r + 8 * i, P_EPI_WARPS);
      mbar_init(bar_t1f + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * K::c; i += P_THREADS) bias_s[i] = i < K::c ? p.bias1[i] : p.bias2[i - K::c];
  PTables tb{p.tile_prefix, p.len, p.pstart};
  if (p.n_utt <= P_TABLE_UTTS) {
    int* tab = reinterpret_cast<int*>(bias_s + 2 * K::c);  // [n+1 | n | n]
    for (int i = threadIdx.x; i <= p.n_utt; i += P_THREADS) tab[i] = p.tile_prefix[i];
    for (int i = threadIdx.x; i < p.n_utt; i += P_THREADS) {
      tab[p.n_utt + 1 + i] = p.len[i];
      tab[2 * p.n_utt + 1 + i] = p.pstart[i];
    }
    tb.prefix = tab;
    tb.len = tab + p.n_utt + 1;
    tb.pstart = tab + 2 * p.n_utt + 1;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // PDL: everything above overlapped the previous kernel's tail; its results are visible after the wait
  pdl_launch_dependents();
  pdl_wait();
  const int n_my = p.n_items > (int)blockIdx.x ? (p.n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;  // items of this CTA

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- activation producer: A chunks for phase 1 ----------------
      uint32_t a_it = 0;
      auto load_a_item = [&](int it) {
        const PTile ti = locate_pair_item(p, tb, (int)blockIdx.x + it * (int)gridDim.x);
        const long long in_row0 = (long long)ti.pstart + ti.t0 - p.h2 - p.h1;
        for (int kc = 0; kc < K::nkc; ++kc, ++a_it) {
          const uint32_t slot = a_it % p.a_slots;
          mbar_wait(bar_ae + 8 * slot, ((a_it / p.a_slots) & 1) ^ 1);
          mbar_expect_tx(bar_af + 8 * slot, slot_bytes);
          for (int q = 0; q < ppc; ++q) {
            const int plane = kc * ppc + q;
            bulk_g2s(sA + slot_bytes * slot + (uint32_t)q * RA * 16, p.in + (size_t)plane * p.out_plane_stride + in_row0 * 8, (uint32_t)RA * 16,
                     bar_af + 8 * slot);
          }
        }
      };
      for (int it = 0; it < n_my; ++it) load_a_item(it);
    }
  } else if (warp == P_B_PRODUCER_WARP) {
    if (lane == 0) {
      // ---------------- weight producer: stages in MMA consumption order ----------------
      uint32_t b_it = 0;
      auto load_b_all = [&](const __half* w) {  // all steps of one conv through the ring
        for (int i = 0; i < p.nloads; ++i, ++b_it) {
          const uint32_t st = b_it % p.nstages;
          mbar_wait(bar_be + 8 * st, ((b_it / p.nstages) & 1) ^ 1);
          const int first_step = i * p.sps;
          const int nsteps = min(p.sps, p.total_steps - first_step);
          const uint32_t bytes = step_bytes * nsteps;
          mbar_expect_tx(bar_bf + 8 * st, bytes);
          bulk_g2s(sB + stage_bytes * st, w + (size_t)first_step * (step_bytes / 2), bytes, bar_bf + 8 * st);
        }
      };
      if (p.b_resident && n_my > 0) {
        const uint32_t wb = step_bytes * p.total_steps;
        mbar_expect_tx(bar_bf, 2 * wb);
        bulk_g2s(sB, p.w1, wb, bar_bf);
        bulk_g2s(sB + wb, p.w2, wb, bar_bf);
      }
      // MMA order: P1(0), then per item it: P1(it+1), P2(it)
      if (!p.b_resident) {
        if (n_my > 0) load_b_all(p.w1);
        for (int it = 0; it < n_my; ++it) {
          if (it + 1 < n_my) load_b_all(p.w1);
          load_b_all(p.w2);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (warp-uniform control flow, one elected lane issues) ----------------
    const int k16 = K::kc / 16, mt = K::mt;
    const uint32_t c_u = (uint32_t)K::c, idesc = p.idesc;
    const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;  // version 1, SBO = 128 B
    const uint64_t a_desc0 = desc_hi | ((uint64_t)((uint32_t)RA & 0x3FFF) << 16);
    const uint64_t t_desc0 = desc_hi | ((uint64_t)((uint32_t)K::t1pitch & 0x3FFF) << 16);
    const uint64_t b_desc0 = desc_hi | ((uint64_t)((uint32_t)K::c & 0x3FFF) << 16);
    const uint32_t a_kstep = 2u * (uint32_t)RA, t_kstep = 2u * (uint32_t)K::t1pitch, b_kstep = 2u * (uint32_t)K::c;
    uint32_t a_slot_i = 0, a_par = 0, b_st = 0, b_par = 0;
    int b_si = 0;
    bool resident_ready = false;
    // B-step bookkeeping shared by both phases (ring or resident)
    auto b_step_addr = [&](int conv, int step) -> uint32_t {
      if (p.b_resident) {
        if (!resident_ready) {
          mbar_wait(bar_bf, 0);
          resident_ready = true;
        }
        return sB + (uint32_t)conv * step_bytes * p.total_steps + step_bytes * step;
      }
      if (b_si == 0) mbar_wait(bar_bf + 8 * b_st, b_par);
      return sB + stage_bytes * b_st + step_bytes * b_si;
    };
    const bool leader = elect_one_sync() != 0;  // the same lane issues every MMA / commit of this CTA
    auto b_step_done = [&](int step) {
      if (p.b_resident) return;
      ++b_si;
      if (b_si == p.sps || step == p.total_steps - 1) {
        if (leader) tc_commit(bar_be + 8 * b_st);
        b_si = 0;
        if (++b_st == (uint32_t)p.nstages) {
          b_st = 0;
          b_par ^= 1;
        }
      }
    };
    // MT / K16 are compile-time (dispatched once, below): a per-tap switch or elect costs ~140 cycles per tap, which the
    // tensor pipe does not hide — it starts each MMA as it is issued (umma_microbench.cu, "issue shape")
    auto issue_p1 = [&](int it, auto mtk) {
      constexpr int MT = decltype(mtk)::mt, K16 = decltype(mtk)::k16;
      const uint32_t buf = it & 1;
      mbar_wait(bar_a1e + 8 * buf, ((it >> 1) & 1) ^ 1);  // epilogue-1 of item it-2 has drained this set
      tc_fence_after();
      if (lane == 0) PTRACE(0, it);
      const uint32_t tacc0 = tmem_base + buf * acc_cols;
      int step = 0;
      for (int kc = 0; kc < K::nkc; ++kc) {
        mbar_wait(bar_af + 8 * a_slot_i, a_par);
        const uint64_t a_chunk = a_desc0 + ((sA + slot_bytes * a_slot_i) >> 4) + (uint32_t)p.h1;
        for (int tap = 0; tap < p.taps; ++tap, ++step) {
          const uint32_t b_addr = b_step_addr(0, step);
          const uint64_t a_tap = a_chunk + (int64_t)p.shift1[tap];
          const uint64_t b_d = b_desc0 + (b_addr >> 4);
          const uint32_t accf = step > 0 ? 1u : 0u;
          if (leader) issue_mmas<MT, K16>(tacc0, a_tap, b_d, a_kstep, b_kstep, c_u, idesc, accf);
          b_step_done(step);
        }
        if (leader) tc_commit(bar_ae + 8 * a_slot_i);
        if (++a_slot_i == (uint32_t)p.a_slots) {
          a_slot_i = 0;
          a_par ^= 1;
        }
      }
      if (leader) tc_commit(bar_a1f + 8 * buf);
      __syncwarp();
      if (lane == 0) PTRACE(1, it);
    };
    auto issue_p2 = [&](int it, auto mtk) {
      constexpr int MT = decltype(mtk)::mt, K16 = decltype(mtk)::k16;
      const uint32_t buf = it & 1;
      if (lane == 0) PTRACE(8, it);
      mbar_wait(bar_a2e + 8 * buf, ((it >> 1) & 1) ^ 1);  // epilogue-2 of item it-2 has drained this set
      if (lane == 0) PTRACE(9, it);
      mbar_wait(bar_t1r + 8 * buf, (it >> 1) & 1);         // t1 tile written and visible to the async proxy
      tc_fence_after();
      if (lane == 0) PTRACE(2, it);
      const uint32_t tacc0 = tmem_base + (2 + buf) * acc_cols;
      const uint64_t t_tile = t_desc0 + ((sT1 + buf * t1_bytes) >> 4);
      int step = 0;
      for (int kc = 0; kc < K::nkc; ++kc) {
        const uint64_t t_chunk = t_tile + (uint32_t)(kc * ppc * K::t1pitch);
        for (int tap = 0; tap < p.taps; ++tap, ++step) {
          const uint32_t b_addr = b_step_addr(1, step);
          const uint64_t t_tap = t_chunk + (uint32_t)tap;  // out row m, tap j reads t1 row m + j
          const uint64_t b_d = b_desc0 + (b_addr >> 4);
          const uint32_t accf = step > 0 ? 1u : 0u;
          if (leader) issue_mmas<MT, K16>(tacc0, t_tap, b_d, t_kstep, b_kstep, c_u, idesc, accf);
          b_step_done(step);
        }
      }
      if (leader) {
        tc_commit(bar_a2f + 8 * buf);
        tc_commit(bar_t1f + 8 * buf);
      }
      __syncwarp();
      if (lane == 0) PTRACE(3, it);
    };
    auto run = [&](auto mtk) {
      if (n_my > 0) issue_p1(0, mtk);
      for (int it = 0; it < n_my; ++it) {
        if (it + 1 < n_my) issue_p1(it + 1, mtk);
        issue_p2(it, mtk);
      }
    };
    run(MtK<K::mt, K::kc / 16>{});
  } else {
    // ---------------- epilogue warps ----------------
    const int wq = warp & 3;
    const int part = (warp - 2) >> 2;  // 4 warps per TMEM lane quarter
    constexpr int NPART = P_EPI_WARPS / 4;
    const float* b1s = bias_s;
    const float* b2s = bias_s + K::c;
    // epilogue 1: t1 = lrelu(acc + b1) -> fp16 -> shared tile; rows outside [0, len) are conv2's zero padding
    auto epi1 = [&](int it) {
      const uint32_t buf = it & 1;
      const PTile ti = locate_pair_item(p, tb, (int)blockIdx.x + it * (int)gridDim.x);
      mbar_wait(bar_a1f + 8 * buf, (it >> 1) & 1);
      mbar_wait(bar_t1f + 8 * buf, ((it >> 1) & 1) ^ 1);  // phase 2 of item it-2 has finished reading this tile
      tc_fence_after();
      if (threadIdx.x == 64) PTRACE(4, it);
      if (threadIdx.x == 64 + 15 * 32) PTRACE(10, it);
      const uint32_t tacc0 = tmem_base + buf * acc_cols;
      uint8_t* t1 = smem + (sT1 - sA) + buf * t1_bytes;
      const int n_sub = K::mt * (K::c / 16);
      for (int sub = part; sub < n_sub; sub += NPART) {
        const int a = sub / (K::c / 16);
        const int c0 = (sub - a * (K::c / 16)) * 16;
        const int r = a * 128 + wq * 32 + lane;  // t1 row of this thread
        const int pos = ti.t0 - p.h2 + r;
        const bool inside = pos >= 0 && pos < ti.len;
        uint32_t v[16];
        tc_ld16(tacc0 + ((uint32_t)(wq * 32) << 16) + (uint32_t)(a * K::c + c0), v);
        tc_wait_ld();
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          const float4 bb0 = *reinterpret_cast<const float4*>(b1s + c0 + 8 * pl), bb1 = *reinterpret_cast<const float4*>(b1s + c0 + 8 * pl + 4);
          float f[8];
          f[0] = __uint_as_float(v[8 * pl + 0]) + bb0.x; f[1] = __uint_as_float(v[8 * pl + 1]) + bb0.y;
          f[2] = __uint_as_float(v[8 * pl + 2]) + bb0.z; f[3] = __uint_as_float(v[8 * pl + 3]) + bb0.w;
          f[4] = __uint_as_float(v[8 * pl + 4]) + bb1.x; f[5] = __uint_as_float(v[8 * pl + 5]) + bb1.y;
          f[6] = __uint_as_float(v[8 * pl + 6]) + bb1.z; f[7] = __uint_as_float(v[8 * pl + 7]) + bb1.w;
          uint4 o = make_uint4(0, 0, 0, 0);
          if (inside) {
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(fmaxf(f[2 * e], f[2 * e] * 0.1f), fmaxf(f[2 * e + 1], f[2 * e + 1] * 0.1f));
          }
          *reinterpret_cast<uint4*>(t1 + ((size_t)((c0 >> 3) + pl) * K::t1pitch + r) * 16) = o;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // t1 writes -> visible to the MMA's async proxy
      tc_fence_before();
      __syncwarp();
      if (threadIdx.x == 64) PTRACE(5, it);
      if (threadIdx.x == 64 + 15 * 32) PTRACE(11, it);
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a1e + 8 * buf) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_t1r + 8 * buf) : "memory");
      }
    };
    // epilogue 2: out = act((acc + b2 + x [+ r0 + r1]) [/ 3])
    auto epi2 = [&](int it) {
      const uint32_t buf = it & 1;
      const PTile ti = locate_pair_item(p, tb, (int)blockIdx.x + it * (int)gridDim.x);
      mbar_wait(bar_a2f + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 64) PTRACE(6, it);
      if (threadIdx.x == 64 + 15 * 32) PTRACE(12, it);
      const uint32_t tacc0 = tmem_base + (2 + buf) * acc_cols;
      const bool wide = (K::c % 32 == 0) && p.has_res != 3;
      const int nch = wide ? 32 : 16;
      const int per_acc = K::c / nch;
      const int n_sub = K::mt * per_acc;
      for (int sub = part; sub < n_sub; sub += NPART) {
        const int a = sub / per_acc;
        const int c0 = (sub - a * per_acc) * nch;
        const int m = a * 128 + wq * 32 + lane;  // output row within the item
        const int t = ti.t0 + m;
        const bool valid = m < p.out_rows && t < ti.len;
        const long long orow = (long long)ti.pstart + t;
        const uint32_t taddr = tacc0 + ((uint32_t)(wq * 32) << 16) + (uint32_t)(a * K::c + c0);
        if (p.has_res == 3) epilogue_item<16, false, 3>(p, taddr, valid, orow, c0, b2s + c0);
        else if (wide) epilogue_item<32, false, 1>(p, taddr, valid, orow, c0, b2s + c0);
        else epilogue_item<16, false, 1>(p, taddr, valid, orow, c0, b2s + c0);
      }
      tc_fence_before();
      __syncwarp();
      if (threadIdx.x == 64) PTRACE(7, it);
      if (threadIdx.x == 64 + 15 * 32) PTRACE(13, it);
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a2e + 8 * buf) : "memory");
    };
    if (n_my > 0) epi1(0);
    for (int it = 0; it < n_my; ++it) {
      if (it + 1 < n_my) epi1(it + 1);
      epi2(it);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

uint16_t f2h_(float f) { return __half_as_ushort(__float2half_rn(f)); }

__half* pack_conv(sbv2_model* owner, const HostConv& c, int kc) {
  // [kc][tap][KC/8][NB][8], NB = Cout = C
  const int C = c.d0, k = c.k, nkc = C / kc;
  std::vector<uint16_t> pk(size_t(C) * C * k);
  size_t o = 0;
  for (int q = 0; q < nkc; ++q)
    for (int tap = 0; tap < k; ++tap)
      for (int pl = 0; pl < kc / 8; ++pl)
        for (int n = 0; n < C; ++n)
          for (int e = 0; e < 8; ++e) pk[o++] = f2h_(c.w[(size_t(n) * C + q * kc + pl * 8 + e) * k + tap]);
  return static_cast<__half*>(owner->upload_bytes(pk.data(), pk.size() * 2));
}

}  // namespace

long long* g_pair_trace = nullptr;  // debug hook, see sbv2_debug_pair_compare

bool make_pair_layer(sbv2_model* owner, const HostConv& c1, int dil, const HostConv& c2, PairLayer* out) {
  PairLayer L;
  const int C = c1.d0, k = c1.k;
  if (c1.d1 != C || c2.d0 != C || c2.d1 != C || c2.k != k || k % 2 == 0 || k > UMMA_MAX_TAPS) return false;
  if (C != 16 && C != 32 && C != 64 && C != 128) return false;  // instantiated shapes; 4 accumulator sets of MT*C = 128 TMEM columns
  L.c = C;
  L.taps = k;
  L.h1 = dil * (k - 1) / 2;
  L.h2 = (k - 1) / 2;
  if (L.h1 + L.h2 > UMMA_GAP) return false;
  for (int j = 0; j < k; ++j) L.shift1[j] = (j - (k - 1) / 2) * dil;
  L.kc = C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : 16);
  L.nkc = C / L.kc;
  L.total_steps = k * L.nkc;
  L.mt = 128 / C;  // MT * C = 128 columns per accumulator set
  if (L.mt < 1) return false;
  L.t1rows = 128 * L.mt;
  L.t1pitch = L.t1rows + 16;  // phase-2 taps of the last M tile read up to k-1 rows past the tile
  L.out_rows = L.t1rows - 2 * L.h2;
  const size_t step_bytes = size_t(C) * L.kc * 2;
  const size_t w_bytes = step_bytes * L.total_steps;
  const size_t slot = size_t(L.kc / 8) * (L.t1rows + 2 * L.h1) * 16;
  const size_t t1 = size_t(C / 8) * L.t1pitch * 16;
  const size_t misc = 384 + size_t(2) * C * 4 + 256 + size_t(3 * P_TABLE_UTTS + 1) * 4;  // barriers, biases, geometry tables
  const size_t budget = P_SMEM_LIMIT;
  const int min_slots = std::max(2, std::min(2 * L.nkc, P_MAX_ASLOTS));
  if (2 * w_bytes + slot * min_slots + 2 * t1 + misc <= budget) {
    L.b_resident = 1;
    L.sps = L.total_steps;
    L.nloads = 1;
    L.nstages = 1;
    L.a_slots = min_slots;
    while (L.a_slots < P_MAX_ASLOTS && L.a_slots < 3 * L.nkc && 2 * w_bytes + slot * (L.a_slots + 1) + 2 * t1 + misc <= budget) ++L.a_slots;
    L.smem = ((slot * L.a_slots + 127) & ~size_t(127)) + 2 * t1 + 2 * w_bytes + misc;
  } else {
    L.b_resident = 0;
    L.sps = int(std::max<size_t>(1, (16 * 1024) / step_bytes));
    L.sps = std::min(L.sps, L.total_steps);
    L.nloads = (L.total_steps + L.sps - 1) / L.sps;
    const size_t stage_bytes = step_bytes * L.sps;
    bool placed = false;
    for (int ns = P_MAX_STAGES; ns >= 2 && !placed; --ns)
      for (int slots = P_MAX_ASLOTS; slots >= min_slots && !placed; --slots) {
        const size_t sm = ((slot * slots + 127) & ~size_t(127)) + 2 * t1 + stage_bytes * ns + misc;
        if (sm <= budget) {
          L.nstages = ns;
          L.a_slots = slots;
          L.smem = sm;
          placed = true;
        }
      }
    if (!placed) return false;
  }
  L.tmem_cols = 512;
  L.idesc = (1u << 4) | ((unsigned)(C >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
  L.w1 = pack_conv(owner, c1, L.kc);
  L.w2 = pack_conv(owner, c2, L.kc);
  L.bias1 = owner->upload_f32(c1.b);
  L.bias2 = owner->upload_f32(c2.b);
  *out = L;
  return true;
}

void launch_umma_pair(const LaunchCtx& ctx, const PairLayer& L, const Geom& g, const PairCall& c, int n_utt) {
  static PerDeviceOnce attr_once;
  attr_once.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(umma_pair_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_LIMIT));
    CUDA_CHECK(cudaFuncSetAttribute(umma_pair_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_LIMIT));
    CUDA_CHECK(cudaFuncSetAttribute(umma_pair_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_LIMIT));
    CUDA_CHECK(cudaFuncSetAttribute(umma_pair_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_LIMIT));
  });
  auto it = g.extra.find(L.out_rows);
  if (it == g.extra.end()) fail(SBV2_ERR_INTERNAL, "geometry lacks the tile table for a fused ResBlock pair");
  PairArgs a;
  a.out = c.out;
  a.out_plane_stride = g.rows_tot * 8;
  a.residual = c.in;  // x itself (stored post-lrelu) is the residual
  a.residual2 = c.residual2;
  a.residual3 = c.residual3;
  a.out_div = c.out_div;
  a.accum = nullptr;
  a.accum_mode = UACC_NONE;
  a.act_on_accum = 0;
  a.accum_div = 1.f;
  a.act_out = c.act_out;
  a.act_slope = c.act_out == ACT_LRELU ? 0.1f : (c.act_out == ACT_LRELU01 ? 0.01f : (c.act_out == ACT_RELU ? 0.f : 1.f));
  a.in = c.in;
  a.w1 = L.w1;
  a.w2 = L.w2;
  a.bias1 = L.bias1;
  a.bias2 = L.bias2;
  a.tile_prefix = it->second.first;
  a.pstart = g.d_pstart;
  a.len = g.d_len;
  a.n_utt = n_utt;
  a.n_items = it->second.second;
  a.c = L.c;
  a.taps = L.taps;
  a.kc = L.kc;
  a.nkc = L.nkc;
  a.mt = L.mt;
  a.t1rows = L.t1rows;
  a.t1pitch = L.t1pitch;
  a.out_rows = L.out_rows;
  a.h1 = L.h1;
  a.h2 = L.h2;
  for (int i = 0; i < UMMA_MAX_TAPS; ++i) a.shift1[i] = L.shift1[i];
  a.sps = L.sps;
  a.nstages = L.nstages;
  a.nloads = L.nloads;
  a.total_steps = L.total_steps;
  a.a_slots = L.a_slots;
  a.b_resident = L.b_resident;
  a.has_res = c.residual2 ? 3 : 1;
  a.tmem_cols = L.tmem_cols;
  a.idesc = L.idesc;
  a.trace = g_pair_trace;
  if (a.n_items <= 0) return;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int grid = std::min(a.n_items, num_sms);
  if (g_pair_trace != nullptr) {
    if (const char* e = getenv("SBV2_B200_PAIR_GRID")) grid = std::max(1, std::min(grid, atoi(e)));  // debugging: fewer CTAs
  }
  void (*kernel)(PairArgs) = nullptr;
  switch (L.c) {
    case 16: kernel = umma_pair_kernel<16>; break;
    case 32: kernel = umma_pair_kernel<32>; break;
    case 64: kernel = umma_pair_kernel<64>; break;
    case 128: kernel = umma_pair_kernel<128>; break;
    default: fail(SBV2_ERR_INTERNAL, "fused ResBlock pair: unsupported channel count");
  }
  launch_pdl(ctx.pdl, kernel, dim3(grid), dim3(P_THREADS), L.smem, ctx.stream, a);
  ctx.count();
}

}  // namespace sbv2
