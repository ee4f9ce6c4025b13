// Glue kernels of the tensor-core transformer flow: fp32 residual stream + LayerNorm, planar fp16
// operand copies, window-relative attention on planar fp16 q|k|v.
#include <cuda_fp16.h>
#include <math_constants.h>

#include <cstdlib>

#include "kernels.h"

namespace sbv2 {
namespace {

// one warp per row, lane = plane (C/8 <= 32)
__global__ void flow_mix_kernel(float* h_out, __half* hp, const float* h_in, const float* src32, const float* vec, int vec_ld, int C,
                                PlanarSegs s) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= s.len[b] || lane >= C / 8) return;
  const size_t row = (size_t)s.start[b] + t;
  const size_t poff = (size_t)lane * s.plane_stride + (size_t)(s.pstart[b] + t) * 8;
  float4 a0, a1;
  if (src32) {
    a0 = *reinterpret_cast<const float4*>(src32 + poff);
    a1 = *reinterpret_cast<const float4*>(src32 + poff + 4);
  } else {
    a0 = *reinterpret_cast<const float4*>(h_in + row * C + lane * 8);
    a1 = *reinterpret_cast<const float4*>(h_in + row * C + lane * 8 + 4);
  }
  if (vec) {
    const float* v = vec + (size_t)b * vec_ld + lane * 8;
    a0.x += v[0]; a0.y += v[1]; a0.z += v[2]; a0.w += v[3];
    a1.x += v[4]; a1.y += v[5]; a1.z += v[6]; a1.w += v[7];
  }
  *reinterpret_cast<float4*>(h_out + row * C + lane * 8) = a0;
  *reinterpret_cast<float4*>(h_out + row * C + lane * 8 + 4) = a1;
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
  oh[0] = __floats2half2_rn(a0.x, a0.y);
  oh[1] = __floats2half2_rn(a0.z, a0.w);
  oh[2] = __floats2half2_rn(a1.x, a1.y);
  oh[3] = __floats2half2_rn(a1.z, a1.w);
  *reinterpret_cast<uint4*>(hp + poff) = o;
}

// One warp normalises R consecutive rows: all of their loads are issued before the first reduction (the kernel is bound by
// memory latency, not bandwidth: 2.7 KB per row), and a lane's R rows are contiguous in the planar buffers (R * 32 B of y32).
template <int R>
__global__ void ln_planar_kernel(float* h, __half* hp, const float* y32, const float* gamma, const float* beta, float eps, int C,
                                 PlanarSegs s) {
  const int b = blockIdx.y;
  const int t0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
  const int lane = threadIdx.x & 31;
  const int len = s.len[b];
  if (t0 >= len) return;  // warp-uniform
  const bool act = lane < C / 8;
  const size_t row0 = (size_t)s.start[b] + t0;
  const size_t poff0 = (size_t)lane * s.plane_stride + (size_t)(s.pstart[b] + t0) * 8;
  float v[R][8];
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[r][e] = 0.f;
    if (act && t0 + r < len) {
      const float4 x0 = *reinterpret_cast<const float4*>(h + (row0 + r) * C + lane * 8);
      const float4 x1 = *reinterpret_cast<const float4*>(h + (row0 + r) * C + lane * 8 + 4);
      const float4 y0 = *reinterpret_cast<const float4*>(y32 + poff0 + r * 8);
      const float4 y1 = *reinterpret_cast<const float4*>(y32 + poff0 + r * 8 + 4);
      v[r][0] = x0.x + y0.x; v[r][1] = x0.y + y0.y; v[r][2] = x0.z + y0.z; v[r][3] = x0.w + y0.w;
      v[r][4] = x1.x + y1.x; v[r][5] = x1.y + y1.y; v[r][6] = x1.z + y1.z; v[r][7] = x1.w + y1.w;
    }
  }
  float g[8], bt[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    g[e] = act ? gamma[lane * 8 + e] : 0.f;
    bt[e] = act ? beta[lane * 8 + e] : 0.f;
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) sum += v[r][e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float sq = 0.f;
    if (act) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[r][e] - mean;
        sq = fmaf(d, d, sq);  // explicit fma here and below: left to ptxas, R = 1 and R = 4 may be contracted differently
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = 1.0f / sqrtf(sq / (float)C + eps);
    if (!act || t0 + r >= len) continue;
    float q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) q[e] = fmaf(__fmul_rn(v[r][e] - mean, rstd), g[e], bt[e]);
    *reinterpret_cast<float4*>(h + (row0 + r) * C + lane * 8) = make_float4(q[0], q[1], q[2], q[3]);
    *reinterpret_cast<float4*>(h + (row0 + r) * C + lane * 8 + 4) = make_float4(q[4], q[5], q[6], q[7]);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(q[2 * e], q[2 * e + 1]);
    *reinterpret_cast<uint4*>(hp + poff0 + r * 8) = o;
  }
}

__global__ void coupling_sub_planar_kernel(float* z, const float* m32, int C, PlanarSegs s) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int H = C / 2;
  if (t >= s.len[b] || lane >= H / 8) return;
  const size_t row = (size_t)s.start[b] + t;
  const size_t poff = (size_t)lane * s.plane_stride + (size_t)(s.pstart[b] + t) * 8;
  float4* zp = reinterpret_cast<float4*>(z + row * C + H + lane * 8);
  const float4 m0 = *reinterpret_cast<const float4*>(m32 + poff);
  const float4 m1 = *reinterpret_cast<const float4*>(m32 + poff + 4);
  float4 z0 = zp[0], z1 = zp[1];
  z0.x -= m0.x; z0.y -= m0.y; z0.z -= m0.z; z0.w -= m0.w;
  z1.x -= m1.x; z1.y -= m1.y; z1.z -= m1.z; z1.w -= m1.w;
  zp[0] = z0;
  zp[1] = z1;
}

// ---------------------------------------------------------------------------------------------
// window-relative attention, flash style, planar fp16 I/O, fp32 math. 64 queries x 64 keys per step.
// ---------------------------------------------------------------------------------------------
constexpr int AQ = 64, AK = 64;

template <int D>
__global__ void __launch_bounds__(256) rel_attention_planar_kernel(__half* out, const __half* qkv, const float* rel_k,
                                                                   const float* rel_v, int heads, int window, PlanarSegs s) {
  extern __shared__ float sm[];
  constexpr int DC = D / 16;   // output columns per thread
  constexpr int DP = D / 8;    // planes per head
  const int R = 2 * window + 1;
  float* Qt = sm;                    // [D][AQ+1]
  float* Kt = Qt + D * (AQ + 1);     // [D][AK+1]
  float* Vs = Kt + D * (AK + 1);     // [AK][D]
  float* Ps = Vs + AK * D;           // [AQ][AK+1]
  float* Ek = Ps + AQ * (AK + 1);    // [R][D]
  float* Ev = Ek + R * D;            // [R][D]
  float* Qrel = Ev + R * D;          // [AQ][R]

  const int b = blockIdx.z, h = blockIdx.y;
  const int len = s.len[b];
  const int q0 = blockIdx.x * AQ;
  if (q0 >= len) return;
  const long long pbase = s.pstart[b];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float scale = 1.0f / sqrtf((float)D);
  const int q_plane0 = h * DP, k_plane0 = heads * DP + h * DP, v_plane0 = 2 * heads * DP + h * DP;

  for (int i = tid; i < R * D; i += 256) {
    Ek[i] = rel_k[i];
    Ev[i] = rel_v[i];
  }
  for (int i = tid; i < AQ * DP; i += 256) {
    const int r = i % AQ, pl = i / AQ;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = 0.f;
    if (q0 + r < len) {
      const uint4 u = *reinterpret_cast<const uint4*>(qkv + (size_t)(q_plane0 + pl) * s.plane_stride + (pbase + q0 + r) * 8);
      const __half2* uh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 y = __half22float2(uh[e]);
        f[2 * e] = y.x * scale;
        f[2 * e + 1] = y.y * scale;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) Qt[(pl * 8 + e) * (AQ + 1) + r] = f[e];
  }
  __syncthreads();
  for (int i = tid; i < AQ * R; i += 256) {
    const int r = i / R, rr = i % R;
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc = fmaf(Qt[d * (AQ + 1) + r], Ek[rr * D + d], acc);
    Qrel[r * R + rr] = acc;
  }

  float m_run[4], l_run[4], o[4][DC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -CUDART_INF_F;
    l_run[i] = 0.f;
#pragma unroll
    for (int c = 0; c < DC; ++c) o[i][c] = 0.f;
  }

  for (int k0 = 0; k0 < len; k0 += AK) {
    __syncthreads();
    for (int i = tid; i < AK * DP; i += 256) {
      const int r = i % AK, pl = i / AK;
      float kf[8], vf[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) kf[e] = vf[e] = 0.f;
      if (k0 + r < len) {
        const uint4 uk = *reinterpret_cast<const uint4*>(qkv + (size_t)(k_plane0 + pl) * s.plane_stride + (pbase + k0 + r) * 8);
        const uint4 uv = *reinterpret_cast<const uint4*>(qkv + (size_t)(v_plane0 + pl) * s.plane_stride + (pbase + k0 + r) * 8);
        const __half2* kh = reinterpret_cast<const __half2*>(&uk);
        const __half2* vh = reinterpret_cast<const __half2*>(&uv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = __half22float2(kh[e]), c = __half22float2(vh[e]);
          kf[2 * e] = a.x; kf[2 * e + 1] = a.y;
          vf[2 * e] = c.x; vf[2 * e + 1] = c.y;
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        Kt[(pl * 8 + e) * (AK + 1) + r] = kf[e];
        Vs[r * D + pl * 8 + e] = vf[e];
      }
    }
    __syncthreads();
    float sc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[i][j] = 0.f;
    for (int d = 0; d < D; ++d) {
      float qa[4], kb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qa[i] = Qt[d * (AQ + 1) + ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) kb[j] = Kt[d * (AK + 1) + tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sc[i][j] = fmaf(qa[i], kb[j], sc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + ty * 4 + i;
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kj = k0 + tx * 4 + j;
        const int rel = kj - qi;
        if (rel >= -window && rel <= window) sc[i][j] += Qrel[(ty * 4 + i) * R + rel + window];
        if (kj >= len) sc[i][j] = -CUDART_INF_F;
        mx = fmaxf(mx, sc[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[i], mx);
      const float corr = (m_run[i] == -CUDART_INF_F) ? 0.f : expf(m_run[i] - m_new);
      float psum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pv = (sc[i][j] == -CUDART_INF_F) ? 0.f : expf(sc[i][j] - m_new);
        Ps[(ty * 4 + i) * (AK + 1) + tx * 4 + j] = pv;
        psum += pv;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, off);
      l_run[i] = l_run[i] * corr + psum;
      m_run[i] = m_new;
#pragma unroll
      for (int c = 0; c < DC; ++c) o[i][c] *= corr;
    }
    __syncthreads();
    for (int j = 0; j < AK; ++j) {
      float pv[4], vv[DC];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = Ps[(ty * 4 + i) * (AK + 1) + j];
#pragma unroll
      for (int c = 0; c < DC; ++c) vv[c] = Vs[j * D + tx * DC + c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < DC; ++c) o[i][c] = fmaf(pv[i], vv[c], o[i][c]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + ty * 4 + i;
      for (int rr = 0; rr < R; ++rr) {
        const int kj = qi + rr - window;
        if (kj >= k0 && kj < k0 + AK && kj < len && kj >= 0) {
          const float pv = Ps[(ty * 4 + i) * (AK + 1) + (kj - k0)];
#pragma unroll
          for (int c = 0; c < DC; ++c) o[i][c] = fmaf(pv, Ev[rr * D + tx * DC + c], o[i][c]);
        }
      }
    }
  }
  // stage the normalised output tile in smem (reuse Ps/Vs region is busy; use Kt: [AQ][D] fits in D*(AK+1))
  __syncthreads();
  float* Os = Kt;  // [AQ][D]
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float inv = 1.0f / l_run[i];
#pragma unroll
    for (int c = 0; c < DC; ++c) Os[(ty * 4 + i) * D + tx * DC + c] = o[i][c] * inv;
  }
  __syncthreads();
  for (int i = tid; i < AQ * DP; i += 256) {
    const int r = i % AQ, pl = i / AQ;
    if (q0 + r >= len) continue;
    uint4 u;
    __half2* uh = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) uh[e] = __floats2half2_rn(Os[r * D + pl * 8 + 2 * e], Os[r * D + pl * 8 + 2 * e + 1]);
    *reinterpret_cast<uint4*>(out + (size_t)(h * DP + pl) * s.plane_stride + (pbase + q0 + r) * 8) = u;
  }
}

}  // namespace

#define POST_LAUNCH(ctx)            \
  do {                              \
    CUDA_CHECK(cudaGetLastError()); \
    (ctx).count();                  \
  } while (0)

void launch_flow_mix(const LaunchCtx& ctx, float* h_out, __half* hp, const float* h_in, const float* src32, const float* vec,
                     int vec_ld, int C, const PlanarSegs& s) {
  if (s.n <= 0 || s.max_len <= 0) return;
  if (C % 8 != 0 || C / 8 > 32) fail(SBV2_ERR_UNSUPPORTED, "flow_mix: C must be a multiple of 8 and <= 256");
  dim3 grid((s.max_len + 7) / 8, s.n);
  flow_mix_kernel<<<grid, 256, 0, ctx.stream>>>(h_out, hp, h_in, src32, vec, vec_ld, C, s);
  POST_LAUNCH(ctx);
}

void launch_ln_planar(const LaunchCtx& ctx, float* h, __half* hp, const float* y32, const float* gamma, const float* beta, float eps,
                      int C, const PlanarSegs& s) {
  if (s.n <= 0 || s.max_len <= 0) return;
  if (C % 8 != 0 || C / 8 > 32) fail(SBV2_ERR_UNSUPPORTED, "ln_planar: C must be a multiple of 8 and <= 256");
  // few rows (batch-1 latency): one row per warp spreads over more SMs; otherwise four rows per warp for loads in flight
  static const int force_r = [] { const char* e = getenv("SBV2_B200_LN_ROWS"); return e ? atoi(e) : 0; }();  // test hook
  if (force_r == 1 || (force_r == 0 && (long long)s.max_len * s.n <= 4096)) {
    dim3 grid((s.max_len + 7) / 8, s.n);
    ln_planar_kernel<1><<<grid, 256, 0, ctx.stream>>>(h, hp, y32, gamma, beta, eps, C, s);
  } else {
    dim3 grid((s.max_len + 31) / 32, s.n);
    ln_planar_kernel<4><<<grid, 256, 0, ctx.stream>>>(h, hp, y32, gamma, beta, eps, C, s);
  }
  POST_LAUNCH(ctx);
}

void launch_coupling_sub_planar(const LaunchCtx& ctx, float* z, const float* m32, int C, const PlanarSegs& s) {
  if (s.n <= 0 || s.max_len <= 0) return;
  dim3 grid((s.max_len + 7) / 8, s.n);
  coupling_sub_planar_kernel<<<grid, 256, 0, ctx.stream>>>(z, m32, C, s);
  POST_LAUNCH(ctx);
}

void launch_rel_attention_planar(const LaunchCtx& ctx, __half* ctx_out, const __half* qkv, const float* rel_k, const float* rel_v,
                                 int heads, int head_dim, int window, const PlanarSegs& s) {
  if (s.n <= 0 || s.max_len <= 0) return;
  if (head_dim != 96) fail(SBV2_ERR_UNSUPPORTED, "rel_attention: head_dim must be 96");
  constexpr int D = 96;
  const int R = 2 * window + 1;
  size_t smem = sizeof(float) * (size_t)(D * (AQ + 1) + D * (AK + 1) + AK * D + AQ * (AK + 1) + 2 * R * D + AQ * R);
  static PerDeviceOnce attr_once;
  attr_once.run([&] { CUDA_CHECK(cudaFuncSetAttribute(rel_attention_planar_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); });
  dim3 grid((s.max_len + AQ - 1) / AQ, heads, s.n);
  rel_attention_planar_kernel<D><<<grid, 256, smem, ctx.stream>>>(ctx_out, qkv, rel_k, rel_v, heads, window, s);
  POST_LAUNCH(ctx);
}

}  // namespace sbv2
