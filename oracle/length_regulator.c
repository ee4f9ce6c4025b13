/* CPU restatement (plain C) of the integer part of the synthesizer graph — TEST INFRASTRUCTURE ONLY.
 *
 * The graph (/root/reference/scripts/convert/convert_model.py:97-110 traces upstream
 * SynthesizerTrn.infer) computes, per utterance:
 *     w      = exp(logw) * x_mask * length_scale
 *     w_ceil = ceil(w)
 *     y_len  = max(1, sum(w_ceil))
 *     attn   = generate_path(w_ceil, mask)      one-hot [T_y, T_x]
 * generate_path builds cum = cumsum(w_ceil) and marks frame j as owned by phoneme i iff
 * cum[i-1] <= j < cum[i].  This file restates that with integers only (durations, inclusive scan,
 * upper_bound), independently of oracle/vits.py (numpy) and of the CUDA kernels
 * (durations_kernel / expand_kernel), and is compared with both in tests/test_oracle.py and
 * tests/test_gpu_parity.py.  PARITY UNPINNED vs the reference's runtime (see oracle/vits.py).
 */
#include <math.h>
#include <stdint.h>

/* w: float32 [t_x] (already exp(logw) * length_scale).  durations: int32 [t_x].
 * Returns T_y = max(1, sum durations). */
int32_t sbv2_oracle_durations(const float* w, int32_t t_x, int32_t* durations) {
  int64_t total = 0;
  for (int32_t i = 0; i < t_x; ++i) {
    float c = ceilf(w[i]);
    if (c < 0.0f) c = 0.0f;
    durations[i] = (int32_t)c;
    total += durations[i];
  }
  return total > 1 ? (int32_t)total : 1;
}

/* frame2ph: int32 [t_y]; -1 for frames beyond the sum of durations (only when the sum is 0). */
void sbv2_oracle_frame2ph(const int32_t* durations, int32_t t_x, int32_t t_y, int32_t* frame2ph) {
  int32_t j = 0;
  for (int32_t i = 0; i < t_x && j < t_y; ++i)
    for (int32_t r = 0; r < durations[i] && j < t_y; ++r) frame2ph[j++] = i;
  while (j < t_y) frame2ph[j++] = -1;
}

/* m_p / logs_p gather + prior sample for one channel-major utterance, fp32, same operation order
 * as the graph: z_p = m + eps * exp(logs) * noise_scale.
 * stats: [t_x, 2*c] (m | logs) time-major; eps, z_p: [t_y, c] time-major. */
void sbv2_oracle_expand(const float* stats, const int32_t* frame2ph, const float* eps, float noise_scale, int32_t t_y,
                        int32_t c, float* z_p) {
  for (int32_t j = 0; j < t_y; ++j) {
    const int32_t i = frame2ph[j];
    for (int32_t k = 0; k < c; ++k) {
      const float m = i >= 0 ? stats[(int64_t)i * 2 * c + k] : 0.0f;
      const float logs = i >= 0 ? stats[(int64_t)i * 2 * c + c + k] : 0.0f;
      z_p[(int64_t)j * c + k] = m + (eps[(int64_t)j * c + k] * expf(logs)) * noise_scale;
    }
  }
}
