// Tensor-core (tcgen05 / TMEM) implicit-GEMM convolution path of the HiFi-GAN decoder.
// Activations live in HBM as fp16 "planar" tiles [C/8][rows][8] so that any row-shifted window
// of any 8-channel plane is a dense run of 16-byte rows: exactly the no-swizzle K-major core-matrix
// layout tcgen05.mma reads, which lets one halo'd tile serve every filter tap through a shifted
// shared-memory descriptor.  See DESIGN.md §kernels.
#pragma once
#include <vector>

#include "common.h"

struct sbv2_model;

namespace sbv2 {

struct HostConv {
  std::vector<float> w, b;  // PyTorch layout: Conv1d [d0=Cout, d1=Cin, k]; ConvTranspose1d [d0=Cin, d1=Cout, k]
  int d0 = 0, d1 = 0, k = 1;
};

struct DecoderHostWeights {
  HostConv pre, cond, post;
  std::vector<HostConv> ups;
  std::vector<int> up_u;
  std::vector<std::vector<HostConv>> res_c1, res_c2;  // [resblock][layer]
  std::vector<std::vector<int>> res_dil;
  int per = 3;  // resblocks per stage
  int gin = 512;
};

struct UmmaDecoder;
// Returns nullptr when the decoder shape is outside what the tensor-core plan supports.
UmmaDecoder* umma_decoder_create(const DecoderHostWeights& w, sbv2_model* owner);
void umma_decoder_free(UmmaDecoder* d);
// z: packed [Ny, Cin] fp32 (time-major), g: [B, gin] fp32, wave: [Ny*hop] fp32 (all device).
void umma_decoder_run(UmmaDecoder* d, sbv2_model* owner, const float* z, const float* g, int B,
                      const std::vector<int>& ystart, const std::vector<int>& ylen, float* wave);

}  // namespace sbv2
