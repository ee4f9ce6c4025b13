#include "umma_conv.h"
namespace sbv2 {
struct UmmaDecoder {};
UmmaDecoder* umma_decoder_create(const DecoderHostWeights&, sbv2_model*) { return nullptr; }
void umma_decoder_free(UmmaDecoder* d) { delete d; }
void umma_decoder_run(UmmaDecoder*, sbv2_model*, const float*, const float*, int, const std::vector<int>&, const std::vector<int>&, float*) {
  fail(SBV2_ERR_INTERNAL, "tensor-core decoder not built");
}
}  // namespace sbv2
