#include "model.h"
namespace sbv2 {
sbv2_model* create_bert_model(const OnnxModel&, int) { fail(SBV2_ERR_UNSUPPORTED, "bert not built yet"); }
void bert_predict(sbv2_model*, const int64_t*, const int64_t*, int, int64_t, float*) { fail(SBV2_ERR_UNSUPPORTED, "bert not built yet"); }
int bert_hidden(const sbv2_model*) { return 0; }
}  // namespace sbv2
