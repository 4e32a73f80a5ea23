// cfgpp_b200 — C ABI of the CLIP text towers (include/cfgpp_b200.h, "CLIP text encoder").
#include "capi_util.h"
#include "text_encoder.cuh"

using namespace cfgpp;

struct cfgpp_clip_handle {
  ClipTextEncoder enc;
  cfgpp_clip_handle(const cfgpp_clip_desc& d, int device) : enc(d, device) {}
};

extern "C" {

CFGPP_API int cfgpp_clip_create(const cfgpp_clip_desc* desc, int device, cfgpp_clip_handle** out) {
  return guarded([&] {
    CFGPP_REQUIRE(desc && out, "null argument");
    *out = new cfgpp_clip_handle(*desc, device);
  });
}

CFGPP_API int cfgpp_clip_destroy(cfgpp_clip_handle* h) {
  return guarded([&] { delete h; });
}

CFGPP_API int cfgpp_clip_load_weight(cfgpp_clip_handle* h, const char* key, const void* data, const int64_t* shape,
                                     int ndim, int dtype, void* stream) {
  return guarded([&] { h->enc.load_weight(key, data, shape, ndim, dtype, (cudaStream_t)stream); });
}

CFGPP_API int cfgpp_clip_finalize_weights(cfgpp_clip_handle* h, void* stream) {
  return guarded([&] { h->enc.finalize_weights((cudaStream_t)stream); });
}

CFGPP_API int cfgpp_clip_encode(cfgpp_clip_handle* h, const int32_t* input_ids, const int32_t* pooled_index, int batch,
                                int n_tokens, int skip, void* hidden_out, void* last_hidden_out, void* pooled_out,
                                void* stream) {
  return guarded([&] {
    h->enc.encode(input_ids, pooled_index, batch, n_tokens, skip, (__half*)hidden_out, (__half*)last_hidden_out,
                  (__half*)pooled_out, (cudaStream_t)stream);
  });
}

CFGPP_API int cfgpp_clip_stats(cfgpp_clip_handle* h, double* flops, size_t* workspace_bytes) {
  return guarded([&] {
    if (flops) *flops = h->enc.flops();
    if (workspace_bytes) *workspace_bytes = h->enc.workspace_bytes();
  });
}

}  // extern "C"
