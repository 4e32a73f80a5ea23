// cfgpp_b200 — C ABI of the AutoencoderKL decoder (include/cfgpp_b200.h, "AutoencoderKL decoder").
#include "capi_util.h"
#include "vae.cuh"

using namespace cfgpp;

struct cfgpp_vae_handle {
  VaeDecoder vae;
  cfgpp_vae_handle(const cfgpp_vae_desc& d, int device) : vae(d, device) {}
};

extern "C" {

CFGPP_API int cfgpp_vae_create(const cfgpp_vae_desc* desc, int device, cfgpp_vae_handle** out) {
  return guarded([&] {
    CFGPP_REQUIRE(desc && out, "null argument");
    *out = new cfgpp_vae_handle(*desc, device);
  });
}

CFGPP_API int cfgpp_vae_destroy(cfgpp_vae_handle* h) {
  return guarded([&] { delete h; });
}

CFGPP_API int cfgpp_vae_load_weight(cfgpp_vae_handle* h, const char* key, const void* data, const int64_t* shape,
                                    int ndim, int dtype, void* stream) {
  return guarded([&] { h->vae.load_weight(key, data, shape, ndim, dtype, (cudaStream_t)stream); });
}

CFGPP_API int cfgpp_vae_finalize_weights(cfgpp_vae_handle* h, void* stream) {
  return guarded([&] { h->vae.finalize_weights((cudaStream_t)stream); });
}

CFGPP_API int cfgpp_vae_decode(cfgpp_vae_handle* h, const void* z, int z_dtype, int batch, int h_lat, int w_lat,
                               void* image, void* stream) {
  return guarded([&] { h->vae.decode(z, z_dtype, batch, h_lat, w_lat, (__half*)image, (cudaStream_t)stream); });
}

CFGPP_API int cfgpp_vae_encode(cfgpp_vae_handle* h, const void* image, int image_dtype, int batch, int height, int width,
                               const void* noise, void* latent, void* stream) {
  return guarded([&] {
    h->vae.encode(image, image_dtype, batch, height, width, (const __half*)noise, (float*)latent, (cudaStream_t)stream);
  });
}

CFGPP_API int cfgpp_vae_stats(cfgpp_vae_handle* h, double* flops, size_t* workspace_bytes) {
  return guarded([&] {
    if (flops) *flops = h->vae.flops();
    if (workspace_bytes) *workspace_bytes = h->vae.workspace_bytes();
  });
}

}  // extern "C"
