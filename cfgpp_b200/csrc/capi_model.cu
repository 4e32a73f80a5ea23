// cfgpp_b200 — C ABI, model-level entry points (handle lifetime, weights, conditioning, forward, fused steps).
#include "capi_util.h"
#include "unet.cuh"

#include <algorithm>
#include <cstring>

using namespace cfgpp;

struct cfgpp_handle {
  Unet unet;
  cfgpp_handle(const cfgpp_model_desc& d, int device) : unet(d, device) {}
};

extern "C" {

CFGPP_API int cfgpp_create(const cfgpp_model_desc* desc, int device, cfgpp_handle** out) {
  return guarded([&] {
    CFGPP_REQUIRE(desc && out, "null argument");
    *out = new cfgpp_handle(*desc, device);
  });
}

CFGPP_API int cfgpp_destroy(cfgpp_handle* h) {
  return guarded([&] { delete h; });
}

CFGPP_API int cfgpp_load_weight(cfgpp_handle* h, const char* key, const void* data, const int64_t* shape, int ndim,
                                int dtype, void* stream) {
  return guarded([&] { h->unet.load_weight(key, data, shape, ndim, dtype, (cudaStream_t)stream); });
}

CFGPP_API int cfgpp_finalize_weights(cfgpp_handle* h, void* stream) {
  return guarded([&] { h->unet.finalize_weights((cudaStream_t)stream); });
}

CFGPP_API int cfgpp_prepare(cfgpp_handle* h, int batch, int h_lat, int w_lat) {
  return guarded([&] { h->unet.prepare(batch, h_lat, w_lat); });
}

CFGPP_API int cfgpp_workspace_bytes(cfgpp_handle* h, size_t* bytes) {
  return guarded([&] { *bytes = h->unet.workspace_bytes(); });
}

CFGPP_API int cfgpp_forward_flops(cfgpp_handle* h, double* flops) {
  return guarded([&] { *flops = h->unet.forward_flops(); });
}

CFGPP_API int cfgpp_launches_per_step(cfgpp_handle* h, int* n) {
  return guarded([&] { *n = h->unet.launches_per_step(); });
}

CFGPP_API int cfgpp_plan_stats(cfgpp_handle* h, double* step_flops, double* prompt_flops, int* prompt_launches) {
  return guarded([&] {
    if (step_flops) *step_flops = h->unet.forward_flops() - h->unet.prompt_flops();
    if (prompt_flops) *prompt_flops = h->unet.prompt_flops();
    if (prompt_launches) *prompt_launches = h->unet.prompt_launches();
  });
}

CFGPP_API int cfgpp_set_prompt(cfgpp_handle* h, const void* ctx, int n_ctx, const void* pooled, const float* time_ids,
                               int add_rows, void* stream) {
  return guarded([&] {
    h->unet.set_prompt((const __half*)ctx, n_ctx, (const __half*)pooled, time_ids, add_rows, (cudaStream_t)stream);
  });
}

CFGPP_API int cfgpp_unet_forward(cfgpp_handle* h, const void* z, int z_dtype, float t, float in_scale, void* eps_uc,
                                 void* eps_c, void* stream) {
  return guarded([&] {
    h->unet.unet_forward(z, z_dtype, t, in_scale, (__half*)eps_uc, (__half*)eps_c, (cudaStream_t)stream);
  });
}

CFGPP_API int cfgpp_set_schedule(cfgpp_handle* h, int method, int state_dtype, const cfgpp_step_state* steps,
                                 int nsteps, void* stream) {
  return guarded([&] { h->unet.set_schedule(method, state_dtype, steps, nsteps, (cudaStream_t)stream); });
}

CFGPP_API int cfgpp_set_state(cfgpp_handle* h, const void* z, int z_dtype, void* stream) {
  return guarded([&] { h->unet.set_state(z, z_dtype, (cudaStream_t)stream); });
}

CFGPP_API int cfgpp_set_noise(cfgpp_handle* h, const void* noise_dev, int slots, void* stream) {
  return guarded([&] { h->unet.set_noise((const __half*)noise_dev, slots, (cudaStream_t)stream); });
}

CFGPP_API int cfgpp_run_steps(cfgpp_handle* h, int first_step, int nsteps, void* stream) {
  return guarded([&] { h->unet.run_steps(first_step, nsteps, (cudaStream_t)stream); });
}

CFGPP_API int cfgpp_get_state(cfgpp_handle* h, int which, void* out, void* stream) {
  return guarded([&] { h->unet.get_state(which, out, (cudaStream_t)stream); });
}

CFGPP_API int cfgpp_apply_step(cfgpp_handle* h, int step, const void* eps_uc, const void* eps_c, void* stream) {
  return guarded([&] { h->unet.apply_step(step, (const __half*)eps_uc, (const __half*)eps_c, (cudaStream_t)stream); });
}

CFGPP_API int cfgpp_profile_forward(cfgpp_handle* h, const void* z, int z_dtype, float t, float in_scale, int max_n,
                                    int* n_out, float* ms_out, double* flops_out, int* kind_out, char* names_out,
                                    int name_stride, void* stream) {
  return guarded([&] {
    auto prof = h->unet.profile_forward(z, z_dtype, t, in_scale, (cudaStream_t)stream);
    const int n = static_cast<int>(prof.size()) < max_n ? static_cast<int>(prof.size()) : max_n;
    *n_out = n;
    for (int i = 0; i < n; ++i) {
      ms_out[i] = prof[i].ms;
      flops_out[i] = prof[i].flops;
      kind_out[i] = prof[i].kind;
      if (names_out && name_stride > 0) {
        const size_t len = std::min<size_t>(prof[i].name.size(), static_cast<size_t>(name_stride - 1));
        memcpy(names_out + static_cast<size_t>(i) * name_stride, prof[i].name.data(), len);
        names_out[static_cast<size_t>(i) * name_stride + len] = 0;
      }
    }
  });
}

}  // extern "C"
