// cfgpp_b200 — host-side helpers: error handling, TMA tensor-map encoding (driver entry point fetched at
// run time so the library does not link libcuda), device properties.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>
#include <utility>

namespace cfgpp {

struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CFGPP_CHECK_CUDA(expr)                                                                     \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      throw ::cfgpp::Error(-2, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                                   __FILE__ + ":" + std::to_string(__LINE__));                     \
  } while (0)

#define CFGPP_REQUIRE(cond, msg)                                                                          \
  do {                                                                                                    \
    if (!(cond))                                                                                          \
      throw ::cfgpp::Error(-1, std::string("requirement failed: ") + #cond + " — " + (msg) + " at " +     \
                                   __FILE__ + ":" + std::to_string(__LINE__));                            \
  } while (0)

int num_sms();
bool pdl_enabled();  // false when CFGPP_NO_PDL=1 is set in the environment (A/B switch)

// Launch with programmatic dependent launch enabled (see common.cuh). Capturable into CUDA graphs.
template <typename... KArgs, typename... Args>
inline void launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                               int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  CFGPP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...));
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
  launch_pdl_cluster(kernel, grid, block, smem, stream, 1, std::forward<Args>(args)...);
}

// Generic fp16 tiled tensor map, 128B swizzle. dims/strides innermost first; strides[i] is the byte
// stride of dim i+1 (dim 0 is contiguous). OOB elements are zero-filled by the hardware.
// `elem_strides` (optional, per dim): traversal stride — with stride s a box extent b loads ceil(b / s) elements
// (every s-th one from the box origin); used by the stride-2 convolution.
CUtensorMap make_tmap_f16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box, int swizzle_bytes = 128, const uint32_t* elem_strides = nullptr);

// 2D row-major [rows][cols] fp16 with leading dimension ld (elements); box = (64 cols, box_rows).
CUtensorMap make_tmap_2d(const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
// epilogue tiles: box = (32 cols = 64 B, box_rows), 64B swizzle (output stores / residual loads)
CUtensorMap make_tmap_2d_sw64(const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

}  // namespace cfgpp
