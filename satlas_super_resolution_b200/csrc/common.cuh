// Shared host-side helpers of libssr_b200: error reporting, launch counting, tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <cuda_bf16.h>

#include "../../include/ssr_b200.h"

namespace ssr {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// optional per-launch CUDA-event timing by kernel class (0 = conv_tc, 1 = wgrad_tc), see ssr_profile_start/stop
void prof_before(int cls, cudaStream_t s);
void prof_after(cudaStream_t s);

// returns false (and sets the error text) when the launch / previous call failed
bool check_cuda(cudaError_t e, const char* what);
bool check_last(const char* what);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
// elem_strides (optional, rank entries): traversal stride per dimension -- with stride s only every s-th element inside the
// bounding box `box` is copied (box[i] / s elements land in shared memory): how a stride-2 convolution gathers its taps.
bool encode_tmap_tiled(CUtensorMap* out, CUtensorMapDataType dtype, uint32_t rank, const void* gaddr,
                       const uint64_t* dims, const uint64_t* strides_bytes /* rank-1 */,
                       const uint32_t* box, CUtensorMapSwizzle swizzle, const uint32_t* elem_strides = nullptr);

// true unless SSR_PDL=0: tensor-core kernels are launched with programmatic stream serialization
bool pdl_enabled();

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace ssr

#define SSR_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ssr::set_error(__VA_ARGS__);    \
      return SSR_E_ARG;               \
    }                                 \
  } while (0)
