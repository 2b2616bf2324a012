// Shared helpers for libaero_b200 (sm_100a).  No torch / ATen types anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/aero_b200.h"

namespace aero {

void set_error(const char* fmt, ...);
int check_launch(const char* what);   // bumps the launch counter, returns AERO_OK / AERO_ERR_LAUNCH

#define AERO_REQUIRE(cond, ...)                         \
    do {                                                \
        if (!(cond)) {                                  \
            aero::set_error(__VA_ARGS__);               \
            return AERO_ERR_INVALID;                    \
        }                                               \
    } while (0)

__device__ __forceinline__ float gelu_exact(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// Round to TF32 (10-bit mantissa), nearest with ties away from zero -- the result of cvt.rna.tf32.f32 for every finite input,
// in two integer instructions (ptxas expands the cvt into ~5 with Inf/NaN special-casing; Inf and NaN also survive this form:
// their low 13 mantissa bits are simply cleared).  Matches aero_b200.engine.tf32_round on the host bit for bit.
__device__ __forceinline__ float round_tf32_rna(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace aero
