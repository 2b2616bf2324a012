// Shared argument block of the tap-GEMM implementations.
#pragma once
#include "common.cuh"

namespace aero {
struct TapGemmArgs {
    const float* a1;
    const float* a2;
    const float* w;
    const float* bias;
    const float* addend_fn;
    const float* colscale;
    const float* residual;
    const float* samp_affine;
    float* out;
    double* stats;
    aero_tapgemm_params p;
    int ntaps;
    int tiles_t;
    int ldw;          // weight row stride (N rounded up to 4)
    int vec_a;        // 16-byte loads of A are legal
    int vec_o;        // 16-byte stores legal
};
int tapgemm_simt_launch(const TapGemmArgs& g, cudaStream_t st);
int tapgemm_tc_launch(const TapGemmArgs& g, cudaStream_t st);
bool tapgemm_tc_eligible(const aero_tapgemm_params& p);
}  // namespace aero
