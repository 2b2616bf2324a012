// Shared argument block of the tap-GEMM implementations.
#pragma once
#include "common.cuh"

namespace aero {
struct TapGemmArgs {
    const void* a1;           // fp32, or FP16 when p.flags & AERO_TG_A_F16
    const void* a2;
    const void* w;            // fp32 (precision 0 / 1) or FP16 (precision 2)
    const float* bias;
    const float* addend_fn;
    const float* colscale;
    const void* residual;     // same type as out
    const float* samp_affine;
    void* out;                // fp32, or FP16 when p.flags & AERO_TG_OUT_F16
    double* stats;
    aero_tapgemm_params p;
    int ntaps;
    int tiles_t;
    int ldw;          // weight row stride (N rounded up to 4)
    int vec_a;        // 16-byte loads of A are legal
    int vec_o;        // 16-byte stores legal
    int vec_o8;       // FP16 outputs: rows are 16-byte aligned in units of 8 halves (direct lane-per-row epilogue)
    // tcgen05 path: exact division of tile indices (< 2^31) by n_tiles, tiles_t, F_out:  q = (n * mul) >> shr
    uint32_t dv_mul[3], dv_shr[3];
    int direct_f16;   // FP16 outputs: same choice
    int direct_f32;   // fp32 outputs: direct lane-per-row epilogue instead of the shared-memory transpose
    int last_tile;    // tiles_total - 1 (reverse walk)
    int grouped_bn;   // tiles with BN <= this use the two-group epilogue
};
int tapgemm_simt_launch(const TapGemmArgs& g, cudaStream_t st);
int tapgemm_tc_launch(const TapGemmArgs& g, cudaStream_t st);
bool tapgemm_tc_eligible(const aero_tapgemm_params& p);
}  // namespace aero
