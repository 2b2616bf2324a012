// Multi-resolution STFT loss, forward reductions (reference src/models/stft_loss.py:11-27,30-63):
// for one resolution, with mag = sqrt(max(re^2 + im^2, 1e-7)) of the un-normalised STFTs of x (estimate) and y (target),
//   sums[0] += sum (mag_y - mag_x)^2      sums[1] += sum mag_y^2      sums[2] += sum |log mag_y - log mag_x|
// so that spectral convergence = sqrt(sums[0] / sums[1]) and the log-magnitude L1 = sums[2] / (B * frames * bins).
// Inputs are the *normalised* spectrograms written by aero_stft_fwd (x n_fft^-1/2): re^2 + im^2 = n_fft |z|^2.
// HBM-bound: both spectrograms are read once; fp32 per-thread partials, fp64 across the grid.
#include "common.cuh"

namespace aero {

__global__ void __launch_bounds__(256) stft_loss_kernel(const float2* __restrict__ zx, const float2* __restrict__ zy,
                                                        double* __restrict__ sums, int64_t n, float n_fft) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 a = zx[i], c = zy[i];
        const float px = fmaxf(n_fft * (a.x * a.x + a.y * a.y), 1e-7f), py = fmaxf(n_fft * (c.x * c.x + c.y * c.y), 1e-7f);
        const float mx = sqrtf(px), my = sqrtf(py);
        s0 += (my - mx) * (my - mx);
        s1 += py;
        s2 += 0.5f * fabsf(logf(py) - logf(px));                  // log mag = log(power) / 2
    }
    __shared__ float red[8][3];
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s0; red[threadIdx.x >> 5][1] = s1; red[threadIdx.x >> 5][2] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += (double)red[w][threadIdx.x];
        atomicAdd(sums + threadIdx.x, t);
    }
}

}  // namespace aero

extern "C" int aero_stft_loss_fwd(const float* z_est, const float* z_ref, double* sums, int32_t B, int32_t bins, int32_t frames,
                                  int32_t n_fft, aero_stream_t stream) {
    using namespace aero;
    AERO_REQUIRE(z_est && z_ref && sums && B >= 1 && bins >= 1 && frames >= 1 && n_fft >= 2, "aero_stft_loss_fwd: bad argument");
    const int64_t n = (int64_t)B * bins * frames;
    int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    stft_loss_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(z_est), reinterpret_cast<const float2*>(z_ref),
                                                              sums, n, (float)n_fft);
    return check_launch("aero_stft_loss_fwd");
}
