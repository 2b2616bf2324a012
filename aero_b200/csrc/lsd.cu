// Log-spectral distance on the GPU (reference src/metrics.py:37-70, `get_lsd` with STFTMag(2048, 512)):
//   LSD = mean_{b,t} sqrt( mean_f ( log10 max(|S_ref|^2, 1e-8) - log10 max(|S_est|^2, 1e-8) )^2 )
// Inputs are the *normalized* spectrograms produced by aero_stft_fwd (x n_fft^-1/2), so |S|^2 = n_fft |z|^2.
// One CTA per (b, t) column; fp64 accumulation of the per-column distances into out[0] (sum) -- the caller divides
// by B * frames.  HBM-bound: reads both spectrograms once.
#include "common.cuh"

namespace aero {

__global__ void __launch_bounds__(256) lsd_kernel(const float2* __restrict__ zr, const float2* __restrict__ ze,
                                                  double* __restrict__ out, int bins, int frames, float n_fft) {
    const int t = blockIdx.x, b = blockIdx.y;
    const float2* pr = zr + ((int64_t)b * bins) * frames + t;
    const float2* pe = ze + ((int64_t)b * bins) * frames + t;
    float acc = 0.f;
    for (int f = threadIdx.x; f < bins; f += 256) {
        const float2 a = pr[(int64_t)f * frames], c = pe[(int64_t)f * frames];
        const float sp = log10f(fmaxf(n_fft * (a.x * a.x + a.y * a.y), 1e-8f));
        const float st = log10f(fmaxf(n_fft * (c.x * c.x + c.y * c.y), 1e-8f));
        acc += (sp - st) * (sp - st);
    }
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        atomicAdd(out, (double)sqrtf(s / (float)bins));
    }
}

}  // namespace aero

extern "C" int aero_lsd_fwd(const float* z_ref, const float* z_est, double* out_sum, int32_t B, int32_t bins, int32_t frames,
                            int32_t n_fft, aero_stream_t stream) {
    using namespace aero;
    AERO_REQUIRE(z_ref && z_est && out_sum && B >= 1 && bins >= 1 && frames >= 1 && n_fft >= 2, "aero_lsd_fwd: bad argument");
    AERO_REQUIRE(B <= 65535, "aero_lsd_fwd: batch too large");
    dim3 grid(frames, B);
    lsd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(z_ref), reinterpret_cast<const float2*>(z_est),
                                                      out_sum, bins, frames, (float)n_fft);
    return check_launch("aero_lsd_fwd");
}
