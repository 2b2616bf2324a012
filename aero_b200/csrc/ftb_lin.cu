// FTB output conv evaluated through a linear (1x1-conv) input -- see include/aero_b200.h, aero_ftb_lin_out_fwd.
//
// In encoder layer 0 the FTB block (reference modules.py:304-325) is fed by `pre_conv`, a 1x1 convolution of the J = 2*C_in
// spectrogram channels (aero.py:89,112).  Everything FTB does before its last ReLU is linear in that input, so the C-channel
// tensors x = pre_conv(z), freq_fc(x * gate) and cat([.., x]) need never exist: with zm = freq_fc applied to z itself,
//   out[b,f,t,n] = relu( sum_j M[b,t][n][j] zm[b,f,t,j] + M[b,t][n][J] s[f] + sum_j V[n][j] z[b,f,t,j] + d[n] )
// where M[b,t] = gate[b,t,:] . Q is a tiny GEMM.  This kernel is that last line: it reads 2J floats per pixel and writes the
// C-channel output once (HBM-bound on the write), instead of three passes over C-channel tensors.
//
// Thread = (frame t, 8 output channels); its 8 x (J+1) slice of M[b,t], V and d stay in registers while it walks down the
// frequency rows.  Consecutive threads write consecutive 16 / 32-byte pieces of one (b, f) row.
#include "common.cuh"

namespace aero {

constexpr int kFtbTT = 32;      // frames per CTA

template <int J, typename TO>
__global__ void __launch_bounds__(512) ftb_lin_out_kernel(const float* __restrict__ z, const float* __restrict__ zm,
                                                          const float* __restrict__ M, const float* __restrict__ s,
                                                          const float* __restrict__ V, const float* __restrict__ d,
                                                          TO* __restrict__ out, const aero_ftb_lin_params p) {
    const int n8 = p.N >> 3;
    const int tl = threadIdx.x / n8, oc = threadIdx.x - tl * n8;
    const int t = blockIdx.x * kFtbTT + tl, b = blockIdx.y;
    if (t >= p.T) return;
    const int n0 = oc * 8;
    float m[8][J + 1], v[8][J], dd[8];
    const float* Mp = M + ((int64_t)b * p.T + t) * p.N * (J + 1) + (int64_t)n0 * (J + 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j <= J; ++j) m[i][j] = Mp[i * (J + 1) + j];
#pragma unroll
        for (int j = 0; j < J; ++j) v[i][j] = V[(n0 + i) * J + j];
        dd[i] = d[n0 + i];
    }
    const float* zp = z + (int64_t)b * p.z_sb + (int64_t)t * J;
    const float* zmp = zm + (int64_t)b * p.zm_sb + (int64_t)t * J;
    TO* op = out + (((int64_t)b * p.F) * p.T + t) * p.N + n0;
    const int64_t ostep = (int64_t)p.T * p.N;
    for (int f = 0; f < p.F; ++f) {
        float a[J], am[J];
#pragma unroll
        for (int j = 0; j < J; j += 2) {
            const float2 q = *reinterpret_cast<const float2*>(zp + j);
            const float2 qm = *reinterpret_cast<const float2*>(zmp + j);
            a[j] = q.x; a[j + 1] = q.y; am[j] = qm.x; am[j + 1] = qm.y;
        }
        const float sf = __ldg(s + f);
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float x = fmaf(m[i][J], sf, dd[i]);
#pragma unroll
            for (int j = 0; j < J; ++j) x = fmaf(m[i][j], am[j], fmaf(v[i][j], a[j], x));
            o[i] = fmaxf(x, 0.f);
            if (sizeof(TO) == 4 && (p.flags & AERO_TG_ROUND_TF32)) o[i] = round_tf32_rna(o[i]);
        }
        st4(op, make_float4(o[0], o[1], o[2], o[3]));
        st4(op + 4, make_float4(o[4], o[5], o[6], o[7]));
        zp += p.z_sf; zmp += p.zm_sf; op += ostep;
    }
}

}  // namespace aero

extern "C" int aero_ftb_lin_out_fwd(const float* z, const float* zm, const float* M, const float* s, const float* V,
                                    const float* d, void* out, const aero_ftb_lin_params* p, aero_stream_t stream) {
    using namespace aero;
    AERO_REQUIRE(z && zm && M && s && V && d && out && p, "aero_ftb_lin_out_fwd: null argument");
    AERO_REQUIRE(p->B >= 1 && p->F >= 1 && p->T >= 1 && p->N >= 8 && p->N % 8 == 0 && p->N <= 128, "aero_ftb_lin_out_fwd: N=%d (multiple of 8, at most 128)", p->N);
    AERO_REQUIRE(p->J == 2 || p->J == 4, "aero_ftb_lin_out_fwd: J=%d (2 or 4 input channels)", p->J);
    AERO_REQUIRE(p->z_sf % 2 == 0 && p->z_sb % 2 == 0 && p->zm_sf % 2 == 0 && p->zm_sb % 2 == 0 &&
                     (((uintptr_t)z | (uintptr_t)zm) & 7) == 0 && ((uintptr_t)out & 15) == 0,
                 "aero_ftb_lin_out_fwd: alignment");
    dim3 grid(cdiv(p->T, kFtbTT), p->B);
    const int threads = kFtbTT * (p->N / 8);
    cudaStream_t st = (cudaStream_t)stream;
    const bool o16 = p->flags & AERO_TG_OUT_F16;
#define AERO_FL(JJ)                                                                                                         \
    if (o16) ftb_lin_out_kernel<JJ, __half><<<grid, threads, 0, st>>>(z, zm, M, s, V, d, static_cast<__half*>(out), *p);     \
    else ftb_lin_out_kernel<JJ, float><<<grid, threads, 0, st>>>(z, zm, M, s, V, d, static_cast<float*>(out), *p)
    if (p->J == 2) { AERO_FL(2); } else { AERO_FL(4); }
#undef AERO_FL
    return check_launch("aero_ftb_lin_out_fwd");
}
