// Training form of the LocalState attention core (reference modules.py:104-124 under autograd): the exact-fp32 forward of
// attention.cu that also returns the per-query log-sum-exp, and its backward as two flash-style passes that recompute the
// scores (no T x T tensor in HBM):
//   pass Q (one thread per query s): dq[s], d(decay logits)[s]   -- keys / values streamed through shared memory
//   pass K (one thread per key t)  : dk[t], dv[t]                -- queries / output gradients streamed through shared memory
// scores  sc[t,s] = k_t . q_s / sqrt(d) - |t - s| * slope_s  (diagonal: the constant -100), softmax over t,  out[s] = sum_t w[t,s] v_t.
// d sc[t,s] = w[t,s] * (v_t . dout_s - out_s . dout_s).
#include "common.cuh"

namespace aero {

constexpr int kTQB = 128;     // queries (or keys) per CTA, one per thread
constexpr int kTKT = 128;     // streamed tile

__device__ __forceinline__ float attn_slope(const float* row, const aero_attn_params& p, int h) {
    float slope = 0.f;
    for (int f = 0; f < p.ndecay; ++f) slope += (float)(f + 1) * 0.5f * sigmoid_f(row[3 * p.H + h * p.ndecay + f]);
    return slope * rsqrtf((float)p.ndecay);
}

template <int D>
__global__ void __launch_bounds__(kTQB) attn_train_fwd_kernel(const float* __restrict__ qkvd, float* __restrict__ out,
                                                              float* __restrict__ lse, const aero_attn_params p) {
    __shared__ __align__(16) float Ks[kTKT * D];
    __shared__ __align__(16) float Vs[kTKT * D];
    const int row = blockIdx.z, h = blockIdx.y;
    const int s = blockIdx.x * kTQB + threadIdx.x;
    const bool valid = s < p.T;
    const int sq = valid ? s : p.T - 1;
    const float* base = qkvd + (int64_t)row * p.T * p.ld;
    float q[D];
    const float inv = rsqrtf((float)D);
#pragma unroll
    for (int c = 0; c < D; ++c) q[c] = base[(int64_t)sq * p.ld + h * D + c] * inv;
    const float slope = attn_slope(base + (int64_t)sq * p.ld, p, h);
    float m = -1e30f, l = 0.f, acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.f;
    for (int k0 = 0; k0 < p.T; k0 += kTKT) {
        const int nk = min(kTKT, p.T - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < nk * D; i += kTQB) {
            const int t = i / D, c = i - t * D;
            const float* src = base + (int64_t)(k0 + t) * p.ld + h * D + c;
            Ks[i] = src[p.H];
            Vs[i] = src[2 * p.H];
        }
        __syncthreads();
        for (int t = 0; t < nk; ++t) {
            float d = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) d = fmaf(q[c], Ks[t * D + c], d);
            const int ta = k0 + t;
            d -= fabsf((float)(ta - sq)) * slope;
            if (ta == sq) d = -100.0f;
            const float mn = fmaxf(m, d);
            const float corr = __expf(m - mn), pr = __expf(d - mn);
            l = l * corr + pr;
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] = fmaf(pr, Vs[t * D + c], acc[c] * corr);
            m = mn;
        }
    }
    if (valid) {
        const float il = 1.0f / l;
        float* o = out + ((int64_t)row * p.T + s) * p.H + h * D;
#pragma unroll
        for (int c = 0; c < D; ++c) o[c] = acc[c] * il;
        lse[((int64_t)row * p.heads + h) * p.T + s] = m + logf(l);
    }
}

template <int D>
__global__ void __launch_bounds__(kTQB) attn_bwd_q_kernel(const float* __restrict__ qkvd, const float* __restrict__ out,
                                                          const float* __restrict__ lse, const float* __restrict__ dout,
                                                          float* __restrict__ dqkvd, const aero_attn_params p) {
    __shared__ __align__(16) float Ks[kTKT * D];
    __shared__ __align__(16) float Vs[kTKT * D];
    const int row = blockIdx.z, h = blockIdx.y;
    const int s = blockIdx.x * kTQB + threadIdx.x;
    const bool valid = s < p.T;
    const int sq = valid ? s : p.T - 1;
    const float* base = qkvd + (int64_t)row * p.T * p.ld;
    const float inv = rsqrtf((float)D);
    float q[D], dq[D], dvec[D];
    float delta = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        q[c] = base[(int64_t)sq * p.ld + h * D + c] * inv;
        dq[c] = 0.f;
        dvec[c] = dout[((int64_t)row * p.T + sq) * p.H + h * D + c];
        delta = fmaf(out[((int64_t)row * p.T + sq) * p.H + h * D + c], dvec[c], delta);
    }
    const float slope = attn_slope(base + (int64_t)sq * p.ld, p, h);
    const float ls = lse[((int64_t)row * p.heads + h) * p.T + sq];
    float dslope = 0.f;
    for (int k0 = 0; k0 < p.T; k0 += kTKT) {
        const int nk = min(kTKT, p.T - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < nk * D; i += kTQB) {
            const int t = i / D, c = i - t * D;
            const float* src = base + (int64_t)(k0 + t) * p.ld + h * D + c;
            Ks[i] = src[p.H];
            Vs[i] = src[2 * p.H];
        }
        __syncthreads();
        for (int t = 0; t < nk; ++t) {
            const int ta = k0 + t;
            float d = 0.f, dw = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                d = fmaf(q[c], Ks[t * D + c], d);
                dw = fmaf(dvec[c], Vs[t * D + c], dw);
            }
            const float dist = fabsf((float)(ta - sq));
            d -= dist * slope;
            if (ta == sq) d = -100.0f;
            const float w = __expf(d - ls);
            const float dsc = (ta == sq) ? 0.f : w * (dw - delta);       // the diagonal score is a constant
#pragma unroll
            for (int c = 0; c < D; ++c) dq[c] = fmaf(dsc, Ks[t * D + c], dq[c]);
            dslope = fmaf(-dsc, dist, dslope);
        }
    }
    if (valid) {
        float* dr = dqkvd + ((int64_t)row * p.T + s) * p.ld;
#pragma unroll
        for (int c = 0; c < D; ++c) dr[h * D + c] = dq[c] * inv;
        const float* xr = base + (int64_t)s * p.ld + 3 * p.H + h * p.ndecay;
        const float k = rsqrtf((float)p.ndecay);
        for (int f = 0; f < p.ndecay; ++f) {
            const float sg = sigmoid_f(xr[f]);
            dr[3 * p.H + h * p.ndecay + f] = dslope * (float)(f + 1) * 0.5f * sg * (1.0f - sg) * k;
        }
    }
}

template <int D>
__global__ void __launch_bounds__(kTQB) attn_bwd_k_kernel(const float* __restrict__ qkvd, const float* __restrict__ out,
                                                          const float* __restrict__ lse, const float* __restrict__ dout,
                                                          float* __restrict__ dqkvd, const aero_attn_params p) {
    __shared__ __align__(16) float Qs[kTKT * D];
    __shared__ __align__(16) float Ds[kTKT * D];
    __shared__ float aux[kTKT][3];                 // lse, delta, slope of the tile's queries
    const int row = blockIdx.z, h = blockIdx.y;
    const int t = blockIdx.x * kTQB + threadIdx.x;
    const bool valid = t < p.T;
    const int tk = valid ? t : p.T - 1;
    const float* base = qkvd + (int64_t)row * p.T * p.ld;
    const float inv = rsqrtf((float)D);
    float kv[D], vv[D], dk[D], dv[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        kv[c] = base[(int64_t)tk * p.ld + p.H + h * D + c];
        vv[c] = base[(int64_t)tk * p.ld + 2 * p.H + h * D + c];
        dk[c] = 0.f;
        dv[c] = 0.f;
    }
    for (int s0 = 0; s0 < p.T; s0 += kTKT) {
        const int ns = min(kTKT, p.T - s0);
        __syncthreads();
        for (int i = threadIdx.x; i < ns * D; i += kTQB) {
            const int s = i / D, c = i - s * D;
            Qs[i] = base[(int64_t)(s0 + s) * p.ld + h * D + c] * inv;
            Ds[i] = dout[((int64_t)row * p.T + s0 + s) * p.H + h * D + c];
        }
        if (threadIdx.x < ns) {
            const int s = s0 + threadIdx.x;
            float delta = 0.f;
            for (int c = 0; c < D; ++c)
                delta = fmaf(out[((int64_t)row * p.T + s) * p.H + h * D + c], dout[((int64_t)row * p.T + s) * p.H + h * D + c], delta);
            aux[threadIdx.x][0] = lse[((int64_t)row * p.heads + h) * p.T + s];
            aux[threadIdx.x][1] = delta;
            aux[threadIdx.x][2] = attn_slope(base + (int64_t)s * p.ld, p, h);
        }
        __syncthreads();
        for (int s = 0; s < ns; ++s) {
            const int sa = s0 + s;
            float d = 0.f, dw = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                d = fmaf(Qs[s * D + c], kv[c], d);
                dw = fmaf(Ds[s * D + c], vv[c], dw);
            }
            d -= fabsf((float)(tk - sa)) * aux[s][2];
            if (sa == tk) d = -100.0f;
            const float w = __expf(d - aux[s][0]);
            const float dsc = (sa == tk) ? 0.f : w * (dw - aux[s][1]);
#pragma unroll
            for (int c = 0; c < D; ++c) {
                dv[c] = fmaf(w, Ds[s * D + c], dv[c]);
                dk[c] = fmaf(dsc, Qs[s * D + c], dk[c]);       // Qs already carries 1/sqrt(d)
            }
        }
    }
    if (valid) {
        float* dr = dqkvd + ((int64_t)row * p.T + t) * p.ld;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            dr[p.H + h * D + c] = dk[c];
            dr[2 * p.H + h * D + c] = dv[c];
        }
    }
}

template <int D>
static int attn_train_fwd_go(const float* qkvd, float* out, float* lse, const aero_attn_params& p, cudaStream_t st) {
    dim3 grid(cdiv(p.T, kTQB), p.heads, p.rows);
    attn_train_fwd_kernel<D><<<grid, kTQB, 0, st>>>(qkvd, out, lse, p);
    return check_launch("aero_local_attn_train_fwd");
}
template <int D>
static int attn_bwd_go(const float* qkvd, const float* out, const float* lse, const float* dout, float* dqkvd, const aero_attn_params& p,
                       cudaStream_t st) {
    dim3 grid(cdiv(p.T, kTQB), p.heads, p.rows);
    attn_bwd_q_kernel<D><<<grid, kTQB, 0, st>>>(qkvd, out, lse, dout, dqkvd, p);
    int rc = check_launch("aero_local_attn_bwd(q)");
    if (rc != AERO_OK) return rc;
    attn_bwd_k_kernel<D><<<grid, kTQB, 0, st>>>(qkvd, out, lse, dout, dqkvd, p);
    return check_launch("aero_local_attn_bwd(k)");
}

static int attn_train_check(const aero_attn_params* p) {
    AERO_REQUIRE(p->heads >= 1 && p->H % p->heads == 0 && p->ndecay >= 1 && p->ndecay <= 16, "aero_local_attn_train: heads/ndecay");
    AERO_REQUIRE(p->ld >= 3 * p->H + p->heads * p->ndecay, "aero_local_attn_train: ld=%d too small", p->ld);
    AERO_REQUIRE(p->rows >= 1 && p->rows <= 65535 && p->T >= 1, "aero_local_attn_train: rows=%d", p->rows);
    return AERO_OK;
}

}  // namespace aero

extern "C" int aero_local_attn_train_fwd(const float* qkvd, float* out, float* lse, const aero_attn_params* p, aero_stream_t stream) {
    using namespace aero;
    AERO_REQUIRE(qkvd && out && lse && p, "aero_local_attn_train_fwd: null argument");
    int rc = attn_train_check(p);
    if (rc != AERO_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (p->H / p->heads) {
        case 3: return attn_train_fwd_go<3>(qkvd, out, lse, *p, st);
        case 6: return attn_train_fwd_go<6>(qkvd, out, lse, *p, st);
        case 12: return attn_train_fwd_go<12>(qkvd, out, lse, *p, st);
        case 24: return attn_train_fwd_go<24>(qkvd, out, lse, *p, st);
        default: set_error("aero_local_attn_train_fwd: head dim %d not instantiated (3, 6, 12, 24)", p->H / p->heads); return AERO_ERR_UNSUPPORTED;
    }
}

extern "C" int aero_local_attn_bwd(const float* qkvd, const float* out, const float* lse, const float* dout, float* dqkvd,
                                   const aero_attn_params* p, aero_stream_t stream) {
    using namespace aero;
    AERO_REQUIRE(qkvd && out && lse && dout && dqkvd && p, "aero_local_attn_bwd: null argument");
    int rc = attn_train_check(p);
    if (rc != AERO_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (p->H / p->heads) {
        case 3: return attn_bwd_go<3>(qkvd, out, lse, dout, dqkvd, *p, st);
        case 6: return attn_bwd_go<6>(qkvd, out, lse, dout, dqkvd, *p, st);
        case 12: return attn_bwd_go<12>(qkvd, out, lse, dout, dqkvd, *p, st);
        case 24: return attn_bwd_go<24>(qkvd, out, lse, dout, dqkvd, *p, st);
        default: set_error("aero_local_attn_bwd: head dim %d not instantiated (3, 6, 12, 24)", p->H / p->heads); return AERO_ERR_UNSUPPORTED;
    }
}
