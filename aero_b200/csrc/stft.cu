// Fused STFT / iSTFT kernels for sm_100a.
//
// aero_stft_fwd  : reflect pad + framing + window + real FFT (N/2-point complex FFT in shared
//                  memory + split post-pass) + n_fft^-1/2 + Nyquist drop + strided (channels-last
//                  or planar-complex) store + per-sample {sum, sumsq}.
//                  Replaces torch.stft as called at reference src/models/spec.py:12-20.
// aero_istft_fwd : strided load + C2R FFT + window + overlap-add + 1/sum(w^2) + centre trim.
//                  Replaces torch.istft as called at reference src/models/spec.py:30-37.
//
// Both are HBM-bound (33.9 MB / 36.9 MB per B=32 forward, FFT flops negligible).
// A CTA owns a run of consecutive frames of one signal so that, for every frequency bin, the
// frames it writes (reads) are contiguous in memory: >=128 B segments for the model's layouts.
#include "common.cuh"

namespace aero {

constexpr int kThreads = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-place radix-2 DIT over `nfr` frames of M = 2^LOGM complex points, input in bit-reversed
// order, twiddles tw[j] = exp(-+2 pi i j / (2M)) (table over N = 2M), sign chosen by table.
template <int LOGM>
__device__ __forceinline__ void fft_inplace(float2* work, const float2* twN, int nfr) {
    constexpr int M = 1 << LOGM;
    const int total = nfr * (M / 2);
#pragma unroll 1
    for (int s = 0; s < LOGM; ++s) {
        const int half = 1 << s;
        const int tw_step = M >> s;          // N / (2*half) = 2M / (2*half)
        for (int i = threadIdx.x; i < total; i += kThreads) {
            const int fr = i / (M / 2), j = i - fr * (M / 2);
            const int pos = j & (half - 1);
            const int i0 = ((j >> s) << (s + 1)) + pos;
            float2* w = work + fr * M;
            const float2 a = w[i0];
            const float2 b = cmul(w[i0 + half], twN[pos * tw_step]);
            w[i0] = make_float2(a.x + b.x, a.y + b.y);
            w[i0 + half] = make_float2(a.x - b.x, a.y - b.y);
        }
        __syncthreads();
    }
}

template <int LOGN>
struct StftCfg {
    static constexpr int N = 1 << LOGN;
    static constexpr int M = N / 2;
    static constexpr int FB = (4096 / M) > 32 ? 32 : (4096 / M);       // frames per CTA (32 KB of work)
};

// ---------------------------------------------------------------------------------- forward
template <int LOGN>
__global__ void __launch_bounds__(kThreads) stft_kernel(const float* __restrict__ x, const float* __restrict__ window,
                                                        float* __restrict__ z, double* __restrict__ stats,
                                                        const aero_stft_params p) {
    using C = StftCfg<LOGN>;
    constexpr int N = C::N, M = C::M, FB = C::FB, LOGM = LOGN - 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* work = reinterpret_cast<float2*>(smem_raw);                 // [FB][M]
    float2* stage = work + FB * M;                                      // [M+1][FB]
    float2* twN = stage + (M + 1) * FB;                                 // [M]  exp(-2 pi i j / N)
    float* wpad = reinterpret_cast<float*>(twN + M);                    // [N]
    float* seg = wpad + N;                                              // [(FB-1)*hop + N]

    const int sig = blockIdx.y;
    const int t0 = blockIdx.x * FB;
    const int nfr = min(FB, p.frames - t0);
    const int L = p.length;

    for (int j = threadIdx.x; j < M; j += kThreads) {
        float s, c;
        sincospif(2.0f * (float)j / (float)N, &s, &c);
        twN[j] = make_float2(c, -s);
    }
    const int wl = (N - p.win) / 2;
    for (int n = threadIdx.x; n < N; n += kThreads) {
        const int k = n - wl;
        wpad[n] = (k >= 0 && k < p.win) ? window[k] : 0.0f;
    }
    const int seg_len = (nfr - 1) * p.hop + N;
    const float* xs = x + (int64_t)sig * L;
    const int q0 = t0 * p.hop - N / 2;
    for (int i = threadIdx.x; i < seg_len; i += kThreads) {
        int src = q0 + i;
        if (src < 0) src = -src;
        if (src >= L) src = 2 * (L - 1) - src;
        seg[i] = xs[src];
    }
    __syncthreads();

    // windowed frames, even/odd packed, bit-reversed placement
    for (int i = threadIdx.x; i < nfr * M; i += kThreads) {
        const int fr = i / M, n = i - fr * M;
        const float* s = seg + fr * p.hop + 2 * n;
        const int r = __brev((unsigned)n) >> (32 - LOGM);
        work[fr * M + r] = make_float2(s[0] * wpad[2 * n], s[1] * wpad[2 * n + 1]);
    }
    __syncthreads();
    fft_inplace<LOGM>(work, twN, nfr);

    // split post-pass: X[k] = Xe[k] + w^k Xo[k], X[M-k] = conj(Xe[k] - w^k Xo[k])
    const float scale = rsqrtf((float)N);
    for (int i = threadIdx.x; i < nfr * (M / 2 + 1); i += kThreads) {
        const int fr = i / (M / 2 + 1), k = i - fr * (M / 2 + 1);
        const float2 a = work[fr * M + k];
        const float2 bq = work[fr * M + ((M - k) & (M - 1))];
        const float2 xe = make_float2(0.5f * (a.x + bq.x), 0.5f * (a.y - bq.y));
        const float2 d = make_float2(0.5f * (a.x - bq.x), 0.5f * (a.y + bq.y));   // (Z[k]-conj(Z[M-k]))/2
        const float2 xo = make_float2(d.y, -d.x);                                  // * (-i)
        const float2 t = cmul(twN[k], xo);
        stage[k * FB + fr] = make_float2(scale * (xe.x + t.x), scale * (xe.y + t.y));
        stage[(M - k) * FB + fr] = make_float2(scale * (xe.x - t.x), -scale * (xe.y - t.y));
    }
    __syncthreads();

    float* zs = z + (int64_t)(sig / p.channels) * p.z_stride_b + (int64_t)(sig % p.channels) * p.z_stride_c;
    float lsum = 0.f, lsq = 0.f;
    for (int i = threadIdx.x; i < p.bins_out * nfr; i += kThreads) {
        const int k = i / nfr, fr = i - k * nfr;
        const float2 v = stage[k * FB + fr];
        *reinterpret_cast<float2*>(zs + (int64_t)k * p.z_stride_k + (int64_t)(t0 + fr) * p.z_stride_t) = v;
        lsum += v.x + v.y;
        lsq += v.x * v.x + v.y * v.y;
    }
    if (stats != nullptr) {
        __shared__ double red[2][kThreads / 32];
        double ds = warp_sum((double)lsum), dq = warp_sum((double)lsq);
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ds; red[1][threadIdx.x >> 5] = dq; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0, b = 0;
            for (int w = 0; w < kThreads / 32; ++w) { a += red[0][w]; b += red[1][w]; }
            atomicAdd(&stats[2 * (sig / p.channels)], a);
            atomicAdd(&stats[2 * (sig / p.channels) + 1], b);
        }
    }
}

template <int LOGN>
static int launch_stft(const float* x, const float* window, float* z, double* stats, const aero_stft_params& p,
                       cudaStream_t st) {
    using C = StftCfg<LOGN>;
    const size_t smem = sizeof(float2) * (C::FB * C::M + (C::M + 1) * C::FB + C::M) +
                        sizeof(float) * (C::N + (size_t)(C::FB - 1) * p.hop + C::N);
    if (smem > 227 * 1024) {
        set_error("aero_stft_fwd: hop %d too large for n_fft %d (smem %zu)", p.hop, p.n_fft, smem);
        return AERO_ERR_UNSUPPORTED;
    }
    cudaFuncSetAttribute(stft_kernel<LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(cdiv(p.frames, C::FB), p.n_signals);
    stft_kernel<LOGN><<<grid, kThreads, smem, st>>>(x, window, z, stats, p);
    return check_launch("aero_stft_fwd");
}

// ---------------------------------------------------------------------------------- inverse
template <int LOGN>
struct IstftCfg {
    static constexpr int N = 1 << LOGN;
    static constexpr int M = N / 2;
    static constexpr int NF = (8192 / M) > 32 ? 32 : (8192 / M);       // frames resident per CTA (64 KB)
};

template <int LOGN>
__global__ void __launch_bounds__(kThreads) istft_kernel(const float* __restrict__ z, const float* __restrict__ window,
                                                         float* __restrict__ y, const aero_istft_params p,
                                                         const int OB, const int halo) {
    using C = IstftCfg<LOGN>;
    constexpr int N = C::N, M = C::M, NF = C::NF, LOGM = LOGN - 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* work = reinterpret_cast<float2*>(smem_raw);                 // [NF][M]  (later: [NF][N] real frames)
    float2* xs = work + NF * M;                                         // [NF][M+1] spectra
    float2* twN = xs + NF * (M + 1);                                    // [M] exp(+2 pi i j / N)
    float* wpad = reinterpret_cast<float*>(twN + M);                    // [N]

    const int sig = blockIdx.y;
    const int blk = blockIdx.x;
    const int t_lo = max(0, blk * OB - halo);
    const int t_hi = min(p.frames - 1, blk * OB + OB - 1);
    const int nfr = t_hi - t_lo + 1;

    for (int j = threadIdx.x; j < M; j += kThreads) {
        float s, c;
        sincospif(2.0f * (float)j / (float)N, &s, &c);
        twN[j] = make_float2(c, s);
    }
    const int wl = (N - p.win) / 2;
    for (int n = threadIdx.x; n < N; n += kThreads) {
        const int k = n - wl;
        wpad[n] = (k >= 0 && k < p.win) ? window[k] : 0.0f;
    }
    const float* zs = z + (int64_t)(sig / p.channels) * p.z_stride_b + (int64_t)(sig % p.channels) * p.z_stride_c;
    for (int i = threadIdx.x; i < (M + 1) * nfr; i += kThreads) {
        const int k = i / nfr, fr = i - k * nfr;
        float2 v = make_float2(0.f, 0.f);
        if (k < p.bins_in)
            v = *reinterpret_cast<const float2*>(zs + (int64_t)k * p.z_stride_k + (int64_t)(t_lo + fr) * p.z_stride_t);
        if (k == 0 || k == M) v.y = 0.f;       // C2R ignores the imaginary part of DC and Nyquist
        xs[fr * (M + 1) + k] = v;
    }
    __syncthreads();

    // Y[k] = Xe[k] + i Xo[k];  Xe = (X[k]+conj(X[M-k]))/2,  Xo = conj(w^k) (X[k]-conj(X[M-k]))/2
    for (int i = threadIdx.x; i < nfr * M; i += kThreads) {
        const int fr = i / M, k = i - fr * M;
        const float2 a = xs[fr * (M + 1) + k];
        const float2 b = xs[fr * (M + 1) + (M - k)];
        const float2 xe = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
        const float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));
        const float2 xo = cmul(twN[k], d);                              // twN holds exp(+i..) = conj(w^k)
        const int r = __brev((unsigned)k) >> (32 - LOGM);
        work[fr * M + r] = make_float2(xe.x - xo.y, xe.y + xo.x);      // xe + i*xo
    }
    __syncthreads();
    fft_inplace<LOGM>(work, twN, nfr);

    // overlap-add; work now holds real frames: frame fr, sample n at ((float*)work)[fr*N + n]
    const float* frames = reinterpret_cast<const float*>(work);
    const float scale = 2.0f * rsqrtf((float)N);                        // sqrt(N) / M
    const int p0 = blk * OB * p.hop;
    const int span = OB * p.hop;
    float* ys = y + (int64_t)sig * p.out_len;
    for (int i = threadIdx.x; i < span; i += kThreads) {
        const int pos = p0 + i;
        const int n_out = pos - N / 2;
        if (n_out < 0 || n_out >= p.out_len) continue;
        int ta = (pos - N + p.hop) / p.hop;                             // ceil((pos-N+1)/hop) for pos-N+1 > 0
        if (pos - N + 1 <= 0) ta = 0;
        ta = max(ta, t_lo);
        const int tb = min(pos / p.hop, t_hi);
        float acc = 0.f, env = 0.f;
        for (int t = ta; t <= tb; ++t) {
            const int n = pos - t * p.hop;
            const float w = wpad[n];
            acc += frames[(t - t_lo) * N + n] * w;
            env += w * w;
        }
        ys[n_out] = acc * scale / env;
    }
}

template <int LOGN>
static int launch_istft(const float* z, const float* window, float* y, const aero_istft_params& p, cudaStream_t st) {
    using C = IstftCfg<LOGN>;
    const int halo = (C::N - 1) / p.hop;
    const int OB = C::NF - halo;
    if (OB < 1) {
        set_error("aero_istft_fwd: hop %d too small for n_fft %d (needs hop >= n_fft/%d)", p.hop, p.n_fft, C::NF - 1);
        return AERO_ERR_UNSUPPORTED;
    }
    const size_t smem = sizeof(float2) * (C::NF * C::M + C::NF * (C::M + 1) + C::M) + sizeof(float) * C::N;
    cudaFuncSetAttribute(istft_kernel<LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // padded positions that can produce output: [N/2, N/2 + out_len)
    const int last_pos = C::N / 2 + p.out_len - 1;
    dim3 grid(last_pos / (OB * p.hop) + 1, p.n_signals);
    istft_kernel<LOGN><<<grid, kThreads, smem, st>>>(z, window, y, p, OB, halo);
    return check_launch("aero_istft_fwd");
}

static int log2_exact(int n) {
    int l = 0;
    while ((1 << l) < n) ++l;
    return (1 << l) == n ? l : -1;
}

}  // namespace aero

extern "C" int aero_stft_fwd(const float* x, const float* window, float* z, double* stats, const aero_stft_params* p,
                             aero_stream_t stream) {
    using namespace aero;
    AERO_REQUIRE(x && window && z && p, "aero_stft_fwd: null argument");
    const int lg = log2_exact(p->n_fft);
    AERO_REQUIRE(lg >= 6 && lg <= 12, "aero_stft_fwd: n_fft=%d must be a power of two in [64,4096]", p->n_fft);
    AERO_REQUIRE(p->win >= 1 && p->win <= p->n_fft && p->hop >= 1, "aero_stft_fwd: bad win/hop %d/%d", p->win, p->hop);
    AERO_REQUIRE(p->length > p->n_fft / 2, "aero_stft_fwd: reflect padding needs length (%d) > n_fft/2", p->length);
    AERO_REQUIRE(p->frames == 1 + p->length / p->hop, "aero_stft_fwd: frames=%d != 1+length/hop", p->frames);
    AERO_REQUIRE(p->bins_out >= 1 && p->bins_out <= p->n_fft / 2 + 1, "aero_stft_fwd: bins_out=%d", p->bins_out);
    AERO_REQUIRE(p->n_signals >= 1 && p->channels >= 1 && p->n_signals % p->channels == 0, "aero_stft_fwd: signals/channels");
    AERO_REQUIRE(((p->z_stride_b | p->z_stride_c | p->z_stride_k | p->z_stride_t) & 1) == 0 && ((uintptr_t)z & 7) == 0,
                 "aero_stft_fwd: output strides must keep float2 alignment");
    cudaStream_t st = (cudaStream_t)stream;
    switch (lg) {
        case 6: return launch_stft<6>(x, window, z, stats, *p, st);
        case 7: return launch_stft<7>(x, window, z, stats, *p, st);
        case 8: return launch_stft<8>(x, window, z, stats, *p, st);
        case 9: return launch_stft<9>(x, window, z, stats, *p, st);
        case 10: return launch_stft<10>(x, window, z, stats, *p, st);
        case 11: return launch_stft<11>(x, window, z, stats, *p, st);
        default: return launch_stft<12>(x, window, z, stats, *p, st);
    }
}

extern "C" int aero_istft_fwd(const float* z, const float* window, float* y, const aero_istft_params* p,
                              aero_stream_t stream) {
    using namespace aero;
    AERO_REQUIRE(z && window && y && p, "aero_istft_fwd: null argument");
    const int lg = log2_exact(p->n_fft);
    AERO_REQUIRE(lg >= 6 && lg <= 12, "aero_istft_fwd: n_fft=%d must be a power of two in [64,4096]", p->n_fft);
    AERO_REQUIRE(p->win >= 1 && p->win <= p->n_fft && p->hop >= 1, "aero_istft_fwd: bad win/hop");
    AERO_REQUIRE(p->bins_in >= 1 && p->bins_in <= p->n_fft / 2 + 1, "aero_istft_fwd: bins_in=%d", p->bins_in);
    AERO_REQUIRE(p->out_len >= 1 && p->out_len <= p->hop * (p->frames - 1), "aero_istft_fwd: out_len=%d > hop*(frames-1)", p->out_len);
    AERO_REQUIRE(p->n_signals >= 1 && p->channels >= 1 && p->n_signals % p->channels == 0, "aero_istft_fwd: signals/channels");
    AERO_REQUIRE(((p->z_stride_b | p->z_stride_c | p->z_stride_k | p->z_stride_t) & 1) == 0 && ((uintptr_t)z & 7) == 0,
                 "aero_istft_fwd: input strides must keep float2 alignment");
    cudaStream_t st = (cudaStream_t)stream;
    switch (lg) {
        case 6: return launch_istft<6>(z, window, y, *p, st);
        case 7: return launch_istft<7>(z, window, y, *p, st);
        case 8: return launch_istft<8>(z, window, y, *p, st);
        case 9: return launch_istft<9>(z, window, y, *p, st);
        case 10: return launch_istft<10>(z, window, y, *p, st);
        case 11: return launch_istft<11>(z, window, y, *p, st);
        default: return launch_istft<12>(z, window, y, *p, st);
    }
}
