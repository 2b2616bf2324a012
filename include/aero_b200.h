/*
 * aero_b200.h -- C ABI of libaero_b200.so: the sm_100a kernels behind the AERO generator forward.
 *
 * The reference (slp-rl/aero) has no FFI: its hot path is the Python class
 * src/models/aero.py:218 `Aero`, whose arithmetic is dispatched to PyTorch library kernels.
 * Each entry point below replaces the library call(s) made at the cited reference lines.  A
 * binding needs nothing but device pointers, plain-old-data parameter blocks and a CUDA stream:
 * no ATen / Python types cross this boundary (see INTEGRATION.md for the ctypes stub).
 *
 * Conventions
 *   - All tensors are fp32, device memory, owned by the caller.  Kernels never allocate, never
 *     synchronise, and enqueue on `stream` only; entry points are re-entrant per stream.
 *   - Activations are channels-last: X[b][f][t][c] ("rows" (b,f) of T frames of C channels),
 *     the natural layout for this model: complex64 [B,F,T] *is* [B,F,T,2].
 *   - Return value: 0 on success, negative aero_status on error; aero_last_error() gives the
 *     text for the calling thread.  Launch errors are reported, asynchronous faults surface at
 *     the caller's next synchronisation.
 *   - Statistics buffers are fp64 pairs {sum, sum of squares} accumulated with atomics; the
 *     caller zeroes them (cudaMemsetAsync) before the producing call.
 */
#ifndef AERO_B200_H
#define AERO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* aero_stream_t; /* cudaStream_t */

enum aero_status {
    AERO_OK = 0,
    AERO_ERR_INVALID = -1,   /* bad parameter block */
    AERO_ERR_UNSUPPORTED = -2,
    AERO_ERR_LAUNCH = -3,    /* cudaGetLastError() after launch */
    AERO_ERR_NO_DEVICE = -4
};

int aero_abi_version(void);            /* 3: training entry points */
const char* aero_last_error(void);
/* compute capability of the current device as 10*major+minor (100 on B200); <0 if no device */
int aero_device_arch(void);
/* number of kernel launches issued through this library by the calling process (bench.py's gpu_launches) */
uint64_t aero_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * STFT  (replaces torch.stft at reference src/models/spec.py:12-20, called from aero.py:420,
 *        plus the Nyquist drop aero.py:420 `[..., :-1, :]`, the complex->channels permute
 *        aero.py:430-434 and the moments for aero.py:462-463)
 * x      : [n_signals][length]           real input, n_signals = B * channels
 * window : [win]                         analysis window (torch.hann_window(win), spec.py:15)
 * z      : element (sig, k, t) is the float2 at  z + (sig / channels) * z_stride_b
 *              + (sig % channels) * z_stride_c + k * z_stride_k + t * z_stride_t   (strides in floats)
 * stats  : optional [B][2] fp64 {sum, sumsq} over every float written for batch item b
 * Semantics: centre=True reflect padding of n_fft/2, window zero-padded (centred) to n_fft,
 * normalized=True (x n_fft^-1/2), onesided; bins 0 .. bins_out-1 are written
 * (bins_out = n_fft/2 drops Nyquist, n_fft/2+1 keeps it).  frames must equal 1 + length / hop.
 */
typedef struct {
    int32_t n_fft, hop, win;
    int32_t n_signals, channels;
    int32_t length, frames, bins_out;
    int64_t z_stride_b, z_stride_c, z_stride_k, z_stride_t;
    int32_t flags;                    /* 0, or AERO_STFT_* (training: this kernel is also the adjoint of the iSTFT) */
    int32_t reserved;
} aero_stft_params;
enum {
    AERO_STFT_ZERO_PAD = 1,           /* samples outside [0, length) are zero instead of reflected */
    AERO_STFT_ADJ_SCALE = 2           /* bins 1 .. n_fft/2-1 are doubled and the imaginary parts of DC / Nyquist zeroed: with ZERO_PAD and
                                         x = dy / envelope (zero-extended to hop*(frames-1)) this is dL/dz of aero_istft_fwd */
};
int aero_stft_fwd(const float* x, const float* window, float* z, double* stats,
                  const aero_stft_params* p, aero_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * iSTFT (replaces torch.istft at reference src/models/spec.py:30-37, called from aero.py:427,
 *        plus the Nyquist zero-pad aero.py:426 and the trim aero.py:513)
 * z addressing as above (bins_in bins present; bins above are taken as zero).
 * y : [n_signals][out_len],  y[n] = OLA(irfft(z * sqrt(n_fft)) * window)[n + n_fft/2] / sum(window^2)
 * out_len <= hop * (frames - 1).
 */
typedef struct {
    int32_t n_fft, hop, win;
    int32_t n_signals, channels;
    int32_t frames, bins_in, out_len;
    int64_t z_stride_b, z_stride_c, z_stride_k, z_stride_t;
    int32_t flags;                    /* 0, or AERO_ISTFT_RAW */
    int32_t reserved;
} aero_istft_params;
enum {
    AERO_ISTFT_RAW = 1                /* y[pos] = OLA(irfft(z * sqrt(n_fft)) * window)[pos], pos < out_len <= hop*(frames-1) + n_fft: no centre
                                         trim and no envelope division.  Applied to dL/dz with the interior bins halved this is the
                                         gradient of the reflect-PADDED input of aero_stft_fwd (the caller folds the padding back). */
};
int aero_istft_fwd(const float* z, const float* window, float* y,
                   const aero_istft_params* p, aero_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Tap-GEMM: every convolution / linear layer of the model as one implicit GEMM
 *   acc[b][fo][t][n] = sum_{tap} sum_{c < C1+C2}  A(b, fi(tap,fo), ti(tap,t), c) * W[slab(tap,fo)][c][n]
 * with A read from up to two channels-last sources (channel concat, reference aero.py:195
 * `torch.cat([x, skip], 1)`); out-of-range fi / ti contribute zero (zero padding).
 *   mode AERO_TAPS_CONV  : fi = fo*stride_f + jf - pad_f,  ti = t + jt*dil_t - pad_t,
 *                          slab = jf*kt + jt                      (nn.Conv2d / nn.Conv1d / nn.Linear)
 *   mode AERO_TAPS_CONVT : fo' = fo + f_out_offset (row in the uncropped output),
 *                          jf in [0, kf/stride_f): kf_idx = fo' % stride_f + jf*stride_f,
 *                          fi = fo' / stride_f - jf,  slab = kf_idx  (nn.ConvTranspose2d [k,1]/[s,1])
 *   mode AERO_TAPS_MIX   : (precision 1 / 2 only) contraction over the ROW axis of a channels-last tensor, FTB's
 *                          `freq_fc` (modules.py:296,317-320):  out[b][n][m] = colscale[b][m] * sum_{k<C1} a1[b][k][m] * W[n][k],
 *                          m < T pixels (contiguous), a1 element (b,k,m) at a1 + b*a1_sb + k*a1_st + m, out element
 *                          at out + b*o_sb + n*o_st + m, colscale at colscale + b*cs_sb + m.  F_out = F_in = 1.
 * Replaces nn.Conv2d/ConvTranspose2d (aero.py:89,95,101,172,179), nn.Conv1d (modules.py:80-92,
 * 206,209,292), nn.Linear (modules.py:29,296) and their cuDNN/cuBLAS kernels.
 *
 * Epilogue, in order:  v = acc + bias[n];  v *= colscale[b][t][n]  (FTB gate, modules.py:314);
 *   act (none | exact GELU | ReLU);
 *   GLU over adjacent column pairs (2j, 2j+1) -> channel j  (weights are stored pair-interleaved);
 *   v += addend_fn[fo][n'] (frequency embedding, aero.py:475-480);
 *   v = residual[b][fo][t][n'] + v  ;  v = v * samp_affine[b][0] + samp_affine[b][1] (aero.py:497-498);
 *   statistics of the stored value:  stats_mode 1: per (b, group), group = n' / (N' / groups);
 *   stats_mode 2: per (b, fo) row.   N' = N/2 with GLU, else N.
 * Generic strides (in elements) let one kernel serve NCHW-free layouts: element (b,f,t,c) of a
 * source is at  src + b*sb + f*sf + t*st + c.
 */
enum { AERO_TAPS_CONV = 0, AERO_TAPS_CONVT = 1, AERO_TAPS_MIX = 2 };
enum { AERO_ACT_NONE = 0, AERO_ACT_GELU = 1, AERO_ACT_RELU = 2 };
typedef struct {
    int32_t B, F_out, T, N;
    int32_t F_in, T_in;
    int32_t C1, C2;
    int32_t mode, kf, kt, stride_f, pad_f, dil_t, pad_t, f_out_offset;
    int32_t act, glu, stats_mode, groups;
    int64_t a1_sb, a1_sf, a1_st;
    int64_t a2_sb, a2_sf, a2_st;
    int64_t w_sb;                     /* weight batch stride (0: shared weights) */
    int64_t o_sb, o_sf, o_st;
    int64_t r_sb, r_sf, r_st;         /* residual strides */
    int64_t cs_sb, cs_st;             /* colscale strides */
    int32_t precision;                /* 0: fp32 SIMT tiles, fp32 weights N-contiguous  W[slab][K][pad4(N)] (sources / outputs of either type);
                                         1: tcgen05 kind::tf32, fp32 sources, fp32 weights K-contiguous W[slab][pad4(N)][K] rounded to TF32;
                                         2: tcgen05 kind::f16, FP16 sources, FP16 weights K-contiguous W[slab][pad4(N)][pad8(K)];
                                            1 / 2: AERO_ERR_UNSUPPORTED unless aero_tapgemm_tc_eligible() */
    int32_t flags;                    /* AERO_TG_* */
} aero_tapgemm_params;
/* Storage types.  Activations that feed a tensor-core GEMM are stored either as fp32 rounded to TF32 or as FP16 (the same
 * 10-bit mantissa at half the bytes; stores saturate at +-65504); every kernel computes in fp32 between load and store.
 * Strides are in ELEMENTS of the tensor they address. */
enum {
    AERO_TG_ROUND_TF32 = 1,           /* fp32 outputs: round stored values to TF32 (round-to-nearest) for a kind::tf32 consumer */
    AERO_TG_A_F16 = 2,                /* a1 / a2 are FP16 */
    AERO_TG_OUT_F16 = 4,              /* out and residual are FP16 */
    AERO_TG_REVERSE = 8               /* tap-GEMM (tcgen05 path) / norm_act: walk tiles / rows from the end.  Results are identical;
                                         a kernel launched right after its producer then starts on the data the producer wrote last,
                                         which is still in the 126 MB L2 (the host alternates the direction from launch to launch) */
};
int aero_tapgemm_fwd(const void* a1, const void* a2, const void* w, const float* bias,
                     const float* addend_fn, const float* colscale, const void* residual,
                     const float* samp_affine, void* out, double* stats,
                     const aero_tapgemm_params* p, aero_stream_t stream);
/* 1 when the shape can run on the tcgen05 path (kind::tf32, or kind::f16 when flags has AERO_TG_A_F16), else 0 */
int aero_tapgemm_tc_eligible(const aero_tapgemm_params* p);

/* ------------------------------------------------------------------------------------------
 * Normalisation / activation passes (HBM-bound elementwise kernels)
 *
 * aero_sample_norm_fwd: per-sample standardisation of the input spectrogram (reference
 *   aero.py:462-464): mean / unbiased std over `count` values from stats[b] = {sum, sumsq};
 *   y = (x - mean) / (1e-5 + std) applied to `extent` (>= count; 0 means count) contiguous floats per sample -- rows may
 *   carry alignment padding that the statistics did not see; also writes samp_affine[b] = {std, mean} for the output
 *   de-normalisation (aero.py:497-498).
 */
int aero_sample_norm_fwd(const float* x, const double* stats, float* y, float* samp_affine,
                         int32_t B, int64_t count, int64_t extent, int32_t round_tf32, aero_stream_t stream);

/* aero_norm_act_fwd: y = op(GroupNorm(x))   (replaces nn.GroupNorm + F.gelu / F.glu / Snake /
 *   LayerScale + residual: aero.py:127,133,198,206-214; modules.py:189,210,232-244; snake.py:67)
 * x : [B][F_in][T][C]; rows f_off .. f_off+F_out-1 are read (decoder crop, aero.py:209).
 * stats : fp64 {sum, sumsq}; scope 1: [B][groups] over F_in*T*(C/groups) values (uncropped);
 *         scope 2: [B*F_in][1] per row over T*C values (DConv's GroupNorm(1, C)).
 * op: AERO_NA_NONE y=g; AERO_NA_GELU; AERO_NA_GLU y[c]=g[c]*sigmoid(g[c+C/2]) (C_out=C/2);
 *     AERO_NA_SNAKE y = g + sin(a[f]*g)^2 / a[f];
 *     AERO_NA_GLU_SCALE_RES y[c] = residual[c] + scale[c] * glu(g)[c].
 */
enum { AERO_NA_NONE = 0, AERO_NA_GELU = 1, AERO_NA_GLU = 2, AERO_NA_SNAKE = 3, AERO_NA_GLU_SCALE_RES = 4,
       AERO_NA_RELU = 5, AERO_NA_LEAKY = 6 /* LeakyReLU(0.2) */ /* 5, 6: training entry points only */ };
enum { AERO_NA_NO_NORM = 16 };      /* aero_norm_act_params.flags, training entry points: skip the normalisation (activation only) */
typedef struct {
    int32_t B, F_in, F_out, f_off, T, C;
    int32_t groups, scope, op;
    float eps;
    int32_t flags;                    /* AERO_TG_ROUND_TF32: round stored fp32 outputs to TF32 for a tensor-core consumer;
                                         AERO_TG_OUT_F16: y and residual are FP16; AERO_TG_A_F16: x is FP16 too (pre-normalisation
                                         tensors stored in FP16; the statistics were taken from the stored values); x == y is
                                         allowed when they have the same type and the op keeps the channel count */
} aero_norm_act_params;
int aero_norm_act_fwd(const void* x, const double* stats, const float* gamma, const float* beta,
                      const float* snake_a, const float* scale, const void* residual, void* y,
                      const aero_norm_act_params* p, aero_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * FTB output conv through a linear input (encoder layer 0: `pre_conv` aero.py:89,112 followed by FTB modules.py:304-325).
 * x = pre_conv(z) is linear in the J = 2*C_in spectrogram channels, and so is everything FTB does before its last ReLU:
 *   out[b,f,t,n] = relu( sum_{j<J} M[b,t][n][j] * zm[b,f,t,j] + M[b,t][n][J] * s[f] + sum_{j<J} V[n][j] * z[b,f,t,j] + d[n] )
 * with zm = freq_fc applied to z (AERO_TAPS_MIX), s[f] = sum_f' Wfc[f][f'], M[b,t] = gate[b,t,:] . Q,
 *   Q[c][n*(J+1)+j] = W2a[n][c] * (j < J ? Wpre[c][j] : bpre[c]),  V = W2b Wpre,  d = W2b bpre + b2
 * (W2 = [W2a | W2b] the BatchNorm-folded FTB conv2 acting on cat([freq_fc out, x]), modules.py:322-324).
 * z, zm : fp32, element (b,f,t,j) at base + b*sb + f*sf + t*J + j;  M : fp32 [B*T][N*(J+1)];  out : [B][F][T][N] fp32 / FP16.
 * The C-channel tensors pre_conv(z), freq_fc(..) and their concatenation are never materialised.
 */
typedef struct {
    int32_t B, F, T, N, J;
    int32_t flags;                    /* AERO_TG_OUT_F16 / AERO_TG_ROUND_TF32 */
    int64_t z_sb, z_sf, zm_sb, zm_sf;
} aero_ftb_lin_params;
int aero_ftb_lin_out_fwd(const float* z, const float* zm, const float* M, const float* s, const float* V, const float* d,
                         void* out, const aero_ftb_lin_params* p, aero_stream_t stream);
/* FTB squeeze (modules.py:286-291,308-309: 1x1 conv C -> r, BatchNorm, ReLU, regrouped to [B][T][F*r]) through the same
 * linearity:  R[b][t][f*r + n] = relu( sum_j W1p[n][j] * z[b,f,t,j] + b1p[n] ),  W1p = W1 Wpre [r][J], b1p = W1 bpre + b1, r <= 8.
 * R is fp32 or FP16 (flags); p->N is unused. */
int aero_ftb_lin_squeeze_fwd(const float* z, const float* W1p, const float* b1p, void* R, int32_t r,
                             const aero_ftb_lin_params* p, aero_stream_t stream);

/* FTB frequency mix (`freq_fc`, modules.py:296,317-320) for the deep layers where only F = 8 / 16 frequency rows are left:
 *   out[b][g][m] = gate[b][m] * sum_f W[g][f] * x[b][f][m],  m < M = T*C contiguous positions (M a multiple of 4), W fp32 [F][F],
 * gate fp32 [B][M] or NULL; x and out are both fp32 or both FP16 (AERO_TG_A_F16 | AERO_TG_OUT_F16).  Same result as
 * AERO_TAPS_MIX, without tensor-core tiles (they are all overhead at this F). */
int aero_freq_mix_small_fwd(const void* x, const float* W, const float* gate, void* out, int32_t B, int32_t F, int64_t M,
                            int32_t flags, aero_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Recurrent half of one bidirectional LSTM layer (replaces the cuDNN RNN behind nn.LSTM,
 * reference modules.py:28,46, together with the overlapping-window framing modules.py:36-44,
 * utils.py:22-35 and the central-crop reassembly modules.py:49-60).
 * The input projections x_t W_ih^T + b_ih + b_hh for BOTH directions are computed beforehand by
 * aero_tapgemm_fwd into `gin` (8H columns: [dir][i,f,g,o][H]).
 *   in_windowed = 0 : gin is [rows][T][8H] over the un-windowed sequence; window k, step p reads
 *                     frame k*win_stride + p; frames >= T are zero inputs, so their gate
 *                     pre-activation is `bias_pad` ([8H] = b_ih + b_hh)   (first layer)
 *   in_windowed = 1 : gin is [rows*n_win][steps][8H]                        (second layer)
 *   out_windowed = 1: h written as [rows*n_win][steps][2H]                  (first layer)
 *   out_windowed = 0: h written de-windowed as [rows][T][2H], keeping for window k the steps
 *                     [keep_lo(k), keep_hi(k)) exactly as modules.py:53-59    (second layer)
 * whh : [2][4H][H] (PyTorch layout weight_hh_l{k}, weight_hh_l{k}_reverse), gate order i,f,g,o.
 */
typedef struct {
    int32_t rows, T, H;
    int32_t n_win, steps, win_stride;
    int32_t in_windowed, out_windowed;
    int32_t flags;                    /* AERO_TG_ROUND_TF32: round the stored fp32 h to TF32; AERO_TG_OUT_F16: hout is FP16;
                                         AERO_TG_A_F16 (precision 1 only): gin is FP16 (bias_pad stays fp32) */
    int32_t precision;                /* 0: fp32 SIMT recurrence (layouts above);
                                         1: tcgen05 recurrence, FP16 operands (h in (-1,1), W_hh O(1): same 10-bit mantissa as TF32), fp32
                                            accumulate.  `gin` / `bias_pad` keep the layout above; only `whh` changes: FP16
                                            [2][(4/GPT)*128][Kp], Kp = 64*ceil(H/64), gate rows re-ordered into 4/GPT tiles of 128 per
                                            direction (GPT = 2 if H <= 64 else 1): GPT=1: tile g = gate g, row = cell; GPT=2: tile t =
                                            gates (2t, 2t+1), in each group of 32 rows rows 0-15 carry gate 2t and rows 16-31 gate 2t+1
                                            of the same 16 cells (zero rows beyond H, zero columns beyond H).  H % 4 == 0, 32 < H <= 96. */
} aero_lstm_params;
int aero_lstm_rec_fwd(const void* gin, const float* bias_pad, const void* whh, void* hout,
                      const aero_lstm_params* p, aero_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * LocalState attention core (replaces the einsum/softmax/einsum chain reference
 * modules.py:104-124; never materialises the T x T matrices).
 * qkvd : [rows][T][ld]  columns: q [0,H) | k [H,2H) | v (content) [2H,3H) | decay logits [3H, 3H+heads*ndecay)
 * out  : [rows][T][H],  out[s][h*d+c] = sum_t softmax_t( k_t.q_s/sqrt(d) - |t-s|*slope_s, diag=-100 ) * v_t[c]
 *        slope_s = sum_{f=1..ndecay} f * sigmoid(decay[h*ndecay+f-1][s]) / 2 / sqrt(ndecay)
 */
typedef struct {
    int32_t rows, T, H, heads, ndecay, ld;
    int32_t flags;                    /* AERO_TG_ROUND_TF32: tensor-core mode -- QK^T and PV on mma.sync TF32 (fp32 accumulate, head dim
                                         12 / 24), outputs rounded to TF32 for the projection GEMM; without it the exact fp32 SIMT kernel.
                                         AERO_TG_OUT_F16: out is FP16 */
} aero_attn_params;
int aero_local_attn_fwd(const float* qkvd, void* out, const aero_attn_params* p, aero_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Log-spectral distance (reference src/metrics.py:37-70 `get_lsd`, STFTMag(2048, 512)).
 * z_ref, z_est : planar complex spectrograms [B][bins][frames] (float2) as written by aero_stft_fwd (normalized);
 * out_sum += sum over (b, t) of sqrt(mean_f (log10 max(n_fft|z_ref|^2, 1e-8) - log10 max(n_fft|z_est|^2, 1e-8))^2).
 * The caller zeroes out_sum and divides by B * frames.
 */
int aero_lsd_fwd(const float* z_ref, const float* z_est, double* out_sum, int32_t B, int32_t bins, int32_t frames,
                 int32_t n_fft, aero_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Multi-resolution STFT loss, forward value (reference src/models/stft_loss.py:11-27,30-63,96-138).
 * z_est, z_ref : planar complex spectrograms [B][bins][frames] (float2) of the estimate and the target as written by
 * aero_stft_fwd (normalised) at ONE resolution; with mag = sqrt(max(n_fft |z|^2, 1e-7)):
 *   sums[0] += sum (mag_ref - mag_est)^2;  sums[1] += sum mag_ref^2;  sums[2] += sum |log mag_ref - log mag_est|
 * -> spectral convergence = sqrt(sums[0]/sums[1]), log-magnitude L1 = sums[2] / (B*bins*frames).  The caller zeroes `sums`.
 */
int aero_stft_loss_fwd(const float* z_est, const float* z_ref, double* sums, int32_t B, int32_t bins, int32_t frames,
                       int32_t n_fft, aero_stream_t stream);

/* Backward of the above for the estimate: with the `sums` of the forward, g_est[B][bins][frames] (float2) = gradient of
 *   c_sc * sqrt(sums[0]/sums[1]) + c_mag * sums[2] / (B*bins*frames)
 * with respect to the normalised spectrogram z_est, the interior bins already halved: feeding g_est to aero_istft_fwd with
 * AERO_ISTFT_RAW gives the gradient of the reflect-padded waveform (the caller folds the padding back).  bins = n_fft/2+1.
 * (reference: autograd through stft_loss.py:11-63, called with gradients at solver.py:470-473) */
int aero_stft_loss_bwd(const float* z_est, const float* z_ref, const double* sums, float* g_est, int32_t B, int32_t bins,
                       int32_t frames, int32_t n_fft, float c_sc, float c_mag, aero_stream_t stream);

/* ==========================================================================================
 * Training (backward of the custom ops + fused Adam; reference callers src/solver.py:292-342,
 * 602-605 `loss.backward(); optimizer.step()`, train.py:83 `torch.optim.Adam`).  All tensors fp32.
 *
 * Data gradients of every convolution / linear layer run on aero_tapgemm_fwd itself: the adjoint of a tap-GEMM is a
 * tap-GEMM (taps flipped, weights transposed, AERO_TAPS_CONV <-> AERO_TAPS_CONVT for strided layers).
 * ========================================================================================== */

/* Weight gradient of a tap-GEMM:  dW(n, k, slab) += sum_{b,fo,t} A(b, fi, ti, k) * dY(b, fo, t, n)  with (fi, ti, slab) the tap
 * geometry of aero_tapgemm_fwd for `p` (mode AERO_TAPS_CONV or AERO_TAPS_CONVT; A = channel concat of a1, a2).  dY is addressed
 * with p->o_sb / o_sf / o_st; element (n, k, slab) of dW lives at dw + n*dw_sn + k*dw_sk + slab*dw_ss (so the gradient can be
 * written straight into PyTorch's [N][K][kf][kt] or ConvTranspose [K][N][kf][1] parameter layout).  The caller zeroes dW.
 * p->precision: 0 = exact fp32 (SIMT; a one-thread-per-weight kernel when the layer has at most 2048 weights); 1 = TF32 on the
 * tensor cores (csrc/wgrad_tc.cu: both operands MN-major through TMA, split over the pixel axis, fp32 atomics) where the shape allows
 * (N >= 16, K >= 16, channel counts and strides multiples of 4, 16-byte aligned bases), exact fp32 otherwise.
 * Replaces the cuDNN wgrad kernels behind autograd of nn.Conv2d / ConvTranspose2d / Conv1d / Linear. */
int aero_tapgemm_wgrad(const float* a1, const float* a2, const float* dy, float* dw, const aero_tapgemm_params* p,
                       int64_t dw_sn, int64_t dw_sk, int64_t dw_ss, aero_stream_t stream);

/* Column sums:  out1[seg][n] += sum_{o < n_outer, i < n_inner} x[seg*seg_stride_x + o*outer_stride + i*inner_stride + n],
 * out2[seg][n] += the same sum of x * z (z addressed like x; NULL: skipped).  out_double: outputs are fp64 (statistics) else fp32.
 * Bias gradients, BatchNorm batch statistics (z = x), frequency-embedding gradients.  The caller zeroes the outputs. */
int aero_colsum(const float* x, const float* z, void* out1, void* out2, int32_t out_double, int32_t N, int64_t n_inner,
                int64_t inner_stride, int64_t n_outer, int64_t outer_stride, int32_t n_seg, int64_t seg_stride_x,
                int64_t seg_stride_out, aero_stream_t stream);

/* out[i][j] += sum_{b, m} P[b][i][m] * gate[b][m] * Q[b][j][m],  i, j < F,  m < M contiguous (batch strides sb_*; gate may be
 * NULL): weight gradient of FTB's `freq_fc` (modules.py:296,317-320), P = dY, Q = x, gate = the FTB gate.  Caller zeroes out. */
int aero_gram(const float* P, const float* Q, const float* gate, float* out, int32_t B, int32_t F, int64_t M, int64_t sb_p,
              int64_t sb_q, int64_t sb_g, aero_stream_t stream);
/* x[b][f][t][c] += addend[f][c]  (frequency embedding, aero.py:475-480; fused into a GEMM epilogue at inference). */
int aero_bcast_add(float* x, const float* addend, int32_t B, int32_t F, int32_t T, int32_t C, aero_stream_t stream);
/* y[b][i] = x[b][i] * s[b*s_stride], i < per_sample  (backward of the per-sample de-normalisation aero.py:497-498). */
int aero_scale_rows(const float* x, float* y, const float* s, int32_t B, int64_t per_sample, int32_t s_stride, aero_stream_t stream);
/* dst[i] += alpha * src[i] (gradient accumulation where a tensor has several consumers). */
int aero_add(float* dst, const float* src, int64_t n, float alpha, aero_stream_t stream);
/* dst[i] += (float) src[i]: fp64 column sums into an fp32 parameter gradient. */
int aero_add_f64(float* dst, const double* src, int64_t n, aero_stream_t stream);

/* Normalisation + activation, training form (nothing folded, fp32): y = act(norm(x)) with
 *   scope 1 / 2: GroupNorm as in aero_norm_act_fwd;  scope 3: BatchNorm with BATCH statistics, stats = [C][2] fp64 {sum, sumsq}
 *   over all B*F_in*T pixels (reference modules.py:287-300 in train mode);  flags & AERO_NA_NO_NORM: activation only.
 * ops: AERO_NA_* (SNAKE needs scope 2).  x [B][F_in][T][C], y [B][F_out][T][C or C/2]. */
int aero_norm_act_train_fwd(const float* x, const double* stats, const float* gamma, const float* beta, const float* snake_a,
                            const float* scale, const float* residual, float* y, const aero_norm_act_params* p,
                            aero_stream_t stream);
/* Backward of the above.  pass 1 accumulates dgamma[C], dbeta[C], dscale[C/2] (GLU_SCALE_RES), dsnake[F_in] (SNAKE) -- all fp64:
 * these are sums over every pixel of the batch whose terms largely cancel, and fp32 atomics across hundreds of CTAs lose
 * ~1e-3 of the result -- and the per-(segment, group) sums ws[slot] = {sum dxh, sum dxh*xh} (fp64, same slots as `stats`;
 * unused for scope 3, which reads dgamma / dbeta instead);
 * pass 2 writes dx[B][F_in][T][C] = rstd * (dxh - mean(dxh) - xh * mean(dxh * xh)) (every input row, cropped rows included).
 * The gradient of the residual input of GLU_SCALE_RES is dy itself (the caller accumulates it).  The caller zeroes the
 * accumulators before pass 1. */
int aero_norm_act_train_bwd(const float* x, const double* stats, const float* gamma, const float* beta, const float* snake_a,
                            const float* scale, const float* dy, float* dx, double* dgamma, double* dbeta, double* dscale,
                            double* dsnake, double* ws, int32_t pass, const aero_norm_act_params* p, aero_stream_t stream);

/* LSTM layer, training form (fp32 SIMT recurrence; reference modules.py:28-65 under autograd, i.e. cuDNN's RNN training
 * forward and backward-data).  Forward = aero_lstm_rec_fwd (precision 0) that also saves, WINDOWED as [rows*n_win][steps][2][..],
 * the post-activation gates (i,f,g,o: 4H), the cell state c (H) and the hidden state h (H) of every step; hout is the
 * de-windowed output (may be NULL when out_windowed: h_s is that output).
 * Backward: dhout in the layout of the forward's output (windowed [n_seq][steps][2H] or de-windowed [rows][T][2H]);
 * dgin_w = gradient of the gate pre-activations, windowed [n_seq][steps][2][4H] (every position, padding frames included).
 * The caller finishes with GEMMs over dgin_w: bias = column sums, W_hh = aero_tapgemm_wgrad against h_s shifted by one step,
 * W_ih / input gradient from dgin (aero_lstm_fold sums the windowed rows back onto un-windowed frames for the first layer). */
int aero_lstm_train_fwd(const float* gin, const float* bias_pad, const float* whh, float* hout, float* gates_s, float* c_s,
                        float* h_s, const aero_lstm_params* p, aero_stream_t stream);
int aero_lstm_bwd(const float* dhout, const float* gates_s, const float* c_s, const float* whh, float* dgin_w,
                  const aero_lstm_params* p, aero_stream_t stream);
int aero_lstm_fold(const float* dgin_w, float* dgin, int32_t rows, int32_t T, int32_t n_win, int32_t steps, int32_t win_stride,
                   int32_t C, aero_stream_t stream);

/* LocalState attention, training form (reference modules.py:104-124 under autograd): exact-fp32 forward that also returns
 * lse[rows][heads][T] (log-sum-exp over keys per query), and the backward: dqkvd[rows][T][ld] (q | k | v | decay-logit columns,
 * every column written) from dout[rows][T][H]; scores are recomputed flash-style (no T x T tensor). */
int aero_local_attn_train_fwd(const float* qkvd, float* out, float* lse, const aero_attn_params* p, aero_stream_t stream);
int aero_local_attn_bwd(const float* qkvd, const float* out, const float* lse, const float* dout, float* dqkvd,
                        const aero_attn_params* p, aero_stream_t stream);

/* MelGAN multi-scale discriminator (reference src/models/discriminators.py:14-78): grouped strided
 * Conv1d (k <= 41, <= 8 input channels per group) on channels-last tensors x [B][Tin][Cin] -> y [B][Tout][Cout], weights in
 * PyTorch's layout w [Cout][Cin/groups][k]; Tout = (Tin + 2 pad - k) / stride + 1.  dgrad writes dx; wgrad ADDS into dw (caller zeroes).
 * The dense layers of the discriminator (k = 15 / 5 / 3) run on aero_tapgemm_fwd / aero_tapgemm_wgrad. */
int aero_gconv1d_fwd(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Tin, int32_t Tout, int32_t Cin,
                     int32_t Cout, int32_t groups, int32_t k, int32_t stride, int32_t pad, aero_stream_t stream);
int aero_gconv1d_dgrad(const float* dy, const float* w, float* dx, int32_t B, int32_t Tin, int32_t Tout, int32_t Cin, int32_t Cout,
                       int32_t groups, int32_t k, int32_t stride, int32_t pad, aero_stream_t stream);
int aero_gconv1d_wgrad(const float* x, const float* dy, float* dw, int32_t B, int32_t Tin, int32_t Tout, int32_t Cin, int32_t Cout,
                       int32_t groups, int32_t k, int32_t stride, int32_t pad, aero_stream_t stream);
/* Weight normalisation (torch.nn.utils.weight_norm, reference modules.py WNConv1d): w[r][:] = g[r] * v[r][:] / ||v[r][:]||, and its
 * backward, which ADDS  dg[r] += <dw[r], v[r]> / ||v[r]||  and  dv[r] += g[r]/||v[r]|| * (dw[r] - v[r] <dw[r], v[r]> / ||v[r]||^2). */
int aero_weight_norm_fwd(const float* v, const float* g, float* w, int32_t rows, int32_t len, aero_stream_t stream);
int aero_weight_norm_bwd(const float* v, const float* g, const float* dw, float* dv, float* dg, int32_t rows, int32_t len,
                         aero_stream_t stream);

/* Weight repack for the tensor-core training modes: w [taps][K][ldn] (the aero_tapgemm_fwd precision-0 layout, ldn = N rounded up to
 * 4) -> out [taps][ldn][K] (the precision-1 layout), every element rounded to TF32 (round to nearest, ties away); out_lo (optional) gets
 * TF32(w - out) in the same layout.
 * aero_split_tf32: the same two-term split of an activation tensor, hi = TF32(x), lo = TF32(x - hi).  Three TF32 tensor-core products
 * hi*hi + hi*lo + lo*hi, summed in fp32, reproduce the fp32 product to ~2^-22 ("3xTF32"): the training engine's fp32-grade tensor-core
 * mode issues aero_tapgemm_fwd / aero_tapgemm_wgrad three times on these halves (aero_b200/train_engine.py, precision 3).
 * aero_tapgemm_wgrad_tc_eligible: 1 when aero_tapgemm_wgrad with p->precision = 1 would run on the tensor cores. */
int aero_pack_kmajor_tf32(const float* w, float* out, float* out_lo, int32_t taps, int32_t K, int32_t ldn, aero_stream_t stream);
int aero_split_tf32(const float* x, float* hi, float* lo, int64_t n, aero_stream_t stream);
int aero_tapgemm_wgrad_tc_eligible(const aero_tapgemm_params* p, const float* a1, const float* a2, const float* dy);

/* Fused multi-tensor Adam (torch.optim.Adam semantics, no amsgrad / weight decay; reference train.py:83).  chunk_table: device
 * array of n_chunks records {float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64 count}; one CTA per
 * record.  step >= 1 is the step count AFTER this update (bias corrections 1 - beta^step); grad_scale multiplies every
 * gradient (1/world_size after a sum all-reduce). */
int aero_adam_step(const void* chunk_table, int32_t n_chunks, float lr, float beta1, float beta2, float eps, int32_t step,
                   float grad_scale, aero_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AERO_B200_H */
