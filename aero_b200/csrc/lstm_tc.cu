// Persistent BiLSTM recurrence on tcgen05 (sm_100a).  See include/aero_b200.h (aero_lstm_rec_fwd,
// precision = 1).
//
// Per step the recurrent term is the small GEMM  D[gate rows x sequences] = W_hh[gate rows x H] * h^T[H x sequences]:
//   A = W_hh, loaded ONCE by TMA into shared memory (K-major, SWIZZLE_128B) and resident for all steps;
//       rows are re-ordered so that TMEM lane r of M-tile m holds (cell 32m + r/4, gate r%4): the four gates
//       of a cell sit in four adjacent lanes of one warp and are combined with quad shuffles;
//   B = h of the previous step, written by the cell-update threads straight into the swizzled
//       shared-memory operand layout (TF32-rounded), 16 sequences per CTA;
//   D = nM x 16 fp32 columns of TMEM, read back with tcgen05.ld by the same threads.
// The gate pre-activations of the input projection (computed by the tap-GEMM in the same re-ordered
// column order, so a warp reads 128 contiguous bytes per sequence) are prefetched while the MMA runs.
// One elected thread issues the MMAs; a pair of mbarriers ping-pongs between "h ready" and
// "accumulators ready".  c stays in registers for the whole sequence.
#include "tc_common.cuh"

namespace aero {

constexpr int kNT = 16;          // sequences per CTA (UMMA N)

struct LstmTcShared {
    uint64_t w_full;
    uint64_t acc_ready;
    uint64_t h_ready;
    uint32_t tmem_base;
};

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 2.0f * fast_sigmoid(2.0f * x) - 1.0f; }

template <int NM>   // number of 128-row M tiles = ceil(H / 32)
__global__ void __launch_bounds__(192, (NM == 3 ? 1 : 2))
lstm_tc_kernel(const __grid_constant__ CUtensorMap mapW, const float* __restrict__ gin, const float* __restrict__ bias_pad,
               float* __restrict__ hout, const aero_lstm_params p, const int nK) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                  // [NM][nK] tiles of 128 rows x 128 B
    uint8_t* sB = smem + NM * nK * 16384;                // [nK] tiles of 16 rows x 128 B
    LstmTcShared* sh = reinterpret_cast<LstmTcShared*>(sB + nK * 2048);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const int seq0 = blockIdx.x * kNT;
    const int n_seq = p.rows * p.n_win;
    const int H = p.H;
    const int ldg = 2 * NM * 128;                        // floats per gin row (both directions, padded)

    if (threadIdx.x == 0) {
        mbar_init(&sh->w_full, 1);
        mbar_init(&sh->acc_ready, 1);
        mbar_init(&sh->h_ready, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < nK * 2048 / 4; i += blockDim.x) reinterpret_cast<float*>(sB)[i] = 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = sh->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
            mbar_expect_tx(&sh->w_full, (uint32_t)(NM * nK * 16384));
            for (int m = 0; m < NM; ++m)
                for (int kc = 0; kc < nK; ++kc)
                    tma_load_2d(sA + (m * nK + kc) * 16384, &mapW, &sh->w_full, kc * 32, (dir * NM + m) * 128);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // UMMA instruction descriptor: D=F32, A=B=TF32, K-major, N=16, M=128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kNT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            mbar_wait(&sh->w_full, 0);
            for (int s = 1; s < p.steps; ++s) {
                mbar_wait(&sh->h_ready, (uint32_t)((s - 1) & 1));
                tcgen05_fence_after();
                for (int m = 0; m < NM; ++m) {
                    for (int kc = 0; kc < nK; ++kc) {
                        const uint64_t da = make_desc_sw128(smem_u32(sA + (m * nK + kc) * 16384));
                        const uint64_t db = make_desc_sw128(smem_u32(sB + kc * 2048));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_tf32(tmem_base + (uint32_t)(m * kNT), da + 2 * k, db + 2 * k, idesc, (kc > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&sh->acc_ready);
            }
        }
    } else {
        // ===================================================== cell update (warps 2..5, 128 threads)
        const int q = warp & 3;
        const int r = q * 32 + lane;                     // TMEM lane inside an M tile
        const int gate = lane & 3;
        const int c_local = r >> 2;
        const int half = p.win_stride / 2;

        // per-sequence addressing (identical for all lanes)
        int seq_row[kNT], seq_k[kNT];
#pragma unroll
        for (int n = 0; n < kNT; ++n) {
            const int s = min(seq0 + n, n_seq - 1);
            seq_row[n] = s / p.n_win;
            seq_k[n] = s - seq_row[n] * p.n_win;
        }
        float c_state[NM][kNT];
#pragma unroll
        for (int m = 0; m < NM; ++m)
#pragma unroll
            for (int n = 0; n < kNT; ++n) c_state[m][n] = 0.f;

        for (int s = 0; s < p.steps; ++s) {
            const int pos = dir ? p.steps - 1 - s : s;
            // ---- prefetch the input-projection gate pre-activations (coalesced: lane r is contiguous)
            float gi[NM][kNT];
#pragma unroll
            for (int n = 0; n < kNT; ++n) {
                const float* src;
                if (p.in_windowed) {
                    src = gin + ((int64_t)min(seq0 + n, n_seq - 1) * p.steps + pos) * ldg;
                } else {
                    const int frame = seq_k[n] * p.win_stride + pos;
                    src = frame < p.T ? gin + ((int64_t)seq_row[n] * p.T + frame) * ldg : bias_pad;
                }
#pragma unroll
                for (int m = 0; m < NM; ++m) gi[m][n] = src[(dir * NM + m) * 128 + r];
            }
            if (s > 0) {
                mbar_wait(&sh->acc_ready, (uint32_t)((s - 1) & 1));
                tcgen05_fence_after();
            }
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                uint32_t acc[16];
                if (s > 0) {
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * kNT), acc);
                } else {
#pragma unroll
                    for (int n = 0; n < kNT; ++n) acc[n] = 0u;
                }
                const int cell = m * 32 + c_local;
                const bool cell_ok = cell < H;
#pragma unroll
                for (int n = 0; n < kNT; ++n) {
                    const float x = __uint_as_float(acc[n]) + gi[m][n];
                    const float sg = fast_sigmoid(gate == 2 ? 2.0f * x : x);
                    const float a = gate == 2 ? 2.0f * sg - 1.0f : sg;
                    const int base = lane & ~3;
                    const float fg = __shfl_sync(0xffffffffu, a, base + 1);
                    const float gg = __shfl_sync(0xffffffffu, a, base + 2);
                    const float og = __shfl_sync(0xffffffffu, a, base + 3);
                    if (gate == 0) {
                        const float c = fg * c_state[m][n] + a * gg;
                        c_state[m][n] = c;
                        const float h = round_tf32_rna(og * fast_tanh(c));
                        if (cell_ok) {
                            // B operand tile kc = cell/32: row n (sequence), 16-byte chunks XOR-swizzled by (row % 8)
                            const int kc = cell >> 5, j = cell & 31;
                            const uint32_t off = (uint32_t)(kc * 2048 + (n >> 3) * 1024 + (n & 7) * 128 + ((((j >> 2) ^ (n & 7)) << 4) | ((j & 3) << 2)));
                            *reinterpret_cast<float*>(sB + off) = h;
                            if (seq0 + n < n_seq) {
                                if (p.out_windowed) {
                                    hout[((int64_t)(seq0 + n) * p.steps + pos) * 2 * H + dir * H + cell] = h;
                                } else {
                                    const int frame = seq_k[n] * p.win_stride + pos;
                                    const int lo = (seq_k[n] == 0) ? 0 : half;
                                    const int hi = (seq_k[n] == p.n_win - 1) ? p.steps : p.steps - half;
                                    if (pos >= lo && pos < hi && frame < p.T)
                                        hout[((int64_t)seq_row[n] * p.T + frame) * 2 * H + dir * H + cell] = h;
                                }
                            }
                        }
                    }
                }
            }
            if (s + 1 < p.steps) {
                fence_proxy_async_smem();                // generic-proxy stores of h -> visible to the tensor core
                tcgen05_fence_before();
                mbar_arrive(&sh->h_ready);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
    }
}

int lstm_tc_launch(const float* gin, const float* bias_pad, const float* whh_r, float* hout, const aero_lstm_params& p,
                   cudaStream_t st) {
    const int H = p.H;
    const int nM = (H + 31) / 32, nK = (H + 31) / 32;
    if (H % 4 || nM > 3) {
        set_error("aero_lstm_rec_fwd(tcgen05): hidden size %d unsupported (multiple of 4, <= 96)", H);
        return AERO_ERR_UNSUPPORTED;
    }
    CUtensorMap mW;
    uint64_t dims[2] = {(uint64_t)H, (uint64_t)(2 * nM * 128)};
    uint64_t strides[1] = {(uint64_t)H * 4};
    uint32_t box[2] = {32, 128};
    int rc = encode_map(&mW, whh_r, 2, dims, strides, box);
    if (rc != AERO_OK) return rc;
    const size_t smem = (size_t)nM * nK * 16384 + (size_t)nK * 2048 + sizeof(LstmTcShared) + 1024;
    const int n_seq = p.rows * p.n_win;
    dim3 grid(cdiv(n_seq, kNT), 2);
    switch (nM) {
        case 1:
            cudaFuncSetAttribute(lstm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            lstm_tc_kernel<1><<<grid, 192, smem, st>>>(mW, gin, bias_pad, hout, p, nK);
            break;
        case 2:
            cudaFuncSetAttribute(lstm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            lstm_tc_kernel<2><<<grid, 192, smem, st>>>(mW, gin, bias_pad, hout, p, nK);
            break;
        default:
            cudaFuncSetAttribute(lstm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            lstm_tc_kernel<3><<<grid, 192, smem, st>>>(mW, gin, bias_pad, hout, p, nK);
            break;
    }
    return check_launch("aero_lstm_rec_fwd(tcgen05)");
}

}  // namespace aero
