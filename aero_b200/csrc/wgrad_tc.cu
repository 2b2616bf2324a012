// Weight gradient of the tap-GEMM on the tensor cores (sm_100a, kind::tf32): the training-side twin of tapgemm_tc.cu.
//
//   dW[slab][k][n] += sum over output pixels (b, fo, t) of  A(b, fi, t + dt, k) * dY(b, fo, t, n)
//
// is, per slab, a GEMM whose reduction axis is the pixel axis -- the axis that is NOT contiguous in the channels-last tensors.  Both
// operands are therefore MN-major for tcgen05 (cute Layout_MN_SW128_32B_Atom, the only MN-major form of tf32): a TMA box of
// 32 channels x 32 consecutive frames of one (b, f) row, written with SWIZZLE_128B_ATOM_32B, is 8 K-atoms (4 frames x 128 B) of one
// 32-channel M atom.  No transposition pass, no im2col: the tap shift is the box coordinate, borders and tails are TMA zero fill.
//
//   D (TMEM)  : 128 lanes = 128 output channels n, BN <= 256 fp32 columns = input channels k (32-channel boxes taken from the
//               concatenation of the two source tensors of a skip connection; a box never straddles them)
//   A (smem)  : dY tile  [32 frames][128 n]  = 4 boxes, 16 KB
//   B (smem)  : act tile [32 frames][BN k]   = BN/32 boxes
// One CTA owns one (slab, n-tile, k-tile) and a strided subset of the 32-frame chunks (split-K over pixels); warp 0 = TMA producer,
// warp 1 = MMA issue, warps 2..5 = epilogue: TMEM -> red.global.add.f32 into dW in the parameter's own layout (dW is zeroed by the
// caller; same contract as the SIMT kernel in train.cu).  The tensor core reads fp32 bit patterns and ignores the low 13 mantissa
// bits (truncation, as cuDNN's TF32 convolutions do): this is the precision-1 training mode, not the parity mode.
#include <cuda.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace aero {

constexpr int kWtThreads = 192;
constexpr int kWtMaxStages = 8;
constexpr int kWtChunk = 32;              // frames per pipeline stage (= 4 UMMAs of K = 8)
constexpr int kWtATile = 128 * 128;       // 4 boxes x 4 KB

struct WgradTcShared {
    uint64_t full[kWtMaxStages];
    uint64_t empty[kWtMaxStages];
    uint64_t acc_full;
    uint32_t tmem_base;
    int has_acc;
};

struct WgradTcArgs {
    float* dw;
    aero_tapgemm_params p;
    int64_t dw_sn, dw_sk, dw_ss;
    int tiles_t, n_tiles, k_tiles, nb1, nb, bpt, BN, splits, stages;   // nb1 / nb: 32-channel boxes of source 1 / of both; bpt: boxes per k-tile
    int d_tt, d_fo, d_b;                                               // `splits` chunks ahead, as (frame-chunk, row, batch) carries
    uint32_t idesc, tmem_cols;
};

// Walks the 32-frame chunks split, split + splits, ... of the (b, fo, frame-chunk) space with carries only: the walkers are single
// threads (TMA producer, MMA issuer) whose every instruction is on the critical path -- a 64-bit division per chunk costs more than the
// chunk's four UMMAs.
struct WtWalker {
    int b, fo, tt;
    __device__ __forceinline__ void init(const WgradTcArgs& g, int split) {
        tt = split % g.tiles_t;
        const int row = split / g.tiles_t;
        b = row / g.p.F_out;
        fo = row - b * g.p.F_out;
    }
    __device__ __forceinline__ bool done(const WgradTcArgs& g) const { return b >= g.p.B; }
    __device__ __forceinline__ void next(const WgradTcArgs& g) {
        tt += g.d_tt;
        int carry = 0;
        if (tt >= g.tiles_t) { tt -= g.tiles_t; carry = 1; }
        fo += g.d_fo + carry;
        carry = 0;
        if (fo >= g.p.F_out) { fo -= g.p.F_out; carry = 1; }
        b += g.d_b + carry;
    }
    // input row of this slab for the current output row; false when the tap has none
    __device__ __forceinline__ bool input_row(const aero_tapgemm_params& p, int jf, int r, int tapi, int& fi) const {
        if (p.mode == AERO_TAPS_CONV) {
            fi = fo * p.stride_f + jf - p.pad_f;
        } else {
            const int fof = fo + p.f_out_offset;
            const int qf = fof / p.stride_f;
            if (fof - qf * p.stride_f != r) return false;
            fi = qf - tapi;
        }
        return fi >= 0 && fi < p.F_in;
    }
};

__global__ void __launch_bounds__(kWtThreads)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
                const __grid_constant__ CUtensorMap mapDy, const WgradTcArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const aero_tapgemm_params& p = g.p;
    const int stage_bytes = kWtATile + (g.BN / 32) * 4096;
    WgradTcShared* sh = reinterpret_cast<WgradTcShared*>(smem + g.stages * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work item of this CTA
    const int kt_i = blockIdx.x % g.k_tiles, nt_i = blockIdx.x / g.k_tiles;
    const int slab = blockIdx.y, split = blockIdx.z;
    const int box0 = kt_i * g.bpt;                                // this k-tile: boxes [box0, box0 + nbox) of the concatenated sources
    const int nbox = min(g.bpt, g.nb - box0);
    const int n0 = nt_i * 128;
    int jf = 0, dt = 0, r = 0, tapi = 0;
    if (p.mode == AERO_TAPS_CONV) {
        jf = slab / p.kt;
        dt = (slab - jf * p.kt) * p.dil_t - p.pad_t;
    } else {
        r = slab % p.stride_f;
        tapi = slab / p.stride_f;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) { mbar_init(&sh->full[s], 1); mbar_init(&sh->empty[s], 1); }
        mbar_init(&sh->acc_full, 1);
        sh->has_acc = 1;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(g.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = sh->tmem_base;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA1) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapDy) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = (uint32_t)(kWtATile + nbox * 4096);
            WtWalker w;
            for (w.init(g, split); !w.done(g); w.next(g)) {
                int fi;
                if (!w.input_row(p, jf, r, tapi, fi)) continue;
                const int t0 = w.tt * kWtChunk;
                mbar_wait(&sh->empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * stage_bytes;
                mbar_expect_tx(&sh->full[stage], tx);
#pragma unroll
                for (int j = 0; j < 4; ++j) tma_load_4d(sa + j * 4096, &mapDy, &sh->full[stage], n0 + 32 * j, t0, w.fo, w.b);
                for (int j = 0; j < nbox; ++j) {
                    const int gb = box0 + j;
                    const bool s2 = gb >= g.nb1;
                    tma_load_4d(sa + kWtATile + j * 4096, s2 ? &mapA2 : &mapA1, &sh->full[stage], 32 * (s2 ? gb - g.nb1 : gb), t0 + dt, fi, w.b);
                }
                if (++stage == g.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        int stage = 0;
        uint32_t phase = 0;
        int iters = 0;
        WtWalker w;
        for (w.init(g, split); !w.done(g); w.next(g)) {
            int fi_unused;
            if (!w.input_row(p, jf, r, tapi, fi_unused)) continue;         // same enumeration as the producer
            mbar_wait(&sh->full[stage], phase);
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t sa = smem_u32(smem + stage * stage_bytes);
                // MN-major tf32 operands (see tapgemm_tc.cu, the frequency-mix mode): LBO = 4096 B between 32-channel atoms (boxes),
                // SBO = 512 B between 4-frame K atoms, layout SWIZZLE_128B_BASE32B; one UMMA (K = 8 frames) = two K atoms = 1024 B
                const uint64_t da = (uint64_t)((sa >> 4) & 0x3FFF) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
                                    ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
                const uint32_t sb = sa + kWtATile;
                const uint64_t db = (uint64_t)((sb >> 4) & 0x3FFF) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
                                    ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_tf32(tmem_base, da + (uint64_t)(k * (1024 >> 4)), db + (uint64_t)(k * (1024 >> 4)), g.idesc, (iters > 0 || k > 0) ? 1u : 0u);
                umma_commit(&sh->empty[stage]);
            }
            __syncwarp();
            ++iters;
            if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) {
            if (iters > 0) {
                umma_commit(&sh->acc_full);
            } else {
                sh->has_acc = 0;
                mbar_arrive(&sh->acc_full);
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue (warps 2..5): TMEM -> atomics on dW
        const int q = warp & 3;                                // TMEM lane quarter of this warp
        const int n = n0 + q * 32 + lane;
        mbar_wait(&sh->acc_full, 0);
        tcgen05_fence_after();
        if (sh->has_acc) {
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16);
            float* dst_n = g.dw + (int64_t)n * g.dw_sn + (int64_t)slab * g.dw_ss;
            for (int j = 0; j < nbox; ++j) {                   // one 32-column block per box
                const int gb = box0 + j;
                const bool s2 = gb >= g.nb1;
                const int ch0 = 32 * (s2 ? gb - g.nb1 : gb);
                const int cnt = min(32, (s2 ? p.C2 : p.C1) - ch0);
                const int kbase = (s2 ? p.C1 : 0) + ch0;
                uint32_t v[32];
                tmem_ld32(tacc + (uint32_t)(32 * j), v);
                if (n < p.N) {
#pragma unroll
                    for (int u = 0; u < 32; ++u)
                        if (u < cnt) atomicAdd(dst_n + (int64_t)(kbase + u) * g.dw_sk, __uint_as_float(v[u]));
                }
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(g.tmem_cols) : "memory");
    }
}

static int wt_map(CUtensorMap* m, const void* base, int C, int T, int F, int B, int64_t sb, int64_t sf, int64_t st) {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)T, (uint64_t)F, (uint64_t)B};
    int64_t s1 = st, s2 = sf, s3 = sb;
    if (s2 <= 0) s2 = s1 * T;                   // size-1 dimensions: any legal stride
    if (s3 <= 0) s3 = s2 * F;
    uint64_t strides[3] = {(uint64_t)s1 * 4, (uint64_t)s2 * 4, (uint64_t)s3 * 4};
    uint32_t box[4] = {32, (uint32_t)kWtChunk, 1, 1};
    return encode_map(m, base, 4, dims, strides, box, 1, 4);
}

bool wgrad_tc_eligible(const aero_tapgemm_params& p, const void* a1, const void* a2, const void* dy) {
    if (p.mode != AERO_TAPS_CONV && p.mode != AERO_TAPS_CONVT) return false;
    if (p.N < 16 || p.N % 4 || p.C1 % 4 || p.C2 % 4 || p.C1 + p.C2 < 16) return false;
    auto ok = [](int64_t sb, int64_t sf, int64_t st) { return sb % 4 == 0 && sf % 4 == 0 && st % 4 == 0 && st > 0; };
    if (p.C1 && !ok(p.a1_sb, p.a1_sf, p.a1_st)) return false;
    if (p.C2 && !ok(p.a2_sb, p.a2_sf, p.a2_st)) return false;
    if (!ok(p.o_sb, p.o_sf, p.o_st)) return false;
    if (((uintptr_t)a1 | (uintptr_t)a2 | (uintptr_t)dy) & 15) return false;
    return true;
}

int wgrad_tc_launch(const float* a1, const float* a2, const float* dy, float* dw, const aero_tapgemm_params& p, int64_t dw_sn, int64_t dw_sk,
                    int64_t dw_ss, cudaStream_t st) {
    WgradTcArgs g;
    g.dw = dw; g.p = p; g.dw_sn = dw_sn; g.dw_sk = dw_sk; g.dw_ss = dw_ss;
    g.tiles_t = cdiv(p.T, kWtChunk);
    g.n_tiles = cdiv(p.N, 128);
    // B tile = up to eight 32-channel boxes taken from the concatenation of the two sources (a box never straddles them)
    g.nb1 = cdiv(p.C1, 32);
    g.nb = g.nb1 + cdiv(p.C2, 32);
    g.k_tiles = cdiv(g.nb, 8);
    g.bpt = cdiv(g.nb, g.k_tiles);
    g.BN = 32 * g.bpt;
    CUtensorMap mA1, mA2, mDy;
    int rc;
    if (p.C1 && (rc = wt_map(&mA1, a1, p.C1, p.T_in, p.F_in, p.B, p.a1_sb, p.a1_sf, p.a1_st)) != AERO_OK) return rc;
    if (p.C2 && (rc = wt_map(&mA2, a2, p.C2, p.T_in, p.F_in, p.B, p.a2_sb, p.a2_sf, p.a2_st)) != AERO_OK) return rc;
    if (!p.C1) mA1 = mA2;
    if (!p.C2) mA2 = mA1;
    if ((rc = wt_map(&mDy, dy, p.N, p.T, p.F_out, p.B, p.o_sb, p.o_sf, p.o_st)) != AERO_OK) return rc;
    g.tmem_cols = 32;
    while ((int)g.tmem_cols < g.BN) g.tmem_cols <<= 1;
    // instruction descriptor: D = F32, A / B = TF32, both MN-major (bits 15, 16), N = BN, M = 128
    g.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(g.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int stage_bytes = kWtATile + (g.BN / 32) * 4096;
    const int fixed = (int)sizeof(WgradTcShared) + 1024;
    g.stages = (226 * 1024 - fixed) / stage_bytes;
    if (g.stages > kWtMaxStages) g.stages = kWtMaxStages;
    if (g.stages < 2) g.stages = 2;
    const size_t smem = (size_t)g.stages * stage_bytes + fixed;
    const int nslab = (p.mode == AERO_TAPS_CONVT) ? p.kf : p.kf * p.kt;
    const int64_t n_chunks = (int64_t)p.B * p.F_out * g.tiles_t;
    const int64_t items = (int64_t)g.k_tiles * g.n_tiles * nslab;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // split the pixel axis: one CTA per SM is resident (the pipeline takes the shared memory), so the grid should be a whole number of
    // waves.  Among the split counts that give 2 .. 6 waves pick the one with the fullest last wave; a split keeps at least 8 chunks so
    // that the atomics stay a small fraction of the work.
    const int64_t max_splits = n_chunks / 8 > 0 ? n_chunks / 8 : 1;
    int64_t lo = cdiv((int64_t)2 * num_sms, items), hi = cdiv((int64_t)6 * num_sms, items);
    if (lo > max_splits) lo = max_splits;
    if (hi > max_splits) hi = max_splits;
    if (hi > 65535) hi = 65535;
    if (lo < 1) lo = 1;
    if (hi < lo) hi = lo;
    int64_t splits = lo;
    double best_eff = 0.0;
    for (int64_t sp = lo; sp <= hi; ++sp) {
        const int64_t ctas = items * sp;
        const double eff = (double)ctas / (double)(cdiv(ctas, (int64_t)num_sms) * num_sms);
        if (eff > best_eff + 0.02) { best_eff = eff; splits = sp; }      // prefer fewer splits unless clearly fuller
    }
    g.splits = (int)splits;
    g.d_tt = g.splits % g.tiles_t;
    const int d_row = g.splits / g.tiles_t;
    g.d_fo = d_row % p.F_out;
    g.d_b = d_row / p.F_out;
    if (nslab > 65535 || items / nslab > 2147483647LL) { set_error("aero_tapgemm_wgrad(tcgen05): grid too large"); return AERO_ERR_INVALID; }
    cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    dim3 grid((unsigned)(items / nslab), (unsigned)nslab, (unsigned)splits);
    wgrad_tc_kernel<<<grid, kWtThreads, smem, st>>>(mA1, mA2, mDy, g);
    return check_launch("aero_tapgemm_wgrad(tcgen05)");
}

}  // namespace aero
