// Register-tiled tensor-product kernels for small irreps (l_max <= 2 layers, diagonal last layer).
//
// Work unit = (centre atom c, chunk of 32 channels): ONE WARP, lane = channel u.
//   1. M_c[u][i][k] = sum_nnz cgw[nnz][u] * gamma[c][j][u]   -- the CG contraction with the centre's
//      environment is done ONCE per centre (83 FMAs at l_max=2) instead of once per edge: every
//      edge of the centre shares it (edge_center = edge_index[0], _allegro.py:238; gather of
//      the scattered sum, _contract.py:205).  Built in shared memory (table indices are
//      run-time data), then held in registers (D_IN*D_OUT <= 81).
//   2. stream the centre's edges (a contiguous CSR row): Vout[z][k][u] = sum_i Vin[z][i][u] M[i][k]
//      -- 2*D coalesced 128-byte (fp32) / 64-byte (bf16) warp accesses per edge, D_IN*D_OUT FMAs.
// Backward additionally accumulates gM[i][k] = sum_z Vin[z][i] gVout[z][k] in registers over the
// row and contracts it with the table once per centre -> ggamma[c][j][u] written exactly once
// (no atomics, no memset, deterministic).  For layer 0 the input features
// Vin = Y (x) w0 (tensorembed.py:95) are formed on the fly and the w0 / Y gradients are produced
// in the same pass (Y gradient: warp-shuffle reduction over channels).
#include "common.cuh"
#include "tp_fast.cuh"

namespace {

template <typename TAcc, int D_IN, int D_OUT, int DG>
struct WarpSmem {
    TAcc M[D_IN * D_OUT][32];
    TAcc G[DG][32];
};

// Build M in shared memory and copy to registers.
template <typename TAcc, int D_IN, int D_OUT, int DG>
__device__ __forceinline__ void build_M(TAcc (&M)[D_IN][D_OUT], WarpSmem<TAcc, D_IN, D_OUT, DG>& sm, int lane, int nnz,
                                        const int32_t* __restrict__ tabp, const TAcc* __restrict__ cgw, int U, int u, bool live,
                                        const TAcc* __restrict__ gam_c /* gamma + c*D*U */) {
    // (i,k)-sorted table: walk it once; each target's segment is accumulated in a register and stored
    // once (no shared-memory read-modify-write chain), then M is pulled into registers.
#pragma unroll
    for (int e = 0; e < D_IN * D_OUT; ++e) sm.M[e][lane] = TAcc(0);
    if (live && nnz > 0) {
        int t_prev = tabp[0] * D_OUT + tabp[2];
        TAcc acc = TAcc(0);
        for (int n = 0; n < nnz; ++n) {
            const int t = tabp[3 * n] * D_OUT + tabp[3 * n + 2];
            if (t != t_prev) {
                sm.M[t_prev][lane] = acc;
                acc = TAcc(0);
                t_prev = t;
            }
            acc += cgw[(int64_t)n * U + u] * gam_c[(int64_t)tabp[3 * n + 1] * U + u];
        }
        sm.M[t_prev][lane] = acc;
    }
#pragma unroll
    for (int i = 0; i < D_IN; ++i)
#pragma unroll
        for (int k = 0; k < D_OUT; ++k) M[i][k] = sm.M[i * D_OUT + k][lane];
}

template <typename TAct, typename TAcc, int D_IN, bool IMPLICIT>
__device__ __forceinline__ void load_vin(TAcc (&v)[D_IN], TAcc (&w0l)[5], TAcc (&Yz)[D_IN], int64_t z, int U, int u, bool live,
                                         const TAct* __restrict__ Vin, const TAcc* __restrict__ Y, const TAct* __restrict__ w0,
                                         int64_t w0_ld) {
    if constexpr (IMPLICIT) {
#pragma unroll
        for (int i = 0; i < D_IN; ++i) Yz[i] = Y[z * D_IN + i];
#pragma unroll
        for (int l = 0; l * l < D_IN; ++l) w0l[l] = live ? to_acc<TAcc>(w0[z * w0_ld + l * U + u]) : TAcc(0);
#pragma unroll
        for (int i = 0; i < D_IN; ++i) v[i] = Yz[i] * w0l[sh_l_of(i)];
    } else {
#pragma unroll
        for (int i = 0; i < D_IN; ++i) v[i] = live ? to_acc<TAcc>(Vin[(z * D_IN + i) * U + u]) : TAcc(0);
    }
}

template <typename TAct, typename TAcc, int D_IN, int D_OUT, int DG, bool IMPLICIT>
__global__ void __launch_bounds__(128) tp_fwd_fast_kernel(int64_t N, int U, int D, int nnz, const int32_t* __restrict__ tabp,
                                                          const TAcc* __restrict__ cgw, const int32_t* __restrict__ row_ptr,
                                                          const TAcc* __restrict__ gamma, const TAct* __restrict__ Vin,
                                                          const TAcc* __restrict__ Y, const TAct* __restrict__ w0, int64_t w0_ld,
                                                          TAct* __restrict__ Vout) {
    __shared__ WarpSmem<TAcc, D_IN, D_OUT, DG> smem[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunk = (U + 31) >> 5;
    const int64_t wid = (int64_t)blockIdx.x * 4 + warp;
    const int64_t c = wid / nchunk;
    if (c >= N) return;
    const int u = (int)(wid % nchunk) * 32 + lane;
    const bool live = u < U;
    const int beg = row_ptr[c], end = row_ptr[c + 1];
    if (beg == end) return;
    TAcc M[D_IN][D_OUT];
    build_M<TAcc, D_IN, D_OUT, DG>(M, smem[warp], lane, nnz, tabp, cgw, U, u, live, gamma + c * D * U);
#pragma unroll(IMPLICIT ? 4 : 2)
    for (int64_t z = beg; z < end; ++z) {
        TAcc v[D_IN], w0l[5], Yz[D_IN];
        load_vin<TAct, TAcc, D_IN, IMPLICIT>(v, w0l, Yz, z, U, u, live, Vin, Y, w0, w0_ld);
        TAcc out[D_OUT];
#pragma unroll
        for (int k = 0; k < D_OUT; ++k) out[k] = TAcc(0);
#pragma unroll
        for (int i = 0; i < D_IN; ++i)
#pragma unroll
            for (int k = 0; k < D_OUT; ++k) out[k] += v[i] * M[i][k];
        if (live) {
#pragma unroll
            for (int k = 0; k < D_OUT; ++k) Vout[(z * D_OUT + k) * U + u] = from_acc<TAct>(out[k]);
        }
    }
}

template <typename TAct, typename TAcc, int D_IN, int D_OUT, int DG, bool IMPLICIT, bool GM_ONLY>
__global__ void __launch_bounds__(128, GM_ONLY ? 3 : 1) tp_bwd_fast_kernel(int64_t N, int U, int D, int nnz, const int32_t* __restrict__ tabp,
                                                          const TAcc* __restrict__ cgw, const int32_t* __restrict__ row_ptr,
                                                          const TAcc* __restrict__ gamma, const TAct* __restrict__ Vin,
                                                          const TAcc* __restrict__ Y, const TAct* __restrict__ w0, int64_t w0_ld,
                                                          const TAct* __restrict__ gVout, TAct* __restrict__ gVin,
                                                          TAct* __restrict__ gw0, int64_t gw0_ld, TAcc* __restrict__ gY,
                                                          TAcc* __restrict__ ggamma) {
    __shared__ WarpSmem<TAcc, D_IN, D_OUT, DG> smem[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunk = (U + 31) >> 5;
    const int64_t wid = (int64_t)blockIdx.x * 4 + warp;
    const int64_t c = wid / nchunk;
    if (c >= N) return;
    const int u = (int)(wid % nchunk) * 32 + lane;
    const bool live = u < U;
    const int beg = row_ptr[c], end = row_ptr[c + 1];
    WarpSmem<TAcc, D_IN, D_OUT, DG>& sm = smem[warp];
    TAcc M[GM_ONLY ? 1 : D_IN][GM_ONLY ? 1 : D_OUT], gM[D_IN][D_OUT];
    if constexpr (!GM_ONLY) build_M<TAcc, D_IN, D_OUT, DG>(M, sm, lane, nnz, tabp, cgw, U, u, live, gamma + c * D * U);
#pragma unroll
    for (int i = 0; i < D_IN; ++i)
#pragma unroll
        for (int k = 0; k < D_OUT; ++k) gM[i][k] = TAcc(0);
#pragma unroll(GM_ONLY ? 2 : 1)
    for (int64_t z = beg; z < end; ++z) {
        TAcc v[D_IN], w0l[5], Yz[D_IN], go[D_OUT];
        load_vin<TAct, TAcc, D_IN, IMPLICIT>(v, w0l, Yz, z, U, u, live, Vin, Y, w0, w0_ld);
#pragma unroll
        for (int k = 0; k < D_OUT; ++k) go[k] = live ? to_acc<TAcc>(gVout[(z * D_OUT + k) * U + u]) : TAcc(0);
        if constexpr (GM_ONLY) {
#pragma unroll
            for (int i = 0; i < D_IN; ++i)
#pragma unroll
                for (int k = 0; k < D_OUT; ++k) gM[i][k] += v[i] * go[k];
            continue;
        }
        TAcc gin[D_IN];
#pragma unroll
        for (int i = 0; i < D_IN; ++i) {
            TAcc s = TAcc(0);
#pragma unroll
            for (int k = 0; k < D_OUT; ++k) {
                if constexpr (!GM_ONLY) s += M[i][k] * go[k];
                gM[i][k] += v[i] * go[k];
            }
            gin[i] = s;
        }
        if constexpr (IMPLICIT) {
            // Vin[i] = Y[i] w0[l(i)]:  gw0[l] = sum_{i in l} Y[i] gin[i];  gY[i] += sum_u w0[l(i)][u] gin[i]
#pragma unroll
            for (int l = 0; l * l < D_IN; ++l) {
                TAcc s = TAcc(0);
#pragma unroll
                for (int i = l * l; i < (l + 1) * (l + 1); ++i) s += Yz[i] * gin[i];
                if (live) gw0[z * gw0_ld + l * U + u] = from_acc<TAct>(s);
            }
#pragma unroll
            for (int i = 0; i < D_IN; ++i) {
                const TAcc s = warp_sum(w0l[sh_l_of(i)] * gin[i]);
                if (lane == i) {
                    if (nchunk == 1) gY[z * D_IN + i] += s;  // single writer per (z, i)
                    else atomicAdd(&gY[z * D_IN + i], s);
                }
            }
        } else {
            if (live) {
#pragma unroll
                for (int i = 0; i < D_IN; ++i) gVin[(z * D_IN + i) * U + u] = from_acc<TAct>(gin[i]);
            }
        }
    }
    // ggamma[c][j][u] = sum_nnz cgw[nnz][u] * gM[i][k]   (adjoint of build_M)
#pragma unroll
    for (int i = 0; i < D_IN; ++i)
#pragma unroll
        for (int k = 0; k < D_OUT; ++k) sm.M[i * D_OUT + k][lane] = gM[i][k];
    for (int j = 0; j < D; ++j) sm.G[j][lane] = TAcc(0);
    if (live) {
        for (int n = 0; n < nnz; ++n) {
            const int i = tabp[3 * n], j = tabp[3 * n + 1], k = tabp[3 * n + 2];
            sm.G[j][lane] += cgw[(int64_t)n * U + u] * sm.M[i * D_OUT + k][lane];
        }
        for (int j = 0; j < D; ++j) ggamma[(c * D + j) * U + u] = sm.G[j][lane];
    }
}

// Backward part B with the gM accumulators split over NS warps of one CTA (warp s owns input
// components i in [s*IW, (s+1)*IW)): IW*D_OUT accumulators per thread instead of D_IN*D_OUT, so
// ~60 registers and 3-4x the resident warps of the monolithic kernel.  One CTA = one (centre,
// 32-channel chunk); the CG contraction gM -> ggamma is done once by warp 0 after a block barrier.
template <typename TAct, typename TAcc, int D_IN, int D_OUT, int DG, bool IMPLICIT, int NS>
__global__ void __launch_bounds__(NS * 32, 8) tp_bwd_gm_split_kernel(int64_t N, int U, int D, int nnz, const int32_t* __restrict__ tabp,
                                                                     const TAcc* __restrict__ cgw, const int32_t* __restrict__ row_ptr,
                                                                     const TAct* __restrict__ Vin, const TAcc* __restrict__ Y,
                                                                     const TAct* __restrict__ w0, int64_t w0_ld,
                                                                     const TAct* __restrict__ gVout, TAcc* __restrict__ ggamma) {
    // explicit Vin: warp s owns input components i in [s*IW, (s+1)*IW) (Vin read once, gVout by all warps);
    // implicit Vin = Y (x) w0 (cheap to rebuild): warp s owns OUTPUT components k instead, so the large
    // gVout tensor is read exactly once and only the small w0 rows are re-read.
    constexpr bool KSPLIT = IMPLICIT && (D_OUT % NS == 0);
    constexpr int IW = KSPLIT ? D_IN : D_IN / NS;
    constexpr int KW = KSPLIT ? D_OUT / NS : D_OUT;
    static_assert(KSPLIT || IW * NS == D_IN, "D_IN must be divisible by the split");
    __shared__ TAcc sGM[D_IN * D_OUT][32];
    __shared__ TAcc sG[DG][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunk = (U + 31) >> 5;
    const int64_t c = blockIdx.x / nchunk;
    const int u = (int)(blockIdx.x % nchunk) * 32 + lane;
    const bool live = u < U;
    const int beg = row_ptr[c], end = row_ptr[c + 1];
    const int i0 = KSPLIT ? 0 : warp * IW;
    const int k0 = KSPLIT ? warp * KW : 0;
    TAcc gM[IW][KW];
#pragma unroll
    for (int i = 0; i < IW; ++i)
#pragma unroll
        for (int k = 0; k < KW; ++k) gM[i][k] = TAcc(0);
#pragma unroll 2
    for (int64_t z = beg; z < end; ++z) {
        TAcc go[KW], v[IW];
#pragma unroll
        for (int k = 0; k < KW; ++k) go[k] = live ? to_acc<TAcc>(gVout[(z * D_OUT + k0 + k) * U + u]) : TAcc(0);
        if constexpr (IMPLICIT) {
            TAcc w0l[5];
#pragma unroll
            for (int l = 0; l * l < D_IN; ++l) w0l[l] = TAcc(0);
#pragma unroll
            for (int l = 0; l * l < D_IN; ++l)
                if (KSPLIT || (l * l < i0 + IW && (l + 1) * (l + 1) > i0)) w0l[l] = live ? to_acc<TAcc>(w0[z * w0_ld + l * U + u]) : TAcc(0);
#pragma unroll
            for (int i = 0; i < IW; ++i) v[i] = Y[z * D_IN + i0 + i] * w0l[sh_l_of(i0 + i)];
        } else {
#pragma unroll
            for (int i = 0; i < IW; ++i) v[i] = live ? to_acc<TAcc>(Vin[(z * D_IN + i0 + i) * U + u]) : TAcc(0);
        }
#pragma unroll
        for (int i = 0; i < IW; ++i)
#pragma unroll
            for (int k = 0; k < KW; ++k) gM[i][k] += v[i] * go[k];
    }
#pragma unroll
    for (int i = 0; i < IW; ++i)
#pragma unroll
        for (int k = 0; k < KW; ++k) sGM[(i0 + i) * D_OUT + k0 + k][lane] = gM[i][k];
    __syncthreads();
    // ggamma[c][j][u] = sum_nnz cgw * gM[i][k]: warp w owns j = w, w+NS, ...; accumulators stay in
    // registers (selected by an if-chain), so there is no shared-memory read-modify-write chain.
    constexpr int JW = (DG + NS - 1) / NS;
    TAcc accj[JW];
#pragma unroll
    for (int q = 0; q < JW; ++q) accj[q] = TAcc(0);
    if (live) {
        for (int n = 0; n < nnz; ++n) {
            const int i = tabp[3 * n], j = tabp[3 * n + 1], k = tabp[3 * n + 2];
            if (j % NS == warp) {
                const TAcc x = cgw[(int64_t)n * U + u] * sGM[i * D_OUT + k][lane];
                const int q = j / NS;
#pragma unroll
                for (int qq = 0; qq < JW; ++qq)
                    if (qq == q) accj[qq] += x;
            }
        }
#pragma unroll
        for (int q = 0; q < JW; ++q) {
            const int j = q * NS + warp;
            if (j < D) ggamma[(c * D + j) * U + u] = accj[q];
        }
    }
    (void)sG;
}

template <typename TAct, typename TAcc, int D_IN, int D_OUT, int DG>
int launch_fwd(int64_t N, int U, int D, int nnz, const int32_t* tabp, const void* cgw, const int32_t* row_ptr, const void* gamma,
               const void* Vin, int implicit_v0, const void* Y, const void* w0, int64_t w0_ld, void* Vout, cudaStream_t st) {
    {
        const int dt = sizeof(TAct) == 4 ? AB2_F32 : AB2_BF16;
        if (g_ab2_opt_tp_fast >= 1 && g_ab2_opt_tp_fast != 2 &&
            ab2_tp_smem(0, dt, N, U, D, D_IN, D_OUT, nnz, tabp, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, Vout, nullptr, nullptr,
                        nullptr, 0, nullptr, st) == 0)
            return 0;
    }
    const int64_t warps = N * ((U + 31) / 32);
    const unsigned grid = ab2_blocks(warps, 4);
    if (implicit_v0) {
        if constexpr (D_IN == 1 || D_IN == 4 || D_IN == 9 || D_IN == 16) {
            tp_fwd_fast_kernel<TAct, TAcc, D_IN, D_OUT, DG, true><<<grid, 128, 0, st>>>(
                N, U, D, nnz, tabp, (const TAcc*)cgw, row_ptr, (const TAcc*)gamma, nullptr, (const TAcc*)Y, (const TAct*)w0, w0_ld,
                (TAct*)Vout);
        } else {
            return -1;
        }
    } else {
        tp_fwd_fast_kernel<TAct, TAcc, D_IN, D_OUT, DG, false><<<grid, 128, 0, st>>>(
            N, U, D, nnz, tabp, (const TAcc*)cgw, row_ptr, (const TAcc*)gamma, (const TAct*)Vin, nullptr, nullptr, 0, (TAct*)Vout);
    }
    return 0;
}

template <typename TAct, typename TAcc, int D_IN, int D_OUT, int DG>
int launch_bwd(int64_t N, int U, int D, int nnz, const int32_t* tabp, const void* cgw, const int32_t* row_ptr, const void* gamma,
               const void* Vin, int implicit_v0, const void* Y, const void* w0, int64_t w0_ld, const void* gVout, void* gVin,
               void* gw0, int64_t gw0_ld, void* gY, void* ggamma, cudaStream_t st) {
    const int64_t warps = N * ((U + 31) / 32);
    const unsigned grid = ab2_blocks(warps, 4);
    if constexpr (D_IN * D_OUT >= 49) {
        // big M: split the backward.  A) gin / gw0 / gY with M in shared memory (tp_smem.cu, high occupancy);
        //                             B) gM -> ggamma with only gM in registers (this file).
        const int dt = sizeof(TAct) == 4 ? AB2_F32 : AB2_BF16;
        if (g_ab2_opt_tp_fast != 2 &&
            ab2_tp_smem(1, dt, N, U, D, D_IN, D_OUT, nnz, tabp, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, nullptr, gVout, gVin,
                        gw0, gw0_ld, gY, st) == 0) {
            constexpr int NS = (D_IN % 3 == 0) ? 3 : ((D_IN % 2 == 0) ? 2 : 1);
            const unsigned gridB = (unsigned)(N * ((U + 31) / 32));
            if (implicit_v0) {
                if constexpr (D_IN == 1 || D_IN == 4 || D_IN == 9 || D_IN == 16) {
                    tp_bwd_gm_split_kernel<TAct, TAcc, D_IN, D_OUT, DG, true, NS><<<gridB, NS * 32, 0, st>>>(
                        N, U, D, nnz, tabp, (const TAcc*)cgw, row_ptr, nullptr, (const TAcc*)Y, (const TAct*)w0, w0_ld, (const TAct*)gVout,
                        (TAcc*)ggamma);
                    return 0;
                }
            } else {
                tp_bwd_gm_split_kernel<TAct, TAcc, D_IN, D_OUT, DG, false, NS><<<gridB, NS * 32, 0, st>>>(
                    N, U, D, nnz, tabp, (const TAcc*)cgw, row_ptr, (const TAct*)Vin, nullptr, nullptr, 0, (const TAct*)gVout, (TAcc*)ggamma);
                return 0;
            }
        }
    }
    if (implicit_v0) {
        if constexpr (D_IN == 1 || D_IN == 4 || D_IN == 9 || D_IN == 16) {
            tp_bwd_fast_kernel<TAct, TAcc, D_IN, D_OUT, DG, true, false><<<grid, 128, 0, st>>>(
                N, U, D, nnz, tabp, (const TAcc*)cgw, row_ptr, (const TAcc*)gamma, nullptr, (const TAcc*)Y, (const TAct*)w0, w0_ld,
                (const TAct*)gVout, nullptr, (TAct*)gw0, gw0_ld, (TAcc*)gY, (TAcc*)ggamma);
        } else {
            return -1;
        }
    } else {
        tp_bwd_fast_kernel<TAct, TAcc, D_IN, D_OUT, DG, false, false><<<grid, 128, 0, st>>>(
            N, U, D, nnz, tabp, (const TAcc*)cgw, row_ptr, (const TAcc*)gamma, (const TAct*)Vin, nullptr, nullptr, 0,
            (const TAct*)gVout, (TAct*)gVin, nullptr, 0, nullptr, (TAcc*)ggamma);
    }
    return 0;
}

#define AB2_FAST_SHAPES(X) X(4, 4, 4) X(4, 1, 4) X(9, 9, 9) X(9, 1, 9) X(16, 1, 16) X(7, 4, 4) X(4, 7, 4) X(7, 7, 4) X(7, 1, 4)

}  // namespace

bool ab2_tp_fast_supported(int dtype, int D, int d_in, int d_out) {
    if (dtype == AB2_F64) return false;
#define X(a, b, dg) \
    if (d_in == a && d_out == b && D == dg) return true;
    AB2_FAST_SHAPES(X)
#undef X
    return false;
}

int ab2_tp_fwd_fast(int dtype, int64_t N, int U, int D, int d_in, int d_out, int nnz, const int32_t* tabp, const void* cgw,
                    const int32_t* row_ptr, const void* gamma, const void* Vin, int implicit_v0, const void* Y, const void* w0,
                    int64_t w0_ld, void* Vout, cudaStream_t st) {
#define X(a, b, dg)                                                                                                               \
    if (d_in == a && d_out == b && D == dg) {                                                                                                \
        if (dtype == AB2_F32)                                                                                                     \
            return launch_fwd<float, float, a, b, dg>(N, U, D, nnz, tabp, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, Vout, st); \
        if (dtype == AB2_BF16)                                                                                                    \
            return launch_fwd<bf16, float, a, b, dg>(N, U, D, nnz, tabp, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, Vout, st);  \
    }
    AB2_FAST_SHAPES(X)
#undef X
    return -1;
}

int ab2_tp_bwd_fast(int dtype, int64_t N, int U, int D, int d_in, int d_out, int nnz, const int32_t* tabp, const void* cgw,
                    const int32_t* row_ptr, const void* gamma, const void* Vin, int implicit_v0, const void* Y, const void* w0,
                    int64_t w0_ld, const void* gVout, void* gVin, void* gw0, int64_t gw0_ld, void* gY, void* ggamma, cudaStream_t st) {
#define X(a, b, dg)                                                                                                              \
    if (d_in == a && d_out == b && D == dg) {                                                                                               \
        if (dtype == AB2_F32)                                                                                                    \
            return launch_bwd<float, float, a, b, dg>(N, U, D, nnz, tabp, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, gVout, \
                                                  gVin, gw0, gw0_ld, gY, ggamma, st);                                            \
        if (dtype == AB2_BF16)                                                                                                   \
            return launch_bwd<bf16, float, a, b, dg>(N, U, D, nnz, tabp, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, gVout,  \
                                                 gVin, gw0, gw0_ld, gY, ggamma, st);                                             \
    }
    AB2_FAST_SHAPES(X)
#undef X
    return -1;
}
