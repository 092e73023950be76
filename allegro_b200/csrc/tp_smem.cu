// Occupancy-oriented tensor-product kernels: the per-centre coupling matrix M_c[u][i][k]
// (built once per centre from the CG table and the centre's environment, see tp_fast.cu) lives
// in SHARED memory for a small group of centres, and the work item is one WARP per
// (edge, 32-channel chunk), lane = channel.  Per item: D coalesced 128-byte loads, D_IN*D_OUT
// FMAs fed by conflict-free LDS (lanes read consecutive words), D coalesced stores -- ~40
// registers per thread, so 40 warps per SM keep enough loads in flight to stream HBM, instead
// of the 8 warps the register-resident M allows.
//
//   MODE 0  forward     Vout[z][k][u] = sum_i Vin[z][i][u] M[i][k]
//   MODE 1  backward A  gVin[z][i][u] = sum_k M[i][k] gVout[z][k][u]
//           (layer 0: Vin = Y (x) w0 is implicit, so instead gw0[z][l][u] = sum_{i in l} Y[z][i] gin[i]
//            and gY[z][i] += sum_u w0[z][l(i)][u] gin[i] by a multi-value warp butterfly)
// The other half of the backward (gM -> ggamma, register-resident, warp per centre) is
// tp_bwd_fast_kernel<..., GM_ONLY> in tp_fast.cu.
#include "common.cuh"
#include "tp_fast.cuh"

extern int g_ab2_opt_tp_variant;

namespace {

constexpr int CPB = 3;  // centres per CTA

template <typename T>
struct alignas(4 * sizeof(T)) Vec4 {
    T v[4];
};

template <typename TAct, typename TAcc, int D_IN, int D_OUT, bool IMPLICIT, int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB) tp_smem_kernel(int64_t N, int U, int D, int nnz, const int32_t* __restrict__ tab,
                                                      const TAcc* __restrict__ cgw, const int32_t* __restrict__ row_ptr,
                                                      const TAcc* __restrict__ gamma, const TAct* __restrict__ Vin,
                                                      const TAcc* __restrict__ Y, const TAct* __restrict__ w0, int64_t w0_ld,
                                                      TAct* __restrict__ Vout, const TAct* __restrict__ gVout,
                                                      TAct* __restrict__ gVin, TAct* __restrict__ gw0, int64_t gw0_ld,
                                                      TAcc* __restrict__ gY) {
    constexpr int KQ = (D_OUT + 3) / 4;  // output components in groups of 4 -> one 16-byte LDS per (i, group)
    __shared__ Vec4<TAcc> sM[CPB][D_IN * KQ][32];
    __shared__ int s_rp[CPB + 1];
    const int tid = threadIdx.x;
    const int64_t c0 = (int64_t)blockIdx.x * CPB;
    const int u0 = blockIdx.y * 32;
    const int nchunk = gridDim.y;
    if (tid <= CPB) s_rp[tid] = row_ptr[min(c0 + tid, N)];
    // ---- build M for (CPB centres) x (32 channels): one thread per column ----
    if (tid < CPB * 32) {
        const int cc = tid >> 5, lu = tid & 31;
        const int64_t c = c0 + cc;
        const int u = u0 + lu;
#pragma unroll
        for (int e = 0; e < D_IN * KQ; ++e) {
#pragma unroll
            for (int t = 0; t < 4; ++t) sM[cc][e][lu].v[t] = TAcc(0);
        }
        if (c < N && u < U) {
            const TAcc* __restrict__ g = gamma + c * D * U + u;
            for (int n = 0; n < nnz; ++n) {
                const int i = tab[3 * n], j = tab[3 * n + 1], k = tab[3 * n + 2];
                sM[cc][i * KQ + (k >> 2)][lu].v[k & 3] += cgw[(int64_t)n * U + u] * g[(int64_t)j * U];
            }
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const int u = u0 + lane;
    const bool live = u < U;
    // EPT consecutive edges of ONE centre per warp-item: every 16-byte LDS of M feeds EPT edges, which
    // divides the shared-memory traffic (the limiter with one edge per item) by EPT.
    constexpr int EPT = 1;
    for (int cc = 0; cc < CPB; ++cc) {
        const int e_beg = s_rp[cc], e_end = s_rp[cc + 1];
        for (int64_t z0 = e_beg + warp * EPT; z0 < e_end; z0 += 8 * EPT) {
            bool ok[EPT];
            int64_t zz[EPT];
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                ok[e] = (z0 + e) < e_end;
                zz[e] = ok[e] ? z0 + e : z0;  // clamp: loads stay in range, stores are predicated
            }
            if constexpr (MODE == 0) {
                TAcc v[EPT][D_IN];
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    const int64_t z = zz[e];
                    if constexpr (IMPLICIT) {
                        TAcc w0l[5];
#pragma unroll
                        for (int l = 0; l * l < D_IN; ++l) w0l[l] = live ? to_acc<TAcc>(w0[z * w0_ld + l * U + u]) : TAcc(0);
#pragma unroll
                        for (int i = 0; i < D_IN; ++i) v[e][i] = Y[z * D_IN + i] * w0l[sh_l_of(i)];
                    } else {
#pragma unroll
                        for (int i = 0; i < D_IN; ++i) v[e][i] = live ? to_acc<TAcc>(Vin[(z * D_IN + i) * U + u]) : TAcc(0);
                    }
                }
                TAcc out[EPT][D_OUT];
#pragma unroll
                for (int e = 0; e < EPT; ++e)
#pragma unroll
                    for (int k = 0; k < D_OUT; ++k) out[e][k] = TAcc(0);
#pragma unroll
                for (int i = 0; i < D_IN; ++i)
#pragma unroll
                    for (int kq = 0; kq < KQ; ++kq) {
                        const Vec4<TAcc> m4 = sM[cc][i * KQ + kq][lane];
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (kq * 4 + t < D_OUT) {
#pragma unroll
                                for (int e = 0; e < EPT; ++e) out[e][kq * 4 + t] += v[e][i] * m4.v[t];
                            }
                    }
#pragma unroll
                for (int e = 0; e < EPT; ++e)
                    if (live && ok[e]) {
#pragma unroll
                        for (int k = 0; k < D_OUT; ++k) Vout[(zz[e] * D_OUT + k) * U + u] = from_acc<TAct>(out[e][k]);
                    }
            } else {
                TAcc go[EPT][D_OUT];
#pragma unroll
                for (int e = 0; e < EPT; ++e)
#pragma unroll
                    for (int k = 0; k < D_OUT; ++k) go[e][k] = live ? to_acc<TAcc>(gVout[(zz[e] * D_OUT + k) * U + u]) : TAcc(0);
                TAcc gin[EPT][D_IN];
#pragma unroll
                for (int i = 0; i < D_IN; ++i) {
                    TAcc s[EPT];
#pragma unroll
                    for (int e = 0; e < EPT; ++e) s[e] = TAcc(0);
#pragma unroll
                    for (int kq = 0; kq < KQ; ++kq) {
                        const Vec4<TAcc> m4 = sM[cc][i * KQ + kq][lane];
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (kq * 4 + t < D_OUT) {
#pragma unroll
                                for (int e = 0; e < EPT; ++e) s[e] += m4.v[t] * go[e][kq * 4 + t];
                            }
                    }
#pragma unroll
                    for (int e = 0; e < EPT; ++e) gin[e][i] = s[e];
                }
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    const int64_t z = zz[e];
                    if constexpr (IMPLICIT) {
                        TAcc part[D_IN];
#pragma unroll
                        for (int l = 0; l * l < D_IN; ++l) {
                            const TAcc w0l = live ? to_acc<TAcc>(w0[z * w0_ld + l * U + u]) : TAcc(0);
                            TAcc s = TAcc(0);
#pragma unroll
                            for (int i = l * l; i < (l + 1) * (l + 1); ++i) {
                                s += Y[z * D_IN + i] * gin[e][i];
                                part[i] = w0l * gin[e][i];
                            }
                            if (live && ok[e]) gw0[z * gw0_ld + l * U + u] = from_acc<TAct>(s);
                        }
                        const TAcc tot = warp_multi_sum<TAcc, D_IN>(part, lane);
                        const int j = lane >> 1;
                        if (ok[e] && !(lane & 1) && j < D_IN) {
                            if (nchunk == 1) gY[z * D_IN + j] += tot;  // single writer per (z, j)
                            else atomicAdd(&gY[z * D_IN + j], tot);
                        }
                    } else {
                        if (live && ok[e]) {
#pragma unroll
                            for (int i = 0; i < D_IN; ++i) gVin[(z * D_IN + i) * U + u] = from_acc<TAct>(gin[e][i]);
                        }
                    }
                }
            }
        }
    }
}

template <typename TAct, typename TAcc, int D_IN, int D_OUT, int MODE, int MINB>
int launch_v(int64_t N, int U, int D, int nnz, const int32_t* tab, const void* cgw, const int32_t* row_ptr, const void* gamma,
           const void* Vin, int implicit_v0, const void* Y, const void* w0, int64_t w0_ld, void* Vout, const void* gVout, void* gVin,
           void* gw0, int64_t gw0_ld, void* gY, cudaStream_t st) {
    dim3 grid(ab2_blocks(N, CPB), (unsigned)((U + 31) / 32));
    if (implicit_v0) {
        if constexpr (D_IN == 1 || D_IN == 4 || D_IN == 9 || D_IN == 16) {
            tp_smem_kernel<TAct, TAcc, D_IN, D_OUT, true, MODE, MINB><<<grid, 256, 0, st>>>(
                N, U, D, nnz, tab, (const TAcc*)cgw, row_ptr, (const TAcc*)gamma, nullptr, (const TAcc*)Y, (const TAct*)w0, w0_ld,
                (TAct*)Vout, (const TAct*)gVout, nullptr, (TAct*)gw0, gw0_ld, (TAcc*)gY);
            return 0;
        }
        return -1;
    }
    tp_smem_kernel<TAct, TAcc, D_IN, D_OUT, false, MODE, MINB><<<grid, 256, 0, st>>>(
        N, U, D, nnz, tab, (const TAcc*)cgw, row_ptr, (const TAcc*)gamma, (const TAct*)Vin, nullptr, nullptr, 0, (TAct*)Vout,
        (const TAct*)gVout, (TAct*)gVin, nullptr, 0, nullptr);
    return 0;
}

template <typename TAct, typename TAcc, int D_IN, int D_OUT, int MODE, typename... A>
int launch(A... a) {
    if (g_ab2_opt_tp_variant == 1) return launch_v<TAct, TAcc, D_IN, D_OUT, MODE, 3>(a...);
    return launch_v<TAct, TAcc, D_IN, D_OUT, MODE, 2>(a...);
}

// shapes whose M (CPB * D_IN * ceil(D_OUT/4)*4 * 32 floats) fits the 48 KB static shared-memory limit
#define AB2_SMEM_SHAPES(X) X(4, 4, 4) X(4, 1, 4) X(9, 9, 9) X(9, 1, 9) X(16, 1, 16) X(7, 4, 4) X(4, 7, 4) X(7, 7, 4) X(7, 1, 4)

}  // namespace

int g_ab2_opt_tp_variant = 0;

int ab2_tp_smem(int mode, int dtype, int64_t N, int U, int D, int d_in, int d_out, int nnz, const int32_t* tab, const void* cgw,
                const int32_t* row_ptr, const void* gamma, const void* Vin, int implicit_v0, const void* Y, const void* w0,
                int64_t w0_ld, void* Vout, const void* gVout, void* gVin, void* gw0, int64_t gw0_ld, void* gY, cudaStream_t st) {
#define X(a, b, dg)                                                                                                              \
    if (d_in == a && d_out == b && D == dg) {                                                                                    \
        if (dtype == AB2_F32 && mode == 0)                                                                                       \
            return launch<float, float, a, b, 0>(N, U, D, nnz, tab, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, Vout, gVout, gVin, gw0, gw0_ld, gY, st); \
        if (dtype == AB2_F32 && mode == 1)                                                                                       \
            return launch<float, float, a, b, 1>(N, U, D, nnz, tab, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, Vout, gVout, gVin, gw0, gw0_ld, gY, st); \
        if (dtype == AB2_BF16 && mode == 0)                                                                                      \
            return launch<bf16, float, a, b, 0>(N, U, D, nnz, tab, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, Vout, gVout, gVin, gw0, gw0_ld, gY, st);  \
        if (dtype == AB2_BF16 && mode == 1)                                                                                      \
            return launch<bf16, float, a, b, 1>(N, U, D, nnz, tab, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, Vout, gVout, gVin, gw0, gw0_ld, gY, st);  \
    }
    AB2_SMEM_SHAPES(X)
#undef X
    return -1;
}
