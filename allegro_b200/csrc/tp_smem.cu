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

template <typename TAct, typename TAcc, int D_IN, int D_OUT, bool IMPLICIT, int MODE, int MINB, int UT>
__global__ void __launch_bounds__(256, MINB) tp_smem_kernel(int64_t N, int U_rt, int D, int nnz, const int32_t* __restrict__ tab,
                                                      const TAcc* __restrict__ cgw, const int32_t* __restrict__ row_ptr,
                                                      const TAcc* __restrict__ gamma, const TAct* __restrict__ Vin,
                                                      const TAcc* __restrict__ Y, const TAct* __restrict__ w0, int64_t w0_ld,
                                                      TAct* __restrict__ Vout, const TAct* __restrict__ gVout,
                                                      TAct* __restrict__ gVin, TAct* __restrict__ gw0, int64_t gw0_ld,
                                                      TAcc* __restrict__ gY) {
    // UT == 32: the common channel count is a compile-time constant (immediate strides, no lane guards)
    const int U = UT ? UT : U_rt;
    constexpr int KQ = (D_OUT + 3) / 4;  // output components in groups of 4 -> one 16-byte LDS per (i, group)
    __shared__ Vec4<TAcc> sM[CPB][D_IN * KQ][32];
    __shared__ int s_rp[CPB + 1];
    const int tid = threadIdx.x;
    const int64_t c0 = (int64_t)blockIdx.x * CPB;
    const int u0 = blockIdx.y * 32;
    const int nchunk = gridDim.y;
    if (tid <= CPB) s_rp[tid] = row_ptr[min(c0 + tid, N)];
    // ---- per-target segment pointers of the (i,k)-sorted table ----
    constexpr int T = D_IN * D_OUT;
    __shared__ int s_ptr[T + 1];
    for (int t = tid; t <= T; t += 256) s_ptr[t] = -1;
    __syncthreads();
    for (int n = tid; n < nnz; n += 256) {
        const int t = tab[3 * n] * D_OUT + tab[3 * n + 2];
        if (n == 0 || (tab[3 * n - 3] * D_OUT + tab[3 * n - 1]) != t) s_ptr[t] = n;
    }
    __syncthreads();
    if (tid == 0) {
        s_ptr[T] = nnz;
        for (int t = T - 1; t >= 0; --t)
            if (s_ptr[t] < 0) s_ptr[t] = s_ptr[t + 1];
    }
    __syncthreads();
    // ---- build M for (CPB centres) x (32 channels): every entry M[i][k] is a register gather over its
    //      table segment (no shared-memory read-modify-write chain), all 256 threads busy ----
    for (int idx = tid; idx < CPB * T * 32; idx += 256) {
        const int lu = idx & 31;
        const int t = (idx >> 5) % T;
        const int cc = (idx >> 5) / T;
        const int64_t c = c0 + cc;
        const int u = u0 + lu;
        TAcc acc = TAcc(0);
        if (c < N && u < U) {
            const TAcc* __restrict__ g = gamma + c * D * U + u;
            for (int n = s_ptr[t]; n < s_ptr[t + 1]; ++n) acc += cgw[(int64_t)n * U + u] * g[(int64_t)tab[3 * n + 1] * U];
        }
        const int i = t / D_OUT, k = t - i * D_OUT;
        sM[cc][i * KQ + (k >> 2)][lu].v[k & 3] = acc;
    }
    if constexpr (D_OUT % 4 != 0) {  // zero the padding lanes of the last k-quad (read by the 16-byte LDS)
        for (int idx = tid; idx < CPB * D_IN * 32; idx += 256) {
            const int lu = idx & 31, i = (idx >> 5) % D_IN, cc = (idx >> 5) / D_IN;
#pragma unroll
            for (int k = D_OUT; k < KQ * 4; ++k) sM[cc][i * KQ + (k >> 2)][lu].v[k & 3] = TAcc(0);
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const int u = u0 + lane;
    const bool live = UT ? true : (u < U);
    // Work item = one warp per (edge, channel chunk).  The global inputs of item n+1 are loaded into
    // registers before item n is computed (manual double buffering): the loads are few (<= 21 per
    // lane) but their latency was the top stall of the single-buffered loop.
    struct In {
        TAcc a[(MODE == 0 && !IMPLICIT) ? D_IN : ((MODE == 1) ? D_OUT : 1)];
        TAcc w0l[IMPLICIT ? 5 : 1];
        TAcc Yz[IMPLICIT ? D_IN : 1];
    };
    auto load_in = [&](int64_t z, In& in) {
        if constexpr (MODE == 0 && !IMPLICIT) {
#pragma unroll
            for (int i = 0; i < D_IN; ++i) in.a[i] = live ? to_acc<TAcc>(Vin[(z * D_IN + i) * U + u]) : TAcc(0);
        }
        if constexpr (MODE == 1) {
#pragma unroll
            for (int k = 0; k < D_OUT; ++k) in.a[k] = live ? to_acc<TAcc>(gVout[(z * D_OUT + k) * U + u]) : TAcc(0);
        }
        if constexpr (IMPLICIT) {
#pragma unroll
            for (int l = 0; l * l < D_IN; ++l) in.w0l[l] = live ? to_acc<TAcc>(w0[z * w0_ld + l * U + u]) : TAcc(0);
#pragma unroll
            for (int i = 0; i < D_IN; ++i) in.Yz[i] = Y[z * D_IN + i];
        }
    };
    const int e_beg = s_rp[0], e_end = s_rp[CPB];
    auto compute = [&](const In& in, int64_t z) {
        int cc = 0;
#pragma unroll
        for (int t = 1; t < CPB; ++t) cc += (z >= s_rp[t]) ? 1 : 0;
        if constexpr (MODE == 0) {
            TAcc v[D_IN];
#pragma unroll
            for (int i = 0; i < D_IN; ++i) {
                if constexpr (IMPLICIT) v[i] = in.Yz[i] * in.w0l[sh_l_of(i)];
                else v[i] = in.a[i];
            }
            TAcc out[D_OUT];
#pragma unroll
            for (int k = 0; k < D_OUT; ++k) out[k] = TAcc(0);
#pragma unroll
            for (int i = 0; i < D_IN; ++i)
#pragma unroll
                for (int kq = 0; kq < KQ; ++kq) {
                    const Vec4<TAcc> m4 = sM[cc][i * KQ + kq][lane];
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (kq * 4 + t < D_OUT) out[kq * 4 + t] += v[i] * m4.v[t];
                }
            if (live) {
#pragma unroll
                for (int k = 0; k < D_OUT; ++k) Vout[(z * D_OUT + k) * U + u] = from_acc<TAct>(out[k]);
            }
        } else {
            TAcc gin[D_IN];
#pragma unroll
            for (int i = 0; i < D_IN; ++i) {
                TAcc s = TAcc(0);
#pragma unroll
                for (int kq = 0; kq < KQ; ++kq) {
                    const Vec4<TAcc> m4 = sM[cc][i * KQ + kq][lane];
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (kq * 4 + t < D_OUT) s += m4.v[t] * in.a[kq * 4 + t];
                }
                gin[i] = s;
            }
            if constexpr (IMPLICIT) {
                TAcc part[D_IN];
#pragma unroll
                for (int l = 0; l * l < D_IN; ++l) {
                    TAcc s = TAcc(0);
#pragma unroll
                    for (int i = l * l; i < (l + 1) * (l + 1); ++i) {
                        s += in.Yz[i] * gin[i];
                        part[i] = in.w0l[l] * gin[i];
                    }
                    if (live) gw0[z * gw0_ld + l * U + u] = from_acc<TAct>(s);
                }
                const TAcc tot = warp_multi_sum<TAcc, D_IN>(part, lane);
                const int j = lane >> 1;
                if (!(lane & 1) && j < D_IN) {
                    if (nchunk == 1) gY[z * D_IN + j] += tot;  // single writer per (z, j)
                    else atomicAdd(&gY[z * D_IN + j], tot);
                }
            } else {
                if (live) {
#pragma unroll
                    for (int i = 0; i < D_IN; ++i) gVin[(z * D_IN + i) * U + u] = from_acc<TAct>(gin[i]);
                }
            }
        }
    };
    // two register buffers used alternately (loop unrolled by two, no buffer copy): the loads of the
    // next item are in flight while the current one is computed
    In bufA, bufB;
    int64_t z = e_beg + warp;
    if (z < e_end) load_in(z, bufA);
    while (z < e_end) {
        if (z + 8 < e_end) load_in(z + 8, bufB);
        compute(bufA, z);
        z += 8;
        if (z >= e_end) break;
        if (z + 8 < e_end) load_in(z + 8, bufA);
        compute(bufB, z);
        z += 8;
    }
}

template <typename TAct, typename TAcc, int D_IN, int D_OUT, int MODE, int MINB, int UT>
int launch_u(int64_t N, int U, int D, int nnz, const int32_t* tab, const void* cgw, const int32_t* row_ptr, const void* gamma,
           const void* Vin, int implicit_v0, const void* Y, const void* w0, int64_t w0_ld, void* Vout, const void* gVout, void* gVin,
           void* gw0, int64_t gw0_ld, void* gY, cudaStream_t st) {
    dim3 grid(ab2_blocks(N, CPB), (unsigned)((U + 31) / 32));
    if (implicit_v0) {
        if constexpr (D_IN == 1 || D_IN == 4 || D_IN == 9 || D_IN == 16) {
            tp_smem_kernel<TAct, TAcc, D_IN, D_OUT, true, MODE, MINB, UT><<<grid, 256, 0, st>>>(
                N, U, D, nnz, tab, (const TAcc*)cgw, row_ptr, (const TAcc*)gamma, nullptr, (const TAcc*)Y, (const TAct*)w0, w0_ld,
                (TAct*)Vout, (const TAct*)gVout, nullptr, (TAct*)gw0, gw0_ld, (TAcc*)gY);
            return 0;
        }
        return -1;
    }
    tp_smem_kernel<TAct, TAcc, D_IN, D_OUT, false, MODE, MINB, UT><<<grid, 256, 0, st>>>(
        N, U, D, nnz, tab, (const TAcc*)cgw, row_ptr, (const TAcc*)gamma, (const TAct*)Vin, nullptr, nullptr, 0, (TAct*)Vout,
        (const TAct*)gVout, (TAct*)gVin, nullptr, 0, nullptr);
    return 0;
}

template <typename TAct, typename TAcc, int D_IN, int D_OUT, int MODE, int MINB>
int launch_v(int64_t N, int U, int D, int nnz, const int32_t* tab, const void* cgw, const int32_t* row_ptr, const void* gamma,
             const void* Vin, int implicit_v0, const void* Y, const void* w0, int64_t w0_ld, void* Vout, const void* gVout, void* gVin,
             void* gw0, int64_t gw0_ld, void* gY, cudaStream_t st) {
    if (U == 32)
        return launch_u<TAct, TAcc, D_IN, D_OUT, MODE, MINB, 32>(N, U, D, nnz, tab, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, Vout,
                                                                gVout, gVin, gw0, gw0_ld, gY, st);
    return launch_u<TAct, TAcc, D_IN, D_OUT, MODE, MINB, 0>(N, U, D, nnz, tab, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, Vout,
                                                           gVout, gVin, gw0, gw0_ld, gY, st);
}

template <typename TAct, typename TAcc, int D_IN, int D_OUT, int MODE, typename... A>
int launch(A... a) {
    if (g_ab2_opt_tp_variant == 1) return launch_v<TAct, TAcc, D_IN, D_OUT, MODE, 3>(a...);
    return launch_v<TAct, TAcc, D_IN, D_OUT, MODE, 2>(a...);
}

// shapes whose M (CPB * D_IN * ceil(D_OUT/4)*4 * 32 floats) fits the 48 KB static shared-memory limit
#define AB2_SMEM_SHAPES(X) X(4, 4, 4) X(4, 1, 4) X(9, 9, 9) X(9, 1, 9) X(16, 1, 16) X(7, 4, 4) X(4, 7, 4) X(7, 7, 4) X(7, 1, 4)

}  // namespace

int g_ab2_opt_tp_variant = 1;  // 1: 3 CTAs/SM (80 registers), 0: 2 CTAs/SM (<=128 registers)

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
