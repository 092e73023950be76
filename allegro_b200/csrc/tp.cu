// Channel-wise Clebsch-Gordan tensor product of the edge features with the centre's
// environment (reference: Contracter._contract, allegro/nn/_strided/_contract.py:213-251,
// after the scatter/gather of :199-205), on the internal component-major layout V[E][d][U].
//
//   Vout[z][k][u] = sum_nnz cgw[nnz][u] * Vin[z][i][u] * gamma[c(z)][j][u]
//
// cgw[nnz][u] = w3j_value[nnz] * weights[u][path(nnz)] is the pre-contracted "ww3j" of
// _contract.py:218-219 in sparse form (83 non-zeros instead of 729 dense entries at l_max=2).
//
// This file holds the shape-generic kernels (any irreps; one thread per (edge, channel),
// channel fastest so all global accesses are coalesced).  The l_max=2 register-tiled
// fast path lives in tp_fast.cu.
#include "common.cuh"
#include "tp_fast.cuh"

#define AB2_TP_MAXD 64

template <typename TAct, typename TAcc>
__device__ __forceinline__ void tp_load_vin(TAcc* vin, int d_in, int implicit_v0, const TAct* __restrict__ Vin,
                                            const TAcc* __restrict__ Y, const TAct* __restrict__ w0, int64_t w0_ld, int lmax_d,
                                            int64_t z, int U, int u) {
    if (implicit_v0) {
        for (int i = 0; i < d_in; ++i) vin[i] = Y[z * lmax_d + i] * to_acc<TAcc>(w0[z * w0_ld + sh_l_of(i) * U + u]);
    } else {
        for (int i = 0; i < d_in; ++i) vin[i] = to_acc<TAcc>(Vin[(z * d_in + i) * U + u]);
    }
}

template <typename TAct, typename TAcc>
__global__ void __launch_bounds__(128) tp_fwd_generic_kernel(int64_t E, int U, int D, int d_in, int d_out, int nnz,
                                                             const int32_t* __restrict__ tab, const TAcc* __restrict__ cgw,
                                                             const int32_t* __restrict__ ctr, const TAcc* __restrict__ gamma,
                                                             const TAct* __restrict__ Vin, int implicit_v0,
                                                             const TAcc* __restrict__ Y, const TAct* __restrict__ w0, int64_t w0_ld,
                                                             TAct* __restrict__ Vout, const int* __restrict__ skip_flag) {
    if (skip_flag && *skip_flag) return;  // the baked-structure kernel launched in front of this one did the work
    const int64_t idx = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (idx >= E * U) return;
    const int64_t z = idx / U;
    const int u = (int)(idx - z * U);
    const int64_t c = ctr[z];
    TAcc vin[AB2_TP_MAXD], out[AB2_TP_MAXD];
    tp_load_vin<TAct, TAcc>(vin, d_in, implicit_v0, Vin, Y, w0, w0_ld, D, z, U, u);
    for (int k = 0; k < d_out; ++k) out[k] = TAcc(0);
    const TAcc* __restrict__ g = gamma + c * D * U + u;
    for (int n = 0; n < nnz; ++n) {
        const int i = tab[3 * n], j = tab[3 * n + 1], k = tab[3 * n + 2];
        out[k] += cgw[(int64_t)n * U + u] * vin[i] * g[(int64_t)j * U];
    }
    for (int k = 0; k < d_out; ++k) Vout[(z * d_out + k) * U + u] = from_acc<TAct>(out[k]);
}

template <typename TAct, typename TAcc>
__global__ void __launch_bounds__(128) tp_bwd_generic_kernel(int64_t E, int U, int D, int d_in, int d_out, int nnz,
                                                             const int32_t* __restrict__ tab, const TAcc* __restrict__ cgw,
                                                             const int32_t* __restrict__ ctr, const TAcc* __restrict__ gamma,
                                                             const TAct* __restrict__ Vin, int implicit_v0,
                                                             const TAcc* __restrict__ Y, const TAct* __restrict__ w0, int64_t w0_ld,
                                                             const TAct* __restrict__ gVout, TAct* __restrict__ gVin,
                                                             TAct* __restrict__ gw0, int64_t gw0_ld, TAcc* __restrict__ gY,
                                                             TAcc* __restrict__ ggamma, const int* __restrict__ skip_flag) {
    if (skip_flag && *skip_flag) return;
    const int64_t idx = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (idx >= E * U) return;
    const int64_t z = idx / U;
    const int u = (int)(idx - z * U);
    const int64_t c = ctr[z];
    TAcc vin[AB2_TP_MAXD], gin[AB2_TP_MAXD], gout[AB2_TP_MAXD], gg[AB2_MAX_LMAX * AB2_MAX_LMAX + 2 * AB2_MAX_LMAX + 1];
    tp_load_vin<TAct, TAcc>(vin, d_in, implicit_v0, Vin, Y, w0, w0_ld, D, z, U, u);
    for (int k = 0; k < d_out; ++k) gout[k] = to_acc<TAcc>(gVout[(z * d_out + k) * U + u]);
    for (int i = 0; i < d_in; ++i) gin[i] = TAcc(0);
    for (int j = 0; j < D; ++j) gg[j] = TAcc(0);
    const TAcc* __restrict__ g = gamma + c * D * U + u;
    for (int n = 0; n < nnz; ++n) {
        const int i = tab[3 * n], j = tab[3 * n + 1], k = tab[3 * n + 2];
        const TAcc t = cgw[(int64_t)n * U + u] * gout[k];
        gin[i] += t * g[(int64_t)j * U];
        gg[j] += t * vin[i];
    }
    for (int j = 0; j < D; ++j) atomicAdd(&ggamma[(c * D + j) * U + u], gg[j]);
    if (implicit_v0) {
        // Vin[i] = Y[i] * w0[l(i)]  ->  gw0[l] = sum_{i in l} Y[i] gin[i];  gY[i] += w0[l(i)] gin[i]
        int i = 0;
        for (int l = 0; i < d_in; ++l) {
            const TAcc wl = to_acc<TAcc>(w0[z * w0_ld + l * U + u]);
            TAcc s = TAcc(0);
            for (; i < (l + 1) * (l + 1) && i < d_in; ++i) {
                s += Y[z * D + i] * gin[i];
                atomicAdd(&gY[z * D + i], wl * gin[i]);
            }
            gw0[z * gw0_ld + l * U + u] = from_acc<TAct>(s);
        }
    } else {
        for (int i = 0; i < d_in; ++i) gVin[(z * d_in + i) * U + u] = from_acc<TAct>(gin[i]);
    }
}

extern "C" int ab2_tp_fwd(int dtype, int lmax, int64_t N, int64_t E, int U, int d_in, int d_out, int nnz, const int32_t* tab_ijk,
                          const void* cgw, const int32_t* row_ptr, const int32_t* ctr, const void* gamma, const void* Vin,
                          int implicit_v0, const void* Y, const void* w0, int64_t w0_ld, void* Vout, void* stream) {
    if (E == 0) return 0;
    const int D = (lmax + 1) * (lmax + 1);
    AB2_CHECK_ARG(lmax >= 0 && lmax <= AB2_MAX_LMAX, "lmax");
    AB2_CHECK_ARG(d_in <= AB2_TP_MAXD && d_out <= AB2_TP_MAXD && d_in > 0 && d_out > 0, "irreps dim exceeds AB2_TP_MAXD");
    AB2_CHECK_ARG(tab_ijk && cgw && ctr && gamma && Vout, "null pointer");
    AB2_CHECK_ARG(implicit_v0 ? (Y && w0 && d_in == D) : (Vin != nullptr), "input features");
    cudaStream_t st = (cudaStream_t)stream;
    if (g_ab2_opt_tp_fast && row_ptr &&
        ab2_tp_stream(0, dtype, N, E, U, D, d_in, d_out, nnz, tab_ijk, cgw, row_ptr, ctr, gamma, Vin, implicit_v0, Y, w0, w0_ld, Vout, nullptr,
                      nullptr, nullptr, 0, nullptr, nullptr, st) == 0) {
        AB2_CUDA_LAUNCH_CHECK();
        return 0;
    }
    if (g_ab2_opt_tp_fast && row_ptr && ab2_tp_fast_supported(dtype, D, d_in, d_out) && (!implicit_v0 || d_in == D)) {
        if (ab2_tp_fwd_fast(dtype, N, U, D, d_in, d_out, nnz, tab_ijk, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, Vout, st) == 0) {
            AB2_CUDA_LAUNCH_CHECK();
            return 0;
        }
    }
    // fp64, l_max = 3 layer shapes: baked-structure kernel first; it reports through a device flag whether the table was its
    // own, the shape-generic kernel behind it stands down if so (no host-side look at device data, graph-capturable)
    int* skip = nullptr;
    ab2_tp_baked64(0, dtype, E, U, D, d_in, d_out, nnz, tab_ijk, cgw, ctr, gamma, Vin, implicit_v0, Y, w0, w0_ld, Vout, nullptr, nullptr, nullptr,
                   0, nullptr, nullptr, &skip, st);
    AB2_DISPATCH_DTYPE(dtype, tp_fwd_generic_kernel<TAct, TAcc><<<ab2_blocks(E * U, 128), 128, 0, st>>>(
                                  E, U, D, d_in, d_out, nnz, tab_ijk, (const TAcc*)cgw, ctr, (const TAcc*)gamma, (const TAct*)Vin,
                                  implicit_v0, (const TAcc*)Y, (const TAct*)w0, w0_ld, (TAct*)Vout, skip));
    AB2_CUDA_LAUNCH_CHECK();
    return 0;
}

extern "C" int ab2_tp_bwd(int dtype, int lmax, int64_t N, int64_t E, int U, int d_in, int d_out, int nnz, const int32_t* tab_ijk,
                          const void* cgw, const int32_t* row_ptr, const int32_t* ctr, const void* gamma, const void* Vin,
                          int implicit_v0, const void* Y, const void* w0, int64_t w0_ld, const void* gVout, void* gVin, void* gw0,
                          int64_t gw0_ld, void* gY, void* ggamma, void* stream) {
    if (E == 0) return 0;
    const int D = (lmax + 1) * (lmax + 1);
    AB2_CHECK_ARG(lmax >= 0 && lmax <= AB2_MAX_LMAX, "lmax");
    AB2_CHECK_ARG(d_in <= AB2_TP_MAXD && d_out <= AB2_TP_MAXD && d_in > 0 && d_out > 0, "irreps dim exceeds AB2_TP_MAXD");
    AB2_CHECK_ARG(tab_ijk && cgw && ctr && gamma && gVout && ggamma, "null pointer");
    AB2_CHECK_ARG(implicit_v0 ? (Y && w0 && gw0 && gY && d_in == D) : (Vin && gVin), "input features / grads");
    cudaStream_t st = (cudaStream_t)stream;
    if (g_ab2_opt_tp_fast && row_ptr &&
        ab2_tp_stream(1, dtype, N, E, U, D, d_in, d_out, nnz, tab_ijk, cgw, row_ptr, ctr, gamma, Vin, implicit_v0, Y, w0, w0_ld, nullptr, gVout,
                      gVin, gw0, gw0_ld, gY, ggamma, st) == 0) {
        AB2_CUDA_LAUNCH_CHECK();
        return 0;
    }
    if (g_ab2_opt_tp_fast && row_ptr && ab2_tp_fast_supported(dtype, D, d_in, d_out) && (!implicit_v0 || d_in == D)) {
        if (ab2_tp_bwd_fast(dtype, N, U, D, d_in, d_out, nnz, tab_ijk, cgw, row_ptr, gamma, Vin, implicit_v0, Y, w0, w0_ld, gVout, gVin,
                            gw0, gw0_ld, gY, ggamma, st) == 0) {
            AB2_CUDA_LAUNCH_CHECK();
            return 0;
        }
    }
    // generic path: ggamma is accumulated with atomics: zero it first
    const size_t acc_size = (dtype == AB2_F64) ? 8 : 4;
    AB2_CUDA_CALL(cudaMemsetAsync(ggamma, 0, (size_t)N * D * U * acc_size, st));
    int* skip = nullptr;
    ab2_tp_baked64(1, dtype, E, U, D, d_in, d_out, nnz, tab_ijk, cgw, ctr, gamma, Vin, implicit_v0, Y, w0, w0_ld, nullptr, gVout, gVin, gw0,
                   gw0_ld, gY, ggamma, &skip, st);
    AB2_DISPATCH_DTYPE(dtype, tp_bwd_generic_kernel<TAct, TAcc><<<ab2_blocks(E * U, 128), 128, 0, st>>>(
                                  E, U, D, d_in, d_out, nnz, tab_ijk, (const TAcc*)cgw, ctr, (const TAcc*)gamma, (const TAct*)Vin,
                                  implicit_v0, (const TAcc*)Y, (const TAct*)w0, w0_ld, (const TAct*)gVout, (TAct*)gVin, (TAct*)gw0,
                                  gw0_ld, (TAcc*)gY, (TAcc*)ggamma, skip));
    AB2_CUDA_LAUNCH_CHECK();
    return 0;
}
