// Operator-level kernels behind the reference's own kernel plug-in point,
// Contracter.forward(x1, x2, idxs, scatter_dim_size)  (allegro/nn/_strided/_contract.py:185-211),
// on the reference's "strided" layout [z][u][i] with arbitrary (unsorted) int64 idxs -- the
// contract the Triton / cuEquivariance back-ends implement (_flashallegro.py:673-755,
// _cueq_contracter.py:84-131).  fp32 / fp64.
#include "common.cuh"

#define AB2_OP_MAXD 64

template <typename T>
__global__ void __launch_bounds__(256) op_scatter_env_kernel(int64_t total, int64_t row, T sf, const T* __restrict__ x2,
                                                             const int64_t* __restrict__ idxs, T* __restrict__ gamma) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= total) return;
    const int64_t z = e / row, r = e - z * row;
    atomicAdd(&gamma[idxs[z] * row + r], sf * x2[e]);
}

template <typename T>
__global__ void __launch_bounds__(256) op_gather_rows_kernel(int64_t total, int64_t row, T sf, const T* __restrict__ src,
                                                             const int64_t* __restrict__ idxs, T* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= total) return;
    const int64_t z = e / row, r = e - z * row;
    out[e] = sf * src[idxs[z] * row + r];
}

// mode 0: out[z][u][k] = sum cgw * a[z][u][i] * b[idx][u][j]        a = x1,   b = gamma
// mode 1: out[z][u][i] = sum cgw * a[z][u][k] * b[idx][u][j]        a = gout, b = gamma
// mode 2: out[idx][u][j] += sum cgw * a[z][u][i] * b[z][u][k]       a = x1,   b = gout   (atomic)
template <typename T, int MODE>
__global__ void __launch_bounds__(128) op_contract_kernel(int64_t E, int U, int d1, int d2, int dout, int nnz,
                                                          const int32_t* __restrict__ tab, const T* __restrict__ cgw,
                                                          const T* __restrict__ a, const T* __restrict__ b,
                                                          const int64_t* __restrict__ idxs, T* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (idx >= E * U) return;
    const int64_t z = idx / U;
    const int u = (int)(idx - z * U);
    const int64_t n_at = idxs[z];
    T acc[AB2_OP_MAXD];
    const int da = (MODE == 1) ? dout : d1;               // width of a
    const int db = (MODE == 2) ? dout : d2;               // width of b
    const int dres = (MODE == 0) ? dout : (MODE == 1 ? d1 : d2);
    const T* __restrict__ pa = a + (z * U + u) * da;
    const T* __restrict__ pb = (MODE == 2) ? b + (z * U + u) * db : b + (n_at * U + u) * db;
    for (int r = 0; r < dres; ++r) acc[r] = T(0);
    for (int n = 0; n < nnz; ++n) {
        const int i = tab[3 * n], j = tab[3 * n + 1], k = tab[3 * n + 2];
        const T c = cgw[(int64_t)n * U + u];
        if (MODE == 0) acc[k] += c * pa[i] * pb[j];
        if (MODE == 1) acc[i] += c * pa[k] * pb[j];
        if (MODE == 2) acc[j] += c * pa[i] * pb[k];
    }
    if (MODE == 2) {
        for (int r = 0; r < dres; ++r) atomicAdd(&out[(n_at * U + u) * dres + r], acc[r]);
    } else {
        for (int r = 0; r < dres; ++r) out[(z * U + u) * dres + r] = acc[r];
    }
}

extern "C" int ab2_op_scatter_env(int dtype, int64_t E, int64_t row, double sf, const void* x2, const int64_t* idxs, void* gamma,
                                  void* stream) {
    const int64_t total = E * row;
    if (total == 0) return 0;
    AB2_CHECK_ARG(x2 && idxs && gamma, "null pointer");
    AB2_CHECK_ARG(dtype == AB2_F64 || dtype == AB2_F32, "operator-level kernels are fp32/fp64");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == AB2_F64)
        op_scatter_env_kernel<double><<<ab2_blocks(total, 256), 256, 0, st>>>(total, row, sf, (const double*)x2, idxs, (double*)gamma);
    else
        op_scatter_env_kernel<float><<<ab2_blocks(total, 256), 256, 0, st>>>(total, row, (float)sf, (const float*)x2, idxs, (float*)gamma);
    AB2_CUDA_LAUNCH_CHECK();
    return 0;
}

extern "C" int ab2_op_gather_rows(int dtype, int64_t E, int64_t row, double sf, const void* src, const int64_t* idxs, void* out,
                                  void* stream) {
    const int64_t total = E * row;
    if (total == 0) return 0;
    AB2_CHECK_ARG(src && idxs && out, "null pointer");
    AB2_CHECK_ARG(dtype == AB2_F64 || dtype == AB2_F32, "operator-level kernels are fp32/fp64");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == AB2_F64)
        op_gather_rows_kernel<double><<<ab2_blocks(total, 256), 256, 0, st>>>(total, row, sf, (const double*)src, idxs, (double*)out);
    else
        op_gather_rows_kernel<float><<<ab2_blocks(total, 256), 256, 0, st>>>(total, row, (float)sf, (const float*)src, idxs, (float*)out);
    AB2_CUDA_LAUNCH_CHECK();
    return 0;
}

template <typename T>
static int op_contract_launch(int mode, int64_t E, int U, int d1, int d2, int dout, int nnz, const int32_t* tab, const void* cgw,
                              const void* a, const void* b, const int64_t* idxs, void* out, cudaStream_t st) {
    const unsigned g = ab2_blocks(E * U, 128);
    if (mode == 0)
        op_contract_kernel<T, 0><<<g, 128, 0, st>>>(E, U, d1, d2, dout, nnz, tab, (const T*)cgw, (const T*)a, (const T*)b, idxs, (T*)out);
    else if (mode == 1)
        op_contract_kernel<T, 1><<<g, 128, 0, st>>>(E, U, d1, d2, dout, nnz, tab, (const T*)cgw, (const T*)a, (const T*)b, idxs, (T*)out);
    else
        op_contract_kernel<T, 2><<<g, 128, 0, st>>>(E, U, d1, d2, dout, nnz, tab, (const T*)cgw, (const T*)a, (const T*)b, idxs, (T*)out);
    return 0;
}

extern "C" int ab2_op_contract(int dtype, int mode, int64_t E, int U, int d1, int d2, int dout, int nnz, const int32_t* tab_ijk,
                               const void* cgw, const void* a, const void* b, const int64_t* idxs, void* out, void* stream) {
    if (E == 0) return 0;
    AB2_CHECK_ARG(mode >= 0 && mode <= 2, "mode");
    AB2_CHECK_ARG(dtype == AB2_F64 || dtype == AB2_F32, "operator-level kernels are fp32/fp64");
    AB2_CHECK_ARG(d1 > 0 && d2 > 0 && dout > 0 && d1 <= AB2_OP_MAXD && d2 <= AB2_OP_MAXD && dout <= AB2_OP_MAXD, "irreps dim");
    AB2_CHECK_ARG(tab_ijk && cgw && a && b && idxs && out && U > 0 && nnz > 0, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == AB2_F64)
        op_contract_launch<double>(mode, E, U, d1, d2, dout, nnz, tab_ijk, cgw, a, b, idxs, out, st);
    else
        op_contract_launch<float>(mode, E, U, d1, d2, dout, nnz, tab_ijk, cgw, a, b, idxs, out, st);
    AB2_CUDA_LAUNCH_CHECK();
    return 0;
}

// Gradient w.r.t. the weighted coupling table (training; the reference's weights are Parameters, _contract.py:170-177, and its
// einsum path gets this product from autograd):  gcgw[n][u] += sum_z a[z][u][i_n] * b[idxs[z]][u][j_n] * g[z][u][k_n].
// grid = (ceil(nnz*U / 128), z-chunks); thread = one (n, u) over a chunk of edges, one atomicAdd per thread and chunk.
template <typename T>
__global__ void __launch_bounds__(128) op_contract_wgrad_kernel(int64_t E, int64_t zchunk, int U, int d1, int d2, int dout, int nnz,
                                                                const int32_t* __restrict__ tab, const T* __restrict__ a,
                                                                const T* __restrict__ b, const T* __restrict__ g,
                                                                const int64_t* __restrict__ idxs, T* __restrict__ out) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    if (t >= nnz * U) return;
    const int n = t / U, u = t - n * U;
    const int i = tab[3 * n], j = tab[3 * n + 1], k = tab[3 * n + 2];
    const int64_t z0 = (int64_t)blockIdx.y * zchunk;
    const int64_t z1 = z0 + zchunk < E ? z0 + zchunk : E;
    T acc = T(0);
    for (int64_t z = z0; z < z1; ++z)
        acc += a[(z * U + u) * d1 + i] * b[(idxs[z] * U + u) * d2 + j] * g[(z * U + u) * dout + k];
    atomicAdd(&out[(int64_t)n * U + u], acc);
}

extern "C" int ab2_op_contract_wgrad(int dtype, int64_t E, int U, int d1, int d2, int dout, int nnz, const int32_t* tab_ijk, const void* x1,
                                     const void* gamma, const void* gout, const int64_t* idxs, void* gcgw, void* stream) {
    if (E == 0) return 0;
    AB2_CHECK_ARG(dtype == AB2_F64 || dtype == AB2_F32, "operator-level kernels are fp32/fp64");
    AB2_CHECK_ARG(d1 > 0 && d2 > 0 && dout > 0 && U > 0 && nnz > 0, "shape");
    AB2_CHECK_ARG(tab_ijk && x1 && gamma && gout && idxs && gcgw, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t zchunk = 512;
    const int64_t ny = (E + zchunk - 1) / zchunk;
    AB2_CHECK_ARG(ny <= 65535, "too many edges for one call");
    const dim3 grid((unsigned)((nnz * U + 127) / 128), (unsigned)ny);
    if (dtype == AB2_F64)
        op_contract_wgrad_kernel<double><<<grid, 128, 0, st>>>(E, zchunk, U, d1, d2, dout, nnz, tab_ijk, (const double*)x1, (const double*)gamma,
                                                              (const double*)gout, idxs, (double*)gcgw);
    else
        op_contract_wgrad_kernel<float><<<grid, 128, 0, st>>>(E, zchunk, U, d1, d2, dout, nnz, tab_ijk, (const float*)x1, (const float*)gamma,
                                                             (const float*)gout, idxs, (float*)gcgw);
    AB2_CUDA_LAUNCH_CHECK();
    return 0;
}
