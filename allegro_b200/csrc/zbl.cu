// ZBL screened nuclear repulsion as a pair term on top of the Allegro energy.
//
// Reference call site: allegro/model/allegro_models.py:270-288 (`pair_potential`, an AddRadialCutoffToData with
// PolynomialCutoff(6) in front of it, the result added to the per-atom energies after the scale/shift).  The module itself
// is nequip's nequip.nn.pair_potential.ZBL, which is not vendored in /root/reference; what is restated here is its
// published algorithm -- LAMMPS pair_style zbl with the constants of pair_zbl_const.h:
//   E_z = (qqr2e / 2) Z_i Z_j / r * phi(r / a) * u(r / r_max),   a = 0.46850 / (Z_i^0.23 + Z_j^0.23),
//   phi(x) = 0.18175 e^{-3.19980 x} + 0.50986 e^{-0.94229 x} + 0.28022 e^{-0.40290 x} + 0.02817 e^{-0.20162 x},
// the factor 1/2 because every pair appears as two directed edges, u the polynomial cutoff.  One thread per edge:
// the edge energy and, in the same pass, its derivative added into the per-edge gradient dE/dvec that the force scatter
// consumes (the upstream gradient of every atomic energy w.r.t. this term is 1: it is added after the per-type scaling).
#include "common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ T zbl_exp(T x);
template <>
__device__ __forceinline__ float zbl_exp<float>(float x) { return expf(x); }
template <>
__device__ __forceinline__ double zbl_exp<double>(double x) { return exp(x); }
__device__ __forceinline__ float zbl_pow(float x, float p) { return powf(x, p); }
__device__ __forceinline__ double zbl_pow(double x, double p) { return pow(x, p); }
__device__ __forceinline__ float zbl_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double zbl_sqrt(double x) { return sqrt(x); }

template <typename T>
__global__ void __launch_bounds__(128) zbl_kernel(int64_t E, T p, T qq, const T* __restrict__ vec, const int32_t* __restrict__ ctr,
                                                  const int32_t* __restrict__ nbr, const int32_t* __restrict__ types, const T* __restrict__ Z,
                                                  const T* __restrict__ rmax_table, int num_types, T* __restrict__ Ez, T* __restrict__ gvec) {
    const int64_t z = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (z >= E) return;
    const T vx = vec[3 * z], vy = vec[3 * z + 1], vz = vec[3 * z + 2];
    const T r = zbl_sqrt(vx * vx + vy * vy + vz * vz);
    const int tc = types[ctr[z]], tn = types[nbr[z]];
    const T zi = Z[tc], zj = Z[tn];
    const T rmax = rmax_table[tc * num_types + tn];
    const T x = r / rmax;
    T e = T(0), dedr = T(0);
    if (x < T(1)) {
        // polynomial cutoff u(x) = 1 - (p+1)(p+2)/2 x^p + p(p+2) x^(p+1) - p(p+1)/2 x^(p+2)
        const T xp = zbl_pow(x, p);
        const T c0 = (p + T(1)) * (p + T(2)) / T(2), c1 = p * (p + T(2)), c2 = p * (p + T(1)) / T(2);
        const T u = T(1) - c0 * xp + c1 * xp * x - c2 * xp * x * x;
        const T du = (-c0 * p * xp / x + c1 * (p + T(1)) * xp - c2 * (p + T(2)) * xp * x) / rmax;  // du/dr
        const T s = (zbl_pow(zi, T(0.23)) + zbl_pow(zj, T(0.23))) / T(0.46850);                    // x_zbl = s r
        const T xs = s * r;
        const T e1 = T(0.02817) * zbl_exp(T(-0.20162) * xs), e2 = T(0.28022) * zbl_exp(T(-0.40290) * xs);
        const T e3 = T(0.50986) * zbl_exp(T(-0.94229) * xs), e4 = T(0.18175) * zbl_exp(T(-3.19980) * xs);
        const T phi = e1 + e2 + e3 + e4;
        const T dphi = s * (T(-0.20162) * e1 + T(-0.40290) * e2 + T(-0.94229) * e3 + T(-3.19980) * e4);  // dphi/dr
        const T pre = qq * zi * zj / r;
        e = pre * phi * u;
        dedr = pre * ((dphi - phi / r) * u + phi * du);
    }
    if (Ez) Ez[z] = e;
    if (gvec) {
        const T f = dedr / r;
        gvec[3 * z] += f * vx;
        gvec[3 * z + 1] += f * vy;
        gvec[3 * z + 2] += f * vz;
    }
}

}  // namespace

extern "C" int ab2_zbl(int acc_dtype, int64_t E, int num_types, double p_cut, double qq, const void* vec, const int32_t* ctr, const int32_t* nbr,
                       const int32_t* types, const void* Z, const void* rmax_table, void* Ez, void* gvec, void* stream) {
    if (E == 0) return 0;
    AB2_CHECK_ARG(acc_dtype == AB2_F64 || acc_dtype == AB2_F32, "ab2_zbl works in the accumulate type (fp32 / fp64)");
    AB2_CHECK_ARG(vec && ctr && nbr && types && Z && rmax_table && (Ez || gvec) && num_types > 0, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (acc_dtype == AB2_F64)
        zbl_kernel<double><<<ab2_blocks(E, 128), 128, 0, st>>>(E, p_cut, qq, (const double*)vec, ctr, nbr, types, (const double*)Z,
                                                              (const double*)rmax_table, num_types, (double*)Ez, (double*)gvec);
    else
        zbl_kernel<float><<<ab2_blocks(E, 128), 128, 0, st>>>(E, (float)p_cut, (float)qq, (const float*)vec, ctr, nbr, types, (const float*)Z,
                                                             (const float*)rmax_table, num_types, (float*)Ez, (float*)gvec);
    AB2_CUDA_LAUNCH_CHECK();
    return 0;
}
