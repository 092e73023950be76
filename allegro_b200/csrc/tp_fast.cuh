// Internal (non-ABI) entry points of the register-tiled tensor-product kernels (tp_fast.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define AB2_FAST_MAXD 25

bool ab2_tp_fast_supported(int dtype, int D, int d_in, int d_out);
// return 0 on launch, -1 if the shape/dtype has no fast instantiation
int ab2_tp_fwd_fast(int dtype, int64_t N, int U, int D, int d_in, int d_out, int nnz, const int32_t* tab, const void* cgw,
                    const int32_t* row_ptr, const void* gamma, const void* Vin, int implicit_v0, const void* Y, const void* w0,
                    int64_t w0_ld, void* Vout, cudaStream_t st);
int ab2_tp_bwd_fast(int dtype, int64_t N, int U, int D, int d_in, int d_out, int nnz, const int32_t* tab, const void* cgw,
                    const int32_t* row_ptr, const void* gamma, const void* Vin, int implicit_v0, const void* Y, const void* w0,
                    int64_t w0_ld, const void* gVout, void* gVin, void* gw0, int64_t gw0_ld, void* gY, void* ggamma, cudaStream_t st);

// shared-memory-M kernels (tp_smem.cu): mode 0 forward, mode 1 backward part A (gin / gw0 / gY)
int ab2_tp_smem(int mode, int dtype, int64_t N, int U, int D, int d_in, int d_out, int nnz, const int32_t* tab, const void* cgw,
                const int32_t* row_ptr, const void* gamma, const void* Vin, int implicit_v0, const void* Y, const void* w0,
                int64_t w0_ld, void* Vout, const void* gVout, void* gVin, void* gw0, int64_t gw0_ld, void* gY, cudaStream_t st);

extern int g_ab2_opt_tp_fast;  // 0 generic, 1 fast (smem-M fwd, split bwd), 2 register-M kernels only

// TMA-staged streaming kernels (tp_stream.cu): mode 0 forward, mode 1 the whole backward in one launch
int ab2_tp_stream(int mode, int dtype, int64_t N, int64_t E, int U, int D, int d_in, int d_out, int nnz, const int32_t* tab,
                  const void* cgw, const int32_t* row_ptr, const int32_t* ctr, const void* gamma, const void* Vin, int implicit_v0, const void* Y,
                  const void* w0, int64_t w0_ld, void* Vout, const void* gVout, void* gVin, void* gw0, int64_t gw0_ld, void* gY,
                  void* ggamma, cudaStream_t st);
extern int g_ab2_opt_tp_stream, g_ab2_opt_tp_stream_te, g_ab2_opt_tp_stream_cps;
// three-consumer-warp layer-0 backward (tp_stream3.cu): fp32, U = 32, 9 x 9 -> 9 with the baked table structure
int ab2_tp_stream3_bwd(int64_t N, int64_t E, const int32_t* tab, const void* cgw, const int32_t* row_ptr, const int32_t* ctr, const void* gamma,
                       const void* Y, const void* w0, const void* gVout, void* gw0, void* gY, void* ggamma, cudaStream_t st);
extern int g_ab2_opt_tp_stream3;
// baked-structure fp64 kernels for the l_max = 3 layer shapes (tp_baked64.cu)
int ab2_tp_baked64(int mode, int dtype, int64_t E, int U, int D, int d_in, int d_out, int nnz, const int32_t* tab, const void* cgw,
                   const int32_t* ctr, const void* gamma, const void* Vin, int implicit_v0, const void* Y, const void* w0, int64_t w0_ld,
                   void* Vout, const void* gVout, void* gVin, void* gw0, int64_t gw0_ld, void* gY, void* ggamma, int** flag_out,
                   cudaStream_t st);
extern int g_ab2_opt_tp_baked64;
