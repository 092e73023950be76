// Streaming adjoint of the per-centre environment sum (ab2_env_bwd), same staging scheme as tp_stream.cu.
//
// Reference semantics: the backward of MakeWeightedChannels + scatter/gather
// (allegro/nn/_strided/_channels.py:44-57, _contract.py:195-205):
//     gw[z][l][u]  = sf * sum_{j in l} Y[z][j] * ggamma[c(z)][j][u]
//     gY[z][j]    += sf * sum_u w[z][l(j)][u] * ggamma[c(z)][j][u]
// Every edge is independent given its centre's ggamma row, so the CTA's contiguous edge range is cut into stages of TE
// edges: one elected producer thread brings the w rows in with 1-D bulk copies (UBLKCP), the producer warp's lanes copy
// the small Y rows and the centre indices with cp.async on the same mbarrier, and NCW consumer warps take the edges of
// a stage round-robin (lane = channel).  The centre's ggamma row (1152 B, shared by ~40 edges, L1/L2 resident) is
// reloaded into registers only when the centre changes.  880 B/edge of HBM traffic at fp32 / l_max = 2 / U = 32.
#include "common.cuh"
#include "stream_common.cuh"

int g_ab2_opt_env_stream = 1;
int g_ab2_opt_env_stream_cps = 0;  // cap on CTAs per SM of the streaming env adjoint (0 = min(occupancy, 8))

namespace {

struct EnvParams {
    int64_t N, E;
    int U;
    const int32_t* ctr;
    const void* Y;
    const void* w;
    const void* ggamma;
    float sf;
    void* gw;
    void* gY;
};

template <typename TAct, int LMAX, int NCH, int UT, int TE, int NS, int NCW>
__global__ void __launch_bounds__((NCW + 1) * 32) env_bwd_stream_kernel(const EnvParams p) {
    constexpr int D = (LMAX + 1) * (LMAX + 1), N_IR = LMAX + 1;
    constexpr int YP = (D + 3) / 4 * 4;  // padded Y row in shared memory (16-byte rows -> LDS.128 broadcasts)
    extern __shared__ __align__(128) uint8_t smem[];
    const int U = UT ? UT : p.U;
    const int row_el = N_IR * U;
    const int offW = 0;
    const int offY = (TE * row_el * (int)sizeof(TAct) + 127) & ~127;
    const int offC = offY + ((TE * YP * 4 + 127) & ~127);
    const int stage_bytes = offC + ((TE * 4 + 127) & ~127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // full[NS], empty[NS]
    uint8_t* ring = smem + 128;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (NS + s); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(full_bar(s), 33);  // expect_tx arrive of lane 0 + 32 cp.async (noinc) arrivals
            mbar_init(empty_bar(s), NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t G = gridDim.x, b = blockIdx.x;
    const int e_lo = (int)(b * p.E / G), e_hi = (int)((b + 1) * p.E / G);

    if (warp == NCW) {
        // =============================== producer ===============================
        int stage = 0;
        uint32_t phase = 0;
        for (int za = e_lo; za < e_hi; za += TE) {
            const int n = (e_hi - za) < TE ? (e_hi - za) : TE;
            if (lane == 0) mbar_wait_backoff(empty_bar(stage), phase ^ 1);
            __syncwarp();
            uint8_t* sb = ring + (size_t)stage * stage_bytes;
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)(n * row_el * sizeof(TAct));
                mbar_expect_tx(full_bar(stage), bytes);
                bulk_g2s(smem_u32(sb + offW), (const TAct*)p.w + (int64_t)za * row_el, bytes, full_bar(stage));
            }
            const float* __restrict__ ysrc = (const float*)p.Y + (int64_t)za * D;
            const uint32_t ydst = smem_u32(sb + offY), cdst = smem_u32(sb + offC);
            for (int e = lane; e < n * D; e += 32) {
                const int r = e / D, i = e - r * D;
                cp_async4(ydst + 4u * (r * YP + i), ysrc + e);
            }
            for (int e = lane; e < n; e += 32) cp_async4(cdst + 4u * e, p.ctr + za + e);
            cp_async_arrive_noinc(full_bar(stage));
            if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        return;
    }

    // =============================== consumers ===============================
    float gg[NCH][D];
    int cur = -1;
    int stage = 0;
    uint32_t phase = 0;
    const float sf = p.sf;
    for (int za = e_lo; za < e_hi; za += TE) {
        const int n = (e_hi - za) < TE ? (e_hi - za) : TE;
        mbar_wait(full_bar(stage), phase);
        const uint8_t* sb = ring + (size_t)stage * stage_bytes;
        const TAct* __restrict__ sW = reinterpret_cast<const TAct*>(sb + offW) + lane;
        const float* __restrict__ sY = reinterpret_cast<const float*>(sb + offY);
        const int* __restrict__ sC = reinterpret_cast<const int*>(sb + offC);
#pragma unroll 2
        for (int t = warp; t < n; t += NCW) {
            const int c = sC[t];
            if (c != cur) {  // warp-uniform: new centre -> its (scaled) ggamma row into registers
                cur = c;
#pragma unroll
                for (int q = 0; q < NCH; ++q)
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        const int u = q * 32 + lane;
                        gg[q][j] = (UT || u < U) ? sf * __ldg((const float*)p.ggamma + ((int64_t)c * D + j) * U + u) : 0.f;
                    }
            }
            float Yr[YP];
#pragma unroll
            for (int i4 = 0; i4 < YP / 4; ++i4) {
                const float4 y4 = *reinterpret_cast<const float4*>(sY + t * YP + 4 * i4);
                Yr[4 * i4] = y4.x; Yr[4 * i4 + 1] = y4.y; Yr[4 * i4 + 2] = y4.z; Yr[4 * i4 + 3] = y4.w;
            }
            float part[D];
#pragma unroll
            for (int j = 0; j < D; ++j) part[j] = 0.f;
            TAct* __restrict__ gw_row = (TAct*)p.gw + (int64_t)(za + t) * row_el + lane;
#pragma unroll
            for (int q = 0; q < NCH; ++q) {
                const bool live = UT ? true : (q * 32 + lane < U);
#pragma unroll
                for (int l = 0; l <= LMAX; ++l) {
                    const float wl = live ? to_acc<float>(sW[(t * N_IR + l) * U + q * 32]) : 0.f;
                    float gwl = 0.f;
#pragma unroll
                    for (int j = l * l; j < (l + 1) * (l + 1); ++j) {
                        gwl = fmaf(Yr[j], gg[q][j], gwl);
                        part[j] = fmaf(wl, gg[q][j], part[j]);
                    }
                    if (live) gw_row[l * U + q * 32] = from_acc<TAct>(gwl);
                }
            }
            // gY[z][j] += sum over channels: multi-value butterflies over groups of <= 8 values, one RED per (z, j)
            float* __restrict__ gy = (float*)p.gY + (int64_t)(za + t) * D;
#pragma unroll
            for (int j0 = 0; j0 < D; j0 += 8) {
                constexpr int DD = D;
                const int cnt = (DD - j0) < 8 ? (DD - j0) : 8;
                if (cnt == 8) {
                    float v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = part[j0 + k];
                    const float tot = MultiSum<8>::run(v, lane);
                    if (MultiSum<8>::is_writer(lane)) atomicAdd(gy + j0 + MultiSum<8>::idx_of(lane), tot);
                } else if (cnt == 4) {
                    float v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = part[(j0 + k) < DD ? (j0 + k) : 0];
                    const float tot = MultiSum<4>::run(v, lane);
                    if (MultiSum<4>::is_writer(lane)) atomicAdd(gy + j0 + MultiSum<4>::idx_of(lane), tot);
                } else {  // cnt == 1 (D = 9: the last component)
                    const float tot = warp_sum(part[j0 < DD ? j0 : 0]);
                    if (lane == 0) atomicAdd(gy + j0, tot);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(stage));
        if (++stage == NS) { stage = 0; phase ^= 1; }
    }
}

template <typename TAct, int LMAX, int NCH, int UT>
int launch(const EnvParams& p, cudaStream_t st) {
    constexpr int TE = 16, NS = 3, NCW = 4;
    constexpr int D = (LMAX + 1) * (LMAX + 1), N_IR = LMAX + 1, YP = (D + 3) / 4 * 4;
    auto kern = env_bwd_stream_kernel<TAct, LMAX, NCH, UT, TE, NS, NCW>;
    const int offY = (TE * N_IR * p.U * (int)sizeof(TAct) + 127) & ~127;
    const int stage_bytes = offY + ((TE * YP * 4 + 127) & ~127) + ((TE * 4 + 127) & ~127);
    const int smem = 128 + NS * stage_bytes;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    int cps = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, kern, (NCW + 1) * 32, smem) != cudaSuccess || cps < 1) {
        cudaGetLastError();
        return -1;
    }
    if (cps > 8) cps = 8;
    if (g_ab2_opt_env_stream_cps > 0 && cps > g_ab2_opt_env_stream_cps) cps = g_ab2_opt_env_stream_cps;
    int64_t grid = (int64_t)num_sms * cps;
    const int64_t max_grid = (p.E + TE - 1) / TE;
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, (NCW + 1) * 32, smem, st>>>(p);
    return 0;
}

template <typename TAct, int LMAX>
int launch_u(const EnvParams& p, cudaStream_t st) {
    if (p.U == 32) return launch<TAct, LMAX, 1, 32>(p, st);
    if (p.U < 32) return launch<TAct, LMAX, 1, 0>(p, st);
    if (p.U == 64) return launch<TAct, LMAX, 2, 64>(p, st);
    return -1;
}

}  // namespace

// returns 0 if launched, -1 if not eligible (caller falls back to the round-1 kernels)
int ab2_env_bwd_stream(int dtype, int lmax, int64_t N, int64_t E, int U, const int32_t* ctr, const void* Y, const void* w, int64_t w_ld,
                       const void* ggamma, double sf, void* gw, int64_t gw_ld, void* gY, cudaStream_t st) {
    if (!g_ab2_opt_env_stream || dtype == AB2_F64 || lmax < 1 || lmax > 2 || E <= 0 || E >= ((int64_t)1 << 31)) return -1;
    const int esz = dtype == AB2_F32 ? 4 : 2;
    const int n_ir = lmax + 1;
    if ((U * esz) % 16 != 0 || w_ld != (int64_t)n_ir * U || gw_ld != (int64_t)n_ir * U) return -1;
    if ((reinterpret_cast<uintptr_t>(w) & 15) != 0) return -1;
    EnvParams p;
    p.N = N; p.E = E; p.U = U; p.ctr = ctr; p.Y = Y; p.w = w; p.ggamma = ggamma; p.sf = (float)sf; p.gw = gw; p.gY = gY;
    if (dtype == AB2_F32) return lmax == 2 ? launch_u<float, 2>(p, st) : launch_u<float, 1>(p, st);
    return lmax == 2 ? launch_u<bf16, 2>(p, st) : launch_u<bf16, 1>(p, st);
}
