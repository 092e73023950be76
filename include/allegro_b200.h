/* allegro_b200 C ABI -- B200 (sm_100a) kernels for Allegro's per-edge hot path.
 *
 * The reference (mir-group/allegro v0.7.1) is pure Python; its "FFI" for this path is the
 * kernel plug-in point Contracter.forward(x1, x2, idxs, scatter_dim_size)
 * (allegro/nn/_strided/_contract.py:185-211, swapped by enable_TritonContracter :253-282 and
 * enable_CuEquivarianceContracter :284-310) plus the graph modules whose forward(data) the
 * fused pipeline replaces (allegro/nn/tensorembed.py:85-96, allegro/nn/_allegro.py:237-301,
 * allegro/nn/edgewise.py:40-60).  Each entry point below cites the reference lines it replaces.
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer unless its name ends in _host;
 *  - the caller owns every buffer (kernels never allocate); work is enqueued on `stream`
 *    (a cudaStream_t passed as void*) and is asynchronous;
 *  - return 0 on success, non-zero on error; ab2_last_error() gives the message
 *    (thread-local);
 *  - dtype: AB2_F64 / AB2_F32 / AB2_BF16 selects the storage type of activations ("TAct").
 *    Accumulation type ("TAcc") is double for AB2_F64, float otherwise.  Geometry, spherical
 *    harmonics, environment sums, energies and gradients w.r.t. them are always TAcc.
 *  - edges are sorted by centre ("CSR": row_ptr[N+1], int32); internal feature layout is
 *    component-major  V[E][d][U]  (channel u fastest), env weights  w[E][n_ir][U].
 */
#ifndef ALLEGRO_B200_H
#define ALLEGRO_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { AB2_F64 = 0, AB2_F32 = 1, AB2_BF16 = 2 };
enum { AB2_ACT_NONE = 0, AB2_ACT_SILU = 1, AB2_ACT_MUL_DSILU = 2 };
enum { AB2_EPI_NONE = 0, AB2_EPI_MUL_DSILU = 1 };

#define AB2_MAX_SEG 4
#define AB2_MAX_LMAX 4

const char* ab2_last_error(void);
int ab2_version(void);
/* 1 if a CUDA device with compute capability 10.x is current, else 0 (no error). */
int ab2_device_ok(void);
/* Kernel-selection switches (for A/B tests and tuning; process-global, not thread-safe):
 *   "tp_fast"    1 shared-memory / register-tiled tensor product (default), 2 register-M variants, 0 shape-generic kernels
 *   "linear_tc"  1 tcgen05 tensor-core linear (default), 0 CUDA-core tile kernel
 *   "tp_variant" 1: 3 CTAs/SM (default), 0: 2 CTAs/SM for the shared-memory tensor-product kernel
 *   "tp_stream"  1 TMA-staged streaming tensor product where instantiated (default), 0 round-1 kernels;
 *                "tp_stream_te" edges per stage (0 = 8), "tp_stream_cps" cap on CTAs per SM (0 = occupancy limit),
 *                "tp_stream3" 1: three consumer warps per centre stream for the layer-0 backward (default), 0: two,
 *                "tp_stream_last" 1: 9 -> 1 (last layer) backward through the streaming kernel (default), 0: tp_smem + split,
 *                "tp_stream_gytile" 1: gY of the layer-0 backward reduced through a shared-memory tile (default), 0: shuffles
 *   "tp_baked64" 1 fp64 tensor-product kernels with the baked l_max = 3 table structure (default), 0 shape-generic kernels
 *   "env_stream" 1 streaming adjoint of the environment sum (default), 0 round-1 kernel
 *   "linear_tma" 1 TMA-producer variant of the tensor-core linear where eligible (default), 0 cp.async producers
 *   "env_split"  warps per (centre, channel chunk) in ab2_env_sum / ab2_env_bwd: 0 auto (default), 1, 2, 4
 *   "tc_debug"   stage knock-out mask of the tcgen05 linear (bit0 no stores, bit1 no loads, bit2 no MMA); results are
 *                wrong when non-zero -- for tools/exp_env.py only
 * Returns 1 (and sets ab2_last_error) for an unknown key. */
int ab2_set_option(const char* key, int value);

/* ---- operator level: the reference's own kernel plug-in point ----------------------- */

/* Contracter.forward, part 1 (_contract.py:195-204): gamma[n][u][j] += sf * x2[z][u][j] for
 * n = idxs[z] (reference "strided" layout [z][u][j], unsorted int64 idxs).  gamma must be
 * zeroed by the caller.  dtype AB2_F64 / AB2_F32. */
int ab2_op_scatter_env(int dtype, int64_t E, int64_t row /* = U*d2 */, double sf,
                       const void* x2, const int64_t* idxs, void* gamma, void* stream);

/* Contracter.forward, part 2 (_contract.py:205-251): out[z][u][k] =
 *   sum_nnz cgw[nnz][u] * x1[z][u][i_nnz] * gamma[idxs[z]][u][j_nnz]
 * with cgw[nnz][u] = w3j_value * weights[u, path]  (pre-contracted, _contract.py:218-219).
 * tab_ijk: int32 [nnz][3].  Also used for both backward products (caller permutes roles):
 *   mode 0: out[k] from (x1[i], g[j])      forward
 *   mode 1: gx1[i] from (gout[k], g[j])    (Triton "bwd1" table, _flashallegro.py:352-355)
 *   mode 2: gg[j]  from (x1[i], gout[k])   per-edge term, atomically added into
 *           ggamma[idxs[z]] (adjoint of the gather _contract.py:205); ggamma pre-zeroed. */
int ab2_op_contract(int dtype, int mode, int64_t E, int U, int d1, int d2, int dout, int nnz,
                    const int32_t* tab_ijk, const void* cgw, const void* a, const void* b,
                    const int64_t* idxs, void* out, void* stream);

/* Training support (the reference's `weights` are Parameters, _contract.py:170-177; its einsum path gets this product
 * from autograd, its Triton path is inference-only, _flashallegro.py:727):
 *   gcgw[n][u] += sum_z x1[z][u][i_n] * gamma[idxs[z]][u][j_n] * gout[z][u][k_n]      (gcgw pre-zeroed, atomics).
 * Together with modes 0-2 above every derivative of the trilinear form is one of these four products, which is how
 * allegro_b200.nn.Contracter provides weight gradients and double backward (forces in the loss). */
int ab2_op_contract_wgrad(int dtype, int64_t E, int U, int d1, int d2, int dout, int nnz,
                          const int32_t* tab_ijk, const void* x1, const void* gamma, const void* gout,
                          const int64_t* idxs, void* gcgw, void* stream);

/* Gather rows: out[z][:] = sf * src[idxs[z]][:]  (adjoint of the scatter; _contract.py:205). */
int ab2_op_gather_rows(int dtype, int64_t E, int64_t row, double sf, const void* src,
                       const int64_t* idxs, void* out, void* stream);

/* ---- fused pipeline (centre-sorted CSR edges, component-major layout) --------------- */

/* tensorembed.py:86,91-93: Y[z][0..d) = SH_{l<=lmax}(vec[z]/|vec[z]|), "component"
 * normalisation; vec, Y are TAcc.  (a1, a2) */
int ab2_sh_fwd(int acc_dtype, int lmax, int64_t E, const void* vec, void* Y, void* stream);
/* backward of the above: gvec[z] (+)= d Y/d vec ^T gY[z]   (SURVEY appendix B step 8). */
int ab2_sh_bwd(int acc_dtype, int lmax, int64_t E, const void* vec, const void* gY, void* gvec,
               int accumulate, void* stream);

/* Generic fused linear layer (nequip ScalarMLPFunction layer, _allegro.py:251,278;
 * tensorembed.py:88-89; allegro_models.py:231-241):
 *   Out[M][N] (split over <=4 column segments) (+)= epi( act(concat_k A_k)[M][K] @ W[K][N] )
 * act: AB2_ACT_SILU applies silu to A on load; AB2_ACT_MUL_DSILU multiplies segment s of A on
 * load by silu'(a_aux[s][m][k]) (a_aux_ptr_host[s] may be NULL = plain segment): the MLP backward's
 * g_pre = g_h * silu'(pre) formed inside the consuming GEMM's (prefetched) prologue.
 * epi: AB2_EPI_MUL_DSILU multiplies the result by silu'(aux[m][n]).  W is [K][N] row-major in TAct (alpha pre-folded on host).
 * A segments: (ptr, leading dim in elements, width); widths sum to K; outputs likewise to N. */
int ab2_linear(int dtype, int64_t M, int K, int N, int n_a, const void* const* a_ptr_host,
               const int64_t* a_ld_host, const int32_t* a_width_host,
               const void* const* a_aux_ptr_host /* nullable */, const int64_t* a_aux_ld_host,
               int act, const void* W,
               const void* W_packed /* nullable: image from ab2_linear_pack -> tcgen05 path */,
               int n_o, void* const* o_ptr_host, const int64_t* o_ld_host,
               const int32_t* o_width_host, const int32_t* o_accum_host, int epi, const void* aux,
               int64_t aux_ld, void* stream);

/* Tensor-core (tcgen05) path of ab2_linear.  ab2_linear_packed_bytes returns the size of the
 * packed weight image (bf16 hi + lo parts in the UMMA canonical K-major core-matrix layout) or 0
 * if (dtype, K, N) is not eligible (fp64, K % 16 != 0, K > 512); ab2_linear_pack builds it on the
 * device from W[K][N].  Any N is accepted: outputs wider than 256 columns, or whose W image does not
 * fit the shared-memory budget (128 KB), run as column slices (one launch per slice, W slice resident).
 * With fp32 storage the kernel computes A_hi W_hi + A_lo W_hi + A_hi W_lo in bf16 MMAs with fp32
 * accumulation (~2^-16 relative). */
int64_t ab2_linear_packed_bytes(int dtype, int K, int N);
int ab2_linear_pack(int dtype, int K, int N, const void* W, void* packed, void* stream);

/* _channels.py:44-57 + _contract.py:195-204 fused: gamma[c][j][u] =
 *   sf * sum_{z in row c} Y[z][j] * w[z][irrep(j)][u]      (a4, a7; deterministic, no atomics) */
int ab2_env_sum(int dtype, int lmax, int64_t N, int U, const int32_t* row_ptr, const void* Y,
                const void* w, int64_t w_ld, double sf, void* gamma, void* stream);

/* Adjoint of ab2_env_sum given ggamma[N][d][U]:
 *   gw[z][r][u] = sf * sum_{j in r} Y[z][j] ggamma[c][j][u]
 *   gY[z][j]   += sf * sum_u w[z][irrep(j)][u] ggamma[c][j][u]     (appendix B steps 3-4) */
int ab2_env_bwd(int dtype, int lmax, int64_t N, int64_t E, int U, const int32_t* row_ptr,
                const int32_t* ctr, const void* Y, const void* w, int64_t w_ld, const void* ggamma,
                double sf, void* gw, int64_t gw_ld, void* gY, void* stream);

/* _contract.py:205-251 for one layer on the fused layout (a8, a10):
 *   Vout[z][k][u] = sum_i Vin[z][i][u] * M_c[u][i][k],
 *   M_c[u][i][k]  = sum_nnz cgw[nnz][u] * gamma[c][j_nnz][u]      (built once per centre)
 * tab_ijk must be sorted by (i, k) (entries of one output target contiguous): the fast kernels
 * gather every M[i][k] from its table segment.
 * implicit_v0 != 0: Vin[z][i][u] = Y[z][i] * w0[z][irrep(i)][u] is formed on the fly
 * (tensorembed.py:95) and never stored. */
int ab2_tp_fwd(int dtype, int lmax, int64_t N, int64_t E, int U, int d_in, int d_out, int nnz,
               const int32_t* tab_ijk, const void* cgw, const int32_t* row_ptr, const int32_t* ctr,
               const void* gamma, const void* Vin, int implicit_v0, const void* Y, const void* w0,
               int64_t w0_ld, void* Vout, void* stream);

/* Backward of ab2_tp_fwd (appendix B steps 1-2): given gVout,
 *   gVin[z][i][u] = sum_k M_c[u][i][k] gVout[z][k][u]
 *   ggamma[c][j][u] = sum_nnz cgw * sum_{z in c} Vin[z][i][u] gVout[z][k][u]
 * implicit_v0: instead of gVin writes gw0[z][r][u] and accumulates into gY[z][i]. */
int ab2_tp_bwd(int dtype, int lmax, int64_t N, int64_t E, int U, int d_in, int d_out, int nnz,
               const int32_t* tab_ijk, const void* cgw, const int32_t* row_ptr, const int32_t* ctr,
               const void* gamma, const void* Vin, int implicit_v0, const void* Y, const void* w0,
               int64_t w0_ld, const void* gVout, void* gVin, void* gw0, int64_t gw0_ld, void* gY,
               void* ggamma, void* stream);

/* edgewise.py:40-60 (a12): Ei[c] = factor * sum_{z in row c} Ez[z]   (TAcc, deterministic). */
int ab2_edge_sum(int acc_dtype, int64_t N, const int32_t* row_ptr, const void* Ez, double factor,
                 void* Ei, void* stream);
/* adjoint: gEz[z] = factor * gEi[ctr[z]] */
int ab2_edge_sum_bwd(int acc_dtype, int64_t E, const int32_t* ctr, const void* gEi, double factor,
                     void* gEz, void* stream);

/* Force assembly (appendix B step 9): F[a] = sum_{z in CSR row a} g[z] - sum_{z: nbr[z]=a} g[z].
 * Both sums are segmented reductions in a fixed order (deterministic, no atomics, F need not be
 * zeroed): the neighbour side walks the TRANSPOSED CSR, col_ptr[n_total+1] / col_perm[E] = edge ids
 * grouped by neighbour atom (built once per neighbour list).  N = number of centres (owned atoms),
 * n_total = rows of F (owned + ghost atoms, allegro/_compile.py:41-61).  F: acc dtype. */
int ab2_force_scatter(int acc_dtype, int64_t N, int64_t n_total, int64_t E, const int32_t* row_ptr,
                      const int32_t* col_ptr, const int32_t* col_perm, const void* gvec, void* F,
                      void* stream);

/* ---- upstream two-body scalar track + geometry ------------------ */

/* nequip with_edge_vectors_ (tensorembed.py:86): vec[z] = pos[nbr[z]] - pos[ctr[z]] (+ shift[z]),
 * computed in the positions' dtype (AB2_F64 / AB2_F32), stored in the accumulate dtype.
 * shift = edge_cell_shift @ cell, nullable. */
int ab2_edge_vec(int pos_dtype, int acc_dtype, int64_t E, const void* pos, const int32_t* ctr,
                 const int32_t* nbr, const void* shift, void* vec, void* stream);

/* EdgeLengthNormalizer + BesselEdgeLengthEncoding*PolynomialCutoff + ProductTypeEmbedding
 * (allegro_models.py:153-157, scalarembed.py:60-81, _edgeembed.py:68-85):
 *   x = |vec| / rmax_table[t_c][t_n];  B_n = sin(pi w_n x)/(pi x) * f_p(x);
 *   e0[z][c] = (c < S_rc/2 ? center_embed[t_c][c] : neighbor_embed[t_n][c - S_rc/2]) * sum_n B_n Wb[n][c]
 * Tables are in the accumulate dtype; e0 in the activation dtype. */
int ab2_radial_fwd(int dtype, int64_t E, int S_rc, int num_bessels, double p_cut, const void* vec,
                   const int32_t* ctr, const int32_t* nbr, const int32_t* types,
                   const void* rmax_table, int num_types, const void* bessel_w, const void* Wb,
                   const void* center_embed, const void* neighbor_embed, void* e0, void* stream);
/* adjoint: gvec[z] += (d e0 / d vec)^T g_e0[z] */
int ab2_radial_bwd(int dtype, int64_t E, int S_rc, int num_bessels, double p_cut, const void* vec,
                   const int32_t* ctr, const int32_t* nbr, const int32_t* types,
                   const void* rmax_table, int num_types, const void* bessel_w, const void* Wb,
                   const void* center_embed, const void* neighbor_embed, const void* g_e0,
                   void* gvec, void* stream);

/* ---- neighbour list on the device, directly in CSR --------------- */

/* Cell-list search on an orthorhombic box with per-axis periodicity.  The reference receives edge_index [2,E] int64
 * from nequip's data pipeline / LAMMPS (allegro/nn/_allegro.py:238, allegro/_compile.py:41-61); these three kernels
 * produce the centre-sorted CSR (row_ptr / nbr int32) and the per-edge shift VECTORS the path consumes, without the
 * int64 COO list.  Host-side geometry: box[3], origin[3] (doubles), pbc[3], ncell[3] with box/ncell >= r_max and
 * >= 3 cells on every periodic axis.  pos: [n][3] fp64 or fp32, raw (unwrapped) coordinates;
 *   r = pos[nbr] + shift - pos[centre]  holds for the raw positions.
 * Call order: ab2_nl_bin -> (host: order = stable argsort(cell_id), cell_start = prefix sum of the cell histogram)
 *             -> ab2_nl_count -> (host: row_ptr = prefix sum) -> ab2_nl_fill.
 * Centres are atoms [0, n_centres) (owned atoms first, ghosts after: only owned atoms get rows). */
int ab2_nl_bin(int pos_dtype, int64_t n, const void* pos, const double* box_host, const double* origin_host,
               const int32_t* pbc_host, const int32_t* ncell_host, double r_max, int32_t* cell_id, void* stream);
int ab2_nl_count(int pos_dtype, int64_t n_centres, const void* pos, const double* box_host,
                 const double* origin_host, const int32_t* pbc_host, const int32_t* ncell_host, double r_max,
                 const int32_t* cell_start, const int32_t* order, int32_t* counts, void* stream);
int ab2_nl_fill(int pos_dtype, int64_t n_centres, const void* pos, const double* box_host,
                const double* origin_host, const int32_t* pbc_host, const int32_t* ncell_host, double r_max,
                const int32_t* cell_start, const int32_t* order, const int32_t* row_ptr, int32_t* nbr,
                void* shift, void* stream);

/* Radial embedding with per-type-pair matrices: out[z][c] = sum_n B_n(x_z) PQ[t_c * T + t_n][n][c], B_n as above
 * (num_bessels must be 8, S <= 128).  PQ: [T*T][8][S] in the accumulate dtype.  The product embedding above is
 * PQ = typeemb(t_c,t_n)[c] * Wb[n][c]; because everything up to the first nonlinearity is linear
 * (allegro/nn/_edgeembed.py:68-85, allegro_models.py:153-183) the host may fold the first scalar_embed_mlp layer in,
 * PQ = Wb diag(typeemb) W_1, and receive that layer's pre-activation directly.
 * bwd: gvec[z] += (d out / d vec)^T (g_out[z] * silu'(aux[z]))   (aux nullable = plain g_out). */
int ab2_radial_pq_fwd(int dtype, int64_t E, int S, int num_bessels, double p_cut, const void* vec,
                      const int32_t* ctr, const int32_t* nbr, const int32_t* types, const void* rmax_table,
                      int num_types, const void* bessel_w, const void* PQ, void* out, void* stream);
int ab2_radial_pq_bwd(int dtype, int64_t E, int S, int num_bessels, double p_cut, const void* vec,
                      const int32_t* ctr, const int32_t* nbr, const int32_t* types, const void* rmax_table,
                      int num_types, const void* bessel_w, const void* PQ, const void* g_out, const void* aux,
                      void* gvec, void* stream);

/* ZBL pair term (reference call site allegro/model/allegro_models.py:270-288; the module is nequip's
 * nequip.nn.pair_potential.ZBL = LAMMPS pair_style zbl, constants of pair_zbl_const.h):
 *   Ez[z] = qq * Z_i Z_j / r * phi((Z_i^0.23 + Z_j^0.23) r / 0.46850) * u(r / rmax_table[t_c][t_n]),
 *   phi(x) = 0.18175 e^{-3.19980x} + 0.50986 e^{-0.94229x} + 0.28022 e^{-0.40290x} + 0.02817 e^{-0.20162x},
 * u = polynomial cutoff of order p_cut, qq = qqr2e / 2 (each pair is two directed edges); and
 *   gvec[z] += dEz/dvec[z].
 * vec, Z [num_types], rmax_table [num_types^2], Ez [E], gvec [E][3] in the accumulate dtype; Ez or gvec may be null. */
int ab2_zbl(int acc_dtype, int64_t E, int num_types, double p_cut, double qq, const void* vec, const int32_t* ctr,
            const int32_t* nbr, const int32_t* types, const void* Z, const void* rmax_table, void* Ez, void* gvec,
            void* stream);

/* ---- ghost-atom halo exchange over NVLink peer memory ------------------- */

/* One mailbox per rank (cudaMalloc'ed here so that it can be exported through CUDA IPC), mapped by its peers.
 * Per step: ab2_p2p_begin (bumps the step counter), ab2_p2p_push_rows into the neighbours' mailboxes (positions of
 * my boundary atoms / gradients of my ghosts), ab2_p2p_wait_unpack of what they pushed into mine, and
 * ab2_p2p_allreduce_energy -- kernels only (peer stores + system-scope release/acquire flags), so the whole step
 * including the halo replays from one CUDA graph with no NCCL call.  Waits are bounded (~2 s): a protocol error sets
 * the mailbox's error word (ab2_p2p_error) instead of hanging the GPU.  There is no reference counterpart (the
 * reference leaves domain decomposition to LAMMPS' MPI, allegro/_compile.py:41-61 only defines the ghost format). */
int64_t ab2_p2p_mailbox_bytes(int max_rows, int world);
int ab2_p2p_alloc(int64_t bytes, void** ptr);
int ab2_p2p_free(void* ptr);
int ab2_p2p_get_handle(void* ptr, void* handle64_host);
int ab2_p2p_open_handle(const void* handle64_host, void** ptr);
int ab2_p2p_close_handle(void* ptr);
int ab2_p2p_error(void* my_mailbox, int max_rows, int world, void* stream);
int ab2_p2p_begin(void* step_counter, void* stream);
int ab2_p2p_push_rows(int src_dtype, int kind, int side, const void* src, const int64_t* idx, int n, double shift_x,
                      void* peer_mailbox, int max_rows, int world, const void* step_counter, void* done_counter,
                      void* stream);
int ab2_p2p_wait_unpack(int dst_dtype, int kind, int side, void* my_mailbox, int max_rows, int world,
                        const void* step_counter, int n, void* dst, const int64_t* idx, int accumulate, void* stream);
int ab2_p2p_allreduce_energy(const void* e_local, int rank, int world, int max_rows, void* const* peers_dev,
                             void* my_mailbox, const void* step_counter, void* e_total, void* stream);

/* layout helpers between the reference strided layout [z][u][i] and the internal [z][i][u] */
int ab2_transpose_ui(int dtype, int64_t E, int U, int d, const void* src, void* dst, int to_internal,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif
