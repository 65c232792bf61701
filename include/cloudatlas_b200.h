/* cloudatlas_b200.h -- C ABI of libcloudatlas_b200.so
 *
 * Drop-in boundary for the data-parallel hot path of johnfgibson/CloudAtlas.jl: building and
 * evaluating  dx/dt = B^-1 (A1 x + A2 x / R + N(x,x))  on NVIDIA B200 (sm_100a).
 * Every entry point names the reference interface it replaces (file:line under the reference
 * repository's src/).  Plain pointers and sizes only; no CUDA, torch or C++ types cross the boundary.
 *
 * Conventions
 *   - every function returns an int32 status, 0 = CA_OK; on failure ca_last_error() (thread-local)
 *     describes it.  The Julia wrapper turns a non-zero status into error(msg), matching the
 *     reference's error() calls (SparseBilinear.jl:8-9).
 *   - matrices are column-major (Julia layout).  Index arrays are 1-based int64, exactly as the
 *     reference's N.ijk (nnz x 3, column-major: all i, then all j, then all k).
 *   - ensembles of states are nstates x m column-major: X[s + nstates*j] (state index fastest).
 *   - the caller owns every array it passes; handles own their device memory; *_destroy frees it.
 *   - a handle lives on the CUDA device that was current when it was created.
 *   - there is NO CPU fallback: without a usable sm_100 device every compute entry point fails.
 *   - entry points ending in _dev take DEVICE pointers and a cudaStream_t (as void*, NULL = default
 *     stream) and are asynchronous; all others take HOST pointers and return when the result is there.
 */
#ifndef CLOUDATLAS_B200_H
#define CLOUDATLAS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  CA_OK = 0,
  CA_ERR_INVALID = 1,   /* bad argument (shape, index out of range, null pointer) */
  CA_ERR_CUDA = 2,      /* CUDA runtime error (sticky; see ca_last_error) */
  CA_ERR_NODEVICE = 3,  /* no CUDA device / not an sm_100 device */
  CA_ERR_NOMEM = 4,
  CA_ERR_UNSUPPORTED = 5
};

/* Symmetry(sx,sy,sz,ax,az,s) -- src/Symmetries.jl:3-14.  ax2 = 2*ax, az2 = 2*az (0 or 1: only half-box
 * shifts are implemented by the reference, Symmetries.jl:33-43). */
typedef struct ca_symmetry {
  int32_t sx, sy, sz;
  int32_t ax2, az2;
  int32_t s;
} ca_symmetry;

typedef struct ca_bilinear ca_bilinear; /* SparseBilinear{Float64}, packed on the device */
typedef struct ca_model ca_model;       /* ODEModel{Float64}: B, A1, A2, N and the folded operators */
typedef struct ca_assembly ca_assembly; /* result of a Galerkin assembly (a slab of output modes) */

/* ---- library ------------------------------------------------------------------------------ */
const char* ca_last_error(void);
int32_t ca_version(void);
/* number of CUDA devices visible, name + SM count + compute capability of the current one */
int32_t ca_device_info(int32_t* ndevices, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor,
                       char* name, int32_t name_len);
/* select / query the CUDA device of the calling host thread (cudaSetDevice / cudaGetDevice): handles are created
 * on the current device, so a single process drives several GPUs by calling this before ca_*_create / ca_assemble,
 * one host thread (Julia task) per device */
int32_t ca_set_device(int32_t device);
int32_t ca_get_device(int32_t* device);
/* page-lock / unlock a caller-owned host buffer so the host-pointer entry points can overlap
 * copies with kernels (Julia: call on the Array's pointer; optional, results identical) */
int32_t ca_pin(void* host_ptr, int64_t bytes);
int32_t ca_unpin(void* host_ptr);
/* FP64 roofline denominators of the current device, measured with register-resident loops of
 * mma.sync.m8n8k4.f64 (DMMA) and of DFMA, best of a few repetitions (MEASURED_PEAKS.json has no FP64
 * entry; SURVEY.md section 8d asks for an on-box measurement in the same run). */
int32_t ca_measure_fp64_peak(double* dmma_tflops, double* dfma_tflops);
/* number of kernels launched by this library in the calling process so far (bench bookkeeping) */
int64_t ca_launch_count(void);

/* ---- several GPUs of one box ---------------------------------------------------------------- */
/* A communicator over NCCL (resolved at run time: dlopen of libnccl.so.2; single-GPU use never loads it).
 *   ca_comm_init_all: ONE process drives ngpus devices (devices == NULL: 0..ngpus-1) -- the Julia case, no NCCL.jl
 *     needed; handles come back one per device, in device order.
 *   ca_comm_unique_id / ca_comm_init_rank: one process per GPU (torchrun, MPI); rank 0 creates the 128-byte id, the
 *     launcher's own transport delivers it, every rank joins with the device that is current in its process.
 * ca_init(ngpus) / ca_shutdown keep a process-wide default communicator that entry points taking `comm == NULL` use
 * (SURVEY.md 8b). */
typedef struct ca_comm ca_comm;
int32_t ca_init(int32_t ngpus);
int32_t ca_shutdown(void);
int32_t ca_comm_init_all(int32_t ngpus, const int32_t* devices, ca_comm** out);
int32_t ca_comm_unique_id(char* id128);
int32_t ca_comm_init_rank(int32_t nranks, int32_t rank, const char* id128, ca_comm** out);
int32_t ca_comm_info(const ca_comm* c, int32_t* nranks, int32_t* nlocal, int32_t* first_rank, int32_t* nccl_version);
int32_t ca_comm_destroy(ca_comm* c);

/* ---- index sets and basis (reference: src/BasisFunctions.jl, src/Symmetries.jl) --------------- */
/* basisIndices(J,K,L[,H]) -- BasisFunctions.jl:489-574 with the symmetric() filter of
 * Symmetries.jl:25-60.  Host only, integer only, bit-exact incl. row order.  nH = 0: unfiltered.
 * ijkl_out == NULL: size query (*m_inout receives the row count).  Otherwise *m_inout is the capacity
 * in rows on entry and the row count on exit; ijkl_out is m x 4 column-major. */
int32_t ca_basis_indices(int32_t J, int32_t K, int32_t L, const ca_symmetry* H, int32_t nH,
                         int64_t* ijkl_out, int64_t* m_inout);
/* symmetric(ijkl, H) -- Symmetries.jl:25-60: *out = 1 iff sigma Psi_ijkl = Psi_ijkl for every generator.  A pure sign
 * computation on one row (i,j,k,l), defined for any row with i in 1:6 whether or not basisIndices would emit it. */
int32_t ca_symmetric(const int64_t* ijkl4, const ca_symmetry* H, int32_t nH, int32_t* out);
/* basisSet(alpha, gamma, ijkl; normalize) in tabular form -- BasisFunctions.jl:630-702: per mode and
 * velocity component (u,v,w) the factor  coef * E_jx(x) E_kz(z) S_l^(d)(y).  Arrays are m x 3
 * column-major; d = 0 for S_l, 1 for S_l'; coef == 0 marks an absent component. */
int32_t ca_basis_components(double alpha, double gamma, const int64_t* ijkl, int64_t m, int32_t normalize,
                            double* coef, int32_t* jx, int32_t* kz, int32_t* l, int32_t* d);

/* Psi_shear of shear_function (ODEModels.jl:65-78): shear(x) = 1 + dot(x, Psi_shear).  Host only. */
int32_t ca_basis_shear(double alpha, double gamma, const int64_t* ijkl, int64_t m, int32_t normalize,
                       double* shear_out);

/* ---- SparseBilinear (reference: src/SparseBilinear.jl) -------------------------------------- */
/* SparseBilinear(ijk, val, m)  -- SparseBilinear.jl:7-11.  ijk: nnz x 3 column-major, 1-based.
 * Copies the COO, builds the device packing.  Fails with CA_ERR_INVALID where the reference
 * constructor calls error() (shape) and additionally on out-of-range indices. */
int32_t ca_bilinear_create(const int64_t* ijk, const double* val, int64_t nnz, int64_t m,
                           ca_bilinear** out);
int32_t ca_bilinear_destroy(ca_bilinear* h);
int32_t ca_bilinear_info(const ca_bilinear* h, int64_t* m, int64_t* nnz, int64_t* nnz_sym,
                         int64_t* npairs, int64_t* padded_fma_per_state);
/* N(x,y): out_i = sum_r val_r x[j_r] y[k_r]  -- SparseBilinear.jl:56-65.  y == NULL means N(x)
 * = N(x,x) (SparseBilinear.jl:70).  x, y, out: m doubles on the host. */
int32_t ca_bilinear_eval(ca_bilinear* h, const double* x, const double* y, double* out);
/* N(x_s) for every row s of X (additive entry point: the reference evaluates one state per call).
 * X, OUT: nstates x m column-major, host.  Chunks of whole kernel waves on three streams: copies overlap kernels. */
int32_t ca_bilinear_eval_batched(ca_bilinear* h, const double* X, int64_t nstates, double* OUT);
/* same on device memory; ldx/ldo = leading dimensions (>= nstates) */
int32_t ca_bilinear_eval_batched_dev(ca_bilinear* h, const double* dX, int64_t nstates, int64_t ldx,
                                     double* dOUT, int64_t ldo, void* stream);
/* N(x_s, y_s) with two ensembles (ordered product x_j y_k), device memory */
int32_t ca_bilinear_eval2_batched_dev(ca_bilinear* h, const double* dX, const double* dY,
                                      int64_t nstates, int64_t ld, double* dOUT, int64_t ldo,
                                      void* stream);
/* derivative(N, x): DN[i,j] = sum_k (N_ijk + N_ikj) x_k, dense m x m column-major
 * -- SparseBilinear.jl:75-91.  Host pointers. */
int32_t ca_bilinear_jacobian(ca_bilinear* h, const double* x, double* DN);
int32_t ca_bilinear_jacobian_dev(ca_bilinear* h, const double* dx, double* dDN, void* stream);
/* derivative(N, x_s) for every state of an ensemble (additive: the reference evaluates one state at a time,
 * SparseBilinear.jl:75-91).  X: nstates x m column-major; DN: nstates x m x m column-major, i.e.
 * DN[s + ld*(i + m*j)] = derivative(N, X[s,:])[i,j] -- the state index is the fastest, like every ensemble here. */
int32_t ca_bilinear_jacobian_batched(ca_bilinear* h, const double* X, int64_t nstates, double* DN);
int32_t ca_bilinear_jacobian_batched_dev(ca_bilinear* h, const double* dX, int64_t nstates, int64_t ldx,
                                         double* dDN, int64_t ldo, void* stream);
/* matrix-free Jacobian action  DN(x) v = N(x,v) + N(v,x)  for ensembles (Newton-Krylov at m ~ 3000) */
int32_t ca_bilinear_jvp_batched_dev(ca_bilinear* h, const double* dX, const double* dV,
                                    int64_t nstates, int64_t ld, double* dOUT, int64_t ldo,
                                    void* stream);

/* ---- ODEModel (reference: src/ODEModels.jl) ------------------------------------------------- */
/* Builds the evaluation side of an ODEModel from its operators: B, A1, A2 (m x m column-major)
 * and N as COO (as ca_bilinear_create).  Plays the role of  Bfact = lu(B)  and the closures f, Df
 * (ODEModels.jl:54-58): B is factored once (block-diagonal LU with partial pivoting) and folded into
 * the operators,  L1 = B^-1 A1,  L2 = B^-1 A2,  Q = B^-1 N,  so that
 *     f(x,R) = (L1 + L2/R) x + Q(x,x)
 * is ONE pass of the batched kernel: each state is read once and written once. */
int32_t ca_model_create(const double* B, const double* A1, const double* A2, int64_t m,
                        const int64_t* ijk, const double* val, int64_t nnz, ca_model** out);
/* The same from DEVICE arrays (as ca_assembly_get_*_dev / the all-gathered slabs deliver them; B, A1, A2 m x m
 * column-major, the COO as three index arrays, 1-based, in the reference's lexicographic (i,j,k) order): the operator
 * never passes through the host.  Folding, symmetrisation, pair-set unions and the placement of the values into the
 * DMMA fragment layout run on the device; the packer's decisions (row clustering, windows, warp schedule) are made by
 * the same host code as in ca_model_create on descriptors of a few MB, so both entry points build the same pack bit for
 * bit.  ca_model_create itself uploads its arguments and takes this path; operators it does not cover (COO not in
 * ascending (i,j,k) order or with duplicates, a block of B with more than 64 modes) go through the host build. */
int32_t ca_model_create_dev(const double* dB, const double* dA1, const double* dA2, int64_t m, const int64_t* d_i,
                            const int64_t* d_j, const int64_t* d_k, const double* d_val, int64_t nnz, void* stream,
                            ca_model** out);
/* ODEModel straight from a full assembly (all rows: ca_assemble(..., 0, m) or an output of ca_assemble_sharded), on
 * the assembly's device: assembled operator -> folded, packed model without leaving the GPU; also installs the
 * observables (shear vector and D of ODEModels.jl:65-137) that the assembly computed. */
int32_t ca_model_create_from_assembly(ca_assembly* a, ca_model** out);
/* f(x_s, R) of ONE host ensemble on several devices of this process: models[d] are replicas of one model on different
 * devices; states are split evenly, one host thread and one copy/compute pipeline per device, no collective. */
int32_t ca_models_rhs_batched(ca_model** models, int32_t nmodels, const double* X, int64_t nstates, double R, double* OUT);
/* How the model was built: *on_device = 1 for the device build; seconds9 = wall seconds {total, blocks of B, fold,
 * symmetrise + CSR, sketches + clustering, tile unions, layout (host), value placement, streaming form}. */
int32_t ca_model_build_info(const ca_model* h, int32_t* on_device, double* seconds9, int64_t* nclusters,
                            int64_t* ntiles, int64_t* nterms);
/* FNV-1a digests of the build products (diagnostics; device build vs host build must agree in all twelve):
 * [0..7] the packed operator (as ca_selftest_pack_digest), [8] L1, [9] L2, [10] linear slots, [11] streaming form. */
int32_t ca_model_digest(ca_model* h, uint64_t* digest12);
int32_t ca_model_destroy(ca_model* h);
int32_t ca_model_info(const ca_model* h, int64_t* m, int64_t* nnz_q_sym, int64_t* nnz_lin,
                      int64_t* npairs, int64_t* padded_fma_per_state, int64_t* nblocks_B);
/* Which kernel serves an ensemble of nstates states: 0 = DMMA row tiles, whole tiles per warp; 1 = few-state CSR
 * stream; 2 = model-specialised straight-line kernel (NVRTC, m <= 48); 3 = DMMA row tiles, k-steps split across
 * warps (unbalanced tile schedules); 4 = windowed DMMA row tiles (large m).  *source / *log (may be NULL) receive
 * pointers to the generated CUDA source and the NVRTC log of path 2, valid while the handle lives. */
int32_t ca_model_kernel_path(ca_model* h, int64_t nstates, int32_t* path, const char** source, const char** log);
/* f(x,R) = Bfact \ (A1*x + (1/R)*(A2*x) + N(x))  -- ODEModels.jl:57.  Host pointers, one state. */
int32_t ca_model_rhs(ca_model* h, const double* x, double R, double* out);
/* f(x_s,R) for every row of X (nstates x m column-major, host) */
int32_t ca_model_rhs_batched(ca_model* h, const double* X, int64_t nstates, double R, double* OUT);
int32_t ca_model_rhs_batched_dev(ca_model* h, const double* dX, int64_t nstates, int64_t ldx, double R,
                                 double* dOUT, int64_t ldo, void* stream);
/* Df(x,R) = Bfact \ (A1 + (1/R)*A2 + derivative(N,x))  -- ODEModels.jl:58.  m x m column-major. */
int32_t ca_model_jac(ca_model* h, const double* x, double R, double* Df);
int32_t ca_model_jac_dev(ca_model* h, const double* dx, double R, double* dDf, void* stream);
/* Df(x_s,R) for every state of an ensemble, dense: Df[s + ld*(i + m*j)] (nstates x m x m column-major). */
int32_t ca_model_jac_batched(ca_model* h, const double* X, int64_t nstates, double R, double* Df);
int32_t ca_model_jac_batched_dev(ca_model* h, const double* dX, int64_t nstates, int64_t ldx, double R,
                                 double* dDf, int64_t ldo, void* stream);
/* Df(x_s,R) v_s for ensembles, matrix-free */
int32_t ca_model_jvp_batched_dev(ca_model* h, const double* dX, const double* dV, int64_t nstates,
                                 int64_t ld, double R, double* dOUT, int64_t ldo, void* stream);

/* ---- Galerkin assembly (reference: src/ODEModels.jl:31-51 over src/BasisFunctions.jl) -------- */
/* Computes, on the device, for the output modes i in [row_begin, row_end) (0-based, half-open):
 *   B_ij  = <Psi_i, Psi_j>                                  ODEModels.jl:32
 *   A1_ij = -<Psi_i, y d/dx Psi_j> - <Psi_i, [v_j,0,0]>     ODEModels.jl:33-36
 *   A2_ij = <Psi_i, laplacian Psi_j>                        ODEModels.jl:35
 *   N_ijk = -<Psi_i, (Psi_j . grad) Psi_k>                  ODEModels.jl:39-49
 * and N's COO in lexicographic (i,j,k) order (SparseBilinear.jl:24-49).  Inner products are exact
 * Gauss-Legendre contractions in y times closed-form Fourier averages (no monomial coefficients, so
 * no loss of digits with L).  An N entry is kept iff |N| > 1e-15 (the reference's cut,
 * ODEModels.jl:46) and |N| > 1e-11 * sum|terms| (discards entries that vanish by cancellation).
 * Slabs of different ranks concatenate, in rank order, to the full operator (sharding by i). */
int32_t ca_assemble(double alpha, double gamma, const int64_t* ijkl /* m x 4 column-major */, int64_t m,
                    int32_t normalize, int64_t row_begin, int64_t row_end, ca_assembly** out);
/* The same operators with N computed by brute-force quadrature on the exact tensor grid (uniform Nx > 3J,
 * Nz > 3K points in x, z; ny >= (3L+13)/2 Gauss-Legendre nodes in y; 0 = smallest exact): a dense FP64
 * contraction  N_ijk = -sum_q W_q Phi_i(q) . (Phi_j . grad) Phi_k (q)  on the tensor pipe, B operand generated on
 * the fly, result compacted to the same COO.  ~2000x the arithmetic of ca_assemble for the same numbers: it is
 * the FP64-roofline benchmark of the assembly stage and an independent check.  An entry is kept iff
 * |N| > max(1e-15, 1e-13 max|N|).  The dense scratch is (row_end-row_begin) m^2 doubles (<= 100 GB). */
int32_t ca_assemble_dense(double alpha, double gamma, const int64_t* ijkl, int64_t m, int32_t normalize,
                          int64_t row_begin, int64_t row_end, int32_t nx, int32_t ny, int32_t nz,
                          ca_assembly** out);
/* Both paths for a basis with coefficients -- the reference's data model  c * Psi  (BasisFunctions.jl:417; its
 * normalisation :697-699 is the special case c_n = 1/||Psi_n||):
 *   mode_scale (m values, may be NULL): Psi_n -> s_n Psi_n, hence N'_ijk = s_i s_j s_k N_ijk, B'_ij = s_i s_j B_ij;
 *   mix (m x m column-major C, may be NULL, dense != 0 only): Phi_n = sum_p Psi_p C[p,n], a basis whose members do
 *   not factorise in x, y, z -- only the dense quadrature applies; N' = N x1 C x2 C x3 C is dense, so assemble row
 *   slabs.  B', A1', A2', D' = C' (.) C.
 * dense = 0: ca_assemble; dense != 0: ca_assemble_dense on the grid nx, ny, nz (0 = smallest exact). */
int32_t ca_assemble_basis(double alpha, double gamma, const int64_t* ijkl, int64_t m, int32_t normalize,
                          const double* mode_scale, const double* mix, int32_t dense, int32_t nx, int32_t ny, int32_t nz,
                          int64_t row_begin, int64_t row_end, ca_assembly** out);
/* The assembly sharded by output mode i over the ranks of `comm` (NULL: the ca_init communicator) with its single
 * collective: every rank assembles a contiguous range of rows, slab sizes are exchanged, and all pieces
 * of all slabs are broadcast into place inside ONE NCCL group over NVLink.  out[d], d < nlocal, receives for local
 * device d a FULL assembly (rows 0..m, lexicographic COO identical on every device and identical to the single-GPU
 * result) -- ready for ca_model_create_from_assembly.  dense != 0: the quadrature path (slabs aligned to 128 rows). */
int32_t ca_assemble_sharded(ca_comm* comm, double alpha, double gamma, const int64_t* ijkl, int64_t m,
                            int32_t normalize, const double* mode_scale, int32_t dense, int32_t nx, int32_t ny,
                            int32_t nz, ca_assembly** out);
/* ranks that took part, CUDA-event seconds of the collective on this device, bytes of the gathered operator */
int32_t ca_assembly_gather_info(const ca_assembly* h, int32_t* world, double* gather_seconds, int64_t* gather_bytes);
/* grid used, CUDA-event seconds of the contraction kernel alone and its algorithmic flops 2 rows m^2 3 Nq */
int32_t ca_assembly_dense_info(const ca_assembly* h, int32_t* grid3, double* contract_seconds,
                               double* contract_flops);
int32_t ca_assembly_destroy(ca_assembly* h);
int32_t ca_assembly_info(const ca_assembly* h, int64_t* m, int64_t* row_begin, int64_t* row_end,
                         int64_t* nnz, double* device_seconds);
/* host copies: B, A1, A2 are m x m column-major, only rows [row_begin,row_end) are written */
int32_t ca_assembly_get_linear(ca_assembly* h, double* B, double* A1, double* A2);
/* ijk: nnz x 3 column-major, 1-based; val: nnz */
int32_t ca_assembly_get_coo(ca_assembly* h, int64_t* ijk, double* val);
/* D_ij = <curl Psi_i, curl Psi_j> of dissipation_function (ODEModels.jl:87-137), rows [row_begin,row_end),
 * m x m column-major host array (computed with B, A1, A2) */
int32_t ca_assembly_get_dissipation(ca_assembly* h, double* D);
/* device copies for the all-gather: linear slabs are (row_end-row_begin) x m ROW-major */
int32_t ca_assembly_get_linear_dev(ca_assembly* h, double* dB, double* dA1, double* dA2, void* stream);
int32_t ca_assembly_get_dissipation_dev(ca_assembly* h, double* dD, void* stream);
int32_t ca_assembly_get_coo_dev(ca_assembly* h, int64_t* d_i, int64_t* d_j, int64_t* d_k, double* d_val,
                                void* stream);

/* cnab2(x0, odemodel, R, dt, Nsteps, nsave) -- src/ODEModels.jl:147-181 (Crank-Nicolson / Adams-Bashforth 2)
 * for an ensemble of initial conditions, on the device:
 *     (B - dt/2 A) x_{n+1} = (B + dt/2 A) x_n + (3dt/2) N(x_n) - (dt/2) N(x_{n-1}),   A = A1 + A2/R,
 * with N(x_{-1}) := N(x_0) as in the reference (:158-159).  Per step: one pass of the sparse-bilinear kernel
 * for N(x_n) and one pass of a dense [M1 M2] operator with M1 = LHS^-1 RHS, M2 = LHS^-1 (factored on the
 * host once per (R, dt)).  dX0: nstates x m (leading dimension ld).  dXsave: Nsave = Nsteps / nsave
 * snapshots, each nstates x m contiguous (leading dimension nstates); snapshot k (0-based) is the state after
 * (k+1)*nsave steps -- the reference stores x0 in row 1 and then overwrites it with its first save because its
 * row counter starts at 1 (:164-177); that behaviour is reproduced.  dXfinal (may be NULL): state after Nsteps. */
int32_t ca_model_cnab2_batched_dev(ca_model* h, const double* dX0, int64_t nstates, int64_t ld, double R,
                                   double dt, int64_t nsteps, int64_t nsave, double* dXsave, double* dXfinal,
                                   void* stream);

/* Observables of ODEModels.jl:65-137 for ensembles: OUT[s,0] = shear(x_s) = 1 + dot(x_s, shear),
 * OUT[s,1] = dissipation(x_s) = 1 + x_s' D x_s.  ca_model_set_observables copies shear (m) and D (m x m
 * column-major, host) into the handle; OUT is nstates x 2 column-major on the device. */
int32_t ca_model_set_observables(ca_model* h, const double* shear, const double* D);
int32_t ca_model_observables_batched_dev(ca_model* h, const double* dX, int64_t nstates, int64_t ld,
                                         double* dOUT, void* stream);

/* Host-pointer forms of the three entry points above (Julia Arrays are host memory): plain copies around the
 * device calls.  X, V, OUT: nstates x m column-major; Xsave: Nsave snapshots of nstates x m; OBS: nstates x 2. */
int32_t ca_model_jvp_batched(ca_model* h, const double* X, const double* V, int64_t nstates, double R, double* OUT);
int32_t ca_model_cnab2_batched(ca_model* h, const double* X0, int64_t nstates, double R, double dt, int64_t nsteps,
                               int64_t nsave, double* Xsave, double* Xfinal);
int32_t ca_model_observables_batched(ca_model* h, const double* X, int64_t nstates, double* OBS);

/* ---- Newton-hookstep (reference: src/Hookstep.jl) --------------------------------------------- */
/* SearchParams -- Hookstep.jl:6-14, same fields, order and defaults
 * (ftol = xtol = 1e-8, delta = 0.02, Nnewton = 20, Nhook = 4, Nmusearch = 6, verbosity = 0). */
typedef struct ca_search_params {
  double ftol, xtol, delta;
  int32_t Nnewton, Nhook, Nmusearch, verbosity;
} ca_search_params;

enum { /* why the search returned (Hookstep.jl line of the return statement) */
  CA_EXIT_CONVERGED = 1,       /* norm(f(x)) <= ftol, the only exit with success = true   :146-149 */
  CA_EXIT_RESIDUAL_UP = 2,     /* dot(fx, Dfx*dx) >= 0, returns x                          :176-183 */
  CA_EXIT_XTOL_NEWTON = 3,     /* norm(dx_newton) < xtol, returns x + dx                   :185-190 */
  CA_EXIT_XTOL_HOOK = 4,       /* norm(dx_hook) < xtol, returns x + dx                     :216-221 */
  CA_EXIT_MAX_NEWTON = 5,      /* Nnewton steps done                                       :349-351 */
  CA_EXIT_SINGULAR = 6         /* a pivot was exactly zero (Julia would throw SingularException) */
};
#define CA_LOG_STEPS 64
typedef struct ca_solve_log {
  int32_t exit_reason, newton_steps, f_evals, df_evals;
  double final_residual;              /* norm(f) at the last evaluated Newton iterate */
  double final_delta;                 /* trust-region radius on exit */
  double device_seconds;              /* CUDA-event time of the solve kernel */
  double residual[CA_LOG_STEPS];      /* norm(f(x_n)) at the start of Newton step n */
  double delta[CA_LOG_STEPS];         /* trust-region radius after step n */
  int32_t hooksteps[CA_LOG_STEPS];    /* hookstep iterations taken in step n */
} ca_solve_log;

/* hookstepsolve(x -> model.f(x,R), x -> model.Df(x,R), xguess, params) -- Hookstep.jl:109-352 with
 * hookstep() :21-74 -- as ONE persistent kernel: a single CTA keeps x, f(x), Df(x), Df'Df and the LU
 * workspaces in shared memory for the whole search (no launches, no host round trips inside the
 * loop).  Same decision tree, constants and exit flags as the reference.  The one-CTA solver is limited by shared
 * memory (3 m^2 + 9 m doubles <= 227 KB, m <= 95); larger models run the same search with f, Df, Df'Df, the LU factors
 * and every vector in device global memory (blocked LU with partial pivoting, csrc/hookstep_large.cu): the host only
 * sees the scalars the decision tree branches on.
 * nsolve independent searches (own R, own guess; one CTA each, or one after the other for m > 95) run in one call:
 * R: nsolve, Xguess / X_out: nsolve x m column-major, success / logs: nsolve (logs may be NULL). */
int32_t ca_hookstep_solve_model(ca_model* h, const double* R, const double* Xguess, int64_t nsolve,
                                const ca_search_params* params, double* X_out, int32_t* success,
                                ca_solve_log* logs);

/* ---- Newton-Krylov (BASELINE.json configs[4]; SURVEY.md 8e, last row) ------------------------------ */
/* Matrix-free Newton-GMRES for f(x,R) = 0, device resident: every Krylov vector costs one Jacobian action
 * Df(x,R) v = (L1 + L2/R) v + Q(x,v) + Q(v,x) by the HBM-bound streaming kernel; restarted GMRES (classical Gram-Schmidt
 * twice), left-preconditioned with the linear operator L1 + L2/R, backtracking line search on |f|.  For models where the dense m x m Jacobian and its LU dominate hookstepsolve
 * (Hookstep.jl:158) -- m ~ 3000.
 * models: nmodels replicas of ONE model, one per local device of `comm` (NULL with nmodels == 1: a single GPU, or
 * NULL = the ca_init communicator).  With N ranks the Jacobian action is sharded by ROW SLAB: every device streams 1/N of
 * the operator and one ncclAllGather of m doubles per Krylov step completes the vector; the Arnoldi process is replicated
 * (deterministic kernels: the replicas stay bit-identical), so no vector is broadcast.  One process per GPU: every rank
 * calls this with its own model and the communicator of ca_comm_init_rank, and every rank receives the solution. */
typedef struct ca_krylov_params {
  double ftol, gmres_rtol;                 /* defaults 1e-10, 1e-8 (relative to |f| of the Newton iterate) */
  int32_t max_newton, restart, max_restarts, verbosity; /* defaults 30, 60, 5, 0 */
  int32_t precondition;                    /* default 1: GMRES on M^-1 Df with M = L1 + L2/R, factored once per solve by the
                                              device LU (without it GMRES stagnates on the viscous spectrum at m >~ 300) */
} ca_krylov_params;
typedef struct ca_krylov_info {
  int32_t converged, newton_steps, jvps, f_evals, world;
  double final_residual, device_seconds;
  double residual[CA_LOG_STEPS];           /* |f| at the start of Newton step n */
} ca_krylov_info;
int32_t ca_newton_krylov(ca_model** models, int32_t nmodels, ca_comm* comm, double R, const double* xguess,
                         const ca_krylov_params* params, double* x_out, int32_t* success, ca_krylov_info* info);

/* ---- diagnostics (host only, no device needed) ---------------------------------------------- */
/* Packs a COO operator exactly as ca_bilinear_create does, expands the packed tiles back into
 * (i,j,k,val) terms and compares them with the (merged) input: *mismatches must come back 0.
 * Also reports the packing statistics the roofline accounting uses.  Structure check only -- it
 * evaluates nothing. */
int32_t ca_selftest_pack(const int64_t* ijk, const double* val, int64_t nnz, int64_t m,
                         int32_t symmetric, int32_t window_modes /* 0 = whole-tile packing */,
                         int64_t* ntiles, int64_t* nterms, int64_t* padded_fma,
                         int64_t* pair_products, int64_t* mismatches, int64_t* nwindows,
                         int64_t* max_window_modes);

/* FNV-1a digests of the eight parts of the pack ca_bilinear_create would build (tile descriptors, pair words, values,
 * row lists, window descriptors, window mode lists, warp schedule, counters): two packs are identical iff all agree.
 * Used to show that the device model build (ca_model_create_dev) and the host packer produce the same pack. */
int32_t ca_selftest_pack_digest(const int64_t* ijk, const double* val, int64_t nnz, int64_t m, int32_t symmetric,
                                int32_t window_modes, uint64_t* digest8);

/* The blocked LU of the large-model solver on a caller-supplied dense system (column-major m x m, host): A is
 * overwritten by its factors (unit lower L below the diagonal, U on and above; rows interchanged as LAPACK's getrf
 * does), b by the solution of A x = b. */
int32_t ca_selftest_lu_solve(double* A, int64_t m, double* b, int32_t* singular);

/* Generates the model-specialised ensemble kernel (m <= 48) for a quadratic operator plus a dense linear pattern and
 * compiles it with NVRTC for sm_100a; reports the cubin size and the CTA shape.  No device needed: the module is
 * not loaded.  Fails with CA_ERR_UNSUPPORTED (and the NVRTC log as error text) if the source does not compile. */
int32_t ca_selftest_small_kernel(const int64_t* ijk, const double* val, int64_t nnz, int64_t m,
                                 int64_t* cubin_bytes, int32_t* threads, int32_t* stages, int64_t* source_bytes);

/* How many CTAs share a block of 32 states in the ensemble kernels (each taking a range of tiles) when the ensemble
 * has `nblocks` blocks and the machine `slots` CTA slots: 1 for ensembles that fill whole passes, more for small ones
 * (csrc/bilinear.cu: choose_parts).  Pure host logic, no device needed. */
int32_t ca_selftest_choose_parts(int64_t nblocks, int64_t slots, int32_t max_parts, int32_t* parts);

#ifdef __cplusplus
}
#endif
#endif /* CLOUDATLAS_B200_H */
