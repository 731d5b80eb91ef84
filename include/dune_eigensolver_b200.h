/* dune_eigensolver_b200.h -- C ABI of the B200-native block-eigensolver hot path.
 *
 * The reference (normallytangent/dune-eigensolver) has no FFI: its API is header-only C++ templates.
 * The drop-in boundary is therefore (1) the source-level signatures of the reference headers, re-created
 * in include/dune/eigensolver/{multivector,kernels_b200,eigensolver}.hh, and (2) this C ABI directly beneath
 * them. Every entry point below names the reference function or type it replaces (file:line relative to
 * the reference root). Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * Conventions
 *   - every function returns a de_status (0 = ok); de_last_error_string() gives the message
 *   - host pointers are borrowed for the duration of the call only; the library never frees caller memory
 *   - a de_context is used by one host thread at a time; several contexts may coexist
 *   - all arithmetic is IEEE fp64; there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with DE_ERR_CUDA
 *   - "panel8" = the reference MultiVector<double,8> layout (multivector.hh:130-133):
 *       element (i,j) at ((j/8)*n + i)*8 + j%8 ; the number of columns must be a multiple of 8
 */
#ifndef DUNE_EIGENSOLVER_B200_H
#define DUNE_EIGENSOLVER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct de_context de_context;
typedef struct de_matrix de_matrix;
typedef struct de_mv de_mv;
typedef struct de_factor de_factor;
typedef struct de_host_factor de_host_factor;

typedef enum de_status
{
  DE_OK = 0,
  DE_ERR_INVALID = 1,     /* shape / block-size / argument violation (reference: std::invalid_argument) */
  DE_ERR_ALLOC = 2,
  DE_ERR_CUDA = 3,
  DE_ERR_NCCL = 4,
  DE_ERR_SINGULAR = 5,    /* factorisation singular (umfpacktools.hh:160-164) or Gram matrix not positive definite */
  DE_ERR_UNSUPPORTED = 6
} de_status;

#define DE_MAX_COLS 64 /* widest column block the fused kernels handle (north_star: p = 8..64) */

int de_version(void);
/* message of the last failure on this context (ctx may be NULL: last failure of the calling thread) */
const char *de_last_error_string(const de_context *ctx);

/* ---- context ------------------------------------------------------------------------------------
 * device: CUDA ordinal. stream: a cudaStream_t to run on (e.g. torch's current stream), or NULL to let
 * the context create its own. */
int de_context_create(int device, void *stream, de_context **out);
int de_context_destroy(de_context *ctx);
int de_context_synchronize(de_context *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int de_context_launch_count(const de_context *ctx, int64_t *count);
/* Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline numbers).
 * de_context_profile synchronises, folds the pending events into per-category totals and returns one category;
 * reset != 0 clears all categories afterwards. */
#define DE_PROF_SPMM 0    /* spmm_kernel (with or without the fused diag-dot) */
#define DE_PROF_GRAM 1    /* gram_kernel */
#define DE_PROF_UPDATE 2  /* update_kernel */
#define DE_PROF_SMALL 3   /* reduce_partials_kernel, chol_inverse_kernel */
#define DE_PROF_DOT 4     /* diag_dot_kernel */
#define DE_PROF_TRSV 5    /* permute / level / chain kernels of the factored apply */
#define DE_PROF_MISC 6    /* layout conversion, halo pack, eigenvector extraction */
#define DE_PROF_SPMM_BOUNDARY 7 /* boundary-row launches of a distributed SpMM (they wait for the halo rows) */
#define DE_PROF_HALO_PUSH 8 /* halo_push_kernel: halo rows stored into the neighbours' windows (runs in front of the interior rows) */
#define DE_PROF_HALO_WAIT 9 /* halo_wait_kernel: what the boundary rows wait for; ~0 when the interior rows hide the exchange */
#define DE_PROF_CATEGORIES 10
/* enable: 0 = off, 1 = every category, otherwise a bit mask of categories shifted left by one (2 << DE_PROF_SPMM | ...) */
int de_context_set_profiling(de_context *ctx, int enable);
/* Tuning / A-B switches of a context. The library itself reads NO environment variable for behaviour; the
 * Python mirror maps DE_B200_<NAME> onto this call for bench.py and the tests. Every option defaults to the faster, measured
 * setting; all settings give the same results to rounding. name (value):
 *   "one_sweep"      (1)  CholQR: one sweep when the first Gram matrix says it is enough, 0 = always the second sweep's test
 *   "cheb_epilogue"  (1)  LOBPCG: Chebyshev update as an epilogue of the tensor-core SpMM, 0 = SpMM + streaming kernel
 *   "lincomb2"       (1)  LOBPCG combination / projection on the tensor-core kernel, 0 = first-generation FMA kernels
 *   "loop_graph"     (1)  StandardLargest: steady-state iterations replayed from a CUDA graph (one GPU), 0 = plain launches
 *   "fused_push"     (0)  halo rows stored by the block-update kernels instead of halo_push_kernel (measured slower)
 *   "brb_plane_points" (16384) PROCESS-WIDE: grid planes with more points are swept in y chunks by the BRB tile order; takes
 *                          effect for matrices created afterwards
 * Unknown names return DE_ERR_INVALID. */
int de_context_set_option(de_context *ctx, const char *name, int64_t value);
int de_context_profile(de_context *ctx, int category, double *total_ms, int64_t *launches, int reset);

/* Multi-GPU (new; the reference is single-threaded): one process per GPU. Rank 0 obtains an
 * id with de_comm_unique_id (128 bytes), the host distributes it (torch.distributed / MPI), every rank calls
 * de_context_init_comm. Reductions of diag-dot / Gram results are then all-reduced over NCCL and distributed
 * matrices exchange their halo rows over NVLink. */
int de_comm_unique_id(void *id128);
int de_context_init_comm(de_context *ctx, int rank, int nranks, const void *id128);
int de_context_rank(const de_context *ctx, int *rank, int *nranks);

/* Optional NVLink fast path of the multi-GPU data path (csrc/kernels_peer.cuh): every rank creates a window in its
 * HBM (flags, all-reduce slots, two halo buffers of halo_bytes each) and receives a 64-byte CUDA IPC handle; the
 * handles of all ranks, in rank order (nranks x 64 bytes), are passed to ..._open, which maps the peers' windows.
 * The caller exchanges the handles (torch.distributed in parallel.py) and runs a barrier after ..._open. From then on
 * the short all-reduces are one-shot peer-memory kernels, and matrices that were told where their rows go in the
 * neighbours' halo blocks (de_matrix_set_peer_deposit) send halo rows as peer stores instead of ncclSend/ncclRecv.
 * Without a window, or if the IPC mapping fails (DE_ERR_UNSUPPORTED), everything runs over NCCL. */
int de_context_peer_window_create(de_context *ctx, int64_t halo_bytes, void *ipc_handle64);
int de_context_peer_window_open(de_context *ctx, const void *ipc_handles);
int de_context_peer_ready(const de_context *ctx, int *ready);

/* ---- sparse matrix: replaces BCRSMatrix<FieldMatrix<double,1,1>> traversal ------------------------
 * (eigensolver.hh:32-66,208-252; kernels_cpp.hh:383-392,644-653). Header code flattens BCRS -> CSR once
 * per solve, AFTER applying shift / axpy / regularisation on the host exactly as the reference does. */
int de_matrix_create_csr(de_context *ctx, int64_t n, int64_t nnz, const int64_t *rowptr, const int64_t *col,
                         const double *val, de_matrix **out);
/* BCSR: a BCRSMatrix<FieldMatrix<double,k,k>> with k > 1 -- nb block rows, nnzb stored blocks, val = nnzb * k * k
 * doubles (block after block, each row-major). The reference's drivers accept such matrices by type
 * (eigensolver.hh:36-39) but every one of its kernels throws for k != 1 (kernels_cpp.hh:362-363, :632-633); here the
 * matrix is the scalar (nb*k) x (nb*k) matrix it denotes: row ib*k + r holds val[e][r][c] at column col[e]*k + c for
 * every block e of block row ib. Vector blocks have nb*k rows (the entries of a BlockVector<FieldVector<double,k>> in
 * storage order). On the device the dense k x k blocks land in the 8 x 4 tensor-core steps of the BRB form like any
 * other entries (a block row of k = 2 or 4 fills them better than a stencil row does). k = 1 is de_matrix_create_csr. */
int de_matrix_create_bcsr(de_context *ctx, int64_t nb, int64_t nnzb, int k, const int64_t *rowptr, const int64_t *col,
                          const double *val, de_matrix **out);
/* Row-partitioned matrix: this rank owns `n_owned` consecutive rows. Columns are already renumbered to
 * [0,n_owned) = owned rows of the vector block, [n_owned, n_owned+n_halo) = halo rows received from peers.
 * Peers are listed in ascending rank order; recv_counts[p] halo rows arrive from peer p (they occupy the
 * halo range in peer order), and the owned rows send_rows[send_offsets[p] .. send_offsets[p+1]) are sent
 * to peer p. Build these with de_halo_plan_* below. */
int de_matrix_create_distributed(de_context *ctx, int64_t n_owned, int64_t n_halo, int64_t nnz,
                                 const int64_t *rowptr, const int64_t *col_local, const double *val, int npeers,
                                 const int *peer_ranks, const int64_t *recv_counts, const int64_t *send_offsets,
                                 const int64_t *send_rows, de_matrix **out);
/* deposit_rows[p]: the row of peer p's halo block (its [owned | halo] numbering minus n_owned) at which the rows this
 * rank sends to peer p start; max_halo_rows_all_ranks: the largest halo block of this matrix over all ranks (whether the
 * halo buffers fit the window must be the same decision on every rank). Enables peer-store halo exchange for this
 * matrix (needs the context's peer window); every rank of the job must call it for its part of the matrix. */
int de_matrix_set_peer_deposit(de_matrix *A, const int64_t *deposit_rows, int64_t max_halo_rows_all_ranks);
int de_matrix_destroy(de_matrix *A);
int de_matrix_rows(const de_matrix *A, int64_t *n_owned, int64_t *nnz);

/* SpMM kernel family used for this matrix (kernels_cpp.hh:626-657 on the device). At creation every matrix gets
 * the CSR form; one whose 8-row blocks fit the shared-memory tile budget also gets the BRB form (8-row blocks x
 * 4-column steps on the FP64 tensor path, X rows staged per tile; csrc/brb_format.hpp). DE_SPMM_AUTO prefers BRB.
 * Both families compute the same products; sums differ in rounding only (different association order). */
enum
{
  DE_SPMM_AUTO = 0,
  DE_SPMM_CSR = 1,
  DE_SPMM_BRB = 2
};
int de_matrix_set_spmm_format(de_matrix *A, int format);
/* format: the family SpMM calls use now (DE_SPMM_CSR / DE_SPMM_BRB); sizes of the BRB form (0 if absent);
 * tile_shape3: grid points per tile if a structured-grid pattern was detected, else zeros. Any pointer may be null. */
int de_matrix_spmm_info(const de_matrix *A, int *format, int64_t *tiles, int64_t *row_blocks, int64_t *steps,
                        int64_t *union_rows_max, int *tile_shape3);
/* The BRB form is built ON THE DEVICE from the uploaded CSR arrays (csrc/kernels_brb_build.cuh). This check rebuilds it
 * with the host builder (csrc/brb_format.hpp) from the caller's CSR arrays and counts the 32-bit words in which the two
 * differ: 0 for matrices without duplicate entries. *mismatches = -1 if the matrix has no BRB form. */
int de_matrix_brb_selfcheck(const de_matrix *A, int64_t n, int64_t ncols, const int64_t *rowptr, const int64_t *col,
                            const double *val, int64_t *mismatches);
/* Host-only self-check of the BRB construction (no GPU needed; not a compute path): builds the BRB form of a CSR
 * matrix with columns [0, ncols) of which [0, n_owned) are owned, decodes every tile the way the kernel does, and
 * returns max_i |(A_brb p)_i - (A_csr p)_i| for a fixed probe vector p. info8 = {has BRB form, structured grid
 * detected, tiles, interior tiles, row blocks, steps, max union rows, tile shape tw | th << 16 | td << 32}.
 * nthreads <= 0: as many builder threads as the library would use. */
int de_brb_format_check(int64_t n, int64_t ncols, int64_t n_owned, const int64_t *rowptr, const int64_t *col,
                        const double *val, int nthreads, int64_t *info8, double *max_abs_diff);

/* Host-only halo planning for a 1-D row partition (no GPU needed; exercised by the gloo CPU tests).
 * Input: this rank's rows [row_begin,row_end) of the global CSR with GLOBAL column indices and the partition
 * offsets part[0..nranks]. Output (caller-allocated): col_local[nnz]; halo_global[] (capacity nnz) = global
 * index of every halo row in halo order (sorted by owner then index); recv_counts[nranks] per owner. */
int de_halo_plan_local(int64_t n_owned, const int64_t *rowptr, const int64_t *col_global, int nranks, int rank,
                       const int64_t *part, int64_t *col_local, int64_t *halo_global, int64_t *n_halo,
                       int64_t *recv_counts);

/* Host-only second half of the planning (no GPU needed; exercised by the gloo CPU tests): from the all-gathered
 * results of de_halo_plan_local -- counts_all[q*nranks + p] = halo rows rank q receives from owner p, lists_all + q *
 * list_stride = rank q's halo_global list -- derive this rank's neighbours. Outputs (caller-allocated, nranks entries;
 * send_offsets nranks + 1; send_rows sum_q counts_all[q*nranks + rank]): peers in ascending rank order, rows received
 * from / sent to each (send_rows are LOCAL row indices in the receiver's halo order), deposit_rows[p] = first row of
 * this rank's rows inside peer p's halo block, the largest halo block over all ranks, and whether the send / receive
 * relation is symmetric between every pair of ranks (the peer-store exchange requires it). */
int de_halo_plan_peers(int nranks, int rank, const int64_t *part, const int64_t *counts_all, const int64_t *lists_all,
                       int64_t list_stride, int *npeers, int *peer_ranks, int64_t *recv_counts, int64_t *send_offsets,
                       int64_t *send_rows, int64_t *deposit_rows, int64_t *max_halo_rows, int *symmetric);

/* One rank's part of a row-partitioned matrix, straight from its rows with GLOBAL column indices: halo planning
 * (de_halo_plan_local), the exchange of the halo lists between the ranks and the peer-deposit offsets happen inside.
 * part[0..nranks] are the partition offsets (this rank owns rows [part[rank], part[rank+1]) = n_owned rows; rank and
 * nranks are the context's, see de_context_init_comm / de_multi_create). The only thing the caller supplies is an
 * all-gather over the ranks of the job: every rank passes `bytes` bytes in `send`, `recv` receives nranks*bytes in
 * rank order; return 0 on success (torch.distributed / MPI_Allgather in a multi-process job; de_multi_* brings its
 * own). With one rank this is de_matrix_create_csr and allgather may be NULL. */
typedef int (*de_allgather_fn)(void *user, const void *send, void *recv, int64_t bytes);
int de_matrix_create_rowblock(de_context *ctx, int64_t n_owned, int64_t nnz, const int64_t *rowptr,
                              const int64_t *col_global, const double *val, const int64_t *part,
                              de_allgather_fn allgather, void *user, de_matrix **out);

/* ---- multivector: replaces MultiVector<double,8> (multivector.hh:17-146) ---------------------------
 * Device-resident n x m block (row-major inside the library; the layout is opaque). m % 8 == 0 is enforced
 * like multivector.hh:48-49; storage is zero-initialised like multivector.hh:52. For a distributed run n is
 * the number of OWNED rows. */
int de_mv_create(de_context *ctx, int64_t n, int m, de_mv **out);
int de_mv_destroy(de_mv *X);
int de_mv_shape(const de_mv *X, int64_t *n, int *m);
int de_mv_upload_panel8(de_mv *X, const double *host_panel8);
int de_mv_download_panel8(const de_mv *X, double *host_panel8);
int de_mv_upload_rowmajor(de_mv *X, const double *host_rowmajor);
int de_mv_download_rowmajor(const de_mv *X, double *host_rowmajor);
int de_mv_copy(de_mv *dst, const de_mv *src);
/* raw device pointer (row-major, leading dimension m) for zero-copy interop. The pointer is INVALIDATED by
 * de_standard_largest_mv / de_standard_inverse_mv on the same block: those drivers alternate between the block's buffer
 * and a work buffer and leave the result in whichever holds it (no copy of n x m doubles); ask again afterwards. */
int de_mv_device_ptr(de_mv *X, void **dptr);

/* ---- kernels --------------------------------------------------------------------------------------*/
/* Y = A X.  replaces matmul_sparse_tallskinny_{naive,blocked,avx2_b8,neon_b8}
 * (kernels_cpp.hh:596-657, kernels_avx2.hh:1021-1059, kernels_neon.hh:1314-1361). All m columns in one pass
 * over A. */
int de_spmm(de_mv *Y, const de_matrix *A, const de_mv *X);
/* Y = A X and dp[j] = sum_i X(i,j) Y(i,j) in the same pass (the SpMM + dot_products_diagonal_blocked pair of
 * eigensolver.hh:84-85, :174-175, :308-309). dp_host has m entries. */
int de_spmm_diag_dot(de_mv *Y, const de_matrix *A, const de_mv *X, double *dp_host);
/* dp[j] = sum_i X(i,j) Y(i,j).  replaces dot_products_diagonal_{blocked,avx2_b8,neon_b8} (kernels_cpp.hh:24-55) */
/* Y = A X, dp = diag(X^T Y) and G = Y^T Y (m x m, row-major) from ONE pass: the Gram matrix the next
 * orthonormalisation of Y needs (kernels_cpp.hh:236-242 forms it panel by panel) is accumulated in the SpMM epilogue
 * when the matrix has the BRB form and m is 8, 16 or 32; otherwise a separate Gram pass produces the same result. */
int de_spmm_gram(de_mv *Y, const de_matrix *A, const de_mv *X, double *dp_host, double *G_host);
int de_diag_dot(double *dp_host, const de_mv *X, const de_mv *Y);
/* G = X^T Y, row-major m x m on the host. replaces dot_products_all_blocked (kernels_cpp.hh:58-96) and the naive
 * dot_products_diagonal(Q) full Gram (kernels_cpp.hh:7-21) */
int de_gram(double *G_host, const de_mv *X, const de_mv *Y);
/* X <- X Q with Q a row-major m x m host matrix: the block update of kernels_cpp.hh:293-305, :514-539 (Q upper
 * triangular there) written for a general Q */
int de_block_update(de_mv *X, const double *Q_host);
/* X[:, j0:j0+w) -= X[:, k0:k0+w) S with S a row-major w x w host matrix: the projection of
 * kernels_cpp.hh:335-348, :570-583 (w = 8 there) */
int de_block_project(de_mv *X, int j0, int k0, int w, const double *S_host);
/* In-place thin QR, X <- X R^-1 with R upper triangular, positive diagonal (so the column order of Gram-Schmidt
 * is preserved). replaces orthonormalize_{naive,blocked,avx2_b8,avx2_b8_v2,neon_b8,neon_b8_v2}
 * (kernels_cpp.hh:121-155, :180-351). Algorithm: CholQR2 over the whole block. */
int de_orthonormalize(de_mv *X);
/* Same in the B inner product, X^T B X = I. replaces B_orthonormalize_{blocked,avx2_b8,neon_b8}
 * (kernels_cpp.hh:356-591). If BX is not NULL it receives B*X for the orthonormalised X (the reference keeps
 * that product in its scratch panel `p`, kernels_cpp.hh:372,527-539). *norm (may be NULL) receives the largest
 * strict-upper entry of the first Gram matrix X^T B X (diagnostic; unused by the reference's callers). */
int de_b_orthonormalize(const de_matrix *B, de_mv *X, de_mv *BX, double *norm);

/* ---- factored inverse apply -----------------------------------------------------------------------
 * replaces UMFPackFactorizedMatrix's field contract (umfpacktools.hh:26-44) and
 * matmul_inverse_tallskinny_{blocked,avx2_b8,neon_b8} (kernels_cpp.hh:660-755). L: CSR, unit diagonal stored
 * last in each row; U: CSC, diagonal last in each column; P, Q, Rs, do_recip as UMFPACK defines them. The upload
 * builds the level schedules once; single-GPU only (triangular solves do not row-shard). */
int de_factor_upload(de_context *ctx, int64_t n, const long *Lp, const long *Lj, const double *Lx, const long *Up,
                     const long *Ui, const double *Ux, const long *P, const long *Q, const double *Rs, long do_recip,
                     de_factor **out);
int de_factor_destroy(de_factor *F);
/* Y = A^-1 X ; X is used as scratch and is clobbered, as in the reference (kernels_cpp.hh:659) */
int de_factor_apply(de_mv *Y, const de_factor *F, de_mv *X);
int de_factor_info(const de_factor *F, int64_t *n, int64_t *lnz, int64_t *unz, int *levels_L, int *levels_U);

/* ---- whole-driver entry points: the iteration loop never leaves the device ------------------------
 * start_panel8: the n x m start block (m = nev rounded up to a multiple of 8) filled by the caller exactly as
 * eigensolver.hh:50-55 does (de_start_block reproduces that stream). eval: nev values; evec: nev vectors of
 * length n, vector j at evec + j*n (both caller-allocated, as eigensolver.hh:105-111 assumes). *iterations
 * receives the reference's loop counter at exit. The shift must already be part of A (the header applies it
 * on the host like eigensolver.hh:57-66); it is passed only to be subtracted from the Rayleigh quotients. */
/* replaces StandardLargest (eigensolver.hh:28-112) */
int de_standard_largest(de_context *ctx, const de_matrix *A, double shift, double tol, int maxiter, int nev,
                        const double *start_panel8, double *eval, double *evec, int verbose, int *iterations);
/* device-resident variants of the two drivers above: Q holds the n x m start block on entry and the eigenvector
 * block on return (all m columns), eval_m receives m Rayleigh quotients; nothing but the m convergence values
 * per iteration crosses PCIe. Used when the caller keeps its blocks on the GPU (and by bench.py's `value`). */
int de_standard_largest_mv(de_context *ctx, const de_matrix *A, double shift, double tol, int maxiter, de_mv *Q,
                           double *eval_m, int verbose, int *iterations);
int de_standard_inverse_mv(de_context *ctx, const de_matrix *A, const de_factor *F, double shift, double tol,
                           int maxiter, de_mv *Q, double *eval_m, int verbose, int *iterations);
/* replaces StandardInverse (eigensolver.hh:116-198); F = factorisation of the shifted A */
int de_standard_inverse(de_context *ctx, const de_matrix *A, const de_factor *F, double shift, double tol,
                        int maxiter, int nev, const double *start_panel8, double *eval, double *evec, int verbose,
                        int *iterations);
/* replaces GeneralizedInverse (eigensolver.hh:204-351); A = inA + shift*B + reg*I (built on the host like
 * eigensolver.hh:241-252), F its factorisation */
int de_generalized_inverse(de_context *ctx, const de_matrix *A, const de_matrix *B, const de_factor *F, double shift,
                           double tol, int maxiter, int nev, const double *start_panel8, double *eval, double *evec,
                           int verbose, int *iterations, double *relerror);

/* ---- single-process multi-GPU front end -----------------------------------------------------------------------
 * SURVEY.md §8b's `de_context_create(const int *device_ids, int ndev, ...)`: ONE process drives ndev GPUs (1..8), one
 * host thread and one de_context per GPU inside the library; the windows of the NVLink data path (halo rows as peer
 * stores, one-shot peer all-reduce of the m and m x m reductions) are peer-mapped allocations of this process, so no
 * NCCL, MPI or Python is involved. halo_bytes: capacity of each of the two halo buffers per GPU (0: 128 MB; a block of
 * halo rows of the widest vector block must fit). The same ordinal may be listed more than once (several ranks on one
 * GPU: used by the tests on single-GPU machines). The drivers take the GLOBAL matrix (host CSR, as
 * de_matrix_create_csr) and the GLOBAL start block and return global eigenvectors, exactly like their single-GPU
 * counterparts de_standard_largest / de_standard_lobpcg / de_generalized_lobpcg; rows are split into ndev contiguous
 * blocks, cut only at multiples of row_align (e.g. one grid plane; <= 1: anywhere). Results agree with the one-GPU
 * run to rounding (different reduction order); every rank takes the same convergence decision. */
typedef struct de_multi de_multi;
int de_multi_create(const int *device_ids, int ndev, int64_t halo_bytes, de_multi **out);
int de_multi_destroy(de_multi *M);
int de_multi_size(const de_multi *M, int *ndev);
/* a rank that waits longer than this for a peer's halo rows / all-reduce contribution gives up; the call then fails with
 * DE_ERR_NCCL instead of hanging the GPU (default ~30 s) */
int de_multi_set_timeout(de_multi *M, double seconds);
/* rank r's context (borrowed; owned by M): for profiling / launch counts or rank-level calls from r's own thread */
int de_multi_context(de_multi *M, int rank, de_context **ctx);
const char *de_multi_last_error(const de_multi *M);
int de_multi_launch_count(const de_multi *M, int64_t *count);
int de_multi_standard_largest(de_multi *M, int64_t n, int64_t nnz, const int64_t *rowptr, const int64_t *col,
                              const double *val, int64_t row_align, double shift, double tol, int maxiter, int nev,
                              const double *start_panel8, double *eval, double *evec, int verbose, int *iterations);
int de_multi_standard_lobpcg(de_multi *M, int64_t n, int64_t nnz, const int64_t *rowptr, const int64_t *col,
                             const double *val, int64_t row_align, double tol, int maxiter, int nev,
                             const double *start_panel8, double *eval, double *evec, int verbose, int *iterations);
int de_multi_generalized_lobpcg(de_multi *M, int64_t n, int64_t nnz, const int64_t *rowptr, const int64_t *col,
                                const double *val, int64_t b_nnz, const int64_t *b_rowptr, const int64_t *b_col,
                                const double *b_val, int64_t row_align, double tol, int maxiter, int nev,
                                const double *start_panel8, double *eval, double *evec, int verbose, int *iterations);

/* ---- LOBPCG drivers -------------------------------------------------------------------------------------
 * NEW: the reference has no LOBPCG (its drivers are eigensolver.hh:28-112, :116-198, :204-351); BASELINE.json names
 * StandardLOBPCG for the 3D configurations, SURVEY.md §8f ranks it first among the adjacent components. Parameter
 * shape of the reference's free functions: start block filled like eigensolver.hh:50-55 (de_start_block), m = nev
 * rounded up to a multiple of 8 (eigensolver.hh:43), eval / evec caller-allocated (nev values, nev vectors of length
 * n). Computes the nev SMALLEST eigenpairs of A x = lambda x (A x = lambda B x, B symmetric positive definite) without
 * any factorisation. Convergence: ||A x_j - theta_j B x_j||_2 <= tol * |theta_j| for every j < nev, with x_j
 * B-normalised; like the reference's drivers the loop falls through silently at maxiter. *iterations = number of
 * basis updates. The Rayleigh-Ritz problem (3m x 3m) is solved on the host (csrc/host_eig.hpp); every n x m operation
 * is a device kernel (csrc/lobpcg_core.hpp lists them). Iteration counts are parity-unpinned (no reference
 * implementation); converged eigenpairs are checked against analytic spectra / the reference's drivers. */
int de_standard_lobpcg(de_context *ctx, const de_matrix *A, double tol, int maxiter, int nev,
                       const double *start_panel8, double *eval, double *evec, int verbose, int *iterations);
int de_generalized_lobpcg(de_context *ctx, const de_matrix *A, const de_matrix *B, double tol, int maxiter, int nev,
                          const double *start_panel8, double *eval, double *evec, int verbose, int *iterations);
/* Preconditioner of the two drivers above: a Jacobi-scaled Chebyshev polynomial in A of this degree (that many extra
 * SpMMs per iteration, no factorisation, works row-partitioned): ~ A^-1 on the upper part of the spectrum of
 * diag(A)^-1 A, bounded by its Gershgorin row sums. A matrix without a positive diagonal is iterated without it. Measured: DESIGN.md §10. */
#define DE_LOBPCG_DEFAULT_CHEB_DEGREE 8
/* device-resident variant with all options: B may be NULL (standard problem); T, if not NULL, is a factorisation used
 * as preconditioner W <- T^-1 W (e.g. of A + shift*B; single GPU only) and takes precedence over cheb_degree;
 * cheb_degree: 0 = unpreconditioned, k > 0 = Chebyshev polynomial preconditioner with k applications of A (smallest
 * eigenvalues only); largest != 0 selects the largest eigenvalues.
 * Q: start block on entry, all m Ritz vectors on return; eval_m / resnorm_m (may be NULL): m Ritz values and residual
 * norms; restarts / converged may be NULL. Works on a row-partitioned matrix (n = owned rows; reductions all-reduced). */
int de_lobpcg_mv(de_context *ctx, const de_matrix *A, const de_matrix *B, const de_factor *T, int largest,
                 int cheb_degree, double tol, int maxiter, int nev, de_mv *Q, double *eval_m, double *resnorm_m, int verbose, int *iterations,
                 int *restarts, int *converged);
/* The fused block combination of an LOBPCG iteration as a kernel of its own (csrc/kernels_lobpcg.cuh):
 * out = sum_{s < ns} S[s] C_s and, if out2 is not NULL and ns > 1, out2 = sum_{1 <= s < ns} S[s] C_s, with ns <= 3
 * blocks of one shape and C_host = ns row-major m x m matrices one after the other. Every source is read once.
 * out may alias S[0]; out2 may alias S[1] or S[2]. Generalises the reference's V <- V U and Q_j -= Q_k S
 * (kernels_cpp.hh:293-305, :335-348) to several sources and two results. */
int de_block_lincomb(de_mv *out, de_mv *out2, int ns, const de_mv *const *S, const double *C_host);
/* Host-only (no GPU): the dense symmetric eigensolvers of the Rayleigh-Ritz step. A = V diag(w) V^T (w ascending,
 * eigenvector j in column j of the row-major V); GA c = w GB c with C^T GB C = I, *min_pivot (may be NULL) = smallest
 * Cholesky pivot of the unit-diagonal-scaled GB. DE_ERR_SINGULAR if GB is not positive definite. */
int de_host_sym_eig(int n, const double *A, double *w, double *V);
int de_host_sym_gen_eig(int n, const double *GA, const double *GB, double *w, double *C, double *min_pivot);


/* ---- host-side helpers (no GPU) -------------------------------------------------------------------*/
/* the reference's start block: std::mt19937{seed} + std::normal_distribution<double>{0,1}, filled
 * panel -> row -> column-in-panel (eigensolver.hh:50-55, :138-143, :232-237) */
int de_start_block(int64_t n, int m, unsigned seed, double *out_panel8);
/* one-time host factorisation filling the UMFPACK field contract (stand-in for umfpacktools.hh:46-199 where
 * UMFPACK is unavailable): ordering 0 = natural, 1 = nested dissection (geometric on
 * structured grids with diagonal neighbours and at least 50 000 rows, else METIS), 2 = RCM, 3 = always the graph partitioner
 * (METIS), 4 = geometric on any detected structured grid */
int de_host_factorize(int64_t n, const int64_t *rowptr, const int64_t *col, const double *val, int ordering,
                      int scale_rows, de_host_factor **out);
int de_host_factor_arrays(const de_host_factor *F, int64_t *n, int64_t *lnz, int64_t *unz, const long **Lp,
                          const long **Lj, const double **Lx, const long **Up, const long **Ui, const double **Ux,
                          const long **P, const long **Q, const double **Rs, long *do_recip);
int de_host_factor_destroy(de_host_factor *F);
/* Second host provider, for LARGE symmetric positive definite matrices (3D pencils, BASELINE.json configs[2]): supernodal
 * multifrontal Cholesky P A P^T = L L^T with nested dissection (include/dune/eigensolver/supernodal_cholesky.hh),
 * multi-threaded (nthreads <= 0: all cores). The factor stays in supernodal form -- dense column blocks, 8 bytes per
 * entry -- which is also what the device apply uses (dense panels instead of scalar level schedules, csrc/kernels_snode.cuh).
 * DE_ERR_SINGULAR if the matrix is not positive definite (use de_host_factorize). de_host_factor_arrays works on such a
 * factor too: the UMFPACK field contract (L unit lower by rows, U = D L^T by columns, P = Q, Rs = 1) is expanded on
 * first use -- meant for factors of moderate size (parity tests against the reference's own apply). */
int de_host_factorize_spd(int64_t n, const int64_t *rowptr, const int64_t *col, const double *val, int ordering, int nthreads,
                          de_host_factor **out);
/* supernodal: 1 for a de_host_factorize_spd factor; lnz: entries of L; stored: doubles held (explicit zeros of the
 * supernodal blocks included); flops of the numeric factorisation; seconds3 = {ordering, symbolic, numeric}. Any may be NULL. */
int de_host_factor_info(const de_host_factor *F, int *supernodal, int64_t *n, int64_t *lnz, int64_t *stored, double *flops,
                        double *seconds3);
/* device copy of a host factorisation of either kind, ready for de_factor_apply and the drivers */
int de_factor_upload_host(de_context *ctx, const de_host_factor *F, de_factor **out);

#ifdef __cplusplus
}
#endif

#endif /* DUNE_EIGENSOLVER_B200_H */
