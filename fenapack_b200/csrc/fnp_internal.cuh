// Internal declarations of libfenapack_cuda (sm_100a).  Not part of the ABI.
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fenapack_cuda.h"

namespace fnp {

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &m);

#define FNP_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      throw ::fnp::Error(FNP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + \
                                           " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

// NCCL entry points, bound with dlopen on first use (nccl_shim.cu)
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
const NcclApi &nccl();

#define FNP_NCCL(call)                                                                               \
  do {                                                                                               \
    ncclResult_t r_ = (call);                                                                        \
    if (r_ != ncclSuccess)                                                                           \
      throw ::fnp::Error(FNP_ERR_NCCL, std::string(#call) + ": " + ::fnp::nccl().GetErrorString(r_)); \
  } while (0)

#define FNP_REQUIRE(cond, code, msg)                 \
  do {                                               \
    if (!(cond)) throw ::fnp::Error((code), (msg));  \
  } while (0)

// ---------------------------------------------------------------------------
// device memory
// ---------------------------------------------------------------------------
template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) FNP_CUDA(cudaMalloc(reinterpret_cast<void **>(&p), count * sizeof(T)));
  }
  void ensure(size_t count) { if (count > n) alloc(count); }
  void upload(const T *h, size_t count, cudaStream_t s) {
    ensure(count);
    if (count) FNP_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void zero(cudaStream_t s) { if (n) FNP_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

// ---------------------------------------------------------------------------
// host CSR (set-up side) and device CSR (apply side)
// ---------------------------------------------------------------------------
struct HostCsr {
  int64_t nrows = 0, ncols = 0;
  std::vector<int32_t> rowptr, col;
  std::vector<double> val;
  int64_t nnz() const { return rowptr.empty() ? 0 : rowptr.back(); }
};

struct Ctx;

// Halo plan of one distributed operator (multi-rank contexts): which owned entries of
// x each peer needs (send_idx, grouped by peer) and where the entries received from
// each peer land in the ghost buffer (ghost columns are sorted by global id, hence
// grouped by owner).
// what a ghost-consuming kernel needs to wait for a peer-memory exchange (lives in device memory)
struct HaloWaitDev {
  const double *arena;                  // [2 ghost slots | flags]
  const unsigned long long *seq;        // exchanges posted so far on this plan
  const unsigned long long *flags;      // one per rank: last sequence number that rank delivered
  const int *recv_cnt;                  // per rank: entries expected (0: not a neighbour)
  int *err;                             // raised by a timed-out wait
  int nranks;
  int nghost;
};

struct HaloPlan {
  std::vector<int> send_count, send_off, recv_count, recv_off;   // per peer rank
  int32_t nsend = 0, nghost = 0;
  DevBuf<int32_t> send_idx;
  DevBuf<double> send_buf, ghost;
  // peer-memory path (one node, NVLink): the pack kernel stores straight into the ghost
  // buffers of the neighbouring GPUs (cudaIpc mappings) and raises a flag there; the
  // receiver spins on its local flags.  Two ghost slots alternate by sequence number.
  bool p2p = false;
  DevBuf<double> arena;                  // [2 * nghost doubles | nranks flags]
  std::vector<void *> peer_base;         // opened IPC mappings (to close)
  DevBuf<double *> d_peer_dst;           // per peer: my segment inside its slot 0
  DevBuf<unsigned long long *> d_peer_flag;
  DevBuf<long long> d_peer_stride;       // per peer: its nghost (slot stride)
  DevBuf<int> d_send_off, d_send_cnt, d_recv_cnt;
  DevBuf<unsigned int> d_counter;
  DevBuf<unsigned long long> d_seq;      // device-resident sequence number: the exchange is CUDA-graph replayable
  DevBuf<HaloWaitDev> d_wait;
  const double *current_ghost = nullptr; // NCCL path: what the SpMV kernels read after the last exchange
  ~HaloPlan();
};
// posts the exchange of the ghost entries of x on `stream`.  NCCL path: complete in stream order.
// Peer-memory path: the consumer kernel waits for the flags (HaloWaitDev, kernels.cu); halo_wait is
// the stand-alone wait for host-side consumers (set-up)
void halo_exchange(Ctx &c, HaloPlan &h, const double *x_own, cudaStream_t stream, ncclComm_t comm);
void halo_wait(Ctx &c, HaloPlan &h, cudaStream_t stream);
const double *halo_ghost_after_wait(Ctx &c, HaloPlan &h);      // host-side: pointer to the current ghost slot
void halo_enable_p2p(Ctx &c, HaloPlan &h);

// Device-resident CSR operator; columns index [x_own | x_ghost].
struct DevCsr {
  int bs = 1;                // Kronecker block size: the operator is S (x) I_bs, S stored (kernels.cu)
  int64_t vec_rows() const { return (int64_t)nrows * bs; }   // length of y
  int64_t vec_cols() const { return (int64_t)ncols_own * bs; }   // owned length of x
  int32_t nrows = 0;
  int32_t ncols_own = 0;     // columns [0, ncols_own) address the owned part of x
  int32_t nghost = 0;        // columns [ncols_own, ncols_own+nghost) address the ghost buffer
  int64_t nnz = 0;
  int lanes = 8;             // lanes per row of the vector kernel, from the row-length histogram
  int sell_warps = 1;        // warps per SELL slice (multi-warp kernel for operators with few rows)
  std::shared_ptr<HaloPlan> halo;   // null on single-rank contexts
  std::string tag;           // name used by the per-kernel timers ("A00", "A00/L1", "A00/P0", ...)
  DevBuf<int32_t> rowptr, col;
  DevBuf<double> val, dinv;
  // SELL-32-sigma copy (short-row operators): slice s holds rows sl_perm[32 s .. 32 s + 31],
  // entries stored column-major inside the slice, padded to the slice's longest row
  bool sell = false;
  int32_t nslices = 0, nslices_b = 0;       // interior part / boundary part (rows with ghost columns)
  int64_t sell_entries = 0, sell_entries_a = 0;   // padded entry count: total / interior part
  DevBuf<int32_t> sl_ptr, sl_col, sl_perm, sl_ptr_b, sl_perm_b;
  DevBuf<double> sl_val;
  std::vector<int32_t> sell_pos;   // host: CSR entry k -> position in sl_val (for value refreshes)
  bool has_dinv = false;
  double mean_row = 0.0, max_row = 0.0;
  // algorithmic bytes of one y = A x  (SURVEY 8d): 12 nnz + 4 (rows+1) + 8 rows + 8 cols
  double spmv_bytes() const {
    return 12.0 * nnz + 4.0 * (nrows + 1) + 8.0 * bs * nrows + 8.0 * bs * (ncols_own + nghost);
  }
};

// epilogues of the SpMV-class kernels ---------------------------------------
// Optional second output of the store / axpby epilogues: y2 = s2 * d2 .* y, the first step of the
// Chebyshev-Jacobi smoother that consumes y as its right-hand side (amg.cu) -- the restriction and
// the residual of a V-cycle level hand the smoother its first iterate instead of leaving it a kernel
// of its own.
struct EpiStore {            // y = A x
  double *y;
  double *y2 = nullptr;
  const double *d2 = nullptr;
  double s2 = 0.0;
};
struct EpiAxpby {            // y = a * (A x) + b * z          (z may alias y)
  double *y;
  const double *z;
  double a, b;
  double *y2 = nullptr;
  const double *d2 = nullptr;
  double s2 = 0.0;
};
struct Jacobi1 {             // second output request: y2 = s2 * d2 .* y
  double *y2 = nullptr;
  const double *d2 = nullptr;
  double s2 = 0.0;
};
struct EpiCheb {             // out = add + c0 p0 + c1 p1 + c2 dinv (b - A p1)   (p0, add optional)
  double *out;
  const double *p0, *p1, *b, *dinv, *add;
  double c0, c1, c2;
};

// Upload one operator: picks the storage format from the row-length histogram
// (SELL-32-sigma for short rows, CSR + sub-warp-per-row kernel for long rows).
// `val` may be null (pattern only); csr_set_values refreshes the numbers later.
// n_own_split >= 0: rows with a ghost column go to a separate boundary part; n_own_cols: number of
// owned columns when the column space is [owned | ghost] (-1: all owned)
void csr_upload_pattern(Ctx &c, DevCsr &A, const HostCsr &h, const std::string &tag, int64_t n_own_split = -1, int bs = 1,
                        int64_t n_own_cols = -1);
void csr_set_values(Ctx &c, DevCsr &A, const HostCsr &pattern, const double *val, bool want_dinv);
void spmv_store(Ctx &c, const DevCsr &A, const double *x, double *y, const Jacobi1 &j = Jacobi1());
void spmv_axpby(Ctx &c, const DevCsr &A, const double *x, double a, double b, const double *z, double *y,
                const Jacobi1 &j = Jacobi1());
void spmv_cheb(Ctx &c, const DevCsr &A, const EpiCheb &e);

// BLAS-1 class kernels ------------------------------------------------------
void vec_copy(Ctx &c, int64_t n, const double *x, double *y);
void vec_scale(Ctx &c, int64_t n, double a, double *x);                       // x *= a
void vec_axpy(Ctx &c, int64_t n, double a, const double *x, double *y);       // y += a x
void vec_axpby(Ctx &c, int64_t n, double a, const double *x, double b, const double *y, double *out);
void vec_zero(Ctx &c, int64_t n, double *x);
void vec_pointwise_scale(Ctx &c, int64_t n, double s, const double *d, const double *b, const double *add, double *out); // out = add + s d b
void vec_copy_bc(Ctx &c, int64_t n, const double *x, double *z, const int32_t *idx, const double *val, int32_t nbc);
void vec_scatter_bc(Ctx &c, double *z, const int32_t *idx, const double *val, int32_t nbc);
void vec_gather(Ctx &c, int64_t n, const int64_t *idx, const double *src, double *dst);   // dst[i] = src[idx[i]]
void vec_scatter(Ctx &c, int64_t n, const int64_t *idx, const double *src, double *dst);  // dst[idx[i]] = src[i]

// deterministic reductions (two-stage, no atomics); results land in device memory.
// Vptrs_dev: device array of pointers to the basis vectors (allocated lazily).
// h[i] = v_i . w for i < nvec
void multi_dot_ptrs(Ctx &c, int64_t n, const double *const *Vptrs_dev, int nvec, const double *w, double *h_dev);
// same plus h[nvec] = w . w (the one-reduction Gram-Schmidt of gmres.cu)
void multi_dot_ww_ptrs(Ctx &c, int64_t n, const double *const *Vptrs_dev, int nvec, const double *w, double *h_dev);
// w -= sum_i h[i] v_i ; nrm2_dev[0] = ||w||^2 (after the update)
void multi_axpy_norm_ptrs(Ctx &c, int64_t n, const double *const *Vptrs_dev, int nvec, const double *h_dev, double *w,
                          double *nrm2_dev);
// x += sum_i y[i] Z_i (y on device)
void multi_axpy_ptrs(Ctx &c, int64_t n, const double *const *Zptrs_dev, int nvec, const double *y_dev, double *x);
// v = w / sqrt(nrm2_dev[0])
void vec_scale_inv_sqrt(Ctx &c, int64_t n, const double *nrm2_dev, const double *w, double *v);
void dot(Ctx &c, int64_t n, const double *x, const double *y, double *out_dev);
// x = (Minv (x) I_bs) b: Minv dense row-major nrows x ncols (scalar), b and x with bs interleaved components
void dense_gemv(Ctx &c, int nrows, int ncols, int bs, const double *Minv, const double *b, double *x);

// ---------------------------------------------------------------------------
// AMG
// ---------------------------------------------------------------------------
struct AmgParams {
  double theta = 0.08;
  int max_levels = 12;
  int coarse_size = 400;
  int smooth_steps = 2;
  double eig_ratio = 10.0;
  double omega_scale = 4.0 / 3.0;
  int lag = 1;                 // rebuild the coarse levels only at every lag-th value refresh (level 0 always
                               // follows the new matrix) -- the role of PETSc's -pc_gamg_reuse_interpolation
  double p_trunc = 0.2;        // drop prolongator entries below p_trunc*max|row|, rescale to the row sum
  int64_t replicate_size = 0;        // multi-rank, > 0: levels with <= this many (global) rows are gathered and
                                     // coarsened/applied redundantly on every rank.  Off by default: +6 % at N=2
                                     // but slower (and 27 instead of 20 iterations) at N=8 with 300000
  double coarse_drop = 0.0;    // > 0: lump coarse entries below drop*sqrt(|a_ii||a_jj|) onto the diagonal
  int refresh = 1;             // value refresh of an existing hierarchy: 0 rebuild on the host (or lag),
                               // 1 frozen prolongators + Galerkin values recomputed on the device (amg_refresh.cu;
                               // single-rank contexts, multi-rank ones fall back to 0)
};

struct HostLevel {
  HostCsr A, P, R;                  // A: local rows, columns [owned | ghost]; P, R: rank local
  HostCsr Pext;                     // multi-rank: P with the rows of A's ghost dofs appended and the columns in the
                                    // local numbering [owned | ghost] of the next level (frozen-P Galerkin refresh)
  std::vector<double> dinv;
  double rho = 1.0;
  std::shared_ptr<HaloPlan> halo;   // null on single-rank contexts
  int64_t n_own = 0;
  std::vector<int64_t> begins;      // ownership offsets of this level
};

struct HostHierarchy {
  std::vector<HostLevel> levels;
  std::vector<double> coarse_inv;   // dense row-major [n_own x coarse_cols]
  int64_t coarse_cols = 0, coarse_maxloc = 0;
  std::vector<int32_t> coarse_gcol; // global column id of every stored entry of the coarsest level
  std::vector<int64_t> coarse_begins;   // ownership offsets of the coarsest level
  // multi-rank: from the first level with <= replicate_size global rows on, the hierarchy is
  // built and applied redundantly on every rank (one all-gather per V-cycle instead of a halo
  // exchange per SpMV on levels whose kernels are far shorter than an exchange)
  std::shared_ptr<HostHierarchy> tail;
  std::vector<int64_t> tail_begins;  // ownership offsets of the first replicated level
};

void host_spgemm(const HostCsr &A, const HostCsr &B, HostCsr &C);     // C = A B, rows sorted
void host_transpose(const HostCsr &A, HostCsr &T);
// dense inverse of the coarsest level from its current values (collective: the rows are gathered)
void amg_coarse_inverse_host(Ctx &c, HostHierarchy &H);
void amg_build_host(Ctx &c, const HostCsr &A_global_cols, std::vector<int64_t> begins, const AmgParams &p,
                    HostHierarchy &H, int level0 = 0);

struct DevLevel {
  DevCsr A_own, P, R;
  DevCsr *Ap = nullptr;      // level 0 aliases the context's operator (no second device copy)
  DevCsr &A() { return *Ap; }
  double rho = 1.0;
  DevBuf<double> x, b, r, w0, w1;
};

struct DevHierarchy {
  std::vector<DevLevel> levels;
  DevBuf<double> coarse_inv, coarse_gather;
  int coarse_n = 0, coarse_cols = 0, coarse_maxloc = 0, coarse_bs = 1;   // vector sizes (bs components per scalar row)
  bool serial = false;                 // replicated tail: applied without communication
  std::unique_ptr<DevHierarchy> tail;
  DevBuf<double> tail_gather, tail_b, tail_x, tail_bloc;
  DevBuf<int64_t> tail_map;            // global dof -> position in the padded all-gather buffer
  int64_t tail_n = 0, tail_nloc = 0, tail_off = 0, tail_maxloc = 0;
  AmgParams params;
  HostHierarchy host;     // kept for introspection / refresh
  bool built = false;
  // device-side Galerkin refresh (amg_refresh.cu): plan matrices W_l with A_{l+1}.val = W_l * A_l.val,
  // diagonal positions of the coarse levels
  std::vector<std::vector<DevCsr>> refresh_W;       // per level: row chunks of the plan matrix
  std::vector<std::vector<int64_t>> refresh_row0;   // first coarse value position of every chunk
  std::vector<DevBuf<int32_t>> refresh_diag;
  bool refresh_built = false;
  bool host_vals_stale = false;       // device-side refreshes since the host mirror was last synchronised
};

void amg_upload(Ctx &c, DevHierarchy &H, const std::string &name, DevCsr *level0, int bs);
// new values on level 0 (same pattern): coarse operators recomputed on the device with frozen P
void amg_refresh_device(Ctx &c, DevHierarchy &H, int bs);
void amg_sync_host(Ctx &c, DevHierarchy &H);     // host mirror of the level values after device-side refreshes
// x = Vcycle(b), zero initial guess; b and x are level-0 sized device vectors (may not alias)
void amg_vcycle(Ctx &c, DevHierarchy &H, const double *b, double *x);

// Chebyshev-Jacobi, zero initial guess, `steps` Jacobi applications:
//   out = add + out_scale * cheb(A, b)
// p1_ready: w0 already holds the first iterate s D^-1 b (written by the kernel that produced b, see
// Jacobi1 / cheb_first_step_scale)
void cheb_jacobi(Ctx &c, const DevCsr &A, const double *b, double emin, double emax, int steps,
                 double out_scale, const double *add, double *out, double *w0, double *w1, bool p1_ready = false);
inline double cheb_first_step_scale(double emin, double emax) { return 2.0 / (emax + emin); }

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
enum KspType { KSP_PREONLY = 0, KSP_RICHARDSON, KSP_CHEBYSHEV, KSP_CG };
enum PcType { PC_NONE = 0, PC_JACOBI, PC_AMG };

struct InnerOpts {
  KspType ksp = KSP_RICHARDSON;
  PcType pc = PC_AMG;
  int max_it = 1;
  double rtol = 0.0;
  double emin = 0.5, emax = 2.0;
  AmgParams amg;
};

struct Timer {
  double ms = 0.0;
  int64_t calls = 0;
  double bytes = 0.0;       // algorithmic bytes moved by the timed launches (0: not a single-kernel timer)
};
struct PendingTimer {
  std::string name;
  cudaEvent_t a, b;
  double bytes;
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 148;
  int64_t launches = 0;

  // distributed
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  ncclComm_t comm_halo = nullptr;          // second communicator: overlapped halo exchanges on comm_stream
  cudaStream_t comm_stream = nullptr;      // halo exchanges overlap the interior rows
  cudaEvent_t ev_x = nullptr, ev_halo = nullptr;

  std::vector<int64_t> u_begins, p_begins;   // ownership offsets of all ranks

  // layout
  bool have_layout = false;
  int64_t n_u = 0, u_begin = 0, n_u_global = 0;
  int64_t n_p = 0, p_begin = 0, n_p_global = 0;

  // options
  int variant = FNP_MAT_COUNT;  // set in ctor
  bool flexible = true;
  int restart = 150;
  double rtol = 1e-6, atol = 1e-50;
  int max_it = 10000;
  InnerOpts opt_u, opt_ap, opt_mp, opt_rp;
  // PCDR (reference preconditioners.py:173-298): Rp = Bt^T diag(Mu)^-1 Bt, built on the host from A01
  std::vector<double> mu_diag;
  bool mu_dirty = false;
  HostCsr h_rp;
  DevCsr rp;
  DevHierarchy amg_rp;
  int spmv_mode = 0;            // 0 auto (by row-length histogram), 1 CSR vector kernel always, 2 SELL always
  int sell_gather = 79;         // SpMV kernel variant bits (kernels.cu): 1 wide gathers, 2 six CTAs/SM, 4 L2 prefetch, 8 wide epilogue (SELL);
                                // 16 L2 prefetch in the CSR sub-warp kernel (experimental)
  int sell_sigma = 1024;        // SELL sorting window (rows)
  int64_t refresh_chunk_terms = (int64_t)1 << 30;   // Galerkin refresh plan: terms per device chunk (option fnp_refresh_chunk_terms)
  int64_t sell_warps_rows = 600000;                  // multi-warp SELL kernel below this many threads (option fnp_sell_warps_rows)
  int sell_warps = 0;           // warps per SELL slice: 0 auto (from rows x mean row), 1 / 2 / 4 / 8 forced
  double sell_max_mean_row = 64.0;   // auto: operators with a longer mean row keep CSR + sub-warp per row
  int timers_on = 0;            // 0 off, 1 stage timers, 2 also one timer per SpMV launch

  // operators
  HostCsr hmat[FNP_MAT_COUNT];          // sorted host copies (pattern always; values for AMG operators)
  std::vector<int32_t> local_cols[FNP_MAT_COUNT];   // multi-rank: columns in local [owned | ghost] numbering
  int kron = 1;                                     // option fnp_kronecker: detect S (x) I_bs velocity blocks
  int prune = 1;                                    // option fnp_prune_zeros: drop stored zeros of A00/P00 at the first upload
  bool pattern_pending[FNP_MAT_COUNT] = {};
  std::vector<int32_t> user_rowptr[FNP_MAT_COUNT], user_col[FNP_MAT_COUNT];
  std::vector<char> prune_mask[FNP_MAT_COUNT];      // per user entry: kept (1) or dropped as a stored zero (0)
  int64_t user_nnz[FNP_MAT_COUNT] = {};
  // device-side value refresh (ingest.cu): stored entry -> position in the caller's value array.
  // d_map: expanded (sorted, pruned, renumbered) entry -> user entry, empty when that is the identity;
  // d_exp_rowptr: row pointers of the expanded pattern in Kronecker mode (the device operator holds the
  // scalar pattern); d_dropped: user entries dropped as stored zeros (checked on every refresh)
  DevBuf<int32_t> d_map[FNP_MAT_COUNT], d_exp_rowptr[FNP_MAT_COUNT], d_dropped[FNP_MAT_COUNT];
  int64_t n_dropped[FNP_MAT_COUNT] = {};
  bool host_vals_valid[FNP_MAT_COUNT] = {};         // hmat[w].val mirrors the device values (fetched lazily)
  DevBuf<double> stage_vals;                         // staging of host value arrays (persistent)
  DevBuf<int> d_flag;                                // error / decision word of the ingest kernels
  int pattern_gen[FNP_MAT_COUNT] = {};               // bumped whenever the stored pattern of an operator is rebuilt
  int kron_bs[FNP_MAT_COUNT] = {1, 1, 1, 1, 1, 1, 1, 1, 1};
  std::vector<int32_t> kron_rowptr[FNP_MAT_COUNT];  // row pointers of the expanded (user) pattern, for value checks
  bool have_pattern[FNP_MAT_COUNT] = {};
  bool have_values[FNP_MAT_COUNT] = {};
  bool dirty[FNP_MAT_COUNT] = {};
  DevCsr dmat[FNP_MAT_COUNT];
  DevBuf<int32_t> bc_idx;
  DevBuf<double> bc_val;
  int32_t nbc = 0;
  DevBuf<int64_t> is_u, is_p;
  bool have_is = false;
  bool is_setup = false;

  DevHierarchy amg_u, amg_ap;
  int amg_u_age = 0;            // value refreshes since the velocity hierarchy was last rebuilt
  int amg_u_gen = -1;           // pattern generation of the velocity block the hierarchy was built for
  int amg_u_which = -1;         // FNP_MAT_A00 or FNP_MAT_P00: the operator the velocity hierarchy belongs to

  // work space
  DevBuf<double> p_w[7];      // pressure-sized work vectors
  DevBuf<double> u_w[5];      // velocity-sized work vectors
  DevBuf<double> red_partial; // block partials of reductions
  DevBuf<double> red_out;     // small device results (h column, norms, CG scalars)
  double *pinned = nullptr;   // host-pinned mirror of red_out
  size_t pinned_n = 0;
  DevBuf<double> io[4];       // staging for host-pointer calls

  // CUDA graph of one block-triangular PC apply on fixed staging buffers (single-rank
  // contexts): ~500 short launches per apply collapse into one graph launch
  int64_t halo_split_rows = 50000;   // operators with at least this many rows are split into interior / boundary rows
                                     // (interior rows overlap the exchange); smaller ones wait in one kernel
  bool split_rows(int64_t nrows) const { return overlap || (p2p && nrows >= halo_split_rows); }
  int p2p = 2;                  // 2 (default): as 1, and the pack + remote-store kernel of a split operator runs on a forked
                                // stream beside the interior rows (53 M dofs on 8 GPUs: 48.1 ms against 56.1 ms with NCCL);
                                // 1: peer-memory halo exchange (cudaIpc stores + flags, wait fused into the consumer
                                // kernel, device-resident sequence numbers); 0: NCCL send/recv
  DevBuf<int> p2p_err;          // set by a timed-out flag wait
  int overlap = 0;              // 1: split SELL operators into interior/boundary rows and overlap the halo exchange
                                // on a second stream/communicator (measured SLOWER with NCCL send/recv: 5.8 vs 4.9 ms per apply)
  int use_graph = 2;            // 0 off, 1 single-rank only, 2 also multi-rank (NCCL send/recv captured in the graph)
  cudaGraphExec_t pc_graph = nullptr;
  int64_t pc_graph_nodes = 0;
  DevBuf<double> g_in, g_out;
  void drop_graph();

  // Krylov basis
  std::vector<DevBuf<double>> V, Z;
  DevBuf<double> kr_w, kr_x, kr_b;
  DevBuf<double> kr_H, kr_small;        // device-resident Hessenberg matrix, rotations, residual estimates (gmres.cu)
  std::vector<cudaEvent_t> ev_iter;     // one event per in-flight iteration (two)
  bool gmres_two_reductions = false;    // multi-rank: explicit norm with its own all-reduce (set after a cancellation)
  int gmres_sync = 0;                   // option fnp_gmres_sync: look at every iteration's result before the next is enqueued
  int converged_reason = 0;             // KSPConvergedReason of the last solve (2 rtol, 3 atol, -3 iterations)
  DevBuf<double> sol_x, sol_b, sol_m;   // staging of the user's b / x (split and monolithic layouts)
  std::vector<double> res_hist;

  // timers
  std::map<std::string, Timer> timers;
  std::vector<PendingTimer> pending;      // recorded, not yet resolved (no sync on the hot path)
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t get_event();
  void resolve_timers();
  cudaEvent_t ev_tic = nullptr, ev_toc = nullptr;

  explicit Ctx(int dev);
  ~Ctx();
  const DevCsr &velocity_pc_matrix() const { return have_values[FNP_MAT_P00] ? dmat[FNP_MAT_P00] : dmat[FNP_MAT_A00]; }
  int velocity_pc_index() const { return have_values[FNP_MAT_P00] ? FNP_MAT_P00 : FNP_MAT_A00; }
};

// scoped stage timer (CUDA events on the context stream); no-op unless timers_on
struct StageTimer {
  Ctx &c;
  std::string name;
  cudaEvent_t a = nullptr;
  double bytes = 0.0;
  StageTimer(Ctx &ctx, const char *nm, int level = 1, double bytes = 0.0);
  StageTimer(Ctx &ctx, const std::string &nm, int level, double bytes = 0.0);
  ~StageTimer();
};

// ingest.cu: pattern finalisation and the device-side value path of fnp_set_values
void ingest_set_pattern(Ctx &c, int which, const int32_t *rowptr, const int32_t *colidx);
void ingest_set_values(Ctx &c, int which, const double *values);
void ensure_host_values(Ctx &c, int which);     // make hmat[which].val current (download from the device copy)

// solver pieces (pcd.cu / gmres.cu)
void setup_all(Ctx &c);
void mp_solve(Ctx &c, const double *b, double out_scale, const double *add, double *x);
void ap_solve(Ctx &c, const double *b, double *x);
void rp_solve(Ctx &c, const double *b, double *x);
void u_solve(Ctx &c, const double *b, double *x);
void schur_apply(Ctx &c, const double *x_p, double *y_p);
void pc_apply(Ctx &c, const double *x_u, const double *x_p, double *y_u, double *y_p);
// same on split vectors [u;p], replayed from a captured CUDA graph when possible
void pc_apply_vec(Ctx &c, const double *x, double *y);
void system_matvec(Ctx &c, const double *x, double *y);   // split vectors [u;p]
void solve_fgmres(Ctx &c, const double *b, double *x, int32_t *its, double *rnorm, int32_t *napply);

}  // namespace fnp
