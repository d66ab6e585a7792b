// Multi-rank plumbing: one process (or host thread) per GPU, rows of every operator
// partitioned in contiguous ownership ranges (PETSc's MPIAIJ layout, which the
// reference inherits from DOLFIN: fenapack/SubfieldBC.h:138-140).  The only data-path
// exchanges are (i) the ghost entries of x before an SpMV (the MatMult VecScatter
// of the reference) and (ii) small all-reduces for the Krylov scalars -- both NCCL
// over NVLink, ordered on the context's stream.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <numeric>

#include "fnp_internal.cuh"

namespace fnp {

double comm_allreduce(Ctx &c, double v, bool max_op);

__global__ void pack_kernel(int32_t n, const int32_t *__restrict__ idx, const double *__restrict__ x,
                            double *__restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = x[idx[i]];
}

// ---- peer-memory exchange ---------------------------------------------------------
// send: every element goes straight into the neighbour's ghost slot (remote store over
// NVLink); the last block to finish publishes the sequence number in the neighbours'
// flag words.  __threadfence_system orders the data before the flag.
__global__ void __launch_bounds__(256)
p2p_send_kernel(int32_t nsend, const int32_t *__restrict__ idx, const double *__restrict__ x, int nranks,
                const int *__restrict__ send_off, const int *__restrict__ send_cnt, double *const *__restrict__ peer_dst,
                const long long *__restrict__ peer_stride, unsigned long long *const *__restrict__ peer_flag,
                unsigned long long *seq_dev, unsigned int *counter) {
  // the sequence number of THIS exchange; *seq_dev is advanced by the last block to finish, i.e.
  // after every block of the launch has read it
  const unsigned long long seq = *((volatile unsigned long long *)seq_dev) + 1ull;
  const int slot = (int)(seq & 1ull);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nsend) {
    int q = 0;
    while (q + 1 < nranks && i >= send_off[q] + send_cnt[q]) ++q;     // segments are ordered by peer
    while (send_cnt[q] == 0 && q + 1 < nranks) ++q;
    double *dst = peer_dst[q] + (long long)slot * peer_stride[q] + (i - send_off[q]);
    *dst = x[idx[i]];
  }
  // block barrier, then ONE system-scope fence per block: the fence is cumulative over the
  // stores the barrier made visible to thread 0 (a fence in every thread costs ~10x more)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {
      *counter = 0;
      *((volatile unsigned long long *)seq_dev) = seq;
      __threadfence_system();
      for (int q = 0; q < nranks; ++q)
        if (send_cnt[q] > 0) *((volatile unsigned long long *)peer_flag[q]) = seq;
    }
  }
}

// stand-alone wait (host-side consumers): one thread per peer spins on the local flag word
// (bounded: ~4 s, then the error flag is raised instead of hanging the GPU)
__global__ void p2p_wait_kernel(const HaloWaitDev *__restrict__ hw) {
  const int q = threadIdx.x;
  if (q >= hw->nranks || hw->recv_cnt[q] == 0) return;
  const unsigned long long seq = *hw->seq;
  const long long t0 = clock64();
  while (*((volatile const unsigned long long *)(hw->flags + q)) < seq) {
    if (clock64() - t0 > 8000000000ll) {
      *hw->err = 1;
      return;
    }
  }
}

HaloPlan::~HaloPlan() {
  for (void *p : peer_base)
    if (p) cudaIpcCloseMemHandle(p);
}

void halo_exchange(Ctx &c, HaloPlan &h, const double *x_own, cudaStream_t stream, ncclComm_t comm) {
  if (h.p2p) {
    if (h.nsend > 0) {
      p2p_send_kernel<<<(h.nsend + 255) / 256, 256, 0, stream>>>(h.nsend, h.send_idx.p, x_own, c.nranks, h.d_send_off.p,
                                                                  h.d_send_cnt.p, h.d_peer_dst.p, h.d_peer_stride.p,
                                                                  h.d_peer_flag.p, h.d_seq.p, h.d_counter.p);
      c.launches++;
      FNP_CUDA(cudaPeekAtLastError());
    }
    h.current_ghost = nullptr;        // the slot follows the device-resident sequence number
    return;
  }
  if (h.nsend > 0) {
    pack_kernel<<<(h.nsend + 255) / 256, 256, 0, stream>>>(h.nsend, h.send_idx.p, x_own, h.send_buf.p);
    c.launches++;
    FNP_CUDA(cudaPeekAtLastError());
  }
  const NcclApi &n = nccl();
  FNP_NCCL(n.GroupStart());
  for (int q = 0; q < c.nranks; ++q) {
    if (h.send_count[q] > 0)
      FNP_NCCL(n.Send(h.send_buf.p + h.send_off[q], (size_t)h.send_count[q], ncclDouble, q, comm, stream));
    if (h.recv_count[q] > 0)
      FNP_NCCL(n.Recv(h.ghost.p + h.recv_off[q], (size_t)h.recv_count[q], ncclDouble, q, comm, stream));
  }
  FNP_NCCL(n.GroupEnd());
  h.current_ghost = h.ghost.p;
}

void halo_wait(Ctx &c, HaloPlan &h, cudaStream_t stream) {
  if (!h.p2p || h.nghost == 0) return;
  p2p_wait_kernel<<<1, 32, 0, stream>>>(h.d_wait.p);
  c.launches++;
  FNP_CUDA(cudaPeekAtLastError());
}

const double *halo_ghost_after_wait(Ctx &c, HaloPlan &h) {
  if (!h.p2p) return h.current_ghost;
  unsigned long long seq = 0;
  FNP_CUDA(cudaMemcpyAsync(&seq, h.d_seq.p, sizeof(seq), cudaMemcpyDeviceToHost, c.stream));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  return h.arena.p + (size_t)(seq & 1ull) * (size_t)h.nghost;
}

// cudaIpcGetMemHandle describes the BASE allocation a pointer lives in (cudaMalloc may
// sub-allocate); the peers need the offset of our buffer inside it.
static long long offset_in_allocation(const void *p) {
  typedef int (*fn_t)(unsigned long long *, size_t *, unsigned long long);
  static fn_t fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (h) fn = reinterpret_cast<fn_t>(dlsym(h, "cuMemGetAddressRange_v2"));
  }
  if (!fn) return -1;
  unsigned long long base = 0;
  size_t size = 0;
  if (fn(&base, &size, (unsigned long long)(uintptr_t)p) != 0) return -1;
  return (long long)((unsigned long long)(uintptr_t)p - base);
}

// Switch a plan to the peer-memory path.  Collective; every rank takes the same decision.
void halo_enable_p2p(Ctx &c, HaloPlan &h) {
  const int R = c.nranks, me = c.rank;
  if (R == 1 || !c.p2p) return;
  FNP_REQUIRE(R <= 32, FNP_ERR_ARG, "peer-memory halo supports up to 32 ranks");
  // neighbour sets must be symmetric: the two-slot scheme relies on "I hear from q
  // whenever I talk to q" to know that q has finished reading the slot I overwrite
  double ok = 1.0;
  for (int q = 0; q < R; ++q)
    if ((h.send_count[q] > 0) != (h.recv_count[q] > 0)) ok = 0.0;
  if (!c.p2p_err.p) {
    c.p2p_err.alloc(1);
    c.p2p_err.zero(c.stream);
  }
  // >= 2 MiB so that the arena is an allocation of its own (small cudaMallocs share a base
  // block, and one IPC handle cannot be opened twice by the same peer)
  h.arena.alloc(std::max<size_t>(2 * (size_t)h.nghost + (size_t)R, (size_t)262144));
  h.arena.zero(c.stream);
  cudaIpcMemHandle_t mine;
  if (cudaIpcGetMemHandle(&mine, h.arena.p) != cudaSuccess) { ok = 0.0; cudaGetLastError(); }
  const long long my_offset = offset_in_allocation(h.arena.p);
  if (my_offset < 0) ok = 0.0;
  ok = -comm_allreduce(c, -ok, true);
  if (ok < 0.5) { h.arena.release(); return; }
  // all-gather: IPC handles (64 B), ghost counts, receive offsets
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  const size_t rec = 64 + sizeof(long long) * (2 + (size_t)R);
  std::vector<char> sendrec(rec), all(rec * R);
  std::memcpy(sendrec.data(), &mine, 64);
  long long *meta = reinterpret_cast<long long *>(sendrec.data() + 64);
  meta[0] = h.nghost;
  meta[1] = my_offset;
  for (int q = 0; q < R; ++q) meta[2 + q] = h.recv_off[q];
  DevBuf<char> d(rec * (R + 1));
  FNP_CUDA(cudaMemcpyAsync(d.p + rec * R, sendrec.data(), rec, cudaMemcpyHostToDevice, c.stream));
  FNP_NCCL(nccl().AllGather(d.p + rec * R, d.p, rec, ncclChar, c.comm, c.stream));
  FNP_CUDA(cudaMemcpyAsync(all.data(), d.p, rec * R, cudaMemcpyDeviceToHost, c.stream));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  std::vector<double *> dst(R, nullptr);
  std::vector<unsigned long long *> flag(R, nullptr);
  std::vector<long long> stride(R, 0);
  h.peer_base.assign(R, nullptr);
  double opened = 1.0;
  for (int q = 0; q < R; ++q) {
    if (q == me || h.send_count[q] == 0) continue;
    cudaIpcMemHandle_t hq;
    std::memcpy(&hq, all.data() + rec * q, 64);
    const long long *mq = reinterpret_cast<const long long *>(all.data() + rec * q + 64);
    void *base = nullptr;
    if (cudaIpcOpenMemHandle(&base, hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      opened = 0.0;
      continue;
    }
    h.peer_base[q] = base;
    double *arena_q = reinterpret_cast<double *>(static_cast<char *>(base) + mq[1]);
    stride[q] = mq[0];
    dst[q] = arena_q + mq[2 + me];                               // where q expects my entries
    flag[q] = reinterpret_cast<unsigned long long *>(arena_q + 2 * mq[0]) + me;
  }
  opened = -comm_allreduce(c, -opened, true);
  if (opened < 0.5) {
    for (void *&p : h.peer_base) { if (p) cudaIpcCloseMemHandle(p); p = nullptr; }
    h.arena.release();
    return;
  }
  h.d_peer_dst.upload(dst.data(), R, c.stream);
  h.d_peer_flag.upload(flag.data(), R, c.stream);
  h.d_peer_stride.upload(stride.data(), R, c.stream);
  h.d_send_off.upload(h.send_off.data(), R, c.stream);
  h.d_send_cnt.upload(h.send_count.data(), R, c.stream);
  h.d_recv_cnt.upload(h.recv_count.data(), R, c.stream);
  h.d_counter.alloc(1);
  h.d_counter.zero(c.stream);
  h.d_seq.alloc(1);
  h.d_seq.zero(c.stream);
  HaloWaitDev hw;
  hw.arena = h.arena.p;
  hw.seq = h.d_seq.p;
  hw.flags = reinterpret_cast<const unsigned long long *>(h.arena.p + 2 * (size_t)h.nghost);
  hw.recv_cnt = h.d_recv_cnt.p;
  hw.err = c.p2p_err.p;
  hw.nranks = R;
  hw.nghost = h.nghost;
  h.d_wait.upload(&hw, 1, c.stream);
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  // nobody may start sending before every rank has zeroed its flags and mapped its peers
  comm_allreduce(c, 0.0, false);
  h.p2p = true;
}

// ---- small host-level collectives (set-up time), staged through device memory ----
std::vector<int64_t> comm_allgather_i64(Ctx &c, int64_t v) {
  std::vector<int64_t> out((size_t)c.nranks, v);
  if (c.nranks == 1) return out;
  DevBuf<int64_t> d((size_t)c.nranks + 1);
  FNP_CUDA(cudaMemcpyAsync(d.p + c.nranks, &v, sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
  FNP_NCCL(nccl().AllGather(d.p + c.nranks, d.p, 1, ncclInt64, c.comm, c.stream));
  FNP_CUDA(cudaMemcpyAsync(out.data(), d.p, c.nranks * sizeof(int64_t), cudaMemcpyDeviceToHost, c.stream));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  return out;
}

double comm_allreduce(Ctx &c, double v, bool max_op) {
  if (c.nranks == 1) return v;
  DevBuf<double> d(1);
  FNP_CUDA(cudaMemcpyAsync(d.p, &v, sizeof(double), cudaMemcpyHostToDevice, c.stream));
  FNP_NCCL(nccl().AllReduce(d.p, d.p, 1, ncclDouble, max_op ? ncclMax : ncclSum, c.comm, c.stream));
  FNP_CUDA(cudaMemcpyAsync(&v, d.p, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  return v;
}

// Padded all-gather of host doubles: every rank contributes `count` (<= maxcount) values;
// result is [nranks x maxcount] row-major.
std::vector<double> comm_allgather_padded(Ctx &c, const double *v, int64_t count, int64_t maxcount) {
  std::vector<double> out((size_t)c.nranks * maxcount, 0.0);
  if (c.nranks == 1) {
    std::copy(v, v + count, out.begin());
    return out;
  }
  DevBuf<double> d((size_t)(c.nranks + 1) * maxcount);
  d.zero(c.stream);
  if (count) FNP_CUDA(cudaMemcpyAsync(d.p + (size_t)c.nranks * maxcount, v, count * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  FNP_NCCL(nccl().AllGather(d.p + (size_t)c.nranks * maxcount, d.p, (size_t)maxcount, ncclDouble, c.comm, c.stream));
  FNP_CUDA(cudaMemcpyAsync(out.data(), d.p, out.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  return out;
}

// Ownership offsets of all ranks from the local begin: begins[q] .. begins[q+1]
std::vector<int64_t> comm_ranges(Ctx &c, int64_t n_local) {
  std::vector<int64_t> counts = comm_allgather_i64(c, n_local);
  std::vector<int64_t> begins((size_t)c.nranks + 1, 0);
  for (int q = 0; q < c.nranks; ++q) begins[q + 1] = begins[q] + counts[q];
  return begins;
}

// Build the halo plan of a matrix whose columns are GLOBAL ids of a space partitioned
// by `begins`, and relabel the columns in place to [owned 0..n_own) | ghosts n_own..).
// Returns null (and only shifts the columns) on single-rank contexts.
std::shared_ptr<HaloPlan> build_halo(Ctx &c, HostCsr &h, const std::vector<int64_t> &begins,
                                     std::vector<int64_t> *ghost_global_out) {
  const int R = c.nranks, me = c.rank;
  const int64_t b0 = begins[me], b1 = begins[me + 1];
  const int64_t n_own = b1 - b0;
  const int64_t nnz = h.nnz();
  if (R == 1) {
    h.ncols = n_own;
    if (ghost_global_out) ghost_global_out->clear();
    return nullptr;
  }
  // ghost columns, sorted unique
  std::vector<int64_t> ghosts;
  for (int64_t k = 0; k < nnz; ++k) {
    const int64_t g = h.col[k];
    if (g < b0 || g >= b1) ghosts.push_back(g);
  }
  std::sort(ghosts.begin(), ghosts.end());
  ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
  const int64_t ng = (int64_t)ghosts.size();
  FNP_REQUIRE(n_own + ng < (int64_t)INT32_MAX, FNP_ERR_ARG, "local column space exceeds 32-bit indices");
  // relabel
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nnz; ++k) {
    const int64_t g = h.col[k];
    if (g >= b0 && g < b1) {
      h.col[k] = (int32_t)(g - b0);
    } else {
      const int64_t pos = std::lower_bound(ghosts.begin(), ghosts.end(), g) - ghosts.begin();
      h.col[k] = (int32_t)(n_own + pos);
    }
  }
  h.ncols = n_own + ng;

  auto plan = std::make_shared<HaloPlan>();
  plan->send_count.assign(R, 0); plan->send_off.assign(R, 0);
  plan->recv_count.assign(R, 0); plan->recv_off.assign(R, 0);
  // what I need from each owner (ghosts are sorted, hence grouped by owner)
  std::vector<int32_t> request((size_t)ng);
  {
    int q = 0;
    for (int64_t i = 0; i < ng; ++i) {
      while (ghosts[i] >= begins[q + 1]) ++q;
      plan->recv_count[q]++;
      request[i] = (int32_t)(ghosts[i] - begins[q]);     // local index at the owner
    }
    for (int q2 = 1; q2 < R; ++q2) plan->recv_off[q2] = plan->recv_off[q2 - 1] + plan->recv_count[q2 - 1];
  }
  plan->nghost = (int32_t)ng;
  // counts matrix: all_need[r * R + q] = what rank r needs from rank q
  DevBuf<int32_t> d_counts((size_t)R * R + R);
  FNP_CUDA(cudaMemcpyAsync(d_counts.p + (size_t)R * R, plan->recv_count.data(), R * sizeof(int32_t), cudaMemcpyHostToDevice, c.stream));
  FNP_NCCL(nccl().AllGather(d_counts.p + (size_t)R * R, d_counts.p, (size_t)R, ncclInt32, c.comm, c.stream));
  std::vector<int32_t> all_need((size_t)R * R);
  FNP_CUDA(cudaMemcpyAsync(all_need.data(), d_counts.p, all_need.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  int64_t nsend = 0;
  for (int q = 0; q < R; ++q) {
    plan->send_count[q] = all_need[(size_t)q * R + me];
    plan->send_off[q] = (int)nsend;
    nsend += plan->send_count[q];
  }
  plan->nsend = (int32_t)nsend;
  // index lists: my requests go to the owners, their requests come to me
  DevBuf<int32_t> d_req((size_t)std::max<int64_t>(ng, 1));
  plan->send_idx.alloc((size_t)std::max<int64_t>(nsend, 1));
  if (ng) FNP_CUDA(cudaMemcpyAsync(d_req.p, request.data(), ng * sizeof(int32_t), cudaMemcpyHostToDevice, c.stream));
  const NcclApi &n = nccl();
  FNP_NCCL(n.GroupStart());
  for (int q = 0; q < R; ++q) {
    if (plan->recv_count[q] > 0)
      FNP_NCCL(n.Send(d_req.p + plan->recv_off[q], (size_t)plan->recv_count[q], ncclInt32, q, c.comm, c.stream));
    if (plan->send_count[q] > 0)
      FNP_NCCL(n.Recv(plan->send_idx.p + plan->send_off[q], (size_t)plan->send_count[q], ncclInt32, q, c.comm, c.stream));
  }
  FNP_NCCL(n.GroupEnd());
  plan->send_buf.alloc((size_t)std::max<int64_t>(nsend, 1));
  plan->ghost.alloc((size_t)std::max<int64_t>(ng, 1));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  if (ghost_global_out) *ghost_global_out = ghosts;
  halo_enable_p2p(c, *plan);
  return plan;
}

// Plan for bs interleaved components per (scalar) entry of `p`.
std::shared_ptr<HaloPlan> expand_plan(Ctx &c, const HaloPlan &p, int bs) {
  auto e = std::make_shared<HaloPlan>();
  const int R = (int)p.send_count.size();
  e->send_count.resize(R); e->send_off.resize(R); e->recv_count.resize(R); e->recv_off.resize(R);
  for (int q = 0; q < R; ++q) {
    e->send_count[q] = p.send_count[q] * bs; e->send_off[q] = p.send_off[q] * bs;
    e->recv_count[q] = p.recv_count[q] * bs; e->recv_off[q] = p.recv_off[q] * bs;
  }
  e->nsend = p.nsend * bs;
  e->nghost = p.nghost * bs;
  std::vector<int32_t> idx((size_t)std::max(p.nsend, 1)), idx2((size_t)std::max(e->nsend, 1));
  if (p.nsend) FNP_CUDA(cudaMemcpyAsync(idx.data(), p.send_idx.p, p.nsend * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  for (int32_t i = 0; i < p.nsend; ++i)
    for (int b = 0; b < bs; ++b) idx2[(size_t)i * bs + b] = idx[i] * bs + b;
  e->send_idx.upload(idx2.data(), idx2.size(), c.stream);
  e->send_buf.alloc((size_t)std::max(e->nsend, 1));
  e->ghost.alloc((size_t)std::max(e->nghost, 1));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  if (p.p2p) halo_enable_p2p(c, *e);
  return e;
}

// Exchange one host vector through a halo plan: returns the ghost values (set-up helper).
std::vector<double> halo_exchange_host(Ctx &c, HaloPlan &plan, const std::vector<double> &x_own) {
  DevBuf<double> d(std::max<size_t>(x_own.size(), 1));
  if (!x_own.empty()) FNP_CUDA(cudaMemcpyAsync(d.p, x_own.data(), x_own.size() * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  halo_exchange(c, plan, d.p, c.stream, c.comm);
  halo_wait(c, plan, c.stream);
  std::vector<double> g((size_t)plan.nghost);
  const double *ghost = halo_ghost_after_wait(c, plan);
  if (plan.nghost) FNP_CUDA(cudaMemcpyAsync(g.data(), ghost, g.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  return g;
}

}  // namespace fnp
