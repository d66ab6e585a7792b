// SpMV-class and BLAS-1-class kernels of libfenapack_cuda (sm_100a, fp64).
//
// Every stage of the PCD apply is HBM-bound (0.17 flop/byte for CSR SpMV), so
// the kernels are organised around memory traffic only:
//   * spmv_kernel<LANES, Epi>: a sub-warp of LANES threads per row (LANES chosen
//     from the operator's row-length histogram), coalesced (col,val) streams,
//     two independent partial sums per lane for memory-level parallelism,
//     warp-shuffle segmented reduction, and a fused epilogue so that residuals,
//     Chebyshev three-term updates and prolongation corrections never cost a
//     second pass over the vectors.
//   * reductions are two-stage and atomic-free: bit-reproducible run to run.
#include <algorithm>

#include "fnp_internal.cuh"

namespace fnp {

static inline int blocks_for(const Ctx &c, int64_t n, int threads, int per_sm) {
  int64_t b = (n + threads - 1) / threads;
  int64_t cap = (int64_t)c.num_sms * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

#define FNP_LAUNCH_CHECK(c)            \
  do {                                 \
    (c).launches++;                    \
    FNP_CUDA(cudaPeekAtLastError());   \
  } while (0)

// ---------------------------------------------------------------------------
// SpMV
// ---------------------------------------------------------------------------
__device__ __forceinline__ void epilogue(const EpiStore &e, int r, double s) {
  e.y[r] = s;
  if (e.y2) e.y2[r] = e.s2 * e.d2[r] * s;
}
__device__ __forceinline__ void epilogue(const EpiAxpby &e, int r, double s) {
  const double v = e.a * s + e.b * e.z[r];
  e.y[r] = v;
  if (e.y2) e.y2[r] = e.s2 * e.d2[r] * v;
}
__device__ __forceinline__ void epilogue(const EpiCheb &e, int r, double s) {
  double v = e.c1 * e.p1[r] + e.c2 * e.dinv[r] * (e.b[r] - s);
  if (e.p0) v += e.c0 * e.p0[r];
  if (e.add) v += e.add[r];
  e.out[r] = v;
}

// x is addressed as [owned entries | ghost buffer]: column c < nown reads the caller's
// vector, c >= nown reads the halo buffer filled by the exchange (dist.cu).  Single-GPU
// contexts pass nown = INT_MAX.
// Kronecker mode (BS > 1): the stored matrix is the scalar operator S of a vector-valued
// block S (x) I_BS with interleaved components (dof = BS*node + comp), the case of the
// Picard/Oseen velocity block and of every level of its AMG hierarchy.  One thread row
// then serves BS vector rows: the (col, val) stream -- 12 B per scalar entry -- is read
// once for BS results, cutting the matrix traffic by BS.
template <int BS>
__device__ __forceinline__ const double *gather_ptr(const double *__restrict__ x, const double *__restrict__ xg, int nown, int c) {
  return c < nown ? x + (int64_t)BS * c : xg + (int64_t)BS * (c - nown);
}

// Peer-memory halo exchange (dist.cu): the ghost entries are stored into this GPU's arena by the
// neighbouring GPUs, which then publish the exchange's sequence number in this GPU's flag words.  A
// kernel that consumes ghosts waits here -- thread 0 of every block spins (bounded) until every
// neighbour's flag has reached the sequence number of the exchange posted just before the launch
// -- and takes the ghost slot of that exchange.  No thread touches the arena before the barrier,
// so no cache of this SM can hold a stale line of it.
__device__ __forceinline__ const double *halo_wait_dev(const HaloWaitDev *__restrict__ hw) {
  const unsigned long long seq = *hw->seq;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int q = 0; q < hw->nranks; ++q) {
      if (hw->recv_cnt[q] == 0) continue;
      while (*((volatile const unsigned long long *)(hw->flags + q)) < seq) {
        if (clock64() - t0 > 8000000000ll) {       // ~4 s: a neighbour died or fell out of step
          *hw->err = 1;
          break;
        }
      }
    }
    __threadfence();
  }
  __syncthreads();
  return hw->arena + (size_t)(seq & 1ull) * (size_t)hw->nghost;
}

// The three components of node c with two 16-byte loads instead of three 8-byte ones: the
// 24-byte group starts either on a 16-byte boundary (take lo.x lo.y hi.x) or 8 bytes past one
// (take lo.y hi.x hi.y).  Every gather instruction of a warp touches the same cache lines whatever
// its width, so the LSU wavefront count of the gather -- the limiter of the Kronecker kernel,
// profiles/r01_spmv_kernel_choice.md -- drops by a third.  The wide loads read 8 bytes beside the
// group: slices that reference the first or last node of the owned vector or of the ghost buffer
// are flagged at build time (bit 0 of their slice pointer) and take the scalar loads.
__device__ __forceinline__ void gather3_wide(const double *__restrict__ p, double &a0, double &a1, double &a2) {
  const bool odd = (reinterpret_cast<uintptr_t>(p) & 8) != 0;
  const double2 *q = reinterpret_cast<const double2 *>(p - (odd ? 1 : 0));
  const double2 lo = __ldg(q), hi = __ldg(q + 1);
  a0 = odd ? lo.y : lo.x;
  a1 = odd ? hi.x : lo.y;
  a2 = odd ? hi.y : hi.x;
}

// PF (option fnp_sell_gather bit 16, off by default; round 2: A10 0.140 -> 0.119 ms in isolation but
// 123 -> 142 us inside the solve): lane 0 of every warp pulls the
// contiguous (col, val) range of the warp's 32 / LANES rows into L2 with two bulk prefetches, as the
// SELL kernel does for its slice (the ends are trimmed to 16-byte boundaries: a prefetch is a hint).
template <int LANES, int BS, class Epi, bool PF>
__global__ void __launch_bounds__(256)
spmv_kernel(int nrows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
            const double *__restrict__ val, const double *__restrict__ x, const double *__restrict__ xg, int nown,
            const HaloWaitDev *__restrict__ hw, Epi epi) {
  if (hw) xg = halo_wait_dev(hw);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = tid / LANES;
  const int lane = tid % LANES;
  if (PF && (threadIdx.x & 31) == 0 && row < nrows) {
    const int k0 = rowptr[row], k1 = rowptr[min(row + 32 / LANES, nrows)];
    const int c0 = (k0 + 3) & ~3, c1 = k1 & ~3;
    if (c1 > c0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(col + c0), "r"((c1 - c0) * 4) : "memory");
    const int v0 = (k0 + 1) & ~1, v1 = k1 & ~1;
    if (v1 > v0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(val + v0), "r"((v1 - v0) * 8) : "memory");
  }
  double s0[BS], s1[BS];
#pragma unroll
  for (int b = 0; b < BS; ++b) s0[b] = s1[b] = 0.0;
  if (row < nrows) {
    const int beg = rowptr[row], end = rowptr[row + 1];
    int k = beg + lane;
    for (; k + LANES < end; k += 2 * LANES) {
      const int c0 = __ldg(col + k), c1 = __ldg(col + k + LANES);
      const double v0 = __ldg(val + k), v1 = __ldg(val + k + LANES);
      const double *p0 = gather_ptr<BS>(x, xg, nown, c0), *p1 = gather_ptr<BS>(x, xg, nown, c1);
#pragma unroll
      for (int b = 0; b < BS; ++b) {
        s0[b] += v0 * __ldg(p0 + b);
        s1[b] += v1 * __ldg(p1 + b);
      }
    }
    if (k < end) {
      const double v0 = __ldg(val + k);
      const double *p0 = gather_ptr<BS>(x, xg, nown, __ldg(col + k));
#pragma unroll
      for (int b = 0; b < BS; ++b) s0[b] += v0 * __ldg(p0 + b);
    }
  }
#pragma unroll
  for (int b = 0; b < BS; ++b) {
    double s = s0[b] + s1[b];
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off, LANES);
    if (row < nrows && lane == 0) epilogue(epi, BS * row + b, s);
  }
}

// ---------------------------------------------------------------------------
// SELL-32-sigma kernel (the default for short-row operators): one thread per row,
// one warp per slice of 32 rows, entries stored column-major inside the slice.
//   * matrix loads go straight to registers, perfectly coalesced (one 128-B line of
//     column indices + two of values per warp instruction), evict-first;
//   * lane r gathers x for the k-th entry of row r: the 32 rows of a slice are
//     neighbouring dofs of the FE operator, so one gather instruction touches 2-5
//     cache lines (15-20 when a warp walks along ONE row as the CSR kernels do);
//   * no shared memory, no shuffles, no barriers; 4 independent entries per thread
//     in flight.
// Why not CSR: on B200 the L1/LSU data pipe issues one 128-B wavefront per cycle per
// SM while HBM delivers ~44 B per cycle per SM, i.e. <= ~8.7 wavefronts per 32
// non-zeros (384 B) at HBM speed.  Sub-warp-per-row CSR spends ~20 (scattered
// gathers), CSR staged through shared memory ~15 (LDGSTS + LDS + gathers); both were
// measured LSU-bound at 35-45 % of HBM peak (profiles/r01_spmv_kernel_choice.md).
// SELL needs ~3 + 3..5.  Rows are sorted by length inside windows of SIGMA rows so
// the padding stays at a few per cent.
// ---------------------------------------------------------------------------
constexpr int SELL_C = 32;
constexpr int SELL_NARROW = 1;   // flag in bit 0 of a slice pointer (pointers are multiples of 32)

// Per-thread sums of one SELL row (entries k = 0 .. len-1 at stride 32).  WIDE selects the
// 16-byte gathers (BS == 3 only); products and their order are the same in both variants.
// (A variant that loaded the column indices of the next group of four ahead of the gathers was
// measured slower in round 2 -- 0.177 / 0.205 ms against 0.170 ms on A00 -- and is gone.)
template <int BS, bool WIDE>
__device__ __forceinline__ void sell_row_sums(const int32_t *__restrict__ cp, const double *__restrict__ vp, int len,
                                              const double *__restrict__ x, const double *__restrict__ xg, int nown,
                                              double (&out)[BS]) {
  double s0[BS], s1[BS];
#pragma unroll
  for (int b = 0; b < BS; ++b) s0[b] = s1[b] = 0.0;
  int k = 0;
  for (; k + 4 <= len; k += 4) {
    const int c0 = __ldcs(cp + (k + 0) * SELL_C);
    const int c1 = __ldcs(cp + (k + 1) * SELL_C);
    const int c2 = __ldcs(cp + (k + 2) * SELL_C);
    const int c3 = __ldcs(cp + (k + 3) * SELL_C);
    const double v0 = __ldcs(vp + (k + 0) * SELL_C), v1 = __ldcs(vp + (k + 1) * SELL_C);
    const double v2 = __ldcs(vp + (k + 2) * SELL_C), v3 = __ldcs(vp + (k + 3) * SELL_C);
    const double *p0 = gather_ptr<BS>(x, xg, nown, c0), *p1 = gather_ptr<BS>(x, xg, nown, c1);
    const double *p2 = gather_ptr<BS>(x, xg, nown, c2), *p3 = gather_ptr<BS>(x, xg, nown, c3);
    double x0[BS], x1[BS], x2[BS], x3[BS];
    if constexpr (WIDE) {
      gather3_wide(p0, x0[0], x0[1], x0[2]);
      gather3_wide(p1, x1[0], x1[1], x1[2]);
      gather3_wide(p2, x2[0], x2[1], x2[2]);
      gather3_wide(p3, x3[0], x3[1], x3[2]);
    } else {
#pragma unroll
      for (int b = 0; b < BS; ++b) {
        x0[b] = __ldg(p0 + b);
        x1[b] = __ldg(p1 + b);
        x2[b] = __ldg(p2 + b);
        x3[b] = __ldg(p3 + b);
      }
    }
#pragma unroll
    for (int b = 0; b < BS; ++b) {
      s0[b] += v0 * x0[b];
      s1[b] += v1 * x1[b];
      s0[b] += v2 * x2[b];
      s1[b] += v3 * x3[b];
    }
  }
  for (; k < len; ++k) {
    const double v0 = __ldcs(vp + k * SELL_C);
    const double *p0 = gather_ptr<BS>(x, xg, nown, __ldcs(cp + k * SELL_C));
    double x0[BS];
    if constexpr (WIDE) {
      gather3_wide(p0, x0[0], x0[1], x0[2]);
    } else {
#pragma unroll
      for (int b = 0; b < BS; ++b) x0[b] = __ldg(p0 + b);
    }
#pragma unroll
    for (int b = 0; b < BS; ++b) s0[b] += v0 * x0[b];
  }
#pragma unroll
  for (int b = 0; b < BS; ++b) out[b] = s0[b] + s1[b];
}

// three consecutive doubles at p (coherent loads: the operand may be written by other rows
// of the same launch, e.g. z aliasing y); `wide` as in gather3_wide
__device__ __forceinline__ void load3(const double *p, bool wide, double (&o)[3]) {
  if (wide) {
    const bool odd = (reinterpret_cast<uintptr_t>(p) & 8) != 0;
    const double2 *q = reinterpret_cast<const double2 *>(p - (odd ? 1 : 0));
    const double2 lo = q[0], hi = q[1];
    o[0] = odd ? lo.y : lo.x;
    o[1] = odd ? hi.x : lo.y;
    o[2] = odd ? hi.y : hi.x;
  } else {
    o[0] = p[0];
    o[1] = p[1];
    o[2] = p[2];
  }
}
__device__ __forceinline__ void epilogue3(const EpiStore &e, int r, const double (&s)[3], bool wide) {
#pragma unroll
  for (int b = 0; b < 3; ++b) e.y[3 * r + b] = s[b];
  if (e.y2) {
    double d[3];
    load3(e.d2 + 3 * r, wide, d);
#pragma unroll
    for (int b = 0; b < 3; ++b) e.y2[3 * r + b] = e.s2 * d[b] * s[b];
  }
}
__device__ __forceinline__ void epilogue3(const EpiAxpby &e, int r, const double (&s)[3], bool wide) {
  double z[3], v[3];
  load3(e.z + 3 * r, wide, z);
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    v[b] = e.a * s[b] + e.b * z[b];
    e.y[3 * r + b] = v[b];
  }
  if (e.y2) {
    double d[3];
    load3(e.d2 + 3 * r, wide, d);
#pragma unroll
    for (int b = 0; b < 3; ++b) e.y2[3 * r + b] = e.s2 * d[b] * v[b];
  }
}
__device__ __forceinline__ void epilogue3(const EpiCheb &e, int r, const double (&s)[3], bool wide) {
  double p1[3], di[3], bb[3], v[3];
  load3(e.p1 + 3 * r, wide, p1);
  load3(e.dinv + 3 * r, wide, di);
  load3(e.b + 3 * r, wide, bb);
#pragma unroll
  for (int b = 0; b < 3; ++b) v[b] = e.c1 * p1[b] + e.c2 * di[b] * (bb[b] - s[b]);
  if (e.p0) {
    double t[3];
    load3(e.p0 + 3 * r, wide, t);
#pragma unroll
    for (int b = 0; b < 3; ++b) v[b] += e.c0 * t[b];
  }
  if (e.add) {
    double t[3];
    load3(e.add + 3 * r, wide, t);
#pragma unroll
    for (int b = 0; b < 3; ++b) v[b] += t[b];
  }
#pragma unroll
  for (int b = 0; b < 3; ++b) e.out[3 * r + b] = v[b];
}

__device__ __forceinline__ void sell_prefetch(const int32_t *__restrict__ col, const double *__restrict__ val, int base, int len) {
  // pull a slice's whole (col, val) stream into L2 with two bulk prefetches: the dependent
  // col -> gather chain then waits on L2, not on HBM
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(col + base), "r"(len * SELL_C * 4) : "memory");
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(val + base), "r"(len * SELL_C * 8) : "memory");
}

// VAR bits (option fnp_sell_gather): 1 16-byte gathers (BS == 3), 2 six CTAs per SM (40 registers),
// 4 L2 bulk prefetch of the slice stream, 8 16-byte loads in the epilogue (BS == 3), 64 default cache
// policy instead of evict-first for operators that fit in L2 (BS == 1).
// Measured on A00 = S (x) I_3 of the 64^3 cavity (profiles/r01_spmv_kernel_choice.md): 0.247 ms
// with none, 0.178 ms with the prefetch alone, 0.170 ms with all.  A persistent variant (warps
// striding over slices, next slice prefetched) was measured 2x slower and is not kept.
template <int BS, class Epi, int VAR>
__global__ void __launch_bounds__(256, (VAR & 2) ? 6 : 0)
spmv_sell_kernel(int nslices, const int32_t *__restrict__ sl_ptr, const int32_t *__restrict__ col,
                 const double *__restrict__ val, const int32_t *__restrict__ perm, const double *__restrict__ x,
                 const double *__restrict__ xg, int nown, const HaloWaitDev *__restrict__ hw, Epi epi) {
  if (hw) xg = halo_wait_dev(hw);
  const int slice = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (slice >= nslices) return;
  const int raw = __ldg(sl_ptr + slice);
  const int base = raw & ~(SELL_C - 1);
  const bool narrow = (raw & SELL_NARROW) != 0;     // slice touches an end of a vector: scalar loads
  const int len = ((__ldg(sl_ptr + slice + 1) & ~(SELL_C - 1)) - base) >> 5;
  if ((VAR & 4) && lane == 0 && len > 0) sell_prefetch(col, val, base, len);
  const int row = __ldg(perm + slice * SELL_C + lane);
  const int32_t *cp = col + base + lane;
  const double *vp = val + base + lane;
  if (BS == 1) {
    // VAR bit 64: operators that fit in L2 and are applied many times per PC apply (Ap, Mp, Kp and
    // their coarse levels) load their stream with the default cache policy instead of evict-first,
    // so that it stays resident between the applies (round 2, 64^3 cavity: Ap SpMV 11.0 -> 9.9 us)
    auto ldm_i = [](const int32_t *p) { return (VAR & 64) ? __ldg(p) : __ldcs(p); };
    auto ldm_d = [](const double *p) { return (VAR & 64) ? __ldg(p) : __ldcs(p); };
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int k = 0;
    for (; k + 4 <= len; k += 4) {
      const int c0 = ldm_i(cp + (k + 0) * SELL_C), c1 = ldm_i(cp + (k + 1) * SELL_C);
      const int c2 = ldm_i(cp + (k + 2) * SELL_C), c3 = ldm_i(cp + (k + 3) * SELL_C);
      const double v0 = ldm_d(vp + (k + 0) * SELL_C), v1 = ldm_d(vp + (k + 1) * SELL_C);
      const double v2 = ldm_d(vp + (k + 2) * SELL_C), v3 = ldm_d(vp + (k + 3) * SELL_C);
      s0 += v0 * __ldg(gather_ptr<1>(x, xg, nown, c0));
      s1 += v1 * __ldg(gather_ptr<1>(x, xg, nown, c1));
      s2 += v2 * __ldg(gather_ptr<1>(x, xg, nown, c2));
      s3 += v3 * __ldg(gather_ptr<1>(x, xg, nown, c3));
    }
    for (; k < len; ++k) s0 += ldm_d(vp + k * SELL_C) * __ldg(gather_ptr<1>(x, xg, nown, ldm_i(cp + k * SELL_C)));
    if (row >= 0) epilogue(epi, row, (s0 + s1) + (s2 + s3));
  } else {
    double s[BS];
    if (BS == 3 && (VAR & 1) && !narrow)
      sell_row_sums<BS, BS == 3>(cp, vp, len, x, xg, nown, s);
    else
      sell_row_sums<BS, false>(cp, vp, len, x, xg, nown, s);
    if (row >= 0) {
      if constexpr (BS == 3 && (VAR & 8) != 0) {
        epilogue3(epi, row, s, !narrow);
      } else {
#pragma unroll
        for (int b = 0; b < BS; ++b) epilogue(epi, BS * row + b, s[b]);
      }
    }
  }
}

// Multi-warp SELL kernel for operators with too few rows to fill the GPU with one thread per row
// (AMG levels 1.., the pressure operators of small meshes): T warps share a slice, warp t takes the
// entries k = t, t + T, ... of its 32 rows -- every load stays a coalesced 32-entry column of the
// slice, the layout is the one of the single-warp kernel -- and the T partial sums meet in shared
// memory (fixed order: deterministic).  Round-1 profile of the case it is for: level 1 of the
// velocity hierarchy, 207 k rows x 51 entries, one thread per row = 44 warps per SM each walking a
// 51-long dependent col -> x[col] chain, 0.21 of the HBM peak.
template <int BS, class Epi, int T, bool L2RES>
__global__ void __launch_bounds__(256)
spmv_sell_mw_kernel(int nslices, const int32_t *__restrict__ sl_ptr, const int32_t *__restrict__ col,
                    const double *__restrict__ val, const int32_t *__restrict__ perm, const double *__restrict__ x,
                    const double *__restrict__ xg, int nown, const HaloWaitDev *__restrict__ hw, Epi epi) {
  if (hw) xg = halo_wait_dev(hw);
  constexpr int SPB = 8 / T;                       // slices per block of 8 warps
  __shared__ double part[8][BS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = warp % T;
  const int slice = blockIdx.x * SPB + warp / T;
  const bool valid = slice < nslices;
  double s0[BS], s1[BS];
#pragma unroll
  for (int b = 0; b < BS; ++b) s0[b] = s1[b] = 0.0;
  bool narrow = false;
  if (valid) {
    const int raw = __ldg(sl_ptr + slice);
    const int base = raw & ~(SELL_C - 1);
    narrow = (raw & SELL_NARROW) != 0;
    const int len = ((__ldg(sl_ptr + slice + 1) & ~(SELL_C - 1)) - base) >> 5;
    const int32_t *cp = col + base + lane;
    const double *vp = val + base + lane;
    auto ldi = [](const int32_t *p) { return L2RES ? __ldg(p) : __ldcs(p); };
    auto ldd = [](const double *p) { return L2RES ? __ldg(p) : __ldcs(p); };
    const bool wide = BS == 3 && !narrow;
    int k = t;
    for (; k + T < len; k += 2 * T) {
      const int c0 = ldi(cp + k * SELL_C), c1 = ldi(cp + (k + T) * SELL_C);
      const double v0 = ldd(vp + k * SELL_C), v1 = ldd(vp + (k + T) * SELL_C);
      const double *p0 = gather_ptr<BS>(x, xg, nown, c0), *p1 = gather_ptr<BS>(x, xg, nown, c1);
      double x0[BS], x1[BS];
      if constexpr (BS == 3) {
        if (wide) {
          gather3_wide(p0, x0[0], x0[1], x0[2]);
          gather3_wide(p1, x1[0], x1[1], x1[2]);
        } else {
#pragma unroll
          for (int b = 0; b < BS; ++b) { x0[b] = __ldg(p0 + b); x1[b] = __ldg(p1 + b); }
        }
      } else {
#pragma unroll
        for (int b = 0; b < BS; ++b) { x0[b] = __ldg(p0 + b); x1[b] = __ldg(p1 + b); }
      }
#pragma unroll
      for (int b = 0; b < BS; ++b) {
        s0[b] += v0 * x0[b];
        s1[b] += v1 * x1[b];
      }
    }
    if (k < len) {
      const double v0 = ldd(vp + k * SELL_C);
      const double *p0 = gather_ptr<BS>(x, xg, nown, ldi(cp + k * SELL_C));
#pragma unroll
      for (int b = 0; b < BS; ++b) s0[b] += v0 * __ldg(p0 + b);
    }
  }
#pragma unroll
  for (int b = 0; b < BS; ++b) part[warp][b][lane] = s0[b] + s1[b];
  __syncthreads();
  if (!valid || t != 0) return;
  const int row = __ldg(perm + slice * SELL_C + lane);
  if (row < 0) return;
  double s[BS];
#pragma unroll
  for (int b = 0; b < BS; ++b) {
    double acc = part[warp][b][lane];
#pragma unroll
    for (int q = 1; q < T; ++q) acc += part[warp + q][b][lane];
    s[b] = acc;
  }
  if constexpr (BS == 3) {
    epilogue3(epi, row, s, !narrow);
  } else {
#pragma unroll
    for (int b = 0; b < BS; ++b) epilogue(epi, BS * row + b, s[b]);
  }
}

// warps per slice of the multi-warp kernel: enough threads to fill the GPU (~600 k), at least
// ~10 entries per lane.  Measured (round 2, 64^3 cavity): level 1 of the velocity hierarchy (207 k
// rows x 51) 99.7 -> 73.6 us and level 2 33.9 -> 14.0 us with 4 warps; two warps on the 15-entry rows of
// Ap / Mp were slower (15.3 -> 17.1 us), and so was the kernel on operators above a million rows.
static int pick_sell_warps(int64_t nrows, double mean_row, int64_t target_threads) {
  int t = 1;
  while (t < 8 && nrows * t < target_threads && mean_row / (2 * t) >= 10.0) t *= 2;
  return t;
}

static int pick_lanes(double mean_row) {
  int lanes = 2;
  while (lanes < 32 && lanes * 2 < mean_row + 0.5) lanes *= 2;
  return lanes;
}

// SELL-32-sigma layout of a subset of rows: permutation (padded with -1 to whole slices),
// slice pointers relative to `base`, and the entry count.
static void sell_layout(const HostCsr &h, const std::vector<int32_t> &rows, std::vector<int32_t> &perm,
                        std::vector<int32_t> &ptr, int64_t &total, int64_t SELL_SIGMA) {
  const int64_t n = (int64_t)rows.size();
  const int64_t nsl = (n + SELL_C - 1) / SELL_C;
  perm.assign((size_t)nsl * SELL_C, -1);
  const int64_t nwin = (n + SELL_SIGMA - 1) / SELL_SIGMA;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t w = 0; w < nwin; ++w) {
    const int64_t w0 = w * SELL_SIGMA, w1 = std::min<int64_t>(n, w0 + SELL_SIGMA);
    for (int64_t i = w0; i < w1; ++i) perm[i] = rows[i];
    std::stable_sort(perm.begin() + w0, perm.begin() + w1, [&](int32_t a, int32_t b) {
      return h.rowptr[a + 1] - h.rowptr[a] > h.rowptr[b + 1] - h.rowptr[b];
    });
  }
  ptr.assign(nsl + 1, 0);
  std::vector<int32_t> slen((size_t)nsl, 0);
#pragma omp parallel for schedule(static)
  for (int64_t s = 0; s < nsl; ++s) {
    int32_t len = 0;
    for (int l = 0; l < SELL_C; ++l) {
      const int32_t r = perm[s * SELL_C + l];
      if (r >= 0) len = std::max(len, h.rowptr[r + 1] - h.rowptr[r]);
    }
    slen[s] = len;
  }
  total = 0;
  for (int64_t s = 0; s < nsl; ++s) {
    total += (int64_t)slen[s] * SELL_C;
    FNP_REQUIRE(total < (int64_t)INT32_MAX, FNP_ERR_ARG, "SELL layout exceeds 2^31 entries on one rank");
    ptr[s + 1] = (int32_t)total;
  }
}

static void sell_fill(const HostCsr &h, const std::vector<int32_t> &perm, const std::vector<int32_t> &ptr, int64_t base,
                      std::vector<int32_t> &col, std::vector<int32_t> &pos) {
  const int64_t nsl = (int64_t)ptr.size() - 1;
#pragma omp parallel for schedule(static)
  for (int64_t s = 0; s < nsl; ++s) {
    const int64_t sbase = base + ptr[s];
    const int32_t len = (ptr[s + 1] - ptr[s]) / SELL_C;
    for (int l = 0; l < SELL_C; ++l) {
      const int32_t r = perm[s * SELL_C + l];
      const int32_t b = r >= 0 ? h.rowptr[r] : 0, e = r >= 0 ? h.rowptr[r + 1] : 0;
      const int32_t fill = e > b ? h.col[b] : 0;       // padding: zero value, harmless in-range column
      for (int32_t k = 0; k < len; ++k) {
        const int64_t dst = sbase + (int64_t)k * SELL_C + l;
        if (b + k < e) {
          col[dst] = h.col[b + k];
          pos[b + k] = (int32_t)dst;
        } else {
          col[dst] = fill;
        }
      }
    }
  }
}

// Build the SELL copy of a pattern.  With n_own_split >= 0 (multi-rank) the rows are split
// into an interior part (no ghost column) and a boundary part, so that the interior part
// can run while the halo exchange is in flight.
static void build_sell(Ctx &c, DevCsr &A, const HostCsr &h, int64_t n_own_split, int64_t n_own_cols) {
  std::vector<int32_t> rows_a, rows_b;
  rows_a.reserve(h.nrows);
  {
    std::vector<char> ghost((size_t)h.nrows, 0);
    if (n_own_split >= 0) {
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < h.nrows; ++i)
        for (int32_t k = h.rowptr[i]; k < h.rowptr[i + 1]; ++k)
          if (h.col[k] >= n_own_split) { ghost[i] = 1; break; }
    }
    for (int64_t i = 0; i < h.nrows; ++i) (ghost[i] ? rows_b : rows_a).push_back((int32_t)i);
  }
  std::vector<int32_t> perm_a, ptr_a, perm_b, ptr_b;
  int64_t tot_a = 0, tot_b = 0;
  sell_layout(h, rows_a, perm_a, ptr_a, tot_a, c.sell_sigma);
  sell_layout(h, rows_b, perm_b, ptr_b, tot_b, c.sell_sigma);
  FNP_REQUIRE(tot_a + tot_b < (int64_t)INT32_MAX, FNP_ERR_ARG, "SELL layout exceeds 2^31 entries on one rank");
  std::vector<int32_t> col((size_t)(tot_a + tot_b));
  A.sell_pos.assign((size_t)h.nnz(), 0);
  sell_fill(h, perm_a, ptr_a, 0, col, A.sell_pos);
  sell_fill(h, perm_b, ptr_b, tot_a, col, A.sell_pos);
  // slices whose entries (padding included) address the first or last node of the owned
  // vector or of the ghost buffer: the 16-byte gathers of the Kronecker kernel would read
  // outside, flag them for the scalar loads
  {
    const int32_t nown = n_own_cols >= 0 ? (int32_t)n_own_cols : (int32_t)h.ncols;
    const int32_t edge[4] = {0, nown - 1, nown, (int32_t)h.ncols - 1};
    auto flag = [&](std::vector<int32_t> &ptr, const std::vector<int32_t> &perm, int64_t off) {
#pragma omp parallel for schedule(static)
      for (int64_t sl = 0; sl < (int64_t)ptr.size() - 1; ++sl) {
        bool hit = false;
        for (int l = 0; l < SELL_C; ++l)     // first / last row: the epilogue's 16-byte loads
          hit = hit || perm[sl * SELL_C + l] == 0 || perm[sl * SELL_C + l] == (int32_t)h.nrows - 1;
        for (int64_t k = off + ptr[sl]; k < off + ptr[sl + 1] && !hit; ++k)
          hit = col[k] == edge[0] || col[k] == edge[1] || col[k] == edge[2] || col[k] == edge[3];
        if (hit) ptr[sl] |= SELL_NARROW;
      }
    };
    flag(ptr_a, perm_a, 0);
    flag(ptr_b, perm_b, tot_a);
  }
  A.nslices = (int32_t)ptr_a.size() - 1;
  A.nslices_b = (int32_t)ptr_b.size() - 1;
  A.sell_entries = tot_a + tot_b;
  A.sell_entries_a = tot_a;
  A.sl_ptr.upload(ptr_a.data(), ptr_a.size(), c.stream);
  A.sl_perm.upload(perm_a.data(), perm_a.size(), c.stream);
  A.sl_ptr_b.upload(ptr_b.data(), ptr_b.size(), c.stream);
  A.sl_perm_b.upload(perm_b.data(), perm_b.size(), c.stream);
  A.sl_col.upload(col.data(), col.size(), c.stream);
  A.sl_val.alloc((size_t)(tot_a + tot_b));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  A.sell = true;
}

void csr_upload_pattern(Ctx &c, DevCsr &A, const HostCsr &h, const std::string &tag, int64_t n_own_split, int bs,
                        int64_t n_own_cols) {
  A.tag = tag;
  A.bs = bs;
  A.nrows = (int32_t)h.nrows;
  A.ncols_own = (int32_t)h.ncols;
  A.nghost = 0;
  A.nnz = h.nnz();
  A.mean_row = h.nrows ? (double)A.nnz / (double)h.nrows : 0.0;
  double maxrow = 0;
  for (int64_t i = 0; i < h.nrows; ++i) maxrow = std::max(maxrow, (double)(h.rowptr[i + 1] - h.rowptr[i]));
  A.max_row = maxrow;
  A.lanes = pick_lanes(A.mean_row);
  A.sell_warps = c.sell_warps > 0 ? c.sell_warps : pick_sell_warps(h.nrows, A.mean_row, c.sell_warps_rows);
  A.has_dinv = false;
  A.sell = false;
  A.rowptr.upload(h.rowptr.data(), h.rowptr.size(), c.stream);
  // format choice from the row-length histogram: thread-per-row SELL needs many short
  // rows; long rows (dense coarse levels, the divergence block) keep CSR + one
  // sub-warp per row, which already has enough loads in flight per row
  const bool short_rows = A.mean_row < c.sell_max_mean_row && maxrow <= 8.0 * std::max(8.0, A.mean_row) && h.nrows >= 4096;
  const bool use_sell = c.spmv_mode == 2 || (c.spmv_mode == 0 && short_rows);
  if (use_sell) {
    build_sell(c, A, h, n_own_split, n_own_cols);
    A.col.release();
    A.val.release();
  } else {
    A.col.upload(h.col.data(), h.col.size(), c.stream);
    A.val.alloc((size_t)A.nnz);
    A.sell_pos.clear();
    A.sell_pos.shrink_to_fit();
  }
  FNP_CUDA(cudaStreamSynchronize(c.stream));
}

void csr_set_values(Ctx &c, DevCsr &A, const HostCsr &h, const double *val, bool want_dinv) {
  const int64_t nnz = h.nnz();
  if (A.sell) {
    std::vector<double> tmp((size_t)A.sell_entries, 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nnz; ++k) tmp[A.sell_pos[k]] = val[k];
    A.sl_val.upload(tmp.data(), tmp.size(), c.stream);
    FNP_CUDA(cudaStreamSynchronize(c.stream));       // tmp is pageable and goes out of scope
  } else {
    A.val.upload(val, (size_t)nnz, c.stream);
  }
  if (want_dinv) {
    const int bs = A.bs;
    std::vector<double> dinv((size_t)h.nrows * bs, 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < h.nrows; ++i) {
      double d = 0.0;
      for (int32_t k = h.rowptr[i]; k < h.rowptr[i + 1]; ++k)
        if (h.col[k] == i) d = val[k];
      for (int b = 0; b < bs; ++b) dinv[(size_t)i * bs + b] = d != 0.0 ? 1.0 / d : 0.0;
    }
    A.dinv.upload(dinv.data(), dinv.size(), c.stream);
    A.has_dinv = true;
  }
  FNP_CUDA(cudaStreamSynchronize(c.stream));
}

// vectors of the fused epilogue that move besides y (written once) and x (read once), which
// DevCsr::spmv_bytes counts: the algorithmic traffic of the fused operation as implemented
static inline int epi_extra_vectors(const EpiStore &e) { return e.y2 ? 2 : 0; }
static inline int epi_extra_vectors(const EpiAxpby &e) { return 1 + (e.y2 ? 2 : 0); }
static inline int epi_extra_vectors(const EpiCheb &e) { return 2 + (e.p0 ? 1 : 0) + (e.add ? 1 : 0); }

template <int BS, class Epi>
static void spmv_launch_bs(Ctx &c, const DevCsr &A, const double *x, const Epi &epi) {
  const double *xg = nullptr;
  int nown = INT32_MAX;
  const int threads = 256;
  // peer-memory exchange of a split operator, option fnp_halo_p2p 2: the pack + remote-store kernel runs
  // on a forked stream beside the interior rows (fork / join by events: capturable)
  const bool fork_send = A.halo && A.halo->p2p && c.p2p >= 2 && A.sell && A.nslices_b > 0 && c.comm_stream && c.ev_x;
  if (A.halo) {
    // post the exchange of the ghost entries of x (peer-memory stores or NCCL send/recv)
    if (fork_send) {
      FNP_CUDA(cudaEventRecord(c.ev_x, c.stream));
      FNP_CUDA(cudaStreamWaitEvent(c.comm_stream, c.ev_x, 0));
      halo_exchange(c, *A.halo, x, c.comm_stream, c.comm);
      FNP_CUDA(cudaEventRecord(c.ev_halo, c.comm_stream));
    } else {
      halo_exchange(c, *A.halo, x, c.stream, c.comm);
    }
    xg = A.halo->current_ghost;
    nown = A.ncols_own;
  }
  // peer-memory exchange: the kernel that reads ghosts waits for the neighbours' flags itself and
  // derives the ghost slot from the device-resident sequence number (graph replayable)
  const HaloWaitDev *hw_ghost = (A.halo && A.halo->p2p && A.halo->nghost > 0) ? A.halo->d_wait.p : nullptr;
  StageTimer kt(c, "spmv " + A.tag, 2, A.spmv_bytes() + 8.0 * (double)A.vec_rows() * epi_extra_vectors(epi));
  if (A.sell) {
    auto launch = [&](int nsl, const int32_t *ptr, const int32_t *perm, int64_t off, const HaloWaitDev *hw) {
      if (nsl <= 0) return;
      if (A.sell_warps > 1) {
        const int T = A.sell_warps;
        const int gridw = (nsl + 8 / T - 1) / (8 / T);
        const bool l2res = (c.sell_gather & 64) && A.spmv_bytes() <= 64e6;
#define FNP_SELL_MW(TT)                                                                                                          \
  do {                                                                                                                       \
    if (l2res)                                                                                                               \
      spmv_sell_mw_kernel<BS, Epi, TT, true><<<gridw, 256, 0, c.stream>>>(nsl, ptr, A.sl_col.p + off, A.sl_val.p + off, perm, x, xg, nown, hw, epi);  \
    else                                                                                                                     \
      spmv_sell_mw_kernel<BS, Epi, TT, false><<<gridw, 256, 0, c.stream>>>(nsl, ptr, A.sl_col.p + off, A.sl_val.p + off, perm, x, xg, nown, hw, epi); \
  } while (0)
        if (T == 2) FNP_SELL_MW(2);
        else if (T == 4) FNP_SELL_MW(4);
        else FNP_SELL_MW(8);
#undef FNP_SELL_MW
        FNP_LAUNCH_CHECK(c);
        return;
      }
      const int grid = (int)(((int64_t)nsl * 32 + threads - 1) / threads);
#define FNP_SELL(V) \
  spmv_sell_kernel<BS, Epi, V><<<grid, threads, 0, c.stream>>>(nsl, ptr, A.sl_col.p + off, A.sl_val.p + off, perm, x, xg, nown, hw, epi)
      if (BS == 3) {
        switch (c.sell_gather & 15) {
          case 0: FNP_SELL(0); break;
          case 4: FNP_SELL(4); break;
          case 7: FNP_SELL(7); break;
          default: FNP_SELL(15); break;
        }
      } else {
        // bit 64: L2-resident policy for operators of at most 64 MB (Ap, Mp, Kp, coarse levels)
        if ((c.sell_gather & 64) && BS == 1 && A.spmv_bytes() <= 64e6) FNP_SELL(68);
        else if (c.sell_gather & 4) FNP_SELL(4);
        else FNP_SELL(0);
      }
#undef FNP_SELL
      FNP_LAUNCH_CHECK(c);
    };
    if (A.halo && A.nslices_b > 0) {
      // split operator: the interior rows (no ghost column) run while the ghost entries are in
      // flight; the boundary rows wait for the flags inside their kernel
      launch(A.nslices, A.sl_ptr.p, A.sl_perm.p, 0, nullptr);
      if (fork_send) FNP_CUDA(cudaStreamWaitEvent(c.stream, c.ev_halo, 0));     // join: x may be overwritten afterwards
      launch(A.nslices_b, A.sl_ptr_b.p, A.sl_perm_b.p, A.sell_entries_a, hw_ghost);
    } else {
      launch(A.nslices, A.sl_ptr.p, A.sl_perm.p, 0, hw_ghost);
      launch(A.nslices_b, A.sl_ptr_b.p, A.sl_perm_b.p, A.sell_entries_a, hw_ghost);
    }
    return;
  }
  const HaloWaitDev *hwc = hw_ghost;
  auto grid = [&](int lanes) { return (int)(((int64_t)A.nrows * lanes + threads - 1) / threads); };
#define FNP_VEC(L)                                                                                                        \
  do {                                                                                                                \
    if (c.sell_gather & 16)                                                                                           \
      spmv_kernel<L, BS, Epi, true><<<grid(L), threads, 0, c.stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, x, xg, nown, hwc, epi);  \
    else                                                                                                              \
      spmv_kernel<L, BS, Epi, false><<<grid(L), threads, 0, c.stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, x, xg, nown, hwc, epi); \
  } while (0)
  switch (A.lanes) {
    case 2: FNP_VEC(2); break;
    case 4: FNP_VEC(4); break;
    case 8: FNP_VEC(8); break;
    case 16: FNP_VEC(16); break;
    default: FNP_VEC(32); break;
  }
#undef FNP_VEC
  FNP_LAUNCH_CHECK(c);
}

template <class Epi>
static void spmv_launch(Ctx &c, const DevCsr &A, const double *x, const Epi &epi) {
  if (A.nrows == 0) {
    // a rank without rows of this operator still owns entries of x that its neighbours need: the
    // exchange is collective (found with a partition that gave one rank no pressure dofs at all --
    // its neighbour waited forever for the velocity ghosts of the divergence block)
    if (A.halo) halo_exchange(c, *A.halo, x, c.stream, c.comm);
    return;
  }
  switch (A.bs) {
    case 1: spmv_launch_bs<1, Epi>(c, A, x, epi); break;
    case 2: spmv_launch_bs<2, Epi>(c, A, x, epi); break;
    case 3: spmv_launch_bs<3, Epi>(c, A, x, epi); break;
    default: throw Error(FNP_ERR_ARG, "unsupported block size");
  }
}

void spmv_store(Ctx &c, const DevCsr &A, const double *x, double *y, const Jacobi1 &j) {
  spmv_launch(c, A, x, EpiStore{y, j.y2, j.d2, j.s2});
}
void spmv_axpby(Ctx &c, const DevCsr &A, const double *x, double a, double b, const double *z, double *y, const Jacobi1 &j) {
  spmv_launch(c, A, x, EpiAxpby{y, z, a, b, j.y2, j.d2, j.s2});
}
void spmv_cheb(Ctx &c, const DevCsr &A, const EpiCheb &e) { spmv_launch(c, A, e.p1, e); }

// ---------------------------------------------------------------------------
// BLAS-1 class
// ---------------------------------------------------------------------------
#define GRID_STRIDE(i, n) \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

__global__ void copy_kernel(int64_t n, const double *__restrict__ x, double *__restrict__ y) { GRID_STRIDE(i, n) y[i] = x[i]; }
__global__ void scale_kernel(int64_t n, double a, double *x) { GRID_STRIDE(i, n) x[i] *= a; }
__global__ void axpy_kernel(int64_t n, double a, const double *__restrict__ x, double *y) { GRID_STRIDE(i, n) y[i] += a * x[i]; }
__global__ void axpby_kernel(int64_t n, double a, const double *x, double b, const double *y, double *out) {
  GRID_STRIDE(i, n) out[i] = a * x[i] + b * y[i];
}
__global__ void zero_kernel(int64_t n, double *x) { GRID_STRIDE(i, n) x[i] = 0.0; }
__global__ void pw_scale_kernel(int64_t n, double s, const double *__restrict__ d, const double *__restrict__ b,
                                const double *add, double *out) {
  GRID_STRIDE(i, n) {
    double v = s * d[i] * b[i];
    if (add) v += add[i];
    out[i] = v;
  }
}
__global__ void scatter_bc_kernel(double *z, const int32_t *__restrict__ idx, const double *__restrict__ val, int32_t n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) z[idx[i]] = val[i];
}
__global__ void gather_kernel(int64_t n, const int64_t *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst) {
  GRID_STRIDE(i, n) dst[i] = src[idx[i]];
}
__global__ void scatter_kernel(int64_t n, const int64_t *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst) {
  GRID_STRIDE(i, n) dst[idx[i]] = src[i];
}

#define VEC_LAUNCH(kern, n, ...)                                            \
  do {                                                                      \
    if ((n) > 0) {                                                          \
      kern<<<blocks_for(c, (n), 256, 16), 256, 0, c.stream>>>(__VA_ARGS__); \
      FNP_LAUNCH_CHECK(c);                                                  \
    }                                                                       \
  } while (0)

void vec_copy(Ctx &c, int64_t n, const double *x, double *y) { VEC_LAUNCH(copy_kernel, n, n, x, y); }
void vec_scale(Ctx &c, int64_t n, double a, double *x) { VEC_LAUNCH(scale_kernel, n, n, a, x); }
void vec_axpy(Ctx &c, int64_t n, double a, const double *x, double *y) { VEC_LAUNCH(axpy_kernel, n, n, a, x, y); }
void vec_axpby(Ctx &c, int64_t n, double a, const double *x, double b, const double *y, double *out) {
  VEC_LAUNCH(axpby_kernel, n, n, a, x, b, y, out);
}
void vec_zero(Ctx &c, int64_t n, double *x) { VEC_LAUNCH(zero_kernel, n, n, x); }
void vec_pointwise_scale(Ctx &c, int64_t n, double s, const double *d, const double *b, const double *add, double *out) {
  VEC_LAUNCH(pw_scale_kernel, n, n, s, d, b, add, out);
}
void vec_scatter_bc(Ctx &c, double *z, const int32_t *idx, const double *val, int32_t nbc) {
  if (nbc > 0) {
    scatter_bc_kernel<<<(nbc + 255) / 256, 256, 0, c.stream>>>(z, idx, val, nbc);
    FNP_LAUNCH_CHECK(c);
  }
}
void vec_copy_bc(Ctx &c, int64_t n, const double *x, double *z, const int32_t *idx, const double *val, int32_t nbc) {
  vec_copy(c, n, x, z);
  vec_scatter_bc(c, z, idx, val, nbc);
}
void vec_gather(Ctx &c, int64_t n, const int64_t *idx, const double *src, double *dst) { VEC_LAUNCH(gather_kernel, n, n, idx, src, dst); }
void vec_scatter(Ctx &c, int64_t n, const int64_t *idx, const double *src, double *dst) { VEC_LAUNCH(scatter_kernel, n, n, idx, src, dst); }

// ---------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------
constexpr int RED_THREADS = 256;
constexpr int DOT_BATCH = 8;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// stage 1: partial[(i) * gridDim.x + blockIdx.x] = sum over this block's elements of V_i[e] * w[e]
__global__ void __launch_bounds__(RED_THREADS)
multidot_kernel(int64_t n, const double *const *__restrict__ V, int nvec, const double *__restrict__ w,
                double *__restrict__ partial, int nbasis) {
  // slots [0, nbasis) are basis vectors, slot nbasis (when nvec = nbasis + 1) is w itself: w . w
  const int i0 = blockIdx.y * DOT_BATCH;
  const double *v[DOT_BATCH];
#pragma unroll
  for (int b = 0; b < DOT_BATCH; ++b) v[b] = (i0 + b < nbasis) ? V[i0 + b] : (nvec > nbasis ? w : V[nbasis - 1]);
  double acc[DOT_BATCH];
#pragma unroll
  for (int b = 0; b < DOT_BATCH; ++b) acc[b] = 0.0;
  GRID_STRIDE(e, n) {
    const double wv = w[e];
#pragma unroll
    for (int b = 0; b < DOT_BATCH; ++b) acc[b] += __ldg(v[b] + e) * wv;
  }
  __shared__ double sm[DOT_BATCH][RED_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int b = 0; b < DOT_BATCH; ++b) {
    double s = warp_sum(acc[b]);
    if (lane == 0) sm[b][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < DOT_BATCH) {
    const int b = threadIdx.x;
    if (i0 + b < nvec) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < RED_THREADS / 32; ++k) s += sm[b][k];
      partial[(int64_t)(i0 + b) * gridDim.x + blockIdx.x] = s;
    }
  }
}

// stage 2: out[i] = sum_k partial[i * nblk + k], fixed order
__global__ void __launch_bounds__(128)
reduce_partials_kernel(const double *__restrict__ partial, int nblk, double *__restrict__ out) {
  const int i = blockIdx.x;
  double s = 0.0;
  for (int k = threadIdx.x; k < nblk; k += blockDim.x) s += partial[(int64_t)i * nblk + k];
  __shared__ double sm[4];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) out[i] = sm[0] + sm[1] + sm[2] + sm[3];
}

static int red_blocks(const Ctx &c, int64_t n) { return blocks_for(c, n, RED_THREADS, 4); }

void multi_dot_ptrs(Ctx &c, int64_t n, const double *const *Vptrs_dev, int nvec, const double *w, double *h_dev) {
  if (nvec <= 0) return;
  const int nblk = red_blocks(c, n);
  // every batch of DOT_BATCH basis vectors re-reads w (SURVEY 8d counts the one-pass ideal 8N(j+2))
  StageTimer kt(c, "multidot", 2, 8.0 * n * (nvec + (nvec + DOT_BATCH - 1) / DOT_BATCH));
  c.red_partial.ensure((size_t)nblk * (nvec + DOT_BATCH));
  dim3 grid(nblk, (nvec + DOT_BATCH - 1) / DOT_BATCH);
  multidot_kernel<<<grid, RED_THREADS, 0, c.stream>>>(n, Vptrs_dev, nvec, w, c.red_partial.p, nvec);
  FNP_LAUNCH_CHECK(c);
  reduce_partials_kernel<<<nvec, 128, 0, c.stream>>>(c.red_partial.p, nblk, h_dev);
  FNP_LAUNCH_CHECK(c);
}

void multi_dot_ww_ptrs(Ctx &c, int64_t n, const double *const *Vptrs_dev, int nbasis, const double *w, double *h_dev) {
  const int nvec = nbasis + 1;
  const int nblk = red_blocks(c, n);
  StageTimer kt(c, "multidot", 2, 8.0 * n * (nvec + (nvec + DOT_BATCH - 1) / DOT_BATCH));
  FNP_REQUIRE(c.red_partial.n >= (size_t)nblk * (nvec + DOT_BATCH), FNP_ERR_STATE, "reduction work space too small");
  dim3 grid(nblk, (nvec + DOT_BATCH - 1) / DOT_BATCH);
  multidot_kernel<<<grid, RED_THREADS, 0, c.stream>>>(n, Vptrs_dev, nvec, w, c.red_partial.p, nbasis);
  FNP_LAUNCH_CHECK(c);
  reduce_partials_kernel<<<nvec, 128, 0, c.stream>>>(c.red_partial.p, nblk, h_dev);
  FNP_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(RED_THREADS)
dot_kernel(int64_t n, const double *__restrict__ x, const double *__restrict__ y, double *__restrict__ partial) {
  double acc = 0.0;
  GRID_STRIDE(e, n) acc += x[e] * y[e];
  __shared__ double sm[RED_THREADS / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < RED_THREADS / 32; ++k) s += sm[k];
    partial[blockIdx.x] = s;
  }
}

void dot(Ctx &c, int64_t n, const double *x, const double *y, double *out_dev) {
  const int nblk = red_blocks(c, n);
  c.red_partial.ensure((size_t)nblk);
  dot_kernel<<<nblk, RED_THREADS, 0, c.stream>>>(n, x, y, c.red_partial.p);
  FNP_LAUNCH_CHECK(c);
  reduce_partials_kernel<<<1, 128, 0, c.stream>>>(c.red_partial.p, nblk, out_dev);
  FNP_LAUNCH_CHECK(c);
}

// w -= sum_i h[i] V_i, partial = block sums of w_new^2.  Two consecutive elements per thread with
// 16-byte loads, four basis vectors in flight: 128 bytes of loads outstanding per thread (round 2:
// the 8-byte version reached 0.59 of the HBM peak at 53 M dofs -- too few bytes in flight).
__global__ void __launch_bounds__(RED_THREADS)
maxpy_norm_kernel(int64_t n, const double *const *__restrict__ V, int nvec, const double *__restrict__ h,
                  double *__restrict__ w, double *__restrict__ partial) {
  extern __shared__ double sh[];          // nvec coefficients
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) sh[i] = h[i];
  __syncthreads();
  double acc = 0.0;
  const int64_t n2 = n >> 1;
  double2 *w2 = reinterpret_cast<double2 *>(w);
  GRID_STRIDE(e, n2) {
    double2 wv = w2[e];
    int i = 0;
    for (; i + 4 <= nvec; i += 4) {
      const double2 a0 = __ldg(reinterpret_cast<const double2 *>(V[i]) + e);
      const double2 a1 = __ldg(reinterpret_cast<const double2 *>(V[i + 1]) + e);
      const double2 a2 = __ldg(reinterpret_cast<const double2 *>(V[i + 2]) + e);
      const double2 a3 = __ldg(reinterpret_cast<const double2 *>(V[i + 3]) + e);
      wv.x -= sh[i] * a0.x;     wv.y -= sh[i] * a0.y;
      wv.x -= sh[i + 1] * a1.x; wv.y -= sh[i + 1] * a1.y;
      wv.x -= sh[i + 2] * a2.x; wv.y -= sh[i + 2] * a2.y;
      wv.x -= sh[i + 3] * a3.x; wv.y -= sh[i + 3] * a3.y;
    }
    for (; i < nvec; ++i) {
      const double2 a0 = __ldg(reinterpret_cast<const double2 *>(V[i]) + e);
      wv.x -= sh[i] * a0.x;
      wv.y -= sh[i] * a0.y;
    }
    w2[e] = wv;
    acc += wv.x * wv.x + wv.y * wv.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {      // odd length: the last element
    const int64_t e = n - 1;
    double wv = w[e];
    for (int i = 0; i < nvec; ++i) wv -= sh[i] * __ldg(V[i] + e);
    w[e] = wv;
    acc += wv * wv;
  }
  __shared__ double sm[RED_THREADS / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < RED_THREADS / 32; ++k) s += sm[k];
    partial[blockIdx.x] = s;
  }
}

void multi_axpy_norm_ptrs(Ctx &c, int64_t n, const double *const *Vptrs_dev, int nvec, const double *h_dev, double *w,
                          double *nrm2_dev) {
  const int nblk = red_blocks(c, n);
  StageTimer kt(c, "maxpy+norm", 2, 8.0 * n * (nvec + 2));
  c.red_partial.ensure((size_t)nblk);
  // + 16 bytes: the compiler pairs the coefficient loads (16-byte shared loads) and may touch one double past the last
  maxpy_norm_kernel<<<nblk, RED_THREADS, (nvec + 2) * sizeof(double), c.stream>>>(n, Vptrs_dev, nvec, h_dev, w, c.red_partial.p);
  FNP_LAUNCH_CHECK(c);
  reduce_partials_kernel<<<1, 128, 0, c.stream>>>(c.red_partial.p, nblk, nrm2_dev);
  FNP_LAUNCH_CHECK(c);
}

// x += sum_i y[i] Z_i
__global__ void __launch_bounds__(RED_THREADS)
maxpy_kernel(int64_t n, const double *const *__restrict__ Z, int nvec, const double *__restrict__ y, double *__restrict__ x) {
  extern __shared__ double sh[];
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) sh[i] = y[i];
  __syncthreads();
  GRID_STRIDE(e, n) {
    double xv = x[e];
    for (int i = 0; i < nvec; ++i) xv += sh[i] * __ldg(Z[i] + e);
    x[e] = xv;
  }
}

void multi_axpy_ptrs(Ctx &c, int64_t n, const double *const *Zptrs_dev, int nvec, const double *y_dev, double *x) {
  if (nvec <= 0 || n <= 0) return;
  maxpy_kernel<<<red_blocks(c, n), RED_THREADS, (nvec + 2) * sizeof(double), c.stream>>>(n, Zptrs_dev, nvec, y_dev, x);
  FNP_LAUNCH_CHECK(c);
}

__global__ void scale_inv_sqrt_kernel(int64_t n, const double *__restrict__ nrm2, const double *__restrict__ w, double *__restrict__ v) {
  const double s = nrm2[0] > 0.0 ? 1.0 / sqrt(nrm2[0]) : 0.0;
  GRID_STRIDE(i, n) v[i] = w[i] * s;
}

void vec_scale_inv_sqrt(Ctx &c, int64_t n, const double *nrm2_dev, const double *w, double *v) {
  VEC_LAUNCH(scale_inv_sqrt_kernel, n, n, nrm2_dev, w, v);
}

// x = (Minv (x) I_BS) b, Minv dense row-major nrows x n (scalar operator): one warp per scalar row,
// the row of the inverse is read once for the BS interleaved components
template <int BS>
__global__ void __launch_bounds__(256)
gemv_kernel(int nrows, int n, const double *__restrict__ M, const double *__restrict__ b, double *__restrict__ x) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  double s[BS];
#pragma unroll
  for (int c = 0; c < BS; ++c) s[c] = 0.0;
  for (int k = lane; k < n; k += 32) {
    const double m = M[(int64_t)row * n + k];
#pragma unroll
    for (int c = 0; c < BS; ++c) s[c] += m * __ldg(b + (int64_t)BS * k + c);
  }
#pragma unroll
  for (int c = 0; c < BS; ++c) {
    const double v = warp_sum(s[c]);
    if (lane == 0) x[(int64_t)BS * row + c] = v;
  }
}

void dense_gemv(Ctx &c, int nrows, int ncols, int bs, const double *Minv, const double *b, double *x) {
  if (nrows <= 0) return;
  const int grid = (nrows * 32 + 255) / 256;
  StageTimer kt(c, "gemv coarse", 2, 8.0 * nrows * (double)ncols + 8.0 * bs * (nrows + ncols));
  switch (bs) {
    case 1: gemv_kernel<1><<<grid, 256, 0, c.stream>>>(nrows, ncols, Minv, b, x); break;
    case 2: gemv_kernel<2><<<grid, 256, 0, c.stream>>>(nrows, ncols, Minv, b, x); break;
    case 3: gemv_kernel<3><<<grid, 256, 0, c.stream>>>(nrows, ncols, Minv, b, x); break;
    default: throw Error(FNP_ERR_ARG, "unsupported block size");
  }
  FNP_LAUNCH_CHECK(c);
}

}  // namespace fnp
