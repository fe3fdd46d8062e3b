// rkb_nearest.cu — batched nearest-neighbour queries over a set of state-space points.
//
// What the planners ask right before they steer: the vertices of the motion graph nearest to a sample, under the
// topology's metric — ReaK::pp::linear_neighbor_search / dvp_tree (ctrl/path_planning/topological_search.hpp:91-112,
// 238-270, 586-596; metric_space_search.hpp; dvp_tree_detail.hpp) with distance(a, b) = norm_2(difference(b, a)) on
// vect_n points (core/lin_alg/vect_alg.hpp:2314-2333: sum += v[i] * v[i] in index order from 0.0, then sqrt).  A
// vantage-point tree returns the same neighbours as the linear scan (it is an exact search); here the scan itself is
// done for a whole batch of queries at once.
//
// Exactness: every vertex that can possibly enter a query's list (a fused-multiply-add scan with a proven error margin
// finds them) has its squared distance accumulated with separately rounded multiply and add (__dmul_rn / __dadd_rn) in
// the reference's order, and is compared on d = sqrt(s) like the reference compares, so distances are bit-identical to
// the CPU scan and so are the chosen vertices; among equal distances the lowest vertex index wins (the first one met by
// min_dist_linear_search).  The sqrt is only taken for the rare candidate that beats the current k-th squared distance
// (sqrt is monotonic: s >= s_k implies d >= d_k).
//
// Mapping: one thread per query, its coordinates in registers (dimension padded to a multiple of 4 with zeros, which add
// +0.0 to a non-negative sum: exact); vertices stream through shared memory in tiles and are read by broadcast, four
// vertices per thread in flight.  When the queries alone cannot fill the GPU the vertex set is split over blockIdx.y and
// the per-chunk lists are merged by a second kernel.  Bound: the FP64 pipe (2 instructions per coordinate and pair).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdint.h>

#include "../../include/reak_b200.h"
#include "rkb_internal.h"

namespace {

constexpr int NN_BLOCK = 128;   // queries per CTA
constexpr int NN_TILE = 64;     // vertices per shared-memory tile
constexpr int NN_MAX_K = RKB_NEAREST_MAX_K;

struct NearestArgs {
  const double* vertices;  // [V][dim]
  const double* queries;   // [Q][dim]
  long long n_vertices, n_queries;
  int dim, k;
  double radius;           // candidates need d < radius
  long long chunk;         // vertices per blockIdx.y
  int32_t* part_idx;       // [Q][n_chunks][k]
  double*  part_dist;      // [Q][n_chunks][k]
  int n_chunks;
};

// sorted insertion of (d, s, idx) into the thread's list of at most k entries (ascending d; a new entry goes behind
// entries of equal d, which have lower indices)
__device__ __forceinline__ void nn_insert(double* dl, double* sl, int32_t* il, int& cnt, int k, double d, double s, int32_t idx) {
  int pos = cnt < k ? cnt : k - 1;
  while (pos > 0 && dl[pos - 1] > d) {
    dl[pos] = dl[pos - 1]; sl[pos] = sl[pos - 1]; il[pos] = il[pos - 1];
    --pos;
  }
  dl[pos] = d; sl[pos] = s; il[pos] = idx;
  if (cnt < k) ++cnt;
}

// the thread's sorted list of at most k candidates for one query, and the gates that spare the sqrt
struct NnList {
  double dl[NN_MAX_K], sl[NN_MAX_K];
  int32_t il[NN_MAX_K];
  int cnt;
  double gate_s, gate_d;
};

__device__ __forceinline__ void nn_offer(NnList& L, int k, double s, int32_t idx) {
  if (s < L.gate_s) {  // rare
    const double d = sqrt(s);
    if (d < L.gate_d) {
      nn_insert(L.dl, L.sl, L.il, L.cnt, k, d, s, idx);
      if (L.cnt == k) { L.gate_d = L.dl[k - 1]; L.gate_s = L.sl[k - 1]; }
    }
  }
}

// RQ queries per thread: a vertex coordinate fetched from shared memory (a broadcast read, but still 16 bytes written
// to every lane's registers — the LSU's return path, not the FP64 pipe, bounds the kernel at RQ = 1) feeds 2 RQ FP64
// instructions.
//
// Two-stage test.  The scan itself forms s~ = sum fma(d, d, .) (2 instructions per coordinate); the reference's sum s
// (separately rounded multiply and add, 3 instructions) differs from it by at most 2 (dim + 2) ulps relative — both
// approximate the same sum of non-negative terms — so only a vertex with s~ below the gate widened by that bound can
// beat the current k-th neighbour.  For those (a handful per query) the exact s is recomputed from the tile in the
// reference's arithmetic and offered to the list: the result is bit-identical, the scan 1.5x shorter.
// the rare path, out of line so that it costs the scan neither registers nor instruction-cache space: the exact squared
// distance of one (query, vertex) pair in the reference's arithmetic, offered to the query's list.  Returns the list's
// squared-distance gate afterwards.
__device__ __noinline__ double nn_exact_offer(const double* __restrict__ query, const double* __restrict__ vertex, int dim, NnList* L, int k,
                                              int32_t idx) {
  double se = 0.0;
  for (int c = 0; c < dim; ++c) { const double d = __dsub_rn(vertex[c], query[c]); se = __dadd_rn(se, __dmul_rn(d, d)); }
  nn_offer(*L, k, se, idx);
  return L->gate_s;
}

template <int DIMP, int RQ, int RV, bool DEFER>
__global__ void __launch_bounds__(NN_BLOCK) nearest_scan_kernel(const NearestArgs A) {
  __shared__ __align__(16) double tile[NN_TILE * DIMP];
  bool live[RQ];
  double q[RQ][DIMP];
  const long long q_first = (long long)blockIdx.x * RQ * NN_BLOCK + threadIdx.x;  // query r of this thread: q_first + r NN_BLOCK
#pragma unroll
  for (int r = 0; r < RQ; ++r) {
    const long long qi = q_first + (long long)r * NN_BLOCK;
    live[r] = qi < A.n_queries;
#pragma unroll
    for (int c = 0; c < DIMP; ++c) q[r][c] = (live[r] && c < A.dim) ? A.queries[qi * A.dim + c] : 0.0;
  }
  const int k = A.k;
  NnList L[RQ];
  // squared-distance gate: anything at or above it is rejected without a sqrt.  Until the list is full the gate is the
  // radius squared, widened by a few ulps so that no candidate with sqrt(s) < radius is lost to rounding.
  const double gate0 = isinf(A.radius) ? A.radius : A.radius * A.radius * (1.0 + 8.0 * 2.220446049250313e-16);
  const double widen = 1.0 + 4.0 * (DIMP + 2) * 2.220446049250313e-16;
  double fast[RQ];  // gate of the fused scan: the list's exact gate, widened
#pragma unroll
  for (int r = 0; r < RQ; ++r) { L[r].cnt = 0; L[r].gate_s = gate0; L[r].gate_d = A.radius; fast[r] = live[r] ? gate0 * widen : -1.0; }

  const long long v0 = (long long)blockIdx.y * A.chunk;
  const long long v1 = v0 + A.chunk < A.n_vertices ? v0 + A.chunk : A.n_vertices;
  for (long long base = v0; base < v1; base += NN_TILE) {
    const int nt = (int)(v1 - base < NN_TILE ? v1 - base : NN_TILE);
    __syncthreads();
    // tile load: nt * dim consecutive doubles, coalesced; padded coordinates are zero
    if (A.dim == DIMP) {
      for (int e = threadIdx.x; e < nt * DIMP; e += NN_BLOCK) tile[e] = A.vertices[base * DIMP + e];
    } else {
      for (int e = threadIdx.x; e < nt * DIMP; e += NN_BLOCK) {
        const int v = e / DIMP, c = e - v * DIMP;
        tile[e] = c < A.dim ? A.vertices[(base + v) * A.dim + c] : 0.0;
      }
    }
    __syncthreads();
    if (!live[0]) continue;  // (queries of a thread are live front to back)
    unsigned long long hits[RQ];  // bit j: vertex j of the tile may enter the list of query r
#pragma unroll
    for (int r = 0; r < RQ; ++r) hits[r] = 0ull;
    for (int j = 0; j < nt; j += RV) {
      double s[RQ][RV];
      const double* t[RV];
#pragma unroll
      for (int m = 0; m < RV; ++m) t[m] = &tile[(j + m < nt ? j + m : j) * DIMP];
#pragma unroll
      for (int r = 0; r < RQ; ++r)
#pragma unroll
        for (int m = 0; m < RV; ++m) s[r][m] = 0.0;
#pragma unroll
      for (int c = 0; c < DIMP; c += 2) {
        double2 a[RV];
#pragma unroll
        for (int m = 0; m < RV; ++m) a[m] = *reinterpret_cast<const double2*>(t[m] + c);
#pragma unroll
        for (int r = 0; r < RQ; ++r) {
#pragma unroll
          for (int m = 0; m < RV; ++m) { const double d = a[m].x - q[r][c]; s[r][m] = fma(d, d, s[r][m]); }
#pragma unroll
          for (int m = 0; m < RV; ++m) { const double d = a[m].y - q[r][c + 1]; s[r][m] = fma(d, d, s[r][m]); }
        }
      }
      if (DEFER) {
#pragma unroll
        for (int r = 0; r < RQ; ++r)
#pragma unroll
          for (int m = 0; m < RV; ++m) hits[r] |= (unsigned long long)(s[r][m] < fast[r] ? 1u : 0u) << (j + m);
      } else {  // k = 1: a query accepts ~ln(chunk) vertices in all; they are handled where they are met
        unsigned now = 0;
#pragma unroll
        for (int r = 0; r < RQ; ++r)
#pragma unroll
          for (int m = 0; m < RV; ++m) now |= (s[r][m] < fast[r] ? 1u : 0u) << (r * RV + m);
        while (now) {
          const int b = __ffs(now) - 1;
          now &= now - 1;
          const int r = b / RV, m = b - r * RV;
          if (j + m >= nt) continue;
          const long long qi = q_first + (long long)r * NN_BLOCK;
          const double g = nn_exact_offer(A.queries + qi * A.dim, &tile[(j + m) * DIMP], A.dim, &L[r], k, (int32_t)(base + j + m)) * widen;
#pragma unroll
          for (int r2 = 0; r2 < RQ; ++r2) if (r2 == r) fast[r2] = g;
        }
      }
    }
    // k > 1: the candidates of the tile, per query in vertex order, decided by the reference's own arithmetic.  Collected
    // per tile rather than handled where they are met: a warp leaves the scan once per tile instead of once per accepting
    // lane (some list of the warp accepts a vertex in most groups of four).  The gate a candidate was admitted by may be
    // stale by then; the exact test inside nn_offer uses the current one.
#pragma unroll
    for (int r = 0; r < RQ && DEFER; ++r) {
      unsigned long long h = hits[r] & (nt < 64 ? ((1ull << nt) - 1ull) : ~0ull);
      if (h) {
        const long long qi = q_first + (long long)r * NN_BLOCK;
        double g = fast[r];
        while (h) {
          const int j = __ffsll((long long)h) - 1;
          h &= h - 1;
          g = nn_exact_offer(A.queries + qi * A.dim, &tile[j * DIMP], A.dim, &L[r], k, (int32_t)(base + j)) * widen;
        }
        fast[r] = g;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RQ; ++r) {
    if (!live[r]) continue;
    const long long qi = q_first + (long long)r * NN_BLOCK;
    int32_t* oi = A.part_idx + (qi * A.n_chunks + blockIdx.y) * k;
    double* od = A.part_dist + (qi * A.n_chunks + blockIdx.y) * k;
    for (int m = 0; m < k; ++m) {
      oi[m] = m < L[r].cnt ? L[r].il[m] : -1;
      od[m] = m < L[r].cnt ? L[r].dl[m] : INFINITY;
    }
  }
}

// the k best of n_chunks sorted lists per query, by (distance, vertex index)
__global__ void __launch_bounds__(128) nearest_merge_kernel(long long n_queries, int n_chunks, int k, const int32_t* __restrict__ part_idx,
                                                             const double* __restrict__ part_dist, int32_t* __restrict__ idx,
                                                             double* __restrict__ dist, int32_t* __restrict__ count) {
  const long long qi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= n_queries) return;
  double dl[NN_MAX_K];
  int32_t il[NN_MAX_K];
  int cnt = 0;
  for (int c = 0; c < n_chunks; ++c) {  // chunks in vertex order: equal distances keep the lower index in front
    const int32_t* pi = part_idx + (qi * n_chunks + c) * k;
    const double* pd = part_dist + (qi * n_chunks + c) * k;
    for (int r = 0; r < k; ++r) {
      const int32_t id = pi[r];
      if (id < 0) break;
      const double d = pd[r];
      if (cnt == k && !(d < dl[k - 1])) break;  // the rest of this list is no better
      int pos = cnt < k ? cnt : k - 1;
      while (pos > 0 && dl[pos - 1] > d) { dl[pos] = dl[pos - 1]; il[pos] = il[pos - 1]; --pos; }
      dl[pos] = d; il[pos] = id;
      if (cnt < k) ++cnt;
    }
  }
  for (int r = 0; r < k; ++r) {
    idx[qi * k + r] = r < cnt ? il[r] : -1;
    if (dist) dist[qi * k + r] = r < cnt ? dl[r] : INFINITY;
  }
  if (count) count[qi] = cnt;
}

// one or two queries per thread (rkb_nearest decides)
template <int DIMP>
cudaError_t launch_scan(const NearestArgs& A, int rq, unsigned n_chunks, cudaStream_t s) {
  const long long per_cta = (long long)NN_BLOCK * rq;
  const dim3 grid((unsigned)((A.n_queries + per_cta - 1) / per_cta), n_chunks);
  if (A.k == 1) {
    if (rq == 2) nearest_scan_kernel<DIMP, (DIMP <= 24 ? 2 : 1), 4, false><<<grid, NN_BLOCK, 0, s>>>(A);
    else nearest_scan_kernel<DIMP, 1, 4, false><<<grid, NN_BLOCK, 0, s>>>(A);
  } else {
    if (rq == 2) nearest_scan_kernel<DIMP, (DIMP <= 24 ? 2 : 1), 4, true><<<grid, NN_BLOCK, 0, s>>>(A);
    else nearest_scan_kernel<DIMP, 1, 4, true><<<grid, NN_BLOCK, 0, s>>>(A);
  }
  return cudaGetLastError();
}

thread_local char g_nn_err[160] = "";

}  // namespace

extern "C" {

const char* rkb_nearest_last_error(void) { return g_nn_err; }

int rkb_nearest(int device, size_t n_vertices, const double* vertices, size_t n_queries, const double* queries, int dim, int k,
                double radius, int32_t* index, double* distance, int32_t* count, unsigned flags, void* stream) {
  g_nn_err[0] = 0;
  if (dim < 1 || dim > RKB_NEAREST_MAX_DIM || k < 1 || k > RKB_NEAREST_MAX_K || !(radius > 0.0)) return RKB_ERR_INVALID;
  if (n_queries == 0) return RKB_OK;
  if (!queries || !index || (n_vertices > 0 && !vertices) || n_vertices > 0x7fffffffull) return RKB_ERR_INVALID;
  if (flags & ~RKB_MEM_DEVICE) return RKB_ERR_UNSUPPORTED;  // AoS only
  const bool on_device = (flags & RKB_MEM_DEVICE) != 0;
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) {
    cudaGetLastError();
    snprintf(g_nn_err, sizeof g_nn_err, "cudaSetDevice(%d) failed", device);
    return RKB_ERR_CUDA;
  }
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  {  // keep the stream-ordered pool's memory across calls (by default it goes back to the driver at every synchronisation)
    static bool pool_set[64] = {};
    if (device >= 0 && device < 64 && !pool_set[device]) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = 1ull << 30;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
      pool_set[device] = true;
    }
  }
  const int dimp = (dim + 3) / 4 * 4;
  // queries per thread: as many as their coordinates leave registers for, while every thread still gets work
  // (measured, profiles/r2_nearest.md: 2 queries per thread reach the rate 4 reach, at 5 instead of 3 CTAs per SM)
  const int rq = (dimp <= 24 && n_queries > (size_t)NN_BLOCK) ? 2 : 1;
  const long long gx = (long long)((n_queries + (size_t)NN_BLOCK * rq - 1) / ((size_t)NN_BLOCK * rq));
  // enough CTAs for every SM several times over, but no chunk shorter than a few tiles
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  long long want = (4LL * sms + gx - 1) / gx;
  // every chunk starts with empty lists and accepts ~k (1 + ln(chunk / k)) vertices per query before its gates are tight
  const long long cmin = k == 1 ? 4 * NN_TILE : 16384;
  const long long max_chunks = (long long)((n_vertices + cmin - 1) / cmin);
  if (want > max_chunks) want = max_chunks;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  long long chunk = ((long long)n_vertices + want - 1) / want;
  chunk = (chunk + NN_TILE - 1) / NN_TILE * NN_TILE;
  if (chunk < NN_TILE) chunk = NN_TILE;
  const int n_chunks = n_vertices ? (int)(((long long)n_vertices + chunk - 1) / chunk) : 1;

  void *dv = nullptr, *dq = nullptr, *dpi = nullptr, *dpd = nullptr, *di = nullptr, *dd = nullptr, *dc = nullptr;
  const size_t bv = n_vertices * (size_t)dim * sizeof(double), bq = n_queries * (size_t)dim * sizeof(double);
  const size_t bo = n_queries * (size_t)k;
  auto fail = [&](const char* where) {
    snprintf(g_nn_err, sizeof g_nn_err, "%s: %s", where, cudaGetErrorString(e));
    cudaGetLastError();
    void* all[] = {on_device ? nullptr : dv, on_device ? nullptr : dq, dpi, dpd, on_device ? nullptr : di, on_device ? nullptr : dd,
                   on_device ? nullptr : dc};
    for (void* p : all) if (p) cudaFreeAsync(p, s);
    cudaStreamSynchronize(s);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return RKB_ERR_CUDA;
  };
  if (on_device) {
    dv = const_cast<double*>(vertices); dq = const_cast<double*>(queries); di = index; dd = distance; dc = count;
  } else {
    if (bv && (e = cudaMallocAsync(&dv, bv, s)) != cudaSuccess) return fail("cudaMallocAsync");
    if ((e = cudaMallocAsync(&dq, bq, s)) != cudaSuccess) return fail("cudaMallocAsync");
    if ((e = cudaMallocAsync(&di, bo * sizeof(int32_t), s)) != cudaSuccess) return fail("cudaMallocAsync");
    if (distance && (e = cudaMallocAsync(&dd, bo * sizeof(double), s)) != cudaSuccess) return fail("cudaMallocAsync");
    if (count && (e = cudaMallocAsync(&dc, n_queries * sizeof(int32_t), s)) != cudaSuccess) return fail("cudaMallocAsync");
    if (bv && (e = cudaMemcpyAsync(dv, vertices, bv, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail("cudaMemcpyAsync");
    if ((e = cudaMemcpyAsync(dq, queries, bq, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail("cudaMemcpyAsync");
  }
  if ((e = cudaMallocAsync(&dpi, bo * n_chunks * sizeof(int32_t), s)) != cudaSuccess) return fail("cudaMallocAsync");
  if ((e = cudaMallocAsync(&dpd, bo * n_chunks * sizeof(double), s)) != cudaSuccess) return fail("cudaMallocAsync");

  NearestArgs A;
  A.vertices = (const double*)dv; A.queries = (const double*)dq;
  A.n_vertices = (long long)n_vertices; A.n_queries = (long long)n_queries;
  A.dim = dim; A.k = k; A.radius = radius; A.chunk = chunk;
  A.part_idx = (int32_t*)dpi; A.part_dist = (double*)dpd; A.n_chunks = n_chunks;
  switch (dimp) {
    case 4: e = launch_scan<4>(A, rq, (unsigned)n_chunks, s); break;
    case 8: e = launch_scan<8>(A, rq, (unsigned)n_chunks, s); break;
    case 12: e = launch_scan<12>(A, rq, (unsigned)n_chunks, s); break;
    case 16: e = launch_scan<16>(A, rq, (unsigned)n_chunks, s); break;
    case 20: e = launch_scan<20>(A, rq, (unsigned)n_chunks, s); break;
    case 24: e = launch_scan<24>(A, rq, (unsigned)n_chunks, s); break;
    case 28: e = launch_scan<28>(A, rq, (unsigned)n_chunks, s); break;
    case 32: e = launch_scan<32>(A, rq, (unsigned)n_chunks, s); break;
    case 36: e = launch_scan<36>(A, rq, (unsigned)n_chunks, s); break;
    case 40: e = launch_scan<40>(A, rq, (unsigned)n_chunks, s); break;
    case 44: e = launch_scan<44>(A, rq, (unsigned)n_chunks, s); break;
    default: e = launch_scan<48>(A, rq, (unsigned)n_chunks, s); break;
  }
  if (e != cudaSuccess) return fail("nearest_scan_kernel");
  nearest_merge_kernel<<<(unsigned)((n_queries + 127) / 128), 128, 0, s>>>((long long)n_queries, n_chunks, k, (const int32_t*)dpi,
                                                                           (const double*)dpd, (int32_t*)di, (double*)dd, (int32_t*)dc);
  if ((e = cudaGetLastError()) != cudaSuccess) return fail("nearest_merge_kernel");
  if (!on_device) {
    if ((e = cudaMemcpyAsync(index, di, bo * sizeof(int32_t), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return fail("cudaMemcpyAsync");
    if (distance && (e = cudaMemcpyAsync(distance, dd, bo * sizeof(double), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return fail("cudaMemcpyAsync");
    if (count && (e = cudaMemcpyAsync(count, dc, n_queries * sizeof(int32_t), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return fail("cudaMemcpyAsync");
    if (dv) cudaFreeAsync(dv, s);
    cudaFreeAsync(dq, s); cudaFreeAsync(di, s);
    if (dd) cudaFreeAsync(dd, s);
    if (dc) cudaFreeAsync(dc, s);
  }
  cudaFreeAsync(dpi, s); cudaFreeAsync(dpd, s);
  if (!on_device && (e = cudaStreamSynchronize(s)) != cudaSuccess) { dv = dq = di = dd = dc = dpi = dpd = nullptr; return fail("cudaStreamSynchronize"); }
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  return RKB_OK;
}

}  // extern "C"
