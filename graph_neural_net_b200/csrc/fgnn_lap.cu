// Batched linear assignment on the device: accuracy_linear_assignment of the reference
// (toolbox/metrics.py:92-116) without the per-graph device->host copy and the host Hungarian.
//
// The reference maximises sum_i log_softmax(scores)[i, col(i)] with scipy.optimize.linear_sum_assignment.
// log_softmax subtracts a per-row constant, which does not change the optimal assignment, so the kernel
// minimises cost[i][j] = -scores[i][j] directly.  Algorithm: shortest augmenting paths with dual variables
// (Jonker-Volgenant as restated by Crouse 2016 -- the algorithm scipy implements), in double precision,
// one CTA per graph: the column scan of every Dijkstra step is spread over the CTA's threads and closed by
// a block-wide arg-min; duals, path and matching arrays live in shared memory.  The result is the exact
// optimum (the same assignment as scipy whenever the optimum is unique).
#include "fgnn_common.cuh"

#include <cfloat>

namespace fgnn {
namespace lap {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct Best {
  double val;
  int idx;     // column
  int free_;   // 1 if the column is unassigned (preferred on ties, as scipy does)
};

__device__ __forceinline__ bool better(const Best& a, const Best& b) {   // is a strictly better than b
  if (a.val != b.val) return a.val < b.val;
  if (a.free_ != b.free_) return a.free_ > b.free_;
  return a.idx < b.idx;
}

__global__ void __launch_bounds__(kThreads)
lap_kernel(const float* __restrict__ scores, int32_t* __restrict__ col_of_row, int32_t* __restrict__ correct,
           double* __restrict__ total_cost, int N, const int32_t* __restrict__ n_per_graph) {
  extern __shared__ __align__(16) unsigned char lap_smem[];
  const int g = blockIdx.x;
  const int n = n_per_graph ? n_per_graph[g] : N;
  double* u = reinterpret_cast<double*>(lap_smem);      // row duals
  double* v = u + N;                                    // column duals
  double* spc = v + N;                                  // shortest path cost to each column
  int* path = reinterpret_cast<int*>(spc + N);          // predecessor row of each column
  int* col4row = path + N;
  int* row4col = col4row + N;
  int* vis = row4col + N;                               // rows visited by the current search (SR)
  unsigned char* sc = reinterpret_cast<unsigned char*>(vis + N);   // columns scanned by the current search (SC)
  __shared__ Best s_best[kWarps];
  __shared__ Best s_pick;
  const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
  const float* S = scores + (long)g * N * N;

  for (int k = tid; k < N; k += kThreads) {
    u[k] = 0.0;
    v[k] = 0.0;
    col4row[k] = -1;
    row4col[k] = -1;
  }
  __syncthreads();

  for (int cur = 0; cur < n; ++cur) {
    for (int j = tid; j < n; j += kThreads) {
      spc[j] = DBL_MAX;
      sc[j] = 0;
    }
    __syncthreads();
    double min_val = 0.0;
    int i = cur, sink = -1, nvis = 0;
    while (sink < 0) {
      if (tid == 0) vis[nvis] = i;
      ++nvis;
      const double ui = u[i];
      const float* row = S + (long)i * N;
      Best b{DBL_MAX, 0x7fffffff, 0};
      for (int j = tid; j < n; j += kThreads) {
        if (sc[j]) continue;
        const double r = min_val - (double)row[j] - ui - v[j];
        double cur_cost = spc[j];
        if (r < cur_cost) {
          spc[j] = r;
          path[j] = i;
          cur_cost = r;
        }
        Best c{cur_cost, j, row4col[j] < 0 ? 1 : 0};
        if (better(c, b)) b = c;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Best c;
        c.val = __shfl_xor_sync(0xffffffffu, b.val, o);
        c.idx = __shfl_xor_sync(0xffffffffu, b.idx, o);
        c.free_ = __shfl_xor_sync(0xffffffffu, b.free_, o);
        if (better(c, b)) b = c;
      }
      if (lane == 0) s_best[warp] = b;
      __syncthreads();
      if (tid == 0) {
        Best p = s_best[0];
        for (int w = 1; w < kWarps; ++w)
          if (better(s_best[w], p)) p = s_best[w];
        s_pick = p;
        sc[p.idx] = 1;
      }
      __syncthreads();
      const Best p = s_pick;
      min_val = p.val;
      if (p.free_) sink = p.idx;
      else i = row4col[p.idx];
    }
    // dual update (Crouse 2016, step 4) and augmentation along the path
    for (int k = tid; k < nvis; k += kThreads) {
      const int r = vis[k];
      if (r == cur) u[r] += min_val;
      else u[r] += min_val - spc[col4row[r]];
    }
    for (int j = tid; j < n; j += kThreads)
      if (sc[j]) v[j] -= min_val - spc[j];
    __syncthreads();
    if (tid == 0) {
      int j = sink;
      while (true) {
        const int r = path[j];
        row4col[j] = r;
        const int prev = col4row[r];
        col4row[r] = j;
        j = prev;
        if (r == cur) break;
      }
    }
    __syncthreads();
  }

  // outputs: matching, #fixed points (the identity is the ground truth, metrics.py:103), optimal cost
  int hits = 0;
  double cost = 0.0;
  for (int r = tid; r < N; r += kThreads) {
    const int c = r < n ? col4row[r] : -1;
    if (col_of_row) col_of_row[(long)g * N + r] = c;
    if (r < n) {
      hits += (c == r);
      cost -= (double)S[(long)r * N + c];
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    hits += __shfl_xor_sync(0xffffffffu, hits, o);
    cost += __shfl_xor_sync(0xffffffffu, cost, o);
  }
  __shared__ int s_hits[kWarps];
  __shared__ double s_cost[kWarps];
  if (lane == 0) { s_hits[warp] = hits; s_cost[warp] = cost; }
  __syncthreads();
  if (tid == 0) {
    int h = 0;
    double c = 0.0;
    for (int w = 0; w < kWarps; ++w) { h += s_hits[w]; c += s_cost[w]; }
    correct[g] = h;
    if (total_cost) total_cost[g] = c;
  }
}

inline size_t smem_bytes(int N) { return (size_t)N * (3 * sizeof(double) + 4 * sizeof(int) + 1) + 16; }

}  // namespace

int lap_fwd(const float* scores, int32_t* col_of_row, int32_t* correct, double* total_cost, int G, int N,
            const int32_t* n_per_graph, cudaStream_t st) {
  FGNN_CHECK_ARG(scores && correct, "null pointer");
  FGNN_CHECK_ARG(G >= 1 && N >= 1, "bad sizes G=%d N=%d", G, N);
  const size_t smem = smem_bytes(N);
  if (smem > 200 * 1024) return fail(FGNN_ERR_UNSUPPORTED, "linear assignment supports N <= 4800 (got %d)", N);
  FGNN_CUDA(cudaFuncSetAttribute(lap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lap_kernel<<<G, kThreads, smem, st>>>(scores, col_of_row, correct, total_cost, N, n_per_graph);
  FGNN_LAUNCHED();
  return FGNN_OK;
}

}  // namespace lap
}  // namespace fgnn
