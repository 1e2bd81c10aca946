// Tensor-core path of libfgnn_b200 (FGNN_BF16 / FGNN_FP16): TMA-fed tcgen05 kernels with TMEM
// accumulators for the two GEMM families of a 2-FGNN block, plus the small CUDA-core kernels
// that glue them (statistics, weight folding, pooling).
//
// Data layout (DESIGN.md "HBM layout"): every activation is a set of 16-bit planes
//   act[g][c][i][j],  i < N rows, row pitch NP = round_up(N, 8) elements (16-byte rows for TMA),
// holding the PRE-GraphNorm output of the MLP that produced it.  GraphNorm is never applied to a
// stored tensor: its per-(graph, channel) scale a and shift s
//   a = w / (2 sqrt(n (var + eps))),   s = beta - a * mean        (layers.py:68-80)
// are folded into the consumer: into the first 1x1-conv weights of the next MLP
// (W diag(a), b + W s), into the epilogue of the N x N matmul
// ((a1 Y1 + s1 J)(a2 Y2 + s2 J) = a1 a2 Y1 Y2 + a1 s2 r1 1^T + s1 a2 1 c2^T + s1 s2 n J), and into the
// final max-pool (max of a*y+s = a*max(y)+s or a*min(y)+s by the sign of a).
#include "fgnn_tc.cuh"
#include "fgnn_ptx.cuh"

#include <cstdlib>
#include <cstring>

namespace fgnn {
namespace tc {

using namespace ptx;

namespace {

constexpr int kMaxN = 1024;   // plane_stats keeps column partials in registers
constexpr int kTileM = 128;   // pixels per MLP tile / rows per matmul tile

__device__ __forceinline__ int graph_n(const int32_t* n_per_graph, int g, int N) {
  return n_per_graph ? n_per_graph[g] : N;
}
// rows of a plane that kernels must keep finite/zero so K-loops of the matmul may over-read
__device__ __forceinline__ int rows_cover(const int32_t* n_per_graph, int g, int N) {
  if (!n_per_graph) return N;
  int n = n_per_graph[g];
  int r = (n + 63) / 64 * 64;
  return r < N ? r : N;
}

// =============================================================================================
// small CUDA-core kernels
// =============================================================================================

// fp32 (G,C,N,N) -> 16-bit planes (G,C,N,NP); padding and invalid positions written as zero
template <typename T>
__global__ void to_planes_kernel(const float* __restrict__ x, T* __restrict__ out, int C, int N, int NP,
                                 const int32_t* __restrict__ n_per_graph) {
  const int gc = blockIdx.y;
  const int g = gc / C;
  const int n = graph_n(n_per_graph, g, N);
  const long Ppl = (long)N * NP;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < Ppl; p += (long)gridDim.x * blockDim.x) {
    int i = (int)(p / NP), j = (int)(p % NP);
    float v = (i < n && j < n) ? x[((long)gc * N + i) * N + j] : 0.f;
    out[(long)gc * Ppl + p] = Elem<T>::from_float(v);
  }
}

// 16-bit planes -> fp32 (G,C,N,N), optionally applying y = a*v + s on valid positions (debug / tests)
template <typename T>
__global__ void from_planes_kernel(const T* __restrict__ in, float* __restrict__ out, const float* __restrict__ coef,
                                   int C, int N, int NP, const int32_t* __restrict__ n_per_graph) {
  const int gc = blockIdx.y;
  const int g = gc / C;
  const int n = graph_n(n_per_graph, g, N);
  const float a = coef ? coef[2 * gc] : 1.f, s = coef ? coef[2 * gc + 1] : 0.f;
  const long P = (long)N * N;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long)gridDim.x * blockDim.x) {
    int i = (int)(p / N), j = (int)(p % N);
    float v = 0.f;
    if (i < n && j < n) v = a * Elem<T>::to_float(in[((long)gc * N + i) * NP + j]) + s;
    out[(long)gc * P + p] = v;
  }
}

// fp32 hidden-layer weights (co, ci) -> 16-bit [co][Kh] zero padded (Kh multiple of 64)
template <typename T>
__global__ void convert_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int co, int ci, int Kh) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= co * Kh) return;
  int o = idx / Kh, k = idx % Kh;
  out[idx] = Elem<T>::from_float(k < ci ? w[o * ci + k] : 0.f);
}

// Per-graph folded first-layer weights.  Source s contributes channels [0,c_s) placed at K offset
// koff_s (K padded to multiples of 16 per source, whole row padded to K1g, a multiple of 64).
//   Wf[g][co][koff_s + ch] = W[co][col_s + ch] * a_s[g][ch]
//   bf[g][co]              = b[co] + sum_s sum_ch W[co][col_s + ch] * s_s[g][ch]
struct FoldArgs {
  const float* w;        // (c_out, c0 + c1)
  const float* b;        // (c_out)
  const float* coef[2];  // [G][c_s][2] or null (identity)
  int c[2], koff[2], nsrc;
  int c_out, K1g;
};
template <typename T>
__global__ void fold_weights_kernel(FoldArgs a, T* __restrict__ wf, float* __restrict__ bf) {
  const int g = blockIdx.x;
  const int cin = a.c[0] + (a.nsrc > 1 ? a.c[1] : 0);
  T* wg = wf + (long)g * a.c_out * a.K1g;
  for (int idx = threadIdx.x; idx < a.c_out * a.K1g; idx += blockDim.x) {
    int co = idx / a.K1g, k = idx % a.K1g;
    float v = 0.f;
    int col = 0;
    for (int s = 0; s < a.nsrc; ++s) {
      int ch = k - a.koff[s];
      if (ch >= 0 && ch < a.c[s]) {
        float sc = a.coef[s] ? a.coef[s][((long)g * a.c[s] + ch) * 2] : 1.f;
        v = a.w[co * cin + col + ch] * sc;
      }
      col += a.c[s];
    }
    wg[idx] = Elem<T>::from_float(v);
  }
  for (int co = threadIdx.x; co < a.c_out; co += blockDim.x) {
    float acc = a.b[co];
    int col = 0;
    for (int s = 0; s < a.nsrc; ++s) {
      if (a.coef[s])
        for (int ch = 0; ch < a.c[s]; ++ch)
          acc = fmaf(a.w[co * cin + col + ch], a.coef[s][((long)g * a.c[s] + ch) * 2 + 1], acc);
      col += a.c[s];
    }
    bf[(long)g * a.c_out + co] = acc;
  }
}

// Per-plane statistics of a stored pre-norm plane over its valid n x n corner:
//   coef[q] = {a, s} (GraphNorm scale/shift), rsum[q][i] = sum_j y[i][j], csum[q][j] = sum_i y[i][j].
template <typename T>
__global__ void __launch_bounds__(256)
plane_stats16_kernel(const T* __restrict__ y, const float* __restrict__ gw, const float* __restrict__ gb,
                     float eps, float* __restrict__ coef, float* __restrict__ rsum, float* __restrict__ csum,
                     int C, int N, int NP, const int32_t* __restrict__ n_per_graph) {
  extern __shared__ float sh_col[];  // [8][NP]
  const int q = blockIdx.x;
  const int g = q / C, c = q % C;
  const int n = graph_n(n_per_graph, g, N);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const T* yp = y + (long)q * N * NP;
  float colacc[kMaxN / 64][2];
#pragma unroll
  for (int t = 0; t < kMaxN / 64; ++t) colacc[t][0] = colacc[t][1] = 0.f;
  float s1 = 0.f, s2 = 0.f;
  for (int i = warp; i < n; i += 8) {
    const T* row = yp + (long)i * NP;
    float rs = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxN / 64; ++t) {
      int j = t * 64 + lane * 2;
      if (j < n) {
        float v0 = Elem<T>::to_float(row[j]);
        float v1 = (j + 1 < n) ? Elem<T>::to_float(row[j + 1]) : 0.f;
        colacc[t][0] += v0;
        colacc[t][1] += v1;
        rs += v0 + v1;
        s2 = fmaf(v0, v0, fmaf(v1, v1, s2));
      }
    }
    s1 += rs;
    for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
    if (lane == 0 && rsum) rsum[(long)q * N + i] = rs;
  }
  if (rsum)
    for (int i = n + threadIdx.x; i < N; i += blockDim.x) rsum[(long)q * N + i] = 0.f;
#pragma unroll
  for (int t = 0; t < kMaxN / 64; ++t) {
    int j = t * 64 + lane * 2;
    if (j < NP) {
      sh_col[warp * NP + j] = colacc[t][0];
      if (j + 1 < NP) sh_col[warp * NP + j + 1] = colacc[t][1];
    }
  }
  __shared__ double red[2][8];
  double d1 = s1, d2 = s2;
  for (int o = 16; o > 0; o >>= 1) {
    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
  }
  if (lane == 0) { red[0][warp] = d1; red[1][warp] = d2; }
  __syncthreads();
  if (csum)
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      float t = 0.f;
      if (j < n)
        for (int w = 0; w < 8; ++w) t += sh_col[w * NP + j];
      csum[(long)q * N + j] = t;
    }
  if (threadIdx.x == 0) {
    double S = 0, SS = 0;
    for (int w = 0; w < 8; ++w) { S += red[0][w]; SS += red[1][w]; }
    double cnt = (double)n * n;
    double mean = S / cnt;
    double var = SS / cnt - mean * mean;
    if (var < 0) var = 0;
    double a = (double)(gw ? gw[c] : 1.f) / (2.0 * sqrt((double)n * (var + (double)eps)));
    coef[2 * q] = (float)a;
    coef[2 * q + 1] = (float)((double)(gb ? gb[c] : 0.f) - a * mean);
  }
}

// emb[g][c][i] = max_{j<n} (a*y[i][j] + s); rows >= n -> 0   (layers.py:194-203 on folded data)
template <typename T>
__global__ void pool_kernel(const T* __restrict__ y, const float* __restrict__ coef, float* __restrict__ emb, int C,
                            int N, int NP, long rows, const int32_t* __restrict__ n_per_graph) {
  const long row = (long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (row >= rows) return;
  const int lane = threadIdx.x % 32;
  const int i = (int)(row % N);
  const long q = row / N;
  const int g = (int)(q / C);
  const int n = graph_n(n_per_graph, g, N);
  float out = 0.f;
  if (i < n) {
    const T* r = y + (q * N + i) * NP;
    float mx = -INFINITY, mn = INFINITY;
    for (int j = lane; j < n; j += 32) {
      float v = Elem<T>::to_float(r[j]);
      mx = fmaxf(mx, v);
      mn = fminf(mn, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    const float a = coef[2 * q], s = coef[2 * q + 1];
    out = (a >= 0.f) ? fmaf(a, mx, s) : fmaf(a, mn, s);
  }
  if (lane == 0) emb[row] = out;
}

// =============================================================================================
// K_A: fused conv chain on tensor cores.
//   tile  = 128 consecutive pixels of one graph (flattened N x NP plane), all channels
//   layer1: D[128 px, COUT] = X[px, K1] * W1f[g]^T     A = TMA-staged smem (MN-major), B = smem (K-major)
//   layer>=2: D = relu(D + b) (16-bit, written back to TMEM) * W^T   A = TMEM, B = smem
//   output: raw last-layer accumulators as 16-bit planes (its bias cancels in GraphNorm).
// Warp roles: warp 0 = TMA producer + MMA issuer (one elected lane), warps 1-4 = TMEM epilogue.
// =============================================================================================
template <typename T>
struct MlpArgs {
  int G, N, NP;
  long Ppl;
  int k_src[2], nsrc, K1, K1g;
  int depth, Kh;
  const float* bias1;                   // [G][COUT] folded layer-1 bias
  const float* bias[FGNN_MAX_DEPTH];    // layer l >= 1 biases
  T* out;                               // [G][COUT][N][NP]
  const int32_t* n_per_graph;
};

template <int COUT>
struct MlpSmem {
  static constexpr int kStages = 2;
  static size_t bytes(int K1, int K1g, int depth, int Kh) {
    return 1024 + (size_t)kStages * K1 * 256 + (size_t)K1g * COUT * 2 + (size_t)(depth - 1) * Kh * COUT * 2 + 256;
  }
};

template <typename T, int COUT>
__global__ void __launch_bounds__(160)
tc_mlp_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1,
              const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_wh,
              const MlpArgs<T> args) {
  constexpr int kStages = MlpSmem<COUT>::kStages;
  constexpr uint32_t kTmemCols = (COUT * 3 / 2 <= 64) ? 64 : (COUT * 3 / 2 <= 128 ? 128 : 256);
  constexpr int kHCol = COUT;  // TMEM column where the packed 16-bit hidden activations start
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int K1 = args.K1, K1g = args.K1g, depth = args.depth, Kh = args.Kh;
  const uint32_t stage_bytes = (uint32_t)K1 * 256u;
  uint8_t* s_in = smem;
  uint8_t* s_w1 = s_in + (size_t)kStages * stage_bytes;
  uint8_t* s_wh = s_w1 + (size_t)K1g * COUT * 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_wh + (size_t)(depth - 1) * Kh * COUT * 2);
  uint64_t* in_full = bars;              // [kStages]
  uint64_t* in_empty = bars + kStages;   // [kStages]
  uint64_t* w1_full = bars + 2 * kStages;
  uint64_t* wh_full = w1_full + 1;
  uint64_t* mma_done = wh_full + 1;
  uint64_t* h_ready = mma_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_ready + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  // ---- tile range of this CTA: contiguous chunk of the flat (graph, tile) list ----------------
  long total = 0;
  for (int g = 0; g < args.G; ++g)
    total += ((long)rows_cover(args.n_per_graph, g, args.N) * args.NP + kTileM - 1) / kTileM;
  const long t_begin = total * blockIdx.x / gridDim.x;
  const long t_end = total * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&in_full[s], 1); mbar_init(&in_empty[s], 1); }
    mbar_init(w1_full, 1);
    mbar_init(wh_full, 1);
    mbar_init(mma_done, 1);
    mbar_init(h_ready, 4);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && t_begin < t_end) {
      // ================= producer + MMA issuer (single thread) =================
      prefetch_tensormap(&map_x0);
      prefetch_tensormap(&map_w1);
      if (depth > 1) {
        const int atoms = Kh / 64;
        mbar_arrive_expect_tx(wh_full, (uint32_t)((depth - 1) * Kh * COUT * 2));
        for (int l = 0; l < depth - 1; ++l)
          for (int at = 0; at < atoms; ++at)
            tma_load_3d(s_wh + ((size_t)l * atoms + at) * COUT * 128, &map_wh, wh_full, at * 64, 0, l);
      }
      // walk to the first tile
      int g = 0;
      long gbase = 0;
      long gtiles = ((long)rows_cover(args.n_per_graph, 0, args.N) * args.NP + kTileM - 1) / kTileM;
      auto seek = [&](long t) {
        while (t >= gbase + gtiles) {
          gbase += gtiles;
          ++g;
          gtiles = ((long)rows_cover(args.n_per_graph, g, args.N) * args.NP + kTileM - 1) / kTileM;
        }
      };
      auto issue_load = [&](long t, int stage) {
        seek(t);
        const int p0 = (int)((t - gbase) * kTileM);
        uint8_t* dst = s_in + (size_t)stage * stage_bytes;
        mbar_arrive_expect_tx(&in_full[stage], stage_bytes);
        for (int u = 0; u < 2; ++u) {
          tma_load_3d(dst + (size_t)u * K1 * 128, &map_x0, &in_full[stage], p0 + u * 64, 0, g);
          if (args.nsrc > 1)
            tma_load_3d(dst + (size_t)u * K1 * 128 + (size_t)args.k_src[0] * 128, &map_x1, &in_full[stage],
                        p0 + u * 64, 0, g);
        }
      };
      const uint32_t idesc1 = make_idesc(Elem<T>::kFmt, /*A MN-major*/ 1, /*B K-major*/ 0, kTileM, COUT);
      const uint32_t idesc2 = make_idesc(Elem<T>::kFmt, 0, 0, kTileM, COUT);
      uint32_t ph_in_full[kStages] = {0, 0}, ph_in_empty[kStages] = {0, 0};
      uint32_t ph_w1 = 0, ph_h = 0;
      int cur_g = -1;
      // save/restore of the walker state around look-ahead loads
      issue_load(t_begin, 0);
      int g_cur_tile = g;
      long gbase_cur = gbase, gtiles_cur = gtiles;
      if (depth > 1) mbar_wait(wh_full, 0);
      for (long t = t_begin; t < t_end; ++t) {
        const int stage = (int)((t - t_begin) % kStages);
        // restore walker for the current tile, then prefetch the next one
        g = g_cur_tile; gbase = gbase_cur; gtiles = gtiles_cur;
        seek(t);
        const int tg = g;
        g_cur_tile = g; gbase_cur = gbase; gtiles_cur = gtiles;
        if (tg != cur_g) {  // (re)load this graph's folded first-layer weights
          mbar_arrive_expect_tx(w1_full, (uint32_t)(K1g * COUT * 2));
          for (int at = 0; at < K1g / 64; ++at)
            tma_load_3d(s_w1 + (size_t)at * COUT * 128, &map_w1, w1_full, at * 64, 0, tg);
          mbar_wait(w1_full, ph_w1);
          ph_w1 ^= 1;
          cur_g = tg;
        }
        if (t + 1 < t_end) {
          const int ns = (int)((t + 1 - t_begin) % kStages);
          if (t + 1 - t_begin >= kStages) {  // stage was used before: wait for its MMAs to retire
            mbar_wait(&in_empty[ns], ph_in_empty[ns]);
            ph_in_empty[ns] ^= 1;
          }
          issue_load(t + 1, ns);
        }
        // ---- layer 1: SS MMA over K1 ----
        mbar_wait(&in_full[stage], ph_in_full[stage]);
        ph_in_full[stage] ^= 1;
        tc_fence_after();
        const uint32_t a_base = smem_u32(s_in + (size_t)stage * stage_bytes);
        const uint32_t w1_base = smem_u32(s_w1);
        for (int s = 0; s < K1 / 16; ++s) {
          const uint64_t ad = smem_desc_sw128(a_base + (uint32_t)s * 2048u, (uint32_t)K1 * 128u, 1024u);
          const uint64_t bd = smem_desc_sw128(w1_base + (uint32_t)(s / 4) * (COUT * 128u) + (uint32_t)(s % 4) * 32u,
                                              16u, 1024u);
          mma_ss(tmem_base, ad, bd, idesc1, s > 0 ? 1u : 0u);
        }
        mma_commit(&in_empty[stage]);
        mma_commit(mma_done);
        // ---- layers 2..depth: A = packed hidden activations in TMEM ----
        for (int l = 1; l < depth; ++l) {
          mbar_wait(h_ready, ph_h);
          ph_h ^= 1;
          tc_fence_after();
          const uint32_t wl = smem_u32(s_wh + (size_t)(l - 1) * Kh * COUT * 2);
          for (int s = 0; s < COUT / 16; ++s) {
            const uint64_t bd = smem_desc_sw128(wl + (uint32_t)(s / 4) * (COUT * 128u) + (uint32_t)(s % 4) * 32u,
                                                16u, 1024u);
            mma_ts(tmem_base, tmem_base + kHCol + (uint32_t)s * 8u, bd, idesc2, s > 0 ? 1u : 0u);
          }
          mma_commit(mma_done);
        }
        // accumulator must be drained before the next tile overwrites it
        mbar_wait(h_ready, ph_h);
        ph_h ^= 1;
      }
    }
  } else {
    // ================= epilogue warps (TMEM lane quadrant = warp % 4) =================
    const int quad = warp % 4;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int pix_in_tile = quad * 32 + lane;
    uint32_t ph_mma = 0;
    int g = 0;
    long gbase = 0;
    long gtiles = ((long)rows_cover(args.n_per_graph, 0, args.N) * args.NP + kTileM - 1) / kTileM;
    for (long t = t_begin; t < t_end; ++t) {
      while (t >= gbase + gtiles) {
        gbase += gtiles;
        ++g;
        gtiles = ((long)rows_cover(args.n_per_graph, g, args.N) * args.NP + kTileM - 1) / kTileM;
      }
      const long p = (t - gbase) * kTileM + pix_in_tile;
      const int n = graph_n(args.n_per_graph, g, args.N);
      const int pi = (int)(p / args.NP), pj = (int)(p % args.NP);
      const bool in_plane = p < args.Ppl;
      const bool valid = in_plane && pi < n && pj < n;
      for (int l = 0; l < depth; ++l) {
        mbar_wait(mma_done, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        const bool last = (l == depth - 1);
        const float* bias = (l == 0) ? (args.bias1 + (long)g * COUT) : args.bias[l];
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(lane_addr + (uint32_t)c0, r);
          tmem_wait_ld();
          if (!last) {
            uint32_t h[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float2 b2 = __ldg(reinterpret_cast<const float2*>(bias + c0 + 2 * u));
              float v0 = fmaxf(__uint_as_float(r[2 * u]) + b2.x, 0.f);
              float v1 = fmaxf(__uint_as_float(r[2 * u + 1]) + b2.y, 0.f);
              h[u] = Elem<T>::pack(v0, v1);
            }
            tmem_st16(lane_addr + (uint32_t)kHCol + (uint32_t)(c0 / 2), h);
          } else if (in_plane) {
            T* op = args.out + ((long)g * COUT + c0) * args.Ppl + p;
#pragma unroll
            for (int u = 0; u < 32; ++u)
              op[(long)u * args.Ppl] = Elem<T>::from_float(valid ? __uint_as_float(r[u]) : 0.f);
          }
        }
        if (!last) tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(h_ready);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =============================================================================================
// K_B: batched per-(graph, channel) N x N matmul with the GraphNorm rank-1 corrections in the
// epilogue.  out = a1 a2 (Y1 Y2) + a1 s2 r1 1^T + s1 a2 1 c2^T + s1 s2 n.
//   A = Y1 plane, K-major (rows i, K = k contiguous); B = Y2 plane, MN-major (rows k, N = j contiguous)
//   tile 128 x BN, K step 64, kNumStages-deep TMA ring, two TMEM accumulator stages.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.
// =============================================================================================
template <typename T>
struct MatmulArgs {
  int G, C, N, NP;
  T* out;                 // [G*C][N][NP]
  const float* coef_a;    // [G*C][2] (a1, s1) or null
  const float* coef_b;    // [G*C][2] (a2, s2) or null
  const float* r1;        // [G*C][N] row sums of Y1 or null
  const float* c2;        // [G*C][N] column sums of Y2 or null
  const int32_t* n_per_graph;
};

template <int BN>
struct MatmulCfg {
  static constexpr int kStageBytes = 128 * 128 + BN * 128;  // A: 128 rows x 64 k, B: 64 k x BN
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr size_t kSmemBytes = 1024 + (size_t)kStages * kStageBytes + 2 * BN * sizeof(float) + 256;
};

struct TileWalker {
  int g = 0;
  long base = 0;
  int mt = 0, nt = 0;
  long tiles_g = 0;
  template <int BN>
  __device__ void init(const int32_t* npg, int N, int C) {
    g = 0;
    base = 0;
    set<BN>(npg, N, C);
  }
  template <int BN>
  __device__ void set(const int32_t* npg, int N, int C) {
    int n = graph_n(npg, g, N);
    mt = (n + 127) / 128;
    nt = (n + BN - 1) / BN;
    tiles_g = (long)mt * nt * C;
  }
  // -> plane q, row tile m, col tile nn for flat tile t (t must not decrease between calls)
  template <int BN>
  __device__ void locate(long t, const int32_t* npg, int N, int C, int& q, int& m, int& nn, int& n) {
    while (t >= base + tiles_g) {
      base += tiles_g;
      ++g;
      set<BN>(npg, N, C);
    }
    long local = t - base;
    int per_plane = mt * nt;
    int c = (int)(local / per_plane);
    int rem = (int)(local % per_plane);
    q = g * C + c;
    m = rem / nt;
    nn = rem % nt;
    n = graph_n(npg, g, N);
  }
};

template <typename T, int BN>
__global__ void __launch_bounds__(192, 1)
tc_matmul_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const MatmulArgs<T> args) {
  using Cfg = MatmulCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  constexpr uint32_t kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_cc = reinterpret_cast<float*>(smem + (size_t)kStages * Cfg::kStageBytes);  // [2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_cc + 2 * BN);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  long total = 0;
  for (int g = 0; g < args.G; ++g) {
    int n = graph_n(args.n_per_graph, g, args.N);
    total += (long)((n + 127) / 128) * ((n + BN - 1) / BN) * args.C;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer =================
      prefetch_tensormap(&map_a);
      prefetch_tensormap(&map_b);
      TileWalker tw;
      tw.init<BN>(args.n_per_graph, args.N, args.C);
      int stage = 0;
      uint32_t phase = 0;
      for (long t = blockIdx.x; t < total; t += gridDim.x) {
        int q, m, nn, n;
        tw.locate<BN>(t, args.n_per_graph, args.N, args.C, q, m, nn, n);
        const int kts = (n + 63) / 64;
        for (int kt = 0; kt < kts; ++kt) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * Cfg::kStageBytes;
          uint8_t* sb = sa + 128 * 128;
          mbar_arrive_expect_tx(&full[stage], (uint32_t)Cfg::kStageBytes);
          tma_load_3d(sa, &map_a, &full[stage], kt * 64, m * 128, q);
          for (int u = 0; u < BN / 64; ++u)
            tma_load_3d(sb + (size_t)u * 8192, &map_b, &full[stage], nn * BN + u * 64, kt * 64, q);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer =================
      const uint32_t idesc = make_idesc(Elem<T>::kFmt, /*A K-major*/ 0, /*B MN-major*/ 1, 128, BN);
      TileWalker tw;
      tw.init<BN>(args.n_per_graph, args.N, args.C);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (long t = blockIdx.x; t < total; t += gridDim.x) {
        int q, m, nn, n;
        tw.locate<BN>(t, args.n_per_graph, args.N, args.C, q, m, nn, n);
        const int kts = (n + 63) / 64;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kt = 0; kt < kts; ++kt) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * Cfg::kStageBytes);
          const uint32_t sb = sa + 128 * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = smem_desc_sw128(sa + (uint32_t)k * 32u, 16u, 1024u);
            const uint64_t bd = smem_desc_sw128(sb + (uint32_t)k * 2048u, 8192u, 1024u);
            mma_ss(d_tmem, ad, bd, idesc, (kt > 0 || k > 0) ? 1u : 0u);
          }
          mma_commit(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        mma_commit(&tmem_full[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ================= epilogue (warps 2..5, TMEM lane quadrant = warp % 4) =================
    const int quad = warp % 4;
    const int et = threadIdx.x - 64;  // 0..127
    TileWalker tw;
    tw.init<BN>(args.n_per_graph, args.N, args.C);
    int as = 0;
    uint32_t aphase = 0;
    for (long t = blockIdx.x; t < total; t += gridDim.x) {
      int q, m, nn, n;
      tw.locate<BN>(t, args.n_per_graph, args.N, args.C, q, m, nn, n);
      float a1 = 1.f, s1 = 0.f, a2 = 1.f, s2 = 0.f;
      if (args.coef_a) { a1 = args.coef_a[2 * q]; s1 = args.coef_a[2 * q + 1]; }
      if (args.coef_b) { a2 = args.coef_b[2 * q]; s2 = args.coef_b[2 * q + 1]; }
      const int row = m * 128 + quad * 32 + lane;
      const int j0 = nn * BN;
      // per-column correction terms for this tile
      float* cc = s_cc + as * BN;
      for (int col = et; col < BN; col += 128) {
        int j = j0 + col;
        cc[col] = (args.c2 && j < n) ? s1 * a2 * args.c2[(long)q * args.N + j] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const float scale = a1 * a2;
      float rc = s1 * s2 * (float)n;
      if (args.r1 && row < n) rc += a1 * s2 * args.r1[(long)q * args.N + row];
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);
      T* orow = args.out + ((long)q * args.N + row) * args.NP + j0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_wait_ld();
        if (row < n && j0 + c0 < args.NP) {
          uint32_t pk[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            float v0 = fmaf(scale, __uint_as_float(r[2 * u]), rc + cc[c0 + 2 * u]);
            float v1 = fmaf(scale, __uint_as_float(r[2 * u + 1]), rc + cc[c0 + 2 * u + 1]);
            pk[u] = Elem<T>::pack(v0, v1);
          }
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (j0 + c0 + v * 8 < args.NP)  // NP is a multiple of 8: 16-byte vectors never straddle the pitch
              *reinterpret_cast<uint4*>(orow + c0 + v * 8) = make_uint4(pk[4 * v], pk[4 * v + 1], pk[4 * v + 2], pk[4 * v + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =============================================================================================
// host side
// =============================================================================================
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

// 3-D tensor map over 16-bit data: dims (d0 contiguous, d1, d2), strides in elements, 128B swizzle.
int make_map3(CUtensorMap* m, int fmt_is_bf16, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
              uint64_t stride1_elems, uint64_t stride2_elems, uint32_t b0, uint32_t b1) {
  EncodeFn enc = get_encode();
  if (!enc) return fail(FGNN_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_elems * 2, stride2_elems * 2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15))
    return fail(FGNN_ERR_INVALID, "tensor map alignment: base %p strides %llu %llu", base,
                (unsigned long long)strides[0], (unsigned long long)strides[1]);
  CUresult r = enc(m, fmt_is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(FGNN_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) dims %llu,%llu,%llu box %u,%u", (int)r,
                (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, b0, b1);
  return FGNN_OK;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

// ---- launchers -------------------------------------------------------------------------------
template <typename T>
int launch_matmul(const T* y1, const T* y2, T* out, const float* coef_a, const float* coef_b, const float* r1,
                  const float* c2, int G, int C, int N, int NP, const int32_t* npg, cudaStream_t st) {
  constexpr int is_bf16 = Elem<T>::kFmt;
  const int BN = (N <= 64) ? 64 : (N <= 128 ? 128 : 256);
  CUtensorMap ma, mb;
  const uint64_t planes = (uint64_t)G * C;
  if (int e = make_map3(&ma, is_bf16, y1, NP, N, planes, NP, (uint64_t)N * NP, 64, 128)) return e;
  if (int e = make_map3(&mb, is_bf16, y2, NP, N, planes, NP, (uint64_t)N * NP, 64, 64)) return e;
  MatmulArgs<T> a{G, C, N, NP, out, coef_a, coef_b, r1, c2, npg};
  const int grid = num_sms();
#define FGNN_MM_LAUNCH(BNV)                                                                              \
  do {                                                                                                   \
    static bool attr = false;                                                                            \
    if (!attr) {                                                                                         \
      FGNN_CUDA(cudaFuncSetAttribute(tc_matmul_kernel<T, BNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)MatmulCfg<BNV>::kSmemBytes));                                  \
      attr = true;                                                                                       \
    }                                                                                                    \
    tc_matmul_kernel<T, BNV><<<grid, 192, MatmulCfg<BNV>::kSmemBytes, st>>>(ma, mb, a);                   \
  } while (0)
  prof::begin(prof::kMatmul, st);
  if (BN == 64) FGNN_MM_LAUNCH(64);
  else if (BN == 128) FGNN_MM_LAUNCH(128);
  else FGNN_MM_LAUNCH(256);
#undef FGNN_MM_LAUNCH
  prof::end(prof::kMatmul, st);
  FGNN_LAUNCHED();
  return FGNN_OK;
}

template <typename T>
struct MlpLaunch {
  const T* src[2];
  int c_src[2];
  int nsrc;
  const T* w1f;        // [G][COUT][K1g]
  const float* bias1;  // [G][COUT]
  const T* wh;         // [depth-1][COUT][Kh]
  const float* bias[FGNN_MAX_DEPTH];
  int depth, c_out;
  T* out;
};

template <typename T, int COUT>
int launch_mlp_t(const MlpLaunch<T>& L, int G, int N, int NP, const int32_t* npg, cudaStream_t st) {
  constexpr int is_bf16 = Elem<T>::kFmt;
  const long Ppl = (long)N * NP;
  MlpArgs<T> a{};
  a.G = G; a.N = N; a.NP = NP; a.Ppl = Ppl;
  a.nsrc = L.nsrc;
  a.k_src[0] = round_up(L.c_src[0], 16);
  a.k_src[1] = L.nsrc > 1 ? round_up(L.c_src[1], 16) : 0;
  a.K1 = a.k_src[0] + a.k_src[1];
  a.K1g = round_up(a.K1, 64);
  a.depth = L.depth;
  a.Kh = COUT < 64 ? 64 : COUT;
  a.bias1 = L.bias1;
  for (int l = 0; l < L.depth; ++l) a.bias[l] = L.bias[l];
  a.out = L.out;
  a.n_per_graph = npg;
  FGNN_CHECK_ARG(a.K1 <= 256, "first-layer K=%d too wide for the tensor-core MLP kernel", a.K1);
  CUtensorMap mx0, mx1, mw1, mwh;
  if (int e = make_map3(&mx0, is_bf16, L.src[0], Ppl, L.c_src[0], G, Ppl, (uint64_t)L.c_src[0] * Ppl, 64, a.k_src[0])) return e;
  if (L.nsrc > 1) {
    if (int e = make_map3(&mx1, is_bf16, L.src[1], Ppl, L.c_src[1], G, Ppl, (uint64_t)L.c_src[1] * Ppl, 64, a.k_src[1])) return e;
  } else {
    mx1 = mx0;
  }
  if (int e = make_map3(&mw1, is_bf16, L.w1f, a.K1g, COUT, G, a.K1g, (uint64_t)COUT * a.K1g, 64, COUT)) return e;
  if (L.depth > 1) {
    if (int e = make_map3(&mwh, is_bf16, L.wh, a.Kh, COUT, L.depth - 1, a.Kh, (uint64_t)COUT * a.Kh, 64, COUT)) return e;
  } else {
    mwh = mw1;
  }
  const size_t smem = MlpSmem<COUT>::bytes(a.K1, a.K1g, a.depth, a.Kh);
  FGNN_CHECK_ARG(smem <= 227 * 1024, "MLP kernel needs %zu bytes of shared memory", smem);
  static size_t attr_bytes = 0;
  if (smem > attr_bytes) {
    FGNN_CUDA(cudaFuncSetAttribute(tc_mlp_kernel<T, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_bytes = smem;
  }
  const int ctas_per_sm = env_int("FGNN_MLP_CTAS_PER_SM", 2);
  long total_tiles = (long)G * ((Ppl + kTileM - 1) / kTileM);
  int grid = (int)std::min<long>((long)num_sms() * ctas_per_sm, total_tiles);
  if (grid < 1) grid = 1;
  prof::begin(prof::kMlp, st);
  tc_mlp_kernel<T, COUT><<<grid, 160, smem, st>>>(mx0, mx1, mw1, mwh, a);
  prof::end(prof::kMlp, st);
  FGNN_LAUNCHED();
  return FGNN_OK;
}

template <typename T>
int launch_mlp(const MlpLaunch<T>& L, int G, int N, int NP, const int32_t* npg, cudaStream_t st) {
  if (L.c_out == 32) return launch_mlp_t<T, 32>(L, G, N, NP, npg, st);
  if (L.c_out == 64) return launch_mlp_t<T, 64>(L, G, N, NP, npg, st);
  return fail(FGNN_ERR_UNSUPPORTED, "tensor-core path supports out_features 32 or 64 (got %d); use FGNN_FP32", L.c_out);
}

template <typename T>
int launch_stats(const T* y, const float* gw, const float* gb, float eps, float* coef, float* rsum, float* csum,
                 int G, int C, int N, int NP, const int32_t* npg, cudaStream_t st) {
  prof::begin(prof::kStats, st);
  plane_stats16_kernel<T><<<G * C, 256, (size_t)8 * NP * sizeof(float), st>>>(y, gw, gb, eps, coef, rsum, csum, C, N,
                                                                              NP, npg);
  prof::end(prof::kStats, st);
  FGNN_LAUNCHED();
  return FGNN_OK;
}

// ---- workspace plan -----------------------------------------------------------------------------
struct Plan {
  int chunk, C, cin0, N, NP, depth_max;
  long Ppl;
  int K1g12_max, K1g3_max;
};

int make_plan(const fgnn_embed_params& p, int G, int N, Plan& pl) {
  pl.N = N;
  pl.NP = round_up(N, 8);
  pl.Ppl = (long)N * pl.NP;
  pl.C = p.block[0].mlp1.c_out;
  pl.cin0 = p.block[0].mlp1.c_in;
  pl.depth_max = 1;
  pl.K1g12_max = 64;
  pl.K1g3_max = 64;
  int cur = pl.cin0;
  for (int b = 0; b < p.num_blocks; ++b) {
    const fgnn_block_params& bp = p.block[b];
    if (bp.mlp1.c_out != pl.C || bp.mlp2.c_out != pl.C || bp.mlp3.c_out != pl.C)
      return fail(FGNN_ERR_UNSUPPORTED, "tensor-core path needs in_features == out_features for every block");
    if (bp.mlp1.c_in != cur || bp.mlp2.c_in != cur || bp.mlp3.c_in != cur + pl.C)
      return fail(FGNN_ERR_INVALID, "block %d: channel counts do not chain", b);
    pl.depth_max = std::max(pl.depth_max, std::max(bp.mlp1.depth, std::max(bp.mlp2.depth, bp.mlp3.depth)));
    pl.K1g12_max = std::max(pl.K1g12_max, round_up(round_up(cur, 16), 64));
    pl.K1g3_max = std::max(pl.K1g3_max, round_up(pl.C + round_up(cur, 16), 64));
    cur = pl.C;
  }
  if (pl.C != 32 && pl.C != 64)
    return fail(FGNN_ERR_UNSUPPORTED, "tensor-core path supports in/out_features 32 or 64 (got %d); use FGNN_FP32", pl.C);
  if (N > kMaxN) return fail(FGNN_ERR_UNSUPPORTED, "tensor-core path supports N <= %d (got %d)", kMaxN, N);
  if (pl.cin0 > 64) return fail(FGNN_ERR_UNSUPPORTED, "original_features_num %d > 64 unsupported", pl.cin0);
  const int chunk_env = env_int("FGNN_TC_CHUNK", 0);
  long per_graph = (long)(5 * pl.C + pl.cin0) * pl.Ppl * 2;
  long budget = (long)6 << 30;
  long chunk = chunk_env > 0 ? chunk_env : std::max<long>(1, budget / std::max<long>(per_graph, 1));
  chunk = std::min<long>(chunk, 65535 / std::max(pl.C, pl.cin0));   // grid.y limits of the helper kernels
  pl.chunk = (int)std::min<long>(G, chunk);
  return FGNN_OK;
}

struct Buffers {
  void *xin, *xa, *xb, *y1, *y2, *mult;        // 16-bit planes
  float *coef1, *coef2, *coef3a, *coef3b;      // [chunk][C][2]
  float *r1, *c2, *scratch_rc;                 // [chunk][C][N]
  void *wf1, *wf2, *wf3;                       // folded first-layer weights
  float *bf1, *bf2, *bf3;                      // folded first-layer biases
  void* wh;                                    // [blocks][3][depth-1][C][Kh]
};

size_t carve(const Plan& pl, int num_blocks, Arena& ar, Buffers& B) {
  const size_t act = (size_t)pl.chunk * pl.C * pl.Ppl;
  B.xin = ar.take<uint16_t>((size_t)pl.chunk * pl.cin0 * pl.Ppl, 1024);
  B.xa = ar.take<uint16_t>(act, 1024);
  B.xb = ar.take<uint16_t>(act, 1024);
  B.y1 = ar.take<uint16_t>(act, 1024);
  B.y2 = ar.take<uint16_t>(act, 1024);
  B.mult = ar.take<uint16_t>(act, 1024);
  const size_t nc = (size_t)pl.chunk * pl.C;
  B.coef1 = ar.take<float>(nc * 2);
  B.coef2 = ar.take<float>(nc * 2);
  B.coef3a = ar.take<float>(nc * 2);
  B.coef3b = ar.take<float>(nc * 2);
  B.r1 = ar.take<float>(nc * pl.N);
  B.c2 = ar.take<float>(nc * pl.N);
  B.scratch_rc = ar.take<float>(nc * pl.N);
  B.wf1 = ar.take<uint16_t>(nc * pl.K1g12_max, 1024);
  B.wf2 = ar.take<uint16_t>(nc * pl.K1g12_max, 1024);
  B.wf3 = ar.take<uint16_t>(nc * pl.K1g3_max, 1024);
  B.bf1 = ar.take<float>(nc);
  B.bf2 = ar.take<float>(nc);
  B.bf3 = ar.take<float>(nc);
  const int Kh = pl.C < 64 ? 64 : pl.C;
  B.wh = ar.take<uint16_t>((size_t)num_blocks * 3 * std::max(pl.depth_max - 1, 1) * pl.C * Kh, 1024);
  return align_up(ar.off, 1024);
}

template <typename T>
int embed_fwd_t(const fgnn_embed_params& p, const float* x, float* emb, int G, int N, const int32_t* npg, void* ws,
                size_t ws_bytes, cudaStream_t st) {
  Plan pl;
  if (int e = make_plan(p, G, N, pl)) return e;
  Arena ar(ws, ws_bytes);
  Buffers B;
  size_t need = carve(pl, p.num_blocks, ar, B);
  if (need > ws_bytes) return fail(FGNN_ERR_WORKSPACE, "embed workspace too small: %zu < %zu", ws_bytes, need);
  const int C = pl.C, NP = pl.NP;
  const int Kh = C < 64 ? 64 : C;
  const int dm1 = std::max(pl.depth_max - 1, 1);
  // hidden-layer weights -> 16-bit, once per call
  for (int b = 0; b < p.num_blocks; ++b) {
    const fgnn_mlp_params* mlps[3] = {&p.block[b].mlp1, &p.block[b].mlp2, &p.block[b].mlp3};
    for (int j = 0; j < 3; ++j)
      for (int l = 1; l < mlps[j]->depth; ++l) {
        T* dst = reinterpret_cast<T*>(B.wh) + (((size_t)b * 3 + j) * dm1 + (l - 1)) * C * Kh;
        convert_weight_kernel<T><<<ceil_div(C * Kh, 256), 256, 0, st>>>(mlps[j]->w[l], dst, C, C, Kh);
        FGNN_LAUNCHED();
      }
  }
  auto run_mlp = [&](const fgnn_mlp_params& mp, int bidx, int j, const T* s0, int c0, const float* coef0, const T* s1,
                     int c1, const float* coef1, T* wf, float* bf, T* out, int gc, const int32_t* n_c) -> int {
    FoldArgs fa{};
    fa.w = mp.w[0];
    fa.b = mp.b[0];
    fa.nsrc = s1 ? 2 : 1;
    fa.c[0] = c0; fa.c[1] = c1;
    fa.coef[0] = coef0; fa.coef[1] = coef1;
    fa.koff[0] = 0; fa.koff[1] = round_up(c0, 16);
    fa.c_out = C;
    fa.K1g = round_up(round_up(c0, 16) + (s1 ? round_up(c1, 16) : 0), 64);
    fold_weights_kernel<T><<<gc, 256, 0, st>>>(fa, wf, bf);
    FGNN_LAUNCHED();
    MlpLaunch<T> L{};
    L.src[0] = s0; L.src[1] = s1;
    L.c_src[0] = c0; L.c_src[1] = c1;
    L.nsrc = fa.nsrc;
    L.w1f = wf;
    L.bias1 = bf;
    L.wh = reinterpret_cast<const T*>(B.wh) + ((size_t)bidx * 3 + j) * dm1 * C * Kh;
    for (int l = 0; l < mp.depth; ++l) L.bias[l] = mp.b[l];
    L.depth = mp.depth;
    L.c_out = C;
    L.out = out;
    return launch_mlp<T>(L, gc, N, NP, n_c, st);
  };
  for (int g0 = 0; g0 < G; g0 += pl.chunk) {
    const int gc = std::min(pl.chunk, G - g0);
    const int32_t* n_c = npg ? npg + g0 : nullptr;
    {
      dim3 grid((unsigned)std::min<long>(64, (pl.Ppl + 255) / 256), gc * pl.cin0);
      to_planes_kernel<T><<<grid, 256, 0, st>>>(x + (size_t)g0 * pl.cin0 * N * N, reinterpret_cast<T*>(B.xin), pl.cin0,
                                                N, NP, n_c);
      FGNN_LAUNCHED();
    }
    const T* cur = reinterpret_cast<const T*>(B.xin);
    int cur_c = pl.cin0;
    const float* cur_coef = nullptr;
    T* nxt = reinterpret_cast<T*>(B.xa);
    float* nxt_coef = B.coef3a;
    T* y1 = reinterpret_cast<T*>(B.y1);
    T* y2 = reinterpret_cast<T*>(B.y2);
    T* mult = reinterpret_cast<T*>(B.mult);
    for (int b = 0; b < p.num_blocks; ++b) {
      const fgnn_block_params& bp = p.block[b];
      if (int e = run_mlp(bp.mlp1, b, 0, cur, cur_c, cur_coef, nullptr, 0, nullptr, reinterpret_cast<T*>(B.wf1), B.bf1,
                          y1, gc, n_c)) return e;
      if (int e = run_mlp(bp.mlp2, b, 1, cur, cur_c, cur_coef, nullptr, 0, nullptr, reinterpret_cast<T*>(B.wf2), B.bf2,
                          y2, gc, n_c)) return e;
      if (int e = launch_stats<T>(y1, bp.mlp1.gn_w, bp.mlp1.gn_b, bp.mlp1.eps, B.coef1, B.r1, nullptr, gc, C, N, NP, n_c, st)) return e;
      if (int e = launch_stats<T>(y2, bp.mlp2.gn_w, bp.mlp2.gn_b, bp.mlp2.eps, B.coef2, nullptr, B.c2, gc, C, N, NP, n_c, st)) return e;
      if (int e = launch_matmul<T>(y1, y2, mult, B.coef1, B.coef2, B.r1, B.c2, gc, C, N, NP, n_c, st)) return e;
      if (int e = run_mlp(bp.mlp3, b, 2, mult, C, nullptr, cur, cur_c, cur_coef, reinterpret_cast<T*>(B.wf3), B.bf3, nxt,
                          gc, n_c)) return e;
      if (int e = launch_stats<T>(nxt, bp.mlp3.gn_w, bp.mlp3.gn_b, bp.mlp3.eps, nxt_coef, nullptr, nullptr, gc, C, N, NP, n_c, st)) return e;
      cur = nxt;
      cur_c = C;
      cur_coef = nxt_coef;
      nxt = (nxt == reinterpret_cast<T*>(B.xa)) ? reinterpret_cast<T*>(B.xb) : reinterpret_cast<T*>(B.xa);
      nxt_coef = (nxt_coef == B.coef3a) ? B.coef3b : B.coef3a;
    }
    const long rows = (long)gc * C * N;
    pool_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(cur, cur_coef, emb + (size_t)g0 * C * N, C, N, NP, rows, n_c);
    FGNN_LAUNCHED();
  }
  return FGNN_OK;
}

}  // namespace

size_t embed_workspace_bytes(const fgnn_embed_params& p, int G, int N) {
  Plan pl;
  if (make_plan(p, G, N, pl)) return 0;
  Arena ar(nullptr, 0);
  Buffers B;
  return carve(pl, p.num_blocks, ar, B);
}

int embed_fwd(const fgnn_embed_params& p, int precision, const float* x, float* emb, int G, int N,
              const int32_t* n_per_graph, const int32_t* /*n_per_graph_host*/, void* ws, size_t ws_bytes,
              cudaStream_t st) {
  if (!fgnn_device_supports_tcgen05())
    return fail(FGNN_ERR_UNSUPPORTED, "FGNN_BF16/FGNN_FP16 need an sm_100 device (tcgen05); there is no fallback");
  if (reinterpret_cast<uintptr_t>(ws) & 1023) return fail(FGNN_ERR_INVALID, "workspace must be 1024-byte aligned");
  if (precision == FGNN_BF16) return embed_fwd_t<__nv_bfloat16>(p, x, emb, G, N, n_per_graph, ws, ws_bytes, st);
  return embed_fwd_t<__half>(p, x, emb, G, N, n_per_graph, ws, ws_bytes, st);
}

// ---- debug: one tensor-core matmul on fp32 host-layout tensors ---------------------------------
size_t debug_matmul_workspace_bytes(int G, int C, int N) {
  const int NP = round_up(N, 8);
  return align_up((size_t)3 * G * C * N * NP * 2 + 4096, 1024);
}

template <typename T>
int debug_matmul_t(const float* a, const float* b, float* out, int G, int C, int N, const int32_t* npg, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  const int NP = round_up(N, 8);
  const size_t act = (size_t)G * C * N * NP;
  if (ws_bytes < debug_matmul_workspace_bytes(G, C, N)) return fail(FGNN_ERR_WORKSPACE, "debug workspace too small");
  Arena ar(ws, ws_bytes);
  T* y1 = ar.take<T>(act, 1024);
  T* y2 = ar.take<T>(act, 1024);
  T* mo = ar.take<T>(act, 1024);
  const long Ppl = (long)N * NP;
  dim3 grid((unsigned)std::min<long>(64, (Ppl + 255) / 256), G * C);
  to_planes_kernel<T><<<grid, 256, 0, st>>>(a, y1, C, N, NP, npg);
  FGNN_LAUNCHED();
  to_planes_kernel<T><<<grid, 256, 0, st>>>(b, y2, C, N, NP, npg);
  FGNN_LAUNCHED();
  if (int e = launch_matmul<T>(y1, y2, mo, nullptr, nullptr, nullptr, nullptr, G, C, N, NP, npg, st)) return e;
  from_planes_kernel<T><<<grid, 256, 0, st>>>(mo, out, nullptr, C, N, NP, npg);
  FGNN_LAUNCHED();
  return FGNN_OK;
}

int debug_matmul(int precision, const float* a, const float* b, float* out, int G, int C, int N,
                 const int32_t* n_per_graph, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!fgnn_device_supports_tcgen05()) return fail(FGNN_ERR_UNSUPPORTED, "needs an sm_100 device");
  FGNN_CHECK_ARG(a && b && out && ws, "null pointer");
  FGNN_CHECK_ARG(N <= kMaxN, "N too large");
  if (precision == FGNN_BF16) return debug_matmul_t<__nv_bfloat16>(a, b, out, G, C, N, n_per_graph, ws, ws_bytes, st);
  if (precision == FGNN_FP16) return debug_matmul_t<__half>(a, b, out, G, C, N, n_per_graph, ws, ws_bytes, st);
  return fail(FGNN_ERR_INVALID, "precision must be FGNN_BF16 or FGNN_FP16");
}

// ---- debug: one tensor-core MlpBlock_Real (fold -> conv chain -> stats -> normalise) ----------------
size_t debug_mlp_workspace_bytes(int G, int c_in, int c_out, int depth, int N) {
  const int NP = round_up(N, 8);
  const size_t Ppl = (size_t)N * NP;
  const int K1g = round_up(round_up(c_in, 16), 64);
  const int Kh = c_out < 64 ? 64 : c_out;
  return align_up((size_t)G * (c_in + c_out) * Ppl * 2 + (size_t)G * c_out * K1g * 2 +
                      (size_t)std::max(depth - 1, 1) * c_out * Kh * 2 + (size_t)G * c_out * 3 * 4 + 16384, 1024);
}

template <typename T>
int debug_mlp_t(const fgnn_mlp_params& mp, const float* x, float* y, int G, int N, const int32_t* npg, void* ws,
                size_t ws_bytes, cudaStream_t st) {
  const int NP = round_up(N, 8);
  const size_t Ppl = (size_t)N * NP;
  const int C = mp.c_out;
  const int K1g = round_up(round_up(mp.c_in, 16), 64);
  const int Kh = C < 64 ? 64 : C;
  if (ws_bytes < debug_mlp_workspace_bytes(G, mp.c_in, C, mp.depth, N)) return fail(FGNN_ERR_WORKSPACE, "debug workspace too small");
  FGNN_CHECK_ARG(mp.c_in <= 128, "c_in too large for the debug entry");
  Arena ar(ws, ws_bytes);
  T* xin = ar.take<T>((size_t)G * mp.c_in * Ppl, 1024);
  T* out = ar.take<T>((size_t)G * C * Ppl, 1024);
  T* wf = ar.take<T>((size_t)G * C * K1g, 1024);
  T* wh = ar.take<T>((size_t)std::max(mp.depth - 1, 1) * C * Kh, 1024);
  float* bf = ar.take<float>((size_t)G * C);
  float* coef = ar.take<float>((size_t)G * C * 2);
  {
    dim3 grid((unsigned)std::min<size_t>(64, (Ppl + 255) / 256), G * mp.c_in);
    to_planes_kernel<T><<<grid, 256, 0, st>>>(x, xin, mp.c_in, N, NP, npg);
    FGNN_LAUNCHED();
  }
  for (int l = 1; l < mp.depth; ++l) {
    convert_weight_kernel<T><<<ceil_div(C * Kh, 256), 256, 0, st>>>(mp.w[l], wh + (size_t)(l - 1) * C * Kh, C, C, Kh);
    FGNN_LAUNCHED();
  }
  FoldArgs fa{};
  fa.w = mp.w[0]; fa.b = mp.b[0]; fa.nsrc = 1; fa.c[0] = mp.c_in; fa.koff[0] = 0; fa.c_out = C; fa.K1g = K1g;
  fold_weights_kernel<T><<<G, 256, 0, st>>>(fa, wf, bf);
  FGNN_LAUNCHED();
  MlpLaunch<T> L{};
  L.src[0] = xin; L.c_src[0] = mp.c_in; L.nsrc = 1; L.w1f = wf; L.bias1 = bf; L.wh = wh;
  for (int l = 0; l < mp.depth; ++l) L.bias[l] = mp.b[l];
  L.depth = mp.depth; L.c_out = C; L.out = out;
  if (int e = launch_mlp<T>(L, G, N, NP, npg, st)) return e;
  if (int e = launch_stats<T>(out, mp.gn_w, mp.gn_b, mp.eps, coef, nullptr, nullptr, G, C, N, NP, npg, st)) return e;
  dim3 grid((unsigned)std::min<size_t>(64, (Ppl + 255) / 256), G * C);
  from_planes_kernel<T><<<grid, 256, 0, st>>>(out, y, coef, C, N, NP, npg);
  FGNN_LAUNCHED();
  return FGNN_OK;
}

int debug_mlp(int precision, const fgnn_mlp_params& mp, const float* x, float* y, int G, int N,
              const int32_t* n_per_graph, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!fgnn_device_supports_tcgen05()) return fail(FGNN_ERR_UNSUPPORTED, "needs an sm_100 device");
  FGNN_CHECK_ARG(x && y && ws, "null pointer");
  FGNN_CHECK_ARG(N <= kMaxN, "N too large");
  if (precision == FGNN_BF16) return debug_mlp_t<__nv_bfloat16>(mp, x, y, G, N, n_per_graph, ws, ws_bytes, st);
  if (precision == FGNN_FP16) return debug_mlp_t<__half>(mp, x, y, G, N, n_per_graph, ws, ws_bytes, st);
  return fail(FGNN_ERR_INVALID, "precision must be FGNN_BF16 or FGNN_FP16");
}

}  // namespace tc
}  // namespace fgnn
