// Embedding-distance stage (K12-K15): Euclidean pdist for complete-linkage clustering and fused
// distance + top-k for the kNN actor classifier, both as one NT GEMM on tcgen05 tensor cores.
//
// Precision: the reference computes these distances in fp64 (scipy pdist / sklearn brute kNN) and
// its results depend on the ORDER of distances, so plain bf16 inputs are not enough.  Each fp32
// row is split into bf16 hi + bf16 lo (x = hi + lo + O(2^-17 |x|)) and the dot product is
//   hi.hi + hi.lo + lo.hi          (three K-segments accumulated into the same fp32 TMEM tile)
// which is fp32-grade.  Norms are fp32 of the original data, d^2 = |a|^2 + |b|^2 - 2 a.b, clamped at 0.
//
// Kernel skeleton = the conv kernel's: TMA producer warp -> smem ring (128-byte swizzle) -> one
// MMA-issuing thread -> double-buffered TMEM accumulators -> four epilogue warps.
#include <stdio.h>
#include <string.h>
#include <algorithm>

#include "host.h"
#include "common.cuh"

namespace frb {

constexpr int kDM = 128;       // tile rows (UMMA M)
constexpr int kDN = 128;       // tile cols (UMMA N)
constexpr int kDK = 64;        // K elements per stage
constexpr int kDStages = 4;
constexpr int kDStageBytes = (kDM + kDN) * kDK * 2;
constexpr int kDEpiWarps = 16;   // 4 warps per TMEM lane quarter, 32 tile columns each (the epilogue is issue-bound)
constexpr int kDSub = kDEpiWarps / 4;
constexpr int kDThreads = (kDEpiWarps + 2) * 32;
constexpr int kTopK = 16;      // per-row candidates kept in registers (k <= 16)

enum DistMode : int { DIST_SQUARE_F32 = 0, DIST_CONDENSED_F32 = 1, DIST_CONDENSED_F64 = 2, DIST_KNN = 3, DIST_SQUARE_SYM_F32 = 4 };

struct DistArgs {
  int mode;
  int rows, row0;        // A rows handled by this call: global rows [row0, row0+rows)
  int N;                 // B rows
  int Dp;                // padded feature length (multiple of 64); split row = [hi(Dp) | lo(Dp)]
  int m_tiles, n_tiles;
  int chunks, tiles_per_chunk;   // kNN: each (m_tile, chunk) item scans tiles_per_chunk column tiles
  int items;
  const float* na;       // |a|^2 indexed by global A row
  const float* nb;       // |b|^2
  void* out;
  float* part_d;         // kNN partial lists [Qrows, chunks, kDSub, kTopK]
  int32_t* part_i;
  int* watchdog;
};

// ------------------------------------------------------------------ fp32 -> (hi | lo) bf16 + norms
__global__ void split_rows_kernel(const float* __restrict__ x, int n, int D, int Dp, __nv_bfloat16* __restrict__ out,
                                  float* __restrict__ norm2) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < n; r += nwarps) {
    const float* row = x + static_cast<size_t>(r) * D;
    __nv_bfloat16* o = out + static_cast<size_t>(r) * 2 * Dp;
    float ss = 0.f;
    for (int c = lane; c < Dp; c += 32) {
      const float v = c < D ? row[c] : 0.f;
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      o[c] = hi;
      o[Dp + c] = lo;
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) norm2[r] = ss;
  }
}

// ------------------------------------------------------------------ item -> tile range
struct ItemRange {
  int m_tile, n_begin, n_end, chunk;
};
__device__ __forceinline__ ItemRange item_range(const DistArgs& p, int item) {
  ItemRange r;
  if (p.mode == DIST_KNN) {
    r.m_tile = item / p.chunks;
    r.chunk = item - r.m_tile * p.chunks;
    r.n_begin = r.chunk * p.tiles_per_chunk;
    r.n_end = min(r.n_begin + p.tiles_per_chunk, p.n_tiles);
  } else {
    r.m_tile = item / p.n_tiles;
    r.chunk = 0;
    r.n_begin = item - r.m_tile * p.n_tiles;
    r.n_end = r.n_begin + 1;
    // condensed layouts only hold j > i: drop tiles entirely on or below the diagonal
    if (p.mode == DIST_SQUARE_SYM_F32) {
      if ((r.n_begin + 1) * kDN <= p.row0 + r.m_tile * kDM) r.n_end = r.n_begin;   // keep tiles touching j >= i
    } else if (p.mode != DIST_SQUARE_F32 && (r.n_begin + 1) * kDN <= p.row0 + r.m_tile * kDM + 1) r.n_end = r.n_begin;
  }
  return r;
}

__device__ __forceinline__ void topk_insert(float (&d)[kTopK], int (&id)[kTopK], float v, int j) {
  if (!(v < d[kTopK - 1])) return;
#pragma unroll
  for (int t = kTopK - 1; t >= 1; --t) {
    const bool shift = v < d[t - 1];
    const bool place = !shift && v < d[t];
    d[t] = shift ? d[t - 1] : (place ? v : d[t]);
    id[t] = shift ? id[t - 1] : (place ? j : id[t]);
  }
  if (v < d[0]) { d[0] = v; id[0] = j; }
}

__global__ void __launch_bounds__(kDThreads, 1)
dist_gemm_kernel(const __grid_constant__ DistArgs p, const __grid_constant__ CUtensorMap tmA,
                 const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kDStages * kDM * kDK * 2;
  float* tr = reinterpret_cast<float*>(smem + kDStages * kDStageBytes);        // [kDEpiWarps][32][33]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tr + kDEpiWarps * 32 * 33);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kDStages;
  uint64_t* tfull_bar = bars + 2 * kDStages;
  uint64_t* tempty_bar = bars + 2 * kDStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kDStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 2 * kDN;

  if (warp == kDEpiWarps + 1 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == kDEpiWarps) {
    if (lane == 0) {
      for (int s = 0; s < kDStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tfull_bar[b], 1);
        mbar_init(&tempty_bar[b], kDEpiWarps);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<kTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int kb_per_seg = p.Dp / kDK;
  const int num_kb = 3 * kb_per_seg;

  if (warp < kDEpiWarps) {
    // ================================================================ epilogue
    const int quarter = warp & 3;            // TMEM lane quarter
    const int csub = warp >> 2;              // 32-column sub-block of the 128-column tile
    float* mytr = tr + warp * 32 * 33;
    int tile_iter = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const ItemRange ir = item_range(p, item);
      const int li = ir.m_tile * kDM + quarter * 32 + lane;   // local A row of this thread
      const int gi = p.row0 + li;                          // global A row
      const bool row_ok = li < p.rows;
      const float na = row_ok ? __ldg(p.na + gi) : 0.f;
      float td[kTopK];
      int tid[kTopK];
#pragma unroll
      for (int t = 0; t < kTopK; ++t) { td[t] = INFINITY; tid[t] = 0x7fffffff; }
      long long rowoff = 0;
      if (p.mode == DIST_SQUARE_F32 || p.mode == DIST_SQUARE_SYM_F32) rowoff = static_cast<long long>(li) * p.N;
      else rowoff = static_cast<long long>(gi) * p.N - (static_cast<long long>(gi) * (gi + 1)) / 2 - gi - 1;

      for (int n_tile = ir.n_begin; n_tile < ir.n_end; ++n_tile, ++tile_iter) {
        const int buf = tile_iter & 1;
        const uint32_t tphase = (tile_iter >> 1) & 1;
        mbar_wait(&tfull_bar[buf], tphase, p.watchdog, 100 + buf);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * kDN);
        const int n_base = n_tile * kDN;
        if (p.mode == DIST_KNN) {
          for (int n0 = csub * 32; n0 < csub * 32 + 32; n0 += 16) {
            uint32_t v[16];
            __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge after the data-dependent inserts
            tmem_ld16(taddr + n0, v);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const int j = n_base + n0 + c;
              if (j < p.N) {
                const float sc = fmaf(-2.f, __uint_as_float(v[c]), __ldg(p.nb + j));
                topk_insert(td, tid, sc, j);
              }
            }
          }
        } else {
          {
            const int n0 = csub * 32;
            uint32_t v[32];
            __syncwarp();
            tmem_ld16(taddr + n0, v);
            tmem_ld16(taddr + n0 + 16, v + 16);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int j = n_base + n0 + c;
              const float nb = j < p.N ? __ldg(p.nb + j) : 0.f;
              const float d2 = fmaf(-2.f, __uint_as_float(v[c]), na + nb);
              mytr[lane * 33 + c] = sqrtf(fmaxf(d2, 0.f));
            }
            __syncwarp();
            const int j = n_base + n0 + lane;
            for (int r = 0; r < 32; ++r) {
              const long long ro = __shfl_sync(0xffffffffu, rowoff, r);
              const int rli = ir.m_tile * kDM + quarter * 32 + r;
              const int rgi = p.row0 + rli;
              const bool ok = rli < p.rows && j < p.N &&
                              (p.mode == DIST_SQUARE_F32 || j > rgi || (p.mode == DIST_SQUARE_SYM_F32 && j == rgi));
              if (ok) {
                const float val = j == rgi ? 0.f : mytr[r * 33 + lane];   // exact zero diagonal (square layout)
                if (p.mode == DIST_CONDENSED_F64) static_cast<double*>(p.out)[ro + j] = static_cast<double>(val);
                else static_cast<float*>(p.out)[ro + j] = val;
              }
            }
            if (p.mode == DIST_SQUARE_SYM_F32) {
              // mirror: D[j][i] = D[i][j] from the SAME value, so the matrix is exactly symmetric and equal to
              // the condensed (upper-triangle) distances; lanes walk i, so these stores are coalesced too
              const int gi_l = p.row0 + ir.m_tile * kDM + quarter * 32 + lane;
              for (int c = 0; c < 32; ++c) {
                const int jj = n_base + n0 + c;
                if (jj < p.N && gi_l < p.N && jj > gi_l)
                  static_cast<float*>(p.out)[static_cast<long long>(jj) * p.N + gi_l] = mytr[lane * 33 + c];
              }
            }
            __syncwarp();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      }
      if (p.mode == DIST_KNN && row_ok && ir.n_end > ir.n_begin) {
        const size_t o = ((static_cast<size_t>(li) * p.chunks + ir.chunk) * kDSub + csub) * kTopK;
#pragma unroll
        for (int t = 0; t < kTopK; ++t) { p.part_d[o + t] = td[t]; p.part_i[o + t] = tid[t]; }
      }
    }
  } else if (warp == kDEpiWarps) {
    // ============================================================== MMA issuer
    const uint32_t idesc = make_idesc_f16(kDM, kDN, 1);
    int stage = 0;
    uint32_t phase = 0;
    int tile_iter = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const ItemRange ir = item_range(p, item);
      for (int n_tile = ir.n_begin; n_tile < ir.n_end; ++n_tile, ++tile_iter) {
        const int buf = tile_iter & 1;
        const uint32_t tphase = (tile_iter >> 1) & 1;
        mbar_wait(&tempty_bar[buf], tphase ^ 1, p.watchdog, 200 + buf);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kDN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase, p.watchdog, 210 + stage);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t a_desc = make_sw128_kmajor_desc(smem_u32(smem_a + stage * (kDM * kDK * 2)));
            const uint64_t b_desc = make_sw128_kmajor_desc(smem_u32(smem_b + stage * (kDN * kDK * 2)));
#pragma unroll
            for (int k4 = 0; k4 < kDK / 16; ++k4)
              tc_mma_f16(d_tmem, a_desc + static_cast<uint64_t>(k4 * 2), b_desc + static_cast<uint64_t>(k4 * 2), idesc,
                         (kb | k4) != 0 ? 1u : 0u);
            tc_commit(&empty_bar[stage]);
            if (kb == num_kb - 1) tc_commit(&tfull_bar[buf]);
          }
          __syncwarp();
          if (++stage == kDStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ============================================================ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const ItemRange ir = item_range(p, item);
      for (int n_tile = ir.n_begin; n_tile < ir.n_end; ++n_tile) {
        for (int kb = 0; kb < num_kb; ++kb) {
          const int seg = kb / kb_per_seg;
          const int kk = (kb - seg * kb_per_seg) * kDK;
          const int a_x = kk + (seg == 2 ? p.Dp : 0);   // hi.hi, hi.lo, lo.hi
          const int b_x = kk + (seg == 1 ? p.Dp : 0);
          mbar_wait(&empty_bar[stage], phase ^ 1, p.watchdog, 300 + stage);
          if (lane == 0) {
            mbar_arrive_expect_tx(&full_bar[stage], kDStageBytes);
            tma_load_2d(smem_a + stage * (kDM * kDK * 2), &tmA, &full_bar[stage], a_x, p.row0 + ir.m_tile * kDM);
            tma_load_2d(smem_b + stage * (kDN * kDK * 2), &tmB, &full_bar[stage], b_x, n_tile * kDN);
          }
          __syncwarp();
          if (++stage == kDStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kDEpiWarps) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------ kNN: merge chunk lists
__global__ void knn_merge_kernel(const float* __restrict__ part_d, const int32_t* __restrict__ part_i, int Q, int chunks,
                                 int k, const float* __restrict__ nq, int32_t* __restrict__ idx, float* __restrict__ dist) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  float d[kTopK];
  int id[kTopK];
#pragma unroll
  for (int t = 0; t < kTopK; ++t) { d[t] = INFINITY; id[t] = 0x7fffffff; }
  const size_t base = static_cast<size_t>(q) * chunks * kTopK;
  for (int c = 0; c < chunks * kTopK; ++c) {
    const float v = part_d[base + c];
    const int j = part_i[base + c];
    // lexicographic (distance, index): ties resolve to the lower database index
    if (!(v < d[kTopK - 1] || (v == d[kTopK - 1] && j < id[kTopK - 1]))) continue;
#pragma unroll
    for (int t = kTopK - 1; t >= 1; --t) {
      const bool shift = v < d[t - 1] || (v == d[t - 1] && j < id[t - 1]);
      const bool place = !shift && (v < d[t] || (v == d[t] && j < id[t]));
      d[t] = shift ? d[t - 1] : (place ? v : d[t]);
      id[t] = shift ? id[t - 1] : (place ? j : id[t]);
    }
    if (v < d[0] || (v == d[0] && j < id[0])) { d[0] = v; id[0] = j; }
  }
  const float n2 = nq[q];
#pragma unroll
  for (int t = 0; t < kTopK; ++t) {
    if (t < k) {
      idx[static_cast<size_t>(q) * k + t] = id[t] == 0x7fffffff ? -1 : id[t];
      dist[static_cast<size_t>(q) * k + t] = sqrtf(fmaxf(n2 + d[t], 0.f));
    }
  }
}

// ------------------------------------------------------------------ K15: neighbour labels -> class histogram / k
__global__ void knn_vote_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ label, int Q, int k,
                                int n_classes, float* __restrict__ proba) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  int lab[kTopK];
  for (int t = 0; t < k; ++t) {
    const int j = idx[static_cast<size_t>(q) * k + t];
    lab[t] = j >= 0 ? label[j] : -1;
  }
  for (int t = 0; t < k; ++t) {
    if (lab[t] < 0 || lab[t] >= n_classes) continue;
    int cnt = 0;
    for (int u = 0; u < k; ++u) cnt += lab[u] == lab[t];
    proba[static_cast<size_t>(q) * n_classes + lab[t]] = static_cast<float>(cnt) / static_cast<float>(k);
  }
}

// ------------------------------------------------------------------ K12: per-segment mean
// numpy `vectors.mean(axis=0)` on float32: sequential fp32 accumulation over rows, then / n.
__global__ void segment_mean_kernel(const float* __restrict__ x, const int32_t* __restrict__ off, int S, int D,
                                    float* __restrict__ out) {
  const int s = blockIdx.x;
  const int b = off[s], e = off[s + 1];
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (int r = b; r < e; ++r) acc = __fadd_rn(acc, x[static_cast<size_t>(r) * D + c]);
    out[static_cast<size_t>(s) * D + c] = __fdiv_rn(acc, static_cast<float>(e - b));
  }
}

// ------------------------------------------------------------------ host side
static int encode_bf16_rows(frb_ctx* ctx, CUtensorMap* tm, const void* base, int rows, int Dp) {
  return encode_2d(ctx, tm, base, static_cast<uint64_t>(2 * Dp), static_cast<uint64_t>(rows),
                   static_cast<uint64_t>(2 * Dp) * 2, kDK, kDM);
}

static int split_rows(frb_ctx* ctx, const float* d_x, int n, int D, int Dp, void* d_split, float* d_norm, cudaStream_t st) {
  const int grid = std::max(1, std::min((n + 7) / 8, ctx->num_sms * 8));
  split_rows_kernel<<<grid, 256, 0, st>>>(d_x, n, D, Dp, static_cast<__nv_bfloat16*>(d_split), d_norm);
  ctx->launches++;
  return check_last(ctx, "split_rows_kernel");
}

static size_t dist_smem_bytes() { return static_cast<size_t>(kDStages) * kDStageBytes + kDEpiWarps * 32 * 33 * 4 + (2 * kDStages + 4) * 8 + 16 + 1024; }

static int launch_dist(frb_ctx* ctx, const DistArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmB, cudaStream_t st) {
  static bool attr_set = false;
  const size_t smem = dist_smem_bytes();
  if (!attr_set) {
    FRB_CUDA(ctx, cudaFuncSetAttribute(dist_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int grid = std::max(1, std::min(a.items, ctx->num_sms));
  dist_gemm_kernel<<<grid, kDThreads, smem, st>>>(a, tmA, tmB);
  ctx->launches++;
  return check_last(ctx, "dist_gemm_kernel");
}

int run_pdist(frb_ctx* ctx, const float* d_X, int N, int D, int row0, int rows, int layout, void* d_out, cudaStream_t st) {
  if (N <= 0 || D <= 0 || row0 < 0 || rows < 0 || row0 + rows > N) return set_err(ctx, FRB_ERR_INVALID, "pdist: bad shape");
  if (layout < 0 || layout > 3) return set_err(ctx, FRB_ERR_INVALID, "pdist: bad layout %d", layout);
  if (layout == FRB_SQUARE_SYM_F32 && (row0 != 0 || rows != N))
    return set_err(ctx, FRB_ERR_INVALID, "pdist: the symmetric square layout covers all rows (row0 = 0, rows = N)");
  if (rows == 0) return FRB_OK;
  const int Dp = (D + kDK - 1) / kDK * kDK;
  void* d_split = nullptr;
  void* d_norm = nullptr;
  int rc;
  if ((rc = ctx_scratch(ctx, 0, static_cast<size_t>(N) * 2 * Dp * 2, &d_split)) != FRB_OK) return rc;
  if ((rc = ctx_scratch(ctx, 1, static_cast<size_t>(N) * 4, &d_norm)) != FRB_OK) return rc;
  if ((rc = split_rows(ctx, d_X, N, D, Dp, d_split, static_cast<float*>(d_norm), st)) != FRB_OK) return rc;
  CUtensorMap tm;
  if ((rc = encode_bf16_rows(ctx, &tm, d_split, N, Dp)) != FRB_OK) return rc;
  DistArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = layout == FRB_SQUARE_F32 ? DIST_SQUARE_F32
           : (layout == FRB_SQUARE_SYM_F32 ? DIST_SQUARE_SYM_F32
              : (layout == FRB_CONDENSED_F32 ? DIST_CONDENSED_F32 : DIST_CONDENSED_F64));
  a.rows = rows; a.row0 = row0; a.N = N; a.Dp = Dp;
  a.m_tiles = (rows + kDM - 1) / kDM;
  a.n_tiles = (N + kDN - 1) / kDN;
  a.chunks = 1; a.tiles_per_chunk = 1;
  const long long items = static_cast<long long>(a.m_tiles) * a.n_tiles;
  if (items > 0x7fffffffLL) return set_err(ctx, FRB_ERR_INVALID, "pdist: too many tiles");
  a.items = static_cast<int>(items);
  a.na = static_cast<const float*>(d_norm);
  a.nb = static_cast<const float*>(d_norm);
  a.out = d_out;
  a.watchdog = ctx->wd_dev;
  return launch_dist(ctx, a, tm, tm, st);
}

int run_knn_topk(frb_ctx* ctx, const float* d_Q, int Q, const float* d_DB, int M, int D, int k, int32_t* d_idx,
                 float* d_dist, cudaStream_t st) {
  if (Q < 0 || M <= 0 || D <= 0 || k <= 0 || k > kTopK) return set_err(ctx, FRB_ERR_INVALID, "knn_topk: bad shape (k must be 1..%d)", kTopK);
  if (Q == 0) return FRB_OK;
  const int Dp = (D + kDK - 1) / kDK * kDK;
  void *qs = nullptr, *qn = nullptr, *ds = nullptr, *dn = nullptr, *pd = nullptr, *pi = nullptr;
  int rc;
  if ((rc = ctx_scratch(ctx, 0, static_cast<size_t>(M) * 2 * Dp * 2, &ds)) != FRB_OK) return rc;
  if ((rc = ctx_scratch(ctx, 1, static_cast<size_t>(M) * 4, &dn)) != FRB_OK) return rc;
  if ((rc = ctx_scratch(ctx, 2, static_cast<size_t>(Q) * 2 * Dp * 2, &qs)) != FRB_OK) return rc;
  if ((rc = ctx_scratch(ctx, 3, static_cast<size_t>(Q) * 4, &qn)) != FRB_OK) return rc;
  DistArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = DIST_KNN;
  a.rows = Q; a.row0 = 0; a.N = M; a.Dp = Dp;
  a.m_tiles = (Q + kDM - 1) / kDM;
  a.n_tiles = (M + kDN - 1) / kDN;
  // split the column range into chunks so that (m_tile, chunk) items fill whole waves of SMs:
  // pick the chunk count (chunks no shorter than 8 tiles) with the best last-wave occupancy
  int chunks = 1;
  double best = -1.0;
  const int max_chunks = std::max(1, std::min(32, a.n_tiles / 8));
  for (int c = 1; c <= max_chunks; ++c) {
    const long long it = static_cast<long long>(a.m_tiles) * c;
    const long long waves = (it + ctx->num_sms - 1) / ctx->num_sms;
    const double eff = static_cast<double>(it) / static_cast<double>(waves * ctx->num_sms) - 0.002 * c;
    if (eff > best) { best = eff; chunks = c; }
  }
  a.tiles_per_chunk = (a.n_tiles + chunks - 1) / chunks;
  a.chunks = (a.n_tiles + a.tiles_per_chunk - 1) / a.tiles_per_chunk;
  const long long items = static_cast<long long>(a.m_tiles) * a.chunks;
  if (items > 0x7fffffffLL) return set_err(ctx, FRB_ERR_INVALID, "knn_topk: too many items");
  a.items = static_cast<int>(items);
  const size_t part = static_cast<size_t>(Q) * a.chunks * kDSub * kTopK;
  if ((rc = ctx_scratch(ctx, 4, part * 4, &pd)) != FRB_OK) return rc;
  if ((rc = ctx_scratch(ctx, 5, part * 4, &pi)) != FRB_OK) return rc;
  if ((rc = split_rows(ctx, d_DB, M, D, Dp, ds, static_cast<float*>(dn), st)) != FRB_OK) return rc;
  if ((rc = split_rows(ctx, d_Q, Q, D, Dp, qs, static_cast<float*>(qn), st)) != FRB_OK) return rc;
  CUtensorMap tmA, tmB;
  if ((rc = encode_bf16_rows(ctx, &tmA, qs, Q, Dp)) != FRB_OK) return rc;
  if ((rc = encode_bf16_rows(ctx, &tmB, ds, M, Dp)) != FRB_OK) return rc;
  a.na = static_cast<const float*>(qn);
  a.nb = static_cast<const float*>(dn);
  a.part_d = static_cast<float*>(pd);
  a.part_i = static_cast<int32_t*>(pi);
  a.watchdog = ctx->wd_dev;
  if ((rc = launch_dist(ctx, a, tmA, tmB, st)) != FRB_OK) return rc;
  knn_merge_kernel<<<(Q + 127) / 128, 128, 0, st>>>(a.part_d, a.part_i, Q, a.chunks * kDSub, k, a.na, d_idx, d_dist);
  ctx->launches++;
  return check_last(ctx, "knn_merge_kernel");
}

int run_knn_vote(frb_ctx* ctx, const int32_t* d_idx, const int32_t* d_label_idx, int Q, int k, int n_classes,
                 float* d_proba, cudaStream_t st) {
  if (Q < 0 || k <= 0 || k > kTopK || n_classes <= 0) return set_err(ctx, FRB_ERR_INVALID, "knn_vote: bad shape");
  if (Q == 0) return FRB_OK;
  FRB_CUDA(ctx, cudaMemsetAsync(d_proba, 0, static_cast<size_t>(Q) * n_classes * 4, st));
  knn_vote_kernel<<<(Q + 127) / 128, 128, 0, st>>>(d_idx, d_label_idx, Q, k, n_classes, d_proba);
  ctx->launches++;
  return check_last(ctx, "knn_vote_kernel");
}

int run_segment_mean(frb_ctx* ctx, const float* d_X, const int32_t* d_offsets, int S, int D, float* d_out, cudaStream_t st) {
  if (S < 0 || D <= 0) return set_err(ctx, FRB_ERR_INVALID, "segment_mean: bad shape");
  if (S == 0) return FRB_OK;
  segment_mean_kernel<<<S, 128, 0, st>>>(d_X, d_offsets, S, D, d_out);
  ctx->launches++;
  return check_last(ctx, "segment_mean_kernel");
}

}  // namespace frb
