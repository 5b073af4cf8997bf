// Embedding-distance stage (K12-K15): Euclidean pdist for complete-linkage clustering and fused
// distance + top-k for the kNN actor classifier, both as one NT GEMM on tcgen05 tensor cores.
//
// Precision.  The reference computes these distances in fp64 (scipy pdist / sklearn brute kNN) and its results depend
// on the ORDER of distances, so plain 16-bit inputs are not enough for the final numbers.  Each matrix is scaled by a
// power of two so that its largest element lands in [2^13, 2^14) (exact, and it keeps the small parts out of the fp16
// subnormals) and every element is split into fp16 hi + fp16 lo (x = hi + lo + O(2^-22 |x|)); the fp32-grade dot product is
//   hi.hi + hi.lo + lo.hi          (three K-segments accumulated into the same fp32 TMEM tile)
// and d^2 = |a|^2 + |b|^2 - 2 a.b with fp32 norms of the original data, clamped at 0.
//   * pdist runs all three segments (every distance is an output).
//   * kNN runs ONE segment (hi.hi) over the whole database and keeps, per query, the kCand best APPROXIMATE scores;
//     the candidates are then re-scored exactly (fp32 FMA dot products of the original rows) and the k best are
//     PROVEN to be the true k nearest: every other database row has approximate score >= tau (the worst kept
//     candidate), so its exact score is >= tau - eps with eps = the worst-case fp16 rounding error of one dot product
//     (u = 2^-11 per operand; bf16 would make eps eight times larger and most queries undecidable);
//     a query whose k-th exact score is not below tau - eps is re-run through the three-segment path.  One third of
//     the tensor work of the exact scheme for (almost) every query, the same answers for all of them.
//     The candidate lists live in registers (sorted insertion, ~6 instructions per list entry): what keeps that cheap
//     is a tight threshold.  Lists restart with every (query block, database chunk) item, so every finished item
//     publishes its kCand-th best score per query (atomicMin) and later chunks of the same queries start from it: any
//     published value is the kCand-th best of a SUBSET of the database, hence an upper bound of the final one, and no
//     true candidate is ever filtered out (the merged result does not depend on which value a chunk happened to see).
//
// Kernel.  A CTA PAIR (thread-block cluster of 2 on the two SMs of a TPC, cta_group::2) computes 256 x 256 tiles:
// each CTA holds 128 rows of A, its 128 accumulator lanes (TMEM, two buffers of 256 columns) and HALF of the B tile;
// the leader CTA issues tcgen05.mma M = 256.  With M = 128 tiles per SM this GEMM is bound by L2 -> SM bandwidth
// (round 1: 107 GB through L2 for a 2.2 TFLOP kNN, tensor pipe 39 % active, profiles/r02_glue_dist_launches_v0.csv);
// the pair halves the B bytes per SM, and in the one-segment kNN pass the A rows of an item (128 KB per CTA) stay
// resident in shared memory, so only B streams: 31 B/clk per SM instead of 94.
// Roles per CTA (320 threads): warps 0-7 epilogue (one TMEM lane quarter x one 128-column half each), warp 8 MMA issuer
// (leader CTA only; allocates TMEM in both), warp 9 TMA producer (both CTAs; completion bytes go to the leader's barrier).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#include "host.h"
#include "common.cuh"

namespace frb {

constexpr int kGM = 128;          // A rows per CTA (256 per pair)
constexpr int kGN = 256;          // tile columns per pair; each CTA loads kGN / 2 rows of B
constexpr int kGK = 64;           // K elements per pipeline stage
constexpr int kGEpiWarps = 8;
constexpr int kGThreads = (kGEpiWarps + 2) * 32;
constexpr int kCand = 16;         // per-row candidates kept in registers by the kNN epilogue (k <= 16)
constexpr int kTopK = 16;         // largest k of frb_knn_topk
constexpr int kGABytes = kGM * kGK * 2;         // one A k-block of a CTA
constexpr int kGBBytes = (kGN / 2) * kGK * 2;   // one B k-block half
constexpr int kGMaxStages = 8;
constexpr int kGWarpScratchPdist = 32 * 33 * 4 + 128 * 4;          // per warp: 32 x 33 transposition tile + |b|^2 of its 128 columns
constexpr int kGScratchPdist = kGEpiWarps * kGWarpScratchPdist;
constexpr int kGScratchKnn = kGEpiWarps * 128 * 4;                 // per-warp |b|^2 of its 128 tile columns
// worst-case |approximate - exact| score of the one-segment pass, relative to |q| * max|x|:
// 2 * (u (2 + u) with u = 2^-11 for the two fp16 roundings + 512 * 2^-21 for the fp32 accumulation of 512 products)
constexpr float kKnnEps = 0.0025f;

enum DistMode : int { DIST_SQUARE_F32 = 0, DIST_CONDENSED_F32 = 1, DIST_CONDENSED_F64 = 2, DIST_KNN = 3, DIST_SQUARE_SYM_F32 = 4 };

struct DistArgs {
  int mode;
  int segs;              // 1: hi.hi only, 3: hi.hi + hi.lo + lo.hi
  int a_resident;        // 1: the item's A rows (segs == 1) are loaded once and stay in shared memory
  int stages;
  int rows, row0;        // A rows handled by this call: global rows [row0, row0+rows)
  int N;                 // B rows
  int Dp;                // padded feature length (multiple of 64); split row = [hi(Dp) | lo(Dp)]
  int m_pairs, n_tiles;  // 256-row A blocks, 256-column B tiles
  int chunks, tiles_per_chunk;   // kNN: each (m_pair, chunk) item scans tiles_per_chunk column tiles
  int items;
  const float* na;       // |a|^2 indexed by global A row (pdist)
  const float* nb;       // |b|^2, padded to n_tiles * 256 entries with +inf (kNN) / 0 (pdist)
  const int* exp_a;      // power-of-two scale exponents of the split A and B matrices (device, written by split_rows_kernel)
  const int* exp_b;
  unsigned int* thr_bits;  // kNN: per-query published threshold (orderable encoding of the best kCand-th score seen so far)
  void* out;
  float* part_d;         // kNN partial lists [Qrows, parts = 2 * chunks, kCand]
  int32_t* part_i;
  int* watchdog;
  int dbg;               // debug build only (FRB_DIST_DBG): 1 = the epilogue only hands the accumulators back (timing the MMA + TMA side)
};

// ------------------------------------------------------------------ fp32 -> scaled (hi | lo) fp16 + norms
// max |x| of a matrix (bits of a non-negative float order like the float)
__global__ void max_abs_kernel(const float* __restrict__ x, size_t n, unsigned int* __restrict__ out_bits) {
  float m = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f && m < INFINITY) atomicMax(out_bits, __float_as_uint(m));
}

// Rows are scaled by 2^e, e chosen so that the largest element of the matrix lands in [2^13, 2^14) (exact; fp16 holds up
// to 65504 and the lo parts of all but vanishing elements stay normal), times `extra` (-2 for kNN queries: the
// accumulator then holds -2 * 2^(ea+eb) q.x and the epilogue is one FMA).  norm2[r] for n <= r < n_pad is filled with
// `pad_norm` (+inf for a kNN database: padding columns never win).  *exp_out = e.
__global__ void split_rows_kernel(const float* __restrict__ x, int n, int D, int Dp, float extra, const unsigned int* __restrict__ max_bits,
                                  __half* __restrict__ out, float* __restrict__ norm2, int n_pad, float pad_norm,
                                  unsigned int* __restrict__ max_norm_bits, int* __restrict__ exp_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float mxa = __uint_as_float(*max_bits);
  int e = 0;
  if (mxa > 0.f) {
    int ex;
    frexpf(mxa, &ex);                 // mxa = f * 2^ex, f in [0.5, 1)  ->  mxa * 2^(14 - ex) in [2^13, 2^14)
    e = 14 - ex;
  }
  const float scale = ldexpf(extra, e);
  if (blockIdx.x == 0 && threadIdx.x == 0) *exp_out = e;
  float mx = 0.f;
  for (int r = warp; r < n_pad; r += nwarps) {
    if (r >= n) {
      if (lane == 0) norm2[r] = pad_norm;
      continue;
    }
    const float* row = x + static_cast<size_t>(r) * D;
    __half* o = out + static_cast<size_t>(r) * 2 * Dp;
    float ss = 0.f;
    for (int c = lane; c < Dp; c += 32) {
      const float v = c < D ? row[c] : 0.f;
      const float sv = v * scale;                             // exact: power of two (times -2)
      const __half hi = __float2half_rn(sv);
      const __half lo = __float2half_rn(sv - __half2float(hi));
      o[c] = hi;
      o[Dp + c] = lo;
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) norm2[r] = ss;
    mx = fmaxf(mx, ss);
  }
  if (max_norm_bits != nullptr && lane == 0 && mx > 0.f) atomicMax(max_norm_bits, __float_as_uint(mx));
}

// MUFU.SQRT (<= 1 ulp): the IEEE sqrtf sequence was a third of the pdist epilogue's instructions
__device__ __forceinline__ float sqrt_fast(float x) {
  float y;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// orderable encoding of a float (unsigned compare == float compare), for atomicMin on scores that may be negative
__device__ __forceinline__ unsigned int enc_f32(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ------------------------------------------------------------------ item -> tile range
struct ItemRange {
  int m_pair, n_begin, n_end, chunk;
};
__device__ __forceinline__ ItemRange item_range(const DistArgs& p, int item) {
  ItemRange r;
  if (p.mode == DIST_KNN) {
    // chunk-major: the clusters running at the same time scan the same part of the database (shared through L2)
    r.chunk = item / p.m_pairs;
    r.m_pair = item - r.chunk * p.m_pairs;
    r.n_begin = r.chunk * p.tiles_per_chunk;
    r.n_end = min(r.n_begin + p.tiles_per_chunk, p.n_tiles);
  } else {
    r.m_pair = item / p.n_tiles;
    r.chunk = 0;
    r.n_begin = item - r.m_pair * p.n_tiles;
    r.n_end = r.n_begin + 1;
    // condensed layouts only hold j > i: drop tiles entirely on or below the diagonal of the pair's 256 rows
    const int first_row = p.row0 + r.m_pair * 2 * kGM;
    if (p.mode == DIST_SQUARE_SYM_F32) {
      if ((r.n_begin + 1) * kGN <= first_row) r.n_end = r.n_begin;   // keep tiles touching j >= i
    } else if (p.mode != DIST_SQUARE_F32 && (r.n_begin + 1) * kGN <= first_row + 1) r.n_end = r.n_begin;
  }
  return r;
}

__device__ __forceinline__ void cand_insert(float (&d)[kCand], int (&id)[kCand], float v, int j) {
#pragma unroll
  for (int t = kCand - 1; t >= 1; --t) {
    const bool shift = v < d[t - 1];
    const bool place = !shift && v < d[t];
    d[t] = shift ? d[t - 1] : (place ? v : d[t]);
    id[t] = shift ? id[t - 1] : (place ? j : id[t]);
  }
  if (v < d[0]) { d[0] = v; id[0] = j; }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGThreads, 1)
dist_gemm_kernel(const __grid_constant__ DistArgs p, const __grid_constant__ CUtensorMap tmA,
                 const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const int kb_per_seg = p.Dp / kGK;
  const int num_kb = p.segs * kb_per_seg;
  const int stage_bytes = (p.a_resident ? 0 : kGABytes) + kGBBytes;
  uint8_t* smem_ares = smem;                                                    // [kb_per_seg][kGABytes] when a_resident
  uint8_t* smem_ring = smem_ares + (p.a_resident ? kb_per_seg * kGABytes : 0);  // [stages][A? | B]
  float* scratch = reinterpret_cast<float*>(smem_ring + p.stages * stage_bytes);
  const int scratch_bytes = p.mode == DIST_KNN ? kGScratchKnn : kGScratchPdist;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(scratch) + scratch_bytes);
  uint64_t* full_bar = bars;                          // [kGMaxStages]  used in the leader CTA
  uint64_t* empty_bar = bars + kGMaxStages;           // [kGMaxStages]  both CTAs (multicast commit)
  uint64_t* tfull_bar = bars + 2 * kGMaxStages;       // [2]            both CTAs (multicast commit)
  uint64_t* tempty_bar = bars + 2 * kGMaxStages + 2;  // [2]            leader: 2 x kGEpiWarps arrivals
  uint64_t* ares_full = bars + 2 * kGMaxStages + 4;   // [1]            leader
  uint64_t* ares_empty = bars + 2 * kGMaxStages + 5;  // [1]            both CTAs (multicast commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kGMaxStages + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  constexpr uint32_t kTmemCols = 2 * kGN;
  constexpr int kMmaWarp = kGEpiWarps, kTmaWarp = kGEpiWarps + 1;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kGMaxStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tfull_bar[b], 1);
        mbar_init(&tempty_bar[b], 2 * kGEpiWarps);
      }
      mbar_init(ares_full, 1);
      mbar_init(ares_empty, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2cta(tmem_slot, kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs: barriers initialised and TMEM allocated before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kGEpiWarps) {
    // ================================================================ epilogue
    const int quarter = warp & 3;            // TMEM lane quarter
    const int half = warp >> 2;              // 128-column half of the 256-column tile
    float* mysc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scratch) + warp * kGWarpScratchPdist);   // pdist: 32 x 33 tile
    float* mynb = p.mode == DIST_KNN ? scratch + warp * 128 : mysc + 32 * 33;                                     // |b|^2 of 128 columns
    const uint32_t tempty_leader = mapa_shared(smem_u32(tempty_bar), 0);
    // accumulator -> score units: 2^-(ea+eb); pdist folds the -2 of  |a|^2 + |b|^2 - 2 a.b  in here, kNN has it in A
    const float unscale = ldexpf(p.mode == DIST_KNN ? 1.f : -2.f, -(__ldg(p.exp_a) + __ldg(p.exp_b)));
    int tile_iter = 0;
    for (int item = cluster_id; item < p.items; item += n_clusters) {
      const ItemRange ir = item_range(p, item);
      const int li = (ir.m_pair * 2 + static_cast<int>(rank)) * kGM + quarter * 32 + lane;   // local A row of this thread
      const int gi = p.row0 + li;                                                              // global A row
      const bool row_ok = li < p.rows;
      float td[kCand];
      int tid[kCand];
      float na = 0.f, seed = INFINITY, thr = INFINITY;
      long long rowoff = 0;
      if (p.mode == DIST_KNN) {
#pragma unroll
        for (int t = 0; t < kCand; ++t) { td[t] = INFINITY; tid[t] = 0x7fffffff; }
        // whatever an earlier chunk of these queries has published: an upper bound of the final kCand-th best score
        if (row_ok && ir.n_end > ir.n_begin) seed = dec_f32(*reinterpret_cast<volatile unsigned int*>(p.thr_bits + li));
        if (!(seed == seed)) seed = INFINITY;
        thr = seed;
      } else {
        na = row_ok ? __ldg(p.na + gi) : 0.f;
        if (p.mode == DIST_SQUARE_F32 || p.mode == DIST_SQUARE_SYM_F32) rowoff = static_cast<long long>(li) * p.N;
        else rowoff = static_cast<long long>(gi) * p.N - (static_cast<long long>(gi) * (gi + 1)) / 2 - gi - 1;
      }
      const int row_first = (ir.m_pair * 2 + static_cast<int>(rank)) * kGM + quarter * 32;   // first row of this warp
      // |x|^2 of the warp's 128 columns of a tile: one coalesced 512-byte load, fetched one tile ahead
      float4 nb_next = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ir.n_end > ir.n_begin) nb_next = __ldg(reinterpret_cast<const float4*>(p.nb + ir.n_begin * kGN + half * 128) + lane);

      for (int n_tile = ir.n_begin; n_tile < ir.n_end; ++n_tile, ++tile_iter) {
        const int buf = tile_iter & 1;
        const uint32_t tphase = (tile_iter >> 1) & 1;
        const int n_base = n_tile * kGN + half * 128;                   // first database row of this warp's columns
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * kGN + half * 128);
        __syncwarp();                                                   // the previous tile's readers are done
        reinterpret_cast<float4*>(mynb)[lane] = nb_next;                // broadcast reads below
        if (n_tile + 1 < ir.n_end) nb_next = __ldg(reinterpret_cast<const float4*>(p.nb + n_base + kGN) + lane);
        __syncwarp();
        mbar_wait(&tfull_bar[buf], tphase, p.watchdog, 100 + buf);
        tc_fence_after();
#ifdef FRB_DEBUG_BUILD
        if (p.dbg & 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader + static_cast<uint32_t>(buf) * 8u);
          continue;
        }
#endif
        if (p.mode == DIST_KNN) {
          uint32_t v0[16], v1[16];
          tmem_ld16(taddr, v0);
          auto scan16 = [&](const uint32_t* v, int c0) {
            float s[16];
            bool any = false;
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              const float4 nb = *reinterpret_cast<const float4*>(mynb + c0 + c);
              s[c] = fmaf(__uint_as_float(v[c]), unscale, nb.x);
              s[c + 1] = fmaf(__uint_as_float(v[c + 1]), unscale, nb.y);
              s[c + 2] = fmaf(__uint_as_float(v[c + 2]), unscale, nb.z);
              s[c + 3] = fmaf(__uint_as_float(v[c + 3]), unscale, nb.w);
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) any = any || (s[c] < thr);
            if (any) {
#pragma unroll
              for (int c = 0; c < 16; ++c)
                if (s[c] < thr) {
                  cand_insert(td, tid, s[c], n_base + c0 + c);
                  thr = fminf(td[kCand - 1], seed);
                }
            }
            __syncwarp();                                               // tcgen05.ld is .sync.aligned: reconverge after the inserts
          };
#pragma unroll 1
          for (int g = 0; g < 8; g += 2) {
            tc_wait_ld16(v0);
            tmem_ld16(taddr + (g + 1) * 16, v1);
            scan16(v0, g * 16);
            tc_wait_ld16(v1);
            if (g + 2 < 8) tmem_ld16(taddr + (g + 2) * 16, v0);
            scan16(v1, (g + 1) * 16);
          }
        } else {
          // every row of the warp valid, every column valid and right of every row's diagonal: no predicates
          const bool off_diag = p.row0 + row_first + 31 < n_base || (p.mode == DIST_SQUARE_F32 && p.row0 + row_first >= n_base + 128);
          const bool interior = row_first + 32 <= p.rows && n_base + 128 <= p.N && off_diag;
#pragma unroll 1
          for (int sub = 0; sub < 4; ++sub) {
            const int c_base = n_base + sub * 32;
            uint32_t v[32];
            __syncwarp();
            tmem_ld16(taddr + sub * 32, v);
            tmem_ld16(taddr + sub * 32 + 16, v + 16);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              const float4 nb = *reinterpret_cast<const float4*>(mynb + sub * 32 + c);
              mysc[lane * 33 + c] = sqrt_fast(fmaxf(fmaf(__uint_as_float(v[c]), unscale, na + nb.x), 0.f));
              mysc[lane * 33 + c + 1] = sqrt_fast(fmaxf(fmaf(__uint_as_float(v[c + 1]), unscale, na + nb.y), 0.f));
              mysc[lane * 33 + c + 2] = sqrt_fast(fmaxf(fmaf(__uint_as_float(v[c + 2]), unscale, na + nb.z), 0.f));
              mysc[lane * 33 + c + 3] = sqrt_fast(fmaxf(fmaf(__uint_as_float(v[c + 3]), unscale, na + nb.w), 0.f));
            }
            __syncwarp();
            const int j = c_base + lane;
            if (interior && p.mode != DIST_CONDENSED_F64) {
              // running row offset: square rows are N apart, condensed row g + 1 starts N - g - 2 entries after row g
              float* o = static_cast<float*>(p.out) + j + __shfl_sync(0xffffffffu, rowoff, 0);
              const bool square = p.mode == DIST_SQUARE_F32 || p.mode == DIST_SQUARE_SYM_F32;
              int step = square ? p.N : p.N - (p.row0 + row_first) - 2;
#pragma unroll 8
              for (int r = 0; r < 32; ++r) {
                *o = mysc[r * 33 + lane];
                o += step;
                step -= square ? 0 : 1;
              }
            } else {
              for (int r = 0; r < 32; ++r) {
                const long long ro = __shfl_sync(0xffffffffu, rowoff, r);
                const int rli = row_first + r;
                const int rgi = p.row0 + rli;
                const bool ok = rli < p.rows && j < p.N &&
                                (p.mode == DIST_SQUARE_F32 || j > rgi || (p.mode == DIST_SQUARE_SYM_F32 && j == rgi));
                if (ok) {
                  const float val = j == rgi ? 0.f : mysc[r * 33 + lane];   // exact zero diagonal (square layouts)
                  if (p.mode == DIST_CONDENSED_F64) static_cast<double*>(p.out)[ro + j] = static_cast<double>(val);
                  else static_cast<float*>(p.out)[ro + j] = val;
                }
              }
            }
            if (p.mode == DIST_SQUARE_SYM_F32) {
              // mirror: D[j][i] = D[i][j] from the SAME value, so the matrix is exactly symmetric and equal to
              // the condensed (upper-triangle) distances; lanes walk i, so these stores are coalesced too
              const int gi_l = p.row0 + row_first + lane;
              for (int c = 0; c < 32; ++c) {
                const int jj = c_base + c;
                if (jj < p.N && gi_l < p.N && jj > gi_l)
                  static_cast<float*>(p.out)[static_cast<long long>(jj) * p.N + gi_l] = mysc[lane * 33 + c];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + static_cast<uint32_t>(buf) * 8u);
      }
      if (p.mode == DIST_KNN && row_ok && ir.n_end > ir.n_begin) {
        const size_t o = (static_cast<size_t>(li) * (2 * p.chunks) + ir.chunk * 2 + half) * kCand;
#pragma unroll
        for (int t = 0; t < kCand; ++t) { p.part_d[o + t] = td[t]; p.part_i[o + t] = tid[t]; }
        if (td[kCand - 1] < seed) atomicMin(p.thr_bits + li, enc_f32(td[kCand - 1]));   // publish for the later chunks
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================================================== MMA issuer (leader CTA)
    if (leader) {
      const uint32_t idesc = make_idesc_f16(2 * kGM, kGN, 0);   // fp16 operands
      const uint32_t hi1024 = desc_hi(1024);
      const uint32_t ring0 = smem_u32(smem_ring), ares0 = smem_u32(smem_ares);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), tfull0 = smem_u32(tfull_bar),
                     tempty0 = smem_u32(tempty_bar);
      int stage = 0;
      uint32_t phase = 0;
      int tile_iter = 0, item_iter = 0;
      for (int item = cluster_id; item < p.items; item += n_clusters) {
        const ItemRange ir = item_range(p, item);
        if (ir.n_end <= ir.n_begin) continue;
        if (p.a_resident) mbar_wait_a(smem_u32(ares_full), static_cast<uint32_t>(item_iter) & 1u, p.watchdog, 230);
        for (int n_tile = ir.n_begin; n_tile < ir.n_end; ++n_tile, ++tile_iter) {
          const int buf = tile_iter & 1;
          const uint32_t tphase = (tile_iter >> 1) & 1;
          mbar_wait_a(tempty0 + buf * 8, tphase ^ 1, p.watchdog, 200 + buf);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kGN);
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait_a(full0 + stage * 8, phase, p.watchdog, 210 + stage);
            tc_fence_after();
            const uint32_t st = ring0 + static_cast<uint32_t>(stage * stage_bytes);
            const uint32_t a_lo = desc_lo(p.a_resident ? ares0 + static_cast<uint32_t>(kb * kGABytes) : st);
            const uint32_t b_lo = desc_lo(p.a_resident ? st : st + kGABytes);
            if (elect_one()) {
              tc_mma_f16_2cta(d_tmem, pack_desc(a_lo, hi1024), pack_desc(b_lo, hi1024), idesc, kb != 0 ? 1u : 0u);
              tc_mma_f16_2cta(d_tmem, pack_desc(a_lo + 2, hi1024), pack_desc(b_lo + 2, hi1024), idesc, 1u);
              tc_mma_f16_2cta(d_tmem, pack_desc(a_lo + 4, hi1024), pack_desc(b_lo + 4, hi1024), idesc, 1u);
              tc_mma_f16_2cta(d_tmem, pack_desc(a_lo + 6, hi1024), pack_desc(b_lo + 6, hi1024), idesc, 1u);
              tc_commit_2cta(empty0 + stage * 8, 3);
              if (kb == num_kb - 1) tc_commit_2cta(tfull0 + buf * 8, 3);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        if (p.a_resident) {
          if (elect_one()) tc_commit_2cta(smem_u32(ares_empty), 3);     // the item's MMAs have read the resident A rows
          __syncwarp();
        }
        ++item_iter;
      }
    }
  } else {
    // ============================================================ TMA producer (both CTAs)
    const uint32_t ring0 = smem_u32(smem_ring), ares0 = smem_u32(smem_ares);
    const uint32_t empty0 = smem_u32(empty_bar);
    const uint32_t full_leader = mapa_shared(smem_u32(full_bar), 0);
    const uint32_t ares_full_leader = mapa_shared(smem_u32(ares_full), 0);
    int stage = 0;
    uint32_t phase = 0;
    int item_iter = 0;
    for (int item = cluster_id; item < p.items; item += n_clusters) {
      const ItemRange ir = item_range(p, item);
      if (ir.n_end <= ir.n_begin) continue;
      const int a_row = p.row0 + (ir.m_pair * 2 + static_cast<int>(rank)) * kGM;
      if (p.a_resident) {
        mbar_wait_a(smem_u32(ares_empty), (static_cast<uint32_t>(item_iter) & 1u) ^ 1u, p.watchdog, 330);
        if (elect_one()) {
          if (leader) mbar_expect_tx_a(smem_u32(ares_full), static_cast<uint32_t>(2 * kb_per_seg * kGABytes));
          for (int kb = 0; kb < kb_per_seg; ++kb)
            tma_load_2d_2cta(ares0 + static_cast<uint32_t>(kb * kGABytes), &tmA, ares_full_leader, kb * kGK, a_row);
        }
        __syncwarp();
      }
      for (int n_tile = ir.n_begin; n_tile < ir.n_end; ++n_tile) {
        const int b_row = n_tile * kGN + static_cast<int>(rank) * (kGN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int seg = kb / kb_per_seg;
          const int kk = (kb - seg * kb_per_seg) * kGK;
          const int a_x = kk + (seg == 2 ? p.Dp : 0);   // hi.hi, hi.lo, lo.hi
          const int b_x = kk + (seg == 1 ? p.Dp : 0);
          mbar_wait_a(empty0 + stage * 8, phase ^ 1, p.watchdog, 300 + stage);
          if (elect_one()) {
            const uint32_t st = ring0 + static_cast<uint32_t>(stage * stage_bytes);
            const uint32_t fb = full_leader + static_cast<uint32_t>(stage) * 8u;
            {
            if (leader) mbar_expect_tx_a(smem_u32(full_bar) + stage * 8, static_cast<uint32_t>(2 * stage_bytes));
            if (!p.a_resident) {
              tma_load_2d_2cta(st, &tmA, fb, a_x, a_row);
              tma_load_2d_2cta(st + kGABytes, &tmB, fb, b_x, b_row);
            } else {
              tma_load_2d_2cta(st, &tmB, fb, b_x, b_row);
            }
            }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
      ++item_iter;
    }
  }

  // -------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer may still be reading this CTA's B halves / arriving on its barriers
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------ kNN: merge the partial lists of a query
// One warp per query.  The `parts` lists (kCand entries each, ascending) are merged by a tournament of their heads:
// lane l owns lists l and l + 32.
//   exact == 0 (after the one-segment pass): the kCand best APPROXIMATE candidates are re-scored exactly,
//        s = |x|^2 - 2 q.x in fp32 from the original rows, the k best by (s, index) are written, and the query is
//        flagged for the exact pass unless  s_k < tau - eps  (see the header of this file).
//   exact == 1 (after the three-segment pass): the k best by (score, index) are written as they are.
// `rows` maps the warp's query to its row in idx / dist (the exact pass works on a compacted list of queries).
__global__ void knn_merge_kernel(const float* __restrict__ part_d, const int32_t* __restrict__ part_i, int Q, int parts, int k,
                                 int exact, const float* __restrict__ Qrows, const float* __restrict__ DB, int D,
                                 const float* __restrict__ nq, const float* __restrict__ nb, const unsigned int* __restrict__ max_nb_bits,
                                 const int32_t* __restrict__ rows, int32_t* __restrict__ idx, float* __restrict__ dist,
                                 int32_t* __restrict__ redo_list, int32_t* __restrict__ redo_count) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= Q) return;
  const float* pd = part_d + static_cast<size_t>(q) * parts * kCand;
  const int32_t* pi = part_i + static_cast<size_t>(q) * parts * kCand;
  // heads of this lane's (up to two) lists
  int h0 = 0, h1 = 0;
  const bool has0 = lane < parts, has1 = lane + 32 < parts;
  float my_s = INFINITY;       // the candidate this lane ends up owning (lane t = t-th best)
  int my_j = 0x7fffffff;
  float tau = INFINITY;
  for (int t = 0; t < kCand; ++t) {
    float s0 = has0 && h0 < kCand ? pd[lane * kCand + h0] : INFINITY;
    int j0 = has0 && h0 < kCand ? pi[lane * kCand + h0] : 0x7fffffff;
    float s1 = has1 && h1 < kCand ? pd[(lane + 32) * kCand + h1] : INFINITY;
    int j1 = has1 && h1 < kCand ? pi[(lane + 32) * kCand + h1] : 0x7fffffff;
    const bool first = s0 < s1 || (s0 == s1 && j0 <= j1);
    float bs = first ? s0 : s1;
    int bj = first ? j0 : j1;
    int bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (os < bs || (os == bs && oj < bj)) { bs = os; bj = oj; bl = ol; }
    }
    if (bl == lane) {
      if (first) ++h0; else ++h1;
    }
    if (lane == t) { my_s = bs; my_j = bj; }
    tau = bs;                  // after the loop: the worst of the kCand kept (approximate) scores
  }
  const int out_row = rows != nullptr ? rows[q] : q;
  const float n2 = nq[q];
  if (!exact) {
    // exact scores of the kCand candidates: the warp walks them one by one, 16 floats of the query per lane
    float qv[16];
    const float* qr = Qrows + static_cast<size_t>(q) * D;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int col = c * 32 + lane;
      qv[c] = col < D ? qr[col] : 0.f;
    }
    float exact_s = INFINITY;
    for (int t = 0; t < kCand; ++t) {
      const int j = __shfl_sync(0xffffffffu, my_j, t);
      if (j == 0x7fffffff) continue;
      const float* xr = DB + static_cast<size_t>(j) * D;
      float acc = 0.f;
      if (D == 512) {
#pragma unroll
        for (int c = 0; c < 16; ++c) acc = fmaf(qv[c], __ldg(xr + c * 32 + lane), acc);
      } else {
        for (int col = lane; col < D; col += 32) acc = fmaf(qr[col], __ldg(xr + col), acc);
      }
      acc = warp_sum(acc);
      if (lane == t) exact_s = fmaf(-2.f, acc, nb[j]);
    }
    my_s = exact_s;
  }
  // rank of this lane's candidate among the kCand by (score, index)
  int rank = 0;
  for (int t = 0; t < kCand; ++t) {
    const float os = __shfl_sync(0xffffffffu, my_s, t);
    const int oj = __shfl_sync(0xffffffffu, my_j, t);
    if (lane < kCand && t != lane && (os < my_s || (os == my_s && oj < my_j))) ++rank;
  }
  if (lane < kCand && rank < k) {
    idx[static_cast<size_t>(out_row) * k + rank] = my_j == 0x7fffffff ? -1 : my_j;
    dist[static_cast<size_t>(out_row) * k + rank] = sqrtf(fmaxf(n2 + my_s, 0.f));
  }
  if (!exact) {
    // the k-th exact score must beat every row that is NOT a candidate: those have approximate score >= tau
    const unsigned has_k = __ballot_sync(0xffffffffu, lane < kCand && rank == k - 1);
    const float sk = __shfl_sync(0xffffffffu, my_s, has_k ? __ffs(has_k) - 1 : 0);
    const float eps = kKnnEps * sqrtf(n2) * sqrtf(__uint_as_float(*max_nb_bits)) + 1e-6f;
    const bool proven = has_k != 0 && (tau == INFINITY || sk < tau - eps);
    if (!proven && lane == 0) redo_list[atomicAdd(redo_count, 1)] = q;
  }
}

// rows of the split matrix / norms of the flagged queries -> compact buffers for the exact pass
__global__ void knn_gather_redo_kernel(const int32_t* __restrict__ redo_list, const int32_t* __restrict__ redo_count,
                                       const float* __restrict__ Q, int D, float* __restrict__ Qc) {
  const int n = *redo_count;
  for (int r = blockIdx.x; r < n; r += gridDim.x) {
    const float* src = Q + static_cast<size_t>(redo_list[r]) * D;
    float* dst = Qc + static_cast<size_t>(r) * D;
    for (int c = threadIdx.x; c < D; c += blockDim.x) dst[c] = src[c];
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
static int encode_f16_rows(frb_ctx* ctx, CUtensorMap* tm, const void* base, int rows, int Dp) {
  // fp16 rows of [hi | lo]; box = 64 elements x 128 rows (one CTA's share of a k-block)
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(2 * Dp), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(2 * Dp) * 2};
  cuuint32_t box[2] = {kGK, kGM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx->encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(ctx, FRB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %d x %d split matrix", (int)r, rows, 2 * Dp);
  return FRB_OK;
}

static int max_abs(frb_ctx* ctx, const float* d_x, size_t n, unsigned int* d_bits, cudaStream_t st) {
  const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>((n + 2047) / 2048, static_cast<size_t>(ctx->num_sms) * 8)));
  max_abs_kernel<<<grid, 256, 0, st>>>(d_x, n, d_bits);
  ctx->launches++;
  return check_last(ctx, "max_abs_kernel");
}

static int split_rows(frb_ctx* ctx, const float* d_x, int n, int D, int Dp, float extra, const unsigned int* d_max_bits, void* d_split,
                      float* d_norm, int n_pad, float pad_norm, unsigned int* d_max_norm_bits, int* d_exp, cudaStream_t st) {
  const int grid = std::max(1, std::min((n_pad + 7) / 8, ctx->num_sms * 8));
  split_rows_kernel<<<grid, 256, 0, st>>>(d_x, n, D, Dp, extra, d_max_bits, static_cast<__half*>(d_split), d_norm, n_pad, pad_norm,
                                          d_max_norm_bits, d_exp);
  ctx->launches++;
  return check_last(ctx, "split_rows_kernel");
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

// shared-memory plan: resident A (one-segment kNN) + stage ring + epilogue scratch + barriers
static size_t dist_plan(DistArgs& a) {
  const int kb_per_seg = a.Dp / kGK;
  const size_t fixed = (a.a_resident ? static_cast<size_t>(kb_per_seg) * kGABytes : 0) +
                       (a.mode == DIST_KNN ? kGScratchKnn : kGScratchPdist) + (2 * kGMaxStages + 8) * 8 + 1024;
  const size_t stage_bytes = (a.a_resident ? 0 : kGABytes) + kGBBytes;
  const size_t budget = 227 * 1024;
  if (fixed + 2 * stage_bytes > budget) return 0;
  a.stages = static_cast<int>(std::min<size_t>(kGMaxStages, (budget - fixed) / stage_bytes));
  return fixed + a.stages * stage_bytes;
}

static int launch_dist(frb_ctx* ctx, DistArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmB, cudaStream_t st) {
  size_t smem = dist_plan(a);
  if (smem == 0 && a.a_resident) {       // very long rows: stream A like the three-segment path does
    a.a_resident = 0;
    smem = dist_plan(a);
  }
  if (smem == 0) return set_err(ctx, FRB_ERR_INVALID, "distance kernel: shared-memory plan does not fit (D = %d)", a.Dp);
#ifdef FRB_DEBUG_BUILD
  if (const char* e = getenv("FRB_DIST_DBG")) a.dbg = atoi(e);
  if (const char* e = getenv("FRB_DIST_STAGES")) a.stages = std::max(2, std::min(a.stages, atoi(e)));
#endif
  if (!ctx->attr_dist) {
    FRB_CUDA(ctx, cudaFuncSetAttribute(dist_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ctx->attr_dist = true;
  }
  // one CTA pair per TPC (the kernel carries __cluster_dims__(2, 1, 1)); a persistent grid of pairs walks the items
  const int pairs = std::max(1, std::min(a.items, ctx->num_sms / 2));
  dist_gemm_kernel<<<2 * pairs, kGThreads, smem, st>>>(a, tmA, tmB);
  ctx->launches++;
  return check_last(ctx, "dist_gemm_kernel");
}

int run_pdist(frb_ctx* ctx, const float* d_X, int N, int D, int row0, int rows, int layout, void* d_out, cudaStream_t st) {
  if (N <= 0 || D <= 0 || row0 < 0 || rows < 0 || row0 + rows > N) return set_err(ctx, FRB_ERR_INVALID, "pdist: bad shape");
  if (layout < 0 || layout > 3) return set_err(ctx, FRB_ERR_INVALID, "pdist: bad layout %d", layout);
  if (layout == FRB_SQUARE_SYM_F32 && (row0 != 0 || rows != N))
    return set_err(ctx, FRB_ERR_INVALID, "pdist: the symmetric square layout covers all rows (row0 = 0, rows = N)");
  if (rows == 0) return FRB_OK;
  const int Dp = round_up(D, kGK);
  const int n_pad = round_up(N, kGN);
  void* d_split = nullptr;
  void* d_norm = nullptr;
  void* misc = nullptr;
  int rc;
  if ((rc = ctx_scratch(ctx, 0, static_cast<size_t>(N) * 2 * Dp * 2, &d_split)) != FRB_OK) return rc;
  if ((rc = ctx_scratch(ctx, 1, static_cast<size_t>(n_pad) * 4, &d_norm)) != FRB_OK) return rc;
  if ((rc = ctx_scratch(ctx, 6, 64, &misc)) != FRB_OK) return rc;
  unsigned int* d_maxabs = static_cast<unsigned int*>(misc) + 2;
  int* d_exp = static_cast<int*>(misc) + 4;
  FRB_CUDA(ctx, cudaMemsetAsync(misc, 0, 64, st));
  if ((rc = max_abs(ctx, d_X, static_cast<size_t>(N) * D, d_maxabs, st)) != FRB_OK) return rc;
  if ((rc = split_rows(ctx, d_X, N, D, Dp, 1.f, d_maxabs, d_split, static_cast<float*>(d_norm), n_pad, 0.f, nullptr, d_exp, st)) != FRB_OK) return rc;
  CUtensorMap tm;
  if ((rc = encode_f16_rows(ctx, &tm, d_split, N, Dp)) != FRB_OK) return rc;
  DistArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = layout == FRB_SQUARE_F32 ? DIST_SQUARE_F32
           : (layout == FRB_SQUARE_SYM_F32 ? DIST_SQUARE_SYM_F32
              : (layout == FRB_CONDENSED_F32 ? DIST_CONDENSED_F32 : DIST_CONDENSED_F64));
  a.segs = 3;
  a.a_resident = 0;
  a.rows = rows; a.row0 = row0; a.N = N; a.Dp = Dp;
  a.m_pairs = (rows + 2 * kGM - 1) / (2 * kGM);
  a.n_tiles = n_pad / kGN;
  a.chunks = 1; a.tiles_per_chunk = 1;
  const long long items = static_cast<long long>(a.m_pairs) * a.n_tiles;
  if (items > 0x7fffffffLL) return set_err(ctx, FRB_ERR_INVALID, "pdist: too many tiles");
  a.items = static_cast<int>(items);
  a.na = static_cast<const float*>(d_norm);
  a.nb = static_cast<const float*>(d_norm);
  a.exp_a = d_exp;
  a.exp_b = d_exp;
  a.out = d_out;
  a.watchdog = ctx->wd_dev;
  return launch_dist(ctx, a, tm, tm, st);
}

// chunk count for the kNN scan: (m_pair, chunk) items should fill whole waves of CTA pairs; every chunk restarts the
// candidate lists, so few long chunks are better than many short ones when the fill is the same
static void knn_chunks(frb_ctx* ctx, DistArgs& a) {
  const int pairs = std::max(1, ctx->num_sms / 2);
  int chunks = 1;
  double best = -1.0;
  const int max_chunks = std::max(1, std::min(32, a.n_tiles / 4));
  for (int c = 1; c <= max_chunks; ++c) {
    const long long it = static_cast<long long>(a.m_pairs) * c;
    const long long waves = (it + pairs - 1) / pairs;
    const double eff = static_cast<double>(it) / static_cast<double>(waves * pairs) - 0.01 * c;
    if (eff > best) { best = eff; chunks = c; }
  }
  a.tiles_per_chunk = (a.n_tiles + chunks - 1) / chunks;
  a.chunks = (a.n_tiles + a.tiles_per_chunk - 1) / a.tiles_per_chunk;
  a.items = a.m_pairs * a.chunks;
}

// KNeighborsClassifier.fit: the database's scaled fp16 planes, |x|^2 (+inf padded), max |x|^2 and scale exponent, kept
// in scratch slots of their own until the next fit
int run_knn_fit(frb_ctx* ctx, const float* d_DB, int M, int D, cudaStream_t st) {
  if (M <= 0 || D <= 0) return set_err(ctx, FRB_ERR_INVALID, "knn_fit: bad shape");
  const int Dp = round_up(D, kGK);
  const int m_pad = round_up(M, kGN);
  void *ds = nullptr, *dn = nullptr, *fm = nullptr;
  int rc;
  ctx->knn_fit_db = nullptr;
  if ((rc = ctx_scratch(ctx, 8, static_cast<size_t>(M) * 2 * Dp * 2, &ds)) != FRB_OK) return rc;
  if ((rc = ctx_scratch(ctx, 9, static_cast<size_t>(m_pad) * 4, &dn)) != FRB_OK) return rc;
  if ((rc = ctx_scratch(ctx, 10, 64, &fm)) != FRB_OK) return rc;
  FRB_CUDA(ctx, cudaMemsetAsync(fm, 0, 64, st));
  unsigned int* bits = static_cast<unsigned int*>(fm);          // [0] max |x|^2, [1] max |x_ij|, [2] exponent
  if ((rc = max_abs(ctx, d_DB, static_cast<size_t>(M) * D, bits + 1, st)) != FRB_OK) return rc;
  if ((rc = split_rows(ctx, d_DB, M, D, Dp, 1.f, bits + 1, ds, static_cast<float*>(dn), m_pad, INFINITY, bits, reinterpret_cast<int*>(bits + 2), st)) != FRB_OK) return rc;
  ctx->knn_fit_db = d_DB;
  ctx->knn_fit_m = M;
  ctx->knn_fit_d = D;
  return FRB_OK;
}

int run_knn_topk(frb_ctx* ctx, const float* d_Q, int Q, const float* d_DB, int M, int D, int k, int32_t* d_idx,
                 float* d_dist, cudaStream_t st) {
  if (Q < 0 || M <= 0 || D <= 0 || k <= 0 || k > kTopK) return set_err(ctx, FRB_ERR_INVALID, "knn_topk: bad shape (k must be 1..%d)", kTopK);
  if (Q == 0) return FRB_OK;
  const int Dp = round_up(D, kGK);
  const int m_pad = round_up(M, kGN);
  void *qs = nullptr, *qn = nullptr, *ds = nullptr, *dn = nullptr, *pd = nullptr, *pi = nullptr, *misc = nullptr;
  int rc;
  const bool fitted = ctx->knn_fit_db == d_DB && ctx->knn_fit_m == M && ctx->knn_fit_d == D && ctx->scratch[8] != nullptr;
  if (fitted) {
    ds = ctx->scratch[8];
    dn = ctx->scratch[9];
  } else {
    if ((rc = ctx_scratch(ctx, 0, static_cast<size_t>(M) * 2 * Dp * 2, &ds)) != FRB_OK) return rc;
    if ((rc = ctx_scratch(ctx, 1, static_cast<size_t>(m_pad) * 4, &dn)) != FRB_OK) return rc;
  }
  if ((rc = ctx_scratch(ctx, 2, static_cast<size_t>(Q) * 2 * Dp * 2, &qs)) != FRB_OK) return rc;
  if ((rc = ctx_scratch(ctx, 3, static_cast<size_t>(Q) * 4, &qn)) != FRB_OK) return rc;
  // misc (32-bit words): [0] max |x|^2 bits, [1] redo count, [2] max|q| bits, [3] max|x| bits, [4] exp(Q), [5] exp(DB),
  //                      [8, 8+Q) redo list, [8+Q, 8+2Q) published thresholds, [8+2Q, 8+3Q) thresholds of the exact pass
  if ((rc = ctx_scratch(ctx, 6, (static_cast<size_t>(Q) * 3 + 8) * 4, &misc)) != FRB_OK) return rc;
  unsigned int* d_max_bits = fitted ? static_cast<unsigned int*>(ctx->scratch[10]) : static_cast<unsigned int*>(misc);
  int32_t* d_redo_count = static_cast<int32_t*>(misc) + 1;
  unsigned int* d_maxabs_q = static_cast<unsigned int*>(misc) + 2;
  unsigned int* d_maxabs_db = static_cast<unsigned int*>(misc) + 3;
  int* d_exp_q = static_cast<int*>(misc) + 4;
  int* d_exp_db = fitted ? static_cast<int*>(ctx->scratch[10]) + 2 : static_cast<int*>(misc) + 5;
  int32_t* d_redo_list = static_cast<int32_t*>(misc) + 8;
  unsigned int* d_thr = static_cast<unsigned int*>(misc) + 8 + Q;
  unsigned int* d_thr2 = d_thr + Q;
  FRB_CUDA(ctx, cudaMemsetAsync(misc, 0, 32, st));
  FRB_CUDA(ctx, cudaMemsetAsync(d_thr, 0xff, static_cast<size_t>(Q) * 2 * 4, st));   // "nothing published yet"

  DistArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = DIST_KNN;
  a.segs = 1;
  a.a_resident = 1;
  a.rows = Q; a.row0 = 0; a.N = M; a.Dp = Dp;
  a.m_pairs = (Q + 2 * kGM - 1) / (2 * kGM);
  a.n_tiles = m_pad / kGN;
  knn_chunks(ctx, a);
  const int parts = 2 * a.chunks;
  const size_t part = static_cast<size_t>(Q) * parts * kCand;
  if ((rc = ctx_scratch(ctx, 4, part * 4, &pd)) != FRB_OK) return rc;
  if ((rc = ctx_scratch(ctx, 5, part * 4, &pi)) != FRB_OK) return rc;
  // database: plain split, |x|^2 padded with +inf; queries: split scaled by -2 (accumulator = -2 q.x in scaled units)
  if (!fitted) {
    if ((rc = max_abs(ctx, d_DB, static_cast<size_t>(M) * D, d_maxabs_db, st)) != FRB_OK) return rc;
    if ((rc = split_rows(ctx, d_DB, M, D, Dp, 1.f, d_maxabs_db, ds, static_cast<float*>(dn), m_pad, INFINITY, d_max_bits, d_exp_db, st)) != FRB_OK) return rc;
  }
  if ((rc = max_abs(ctx, d_Q, static_cast<size_t>(Q) * D, d_maxabs_q, st)) != FRB_OK) return rc;
  if ((rc = split_rows(ctx, d_Q, Q, D, Dp, -2.f, d_maxabs_q, qs, static_cast<float*>(qn), Q, 0.f, nullptr, d_exp_q, st)) != FRB_OK) return rc;
  CUtensorMap tmA, tmB;
  if ((rc = encode_f16_rows(ctx, &tmA, qs, Q, Dp)) != FRB_OK) return rc;
  if ((rc = encode_f16_rows(ctx, &tmB, ds, M, Dp)) != FRB_OK) return rc;
  a.na = static_cast<const float*>(qn);
  a.nb = static_cast<const float*>(dn);
  a.exp_a = d_exp_q;
  a.exp_b = d_exp_db;
  a.thr_bits = d_thr;
  a.part_d = static_cast<float*>(pd);
  a.part_i = static_cast<int32_t*>(pi);
  a.watchdog = ctx->wd_dev;
  if ((rc = launch_dist(ctx, a, tmA, tmB, st)) != FRB_OK) return rc;
  knn_merge_kernel<<<(Q + 3) / 4, 128, 0, st>>>(a.part_d, a.part_i, Q, parts, k, 0, d_Q, d_DB, D, a.na, a.nb, d_max_bits, nullptr,
                                                d_idx, d_dist, d_redo_list, d_redo_count);
  ctx->launches++;
  if ((rc = check_last(ctx, "knn_merge_kernel")) != FRB_OK) return rc;

  // queries whose top-k could not be proven from the one-segment pass: the exact three-segment pass, on them only
  int32_t redo = 0;
  FRB_CUDA(ctx, cudaMemcpyAsync(&redo, d_redo_count, 4, cudaMemcpyDeviceToHost, st));
  FRB_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->knn_redo_last = redo;
  if (redo == 0) return FRB_OK;
#ifdef FRB_DEBUG_BUILD
  if (a.dbg) return FRB_OK;              // timing experiments produce garbage candidates: do not chase them
#endif
  void *qc = nullptr, *qcs = nullptr, *qcn = nullptr;
  if ((rc = ctx_scratch(ctx, 7, static_cast<size_t>(redo) * D * 4 + static_cast<size_t>(redo) * 2 * Dp * 2 + static_cast<size_t>(redo) * 4 + 512, &qc)) != FRB_OK) return rc;
  qcs = static_cast<uint8_t*>(qc) + ((static_cast<size_t>(redo) * D * 4 + 255) & ~static_cast<size_t>(255));
  qcn = static_cast<uint8_t*>(qcs) + static_cast<size_t>(redo) * 2 * Dp * 2;
  knn_gather_redo_kernel<<<std::min(redo, ctx->num_sms * 4), 128, 0, st>>>(d_redo_list, d_redo_count, d_Q, D, static_cast<float*>(qc));
  ctx->launches++;
  // (same power-of-two scale as the whole query matrix: the subset's largest element is not larger)
  if ((rc = split_rows(ctx, static_cast<const float*>(qc), redo, D, Dp, -2.f, d_maxabs_q, qcs, static_cast<float*>(qcn), redo, 0.f, nullptr, d_exp_q, st)) != FRB_OK) return rc;
  DistArgs e = a;
  e.segs = 3;
  e.a_resident = 0;
  e.rows = redo;
  e.m_pairs = (redo + 2 * kGM - 1) / (2 * kGM);
  knn_chunks(ctx, e);
  const int eparts = 2 * e.chunks;
  const size_t epart = static_cast<size_t>(redo) * eparts * kCand;
  if (epart > part) {                    // few queries, many chunks: the partial lists may outgrow the first pass's
    if ((rc = ctx_scratch(ctx, 4, epart * 4, &pd)) != FRB_OK) return rc;
    if ((rc = ctx_scratch(ctx, 5, epart * 4, &pi)) != FRB_OK) return rc;
    e.part_d = static_cast<float*>(pd);
    e.part_i = static_cast<int32_t*>(pi);
  }
  e.na = static_cast<const float*>(qcn);
  e.thr_bits = d_thr2;
  CUtensorMap tmC;
  if ((rc = encode_f16_rows(ctx, &tmC, qcs, redo, Dp)) != FRB_OK) return rc;
  if ((rc = launch_dist(ctx, e, tmC, tmB, st)) != FRB_OK) return rc;
  knn_merge_kernel<<<(redo + 3) / 4, 128, 0, st>>>(e.part_d, e.part_i, redo, eparts, k, 1, nullptr, nullptr, D, e.na, e.nb, d_max_bits,
                                                   d_redo_list, d_idx, d_dist, nullptr, nullptr);
  ctx->launches++;
  return check_last(ctx, "knn_merge_kernel (exact pass)");
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
