// K2 (main variant) — implicit-GEMM convolution with BOTH operands and the output moved by TMA.
//
// The im2col never exists in memory and no thread computes an address for it: the output is tiled
// into spatial patches of bb x bh x bw = 128 pixels (or 128 consecutive pixels for 1x1 convs), and for
// every filter tap (r, s) and 64-channel block the A operand of the patch is ONE 4-D tensor-map box
//   [64 ch, bw (stride sw), bh (stride sh), bb]   at   (c0, w0*sw - pw + s, h0*sh - ph + r, b0)
// which TMA delivers straight into the 128-byte-swizzled K-major layout tcgen05.mma consumes; padding
// and ragged image borders are the tensor map's out-of-bounds zero fill.  Weights come the same way
// (2-D box of the [Cout, K] matrix); when the whole weight matrix is small (<= 80 KB: the bandwidth-bound
// early layers) it is loaded ONCE per CTA and stays resident, so the stage ring carries only A and the
// L2 -> SM traffic drops by the weight re-fetch of every tile.  Accumulators live in TMEM (double
// buffered, 2 x BN columns), and the epilogue (bias, residual, per-channel slope) packs 64-column slabs
// into a swizzled staging tile that TMA stores back as the same patch box — partial tiles are clipped by
// the hardware, and every global store is a full 128-byte line.  The residual tile (ResNet / Inception
// skip) is a producer-side TMA prefetch, double buffered across tiles, so its latency hides behind the
// previous tile's epilogue.
//
// Halo mode (stride-1 k x k convs): instead of one A box per filter tap (k*k reads of the same pixels through
// L2), the producer loads ONE (bh+kh-1) x (bw+kw-1) halo tile per 64-channel block and the MMA issuer points
// the A descriptor of tap (r, s) at pixel (r, s) of that tile: 8-row groups are 8 consecutive pixels of a halo
// row (bw = 8), the stride between groups is the halo row pitch.  The 128-byte swizzle is a function of the
// shared-memory ADDRESS (tools/umma_shift_probe.cu: any 128-byte row offset and any group stride read back
// exactly what TMA wrote), so shifted windows need no copy.  A traffic through L2 drops k*k-fold.
//
// Warp roles (576 threads): 0-15 epilogue, 16 MMA issuer + TMEM allocator, 17 TMA producer.
// The epilogue is instruction-issue bound when one warp per scheduler walks a whole 64-column row
// (ncu: ~2 400 cycles per slab, issue slots mostly idle), so each TMEM lane quarter (warp % 4) is
// served by four warps that take 16 columns of the slab each: four warps per scheduler hide each
// other's dependent-issue latency.
// Persistent: one CTA per SM walks tiles round-robin (n fastest, so CTAs running at the same time
// share the A patch through L2).
#pragma once
#include "conv_igemm.cuh"

namespace frb {

constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kTmaThreads = kEpiThreads + 64;
constexpr int kSlab = 64;                  // epilogue / store granularity in output channels
constexpr int kSlabBytes = kBM * kSlab * 2;
constexpr int kColsPerWarp = kSlab / (kEpiWarps / 4);   // 16
constexpr int kOutBufs = 2;                // staging slabs (TMA stores drain asynchronously)
constexpr int kResSets = 2;                // residual tiles in flight (each = up to 2 slabs)
constexpr int kResidentWeightBytes = 80 * 1024;

struct ConvTmaArgs {
  int patch;             // 0: linear M tiling (1x1, stride 1, no padding); 1: spatial patch tiling
  int kw, sh, sw, ph, pw;
  int cin_blocks;        // 64-channel blocks per filter tap
  int num_kb;            // kh*kw*cin_blocks
  int BN, n_tiles, total_tiles;
  int tiles_w, tiles_h;  // patch grid inside an image group
  int bw, bh, bb;        // patch = bb images x bh rows x bw cols (<= 128 pixels)
  int a_bytes;           // bytes one A box delivers (= patch pixels * 128)
  int stages;
  int kbps;              // 64-wide k-blocks per pipeline stage (2 halves the barrier round trips per unit of K)
  int b_resident;        // whole weight matrix stays in shared memory (n_tiles * num_kb blocks of BN x 64)
  int halo;              // 1: stride-1 k x k conv fed from ONE halo tile per 64-channel block (see below)
  int kh;                // filter rows (halo mode walks taps itself)
  int halo_w, halo_h;    // halo tile = (bh + kh - 1) x (bw + kw - 1) pixels of 128 B
  int halo_bytes;        // bytes of one halo buffer (1024-aligned)
  int a_stages;          // halo buffers in the ring
  int act_kind;          // 0: per-channel slope (general), 1: ReLU everywhere, 2: linear everywhere
  const float* bias;     // [n_tiles*BN]
  const float* slope;    // [n_tiles*BN]
  int res_mode;          // ConvRes: PRE_ACT comes through tmR; POST_ACT_UPSAMPLE (linear tiling only) from global
  const void* res;       // POST_ACT_UPSAMPLE source
  int res_ld, rH, rW, Ho, Wo, M;
  int* watchdog;
  long long* trace;      // debug: (event, tile_iter, clock64) records of CTA 0, or null
  long long* stall;      // debug: per-CTA stall-cycle counters of the three roles (16 words per CTA), or null
};

// Stall accounting (FRB_OPT_TRACE = 2): clock64 deltas around the blocking waits of one thread per role, summed in
// registers and written once at the end (cheap enough not to move the timing, unlike the timeline below).
struct StallClock {
  bool on;
  long long t0;
  __device__ __forceinline__ void start() { if (on) t0 = clock64(); }
  __device__ __forceinline__ void stop(long long& acc) { if (on) acc += clock64() - t0; }
};

// Timeline probe (FRB_OPT_TRACE): CTA 0 appends (event id, tile_iter, clock) triples; trace[0] is the record count.
__device__ __forceinline__ void trace_event(long long* trace, int event, int tile_iter) {
  if (trace == nullptr || blockIdx.x != 0) return;
  const unsigned long long i = atomicAdd(reinterpret_cast<unsigned long long*>(trace), 1ULL);
  if (i < 4000) {
    trace[1 + 3 * i] = event;
    trace[2 + 3 * i] = tile_iter;
    trace[3 + 3 * i] = clock64();
  }
}

template <bool kBF16>
__global__ void __launch_bounds__(kTmaThreads, 1)
conv_tma_kernel(const __grid_constant__ ConvTmaArgs p, const __grid_constant__ CUtensorMap tmA,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                const __grid_constant__ CUtensorMap tmR) {
  using E = Elem16<kBF16>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  const int b_stage_bytes = p.BN * 128;
  const bool has_res = p.res_mode == RES_PRE_ACT;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + (p.halo ? p.a_stages * p.halo_bytes : p.stages * p.kbps * kAStageBytes);   // ring (or resident matrix)
  uint8_t* smem_o = smem_b + (p.b_resident ? p.n_tiles * p.num_kb : p.stages * (p.halo ? 1 : p.kbps)) * b_stage_bytes;
  uint8_t* smem_r = smem_o + kOutBufs * kSlabBytes;                                    // kResSets x 2 slabs (PRE_ACT only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_r + (has_res ? kResSets * 2 * kSlabBytes : 0));
  uint64_t* full_bar = bars;                          // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;            // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;        // [2]
  uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;   // [2]
  uint64_t* rfull_bar = bars + 2 * kMaxStages + 4;    // [kResSets]
  uint64_t* rempty_bar = bars + 2 * kMaxStages + 6;   // [kResSets]
  uint64_t* bres_bar = bars + 2 * kMaxStages + 8;     // [1]
  uint64_t* afull_bar = bars + 2 * kMaxStages + 9;    // [kMaxStages] halo ring
  uint64_t* aempty_bar = bars + 3 * kMaxStages + 9;   // [kMaxStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * kMaxStages + 9);
  float* s_bias = reinterpret_cast<float*>(bars + 4 * kMaxStages + 10);   // [256]
  float* s_slope = s_bias + 256;                                          // [256]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kMmaWarp = kEpiWarps, kTmaWarp = kEpiWarps + 1;
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(2 * p.BN)) tmem_cols <<= 1;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    if (has_res) tma_prefetch_desc(&tmR);
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tfull_bar[b], 1);
        mbar_init(&tempty_bar[b], kEpiWarps);
        mbar_init(&rfull_bar[b], 1);
        mbar_init(&rempty_bar[b], 1);
      }
      mbar_init(bres_bar, 1);
      for (int s = 0; s < kMaxStages; ++s) {
        mbar_init(&afull_bar[s], 1);
        mbar_init(&aempty_bar[s], 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_dyn(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nslab = (p.BN + kSlab - 1) / kSlab;

  // tile -> coordinates of its first pixel
  auto decode = [&](int tile, int& n_tile, int& c1, int& c2, int& c3) {
    const int m_tile = tile / p.n_tiles;
    n_tile = tile - m_tile * p.n_tiles;
    if (p.patch) {
      const int tw = m_tile % p.tiles_w;
      const int t = m_tile / p.tiles_w;
      c1 = tw * p.bw;
      c2 = (t % p.tiles_h) * p.bh;
      c3 = (t / p.tiles_h) * p.bb;
    } else {
      c1 = m_tile * kBM;
      c2 = 0;
      c3 = 0;
    }
  };

  if (warp < kEpiWarps) {
    // ================================================================ epilogue
    const int quarter = warp & 3;              // TMEM lane quarter this warp may read
    const int csub = warp >> 2;                // which 16 columns of every 64-column slab
    const int row = quarter * 32 + lane;
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t sw7 = static_cast<uint32_t>(row & 7);
    const int col0 = csub * kColsPerWarp;
    int tile_iter = 0;
    uint32_t slab_ctr = 0;
    StallClock sc{p.stall != nullptr && threadIdx.x == 0, 0};
    long long st_tfull = 0, st_res = 0, st_slab = 0, st_bar = 0;
    const long long st_begin = sc.on ? clock64() : 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tile_iter) {
      int n_tile, c1, c2, c3;
      decode(tile, n_tile, c1, c2, c3);
      const int n_base = n_tile * p.BN;
      const int buf = tile_iter & 1;
      const uint32_t tphase = (tile_iter >> 1) & 1;
      const int rset = tile_iter % kResSets;
      const uint32_t rphase = (tile_iter / kResSets) & 1;

      const typename E::T* up_row = nullptr;
      if (p.res_mode == RES_POST_ACT_UPSAMPLE) {
        const int m = c1 + row;
        if (m < p.M) {
          const int ow = m % p.Wo;
          const int t = m / p.Wo;
          const int oh = t % p.Ho;
          const int b = t / p.Ho;
          const int rh = (oh * p.rH) / p.Ho;   // nearest: floor(dst * in / out)
          const int rw = (ow * p.rW) / p.Wo;
          up_row = static_cast<const typename E::T*>(p.res) +
                   (static_cast<size_t>(b) * p.rH * p.rW + static_cast<size_t>(rh) * p.rW + rw) * p.res_ld;
        }
      }
      // bias / slope of this N tile -> shared memory (broadcast reads below); the previous tile's readers
      // are all past their slab barriers, and the first slab barrier below orders these writes before use
      for (int t = threadIdx.x; t < p.BN; t += kEpiThreads) {
        s_bias[t] = __ldg(p.bias + n_base + t);
        s_slope[t] = __ldg(p.slope + n_base + t);
      }

      if (threadIdx.x == 0) trace_event(p.trace, 10, tile_iter);        // epilogue ready for this tile
      sc.start();
      mbar_wait(&tfull_bar[buf], tphase, p.watchdog, 100 + buf);
      sc.stop(st_tfull);
      tc_fence_after();
      if (threadIdx.x == 0) trace_event(p.trace, 11, tile_iter);        // accumulator available
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * p.BN);
      sc.start();
      if (has_res) mbar_wait(&rfull_bar[rset], rphase, p.watchdog, 500 + rset);
      sc.stop(st_res);

      for (int s = 0; s < nslab; ++s) {
        const int ncols = min(kSlab, p.BN - s * kSlab);
        const int n0 = n_base + s * kSlab;
        const bool mine = col0 < ncols;            // warp-uniform (BN is a multiple of 16)
        float x[kColsPerWarp];
        if (mine) {
          uint32_t v[kColsPerWarp];
          tmem_ld16(taddr + s * kSlab + col0, v);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < kColsPerWarp; ++c) x[c] = __uint_as_float(v[c]);
        }
        if (s == nslab - 1) {                     // accumulator fully read: hand the TMEM buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        }
        // staging slab `ob` is free once the store issued kOutBufs slabs ago has finished reading it
        const uint32_t ob = slab_ctr % kOutBufs;
        if (threadIdx.x == 0) {
          trace_event(p.trace, 13, tile_iter);                              // TMEM read out, waiting for a staging slab
          sc.start();
          tma_store_wait_read<kOutBufs - 1>();
          sc.stop(st_slab);
          trace_event(p.trace, 14, tile_iter);
        }
        sc.start();
        named_bar_sync(1, kEpiThreads);
        sc.stop(st_bar);

        if (mine) {
          const uint8_t* rrow = smem_r + (rset * 2 + s) * kSlabBytes + row_off;
          uint8_t* orow = smem_o + ob * kSlabBytes + row_off;
#pragma unroll
          for (int c = 0; c < kColsPerWarp / 8; ++c) {
            float* y = x + c * 8;
            const int cc = col0 + c * 8;                                   // column inside the slab
            const uint32_t chunk = (static_cast<uint32_t>(cc >> 3) ^ sw7) << 4;
            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + s * kSlab + cc);
            const float4 b1 = *reinterpret_cast<const float4*>(s_bias + s * kSlab + cc + 4);
            y[0] += b0.x; y[1] += b0.y; y[2] += b0.z; y[3] += b0.w;
            y[4] += b1.x; y[5] += b1.y; y[6] += b1.z; y[7] += b1.w;
            if (has_res) {
              const uint4 r = *reinterpret_cast<const uint4*>(rrow + chunk);
              float2 f;
              f = E::to_f2(r.x); y[0] += f.x; y[1] += f.y;
              f = E::to_f2(r.y); y[2] += f.x; y[3] += f.y;
              f = E::to_f2(r.z); y[4] += f.x; y[5] += f.y;
              f = E::to_f2(r.w); y[6] += f.x; y[7] += f.y;
            }
            if (p.act_kind == 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) y[j] = fmaxf(y[j], 0.f);
            } else if (p.act_kind == 0) {
              const float4 s0 = *reinterpret_cast<const float4*>(s_slope + s * kSlab + cc);
              const float4 s1 = *reinterpret_cast<const float4*>(s_slope + s * kSlab + cc + 4);
              // leaky(x) = max(x,0) + slope * min(x,0): branch-free, slope 1 = identity, 0 = ReLU
              y[0] = fmaf(s0.x, fminf(y[0], 0.f), fmaxf(y[0], 0.f)); y[1] = fmaf(s0.y, fminf(y[1], 0.f), fmaxf(y[1], 0.f));
              y[2] = fmaf(s0.z, fminf(y[2], 0.f), fmaxf(y[2], 0.f)); y[3] = fmaf(s0.w, fminf(y[3], 0.f), fmaxf(y[3], 0.f));
              y[4] = fmaf(s1.x, fminf(y[4], 0.f), fmaxf(y[4], 0.f)); y[5] = fmaf(s1.y, fminf(y[5], 0.f), fmaxf(y[5], 0.f));
              y[6] = fmaf(s1.z, fminf(y[6], 0.f), fmaxf(y[6], 0.f)); y[7] = fmaf(s1.w, fminf(y[7], 0.f), fmaxf(y[7], 0.f));
            }
            if (p.res_mode == RES_POST_ACT_UPSAMPLE && up_row != nullptr) {
              const uint4 r = __ldg(reinterpret_cast<const uint4*>(up_row + n0 + cc));
              float2 f;
              f = E::to_f2(r.x); y[0] += f.x; y[1] += f.y;
              f = E::to_f2(r.y); y[2] += f.x; y[3] += f.y;
              f = E::to_f2(r.z); y[4] += f.x; y[5] += f.y;
              f = E::to_f2(r.w); y[6] += f.x; y[7] += f.y;
            }
            uint4 o;
            o.x = E::from_f2(y[0], y[1]);
            o.y = E::from_f2(y[2], y[3]);
            o.z = E::from_f2(y[4], y[5]);
            o.w = E::from_f2(y[6], y[7]);
            *reinterpret_cast<uint4*>(orow + chunk) = o;
          }
        }
        fence_proxy_async_smem();
        sc.start();
        named_bar_sync(1, kEpiThreads);
        sc.stop(st_bar);
        if (threadIdx.x == 0) {
          if (p.patch) tma_store_4d(&tmO, smem_o + ob * kSlabBytes, n0, c1, c2, c3);
          else tma_store_2d(&tmO, smem_o + ob * kSlabBytes, n0, c1);
          tma_store_commit();
          if (s == nslab - 1) trace_event(p.trace, 12, tile_iter);        // last store of the tile issued
          // every epilogue thread is past its reads of this tile's residual slabs: give the set back
          if (has_res && s == nslab - 1) mbar_arrive(&rempty_bar[rset]);
        }
        ++slab_ctr;
      }
    }
    if (sc.on) {
      long long* o = p.stall + 1 + blockIdx.x * 16;
      o[0] = clock64() - st_begin; o[1] = st_tfull; o[2] = st_res; o[3] = st_slab; o[4] = st_bar; o[5] = tile_iter;
    }
    if (threadIdx.x == 0) tma_store_wait<0>();     // stores must land before the CTA (and its smem) goes away
  } else if (warp == kMmaWarp) {
    // ============================================================== MMA issuer
    // Everything here is warp-uniform and incremental: descriptor LOW words (start address >> 4) advance by
    // constant steps, high words are constants, barriers are 32-bit shared addresses, and only the tcgen05
    // instructions sit under the elected lane.  (Measured before this rewrite: ~600 cycles of dependent
    // single-warp instructions per k-block in both this loop and the producer's, i.e. the pipeline could not
    // deliver a k-block faster than that whatever the tile shape — 4 MMAs are 225-512 cycles of tensor time.)
    const uint32_t idesc = make_idesc_f16(kBM, p.BN, kBF16 ? 1 : 0);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a)), b_lo0 = desc_lo(smem_u32(smem_b));
    const uint32_t a_step = kAStageBytes >> 4, b_step = static_cast<uint32_t>(b_stage_bytes) >> 4;
    const uint32_t halo_step = static_cast<uint32_t>(p.halo_bytes) >> 4;
    const uint32_t hi1024 = desc_hi(1024), hi_halo = desc_hi(static_cast<uint32_t>(p.halo_w) * 128u);
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), afull0 = smem_u32(afull_bar),
                   aempty0 = smem_u32(aempty_bar), tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);
    int stage = 0, astage = 0;
    uint32_t phase = 0, aphase = 0;
    uint32_t a_ring = a_lo0, b_ring = b_lo0;          // descriptor low words of the current ring slots
    int tile_iter = 0;
    StallClock sc{p.stall != nullptr, 0};
    long long st_tempty = 0, st_full = 0;
    const long long st_begin = sc.on ? clock64() : 0;
    if (p.b_resident) mbar_wait(bres_bar, 0, p.watchdog, 250);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tile_iter) {
      const int buf = tile_iter & 1;
      const uint32_t tphase = (tile_iter >> 1) & 1;
      const int n_tile = tile % p.n_tiles;
      sc.start();
      mbar_wait_a(tempty0 + buf * 8, tphase ^ 1, p.watchdog, 200 + buf);
      sc.stop(st_tempty);
      tc_fence_after();
      if (lane == 0) trace_event(p.trace, 21, tile_iter);                 // TMEM buffer free
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * p.BN);
      const uint32_t b_res = b_lo0 + static_cast<uint32_t>(n_tile * p.num_kb) * b_step;   // resident weights of this N tile
      if (p.halo) {
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          sc.start();
          mbar_wait_a(afull0 + astage * 8, aphase, p.watchdog, 220 + astage);
          sc.stop(st_full);
          tc_fence_after();
          uint32_t a_tap = a_lo0 + static_cast<uint32_t>(astage) * halo_step;          // pixel (0,0) of the halo tile
          uint32_t b_tap = b_res + static_cast<uint32_t>(cb) * b_step;                   // resident: block (tap 0, cb)
          uint32_t acc = cb != 0 ? 1u : 0u;
          for (int tr = 0; tr < p.kh; ++tr) {
            for (int ts = 0; ts < p.kw; ++ts) {
              if (!p.b_resident) {
                sc.start();
                mbar_wait_a(full0 + stage * 8, phase, p.watchdog, 210 + stage);
                sc.stop(st_full);
                tc_fence_after();
                b_tap = b_ring;
              }
              const bool last_tap = tr == p.kh - 1 && ts == p.kw - 1;
              if (elect_one()) {
                tc_mma_f16(d_tmem, pack_desc(a_tap, hi_halo), pack_desc(b_tap, hi1024), idesc, acc);
                tc_mma_f16(d_tmem, pack_desc(a_tap + 2, hi_halo), pack_desc(b_tap + 2, hi1024), idesc, 1u);
                tc_mma_f16(d_tmem, pack_desc(a_tap + 4, hi_halo), pack_desc(b_tap + 4, hi1024), idesc, 1u);
                tc_mma_f16(d_tmem, pack_desc(a_tap + 6, hi_halo), pack_desc(b_tap + 6, hi1024), idesc, 1u);
                if (!p.b_resident) tc_commit_a(empty0 + stage * 8);
                if (last_tap) {
                  tc_commit_a(aempty0 + astage * 8);
                  if (cb == p.cin_blocks - 1) tc_commit_a(tfull0 + buf * 8);
                }
              }
              __syncwarp();
              acc = 1u;
              a_tap += 8;                                                // next pixel of the halo row (128 B)
              if (p.b_resident) {
                b_tap += static_cast<uint32_t>(p.cin_blocks) * b_step;
              } else {
                b_ring += b_step;
                if (++stage == p.stages) { stage = 0; phase ^= 1; b_ring = b_lo0; }
              }
            }
            a_tap += static_cast<uint32_t>(p.halo_w - p.kw) * 8;         // first pixel of the next halo row
          }
          if (++astage == p.a_stages) { astage = 0; aphase ^= 1; }
        }
      } else {
        uint32_t b_cur = b_res;
        for (int kb0 = 0; kb0 < p.num_kb; kb0 += p.kbps) {
          const bool two = p.kbps == 2 && kb0 + 1 < p.num_kb;            // this stage carries two k-blocks
          sc.start();
          mbar_wait_a(full0 + stage * 8, phase, p.watchdog, 210 + stage);
          sc.stop(st_full);
          tc_fence_after();
          if (!p.b_resident) b_cur = b_ring;
          if (elect_one()) {
            tc_mma_f16(d_tmem, pack_desc(a_ring, hi1024), pack_desc(b_cur, hi1024), idesc, kb0 != 0 ? 1u : 0u);
            tc_mma_f16(d_tmem, pack_desc(a_ring + 2, hi1024), pack_desc(b_cur + 2, hi1024), idesc, 1u);
            tc_mma_f16(d_tmem, pack_desc(a_ring + 4, hi1024), pack_desc(b_cur + 4, hi1024), idesc, 1u);
            tc_mma_f16(d_tmem, pack_desc(a_ring + 6, hi1024), pack_desc(b_cur + 6, hi1024), idesc, 1u);
            if (two) {
              const uint32_t a2 = a_ring + a_step, b2 = b_cur + b_step;
              tc_mma_f16(d_tmem, pack_desc(a2, hi1024), pack_desc(b2, hi1024), idesc, 1u);
              tc_mma_f16(d_tmem, pack_desc(a2 + 2, hi1024), pack_desc(b2 + 2, hi1024), idesc, 1u);
              tc_mma_f16(d_tmem, pack_desc(a2 + 4, hi1024), pack_desc(b2 + 4, hi1024), idesc, 1u);
              tc_mma_f16(d_tmem, pack_desc(a2 + 6, hi1024), pack_desc(b2 + 6, hi1024), idesc, 1u);
            }
            tc_commit_a(empty0 + stage * 8);
            if (kb0 + p.kbps >= p.num_kb) tc_commit_a(tfull0 + buf * 8);
          }
          __syncwarp();
          b_cur += static_cast<uint32_t>(p.kbps) * b_step;
          a_ring += static_cast<uint32_t>(p.kbps) * a_step;
          b_ring += static_cast<uint32_t>(p.kbps) * b_step;
          if (++stage == p.stages) { stage = 0; phase ^= 1; a_ring = a_lo0; b_ring = b_lo0; }
        }
      }
      if (lane == 0) trace_event(p.trace, 23, tile_iter);                 // all MMAs of the tile issued
    }
    if (sc.on && lane == 0) {
      long long* o = p.stall + 1 + blockIdx.x * 16;
      o[6] = clock64() - st_begin; o[7] = st_tempty; o[8] = st_full;
    }
  } else {
    // ============================================================ TMA producer
    const uint32_t a_s0 = smem_u32(smem_a), b_s0 = smem_u32(smem_b), r_s0 = smem_u32(smem_r);
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), afull0 = smem_u32(afull_bar),
                   aempty0 = smem_u32(aempty_bar), rfull0 = smem_u32(rfull_bar), rempty0 = smem_u32(rempty_bar);
    int stage = 0, astage = 0;
    uint32_t phase = 0, aphase = 0;
    uint32_t a_dst = a_s0, b_dst = b_s0;
    int tile_iter = 0;
    StallClock sc{p.stall != nullptr, 0};
    long long st_empty = 0, st_rempty = 0;
    const long long st_begin = sc.on ? clock64() : 0;
    if (p.b_resident && lane == 0) {
      mbar_arrive_expect_tx(bres_bar, static_cast<uint32_t>(p.n_tiles * p.num_kb * b_stage_bytes));
      for (int nt = 0; nt < p.n_tiles; ++nt)
        for (int kb = 0; kb < p.num_kb; ++kb)
          tma_load_2d(smem_b + (nt * p.num_kb + kb) * b_stage_bytes, &tmB, bres_bar, kb * kBK, nt * p.BN);
    }
    const uint32_t tx_bytes = static_cast<uint32_t>((p.b_resident ? 0 : b_stage_bytes) + p.a_bytes);
    const uint32_t halo_tx = static_cast<uint32_t>(p.halo_w * p.halo_h * 128);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tile_iter) {
      int n_tile, c1, c2, c3;
      decode(tile, n_tile, c1, c2, c3);
      const int n_coord = n_tile * p.BN;
      if (has_res) {
        // residual tile of THIS tile into set tile_iter % kResSets, as soon as the epilogue released it
        const int rset = tile_iter % kResSets;
        sc.start();
        mbar_wait_a(rempty0 + rset * 8, ((tile_iter / kResSets) & 1) ^ 1, p.watchdog, 600 + rset);
        sc.stop(st_rempty);
        if (elect_one()) {
          mbar_expect_tx_a(rfull0 + rset * 8, static_cast<uint32_t>(nslab * p.a_bytes));
          for (int rs = 0; rs < nslab; ++rs) {
            const uint32_t dst = r_s0 + static_cast<uint32_t>((rset * 2 + rs) * kSlabBytes);
            if (p.patch) tma_load_4d_a(dst, &tmR, rfull0 + rset * 8, n_coord + rs * kSlab, c1, c2, c3);
            else tma_load_2d_a(dst, &tmR, rfull0 + rset * 8, n_coord + rs * kSlab, c1);
          }
        }
        __syncwarp();
      }
      if (p.halo) {
        const int taps = p.kh * p.kw;
        for (int hb = 0; hb < p.cin_blocks; ++hb) {
          sc.start();
          mbar_wait_a(aempty0 + astage * 8, aphase ^ 1, p.watchdog, 320 + astage);
          sc.stop(st_empty);
          if (elect_one()) {
            mbar_expect_tx_a(afull0 + astage * 8, halo_tx);
            tma_load_4d_a(a_s0 + static_cast<uint32_t>(astage * p.halo_bytes), &tmA, afull0 + astage * 8, hb * kBK, c1 - p.pw,
                          c2 - p.ph, c3);
          }
          __syncwarp();
          if (lane == 0) trace_event(p.trace, 30, tile_iter);             // halo load issued
          if (++astage == p.a_stages) { astage = 0; aphase ^= 1; }
          if (!p.b_resident) {
            int kcoord = hb * kBK;                                        // K coordinate of (tap 0, hb)
            for (int tap = 0; tap < taps; ++tap) {
              sc.start();
              mbar_wait_a(empty0 + stage * 8, phase ^ 1, p.watchdog, 300 + stage);
              sc.stop(st_empty);
              if (elect_one()) {
                mbar_expect_tx_a(full0 + stage * 8, static_cast<uint32_t>(b_stage_bytes));
                tma_load_2d_a(b_dst, &tmB, full0 + stage * 8, kcoord, n_coord);
              }
              __syncwarp();
              kcoord += p.cin_blocks * kBK;
              b_dst += static_cast<uint32_t>(b_stage_bytes);
              if (++stage == p.stages) { stage = 0; phase ^= 1; b_dst = b_s0; }
            }
          }
        }
      } else {
        const int aw = c1 * p.sw - p.pw, ah = c2 * p.sh - p.ph;
        int tap_r = 0, tap_s = 0, cb = 0, kcoord = 0;
        for (int kb0 = 0; kb0 < p.num_kb; kb0 += p.kbps) {
          const int nk = min(p.kbps, p.num_kb - kb0);
          sc.start();
          mbar_wait_a(empty0 + stage * 8, phase ^ 1, p.watchdog, 300 + stage);
          sc.stop(st_empty);
          if (elect_one()) mbar_expect_tx_a(full0 + stage * 8, tx_bytes * static_cast<uint32_t>(nk));
          for (int j = 0; j < nk; ++j) {
            if (elect_one()) {
              if (!p.b_resident) tma_load_2d_a(b_dst + static_cast<uint32_t>(j * b_stage_bytes), &tmB, full0 + stage * 8, kcoord, n_coord);
              if (p.patch) tma_load_4d_a(a_dst + static_cast<uint32_t>(j * kAStageBytes), &tmA, full0 + stage * 8, cb * kBK, aw + tap_s, ah + tap_r, c3);
              else tma_load_2d_a(a_dst + static_cast<uint32_t>(j * kAStageBytes), &tmA, full0 + stage * 8, kcoord, c1);
            }
            kcoord += kBK;
            if (++cb == p.cin_blocks) {
              cb = 0;
              if (++tap_s == p.kw) { tap_s = 0; ++tap_r; }
            }
          }
          __syncwarp();
          a_dst += static_cast<uint32_t>(p.kbps) * kAStageBytes;
          b_dst += static_cast<uint32_t>(p.kbps * b_stage_bytes);
          if (++stage == p.stages) { stage = 0; phase ^= 1; a_dst = a_s0; b_dst = b_s0; }
        }
      }
    }
    if (sc.on && lane == 0) {
      long long* o = p.stall + 1 + blockIdx.x * 16;
      o[9] = clock64() - st_begin; o[10] = st_empty; o[11] = st_rempty;
    }
  }

  // -------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, tmem_cols);
  }
}

}  // namespace frb
