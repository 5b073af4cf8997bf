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
// L2 -> SM traffic drops by the weight re-fetch of every tile.  Accumulators live in TMEM (four buffers of BN
// columns for BN <= 128, two above), and the epilogue (bias, residual, per-channel slope) packs 64-column slabs
// into swizzled staging slabs that TMA stores back as the same patch box — partial tiles are clipped by
// the hardware, and every global store is a full 128-byte line.  A pre-activation residual (ResNet / Inception
// skip) is added BY THE TENSOR CORE: after the last weight k-block the producer sends each 64-channel residual slab
// of the tile through the same stage ring as one more A block, and the issuer multiplies it by a resident 64 x 64
// identity into the slab's accumulator columns (16-bit x 1.0 is exact in the fp32 accumulator).  The skip tensor
// thus gets the deep TMA prefetch of the main loop and costs the epilogue nothing — as per-thread global loads it
// left the epilogue groups waiting on HBM latency (profiles/r01_ncu_epilogue_v7.md).
//
// Halo mode (stride-1 k x k convs): instead of one A box per filter tap (k*k reads of the same pixels through
// L2), the producer loads ONE (bh+kh-1) x (bw+kw-1) halo tile per 64-channel block and the MMA issuer points
// the A descriptor of tap (r, s) at pixel (r, s) of that tile: 8-row groups are 8 consecutive pixels of a halo
// row (bw = 8), the stride between groups is the halo row pitch.  The 128-byte swizzle is a function of the
// shared-memory ADDRESS (tools/umma_shift_probe.cu: any 128-byte row offset and any group stride read back
// exactly what TMA wrote), so shifted windows need no copy.  A traffic through L2 drops k*k-fold, and the issuer
// waits once per channel block (resident weights: all taps under one wait) or once per filter row (streamed weights,
// kw taps per B stage) instead of once per tap.  Default where it pays (conv_launch_tma), FRB_OPT_HALO overrides.
// Paired tiles (FRB_OPT_MPAIR, streamed weights with N tiles >= 192): a tile is TWO 8 x 16 patches side by side under
// one (16+kw-1)-wide halo tile; every weight stage feeds two M = 128 MMAs into two accumulator buffers, which the
// epilogue treats as two ordinary tiles.  Half the weight bytes through L2 per pixel, twice the MMAs per wait.
//
// Weight slice (FRB_OPT_WEIGHT_SLICE): when the grid is a multiple of the number of N tiles, tile % n_tiles is the
// same for every tile of a CTA, and a slice of up to 48 KB is loaded once like the small matrices above.
//
// CTA pairs (FRB_OPT_CTA2; every path except 1x1 layers with resident weights / weight slices, which are HBM-bound): two CTAs on the SMs of one TPC (cluster of 2,
// cta_group::2) work on two M tiles of the SAME N tile with ONE tcgen05.mma of M = 256: each CTA loads its own A blocks
// and residual slabs and HALF of every weight block (BN/2 rows), the leader CTA (cluster rank 0) issues the MMAs, commits
// are multicast to both CTAs' barriers, the epilogues hand the accumulators back on the leader's barrier.  The layers that
// take this path (1x1 convs of layer3/4 and the FPN, strided convs) stream 94 B/clk of operands per SM at BN = 256
// against ~53 B/clk an SM can pull through L2 beside a busy tensor pipe; half the weight bytes is 62 B/clk.
//
// Programmatic dependent launch (FRB_OPT_PDL, short launches only): barrier init, TMEM allocation, descriptor
// prefetch and the resident-weight loads run before griddepcontrol.wait, i.e. under the previous kernel's tail.
//
// Warp roles (480 threads, 15 warps = at most 4 per scheduler, i.e. up to 128 registers per thread: with ~227 KB
// of shared memory there is no L1 left and a spill is an L2 round trip): warps 0-11 are
// three independent epilogue groups (one warp per TMEM lane quarter each), 12 and 14 two MMA issuers taking tiles
// alternately (12 also allocates TMEM), 13 the TMA producer.  Profiles that led here: profiles/r01_stall_counters_v6.txt (16 warps sharing every slab
// behind two 512-thread barriers: 2 000 cycles per slab, the MMA issuer waiting for a TMEM buffer 78 % of the time).
// Persistent: one CTA per SM walks tiles round-robin (n fastest, so CTAs running at the same time
// share the A patch through L2).
#pragma once
#include "conv_igemm.cuh"

namespace frb {

constexpr int kEpiWarps = 12;             // 14 warps in all: at most 4 per scheduler (measured: 16 + 2 warps is 2-3 % slower end to end)
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kTmaThreads = kEpiThreads + 96;   // + MMA issuer A, TMA producer, MMA issuer B
constexpr int kSlab = 64;                  // epilogue / store granularity in output channels
constexpr int kSlabBytes = kBM * kSlab * 2;
constexpr int kEpiGroups = kEpiWarps / 4;  // independent epilogue groups: one warp per TMEM lane quarter each
constexpr int kGroupThreads = 128;
constexpr int kOutBufs = kEpiGroups;       // one staging slab per group
constexpr int kMaxAcc = 4;                 // TMEM accumulator buffers: 4 for BN <= 128, else 2
constexpr int kResidentWeightBytes = 80 * 1024;
// One N tile's weights resident per CTA (b_resident = 2).  The 32 KB slices of layer2's conv3 + skip take those launches
// from 0.97 to 0.75 ms (HBM-bound at 6.4 TB/s); the 64 KB slice of layer3's conv3 measured neutral: it costs ring depth
// and with it the second MMA issuer (profiles/r01_pdl_ab.txt).
constexpr int kResidentSliceBytes = 48 * 1024;
constexpr int kIdentBytes = 64 * 128;      // identity B operand of the residual MMAs

struct ConvTmaArgs {
  int patch;             // 0: linear M tiling (1x1, stride 1, no padding); 1: spatial patch tiling
  int kw, sh, sw, ph, pw;
  int cin_blocks;        // 64-channel blocks per filter tap
  int num_kb;            // kh*kw*cin_blocks
  int BN, n_tiles, total_tiles;
  uint32_t fd_n[2], fd_w[2], fd_h[2];   // fast-division (multiplier, shift) pairs for n_tiles, tiles_w, tiles_h
  int tiles_w, tiles_h;  // patch grid inside an image group
  int bw, bh, bb;        // patch = bb images x bh rows x bw cols (<= 128 pixels)
  int a_bytes;           // bytes one A box delivers (= patch pixels * 128)
  int stages;
  int kbps;              // 64-wide k-blocks per pipeline stage (2 halves the barrier round trips per unit of K)
  int b_resident;        // 1: whole weight matrix stays in shared memory (n_tiles * num_kb blocks of BN x 64); 2: only the
                         // N tile of this CTA (the grid is a multiple of n_tiles, so a CTA never sees another one)
  int halo;              // 1: stride-1 k x k conv fed from ONE halo tile per 64-channel block (see below)
  int kh;                // filter rows (halo mode walks taps itself)
  int halo_w, halo_h;    // halo tile = (bh + kh - 1) x (bw + kw - 1) pixels of 128 B
  int halo_bytes;        // bytes of one halo buffer (1024-aligned)
  int a_stages;          // halo buffers in the ring
  int brow;              // halo mode, streamed weights: filter taps per B stage (1, or kw = a whole filter row under one wait)
  int act_kind;          // 0: per-channel slope (general), 1: ReLU everywhere, 2: linear everywhere
  const float* bias;     // [n_tiles*BN]
  const float* slope;    // [n_tiles*BN]
  int res_mode;          // ConvRes: both kinds are read from global by the epilogue threads (PRE_ACT one slab ahead)
  const void* res;       // residual tensor [B, Ho, Wo, res_ld] (PRE_ACT) or [B, rH, rW, res_ld] (POST_ACT_UPSAMPLE)
  int B;                 // images (patch tiling bounds check of the residual rows)
  int res_ld, rH, rW, Ho, Wo, M;
  int* watchdog;
  int dual_issue;        // 1: two MMA issuer warps take tiles alternately where the rings allow it (FRB_OPT_DUAL_ISSUE)
  int mpair;             // 1: halo mode, a tile is two 8-wide patches side by side under one halo tile and one weight stream
  int pdl;               // 1: launched with programmatic stream serialization (FRB_OPT_PDL)
  int cta2;              // 1: CTA pairs (cluster of 2): M = 256 MMAs, weights split over the pair (tmB / tmI boxes are half size)
  int skip_loads;        // debug build, stall variants only (FRB_OPT_SKIP_LOADS): the producer only arrives on the full barriers
  long long* stall;      // debug: per-CTA stall-cycle counters of the three roles (16 words per CTA), or null
  // fused RetinaFace heads (1x1, 32 columns = cls 4 | box 8 | ldm 20 per pixel): fp32 rows written straight from the
  // accumulator registers to loc [B,P,4] / conf [B,P,2] / landm [B,P,10] at prior P_off + 2 * pixel (linear tiling only)
  float* head_loc;
  float* head_conf;
  float* head_landm;
  int head_P_total, head_P_off, head_hw;
  // fused 3x3 / stride 2 / pad 1 max pool of the (ReLU) output (kPool variant: halo mode, 8 x 16 patches): the conv output
  // itself is never stored; pool_out is [B, pool_Ho, pool_Wo, pool_ld] 16-bit, zero at every position with row % 8 == 0 or
  // col % 4 == 0 before the launch (those take max-reductions from up to four tiles)
  void* pool_out;
  int pool_Ho, pool_Wo, pool_ld;
};

// Host entry points (conv_tma_f16.cu / conv_tma_bf16.cu / conv_tma_stall.cu): pick the kernel variant for
// (res = ConvRes, act = 0 per-channel slope / 1 ReLU / 2 linear) and launch it.
cudaError_t conv_tma_launch_f16(int res, int act, int grid, size_t smem, cudaStream_t st, const ConvTmaArgs& a,
                                const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR,
    const CUtensorMap& tmI);
cudaError_t conv_tma_launch_bf16(int res, int act, int grid, size_t smem, cudaStream_t st, const ConvTmaArgs& a,
                                 const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR,
    const CUtensorMap& tmI);
cudaError_t conv_tma_launch_f16_pair(int res, int act, int grid, size_t smem, cudaStream_t st, const ConvTmaArgs& a,
                                     const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR,
    const CUtensorMap& tmI);
cudaError_t conv_tma_launch_bf16_pair(int res, int act, int grid, size_t smem, cudaStream_t st, const ConvTmaArgs& a,
                                      const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR,
    const CUtensorMap& tmI);
cudaError_t conv_tma_launch_stall(int res, int act, int grid, size_t smem, cudaStream_t st, const ConvTmaArgs& a,
                                  const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR,
    const CUtensorMap& tmI);

// Stall accounting (FRB_OPT_TRACE = 2): clock64 deltas around the blocking waits of one thread per role, summed in
// registers and written once at the end (cheap enough not to move the timing, unlike the timeline below).
template <bool kOn>
struct StallClock {
  bool on;
  long long t0;
  __device__ __forceinline__ void start() { if (kOn && on) t0 = clock64(); }
  __device__ __forceinline__ void stop(long long& acc) { if (kOn && on) acc += clock64() - t0; }
};

// n / d for n < 2^31 with (mul, shift) from fastdiv_make (host): three instructions instead of ~25
__device__ __forceinline__ int fastdiv(int n, const uint32_t* fd) {
  return static_cast<int>((__umulhi(static_cast<uint32_t>(n), fd[0]) + static_cast<uint32_t>(n)) >> fd[1]);
}

// kRes: 0 no residual, 1 residual added before the activation (ResNet / Inception skip), 2 nearest-upsampled
// residual added after it (FPN top-down).  kAct: 0 per-channel slope, 1 ReLU everywhere, 2 linear everywhere.
// 16-byte max-reduction of eight 16-bit floats into global memory (REDG.E.MAX.F16x8 / BF16x8)
template <bool kBF16>
__device__ __forceinline__ void red_max8(void* dst, const uint4& v) {
  if constexpr (kBF16)
    asm volatile("red.global.max.noftz.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else
    asm volatile("red.global.max.noftz.v4.f16x2 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <bool kBF16, bool kStall, int kRes, int kAct, bool kPair, bool kPool = false>
__global__ void __launch_bounds__(kTmaThreads, 1)
conv_tma_kernel(const __grid_constant__ ConvTmaArgs p, const __grid_constant__ CUtensorMap tmA,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmI) {
  using E = Elem16<kBF16>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  // CTA pairs are a separate instantiation: every tcgen05 instruction of a kernel must carry the same cta_group, and a
  // kernel that uses cta_group::2 can only be launched as clusters of two
  constexpr bool cta2 = kPair;
  uint32_t crank = 0u;                                            // rank in the CTA pair (leader = 0)
  if constexpr (cta2) crank = cluster_ctarank();
  const int tile_unit = cta2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);   // who walks the tile list
  const int grid_units = cta2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int b_rows = cta2 ? p.BN / 2 : p.BN;                     // weight rows this CTA holds per k-block
  const int b_stage_bytes = b_rows * 128;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + (p.halo ? p.a_stages * p.halo_bytes : p.stages * p.kbps * kAStageBytes);   // ring (or resident matrix)
  const int res_nts = p.b_resident == 2 ? 1 : p.n_tiles;                                // resident N tiles and the first of them
  const int res_nt0 = p.b_resident == 2 ? tile_unit % p.n_tiles : 0;
  uint8_t* smem_o = smem_b + (p.b_resident ? res_nts * p.num_kb : p.stages * (p.halo ? p.brow : p.kbps)) * b_stage_bytes;
  uint8_t* smem_i = smem_o + kOutBufs * kSlabBytes;                                    // 64 x 64 identity (kRes = 1)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_i + (kRes == 1 ? kIdentBytes : 0));
  uint64_t* full_bar = bars;                          // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;            // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;        // [kMaxAcc]
  uint64_t* tempty_bar = bars + 2 * kMaxStages + 4;   // [kMaxAcc]
  uint64_t* bres_bar = bars + 2 * kMaxStages + 16;    // [1] resident weights + identity have landed
  uint64_t* afull_bar = bars + 2 * kMaxStages + 17;   // [kMaxStages] halo ring
  uint64_t* aempty_bar = bars + 3 * kMaxStages + 17;  // [kMaxStages]
  uint64_t* turn_bar = bars + 4 * kMaxStages + 17;    // [2] issuer w has finished issuing a tile (dual issue hand-over)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * kMaxStages + 19);
  float* s_bias = reinterpret_cast<float*>(bars + 4 * kMaxStages + 20);   // [kEpiGroups][kSlab]
  float* s_slope = s_bias + kEpiGroups * kSlab;                           // [kEpiGroups][kSlab]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kMmaWarp = kEpiWarps, kTmaWarp = kEpiWarps + 1, kMmaWarp2 = kEpiWarps + 2;
  const int nslab = (p.BN + kSlab - 1) / kSlab;
  const int nacc = p.BN <= 128 ? kMaxAcc : 2;          // accumulator buffers in TMEM (nacc * BN <= 512 columns)
  const int nacc_shift = p.BN <= 128 ? 2 : 1;          // log2(nacc): the per-tile bookkeeping of the single-warp roles avoids divisions
  // Two MMA issuer warps take the CTA's tiles alternately when two consecutive tiles fit in one pass over each ring
  // (see the issuer role).  Ring slots a tile consumes:
  const int tile_stages = p.halo ? (p.b_resident ? 0 : p.cin_blocks * (p.brow > 1 ? p.kh : p.kh * p.kw))
                                 : (p.num_kb + p.kbps - 1) / p.kbps + (kRes == 1 ? nslab : 0);
  const int tile_astages = p.halo ? p.cin_blocks : 0;
  const bool dual = p.dual_issue == 1 && !p.mpair && 2 * tile_stages <= p.stages && (!p.halo || 2 * tile_astages <= p.a_stages);
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(nacc * p.BN)) tmem_cols <<= 1;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    if (kRes == 1) {
      tma_prefetch_desc(&tmR);
      tma_prefetch_desc(&tmI);
    }
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < kMaxAcc; ++b) {
        mbar_init(&tfull_bar[b], 1);
        mbar_init(&tempty_bar[b], 4 * nslab * (cta2 ? 2 : 1));   // every warp of the nslab groups that read the buffer (both CTAs of a pair)
      }
      mbar_init(bres_bar, 1);
      mbar_init(&turn_bar[0], 1);
      mbar_init(&turn_bar[1], 1);
      for (int s = 0; s < kMaxStages; ++s) {
        mbar_init(&afull_bar[s], 1);
        mbar_init(&aempty_bar[s], 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    if constexpr (cta2) {
      tmem_alloc_2cta(tmem_slot, tmem_cols);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc_dyn(tmem_slot, tmem_cols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (cta2) cluster_sync_all();       // barriers initialised and TMEM allocated in both CTAs before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) touches no
  // global memory and may run under the tail of the previous kernel of the stream.  The roles that touch global
  // memory (TMA producer, epilogue) wait for that kernel below; the producer fetches the constant operands first.
  if (p.pdl) griddep_launch_dependents();

  // tile -> coordinates of its first pixel
  auto decode = [&](int tile, int& n_tile, int& c1, int& c2, int& c3) {
    int m_tile = fastdiv(tile, p.fd_n);
    n_tile = tile - m_tile * p.n_tiles;
    if (cta2) m_tile = 2 * m_tile + static_cast<int>(crank);   // the pair's two M tiles (one past the end: all loads and stores clip)
    if (p.patch) {
      const int t = fastdiv(m_tile, p.fd_w);
      const int tb = fastdiv(t, p.fd_h);
      c1 = (m_tile - t * p.tiles_w) * p.bw;
      c2 = (t - tb * p.tiles_h) * p.bh;
      c3 = tb * p.bb;
    } else {
      c1 = m_tile * kBM;
      c2 = 0;
      c3 = 0;
    }
  };

  if (warp < kEpiWarps) {
    // ================================================================ epilogue
    // Independent groups of four warps (one warp per TMEM lane quarter).  Group g owns staging slab g, named
    // barrier 1 + g and every 64-column output slab whose running index q (= tile_iter * nslab + slab) has
    // q % kEpiGroups == g, so several slabs — of different tiles when BN = 64 — are in flight at once and nothing
    // in here synchronises more than 128 threads.  The bias / slope values of a slab are fetched ONE SLAB AHEAD,
    // right after the previous slab's math.  A pre-activation residual never shows up here: the MMA issuer has
    // already added it into the accumulator (see there).  Residual kind and activation are template parameters:
    // as run-time flags they cost ~450 predicated-off or branch instructions per warp and slab
    // (profiles/r01_ncu_epilogue_v7.md).
    const int quarter = warp & 3;              // TMEM lane quarter this warp may read
    const int grp = warp >> 2;
    const int gtid = threadIdx.x & (kGroupThreads - 1);
    const bool leader = gtid == 0;
    const int row = quarter * 32 + lane;
    const uint32_t sw7 = static_cast<uint32_t>(row & 7);
    uint8_t* const orow = smem_o + grp * kSlabBytes + static_cast<uint32_t>(row) * 128u;
    float* const g_bias = s_bias + grp * kSlab;
    float* const g_slope = s_slope + grp * kSlab;
    // (a paired tile is two accumulators = two "tiles" of this role: the same slabs, the second 8 pixels to the right)
    const int my_tiles = ((p.total_tiles - tile_unit + grid_units - 1) / grid_units) << (p.mpair ? 1 : 0);
    uint32_t tempty_leader = 0u;
    if constexpr (cta2) tempty_leader = mapa_shared(smem_u32(tempty_bar), 0);
    const int my_slabs = my_tiles * nslab;
    StallClock<kStall> sc{kStall && p.stall != nullptr && threadIdx.x == 0, 0};
    long long st_tfull = 0, st_res = 0, st_slab = 0, st_bar = 0;
    const long long st_begin = kStall && sc.on ? clock64() : 0;
    griddep_wait();                            // the upsampled residual is read, the output written, from here

    struct Slab { int tile_iter, s, n0, ncols, c1, c2, c3; };
    auto slab_at = [&](int ti, int s2) {
      Slab sl;
      int nt;
      decode(tile_unit + (p.mpair ? ti >> 1 : ti) * grid_units, nt, sl.c1, sl.c2, sl.c3);
      if (p.mpair) sl.c1 += (ti & 1) * 8;
      sl.tile_iter = ti;
      sl.s = s2;
      sl.n0 = nt * p.BN + s2 * kSlab;
      sl.ncols = min(kSlab, p.BN - s2 * kSlab);
      return sl;
    };
    // kPool: a thread's pooling items (pooled position, 16-byte channel chunk) are the same for every patch, so their
    // shared-memory offsets (window clamped into the patch) and the output offset are worked out once
    // (three items per thread at most: 45 positions x 8 chunks over 128 threads), packed into three registers per item
    uint32_t pk_row[3] = {0u, 0u, 0u}, pk_col[3] = {0u, 0u, 0u}, pk_dst[3] = {0u, 0u, 0u};
    if constexpr (kPool) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int it = gtid + k * kGroupThreads;
        if (it < 45 * 8) {
          const int pos = it >> 3, ch = it & 7;
          const int pi = pos / 5, pj = pos - pi * 5;
          uint32_t rows = 0u, cols = 0u;
#pragma unroll
          for (int d = 0; d < 3; ++d) {                    // window rows / columns clamped into the patch
            const int hc = min(max(2 * pi - 1 + d, 0), 15);
            const int wc = min(max(2 * pj - 1 + d, 0), 7);
            rows |= static_cast<uint32_t>(hc) << (4 * d);
            cols |= static_cast<uint32_t>(wc * 128 + ((ch ^ wc) << 4)) << (10 * d);
          }
          const uint32_t inner = (pi >= 1 && pi <= 7 && pj >= 1 && pj <= 3) ? 1u : 0u;
          pk_row[k] = rows | (inner << 21) | (1u << 22);
          pk_col[k] = cols;
          pk_dst[k] = static_cast<uint32_t>((pi * p.pool_Wo + pj) * p.pool_ld + ch * 8);
        }
      }
    }
    float next_bs = 0.f;                       // bias (threads 0-63 of the group) / slope (64-127) of the next slab
    auto prefetch = [&](const Slab& sl) {
      const int col = gtid & (kSlab - 1);
      next_bs = 0.f;
      if (col < sl.ncols && (kAct == 0 || gtid < kSlab)) next_bs = __ldg((gtid < kSlab ? p.bias : p.slope) + sl.n0 + col);
    };
    int ti = grp / nslab, s_in = grp - ti * nslab;
    Slab cur = {};
    if (grp < my_slabs) {
      cur = slab_at(ti, s_in);
      prefetch(cur);
    }
    for (int q = grp; q < my_slabs; q += kEpiGroups) {
      const int buf = cur.tile_iter & (nacc - 1);                       // nacc is 2 or 4
      const uint32_t tphase = static_cast<uint32_t>(cur.tile_iter >> nacc_shift) & 1u;
      const int ncols = cur.ncols;

      const typename E::T* up_row = nullptr;
      if constexpr (kRes == 2) {
        const int m = cur.c1 + row;
        if (m < p.M) {
          const int ow = m % p.Wo;
          const int t = m / p.Wo;
          const int oh = t % p.Ho;
          const int b = t / p.Ho;
          const int rh = (oh * p.rH) / p.Ho;   // nearest: floor(dst * in / out)
          const int rw = (ow * p.rW) / p.Wo;
          up_row = static_cast<const typename E::T*>(p.res) +
                   (static_cast<size_t>(b) * p.rH * p.rW + static_cast<size_t>(rh) * p.rW + rw) * p.res_ld + cur.n0;
        }
      }
      // bias / slope of this slab -> the group's shared copy (broadcast reads below); the group's previous readers
      // are past barrier B of the previous slab, barrier A below orders these writes before use
      if (gtid < kSlab) g_bias[gtid] = next_bs;
      else if (kAct == 0) g_slope[gtid - kSlab] = next_bs;

      if (kStall && leader && p.watchdog) p.watchdog[16 + 8 * blockIdx.x + 3 + grp] = q * 10 + 1;
      sc.start();
      mbar_wait(&tfull_bar[buf], tphase, p.watchdog, 100 + buf);
      sc.stop(st_tfull);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * p.BN + cur.s * kSlab);
      uint32_t v0[16], v1[16];
      tmem_ld16(taddr, v0);                    // in flight while the staging slab is being secured
      if (leader) {
        sc.start();
        tma_store_wait_read<0>();              // the previous store of this group has finished reading the slab
        sc.stop(st_slab);
      }
      sc.start();
      named_bar_sync(1 + grp, kGroupThreads);  // barrier A
      sc.stop(st_bar);

      auto finish16 = [&](const uint32_t* v, int cbase) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float y[8];
          const int cc = cbase + c * 8;                                  // column inside the slab
          uint8_t* const dst = orow + ((static_cast<uint32_t>(cc >> 3) ^ sw7) << 4);
          const float4 b0 = *reinterpret_cast<const float4*>(g_bias + cc);
          const float4 b1 = *reinterpret_cast<const float4*>(g_bias + cc + 4);
          y[0] = __uint_as_float(v[c * 8 + 0]) + b0.x; y[1] = __uint_as_float(v[c * 8 + 1]) + b0.y;
          y[2] = __uint_as_float(v[c * 8 + 2]) + b0.z; y[3] = __uint_as_float(v[c * 8 + 3]) + b0.w;
          y[4] = __uint_as_float(v[c * 8 + 4]) + b1.x; y[5] = __uint_as_float(v[c * 8 + 5]) + b1.y;
          y[6] = __uint_as_float(v[c * 8 + 6]) + b1.z; y[7] = __uint_as_float(v[c * 8 + 7]) + b1.w;
          if constexpr (kAct == 0) {
            const float4 s0 = *reinterpret_cast<const float4*>(g_slope + cc);
            const float4 s1 = *reinterpret_cast<const float4*>(g_slope + cc + 4);
            // leaky(x) = max(x,0) + slope * min(x,0): branch-free, slope 1 = identity, 0 = ReLU
            y[0] = fmaf(s0.x, fminf(y[0], 0.f), fmaxf(y[0], 0.f)); y[1] = fmaf(s0.y, fminf(y[1], 0.f), fmaxf(y[1], 0.f));
            y[2] = fmaf(s0.z, fminf(y[2], 0.f), fmaxf(y[2], 0.f)); y[3] = fmaf(s0.w, fminf(y[3], 0.f), fmaxf(y[3], 0.f));
            y[4] = fmaf(s1.x, fminf(y[4], 0.f), fmaxf(y[4], 0.f)); y[5] = fmaf(s1.y, fminf(y[5], 0.f), fmaxf(y[5], 0.f));
            y[6] = fmaf(s1.z, fminf(y[6], 0.f), fmaxf(y[6], 0.f)); y[7] = fmaf(s1.w, fminf(y[7], 0.f), fmaxf(y[7], 0.f));
          }
          if constexpr (kRes == 2) {
            if (up_row != nullptr) {
              const uint4 r = __ldg(reinterpret_cast<const uint4*>(up_row + cc));
              float2 f;
              f = E::to_f2(r.x); y[0] += f.x; y[1] += f.y;
              f = E::to_f2(r.y); y[2] += f.x; y[3] += f.y;
              f = E::to_f2(r.z); y[4] += f.x; y[5] += f.y;
              f = E::to_f2(r.w); y[6] += f.x; y[7] += f.y;
            }
          }
          uint4 o;
          if constexpr (kAct == 1) {
            // ReLU inside the conversion (cvt.rn.relu): rounding is monotonic and keeps the sign, so this is max(x, 0) rounded
            o.x = E::from_f2_relu(y[0], y[1]);
            o.y = E::from_f2_relu(y[2], y[3]);
            o.z = E::from_f2_relu(y[4], y[5]);
            o.w = E::from_f2_relu(y[6], y[7]);
          } else {
            o.x = E::from_f2(y[0], y[1]);
            o.y = E::from_f2(y[2], y[3]);
            o.z = E::from_f2(y[4], y[5]);
            o.w = E::from_f2(y[6], y[7]);
          }
          *reinterpret_cast<uint4*>(dst) = o;
        }
      };
      if constexpr (kRes == 0 && kAct == 2) {
        if (p.head_loc != nullptr) {
          // head layer: 32 fp32 columns per pixel go from the registers to the three prior-major outputs (16-byte
          // stores, consecutive lanes = consecutive pixels = consecutive addresses); no staging slab, no TMA store
          float y[32];
          tc_wait_ld16(v0);
          tmem_ld16(taddr + 16, v1);
#pragma unroll
          for (int c = 0; c < 16; ++c) y[c] = __uint_as_float(v0[c]) + g_bias[c];
          tc_wait_ld16(v1);
#pragma unroll
          for (int c = 0; c < 16; ++c) y[16 + c] = __uint_as_float(v1[c]) + g_bias[16 + c];
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (cta2) mbar_arrive_cluster(tempty_leader + static_cast<uint32_t>(buf) * 8u);
            else mbar_arrive(&tempty_bar[buf]);
          }
          const int m = cur.c1 + row;
          if (m < p.M) {
            const int hb = m / p.head_hw;
            const size_t prior0 = static_cast<size_t>(hb) * p.head_P_total + p.head_P_off + 2 * static_cast<size_t>(m - hb * p.head_hw);
            *reinterpret_cast<float4*>(p.head_conf + prior0 * 2) = make_float4(y[0], y[1], y[2], y[3]);
            float4* lc = reinterpret_cast<float4*>(p.head_loc + prior0 * 4);
            lc[0] = make_float4(y[4], y[5], y[6], y[7]);
            lc[1] = make_float4(y[8], y[9], y[10], y[11]);
            float4* lm = reinterpret_cast<float4*>(p.head_landm + prior0 * 10);
#pragma unroll
            for (int t = 0; t < 5; ++t) lm[t] = make_float4(y[12 + 4 * t], y[13 + 4 * t], y[14 + 4 * t], y[15 + 4 * t]);
          }
          s_in += kEpiGroups;
          while (s_in >= nslab) { s_in -= nslab; ++ti; }
          if (q + kEpiGroups < my_slabs) {
            cur = slab_at(ti, s_in);
            prefetch(cur);
          }
          named_bar_sync(1 + grp, kGroupThreads);   // barrier B: the group's bias copy may be overwritten
          continue;
        }
      }
      // 16 columns at a time, the next tcgen05.ld in flight while the current chunk is finished (ncols % 16 == 0;
      // columns past ncols of a partial last slab are never stored: the output tensor map clips them)
      tc_wait_ld16(v0);
      if (ncols > 16) tmem_ld16(taddr + 16, v1);
      finish16(v0, 0);
      if (ncols > 16) {
        tc_wait_ld16(v1);
        if (ncols > 32) tmem_ld16(taddr + 32, v0);
        finish16(v1, 16);
        if (ncols > 32) {
          tc_wait_ld16(v0);
          if (ncols > 48) tmem_ld16(taddr + 48, v1);
          finish16(v0, 32);
          if (ncols > 48) {
            tc_wait_ld16(v1);
            finish16(v1, 48);
          }
        }
      }
      // this warp's part of the accumulator slab is in registers / staged: hand the TMEM buffer back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (cta2) mbar_arrive_cluster(tempty_leader + static_cast<uint32_t>(buf) * 8u);   // the leader's issuer owns the MMAs
        else mbar_arrive(&tempty_bar[buf]);
      }
      // global reads of the group's next slab
      s_in += kEpiGroups;
      while (s_in >= nslab) { s_in -= nslab; ++ti; }
      const Slab done = cur;
      if (q + kEpiGroups < my_slabs) {
        cur = slab_at(ti, s_in);
        prefetch(cur);
      }
      fence_proxy_async_smem();
      if (kStall && leader && p.watchdog) p.watchdog[16 + 8 * blockIdx.x + 3 + grp] = q * 10 + 5;
      sc.start();
      named_bar_sync(1 + grp, kGroupThreads);  // barrier B
      sc.stop(st_bar);
      if constexpr (kPool) {
        // Max pool 3x3 / stride 2 / pad 1 straight from the staged 8 x 16 patch (row = 8 h + w, 128 B per pixel, chunk c of
        // row r at (c ^ (r & 7)) * 16): the patch touches 9 x 5 pooled positions; the 7 x 3 inner ones see their whole
        // window here and are stored, the others are max-REDUCED into the pre-zeroed output together with the partial
        // maxima of the neighbouring tiles (values are >= 0 after the ReLU, max is exact and order independent: the pooled
        // tensor has the bits of the separate max-pool launch).  The conv output itself never reaches memory.
        const uint8_t* const slab = smem_o + grp * kSlabBytes;
        const int w0 = done.c1, h0 = done.c2;
        const int pi0 = h0 >> 1, pj0 = w0 >> 1;
        typename E::T* const pout = static_cast<typename E::T*>(p.pool_out) + static_cast<size_t>(done.c3) * p.pool_Ho * p.pool_Wo * p.pool_ld + done.n0;
        const int nchunk = done.ncols >> 3;
        const bool whole = h0 + 16 <= p.Ho && w0 + 8 <= p.Wo && pi0 + 9 <= p.pool_Ho && pj0 + 5 <= p.pool_Wo && nchunk == 8;
        if (whole) {
          // the patch lies inside the image: only the precomputed in-patch validity matters
          typename E::T* const ptile = pout + (static_cast<size_t>(pi0) * p.pool_Wo + pj0) * p.pool_ld;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const uint32_t pr = pk_row[k], pc = pk_col[k];
            if (pr == 0u) continue;
            const uint8_t* const rp[3] = {slab + ((pr & 15u) << 10), slab + (((pr >> 4) & 15u) << 10), slab + (((pr >> 8) & 15u) << 10)};
            const uint32_t cp[3] = {pc & 1023u, (pc >> 10) & 1023u, pc >> 20};
            // a window row / column outside the patch was clamped onto the nearest one inside, which belongs to the same
            // window: duplicates do not change a max, so all nine loads are unconditional (and in flight together)
            uint4 v[9];
#pragma unroll
            for (int dh = 0; dh < 3; ++dh)
#pragma unroll
              for (int dw = 0; dw < 3; ++dw) v[dh * 3 + dw] = *reinterpret_cast<const uint4*>(rp[dh] + cp[dw]);
            uint4 m = v[0];
#pragma unroll
            for (int t9 = 1; t9 < 9; ++t9) {
              m.x = E::max2(m.x, v[t9].x); m.y = E::max2(m.y, v[t9].y); m.z = E::max2(m.z, v[t9].z); m.w = E::max2(m.w, v[t9].w);
            }
            typename E::T* const dst = ptile + pk_dst[k];
            if (pr & (1u << 21)) *reinterpret_cast<uint4*>(dst) = m;
            else red_max8<kBF16>(dst, m);
          }
        } else
        for (int it = gtid; it < 45 * 8; it += kGroupThreads) {
          const int pos = it >> 3, ch = it & 7;
          const int pi = pos / 5, pj = pos - pi * 5;
          const int gi = pi0 + pi, gj = pj0 + pj;
          if (gi >= p.pool_Ho || gj >= p.pool_Wo || ch >= nchunk) continue;
          // nine unconditional loads from clamped coordinates (all in flight at once), then a masked max: a window pixel
          // outside the patch or the image counts as 0, the identity for non-negative values
          uint4 v[9];
          bool ok[9];
#pragma unroll
          for (int dh = 0; dh < 3; ++dh) {
            const int h = 2 * pi - 1 + dh;
            const int hc = min(max(h, 0), 15);
            const bool okh = h == hc && h0 + h < p.Ho;
#pragma unroll
            for (int dw = 0; dw < 3; ++dw) {
              const int w = 2 * pj - 1 + dw;
              const int wc = min(max(w, 0), 7);
              const int r = hc * 8 + wc;
              v[dh * 3 + dw] = *reinterpret_cast<const uint4*>(slab + r * 128 + ((ch ^ (r & 7)) << 4));
              ok[dh * 3 + dw] = okh && w == wc && w0 + w < p.Wo;
            }
          }
          uint4 m = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int t9 = 0; t9 < 9; ++t9) {
            if (ok[t9]) {
              m.x = E::max2(m.x, v[t9].x); m.y = E::max2(m.y, v[t9].y); m.z = E::max2(m.z, v[t9].z); m.w = E::max2(m.w, v[t9].w);
            }
          }
          typename E::T* const dst = pout + (static_cast<size_t>(gi) * p.pool_Wo + gj) * p.pool_ld + ch * 8;
          if (pi >= 1 && pi <= 7 && pj >= 1 && pj <= 3) *reinterpret_cast<uint4*>(dst) = m;
          else red_max8<kBF16>(dst, m);
        }
      } else if (leader) {
        if (p.patch) tma_store_4d(&tmO, smem_o + grp * kSlabBytes, done.n0, done.c1, done.c2, done.c3);
        else tma_store_2d(&tmO, smem_o + grp * kSlabBytes, done.n0, done.c1);
        tma_store_commit();
      }
    }
    if (kStall && sc.on) {
      long long* o = p.stall + 1 + blockIdx.x * 16;
      o[0] = clock64() - st_begin; o[1] = st_tfull; o[2] = st_res; o[3] = st_slab; o[4] = st_bar; o[5] = my_tiles;
    }
    if (leader) tma_store_wait<0>();           // stores must land before the CTA (and its smem) goes away
  } else if (warp == kMmaWarp || warp == kMmaWarp2) {
    // ============================================================== MMA issuer
    // Everything here is warp-uniform and incremental: descriptor LOW words (start address >> 4) advance by
    // constant steps, high words are constants, barriers are 32-bit shared addresses, and only the tcgen05
    // instructions sit under the elected lane.  (Measured before this rewrite: ~600 cycles of dependent
    // single-warp instructions per k-block in both this loop and the producer's, i.e. the pipeline could not
    // deliver a k-block faster than that whatever the tile shape — 4 MMAs are 225-512 cycles of tensor time.)
    const uint32_t idesc = make_idesc_f16(cta2 ? 2 * kBM : kBM, p.BN, kBF16 ? 1 : 0);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a)), b_lo0 = desc_lo(smem_u32(smem_b));
    const uint32_t a_step = kAStageBytes >> 4, b_step = static_cast<uint32_t>(b_stage_bytes) >> 4;
    const uint32_t halo_step = static_cast<uint32_t>(p.halo_bytes) >> 4;
    const uint32_t hi1024 = desc_hi(1024), hi_halo = desc_hi(static_cast<uint32_t>(p.halo_w) * 128u);
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), afull0 = smem_u32(afull_bar),
                   aempty0 = smem_u32(aempty_bar), tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);
    int stage = 0, astage = 0;
    uint32_t phase = 0, aphase = 0;
    uint32_t a_ring = a_lo0, b_ring = b_lo0;          // descriptor low words of the current ring slots
    // Two issuer warps take the CTA's tiles alternately (A the even tile_iters, B the odd ones): the barrier waits and
    // the per-tile bookkeeping of one (650-850 cycles per tile, as long as 16 N=64 MMAs) overlap with the other's MMAs.
    // They use different TMEM buffers, their own commits, and ring positions that each warp advances past the other's
    // tile (a tile always consumes the same number of ring slots).  The MMAs themselves still go out in tile order —
    // a warp waits for the other's "tile issued" arrival right before its first MMA — because every other barrier in
    // here tells phases apart by parity only and relies on tiles completing in order (an epilogue group that ran
    // ahead of a lagging issuer would take the previous phase of a TMEM buffer for its own).
    const int iw = warp == kMmaWarp ? 0 : 1;
    const int b_slots = p.halo ? p.brow : p.kbps;     // k-blocks (taps) of weights per ring stage
    auto skip_tile = [&]() {                           // move the ring positions past one tile issued by the other warp
      stage += tile_stages;
      while (stage >= p.stages) { stage -= p.stages; phase ^= 1; }
      a_ring = a_lo0 + static_cast<uint32_t>(stage * p.kbps) * a_step;
      b_ring = b_lo0 + static_cast<uint32_t>(stage * b_slots) * b_step;
      astage += tile_astages;
      while (p.halo && astage >= p.a_stages) { astage -= p.a_stages; aphase ^= 1; }
    };
    // mbarrier waits tell phases apart by parity only: a waiter must never be two phases ahead of a barrier.  With two
    // issuers that holds as long as two consecutive tiles fit in one pass over each ring; longer tiles keep one issuer
    // (their per-tile overheads are amortised over many MMAs anyway).
    const int solo = p.dual_issue == 2 ? 1 : 0;      // debug: which warp issues when there is only one issuer
    if (iw == 1 && dual) skip_tile();
    const uint32_t turn_other = smem_u32(&turn_bar[iw ^ 1]);
    int tile_iter = dual ? iw : 0;
    const int tstep = dual ? 2 : 1;
    StallClock<kStall> sc{kStall && p.stall != nullptr && iw == 0, 0};
    long long st_tempty = 0, st_full = 0;
    const long long st_begin = kStall && sc.on ? clock64() : 0;
    const bool issues = !cta2 || crank == 0;          // in a pair only the leader's issuer warps have work
    if (issues && (p.b_resident || kRes == 1)) mbar_wait(bres_bar, 0, p.watchdog, 250);
    const uint32_t ident_lo = desc_lo(smem_u32(smem_i));
    // the pair's MMAs and commits (per-tap / linear path only; the halo paths never run as pairs)
    auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
      if constexpr (cta2) tc_mma_f16_2cta(d, a, b, id, acc);
      else tc_mma_f16(d, a, b, id, acc);
    };
    auto commit = [&](uint32_t bar) {
      if constexpr (cta2) tc_commit_2cta(bar, 3);
      else tc_commit_a(bar);
    };
    for (int tile = !issues ? p.total_tiles : (dual ? tile_unit + iw * grid_units : (iw == solo ? tile_unit : p.total_tiles));
         tile < p.total_tiles; tile += tstep * grid_units, tile_iter += tstep) {
      const int acc_iter = p.mpair ? 2 * tile_iter : tile_iter;          // a paired tile fills two accumulator buffers
      const int buf = acc_iter & (nacc - 1);
      const uint32_t tphase = static_cast<uint32_t>(acc_iter >> nacc_shift) & 1u;
      const int n_tile = tile - fastdiv(tile, p.fd_n) * p.n_tiles;
      if (kStall && lane == 0 && p.watchdog) p.watchdog[16 + 8 * blockIdx.x + iw] = tile_iter * 10 + 1;
      sc.start();
      mbar_wait_a(tempty0 + buf * 8, tphase ^ 1, p.watchdog, 200 + buf + 1000 * iw);
      if (p.mpair) mbar_wait_a(tempty0 + (buf + 1) * 8, tphase ^ 1, p.watchdog, 201 + buf + 1000 * iw);
      sc.stop(st_tempty);
      tc_fence_after();
      if (kStall && lane == 0 && p.watchdog) p.watchdog[16 + 8 * blockIdx.x + iw] = tile_iter * 10 + 2;
      auto my_turn = [&]() {                 // dual issue: the other warp has issued the tile before this one
        if (dual && tile_iter > 0) mbar_wait_a(turn_other, static_cast<uint32_t>((tile_iter - 1) >> 1) & 1u, p.watchdog, 260 + 1000 * iw);
      };
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * p.BN);
      const uint32_t b_res = b_lo0 + static_cast<uint32_t>((n_tile - res_nt0) * p.num_kb) * b_step;   // resident weights of this N tile
      auto mma4 = [&](uint32_t a, uint32_t b, uint32_t acc0) {          // one 64-channel tap out of a halo tile
        mma(d_tmem, pack_desc(a, hi_halo), pack_desc(b, hi1024), idesc, acc0);
        mma(d_tmem, pack_desc(a + 2, hi_halo), pack_desc(b + 2, hi1024), idesc, 1u);
        mma(d_tmem, pack_desc(a + 4, hi_halo), pack_desc(b + 4, hi1024), idesc, 1u);
        mma(d_tmem, pack_desc(a + 6, hi_halo), pack_desc(b + 6, hi1024), idesc, 1u);
        if (p.mpair) {                                                   // the patch 8 pixels (8 x 128 B) to the right
          const uint32_t d2 = d_tmem + static_cast<uint32_t>(p.BN);
          mma(d2, pack_desc(a + 64, hi_halo), pack_desc(b, hi1024), idesc, acc0);
          mma(d2, pack_desc(a + 66, hi_halo), pack_desc(b + 2, hi1024), idesc, 1u);
          mma(d2, pack_desc(a + 68, hi_halo), pack_desc(b + 4, hi1024), idesc, 1u);
          mma(d2, pack_desc(a + 70, hi_halo), pack_desc(b + 6, hi1024), idesc, 1u);
        }
      };
      auto commit_tile = [&]() {                                         // accumulator(s) complete -> epilogue
        commit(tfull0 + buf * 8);
        if (p.mpair) commit(tfull0 + (buf + 1) * 8);
      };
      const uint32_t row_a = static_cast<uint32_t>(p.halo_w) * 8u;             // next halo row (descriptor units of 16 B)
      const bool halo_done = p.halo != 0;
      {
      if (p.halo && p.b_resident) {
        // Resident weights: nothing but the halo tile has to arrive, so ALL taps of a channel block go out under one
        // barrier wait, back to back (an mbarrier wait costs the pipe ~140 cycles, tools/umma_issue_probe.cu; the
        // per-tap path below pays elect + __syncwarp + bookkeeping between every four MMAs).
        const uint32_t tap_b = static_cast<uint32_t>(p.cin_blocks) * b_step;   // next tap of the same channel block
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          sc.start();
          mbar_wait_a(afull0 + astage * 8, aphase, p.watchdog, 220 + astage + 1000 * iw);
          sc.stop(st_full);
          if (cb == 0) my_turn();
          tc_fence_after();
          const uint32_t a0 = a_lo0 + static_cast<uint32_t>(astage) * halo_step;          // pixel (0,0) of the halo tile
          const uint32_t b0 = b_res + static_cast<uint32_t>(cb) * b_step;                  // block (tap 0, cb)
          const uint32_t acc = cb != 0 ? 1u : 0u;
          if (elect_one()) {
            if (p.kh == 3 && p.kw == 3) {
              mma4(a0, b0, acc);
              mma4(a0 + 8, b0 + tap_b, 1u);
              mma4(a0 + 16, b0 + 2 * tap_b, 1u);
              mma4(a0 + row_a, b0 + 3 * tap_b, 1u);
              mma4(a0 + row_a + 8, b0 + 4 * tap_b, 1u);
              mma4(a0 + row_a + 16, b0 + 5 * tap_b, 1u);
              mma4(a0 + 2 * row_a, b0 + 6 * tap_b, 1u);
              mma4(a0 + 2 * row_a + 8, b0 + 7 * tap_b, 1u);
              mma4(a0 + 2 * row_a + 16, b0 + 8 * tap_b, 1u);
            } else {
              uint32_t a_row = a0, b_tap = b0, acc_t = acc;
              for (int tr = 0; tr < p.kh; ++tr, a_row += row_a) {
                uint32_t a_tap = a_row;
                for (int ts = 0; ts < p.kw; ++ts, a_tap += 8, b_tap += tap_b) {
                  mma4(a_tap, b_tap, acc_t);
                  acc_t = 1u;
                }
              }
            }
            commit(aempty0 + astage * 8);
            if (kRes != 1 && cb == p.cin_blocks - 1) commit_tile();
          }
          __syncwarp();
          if (++astage == p.a_stages) { astage = 0; aphase ^= 1; }
        }
      } else if (p.halo && p.brow > 1) {
        // Streamed weights, one filter ROW (kw taps) per B stage: a wait per row instead of a wait per tap
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          sc.start();
          mbar_wait_a(afull0 + astage * 8, aphase, p.watchdog, 220 + astage + 1000 * iw);
          sc.stop(st_full);
          uint32_t a_row = a_lo0 + static_cast<uint32_t>(astage) * halo_step;
          uint32_t acc = cb != 0 ? 1u : 0u;
          for (int tr = 0; tr < p.kh; ++tr) {
            sc.start();
            mbar_wait_a(full0 + stage * 8, phase, p.watchdog, 210 + stage + 1000 * iw);
            sc.stop(st_full);
            if (cb == 0 && tr == 0) my_turn();
            tc_fence_after();
            if (elect_one()) {
              if (p.kw == 3) {
                mma4(a_row, b_ring, acc);
                mma4(a_row + 8, b_ring + b_step, 1u);
                mma4(a_row + 16, b_ring + 2 * b_step, 1u);
              } else {
                uint32_t acc_t = acc;
                for (int ts = 0; ts < p.kw; ++ts) {
                  mma4(a_row + static_cast<uint32_t>(ts) * 8u, b_ring + static_cast<uint32_t>(ts) * b_step, acc_t);
                  acc_t = 1u;
                }
              }
              commit(empty0 + stage * 8);
              if (tr == p.kh - 1) {
                commit(aempty0 + astage * 8);
                if (kRes != 1 && cb == p.cin_blocks - 1) commit_tile();
              }
            }
            __syncwarp();
            acc = 1u;
            a_row += row_a;
            b_ring += static_cast<uint32_t>(p.brow) * b_step;
            if (++stage == p.stages) { stage = 0; phase ^= 1; b_ring = b_lo0; }
          }
          if (++astage == p.a_stages) { astage = 0; aphase ^= 1; }
        }
      } else if (p.halo) {
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          sc.start();
          mbar_wait_a(afull0 + astage * 8, aphase, p.watchdog, 220 + astage + 1000 * iw);
          sc.stop(st_full);
          tc_fence_after();
          uint32_t a_tap = a_lo0 + static_cast<uint32_t>(astage) * halo_step;          // pixel (0,0) of the halo tile
          uint32_t b_tap = b_res + static_cast<uint32_t>(cb) * b_step;                   // resident: block (tap 0, cb)
          uint32_t acc = cb != 0 ? 1u : 0u;
          for (int tr = 0; tr < p.kh; ++tr) {
            for (int ts = 0; ts < p.kw; ++ts) {
              if (!p.b_resident) {
                sc.start();
                mbar_wait_a(full0 + stage * 8, phase, p.watchdog, 210 + stage + 1000 * iw);
                sc.stop(st_full);
                tc_fence_after();
                b_tap = b_ring;
              }
              if (cb == 0 && tr == 0 && ts == 0) {
                my_turn();
                tc_fence_after();
              }
              const bool last_tap = tr == p.kh - 1 && ts == p.kw - 1;
              if (elect_one()) {
                mma4(a_tap, b_tap, acc);
                if (!p.b_resident) commit(empty0 + stage * 8);
                if (last_tap) {
                  commit(aempty0 + astage * 8);
                  if (kRes != 1 && cb == p.cin_blocks - 1) commit_tile();
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
      }
      }
      if (!halo_done) {
        // One walk over the tile's ring stages: n_k weight stages (one or two k-blocks each), then for kRes = 1 the
        // residual slabs (accumulator columns [rs * 64, rs * 64 + ncols) += residual slab x identity).
        // (Splitting the stages of a tile between the two issuers was tried: 2-4 % on the long tiles, but the order in
        // which the k-blocks reach the accumulator then depends on timing and results stop being reproducible.)
        const int n_k = (p.num_kb + p.kbps - 1) / p.kbps;
        uint32_t b_cur = b_res;
        int kb0 = 0;
        for (int j = 0; j < tile_stages; ++j) {
          {
            if (kStall && lane == 0 && p.watchdog) p.watchdog[16 + 8 * blockIdx.x + iw] = tile_iter * 10 + 3 + 10000 * stage + 100000 * phase;
            sc.start();
            mbar_wait_a(full0 + stage * 8, phase, p.watchdog, 210 + stage + 1000 * iw);
            sc.stop(st_full);
            if (j == 0) my_turn();
            tc_fence_after();
            const bool last_mine = j + 1 >= tile_stages;
            if (j < n_k) {
              const bool two = p.kbps == 2 && kb0 + 1 < p.num_kb;          // this stage carries two k-blocks
              if (!p.b_resident) b_cur = b_ring;
              if (elect_one()) {
                mma(d_tmem, pack_desc(a_ring, hi1024), pack_desc(b_cur, hi1024), idesc, j != 0 ? 1u : 0u);
                mma(d_tmem, pack_desc(a_ring + 2, hi1024), pack_desc(b_cur + 2, hi1024), idesc, 1u);
                mma(d_tmem, pack_desc(a_ring + 4, hi1024), pack_desc(b_cur + 4, hi1024), idesc, 1u);
                mma(d_tmem, pack_desc(a_ring + 6, hi1024), pack_desc(b_cur + 6, hi1024), idesc, 1u);
                if (two) {
                  const uint32_t a2 = a_ring + a_step, b2 = b_cur + b_step;
                  mma(d_tmem, pack_desc(a2, hi1024), pack_desc(b2, hi1024), idesc, 1u);
                  mma(d_tmem, pack_desc(a2 + 2, hi1024), pack_desc(b2 + 2, hi1024), idesc, 1u);
                  mma(d_tmem, pack_desc(a2 + 4, hi1024), pack_desc(b2 + 4, hi1024), idesc, 1u);
                  mma(d_tmem, pack_desc(a2 + 6, hi1024), pack_desc(b2 + 6, hi1024), idesc, 1u);
                }
                commit(empty0 + stage * 8);
                if (last_mine) commit(tfull0 + buf * 8);
              }
            } else {
              const int rs = j - n_k;
              const uint32_t idesc_r = make_idesc_f16(cta2 ? 2 * kBM : kBM, min(kSlab, p.BN - rs * kSlab), kBF16 ? 1 : 0);
              const uint32_t d_r = d_tmem + static_cast<uint32_t>(rs * kSlab);
              if (elect_one()) {
                mma(d_r, pack_desc(a_ring, hi1024), pack_desc(ident_lo, hi1024), idesc_r, 1u);
                mma(d_r, pack_desc(a_ring + 2, hi1024), pack_desc(ident_lo + 2, hi1024), idesc_r, 1u);
                mma(d_r, pack_desc(a_ring + 4, hi1024), pack_desc(ident_lo + 4, hi1024), idesc_r, 1u);
                mma(d_r, pack_desc(a_ring + 6, hi1024), pack_desc(ident_lo + 6, hi1024), idesc_r, 1u);
                commit(empty0 + stage * 8);
                if (last_mine) commit(tfull0 + buf * 8);
              }
            }
            __syncwarp();
          }
          if (j < n_k) {
            kb0 += p.kbps;
            b_cur += static_cast<uint32_t>(p.kbps) * b_step;
          }
          a_ring += static_cast<uint32_t>(p.kbps) * a_step;
          b_ring += static_cast<uint32_t>(p.kbps) * b_step;
          if (++stage == p.stages) { stage = 0; phase ^= 1; a_ring = a_lo0; b_ring = b_lo0; }
        }
      }
      if (dual) {
        if (elect_one()) mbar_arrive(&turn_bar[iw]);     // everything of this tile has been issued: the other warp's turn
        __syncwarp();
        skip_tile();
      }
    }
    if (kStall && lane == 0 && p.watchdog) p.watchdog[16 + 8 * blockIdx.x + iw] = 99999;
    if (kStall && sc.on && lane == 0) {
      long long* o = p.stall + 1 + blockIdx.x * 16;
      o[6] = clock64() - st_begin; o[7] = st_tempty; o[8] = st_full;
    }
  } else {
    // ============================================================ TMA producer
    const uint32_t a_s0 = smem_u32(smem_a), b_s0 = smem_u32(smem_b);
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), afull0 = smem_u32(afull_bar),
                   aempty0 = smem_u32(aempty_bar);
    int stage = 0, astage = 0;
    uint32_t phase = 0, aphase = 0;
    uint32_t a_dst = a_s0, b_dst = b_s0;
    int tile_iter = 0;
    StallClock<kStall> sc{kStall && p.stall != nullptr, 0};
    long long st_empty = 0, st_rempty = 0;
    const long long st_begin = kStall && sc.on ? clock64() : 0;
    // pair: completion bytes of both CTAs' loads are signalled on the LEADER's barriers (its issuer waits on them)
    // pair: completion bytes of BOTH CTAs' loads are signalled on the LEADER's barriers (its issuer waits on them); the
    // leader announces the sum.  `lead` turns a local barrier address into the leader's (same offset in its window).
    uint32_t lead = 0u;
    if constexpr (cta2) lead = mapa_shared(full0, 0) - full0;
    const uint32_t full_dst0 = full0 + lead;
    auto expect = [&](uint32_t bar, uint32_t bytes) {
      if constexpr (cta2) {
        if (crank == 0) mbar_expect_tx_a(bar, 2u * bytes);
      } else {
        mbar_expect_tx_a(bar, bytes);
      }
    };
    auto ld2 = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int x, int y) {
      if constexpr (cta2) tma_load_2d_2cta(dst, m, bar + lead, x, y);
      else tma_load_2d_a(dst, m, bar, x, y);
    };
    auto ld4 = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1_, int c2_, int c3_) {
      if constexpr (cta2) tma_load_4d_2cta(dst, m, bar + lead, c0, c1_, c2_, c3_);
      else tma_load_4d_a(dst, m, bar, c0, c1_, c2_, c3_);
    };
    if ((p.b_resident || kRes == 1) && lane == 0) {
      // a pair holds half of every resident weight block (b_rows rows) and half of the identity (32 rows) per CTA
      const uint32_t bres = smem_u32(bres_bar);
      expect(bres, static_cast<uint32_t>((p.b_resident ? res_nts * p.num_kb * b_stage_bytes : 0) +
                                          (kRes == 1 ? (cta2 ? kIdentBytes / 2 : kIdentBytes) : 0)));
      if (p.b_resident)
        for (int nt = 0; nt < res_nts; ++nt)
          for (int kb = 0; kb < p.num_kb; ++kb)
            ld2(smem_u32(smem_b) + static_cast<uint32_t>((nt * p.num_kb + kb) * b_stage_bytes), &tmB, bres, kb * kBK,
                (res_nt0 + nt) * p.BN + static_cast<int>(crank) * b_rows);
      if (kRes == 1) ld2(smem_u32(smem_i), &tmI, bres, 0, static_cast<int>(crank) * 32);
    }
    griddep_wait();                      // weights and identity are constants; activations are the previous kernel's
    const uint32_t tx_bytes = static_cast<uint32_t>((p.b_resident ? 0 : b_stage_bytes) + p.a_bytes);
    const uint32_t halo_tx = static_cast<uint32_t>(p.halo_w * p.halo_h * 128);
    for (int tile = tile_unit; tile < p.total_tiles; tile += grid_units, ++tile_iter) {
      int n_tile, c1, c2, c3;
      decode(tile, n_tile, c1, c2, c3);
      const int n_coord = n_tile * p.BN + static_cast<int>(crank) * b_rows;      // pair: this CTA's half of the weight rows
      if (kStall && lane == 0 && p.watchdog) p.watchdog[16 + 8 * blockIdx.x + 2] = tile_iter;
      if (p.halo) {
        const int taps = p.kh * p.kw;
        for (int hb = 0; hb < p.cin_blocks; ++hb) {
          sc.start();
          mbar_wait_a(aempty0 + astage * 8, aphase ^ 1, p.watchdog, 320 + astage);
          sc.stop(st_empty);
          if (elect_one()) {
            expect(afull0 + astage * 8, halo_tx);
            ld4(a_s0 + static_cast<uint32_t>(astage * p.halo_bytes), &tmA, afull0 + astage * 8, hb * kBK, c1 - p.pw, c2 - p.ph, c3);
          }
          __syncwarp();
          if (++astage == p.a_stages) { astage = 0; aphase ^= 1; }
          if (!p.b_resident && p.brow > 1) {
            int kcoord = hb * kBK;                                        // K coordinate of (tap 0, hb)
            for (int tr = 0; tr < p.kh; ++tr) {
              sc.start();
              mbar_wait_a(empty0 + stage * 8, phase ^ 1, p.watchdog, 300 + stage);
              sc.stop(st_empty);
              if (elect_one()) {
                expect(full0 + stage * 8, static_cast<uint32_t>(p.brow * b_stage_bytes));
                for (int ts = 0; ts < p.brow; ++ts)
                  ld2(b_dst + static_cast<uint32_t>(ts * b_stage_bytes), &tmB, full0 + stage * 8, kcoord + ts * p.cin_blocks * kBK, n_coord);
              }
              __syncwarp();
              kcoord += p.brow * p.cin_blocks * kBK;
              b_dst += static_cast<uint32_t>(p.brow * b_stage_bytes);
              if (++stage == p.stages) { stage = 0; phase ^= 1; b_dst = b_s0; }
            }
          } else if (!p.b_resident) {
            int kcoord = hb * kBK;                                        // K coordinate of (tap 0, hb)
            for (int tap = 0; tap < taps; ++tap) {
              sc.start();
              mbar_wait_a(empty0 + stage * 8, phase ^ 1, p.watchdog, 300 + stage);
              sc.stop(st_empty);
              if (elect_one()) {
                expect(full0 + stage * 8, static_cast<uint32_t>(b_stage_bytes));
                ld2(b_dst, &tmB, full0 + stage * 8, kcoord, n_coord);
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
          if (kStall && p.skip_loads) {
            if (elect_one()) mbar_arrive(&full_bar[stage]);
          } else if (elect_one()) {
            if (!cta2) mbar_expect_tx_a(full0 + stage * 8, tx_bytes * static_cast<uint32_t>(nk));
            else if (crank == 0) mbar_expect_tx_a(full0 + stage * 8, 2u * tx_bytes * static_cast<uint32_t>(nk));
          }
          for (int j = 0; j < nk; ++j) {
            if (!(kStall && p.skip_loads) && elect_one()) {
              const uint32_t fb = full_dst0 + static_cast<uint32_t>(stage) * 8u;
              if constexpr (cta2) {
                tma_load_2d_2cta(b_dst + static_cast<uint32_t>(j * b_stage_bytes), &tmB, fb, kcoord, n_coord);
                if (p.patch) tma_load_4d_2cta(a_dst + static_cast<uint32_t>(j * kAStageBytes), &tmA, fb, cb * kBK, aw + tap_s, ah + tap_r, c3);
                else tma_load_2d_2cta(a_dst + static_cast<uint32_t>(j * kAStageBytes), &tmA, fb, kcoord, c1);
              } else {
                if (!p.b_resident) tma_load_2d_a(b_dst + static_cast<uint32_t>(j * b_stage_bytes), &tmB, fb, kcoord, n_coord);
                if (p.patch) tma_load_4d_a(a_dst + static_cast<uint32_t>(j * kAStageBytes), &tmA, fb, cb * kBK, aw + tap_s, ah + tap_r, c3);
                else tma_load_2d_a(a_dst + static_cast<uint32_t>(j * kAStageBytes), &tmA, fb, kcoord, c1);
              }
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
      if constexpr (kRes == 1) {
        // the tile's residual slabs ride the same ring as extra A blocks (host: one k-block per stage, no halo)
        for (int rs = 0; rs < nslab; ++rs) {
          sc.start();
          mbar_wait_a(empty0 + stage * 8, phase ^ 1, p.watchdog, 300 + stage);
          sc.stop(st_empty);
          if (elect_one()) {
            // (the residual's channel coordinate is the FULL N tile's, not this CTA's half of the weight rows)
            const int r_coord = n_tile * p.BN + rs * kSlab;
            const uint32_t fb = full_dst0 + static_cast<uint32_t>(stage) * 8u;
            if constexpr (cta2) {
              if (crank == 0) mbar_expect_tx_a(full0 + stage * 8, 2u * static_cast<uint32_t>(p.a_bytes));
              if (p.patch) tma_load_4d_2cta(a_dst, &tmR, fb, r_coord, c1, c2, c3);
              else tma_load_2d_2cta(a_dst, &tmR, fb, r_coord, c1);
            } else {
              mbar_expect_tx_a(full0 + stage * 8, static_cast<uint32_t>(p.a_bytes));
              if (p.patch) tma_load_4d_a(a_dst, &tmR, fb, r_coord, c1, c2, c3);
              else tma_load_2d_a(a_dst, &tmR, fb, r_coord, c1);
            }
          }
          __syncwarp();
          a_dst += kAStageBytes;
          b_dst += static_cast<uint32_t>(b_stage_bytes);
          if (++stage == p.stages) { stage = 0; phase ^= 1; a_dst = a_s0; b_dst = b_s0; }
        }
      }
    }
    if (kStall && lane == 0 && p.watchdog) p.watchdog[16 + 8 * blockIdx.x + 2] = 99999;
    if (kStall && sc.on && lane == 0) {
      long long* o = p.stall + 1 + blockIdx.x * 16;
      o[9] = clock64() - st_begin; o[10] = st_empty; o[11] = st_rempty;
    }
  }

  // -------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if constexpr (cta2) cluster_sync_all();   // the peer may still read this CTA's weight halves or arrive on its barriers
  if (warp == kMmaWarp) {
    tc_fence_after();
    if constexpr (cta2) tmem_dealloc_2cta(tmem_base, tmem_cols);
    else tmem_dealloc_dyn(tmem_base, tmem_cols);
  }
}

}  // namespace frb
