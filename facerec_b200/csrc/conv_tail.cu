// K2t — the tail of a ResNet bottleneck fused with the head of the next one (layer1 of RetinaFace-R50; layer2: K2t2 below):
//
//   y  = relu(conv3_1x1(a) + skip)          [M, 256]   (what conv_tma_kernel<res = 1, act = 1> writes)
//   y1 = relu(conv1_1x1_next(y))            [M, N2]    (what the next block's conv_tma_kernel<0, 1> writes)
//
// Both layers are HBM-bound at 1080p (profiles/r02_layer_times_r50_ab_pairs_res256.txt: conv3 + skip 1.59 ms at 6.0 TB/s,
// the next conv1 0.77 ms at 6.9 TB/s of which 4.25 GB is re-reading y), so the only way to make them faster is not to read
// y again: the staged output slabs of y — which the epilogue packs into the 128-byte-swizzled K-major layout for their TMA
// store anyway — are at the same time the A operand of the second GEMM.  Per 128-pixel tile:
//
//   producer    A block(s) of a and the four 64-channel skip slabs -> stage ring (16 KB boxes)
//   G1 issuer   acc1[buf] (256 TMEM columns)  = a x W3   (N = 256 MMAs), then += skip slab x I (N = 64 MMAs on the slab's
//               columns: the skip is added by the tensor core exactly as in conv_tma_kernel, same order, same bits); with
//               FRB_OPT_FUSE_TAIL = 3 and a 1x1 skip conv (layer1.0) instead += x x Wd: the skip conv as one more k-block
//   epilogue    e1: per 64-column slab: TMEM -> +bias -> 16-bit with ReLU -> slab slot (swizzled), "slab written" to the M2
//               issuer, "slab staged" to the store warp
//   M2 issuer   acc2 (columns [0, N2) of the SAME buffer, drained by then) (+)= slab x W1n block, slabs in channel
//               order (the accumulation order of the unfused conv1: its output bits do not change either)
//   epilogue    e2: TMEM -> +bias2 -> 16-bit with ReLU -> slab slot; buffer back to the G1 issuer
//   store warp  TMA-stores the staged slabs (e1 into y, e2 into y1, in item order) and hands a slot back once its store has
//               drained, one store behind
//
// Weights of both layers and the identity (or the skip conv's weights) stay resident in shared memory.  Epilogue items (4 e1
// slabs + N2/64 e2 slabs per tile) are dealt round-robin to three independent 4-warp groups that take the slab slots in item
// order; a slot is reused when its TMA store has finished reading it and (e1) the M2 MMAs on it have retired.  G1 runs
// ahead of M2 by up to a tile (two TMEM buffers), so the epilogue of tile t overlaps G1 of tile t + 1; what orders the two
// issuers' MMAs on a buffer are barriers, and every wait depends only on earlier items, so the pipeline cannot deadlock
// (tools/conv_tail_protocol_sim.py replays the barrier protocol under random interleavings and checks every parity wait).
// Linear M tiling (1x1, stride 1, no padding); ragged last tile: loads zero-fill, stores clip.
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "conv_igemm.cuh"
#include "host.h"

namespace frb {

constexpr int kTailEpiWarps = 12;             // (16 = four groups measured no faster: profiles/r02_layer_times_r50_tail2_ab.txt)
constexpr int kTailGroups = kTailEpiWarps / 4;
constexpr int kTailThreads = (kTailEpiWarps + 4) * 32;     // + G1 issuer + TMA producer + store warp + M2 issuer
constexpr int kTailN1 = 256;                               // conv3 output channels = one MMA N
constexpr int kTailSlab = 64;
constexpr int kTailSlabBytes = kBM * kTailSlab * 2;        // 16 KB: one A box, one skip slab, one output slab
constexpr int kTailMaxRing = 8, kTailMaxSlots = 8;
constexpr int kTailIdentBytes = 64 * 128;

struct ConvTailArgs {
  int M, m_tiles;
  int kb1;               // 64-channel blocks of conv3's input (1 or 2)
  int N2;                // output channels of the next conv1: 64 or 128
  int ring, slots;
  int kbx;               // 0: the skip is a tensor (four 64-channel slabs per tile, added through the identity); > 0: the skip is
                         // a 1x1 conv of x computed right here, kbx 64-channel blocks of x per tile (tmR = x, tmI = its weights)
  const float* bias1;    // [256] (skip conv: conv3's bias + the skip conv's)
  const float* bias2;    // [N2]
  int* watchdog;
};

template <bool kBF16>
__global__ void __launch_bounds__(kTailThreads, 1)
conv_tail_kernel(const __grid_constant__ ConvTailArgs p, const __grid_constant__ CUtensorMap tmA,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO1,
                 const __grid_constant__ CUtensorMap tmO2, const __grid_constant__ CUtensorMap tmW3,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmI) {
  using E = Elem16<kBF16>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  const int w1_block = p.N2 * 128;                                   // one 64-wide k-block of the next conv1's weights
  uint8_t* smem_w3 = smem;                                           // [kb1][256 x 64]
  uint8_t* smem_w1 = smem_w3 + p.kb1 * (kTailN1 * 128);              // [4][N2 x 64]
  uint8_t* smem_i = smem_w1 + 4 * w1_block;                          // 64 x 64 identity, or [kbx][256 x 64] skip-conv weights
  uint8_t* smem_ring = smem_i + (p.kbx ? p.kbx * (kTailN1 * 128) : kTailIdentBytes);
  uint8_t* smem_slot = smem_ring + p.ring * kTailSlabBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_slot + p.slots * kTailSlabBytes);
  uint64_t* full_bar = bars;                          // [8] ring slot filled by TMA
  uint64_t* empty_bar = bars + 8;                     // [8] ring slot consumed by the MMAs
  uint64_t* slabw_bar = bars + 16;                    // [8] slab slot written by its epilogue group (e1)
  uint64_t* sfree_bar = bars + 24;                    // [8] slab slot free again: 2 arrivals (store drained, M2 retired)
  uint64_t* acc1_bar = bars + 32;                     // [2] G1 + skip complete
  uint64_t* acc2_bar = bars + 34;                     // [2] M2 complete
  uint64_t* afree_bar = bars + 36;                    // [2] e2 has drained the buffer
  uint64_t* bres_bar = bars + 38;                     // resident operands have landed
  uint64_t* turn_bar = bars + 39;                     // [kTailGroups <= 4] group g has taken the slab slot of its current item
  uint64_t* staged_bar = bars + 43;                   // [8] slab slot staged (e1 and e2): the store warp may send it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 51);
  float* s_bias1 = reinterpret_cast<float*>(bars + 52);   // [256]
  float* s_bias2 = s_bias1 + kTailN1;                     // [128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kMmaWarp = kTailEpiWarps, kTmaWarp = kTailEpiWarps + 1, kStoreWarp = kTailEpiWarps + 2, kM2Warp = kTailEpiWarps + 3;
  const int ns2 = p.N2 / kTailSlab;
  const int ipt = 4 + ns2;                                           // epilogue items per tile
  const int my_tiles = (p.m_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmR); tma_prefetch_desc(&tmO1); tma_prefetch_desc(&tmO2);
    tma_prefetch_desc(&tmW3); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmI);
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < 8; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
        mbar_init(&slabw_bar[s], 1);
        mbar_init(&sfree_bar[s], 2);
        mbar_init(&staged_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&acc1_bar[b], 1);
        mbar_init(&acc2_bar[b], 1);
        mbar_init(&afree_bar[b], 4 * ns2);
      }
      mbar_init(bres_bar, 1);
      for (int g = 0; g < kTailGroups; ++g) mbar_init(&turn_bar[g], 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_dyn(tmem_slot, 512);
    tmem_relinquish();
  }
  if (threadIdx.x < kTailN1) s_bias1[threadIdx.x] = __ldg(p.bias1 + threadIdx.x);
  else if (threadIdx.x < kTailN1 + p.N2) s_bias2[threadIdx.x - kTailN1] = __ldg(p.bias2 + threadIdx.x - kTailN1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kTailEpiWarps) {
    // ================================================================ epilogue groups
    const int quarter = warp & 3;
    const int grp = warp >> 2;
    const int gtid = threadIdx.x & 127;
    const bool leader = gtid == 0;
    const int row = quarter * 32 + lane;
    const uint32_t sw7 = static_cast<uint32_t>(row & 7);
    const int my_items = my_tiles * ipt;
    int t = grp / ipt, j = grp - t * ipt;                            // tile iteration, item inside the tile
    int slot = grp % p.slots, use = grp / p.slots;                   // slab slot of the item, how often it has been used
    for (int q = grp; q < my_items; q += kTailGroups) {
      const int buf = t & 1;
      const uint32_t aph = static_cast<uint32_t>(t >> 1) & 1u;
      const bool second = j >= 4;                                    // e2 item
      const int s = second ? j - 4 : j;
      if (leader) {
        // Slots are taken in item order (item q after item q - 1, whichever groups they belong to): barrier phases are told
        // apart by parity only, and a group that ran two uses of a slot ahead of another group would take the slot's
        // phase before last for its own (tools/conv_tail_protocol_sim.py finds that interleaving when groups take slots
        // independently).  One thread waits for the group: the turn barrier below may be a phase further on by the
        // time a slower thread would look at it.
        if (q > 0) mbar_wait(&turn_bar[(grp + kTailGroups - 1) % kTailGroups], static_cast<uint32_t>((q - 1) / kTailGroups) & 1u, p.watchdog, 130 + grp);
        mbar_wait(&sfree_bar[slot], (static_cast<uint32_t>(use) & 1u) ^ 1u, p.watchdog, 100 + slot);
        mbar_arrive(&turn_bar[grp]);
      }
      named_bar_sync(1 + grp, 128);
      mbar_wait(second ? &acc2_bar[buf] : &acc1_bar[buf], aph, p.watchdog, second ? 120 + buf : 110 + buf);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * kTailN1 + s * kTailSlab);
      const float* bias = second ? s_bias2 + s * kTailSlab : s_bias1 + s * kTailSlab;
      uint8_t* const orow = smem_slot + slot * kTailSlabBytes + static_cast<uint32_t>(row) * 128u;
      uint32_t v0[16], v1[16];
      auto finish16 = [&](const uint32_t* v, int cbase) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int cc = cbase + c * 8;
          const float4 b0 = *reinterpret_cast<const float4*>(bias + cc);
          const float4 b1 = *reinterpret_cast<const float4*>(bias + cc + 4);
          uint4 o;
          o.x = E::from_f2_relu(__uint_as_float(v[c * 8 + 0]) + b0.x, __uint_as_float(v[c * 8 + 1]) + b0.y);
          o.y = E::from_f2_relu(__uint_as_float(v[c * 8 + 2]) + b0.z, __uint_as_float(v[c * 8 + 3]) + b0.w);
          o.z = E::from_f2_relu(__uint_as_float(v[c * 8 + 4]) + b1.x, __uint_as_float(v[c * 8 + 5]) + b1.y);
          o.w = E::from_f2_relu(__uint_as_float(v[c * 8 + 6]) + b1.z, __uint_as_float(v[c * 8 + 7]) + b1.w);
          *reinterpret_cast<uint4*>(orow + ((static_cast<uint32_t>(cc >> 3) ^ sw7) << 4)) = o;
        }
      };
      tmem_ld16(taddr, v0);
      tc_wait_ld16(v0);
      tmem_ld16(taddr + 16, v1);
      finish16(v0, 0);
      tc_wait_ld16(v1);
      tmem_ld16(taddr + 32, v0);
      finish16(v1, 16);
      tc_wait_ld16(v0);
      tmem_ld16(taddr + 48, v1);
      finish16(v0, 32);
      tc_wait_ld16(v1);
      tc_fence_before();                       // the accumulator columns of this warp are in registers
      if (second) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&afree_bar[buf]);
      }
      finish16(v1, 48);
      fence_proxy_async_smem();                // the slab is read by TMA (store) and by the tensor core (M2)
      named_bar_sync(1 + grp, 128);
      if (leader) {
        if (!second) mbar_arrive(&slabw_bar[slot]);          // -> M2
        mbar_arrive(&staged_bar[slot]);                      // -> the store warp
      }
      j += kTailGroups;
      while (j >= ipt) { j -= ipt; ++t; }
      slot += kTailGroups;
      while (slot >= p.slots) { slot -= p.slots; ++use; }
    }
  } else if (warp == kStoreWarp) {
    // ================================================================ store warp
    // One thread sends every staged slab (items in order) and gives the slot back once the store has finished READING it, one
    // store behind: the epilogue groups never wait for a store to drain (as leaders of their own stores they spent a
    // store-read latency per item: 1.57 -> ... ms per launch, profiles/r02_layer_times_r50_fuse_ab.txt).
    if (lane == 0) {
      const int my_items = my_tiles * ipt;
      int t = 0, j = 0, slot = 0;
      uint32_t par = 0;                        // per slot: parity of its uses so far
      int pend_slot = -1;                      // the store before the last one: slot | 16 for an e2 slot
      int last_slot = -1;
      auto give_back = [&](int ps) {
        // a slot's two "free" arrivals: store drained, M2 retired (an e2 slot is read by no MMA: both from here)
        mbar_arrive(&sfree_bar[ps & 15]);
        if (ps & 16) mbar_arrive(&sfree_bar[ps & 15]);
      };
      for (int q = 0; q < my_items; ++q) {
        const bool second = j >= 4;
        const int s = second ? j - 4 : j;
        const int m0 = (static_cast<int>(blockIdx.x) + t * static_cast<int>(gridDim.x)) * kBM;
        mbar_wait(&staged_bar[slot], (par >> slot) & 1u, p.watchdog, 400 + slot);
        par ^= 1u << slot;
        tma_store_2d(second ? &tmO2 : &tmO1, smem_slot + slot * kTailSlabBytes, s * kTailSlab, m0);
        tma_store_commit();
        pend_slot = last_slot;
        last_slot = slot | (second ? 16 : 0);
        if (pend_slot >= 0) {
          tma_store_wait_read<1>();            // everything but the store just committed has been read
          give_back(pend_slot);
        }
        if (++j == ipt) { j = 0; ++t; }
        if (++slot == p.slots) slot = 0;
      }
      if (last_slot >= 0) {
        tma_store_wait_read<0>();
        give_back(last_slot);
      }
      tma_store_wait<0>();                     // stores must land before the CTA (and its smem) goes away
    }
  } else if (warp == kMmaWarp) {
    // ================================================================ G1 issuer
    const uint32_t idesc1 = make_idesc_f16(kBM, kTailN1, kBF16 ? 1 : 0);
    const uint32_t idesc_r = make_idesc_f16(kBM, kTailSlab, kBF16 ? 1 : 0);
    const uint32_t hi = desc_hi(1024);
    const uint32_t ring_lo = desc_lo(smem_u32(smem_ring));
    const uint32_t w3_lo = desc_lo(smem_u32(smem_w3)), id_lo = desc_lo(smem_u32(smem_i));
    const uint32_t step = kTailSlabBytes >> 4;
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), acc10 = smem_u32(acc1_bar), afree0 = smem_u32(afree_bar);
    int stage = 0;
    uint32_t phase = 0;
    mbar_wait(bres_bar, 0, p.watchdog, 250);
    auto mma4 = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, uint32_t acc0) {
      tc_mma_f16(d, pack_desc(a, hi), pack_desc(b, hi), idesc, acc0);
      tc_mma_f16(d, pack_desc(a + 2, hi), pack_desc(b + 2, hi), idesc, 1u);
      tc_mma_f16(d, pack_desc(a + 4, hi), pack_desc(b + 4, hi), idesc, 1u);
      tc_mma_f16(d, pack_desc(a + 6, hi), pack_desc(b + 6, hi), idesc, 1u);
    };
    auto issue_g1 = [&](int t) {
      const int buf = t & 1;
      const uint32_t d = tmem_base + static_cast<uint32_t>(buf * kTailN1);
      mbar_wait_a(afree0 + buf * 8, (static_cast<uint32_t>(t >> 1) & 1u) ^ 1u, p.watchdog, 200 + buf);
      tc_fence_after();
      for (int kb = 0; kb < p.kb1; ++kb) {
        mbar_wait_a(full0 + stage * 8, phase, p.watchdog, 210 + stage);
        tc_fence_after();
        if (elect_one()) {
          mma4(d, ring_lo + static_cast<uint32_t>(stage) * step, w3_lo + static_cast<uint32_t>(kb) * ((kTailN1 * 128) >> 4), idesc1,
               kb != 0 ? 1u : 0u);
          tc_commit_a(empty0 + stage * 8);
        }
        __syncwarp();
        if (++stage == p.ring) { stage = 0; phase ^= 1; }
      }
      if (p.kbx) {
        // the skip is a 1x1 conv of x: more k-blocks into the same accumulator (fp32, not rounded to 16 bits in between)
        for (int kb = 0; kb < p.kbx; ++kb) {
          mbar_wait_a(full0 + stage * 8, phase, p.watchdog, 210 + stage);
          tc_fence_after();
          if (elect_one()) {
            mma4(d, ring_lo + static_cast<uint32_t>(stage) * step, id_lo + static_cast<uint32_t>(kb) * ((kTailN1 * 128) >> 4), idesc1, 1u);
            tc_commit_a(empty0 + stage * 8);
            if (kb == p.kbx - 1) tc_commit_a(acc10 + buf * 8);
          }
          __syncwarp();
          if (++stage == p.ring) { stage = 0; phase ^= 1; }
        }
        return;
      }
      for (int rs = 0; rs < 4; ++rs) {
        mbar_wait_a(full0 + stage * 8, phase, p.watchdog, 210 + stage);
        tc_fence_after();
        if (elect_one()) {
          mma4(d + static_cast<uint32_t>(rs * kTailSlab), ring_lo + static_cast<uint32_t>(stage) * step, id_lo, idesc_r, 1u);
          tc_commit_a(empty0 + stage * 8);
          if (rs == 3) tc_commit_a(acc10 + buf * 8);
        }
        __syncwarp();
        if (++stage == p.ring) { stage = 0; phase ^= 1; }
      }
    };
    for (int t = 0; t < my_tiles; ++t) issue_g1(t);
  } else if (warp == kM2Warp) {
    // ================================================================ M2 issuer (a warp of its own: a G1 that waits for a buffer
    // or for its loads must not hold up the M2 MMAs of slabs that are staged; barriers, not program order, order the two
    // issuers' MMAs on a TMEM buffer)
    const uint32_t idesc2 = make_idesc_f16(kBM, p.N2, kBF16 ? 1 : 0);
    const uint32_t hi = desc_hi(1024);
    const uint32_t slot_lo = desc_lo(smem_u32(smem_slot)), w1_lo = desc_lo(smem_u32(smem_w1));
    const uint32_t step = kTailSlabBytes >> 4;
    const uint32_t slabw0 = smem_u32(slabw_bar), sfree0 = smem_u32(sfree_bar), acc20 = smem_u32(acc2_bar);
    mbar_wait(bres_bar, 0, p.watchdog, 251);
    auto mma4 = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, uint32_t acc0) {
      tc_mma_f16(d, pack_desc(a, hi), pack_desc(b, hi), idesc, acc0);
      tc_mma_f16(d, pack_desc(a + 2, hi), pack_desc(b + 2, hi), idesc, 1u);
      tc_mma_f16(d, pack_desc(a + 4, hi), pack_desc(b + 4, hi), idesc, 1u);
      tc_mma_f16(d, pack_desc(a + 6, hi), pack_desc(b + 6, hi), idesc, 1u);
    };
    int slot = 0;                             // slab slot of the next e1 item M2 consumes
    uint32_t e1par = 0;                       // per slot: parity of the number of e1 slabs consumed from it ("written" phases
                                              // count e1 items only; e2 items share the slots but never arrive there)
    auto issue_m2 = [&](int t) {
      const int buf = t & 1;
      const uint32_t d = tmem_base + static_cast<uint32_t>(buf * kTailN1);
      for (int s = 0; s < 4; ++s) {
        mbar_wait_a(slabw0 + slot * 8, (e1par >> slot) & 1u, p.watchdog, 230 + slot);
        if (s == 0 && ns2 == 2) {
          // acc2 of a 128-wide conv1 covers the columns of slabs 0 AND 1: both must have left TMEM
          const int s1 = slot + 1 == p.slots ? 0 : slot + 1;
          mbar_wait_a(slabw0 + s1 * 8, (e1par >> s1) & 1u, p.watchdog, 240 + s1);
        }
        tc_fence_after();
        if (elect_one()) {
          mma4(d, slot_lo + static_cast<uint32_t>(slot) * step, w1_lo + static_cast<uint32_t>(s) * (static_cast<uint32_t>(w1_block) >> 4),
               idesc2, s != 0 ? 1u : 0u);
          tc_commit_a(sfree0 + slot * 8);
          if (s == 3) tc_commit_a(acc20 + buf * 8);
        }
        __syncwarp();
        e1par ^= 1u << slot;
        if (++slot == p.slots) slot = 0;
      }
      // the tile's e2 items take the next ns2 slots
      for (int e = 0; e < ns2; ++e)
        if (++slot == p.slots) slot = 0;
    };
    for (int t = 0; t < my_tiles; ++t) issue_m2(t);
  } else if (warp == kTmaWarp) {
    // ================================================================ TMA producer
    const uint32_t ring0 = smem_u32(smem_ring);
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    if (lane == 0) {
      const uint32_t bres = smem_u32(bres_bar);
      mbar_expect_tx_a(bres, static_cast<uint32_t>(p.kb1 * kTailN1 * 128 + 4 * w1_block + (p.kbx ? p.kbx * kTailN1 * 128 : kTailIdentBytes)));
      for (int kb = 0; kb < p.kb1; ++kb) tma_load_2d_a(smem_u32(smem_w3) + static_cast<uint32_t>(kb * kTailN1 * 128), &tmW3, bres, kb * kBK, 0);
      for (int s = 0; s < 4; ++s) tma_load_2d_a(smem_u32(smem_w1) + static_cast<uint32_t>(s * w1_block), &tmW1, bres, s * kBK, 0);
      if (p.kbx == 0) tma_load_2d_a(smem_u32(smem_i), &tmI, bres, 0, 0);
      for (int kb = 0; kb < p.kbx; ++kb) tma_load_2d_a(smem_u32(smem_i) + static_cast<uint32_t>(kb * kTailN1 * 128), &tmI, bres, kb * kBK, 0);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int m0 = (static_cast<int>(blockIdx.x) + t * static_cast<int>(gridDim.x)) * kBM;
      const int n_skip = p.kbx ? p.kbx : 4;                          // x blocks, or the four slabs of the skip tensor
      for (int i = 0; i < p.kb1 + n_skip; ++i) {
        mbar_wait_a(empty0 + stage * 8, phase ^ 1u, p.watchdog, 300 + stage);
        if (elect_one()) {
          mbar_expect_tx_a(full0 + stage * 8, kTailSlabBytes);
          if (i < p.kb1) tma_load_2d_a(ring0 + static_cast<uint32_t>(stage * kTailSlabBytes), &tmA, full0 + stage * 8, i * kBK, m0);
          else tma_load_2d_a(ring0 + static_cast<uint32_t>(stage * kTailSlabBytes), &tmR, full0 + stage * 8, (i - p.kb1) * kTailSlab, m0);
        }
        __syncwarp();
        if (++stage == p.ring) { stage = 0; phase ^= 1; }
      }
    }
  }

  // -------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K2t2 — the same fusion for layer2 of RetinaFace-R50: conv3 (128 -> 512) + skip and the next block's conv1 (512 -> 128).
// Neither weight matrix fits beside the pipeline buffers (128 KB each), so both are streamed, and the 512 output channels
// of conv3 are worked on as two halves of 256 (TMEM holds two accumulators of 256 columns):
//
//   G ring (16 KB units, G producer)   per half: for each of conv3's two k-blocks: a, W3 rows 0-127, W3 rows 128-255 of
//                                      the half; then the half's four skip slabs.  20 units per tile.
//   W ring (16 KB units, W producer)   the eight 64-wide k-blocks of the next conv1's weights [128 x 512], 8 per tile
//   issuer    G1(t, h): acc1 = a x W3_h (two N = 128 MMAs per k16 step) + skip slabs x I, into buffer X_t = t & 1 for
//             h = 0 and Y_t = 1 - X_t for h = 1;  M2(t, h): columns [0, 128) of X_t (+)= staged slab x W1n k-block, slabs in
//             channel order over both halves.  G1 and M2 have an issuer warp each.
//   epilogue  ten items per tile: e1 of half 0 (4 slabs), e1 of half 1 (4), e2 (2 slabs of the 128 conv1 channels).
//             Y_t is free for G1(t+1, 0) once e1(t, 1) has drained it, X_t for G1(t+1, 1) once e2(t) has.
//   store warp as in K2t.
// (With ONE issuer walking M2(t,0); G1(t+1,0); M2(t,1); G1(t+1,1) three slab slots deadlock - G1(t+1, 0) waits for all four e1
// items of half 1, whose fourth then needs a slot that only M2(t, 1) frees - which tools/conv_tail_protocol_sim.py found
// before the GPU did.)  Accumulation orders are those of the separate launches: same bits.
constexpr int kTail2Threads = (kTailEpiWarps + 5) * 32;    // + G1 issuer, G producer, store warp, W producer, M2 issuer
constexpr int kTail2N1 = 512, kTail2K1 = 128, kTail2N2 = 128;
constexpr int kTail2Ipt = 10;                              // epilogue items per tile

struct ConvTail2Args {
  int M, m_tiles;
  int gring, wring, slots;
  const float* bias1;    // [512]
  const float* bias2;    // [128]
  int* watchdog;
};

template <bool kBF16>
__global__ void __launch_bounds__(kTail2Threads, 1)
conv_tail2_kernel(const __grid_constant__ ConvTail2Args p, const __grid_constant__ CUtensorMap tmA,
                  const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO1,
                  const __grid_constant__ CUtensorMap tmO2, const __grid_constant__ CUtensorMap tmW3,
                  const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmI) {
  using E = Elem16<kBF16>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  uint8_t* smem_g = smem;                                            // G ring
  uint8_t* smem_w = smem_g + p.gring * kTailSlabBytes;               // W ring
  uint8_t* smem_slot = smem_w + p.wring * kTailSlabBytes;            // output slab slots
  uint8_t* smem_i = smem_slot + p.slots * kTailSlabBytes;            // 64 x 64 identity
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_i + kTailIdentBytes);
  uint64_t* gfull_bar = bars;            // [8]
  uint64_t* gempty_bar = bars + 8;       // [8]
  uint64_t* wfull_bar = bars + 16;       // [4]
  uint64_t* wempty_bar = bars + 20;      // [4]
  uint64_t* slabw_bar = bars + 24;       // [8] e1 slab written -> M2
  uint64_t* sfree_bar = bars + 32;       // [8] slot free: store drained + M2 retired
  uint64_t* staged_bar = bars + 40;      // [8] slab staged -> store warp
  uint64_t* acc1_bar = bars + 48;        // [2] G1 + skip of a half complete (one phase per tile and buffer)
  uint64_t* acc2_bar = bars + 50;        // [2] M2 of both halves complete (buffer X_t: one phase every two tiles)
  uint64_t* free1_bar = bars + 52;       // [2] buffer drained by the four e1 items of a second half (16 warps)
  uint64_t* free2_bar = bars + 54;       // [2] buffer drained by the two e2 items (8 warps)
  uint64_t* bres_bar = bars + 56;
  uint64_t* turn_bar = bars + 57;        // [kTailGroups <= 4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 61);
  float* s_bias1 = reinterpret_cast<float*>(bars + 62);   // [512]
  float* s_bias2 = s_bias1 + kTail2N1;                    // [128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kMmaWarp = kTailEpiWarps, kGWarp = kTailEpiWarps + 1, kStoreWarp = kTailEpiWarps + 2, kWWarp = kTailEpiWarps + 3,
                kM2Warp = kTailEpiWarps + 4;
  const int my_tiles = (p.m_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == kGWarp && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmR); tma_prefetch_desc(&tmO1); tma_prefetch_desc(&tmO2);
    tma_prefetch_desc(&tmW3); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmI);
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int i = 0; i < 8; ++i) {
        mbar_init(&gfull_bar[i], 1);
        mbar_init(&gempty_bar[i], 1);
        mbar_init(&slabw_bar[i], 1);
        mbar_init(&sfree_bar[i], 2);
        mbar_init(&staged_bar[i], 1);
      }
      for (int i = 0; i < 4; ++i) {
        mbar_init(&wfull_bar[i], 1);
        mbar_init(&wempty_bar[i], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&acc1_bar[b], 1);
        mbar_init(&acc2_bar[b], 1);
        mbar_init(&free1_bar[b], 16);
        mbar_init(&free2_bar[b], 8);
      }
      mbar_init(bres_bar, 1);
      for (int g = 0; g < kTailGroups; ++g) mbar_init(&turn_bar[g], 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_dyn(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kTail2N1 + kTail2N2; i += kTail2Threads)
    s_bias1[i] = i < kTail2N1 ? __ldg(p.bias1 + i) : __ldg(p.bias2 + i - kTail2N1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kTailEpiWarps) {
    // ================================================================ epilogue groups
    const int quarter = warp & 3;
    const int grp = warp >> 2;
    const int gtid = threadIdx.x & 127;
    const bool leader = gtid == 0;
    const int row = quarter * 32 + lane;
    const uint32_t sw7 = static_cast<uint32_t>(row & 7);
    const int my_items = my_tiles * kTail2Ipt;
    int t = grp / kTail2Ipt, j = grp - t * kTail2Ipt;
    int slot = grp % p.slots, use = grp / p.slots;
    for (int q = grp; q < my_items; q += kTailGroups) {
      const int X = t & 1;
      const bool second = j >= 8;                                    // e2 item
      const int h = j < 4 ? 0 : 1;                                   // half (e1 items)
      const int s = second ? j - 8 : (j & 3);
      const int buf = (second || h == 0) ? X : 1 - X;
      if (leader) {                            // slots in item order, one waiting thread per group: see K2t
        if (q > 0) mbar_wait(&turn_bar[(grp + kTailGroups - 1) % kTailGroups], static_cast<uint32_t>((q - 1) / kTailGroups) & 1u, p.watchdog, 130 + grp);
        mbar_wait(&sfree_bar[slot], (static_cast<uint32_t>(use) & 1u) ^ 1u, p.watchdog, 100 + slot);
        mbar_arrive(&turn_bar[grp]);
      }
      named_bar_sync(1 + grp, 128);
      if (second) mbar_wait(&acc2_bar[buf], static_cast<uint32_t>(t >> 1) & 1u, p.watchdog, 120 + buf);
      else mbar_wait(&acc1_bar[buf], static_cast<uint32_t>(t) & 1u, p.watchdog, 110 + buf);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * kTailN1 + s * kTailSlab);
      const float* bias = second ? s_bias2 + s * kTailSlab : s_bias1 + h * kTailN1 + s * kTailSlab;
      uint8_t* const orow = smem_slot + slot * kTailSlabBytes + static_cast<uint32_t>(row) * 128u;
      uint32_t v0[16], v1[16];
      auto finish16 = [&](const uint32_t* v, int cbase) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int cc = cbase + c * 8;
          const float4 b0 = *reinterpret_cast<const float4*>(bias + cc);
          const float4 b1 = *reinterpret_cast<const float4*>(bias + cc + 4);
          uint4 o;
          o.x = E::from_f2_relu(__uint_as_float(v[c * 8 + 0]) + b0.x, __uint_as_float(v[c * 8 + 1]) + b0.y);
          o.y = E::from_f2_relu(__uint_as_float(v[c * 8 + 2]) + b0.z, __uint_as_float(v[c * 8 + 3]) + b0.w);
          o.z = E::from_f2_relu(__uint_as_float(v[c * 8 + 4]) + b1.x, __uint_as_float(v[c * 8 + 5]) + b1.y);
          o.w = E::from_f2_relu(__uint_as_float(v[c * 8 + 6]) + b1.z, __uint_as_float(v[c * 8 + 7]) + b1.w);
          *reinterpret_cast<uint4*>(orow + ((static_cast<uint32_t>(cc >> 3) ^ sw7) << 4)) = o;
        }
      };
      tmem_ld16(taddr, v0);
      tc_wait_ld16(v0);
      tmem_ld16(taddr + 16, v1);
      finish16(v0, 0);
      tc_wait_ld16(v1);
      tmem_ld16(taddr + 32, v0);
      finish16(v1, 16);
      tc_wait_ld16(v0);
      tmem_ld16(taddr + 48, v1);
      finish16(v0, 32);
      tc_wait_ld16(v1);
      tc_fence_before();                       // the accumulator columns of this warp are in registers
      if (second || h == 1) {                  // this warp's part of a buffer hand-over (e2: X_t, e1 of half 1: Y_t)
        __syncwarp();
        if (lane == 0) mbar_arrive(second ? &free2_bar[buf] : &free1_bar[buf]);
      }
      finish16(v1, 48);
      fence_proxy_async_smem();
      named_bar_sync(1 + grp, 128);
      if (leader) {
        if (!second) mbar_arrive(&slabw_bar[slot]);
        mbar_arrive(&staged_bar[slot]);
      }
      j += kTailGroups;
      while (j >= kTail2Ipt) { j -= kTail2Ipt; ++t; }
      slot += kTailGroups;
      while (slot >= p.slots) { slot -= p.slots; ++use; }
    }
  } else if (warp == kStoreWarp) {
    // ================================================================ store warp
    if (lane == 0) {
      const int my_items = my_tiles * kTail2Ipt;
      int t = 0, j = 0, slot = 0;
      uint32_t par = 0;
      int pend_slot = -1, last_slot = -1;
      auto give_back = [&](int ps) {
        mbar_arrive(&sfree_bar[ps & 15]);
        if (ps & 16) mbar_arrive(&sfree_bar[ps & 15]);
      };
      for (int q = 0; q < my_items; ++q) {
        const bool second = j >= 8;
        const int m0 = (static_cast<int>(blockIdx.x) + t * static_cast<int>(gridDim.x)) * kBM;
        mbar_wait(&staged_bar[slot], (par >> slot) & 1u, p.watchdog, 400 + slot);
        par ^= 1u << slot;
        if (second) tma_store_2d(&tmO2, smem_slot + slot * kTailSlabBytes, (j - 8) * kTailSlab, m0);
        else tma_store_2d(&tmO1, smem_slot + slot * kTailSlabBytes, j * kTailSlab, m0);       // j = 4 h + s: channel 64 j
        tma_store_commit();
        pend_slot = last_slot;
        last_slot = slot | (second ? 16 : 0);
        if (pend_slot >= 0) {
          tma_store_wait_read<1>();
          give_back(pend_slot);
        }
        if (++j == kTail2Ipt) { j = 0; ++t; }
        if (++slot == p.slots) slot = 0;
      }
      if (last_slot >= 0) {
        tma_store_wait_read<0>();
        give_back(last_slot);
      }
      tma_store_wait<0>();
    }
  } else if (warp == kMmaWarp) {
    // ================================================================ G1 issuer
    const uint32_t idesc_h = make_idesc_f16(kBM, 128, kBF16 ? 1 : 0);
    const uint32_t idesc_r = make_idesc_f16(kBM, kTailSlab, kBF16 ? 1 : 0);
    const uint32_t hi = desc_hi(1024);
    const uint32_t g_lo = desc_lo(smem_u32(smem_g)), id_lo = desc_lo(smem_u32(smem_i));
    const uint32_t step = kTailSlabBytes >> 4;
    const uint32_t gfull0 = smem_u32(gfull_bar), gempty0 = smem_u32(gempty_bar), acc10 = smem_u32(acc1_bar),
                   free10 = smem_u32(free1_bar), free20 = smem_u32(free2_bar);
    int gs = 0;
    uint32_t gph = 0;
    mbar_wait(bres_bar, 0, p.watchdog, 250);
    auto mma4 = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, uint32_t acc0) {
      tc_mma_f16(d, pack_desc(a, hi), pack_desc(b, hi), idesc, acc0);
      tc_mma_f16(d, pack_desc(a + 2, hi), pack_desc(b + 2, hi), idesc, 1u);
      tc_mma_f16(d, pack_desc(a + 4, hi), pack_desc(b + 4, hi), idesc, 1u);
      tc_mma_f16(d, pack_desc(a + 6, hi), pack_desc(b + 6, hi), idesc, 1u);
    };
    auto g_take = [&](int& unit) {             // wait for the next G-ring unit, return its index, advance
      mbar_wait_a(gfull0 + gs * 8, gph, p.watchdog, 210 + gs);
      unit = gs;
      if (++gs == p.gring) { gs = 0; gph ^= 1; }
    };
    auto issue_g1 = [&](int t, int h) {
      const int X = t & 1;
      const int buf = h == 0 ? X : 1 - X;
      const uint32_t d = tmem_base + static_cast<uint32_t>(buf * kTailN1);
      if (t >= 1) {
        const uint32_t par = static_cast<uint32_t>((t - 1) >> 1) & 1u;
        mbar_wait_a((h == 0 ? free10 : free20) + buf * 8, par, p.watchdog, 200 + buf + 2 * h);
      }
      tc_fence_after();
      for (int kb = 0; kb < 2; ++kb) {
        int ua, ub0, ub1;
        g_take(ua); g_take(ub0); g_take(ub1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a = g_lo + static_cast<uint32_t>(ua) * step;
          mma4(d, a, g_lo + static_cast<uint32_t>(ub0) * step, idesc_h, kb != 0 ? 1u : 0u);
          mma4(d + 128u, a, g_lo + static_cast<uint32_t>(ub1) * step, idesc_h, kb != 0 ? 1u : 0u);
          tc_commit_a(gempty0 + ua * 8);
          tc_commit_a(gempty0 + ub0 * 8);
          tc_commit_a(gempty0 + ub1 * 8);
        }
        __syncwarp();
      }
      for (int rs = 0; rs < 4; ++rs) {
        int u;
        g_take(u);
        tc_fence_after();
        if (elect_one()) {
          mma4(d + static_cast<uint32_t>(rs * kTailSlab), g_lo + static_cast<uint32_t>(u) * step, id_lo, idesc_r, 1u);
          tc_commit_a(gempty0 + u * 8);
          if (rs == 3) tc_commit_a(acc10 + buf * 8);
        }
        __syncwarp();
      }
    };
    if (my_tiles > 0) {
      issue_g1(0, 0);
      issue_g1(0, 1);
    }
    for (int t = 1; t < my_tiles; ++t) {
      issue_g1(t, 0);
      issue_g1(t, 1);
    }
  } else if (warp == kM2Warp) {
    // ================================================================ M2 issuer
    // A warp of its own: a G1 that waits for its loads (the G ring turns over three times per tile) must not hold up the
    // M2 MMAs of slabs that are ready, and with them e2 and the hand-over of the buffer the next G1 needs.  What orders the
    // two issuers' MMAs on a TMEM buffer are the barriers (slab written, buffer drained), never program order.
    const uint32_t idesc_h = make_idesc_f16(kBM, 128, kBF16 ? 1 : 0);
    const uint32_t hi = desc_hi(1024);
    const uint32_t w_lo = desc_lo(smem_u32(smem_w)), slot_lo = desc_lo(smem_u32(smem_slot));
    const uint32_t step = kTailSlabBytes >> 4;
    const uint32_t wfull0 = smem_u32(wfull_bar), wempty0 = smem_u32(wempty_bar), slabw0 = smem_u32(slabw_bar), sfree0 = smem_u32(sfree_bar),
                   acc20 = smem_u32(acc2_bar);
    int ws = 0;
    uint32_t wph = 0;
    auto mma4 = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, uint32_t acc0) {
      tc_mma_f16(d, pack_desc(a, hi), pack_desc(b, hi), idesc, acc0);
      tc_mma_f16(d, pack_desc(a + 2, hi), pack_desc(b + 2, hi), idesc, 1u);
      tc_mma_f16(d, pack_desc(a + 4, hi), pack_desc(b + 4, hi), idesc, 1u);
      tc_mma_f16(d, pack_desc(a + 6, hi), pack_desc(b + 6, hi), idesc, 1u);
    };
    int slot = 0;
    uint32_t e1par = 0;
    auto issue_m2 = [&](int t, int h) {
      const int X = t & 1;
      const uint32_t d = tmem_base + static_cast<uint32_t>(X * kTailN1);
      for (int s = 0; s < 4; ++s) {
        mbar_wait_a(slabw0 + slot * 8, (e1par >> slot) & 1u, p.watchdog, 230 + slot);
        if (s == 0 && h == 0) {                // acc2 covers the columns of slabs 0 and 1 of half 0
          const int s1 = slot + 1 == p.slots ? 0 : slot + 1;
          mbar_wait_a(slabw0 + s1 * 8, (e1par >> s1) & 1u, p.watchdog, 240 + s1);
        }
        mbar_wait_a(wfull0 + ws * 8, wph, p.watchdog, 260 + ws);
        tc_fence_after();
        if (elect_one()) {
          mma4(d, slot_lo + static_cast<uint32_t>(slot) * step, w_lo + static_cast<uint32_t>(ws) * step, idesc_h, (h != 0 || s != 0) ? 1u : 0u);
          tc_commit_a(sfree0 + slot * 8);
          tc_commit_a(wempty0 + ws * 8);
          if (h == 1 && s == 3) tc_commit_a(acc20 + X * 8);
        }
        __syncwarp();
        e1par ^= 1u << slot;
        if (++slot == p.slots) slot = 0;
        if (++ws == p.wring) { ws = 0; wph ^= 1; }
      }
      if (h == 1)                              // the tile's two e2 items take the next two slots
        for (int e = 0; e < 2; ++e)
          if (++slot == p.slots) slot = 0;
    };
    for (int t = 0; t < my_tiles; ++t) {
      issue_m2(t, 0);
      issue_m2(t, 1);
    }
  } else if (warp == kGWarp) {
    // ================================================================ G producer
    const uint32_t g0 = smem_u32(smem_g);
    const uint32_t gfull0 = smem_u32(gfull_bar), gempty0 = smem_u32(gempty_bar);
    if (lane == 0) {
      const uint32_t bres = smem_u32(bres_bar);
      mbar_expect_tx_a(bres, kTailIdentBytes);
      tma_load_2d_a(smem_u32(smem_i), &tmI, bres, 0, 0);
    }
    int gs = 0;
    uint32_t gph = 0;
    auto put = [&](const CUtensorMap* m, int x, int y) {
      mbar_wait_a(gempty0 + gs * 8, gph ^ 1u, p.watchdog, 300 + gs);
      if (elect_one()) {
        mbar_expect_tx_a(gfull0 + gs * 8, kTailSlabBytes);
        tma_load_2d_a(g0 + static_cast<uint32_t>(gs * kTailSlabBytes), m, gfull0 + gs * 8, x, y);
      }
      __syncwarp();
      if (++gs == p.gring) { gs = 0; gph ^= 1; }
    };
    for (int t = 0; t < my_tiles; ++t) {
      const int m0 = (static_cast<int>(blockIdx.x) + t * static_cast<int>(gridDim.x)) * kBM;
      for (int h = 0; h < 2; ++h) {
        for (int kb = 0; kb < 2; ++kb) {
          put(&tmA, kb * kBK, m0);
          put(&tmW3, kb * kBK, h * kTailN1);
          put(&tmW3, kb * kBK, h * kTailN1 + 128);
        }
        for (int rs = 0; rs < 4; ++rs) put(&tmR, h * kTailN1 + rs * kTailSlab, m0);
      }
    }
  } else {
    // ================================================================ W producer
    const uint32_t w0 = smem_u32(smem_w);
    const uint32_t wfull0 = smem_u32(wfull_bar), wempty0 = smem_u32(wempty_bar);
    int ws = 0;
    uint32_t wph = 0;
    for (int t = 0; t < my_tiles; ++t)
      for (int kb = 0; kb < 8; ++kb) {
        mbar_wait_a(wempty0 + ws * 8, wph ^ 1u, p.watchdog, 320 + ws);
        if (elect_one()) {
          mbar_expect_tx_a(wfull0 + ws * 8, kTailSlabBytes);
          tma_load_2d_a(w0 + static_cast<uint32_t>(ws * kTailSlabBytes), &tmW1, wfull0 + ws * 8, kb * kBK, 0);
        }
        __syncwarp();
        if (++ws == p.wring) { ws = 0; wph ^= 1; }
      }
  }

  // -------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, 512);
  }
}

bool conv_tail2_applicable(const ConvLayer& L3, const ConvLayer& L1) {
  auto lin = [](const ConvLayer& L) { return L.kh == 1 && L.kw == 1 && L.sh == 1 && L.sw == 1 && L.ph == 0 && L.pw == 0 && !L.special; };
  return lin(L3) && lin(L1) && L3.Cout == kTail2N1 && L3.Cout_pad == kTail2N1 && L3.Cin == kTail2K1 && L3.K_pad == kTail2K1 &&
         L1.Cin == kTail2N1 && L1.K_pad == kTail2N1 && L1.Cout == kTail2N2 && L1.Cout_pad == kTail2N2 && L3.act_kind == 1 &&
         L1.act_kind == 1;
}

int conv_tail2_launch(frb_ctx* ctx, const ConvLayer& L3, const ConvLayer& L1, const void* in, int in_ld, long long M,
                      const void* res, int res_ld, void* out1, int out1_ld, void* out2, int out2_ld, cudaStream_t stream) {
  if (!conv_tail2_applicable(L3, L1)) return set_err(ctx, FRB_ERR_INVALID, "conv_tail2: layer pair not supported");
  if (M <= 0 || M > 0x7fffffffLL) return set_err(ctx, FRB_ERR_INVALID, "conv_tail2: bad problem size");
  const void* ptrs[4] = {in, res, out1, out2};
  const int lds[4] = {in_ld, res_ld, out1_ld, out2_ld};
  for (int i = 0; i < 4; ++i)
    if ((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) != 0 || (lds[i] % 8) != 0)
      return set_err(ctx, FRB_ERR_INVALID, "conv_tail2: operands must be 16-byte aligned with a pixel stride that is a multiple of 8");
  ConvTail2Args a;
  memset(&a, 0, sizeof(a));
  a.M = static_cast<int>(M);
  a.m_tiles = static_cast<int>((M + kBM - 1) / kBM);
  a.bias1 = L3.d_bias;
  a.bias2 = L1.d_bias;
  a.watchdog = ctx->wd_dev;
  const int fixed = kTailIdentBytes + 62 * 8 + (kTail2N1 + kTail2N2) * 4 + 1024;
  const int units = (227 * 1024 - fixed) / kTailSlabBytes;               // 13
  // the G ring is what hides the load latency (20 units per tile): measured per launch with 7 / 6 / 5 units 1.30 / 1.36 / 1.52 ms
  a.slots = 4;
  a.wring = 2;
  a.gring = std::min(8, units - a.slots - a.wring);
#ifdef FRB_DEBUG_BUILD
  if (const char* e = getenv("FRB_TAIL2_SPLIT")) {                       // tuning experiment: "SLOTSxWRING"
    int sl = 4, wr = 3;
    if (sscanf(e, "%dx%d", &sl, &wr) == 2 && sl >= 2 && sl <= 8 && wr >= 2 && wr <= 4 && units - sl - wr >= 4) {
      a.slots = sl; a.wring = wr; a.gring = std::min(8, units - sl - wr);
    }
  }
#endif
  if (a.gring < 4) return set_err(ctx, FRB_ERR_INVALID, "conv_tail2: shared-memory plan does not fit");
  const size_t smem = static_cast<size_t>(fixed) + static_cast<size_t>(a.gring + a.wring + a.slots) * kTailSlabBytes;

  void* ident = nullptr;
  int rc = ctx_identity(ctx, &ident);
  if (rc != FRB_OK) return rc;
  CUtensorMap tmA, tmR, tmO1, tmO2, tmW3, tmW1, tmI;
  if ((rc = encode_2d(ctx, &tmA, in, kTail2K1, M, static_cast<uint64_t>(in_ld) * 2, kBK, kBM)) != FRB_OK) return rc;
  if ((rc = encode_2d(ctx, &tmR, res, kTail2N1, M, static_cast<uint64_t>(res_ld) * 2, kTailSlab, kBM)) != FRB_OK) return rc;
  if ((rc = encode_2d(ctx, &tmO1, out1, kTail2N1, M, static_cast<uint64_t>(out1_ld) * 2, kTailSlab, kBM)) != FRB_OK) return rc;
  if ((rc = encode_2d(ctx, &tmO2, out2, kTail2N2, M, static_cast<uint64_t>(out2_ld) * 2, kTailSlab, kBM)) != FRB_OK) return rc;
  if ((rc = encode_2d(ctx, &tmW3, L3.d_w, L3.K_pad, L3.Cout_pad, static_cast<uint64_t>(L3.K_pad) * 2, kBK, 128)) != FRB_OK) return rc;
  if ((rc = encode_2d(ctx, &tmW1, L1.d_w, L1.K_pad, L1.Cout_pad, static_cast<uint64_t>(L1.K_pad) * 2, kBK, 128)) != FRB_OK) return rc;
  if ((rc = encode_2d(ctx, &tmI, ident, 64, 64, 128, 64, 64)) != FRB_OK) return rc;

  const int grid = std::min(a.m_tiles, ctx->num_sms);
  if (ctx->use_bf16) {
    FRB_CUDA(ctx, cudaFuncSetAttribute(conv_tail2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    conv_tail2_kernel<true><<<grid, kTail2Threads, smem, stream>>>(a, tmA, tmR, tmO1, tmO2, tmW3, tmW1, tmI);
  } else {
    FRB_CUDA(ctx, cudaFuncSetAttribute(conv_tail2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    conv_tail2_kernel<false><<<grid, kTail2Threads, smem, stream>>>(a, tmA, tmR, tmO1, tmO2, tmW3, tmW1, tmI);
  }
  ctx->launches++;
  return check_last(ctx, "conv_tail2_kernel launch");
}

// Which pairs of layers the fused kernel takes: conv3 (1x1, Cin 64 or 128 -> 256, ReLU after the skip) followed by the next
// block's conv1 (1x1, 256 -> 64 or 128, ReLU), both with every channel on ReLU.
bool conv_tail_applicable(const ConvLayer& L3, const ConvLayer& L1) {
  auto lin = [](const ConvLayer& L) { return L.kh == 1 && L.kw == 1 && L.sh == 1 && L.sw == 1 && L.ph == 0 && L.pw == 0 && !L.special; };
  return lin(L3) && lin(L1) && L3.Cout == kTailN1 && L3.Cout_pad == kTailN1 && (L3.Cin == 64 || L3.Cin == 128) && L3.K_pad == L3.Cin &&
         L1.Cin == kTailN1 && L1.K_pad == kTailN1 && (L1.Cout == 64 || L1.Cout == 128) && L1.Cout_pad == L1.Cout && L3.act_kind == 1 &&
         L1.act_kind == 1;
}

// A skip conv that K2t can compute itself: 1x1 / stride 1, 64 or 128 -> 256 channels, linear (the bottleneck's ReLU comes
// after the add).  Its output then never exists in memory, nor in 16 bits: one rounding fewer than the separate launches.
bool conv_tail_skip_applicable(const ConvLayer& Ld) {
  return Ld.kh == 1 && Ld.kw == 1 && Ld.sh == 1 && Ld.sw == 1 && Ld.ph == 0 && Ld.pw == 0 && !Ld.special && Ld.Cout == kTailN1 &&
         Ld.Cout_pad == kTailN1 && Ld.Cin == 64 && Ld.K_pad == Ld.Cin && Ld.act_kind == 2;
}

// res: the skip tensor [M, 256], or with Ld != null the INPUT x [M, Ld->Cin] of the skip conv Ld (bias1 = conv3's bias + Ld's)
int conv_tail_launch(frb_ctx* ctx, const ConvLayer& L3, const ConvLayer& L1, const void* in, int in_ld, long long M,
                     const void* res, int res_ld, void* out1, int out1_ld, void* out2, int out2_ld, const ConvLayer* Ld,
                     const float* bias1, cudaStream_t stream) {
  if (!conv_tail_applicable(L3, L1) || (Ld != nullptr && (!conv_tail_skip_applicable(*Ld) || bias1 == nullptr)))
    return set_err(ctx, FRB_ERR_INVALID, "conv_tail: layer pair not supported");
  if (M <= 0 || M > 0x7fffffffLL) return set_err(ctx, FRB_ERR_INVALID, "conv_tail: bad problem size");
  const void* ptrs[4] = {in, res, out1, out2};
  const int lds[4] = {in_ld, res_ld, out1_ld, out2_ld};
  for (int i = 0; i < 4; ++i)
    if ((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) != 0 || (lds[i] % 8) != 0)
      return set_err(ctx, FRB_ERR_INVALID, "conv_tail: operands must be 16-byte aligned with a pixel stride that is a multiple of 8");
  ConvTailArgs a;
  memset(&a, 0, sizeof(a));
  a.M = static_cast<int>(M);
  a.m_tiles = static_cast<int>((M + kBM - 1) / kBM);
  a.kb1 = L3.Cin / kBK;
  a.N2 = L1.Cout;
  a.kbx = Ld != nullptr ? Ld->Cin / kBK : 0;
  a.bias1 = Ld != nullptr ? bias1 : L3.d_bias;
  a.bias2 = L1.d_bias;
  a.watchdog = ctx->wd_dev;
  const int fixed = a.kb1 * kTailN1 * 128 + 4 * a.N2 * 128 + (a.kbx ? a.kbx * kTailN1 * 128 : kTailIdentBytes) + 52 * 8 +
                    (kTailN1 + 128) * 4 + 1024;
  const int units = (227 * 1024 - fixed) / kTailSlabBytes;               // 16 KB buffers left for the ring and the slab slots
  // The ring is what hides the HBM latency (a tile's loads: conv3's input and the four skip slabs); the output slabs need
  // a slot for an epilogue pass plus a store drain.  Measured on one box (profiles/r02_layer_times_r50_fuse_ab.txt), 9 units:
  // ring 4 / 5 / 6 / 7 = 1.72 / 1.66 / 1.65 / 1.91 ms per launch; 7 units (N2 = 128): ring 3 / 4 / 5 = 2.42 / 2.23 / 2.19 ms.
  a.slots = units >= 9 ? 4 : 3;
  if (a.kbx) a.slots = units - 4;                                        // two loads per tile: a short ring is enough
  a.ring = std::min(kTailMaxRing, units - a.slots);
#ifdef FRB_DEBUG_BUILD
  if (const char* e = getenv("FRB_TAIL_RING")) { a.ring = std::max(2, std::min(units - 2, atoi(e))); a.slots = units - a.ring; }   // tuning experiment
#endif
  a.slots = std::min(kTailMaxSlots, a.slots);
  if (a.ring < 2 || a.slots < 2) return set_err(ctx, FRB_ERR_INVALID, "conv_tail: shared-memory plan does not fit");
  const size_t smem = static_cast<size_t>(fixed) + static_cast<size_t>(a.ring + a.slots) * kTailSlabBytes;

  void* ident = nullptr;
  int rc = ctx_identity(ctx, &ident);
  if (rc != FRB_OK) return rc;
  CUtensorMap tmA, tmR, tmO1, tmO2, tmW3, tmW1, tmI;
  if ((rc = encode_2d(ctx, &tmA, in, L3.Cin, M, static_cast<uint64_t>(in_ld) * 2, kBK, kBM)) != FRB_OK) return rc;
  if ((rc = encode_2d(ctx, &tmR, res, Ld != nullptr ? Ld->Cin : kTailN1, M, static_cast<uint64_t>(res_ld) * 2, kTailSlab, kBM)) != FRB_OK) return rc;
  if ((rc = encode_2d(ctx, &tmO1, out1, kTailN1, M, static_cast<uint64_t>(out1_ld) * 2, kTailSlab, kBM)) != FRB_OK) return rc;
  if ((rc = encode_2d(ctx, &tmO2, out2, a.N2, M, static_cast<uint64_t>(out2_ld) * 2, kTailSlab, kBM)) != FRB_OK) return rc;
  if ((rc = encode_2d(ctx, &tmW3, L3.d_w, L3.K_pad, L3.Cout_pad, static_cast<uint64_t>(L3.K_pad) * 2, kBK, kTailN1)) != FRB_OK) return rc;
  if ((rc = encode_2d(ctx, &tmW1, L1.d_w, L1.K_pad, L1.Cout_pad, static_cast<uint64_t>(L1.K_pad) * 2, kBK, a.N2)) != FRB_OK) return rc;
  if (Ld != nullptr) rc = encode_2d(ctx, &tmI, Ld->d_w, Ld->K_pad, Ld->Cout_pad, static_cast<uint64_t>(Ld->K_pad) * 2, kBK, kTailN1);
  else rc = encode_2d(ctx, &tmI, ident, 64, 64, 128, 64, 64);
  if (rc != FRB_OK) return rc;

  const int grid = std::min(a.m_tiles, ctx->num_sms);
  if (ctx->use_bf16) {
    FRB_CUDA(ctx, cudaFuncSetAttribute(conv_tail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    conv_tail_kernel<true><<<grid, kTailThreads, smem, stream>>>(a, tmA, tmR, tmO1, tmO2, tmW3, tmW1, tmI);
  } else {
    FRB_CUDA(ctx, cudaFuncSetAttribute(conv_tail_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    conv_tail_kernel<false><<<grid, kTailThreads, smem, stream>>>(a, tmA, tmR, tmO1, tmO2, tmW3, tmW1, tmI);
  }
  ctx->launches++;
  return check_last(ctx, "conv_tail_kernel launch");
}

}  // namespace frb
