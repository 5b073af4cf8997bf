// Complete-linkage agglomerative clustering on the GPU: replaces the host side of
// scipy.cluster.hierarchy.linkage(y, "complete") at facerec/cluster.py:41,199 once the distance matrix is
// already in HBM (frb_pdist, FRB_SQUARE_F32), so the 5-10 GB device->host copy, the fp64 up-cast and the
// ~2 minute single-thread linkage at N = 50 000 disappear.
//
// Algorithm: the nearest-neighbour chain (the one scipy and fastcluster use for complete linkage), kept
// decision-for-decision identical so that the dendrogram — including every tie — is the one scipy builds
// from the same distances: the chain grows from its tip x to argmin_i D[x,i] over live clusters, ties go
// to the lowest index, the previous chain element is preferred on equality, reciprocal neighbours merge
// (lower index dies, higher index becomes the union), D[i, x u y] = max(D[i,x], D[i,y]).  max never
// creates a new value, so fp32 distances are exact under the whole run.
//
// One persistent cooperative kernel: every block scans a slice of row x (coalesced), a grid-wide sync,
// every block redundantly reduces the per-block minima and takes the same decision, then (on a merge) all
// blocks update row and column y.  Two grid syncs per chain step.
#include <cooperative_groups.h>

#include "host.h"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace frb {

constexpr int kLinkThreads = 256;

struct LinkPartial {
  float d;
  int i;
};

__device__ __forceinline__ bool lp_less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }

__global__ void __launch_bounds__(kLinkThreads)
linkage_complete_kernel(float* D, int n, int* size, int* chain, LinkPartial* partial /*[2][gridDim.x]*/,
                        int* __restrict__ z_xy /*[n-1][2]*/, float* __restrict__ z_d /*[n-1]*/) {
  // D, size, chain and partial are rewritten by other blocks between grid syncs: read them through L2 (__ldcg)
  cg::grid_group grid = cg::this_grid();
  __shared__ float s_d[kLinkThreads / 32];
  __shared__ int s_i[kLinkThreads / 32];
  __shared__ float s_bd;
  __shared__ int s_bi;
  const int nb = gridDim.x;
  const long long N = n;
  // control state, replicated (identically) in every block
  int chain_len = 0, merges = 0, cursor = 0, iter = 0;
  while (merges < n - 1) {
    int tip, prev;
    if (chain_len == 0) {
      while (__ldcg(size + cursor) == 0) ++cursor;        // the lowest live cluster starts a new chain
      tip = cursor;
      prev = -1;
      chain_len = 1;
      if (blockIdx.x == 0 && threadIdx.x == 0) chain[0] = cursor;
    } else {                                     // chain[] entries read here were written >= 1 grid sync ago
      tip = __ldcg(chain + chain_len - 1);
      prev = chain_len >= 2 ? __ldcg(chain + chain_len - 2) : -1;
    }
    const float* row = D + static_cast<long long>(tip) * N;
    float cur_min = prev >= 0 ? __ldcg(row + prev) : INFINITY;
    // ---- phase 1: per-block argmin (distance, then index) over a slice of row `tip`
    float bd = INFINITY;
    int bi = 0x7fffffff;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nb * blockDim.x) {
      if (i == tip || __ldcg(size + i) == 0) continue;
      const float d = __ldcg(row + i);
      if (lp_less(d, i, bd, bi)) { bd = d; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (lp_less(od, oi, bd, bi)) { bd = od; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_d[threadIdx.x >> 5] = bd; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
      bd = threadIdx.x < kLinkThreads / 32 ? s_d[threadIdx.x] : INFINITY;
      bi = threadIdx.x < kLinkThreads / 32 ? s_i[threadIdx.x] : 0x7fffffff;
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (lp_less(od, oi, bd, bi)) { bd = od; bi = oi; }
      }
      if (threadIdx.x == 0) {
        partial[(iter & 1) * nb + blockIdx.x].d = bd;
        partial[(iter & 1) * nb + blockIdx.x].i = bi;
      }
    }
    grid.sync();
    // ---- phase 2: every block reduces all per-block minima and takes the same decision
    if (threadIdx.x < 32) {
      bd = INFINITY;
      bi = 0x7fffffff;
      for (int b = threadIdx.x; b < nb; b += 32) {
        const float pd = __ldcg(&partial[(iter & 1) * nb + b].d);
        const int pi = __ldcg(&partial[(iter & 1) * nb + b].i);
        if (lp_less(pd, pi, bd, bi)) { bd = pd; bi = pi; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (lp_less(od, oi, bd, bi)) { bd = od; bi = oi; }
      }
      if (threadIdx.x == 0) { s_bd = bd; s_bi = bi; }
    }
    __syncthreads();
    int y = prev;
    if (s_bd < cur_min) {        // strict: the previous chain element wins ties (no cycles)
      cur_min = s_bd;
      y = s_bi;
    }
    __syncthreads();             // s_bd / s_bi are rewritten next iteration
    ++iter;
    if (chain_len > 1 && y == prev) {
      // ---- reciprocal nearest neighbours: merge; the lower index dies, the higher index becomes the union
      chain_len -= 2;
      const int mx = min(tip, y), my = max(tip, y);
      const float* rx = D + static_cast<long long>(mx) * N;
      float* ry = D + static_cast<long long>(my) * N;
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nb * blockDim.x) {
        if (i == my || i == mx || __ldcg(size + i) == 0) continue;
        const float v = fmaxf(__ldcg(rx + i), __ldcg(ry + i));
        ry[i] = v;                                        // row my (coalesced)
        D[static_cast<long long>(i) * N + my] = v;        // column my (keeps every live row exact)
      }
      if (blockIdx.x == 0 && threadIdx.x == 0) {          // size[mx], size[my] are not read by the loop above
        z_xy[2 * merges] = mx;
        z_xy[2 * merges + 1] = my;
        z_d[merges] = cur_min;
        size[my] += size[mx];
        size[mx] = 0;
      }
      ++merges;
    } else {
      if (blockIdx.x == 0 && threadIdx.x == 0) chain[chain_len] = y;
      ++chain_len;
    }
    grid.sync();                 // D / size[] / chain[] updates visible before the next scan
  }
}

int run_linkage_complete(frb_ctx* ctx, float* d_D, int n, int32_t* d_zxy, float* d_zd, cudaStream_t st) {
  if (n < 2) return set_err(ctx, FRB_ERR_INVALID, "linkage: need at least 2 observations");
  int per_sm = 0;
  FRB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, linkage_complete_kernel, kLinkThreads, 0));
  if (per_sm < 1) return set_err(ctx, FRB_ERR_CUDA, "linkage: kernel does not fit an SM");
  int grid = ctx->num_sms * std::min(per_sm, 2);
  grid = std::min(grid, std::max(1, (n + kLinkThreads - 1) / kLinkThreads));
  void *d_size = nullptr, *d_chain = nullptr, *d_part = nullptr;
  int rc;
  if ((rc = ctx_scratch(ctx, 6, static_cast<size_t>(n) * 8, &d_size)) != FRB_OK) return rc;
  if ((rc = ctx_scratch(ctx, 7, static_cast<size_t>(grid) * 2 * sizeof(LinkPartial) + 64, &d_part)) != FRB_OK) return rc;
  d_chain = static_cast<int*>(d_size) + n;
  std::vector<int> ones(n, 1);
  FRB_CUDA(ctx, cudaMemcpyAsync(d_size, ones.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  FRB_CUDA(ctx, cudaStreamSynchronize(st));      // `ones` is a stack-lifetime host buffer
  int* p_size = static_cast<int*>(d_size);
  int* p_chain = static_cast<int*>(d_chain);
  LinkPartial* p_part = static_cast<LinkPartial*>(d_part);
  void* args[] = {&d_D, &n, &p_size, &p_chain, &p_part, &d_zxy, &d_zd};
  FRB_CUDA(ctx, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(linkage_complete_kernel), dim3(grid), dim3(kLinkThreads), args, 0, st));
  ctx->launches++;
  return check_last(ctx, "linkage_complete_kernel");
}

}  // namespace frb
