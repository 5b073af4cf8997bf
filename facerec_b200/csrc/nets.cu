// Network graphs of the per-frame path, expressed over the K2 conv kernel and the glue kernels:
//   RetinaFace (MobileNetV1-0.25 or ResNet-50 backbone, FPN, SSH, fused heads)  -> loc / conf / landm
//   FaceNet InceptionResnetV1                                                   -> 512-d unit embeddings
//
// Weights arrive as one fp32 blob (BatchNorm folded on the host, see facerec_b200/weights.py).  At
// load time sibling convolutions that read the same tensor are assembled into ONE wider layer
// (SSH c3+c5_1 and c5+c7_2, the three head convs, the parallel 1x1 branch convs of the
// Inception blocks, a block-diagonal pair of 3x3 convs in Block35) and concat buffers are laid out
// so that every producer writes, and every consumer reads, one contiguous channel range: concat,
// channel reorder, residual scale and the final ReLU of SSH cost nothing at run time.
//
// A forward for a given (batch, H, W) is planned once: tensors get arena offsets by liveness
// (greedy interval placement), ops are replayed on the caller's stream.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <stdexcept>
#include <tuple>

#include "host.h"
#include "conv_igemm.cuh"

namespace frb {

// ------------------------------------------------------------------------------------ blob
struct BlobEntry {
  std::vector<int> dims;
  const float* data = nullptr;
  size_t count = 0;
};
struct WeightBlob {
  std::map<std::string, BlobEntry> entries;
  const BlobEntry& get(const std::string& name) const {
    auto it = entries.find(name);
    if (it == entries.end()) throw std::runtime_error("weight blob has no entry '" + name + "'");
    return it->second;
  }
};

static void parse_blob(const void* blob, size_t bytes, WeightBlob& out) {
  const uint8_t* p = static_cast<const uint8_t*>(blob);
  if (bytes < 20 || memcmp(p, "FRBW", 4) != 0) throw std::runtime_error("weight blob: bad magic (expected FRBW)");
  uint32_t version, n;
  uint64_t data_off;
  memcpy(&version, p + 4, 4);
  memcpy(&n, p + 8, 4);
  memcpy(&data_off, p + 12, 8);
  if (version != 1) throw std::runtime_error("weight blob: unsupported version");
  if (data_off > bytes || (data_off & 3)) throw std::runtime_error("weight blob: bad data offset");
  size_t pos = 20;
  for (uint32_t i = 0; i < n; ++i) {
    if (pos + 2 > data_off) throw std::runtime_error("weight blob: truncated entry table");
    uint16_t nl;
    memcpy(&nl, p + pos, 2);
    pos += 2;
    if (pos + nl + 4 > data_off) throw std::runtime_error("weight blob: truncated entry table");
    std::string name(reinterpret_cast<const char*>(p + pos), nl);
    pos += nl;
    uint32_t nd;
    memcpy(&nd, p + pos, 4);
    pos += 4;
    if (nd > 8 || pos + 4 * nd + 8 > data_off) throw std::runtime_error("weight blob: truncated entry table");
    BlobEntry e;
    e.count = 1;
    for (uint32_t d = 0; d < nd; ++d) {
      uint32_t v;
      memcpy(&v, p + pos, 4);
      pos += 4;
      e.dims.push_back(static_cast<int>(v));
      e.count *= v;
    }
    uint64_t eo;
    memcpy(&eo, p + pos, 8);
    pos += 8;
    if (data_off + (eo + e.count) * 4 > bytes) throw std::runtime_error("weight blob: entry '" + name + "' out of range");
    e.data = reinterpret_cast<const float*>(p + data_off) + eo;
    out.entries[name] = e;
  }
}

// ------------------------------------------------------------------------------------ layer assembly
struct HostConv {
  int Cout = 0, kh = 1, kw = 1, Cin = 0;
  std::vector<float> w, b, slope;
  HostConv(int co, int kh_, int kw_, int ci) : Cout(co), kh(kh_), kw(kw_), Cin(ci) {
    w.assign(static_cast<size_t>(co) * kh_ * kw_ * ci, 0.f);
    b.assign(co, 0.f);
    slope.assign(co, 1.f);
  }
};

// Copies source conv `name` (all of it, or the input-channel range [src_ci0, src_ci0+src_cin)) into
// dst at output-channel offset dst_co / input-channel offset dst_ci, scaling weights and bias.
static void place(HostConv& dst, const WeightBlob& blob, const std::string& name, int dst_co, int dst_ci, float slope,
                  float scale = 1.f, int src_ci0 = 0, int src_cin = -1, bool with_bias = true) {
  const BlobEntry& w = blob.get(name + ".w");
  const BlobEntry& b = blob.get(name + ".b");
  if (w.dims.size() != 4) throw std::runtime_error("'" + name + ".w' must be 4-d [Cout,kh,kw,Cin]");
  const int co = w.dims[0], kh = w.dims[1], kw = w.dims[2], ci = w.dims[3];
  if (src_cin < 0) src_cin = ci - src_ci0;
  if (kh != dst.kh || kw != dst.kw || dst_co + co > dst.Cout || dst_ci + src_cin > dst.Cin || src_ci0 + src_cin > ci ||
      static_cast<int>(b.count) != co)
    throw std::runtime_error("layer assembly: '" + name + "' does not fit its slot");
  for (int o = 0; o < co; ++o) {
    for (int t = 0; t < kh * kw; ++t)
      for (int c = 0; c < src_cin; ++c)
        dst.w[(static_cast<size_t>(dst_co + o) * kh * kw + t) * dst.Cin + dst_ci + c] =
            w.data[(static_cast<size_t>(o) * kh * kw + t) * ci + src_ci0 + c] * scale;
    if (with_bias) dst.b[dst_co + o] = b.data[o] * scale;
    dst.slope[dst_co + o] = slope;
  }
}

struct DwLayer {
  int C = 0, stride = 1;
  float slope = 0.f;
  float* d_w = nullptr;   // [9][C]
  float* d_bias = nullptr;
};

// ------------------------------------------------------------------------------------ plan
struct TensorInfo {
  int H = 0, W = 0, C = 0, elem = 2, ext = 0;   // ext: 0 arena, 1 network input, 2 embedding output
  size_t bytes = 0, off = 0;
  int first = -1, last = -1;
};
struct View {
  int t = -1, coff = 0;
};
// OP_CONV_TAIL: conv3 + skip (layer) and the next block's conv1 (layer2) in one launch (conv_tail.cu); out2 = conv1's output
// OP_DWPW: depthwise 3x3 (dws[layer]) + pointwise 1x1 (convs[layer2]) of a MobileNet block in one launch (glue.cu K4f)
enum OpKind : int { OP_CONV = 0, OP_DW = 1, OP_MAXPOOL = 2, OP_AVGPOOL = 3, OP_L2NORM = 4, OP_CONV_TAIL = 5, OP_DWPW = 6 };
struct Op {
  int kind = 0, layer = -1, layer2 = -1, layer3 = -1;   // layer3: skip conv computed inside OP_CONV_TAIL (res = its input)
  View in, out, res, out2;
  int res_mode = 0, out_mode = 0, pad = 0, C = 0, P_off = 0;
};
struct Plan {
  int B = 0, H = 0, W = 0, P_total = 0;
  std::vector<TensorInfo> tensors;
  std::vector<Op> ops;
  size_t arena_bytes = 0;
  // CUDA graph of the forward (short forwards only, see use_graph): captured on the second call of a plan, replayed
  // afterwards.  The graph reads / writes net-owned staging buffers, so it does not depend on the caller's pointers.
  int calls = 0;
  cudaGraphExec_t graph = nullptr;
  long long graph_launches = 0;     // kernels per replay (bench bookkeeping)
  void* graph_arena = nullptr;      // arena / staging addresses baked into the graph
  void* graph_io = nullptr;
};

struct Net {
  int model = 0;
  frb_ctx* ctx = nullptr;
  std::vector<ConvLayer> convs;
  std::map<std::string, int> conv_idx;
  std::vector<DwLayer> dws;
  std::map<std::tuple<int, int, int>, Plan> plans;
  void* arena = nullptr;
  size_t arena_bytes = 0;
  void* io = nullptr;               // staging for the network input and outputs of graph-replayed forwards
  size_t io_bytes = 0;
};

static void drop_graphs(Net* net) {
  for (auto& kv : net->plans)
    if (kv.second.graph) {
      cudaGraphExecDestroy(kv.second.graph);
      kv.second.graph = nullptr;
    }
}

static void add_conv(Net* net, const std::string& name, const HostConv& h, int sh, int sw, int ph, int pw, int bn_cap = 0) {
  ConvLayer L;
  L.bn_cap = bn_cap;
  L.Cin = h.Cin; L.Cout = h.Cout; L.kh = h.kh; L.kw = h.kw; L.sh = sh; L.sw = sw; L.ph = ph; L.pw = pw;
  if (conv_prepare(net->ctx, L, h.w.data(), h.b.data(), h.slope.data()) != FRB_OK) {
    conv_release(L);
    throw std::runtime_error("conv_prepare('" + name + "'): " + net->ctx->err);
  }
  net->conv_idx[name] = static_cast<int>(net->convs.size());
  net->convs.push_back(L);
}

// Single source conv, optional zero-padding of input channels up to cin_pad.
static void add_plain(Net* net, const WeightBlob& blob, const std::string& name, float slope, int sh, int sw, int ph, int pw,
                      int cin_pad = 0, float scale = 1.f, int bn_cap = 0) {
  const BlobEntry& w = blob.get(name + ".w");
  if (w.dims.size() != 4) throw std::runtime_error("'" + name + ".w' must be 4-d");
  HostConv h(w.dims[0], w.dims[1], w.dims[2], std::max(cin_pad, w.dims[3]));
  place(h, blob, name, 0, 0, slope, scale);
  add_conv(net, name, h, sh, sw, ph, pw, bn_cap);
}

static void add_dw(Net* net, const WeightBlob& blob, const std::string& name, int stride, float slope) {
  const BlobEntry& w = blob.get(name + ".w");
  const BlobEntry& b = blob.get(name + ".b");
  if (w.dims.size() != 4 || w.dims[1] != 3 || w.dims[2] != 3 || w.dims[3] != 1) throw std::runtime_error("'" + name + "' is not a depthwise 3x3");
  DwLayer L;
  L.C = w.dims[0]; L.stride = stride; L.slope = slope;
  std::vector<float> w9c(static_cast<size_t>(9) * L.C);
  for (int c = 0; c < L.C; ++c)
    for (int t = 0; t < 9; ++t) w9c[static_cast<size_t>(t) * L.C + c] = w.data[static_cast<size_t>(c) * 9 + t];
  if (cudaMalloc(&L.d_w, w9c.size() * 4) != cudaSuccess || cudaMalloc(&L.d_bias, L.C * 4) != cudaSuccess)
    throw std::runtime_error("cudaMalloc failed for depthwise layer");
  cudaMemcpy(L.d_w, w9c.data(), w9c.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(L.d_bias, b.data, L.C * 4, cudaMemcpyHostToDevice);
  net->dws.push_back(L);
}

// Layers whose epilogue adds a residual tile stage two 64-column slabs of it in shared memory.
// N tile cap of the layers that add a residual: 256 where K is long enough to pay for the wider tile (every N tile of an
// M tile re-reads the A blocks: layer3 / layer4 conv3 + skip went 0.62 -> 0.55 and 0.53 -> 0.36 ms per launch), 128 for
// the short-K ones, which keep a resident weight slice per CTA instead (layer2 conv3, Block35).
static int res_bn_cap(int cin) { return cin >= 256 ? 256 : 128; }

// ------------------------------------------------------------------------------------ architecture tables
static const int kMnetStride[13] = {1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1};
static const int kMnetCout[13] = {16, 32, 32, 64, 64, 128, 128, 128, 128, 128, 128, 256, 256};
static const int kR50Blocks[4] = {3, 4, 6, 3};
static const int kR50Planes[4] = {64, 128, 256, 512};

static std::string sfmt(const char* fmt, ...) {
  char buf[256];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return buf;
}

static void load_retina(Net* net, const WeightBlob& blob, bool r50) {
  const float neck_slope = r50 ? 0.f : 0.1f;   // out_channels <= 64 -> LeakyReLU(0.1), else ReLU
  const int C = r50 ? 256 : 64;
  if (r50) {
    {
      // 7x7 / stride 2 / pad 3 stem on the space-to-depth input (frb_preprocess_u8, FRB_LAYOUT_S2D16): with the
      // kernel zero-extended to 8x8 it is a 4x4 stride-1 conv over 2x2 pixel blocks; the four horizontally
      // adjacent blocks of a window are 64 contiguous elements, so the layer runs as a 4-tap (rows) x 64-channel
      // conv whose A operand is a TMA box over overlapping windows.  K = 256 instead of 7*7*8 = 392.
      const BlobEntry& w = blob.get("conv1.w");
      const BlobEntry& b = blob.get("conv1.b");
      if (w.dims.size() != 4 || w.dims[0] != 64 || w.dims[1] != 7 || w.dims[2] != 7 || w.dims[3] != 3)
        throw std::runtime_error("conv1.w must be [64,7,7,3]");
      HostConv h(64, 4, 1, 64);
      for (int n = 0; n < 64; ++n) {
        for (int r = 0; r < 4; ++r)
          for (int s = 0; s < 4; ++s)
            for (int dy = 0; dy < 2; ++dy)
              for (int dx = 0; dx < 2; ++dx) {
                const int kr = 2 * r + dy - 1, kc = 2 * s + dx - 1;
                if (kr < 0 || kr > 6 || kc < 0 || kc > 6) continue;
                for (int c = 0; c < 3; ++c)
                  h.w[(static_cast<size_t>(n) * 4 + r) * 64 + s * 16 + dy * 8 + dx * 4 + c] =
                      w.data[((static_cast<size_t>(n) * 7 + kr) * 7 + kc) * 3 + c];
              }
        h.b[n] = b.data[n];
        h.slope[n] = 0.f;
      }
      ConvLayer L;
      L.special = 1;
      L.Cin = 64; L.Cout = 64; L.kh = 4; L.kw = 1; L.sh = 1; L.sw = 1; L.ph = 2; L.pw = 0;
      if (conv_prepare(net->ctx, L, h.w.data(), h.b.data(), h.slope.data()) != FRB_OK) {
        conv_release(L);
        throw std::runtime_error("conv_prepare('conv1'): " + net->ctx->err);
      }
      net->conv_idx["conv1"] = static_cast<int>(net->convs.size());
      net->convs.push_back(L);
    }
    for (int l = 0; l < 4; ++l)
      for (int i = 0; i < kR50Blocks[l]; ++i) {
        const int s = (i == 0 && l > 0) ? 2 : 1;
        const std::string p = sfmt("layer%d.%d", l + 1, i);
        add_plain(net, blob, p + ".conv1", 0.f, 1, 1, 0, 0);
        add_plain(net, blob, p + ".conv2", 0.f, s, s, 1, 1);
        add_plain(net, blob, p + ".conv3", 0.f, 1, 1, 0, 0, 0, 1.f, res_bn_cap(kR50Planes[l]));   // ReLU after the residual add
        if (i == 0) add_plain(net, blob, p + ".down", 1.f, s, s, 0, 0);
        if (i == 0) {
          // where K2t can compute the skip conv itself (layer1.0: stride 1), conv3 gets the sum of the two biases
          ConvLayer& L3 = net->convs[net->conv_idx[p + ".conv3"]];
          const ConvLayer& Ld = net->convs[net->conv_idx[p + ".down"]];
          if (conv_tail_skip_applicable(Ld) && L3.Cout_pad == Ld.Cout_pad) {
            std::vector<float> b3(L3.Cout_pad), bd(Ld.Cout_pad);
            if (cudaMemcpy(b3.data(), L3.d_bias, b3.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(bd.data(), Ld.d_bias, bd.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
              throw std::runtime_error("bias read-back failed");
            for (size_t c = 0; c < b3.size(); ++c) b3[c] += bd[c];
            if (cudaMalloc(&L3.d_bias_skip, b3.size() * 4) != cudaSuccess ||
                cudaMemcpy(L3.d_bias_skip, b3.data(), b3.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess)
              throw std::runtime_error("cudaMalloc failed for the fused skip bias");
          }
        }
      }
  } else {
    add_plain(net, blob, "body.0", 0.1f, 2, 2, 1, 1, 8);
    net->convs.back().direct3 = 1;      // K = 27: CUDA cores (glue.cu K3)
    for (int i = 1; i <= 13; ++i) {
      add_dw(net, blob, sfmt("body.%d.dw", i), kMnetStride[i - 1], 0.1f);
      add_plain(net, blob, sfmt("body.%d.pw", i), 0.1f, 1, 1, 0, 0);
    }
  }
  for (int k = 1; k <= 3; ++k) add_plain(net, blob, sfmt("fpn.out%d", k), neck_slope, 1, 1, 0, 0);
  for (int k = 1; k <= 2; ++k) add_plain(net, blob, sfmt("fpn.merge%d", k), neck_slope, 1, 1, 1, 1);
  const int q = C / 4;
  for (int k = 1; k <= 3; ++k) {
    // SSH buffer layout (channels): [c5_1 (q) | c3 (2q) | c7 (q) | c5 (q) | c7_2 (q)]; the block's
    // final ReLU over cat[c3,c5,c7] is applied by the producers of those three slices.
    const std::string s = sfmt("ssh%d", k);
    HostConv a(3 * q, 3, 3, C);
    place(a, blob, s + ".c5_1", 0, 0, neck_slope);
    place(a, blob, s + ".c3", q, 0, 0.f);
    add_conv(net, s + ".A", a, 1, 1, 1, 1);
    HostConv b(2 * q, 3, 3, q);
    place(b, blob, s + ".c5_2", 0, 0, 0.f);
    place(b, blob, s + ".c7_2", q, 0, neck_slope);
    add_conv(net, s + ".B", b, 1, 1, 1, 1);
    HostConv c(q, 3, 3, q);
    place(c, blob, s + ".c7_3", 0, 0, 0.f);
    add_conv(net, s + ".C", c, 1, 1, 1, 1);
    // fused head: columns [cls 4 | box 8 | ldm 20]; input order (c3, c7, c5) vs source (c3, c5, c7)
    HostConv h(32, 1, 1, C);
    const char* parts[3] = {"cls", "box", "ldm"};
    const int co[3] = {0, 4, 12};
    for (int t = 0; t < 3; ++t) {
      const std::string hn = sfmt("head%d.%s", k, parts[t]);
      place(h, blob, hn, co[t], 0, 1.f, 1.f, 0, 2 * q);            // c3
      place(h, blob, hn, co[t], 2 * q, 1.f, 1.f, 3 * q, q);         // c7
      place(h, blob, hn, co[t], 3 * q, 1.f, 1.f, 2 * q, q);         // c5
    }
    add_conv(net, sfmt("head%d", k), h, 1, 1, 0, 0);
  }
}

static void load_irv1(Net* net, const WeightBlob& blob) {
  // The first three convs have 32-channel tensors between them; the TMA conv path moves whole 64-channel blocks.
  // They are assembled with 64 channels (32 real + 32 that are exactly zero: zero weights, zero bias) so that 2a / 2b
  // run on the tcgen05 path with halo tiles and resident weights instead of the gather fallback (at 512 faces 0.76 +
  // 0.79 ms of a 5.1 ms forward); the zero half of K costs MMAs the tensor pipe has to spare.
  {
    HostConv c1(64, 3, 3, 8);
    place(c1, blob, "conv2d_1a", 0, 0, 0.f);
    add_conv(net, "conv2d_1a", c1, 2, 2, 0, 0);       // (K3's direct kernel measured slower here: 64 output channels, 0.38 vs 0.14 ms at 128 faces)
    HostConv c2(64, 3, 3, 64);
    place(c2, blob, "conv2d_2a", 0, 0, 0.f);
    add_conv(net, "conv2d_2a", c2, 1, 1, 0, 0);
    HostConv c3(64, 3, 3, 64);
    place(c3, blob, "conv2d_2b", 0, 0, 0.f);
    add_conv(net, "conv2d_2b", c3, 1, 1, 1, 1);
  }
  {
    // same trick for the 80-channel tensor between 3b and 4a: 128 channels, 48 of them exactly zero
    HostConv c3b(128, 1, 1, 64);
    place(c3b, blob, "conv2d_3b", 0, 0, 0.f);
    add_conv(net, "conv2d_3b", c3b, 1, 1, 0, 0);
    HostConv c4a(192, 3, 3, 128);
    place(c4a, blob, "conv2d_4a", 0, 0, 0.f);
    add_conv(net, "conv2d_4a", c4a, 1, 1, 0, 0);
  }
  add_plain(net, blob, "conv2d_4b", 0.f, 2, 2, 0, 0);
  for (int i = 0; i < 5; ++i) {
    // Block35 buffer: [b1.0 | b2.0 | b0 | b2.2 | b1.1 | b2.1], 32 channels each
    const std::string p = sfmt("repeat_1.%d", i);
    HostConv a(96, 1, 1, 256);
    place(a, blob, p + ".branch1.0", 0, 0, 0.f);
    place(a, blob, p + ".branch2.0", 32, 0, 0.f);
    place(a, blob, p + ".branch0", 64, 0, 0.f);
    add_conv(net, p + ".A", a, 1, 1, 0, 0);
    HostConv g(64, 3, 3, 64);   // block diagonal: (b1.0 -> b1.1), (b2.0 -> b2.1)
    place(g, blob, p + ".branch1.1", 0, 0, 0.f);
    place(g, blob, p + ".branch2.1", 32, 32, 0.f);
    add_conv(net, p + ".G", g, 1, 1, 1, 1);
    add_plain(net, blob, p + ".branch2.2", 0.f, 1, 1, 1, 1);
    HostConv u(256, 1, 1, 96);  // input order (b0, b2.2, b1.1) vs source (b0, b1, b2); residual scale folded
    place(u, blob, p + ".conv2d", 0, 0, 0.f, 0.17f, 0, 32);
    place(u, blob, p + ".conv2d", 0, 32, 0.f, 0.17f, 64, 32);
    place(u, blob, p + ".conv2d", 0, 64, 0.f, 0.17f, 32, 32);
    add_conv(net, p + ".up", u, 1, 1, 0, 0, res_bn_cap(96));
  }
  add_plain(net, blob, "mixed_6a.branch0", 0.f, 2, 2, 0, 0);
  add_plain(net, blob, "mixed_6a.branch1.0", 0.f, 1, 1, 0, 0);
  add_plain(net, blob, "mixed_6a.branch1.1", 0.f, 1, 1, 1, 1);
  add_plain(net, blob, "mixed_6a.branch1.2", 0.f, 2, 2, 0, 0);
  auto add_block_17_8 = [&](const std::string& p, int cin, int cb, float scale, float out_slope, bool is17) {
    // buffer: [b1.0 | b0 | b1.2 | b1.1], cb channels each
    HostConv a(2 * cb, 1, 1, cin);
    place(a, blob, p + ".branch1.0", 0, 0, 0.f);
    place(a, blob, p + ".branch0", cb, 0, 0.f);
    add_conv(net, p + ".A", a, 1, 1, 0, 0);
    if (is17) {
      add_plain(net, blob, p + ".branch1.1", 0.f, 1, 1, 0, 3);   // 1x7
      add_plain(net, blob, p + ".branch1.2", 0.f, 1, 1, 3, 0);   // 7x1
    } else {
      add_plain(net, blob, p + ".branch1.1", 0.f, 1, 1, 0, 1);   // 1x3
      add_plain(net, blob, p + ".branch1.2", 0.f, 1, 1, 1, 0);   // 3x1
    }
    add_plain(net, blob, p + ".conv2d", out_slope, 1, 1, 0, 0, 0, scale, res_bn_cap(2 * cb));   // input order (b0, b1.2) == source order
    net->conv_idx[p + ".up"] = net->conv_idx[p + ".conv2d"];
  };
  for (int i = 0; i < 10; ++i) add_block_17_8(sfmt("repeat_2.%d", i), 896, 128, 0.10f, 0.f, true);
  {
    HostConv a(768, 1, 1, 896);
    place(a, blob, "mixed_7a.branch0.0", 0, 0, 0.f);
    place(a, blob, "mixed_7a.branch1.0", 256, 0, 0.f);
    place(a, blob, "mixed_7a.branch2.0", 512, 0, 0.f);
    add_conv(net, "mixed_7a.A", a, 1, 1, 0, 0);
    add_plain(net, blob, "mixed_7a.branch0.1", 0.f, 2, 2, 0, 0);
    add_plain(net, blob, "mixed_7a.branch1.1", 0.f, 2, 2, 0, 0);
    add_plain(net, blob, "mixed_7a.branch2.1", 0.f, 1, 1, 1, 1);
    add_plain(net, blob, "mixed_7a.branch2.2", 0.f, 2, 2, 0, 0);
  }
  for (int i = 0; i < 5; ++i) add_block_17_8(sfmt("repeat_3.%d", i), 1792, 192, 0.20f, 0.f, false);
  add_block_17_8("block8", 1792, 192, 1.0f, 1.f, false);
  add_plain(net, blob, "last_linear", 1.f, 1, 1, 0, 0);   // Linear(1792->512) with BatchNorm1d folded
}

// ------------------------------------------------------------------------------------ plan builder
struct Builder {
  Net* net;
  Plan* plan;
  int tensor(int H, int W, int C, int elem = 2, int ext = 0) {
    TensorInfo t;
    t.H = H; t.W = W; t.C = C; t.elem = elem; t.ext = ext;
    t.bytes = (static_cast<size_t>(plan->B) * H * W * C * elem + 255) & ~static_cast<size_t>(255);
    plan->tensors.push_back(t);
    return static_cast<int>(plan->tensors.size()) - 1;
  }
  void touch(View v) {
    if (v.t < 0) return;
    TensorInfo& t = plan->tensors[v.t];
    const int i = static_cast<int>(plan->ops.size());
    if (t.first < 0) t.first = i;
    t.last = i;
  }
  int layer(const std::string& name) {
    auto it = net->conv_idx.find(name);
    if (it == net->conv_idx.end()) throw std::runtime_error("plan: unknown layer '" + name + "'");
    return it->second;
  }
  void conv(const std::string& name, View in, View out, int res_mode = RES_NONE, View res = View(), int out_mode = OUT_16BIT,
            int P_off = 0) {
    Op op;
    op.kind = OP_CONV; op.layer = layer(name); op.in = in; op.out = out; op.res = res; op.res_mode = res_mode;
    op.out_mode = out_mode; op.P_off = P_off;
    const ConvLayer& L = net->convs[op.layer];
    const TensorInfo& ti = plan->tensors[in.t];
    if (!L.special && in.coff + L.Cin > ti.C) throw std::runtime_error("plan: layer '" + name + "' reads past its input tensor");
    if (out.t >= 0 && !L.special) {
      const TensorInfo& to = plan->tensors[out.t];
      if (to.H != conv_out_dim(ti.H, L.kh, L.sh, L.ph) || to.W != conv_out_dim(ti.W, L.kw, L.sw, L.pw) ||
          out.coff + L.Cout > to.C)
        throw std::runtime_error("plan: layer '" + name + "' output shape mismatch");
    }
    touch(in); touch(out); touch(res);
    plan->ops.push_back(op);
  }
  void conv_tail(const std::string& name3, const std::string& name1, View in, View res, View out, View out2,
                 const std::string& skip_conv = std::string()) {
    Op op;
    op.kind = OP_CONV_TAIL; op.layer = layer(name3); op.layer2 = layer(name1); op.in = in; op.res = res; op.out = out; op.out2 = out2;
    if (!skip_conv.empty()) op.layer3 = layer(skip_conv);
    op.res_mode = RES_PRE_ACT;
    touch(in); touch(res); touch(out); touch(out2);
    plan->ops.push_back(op);
  }
  void dw(int idx, View in, View out) {
    Op op;
    op.kind = OP_DW; op.layer = idx; op.in = in; op.out = out;
    touch(in); touch(out);
    plan->ops.push_back(op);
  }
  void conv_pool(const std::string& name, View in, View pooled) {      // conv whose epilogue max-pools (3x3 / 2 / pad 1) into `pooled`
    Op op;
    op.kind = OP_CONV; op.layer = layer(name); op.in = in; op.out2 = pooled;
    touch(in); touch(pooled);
    plan->ops.push_back(op);
  }
  void dwpw(int idx, const std::string& pw, View in, View out) {
    Op op;
    op.kind = OP_DWPW; op.layer = idx; op.layer2 = layer(pw); op.in = in; op.out = out;
    touch(in); touch(out);
    plan->ops.push_back(op);
  }
  void maxpool(View in, View out, int C, int pad) {
    Op op;
    op.kind = OP_MAXPOOL; op.in = in; op.out = out; op.C = C; op.pad = pad;
    touch(in); touch(out);
    plan->ops.push_back(op);
  }
  void avgpool(View in, View out) {
    Op op;
    op.kind = OP_AVGPOOL; op.in = in; op.out = out;
    touch(in); touch(out);
    plan->ops.push_back(op);
  }
  void l2norm(View io) {
    Op op;
    op.kind = OP_L2NORM; op.in = io; op.out = io;
    touch(io);
    plan->ops.push_back(op);
  }
};

static View V(int t, int coff = 0) {
  View v;
  v.t = t; v.coff = coff;
  return v;
}

static void plan_retina(Net* net, Plan& plan, bool r50) {
  Builder b{net, &plan};
  const int H = plan.H, W = plan.W;
  // network input: NHWC8 for MobileNet, padded space-to-depth [H/2, W/2 + 3, 16] for ResNet-50 (glue.cu K1)
  // (odd frame sizes: the missing row / column of the last 2x2 block is zero, which is exactly the stem's padding)
  const int x0 = r50 ? b.tensor((H + 1) / 2, (W + 1) / 2 + 3, 16, 2, 1) : b.tensor(H, W, 8, 2, 1);
  int c[3];
  if (r50) {
    const int H1 = (H + 1) / 2, W1 = (W + 1) / 2;
    int h = conv_out_dim(H1, 3, 2, 1), w = conv_out_dim(W1, 3, 2, 1);
    int x = b.tensor(h, w, 64);
    if (net->ctx->fuse_pool && conv_pool_fusable(net->ctx, net->convs[b.layer("conv1")], H1, W1 + 3)) {
      // stem conv with the max pool in its epilogue: the 64-channel stem output (4.25 GB at 64 x 1080p) never exists
      b.conv_pool("conv1", V(x0), V(x));
    } else {
      const int t1 = b.tensor(H1, W1, 64);
      b.conv("conv1", V(x0), V(t1));
      b.maxpool(V(t1), V(x), 64, 1);
    }
    int y1_fused = -1;
    for (int l = 0; l < 4; ++l) {
      const int planes = kR50Planes[l];
      for (int i = 0; i < kR50Blocks[l]; ++i) {
        const int s = (i == 0 && l > 0) ? 2 : 1;
        const std::string p = sfmt("layer%d.%d", l + 1, i);
        const int ho = conv_out_dim(h, 3, s, 1), wo = conv_out_dim(w, 3, s, 1);
        int y1 = y1_fused;                       // already written by the previous block's fused tail
        if (y1 < 0) {
          y1 = b.tensor(h, w, planes);
          b.conv(p + ".conv1", V(x), V(y1));
        }
        y1_fused = -1;
        const int y2 = b.tensor(ho, wo, planes);
        b.conv(p + ".conv2", V(y1), V(y2));
        const bool last = i + 1 == kR50Blocks[l];
        const std::string pn = last ? sfmt("layer%d.0", l + 2) : sfmt("layer%d.%d", l + 1, i + 1);
        const bool fuse = net->ctx->fuse_tail && !(last && l == 3) && !(net->ctx->fuse_tail == 2 && last) &&
                          (conv_tail_applicable(net->convs[b.layer(p + ".conv3")], net->convs[b.layer(pn + ".conv1")]) ||
                           (net->ctx->fuse_tail2 && conv_tail2_applicable(net->convs[b.layer(p + ".conv3")], net->convs[b.layer(pn + ".conv1")])));
        // FRB_OPT_FUSE_TAIL = 3: the skip conv of the layer's first block is computed inside the fused tail as well
        const bool fuse_skip = fuse && i == 0 && net->ctx->fuse_tail == 3 && net->convs[b.layer(p + ".conv3")].d_bias_skip != nullptr &&
                               conv_tail_skip_applicable(net->convs[b.layer(p + ".down")]);
        int idn = x;
        if (i == 0 && !fuse_skip) {
          idn = b.tensor(ho, wo, planes * 4);
          b.conv(p + ".down", V(x), V(idn));
        }
        const int y = b.tensor(ho, wo, planes * 4);
        // conv3 + skip together with the next block's conv1 (always 1x1 / stride 1: the stride of a layer's first block
        // sits in its 3x3) where K2t takes the pair
        if (fuse) {
          y1_fused = b.tensor(ho, wo, net->convs[b.layer(pn + ".conv1")].Cout);
          b.conv_tail(p + ".conv3", pn + ".conv1", V(y2), V(idn), V(y), V(y1_fused), fuse_skip ? p + ".down" : std::string());
        } else {
          b.conv(p + ".conv3", V(y2), V(y), RES_PRE_ACT, V(idn));
        }
        x = y; h = ho; w = wo;
      }
      if (l >= 1) c[l - 1] = x;
    }
  } else {
    int h = conv_out_dim(H, 3, 2, 1), w = conv_out_dim(W, 3, 2, 1);
    int x = b.tensor(h, w, 8);
    b.conv("body.0", V(x0), V(x));
    int cin = 8;
    for (int i = 1; i <= 13; ++i) {
      const int s = kMnetStride[i - 1];
      const int ho = conv_out_dim(h, 3, s, 1), wo = conv_out_dim(w, 3, s, 1);
      const int y = b.tensor(ho, wo, kMnetCout[i - 1]);
      // fused where it measured faster (profiles/r02_layer_times_mnet_360p_dwpw.txt): the blocks with up to 32 input channels,
      // whose pointwise GEMM (K <= 32) wastes a tcgen05 tile; from 64 channels on the two launches are level or ahead
      if (net->ctx->fuse_dwpw && dwpw_fusable(cin, kMnetCout[i - 1]) && (cin <= 32 || net->ctx->fuse_dwpw == 2)) {
        b.dwpw(i - 1, sfmt("body.%d.pw", i), V(x), V(y));
      } else {
        const int d = b.tensor(ho, wo, cin);
        b.dw(i - 1, V(x), V(d));
        b.conv(sfmt("body.%d.pw", i), V(d), V(y));
      }
      x = y; h = ho; w = wo; cin = kMnetCout[i - 1];
      if (i == 5) c[0] = x;
      if (i == 11) c[1] = x;
      if (i == 13) c[2] = x;
    }
  }
  const int C = r50 ? 256 : 64, q = C / 4;
  auto dims = [&](int t, int& h, int& w) { h = plan.tensors[t].H; w = plan.tensors[t].W; };
  int h1, w1, h2, w2, h3, w3;
  dims(c[0], h1, w1); dims(c[1], h2, w2); dims(c[2], h3, w3);
  const int o3 = b.tensor(h3, w3, C);
  b.conv("fpn.out3", V(c[2]), V(o3));
  const int m2 = b.tensor(h2, w2, C);
  b.conv("fpn.out2", V(c[1]), V(m2), RES_POST_ACT_UPSAMPLE, V(o3));
  const int p2 = b.tensor(h2, w2, C);
  b.conv("fpn.merge2", V(m2), V(p2));
  const int m1 = b.tensor(h1, w1, C);
  b.conv("fpn.out1", V(c[0]), V(m1), RES_POST_ACT_UPSAMPLE, V(p2));
  const int p1 = b.tensor(h1, w1, C);
  b.conv("fpn.merge1", V(m1), V(p1));
  const int feat[3] = {p1, p2, o3};
  plan.P_total = 2 * (h1 * w1 + h2 * w2 + h3 * w3);
  int P_off = 0;
  for (int k = 0; k < 3; ++k) {
    int h, w;
    dims(feat[k], h, w);
    const int S = b.tensor(h, w, 6 * q);
    const std::string s = sfmt("ssh%d", k + 1);
    b.conv(s + ".A", V(feat[k]), V(S, 0));          // -> [c5_1 | c3]
    b.conv(s + ".B", V(S, 0), V(S, 4 * q));         // c5_1 -> [c5 | c7_2]
    b.conv(s + ".C", V(S, 5 * q), V(S, 3 * q));     // c7_2 -> c7
    b.conv(sfmt("head%d", k + 1), V(S, q), View(), RES_NONE, View(), OUT_HEAD_SCATTER, P_off);
    P_off += 2 * h * w;
  }
}

static void plan_irv1(Net* net, Plan& plan) {
  Builder b{net, &plan};
  const int x0 = b.tensor(160, 160, 8, 2, 1);
  const int t1 = b.tensor(79, 79, 64);
  b.conv("conv2d_1a", V(x0), V(t1));
  const int t2 = b.tensor(77, 77, 64);
  b.conv("conv2d_2a", V(t1), V(t2));
  const int t3 = b.tensor(77, 77, 64);
  b.conv("conv2d_2b", V(t2), V(t3));
  const int t4 = b.tensor(38, 38, 64);
  b.maxpool(V(t3), V(t4), 64, 0);
  const int t5 = b.tensor(38, 38, 128);
  b.conv("conv2d_3b", V(t4), V(t5));
  const int t6 = b.tensor(36, 36, 192);
  b.conv("conv2d_4a", V(t5), V(t6));
  int x = b.tensor(17, 17, 256);
  b.conv("conv2d_4b", V(t6), V(x));
  for (int i = 0; i < 5; ++i) {
    const std::string p = sfmt("repeat_1.%d", i);
    const int S = b.tensor(17, 17, 192);
    b.conv(p + ".A", V(x), V(S, 0));
    b.conv(p + ".G", V(S, 0), V(S, 128));
    b.conv(p + ".branch2.2", V(S, 160), V(S, 96));
    const int y = b.tensor(17, 17, 256);
    b.conv(p + ".up", V(S, 64), V(y), RES_PRE_ACT, V(x));
    x = y;
  }
  {
    const int y = b.tensor(8, 8, 896);
    b.conv("mixed_6a.branch0", V(x), V(y, 0));
    const int u1 = b.tensor(17, 17, 192);
    b.conv("mixed_6a.branch1.0", V(x), V(u1));
    const int u2 = b.tensor(17, 17, 192);
    b.conv("mixed_6a.branch1.1", V(u1), V(u2));
    b.conv("mixed_6a.branch1.2", V(u2), V(y, 384));
    b.maxpool(V(x), V(y, 640), 256, 0);
    x = y;
  }
  auto block_17_8 = [&](const std::string& p, int hw, int cb, int cout) {
    const int S = b.tensor(hw, hw, 4 * cb);
    b.conv(p + ".A", V(x), V(S, 0));
    b.conv(p + ".branch1.1", V(S, 0), V(S, 3 * cb));
    b.conv(p + ".branch1.2", V(S, 3 * cb), V(S, 2 * cb));
    const int y = b.tensor(hw, hw, cout);
    b.conv(p + ".up", V(S, cb), V(y), RES_PRE_ACT, V(x));
    x = y;
  };
  for (int i = 0; i < 10; ++i) block_17_8(sfmt("repeat_2.%d", i), 8, 128, 896);
  {
    const int y = b.tensor(3, 3, 1792);
    const int U = b.tensor(8, 8, 768);
    b.conv("mixed_7a.A", V(x), V(U));
    b.conv("mixed_7a.branch0.1", V(U, 0), V(y, 0));
    b.conv("mixed_7a.branch1.1", V(U, 256), V(y, 384));
    const int Vt = b.tensor(8, 8, 256);
    b.conv("mixed_7a.branch2.1", V(U, 512), V(Vt));
    b.conv("mixed_7a.branch2.2", V(Vt), V(y, 640));
    b.maxpool(V(x), V(y, 896), 896, 0);
    x = y;
  }
  for (int i = 0; i < 5; ++i) block_17_8(sfmt("repeat_3.%d", i), 3, 192, 1792);
  block_17_8("block8", 3, 192, 1792);
  const int a = b.tensor(1, 1, 1792);
  b.avgpool(V(x), V(a));
  const int e = b.tensor(1, 1, 512, 4, 2);
  b.conv("last_linear", V(a), V(e), RES_NONE, View(), OUT_FP32);
  b.l2norm(V(e));
}

// Greedy interval placement: largest tensors first, each at the lowest offset that does not
// collide with an already placed tensor whose live range overlaps.
static void assign_arena(Plan& plan) {
  std::vector<int> order;
  for (size_t i = 0; i < plan.tensors.size(); ++i)
    if (plan.tensors[i].ext == 0 && plan.tensors[i].first >= 0) order.push_back(static_cast<int>(i));
  std::sort(order.begin(), order.end(), [&](int a, int b) { return plan.tensors[a].bytes > plan.tensors[b].bytes; });
  std::vector<int> placed;
  size_t total = 0;
  for (int id : order) {
    TensorInfo& t = plan.tensors[id];
    std::vector<std::pair<size_t, size_t>> busy;
    for (int pid : placed) {
      const TensorInfo& o = plan.tensors[pid];
      if (o.first <= t.last && t.first <= o.last) busy.push_back({o.off, o.off + o.bytes});
    }
    std::sort(busy.begin(), busy.end());
    size_t off = 0;
    for (auto& iv : busy) {
      if (off + t.bytes <= iv.first) break;
      off = std::max(off, iv.second);
    }
    t.off = off;
    total = std::max(total, off + t.bytes);
    placed.push_back(id);
  }
  plan.arena_bytes = total;
}

static Plan* get_plan(Net* net, int B, int H, int W) {
  const auto key = std::make_tuple(B, H, W);
  auto it = net->plans.find(key);
  if (it != net->plans.end()) return &it->second;
  Plan plan;
  plan.B = B; plan.H = H; plan.W = W;
  if (net->model == FRB_IRV1) plan_irv1(net, plan);
  else plan_retina(net, plan, net->model == FRB_RETINA_R50);
  assign_arena(plan);
  if (plan.arena_bytes > net->arena_bytes) {
    if (net->arena) {
      if (cudaDeviceSynchronize() != cudaSuccess) throw std::runtime_error("device sync failed before arena growth");
      drop_graphs(net);             // they address the old arena
      cudaFree(net->arena);
      net->arena = nullptr;
      net->arena_bytes = 0;
    }
    if (cudaMalloc(&net->arena, plan.arena_bytes) != cudaSuccess) {
      cudaGetLastError();
      throw std::runtime_error(sfmt("cudaMalloc of %.2f GB activation arena failed", plan.arena_bytes / 1e9));
    }
    net->arena_bytes = plan.arena_bytes;
  }
  auto res = net->plans.emplace(key, std::move(plan));
  return &res.first->second;
}

// Bytes of the activation arena a forward of this shape needs (the arena is allocated, and only ever grows, on the first
// forward of a shape; a caller that wants no allocation on its hot path runs one warm-up forward per shape it uses).
long long net_workspace_bytes(Net* net, int B, int H, int W) {
  try {
    Plan plan;
    plan.B = B; plan.H = H; plan.W = W;
    if (net->model == FRB_IRV1) plan_irv1(net, plan);
    else plan_retina(net, plan, net->model == FRB_RETINA_R50);
    assign_arena(plan);
    return static_cast<long long>(plan.arena_bytes);
  } catch (const std::exception&) {
    return -1;
  }
}

// ------------------------------------------------------------------------------------ execution
struct ExecPtrs {
  const void* in = nullptr;
  float* emb = nullptr;
  float* loc = nullptr;
  float* conf = nullptr;
  float* landm = nullptr;
};

static int run_plan(frb_ctx* ctx, Net* net, const Plan& plan, const ExecPtrs& ep, cudaStream_t st) {
  auto base = [&](View v) -> uint8_t* {
    const TensorInfo& t = plan.tensors[v.t];
    uint8_t* p = t.ext == 1 ? static_cast<uint8_t*>(const_cast<void*>(ep.in))
                            : (t.ext == 2 ? reinterpret_cast<uint8_t*>(ep.emb) : static_cast<uint8_t*>(net->arena) + t.off);
    return p + static_cast<size_t>(v.coff) * t.elem;
  };
  const bool timed = ctx->layer_times != 0;
  size_t ev = 0;
  auto mark = [&]() {
    if (ev >= ctx->lt_events.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ctx->lt_events.push_back(e);
    }
    cudaEventRecord(ctx->lt_events[ev++], st);
  };
  if (timed) {
    ctx->lt_info.clear();
    mark();
  }
  for (const Op& op : plan.ops) {
    int rc = FRB_OK;
    const TensorInfo& ti = plan.tensors[op.in.t];
    switch (op.kind) {
      case OP_CONV: {
        ConvIO io;
        io.in = base(op.in); io.in_ld = ti.C; io.B = plan.B; io.H = ti.H; io.W = ti.W;
        io.out_mode = op.out_mode;
        if (op.out.t >= 0) { io.out = base(op.out); io.out_ld = plan.tensors[op.out.t].C; }
        io.res_mode = op.res_mode;
        if (op.res.t >= 0) {
          const TensorInfo& tr = plan.tensors[op.res.t];
          io.res = base(op.res); io.res_ld = tr.C; io.rH = tr.H; io.rW = tr.W;
        }
        io.loc = ep.loc; io.conf = ep.conf; io.landm = ep.landm; io.P_total = plan.P_total; io.P_off = op.P_off;
        if (op.out2.t >= 0) {                      // max pool in the epilogue (the ResNet stem)
          const TensorInfo& tp = plan.tensors[op.out2.t];
          io.pool_out = base(op.out2); io.pool_ld = tp.C;
          rc = launch_pool_border_zero(ctx, io.pool_out, plan.B, tp.H, tp.W, net->convs[op.layer].Cout, tp.C, st);
          if (rc != FRB_OK) return rc;
        }
        {
          const ConvLayer& L = net->convs[op.layer];
          if (L.direct3 && io.out_mode == OUT_16BIT && io.res_mode == RES_NONE && ti.C == 8 && op.in.coff == 0)
            rc = launch_conv3x3_c3_direct(ctx, io.in, plan.B, ti.H, ti.W, L.sh, L.ph, L.d_w, L.K_pad, L.Cout, L.d_bias, L.d_slope,
                                          io.out, io.out_ld, st);
          else
            rc = conv_launch(ctx, L, io, st);
        }
        break;
      }
      case OP_CONV_TAIL: {
        const TensorInfo& tr = plan.tensors[op.res.t];
        if (conv_tail2_applicable(net->convs[op.layer], net->convs[op.layer2]))
          rc = conv_tail2_launch(ctx, net->convs[op.layer], net->convs[op.layer2], base(op.in), ti.C,
                                 static_cast<long long>(plan.B) * ti.H * ti.W, base(op.res), tr.C, base(op.out),
                                 plan.tensors[op.out.t].C, base(op.out2), plan.tensors[op.out2.t].C, st);
        else
        rc = conv_tail_launch(ctx, net->convs[op.layer], net->convs[op.layer2], base(op.in), ti.C,
                              static_cast<long long>(plan.B) * ti.H * ti.W, base(op.res), tr.C, base(op.out),
                              plan.tensors[op.out.t].C, base(op.out2), plan.tensors[op.out2.t].C,
                              op.layer3 >= 0 ? &net->convs[op.layer3] : nullptr, net->convs[op.layer].d_bias_skip, st);
        break;
      }
      case OP_DWPW: {
        const DwLayer& D = net->dws[op.layer];
        const ConvLayer& P = net->convs[op.layer2];
        rc = launch_dwpw_fused(ctx, base(op.in), plan.B, ti.H, ti.W, D.C, D.stride, D.d_w, D.d_bias, D.slope, P.d_w, P.K_pad,
                               P.Cout, P.d_bias, P.d_slope, base(op.out), plan.tensors[op.out.t].C, st);
        break;
      }
      case OP_DW: {
        const DwLayer& L = net->dws[op.layer];
        rc = launch_dwconv3x3(ctx, base(op.in), plan.B, ti.H, ti.W, L.C, L.stride, L.d_w, L.d_bias, 2, L.slope, base(op.out), st);
        break;
      }
      case OP_MAXPOOL:
        rc = launch_maxpool3x3s2(ctx, base(op.in), ti.C, plan.B, ti.H, ti.W, op.C, op.pad, base(op.out),
                                 plan.tensors[op.out.t].C, st);
        break;
      case OP_AVGPOOL:
        rc = launch_avgpool(ctx, base(op.in), plan.B, ti.H * ti.W, ti.C, base(op.out), st);
        break;
      case OP_L2NORM:
        rc = launch_l2norm(ctx, reinterpret_cast<float*>(base(op.in)), plan.B, ti.C, st);
        break;
    }
    if (rc != FRB_OK) return rc;
    if (timed) {
      double M = 0, N = 0, K = 0;
      if (op.kind == OP_CONV) {
        const ConvLayer& L = net->convs[op.layer];
        M = L.special ? static_cast<double>(plan.B) * ti.H * (ti.W - 3)
                      : static_cast<double>(plan.B) * conv_out_dim(ti.H, L.kh, L.sh, L.ph) * conv_out_dim(ti.W, L.kw, L.sw, L.pw);
        N = L.Cout;
        K = static_cast<double>(L.kh) * L.kw * L.Cin;
      } else if (op.kind == OP_DWPW) {
        const TensorInfo& to = plan.tensors[op.out.t];
        M = static_cast<double>(plan.B) * to.H * to.W;
        N = net->convs[op.layer2].Cout;
        K = net->dws[op.layer].C + 9.0 * net->dws[op.layer].C / N;    // pointwise + depthwise FMAs per output element
      } else if (op.kind == OP_CONV_TAIL) {
        // reported as one conv row: M pixels, N = both outputs, flop of both GEMMs
        const ConvLayer& L3 = net->convs[op.layer];
        const ConvLayer& L1 = net->convs[op.layer2];
        M = static_cast<double>(plan.B) * ti.H * ti.W;
        N = L3.Cout + L1.Cout;
        K = (static_cast<double>(L3.Cout) * (L3.Cin + (op.layer3 >= 0 ? net->convs[op.layer3].Cin : 0)) +
             static_cast<double>(L1.Cout) * L1.Cin) / N;
      } else {
        M = static_cast<double>(plan.B) * ti.H * ti.W;
        N = ti.C;
      }
      const double row[6] = {static_cast<double>(op.kind == OP_CONV_TAIL ? OP_CONV : (op.kind == OP_DWPW ? OP_DW : op.kind)), M, N, K, 2.0 * M * N * K, 0.0};
      ctx->lt_info.insert(ctx->lt_info.end(), row, row + 6);
      mark();
    }
  }
  return FRB_OK;
}

// ------------------------------------------------------------------------------------ graph replay of short forwards
// InceptionResnetV1 is 105 launches of 10-35 us, RetinaFace-MobileNet0.25 at 360p 56 launches: per launch the host
// encodes 3-4 tensor maps and the GPU idles between kernels.  Such a forward is captured once per plan into a CUDA graph
// (tensor maps and arguments become node parameters) and replayed; network input and outputs go through net-owned
// staging buffers so that the graph does not depend on the caller's pointers (one small D2D copy on each side).
static size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

struct IoLayout {
  size_t in_bytes = 0, out_bytes[3] = {0, 0, 0}, off[4] = {0, 0, 0, 0}, total = 0;
};

static IoLayout io_layout(const Net* net, const Plan& plan) {
  IoLayout l;
  for (const TensorInfo& t : plan.tensors)
    if (t.ext == 1) l.in_bytes = static_cast<size_t>(plan.B) * t.H * t.W * t.C * t.elem;
  if (net->model == FRB_IRV1) {
    l.out_bytes[0] = static_cast<size_t>(plan.B) * 512 * 4;
  } else {
    l.out_bytes[0] = static_cast<size_t>(plan.B) * plan.P_total * 4 * 4;
    l.out_bytes[1] = static_cast<size_t>(plan.B) * plan.P_total * 2 * 4;
    l.out_bytes[2] = static_cast<size_t>(plan.B) * plan.P_total * 10 * 4;
  }
  l.off[0] = 0;
  l.off[1] = align256(l.in_bytes);
  l.off[2] = l.off[1] + align256(l.out_bytes[0]);
  l.off[3] = l.off[2] + align256(l.out_bytes[1]);
  l.total = l.off[3] + align256(l.out_bytes[2]);
  return l;
}

static bool use_graph(frb_ctx* ctx, const Net* net, const Plan& plan) {
  if (!ctx->graphs || ctx->layer_times || ctx->trace_mode) return false;
  const IoLayout l = io_layout(net, plan);
  if (l.total > (static_cast<size_t>(1) << 30)) return false;                           // staging copies must stay cheap
  if (net->model == FRB_IRV1) return true;
  return static_cast<double>(plan.B) * plan.H * plan.W <= 64.0 * 360 * 640 && net->model == FRB_RETINA_MNET025;
}

static int run_plan(frb_ctx* ctx, Net* net, const Plan& plan, const ExecPtrs& ep, cudaStream_t st);

static int run_plan_graphed(frb_ctx* ctx, Net* net, Plan& plan, const ExecPtrs& ep, cudaStream_t st) {
  plan.calls++;
  if (!use_graph(ctx, net, plan) || plan.calls < 2) return run_plan(ctx, net, plan, ep, st);   // the first call also warms caches
  const IoLayout l = io_layout(net, plan);
  if (net->io_bytes < l.total) {
    if (net->io) {
      FRB_CUDA(ctx, cudaDeviceSynchronize());
      drop_graphs(net);
      FRB_CUDA(ctx, cudaFree(net->io));
      net->io = nullptr;
      net->io_bytes = 0;
    }
    FRB_CUDA(ctx, cudaMalloc(&net->io, l.total));
    net->io_bytes = l.total;
  }
  uint8_t* io = static_cast<uint8_t*>(net->io);
  if (plan.graph != nullptr && (plan.graph_arena != net->arena || plan.graph_io != net->io)) {
    cudaGraphExecDestroy(plan.graph);
    plan.graph = nullptr;
  }
  ExecPtrs gp;
  gp.in = io + l.off[0];
  if (net->model == FRB_IRV1) {
    gp.emb = reinterpret_cast<float*>(io + l.off[1]);
  } else {
    gp.loc = reinterpret_cast<float*>(io + l.off[1]);
    gp.conf = reinterpret_cast<float*>(io + l.off[2]);
    gp.landm = reinterpret_cast<float*>(io + l.off[3]);
  }
  if (plan.graph == nullptr) {
    const long long l0 = ctx->launches;
    cudaGraph_t g = nullptr;
    // captured on a stream of our own: the caller's may be the legacy default stream, which cannot be captured
    if (ctx->capture_stream == nullptr) FRB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->capture_stream, cudaStreamNonBlocking));
    FRB_CUDA(ctx, cudaStreamBeginCapture(ctx->capture_stream, cudaStreamCaptureModeRelaxed));
    const int rc = run_plan(ctx, net, plan, gp, ctx->capture_stream);
    const cudaError_t ce = cudaStreamEndCapture(ctx->capture_stream, &g);
    if (rc != FRB_OK) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    FRB_CUDA(ctx, ce);
    const cudaError_t ie = cudaGraphInstantiate(&plan.graph, g, 0);
    cudaGraphDestroy(g);
    FRB_CUDA(ctx, ie);
    plan.graph_launches = ctx->launches - l0;
    ctx->launches = l0;                          // counted per replay below
    plan.graph_arena = net->arena;
    plan.graph_io = net->io;
  }
  FRB_CUDA(ctx, cudaMemcpyAsync(io + l.off[0], ep.in, l.in_bytes, cudaMemcpyDeviceToDevice, st));
  FRB_CUDA(ctx, cudaGraphLaunch(plan.graph, st));
  ctx->launches += plan.graph_launches;
  if (net->model == FRB_IRV1) {
    FRB_CUDA(ctx, cudaMemcpyAsync(ep.emb, io + l.off[1], l.out_bytes[0], cudaMemcpyDeviceToDevice, st));
  } else {
    FRB_CUDA(ctx, cudaMemcpyAsync(ep.loc, io + l.off[1], l.out_bytes[0], cudaMemcpyDeviceToDevice, st));
    FRB_CUDA(ctx, cudaMemcpyAsync(ep.conf, io + l.off[2], l.out_bytes[1], cudaMemcpyDeviceToDevice, st));
    FRB_CUDA(ctx, cudaMemcpyAsync(ep.landm, io + l.off[3], l.out_bytes[2], cudaMemcpyDeviceToDevice, st));
  }
  return FRB_OK;
}

// An option that changes which ops a forward consists of (FRB_OPT_FUSE_TAIL) invalidates the cached plans.
void net_drop_plans(Net* net) {
  if (net == nullptr) return;
  cudaDeviceSynchronize();
  drop_graphs(net);
  net->plans.clear();
}

Net* build_net(frb_ctx* ctx, int model_id, const void* blob, size_t bytes) {
  Net* net = new Net();
  net->model = model_id;
  net->ctx = ctx;
  try {
    WeightBlob wb;
    parse_blob(blob, bytes, wb);
    if (model_id == FRB_IRV1) load_irv1(net, wb);
    else if (model_id == FRB_RETINA_MNET025 || model_id == FRB_RETINA_R50) load_retina(net, wb, model_id == FRB_RETINA_R50);
    else throw std::runtime_error("unknown model id");
  } catch (const std::exception& e) {
    set_err(ctx, FRB_ERR_INVALID, "frb_load_model: %s", e.what());
    destroy_net(net);
    return nullptr;
  }
  return net;
}

void destroy_net(Net* net) {
  if (net == nullptr) return;
  drop_graphs(net);
  if (net->io) cudaFree(net->io);
  std::vector<bool> seen(net->convs.size(), false);
  for (auto& L : net->convs) conv_release(L);
  for (auto& L : net->dws) {
    if (L.d_w) cudaFree(L.d_w);
    if (L.d_bias) cudaFree(L.d_bias);
  }
  if (net->arena) cudaFree(net->arena);
  delete net;
}

int net_forward_retina(frb_ctx* ctx, Net* net, const void* d_in, int B, int H, int W, float* d_loc, float* d_conf,
                       float* d_landm, cudaStream_t stream) {
  if (B <= 0 || H < 32 || W < 32) return set_err(ctx, FRB_ERR_INVALID, "retina_forward: bad shape B=%d H=%d W=%d", B, H, W);
  try {
    Plan* plan = get_plan(net, B, H, W);
    ExecPtrs ep;
    ep.in = d_in; ep.loc = d_loc; ep.conf = d_conf; ep.landm = d_landm;
    return run_plan_graphed(ctx, net, *plan, ep, stream);
  } catch (const std::exception& e) {
    return set_err(ctx, FRB_ERR_INVALID, "retina_forward: %s", e.what());
  }
}

int net_forward_irv1(frb_ctx* ctx, Net* net, const void* d_in, int F, float* d_emb, cudaStream_t stream) {
  if (F <= 0) return set_err(ctx, FRB_ERR_INVALID, "irv1_forward: F must be positive");
  try {
    Plan* plan = get_plan(net, F, 160, 160);
    ExecPtrs ep;
    ep.in = d_in; ep.emb = d_emb;
    return run_plan_graphed(ctx, net, *plan, ep, stream);
  } catch (const std::exception& e) {
    return set_err(ctx, FRB_ERR_INVALID, "irv1_forward: %s", e.what());
  }
}

}  // namespace frb
