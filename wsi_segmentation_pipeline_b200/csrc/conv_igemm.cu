// conv_igemm.cu — host side of the tcgen05 implicit-GEMM convolution (see conv_igemm.cuh).
#include "conv_igemm.cuh"
#include "conv_pair.cuh"
#include "conv_halo.cuh"
#include "conv_upstream.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "conv_rowstream.cuh"
#include "conv_rowtile.cuh"

namespace wsi {

ConvOp::ConvOp() = default;
ConvOp::~ConvOp() = default;
ConvOp::ConvOp(ConvOp&&) noexcept = default;
ConvOp& ConvOp::operator=(ConvOp&&) noexcept = default;

// ---- cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda) ----
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;

void init_tensor_map_api() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  });
  WSI_REQUIRE(g_encode != nullptr, WSI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
}

static CUtensorMapSwizzle swizzle_for(int block_k) {
  return block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (block_k == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// 4-D bf16 map: dims (d0 contiguous .. d3), byte strides for d1..d3
static void encode4(CUtensorMap* m, const void* base, const uint64_t dims[4], const uint64_t strides[3],
                    const uint32_t box[4], int block_k) {
  init_tensor_map_api();
  const cuuint32_t es[4] = {1, 1, 1, 1};
  WSI_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, WSI_ERR_INVALID, "TMA base not 16-byte aligned");
  for (int i = 0; i < 3; ++i) WSI_REQUIRE(strides[i] % 16 == 0, WSI_ERR_INVALID, "TMA stride %d = %llu not a multiple of 16", i, (unsigned long long)strides[i]);
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(block_k), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WSI_REQUIRE(r == CUDA_SUCCESS, WSI_ERR_CUDA,
              "cuTensorMapEncodeTiled(4d) failed: %d dims=(%llu,%llu,%llu,%llu) box=(%u,%u,%u,%u) strides=(%llu,%llu,%llu)", (int)r,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)dims[3],
              box[0], box[1], box[2], box[3], (unsigned long long)strides[0], (unsigned long long)strides[1], (unsigned long long)strides[2]);
}

static void encode2(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t stride1, uint32_t b0, uint32_t b1, int block_k) {
  init_tensor_map_api();
  const cuuint64_t dims[2] = {d0, d1};
  const cuuint64_t strides[1] = {stride1};
  const cuuint32_t box[2] = {b0, b1};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(block_k), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WSI_REQUIRE(r == CUDA_SUCCESS, WSI_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed: %d dims=(%llu,%llu) box=(%u,%u)", (int)r,
              (unsigned long long)d0, (unsigned long long)d1, b0, b1);
}

static int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }
static int pow2_ceil(int v) { int p = 1; while (p < v) p *= 2; return p; }

static void choose_box(int A_h, int A_w, int* bw, int* bh, int* bn) {
  *bw = std::min(16, pow2_floor(std::max(A_w, 1)));
  *bh = std::min(kBlockM / *bw, pow2_ceil(std::max(A_h, 1)));
  *bn = kBlockM / (*bw * *bh);
}

// plain NHWC view (planes: 1, or 3 channel blocks [a | b | c] of the fp32 emulation)
static void map_plain(CUtensorMap* m, const TensorView& t, int block_k, int bw, int bh, int bn, int planes = 1) {
  const uint64_t C = (uint64_t)t.C * planes;
  const uint64_t dims[4] = {C, (uint64_t)t.W, (uint64_t)t.H, (uint64_t)t.N};
  const uint64_t st[3] = {C * 2, (uint64_t)t.W * C * 2, (uint64_t)t.H * t.W * C * 2};
  const uint32_t box[4] = {(uint32_t)block_k, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
  encode4(m, t.ptr, dims, st, box, block_k);
}
// (hp, wp) parity view: element (a, b) = t[2a+hp, 2b+wp]
static void map_parity(CUtensorMap* m, const TensorView& t, int hp, int wp, int block_k, int bw, int bh, int bn, int planes = 1) {
  const uint64_t C = (uint64_t)t.C * planes;
  const char* base = static_cast<const char*>(t.ptr) + ((size_t)hp * t.W + wp) * C * 2;
  const uint64_t dims[4] = {C, (uint64_t)((t.W - wp + 1) / 2), (uint64_t)((t.H - hp + 1) / 2), (uint64_t)t.N};
  const uint64_t st[3] = {C * 4, (uint64_t)t.W * C * 4, (uint64_t)t.H * t.W * C * 2};
  const uint32_t box[4] = {(uint32_t)block_k, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
  encode4(m, base, dims, st, box, block_k);
}

// fp32 -> three bf16 planes a + b + c == v exactly (8 + 8 + 8 mantissa bits)
static inline void split3(float v, uint16_t out[3]) {
  float r = v;
  for (int j = 0; j < 3; ++j) {
    out[j] = f32_to_bf16_bits(r);
    r -= bf16_bits_to_f32(out[j]);
  }
}
static const int kSplitBHost[6] = {2, 0, 1, 1, 0, 0};   // weight plane of product pr (kSplitB in conv_igemm.cuh)

static inline int floor_div2(int t) { return (t >= 0) ? t / 2 : -((-t + 1) / 2); }

bool ConvOp::routes_to_rowtile(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const void* residual) {
  return RowConvOp::eligible(parts, spec, residual) && getenv("WSI_NO_ROWTILE") == nullptr;
}

bool ConvOp::routes_to_upstream(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const void* residual, int out_layout) {
  return routes_to_rowtile(parts, spec, residual) && UpStreamOp::eligible(parts, spec, residual, out_layout) &&
         getenv("WSI_NO_UPSTREAM") == nullptr;
}

void ConvOp::build(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const float* w_oihw,
                   const float* scale, const float* bias, const void* residual, void* out,
                   const float* head_w, const float* head_b, float* head_out, int* error_flag, int num_sms, int out_layout,
                   int res_layout, int split) {
  WSI_REQUIRE(!parts.empty() && parts.size() <= 2, WSI_ERR_INVALID, "conv: 1 or 2 input parts");
  stream_.reset();
  upstream_.reset();
  split_ = split ? 1 : 0;
  const int planes = split_ ? 3 : 1, nprod = split_ ? 6 : 1;
  if (split_) {
    // fp32 emulation: always the table-driven TMA kernel on NHWC [a | b | c] tensors
    WSI_REQUIRE(out_layout == LAYOUT_NHWC && res_layout == LAYOUT_NHWC, WSI_ERR_UNSUPPORTED, "conv (fp32 emulation): NHWC only");
    row_.reset();
    stem_.reset();
    halo_ = false;
  }
  if (!split_ && head_out == nullptr && routes_to_upstream(parts, spec, residual, out_layout)) {
    upstream_.reset(new UpStreamOp());
    upstream_->build(parts, spec, w_oihw, scale, bias, out, error_flag, num_sms);
    flops_ = upstream_->flops();
    block_n_ = spec.cout;
    block_k_ = 16;
    row_.reset();
    stem_.reset();
    return;
  }
  if (!split_ && routes_to_rowtile(parts, spec, residual) && RowStreamOp::eligible(parts, spec) && getenv("WSI_NO_ROWSTREAM") == nullptr) {
    stream_.reset(new RowStreamOp());
    stream_->build(parts[0], spec, w_oihw, scale, bias, residual, res_layout, out, out_layout, head_w, head_b, head_out, error_flag, num_sms);
    flops_ = stream_->flops();
    block_n_ = spec.cout;
    block_k_ = 16;
    row_.reset();
    stem_.reset();
    return;
  }
  if (!split_ && routes_to_rowtile(parts, spec, residual)) {
    row_.reset(new RowConvOp());
    row_->build(parts, spec, w_oihw, scale, bias, residual, res_layout, out, out_layout, head_w, head_b, head_out, error_flag, num_sms);
    flops_ = row_->flops();
    block_n_ = spec.cout;
    block_k_ = 16;
    stem_.reset();
    return;
  }
  row_.reset();
  stem_.reset();
  WSI_REQUIRE((out_layout == LAYOUT_NHWC || (out_layout == LAYOUT_PLANAR && head_out == nullptr)) &&
                  (residual == nullptr || res_layout == LAYOUT_NHWC),
              WSI_ERR_UNSUPPORTED, "conv: the TMA kernel reads NHWC tensors and writes NHWC or plain planar ones");
  out_planar_ = (out_layout == LAYOUT_PLANAR);
  for (auto& q : parts) WSI_REQUIRE(q.t.layout == LAYOUT_NHWC, WSI_ERR_UNSUPPORTED, "conv: the TMA kernel reads NHWC operands");
  halo_ = false;
  if (!split_ && halo_eligible(parts, spec, out_layout, head_out != nullptr, num_sms)) {
    build_halo(parts, spec, w_oihw, scale, bias, residual, out, out_layout, error_flag, num_sms);
    return;
  }
  bool any_up = false;
  for (auto& q : parts) any_up |= q.up2;
  const int k = spec.ksize;
  int Hin = 0, Win = 0, N = parts[0].t.N, cin = 0;
  for (auto& q : parts) {
    const int h = q.up2 ? 2 * q.t.H : q.t.H, w = q.up2 ? 2 * q.t.W : q.t.W;
    if (Hin == 0) { Hin = h; Win = w; }
    WSI_REQUIRE(h == Hin && w == Win && q.t.N == N, WSI_ERR_INVALID, "conv: input parts disagree on shape");
    WSI_REQUIRE(q.t.C % 16 == 0, WSI_ERR_UNSUPPORTED, "conv: channels (%d) must be a multiple of 16", q.t.C);
    cin += q.t.C;
  }
  int bk = 64;
  for (auto& q : parts) while (q.t.C % bk) bk /= 2;
  const int OH = (Hin + 2 * spec.pad - k) / spec.stride + 1, OW = (Win + 2 * spec.pad - k) / spec.stride + 1;
  WSI_REQUIRE(spec.cout % 16 == 0 && spec.cout <= 512, WSI_ERR_UNSUPPORTED, "conv: cout (%d) must be a multiple of 16, at most 512", spec.cout);
  // Output-channel tile.  The MMA rate is bound by the shared-memory operand fetch (~50 B/clk measured), so a
  // 128x256 tile (12 KB of operands per 128 math cycles) beats 128x128 (8 KB per 64) whenever Cout allows it
  // and there are still enough tiles for every SM.
  int bn_out = 128;
  while (spec.cout % bn_out) bn_out /= 2;
  if (!split_ && spec.cout % 256 == 0 && bk == 64 && getenv("WSI_NO_BN256") == nullptr) {
    // a 128x256 tile does twice the math of a 128x128 one in ~1.6x the time (operand bytes 48 KB vs 32 KB per
    // K block); take it unless wave quantisation over the SMs eats that
    const long long m_tiles = ceil_div((long long)N * OH * OW, kBlockM);
    const long long w128 = ceil_div(m_tiles * (spec.cout / 128), num_sms), w256 = ceil_div(m_tiles * (spec.cout / 256), num_sms);
    if (w256 * 16 <= w128 * 10) bn_out = 256;
  }
  block_n_ = bn_out;
  block_k_ = bk;

  ConvParams& p = p_;
  p = ConvParams{};
  p.N = N; p.OH = OH; p.OW = OW; p.Cout = spec.cout;
  p.relu = spec.relu ? 1 : 0;
  p.res = static_cast<const bf16*>(residual);
  p.out = static_cast<bf16*>(out);
  p.error_flag = error_flag;
  if (out_planar_) {
    WSI_REQUIRE(spec.cout % 8 == 0, WSI_ERR_UNSUPPORTED, "conv: planar output needs Cout %% 8 == 0");
    const PlanarDims od = PlanarDims::make(OH, OW, spec.cout, LAYOUT_PLANAR);
    p.out_planar = 1;
    p.pl_chunk = (long long)od.Wrow * 16;
    p.pl_row = (long long)od.KC * p.pl_chunk;
    p.pl_img = (long long)(OH + 2) * p.pl_row;
  }

  std::vector<KBlock> table;
  int num_parity = 1;
  if (any_up) {
    WSI_REQUIRE(k == 3 && spec.stride == 1 && spec.pad == 1, WSI_ERR_UNSUPPORTED, "up2 conv must be 3x3/s1/p1");
    WSI_REQUIRE(parts[0].up2 && (parts.size() == 1 || !parts[1].up2), WSI_ERR_UNSUPPORTED, "up2 conv: parts = [up2 x, skip]");
    p.sigma = 2; num_parity = 4;
    p.A_h = OH / 2; p.A_w = OW / 2;
  } else {
    WSI_REQUIRE(spec.stride == 1 || (spec.stride == 2 && parts.size() == 1), WSI_ERR_UNSUPPORTED, "conv: stride %d with %zu parts", spec.stride, parts.size());
    p.sigma = 1;
    p.A_h = OH; p.A_w = OW;
  }
  choose_box(p.A_h, p.A_w, &p.bw, &p.bh, &p.bn);

  // tensor maps
  if (any_up) {
    map_plain(&amaps_.m[0], parts[0].t, bk, p.bw, p.bh, p.bn, planes);
    p.a_plane[0] = parts[0].t.C;
    if (parts.size() == 2)
      for (int hp = 0; hp < 2; ++hp)
        for (int wp = 0; wp < 2; ++wp) {
          map_parity(&amaps_.m[1 + hp * 2 + wp], parts[1].t, hp, wp, bk, p.bw, p.bh, p.bn, planes);
          p.a_plane[1 + hp * 2 + wp] = parts[1].t.C;
        }
  } else if (spec.stride == 1) {
    for (size_t i = 0; i < parts.size(); ++i) { map_plain(&amaps_.m[i], parts[i].t, bk, p.bw, p.bh, p.bn, planes); p.a_plane[i] = parts[i].t.C; }
  } else {
    for (int hp = 0; hp < 2; ++hp)
      for (int wp = 0; wp < 2; ++wp) {
        map_parity(&amaps_.m[hp * 2 + wp], parts[0].t, hp, wp, bk, p.bw, p.bh, p.bn, planes);
        p.a_plane[hp * 2 + wp] = parts[0].t.C;
      }
  }
  for (int i = 0; i < kMaxAMaps; ++i)  // unused slots alias map 0 so every descriptor is valid
    if (i >= (any_up ? (parts.size() == 2 ? 5 : 1) : (spec.stride == 1 ? (int)parts.size() : 4))) amaps_.m[i] = amaps_.m[0];

  // K-block table and packed weights.  Plain / strided convs: one table, one weight matrix.  x2-upsample
  // convs: per parity class, the upsampled operand contributes its 4 collapsed positions (weights = fp32 sum
  // of the taps that read that half-resolution pixel), the skip operand its 9 taps; weights are packed per parity.
  const int taps = k * k;
  int kb_per_par = 0;
  for (auto& q : parts) kb_per_par += (q.t.C / bk) * ((any_up && q.up2) ? 4 : taps);
  const int num_kb = kb_per_par;
  WSI_REQUIRE(num_kb <= 128, WSI_ERR_UNSUPPORTED, "conv: %d K blocks > 128", num_kb);
  const int K = num_kb * nprod * bk;               // per parity class; fp32 emulation: 6 plane products per K block
  const int wsets = any_up ? 4 : 1;
  // weight element (co, K block kbi, j) = v: bf16(v), or its three planes laid out per product
  auto put_w = [&](std::vector<uint16_t>& wp_, size_t row_base, int kbi, int j, float v) {
    if (!split_) { wp_[row_base + (size_t)kbi * bk + j] = f32_to_bf16_bits(v); return; }
    uint16_t pl[3];
    split3(v, pl);
    for (int pr = 0; pr < 6; ++pr) wp_[row_base + ((size_t)kbi * 6 + pr) * bk + j] = pl[kSplitBHost[pr]];
  };
  std::vector<uint16_t> wp((size_t)spec.cout * K * wsets);
  const size_t Ktot = (size_t)K * wsets;          // row length of the packed [Cout][wsets*K] matrix
  auto w_at = [&](int co, int ci, int r, int s) { return w_oihw[(((size_t)co * cin + ci) * k + r) * k + s]; };
  for (int par = 0; par < num_parity; ++par) {
    const int py = par >> 1, px = par & 1;
    int kbi = 0;
    int part_off = 0;
    for (size_t pi = 0; pi < parts.size(); ++pi) {
      const auto& q = parts[pi];
      if (any_up && q.up2) {
        for (int pos = 0; pos < 4; ++pos)
          for (int c0 = 0; c0 < q.t.C; c0 += bk) {
            KBlock e{};
            e.c0 = c0; e.map = 0;
            e.da = (int8_t)(floor_div2(py - 1) + (pos >> 1));
            e.db = (int8_t)(floor_div2(px - 1) + (pos & 1));
            for (int co = 0; co < spec.cout; ++co)
              for (int j = 0; j < bk; ++j) {
                float v = 0.f;
                for (int r = 0; r < 3; ++r)
                  for (int s2 = 0; s2 < 3; ++s2) {
                    const int dy = floor_div2(py + r - 1) - floor_div2(py - 1), dx = floor_div2(px + s2 - 1) - floor_div2(px - 1);
                    if (dy * 2 + dx == pos) v += w_at(co, part_off + c0 + j, r, s2);
                  }
                put_w(wp, (size_t)co * Ktot + (size_t)par * K, kbi, j, v);
              }
            table.push_back(e);
            ++kbi;
          }
      } else {
        for (int r = 0; r < k; ++r)
          for (int s2 = 0; s2 < k; ++s2)
            for (int c0 = 0; c0 < q.t.C; c0 += bk) {
              KBlock e{};
              e.c0 = c0;
              if (any_up) {
                const int ty = py + r - 1, tx = px + s2 - 1;
                const int hp = ty & 1, wpp = tx & 1;
                e.map = (int8_t)(1 + hp * 2 + wpp); e.da = (int8_t)((ty - hp) / 2); e.db = (int8_t)((tx - wpp) / 2);
              } else if (spec.stride == 1) {
                e.map = (int8_t)pi; e.da = (int8_t)(r - spec.pad); e.db = (int8_t)(s2 - spec.pad);
              } else {
                const int ty = r - spec.pad, tx = s2 - spec.pad;
                const int hp = ty & 1, wpp = tx & 1;
                e.map = (int8_t)(hp * 2 + wpp); e.da = (int8_t)((ty - hp) / 2); e.db = (int8_t)((tx - wpp) / 2);
              }
              for (int co = 0; co < spec.cout; ++co)
                for (int j = 0; j < bk; ++j)
                  put_w(wp, (size_t)co * Ktot + (size_t)par * K, kbi, j, w_at(co, part_off + c0 + j, r, s2));
              table.push_back(e);
              ++kbi;
            }
      }
      part_off += q.t.C;
    }
  }
  p.b_parity_stride = any_up ? K : 0;
  p.num_kb = num_kb;
  flops_ = 2.0 * N * OH * OW * (double)spec.cout * cin * taps;

  if (spec.head) {
    WSI_REQUIRE(spec.cout == 16 && head_w && head_b && head_out, WSI_ERR_INVALID, "fused head needs cout == 16");
    std::vector<float> hw(head_w, head_w + 64), hb(head_b, head_b + 4);
    upload(headw_, hw); upload(headb_, hb);
    p.head_w = headw_.as<float>(); p.head_b = headb_.as<float>(); p.head_out = head_out;
    flops_ += 2.0 * N * OH * OW * 16 * 4;
  }
  finish(table, num_parity, wp, K * wsets, scale, bias, num_sms);
}

void ConvOp::set_head_out(float* ptr) {
  WSI_REQUIRE(ptr != nullptr, WSI_ERR_INVALID, "set_head_out: NULL");
  if (stream_) { stream_->set_head_out(ptr); return; }
  if (row_) { row_->set_head_out(ptr); return; }
  WSI_REQUIRE(p_.head_out != nullptr, WSI_ERR_INVALID, "set_head_out: this conv has no fused head");
  p_.head_out = ptr;
}

std::string ConvOp::kernel_name() const {
  char b[96];
  if (stream_) return stream_->kernel_name();
  if (upstream_) { snprintf(b, sizeof(b), "conv_upstream_kernel<%d>", block_n_); return b; }
  if (row_) { snprintf(b, sizeof(b), "conv_rowtile_kernel<%d>", block_n_); return b; }
  if (stem_) return "stem_rowtile_kernel";
  if (halo_) { snprintf(b, sizeof(b), "conv_halo_pair_kernel<%d,%s>", block_n_, (p_.halo_plain && !p_.out_planar) ? "plain" : "table"); return b; }
  if (pair_) { snprintf(b, sizeof(b), "conv_igemm_pair_kernel<%d>", block_n_); return b; }
  snprintf(b, sizeof(b), "conv_igemm_kernel<%d,%d%s%s>", block_n_, block_k_, resb_ ? ",RESB" : "", split_ ? ",NSPLIT=3" : "");
  return b;
}

bool ConvOp::stem_routes_to_rowtile() { return getenv("WSI_NO_ROWTILE") == nullptr; }

void ConvOp::build_stem(const void* padded_tiles, int n, int ph, int pw, const float* w_oihw, const float* scale,
                        const float* bias, void* out, int* error_flag, int num_sms, int out_layout, int split) {
  WSI_REQUIRE(ph % 2 == 0 && pw % 2 == 0, WSI_ERR_UNSUPPORTED, "stem: tile size must be even");
  row_.reset();
  stream_.reset();
  split_ = split ? 1 : 0;
  if (!split_ && stem_routes_to_rowtile()) {
    stem_.reset(new RowStemOp());
    stem_->build(padded_tiles, n, ph, pw, w_oihw, scale, bias, out, out_layout, error_flag, num_sms);
    flops_ = stem_->flops();
    block_n_ = 64; block_k_ = 16;
    return;
  }
  stem_.reset();
  WSI_REQUIRE(out_layout == LAYOUT_NHWC, WSI_ERR_UNSUPPORTED, "stem: only the row-tile stem writes the planar layout");
  block_n_ = 64; block_k_ = 32;
  ConvParams& p = p_;
  p = ConvParams{};
  p.N = n; p.OH = ph / 2; p.OW = pw / 2; p.Cout = 64;
  p.sigma = 1; p.A_h = p.OH; p.A_w = p.OW;
  p.relu = 1; p.out = static_cast<bf16*>(out); p.error_flag = error_flag;
  choose_box(p.A_h, p.A_w, &p.bw, &p.bh, &p.bn);
  const uint64_t pitch = (uint64_t)(pw + 8) * 8, tile_bytes = (uint64_t)(ph + 6) * pitch;
  // fp32 emulation: the gather writes three padded-tile buffers (planes a, b, c) back to back; plane j = maps 2j, 2j+1
  const int planes = split_ ? 3 : 1, nprod = split_ ? 6 : 1;
  for (int pl = 0; pl < planes; ++pl)
    for (int hp = 0; hp < 2; ++hp) {
      // overlapping-window view: dim0 = 32 elements (8 px x 4 ch) starting every 2 px (16 B)
      const char* base = static_cast<const char*>(padded_tiles) + (size_t)pl * n * tile_bytes + hp * pitch;
      const uint64_t dims[4] = {32, (uint64_t)p.OW, (uint64_t)((ph + 6) / 2), (uint64_t)n};
      const uint64_t st[3] = {16, 2 * pitch, tile_bytes};
      const uint32_t box[4] = {32, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
      encode4(&amaps_.m[pl * 2 + hp], base, dims, st, box, 32);
    }
  for (int i = 2 * planes; i < kMaxAMaps; ++i) amaps_.m[i] = amaps_.m[0];
  p.split_map_step = split_ ? 2 : 0;
  std::vector<KBlock> table;
  const int K = 7 * 32 * nprod;
  std::vector<uint16_t> wp((size_t)64 * K, 0);
  for (int r = 0; r < 7; ++r) {
    KBlock e{};
    e.map = (int8_t)(r & 1); e.da = (int8_t)(r >> 1); e.db = 0; e.c0 = 0;
    table.push_back(e);
    for (int co = 0; co < 64; ++co)
      for (int s = 0; s < 7; ++s)
        for (int c = 0; c < 3; ++c) {
          const float v = w_oihw[(((size_t)co * 3 + c) * 7 + r) * 7 + s];
          if (!split_) { wp[(size_t)co * K + r * 32 + s * 4 + c] = f32_to_bf16_bits(v); continue; }
          uint16_t pl3[3];
          split3(v, pl3);
          for (int pr = 0; pr < 6; ++pr) wp[(size_t)co * K + (r * 6 + pr) * 32 + s * 4 + c] = pl3[kSplitBHost[pr]];
        }
  }
  p.num_kb = 7;
  flops_ = 2.0 * n * p.OH * p.OW * 64.0 * 147.0;
  finish(table, 1, wp, K, scale, bias, num_sms);
}

void ConvOp::finish(const std::vector<KBlock>& table, int num_parity, const std::vector<uint16_t>& wpacked, int K,
                    const float* scale, const float* bias, int num_sms) {
  ConvParams& p = p_;
  p.num_parity = num_parity;
  upload(w_, wpacked);
  std::vector<float> sc(p.Cout, 1.f), bi(p.Cout, 0.f);
  if (scale) sc.assign(scale, scale + p.Cout);
  if (bias) bi.assign(bias, bias + p.Cout);
  upload(scale_, sc); upload(bias_, bi);
  upload(tbl_, table);
  p.scale = scale_.as<float>(); p.bias = bias_.as<float>(); p.kblocks = tbl_.as<KBlock>();
  p.tiles_w = (int)ceil_div(p.A_w, p.bw);
  p.tiles_h = (int)ceil_div(p.A_h, p.bh);
  p.tiles_n = (int)ceil_div(p.N, p.bn);
  p.tiles_co = p.Cout / block_n_;
  const long long tiles_m = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
  long long total = tiles_m * p.tiles_co * num_parity;
  WSI_REQUIRE(total < (1LL << 31), WSI_ERR_UNSUPPORTED, "conv: too many tiles");
  grid_ = (int)std::min<long long>(total, num_sms);
  {
    const int bbytes = (block_n_ * block_k_ * 2 + 1023) / 1024 * 1024;
    resb_ = !split_ && (p.tiles_co == 1) && (block_k_ == 64) && (block_n_ == 64 || block_n_ == 128) &&
            ((long long)p.num_kb * bbytes <= kResidentBBytes) && p.b_parity_stride == 0 && getenv("WSI_NO_RESB") == nullptr;
  }
  // CTA pairs (conv_pair.cuh): a 128 x N x 16 MMA costs ~64 + N/2 cycles on one CTA and ~64 + N/4 per CTA of a pair
  // (operand-fetch bound, measured); take the pair kernel when that beats the single-CTA schedule after wave
  // quantisation (pairs run on num_sms / 2 SM pairs)
  pair_ = false;
  if (!split_ && block_k_ == 64 && p.Cout % 128 == 0 && !resb_ && p.head_out == nullptr && !p.out_planar && num_sms >= 2 && getenv("WSI_NO_PAIR") == nullptr) {
    const int bnp = (p.Cout % 256 == 0) ? 256 : 128;
    const long long pair_tiles = (tiles_m + 1) / 2 * (p.Cout / bnp) * num_parity;
    const long long cost_single = ceil_div(total, num_sms) * (64 + block_n_ / 2);
    const long long cost_pair = ceil_div(pair_tiles, num_sms / 2) * (64 + bnp / 4);
    if (cost_pair < cost_single) {
      pair_ = true;
      block_n_ = bnp;
      p.tiles_co = p.Cout / bnp;
      grid_ = 2 * (int)std::min<long long>(pair_tiles, num_sms / 2);
    }
  }
  // B tile of one CTA: block_n_ weight rows, or half of them in the pair kernel
  encode2(&bmap_, w_.p, (uint64_t)K, (uint64_t)p.Cout, (uint64_t)K * 2, (uint32_t)block_k_, (uint32_t)(pair_ ? block_n_ / 2 : block_n_), block_k_);
  CUDA_CHECK(cudaStreamSynchronize(0));   // uploads above used the default stream
}

template <int BN, int BK, bool RESB = false, int NSPLIT = 1>
static void launch_inst(const AMaps& am, const CUtensorMap& bm, const ConvParams& p, int grid, cudaStream_t s) {
  using S = ConvSmem<BN, BK, RESB>;
  static_assert(S::kTotal <= 227 * 1024, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_kernel<BN, BK, RESB, NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  conv_igemm_kernel<BN, BK, RESB, NSPLIT><<<grid, kNumThreads, S::kTotal, s>>>(am, bm, p);
  CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// halo-resident pair kernel (conv_halo.cuh): 3x3 / s1 convs on >= 64-channel NHWC tensors, Cout % 128 == 0
// ---------------------------------------------------------------------------------------------
bool ConvOp::halo_eligible(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, int out_layout, bool head, int num_sms) {
  if (getenv("WSI_NO_HALO") != nullptr || num_sms < 2 || head) return false;
  if (out_layout != LAYOUT_NHWC && out_layout != LAYOUT_PLANAR) return false;
  if (spec.ksize != 3 || spec.pad != 1 || parts.empty() || parts.size() > 2) return false;
  if (spec.cout % 128 != 0 && spec.cout != 64) return false;
  for (auto& q : parts)
    if (q.t.layout != LAYOUT_NHWC || q.t.C % 64 != 0 || q.t.C < 64) return false;
  int A_h, A_w;                                       // output lattice of one parity class
  if (parts[0].up2) {
    if (spec.stride != 1 || (parts.size() == 2 && parts[1].up2)) return false;
    // measured (same box A/B): 5-7 % faster than the TMA pair kernel once 4+ halo tiles are in flight (with two they
    // were 3-10 % slower: a 4-tap group is ~1 500 cycles of MMAs, less than the TMA latency)
    if (getenv("WSI_NO_HALO_UP2") != nullptr) return false;
    A_h = parts[0].t.H; A_w = parts[0].t.W;
  } else {
    if (parts.size() != 1) return false;
    if (spec.stride == 2) {
      if (getenv("WSI_HALO_S2") == nullptr) return false;      // 1-4 taps per parity-plane halo: no gain (+-2 %), opt-in
      A_h = (parts[0].t.H - 1) / 2 + 1; A_w = (parts[0].t.W - 1) / 2 + 1;
    } else if (spec.stride == 1) {
      A_h = parts[0].t.H; A_w = parts[0].t.W;
    } else {
      return false;
    }
  }
  // 16 x 8 lattice tiles: smaller maps waste most of the tile, the TMA kernel keeps those
  return A_h >= kHaloTileH && A_w >= kHaloTileW;
}

void ConvOp::build_halo(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const float* w_oihw, const float* scale, const float* bias,
                        const void* residual, void* out, int out_layout, int* error_flag, int num_sms) {
  const bool up2 = parts[0].up2, s2 = (spec.stride == 2);
  const int N = parts[0].t.N, cout = spec.cout;
  int cin = 0;
  for (auto& q : parts) cin += q.t.C;
  const int Hin = up2 ? 2 * parts[0].t.H : parts[0].t.H, Win = up2 ? 2 * parts[0].t.W : parts[0].t.W;
  const int OH = s2 ? (Hin - 1) / 2 + 1 : Hin, OW = s2 ? (Win - 1) / 2 + 1 : Win;
  halo_ = true;
  pair_ = false; resb_ = false;
  out_planar_ = (out_layout == LAYOUT_PLANAR);
  block_n_ = (cout % 256 == 0) ? 256 : (cout % 128 == 0 ? 128 : 64);
  block_k_ = 64;
  ConvParams& p = p_;
  p = ConvParams{};
  p.N = N; p.OH = OH; p.OW = OW; p.Cout = cout;
  p.sigma = up2 ? 2 : 1;
  p.num_parity = up2 ? 4 : 1;
  p.A_h = up2 ? OH / 2 : OH; p.A_w = up2 ? OW / 2 : OW;
  p.bw = kHaloTileW; p.bh = kHaloTileH; p.bn = 1;
  p.tiles_w = (int)ceil_div(p.A_w, kHaloTileW);
  p.tiles_h = (int)ceil_div(p.A_h, kHaloTileH);
  p.tiles_n = N;
  p.tiles_co = cout / block_n_;
  p.relu = spec.relu ? 1 : 0;
  p.res = static_cast<const bf16*>(residual);
  p.out = static_cast<bf16*>(out);
  p.error_flag = error_flag;
  if (out_planar_) {
    const PlanarDims od = PlanarDims::make(OH, OW, cout, LAYOUT_PLANAR);
    p.out_planar = 1;
    p.pl_chunk = (long long)od.Wrow * 16;
    p.pl_row = (long long)od.KC * p.pl_chunk;
    p.pl_img = (long long)(OH + 2) * p.pl_row;
  }
  // tensor maps: every halo load is a box {64 ch, 10 px, 18 rows, 1 image}
  if (up2) {
    map_plain(&amaps_.m[0], parts[0].t, 64, kHaloW, kHaloH, 1);
    if (parts.size() == 2)
      for (int hp = 0; hp < 2; ++hp)
        for (int wp = 0; wp < 2; ++wp) map_parity(&amaps_.m[1 + hp * 2 + wp], parts[1].t, hp, wp, 64, kHaloW, kHaloH, 1);
  } else if (s2) {
    for (int hp = 0; hp < 2; ++hp)
      for (int wp = 0; wp < 2; ++wp) map_parity(&amaps_.m[hp * 2 + wp], parts[0].t, hp, wp, 64, kHaloW, kHaloH, 1);
  } else {
    map_plain(&amaps_.m[0], parts[0].t, 64, kHaloW, kHaloH, 1);
  }
  for (int i = 0; i < kMaxAMaps; ++i)
    if (i >= (up2 ? (parts.size() == 2 ? 5 : 1) : (s2 ? 4 : 1))) amaps_.m[i] = amaps_.m[0];

  // groups and weights.  Weight matrix [Cout][num_parity * Kpar], per parity class the (group, tap) blocks of 64 in
  // table order; taps of a x2-upsampled operand are pre-summed in fp32 (as in the TMA kernel).
  auto w_at = [&](int co, int ci, int r, int s) { return w_oihw[(((size_t)co * cin + ci) * 3 + r) * 3 + s]; };
  std::vector<HaloGroup> table;
  std::vector<std::vector<float>> wblocks;          // per parity: [block][co][64] fp32, flattened
  int groups_per_parity = -1;
  size_t Kpar = 0;
  for (int par = 0; par < p.num_parity; ++par) {
    const int py = par >> 1, px = par & 1;
    std::vector<float> wb;
    int ng = 0, wk = 0;
    auto add_block = [&](auto&& weight_of /* (co, j) -> float */) {
      const size_t base = wb.size();
      wb.resize(base + (size_t)cout * 64);
      for (int co = 0; co < cout; ++co)
        for (int j = 0; j < 64; ++j) wb[base + (size_t)co * 64 + j] = weight_of(co, j);
    };
    int part_off = 0;
    for (size_t pi = 0; pi < parts.size(); ++pi) {
      const auto& q = parts[pi];
      for (int c0 = 0; c0 < q.t.C; c0 += 64) {
        if (up2 && q.up2) {
          // upsampled operand: the 3x3 taps of this parity class fall on 2 x 2 half-resolution pixels
          HaloGroup g{};
          g.map = 0; g.c0 = c0; g.wk = wk;
          g.oy = (int8_t)floor_div2(py - 1); g.ox = (int8_t)floor_div2(px - 1);
          for (int pos = 0; pos < 4; ++pos) {
            g.tap_off[g.ntaps++] = (uint8_t)((pos >> 1) * kHaloW + (pos & 1));
            add_block([&](int co, int j) {
              float v = 0.f;
              for (int r = 0; r < 3; ++r)
                for (int s2_ = 0; s2_ < 3; ++s2_) {
                  const int dy = floor_div2(py + r - 1) - floor_div2(py - 1), dx = floor_div2(px + s2_ - 1) - floor_div2(px - 1);
                  if (dy * 2 + dx == pos) v += w_at(co, part_off + c0 + j, r, s2_);
                }
              return v;
            });
            wk += 64;
          }
          table.push_back(g); ++ng;
        } else if (up2 || s2) {
          // full-resolution operand seen through its (row, col) parity planes: source row = 2 * (a + ry) + hp
          const int qy = up2 ? py : 0, qx = up2 ? px : 0;      // stride 2: output (a, b) reads rows 2a + r - 1
          for (int hp = 0; hp < 2; ++hp)
            for (int wp = 0; wp < 2; ++wp) {
              HaloGroup g{};
              g.map = (int8_t)((up2 ? 1 : 0) + hp * 2 + wp); g.c0 = c0; g.wk = wk;
              int min_ry = 9, min_cx = 9;
              for (int r = 0; r < 3; ++r) if (((qy + r - 1) & 1) == hp) min_ry = std::min(min_ry, floor_div2(qy + r - 1));
              for (int s_ = 0; s_ < 3; ++s_) if (((qx + s_ - 1) & 1) == wp) min_cx = std::min(min_cx, floor_div2(qx + s_ - 1));
              if (min_ry == 9 || min_cx == 9) continue;
              g.oy = (int8_t)min_ry; g.ox = (int8_t)min_cx;
              for (int r = 0; r < 3; ++r)
                for (int s_ = 0; s_ < 3; ++s_) {
                  if (((qy + r - 1) & 1) != hp || ((qx + s_ - 1) & 1) != wp) continue;
                  g.tap_off[g.ntaps++] = (uint8_t)((floor_div2(qy + r - 1) - min_ry) * kHaloW + (floor_div2(qx + s_ - 1) - min_cx));
                  add_block([&](int co, int j) { return w_at(co, part_off + c0 + j, r, s_); });
                  wk += 64;
                }
              table.push_back(g); ++ng;
            }
        } else {
          HaloGroup g{};
          g.map = 0; g.c0 = c0; g.wk = wk; g.oy = -1; g.ox = -1;
          for (int tap = 0; tap < 9; ++tap) {
            g.tap_off[g.ntaps++] = (uint8_t)((tap / 3) * kHaloW + tap % 3);
            add_block([&](int co, int j) { return w_at(co, part_off + c0 + j, tap / 3, tap % 3); });
            wk += 64;
          }
          table.push_back(g); ++ng;
        }
      }
      part_off += q.t.C;
    }
    WSI_REQUIRE(ng <= kHaloMaxGroups, WSI_ERR_UNSUPPORTED, "halo conv: %d groups > %d", ng, kHaloMaxGroups);
    WSI_REQUIRE(groups_per_parity < 0 || groups_per_parity == ng, WSI_ERR_UNSUPPORTED, "halo conv: parity classes differ in group count");
    groups_per_parity = ng;
    Kpar = (size_t)wk;
    wblocks.push_back(std::move(wb));
  }
  p.num_kb = groups_per_parity;
  p.halo_plain = (!up2 && !s2) ? 1 : 0;
  p.b_parity_stride = (p.num_parity > 1) ? (int)Kpar : 0;
  const size_t Ktot = Kpar * p.num_parity;
  std::vector<uint16_t> wp((size_t)cout * Ktot);
  for (int par = 0; par < p.num_parity; ++par) {
    const std::vector<float>& wb = wblocks[par];
    const size_t nblk = Kpar / 64;
    for (size_t blk = 0; blk < nblk; ++blk)
      for (int co = 0; co < cout; ++co)
        for (int j = 0; j < 64; ++j)
          wp[(size_t)co * Ktot + (size_t)par * Kpar + blk * 64 + j] = f32_to_bf16_bits(wb[(blk * cout + co) * 64 + j]);
  }
  upload(w_, wp);
  upload(hgroups_, table);
  p.hgroups = hgroups_.as<HaloGroup>();
  std::vector<float> sc(cout, 1.f), bi(cout, 0.f);
  if (scale) sc.assign(scale, scale + cout);
  if (bias) bi.assign(bias, bias + cout);
  upload(scale_, sc); upload(bias_, bi);
  p.scale = scale_.as<float>(); p.bias = bias_.as<float>();
  encode2(&bmap_, w_.p, (uint64_t)Ktot, (uint64_t)cout, (uint64_t)Ktot * 2, 64, (uint32_t)(block_n_ / 2), 64);
  const long long tiles_m = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
  const long long pair_tiles = (tiles_m + 1) / 2 * p.tiles_co * p.num_parity;
  WSI_REQUIRE(pair_tiles < (1LL << 30), WSI_ERR_UNSUPPORTED, "conv: too many tiles");
  grid_ = 2 * (int)std::min<long long>(pair_tiles, num_sms / 2);
  const int taps = 9;
  flops_ = 2.0 * N * OH * OW * (double)cout * cin * taps;
  CUDA_CHECK(cudaStreamSynchronize(0));
}

template <int BN, bool PLAIN, bool PLAIN_EPI = PLAIN>
static void launch_halo(const AMaps& am, const CUtensorMap& bm, const ConvParams& p, int grid, cudaStream_t s) {
  using S = HaloSmem<BN, PLAIN>;
  static_assert(S::kTotal <= 227 * 1024, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(conv_halo_pair_kernel<BN, PLAIN, PLAIN_EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  conv_halo_pair_kernel<BN, PLAIN, PLAIN_EPI><<<grid, kNumThreads, S::kTotal, s>>>(am, bm, p);   // __cluster_dims__(2, 1, 1)
  CUDA_CHECK(cudaGetLastError());
}

template <int BN>
static void launch_pair(const AMaps& am, const CUtensorMap& bm, const ConvParams& p, int grid, cudaStream_t s) {
  using S = PairSmem<BN>;
  static_assert(S::kTotal <= 227 * 1024, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  conv_igemm_pair_kernel<BN><<<grid, kNumThreads, S::kTotal, s>>>(am, bm, p);   // __cluster_dims__(2, 1, 1)
  CUDA_CHECK(cudaGetLastError());
}

void ConvOp::launch(cudaStream_t stream, LaunchCounter* lc) const {
  if (stream_) { stream_->launch(stream, lc); return; }
  if (upstream_) { upstream_->launch(stream, lc); return; }
  if (row_) { row_->launch(stream, lc); return; }
  if (stem_) { stem_->launch(stream, lc); return; }
  if (halo_) {
    const bool plain = p_.halo_plain && !p_.out_planar;
    if (block_n_ == 256) { if (plain) launch_halo<256, true>(amaps_, bmap_, p_, grid_, stream); else launch_halo<256, false>(amaps_, bmap_, p_, grid_, stream); }
    else if (block_n_ == 128) { if (plain) launch_halo<128, true>(amaps_, bmap_, p_, grid_, stream); else launch_halo<128, false>(amaps_, bmap_, p_, grid_, stream); }
    else launch_halo<64, false>(amaps_, bmap_, p_, grid_, stream);
    if (lc) lc->n++;
    return;
  }
  if (pair_) {
    if (block_n_ == 256) launch_pair<256>(amaps_, bmap_, p_, grid_, stream);
    else launch_pair<128>(amaps_, bmap_, p_, grid_, stream);
    if (lc) lc->n++;
    return;
  }
  if (split_) {
#define WSI_SPLIT_CASE(BN, BK) \
  if (block_n_ == BN && block_k_ == BK) { launch_inst<BN, BK, false, 3>(amaps_, bmap_, p_, grid_, stream); if (lc) lc->n++; return; }
    WSI_SPLIT_CASE(128, 64) WSI_SPLIT_CASE(64, 64) WSI_SPLIT_CASE(32, 64) WSI_SPLIT_CASE(16, 64)
    WSI_SPLIT_CASE(128, 32) WSI_SPLIT_CASE(64, 32) WSI_SPLIT_CASE(32, 32) WSI_SPLIT_CASE(16, 32)
    WSI_SPLIT_CASE(128, 16) WSI_SPLIT_CASE(64, 16) WSI_SPLIT_CASE(32, 16) WSI_SPLIT_CASE(16, 16)
#undef WSI_SPLIT_CASE
    WSI_THROW(WSI_ERR_UNSUPPORTED, "conv (fp32 emulation): no kernel instance for BLOCK_N=%d BLOCK_K=%d", block_n_, block_k_);
  }
  if (resb_) {
    if (block_n_ == 64) launch_inst<64, 64, true>(amaps_, bmap_, p_, grid_, stream);
    else launch_inst<128, 64, true>(amaps_, bmap_, p_, grid_, stream);
    if (lc) lc->n++;
    return;
  }
#define WSI_CASE(BN, BK) \
  if (block_n_ == BN && block_k_ == BK) { launch_inst<BN, BK>(amaps_, bmap_, p_, grid_, stream); if (lc) lc->n++; return; }
  WSI_CASE(256, 64) WSI_CASE(128, 64) WSI_CASE(64, 64) WSI_CASE(32, 64) WSI_CASE(16, 64)
  WSI_CASE(128, 32) WSI_CASE(64, 32) WSI_CASE(32, 32) WSI_CASE(16, 32)
  WSI_CASE(128, 16) WSI_CASE(64, 16) WSI_CASE(32, 16) WSI_CASE(16, 16)
#undef WSI_CASE
  WSI_THROW(WSI_ERR_UNSUPPORTED, "conv: no kernel instance for BLOCK_N=%d BLOCK_K=%d", block_n_, block_k_);
}

// ---------------------------------------------------------------------------------------------
// Hardware probe (dev tool, tools/umma_shift_probe.py): can a SWIZZLE_128B K-major A operand start at an arbitrary
// 128-byte row of a TMA-written tile?  A halo tile [18 rows][pitch px][64 ch] is loaded once; the MMA reads the
// 16 x 8 pixel window shifted by (r, s): start = tile + (r*pitch + s) * 128 B, 8-row groups `pitch` rows apart
// (SBO = pitch * 128 B), descriptor base_offset = (start >> 7) & 7 or 0.  If this works, the 9 taps of a 3x3 conv
// can share ONE halo load (the per-CTA-unique A traffic of the mid layers drops 4-6x).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_shift_probe_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                                                                   int r, int s, int pitch, int use_base_offset, float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = 18 * pitch * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((a_bytes + 1023) & ~1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2048);
  uint32_t* holder = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc(holder, 32);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *holder;
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(&bars[0], (uint32_t)a_bytes + 2048u);
    ptx::tma_load_4d(sA, &amap, &bars[0], 0, 0, 0, 0);
    ptx::tma_load_2d(sB, &bmap, &bars[0], 0, 0);
  }
  ptx::mbar_wait(&bars[0], 0, nullptr, 90);
  ptx::tc_fence_after();
  if (warp == 0 && ptx::elect_one()) {
    const uint32_t start = ptx::smem_u32(sA) + (uint32_t)((r * pitch + s) * 128);
    uint64_t adesc = 0;
    adesc |= (uint64_t)((start & 0x3FFFFu) >> 4);
    adesc |= (uint64_t)1 << 16;
    adesc |= (uint64_t)(((uint32_t)pitch * 128u) >> 4) << 32;          // SBO: next 8-pixel group = next image row
    adesc |= (uint64_t)1 << 46;
    if (use_base_offset) adesc |= (uint64_t)((start >> 7) & 7u) << 49;
    adesc |= (uint64_t)2 << 61;                                        // SWIZZLE_128B
    const uint64_t bdesc = make_kmajor_desc<64>(ptx::smem_u32(sB));
    constexpr uint32_t idesc = make_idesc_bf16<16>();
#pragma unroll
    for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
    ptx::umma_commit(&bars[1]);
  }
  ptx::mbar_wait(&bars[1], 0, nullptr, 91);
  ptx::tc_fence_after();
  uint32_t v[16];
  ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
  ptx::tmem_ld_wait();
  for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 32);
  }
}

void debug_umma_shift(const void* A_dev, const void* B_dev, int r, int s, int pitch, int use_base_offset, float* D_dev, cudaStream_t st) {
  WSI_REQUIRE(pitch >= 10 && pitch <= 16 && r >= 0 && r <= 2 && s >= 0 && s <= 2, WSI_ERR_INVALID, "probe: bad arguments");
  CUtensorMap amap, bmap;
  const uint64_t dims[4] = {64, (uint64_t)pitch, 18, 1};
  const uint64_t strides[3] = {128, (uint64_t)pitch * 128, (uint64_t)pitch * 128 * 18};
  const uint32_t box[4] = {64, (uint32_t)pitch, 18, 1};
  encode4(&amap, A_dev, dims, strides, box, 64);
  encode2(&bmap, B_dev, 64, 16, 128, 64, 16, 64);
  const int smem = 1024 + ((18 * pitch * 128 + 1023) & ~1023) + 2048 + 64;
  CUDA_CHECK(cudaFuncSetAttribute(umma_shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_shift_probe_kernel<<<1, 128, smem, st>>>(amap, bmap, r, s, pitch, use_base_offset, D_dev);
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace wsi
