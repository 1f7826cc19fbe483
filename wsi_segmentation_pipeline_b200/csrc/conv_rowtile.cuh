// conv_rowtile.cuh — 3x3/s1 convolution for the high-resolution, small-channel layers (sm_100a).
//
// The decoder's last levels (smp Unet: 128->32->32 @ H/2, 32->16->16 @ H) and the
// 64->64 BasicBlocks of layer1 / decoder level 3 (resnets_shift.py:49-65) have 16..64 channels per operand
// and 16/32/64 output channels.  Per-tap TMA boxes move one 32..128-byte
// row per pixel per tap (9x re-reads, bound by the TMA box-row rate: measured ~4 cycles/row), and
// 16-byte cp.async producers starve for memory-level parallelism (measured 2 TB/s).  Instead:
//
//   * these activations live in HBM in a zero-padded CHANNEL-CHUNK-PLANAR layout
//         [N][H+2][C/8][Wrow][8 ch],  entry index = x + 8,  Wrow = roundup8(W + 9)     (LAYOUT_PLANAR)
//     so that a halo row of one 8-channel chunk is ONE contiguous, 128-byte aligned run, and the conv
//     zero padding is simply there.  The skip operand of an upsample+concat conv is additionally
//     de-interleaved by column parity: [N][H+2][C/8][2][Wq][8], index = floor(x/2) + 8  (LAYOUT_PLANAR_PARITY)
//   * a tile = 128 consecutive pixels of ONE output row (x2-upsampling convs: 256 pixels = two column
//     parity groups).  Its 3-row halo (144 entries per row: x in [x0-8, x0+136), 128-byte aligned at both
//     ends — misaligned bulk copies run at ~14 B/cycle) is fetched ONCE per 16-channel slab by 6 (12)
//     bulk copies (cp.async.bulk, TMA engine, mbarrier complete_tx) issued by the lanes of one warp into
//         plane[kc][row 0..2][col 0..143][16 B]
//     — the canonical NO-SWIZZLE K-major UMMA layout with SBO = 128 B (rows linear at 16 B pitch),
//     LBO = plane stride.  Every filter tap (r,s) is then only a different descriptor START ADDRESS
//     (+ (r*144+s+7)*16 B): nothing is re-read or re-arranged.  For x2-nearest sources the planes hold
//     the half-resolution rows and the 9 taps collapse to the 4 distinct source positions.
//   * the whole weight tensor (<= 74 KB) stays resident in smem for the life of the persistent CTA.
//   * warps 0-2: bulk-copy producers (one halo row each); warps 3-4: MMA issuers (one accumulator each; warp 3
//     owns the TMEM allocation); then 4 epilogue warps (folded BN +
//     ReLU + bf16 store in planar or NHWC layout, or the fused 1x1 `final_conv` head -> fp32 logits).
#pragma once
#include "conv_igemm.cuh"

namespace wsi {

constexpr int kRowPad = 8;                               // left padding entries of a planar row (128 B)
constexpr int kRowHaloCols = 144;                        // halo entries per row: x in [x0-8, x0+136)
constexpr int kRowPlaneBytes = 3 * kRowHaloCols * 16;    // 6912: [3 rows][144 cols][16 B], a multiple of 128
constexpr int kRowProducerWarps = 3;                     // one per halo row
constexpr int kRowMmaWarps = 2;                          // one per accumulator of a tile
constexpr int kRowEpiGroups = 2;                         // epilogue warp-groups (4 warps each) alternating tiles
constexpr int kRowThreads = (kRowProducerWarps + kRowMmaWarps + 4 * kRowEpiGroups) * 32;
constexpr int kRowMaxSlabs = 8;
constexpr int kRowAccStages = 4;                         // TMEM accumulator ring depth

// geometry of the planar layouts (host + device)
struct PlanarDims {
  int H, W, KC, P, Wrow;   // P = 1 (plain) or 2 (column-parity planes); Wrow = entries per chunk row
  __host__ __device__ static PlanarDims make(int H_, int W_, int C_, int layout) {
    PlanarDims d;
    d.H = H_; d.W = W_; d.KC = C_ / 8;
    d.P = (layout == LAYOUT_PLANAR_PARITY) ? 2 : 1;
    d.Wrow = (((layout == LAYOUT_PLANAR_PARITY) ? (W_ / 2) : W_) + kRowPad + 1 + 7) & ~7;   // x = -1 .. W, rows 128 B aligned
    return d;
  }
  // byte offset of entry 0 of chunk row (n, y in [-1,H], kc, par); pixel x sits at entry x + kRowPad
  // (parity planes: floor(x/2) + kRowPad, with x = -1 at entry kRowPad - 1 of the odd plane)
  __host__ __device__ size_t row_off(int n, int y, int kc, int par) const {
    return (((((size_t)n * (H + 2) + (size_t)(y + 1)) * KC + kc) * P + par) * (size_t)Wrow) * 16;
  }
  __host__ __device__ size_t bytes(int N_) const { return (size_t)N_ * (H + 2) * KC * P * Wrow * 16; }
};

// Epilogue constants carried IN the kernel parameter struct: folded-BN scale / bias and the fused 1x1 head live in the
// constant bank, so `fmaf(acc, k.scale[j], k.bias[j])` with a compile-time j is an FFMA with a c[][] operand — no load
// instruction.  (Round 1 broadcast them from shared memory: ncu showed the LSU shared-memory pipe at 88 % in the d5b
// kernel — 52 broadcast LDS per output row and warp — and at 56-59 % in the other row kernels.)
struct EpiConst {
  float scale[64];
  float bias[64];
  float head[68];      // [4][16] head weights, then the 4 head biases
};

struct RowPart {
  const uint8_t* base;   // planar tensor
  PlanarDims d;
  int mode;              // 0 plain, 1 nearest-x2 source (half res), 2 skip operand of an x2 conv (parity planes)
  int nslabs;            // C / 16
};

struct RowParams {
  RowPart part[2];
  int nparts;
  int N, OH, OW, Cout;
  int up2;                       // output tile = 256 px (two column parities) instead of 128
  int tiles_x, total_tiles;
  int nslabs;                    // 16-channel slabs per tile (all parts)
  int stages, stage_bytes;
  uint16_t adelta[3][2][2][9];   // [mode][output row parity][column parity group][tap]: A start offset, 16-byte units
  uint16_t adelta_up[2][2][4];   // nearest-x2 slabs: [row parity][column parity][collapsed 2x2 position]
  int slab_wblock[kRowMaxSlabs]; // first weight block of each slab
  int w_blocks;                  // total weight blocks (each [2][BN][8] bf16)
  const bf16* w;                 // per slab: 9 tap blocks, or 16 = (parities) x 4 collapsed positions for x2 slabs
  const float* scale;            // [BN]
  const float* bias;             // [BN]
  int relu;
  uint8_t* out;                  // bf16 output, or nullptr
  int out_layout;                // LAYOUT_NHWC or LAYOUT_PLANAR
  const uint8_t* res;            // bf16 residual [N,OH,OW,BN] added before the ReLU, or nullptr
  int res_layout;                // LAYOUT_NHWC or LAYOUT_PLANAR (same geometry as a planar output)
  PlanarDims od;                 // geometry of a planar output
  const float* head_w;           // [4][16] or nullptr
  const float* head_b;
  float* head_out;               // [N,OH,OW,4] fp32
  int* error_flag;
};

class RowConvOp {
 public:
  // true if this conv can run on the row-tile kernel
  static bool eligible(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const void* residual);
  // parts / out may be NHWC (converted by an internal relayout launch; output written NHWC) or planar.
  void build(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const float* w_oihw, const float* scale,
             const float* bias, const void* residual, int res_layout, void* out, int out_layout, const float* head_w,
             const float* head_b, float* head_out, int* error_flag, int num_sms);
  void launch(cudaStream_t stream, LaunchCounter* lc) const;
  double flops() const { return flops_; }
  int block_n() const { return p_.Cout; }
  void set_head_out(float* p) { p_.head_out = p; }

 private:
  struct Relayout { const void* src; void* dst; int N, H, W, C, layout; };
  RowParams p_{};
  DevBuf w_, scale_, bias_, headw_, headb_, stage_in_[2];
  std::vector<Relayout> relayouts_;
  int grid_ = 0, smem_ = 0;
  double flops_ = 0;
};

// ---------------------------------------------------------------------------------------------
// Stem (conv1 7x7/s2/p3, 3 -> 64, + bn1 + relu; resnets_shift.py:196-198) in the same style.
// Input: the gather kernel's zero-padded tiles [n][ph+6][pw+8][4] bf16 (8 B per pixel).  A tile = 128
// output pixels of one output row; its 7 input rows (131 16-byte chunks of 2 pixels each) are fetched
// once by 7 bulk copies.  Output pixel i, filter row r reads padded pixels 2i..2i+7 = chunks i..i+3 of
// row r: consecutive output pixels advance by ONE chunk, so the A operand of (r, k-step j) is the
// no-swizzle descriptor {start = row r + 2j chunks, row pitch 16 B (SBO 128), LBO 16 B} — an
// overlapping (Hankel) view, nothing is im2col'ed.  K = 7 x 32 (8th pixel and 4th channel: zero
// weights), 14 MMAs of 128x64x16 per tile, alternating two partial accumulators.
// ---------------------------------------------------------------------------------------------
constexpr int kStemRowBytes = 132 * 16;                  // 131 chunks needed, padded to 132
constexpr int kStemStageBytes = 7 * kStemRowBytes;       // 14 784

struct StemParams {
  const uint8_t* in;             // padded tiles
  int N, PH, PW, OH, OW;
  int tiles_x, total_tiles, stages;
  const bf16* w;                 // [28 k-chunks][64][8] bf16
  const float* scale;            // [64]
  const float* bias;             // [64]
  EpiConst k;                    // the same constants in the parameter (constant) bank
  uint8_t* out;                  // bf16, NHWC [N,OH,OW,64] or parity-planar
  int out_layout;                // LAYOUT_NHWC or LAYOUT_PLANAR_PARITY
  PlanarDims od;
  int* error_flag;
};

class RowStemOp {
 public:
  void build(const void* padded_tiles, int n, int ph, int pw, const float* w_oihw, const float* scale, const float* bias,
             void* out, int out_layout, int* error_flag, int num_sms);
  void launch(cudaStream_t stream, LaunchCounter* lc) const;
  double flops() const { return flops_; }

 private:
  StemParams p_{};
  DevBuf w_, scale_, bias_;
  int grid_ = 0, smem_ = 0;
  double flops_ = 0;
};

// 3x3/s2/p1 max pool reading the stem's parity-planar output (values are post-ReLU >= 0, so the layout's
// zero border is equivalent to the reference's -inf padding), writing NHWC or the planar layout
void launch_maxpool_planar(const void* x_parity_planar, int n, int h, int w, int c, void* y, int y_layout, cudaStream_t s,
                           LaunchCounter* lc);

// NHWC bf16 -> padded planar (interior only; the zero border is written once at allocation)
void launch_relayout_planar(const void* src_nhwc, void* dst, int N, int H, int W, int C, int layout, cudaStream_t s,
                            LaunchCounter* lc);
void launch_relayout_nhwc(const void* src_planar, void* dst_nhwc, int N, int H, int W, int C, cudaStream_t s, LaunchCounter* lc);

}  // namespace wsi
