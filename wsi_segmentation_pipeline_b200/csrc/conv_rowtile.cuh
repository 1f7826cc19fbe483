// conv_rowtile.cuh — 3x3/s1 convolution for the high-resolution, small-channel layers (sm_100a).
//
// The decoder's last levels (smp Unet: 128->32->32 @ H/2, 32->16->16 @ H) have
// 16..64 channels per operand and 16/32 output channels.  Per-tap TMA boxes would move one
// 32..128-byte row per pixel per tap (9x re-reads, TMA box-row rate bound: measured 4 cycles/row).
// Here a tile is 128 consecutive pixels of ONE output row and its 3 x 130 pixel halo is loaded
// ONCE with 16-byte cp.async (LDGSTS, zero-fill = conv padding) into channel-chunk planes
//     plane[kc][row 0..2][col 0..129][16 B]            (kc = 8-channel chunk)
// which is the canonical NO-SWIZZLE K-major UMMA layout with SBO = 128 B (rows are linear at 16 B
// pitch) and LBO = plane stride.  Every filter tap (r,s) is then just a different descriptor START
// ADDRESS (+ (r*130+s)*16 B) into the same smem — no data is re-read or re-arranged.
//   * nearest x2 upsample: output pixels are processed per column parity; the source planes hold
//     the half-resolution rows, tap offsets become floor((parity+tap-1)/2).
//   * the skip operand of an upsample+concat conv is de-interleaved into even/odd column planes.
//   * the whole weight tensor (<= 74 KB) stays resident in smem for the life of the persistent CTA.
//   * warps 0-5: cp.async producers (one halo-row task each); warp 6: MMA issuer (+TMEM alloc); warps 7-10: epilogue
//     (folded BN + ReLU + bf16 store, or the fused 1x1 `final_conv` head -> fp32 logits).
#pragma once
#include "conv_igemm.cuh"

namespace wsi {

constexpr int kRowHaloCols = 130;
constexpr int kRowPlanePx = 3 * kRowHaloCols;            // 390 pixels
constexpr int kRowPlaneBytes = 6304;                     // 390*16 = 6240, padded so planes land 32 B apart mod 128 (banks)
constexpr int kRowProducerWarps = 6;
constexpr int kRowProducers = kRowProducerWarps * 32;
constexpr int kRowThreads = (kRowProducerWarps + 1 + 4) * 32;
constexpr int kRowMaxSlabs = 8;
constexpr int kRowAccStages = 4;                        // TMEM accumulator ring depth

struct RowPart {
  const bf16* ptr;   // NHWC source
  int H, W, C;       // source extents (half resolution for mode 1)
  int mode;          // 0 plain, 1 nearest-x2 source, 2 skip operand of an x2 conv (parity planes)
};

struct RowParams {
  RowPart part[2];
  int nparts;
  int N, OH, OW, Cout;
  int up2;                       // output tile = 256 px (two column parities) instead of 128
  int tiles_x, total_tiles;
  int nslabs;                    // 16-channel slabs per tile
  int8_t slab_part[kRowMaxSlabs];
  int16_t slab_kc0[kRowMaxSlabs];   // first 8-channel chunk of the slab inside its part
  int stages, stage_bytes;
  const bf16* w;                 // [slab][tap][2][BN][8] bf16
  const float* scale;            // [BN]
  const float* bias;             // [BN]
  int relu;
  bf16* out;                     // NHWC [N,OH,OW,BN] or nullptr
  const float* head_w;           // [4][16] or nullptr
  const float* head_b;
  float* head_out;               // [N,OH,OW,4] fp32
  int* error_flag;
};

class RowConvOp {
 public:
  // true if this conv can run on the row-tile kernel
  static bool eligible(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const void* residual);
  void build(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const float* w_oihw, const float* scale,
             const float* bias, void* out, const float* head_w, const float* head_b, float* head_out, int* error_flag, int num_sms);
  void launch(cudaStream_t stream, LaunchCounter* lc) const;
  double flops() const { return flops_; }
  int block_n() const { return p_.Cout; }

 private:
  RowParams p_{};
  DevBuf w_, scale_, bias_, headw_, headb_;
  int grid_ = 0, smem_ = 0;
  double flops_ = 0;
};

}  // namespace wsi
