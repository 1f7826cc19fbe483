// conv_upstream.cuh — row-streaming 3x3/s1 convolution over [nearest-x2 upsample(u) ++ skip] (sm_100a).
//
// The decoder's two highest-resolution "a" convs (smp DecoderBlock conv1: 128->32 @256^2 and 32->16 @512^2 per
// 512 px tile) have Cout = 32 / 16: an MMA of M128 x N32 x K16 is bound by the 4 KB A-operand fetch, not by math,
// and the row-tile kernel (conv_rowtile.cuh) issues one such MMA per (tap, slab, output row).  This kernel applies
// the row-stream idea (conv_rowstream.cuh) to the x2-upsampling case:
//
//   * a CTA streams SOURCE rows a of u down a strip of 128 source columns (= 256 output columns); step a loads
//     u row a and, when there is a skip operand, skip rows 2a and 2a+1 — every input row is read ONCE;
//   * output pixels are processed per column parity p (ox = 2b + p): for a fixed parity the 3 horizontal taps of
//     the upsampled operand collapse onto 2 source columns (weights pre-summed in fp32 on the host) and the
//     skip operand is read through its column-parity planes, so the 128 M rows of every MMA are 128 CONSECUTIVE
//     smem entries and a tap is a descriptor start offset (no-swizzle K-major, as in the other row kernels);
//   * vertical taps are stacked along N: u row a feeds the FOUR output rows 2a-1 .. 2a+2 with the collapsed
//     vertical weights [W2 | W1+W2 | W0+W1 | W0] in one MMA of N = 4*Cout; a skip row y feeds y-1 .. y+1 with
//     [W2 | W1 | W0] (N = 3*Cout).  Per step and slab: 2 x 2 MMAs for u, 2 x 2 x 3 for the skip rows, instead of
//     2 x 9 + 2 x 9 per output row pair in the row-tile kernel;
//   * the two parities are independent accumulator chains and their MMAs are issued alternately (a dependent
//     tcgen05.mma chain on one accumulator runs at ~160 cycles per instruction, measured);
//   * accumulators: TMEM ring of `ring` output rows x 2 parities x Cout columns (512 columns); rows 2a-1 and 2a are
//     complete after step a, signalled by a commit on row_done[slot] each; 4 epilogue groups of 4 warps
//     drain them (thread = source column b: pixels 2b, 2b+1 -> 32 contiguous bytes per channel chunk of the
//     planar output) and return the slot through slot_free[].
//
// Operands: u planar [N][h+2][Cu/8][Wrow][8], skip column-parity planar [N][OH+2][Cs/8][2][Wq][8], output planar
// (conv_rowtile.cuh: PlanarDims).  Reference: smp DecoderBlock (restated in oracle/wsi_oracle.py) —
// F.interpolate(scale_factor=2, mode="nearest"), torch.cat([x, skip], 1), conv3x3 + BN + ReLU.
#pragma once
#include "conv_rowstream.cuh"

namespace wsi {

constexpr int kUpThreads = (1 + 1 + 16) * 32;      // producer, MMA issuer, 4 x 4 epilogue warps

struct UpParams {
  const uint8_t* u;              // planar half-resolution operand
  PlanarDims du;
  int nslabs_u;                  // Cu / 16
  const uint8_t* skip;           // column-parity planar full-resolution operand, or nullptr
  PlanarDims ds;
  int nslabs_s;                  // Cs / 16
  int N, h, w, OH, OW, Cout;
  int tiles_x, total_rows;       // strips of 128 source columns per image; N * tiles_x * h source rows
  int stages, stage_bytes;       // pipeline stage = one input row item (u row, or one skip row)
  const bf16* wts;               // u blocks [slab][parity][tap2][2 chunks][4*BN][8], then skip blocks [slab][s][2][3*BN][8]
  int w_bytes, wskip_off;
  const float* scale;
  const float* bias;
  EpiConst k;                    // scale / bias in the parameter (constant) bank: what the epilogue reads
  int relu;
  uint8_t* out;                  // planar output
  PlanarDims od;
  int* error_flag;
  int dbg;                       // timing experiments only (WSI_UP_DBG): 1 no MMAs
};

class UpStreamOp {
 public:
  static bool eligible(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const void* residual, int out_layout);
  void build(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const float* w_oihw, const float* scale,
             const float* bias, void* out, int* error_flag, int num_sms);
  void launch(cudaStream_t stream, LaunchCounter* lc) const;
  double flops() const { return flops_; }

 private:
  struct Relayout { const void* src; void* dst; int N, H, W, C, layout; };
  UpParams p_{};
  DevBuf w_, scale_, bias_, stage_in_[2];
  std::vector<Relayout> relayouts_;
  int grid_ = 0, smem_ = 0;
  double flops_ = 0;
};

}  // namespace wsi
