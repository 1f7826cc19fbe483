// conv_rowstream.cuh — row-streaming 3x3/s1 convolution for <= 64-channel planar operands (sm_100a).
//
// Successor of the row-tile kernel (conv_rowtile.cuh) for plain (non-upsampling) convs.  The row-tile
// kernel fetches a 3-row halo per output row (every input row 3x) and issues 9 small MMAs per 16-channel
// slab — both the L2->SM fabric and the per-MMA operand fetch (M128 x K16 = 4 KB at ~45 B/clk, whatever N
// is) bound it.  Here a CTA streams INPUT rows down a strip of 128 output columns:
//
//   * each input row is loaded ONCE (2 bulk copies per 16-channel slab, 4.6 KB);
//   * the three vertical taps are stacked along N: one MMA per (slab, horizontal shift s)
//         D[128 px, 3*Cout] += A_s[128 px, 16 ch] * [W(r=2,s) | W(r=1,s) | W(r=0,s)]
//     adds the row's contribution to the THREE output rows it touches (y_in-1, y_in, y_in+1), whose
//     accumulators sit in consecutive slots of a TMEM ring — 3*slabs+1 MMAs per row instead of 9*slabs
//     (the "+1": the first touch of a new output row must overwrite, so its block is issued separately);
//   * a pipeline stage is one input row (all slabs) and there is exactly ONE tcgen05.commit per input row
//     (a commit drains the tensor pipe, ~600 cycles measured): it releases the stage to the producer and,
//     because output row i is complete exactly when input row i+2 has been consumed, it is also what the
//     epilogue warps (folded BN + residual + ReLU + store / fused 1x1 head) wait on; they hand the slot
//     back through a second barrier ring.
//
// Operands and outputs use the padded channel-chunk-planar layout of conv_rowtile.cuh (entry x+8,
// 128-byte aligned rows); shifts s are descriptor start offsets into the same smem row, as before.
#pragma once
#include "conv_rowtile.cuh"

namespace wsi {

constexpr int kRowRunBytes = kRowHaloCols * 16;               // 2304: one halo row of one 8-channel chunk
constexpr int kStreamStageBytes = 2 * kRowRunBytes;        // one input row of one 16-channel slab: 2 chunk runs
constexpr int kStreamThreads = (1 + 1 + 8) * 32;           // producer, MMA issuer, 2 x 4 epilogue warps

struct StreamParams {
  const uint8_t* in;             // planar input
  PlanarDims d;
  int nslabs;                    // C / 16
  int N, OH, OW, Cout;
  int tiles_x, total_rows;        // strips of 128 output columns per image; N * tiles_x * OH
  int stages;                    // pipeline stages, one input row (all slabs) each
  int ring;                      // TMEM accumulator slots (output rows in flight): 8 or 16
  int lanes;                     // 1: 128-column strips (conv_rowstream_kernel); 2: 256-column strips, two lanes (Cout <= 32)
  const bf16* w;                 // [slab][s][2 chunks][3*BN][8] bf16, N order r = 2 | 1 | 0
  const float* scale;
  const float* bias;
  int relu;
  uint8_t* out;
  int out_layout;
  PlanarDims od;
  const uint8_t* res;
  int res_layout;
  const float* head_w;
  const float* head_b;
  float* head_out;
  EpiConst k;                    // scale / bias / head in the parameter (constant) bank: what the epilogues read
  int* error_flag;
  int dbg;                       // timing experiments only (WSI_STREAM_DBG; results are garbage): 1 no MMAs, 2 no loads, 4 independent MMAs,
                                 // 3 epilogue drains nothing
};

class RowStreamOp {
 public:
  static bool eligible(const std::vector<ConvInputPart>& parts, const ConvSpec& spec);
  void build(const ConvInputPart& part, const ConvSpec& spec, const float* w_oihw, const float* scale, const float* bias,
             const void* residual, int res_layout, void* out, int out_layout, const float* head_w, const float* head_b,
             float* head_out, int* error_flag, int num_sms);
  void launch(cudaStream_t stream, LaunchCounter* lc) const;
  double flops() const { return flops_; }
  void set_head_out(float* p) { p_.head_out = p; }
  std::string kernel_name() const {
    char b[96];
    snprintf(b, sizeof(b), "%s<%d%s>", p_.lanes == 2 ? "conv_rowstream2_kernel" : "conv_rowstream_kernel", p_.Cout, p_.head_out ? ",HEAD" : "");
    return b;
  }

 private:
  StreamParams p_{};
  DevBuf w_, scale_, bias_, headw_, headb_, stage_in_;
  bool relayout_ = false;
  const void* relayout_src_ = nullptr;
  int grid_ = 0, smem_ = 0;
  double flops_ = 0;
};

}  // namespace wsi
