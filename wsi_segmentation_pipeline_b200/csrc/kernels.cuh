// kernels.cuh — the HBM-bound stages of the path: tile gather + normalise (K0), max-pool, global
// average pool + heads (K4), overlap-accumulate stitch (K6), softmax/argmax/heatmap finalise (K7),
// and the synthetic-slide generator.  Launch wrappers; kernels live in kernels.cu.
#pragma once
#include "common.cuh"

namespace wsi {

// Rectangles (tiles mapped to canvas coordinates) sorted by (ty, tx) with a row index, the
// lookup structure of the atomic-free gather-formulated stitch.
struct RowInfo {            // one tile row (all sorted rects sharing a canvas y origin), 32 bytes
  int32_t ry;               // canvas y origin
  int32_t a, b;             // sorted rects [a, b)
  int32_t tx0, step, nreg;  // the first nreg rects sit at tx0 + k * step (the regular grid of the tile planner; nreg >= 1,
                            // step = 0 when nreg == 1): their x origin needs no memory access and the first rect right
                            // of a given x is found in O(1); rects [a + nreg, b) (the reference's extra right-column
                            // tile, or anything after a gap left by the foreground filter) are looked up in `tx`
  int32_t pad0, pad1;
};

struct RectIndex {
  const int32_t* tx;        // [T] canvas x origin of sorted rect i
  const int32_t* row_y;     // [R] distinct canvas y origins, ascending
  const int32_t* row_start; // [R+1] first sorted rect of each row
  const RowInfo* rows;      // [R]
  int32_t R;
  int32_t dx, dy;           // rectangle size on the canvas
  int32_t up = 1;           // SEG: the logit tile is (dy / up) x (dx / up), nearest-upsampled onto the rectangle (scan_resize, utils/eval.py:202-206)
};

struct FinaliseArgs {
  int64_t W2;               // canvas width
  int64_t own0, own1;       // canvas rows handled
  const uint8_t* mask;      // [own1-own0, W2] or nullptr (= ones)
  double class_probs[4];    // per-class floor of threshold_probs (compared in double, as the reference does)
  int heat_mode;            // 0: p[2]+p[3] (seg), 1: p[1] (cls)
  uint8_t* classes;         // [rows, W2]
  uint8_t* heatmap;         // [rows, W2]
  float* canvas_out;        // [4, rows, W2] or nullptr
  float* probs_out;         // [4, rows, W2] or nullptr
};

// K0: raster u8 [rows, iw, 3] -> zero-padded normalised bf16 tiles [n][ph+6][pw+8][4] (interior only)
void launch_gather(const uint8_t* rgb, int64_t row_stride, int64_t row0, const int32_t* tiles_xy_dev, int n,
                   int ph, int pw, const float* lut_dev /*f32 [3][256]*/, bf16* padded_or_null, float* norm_out_or_null,
                   cudaStream_t s, LaunchCounter* lc, int planes = 1, int64_t plane_stride = 0);
// (planes == 3, fp32-emulated precision: three padded buffers plane_stride elements apart hold the bf16 expansion a + b + c)
// K0r: PIL-exact bicubic resize of every tile window pw x ph -> tw x th (scan_resize != 1); tmp u8 [n][ph][tw][3] scratch,
// out u8 [n][th][tw][3], out_xy int32 [n][2] = origins of the resized tiles inside `out` seen as one raster of width tw
void launch_resample_tiles(const uint8_t* rgb, int64_t row_stride, int64_t row0, const int32_t* tiles_xy_dev, int n, int ph, int pw, int th,
                           int tw, const int32_t* hb, const int32_t* hk, int hks, const int32_t* vb, const int32_t* vk, int vks,
                           uint8_t* tmp, uint8_t* out, int32_t* out_xy, cudaStream_t s, LaunchCounter* lc);
// normalised f32 NCHW -> padded bf16 tiles (nn.Module shim forward)
void launch_pack_nchw(const float* x, int n, int h, int w, bf16* padded, cudaStream_t s, LaunchCounter* lc, int view = 0, int planes = 1,
                      int64_t plane_stride = 0);
// TTA mean over views in the reference's fp32 order (utils/eval.py:311-334)
void launch_tta_accumulate(float* acc, const float* v, int64_t n, bool first, float final_div, cudaStream_t s, LaunchCounter* lc);
// 3x3/s2/p1 max pool, NHWC bf16, C multiple of 8
void launch_maxpool(const bf16* x, int n, int h, int w, int c, bf16* y, cudaStream_t s, LaunchCounter* lc);
// same on the three-plane tensors of the fp32-emulated precision: x [n,h,w,3c] -> y [n,oh,ow,3c]
void launch_maxpool_split(const bf16* x, int n, int h, int w, int c, bf16* y, cudaStream_t s, LaunchCounter* lc);
// fp32 NHWC <-> three bf16 planes per pixel [a | b | c]
void launch_split_planes(const float* x, int64_t px, int c, bf16* y, cudaStream_t s, LaunchCounter* lc);
void launch_merge_planes(const bf16* x, int64_t px, int c, float* y, cudaStream_t s, LaunchCounter* lc);
// global average pool + up to two Linear layers: feat[512] -> (W1,b1)[n1] (-> ReLU -> (W2,b2)[n2])
void launch_pool_head(const bf16* x4, int n, int hw, int c, const float* w1, const float* b1, int n1,
                      const float* w2, const float* b2, int n2, float* feat_out_or_null, float* out,
                      cudaStream_t s, LaunchCounter* lc, int planes = 1);
// logits f32 NHWC4 [n,h,w,4] -> NCHW [n,4,h,w]
void launch_nhwc4_to_nchw(const float* x, int n, int h, int w, float* y, cudaStream_t s, LaunchCounter* lc);

// K6 + K7 fused (seg): rows [y0, y1) of the canvas are final (every tile that touches them is among the sorted tiles
// [t_lo, t_hi), whose fp32 logits [dy][dx][4] sit in slot (i % ring_cap) of `ring`): sum the covering tiles per pixel in
// double in sorted order, softmax / floor / argmax / heat, write classes + heatmap (+ optional fp32 canvas / probs).
// One owner thread per pixel; no canvas in memory, no atomics.
void launch_stitch_finalise_seg(const RectIndex& ri, const float4* ring, int ring_cap, int t_lo, int t_hi, int r_lo, int r_hi, int64_t y0, int64_t y1,
                                const FinaliseArgs& a, cudaStream_t s, LaunchCounter* lc);
// ([r_lo, r_hi): the tile rows that can touch canvas rows [y0, y1), found by the host)
// K6+K7 (cls): per-tile logits [T][4] (sorted order) broadcast over rectangles.  bx [nbx] / by [nby]: the canvas cut at every
// tile edge (first pixel of each cell, ascending; the host builds them); cellx [W2] / celly [owned rows]: cell index of every
// column / owned row; cells: scratch of nbx * nby * cls_cell_bytes().  One sum + softmax per CELL, 2 bytes written per pixel.
size_t cls_cell_bytes();
void launch_stitch_finalise_cls(const RectIndex& ri, const float4* tile_logits, int T, const int32_t* bx, int nbx, const int32_t* by, int nby,
                                const int32_t* cellx, const int32_t* celly, void* cells, const FinaliseArgs& a, cudaStream_t s, LaunchCounter* lc);
// coverage counts from the rect index
void launch_counts(const RectIndex& ri, int T, int64_t W2, int64_t own0, int64_t own1, int32_t* counts,
                   cudaStream_t s, LaunchCounter* lc);

// synthetic slide rows [y0, y1) (twin of synth.py)
void launch_synth(int64_t ih, int64_t iw, uint32_t seed, int64_t y0, int64_t y1, const uint8_t* lut_dev,
                  uint8_t* rgb, int64_t row_stride, uint8_t* mask_or_null, cudaStream_t s, LaunchCounter* lc);

// A9: cv2.resize(INTER_LINEAR) of the [4][H][W] summed-logit canvas to [H2][W2] + argmax (predict_wsis, utils/eval.py:66-81)
void launch_resize_argmax(const float* src, int64_t H, int64_t W, int64_t H2, int64_t W2, uint8_t* classes, float* pred_or_null,
                          cudaStream_t s, LaunchCounter* lc);

// A12: find_nuclei(mode='hsv') — HSV saturation threshold through a (max, min) bit table built in float64 on the host
void launch_find_nuclei(const uint8_t* rgb, int64_t row_stride, int64_t H, int64_t W, const uint32_t* lut_bits, uint8_t* mask, cudaStream_t s,
                        LaunchCounter* lc);
// A1/A12: isforeground window counts of the tile planner, one CTA per candidate window
void launch_window_count(const uint8_t* mask, int64_t mh, int64_t mw, const int64_t* win, int64_t n, int64_t dx, int64_t dy, uint32_t* counts,
                         int64_t* sizes, cudaStream_t s, LaunchCounter* lc);

// multi-patch ResNet ensemble head `fc` (resnets_shift.py:133-139, :213-215) on patch-major pooled features
void launch_ensemble_head(const float* feats, int B, int P, const float* w1, const float* b1, int n_hid, const float* w2, const float* b2,
                          int n_out, float* hid, float* out, cudaStream_t s, LaunchCounter* lc);

// ---- tumour-bed post-processing (postproc.cu) ----------------------------------------------------
void launch_lut(const uint8_t* src, int64_t n, const uint8_t* lut_dev, uint8_t* dst, unsigned long long* nonzero_or_null, cudaStream_t s, LaunchCounter* lc);
// one k x k erosion (is_max = false) / dilation (true) with cv2's window and border rule; tmp: H*W scratch; src may equal dst
void launch_morph(const uint8_t* src, int64_t H, int64_t W, int k, bool is_max, uint8_t* tmp, uint8_t* dst, unsigned long long* nonzero_or_null,
                  cudaStream_t s, LaunchCounter* lc);
void launch_row_extent(const uint8_t* mask, int64_t H, int64_t W, int32_t* xmin, int32_t* xmax, cudaStream_t s, LaunchCounter* lc);
void launch_fill_rows(const int32_t* xl, const int32_t* xr, int64_t H, int64_t W, uint8_t* out, cudaStream_t s, LaunchCounter* lc);
void launch_bwperim(const uint8_t* bw, int64_t H, int64_t W, uint8_t* out, cudaStream_t s, LaunchCounter* lc);
void launch_overlay_heat(const uint8_t* rgb, const uint8_t* heat, int64_t n_px, const uint8_t* on_lut_dev, uint8_t* out, cudaStream_t s, LaunchCounter* lc);
void launch_overlay_bed(const uint8_t* rgb, const uint8_t* heat, const uint8_t* im_or_null, const uint8_t* perim_or_null, int64_t n_px, uint8_t* out,
                        cudaStream_t s, LaunchCounter* lc);

}  // namespace wsi
