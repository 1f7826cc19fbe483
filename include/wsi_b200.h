/* wsi_b200.h — C-ABI of libwsi_b200.so: the B200-native sliding-window whole-slide inference path.
 *
 * The reference (acproject/wsi-segmentation-pipeline) is pure Python and has no FFI layer; its
 * hot path is the duck-typed Python surface described in SURVEY.md §8(b).  This header is the
 * boundary a maintainer would bind with ctypes/cffi; each entry point cites
 * the reference code it replaces (paths relative to the reference root).
 *
 * Conventions: every function returns 0 (WSI_OK) or a negative wsi_status; no exceptions cross
 * the ABI; the caller owns every buffer it passes in; the library owns wsi_ctx and anything it
 * returns through an out-pointer (release with wsi_free).  A wsi_ctx is bound to one CUDA device
 * and is not thread-safe (one per GPU / process).  All device work is ordered on the `stream`
 * argument (a cudaStream_t passed as void*, NULL = the legacy default stream); the context's two
 * copy streams (chunked raster upload, strip-wise result download) are ordered against it with
 * events.  Calls whose inputs AND outputs are device pointers never synchronise the host (small index
 * arrays travel through the context's pinned staging); calls with WSI_MEM_HOST outputs return when
 * those outputs are complete.  A device-side failure inside an asynchronous call (the conv kernels'
 * pipeline barriers time out instead of hanging) raises a sticky flag that the next synchronising
 * call, or wsi_check, reports.  Device pointers need no particular alignment.  There is no CPU
 * fallback: without a CUDA device wsi_ctx_create fails with WSI_ERR_CUDA.
 */
#ifndef WSI_B200_H
#define WSI_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define WSI_API __attribute__((visibility("default")))
#else
#define WSI_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wsi_ctx wsi_ctx;
typedef struct wsi_tiff wsi_tiff;       /* an open TIFF / SVS file (slide ingestion) */

typedef enum {
  WSI_OK = 0,
  WSI_ERR_INVALID = -1,      /* bad argument */
  WSI_ERR_CUDA = -2,         /* CUDA runtime / driver error (see wsi_last_error) */
  WSI_ERR_NOMODEL = -3,      /* wsi_model_load has not been called / tensor missing */
  WSI_ERR_DEGENERATE = -4,   /* geometry on which the reference itself raises (empty mask window) */
  WSI_ERR_UNSUPPORTED = -5,  /* e.g. num_classes != 4, seg mode with m != 1 */
  WSI_ERR_NOMEM = -6
} wsi_status;

/* Model families on the path.  RESNET18: resnets_shift.py:111-217 (trunk + fc0 [+ fc]).
 * UNET_R18: smp.Unet('resnet18') + Classifier/Regressor heads, eval_tumorbed.py:21-28. */
typedef enum { WSI_ARCH_RESNET18 = 0, WSI_ARCH_UNET_R18 = 1 } wsi_arch;

/* What is computed per tile (utils/eval.py:196-200):
 *   SEG        model.decoder(model.encoder(x))            -> [C, ph, pw] logits, stitched per pixel
 *   CLS        model.classifier(model.encoder(x)[0])      -> [C] logits broadcast over the tile
 *              (for WSI_ARCH_RESNET18: fc0(flatten(avgpool(trunk(x)))), resnets_shift.py:206-208)
 *   REG        model.regressor(model.encoder(x)[0])       -> [1]   (utils/eval.py:315-317,399-400)
 *   FEATURES   flatten(avgpool(trunk(x)))                 -> [512] (resnets_shift.py:211-212)    */
typedef enum { WSI_HEAD_SEG = 0, WSI_HEAD_CLS = 1, WSI_HEAD_REG = 2, WSI_HEAD_FEATURES = 3 } wsi_head;

typedef enum { WSI_MEM_HOST = 0, WSI_MEM_DEVICE = 1 } wsi_mem_kind;

/* Arithmetic of the network (wsi_set_option "precision").  The reference runs fp32 end to end (utils/eval.py:196-200;
 * AMP is commented out, train.py:14-15,101-102).
 *   WSI_PRECISION_BF16  bf16 operands (activations and weights rounded once), fp32 accumulation in TMEM, fp32 BN /
 *                       residual / ReLU epilogues: the throughput mode (north_star tolerance 1e-2 / 99.9 % argmax).
 *   WSI_PRECISION_FP32  fp32 emulated on the same bf16 tensor cores: every activation and weight is the exact sum of
 *                       three bf16 planes (8+8+8 mantissa bits = the fp32 value) and every K block issues the six plane
 *                       products of weight >= 2^-24; fp32 accumulation.  Matches an fp32 evaluation to fp32 rounding
 *                       noise (north_star tolerance 1e-4 on probabilities); ~6x the MMA work, 3x the activation bytes. */
typedef enum { WSI_PRECISION_BF16 = 0, WSI_PRECISION_FP32 = 1 } wsi_precision;

/* One fp32 state_dict entry (host memory), named exactly as in the reference's checkpoints
 * (utils/networks.py:6-10 -> model.load_state_dict(state['state_dict'])). */
typedef struct {
  const char* name;
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} wsi_tensor_desc;

/* A slide (or one row band of it) at scan level.  Replaces the OpenSlide handle + DataLoader
 * of utils/dataset.py:110-201 for rasters that are already decoded. */
typedef struct {
  const uint8_t* rgb;      /* u8 [rows, iw, 3], row `row0` of the scan-level raster first      */
  int64_t row_stride;      /* bytes between raster rows (>= 3*iw)                               */
  int32_t rgb_mem;         /* wsi_mem_kind                                                      */
  int64_t ih, iw;          /* full scan-level dimensions (level_dimensions[scan_level])         */
  int64_t row0, rows;      /* raster rows present: [row0, row0+rows) (band + halo); 0, ih = all */
  int32_t ph, pw;          /* tile size (utils/dataset.py params.ph/pw)                         */
  double m;                /* level_downsamples[scan_level] / level_downsamples[2]              */
  int64_t H2, W2;          /* level-2 canvas dimensions (utils/eval.py:182)                     */
  int64_t own0, own1;      /* canvas rows this call owns and writes: [own0, own1); 0, H2 = all  */
  const uint8_t* mask;     /* level-2 foreground mask rows [own0, own1), u8 {0,1} [own1-own0,W2];
                              NULL = all ones (utils/eval.py:219,225)                           */
  int32_t mask_mem;        /* wsi_mem_kind                                                      */
  int32_t resize;          /* myargs.py:115 scan_resize r; 0 or 1 = off.  r > 1: ph x pw (= tile_h*r x tile_w*r,
                              eval_tumorbed.py:39-40) windows are resized to tile_h x tile_w exactly as
                              PIL.Image.resize((tile_w, tile_h)) does at its default filter
                              (utils/dataset.py:180-181), the network runs on those, and SEG logits are
                              nearest-upsampled x r before the slice-add (F.interpolate, utils/eval.py:202-206) */
} wsi_slide_desc;

/* Outputs of one slide/band, all optional except classes+heatmap; rows [own0, own1) only. */
typedef struct {
  int32_t mem;             /* wsi_mem_kind of every non-NULL pointer below                      */
  uint8_t* classes;        /* u8 [rows, W2]   argmax of threshold_probs (preprocessing.py:172)  */
  uint8_t* heatmap;        /* u8 [rows, W2]   uint8(255*mask*h) (utils/eval.py:220-228)         */
  float* canvas;           /* f32 [C, rows, W2] summed logits (utils/eval.py:213-215), or NULL  */
  float* probs;            /* f32 [C, rows, W2] softmax after class_probs floor, or NULL        */
  int32_t* counts;         /* i32 [rows, W2]  number of tiles covering each pixel, or NULL      */
  float* tile_logits;      /* f32 [T, C] (CLS) per-tile logits in the order of tiles_xy, or NULL */
} wsi_out_desc;

/* ---- context ------------------------------------------------------------------------------ */
WSI_API int wsi_ctx_create(int device, wsi_ctx** out);
WSI_API int wsi_ctx_destroy(wsi_ctx* ctx);
WSI_API const char* wsi_last_error(wsi_ctx* ctx);           /* ctx may be NULL: last global error       */
WSI_API const char* wsi_version(void);
/* knobs: "batch_tiles" (tiles per forward batch, 0 = auto), "stage_timing" (0/1), "precision" (wsi_precision),
 * "op_trace" (0/1, see wsi_op_stats) */
WSI_API int wsi_set_option(wsi_ctx* ctx, const char* key, int64_t value);
WSI_API int wsi_set_class_probs(wsi_ctx* ctx, const float* p, int n);   /* myargs.py:15 class_probs     */
WSI_API int64_t wsi_kernel_launches(wsi_ctx* ctx);          /* kernels launched by this ctx so far      */
/* Synchronise `stream` and report a device-side failure of an earlier asynchronous call (the conv kernels' pipeline
 * barriers time out instead of hanging and raise a sticky flag).  Calls with WSI_MEM_HOST outputs do this themselves. */
WSI_API int wsi_check(wsi_ctx* ctx, void* stream);

/* ---- model (replaces model.load_state_dict + .cuda(), eval_tumorbed.py:30-46) ---------------- */
WSI_API int wsi_model_load(wsi_ctx* ctx, int arch, const wsi_tensor_desc* tensors, int n, int num_classes);

/* ---- tile planner (replaces Dataset_wsi.__init__, utils/dataset.py:143-166) ------------------ */
/* xy_out: malloc'd int32 [T,2] (x,y) in the reference's enumeration order; free with wsi_free.
 * mask NULL = all foreground.  Host-only: needs no ctx and no GPU.                              */
WSI_API int wsi_plan_tiles(int64_t ih, int64_t iw, int32_t ph, int32_t pw, int32_t sh, int32_t sw,
                   const uint8_t* mask, int64_t mh, int64_t mw, double m,
                   int32_t** xy_out, int64_t* n_out);
WSI_API void wsi_free(void* p);

/* ---- tile resize tables of the scan_resize branch (host-only; utils/dataset.py:180-181 -> PIL.Image.resize) ---
 * One axis of Pillow's antialiased bicubic resample in_size -> out_size as the library computes it (Pillow 12.2,
 * libImaging/Resample.c): per output sample the window [bounds[2i], bounds[2i] + bounds[2i+1]) of input samples and
 * its 22-bit fixed-point weights kk[i*ksize ...]; out = clip8((2^21 + sum(kk * in)) >> 22), horizontal pass first with a
 * u8 intermediate.  wsi_run_slide uses exactly these tables on the device (wsi_slide_desc.resize).               */
WSI_API int wsi_resample_ksize(int32_t in_size, int32_t out_size);
WSI_API int wsi_resample_coeffs(int32_t in_size, int32_t out_size, int32_t* bounds /*[out_size][2]*/, int32_t* kk /*[out_size][ksize]*/);

/* ---- row-band partition for multi-GPU (no reference counterpart) ----------------- */
/* bands[k] = {own0, own1, row0, row1}: canvas rows owned, and scan-level raster rows needed
 * (own rows + halo so that every tile intersecting the band is complete).  seg/m==1 geometry.  */
WSI_API int wsi_band_partition(int64_t ih, int32_t ph, int32_t sh, int32_t nranks, int64_t* bands /*[nranks*4]*/);
/* keep the tiles that intersect canvas rows [own0, own1); idx_out malloc'd int64 indices        */
WSI_API int wsi_band_tiles(const int32_t* xy, int64_t n, int32_t ph, double m, int64_t own0, int64_t own1,
                   int64_t** idx_out, int64_t* n_out);

/* ---- peer-mapped result (multi-GPU, one process per GPU of one node; no reference counterpart) -----
 * The band outputs of every rank are the u8 rows [own0, own1) of ONE result [2][H2][W2] (classes, heatmap) that lives
 * in rank 0's HBM.  Rank 0 allocates it (wsi_ipc_alloc -> device pointer + a 64-byte CUDA IPC handle to send to the
 * other processes), every other rank maps it (wsi_ipc_open) and passes pointers INTO the mapping as the device outputs
 * of wsi_run_slide: the fused stitch + finalise kernel then stores its u8 rows straight into rank 0's memory over
 * NVLink / NVSwitch — compute and "gather" are one kernel, no staging, no collective besides a closing barrier.
 * The caller synchronises its stream and runs a barrier across ranks before rank 0 reads the result.             */
WSI_API int wsi_ipc_alloc(wsi_ctx* ctx, int64_t bytes, void** dev_ptr, uint8_t* handle /*[64]*/);
WSI_API int wsi_ipc_open(wsi_ctx* ctx, const uint8_t* handle /*[64]*/, void** dev_ptr);
WSI_API int wsi_ipc_close(wsi_ctx* ctx, void* dev_ptr);      /* on the ranks that opened it */
WSI_API int wsi_ipc_free(wsi_ctx* ctx, void* dev_ptr);       /* on the rank that allocated it, after the others closed */

/* Page-lock / release caller-owned host memory (a result buffer shared between the ranks' processes): wsi_run_slide's
 * strip-wise downloads into it are then asynchronous DMA.  WSI_ERR_NOMEM when the platform refuses (nothing stays locked). */
WSI_API int wsi_host_register(void* ptr, int64_t bytes);
WSI_API int wsi_host_unregister(void* ptr);

/* ---- the hot path (replaces the loop of predict_tumorbed, utils/eval.py:190-228) ------------- */
WSI_API int wsi_run_slide(wsi_ctx* ctx, const wsi_slide_desc* slide, const int32_t* tiles_xy, int64_t n_tiles,
                  int head, const wsi_out_desc* out, void* stream);

/* ---- 4-view test-time augmentation of predict_reg / predict_breastpathq (utils/eval.py:303-334, :384-405) ------
 * x: f32 [n, 3, h, w] normalised square tiles; out: f32 [n, dim] = mean over the views {x, x.transpose(2,3), x.flip(2),
 * x.transpose(2,3).flip(3)} of the REG (regressor, dim 1) or CLS head, accumulated in the reference's fp32 order.    */
WSI_API int wsi_forward_batch_tta(wsi_ctx* ctx, const float* x, int64_t n, int32_t h, int32_t w, int head, float* out,
                          int mem, void* stream);

/* ---- resnets_shift.ResNet.forward (resnets_shift.py:189-217): multi-patch classifier with the ensemble head ---
 * xs: f32 [P*B, 3, h, w] normalised patches, PATCH-MAJOR (xs.transpose(0, 1) of the reference's [B, P, 3, h, w]);
 * y: f32 [P*B, 4] = cat(y_list, 0), the per-patch fc0 logits; ens: f32 [B, 4] = fc(cat(x_list, 1)), the ensemble head
 * Linear(P*512, P*256) + ReLU + Linear(P*256, 4) on the pooled trunk features.  Needs WSI_ARCH_RESNET18 with
 * fc.0.* / fc.2.* in the loaded state dict.  All buffers in `mem` memory.                                        */
WSI_API int wsi_forward_patches(wsi_ctx* ctx, const float* xs, int64_t B, int32_t P, int32_t h, int32_t w, float* y,
                        float* ens, int mem, void* stream);

/* ---- foreground mask and masked tile plan on the GPU ------------------------------------
 * wsi_find_nuclei: find_nuclei(wsi, mu_percent, mode='hsv', fill_mask=False) (utils/preprocessing.py:74-110):
 * mask = u8 {0,1}, HSV saturation of the level-2 thumbnail > mu_percent, evaluated as skimage.color.rgb2hsv does
 * (float64, S = (max-min)/max on u8/255, 0 where max == min) — bit-exact through a (max,min) table.
 * rgb: u8 [H][row_stride] interleaved RGB in rgb_mem memory; mask: u8 [H][W] in mask_mem memory.                  */
WSI_API int wsi_find_nuclei(wsi_ctx* ctx, const uint8_t* rgb, int64_t row_stride, int rgb_mem, int64_t H, int64_t W,
                    double mu_percent, uint8_t* mask, int mask_mem, void* stream);
/* wsi_plan_tiles_gpu: same result as wsi_plan_tiles (Dataset_wsi.__init__, utils/dataset.py:143-166) with the
 * isforeground window counts (utils/preprocessing.py:60-71) evaluated on the device; `mask` (required) may be
 * device-resident, e.g. the output of wsi_find_nuclei.  xy_out is malloc'd (wsi_free).                            */
WSI_API int wsi_plan_tiles_gpu(wsi_ctx* ctx, int64_t ih, int64_t iw, int32_t ph, int32_t pw, int32_t sh, int32_t sw,
                       const uint8_t* mask, int mask_mem, int64_t mh, int64_t mw, double m, int32_t** xy_out,
                       int64_t* n_out, void* stream);

/* ---- predict_wsis tail (utils/eval.py:66-71 cv2.resize per class to the level-2 size, :81 np.argmax) ----------
 * canvas: f32 [4][H][W] summed logits at scan-level resolution (the `canvas` output of wsi_run_slide run with m = 1,
 * H2 = ih, W2 = iw); classes: u8 [H2][W2]; pred_or_null: f32 [4][H2][W2] (the reference's resized `pred`).
 * INTER_LINEAR exactly as OpenCV evaluates it (half-pixel centres, border clamp, float weights), fp32 accumulation.
 * All three buffers in `mem` memory.                                                                              */
WSI_API int wsi_resize_argmax(wsi_ctx* ctx, const float* canvas, int64_t H, int64_t W, int64_t H2, int64_t W2,
                      uint8_t* classes, float* pred_or_null, int mem, void* stream);

/* ---- slide ingestion: JPEG-compressed tiled / stripped TIFF and Aperio SVS -> raster rows in HBM -------
 * Replaces openslide.OpenSlide(path) + read_region(...).convert('RGB') (utils/dataset.py:121,175-178; utils/eval.py:263) for
 * rasters still on disk: a row band of one pyramid level (TIFF directory `level`) is decoded by nvJPEG on the GPU straight
 * into the u8 [rows][W][3] device raster that wsi_run_slide / wsi_find_nuclei read.  Classic TIFF and BigTIFF, little-endian,
 * compression 7 (JPEG) with shared JPEGTables, photometric YCbCr or RGB.  Decoded pixels may differ from libjpeg's by a
 * level or two (IDCT / chroma upsampling are decoder-specific).  WSI_ERR_UNSUPPORTED without libnvjpeg or for other codecs. */
WSI_API int wsi_tiff_open(const char* path, wsi_tiff** out);
WSI_API int wsi_tiff_close(wsi_tiff* t);
WSI_API const char* wsi_tiff_last_error(wsi_tiff* t);
WSI_API int wsi_tiff_levels(wsi_tiff* t);                      /* number of image directories */
WSI_API int wsi_tiff_level_info(wsi_tiff* t, int level, int64_t* W, int64_t* H, int32_t* tile_w, int32_t* tile_h, int32_t* compression,
                        int32_t* photometric);
/* host-only: the complete JPEG stream of tile / strip `unit` (JPEGTables spliced in); *len = its size (buf may be NULL) */
WSI_API int wsi_tiff_unit_stream(wsi_tiff* t, int level, int64_t unit, uint8_t* buf, int64_t cap, int64_t* len);
/* rows [row0, row0 + rows) of `level` -> rgb_dev (device, row_stride >= 3 * W bytes); returns when the rows are decoded */
WSI_API int wsi_tiff_read_rows(wsi_ctx* ctx, wsi_tiff* t, int level, int64_t row0, int64_t rows, uint8_t* rgb_dev, int64_t row_stride,
                       void* stream);

/* ---- tumour-bed post-processing of the outputs -------------------------------------------------
 * cv2.morphologyEx / cv2.erode / cv2.dilate with an np.ones((k, k)) kernel, default anchor and border (utils/eval.py:93,96;
 * paper_tools/overlay_tb_wsi.py:50-54,63-67; paper_tools/check_for_false_positives.py:66-70): bit-exact with OpenCV.
 * src / dst: u8 [H][W] in `mem` memory (dst may alias src). */
typedef enum { WSI_MORPH_ERODE = 0, WSI_MORPH_DILATE = 1, WSI_MORPH_OPEN = 2, WSI_MORPH_CLOSE = 3 } wsi_morph_op;
WSI_API int wsi_morph(wsi_ctx* ctx, const uint8_t* src, int64_t H, int64_t W, int op, int k, uint8_t* dst, int mem, void* stream);
/* The tumour-bed chain of utils/eval.py:90-96 (and of the two paper_tools scripts):
 *   tb = rule(src)                      rule_lut: host u8 [256], the reference's threshold evaluated for every u8 level,
 *                                       e.g. classes >= 2 (:91), heatmap >= 0.99 * 255, heatmap / 255 >= 0.9
 *   tb = MORPH_OPEN(tb, ones(open_k))                    -> opened_out (u8 {0,1}), *n_open_out = count_nonzero
 *   tb_pred = convex_hull_image(tb)                      -> hull_out   (skimage, restated: boundary-inclusive, exact integer test)
 *   outline = dilate(bwperim(tb_pred), ones(dilate_k))   -> outline_out (mahotas.bwperim n=4 restated; dilate_k <= 1: no dilation)
 * Any output pointer may be NULL; all buffers in `mem` memory. */
WSI_API int wsi_tumor_bed(wsi_ctx* ctx, const uint8_t* src, int64_t H, int64_t W, const uint8_t* rule_lut, int open_k, int dilate_k,
                  uint8_t* opened_out, uint8_t* hull_out, uint8_t* outline_out, int64_t* n_open_out, int mem, void* stream);
/* host-only part of convex_hull_image (no ctx, no GPU): per-row first / last set column (-1 / -1 for an empty row) ->
 * per-row inclusive pixel range [xl, xr] of the hull image (xl > xr: empty row).  Exposed for CPU tests. */
WSI_API int wsi_hull_rows(const int32_t* xmin, const int32_t* xmax, int64_t H, int32_t* xl, int32_t* xr);
/* Overlays.  WSI_OVERLAY_HEAT: np.uint8(img * 0.75 + 255 * rule(heat) * 0.25) (utils/eval.py:262-267; on_lut = host u8 [256],
 * heat > 255 * 0.99 per level).  WSI_OVERLAY_BED: np.uint8(0.65 * img + 0.35 * (heat * im)) with outline pixels set to 0
 * (paper_tools/overlay_tb_wsi.py:56-72; im / perim may be NULL).  rgb / out: u8 [H][W][3]; heat / im / perim: u8 [H][W]. */
typedef enum { WSI_OVERLAY_HEAT = 0, WSI_OVERLAY_BED = 1 } wsi_overlay_mode;
WSI_API int wsi_overlay(wsi_ctx* ctx, const uint8_t* rgb, const uint8_t* heat, int64_t H, int64_t W, int mode, const uint8_t* on_lut,
                const uint8_t* im, const uint8_t* perim, uint8_t* out, int mem, void* stream);

/* ---- one batch through the network (nn.Module shim forward; utils/eval.py:196-200) ----------- */
/* x: f32 [n, 3, h, w] already normalised (standard_augmentor output); out: SEG f32 [n,C,h,w],
 * CLS [n,C], REG [n,1], FEATURES [n,512].  Both in `mem` memory.                                 */
WSI_API int wsi_forward_batch(wsi_ctx* ctx, const float* x, int64_t n, int32_t h, int32_t w, int head,
                      float* out, int mem, void* stream);
/* same, but tiles are cut from a u8 raster with the fused gather+normalise kernel               */
WSI_API int wsi_forward_tiles(wsi_ctx* ctx, const wsi_slide_desc* slide, const int32_t* tiles_xy, int64_t n_tiles,
                      int head, float* out, int mem, void* stream);

/* ---- synthetic slide on device (twin of synth.py) -------------------------------- */
WSI_API int wsi_synth_slide(wsi_ctx* ctx, int64_t ih, int64_t iw, uint32_t seed, int64_t y0, int64_t y1,
                    const uint8_t* lut /*host u8[16*8*3]*/, uint8_t* rgb_dev, int64_t row_stride,
                    uint8_t* mask_dev_or_null, void* stream);

/* ---- per-kernel debug/bench entry points (tests/, bench.py roofline legs) -------------------- */
/* generic conv through the tcgen05 implicit-GEMM kernel: x bf16 NHWC [n,h,w,cin] (device),
 * w f32 OIHW (host), scale/bias f32 [cout] (host, may be NULL = 1/0), res bf16 NHWC or NULL,
 * y bf16 NHWC [n,oh,ow,cout] (device).  up2: x is nearest-upsampled x2 and concatenated with
 * `skip` (bf16 NHWC [n,2h,2w,cskip] or NULL) before a 3x3/s1 conv (smp DecoderBlock).           */
/* hardware probe (dev tool, tools/umma_shift_probe.py): one MMA on a shifted window of a TMA-written SWIZZLE_128B tile */
WSI_API int wsi_debug_umma_shift(wsi_ctx* ctx, const void* A_dev, const void* B_dev, int r, int s, int pitch,
                         int use_base_offset, float* D_dev, void* stream);
WSI_API int wsi_debug_conv(wsi_ctx* ctx, const void* x, int n, int h, int w, int cin,
                   const float* wt, int cout, int ksize, int stride, int pad,
                   const float* scale, const float* bias, const void* res, int relu,
                   int up2, const void* skip, int cskip, void* y, void* stream);
/* the same conv in the fp32-emulated precision (WSI_PRECISION_FP32): x / res / skip / y are f32 NHWC device tensors
 * (split into three bf16 planes, convolved with the six plane products, merged back) — per-layer accuracy tests */
WSI_API int wsi_debug_conv_f32(wsi_ctx* ctx, const float* x, int n, int h, int w, int cin,
                       const float* wt, int cout, int ksize, int stride, int pad,
                       const float* scale, const float* bias, const float* res, int relu,
                       int up2, const float* skip, int cskip, float* y, void* stream);
/* K0 alone: tiles cut from the raster by the fused gather + normalise kernel.
 * norm_out: f32 [n,3,ph,pw] (device) = standard_augmentor(True) output, bit-exact
 * (utils/preprocessing.py:209-212); padded_out: bf16 [n,ph+6,pw+8,4] (device) = the stem operand
 * (RN-to-bf16 of the same values, zero border, 4th channel 0).  Either may be NULL.            */
WSI_API int wsi_debug_gather(wsi_ctx* ctx, const wsi_slide_desc* slide, const int32_t* tiles_xy, int n,
                     float* norm_out, void* padded_out, void* stream);
/* stem: tiles u8 raster -> gather/normalise -> 7x7/s2 conv + BN + ReLU; y bf16 NHWC [n,ph/2,pw/2,64] */
WSI_API int wsi_debug_stem(wsi_ctx* ctx, const wsi_slide_desc* slide, const int32_t* tiles_xy, int n,
                   const float* wt /*[64,3,7,7]*/, const float* scale, const float* bias, void* y,
                   void* stream);
/* max pool 3x3/s2/p1 on bf16 NHWC (resnets_shift.py:126) */
WSI_API int wsi_debug_maxpool(wsi_ctx* ctx, const void* x, int n, int h, int w, int c, void* y, void* stream);

/* ---- stage statistics (bench.py roofline legs) ------------------------------------------------ */
/* With option "stage_timing" = 1, wsi_run_slide brackets every stage with CUDA events on the
 * launching stream and accumulates, per stage name ("gather", "stem", "maxpool", "conv", "head",
 * "stitch", "finalise", "h2d", "d2h"): device milliseconds, kernel launches and algorithmic work
 * (FLOPs for stem/conv, bytes otherwise).  Reset with wsi_stage_reset.            */
WSI_API int wsi_stage_stats(wsi_ctx* ctx, const char* stage, double* ms_out, int64_t* launches_out, double* work_out);
WSI_API int wsi_stage_reset(wsi_ctx* ctx);
/* Per-kernel evidence (bench.py roofline table): with option "op_trace" = 1 (set before the first run) every conv launch
 * is bracketed by CUDA events on the launching stream.  idx walks the convs of the current network plan in execution
 * order (stem first); returns WSI_ERR_INVALID past the end.  flops / bytes = ALGORITHMIC work of one launch (2*MAC;
 * every operand read once + output written once), ms_per_launch = mean device time.  Adds ~2 events per kernel: not for
 * the timed region of a benchmark. */
WSI_API int wsi_op_stats(wsi_ctx* ctx, int idx, char* desc, int desc_cap, char* kernel, int kernel_cap, double* ms_per_launch,
                 double* flops, double* bytes, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* WSI_B200_H */
