// postproc.cu — tumour-bed post-processing of the heatmap / class mask on the device.
//
// Reference call sites (all on the host, numpy / cv2 / mahotas / skimage):
//   utils/eval.py:90-96                  tb = (argmax >= 2); cv2 MORPH_OPEN 20x20; convex_hull_image; bwperim; cv2.dilate 20x20
//   utils/eval.py:262-267                overlay = img * 0.75 + 255 * (heatmap > 255 * 0.99) * 0.25  -> np.uint8
//   paper_tools/overlay_tb_wsi.py:48-67  im = heatmap / 255 >= 0.9; MORPH_OPEN 30x30; dilate(bwperim(chull(im)), 20x20)
//   paper_tools/check_for_false_positives.py:62-72   im = heatmap >= 0.99 * 255; MORPH_OPEN 50x50; count_nonzero(im) / im.size > 0
//
// Kernels (all HBM-bound u8 streams; algorithmic bytes = 1 B read + 1 B written per pixel and pass):
//   lut_kernel            dst = lut[src] (any per-level threshold rule, evaluated on the host in float64 for the 256 levels)
//   morph_row / morph_col separable k x k min / max with cv2's window [x - k/2, x - k/2 + k) and BORDER_CONSTANT = "ignore
//                         outside" (cv2.morphologyDefaultBorderValue): bit-exact with cv2.erode / cv2.dilate / morphologyEx
//                         for np.ones((k, k)) kernels; 4 pixels per thread with byte-wise SIMD min / max (__vminu4)
//   row_extent_kernel     first / last non-zero column of every row (the only points a 2-D convex hull can use)
//   fill_rows_kernel      mask[y][x] = xl[y] <= x <= xr[y]   (rasterised hull)
//   bwperim_kernel        set pixels with a zero (or outside) 4-neighbour
//   overlay kernels       the two blends, truncating like np.uint8
#include <algorithm>

#include "kernels.cuh"

namespace wsi {

__global__ void __launch_bounds__(256) lut_kernel(const uint8_t* __restrict__ src, int64_t n, const uint8_t* __restrict__ lut, uint8_t* __restrict__ dst,
                                                   unsigned long long* __restrict__ nonzero) {
  __shared__ uint8_t s_lut[256];
  s_lut[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  unsigned cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t v = s_lut[src[i]];
    dst[i] = v;
    cnt += v != 0;
  }
  if (nonzero) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(nonzero, (unsigned long long)cnt);
  }
}

void launch_lut(const uint8_t* src, int64_t n, const uint8_t* lut_dev, uint8_t* dst, unsigned long long* nonzero, cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return;
  lut_kernel<<<(int)std::min<int64_t>(ceil_div(n, 256), 148 * 16), 256, 0, s>>>(src, n, lut_dev, dst, nonzero);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// ---- separable morphology ------------------------------------------------------------------------------------
// row pass: a CTA owns 1024 consecutive pixels of one row; the segment [x0 - k/2, x0 - k/2 + 1024 + k) is staged in
// shared memory (outside the image = the neutral element), thread t computes output pixels 4t .. 4t+3 as one u32.
constexpr int kMorphSeg = 1024;
constexpr int kMorphMaxK = 128;

template <bool MAX>
__global__ void __launch_bounds__(256) morph_row_kernel(const uint8_t* __restrict__ src, int64_t H, int64_t W, int k, uint8_t* __restrict__ dst) {
  __shared__ uint32_t s_w[(kMorphSeg + kMorphMaxK + 8) / 4];
  uint8_t* s_b = reinterpret_cast<uint8_t*>(s_w);
  const int64_t y = blockIdx.y;
  const int64_t x0 = (int64_t)blockIdx.x * kMorphSeg;
  const int64_t in0 = x0 - k / 2;                         // cv2: anchor = k / 2, window [x - anchor, x - anchor + k)
  const int n_in = kMorphSeg + k - 1;
  const uint8_t neutral = MAX ? 0 : 255;
  const uint8_t* row = src + y * W;
  for (int i = threadIdx.x; i < n_in + 4; i += blockDim.x) {
    const int64_t x = in0 + i;
    s_b[i] = (i < n_in && x >= 0 && x < W) ? row[x] : neutral;
  }
  __syncthreads();
  const int t = threadIdx.x;
  uint32_t acc = MAX ? 0u : 0xffffffffu;
  for (int jw = 0; 4 * jw < k; ++jw) {
    const uint32_t a = s_w[t + jw], b = s_w[t + jw + 1];
#pragma unroll
    for (int js = 0; js < 4; ++js) {
      if (4 * jw + js < k) {
        const uint32_t w = __funnelshift_r(a, b, 8 * js);
        acc = MAX ? __vmaxu4(acc, w) : __vminu4(acc, w);
      }
    }
  }
  const int64_t x = x0 + 4 * t;
  uint8_t* o = dst + y * W + x;
  if (x + 3 < W && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
    *reinterpret_cast<uint32_t*>(o) = acc;
  } else {
    for (int j = 0; j < 4; ++j)
      if (x + j < W) o[j] = (uint8_t)(acc >> (8 * j));
  }
}

// column pass: a CTA owns a block of 128 u32 word-columns x 64 output rows; rows [y0 - k/2, y0 - k/2 + 64 + k) staged.
constexpr int kMorphRows = 64;

template <bool MAX>
__global__ void __launch_bounds__(128) morph_col_kernel(const uint8_t* __restrict__ src, int64_t H, int64_t W, int k, uint8_t* __restrict__ dst,
                                                         unsigned long long* __restrict__ nonzero) {
  extern __shared__ uint32_t s_rows[];                    // [kMorphRows + k - 1][128]
  const int64_t xw = ((int64_t)blockIdx.x * 128 + threadIdx.x) * 4;       // first pixel of this thread's word
  const int64_t y0 = (int64_t)blockIdx.y * kMorphRows;
  const int64_t in0 = y0 - k / 2;
  const int n_in = kMorphRows + k - 1;
  const uint32_t neutral = MAX ? 0u : 0xffffffffu;
  const bool aligned = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 3) == 0) && (xw + 3 < W);
  for (int i = 0; i < n_in; ++i) {
    const int64_t y = in0 + i;
    uint32_t w = neutral;
    if (y >= 0 && y < H && xw < W) {
      const uint8_t* p = src + y * W + xw;
      if (aligned) {
        w = *reinterpret_cast<const uint32_t*>(p);
      } else {
        w = 0;
        for (int j = 0; j < 4; ++j) w |= (uint32_t)((xw + j < W) ? p[j] : (uint8_t)neutral) << (8 * j);
      }
    }
    s_rows[i * 128 + threadIdx.x] = w;
  }
  __syncthreads();
  const bool aligned_o = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0) && (xw + 3 < W);
  unsigned cnt = 0;
  for (int r = 0; r < kMorphRows; ++r) {
    const int64_t y = y0 + r;
    if (y >= H || xw >= W) break;
    uint32_t acc = neutral;
    for (int j = 0; j < k; ++j) {
      const uint32_t w = s_rows[(r + j) * 128 + threadIdx.x];
      acc = MAX ? __vmaxu4(acc, w) : __vminu4(acc, w);
    }
    uint8_t* o = dst + y * W + xw;
    if (aligned_o) {
      *reinterpret_cast<uint32_t*>(o) = acc;
      cnt += __popc(__vcmpne4(acc, 0u) & 0x01010101u);
    } else {
      for (int j = 0; j < 4; ++j)
        if (xw + j < W) { o[j] = (uint8_t)(acc >> (8 * j)); cnt += ((acc >> (8 * j)) & 0xffu) != 0; }
    }
  }
  if (nonzero) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(nonzero, (unsigned long long)cnt);
  }
}

// one k x k erosion (MAX = false) or dilation (MAX = true): src -> tmp (rows) -> dst (columns); src may equal dst
void launch_morph(const uint8_t* src, int64_t H, int64_t W, int k, bool is_max, uint8_t* tmp, uint8_t* dst, unsigned long long* nonzero,
                  cudaStream_t s, LaunchCounter* lc) {
  WSI_REQUIRE(k >= 1 && k <= kMorphMaxK, WSI_ERR_UNSUPPORTED, "morphology: kernel size %d outside [1, %d]", k, kMorphMaxK);
  if (H <= 0 || W <= 0) return;
  WSI_REQUIRE(H < 65536LL * kMorphRows, WSI_ERR_UNSUPPORTED, "morphology: image too tall");
  // grid.y is limited to 65535: taller images run the row pass in slabs
  for (int64_t yb = 0; yb < H; yb += 65535) {
    const int64_t hb = std::min<int64_t>(65535, H - yb);
    dim3 g((unsigned)ceil_div(W, kMorphSeg), (unsigned)hb);
    if (is_max) morph_row_kernel<true><<<g, 256, 0, s>>>(src + yb * W, hb, W, k, tmp + yb * W);
    else morph_row_kernel<false><<<g, 256, 0, s>>>(src + yb * W, hb, W, k, tmp + yb * W);
    CUDA_CHECK(cudaGetLastError());
    if (lc) lc->n++;
  }
  const size_t smem = (size_t)(kMorphRows + k - 1) * 128 * 4;
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(morph_col_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (kMorphRows + kMorphMaxK) * 128 * 4));
    CUDA_CHECK(cudaFuncSetAttribute(morph_col_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (kMorphRows + kMorphMaxK) * 128 * 4));
    configured = true;
  }
  dim3 gcol((unsigned)ceil_div(ceil_div(W, 4), 128), (unsigned)ceil_div(H, kMorphRows));
  if (is_max) morph_col_kernel<true><<<gcol, 128, smem, s>>>(tmp, H, W, k, dst, nonzero);
  else morph_col_kernel<false><<<gcol, 128, smem, s>>>(tmp, H, W, k, dst, nonzero);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// ---- convex hull support: first / last non-zero column per row (one warp per row) ----------------------------------
__global__ void __launch_bounds__(256) row_extent_kernel(const uint8_t* __restrict__ mask, int64_t H, int64_t W, int32_t* __restrict__ xmin,
                                                          int32_t* __restrict__ xmax) {
  const int64_t y = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (y >= H) return;
  const int lane = threadIdx.x & 31;
  const uint8_t* row = mask + y * W;
  int lo = INT32_MAX, hi = -1;
  for (int64_t x = lane; x < W; x += 32)
    if (row[x]) { lo = min(lo, (int)x); hi = max(hi, (int)x); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { xmin[y] = (hi < 0) ? -1 : lo; xmax[y] = hi; }
}

void launch_row_extent(const uint8_t* mask, int64_t H, int64_t W, int32_t* xmin, int32_t* xmax, cudaStream_t s, LaunchCounter* lc) {
  if (H <= 0) return;
  row_extent_kernel<<<(unsigned)ceil_div(H, 8), 256, 0, s>>>(mask, H, W, xmin, xmax);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

__global__ void __launch_bounds__(256) fill_rows_kernel(const int32_t* __restrict__ xl, const int32_t* __restrict__ xr, int64_t H, int64_t W,
                                                         uint8_t* __restrict__ out) {
  const int64_t total = H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t y = i / W;
    const int x = (int)(i - y * W);
    out[i] = (x >= __ldg(xl + y) && x <= __ldg(xr + y)) ? 1 : 0;
  }
}

void launch_fill_rows(const int32_t* xl, const int32_t* xr, int64_t H, int64_t W, uint8_t* out, cudaStream_t s, LaunchCounter* lc) {
  if (H * W <= 0) return;
  fill_rows_kernel<<<(int)std::min<int64_t>(ceil_div(H * W, 256), 148 * 16), 256, 0, s>>>(xl, xr, H, W, out);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// mahotas.bwperim(bw, n=4): set pixels with at least one zero 4-neighbour; outside the image counts as zero
__global__ void __launch_bounds__(256) bwperim_kernel(const uint8_t* __restrict__ bw, int64_t H, int64_t W, uint8_t* __restrict__ out) {
  const int64_t total = H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t y = i / W, x = i - y * W;
    uint8_t v = 0;
    if (bw[i]) {
      const bool n = (y > 0) && bw[i - W], s_ = (y + 1 < H) && bw[i + W], w = (x > 0) && bw[i - 1], e = (x + 1 < W) && bw[i + 1];
      v = (n && s_ && w && e) ? 0 : 1;
    }
    out[i] = v;
  }
}

void launch_bwperim(const uint8_t* bw, int64_t H, int64_t W, uint8_t* out, cudaStream_t s, LaunchCounter* lc) {
  if (H * W <= 0) return;
  bwperim_kernel<<<(int)std::min<int64_t>(ceil_div(H * W, 256), 148 * 16), 256, 0, s>>>(bw, H, W, out);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// utils/eval.py:262-267: np.uint8(img * 0.75 + 255 * (heat > 255 * 0.99) * 0.25): every term is exact in float64, so
// the truncated result is floor((3 * c + 255 * on) / 4).  `on` comes from the caller's 256-level rule table.
__global__ void __launch_bounds__(256) overlay_heat_kernel(const uint8_t* __restrict__ rgb, const uint8_t* __restrict__ heat, int64_t n_px,
                                                            const uint8_t* __restrict__ on_lut, uint8_t* __restrict__ out) {
  __shared__ uint8_t s_lut[256];
  s_lut[threadIdx.x] = on_lut[threadIdx.x];
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += (int64_t)gridDim.x * blockDim.x) {
    const int add = s_lut[heat[i]] ? 255 : 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[3 * i + c] = (uint8_t)((3 * (int)rgb[3 * i + c] + add) >> 2);
  }
}

void launch_overlay_heat(const uint8_t* rgb, const uint8_t* heat, int64_t n_px, const uint8_t* on_lut_dev, uint8_t* out, cudaStream_t s, LaunchCounter* lc) {
  if (n_px <= 0) return;
  overlay_heat_kernel<<<(int)std::min<int64_t>(ceil_div(n_px, 256), 148 * 16), 256, 0, s>>>(rgb, heat, n_px, on_lut_dev, out);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// paper_tools/overlay_tb_wsi.py:56-72: overlay = 0.65 * wsi + 0.35 * (heatmap * im); overlay[tb_perim] = 0; np.uint8(overlay)
// — float64 products and sum in that order (no FMA contraction), truncation
__global__ void __launch_bounds__(256) overlay_bed_kernel(const uint8_t* __restrict__ rgb, const uint8_t* __restrict__ heat,
                                                           const uint8_t* __restrict__ im, const uint8_t* __restrict__ perim, int64_t n_px,
                                                           uint8_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += (int64_t)gridDim.x * blockDim.x) {
    const bool zero = perim && perim[i];
    const double h = __dmul_rn(0.35, (double)(uint8_t)(heat[i] * (im ? im[i] : (uint8_t)1)));      // uint8 * uint8 wraps in numpy too
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double v = __dadd_rn(__dmul_rn(0.65, (double)rgb[3 * i + c]), h);
      out[3 * i + c] = zero ? (uint8_t)0 : (uint8_t)(int)v;
    }
  }
}

void launch_overlay_bed(const uint8_t* rgb, const uint8_t* heat, const uint8_t* im, const uint8_t* perim, int64_t n_px, uint8_t* out, cudaStream_t s,
                        LaunchCounter* lc) {
  if (n_px <= 0) return;
  overlay_bed_kernel<<<(int)std::min<int64_t>(ceil_div(n_px, 256), 148 * 16), 256, 0, s>>>(rgb, heat, im, perim, n_px, out);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

}  // namespace wsi
