// kernels.cu — HBM-bound stages (see kernels.cuh).  Reference lines cited per kernel.
#include "kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace wsi {

// =============================================================================================
// K0: fused tile gather + normalise + bf16 cast
//   reference: read_region(...).convert('RGB') (utils/dataset.py:175-178), ToTensor + Normalize
//   (utils/preprocessing.py:209-212), collate + .cuda() (utils/eval.py:192).
//   ((u8/255) - mean_c)/std_c has only 256 values per channel: the host builds the table with the
//   reference's exact fp32 op order and rounds it to bf16 once, so the kernel is a byte gather.
//   Output is the stem's operand layout: [n][ph+6][pw+8][4] bf16, zero border (3 top/left), ch 3 = 0.
// =============================================================================================
// ARITH (throughput mode only): bf16(fmaf(float(u8), a_c, b_c)) with a_c = 1 / (255 std_c), b_c = -mean_c / std_c instead
// of the table lookup — the host has verified that this reproduces ALL 768 table entries bit for bit (launch_gather),
// so the output is unchanged; it removes the three data-dependent shared-memory lookups per pixel (bank conflicts on
// tissue pixels) and leaves the byte loads and one 8-byte store.
struct GatherAffine { float a[3], b[3]; };

template <int PLANES, bool ARITH = false>
__global__ void __launch_bounds__(256) gather_kernel(const uint8_t* __restrict__ rgb, int64_t row_stride, int64_t row0,
                                                      const int32_t* __restrict__ tiles_xy, int n_tiles, int ph, int pw,
                                                      const float* __restrict__ lut, bf16* __restrict__ padded,
                                                      float* __restrict__ norm_out, int64_t plane_stride, GatherAffine aff) {
  // s_lut: the 3x256 possible outputs of Normalize(ToTensor(u8)) in fp32 (host-built, reference op order), rounded to
  // bf16 — and, for the fp32-emulated precision (planes == 3), the bf16 remainders b = rn(v - a), c = rn(v - a - b)
  __shared__ float s_f32[768];
  __shared__ uint16_t s_lut[PLANES][768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const float v = lut[i];
    s_f32[i] = v;
    float r = v;
#pragma unroll
    for (int j = 0; j < PLANES; ++j) {
      const bf16 h = __float2bfloat16_rn(r);
      s_lut[j][i] = __bfloat16_as_ushort(h);
      r -= __bfloat162float(h);
    }
  }
  __syncthreads();
  // Persistent blocks; ONE WARP per (tile, row) job.  (Round 1 gave a whole 256-thread block one row of 512 pixels: two
  // pixels per thread per job, so the per-job index arithmetic — a 64-bit division, the tile lookup, 64-bit addressing —
  // dominated: ncu counted ~93 instructions per pixel at 69 % issue utilisation.  A lane now carries pw / 32 pixels per
  // job and walks jobs without dividing.)
  const int64_t pitch = (int64_t)(pw + 8) * 4;
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int64_t n_jobs = (int64_t)n_tiles * ph;
  int64_t job = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int t = (int)(job / ph), r = (int)(job - (int64_t)t * ph);
  const int dt = warps_total / ph, dr = warps_total - dt * ph;          // job += warps_total  ==  (t, r) += (dt, dr) with carry
  for (; job < n_jobs; job += warps_total) {
    const int x0 = __ldg(tiles_xy + 2 * t), y0 = __ldg(tiles_xy + 2 * t + 1);
    const uint8_t* src = rgb + (int64_t)(y0 + r - row0) * row_stride + (int64_t)x0 * 3;
    bf16* dst = padded ? padded + ((int64_t)t * (ph + 6) + (r + 3)) * pitch + 3 * 4 : nullptr;
    float* q0 = norm_out ? norm_out + (int64_t)t * 3 * ph * pw + (int64_t)r * pw : nullptr;
    constexpr int kPx = 8;                                                 // pixels per lane in flight (24 byte loads): the kernel is
    for (int xb = 0; xb < pw; xb += 32 * kPx) {                            // latency-bound (ncu: issue 45 %, DRAM 30 %, nothing saturated)
      uint8_t c[kPx][3];
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
        const int x = xb + 32 * k + lane;
        if (x < pw) { c[k][0] = __ldg(src + 3 * x); c[k][1] = __ldg(src + 3 * x + 1); c[k][2] = __ldg(src + 3 * x + 2); }
      }
#pragma unroll
      for (int k = 0; k < kPx; ++k) {
        const int x = xb + 32 * k + lane;
        if (x >= pw) continue;
        if (ARITH) {
          // exact u8 -> float without a conversion instruction: 2^23 + c is representable, subtract 2^23
          const float f0 = __uint_as_float(0x4B000000u | c[k][0]) - 8388608.f, f1 = __uint_as_float(0x4B000000u | c[k][1]) - 8388608.f,
                      f2 = __uint_as_float(0x4B000000u | c[k][2]) - 8388608.f;
          const __nv_bfloat162 lo2 = __floats2bfloat162_rn(fmaf(f0, aff.a[0], aff.b[0]), fmaf(f1, aff.a[1], aff.b[1]));
          const __nv_bfloat162 hi2 = __floats2bfloat162_rn(fmaf(f2, aff.a[2], aff.b[2]), 0.f);
          uint2 o;
          o.x = *reinterpret_cast<const uint32_t*>(&lo2);
          o.y = *reinterpret_cast<const uint32_t*>(&hi2);
          *reinterpret_cast<uint2*>(dst + 4 * x) = o;
          continue;
        }
        if (dst) {
#pragma unroll
          for (int j = 0; j < PLANES; ++j) {
            const uint16_t v0 = s_lut[j][c[k][0]], v1 = s_lut[j][256 + c[k][1]], v2 = s_lut[j][512 + c[k][2]];
            uint2 o;
            o.x = (uint32_t)v0 | ((uint32_t)v1 << 16);
            o.y = (uint32_t)v2;
            *reinterpret_cast<uint2*>(dst + (int64_t)j * plane_stride + 4 * x) = o;
          }
        }
        if (q0) {
          const int64_t plane = (int64_t)ph * pw;
          q0[x] = s_f32[c[k][0]];
          q0[plane + x] = s_f32[256 + c[k][1]];
          q0[2 * plane + x] = s_f32[512 + c[k][2]];
        }
      }
    }
    t += dt; r += dr;
    if (r >= ph) { r -= ph; ++t; }
  }
}

// Does bf16(fmaf(v, a, b)) reproduce the reference-order table ((v / 255) - mean) / std rounded to bf16 for every u8 value?
// (It does for the ImageNet statistics of myargs.py:127-130; checked here once, so the arithmetic path can never change
// a result.)
static bool gather_affine(GatherAffine* aff) {
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int ch = 0; ch < 3; ++ch) {
    volatile float a = 1.0f / (255.0f * stdv[ch]);
    volatile float b = -mean[ch] / stdv[ch];
    aff->a[ch] = a; aff->b[ch] = b;
    for (int v = 0; v < 256; ++v) {
      volatile float t0 = (float)v / 255.0f;
      volatile float t1 = t0 - mean[ch];
      volatile float ref = t1 / stdv[ch];
      const float y = std::fma((float)v, (float)a, (float)b);
      if (f32_to_bf16_bits(y) != f32_to_bf16_bits(ref)) return false;
    }
  }
  return true;
}

void launch_gather(const uint8_t* rgb, int64_t row_stride, int64_t row0, const int32_t* tiles_xy_dev, int n, int ph,
                   int pw, const float* lut_dev, bf16* padded, float* norm_out, cudaStream_t s, LaunchCounter* lc, int planes,
                   int64_t plane_stride) {
  if (n <= 0) return;
  static GatherAffine aff;
  static const bool affine_ok = gather_affine(&aff) && getenv("WSI_GATHER_LUT") == nullptr;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div((int64_t)n * ph, 8), 148 * 8);
  // (a per-lane, bank-conflict-free copy of the table — 96 KB, 2 blocks per SM — was slower: 7.5 vs 6.1 ms per 8 280 tiles)
  if (planes == 3) gather_kernel<3><<<grid, 256, 0, s>>>(rgb, row_stride, row0, tiles_xy_dev, n, ph, pw, lut_dev, padded, norm_out, plane_stride, aff);
  else if (affine_ok && padded && !norm_out) gather_kernel<1, true><<<grid, 256, 0, s>>>(rgb, row_stride, row0, tiles_xy_dev, n, ph, pw, lut_dev, padded, norm_out, plane_stride, aff);
  else gather_kernel<1><<<grid, 256, 0, s>>>(rgb, row_stride, row0, tiles_xy_dev, n, ph, pw, lut_dev, padded, norm_out, plane_stride, aff);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// =============================================================================================
// K0r: tile resize of the scan_resize != 1 branch — `image.resize((tile_w, tile_h))`, utils/dataset.py:180-181
//   = PIL.Image.resize at its default filter on the pw x ph window of every tile (a standalone RGB image): antialiased
//   bicubic in 22-bit fixed point, horizontal pass first, u8 intermediate (libImaging/Resample.c, 8bpc paths).  The
//   per-axis windows / weights come from the host (wsi_resample_coeffs, planner.cpp) and are the same for every tile.
//   Integer arithmetic end to end: bit-exact with PIL.  Output: u8 tiles [n][th][tw][3], which the gather kernel then
//   reads as a raster of stacked tiles at origins (0, i * th) written to out_xy.
// =============================================================================================
__device__ __forceinline__ uint8_t resample_clip8(int v) {
  v >>= 22;                                             // PRECISION_BITS = 32 - 8 - 2; arithmetic shift like the C code
  return (uint8_t)min(max(v, 0), 255);
}

// Weight rows are padded to KS = a multiple of 4 taps (zeros), so a thread fetches the weights of its output sample with
// 16-byte loads and runs 4 taps x 3 channels per loop trip; taps past the window read the zero-filled tail of the staged row
// (h) or a clamped row (v) under a zero weight.  Both passes are instruction-bound (one byte load + one IMAD per tap and
// channel), not HBM-bound: the window bytes are read once, the half-resized tile is written and read once.
__global__ void __launch_bounds__(128) resample_h_kernel(const uint8_t* __restrict__ rgb, int64_t row_stride, int64_t row0,
                                                          const int32_t* __restrict__ tiles_xy, int n_tiles, int ph, int pw, int tw,
                                                          const int32_t* __restrict__ bounds, const int4* __restrict__ kk, int ks4,
                                                          uint8_t* __restrict__ tmp) {
  extern __shared__ __align__(16) uint8_t s_raw[];      // one window row (pw * 3 bytes) + 12 * ks4 + 8 zero bytes + 4 bytes of alignment slack
  const int row_bytes = pw * 3;
  const int64_t n_jobs = (int64_t)n_tiles * ph;
  for (int64_t job = blockIdx.x; job < n_jobs; job += gridDim.x) {
    const int t = (int)(job / ph), r = (int)(job - (int64_t)t * ph);
    const int x0 = __ldg(tiles_xy + 2 * t), y0 = __ldg(tiles_xy + 2 * t + 1);
    const uint8_t* src = rgb + (int64_t)(y0 + r - row0) * row_stride + (int64_t)x0 * 3;
    // The raster has no alignment contract: bytes up to the first 4-byte boundary of the source, whole words, tail bytes.
    // The row is staged at the same offset modulo 4, so the word loads are stored as words.
    const int head = (int)((4 - (reinterpret_cast<uintptr_t>(src) & 3)) & 3);
    uint8_t* s_row = s_raw + ((4 - head) & 3);
    __syncthreads();                                             // the previous job's readers are done
    {
      const int words = (row_bytes - head) >> 2;
      if (threadIdx.x < head) s_row[threadIdx.x] = __ldg(src + threadIdx.x);
      const uint32_t* src4 = reinterpret_cast<const uint32_t*>(src + head);
      uint32_t* dst4 = reinterpret_cast<uint32_t*>(s_row + head);
      for (int i = threadIdx.x; i < words; i += blockDim.x) dst4[i] = __ldg(src4 + i);
      const int done = head + 4 * words;                         // tail bytes, then the zero weights' bytes
      for (int i = done + threadIdx.x; i < row_bytes + 12 * ks4 + 8; i += blockDim.x) s_row[i] = (i < row_bytes) ? __ldg(src + i) : (uint8_t)0;
    }
    __syncthreads();
    uint8_t* dst = tmp + job * (int64_t)tw * 3;
    const int off = (4 - head) & 3;
    for (int xx = threadIdx.x; xx < tw; xx += blockDim.x) {
      // The window of 4 * ks4 pixels starts at an arbitrary byte of the staged row.  Byte loads would make the kernel
      // LSU-bound (36 LDS.U8 per output pixel at 9 taps): aligned words + a funnel shift put the window on word boundaries,
      // then one PRMT per byte (ALU pipe) feeds the IMAD.
      const int b = off + __ldg(bounds + 2 * xx) * 3;
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_raw) + (b >> 2);
      const int sh = (b & 3) * 8;
      const int4* k = kk + (int64_t)xx * ks4;
      int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
      uint32_t w0 = wp[0];
#define WSI_B(v, i) ((int)__byte_perm((v), 0u, 0x4440u + (i)))
      for (int q = 0; q < ks4; ++q, wp += 3) {
        const int4 c = __ldg(k + q);
        const uint32_t w1 = wp[1], w2 = wp[2], w3 = wp[3];
        const uint32_t A0 = __funnelshift_r(w0, w1, sh), A1 = __funnelshift_r(w1, w2, sh), A2 = __funnelshift_r(w2, w3, sh);
        a0 += WSI_B(A0, 0) * c.x + WSI_B(A0, 3) * c.y + WSI_B(A1, 2) * c.z + WSI_B(A2, 1) * c.w;
        a1 += WSI_B(A0, 1) * c.x + WSI_B(A1, 0) * c.y + WSI_B(A1, 3) * c.z + WSI_B(A2, 2) * c.w;
        a2 += WSI_B(A0, 2) * c.x + WSI_B(A1, 1) * c.y + WSI_B(A2, 0) * c.z + WSI_B(A2, 3) * c.w;
        w0 = w3;
      }
#undef WSI_B
      dst[3 * xx] = resample_clip8(a0); dst[3 * xx + 1] = resample_clip8(a1); dst[3 * xx + 2] = resample_clip8(a2);
    }
  }
}

// VEC: the half-resized rows are read as 32-bit words, four output bytes per thread (needs tw * 3 % 4 == 0)
template <bool VEC>
__global__ void __launch_bounds__(128) resample_v_kernel(const uint8_t* __restrict__ tmp, int n_tiles, int ph, int th, int tw,
                                                          const int32_t* __restrict__ bounds, const int4* __restrict__ kk, int ks4,
                                                          uint8_t* __restrict__ out, int32_t* __restrict__ out_xy) {
  const int64_t n_jobs = (int64_t)n_tiles * th;
  const int row = tw * 3;
  for (int64_t job = blockIdx.x; job < n_jobs; job += gridDim.x) {
    const int t = (int)(job / th), yy = (int)(job - (int64_t)t * th);
    const int ymin = __ldg(bounds + 2 * yy);
    const int4* k = kk + (int64_t)yy * ks4;
    const uint8_t* src = tmp + (int64_t)t * ph * row;
    uint8_t* dst = out + job * (int64_t)row;
    if (VEC) {
      for (int e = threadIdx.x; e < (row >> 2); e += blockDim.x) {
        int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21, a3 = 1 << 21;
        for (int q = 0; q < ks4; ++q) {
          const int4 c = __ldg(k + q);
          const int cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int y = min(ymin + 4 * q + j, ph - 1);           // past the window: weight 0
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src + (int64_t)y * row) + e);
            a0 += (int)(v & 255u) * cc[j]; a1 += (int)((v >> 8) & 255u) * cc[j]; a2 += (int)((v >> 16) & 255u) * cc[j]; a3 += (int)(v >> 24) * cc[j];
          }
        }
        const uint32_t o = (uint32_t)resample_clip8(a0) | ((uint32_t)resample_clip8(a1) << 8) | ((uint32_t)resample_clip8(a2) << 16) |
                           ((uint32_t)resample_clip8(a3) << 24);
        reinterpret_cast<uint32_t*>(dst)[e] = o;
      }
    } else {
      for (int e = threadIdx.x; e < row; e += blockDim.x) {
        int acc = 1 << 21;
        for (int q = 0; q < ks4; ++q) {
          const int4 c = __ldg(k + q);
          const int cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) acc += (int)src[(int64_t)min(ymin + 4 * q + j, ph - 1) * row + e] * cc[j];
        }
        dst[e] = resample_clip8(acc);
      }
    }
    if (yy == 0 && threadIdx.x == 0) { out_xy[2 * t] = 0; out_xy[2 * t + 1] = t * th; }
  }
}

void launch_resample_tiles(const uint8_t* rgb, int64_t row_stride, int64_t row0, const int32_t* tiles_xy_dev, int n, int ph, int pw, int th,
                           int tw, const int32_t* hb, const int32_t* hk, int hks, const int32_t* vb, const int32_t* vk, int vks,
                           uint8_t* tmp, uint8_t* out, int32_t* out_xy, cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return;
  // hks / vks: taps per weight row, padded by the caller to a multiple of 4; tmp / out: 4-byte aligned (device allocations)
  const unsigned gh = (unsigned)std::min<int64_t>((int64_t)n * ph, 148 * 16), gv = (unsigned)std::min<int64_t>((int64_t)n * th, 148 * 16);
  resample_h_kernel<<<gh, 128, (size_t)pw * 3 + 3 * hks + 16, s>>>(rgb, row_stride, row0, tiles_xy_dev, n, ph, pw, tw, hb,
                                                              reinterpret_cast<const int4*>(hk), hks / 4, tmp);
  CUDA_CHECK(cudaGetLastError());
  if ((tw * 3) % 4 == 0)
    resample_v_kernel<true><<<gv, 128, 0, s>>>(tmp, n, ph, th, tw, vb, reinterpret_cast<const int4*>(vk), vks / 4, out, out_xy);
  else
    resample_v_kernel<false><<<gv, 128, 0, s>>>(tmp, n, ph, th, tw, vb, reinterpret_cast<const int4*>(vk), vks / 4, out, out_xy);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n += 2;
}

// view: the test-time-augmentation views of predict_reg / predict_breastpathq (utils/eval.py:305-310, square tiles):
//   0 image, 1 image.transpose(2, 3), 2 image.flip(2), 3 image.transpose(2, 3).flip(3) — folded into the read address
__global__ void __launch_bounds__(256) pack_nchw_kernel(const float* __restrict__ x, int h, int w, int view, bf16* __restrict__ padded, int planes,
                                                        int64_t plane_stride) {
  const int t = blockIdx.x / h, r = blockIdx.x % h;
  const int64_t plane = (int64_t)h * w;
  const float* img = x + (int64_t)t * 3 * plane;
  const int64_t pitch = (int64_t)(w + 8) * 4;
  bf16* dst = padded + ((int64_t)t * (h + 6) + (r + 3)) * pitch + 3 * 4;
  for (int c = threadIdx.x; c < w; c += blockDim.x) {
    // output pixel (r, c) <- source pixel (sy, sx)
    int sy = r, sx = c;
    if (view == 1) { sy = c; sx = r; }
    else if (view == 2) { sy = h - 1 - r; }
    else if (view == 3) { sy = w - 1 - c; sx = r; }
    const float* src = img + (int64_t)sy * w + sx;
    float v0 = src[0], v1 = src[plane], v2 = src[2 * plane];
    for (int j = 0; j < planes; ++j) {          // planes == 3: bf16 expansion a + b + c of the fp32 value
      __nv_bfloat162 a = __floats2bfloat162_rn(v0, v1);
      __nv_bfloat162 b = __floats2bfloat162_rn(v2, 0.f);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(dst + (int64_t)j * plane_stride + 4 * c) = o;
      v0 -= __low2float(a); v1 -= __high2float(a); v2 -= __low2float(b);
    }
  }
}

void launch_pack_nchw(const float* x, int n, int h, int w, bf16* padded, cudaStream_t s, LaunchCounter* lc, int view, int planes,
                      int64_t plane_stride) {
  if (n <= 0) return;
  WSI_REQUIRE(view == 0 || h == w, WSI_ERR_UNSUPPORTED, "TTA views need square tiles (%dx%d)", h, w);
  pack_nchw_kernel<<<(unsigned)((int64_t)n * h), 256, 0, s>>>(x, h, w, view, padded, planes, plane_stride);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// TTA mean (utils/eval.py:311-334): acc = v0; acc += v1; acc += v2; acc += v3; out = acc / 4 — same fp32 order
__global__ void __launch_bounds__(256) tta_accumulate_kernel(float* __restrict__ acc, const float* __restrict__ v, int64_t n, int first,
                                                              float final_div) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float a = first ? v[i] : acc[i] + v[i];
    if (final_div != 0.f) a = a / final_div;
    acc[i] = a;
  }
}

void launch_tta_accumulate(float* acc, const float* v, int64_t n, bool first, float final_div, cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return;
  tta_accumulate_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), 1024), 256, 0, s>>>(acc, v, n, first ? 1 : 0, final_div);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// =============================================================================================
// max pool 3x3 / s2 / p1 (resnets_shift.py:126), NHWC bf16; one thread = 8 channels of one output px
// =============================================================================================
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__global__ void __launch_bounds__(256) maxpool_kernel(const bf16* __restrict__ x, int n, int h, int w, int c,
                                                       bf16* __restrict__ y) {
  const int oh = (h - 1) / 2 + 1, ow = (w - 1) / 2 + 1, cg = c / 8;
  const int64_t total = (int64_t)n * oh * ow * cg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    int64_t r = i / cg;
    const int ox = (int)(r % ow); r /= ow;
    const int oy = (int)(r % oh);
    const int b = (int)(r / oh);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int iy = 2 * oy + dy;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int ix = 2 * ox + dx;
        if (ix < 0 || ix >= w) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)b * h + iy) * w + ix) * c + g * 8));
        const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          m[2 * t] = fmaxf(m[2 * t], bf_lo(ww[t]));
          m[2 * t + 1] = fmaxf(m[2 * t + 1], bf_hi(ww[t]));
        }
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(m[2 * t], m[2 * t + 1]);
      o[t] = *reinterpret_cast<uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(y + (((int64_t)b * oh + oy) * ow + ox) * c + g * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// fp32-emulated precision: x is [n, h, w, 3c] = three bf16 planes [a | b | c] whose sum is the fp32 activation; the
// maximum is taken over the sums (exact in fp32) and its three planes are copied through
__global__ void __launch_bounds__(256) maxpool_split_kernel(const bf16* __restrict__ x, int n, int h, int w, int c, bf16* __restrict__ y) {
  const int oh = (h - 1) / 2 + 1, ow = (w - 1) / 2 + 1, cg = c / 8;
  const int64_t total = (int64_t)n * oh * ow * cg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    int64_t r = i / cg;
    const int ox = (int)(r % ow); r /= ow;
    const int oy = (int)(r % oh);
    const int b = (int)(r / oh);
    float m[8];
    uint16_t best[3][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; best[0][j] = 0xff80; best[1][j] = 0; best[2][j] = 0; }
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int iy = 2 * oy + dy;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int ix = 2 * ox + dx;
        if (ix < 0 || ix >= w) continue;
        const bf16* px = x + (((int64_t)b * h + iy) * w + ix) * (3 * c) + g * 8;
        uint4 v[3];
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) v[pl] = __ldg(reinterpret_cast<const uint4*>(px + pl * c));
        const uint16_t* e0 = reinterpret_cast<const uint16_t*>(&v[0]);
        const uint16_t* e1 = reinterpret_cast<const uint16_t*>(&v[1]);
        const uint16_t* e2 = reinterpret_cast<const uint16_t*>(&v[2]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float f = (__uint_as_float((uint32_t)e0[j] << 16) + __uint_as_float((uint32_t)e1[j] << 16)) + __uint_as_float((uint32_t)e2[j] << 16);
          if (f > m[j]) { m[j] = f; best[0][j] = e0[j]; best[1][j] = e1[j]; best[2][j] = e2[j]; }
        }
      }
    }
    bf16* o = y + (((int64_t)b * oh + oy) * ow + ox) * (3 * c) + g * 8;
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      uint4 q;
      q.x = (uint32_t)best[pl][0] | ((uint32_t)best[pl][1] << 16);
      q.y = (uint32_t)best[pl][2] | ((uint32_t)best[pl][3] << 16);
      q.z = (uint32_t)best[pl][4] | ((uint32_t)best[pl][5] << 16);
      q.w = (uint32_t)best[pl][6] | ((uint32_t)best[pl][7] << 16);
      *reinterpret_cast<uint4*>(o + pl * c) = q;
    }
  }
}

void launch_maxpool_split(const bf16* x, int n, int h, int w, int c, bf16* y, cudaStream_t s, LaunchCounter* lc) {
  const int64_t total = (int64_t)n * ((h - 1) / 2 + 1) * ((w - 1) / 2 + 1) * (c / 8);
  if (total <= 0) return;
  const int grid = (int)std::min<int64_t>(ceil_div(total, 256), 148 * 16);
  maxpool_split_kernel<<<grid, 256, 0, s>>>(x, n, h, w, c, y);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

void launch_maxpool(const bf16* x, int n, int h, int w, int c, bf16* y, cudaStream_t s, LaunchCounter* lc) {
  const int64_t total = (int64_t)n * ((h - 1) / 2 + 1) * ((w - 1) / 2 + 1) * (c / 8);
  if (total <= 0) return;
  const int grid = (int)std::min<int64_t>(ceil_div(total, 256), 148 * 16);
  maxpool_kernel<<<grid, 256, 0, s>>>(x, n, h, w, c, y);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// =============================================================================================
// global average pool + head MLP (models/models.py:32-38, :52-58; resnets_shift.py:206-208)
// one block (512 threads) per tile
// =============================================================================================
__global__ void __launch_bounds__(512) pool_head_kernel(const bf16* __restrict__ x4, int hw, int c,
                                                         const float* __restrict__ w1, const float* __restrict__ b1, int n1,
                                                         const float* __restrict__ w2, const float* __restrict__ b2, int n2,
                                                         float* __restrict__ feat_out, float* __restrict__ out, int planes) {
  __shared__ float s_feat[512];
  __shared__ float s_hid[512];
  const int t = blockIdx.x;
  const int cs = c * planes;                       // planes == 3: pixel = [a | b | c] bf16 planes of the fp32 activation
  const bf16* src = x4 + (int64_t)t * hw * cs;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < hw; ++i) {
      float v = __bfloat162float(src[(int64_t)i * cs + ch]);
      if (planes == 3) v = (v + __bfloat162float(src[(int64_t)i * cs + c + ch])) + __bfloat162float(src[(int64_t)i * cs + 2 * c + ch]);
      s += v;
    }
    s = s / (float)hw;
    s_feat[ch] = s;
    if (feat_out) feat_out[(int64_t)t * c + ch] = s;
  }
  __syncthreads();
  if (n1 <= 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int j = warp; j < n1; j += nwarp) {
    float s = 0.f;
    for (int ch = lane; ch < c; ch += 32) s = fmaf(s_feat[ch], __ldg(w1 + (int64_t)j * c + ch), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      s += b1[j];
      if (n2 > 0) s_hid[j] = fmaxf(s, 0.f);
      else out[(int64_t)t * n1 + j] = s;
    }
  }
  if (n2 <= 0) return;
  __syncthreads();
  for (int j = warp; j < n2; j += nwarp) {
    float s = 0.f;
    for (int ch = lane; ch < n1; ch += 32) s = fmaf(s_hid[ch], __ldg(w2 + (int64_t)j * n1 + ch), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[(int64_t)t * n2 + j] = s + b2[j];
  }
}

void launch_pool_head(const bf16* x4, int n, int hw, int c, const float* w1, const float* b1, int n1, const float* w2,
                      const float* b2, int n2, float* feat_out, float* out, cudaStream_t s, LaunchCounter* lc, int planes) {
  WSI_REQUIRE(c <= 512 && n1 <= 512, WSI_ERR_UNSUPPORTED, "pool_head: c=%d n1=%d", c, n1);
  if (n <= 0) return;
  pool_head_kernel<<<n, 512, 0, s>>>(x4, hw, c, w1, b1, n1, w2, b2, n2, feat_out, out, planes);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// fp32 NHWC [px][c] <-> three bf16 planes [px][a(c) | b(c) | c(c)] (per-layer tests of the fp32-emulated convs)
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, int64_t px, int c, bf16* __restrict__ y) {
  const int64_t total = px * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / c;
    const int ch = (int)(i - p * c);
    float r = x[i];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const bf16 h = __float2bfloat16_rn(r);
      y[p * 3 * c + (int64_t)j * c + ch] = h;
      r -= __bfloat162float(h);
    }
  }
}
__global__ void __launch_bounds__(256) merge_planes_kernel(const bf16* __restrict__ x, int64_t px, int c, float* __restrict__ y) {
  const int64_t total = px * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / c;
    const int ch = (int)(i - p * c);
    const bf16* q = x + p * 3 * c + ch;
    y[i] = (__bfloat162float(q[0]) + __bfloat162float(q[c])) + __bfloat162float(q[2 * c]);
  }
}
void launch_split_planes(const float* x, int64_t px, int c, bf16* y, cudaStream_t s, LaunchCounter* lc) {
  if (px * c <= 0) return;
  split_planes_kernel<<<(int)std::min<int64_t>(ceil_div(px * c, 256), 148 * 16), 256, 0, s>>>(x, px, c, y);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}
void launch_merge_planes(const bf16* x, int64_t px, int c, float* y, cudaStream_t s, LaunchCounter* lc) {
  if (px * c <= 0) return;
  merge_planes_kernel<<<(int)std::min<int64_t>(ceil_div(px * c, 256), 148 * 16), 256, 0, s>>>(x, px, c, y);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

__global__ void __launch_bounds__(256) nhwc4_to_nchw_kernel(const float4* __restrict__ x, int64_t plane, int64_t total, float* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / plane, r = i % plane;
    const float4 v = x[i];
    float* o = y + b * 4 * plane + r;
    o[0] = v.x; o[plane] = v.y; o[2 * plane] = v.z; o[3 * plane] = v.w;
  }
}
void launch_nhwc4_to_nchw(const float* x, int n, int h, int w, float* y, cudaStream_t s, LaunchCounter* lc) {
  const int64_t plane = (int64_t)h * w, total = plane * n;
  const int grid = (int)std::min<int64_t>(ceil_div(total, 256), 148 * 16);
  nhwc4_to_nchw_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(x), plane, total, y);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// =============================================================================================
// rect lookup shared by the stitch kernels: visit every sorted rect in [t0, t1) covering (X, Y)
// in a fixed (row, x) order -> deterministic sums, no atomics (each pixel has one owner thread).
// =============================================================================================
template <class F>
__device__ __forceinline__ void for_each_cover(const RectIndex& ri, int t0, int t1, int X, int Y, F&& f) {
  // rows with row_y in (Y - dy, Y]
  int lo = 0, hi = ri.R;
  while (lo < hi) {  // first row with row_y > Y - dy
    const int mid = (lo + hi) >> 1;
    if (__ldg(ri.row_y + mid) > Y - ri.dy) hi = mid; else lo = mid + 1;
  }
  for (int r = lo; r < ri.R; ++r) {
    const int ry = __ldg(ri.row_y + r);
    if (ry > Y) break;
    int a = max(__ldg(ri.row_start + r), t0), b = min(__ldg(ri.row_start + r + 1), t1);
    if (a >= b) continue;
    int l2 = a, h2 = b;
    while (l2 < h2) {  // first rect with tx > X - dx
      const int mid = (l2 + h2) >> 1;
      if (__ldg(ri.tx + mid) > X - ri.dx) h2 = mid; else l2 = mid + 1;
    }
    for (int i = l2; i < b; ++i) {
      const int tx = __ldg(ri.tx + i);
      if (tx > X) break;
      f(i, X - tx, Y - ry);
    }
  }
}

// =============================================================================================
// K6 + K7 fused, canvas-free (north_star (4)).
//   reference: pred[:, ty:ty+dy, tx:tx+dx] += pred_src[bj] into a float64 canvas (utils/eval.py:183,213-215), then
//   threshold_probs (utils/preprocessing.py:156-172: torch.softmax of the SUMMED logits in float64, per-class floor,
//   np.argmax = first maximum) and the heatmap (utils/eval.py:220-228: p[2]+p[3] | p[1], x mask, np.uint8(255 * h)
//   = truncation).
//   Gather formulation: one owner thread per canvas pixel sums the logits of every tile covering it, in the fixed
//   sorted (ty, tx) order, starting from zero, in DOUBLE like the reference's canvas (a sum of <= 64 fp32 values is
//   exact in double unless their magnitudes differ by > 2^25, so the result does not depend on the order the
//   reference's shuffling DataLoader presents the tiles in), then softmax / floor / argmax / heat in double, and writes
//   the u8 mask + u8 heatmap (+ optional fp32 summed logits and probabilities).  No canvas, no atomics: every logit is
//   read exactly once (T*P*16 B) and every output pixel written once (2 B) — SURVEY §8d's fused figure.
//   A CTA owns 256 consecutive pixels of one canvas row: the covering tile rows and the candidate tiles of each are
//   found ONCE per CTA (uniform binary searches on the sorted index), not once per pixel.
// =============================================================================================
struct PixelOut { uint8_t cls, heat; double p[4]; };

__device__ __forceinline__ PixelOut finalise_pixel(double l0, double l1, double l2, double l3, double maskv, const double* cp, int heat_mode) {
  PixelOut o;
  const double mx = fmax(fmax(l0, l1), fmax(l2, l3));
  const double e0 = exp(l0 - mx), e1 = exp(l1 - mx), e2 = exp(l2 - mx), e3 = exp(l3 - mx);
  const double sum = ((e0 + e1) + e2) + e3;                 // torch's vec_softmax: sequential over the class dim
  double p0 = e0 / sum, p1 = e1 / sum, p2 = e2 / sum, p3 = e3 / sum;
  if (p0 < cp[0]) p0 = 0.0;
  if (p1 < cp[1]) p1 = 0.0;
  if (p2 < cp[2]) p2 = 0.0;
  if (p3 < cp[3]) p3 = 0.0;
  int am = 0; double best = p0;
  if (p1 > best) { best = p1; am = 1; }
  if (p2 > best) { best = p2; am = 2; }
  if (p3 > best) { best = p3; am = 3; }
  const double h = (heat_mode == 1) ? p1 : (p2 + p3);
  const double hv = 255.0 * (maskv * h);
  o.cls = (uint8_t)am;
  o.heat = (uint8_t)(int)fmin(hv, 255.0);
  o.p[0] = p0; o.p[1] = p1; o.p[2] = p2; o.p[3] = p3;
  return o;
}

__device__ __forceinline__ void write_pixel(const FinaliseArgs& a, int64_t idx, int64_t plane, double l0, double l1, double l2, double l3,
                                            const PixelOut& o) {
  a.classes[idx] = o.cls;
  a.heatmap[idx] = o.heat;
  if (a.canvas_out) {
    a.canvas_out[idx] = (float)l0; a.canvas_out[plane + idx] = (float)l1; a.canvas_out[2 * plane + idx] = (float)l2; a.canvas_out[3 * plane + idx] = (float)l3;
  }
  if (a.probs_out) {
    a.probs_out[idx] = (float)o.p[0]; a.probs_out[plane + idx] = (float)o.p[1]; a.probs_out[2 * plane + idx] = (float)o.p[2];
    a.probs_out[3 * plane + idx] = (float)o.p[3];
  }
}

// x origin of sorted rect i of row q
__device__ __forceinline__ int rect_tx(const RectIndex& ri, const RowInfo& q, int i) {
  return (i < q.a + q.nreg) ? q.tx0 + (i - q.a) * q.step : __ldg(ri.tx + i);
}
// first rect i in [a2, b2) of row q with tx(i) > lim: O(1) on the regular prefix, a short scan / bisection on the rest
__device__ __forceinline__ int row_first_right_of(const RectIndex& ri, const RowInfo& q, int lim, int a2, int b2) {
  int i = q.a;
  if (lim >= q.tx0) i = q.a + ((q.step > 0) ? min((lim - q.tx0) / q.step + 1, q.nreg) : 1);
  i = min(max(i, a2), b2);
  if (i >= q.a + q.nreg) {                       // irregular tail
    if (b2 - i > 8) {
      int lo = i, hi = b2;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(ri.tx + mid) > lim) hi = mid; else lo = mid + 1;
      }
      i = lo;
    } else {
      while (i < b2 && __ldg(ri.tx + i) <= lim) ++i;
    }
  }
  return i;
}

// visit, in sorted order, the tiles of [t_lo, t_hi) in tile rows [r_lo, r_hi) that cover row Y and may cover columns
// [X0, X0 + span): f(i, tx, oy).  All control flow is uniform across the CTA; a regular tile grid needs one 32-byte
// RowInfo load per tile row and no other index access (the first version bisected row_y and tx per CTA: ~70 dependent
// L1 round trips, ~3 000 cycles of latency per 256 pixels).
template <class F>
__device__ __forceinline__ void for_each_candidate(const RectIndex& ri, int t_lo, int t_hi, int r_lo, int r_hi, int X0, int span, int Y, F&& f) {
  for (int r = r_lo; r < r_hi; ++r) {
    const int4 w0 = __ldg(reinterpret_cast<const int4*>(ri.rows + r));
    const int4 w1 = __ldg(reinterpret_cast<const int4*>(ri.rows + r) + 1);
    RowInfo q;
    q.ry = w0.x; q.a = w0.y; q.b = w0.z; q.tx0 = w0.w; q.step = w1.x; q.nreg = w1.y;
    if (q.ry > Y) break;
    if (q.ry + ri.dy <= Y) continue;
    const int a2 = max(q.a, t_lo), b2 = min(q.b, t_hi);
    if (a2 >= b2) continue;
    for (int i = row_first_right_of(ri, q, X0 - ri.dx, a2, b2); i < b2; ++i) {
      const int tx = rect_tx(ri, q, i);
      if (tx >= X0 + span) break;
      f(i, tx, Y - q.ry);
    }
  }
}

// seg: logits of sorted tile i live in slot i % ring_cap of `ring` (f32 [ring_cap][dy][dx][4]); rows [y0, y0 + gridDim.x).
// A CTA owns PX * 256 consecutive pixels of one canvas row, a thread PX of them 256 apart (coalesced per access).
// Measured (12k x 12k slide, same box, T*P*16 + 2*S algorithmic bytes): PX = 1 / 2 / 4 -> 2.20 / 2.35 / 1.55 TB/s; an
// L1-prefetch pass ahead of the accumulate pass made every variant slower; O(1) tile lookup (RowInfo) -> 2.62 TB/s; a
// variant that staged every covering run in shared memory with cp.async.bulk (192 KB in flight per SM, no load
// scoreboard) was SLOWER (2.27 TB/s: one thread walks the candidates per CTA) — the stage is not starved for bytes in
// flight.  Cost split by elimination: float accumulation instead of double -0.9 ms of 13.3, float finalise -2.0 ms,
// both -3.4 ms (3.55 TB/s): the float64 arithmetic that makes the stage bit-compatible with the reference's float64
// canvas costs a quarter of it; the rest is the 16-way gather itself (28 address streams per CTA, 92 strip launches).
// UP: the scan_resize != 1 branch — F.interpolate(pred_src, (tile_h * r, tile_w * r)) at its default mode 'nearest'
// (utils/eval.py:202-206) before the slice-add: canvas pixel (oy, ox) of a rectangle reads logit (oy / r, ox / r) of the
// (dy / r) x (dx / r) tile (integer r: floor(dst * in / out) == dst / r).
template <int PX, bool UP>
__global__ void __launch_bounds__(256) stitch_finalise_seg_kernel(RectIndex ri, const float4* __restrict__ ring, int ring_cap, int t_lo, int t_hi,
                                                                   int r_lo, int r_hi, int y0, FinaliseArgs a) {
  const int Y = y0 + blockIdx.x;
  const int X0 = blockIdx.y * (256 * PX);
  const int Xt = X0 + threadIdx.x;
  const int lw = UP ? ri.dx / ri.up : ri.dx;                       // logit tile width
  const int64_t tile_px = UP ? (int64_t)lw * (ri.dy / ri.up) : (int64_t)ri.dx * ri.dy;
  double s[PX][4];
#pragma unroll
  for (int k = 0; k < PX; ++k) s[k][0] = s[k][1] = s[k][2] = s[k][3] = 0.0;
  for_each_candidate(ri, t_lo, t_hi, r_lo, r_hi, X0, 256 * PX, Y, [&](int i, int tx, int oy) {
    const float4* src = ring + (int64_t)(i % ring_cap) * tile_px + (int64_t)(UP ? oy / ri.up : oy) * lw;
    float4 v[PX];
    bool hit[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) {
      const int ox = Xt + 256 * k - tx;
      hit[k] = (ox >= 0) && (ox < ri.dx) && (Xt + 256 * k < a.W2);
      if (hit[k]) v[k] = __ldg(src + (UP ? ox / ri.up : ox));
    }
#pragma unroll
    for (int k = 0; k < PX; ++k)
      if (hit[k]) { s[k][0] += (double)v[k].x; s[k][1] += (double)v[k].y; s[k][2] += (double)v[k].z; s[k][3] += (double)v[k].w; }
  });
  const int64_t plane = (a.own1 - a.own0) * a.W2;
#pragma unroll
  for (int k = 0; k < PX; ++k) {
    const int X = Xt + 256 * k;
    if (X >= a.W2) break;
    const int64_t idx = (int64_t)(Y - a.own0) * a.W2 + X;
    const double mv = a.mask ? (double)a.mask[idx] : 1.0;
    const PixelOut o = finalise_pixel(s[k][0], s[k][1], s[k][2], s[k][3], mv, a.class_probs, a.heat_mode);
    write_pixel(a, idx, plane, s[k][0], s[k][1], s[k][2], s[k][3], o);
  }
}

void launch_stitch_finalise_seg(const RectIndex& ri, const float4* ring, int ring_cap, int t_lo, int t_hi, int r_lo, int r_hi, int64_t y0, int64_t y1,
                                const FinaliseArgs& a, cudaStream_t s, LaunchCounter* lc) {
  if (y1 <= y0 || a.W2 <= 0) return;
  constexpr int px = 2;
  dim3 grid((unsigned)(y1 - y0), (unsigned)ceil_div(a.W2, 256 * px));
  if (ri.up > 1) stitch_finalise_seg_kernel<px, true><<<grid, 256, 0, s>>>(ri, ring, ring_cap > 0 ? ring_cap : 1, t_lo, t_hi, r_lo, r_hi, (int)y0, a);
  else stitch_finalise_seg_kernel<px, false><<<grid, 256, 0, s>>>(ri, ring, ring_cap > 0 ? ring_cap : 1, t_lo, t_hi, r_lo, r_hi, (int)y0, a);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// cls: pred_src [C] broadcast over the tile rectangle (utils/eval.py:210-215); tile_logits f32 [T][4] in sorted order.
// The summed logits are piecewise constant: they only change where a tile starts or ends.  The host cuts the canvas at
// every tile edge (x breakpoints Bx, y breakpoints By) and gives every column / row its cell index; ONE thread per CELL
// sums the covering tiles (double, sorted order) and runs the softmax / floor / argmax once, then the paint kernel writes
// 2 bytes per pixel from the cell table.  (The first version re-did the sum and the double-precision softmax per PIXEL:
// 189 ms at 4x coverage and 804 ms at 64x coverage on a 50k x 50k canvas for 5 GB of output — 0.1-0.4 % of HBM speed.)
struct ClsCell { double s[4]; double p[4]; double h; uint8_t cls; uint8_t pad[7]; };      // 80 bytes

__global__ void __launch_bounds__(128) cls_cells_kernel(RectIndex ri, const float4* __restrict__ tile_logits, int T, const int32_t* __restrict__ bx, int nbx,
                                                         const int32_t* __restrict__ by, int nby, FinaliseArgs a, ClsCell* __restrict__ cells) {
  const int64_t n = (int64_t)nbx * nby;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(c / nbx), i = (int)(c - (int64_t)j * nbx);
    const int X = __ldg(bx + i), Y = __ldg(by + j);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for_each_cover(ri, 0, T, X, Y, [&](int t, int, int) {
      const float4 v = __ldg(tile_logits + t);
      s0 += (double)v.x; s1 += (double)v.y; s2 += (double)v.z; s3 += (double)v.w;
    });
    const PixelOut o = finalise_pixel(s0, s1, s2, s3, 1.0, a.class_probs, a.heat_mode);
    ClsCell q;
    q.s[0] = s0; q.s[1] = s1; q.s[2] = s2; q.s[3] = s3;
    q.p[0] = o.p[0]; q.p[1] = o.p[1]; q.p[2] = o.p[2]; q.p[3] = o.p[3];
    q.h = (a.heat_mode == 1) ? o.p[1] : (o.p[2] + o.p[3]);
    q.cls = o.cls;
    cells[c] = q;
  }
}

__global__ void __launch_bounds__(256) cls_paint_kernel(const ClsCell* __restrict__ cells, const int32_t* __restrict__ cellx, const int32_t* __restrict__ celly,
                                                         int nbx, int y0, FinaliseArgs a) {
  const int Y = y0 + blockIdx.x;
  const int X = blockIdx.y * 256 + threadIdx.x;
  if (X >= a.W2) return;
  const ClsCell* q = cells + (int64_t)__ldg(celly + (Y - (int)a.own0)) * nbx + __ldg(cellx + X);
  const int64_t idx = (int64_t)(Y - a.own0) * a.W2 + X, plane = (a.own1 - a.own0) * a.W2;
  const double mv = a.mask ? (double)a.mask[idx] : 1.0;
  const double hv = 255.0 * (mv * q->h);
  a.classes[idx] = q->cls;
  a.heatmap[idx] = (uint8_t)(int)fmin(hv, 255.0);
  if (a.canvas_out) {
    a.canvas_out[idx] = (float)q->s[0]; a.canvas_out[plane + idx] = (float)q->s[1]; a.canvas_out[2 * plane + idx] = (float)q->s[2];
    a.canvas_out[3 * plane + idx] = (float)q->s[3];
  }
  if (a.probs_out) {
    a.probs_out[idx] = (float)q->p[0]; a.probs_out[plane + idx] = (float)q->p[1]; a.probs_out[2 * plane + idx] = (float)q->p[2];
    a.probs_out[3 * plane + idx] = (float)q->p[3];
  }
}

size_t cls_cell_bytes() { return sizeof(ClsCell); }

void launch_stitch_finalise_cls(const RectIndex& ri, const float4* tile_logits, int T, const int32_t* bx, int nbx, const int32_t* by, int nby,
                                const int32_t* cellx, const int32_t* celly, void* cells, const FinaliseArgs& a, cudaStream_t s, LaunchCounter* lc) {
  if (a.own1 <= a.own0 || a.W2 <= 0) return;
  const int64_t n = (int64_t)nbx * nby;
  cls_cells_kernel<<<(int)std::min<int64_t>(ceil_div(n, 128), 148 * 16), 128, 0, s>>>(ri, tile_logits, T, bx, nbx, by, nby, a, static_cast<ClsCell*>(cells));
  CUDA_CHECK(cudaGetLastError());
  dim3 grid((unsigned)(a.own1 - a.own0), (unsigned)ceil_div(a.W2, 256));
  cls_paint_kernel<<<grid, 256, 0, s>>>(static_cast<const ClsCell*>(cells), cellx, celly, nbx, (int)a.own0, a);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n += 2;
}

__global__ void __launch_bounds__(256) counts_kernel(RectIndex ri, int T, int64_t W2, int64_t own0, int64_t own1, int32_t* __restrict__ counts) {
  const int64_t plane = (own1 - own0) * W2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < plane; idx += (int64_t)gridDim.x * blockDim.x) {
    const int Y = (int)(own0 + idx / W2), X = (int)(idx % W2);
    int n = 0;
    for_each_cover(ri, 0, T, X, Y, [&](int, int, int) { ++n; });
    counts[idx] = n;
  }
}

void launch_counts(const RectIndex& ri, int T, int64_t W2, int64_t own0, int64_t own1, int32_t* counts, cudaStream_t s, LaunchCounter* lc) {
  const int64_t plane = (own1 - own0) * W2;
  if (plane <= 0) return;
  const int grid = (int)std::min<int64_t>(ceil_div(plane, 256), 148 * 32);
  counts_kernel<<<grid, 256, 0, s>>>(ri, T, W2, own0, own1, counts);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// =============================================================================================
// synthetic H&E slide (integer-only twin of wsi_segmentation_pipeline_b200/synth.py)
// =============================================================================================
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t hash2(uint32_t seed, uint32_t a, uint32_t b, uint32_t salt) {
  const uint32_t k = seed * 0xC2B2AE3Du + salt * 0x27D4EB2Fu;
  return mix32((a * 0x9E3779B1u) ^ (b * 0x85EBCA77u) ^ k);
}
__device__ __forceinline__ uint32_t value_noise(uint32_t seed, int64_t y, int64_t x, int shift, uint32_t salt) {
  const uint32_t cell = 1u << shift;
  const uint32_t gy = (uint32_t)(y >> shift), gx = (uint32_t)(x >> shift);
  const uint32_t fy = (uint32_t)(y & (cell - 1)), fx = (uint32_t)(x & (cell - 1));
  const uint32_t v00 = hash2(seed, gx, gy, salt) & 255u, v10 = hash2(seed, gx + 1, gy, salt) & 255u;
  const uint32_t v01 = hash2(seed, gx, gy + 1, salt) & 255u, v11 = hash2(seed, gx + 1, gy + 1, salt) & 255u;
  const uint32_t top = v00 * (cell - fx) + v10 * fx, bot = v01 * (cell - fx) + v11 * fx;
  const uint32_t val = top * (cell - fy) + bot * fy;
  return val >> (2 * shift - 8);
}

__global__ void __launch_bounds__(256) synth_kernel(int64_t iw, uint32_t seed, int64_t y0, int64_t y1, const uint8_t* __restrict__ lut,
                                                     uint8_t* __restrict__ rgb, int64_t row_stride, uint8_t* __restrict__ mask) {
  __shared__ uint8_t s_lut[16 * 8 * 3];
  for (int i = threadIdx.x; i < 16 * 8 * 3; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const int64_t total = (y1 - y0) * iw;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t y = y0 + idx / iw, x = idx % iw;
    const uint32_t f1 = value_noise(seed, y, x, 9, 11), f2 = value_noise(seed, y, x, 7, 23);
    const bool tissue = ((f1 * 3u + f2) >> 2) > 30200u;
    const uint32_t hp = hash2(seed, (uint32_t)x, (uint32_t)y, 37);
    uint8_t* o = rgb + (y - y0) * row_stride + x * 3;
    if (mask) mask[(y - y0) * iw + x] = tissue ? 1 : 0;
    int out[3];
    if (tissue) {
      const uint32_t ef = value_noise(seed, y, x, 6, 41);
      const uint32_t e_lvl = min((ef >> 12) + ((hp >> 28) & 3u), 15u);
      const int64_t cy = y / 12, cx = x / 12;
      uint32_t h_lvl = 0;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int64_t gy = cy + dy, gx = cx + dx;
          const uint32_t hc = hash2(seed, (uint32_t)gx, (uint32_t)gy, 53);
          if ((hc & 7u) >= 5u) continue;
          const int64_t ox = (hc >> 3) % 12u, oy = (hc >> 9) % 12u;
          const int64_t r = 3 + ((hc >> 15) & 3u);
          const int64_t ddx = x - (gx * 12 + ox), ddy = y - (gy * 12 + oy);
          if (ddx * ddx + ddy * ddy <= r * r) h_lvl = max(h_lvl, 1u + ((hc >> 17) % 7u));
        }
      const uint8_t* base = s_lut + (e_lvl * 8 + h_lvl) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint32_t bits = (hp >> (6 * c)) & 63u;
        out[c] = (int)base[c] + (int)((bits & 7u) + (bits >> 3)) - 7;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) out[c] = 240 + (int)((hp >> (4 * c + 8)) % 9u) - 4;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = (uint8_t)min(max(out[c], 0), 255);
  }
}

void launch_synth(int64_t ih, int64_t iw, uint32_t seed, int64_t y0, int64_t y1, const uint8_t* lut_dev, uint8_t* rgb,
                  int64_t row_stride, uint8_t* mask, cudaStream_t s, LaunchCounter* lc) {
  (void)ih;
  const int64_t total = (y1 - y0) * iw;
  if (total <= 0) return;
  const int grid = (int)std::min<int64_t>(ceil_div(total, 256), 148 * 32);
  synth_kernel<<<grid, 256, 0, s>>>(iw, seed, y0, y1, lut_dev, rgb, row_stride, mask);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// A9 (predict_wsis, utils/eval.py:66-81): per-class cv2.resize(pred[c], level-2 size) of the summed-logit canvas
// — INTER_LINEAR as OpenCV computes it for CV_64F: source coordinate (d + 0.5) * scale - 0.5 in double, floor +
// fraction, clamped to the border (fraction 0 there), weights 1 - f and f, horizontal pass then vertical pass —
// followed by np.argmax over the classes (first maximum).  OpenCV blends in double, this kernel in fp32 (the
// canvas is fp32): values agree to ~1e-6 relative, the argmax except at near-ties.  HBM-bound: 16 B read per
// source pixel (each read once when downscaling), 17 B written per destination pixel.
__global__ void __launch_bounds__(256) resize_argmax_kernel(const float* __restrict__ src, int64_t H, int64_t W, int64_t H2, int64_t W2,
                                                             double scale_x, double scale_y, uint8_t* __restrict__ classes,
                                                             float* __restrict__ pred) {
  const int64_t plane_s = H * W, plane_d = H2 * W2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < plane_d; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t dy = idx / W2, dx = idx - dy * W2;
    const double fxd = ((double)dx + 0.5) * scale_x - 0.5, fyd = ((double)dy + 0.5) * scale_y - 0.5;
    int64_t sx = (int64_t)floor(fxd);
    float fx = (float)(fxd - (double)sx);
    if (sx < 0) { sx = 0; fx = 0.f; }
    if (sx >= W - 1) { sx = W - 1; fx = 0.f; }
    const int64_t sy = (int64_t)floor(fyd);
    const float fy = (float)(fyd - (double)sy);
    const int64_t y0 = min(max(sy, (int64_t)0), H - 1), y1 = min(max(sy + 1, (int64_t)0), H - 1);
    const int64_t x1 = min(sx + 1, W - 1);
    const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
    float best = 0.f;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float* pl = src + (int64_t)c * plane_s;
      const float h0 = __fadd_rn(__fmul_rn(__ldg(pl + y0 * W + sx), a0), __fmul_rn(__ldg(pl + y0 * W + x1), a1));
      const float h1 = __fadd_rn(__fmul_rn(__ldg(pl + y1 * W + sx), a0), __fmul_rn(__ldg(pl + y1 * W + x1), a1));
      const float v = __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, b1));
      if (pred) pred[(int64_t)c * plane_d + idx] = v;
      if (c == 0 || v > best) { best = v; arg = c; }
    }
    classes[idx] = (uint8_t)arg;
  }
}

void launch_resize_argmax(const float* src, int64_t H, int64_t W, int64_t H2, int64_t W2, uint8_t* classes, float* pred, cudaStream_t s,
                          LaunchCounter* lc) {
  const int64_t plane = H2 * W2;
  if (plane <= 0) return;
  const int grid = (int)std::min<int64_t>(ceil_div(plane, 256), 148 * 32);
  // OpenCV: inv_scale = dsize / ssize (double), scale = 1 / inv_scale
  const double scale_x = 1.0 / ((double)W2 / (double)W), scale_y = 1.0 / ((double)H2 / (double)H);
  resize_argmax_kernel<<<grid, 256, 0, s>>>(src, H, W, H2, W2, scale_x, scale_y, classes, pred);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// A12 (find_nuclei mode='hsv', utils/preprocessing.py:74-110): mask = HSV saturation > mu_percent.  skimage's rgb2hsv
// computes S = (max - min) / max on the float64 image u8/255 (0 where max == min); the comparison only depends on
// (max, min), so the host evaluates exactly those float64 operations for all 256 x 256 pairs into a bit table and
// the kernel is a max/min + lookup: bit-exact with the float64 formula by construction.  HBM-bound: 3 B read + 1 B
// written per pixel.
__global__ void __launch_bounds__(256) find_nuclei_kernel(const uint8_t* __restrict__ rgb, int64_t row_stride, int64_t H, int64_t W,
                                                           const uint32_t* __restrict__ lut_bits, uint8_t* __restrict__ mask) {
  __shared__ uint32_t s_lut[2048];      // 65536 bits: index mx * 256 + mn
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) s_lut[i] = lut_bits[i];
  __syncthreads();
  const int64_t total = H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t y = idx / W, x = idx - y * W;
    const uint8_t* px = rgb + y * row_stride + 3 * x;
    const int r = px[0], g = px[1], b = px[2];
    const int mx = max(r, max(g, b)), mn = min(r, min(g, b));
    const int bit = mx * 256 + mn;
    mask[idx] = (uint8_t)((s_lut[bit >> 5] >> (bit & 31)) & 1u);
  }
}

void launch_find_nuclei(const uint8_t* rgb, int64_t row_stride, int64_t H, int64_t W, const uint32_t* lut_bits, uint8_t* mask, cudaStream_t s,
                        LaunchCounter* lc) {
  const int64_t total = H * W;
  if (total <= 0) return;
  const int grid = (int)std::min<int64_t>(ceil_div(total, 256), 148 * 16);
  find_nuclei_kernel<<<grid, 256, 0, s>>>(rgb, row_stride, H, W, lut_bits, mask);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// A1/A12 (isforeground, utils/preprocessing.py:60-71, called per candidate tile from utils/dataset.py:147-166):
// count_nonzero(mask[yp:yp+dy, xp:xp+dx]) with numpy's silent clipping at the mask edge; one CTA per candidate window,
// 4-byte loads over the aligned body of each row.  counts[i] = nonzero pixels, sizes[i] = clipped window size.
__global__ void __launch_bounds__(256) window_count_kernel(const uint8_t* __restrict__ mask, int64_t mh, int64_t mw,
                                                            const int64_t* __restrict__ win /* [n][2] = (xp, yp) */, int64_t dx, int64_t dy,
                                                            uint32_t* __restrict__ counts, int64_t* __restrict__ sizes) {
  const int64_t i = blockIdx.x;
  const int64_t xp = win[2 * i], yp = win[2 * i + 1];
  const int64_t x0 = min(xp, mw), x1 = min(xp + dx, mw), y0 = min(yp, mh), y1 = min(yp + dy, mh);
  const int64_t w = x1 - x0, h = y1 - y0;
  uint32_t cnt = 0;
  if (w > 0 && h > 0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int64_t y = y0 + warp; y < y1; y += nwarps) {
      const uint8_t* row = mask + y * mw;
      const uintptr_t a0 = reinterpret_cast<uintptr_t>(row + x0);
      const int64_t head = min((int64_t)((4 - (int64_t)(a0 & 3)) & 3), w);
      const int64_t body = (w - head) >> 2;
      if (lane < head) cnt += row[x0 + lane] != 0;
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(row + x0 + head);
      for (int64_t k = lane; k < body; k += 32) {
        uint32_t t = __ldg(rw + k);
        t |= t >> 4; t |= t >> 2; t |= t >> 1;      // bit 0 of every byte = OR of its 8 bits
        cnt += __popc(t & 0x01010101u);
      }
      const int64_t tail0 = head + 4 * body;
      if (lane < w - tail0) cnt += row[x0 + tail0 + lane] != 0;
    }
  }
  __shared__ uint32_t s_cnt[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t c = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) c += s_cnt[k];
    counts[i] = c;
    sizes[i] = (w > 0 && h > 0) ? w * h : 0;
  }
}

void launch_window_count(const uint8_t* mask, int64_t mh, int64_t mw, const int64_t* win, int64_t n, int64_t dx, int64_t dy, uint32_t* counts,
                         int64_t* sizes, cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return;
  window_count_kernel<<<(unsigned)n, 256, 0, s>>>(mask, mh, mw, win, dx, dy, counts, sizes);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// Ensemble head of the multi-patch ResNet (resnets_shift.py:133-139, :213-215): features = cat over the P patches
// of the pooled 512-vectors -> Linear(P*512, P*256) + ReLU -> Linear(P*256, 4).  The engine holds the pooled features
// patch-major (feats[(p*B + b)*512 + c]), the reference concatenates along dim 1 (k = p*512 + c).  fp32 weights and
// accumulation (33.6 M parameters at P = 16: the kernel streams the weight matrix once per group of <= 8 batch
// elements — weight-bandwidth bound, 134 MB).  One warp per output neuron, fixed lane-strided summation order.
__global__ void __launch_bounds__(256) ensemble_fc1_kernel(const float* __restrict__ feats, int B, int P, const float* __restrict__ w,
                                                            const float* __restrict__ bias, int n_out, float* __restrict__ hid) {
  const int K = P * 512;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + warp;
  if (j >= n_out) return;
  const float4* wr = reinterpret_cast<const float4*>(w + (size_t)j * K);
  for (int b0 = 0; b0 < B; b0 += 8) {
    const int nb = min(8, B - b0);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k4 = lane; k4 < K / 4; k4 += 32) {
      const float4 wv = __ldg(wr + k4);
      const int k = 4 * k4, pch = k >> 9, c = k & 511;
#pragma unroll
      for (int bb = 0; bb < 8; ++bb) {
        if (bb < nb) {
          const float4 fv = __ldg(reinterpret_cast<const float4*>(feats + ((size_t)pch * B + (b0 + bb)) * 512 + c));
          acc[bb] = fmaf(fv.x, wv.x, acc[bb]);
          acc[bb] = fmaf(fv.y, wv.y, acc[bb]);
          acc[bb] = fmaf(fv.z, wv.z, acc[bb]);
          acc[bb] = fmaf(fv.w, wv.w, acc[bb]);
        }
      }
    }
#pragma unroll
    for (int bb = 0; bb < 8; ++bb) {
      float v = acc[bb];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && bb < nb) hid[(size_t)(b0 + bb) * n_out + j] = fmaxf(v + bias[j], 0.f);
    }
  }
}

__global__ void __launch_bounds__(128) ensemble_fc2_kernel(const float* __restrict__ hid, int n_hid, const float* __restrict__ w,
                                                            const float* __restrict__ bias, int n_out, float* __restrict__ out) {
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < n_out; j += (blockDim.x >> 5)) {
    float v = 0.f;
    for (int k = lane; k < n_hid; k += 32) v = fmaf(hid[(size_t)b * n_hid + k], __ldg(w + (size_t)j * n_hid + k), v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) out[(size_t)b * n_out + j] = v + bias[j];
  }
}

void launch_ensemble_head(const float* feats, int B, int P, const float* w1, const float* b1, int n_hid, const float* w2, const float* b2,
                          int n_out, float* hid, float* out, cudaStream_t s, LaunchCounter* lc) {
  if (B <= 0) return;
  ensemble_fc1_kernel<<<(unsigned)ceil_div(n_hid, 8), 256, 0, s>>>(feats, B, P, w1, b1, n_hid, hid);
  CUDA_CHECK(cudaGetLastError());
  ensemble_fc2_kernel<<<(unsigned)B, 128, 0, s>>>(hid, n_hid, w2, b2, n_out, out);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n += 2;
}

}  // namespace wsi
