// planner.cpp — host-side tile planner and row-band partitioner (no CUDA needed).
//
// wsi_plan_tiles is bit-exact with the reference's enumeration (utils/dataset.py:143-166) including
// its quirks: tiles start at (1,1); main grid, then the right column x = iw-1-pw, then the bottom
// row y = ih-1-ph; the bottom-right corner tile is never generated; a tile is kept iff
// count_nonzero(window)/window.size >= 0.05 (utils/preprocessing.py:60-71) on the level-2 mask
// window [int(y*m) : +int(ph*m), int(x*m) : +int(pw*m)] with numpy's silent clipping at the mask
// edge.  Geometries on which the reference itself raises (empty window -> ZeroDivisionError) or
// indexes with a negative origin return WSI_ERR_DEGENERATE.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/wsi_b200.h"

namespace {

// python: list(range(start, stop, step)) for step > 0
std::vector<int64_t> py_range(int64_t start, int64_t stop, int64_t step) {
  std::vector<int64_t> v;
  for (int64_t i = start; i < stop; i += step) v.push_back(i);
  return v;
}

struct RowCounter {
  // column prefix sums of the mask over rows [yp, yp+dy) (clipped), rebuilt per distinct yp
  const uint8_t* mask;
  int64_t mh, mw;
  int64_t cur_yp = -1, cur_h = 0;
  std::vector<uint32_t> prefix;  // [mw+1]
  void set_row(int64_t yp, int64_t dy) {
    if (yp == cur_yp) return;
    cur_yp = yp;
    const int64_t y0 = std::min(yp, mh), y1 = std::min(yp + dy, mh);
    cur_h = y1 - y0;
    std::vector<uint32_t> col((size_t)mw, 0);
    for (int64_t y = y0; y < y1; ++y) {
      const uint8_t* r = mask + y * mw;
      for (int64_t x = 0; x < mw; ++x) col[(size_t)x] += (r[x] != 0);
    }
    prefix.assign((size_t)mw + 1, 0);
    for (int64_t x = 0; x < mw; ++x) prefix[(size_t)x + 1] = prefix[(size_t)x] + col[(size_t)x];
  }
  // returns 1 keep, 0 drop, -1 degenerate
  int foreground(int64_t xp, int64_t dx) const {
    const int64_t x0 = std::min(xp, mw), x1 = std::min(xp + dx, mw);
    const int64_t size = cur_h * (x1 - x0);
    if (size <= 0) return -1;
    const uint32_t cnt = prefix[(size_t)x1] - prefix[(size_t)x0];
    return ((double)cnt / (double)size >= 0.05) ? 1 : 0;
  }
};

}  // namespace

extern "C" int wsi_plan_tiles(int64_t ih, int64_t iw, int32_t ph, int32_t pw, int32_t sh, int32_t sw,
                              const uint8_t* mask, int64_t mh, int64_t mw, double m, int32_t** xy_out,
                              int64_t* n_out) {
  if (!xy_out || !n_out || ih <= 0 || iw <= 0 || ph <= 0 || pw <= 0 || sh <= 0 || sw <= 0 || !(m > 0)) return WSI_ERR_INVALID;
  if (mask && (mh <= 0 || mw <= 0)) return WSI_ERR_INVALID;
  *xy_out = nullptr;
  *n_out = 0;
  const std::vector<int64_t> ys = py_range(1, ih - 1 - ph, sh), xs = py_range(1, iw - 1 - pw, sw);
  const int64_t x_last = iw - 1 - pw, y_last = ih - 1 - ph;
  const int64_t dx = (int64_t)((double)pw * m), dy = (int64_t)((double)ph * m);
  std::vector<int32_t> out;
  out.reserve((ys.size() * (xs.size() + 1) + xs.size()) * 2);
  RowCounter rc{mask, mh, mw};

  auto test = [&](int64_t xpos, int64_t ypos) -> int {
    if (xpos < 0 || ypos < 0) return -1;   // the reference would index the mask with a negative origin
    if (!mask) return 1;
    const int64_t yp = (int64_t)((double)ypos * m), xp = (int64_t)((double)xpos * m);
    rc.set_row(yp, dy);
    return rc.foreground(xp, dx);
  };
  // pass 1: main grid rows, each followed later by its right-column tile; to reuse the per-row
  // column sums we evaluate the right-column tile together with its row but emit it in the
  // reference's order (all grid tiles, then the right column, then the bottom row).
  std::vector<int32_t> right;
  for (int64_t ypos : ys) {
    for (int64_t xpos : xs) {
      const int k = test(xpos, ypos);
      if (k < 0) return WSI_ERR_DEGENERATE;
      if (k) { out.push_back((int32_t)xpos); out.push_back((int32_t)ypos); }
    }
    const int k = test(x_last, ypos);
    if (k < 0) return WSI_ERR_DEGENERATE;
    if (k) { right.push_back((int32_t)x_last); right.push_back((int32_t)ypos); }
  }
  out.insert(out.end(), right.begin(), right.end());
  for (int64_t xpos : xs) {
    const int k = test(xpos, y_last);
    if (k < 0) return WSI_ERR_DEGENERATE;
    if (k) { out.push_back((int32_t)xpos); out.push_back((int32_t)y_last); }
  }
  const int64_t n = (int64_t)out.size() / 2;
  int32_t* buf = (int32_t*)malloc(std::max<size_t>(out.size(), 2) * sizeof(int32_t));
  if (!buf) return WSI_ERR_NOMEM;
  if (!out.empty()) memcpy(buf, out.data(), out.size() * sizeof(int32_t));
  *xy_out = buf;
  *n_out = n;
  return WSI_OK;
}

extern "C" void wsi_free(void* p) { free(p); }

// Row bands over canvas rows, boundaries on the tile grid: r_k = 1 + sh*round(k*ny/G).
// Band k owns canvas rows [own0, own1) and needs raster rows [row0, row1) = the union of every tile
// row intersecting it (tiles straddling a boundary are evaluated by both neighbours; no exchange).
extern "C" int wsi_band_partition(int64_t ih, int32_t ph, int32_t sh, int32_t nranks, int64_t* bands) {
  if (!bands || nranks <= 0 || ih <= 0 || ph <= 0 || sh <= 0) return WSI_ERR_INVALID;
  std::vector<int64_t> ys = py_range(1, ih - 1 - ph, sh);
  const int64_t ny = (int64_t)ys.size();
  const int64_t y_last = ih - 1 - ph;
  if (y_last >= 0) ys.push_back(y_last);   // bottom-row tiles
  for (int k = 0; k < nranks; ++k) {
    auto bound = [&](int kk) -> int64_t {
      if (kk <= 0) return 0;
      if (kk >= nranks) return ih;
      const int64_t q = (2 * (int64_t)kk * ny + nranks) / (2 * (int64_t)nranks);  // round(k*ny/G), half up
      return std::min<int64_t>(ih, 1 + (int64_t)sh * q);
    };
    const int64_t own0 = bound(k), own1 = bound(k + 1);
    int64_t row0 = -1, row1 = -1;
    for (int64_t y : ys) {
      if (y < own1 && y + ph > own0) {
        if (row0 < 0 || y < row0) row0 = y;
        if (y + ph > row1) row1 = y + ph;
      }
    }
    if (row0 < 0) { row0 = own0; row1 = own0; }
    bands[4 * k + 0] = own0;
    bands[4 * k + 1] = own1;
    bands[4 * k + 2] = row0;
    bands[4 * k + 3] = row1;
  }
  return WSI_OK;
}

extern "C" int wsi_band_tiles(const int32_t* xy, int64_t n, int32_t ph, double m, int64_t own0, int64_t own1,
                              int64_t** idx_out, int64_t* n_out) {
  if (!xy || !idx_out || !n_out || n < 0) return WSI_ERR_INVALID;
  const int64_t dy = (int64_t)((double)ph * m);
  std::vector<int64_t> keep;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t ty = (int64_t)(m * (double)xy[2 * i + 1]);
    if (ty < own1 && ty + dy > own0) keep.push_back(i);
  }
  int64_t* buf = (int64_t*)malloc(std::max<size_t>(keep.size(), 1) * sizeof(int64_t));
  if (!buf) return WSI_ERR_NOMEM;
  if (!keep.empty()) memcpy(buf, keep.data(), keep.size() * sizeof(int64_t));
  *idx_out = buf;
  *n_out = (int64_t)keep.size();
  return WSI_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Coefficients of the tile resize of the scan_resize != 1 branch (utils/dataset.py:180-181):
// `image.resize((tile_w, tile_h))` = PIL.Image.resize at its default filter.  Pillow is a third-party dependency of
// the reference (not vendored, no version pinned in the tree); this restates the published algorithm of the Pillow
// installed in this image (12.2: default filter BICUBIC for RGB images; libImaging/Resample.c precompute_coeffs +
// normalize_coeffs_8bpc): antialiased bicubic (a = -0.5), support 2 * max(scale, 1), window
// [int(center - support + 0.5), int(center + support + 0.5)) clipped to the image, weights normalised in double and
// rounded to 22-bit fixed point.  One axis; the tile is a standalone image, so every tile shares one table per axis.
// ------------------------------------------------------------------------------------------------------------
namespace {
double bicubic_weight(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
}  // namespace

extern "C" int wsi_resample_ksize(int32_t in_size, int32_t out_size) {
  if (in_size <= 0 || out_size <= 0) return 0;
  double fs = (double)in_size / (double)out_size;
  if (fs < 1.0) fs = 1.0;
  return (int)ceil(2.0 * fs) * 2 + 1;
}

extern "C" int wsi_resample_coeffs(int32_t in_size, int32_t out_size, int32_t* bounds /*[out_size][2]*/, int32_t* kk /*[out_size][ksize]*/) {
  if (in_size <= 0 || out_size <= 0 || !bounds || !kk) return WSI_ERR_INVALID;
  const int precision_bits = 32 - 8 - 2;
  const double scale = (double)in_size / (double)out_size;
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * fs;
  const int ksize = (int)ceil(support) * 2 + 1;
  const double ss = 1.0 / fs;
  std::vector<double> k((size_t)ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    int x = 0;
    for (; x < xmax; ++x) {
      const double w = bicubic_weight((x + xmin - center + 0.5) * ss);
      k[(size_t)x] = w;
      ww += w;
    }
    for (int i = 0; i < xmax; ++i)
      if (ww != 0.0) k[(size_t)i] /= ww;
    for (; x < ksize; ++x) k[(size_t)x] = 0.0;
    for (int i = 0; i < ksize; ++i) {
      const double v = k[(size_t)i];
      kk[(size_t)xx * ksize + i] = (v < 0) ? (int)(-0.5 + v * (1 << precision_bits)) : (int)(0.5 + v * (1 << precision_bits));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return WSI_OK;
}
