// tiff_ingest.cu — slide ingestion: JPEG-compressed (tiled or stripped) TIFF / Aperio SVS pyramid level -> u8 RGB raster rows
// in HBM.
//
// Replaces, for rasters that are still on disk, what the reference does through OpenSlide:
//   utils/dataset.py:121      self.scan = openslide.OpenSlide(wsipth)
//   utils/dataset.py:175-178  self.scan.read_region((x, y), level, (pw, ph)).convert('RGB')      (one tile at a time, on the CPU)
//   utils/eval.py:263         scan.read_region((0, 0), 2, scan.level_dimensions[2])             (the level-2 thumbnail)
// Here a row band of a pyramid level is decoded ONCE, straight into the device raster that the gather kernel cuts tiles
// from: the host parses the TIFF directory (classic and BigTIFF, little-endian), splices each tile's abbreviated JPEG
// stream with the shared JPEGTables (tag 347), nvJPEG decodes it on the GPU, and a small kernel (or a 2-D copy) places the
// pixels — clipped at the image edge — into [rows][iw][3].
//   Photometric = YCbCr (6): nvJPEG converts to interleaved RGB (NVJPEG_OUTPUT_RGBI).
//   Photometric = RGB (2), as Aperio writes "JPEG/RGB" and libtiff writes RGB input: the three JPEG components ARE R, G, B;
//   they are decoded unconverted (NVJPEG_OUTPUT_UNCHANGED) and interleaved here.
// There is no reference arithmetic to match beyond the decoded bytes; JPEG decoders may differ by a level or two in their
// IDCT / chroma upsampling, so parity against libtiff + libjpeg (PIL) is a tolerance, stated in the tests.
// nvJPEG is loaded lazily with dlopen (no link-time dependency): without it these entry points return WSI_ERR_UNSUPPORTED.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <algorithm>
#include <memory>
#include <mutex>

#include "kernels.cuh"

namespace wsi {

int ctx_device(const wsi_ctx* c);            // engine.cu
LaunchCounter* ctx_launch_counter(wsi_ctx* c);

// ---- nvJPEG through dlopen ----------------------------------------------------------------------------------------
struct NvJpegApi {
  void* lib = nullptr;
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*, cudaStream_t) = nullptr;
};
static NvJpegApi g_nvjpeg;
static std::once_flag g_nvjpeg_once;

static const NvJpegApi& nvjpeg_api() {
  std::call_once(g_nvjpeg_once, [] {
    for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
      g_nvjpeg.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (g_nvjpeg.lib) break;
    }
    if (!g_nvjpeg.lib) return;
    auto sym = [](const char* n) { return dlsym(g_nvjpeg.lib, n); };
    g_nvjpeg.CreateSimple = reinterpret_cast<decltype(g_nvjpeg.CreateSimple)>(sym("nvjpegCreateSimple"));
    g_nvjpeg.Destroy = reinterpret_cast<decltype(g_nvjpeg.Destroy)>(sym("nvjpegDestroy"));
    g_nvjpeg.JpegStateCreate = reinterpret_cast<decltype(g_nvjpeg.JpegStateCreate)>(sym("nvjpegJpegStateCreate"));
    g_nvjpeg.JpegStateDestroy = reinterpret_cast<decltype(g_nvjpeg.JpegStateDestroy)>(sym("nvjpegJpegStateDestroy"));
    g_nvjpeg.GetImageInfo = reinterpret_cast<decltype(g_nvjpeg.GetImageInfo)>(sym("nvjpegGetImageInfo"));
    g_nvjpeg.Decode = reinterpret_cast<decltype(g_nvjpeg.Decode)>(sym("nvjpegDecode"));
  });
  WSI_REQUIRE(g_nvjpeg.lib && g_nvjpeg.CreateSimple && g_nvjpeg.Decode && g_nvjpeg.JpegStateCreate && g_nvjpeg.GetImageInfo, WSI_ERR_UNSUPPORTED,
              "libnvjpeg is not available on this machine (%s)", g_nvjpeg.lib ? "symbols missing" : "dlopen failed");
  return g_nvjpeg;
}

// ---- TIFF directory ----------------------------------------------------------------------------------------------
struct TiffLevel {
  int64_t W = 0, H = 0;
  int tile_w = 0, tile_h = 0;      // tiled layout; 0 = strips
  int rows_per_strip = 0;
  int compression = 1, photometric = 2, spp = 1, bits = 8, planar = 1;
  std::vector<uint64_t> offsets, counts;
  std::vector<uint8_t> jpeg_tables;
  int64_t units_across() const { return tile_w ? ceil_div(W, tile_w) : 1; }
  int64_t unit_w() const { return tile_w ? tile_w : W; }
  int64_t unit_h() const { return tile_w ? tile_h : rows_per_strip; }
};

}  // namespace wsi

struct wsi_tiff {
  std::string path, err;
  FILE* f = nullptr;
  bool big = false;
  std::vector<wsi::TiffLevel> levels;
  // decode state (created on first read)
  nvjpegHandle_t nvh = nullptr;
  nvjpegJpegState_t nvs = nullptr;
  wsi::DevBuf unit_buf;
  ~wsi_tiff() {
    if (nvs && wsi::g_nvjpeg.JpegStateDestroy) wsi::g_nvjpeg.JpegStateDestroy(nvs);
    if (nvh && wsi::g_nvjpeg.Destroy) wsi::g_nvjpeg.Destroy(nvh);
    if (f) fclose(f);
  }
};

namespace wsi {

static void read_at(wsi_tiff* t, uint64_t off, void* dst, size_t n) {
  WSI_REQUIRE(fseeko(t->f, (off_t)off, SEEK_SET) == 0 && fread(dst, 1, n, t->f) == n, WSI_ERR_INVALID, "%s: short read of %zu bytes at %llu",
              t->path.c_str(), n, (unsigned long long)off);
}

static size_t type_size(int type) {
  switch (type) {
    case 1: case 2: case 6: case 7: return 1;
    case 3: case 8: return 2;
    case 4: case 9: case 11: case 13: return 4;
    case 5: case 10: case 12: case 16: case 17: case 18: return 8;
    default: return 0;
  }
}

// values of one directory entry as u64 (BYTE / SHORT / LONG / LONG8) or raw bytes (UNDEFINED)
static std::vector<uint64_t> entry_values(wsi_tiff* t, int type, uint64_t count, const uint8_t* inline_bytes, size_t inline_cap, std::vector<uint8_t>* raw) {
  const size_t ts = type_size(type);
  WSI_REQUIRE(ts > 0 && count < (1ULL << 32), WSI_ERR_UNSUPPORTED, "%s: TIFF field type %d / count %llu", t->path.c_str(), type, (unsigned long long)count);
  std::vector<uint8_t> buf((size_t)count * ts);
  if (buf.size() <= inline_cap) {
    memcpy(buf.data(), inline_bytes, buf.size());
  } else {
    uint64_t off = 0;
    memcpy(&off, inline_bytes, inline_cap);          // little-endian: the offset occupies the first inline_cap bytes
    read_at(t, off, buf.data(), buf.size());
  }
  std::vector<uint64_t> v((size_t)count);
  for (size_t i = 0; i < (size_t)count; ++i) {
    uint64_t x = 0;
    memcpy(&x, buf.data() + i * ts, ts);
    v[i] = x;
  }
  if (raw) *raw = std::move(buf);
  return v;
}

static void parse_directories(wsi_tiff* t) {
  uint8_t hdr[16];
  read_at(t, 0, hdr, 8);
  WSI_REQUIRE(hdr[0] == 'I' && hdr[1] == 'I', WSI_ERR_UNSUPPORTED, "%s: only little-endian TIFF ('II') is supported", t->path.c_str());
  uint16_t magic;
  memcpy(&magic, hdr + 2, 2);
  uint64_t ifd = 0;
  if (magic == 42) {
    uint32_t o;
    memcpy(&o, hdr + 4, 4);
    ifd = o;
  } else if (magic == 43) {
    t->big = true;
    read_at(t, 0, hdr, 16);
    memcpy(&ifd, hdr + 8, 8);
  } else {
    WSI_THROW(WSI_ERR_INVALID, "%s: not a TIFF file (magic %u)", t->path.c_str(), magic);
  }
  const size_t esz = t->big ? 20 : 12, inl = t->big ? 8 : 4;
  int guard = 0;
  while (ifd != 0 && guard++ < 256) {
    uint64_t n = 0;
    if (t->big) { read_at(t, ifd, &n, 8); } else { uint16_t n16; read_at(t, ifd, &n16, 2); n = n16; }
    WSI_REQUIRE(n > 0 && n < 4096, WSI_ERR_INVALID, "%s: corrupt TIFF directory", t->path.c_str());
    std::vector<uint8_t> ents((size_t)n * esz + 8);
    read_at(t, ifd + (t->big ? 8 : 2), ents.data(), (size_t)n * esz + inl);
    TiffLevel L;
    for (uint64_t i = 0; i < n; ++i) {
      const uint8_t* e = ents.data() + i * esz;
      uint16_t tag, type;
      memcpy(&tag, e, 2);
      memcpy(&type, e + 2, 2);
      uint64_t count = 0;
      memcpy(&count, e + 4, t->big ? 8 : 4);
      const uint8_t* val = e + (t->big ? 12 : 8);
      auto vals = [&](std::vector<uint8_t>* raw = nullptr) { return entry_values(t, type, count, val, inl, raw); };
      switch (tag) {
        case 256: L.W = (int64_t)vals()[0]; break;
        case 257: L.H = (int64_t)vals()[0]; break;
        case 258: L.bits = (int)vals()[0]; break;
        case 259: L.compression = (int)vals()[0]; break;
        case 262: L.photometric = (int)vals()[0]; break;
        case 273: case 324: L.offsets = vals(); break;
        case 277: L.spp = (int)vals()[0]; break;
        case 278: L.rows_per_strip = (int)std::min<uint64_t>(vals()[0], 1u << 30); break;
        case 279: case 325: L.counts = vals(); break;
        case 284: L.planar = (int)vals()[0]; break;
        case 322: L.tile_w = (int)vals()[0]; break;
        case 323: L.tile_h = (int)vals()[0]; break;
        case 347: vals(&L.jpeg_tables); break;
        default: break;
      }
    }
    if (L.rows_per_strip <= 0 || L.rows_per_strip > L.H) L.rows_per_strip = (int)L.H;
    t->levels.push_back(std::move(L));
    uint64_t next = 0;
    memcpy(&next, ents.data() + (size_t)n * esz, inl);
    ifd = next;
  }
  WSI_REQUIRE(!t->levels.empty(), WSI_ERR_INVALID, "%s: no image directory", t->path.c_str());
}

// the complete JPEG stream of tile / strip k: JPEGTables without its EOI + the abbreviated stream without its SOI
static void unit_stream(wsi_tiff* t, const TiffLevel& L, int64_t k, std::vector<uint8_t>& out) {
  WSI_REQUIRE(k >= 0 && (size_t)k < L.offsets.size() && (size_t)k < L.counts.size(), WSI_ERR_INVALID, "tile %lld out of range", (long long)k);
  const size_t n = (size_t)L.counts[(size_t)k];
  std::vector<uint8_t> data(n);
  read_at(t, L.offsets[(size_t)k], data.data(), n);
  out.clear();
  const auto& tb = L.jpeg_tables;
  if (tb.size() >= 4 && n >= 2 && data[0] == 0xFF && data[1] == 0xD8) {
    size_t tl = tb.size();
    if (tb[tl - 2] == 0xFF && tb[tl - 1] == 0xD9) tl -= 2;
    out.insert(out.end(), tb.begin(), tb.begin() + (long)tl);
    out.insert(out.end(), data.begin() + 2, data.end());
  } else {
    out = std::move(data);
  }
}

static void check_level(const wsi_tiff* t, int level) {
  WSI_REQUIRE(level >= 0 && level < (int)t->levels.size(), WSI_ERR_INVALID, "level %d out of range (%zu directories)", level, t->levels.size());
  const TiffLevel& L = t->levels[(size_t)level];
  WSI_REQUIRE(L.compression == 7, WSI_ERR_UNSUPPORTED, "level %d: compression %d (only JPEG, 7, is decoded here)", level, L.compression);
  WSI_REQUIRE(L.spp == 3 && L.bits == 8 && L.planar == 1 && (L.photometric == 2 || L.photometric == 6), WSI_ERR_UNSUPPORTED,
              "level %d: need 8-bit chunky RGB / YCbCr (spp %d, bits %d, photometric %d)", level, L.spp, L.bits, L.photometric);
  WSI_REQUIRE(!L.offsets.empty() && L.offsets.size() == L.counts.size(), WSI_ERR_INVALID, "level %d: tile table missing", level);
}

// three u8 planes (pitch pp) -> interleaved RGB rows of the raster, clipped to [w, h]
__global__ void __launch_bounds__(256) interleave_rgb_kernel(const uint8_t* __restrict__ p0, const uint8_t* __restrict__ p1, const uint8_t* __restrict__ p2,
                                                              int pp, int w, int h, uint8_t* __restrict__ dst, int64_t dst_stride) {
  const int total = w * h;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / w, x = i - y * w;
    uint8_t* o = dst + (int64_t)y * dst_stride + 3 * x;
    o[0] = p0[(size_t)y * pp + x];
    o[1] = p1[(size_t)y * pp + x];
    o[2] = p2[(size_t)y * pp + x];
  }
}

}  // namespace wsi

using namespace wsi;

#define TIFF_API_BEGIN try {
#define TIFF_API_END(tp)                                   \
  }                                                        \
  catch (const ::wsi::Error& e) {                          \
    if (tp) (tp)->err = e.msg;                             \
    ::wsi::set_global_error(e.msg);                        \
    return e.status;                                       \
  }                                                        \
  catch (const std::exception& e) {                        \
    if (tp) (tp)->err = e.what();                          \
    ::wsi::set_global_error(e.what());                     \
    return WSI_ERR_INVALID;                                \
  }                                                        \
  return WSI_OK;

extern "C" {

int wsi_tiff_open(const char* path, wsi_tiff** out) {
  wsi_tiff* none = nullptr;
  TIFF_API_BEGIN
  WSI_REQUIRE(path && out, WSI_ERR_INVALID, "NULL argument");
  *out = nullptr;
  std::unique_ptr<wsi_tiff> t(new wsi_tiff());
  t->path = path;
  t->f = fopen(path, "rb");
  WSI_REQUIRE(t->f != nullptr, WSI_ERR_INVALID, "cannot open %s", path);
  parse_directories(t.get());
  *out = t.release();
  TIFF_API_END(none)
}

int wsi_tiff_close(wsi_tiff* t) {
  delete t;
  return WSI_OK;
}

const char* wsi_tiff_last_error(wsi_tiff* t) { return t ? t->err.c_str() : ""; }

int wsi_tiff_levels(wsi_tiff* t) { return t ? (int)t->levels.size() : 0; }

int wsi_tiff_level_info(wsi_tiff* t, int level, int64_t* W, int64_t* H, int32_t* tile_w, int32_t* tile_h, int32_t* compression, int32_t* photometric) {
  TIFF_API_BEGIN
  WSI_REQUIRE(t && level >= 0 && level < (int)t->levels.size(), WSI_ERR_INVALID, "bad level");
  const TiffLevel& L = t->levels[(size_t)level];
  if (W) *W = L.W;
  if (H) *H = L.H;
  if (tile_w) *tile_w = (int32_t)L.unit_w();
  if (tile_h) *tile_h = (int32_t)L.unit_h();
  if (compression) *compression = L.compression;
  if (photometric) *photometric = L.photometric;
  TIFF_API_END(t)
}

int wsi_tiff_unit_stream(wsi_tiff* t, int level, int64_t unit, uint8_t* buf, int64_t cap, int64_t* len) {
  TIFF_API_BEGIN
  WSI_REQUIRE(t && len, WSI_ERR_INVALID, "NULL argument");
  check_level(t, level);
  std::vector<uint8_t> s;
  unit_stream(t, t->levels[(size_t)level], unit, s);
  *len = (int64_t)s.size();
  if (buf && cap >= (int64_t)s.size()) memcpy(buf, s.data(), s.size());
  TIFF_API_END(t)
}

int wsi_tiff_read_rows(wsi_ctx* ctx, wsi_tiff* t, int level, int64_t row0, int64_t rows, uint8_t* rgb_dev, int64_t row_stride, void* stream) {
  TIFF_API_BEGIN
  WSI_REQUIRE(ctx && t && rgb_dev && rows > 0 && row0 >= 0, WSI_ERR_INVALID, "bad argument");
  check_level(t, level);
  const TiffLevel& L = t->levels[(size_t)level];
  WSI_REQUIRE(row0 + rows <= L.H && row_stride >= 3 * L.W, WSI_ERR_INVALID, "rows [%lld, +%lld) / stride %lld do not fit the %lld x %lld level",
              (long long)row0, (long long)rows, (long long)row_stride, (long long)L.W, (long long)L.H);
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_CHECK(cudaSetDevice(ctx_device(ctx)));
  const NvJpegApi& nv = nvjpeg_api();
  if (!t->nvh) {
    WSI_REQUIRE(nv.CreateSimple(&t->nvh) == NVJPEG_STATUS_SUCCESS, WSI_ERR_CUDA, "nvjpegCreateSimple failed");
    WSI_REQUIRE(nv.JpegStateCreate(t->nvh, &t->nvs) == NVJPEG_STATUS_SUCCESS, WSI_ERR_CUDA, "nvjpegJpegStateCreate failed");
  }
  const int64_t uw = L.unit_w(), uh = L.unit_h(), across = L.units_across();
  const int64_t pitch = round_up(uw, 256);
  t->unit_buf.alloc((size_t)pitch * 3 * (size_t)uh + 1024);
  std::vector<uint8_t> jpg;
  for (int64_t uy = row0 / uh; uy * uh < row0 + rows; ++uy) {
    for (int64_t ux = 0; ux < across; ++ux) {
      unit_stream(t, L, uy * across + ux, jpg);
      int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
      nvjpegChromaSubsampling_t ss;
      WSI_REQUIRE(nv.GetImageInfo(t->nvh, jpg.data(), jpg.size(), &ncomp, &ss, ws, hs) == NVJPEG_STATUS_SUCCESS, WSI_ERR_INVALID,
                  "%s: tile (%lld, %lld) is not a decodable JPEG stream", t->path.c_str(), (long long)ux, (long long)uy);
      WSI_REQUIRE(ncomp == 3 && ws[0] <= uw && hs[0] <= uh, WSI_ERR_UNSUPPORTED, "tile (%lld, %lld): %d components, %d x %d", (long long)ux, (long long)uy, ncomp, ws[0], hs[0]);
      // rows / columns of this unit that land in the requested band
      const int64_t y_top = uy * uh, x_left = ux * uw;
      const int64_t ya = std::max(y_top, row0), yb = std::min<int64_t>({y_top + hs[0], row0 + rows, L.H});
      const int64_t cw = std::min<int64_t>(ws[0], L.W - x_left);
      if (yb <= ya || cw <= 0) continue;
      uint8_t* dst = rgb_dev + (ya - row0) * row_stride + 3 * x_left;
      nvjpegImage_t img;
      memset(&img, 0, sizeof(img));
      if (L.photometric == 6) {
        img.channel[0] = t->unit_buf.as<uint8_t>();
        img.pitch[0] = (size_t)pitch * 3;
        WSI_REQUIRE(nv.Decode(t->nvh, t->nvs, jpg.data(), jpg.size(), NVJPEG_OUTPUT_RGBI, &img, s) == NVJPEG_STATUS_SUCCESS, WSI_ERR_CUDA,
                    "nvjpegDecode failed on tile (%lld, %lld)", (long long)ux, (long long)uy);
        CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)row_stride, img.channel[0] + (ya - y_top) * img.pitch[0], img.pitch[0], (size_t)cw * 3, (size_t)(yb - ya),
                                     cudaMemcpyDeviceToDevice, s));
      } else {
        WSI_REQUIRE(ss == NVJPEG_CSS_444, WSI_ERR_UNSUPPORTED, "RGB-photometric JPEG with subsampled components");
        for (int c = 0; c < 3; ++c) { img.channel[c] = t->unit_buf.as<uint8_t>() + (size_t)c * pitch * uh; img.pitch[c] = (size_t)pitch; }
        WSI_REQUIRE(nv.Decode(t->nvh, t->nvs, jpg.data(), jpg.size(), NVJPEG_OUTPUT_UNCHANGED, &img, s) == NVJPEG_STATUS_SUCCESS, WSI_ERR_CUDA,
                    "nvjpegDecode failed on tile (%lld, %lld)", (long long)ux, (long long)uy);
        const size_t skip = (size_t)(ya - y_top) * pitch;
        const int total = (int)(cw * (yb - ya));
        interleave_rgb_kernel<<<std::min(ceil_div(total, 256), (int64_t)148 * 8), 256, 0, s>>>(img.channel[0] + skip, img.channel[1] + skip, img.channel[2] + skip,
                                                                                              (int)pitch, (int)cw, (int)(yb - ya), dst, row_stride);
        CUDA_CHECK(cudaGetLastError());
        ctx_launch_counter(ctx)->n++;
      }
      // the JPEG stream buffer and the unit buffer are reused by the next tile: finish this one first (ingestion is not on
      // the timed path; a batched nvjpegDecodeBatched pipeline is the obvious next step)
      CUDA_CHECK(cudaStreamSynchronize(s));
    }
  }
  TIFF_API_END(t)
}

}  // extern "C"
