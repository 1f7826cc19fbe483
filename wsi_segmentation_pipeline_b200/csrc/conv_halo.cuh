// conv_halo.cuh — halo-resident CTA-pair implicit GEMM for 3x3 / stride-1 convolutions on >= 64-channel NHWC tensors.
//
// Measured (profiles/r01_notes.md): the 128/256-channel layers of the TMA implicit-GEMM kernels are bound by the
// L2 -> SM fill of their A tiles — unique per CTA, ~25 B/clk/SM whatever the layout — because every input pixel is
// loaded once per filter tap (9x).  Probe (tools/umma_shift_probe.py): the tensor core applies the 128-byte swizzle
// to ABSOLUTE shared-memory address bits, so a K-major SWIZZLE_128B operand may start at any 128-byte row of a
// TMA-written tile and its 8-row groups may be any multiple of 128 bytes apart (no descriptor base_offset needed).
// Hence:
//   * the M tile is 16 rows x 8 columns of output pixels of one image; per 64-channel K chunk ONE TMA box
//     {64 ch, 10 px, 18 rows} brings the 18 x 10 halo (23 KB; out-of-bounds = the conv zero padding);
//   * the A operand of tap (r, s) is that tile read from row r * 10 + s on, 8-pixel groups 10 rows apart
//     (SBO = 1280 B): nine descriptor offsets instead of nine 16 KB loads — per-CTA-unique traffic drops 6.4x;
//   * weights stream through the stage ring one tap at a time ([Cout][K] with K = (chunk, tap, 64 ch)); they are
//     shared by all CTAs (L2 broadcast) and, as in conv_pair.cuh, each CTA of a pair stages only half of the rows;
//   * everything else (cta_group::2 MMAs with M = 256, multicast commits, 16-arrival tmem_empty, the epilogue) is
//     the protocol of conv_pair.cuh.
#pragma once
#include "conv_pair.cuh"

namespace wsi {

// The kernel is table driven (HaloGroup: one halo load + the taps that read it), which also covers
//   * stride-2 convs: the input's 4 (row, col)-parity views are 4 groups per K chunk (4 loads instead of 9);
//   * the decoder's x2-upsample + concat convs, per output parity class as in conv_igemm.cuh: the upsampled operand
//     is ONE group of 4 collapsed taps per K chunk, the skip operand one group per parity plane it touches.
constexpr int kHaloW = 10, kHaloH = 18, kHaloTileW = 8, kHaloTileH = 16;
constexpr int kHaloMaxGroups = 40;            // per parity class

struct HaloGroup {        // one halo tile of one 64-channel chunk and the filter taps that read it
  int8_t map;             // which A tensor map (box {64 ch, 10 px, 18 rows, 1})
  int8_t oy, ox;          // box origin relative to the tile's lattice origin (-1, 0)
  int8_t ntaps;
  int32_t c0;             // channel origin in that map
  int32_t wk;             // K offset of the first tap's 64 weights (taps are consecutive blocks of 64)
  uint8_t tap_off[12];    // per tap: (row * kHaloW + col) of its 16 x 8 window inside the halo tile
};
static_assert(sizeof(HaloGroup) == 24, "HaloGroup layout");
constexpr int kHaloBytes = kHaloW * kHaloH * 128;                      // 23 040
constexpr int kHaloBuf = (kHaloBytes + 1023) / 1024 * 1024;            // 23 552: buffers stay 1024-byte aligned

template <int BN, bool PLAIN>
struct HaloSmem {
  // The ring is a latency buffer (Little's law: bytes in flight / ~2 us of loaded TMA latency).  A plain group keeps the
  // tensor cores busy for 36 MMAs (~3 500 cycles), so 2 halo tiles in flight are enough and the weight ring gets the
  // rest; the groups of stride-2 / x2 convs have 1-4 taps (<= 1 500 cycles), they need more halo tiles in flight.
  static constexpr int kBufs = PLAIN ? 2 : (BN == 64 ? 6 : (BN == 128 ? 5 : 4));
  static constexpr int kBBytes = (BN / 2) * 128;                        // this CTA's half of one tap's weight rows
  static constexpr int kStagesWanted = ((PLAIN ? 144 : (BN == 128 ? 80 : 96)) * 1024) / kBBytes;
  static constexpr int kStages = kStagesWanted > 12 ? 12 : kStagesWanted;
  static constexpr int kRing = kBufs * kHaloBuf + kStages * kBBytes;
  static constexpr int kBarBytes = 512;
  static constexpr int kScaleBytes = 2 * 512 * (int)sizeof(float);
  static constexpr int kTableBytes = 4 * kHaloMaxGroups * (int)sizeof(HaloGroup);
  static constexpr int kTotal = 1024 + kRing + kBarBytes + kScaleBytes + kTableBytes;
  static constexpr int kTmemCols = 2 * BN;
  static_assert(kTmemCols <= 512, "TMEM");
  static_assert(kBBytes % 1024 == 0, "swizzle atom alignment");
};

// PLAIN: 3x3 / stride-1 conv of one NHWC operand into an NHWC tensor — the hot case.  Its groups are implicit (chunk gi
// = channels [64 gi, 64 gi + 64), origin (-1, -1), taps in (r, s) order), so neither the producer nor the MMA issuer
// reads the table, and the epilogue has no parity / planar addressing: 25 % faster on the 128-channel layers than the
// table-driven instance (same-box A/B).
template <int BN, bool PLAIN, bool PLAIN_EPI = PLAIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
conv_halo_pair_kernel(const __grid_constant__ AMaps amaps, const __grid_constant__ CUtensorMap bmap, const ConvParams p) {
  using S = HaloSmem<BN, PLAIN>;
  constexpr int kHaloBufs = S::kBufs;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* halo_base = smem;                            // kHaloBufs halo buffers
  uint8_t* stage_base = smem + kHaloBufs * kHaloBuf;    // weight ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kRing);
  uint64_t* full = bars;                               // [kStages]  leader only: one tap's weight rows of both CTAs landed
  uint64_t* empty = bars + S::kStages;                 // [kStages]  per CTA
  uint64_t* a_full = bars + 2 * S::kStages;                          // [kHaloBufs] leader only: the halo tiles of both CTAs landed
  uint64_t* a_empty = bars + 2 * S::kStages + kHaloBufs;             // [kHaloBufs] per CTA
  uint64_t* tmem_full = bars + 2 * S::kStages + 2 * kHaloBufs;       // [2]        per CTA
  uint64_t* tmem_empty = bars + 2 * S::kStages + 2 * kHaloBufs + 2;  // [2]        leader only (16 warp arrivals)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * S::kStages + 2 * kHaloBufs + 4);
  float* s_scale = reinterpret_cast<float*>(smem + S::kRing + S::kBarBytes);
  float* s_bias = s_scale + 512;
  HaloGroup* tbl = reinterpret_cast<HaloGroup*>(smem + S::kRing + S::kBarBytes + S::kScaleBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ngroups = p.num_kb;                        // halo groups per parity class
  const uint32_t rank = pptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) { s_scale[i] = p.scale[i]; s_bias[i] = p.bias[i]; }
  for (int i = threadIdx.x; i < p.num_parity * ngroups; i += blockDim.x) tbl[i] = p.hgroups[i];

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < S::kStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kHaloBufs; ++i) {
      ptx::mbar_init(&a_full[i], 1);
      ptx::mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 16);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&bmap);
    ptx::prefetch_tmap(&amaps.m[0]);
  }
  if (warp == 1) pptx::tmem_alloc_pair(tmem_holder, S::kTmemCols);
  ptx::tc_fence_before();
  pptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int tiles_m = p.tiles_n * p.tiles_h * p.tiles_w;
  const int pairs_m = (tiles_m + 1) >> 1;
  const int total_tiles = pairs_m * p.num_parity * p.tiles_co;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (ptx::elect_one()) {
      int stage = 0, abuf = 0;
      uint32_t phase = 0, aphase = 0;
      const uint32_t full0 = pptx::mapa(ptx::smem_u32(&full[0]), 0);
      const uint32_t afull0 = pptx::mapa(ptx::smem_u32(&a_full[0]), 0);
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        int r = tile;
        const int ct = r % p.tiles_co; r /= p.tiles_co;
        const int par = PLAIN ? 0 : r % p.num_parity;
        if (!PLAIN) r /= p.num_parity;
        int mt = 2 * r + (int)rank;
        const int tw = mt % p.tiles_w; mt /= p.tiles_w;
        const int th = mt % p.tiles_h; mt /= p.tiles_h;
        const int tn = mt;                                   // == tiles_n for the out-of-range tail tile
        const int b0 = tw * kHaloTileW, a0 = th * kHaloTileH;
        const int co0 = ct * BN + (int)rank * (BN / 2);
        const int wpar = PLAIN ? 0 : par * p.b_parity_stride;
        const HaloGroup* gt = tbl + par * ngroups;
        for (int gi = 0; gi < ngroups; ++gi) {
          int g_c0 = gi * 64, g_ox = -1, g_oy = -1, g_ntaps = 9, g_wk = gi * 9 * 64;
          const CUtensorMap* am = &amaps.m[0];
          if (!PLAIN) {
            const HaloGroup& g = gt[gi];
            g_c0 = g.c0; g_ox = g.ox; g_oy = g.oy; g_ntaps = g.ntaps; g_wk = g.wk;
            switch (g.map) {
              case 1: am = &amaps.m[1]; break;
              case 2: am = &amaps.m[2]; break;
              case 3: am = &amaps.m[3]; break;
              case 4: am = &amaps.m[4]; break;
              default: break;
            }
          }
          ptx::mbar_wait(&a_empty[abuf], aphase ^ 1u, p.error_flag, 61);
          if (rank == 0) ptx::mbar_expect_tx(&a_full[abuf], 2u * (uint32_t)kHaloBytes);
          pptx::tma_load_4d_pair(halo_base + abuf * kHaloBuf, am, afull0 + (uint32_t)(abuf * 8), g_c0, b0 + g_ox, a0 + g_oy, tn);
          if (++abuf == kHaloBufs) { abuf = 0; aphase ^= 1u; }
          for (int tap = 0; tap < g_ntaps; ++tap) {
            ptx::mbar_wait(&empty[stage], phase ^ 1u, p.error_flag, 62);
            if (rank == 0) ptx::mbar_expect_tx(&full[stage], 2u * (uint32_t)S::kBBytes);
            pptx::tma_load_2d_pair(stage_base + stage * S::kBBytes, &bmap, full0 + (uint32_t)(stage * 8), wpar + g_wk + tap * 64, co0);
            if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader only) ================================
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      // A: SWIZZLE_128B K-major, 8-pixel groups one halo row (10 px = 1280 B) apart; the tap offset is added below
      uint64_t adesc0 = 0;
      adesc0 |= (uint64_t)1 << 16;
      adesc0 |= (uint64_t)((kHaloW * 128) >> 4) << 32;
      adesc0 |= (uint64_t)1 << 46;
      adesc0 |= (uint64_t)2 << 61;
      int stage = 0, abuf = 0;
      uint32_t phase = 0, aphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        pptx::mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1u, p.error_flag, 63);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        const int par = PLAIN ? 0 : (tile / p.tiles_co) % p.num_parity;
        const HaloGroup* gt = tbl + par * ngroups;
        uint32_t started = 0;
        constexpr bool plain = PLAIN;
        for (int gi = 0; gi < ngroups; ++gi) {
          // the table stays in shared memory: a register copy indexed by `tap` would live in local memory
          const int ntaps = plain ? 9 : gt[gi].ntaps;
          const uint8_t* tap_off = gt[gi].tap_off;
          pptx::mbar_wait_cluster(&a_full[abuf], aphase, p.error_flag, 65);
          const uint32_t a_units = (ptx::smem_u32(halo_base + abuf * kHaloBuf) & 0x3FFFFu) >> 4;
#pragma unroll 1
          for (int tap = 0; tap < ntaps; ++tap) {
            pptx::mbar_wait_cluster(&full[stage], phase, p.error_flag, 66);
            ptx::tc_fence_after();
            // plain 3x3 (the hot case): tap (r, s) -> row r * 10 + s, no table read in the issue loop
            const uint32_t toff = plain ? (uint32_t)((tap / 3) * kHaloW + tap % 3) : (uint32_t)tap_off[tap];
            const uint64_t adesc = adesc0 | (uint64_t)(a_units + toff * 8u);
            const uint64_t bdesc = make_kmajor_desc<64>(ptx::smem_u32(stage_base + stage * S::kBBytes));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              pptx::umma_bf16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, started);
              started = 1u;
            }
            pptx::umma_commit_pair(&empty[stage]);
            if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
          }
          pptx::umma_commit_pair(&a_empty[abuf]);       // the halo buffers of both CTAs are free when these MMAs retire
          if (++abuf == kHaloBufs) { abuf = 0; aphase ^= 1u; }
        }
        pptx::umma_commit_pair(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ================================ epilogue (8 warps, both CTAs) ===========================
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    constexpr int CH = BN / 2;
    constexpr int STEP = 32;
    const int c_lo = hsel * CH;
    const uint32_t tmem_empty0 = pptx::mapa(ptx::smem_u32(&tmem_empty[0]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs) {
      int r = tile;
      const int ct = r % p.tiles_co; r /= p.tiles_co;
      const int par = PLAIN_EPI ? 0 : r % p.num_parity;
      if (!PLAIN_EPI) r /= p.num_parity;
      int mt = 2 * r + (int)rank;
      const int tw = mt % p.tiles_w; mt /= p.tiles_w;
      const int th = mt % p.tiles_h; mt /= p.tiles_h;
      const int tn = mt;
      const int co0 = ct * BN;
      const int wl = row % p.bw;
      const int hl = (row / p.bw) % p.bh;
      const int nl = row / (p.bw * p.bh);
      const int n = tn * p.bn + nl, a = th * p.bh + hl, b = tw * p.bw + wl;
      const bool valid = (n < p.N) && (a < p.A_h) && (b < p.A_w);
      const int oh = PLAIN_EPI ? a : p.sigma * a + (par >> 1), ow = PLAIN_EPI ? b : p.sigma * b + (par & 1);
      const size_t pix = ((size_t)n * p.OH + oh) * p.OW + ow;
      const size_t off0 = pix * p.Cout + co0 + c_lo;
      const size_t pl_off = (size_t)n * (size_t)p.pl_img + (size_t)(oh + 1) * (size_t)p.pl_row + (size_t)(ow + 8) * 16;
      const bool has_res = (p.res != nullptr) && valid;

      uint4 rcur[STEP / 8], rnext[STEP / 8];
      if (has_res) {
#pragma unroll
        for (int j = 0; j < STEP / 8; j += 2) ptx::ld_global_nc_256(p.res + off0 + 8 * j, rcur[j], rcur[j + 1]);
      }

      ptx::mbar_wait(&tmem_full[acc], acc_phase, p.error_flag, 64);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c_lo);
#pragma unroll 1
      for (int c = 0; c < CH; c += STEP) {
        uint32_t v[STEP];
#pragma unroll
        for (int j = 0; j < STEP; j += 16) ptx::tmem_ld16(t_row + (uint32_t)(c + j), *reinterpret_cast<uint32_t(*)[16]>(&v[j]));
        if (has_res && c + STEP < CH) {
#pragma unroll
          for (int j = 0; j < STEP / 8; j += 2) ptx::ld_global_nc_256(p.res + off0 + c + STEP + 8 * j, rnext[j], rnext[j + 1]);
        }
        ptx::tmem_ld_wait();
        float y[STEP];
        const float4* sc4 = reinterpret_cast<const float4*>(s_scale + co0 + c_lo + c);
        const float4* bi4 = reinterpret_cast<const float4*>(s_bias + co0 + c_lo + c);
#pragma unroll
        for (int j = 0; j < STEP / 4; ++j) {
          const float4 sc = sc4[j], bb = bi4[j];
          y[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), sc.x, bb.x);
          y[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), sc.y, bb.y);
          y[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), sc.z, bb.z);
          y[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), sc.w, bb.w);
        }
        if (has_res) {
#pragma unroll
          for (int j = 0; j < STEP / 8; ++j) {
            const uint32_t w[4] = {rcur[j].x, rcur[j].y, rcur[j].z, rcur[j].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              y[8 * j + 2 * t + 0] += __uint_as_float(w[t] << 16);
              y[8 * j + 2 * t + 1] += __uint_as_float(w[t] & 0xffff0000u);
            }
          }
#pragma unroll
          for (int j = 0; j < STEP / 8; ++j) rcur[j] = rnext[j];
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < STEP; ++j) y[j] = fmaxf(y[j], 0.f);
        }
        if (valid) {
          // NHWC: 2 * STEP contiguous bytes; planar (consumer = a row kernel): one 16-byte entry per 8-channel chunk row
          const bool planar = !PLAIN_EPI && p.out_planar;
          uint8_t* ob = planar ? reinterpret_cast<uint8_t*>(p.out) + pl_off + (size_t)((co0 + c_lo + c) >> 3) * (size_t)p.pl_chunk
                               : reinterpret_cast<uint8_t*>(p.out + off0 + c);
          const size_t ostep = planar ? (size_t)p.pl_chunk : 16;
          if (!planar) {
#pragma unroll
            for (int j = 0; j < STEP / 16; ++j) {      // NHWC: 256-bit stores, one full sector each
              uint32_t w[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(y[16 * j + 2 * t], y[16 * j + 2 * t + 1]);
                w[t] = *reinterpret_cast<uint32_t*>(&h2);
              }
              ptx::st_global_256(ob + (size_t)j * 32, w);
            }
          } else {
#pragma unroll
          for (int j = 0; j < STEP / 8; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(y[8 * j + 2 * t], y[8 * j + 2 * t + 1]);
              w[t] = *reinterpret_cast<uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(ob + (size_t)j * ostep) = make_uint4(w[0], w[1], w[2], w[3]);
          }
          }
        }
      }
      // this warp's TMEM reads are done: one arrival per warp on the leader's barrier
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) pptx::mbar_arrive_remote(tmem_empty0 + (uint32_t)(acc * 8));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  // neither CTA may exit (or free TMEM) while the other can still signal its barriers or read its smem
  ptx::tc_fence_before();
  pptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    pptx::tmem_dealloc_pair(tmem_base, S::kTmemCols);
  }
}

}  // namespace wsi
