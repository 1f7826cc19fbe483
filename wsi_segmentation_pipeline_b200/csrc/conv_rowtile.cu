// conv_rowtile.cu — see conv_rowtile.cuh.
#include "conv_rowtile.cuh"

#include <algorithm>
#include <cstdlib>

namespace wsi {

namespace ptx {
// contiguous global -> shared bulk copy on the TMA engine, completion signalled on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace ptx

// K-major, no swizzle: 8-row core matrices of 16-byte rows; LBO = distance between the two 16-byte
// K chunks of one MMA (K = 16 bf16), SBO = distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}

struct TileCoord {
  int n, y, xb;
  __device__ __forceinline__ void init(int tile, int tiles_x, int OH) {
    xb = tile % tiles_x;
    const int t = tile / tiles_x;
    y = t % OH;
    n = t / OH;
  }
  __device__ __forceinline__ void next(int tiles_x, int OH) {
    if (++xb == tiles_x) {
      xb = 0;
      if (++y == OH) { y = 0; ++n; }
    }
  }
  __device__ __forceinline__ void advance(int k, int tiles_x, int OH) {
    for (int i = 0; i < k; ++i) next(tiles_x, OH);
  }
};

template <int BN, bool HEAD, int G>
__global__ void __launch_bounds__(kRowThreads, 1) conv_rowtile_kernel(const __grid_constant__ RowParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int w_bytes = p.w_blocks * 2 * BN * 16;
  uint8_t* s_w = smem;
  float* s_scale = reinterpret_cast<float*>(smem + w_bytes);
  float* s_bias = s_scale + BN;
  float* s_hw = s_bias + BN;
  float* s_hb = s_hw + 64;
  uint8_t* s_stage = smem + ((w_bytes + (2 * BN + 68) * 4 + 127) & ~127);
  const int S = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + (size_t)S * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tmem_full = bars + 2 * S;
  uint64_t* tmem_empty = bars + 2 * S + kRowAccStages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * S + 2 * kRowAccStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // accumulator ring: kRowAccStages tiles in flight between the MMA issuer and the epilogue.  Back-to-back
  // MMAs into the same TMEM columns serialise on the accumulate dependency (~120 vs ~60 cycles per small
  // MMA, measured), so consecutive issues always alternate accumulators: the two column-parity groups
  // when G == 2, two partial sums (added in the epilogue) when G == 1.
  constexpr int NA = (G == 1) ? 2 : 1;
  static_assert(G * NA == 2, "two accumulators per tile, one per MMA warp");
  constexpr uint32_t kTmemCols = kRowAccStages * G * NA * BN;
  static_assert(kTmemCols >= 32 && kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns must be a power of two");

  // resident weights + epilogue constants (generic-proxy writes, read by the async proxy -> fence)
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = threadIdx.x; i < w_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < BN; i += blockDim.x) { s_scale[i] = p.scale[i]; s_bias[i] = p.bias[i]; }
    if (HEAD)
      for (int i = threadIdx.x; i < 68; i += blockDim.x) s_hw[i] = (i < 64) ? p.head_w[i] : p.head_b[i - 64];
    // stage planes are only partially overwritten for ragged tiles: start from zeros (no NaN garbage)
    uint4* st = reinterpret_cast<uint4*>(s_stage);
    for (int i = threadIdx.x; i < S * p.stage_bytes / 16; i += blockDim.x) st[i] = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      ptx::mbar_init(&full[i], kRowProducerWarps);   // one arrive.expect_tx per producer warp
      ptx::mbar_init(&empty[i], kRowMmaWarps);       // one tcgen05.commit per MMA warp
    }
    for (int i = 0; i < kRowAccStages; ++i) {
      ptx::mbar_init(&tmem_full[i], kRowMmaWarps);
      ptx::mbar_init(&tmem_empty[i], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == kRowProducerWarps) ptx::tmem_alloc(tmem_holder, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

  // contiguous tile range per CTA: coordinates advance incrementally (no divisions in the loops) and a
  // CTA walks down consecutive rows, so two of its three halo rows were just read by itself (L2 hits)
  const int t_begin = (int)((long long)p.total_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)p.total_tiles * (blockIdx.x + 1) / gridDim.x);

  if (warp < kRowProducerWarps) {
    // ================================ bulk-copy producers ================================
    // A lone warp issues ~1 dependent instruction per 8-10 cycles, so the roles are spread: producer warp w
    // owns halo row w of every slab; its lanes (chunk kc, column parity) issue the row's 2 (4) bulk copies.
    const int ry = warp, kc = lane & 1, par = lane >> 1;
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t stage0 = ptx::smem_u32(s_stage);
    TileCoord tc;
    tc.init(t_begin, p.tiles_x, p.OH);
    for (int tile = t_begin; tile < t_end; ++tile, tc.next(p.tiles_x, p.OH)) {
      const int b0 = tc.xb * 128;
#pragma unroll 1
      for (int pi = 0; pi < p.nparts; ++pi) {
        const RowPart& pt = p.part[pi];
        const int mode = pt.mode;
        const int ncopies = (mode == 2) ? 4 : 2;
        const int sy = ((mode == 1) ? (tc.y >> 1) : tc.y) - 1 + ry;            // in [-1, H]: the padded layout holds it
        // halo col c <-> source column b0 - 8 + c = entry b0 + c of the padded row: source and destination
        // are both 128-byte aligned and the length is a multiple of 128 B (Wrow is a multiple of 8)
        const int n_ent = min(kRowHaloCols, pt.d.Wrow - b0);
        const uint32_t bytes = (uint32_t)n_ent * 16u;
        const bool active = lane < ncopies;
        const uint32_t dst_l = (uint32_t)(((active ? par : 0) * 2 + kc) * kRowPlaneBytes + ry * (kRowHaloCols * 16));
        const uint8_t* src_l = pt.base + pt.d.row_off(tc.n, sy, kc, (mode == 2 && active) ? par : 0) + (size_t)b0 * 16;
        const size_t slab_step = (size_t)2 * pt.d.P * pt.d.Wrow * 16;            // two 8-channel chunks further
        for (int j = 0; j < pt.nslabs; ++j) {
          ptx::mbar_wait(&empty[stage], phase ^ 1u, p.error_flag, 11);
          if (lane == 0) ptx::mbar_expect_tx(&full[stage], bytes * (uint32_t)ncopies);
          __syncwarp();
          if (active) ptx::bulk_g2s(stage0 + (uint32_t)stage * p.stage_bytes + dst_l, src_l + j * slab_step, bytes, &full[stage]);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp < kRowProducerWarps + kRowMmaWarps) {
    // ================================ MMA issuers (2 warps) =============================
    // MMA warp m owns accumulator m of every tile: the column-parity group g = m when G == 2, the partial
    // sum over taps of parity m when G == 1 (the epilogue adds the two partials).  Consecutive MMAs into one
    // accumulator serialise, two issuing threads interleave naturally.  All operands come from kernel
    // parameters / constants (uniform datapath); the TMEM base is made warp-uniform with a redux.
    const int m = warp - kRowProducerWarps;
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_holder);
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16<BN>();
      const uint64_t a_desc0 = make_nosw_desc(ptx::smem_u32(s_stage), kRowPlaneBytes, 128);
      const uint64_t b_desc0 = make_nosw_desc(ptx::smem_u32(s_w), BN * 16, 128);
      const uint32_t stage_units = (uint32_t)p.stage_bytes >> 4;
      const int g = (G == 2) ? m : 0;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      TileCoord tc;
      tc.init(t_begin, p.tiles_x, p.OH);
      for (int tile = t_begin; tile < t_end; ++tile, tc.next(p.tiles_x, p.OH)) {
        const int py = tc.y & 1;
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u, p.error_flag, 12);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)((acc * 2 + m) * BN);
        int sl = 0;
#pragma unroll 1
        for (int pi = 0; pi < p.nparts; ++pi) {
          const int mode = p.part[pi].mode;
          for (int j = 0; j < p.part[pi].nslabs; ++j, ++sl) {
            const uint64_t a_st = a_desc0 + (uint64_t)((uint32_t)stage * stage_units);
            const uint64_t b_sl = b_desc0 + (uint64_t)(p.slab_wblock[sl] * 2 * BN);
            ptx::mbar_wait(&full[stage], phase, p.error_flag, 13);
            ptx::tc_fence_after();
            if (mode == 1) {
              // nearest x2 source (G == 2): the 9 taps touch only 2 x 2 distinct half-resolution positions; their
              // weights were summed per (row parity, column parity) on the host -> 4 MMAs instead of 9
#pragma unroll
              for (int pos = 0; pos < 4; ++pos)
                ptx::umma_bf16(d_tmem, a_st + (uint64_t)p.adelta_up[py][g][pos], b_sl + (uint64_t)((((py * 2 + g) * 4) + pos) * 2 * BN), idesc,
                               (uint32_t)((sl | pos) != 0));
            } else if (G == 2) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap)
                ptx::umma_bf16(d_tmem, a_st + (uint64_t)p.adelta[mode][py][g][tap], b_sl + (uint64_t)(tap * 2 * BN), idesc,
                               (uint32_t)((sl | tap) != 0));
            } else {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap)
                if ((tap & 1) == m)
                  ptx::umma_bf16(d_tmem, a_st + (uint64_t)p.adelta[mode][py][0][tap], b_sl + (uint64_t)(tap * 2 * BN), idesc,
                                 (uint32_t)(sl != 0 || tap >= 2));
            }
            ptx::umma_commit(&empty[stage]);
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
        }
        ptx::umma_commit(&tmem_full[acc]);
        if (++acc == kRowAccStages) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ================================ epilogue (2 x 4 warps) ============================
    // Two groups of 4 warps alternate tiles (group e drains accumulator stages e, e+2, ..).  Columns are
    // drained 16 at a time: sum of the two partial accumulators (G == 1), folded BN, residual, ReLU, store.
    // Folded-BN constants and the fused 1x1 head live in registers (BN <= 32) or smem (BN == 64).
    const uint32_t tmem_base = *tmem_holder;
    const int q = warp & 3;
    const int egrp = (warp - kRowProducerWarps - kRowMmaWarps) >> 2;
    const int row = q * 32 + lane;
    constexpr int RB = (BN <= 32) ? BN : 1;
    float r_scale[RB], r_bias[RB];
    if (BN <= 32) {
#pragma unroll
      for (int j = 0; j < RB; ++j) { r_scale[j] = s_scale[j]; r_bias[j] = s_bias[j]; }
    }
    float r_hw[HEAD ? 64 : 1], r_hb[HEAD ? 4 : 1];
    if (HEAD) {
#pragma unroll
      for (int j = 0; j < 64; ++j) r_hw[HEAD ? j : 0] = s_hw[j];
#pragma unroll
      for (int j = 0; j < 4; ++j) r_hb[HEAD ? j : 0] = s_hb[j];
    }
    const float lo = p.relu ? 0.f : -INFINITY;
    const bool planar_out = (p.out_layout == LAYOUT_PLANAR);
    const bool planar_res = (p.res_layout == LAYOUT_PLANAR);
    const size_t chunk_step = (size_t)p.od.Wrow * 16;
    int acc = egrp;
    uint32_t acc_phase = 0;
    TileCoord tc;
    tc.init(t_begin + egrp, p.tiles_x, p.OH);
    for (int tile = t_begin + egrp; tile < t_end; tile += kRowEpiGroups, tc.advance(kRowEpiGroups, p.tiles_x, p.OH)) {
      const size_t rowpix = ((size_t)tc.n * p.OH + tc.y) * p.OW;
      const size_t prow = p.od.row_off(tc.n, tc.y, 0, 0);
      ptx::mbar_wait(&tmem_full[acc], acc_phase, p.error_flag, 14);
      ptx::tc_fence_after();
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int x = (G == 2) ? (2 * (tc.xb * 128 + row) + g) : (tc.xb * 128 + row);
        const bool valid = x < p.OW;
        const size_t pix = rowpix + x;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * G + g) * NA * BN);
        // byte offsets of this pixel's first 8-channel chunk and the step to the next chunk
        const size_t o_off = planar_out ? prow + (size_t)(x + kRowPad) * 16 : pix * (size_t)(BN * 2);
        const size_t o_step = planar_out ? chunk_step : 16;
        const size_t r_off = planar_res ? prow + (size_t)(x + kRowPad) * 16 : pix * (size_t)(BN * 2);
        const size_t r_step = planar_res ? chunk_step : 16;
        float4 hacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < BN; c += 16) {
          uint32_t v[NA * 16];
          ptx::tmem_ld16(t_row + (uint32_t)c, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          if (NA == 2) ptx::tmem_ld16(t_row + (uint32_t)(BN + c), *reinterpret_cast<uint32_t(*)[16]>(&v[NA == 2 ? 16 : 0]));
          uint4 rv[2];
          const bool has_res = (p.res != nullptr) && valid;
          if (has_res) {
            rv[0] = __ldg(reinterpret_cast<const uint4*>(p.res + r_off + (size_t)(c / 8) * r_step));
            rv[1] = __ldg(reinterpret_cast<const uint4*>(p.res + r_off + (size_t)(c / 8 + 1) * r_step));
          }
          ptx::tmem_ld_wait();
          float yv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = (NA == 2) ? (__uint_as_float(v[j]) + __uint_as_float(v[(NA == 2 ? 16 : 0) + j])) : __uint_as_float(v[j]);
            const float sc = (BN <= 32) ? r_scale[(BN <= 32) ? c + j : 0] : s_scale[c + j];
            const float bi = (BN <= 32) ? r_bias[(BN <= 32) ? c + j : 0] : s_bias[c + j];
            yv[j] = fmaf(a, sc, bi);
          }
          if (has_res) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint32_t w[4] = {rv[k].x, rv[k].y, rv[k].z, rv[k].w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                yv[8 * k + 2 * t + 0] += __uint_as_float(w[t] << 16);
                yv[8 * k + 2 * t + 1] += __uint_as_float(w[t] & 0xffff0000u);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) yv[j] = fmaxf(yv[j], lo);
          if (HEAD) {
            float* hp = reinterpret_cast<float*>(&hacc);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // two independent chains per logit (ILP), summed in a fixed order
              float s0 = 0.f, s1 = 0.f;
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                s0 = fmaf(yv[j], r_hw[HEAD ? k * 16 + j : 0], s0);
                s1 = fmaf(yv[j + 1], r_hw[HEAD ? k * 16 + j + 1 : 0], s1);
              }
              hp[k] = (s0 + s1) + r_hb[HEAD ? k : 0];
            }
          }
          if (valid && p.out != nullptr) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              uint32_t w[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(yv[8 * k + 2 * t], yv[8 * k + 2 * t + 1]);
                w[t] = *reinterpret_cast<uint32_t*>(&h2);
              }
              *reinterpret_cast<uint4*>(p.out + o_off + (size_t)(c / 8 + k) * o_step) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
        if (HEAD && valid) reinterpret_cast<float4*>(p.head_out)[pix] = hacc;
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[acc]);
      acc += kRowEpiGroups;
      if (acc >= kRowAccStages) { acc -= kRowAccStages; acc_phase ^= 1u; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kRowProducerWarps) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(*tmem_holder, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// stem kernel (see conv_rowtile.cuh)
// ---------------------------------------------------------------------------------------------
constexpr int kStemThreads = (1 + 2 + 8) * 32;   // producer warp, 2 MMA warps, 2 x 4 epilogue warps

__global__ void __launch_bounds__(kStemThreads, 1) stem_rowtile_kernel(const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  constexpr int kWBytes = 28 * 64 * 16;           // 28 672
  uint8_t* s_w = smem;
  float* s_scale = reinterpret_cast<float*>(smem + kWBytes);
  float* s_bias = s_scale + 64;
  uint8_t* s_stage = smem + kWBytes + 512;
  const int S = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + (size_t)S * kStemStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tmem_full = bars + 2 * S;
  uint64_t* tmem_empty = bars + 2 * S + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * S + 8);
  constexpr int kAcc = 4;                          // accumulator ring (tiles in flight)
  constexpr uint32_t kTmemCols = kAcc * 2 * 64;    // 512: two partial accumulators of 64 columns per tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = threadIdx.x; i < kWBytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) { s_scale[i] = p.scale[i]; s_bias[i] = p.bias[i]; }
    uint4* st = reinterpret_cast<uint4*>(s_stage);
    for (int i = threadIdx.x; i < S * kStemStageBytes / 16; i += blockDim.x) st[i] = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 2);
    }
    for (int i = 0; i < kAcc; ++i) {
      ptx::mbar_init(&tmem_full[i], 2);
      ptx::mbar_init(&tmem_empty[i], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_holder, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

  const int t_begin = (int)((long long)p.total_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)p.total_tiles * (blockIdx.x + 1) / gridDim.x);
  const size_t in_pitch = (size_t)(p.PW + 8) * 8;
  const int row_chunks = (p.PW + 8) / 2;

  if (warp == 0) {
    // ---- producer: lane r copies input row r of the tile (7 bulk copies per tile) ----
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t stage0 = ptx::smem_u32(s_stage);
    TileCoord tc;
    tc.init(t_begin, p.tiles_x, p.OH);
    for (int tile = t_begin; tile < t_end; ++tile, tc.next(p.tiles_x, p.OH)) {
      const int b0 = tc.xb * 128;
      const int n_chunks = min(131, row_chunks - b0);
      const uint32_t bytes = (uint32_t)n_chunks * 16u;
      const bool active = lane < 7;
      const uint8_t* src = p.in + ((size_t)tc.n * (p.PH + 6) + (size_t)(2 * tc.y + (active ? lane : 0))) * in_pitch + (size_t)b0 * 16;
      ptx::mbar_wait(&empty[stage], phase ^ 1u, p.error_flag, 21);
      if (lane == 0) ptx::mbar_expect_tx(&full[stage], bytes * 7u);
      __syncwarp();
      if (active) ptx::bulk_g2s(stage0 + (uint32_t)(stage * kStemStageBytes + lane * kStemRowBytes), src, bytes, &full[stage]);
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
  } else if (warp < 3) {
    // ---- MMA warps: warp m issues the k-steps of parity m into partial accumulator m ----
    const int m = warp - 1;
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_holder);
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16<64>();
      const uint64_t a_desc0 = make_nosw_desc(ptx::smem_u32(s_stage), 16, 128);          // K chunks 16 B apart: overlapping windows
      const uint64_t b_desc0 = make_nosw_desc(ptx::smem_u32(s_w), 64 * 16, 128);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u, p.error_flag, 22);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)((acc * 2 + m) * 64);
        const uint64_t a_st = a_desc0 + (uint64_t)((uint32_t)(stage * kStemStageBytes) >> 4);
        ptx::mbar_wait(&full[stage], phase, p.error_flag, 23);
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 14; ++ks) {
          if ((ks & 1) != m) continue;
          const int r = ks >> 1, j = ks & 1;
          ptx::umma_bf16(d_tmem, a_st + (uint64_t)((r * kStemRowBytes + 2 * j * 16) >> 4), b_desc0 + (uint64_t)((r * 4 + 2 * j) * 64), idesc,
                         (uint32_t)(ks >= 2));
        }
        ptx::umma_commit(&empty[stage]);
        if (++stage == S) { stage = 0; phase ^= 1u; }
        ptx::umma_commit(&tmem_full[acc]);
        if (++acc == kAcc) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ---- epilogue: two groups of 4 warps alternate tiles ----
    const uint32_t tmem_base = *tmem_holder;
    const int q = warp & 3;
    const int egrp = (warp - 3) >> 2;
    const int row = q * 32 + lane;
    const bool planar = (p.out_layout == LAYOUT_PLANAR_PARITY);
    int acc = egrp;
    uint32_t acc_phase = 0;
    TileCoord tc;
    tc.init(t_begin + egrp, p.tiles_x, p.OH);
    for (int tile = t_begin + egrp; tile < t_end; tile += 2, tc.advance(2, p.tiles_x, p.OH)) {
      const int x = tc.xb * 128 + row;
      const bool valid = x < p.OW;
      ptx::mbar_wait(&tmem_full[acc], acc_phase, p.error_flag, 24);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * 64);
      uint8_t* obase;
      size_t cstep;
      if (planar) {
        obase = p.out + p.od.row_off(tc.n, tc.y, 0, x & 1) + (size_t)((x >> 1) + kRowPad) * 16;
        cstep = (size_t)2 * p.od.Wrow * 16;                    // next 8-channel chunk (both parity runs)
      } else {
        obase = p.out + (((size_t)tc.n * p.OH + tc.y) * p.OW + x) * 128;
        cstep = 16;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {                              // 32 output channels at a time
        uint32_t v0[32], v1[32];
        ptx::tmem_ld16(t_row + (uint32_t)(h * 32), *reinterpret_cast<uint32_t(*)[16]>(&v0[0]));
        ptx::tmem_ld16(t_row + (uint32_t)(h * 32 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v0[16]));
        ptx::tmem_ld16(t_row + (uint32_t)(64 + h * 32), *reinterpret_cast<uint32_t(*)[16]>(&v1[0]));
        ptx::tmem_ld16(t_row + (uint32_t)(64 + h * 32 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v1[16]));
        ptx::tmem_ld_wait();
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          y[j] = fmaxf(fmaf(__uint_as_float(v0[j]) + __uint_as_float(v1[j]), p.k.scale[h * 32 + j], p.k.bias[h * 32 + j]), 0.f);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(y[8 * j + 2 * t], y[8 * j + 2 * t + 1]);
              w[t] = *reinterpret_cast<uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(obase + (size_t)(h * 4 + j) * cstep) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[acc]);
      acc += 2;
      if (acc >= kAcc) { acc -= kAcc; acc_phase ^= 1u; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(*tmem_holder, kTmemCols);
  }
}

void RowStemOp::build(const void* padded_tiles, int n, int ph, int pw, const float* w_oihw, const float* scale, const float* bias,
                      void* out, int out_layout, int* error_flag, int num_sms) {
  WSI_REQUIRE(ph % 2 == 0 && pw % 2 == 0, WSI_ERR_UNSUPPORTED, "stem: tile size must be even");
  WSI_REQUIRE(out_layout == LAYOUT_NHWC || out_layout == LAYOUT_PLANAR_PARITY, WSI_ERR_INVALID, "stem: bad output layout");
  StemParams& p = p_;
  p = StemParams{};
  p.in = static_cast<const uint8_t*>(padded_tiles);
  p.N = n; p.PH = ph; p.PW = pw; p.OH = ph / 2; p.OW = pw / 2;
  p.tiles_x = (int)ceil_div(p.OW, 128);
  const long long total = (long long)n * p.OH * p.tiles_x;
  WSI_REQUIRE(total < (1LL << 31), WSI_ERR_UNSUPPORTED, "stem: too many tiles");
  p.total_tiles = (int)total;
  p.out = static_cast<uint8_t*>(out);
  p.out_layout = out_layout;
  p.od = PlanarDims::make(p.OH, p.OW, 64, LAYOUT_PLANAR_PARITY);
  p.error_flag = error_flag;
  // weights: k-chunk kc = r*4 + (pixel pair), element e = (pixel in pair)*4 + channel; [kc][cout][8]
  std::vector<uint16_t> wp((size_t)28 * 64 * 8, 0);
  for (int r = 0; r < 7; ++r)
    for (int s = 0; s < 7; ++s)
      for (int c = 0; c < 3; ++c)
        for (int co = 0; co < 64; ++co) {
          const int kc = r * 4 + s / 2, e = (s & 1) * 4 + c;
          wp[((size_t)kc * 64 + co) * 8 + e] = f32_to_bf16_bits(w_oihw[(((size_t)co * 3 + c) * 7 + r) * 7 + s]);
        }
  upload(w_, wp);
  std::vector<float> sc(64, 1.f), bi(64, 0.f);
  if (scale) sc.assign(scale, scale + 64);
  if (bias) bi.assign(bias, bias + 64);
  upload(scale_, sc);
  upload(bias_, bi);
  p.w = w_.as<bf16>(); p.scale = scale_.as<float>(); p.bias = bias_.as<float>();
  for (int jj = 0; jj < 64; ++jj) { p.k.scale[jj] = sc[jj]; p.k.bias[jj] = bi[jj]; }
  flops_ = 2.0 * n * p.OH * p.OW * 64.0 * 147.0;
  p.stages = 8;
  smem_ = 128 + 28 * 64 * 16 + 512 + p.stages * kStemStageBytes + 256;
  grid_ = (int)std::min<long long>(total, num_sms);
  CUDA_CHECK(cudaStreamSynchronize(0));
}

void RowStemOp::launch(cudaStream_t stream, LaunchCounter* lc) const {
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(stem_rowtile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  stem_rowtile_kernel<<<grid_, kStemThreads, smem_, stream>>>(p_);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// ---------------------------------------------------------------------------------------------
// max pool 3x3/s2/p1 over the parity-planar stem output -> NHWC (resnets_shift.py:126)
// one thread = 8 channels of one output pixel; all nine taps are in range thanks to the zero border
// ---------------------------------------------------------------------------------------------
// One block = one (image, output row, 8-channel chunk) run: the index arithmetic is per block (no per-thread 64-bit
// divisions), the nine taps are reduced with packed bf16 max (HMNMX2.BF16: exact, inputs are bf16) — ncu had the first
// version at 67 % issue utilisation with 58 % DRAM: ~150 unpack / max instructions and three 64-bit divisions per 16 bytes.
__global__ void __launch_bounds__(128) maxpool_planar_kernel(const uint8_t* __restrict__ x, int n, int h, int w, int kcs,
                                                              uint8_t* __restrict__ y, int y_layout) {
  const PlanarDims d = PlanarDims::make(h, w, kcs * 8, LAYOUT_PLANAR_PARITY);
  const int oh = (h - 1) / 2 + 1, ow = (w - 1) / 2 + 1;
  const PlanarDims od = PlanarDims::make(oh, ow, kcs * 8, LAYOUT_PLANAR);
  const int64_t runs = (int64_t)n * oh * kcs;
  for (int64_t run = blockIdx.x; run < runs; run += gridDim.x) {
    const int g = (int)(run % kcs);
    const int64_t r = run / kcs;
    const int oy = (int)(r % oh);
    const int b = (int)(r / oh);
    const uint8_t* even[3];
    const uint8_t* odd[3];
    bool row_ok[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int iy = 2 * oy + k - 1;
      row_ok[k] = iy < h;                            // row -1 is the layout's zero border; inputs are post-ReLU (>= 0)
      even[k] = x + d.row_off(b, row_ok[k] ? iy : 0, g, 0);
      odd[k] = x + d.row_off(b, row_ok[k] ? iy : 0, g, 1);
    }
    uint8_t* orow = y + ((y_layout == LAYOUT_PLANAR) ? od.row_off(b, oy, g, 0) + (size_t)kRowPad * 16 : ((((size_t)b * oh + oy) * ow) * kcs + g) * 16);
    const size_t ostep = (y_layout == LAYOUT_PLANAR) ? 16 : (size_t)kcs * 16;
    for (int ox = threadIdx.x; ox < ow; ox += blockDim.x) {
      __nv_bfloat162 m[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) m[t] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (!row_ok[k]) continue;
        // columns 2ox-1 (odd run, half-index ox-1), 2ox (even run, ox), 2ox+1 (odd run, ox)
        const uint4 v[3] = {__ldg(reinterpret_cast<const uint4*>(odd[k] + (size_t)(ox - 1 + kRowPad) * 16)),
                            __ldg(reinterpret_cast<const uint4*>(even[k] + (size_t)(ox + kRowPad) * 16)),
                            __ldg(reinterpret_cast<const uint4*>(odd[k] + (size_t)(ox + kRowPad) * 16))};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const uint32_t ww[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) m[t] = __hmax2(m[t], *reinterpret_cast<const __nv_bfloat162*>(&ww[t]));
        }
      }
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&m[0]); o.y = *reinterpret_cast<uint32_t*>(&m[1]);
      o.z = *reinterpret_cast<uint32_t*>(&m[2]); o.w = *reinterpret_cast<uint32_t*>(&m[3]);
      *reinterpret_cast<uint4*>(orow + (size_t)ox * ostep) = o;
    }
  }
}

void launch_maxpool_planar(const void* x, int n, int h, int w, int c, void* y, int y_layout, cudaStream_t s, LaunchCounter* lc) {
  const int64_t runs = (int64_t)n * ((h - 1) / 2 + 1) * (c / 8);
  if (runs <= 0) return;
  const int grid = (int)std::min<int64_t>(runs, 148 * 64);
  maxpool_planar_kernel<<<grid, 128, 0, s>>>(static_cast<const uint8_t*>(x), n, h, w, c / 8, static_cast<uint8_t*>(y), y_layout);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// ---------------------------------------------------------------------------------------------
// NHWC -> planar relayout (one thread per 16-byte chunk, chunk index fastest: reads are fully
// coalesced, writes are 16-byte entries of C/8 planar rows)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) relayout_planar_kernel(const uint4* __restrict__ src, uint8_t* __restrict__ dst, int N, int H,
                                                               int W, int KC, int layout) {
  const PlanarDims d = PlanarDims::make(H, W, KC * 8, layout);
  const int64_t total = (int64_t)N * H * KC * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kc = (int)(i % KC);
    int64_t r = i / KC;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int n = (int)(r / H);
    const uint4 v = __ldg(src + i);
    size_t off;
    if (layout == LAYOUT_PLANAR_PARITY) off = d.row_off(n, y, kc, x & 1) + (size_t)((x >> 1) + kRowPad) * 16;
    else off = d.row_off(n, y, kc, 0) + (size_t)(x + kRowPad) * 16;
    *reinterpret_cast<uint4*>(dst + off) = v;
  }
}

void launch_relayout_planar(const void* src_nhwc, void* dst, int N, int H, int W, int C, int layout, cudaStream_t s, LaunchCounter* lc) {
  const int64_t total = (int64_t)N * H * (C / 8) * W;
  if (total <= 0) return;
  const int grid = (int)std::min<int64_t>(ceil_div(total, 256), 148 * 32);
  relayout_planar_kernel<<<grid, 256, 0, s>>>(static_cast<const uint4*>(src_nhwc), static_cast<uint8_t*>(dst), N, H, W, C / 8, layout);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// inverse (plain planar -> NHWC): only used where a planar-only kernel must hand an NHWC tensor back (debug entry point)
__global__ void __launch_bounds__(256) relayout_nhwc_kernel(const uint8_t* __restrict__ src, uint4* __restrict__ dst, int N, int H, int W,
                                                             int KC) {
  const PlanarDims d = PlanarDims::make(H, W, KC * 8, LAYOUT_PLANAR);
  const int64_t total = (int64_t)N * H * KC * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kc = (int)(i % KC);
    int64_t r = i / KC;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int n = (int)(r / H);
    dst[i] = *reinterpret_cast<const uint4*>(src + d.row_off(n, y, kc, 0) + (size_t)(x + kRowPad) * 16);
  }
}

void launch_relayout_nhwc(const void* src_planar, void* dst_nhwc, int N, int H, int W, int C, cudaStream_t s, LaunchCounter* lc) {
  const int64_t total = (int64_t)N * H * (C / 8) * W;
  if (total <= 0) return;
  const int grid = (int)std::min<int64_t>(ceil_div(total, 256), 148 * 32);
  relayout_nhwc_kernel<<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(src_planar), static_cast<uint4*>(dst_nhwc), N, H, W, C / 8);
  CUDA_CHECK(cudaGetLastError());
  if (lc) lc->n++;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool RowConvOp::eligible(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const void* residual) {
  if (spec.ksize != 3 || spec.stride != 1 || spec.pad != 1) return false;
  if (spec.cout != 16 && spec.cout != 32 && spec.cout != 64) return false;
  if (spec.cout == 64 && parts[0].up2) return false;            // TMEM: 4 stages x 2 accumulators x 64 columns = 512
  (void)residual;
  if (spec.head && spec.cout != 16) return false;
  if (parts.empty() || parts.size() > 2) return false;
  int slabs = 0;
  for (auto& q : parts) {
    if (q.t.C % 16 != 0 || q.t.C > 64 || q.t.C <= 0) return false;
    slabs += q.t.C / 16;
  }
  if (slabs > kRowMaxSlabs) return false;
  if (parts[0].up2) {
    if (parts.size() == 2 && (parts[1].up2 || parts[1].t.W % 2 != 0)) return false;
    return true;
  }
  return parts.size() == 1;
}

void RowConvOp::build(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const float* w_oihw, const float* scale,
                      const float* bias, const void* residual, int res_layout, void* out, int out_layout, const float* head_w,
                      const float* head_b, float* head_out, int* error_flag, int num_sms) {
  WSI_REQUIRE(eligible(parts, spec, residual), WSI_ERR_UNSUPPORTED, "conv is not eligible for the row-tile kernel");
  WSI_REQUIRE(res_layout == LAYOUT_NHWC || res_layout == LAYOUT_PLANAR, WSI_ERR_INVALID, "row conv: bad residual layout");
  RowParams& p = p_;
  p = RowParams{};
  relayouts_.clear();
  const bool up2 = parts[0].up2;
  const int N = parts[0].t.N;
  const int OH = up2 ? 2 * parts[0].t.H : parts[0].t.H, OW = up2 ? 2 * parts[0].t.W : parts[0].t.W;
  const int BN = spec.cout;
  int cin = 0;
  p.nparts = (int)parts.size();
  p.nslabs = 0;
  bool has_skip = false;
  for (size_t i = 0; i < parts.size(); ++i) {
    const auto& q = parts[i];
    RowPart& rp = p.part[i];
    rp.mode = up2 ? (q.up2 ? 1 : 2) : 0;
    const int want_layout = (rp.mode == 2) ? LAYOUT_PLANAR_PARITY : LAYOUT_PLANAR;
    rp.d = PlanarDims::make(q.t.H, q.t.W, q.t.C, want_layout);
    if (q.t.layout == LAYOUT_NHWC) {
      // operand produced by an NHWC kernel: converted before every launch into a private planar buffer
      stage_in_[i].alloc(rp.d.bytes(N));
      CUDA_CHECK(cudaMemset(stage_in_[i].p, 0, stage_in_[i].bytes));
      relayouts_.push_back(Relayout{q.t.ptr, stage_in_[i].p, N, q.t.H, q.t.W, q.t.C, want_layout});
      rp.base = stage_in_[i].as<uint8_t>();
    } else {
      WSI_REQUIRE(q.t.layout == want_layout, WSI_ERR_INVALID, "row conv: operand %zu has layout %d, needs %d", i, q.t.layout, want_layout);
      rp.base = static_cast<const uint8_t*>(q.t.ptr);
    }
    if (rp.mode == 2) {
      WSI_REQUIRE(q.t.H == OH && q.t.W == OW && q.t.N == N, WSI_ERR_INVALID, "row conv: skip shape mismatch");
      has_skip = true;
    }
    rp.nslabs = q.t.C / 16;
    p.nslabs += rp.nslabs;
    cin += q.t.C;
  }
  p.N = N; p.OH = OH; p.OW = OW; p.Cout = BN; p.up2 = up2 ? 1 : 0;
  p.tiles_x = (int)ceil_div(OW, up2 ? 256 : 128);
  const long long total = (long long)N * OH * p.tiles_x;
  WSI_REQUIRE(total < (1LL << 31), WSI_ERR_UNSUPPORTED, "row conv: too many tiles");
  p.total_tiles = (int)total;
  p.relu = spec.relu ? 1 : 0;
  p.out = static_cast<uint8_t*>(out);
  p.out_layout = out_layout;
  p.res = static_cast<const uint8_t*>(residual);
  p.res_layout = res_layout;
  WSI_REQUIRE(out_layout == LAYOUT_NHWC || out_layout == LAYOUT_PLANAR, WSI_ERR_INVALID, "row conv: bad output layout");
  p.od = PlanarDims::make(OH, OW, BN, LAYOUT_PLANAR);
  p.error_flag = error_flag;

  // A-operand start offsets of every tap relative to the stage base, in 16-byte units
  auto fl2 = [](int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); };
  for (int mode = 0; mode < 3; ++mode)
    for (int py = 0; py < 2; ++py)
      for (int g = 0; g < 2; ++g)
        for (int tap = 0; tap < 9; ++tap) {
          const int r = tap / 3, s = tap % 3;
          int par = 0, poff;
          // halo col c holds source column b0 - kRowPad + c (parity planes: half-index b0 - kRowPad + c)
          if (mode == 0) {
            poff = r * kRowHaloCols + (s - 1 + kRowPad);
          } else if (mode == 1) {
            poff = (fl2(py + r - 1) + 1) * kRowHaloCols + (fl2(g + s - 1) + kRowPad);
          } else {
            const int q = g + s - 1;          // source column 2(b0+i) + q: parity q&1, half-index b0 + i + floor(q/2)
            par = q & 1;
            poff = r * kRowHaloCols + (fl2(q) + kRowPad);
          }
          p.adelta[mode][py][g][tap] = (uint16_t)((par * 2 * kRowPlaneBytes + poff * 16) >> 4);
        }

  // weights: per slab a run of [block][2 chunks][BN][8] bf16; slabs enumerate the concatenated input channels in
  // order.  Plain / skip slabs: 9 blocks = the 9 taps.  Nearest-x2 slabs: 16 blocks = (row parity, column
  // parity) x the 4 distinct half-resolution positions, each the fp32 sum of the taps that read that position.
  std::vector<uint16_t> wp;
  int wblock = 0;
  {
    int sl = 0;
    for (int pi = 0; pi < p.nparts; ++pi)
      for (int j = 0; j < p.part[pi].nslabs; ++j, ++sl) {
        p.slab_wblock[sl] = wblock;
        const int nblk = (p.part[pi].mode == 1) ? 16 : 9;
        wp.resize((size_t)(wblock + nblk) * 2 * BN * 8);
        for (int b = 0; b < nblk; ++b)
          for (int ch = 0; ch < 2; ++ch)
            for (int n = 0; n < BN; ++n)
              for (int e = 0; e < 8; ++e) {
                const int ci = sl * 16 + ch * 8 + e;
                float v = 0.f;
                if (p.part[pi].mode != 1) {
                  v = w_oihw[(((size_t)n * cin + ci) * 3 + b / 3) * 3 + b % 3];
                } else {
                  const int py = b >> 3, g = (b >> 2) & 1, pos = b & 3;
                  for (int tap = 0; tap < 9; ++tap) {
                    const int r = tap / 3, sft = tap % 3;
                    const int dy = fl2(py + r - 1) - fl2(py - 1), dx = fl2(g + sft - 1) - fl2(g - 1);
                    if (dy * 2 + dx == pos) v += w_oihw[(((size_t)n * cin + ci) * 3 + r) * 3 + sft];
                  }
                }
                wp[((((size_t)(wblock + b)) * 2 + ch) * BN + n) * 8 + e] = f32_to_bf16_bits(v);
              }
        wblock += nblk;
      }
  }
  p.w_blocks = wblock;
  // A start offsets of the 4 collapsed positions of a nearest-x2 slab
  for (int py = 0; py < 2; ++py)
    for (int g = 0; g < 2; ++g)
      for (int pos = 0; pos < 4; ++pos) {
        const int ry = fl2(py - 1) + (pos >> 1) + 1, cx = fl2(g - 1) + (pos & 1) + kRowPad;
        p.adelta_up[py][g][pos] = (uint16_t)(ry * kRowHaloCols + cx);
      }
  upload(w_, wp);
  std::vector<float> sc(BN, 1.f), bi(BN, 0.f);
  if (scale) sc.assign(scale, scale + BN);
  if (bias) bi.assign(bias, bias + BN);
  upload(scale_, sc);
  upload(bias_, bi);
  p.w = w_.as<bf16>(); p.scale = scale_.as<float>(); p.bias = bias_.as<float>();
  flops_ = 2.0 * N * OH * OW * (double)BN * cin * 9;
  if (spec.head) {
    WSI_REQUIRE(head_w && head_b && head_out, WSI_ERR_INVALID, "fused head needs weights and an output");
    std::vector<float> hw(head_w, head_w + 64), hb(head_b, head_b + 4);
    upload(headw_, hw);
    upload(headb_, hb);
    p.head_w = headw_.as<float>(); p.head_b = headb_.as<float>(); p.head_out = head_out;
    flops_ += 2.0 * N * OH * OW * 16 * 4;
  }
  p.stage_bytes = (has_skip ? 4 : 2) * kRowPlaneBytes;
  const int w_bytes = p.w_blocks * 2 * BN * 16;
  const int fixed = 128 + ((w_bytes + (2 * BN + 68) * 4 + 127) & ~127) + 512;
  int max_stages = 8;
  if (const char* e = getenv("WSI_ROW_STAGES")) max_stages = std::max(2, atoi(e));
  p.stages = std::min(max_stages, (226 * 1024 - fixed) / p.stage_bytes);
  WSI_REQUIRE(p.stages >= 2, WSI_ERR_UNSUPPORTED, "row conv: not enough shared memory for 2 stages");
  smem_ = fixed + p.stages * p.stage_bytes;
  grid_ = (int)std::min<long long>(total, num_sms);
  CUDA_CHECK(cudaStreamSynchronize(0));
}

template <int BN, bool HEAD, int G>
static void launch_row(const RowParams& p, int grid, int smem, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(conv_rowtile_kernel<BN, HEAD, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured = true;
  }
  conv_rowtile_kernel<BN, HEAD, G><<<grid, kRowThreads, smem, s>>>(p);
  CUDA_CHECK(cudaGetLastError());
}

void RowConvOp::launch(cudaStream_t stream, LaunchCounter* lc) const {
  for (auto& r : relayouts_) launch_relayout_planar(r.src, r.dst, r.N, r.H, r.W, r.C, r.layout, stream, lc);
  const bool head = p_.head_out != nullptr;
  if (p_.Cout == 16) {
    if (p_.up2) { if (head) launch_row<16, true, 2>(p_, grid_, smem_, stream); else launch_row<16, false, 2>(p_, grid_, smem_, stream); }
    else        { if (head) launch_row<16, true, 1>(p_, grid_, smem_, stream); else launch_row<16, false, 1>(p_, grid_, smem_, stream); }
  } else if (p_.Cout == 32) {
    if (p_.up2) launch_row<32, false, 2>(p_, grid_, smem_, stream);
    else        launch_row<32, false, 1>(p_, grid_, smem_, stream);
  } else {
    launch_row<64, false, 1>(p_, grid_, smem_, stream);
  }
  if (lc) lc->n++;
}

}  // namespace wsi
