// conv_rowstream.cu — see conv_rowstream.cuh.
#include "conv_rowstream.cuh"

#include <algorithm>
#include <cstdlib>

namespace wsi {

namespace sptx {
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace sptx

__device__ __forceinline__ uint64_t stream_nosw_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// instruction descriptor: D fp32, A/B bf16 K-major, M = 128, N given at run time
__device__ __forceinline__ uint32_t stream_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}

// Work split: the (strip, output row) space — strips of 128 output columns, strip-major — is cut into gridDim.x
// equal contiguous ranges, so every CTA streams the same number of rows (+ 2 halo rows per strip it touches).
// A unit is the part of one strip inside the CTA's range.
struct UnitIter {
  int g, g_end, OH, tiles_x;
  int n, xb, y0, Lu;
  __device__ __forceinline__ void init(const StreamParams& p) {
    g = (int)((long long)p.total_rows * blockIdx.x / gridDim.x);
    g_end = (int)((long long)p.total_rows * (blockIdx.x + 1) / gridDim.x);
    OH = p.OH;
    tiles_x = p.tiles_x;
  }
  __device__ __forceinline__ bool next() {
    if (g >= g_end) return false;
    const int strip = g / OH;
    y0 = g - strip * OH;
    Lu = min(OH - y0, g_end - g);
    n = strip / tiles_x;
    xb = strip - n * tiles_x;
    g += Lu;
    return true;
  }
};

template <int BN, bool HEAD, int RS>
__global__ void __launch_bounds__(kStreamThreads, (BN <= 32) ? 2 : 1) conv_rowstream_kernel(const __grid_constant__ StreamParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int w_bytes = p.nslabs * 3 * 2 * 3 * BN * 16;
  uint8_t* s_w = smem;
  float* s_scale = reinterpret_cast<float*>(smem + w_bytes);
  float* s_bias = s_scale + BN;
  float* s_hw = s_bias + BN;
  float* s_hb = s_hw + 64;
  uint8_t* s_stage = smem + ((w_bytes + (2 * BN + 68) * 4 + 127) & ~127);
  const int S = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + (size_t)S * (p.nslabs * kStreamStageBytes));
  uint64_t* full = bars;                 // producer -> MMA: input row landed
  uint64_t* empty = bars + S;            // MMA -> producer AND epilogue: all MMAs of this input row retired
  uint64_t* slot_free = bars + 2 * S;    // epilogue -> MMA: accumulator slot drained
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * S + RS);
  const int stage_bytes = p.nslabs * kStreamStageBytes;
  constexpr uint32_t kTmemCols = RS * BN;
  static_assert(kTmemCols >= 32 && kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns must be a power of two");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = threadIdx.x; i < w_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < BN; i += blockDim.x) { s_scale[i] = p.scale[i]; s_bias[i] = p.bias[i]; }
    if (HEAD)
      for (int i = threadIdx.x; i < 68; i += blockDim.x) s_hw[i] = (i < 64) ? p.head_w[i] : p.head_b[i - 64];
    uint4* st = reinterpret_cast<uint4*>(s_stage);
    for (int i = threadIdx.x; i < S * stage_bytes / 16; i += blockDim.x) st[i] = make_uint4(0, 0, 0, 0);
    sptx::fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < RS; ++i) ptx::mbar_init(&slot_free[i], 128);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_holder, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // every accumulator slot starts (and is handed back by the epilogue) zeroed: see the MMA issuer's interior path
  if (warp >= 2) {
    ptx::tmem_zero_all(*tmem_holder, warp, 2, (int)kTmemCols);
    ptx::tc_fence_before();
  }
  __syncthreads();
  ptx::tc_fence_after();


  if (warp == 0) {
    // ================================ producer ==========================================
    // one pipeline stage = one input row, all slabs (2 chunk runs per 16-channel slab).  One elected lane issues
    // everything: its operands stay on the uniform datapath and the per-row instruction stream is short — the
    // three roles are each a single dependent instruction chain, which is what bounds the small-channel layers.
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage0 = ptx::smem_u32(s_stage);
      uint32_t dst = stage0;
      const size_t run_step = (size_t)p.d.Wrow * 16;
      const size_t row_step = (size_t)p.d.KC * run_step;
      const int nruns = 2 * p.nslabs;
      UnitIter uc;
      uc.init(p);
      while (uc.next()) {
        const int b0 = uc.xb * 128;
        const uint32_t bytes = (uint32_t)min(kRowHaloCols, p.d.Wrow - b0) * 16u;
        const uint32_t row_tx = (WSI_DBG(p) == 2) ? 0u : (uint32_t)nruns * bytes;
        const uint8_t* rowp = p.in + p.d.row_off(uc.n, uc.y0 - 1, 0, 0) + (size_t)b0 * 16;
        const int rows = uc.Lu + 2;
        for (int t = 0; t < rows; ++t) {
          ptx::mbar_wait(&empty[stage], phase ^ 1u, p.error_flag, 31);
          ptx::mbar_expect_tx(&full[stage], row_tx);
          if (WSI_DBG(p) != 2) {
            const uint8_t* src = rowp;
            uint32_t d = dst;
#pragma unroll 2
            for (int r = 0; r < nruns; ++r) {
              sptx::bulk_g2s(d, src, bytes, &full[stage]);
              d += kRowRunBytes;
              src += run_step;
            }
          }
          rowp += row_step;
          dst += (uint32_t)stage_bytes;
          if (++stage == S) { stage = 0; phase ^= 1u; dst = stage0; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ========================================
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_holder);
    if (ptx::elect_one()) {
      constexpr uint32_t RM = RS - 1;                              // RS is a power of two
      const uint64_t a_desc0 = stream_nosw_desc(ptx::smem_u32(s_stage), kRowRunBytes, 128) + (uint64_t)(kRowPad - 1);
      const uint64_t b_desc0 = stream_nosw_desc(ptx::smem_u32(s_w), 3 * BN * 16, 128);
      const uint32_t stage_units = (uint32_t)stage_bytes >> 4;
      constexpr uint32_t kSlabUnits = kStreamStageBytes >> 4;      // A: next 16-channel slab of the row
      constexpr uint32_t kWSlab = 3 * 2 * 3 * BN, kWShift = 2 * 3 * BN;   // B: next slab / next horizontal shift (16-byte units)
      const uint32_t id1 = stream_idesc(BN), id2 = stream_idesc(2 * BN), id3 = stream_idesc(3 * BN);
      const int nslabs = p.nslabs;
      const bool no_mma = (WSI_DBG(p) == 1);
      int stage = 0;
      uint32_t phase = 0, a_off = 0;                   // a_off: stage * stage_units
      uint32_t jb = 0;                                 // output-row jobs issued by this CTA before the current unit
      // blocks [a, b] of the stacked weights (block 2-r <-> vertical tap r) into the slots of jobs jrow_a ...
      auto issue = [&](uint64_t a_desc, uint64_t b_desc, uint32_t jrow_a, int a, int b, uint32_t accumulate) {
        int cnt = b - a + 1;
        int slot = (int)(jrow_a & RM);
        int blk = a;
        while (cnt > 0) {
          const int n_here = min(cnt, RS - slot);      // split where the ring wraps
          ptx::umma_bf16(tmem_base + (uint32_t)(slot * BN), a_desc, b_desc + (uint64_t)(blk * BN), stream_idesc(n_here * BN), accumulate);
          cnt -= n_here;
          blk += n_here;
          slot = 0;
        }
      };
      UnitIter uc;
      uc.init(p);
      while (uc.next()) {
        const int Lu = uc.Lu;
        for (int t = 0; t < Lu + 2; ++t) {
          const bool opens = (t < Lu);                 // output row i = t gets its first contribution from this input row
          const uint32_t j = jb + (uint32_t)t;
          if (opens) ptx::mbar_wait(&slot_free[j & RM], ((j / RS) & 1u) ^ 1u, p.error_flag, 32);
          const uint64_t a_row = a_desc0 + (uint64_t)a_off;
          ptx::mbar_wait(&full[stage], phase, p.error_flag, 33);
          ptx::tc_fence_after();
          const uint32_t slot_lo = (j - 2u) & RM;
          if (no_mma) {
          } else if (WSI_DBG(p) == 4 && BN == 16) {
            // timing experiment: the same 4 MMAs, each into its own TMEM region -> no dependent chain inside a row
            ptx::umma_bf16(tmem_base, a_row, b_desc0 + (uint64_t)(2 * BN), id1, 1u);
            ptx::umma_bf16(tmem_base + 48, a_row, b_desc0, id2, 1u);
            ptx::umma_bf16(tmem_base + 96, a_row + 1, b_desc0 + (uint64_t)kWShift, id3, 1u);
            ptx::umma_bf16(tmem_base + 144, a_row + 2, b_desc0 + (uint64_t)(2 * kWShift), id3, 1u);
          } else if (t >= 2 && opens && slot_lo <= RS - 3) {
            // fast path (interior row, the three target slots are contiguous): 3 MMAs of N = 3*BN per slab.  The slot
            // of the newly opened row was zeroed by the epilogue when it drained it (or at kernel start), so its first
            // contribution accumulates like the others — no separate overwriting MMA (every MMA costs a 4 KB A fetch
            // and a place in the dependent chain whatever its N)
            const uint32_t d_lo = tmem_base + slot_lo * BN;
            ptx::umma_bf16(d_lo, a_row, b_desc0, id3, 1u);
            ptx::umma_bf16(d_lo, a_row + 1, b_desc0 + (uint64_t)kWShift, id3, 1u);
            ptx::umma_bf16(d_lo, a_row + 2, b_desc0 + (uint64_t)(2 * kWShift), id3, 1u);
            uint64_t a_sl = a_row, b_sl = b_desc0;
#pragma unroll 1
            for (int sl = 1; sl < nslabs; ++sl) {
              a_sl += kSlabUnits;
              b_sl += kWSlab;
              ptx::umma_bf16(d_lo, a_sl, b_sl, id3, 1u);
              ptx::umma_bf16(d_lo, a_sl + 1, b_sl + (uint64_t)kWShift, id3, 1u);
              ptx::umma_bf16(d_lo, a_sl + 2, b_sl + (uint64_t)(2 * kWShift), id3, 1u);
            }
          } else {
            const int i_lo = max(0, t - 2), i_hi = min(Lu - 1, t);
            const int blk_lo = 2 - t + i_lo, blk_hi = 2 - t + i_hi;
            for (int sl = 0; sl < nslabs; ++sl) {
              for (int sft = 0; sft < 3; ++sft) {
                const uint64_t a_desc = a_row + (uint64_t)(sl * kSlabUnits + sft);
                const uint64_t b_desc = b_desc0 + (uint64_t)(sl * kWSlab + sft * kWShift);
                if (sl == 0 && sft == 0 && opens) {
                  issue(a_desc, b_desc, j, 2, 2, 0u);                                    // overwrite the new row
                  if (blk_lo <= 1) issue(a_desc, b_desc, jb + (uint32_t)i_lo, blk_lo, 1, 1u);
                } else {
                  issue(a_desc, b_desc, jb + (uint32_t)i_lo, blk_lo, blk_hi, 1u);
                }
              }
            }
          }
          // ONE commit per input row: it frees the stage for the producer and tells the epilogue that output
          // row i = t-2 is complete (a tcgen05.commit drains the tensor pipe — ~600 cycles measured — so the
          // number of commits, not of MMAs, bounded the previous schedules)
          ptx::umma_commit(&empty[stage]);
          a_off += stage_units;
          if (++stage == S) { stage = 0; phase ^= 1u; a_off = 0; }
        }
        jb += (uint32_t)Lu;
      }
    }
  } else {
    // ================================ epilogue (2 x 4 warps) ============================
    const uint32_t tmem_base = *tmem_holder;
    const int q = warp & 3;
    const int egrp = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    // folded-BN constants and the fused head are read from the parameter (constant) bank: p.k
    const float lo = p.relu ? 0.f : -INFINITY;
    const bool planar_out = (p.out_layout == LAYOUT_PLANAR);
    const bool planar_res = (p.res_layout == LAYOUT_PLANAR);
    const size_t chunk_step = (size_t)p.od.Wrow * 16;
    const size_t plane_row_step = (size_t)p.od.KC * p.od.P * chunk_step;      // next image row of a planar tensor
    const size_t o_step = planar_out ? chunk_step : 16, o_inc = planar_out ? plane_row_step : (size_t)p.OW * (BN * 2);
    const size_t r_step = planar_res ? chunk_step : 16, r_inc = planar_res ? plane_row_step : (size_t)p.OW * (BN * 2);
    uint32_t job = 0;
    // stage / phase of the input row whose commit completes the next output row (input row i+2 for output row i);
    // everything per row is incremental — this loop is one dependent instruction chain per warp
    int gs = 2;
    uint32_t gph = 0;
    UnitIter uc;
    uc.init(p);
    while (uc.next()) {
      const int Lu = uc.Lu;
      const int x = uc.xb * 128 + row;
      const bool valid = x < p.OW;
      size_t pix = ((size_t)uc.n * p.OH + uc.y0) * p.OW + x;
      const size_t prow = p.od.row_off(uc.n, uc.y0, 0, 0) + (size_t)(x + kRowPad) * 16;
      size_t o_off = planar_out ? prow : pix * (size_t)(BN * 2);
      size_t r_off = planar_res ? prow : pix * (size_t)(BN * 2);
      const bool has_res = (p.res != nullptr) && valid;
      for (int i = 0; i < Lu; ++i, ++job, pix += (size_t)p.OW, o_off += o_inc, r_off += r_inc) {
        const int gs_row = gs;
        const uint32_t gph_row = gph;
        if (++gs == S) { gs = 0; gph ^= 1u; }
        if ((job & 1u) != (uint32_t)egrp) continue;
        const uint32_t slot = job & (uint32_t)(RS - 1);
        // the whole residual row of this pixel is requested before waiting for the accumulator
        uint4 rall[BN / 8];
        if (has_res) {
#pragma unroll
          for (int k = 0; k < BN / 8; ++k) rall[k] = __ldg(reinterpret_cast<const uint4*>(p.res + r_off + (size_t)k * r_step));
        }
        // output row i is complete when the MMAs of input row t = i+2 have retired: that is the commit on that
        // row's stage barrier (the ring cannot lap: the MMA warp needs this slot back before it gets S rows ahead)
        ptx::mbar_wait(&empty[gs_row], gph_row, p.error_flag, 34);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * BN);
        float4 hacc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (WSI_DBG(p) != 3)
#pragma unroll
        for (int c = 0; c < BN; c += 16) {
          uint32_t v[16];
          ptx::tmem_ld16(t_row + (uint32_t)c, v);
          ptx::tmem_ld_wait();
          ptx::tmem_st16_zero(t_row + (uint32_t)c);              // the slot goes back to the MMA issuer zeroed
          float yv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            yv[j] = fmaf(__uint_as_float(v[j]), p.k.scale[c + j], p.k.bias[c + j]);
          }
          if (has_res) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint4 rv = rall[c / 8 + k];
              const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int tt = 0; tt < 4; ++tt) {
                yv[8 * k + 2 * tt + 0] += __uint_as_float(w[tt] << 16);
                yv[8 * k + 2 * tt + 1] += __uint_as_float(w[tt] & 0xffff0000u);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) yv[j] = fmaxf(yv[j], lo);
          if (HEAD) {
            float* hp = reinterpret_cast<float*>(&hacc);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float s0 = 0.f, s1 = 0.f;      // two independent chains per logit, summed in a fixed order
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                s0 = fmaf(yv[j], p.k.head[k * 16 + j], s0);
                s1 = fmaf(yv[j + 1], p.k.head[k * 16 + j + 1], s1);
              }
              hp[k] = (s0 + s1) + p.k.head[64 + k];
            }
          }
          if (valid && p.out != nullptr) {
            if (!planar_out) {
              // NHWC: the 16 channels of this step are 32 contiguous bytes — one 256-bit store (a full sector)
              uint32_t w[8];
#pragma unroll
              for (int tt = 0; tt < 8; ++tt) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(yv[2 * tt], yv[2 * tt + 1]);
                w[tt] = *reinterpret_cast<uint32_t*>(&h2);
              }
              ptx::st_global_256(p.out + o_off + (size_t)c * 2, w);
            } else {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              uint32_t w[4];
#pragma unroll
              for (int tt = 0; tt < 4; ++tt) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(yv[8 * k + 2 * tt], yv[8 * k + 2 * tt + 1]);
                w[tt] = *reinterpret_cast<uint32_t*>(&h2);
              }
              *reinterpret_cast<uint4*>(p.out + o_off + (size_t)(c / 8 + k) * o_step) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            }
          }
        }
        if (HEAD && valid) reinterpret_cast<float4*>(p.head_out)[pix] = hacc;
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&slot_free[slot]);
      }
      for (int k = 0; k < 2; ++k)      // the two halo rows of the unit
        if (++gs == S) { gs = 0; gph ^= 1u; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(*tmem_holder, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// Two-lane variant for Cout = 16 / 32.
//
// With N = 3*Cout <= 96 the MMAs of one strip form ONE dependent accumulator chain, and a dependent tcgen05.mma
// chain retires one instruction per ~160 cycles whatever its size (measured: 4 MMAs of N <= 48 per 128-pixel row
// took ~690 cycles; the same MMAs into unrelated accumulators ~18 % less, with the other roles then bounding).
// Here a CTA streams a strip of 256 output columns as two 128-column lanes whose MMAs are issued alternately
// (independent chains, the second lane's A operand is the same smem run 128 entries further), each lane has its
// own TMEM region and its own two epilogue groups (4 x 4 epilogue warps in all), and a completed output row is
// signalled per ring slot (row_done[slot]) so the stage ring and the accumulator ring are decoupled.
// ---------------------------------------------------------------------------------------------
// packed fp32 FMA (FFMA2): two independent IEEE fmas per instruction — halves the epilogue's FMA count without
// changing any rounding
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ua, ub, uc, ud;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(uc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ua), "l"(ub), "l"(uc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
  return d;
}

constexpr int kRun2Cols = 256 + 2 * kRowPad;              // 272 entries: x = b0-8 .. b0+263
constexpr int kRun2Bytes = kRun2Cols * 16;                // 4352
constexpr int kStream2Threads = (1 + 1 + 16) * 32;

template <int BN, bool HEAD>
__global__ void __launch_bounds__(kStream2Threads, 1) conv_rowstream2_kernel(const __grid_constant__ StreamParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  constexpr int RS = 256 / BN;                             // ring slots per lane: 2 lanes x RS x BN = 512 TMEM columns
  constexpr uint32_t RM = RS - 1;
  constexpr uint32_t kLane = RS * BN;                      // TMEM columns of one lane
  constexpr uint32_t kTmemCols = 2 * kLane;
  const int w_bytes = p.nslabs * 3 * 2 * 3 * BN * 16;
  uint8_t* s_w = smem;
  float* s_scale = reinterpret_cast<float*>(smem + w_bytes);
  float* s_bias = s_scale + BN;          // (the fused head and, for BN > 16, scale / bias are read from the parameter bank: p.k)
  uint8_t* s_stage = smem + ((w_bytes + (2 * BN + 68) * 4 + 127) & ~127);
  const int S = p.stages;
  const int stage_bytes = p.nslabs * 2 * kRun2Bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + (size_t)S * stage_bytes);
  uint64_t* full = bars;                   // producer -> MMA: input row landed
  uint64_t* empty = bars + S;              // MMA -> producer: the MMAs reading this input row retired
  uint64_t* row_done = bars + 2 * S;       // MMA -> epilogue: the output row in this slot is complete (both lanes)
  uint64_t* slot_free = bars + 2 * S + RS; // epilogue -> MMA: slot drained by both lanes (256 arrivals)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * S + 2 * RS);
  uint64_t* dummy_bar = bars + 2 * S + 2 * RS + 1;        // timing experiment dbg = 6 only

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = threadIdx.x; i < w_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < BN; i += blockDim.x) { s_scale[i] = p.scale[i]; s_bias[i] = p.bias[i]; }
    uint4* st = reinterpret_cast<uint4*>(s_stage);
    for (int i = threadIdx.x; i < S * stage_bytes / 16; i += blockDim.x) st[i] = make_uint4(0, 0, 0, 0);
    sptx::fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < RS; ++i) {
      ptx::mbar_init(&row_done[i], 1);
      ptx::mbar_init(&slot_free[i], 256);
    }
    ptx::mbar_init(dummy_bar, 1u << 19);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_holder, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // every accumulator slot starts (and is handed back by the epilogue) zeroed: see the MMA issuer's interior path
  if (warp >= 2) {
    ptx::tmem_zero_all(*tmem_holder, warp, 4, (int)kTmemCols);
    ptx::tc_fence_before();
  }
  __syncthreads();
  ptx::tc_fence_after();

  if (warp == 0) {
    // ================================ producer ==========================================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage0 = ptx::smem_u32(s_stage);
      uint32_t dst = stage0;
      const size_t run_step = (size_t)p.d.Wrow * 16;
      const size_t row_step = (size_t)p.d.KC * run_step;
      const int nruns = 2 * p.nslabs;
      UnitIter uc;
      uc.init(p);
      while (uc.next()) {
        const int b0 = uc.xb * 256;
        const uint32_t bytes = (uint32_t)min(kRun2Cols, p.d.Wrow - b0) * 16u;
        const uint32_t row_tx = (uint32_t)nruns * bytes;
        const uint8_t* rowp = p.in + p.d.row_off(uc.n, uc.y0 - 1, 0, 0) + (size_t)b0 * 16;
        const int rows = uc.Lu + 2;
        for (int t = 0; t < rows; ++t) {
          ptx::mbar_wait(&empty[stage], phase ^ 1u, p.error_flag, 51);
          ptx::mbar_expect_tx(&full[stage], row_tx);
          const uint8_t* src = rowp;
          uint32_t d = dst;
#pragma unroll 2
          for (int r = 0; r < nruns; ++r) {
            sptx::bulk_g2s(d, src, bytes, &full[stage]);
            d += kRun2Bytes;
            src += run_step;
          }
          rowp += row_step;
          dst += (uint32_t)stage_bytes;
          if (++stage == S) { stage = 0; phase ^= 1u; dst = stage0; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ========================================
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_holder);
    if (ptx::elect_one()) {
      const uint64_t a_desc0 = stream_nosw_desc(ptx::smem_u32(s_stage), kRun2Bytes, 128) + (uint64_t)(kRowPad - 1);
      const uint64_t b_desc0 = stream_nosw_desc(ptx::smem_u32(s_w), 3 * BN * 16, 128);
      const uint32_t stage_units = (uint32_t)stage_bytes >> 4;
      constexpr uint32_t kSlabUnits = (2 * kRun2Bytes) >> 4;       // A: next 16-channel slab of the row
      constexpr uint32_t kLaneUnits = 128;                         // A: the second lane, 128 entries further
      constexpr uint32_t kWSlab = 3 * 2 * 3 * BN, kWShift = 2 * 3 * BN;
      const uint32_t id1 = stream_idesc(BN), id2 = stream_idesc(2 * BN), id3 = stream_idesc(3 * BN);
      const int nslabs = p.nslabs;
      const bool no_mma = (WSI_DBG(p) == 1);
      int stage = 0;
      uint32_t phase = 0, a_off = 0;
      uint32_t jb = 0;
      auto issue = [&](uint32_t d_base, uint64_t a_desc, uint64_t b_desc, uint32_t jrow_a, int a, int b, uint32_t accumulate) {
        int cnt = b - a + 1;
        int slot = (int)(jrow_a & RM);
        int blk = a;
        while (cnt > 0) {
          const int n_here = min(cnt, RS - slot);      // split where the ring wraps
          ptx::umma_bf16(d_base + (uint32_t)(slot * BN), a_desc, b_desc + (uint64_t)(blk * BN), stream_idesc(n_here * BN), accumulate);
          cnt -= n_here;
          blk += n_here;
          slot = 0;
        }
      };
      UnitIter uc;
      uc.init(p);
      while (uc.next()) {
        const int Lu = uc.Lu;
        for (int t = 0; t < Lu + 2; ++t) {
          const bool opens = (t < Lu);
          const uint32_t j = jb + (uint32_t)t;
          if (opens) ptx::mbar_wait(&slot_free[j & RM], ((j / RS) & 1u) ^ 1u, p.error_flag, 52);
          const uint64_t a_row = a_desc0 + (uint64_t)a_off;
          ptx::mbar_wait(&full[stage], phase, p.error_flag, 53);
          ptx::tc_fence_after();
          const uint32_t slot_lo = (j - 2u) & RM;
          if (no_mma) {
          } else if (t >= 2 && opens && slot_lo <= RS - 3) {
            // interior row, contiguous slots: lanes alternate, so consecutive MMAs never share an accumulator
            const uint32_t dA = tmem_base + slot_lo * BN, dB = dA + kLane;
            const uint64_t a_B = a_row + kLaneUnits;
            ptx::umma_bf16(dA, a_row, b_desc0, id3, 1u);       // the newly opened row's slot is zero (epilogue hands it back zeroed)
            ptx::umma_bf16(dB, a_B, b_desc0, id3, 1u);
            ptx::umma_bf16(dA, a_row + 1, b_desc0 + (uint64_t)kWShift, id3, 1u);
            ptx::umma_bf16(dB, a_B + 1, b_desc0 + (uint64_t)kWShift, id3, 1u);
            ptx::umma_bf16(dA, a_row + 2, b_desc0 + (uint64_t)(2 * kWShift), id3, 1u);
            ptx::umma_bf16(dB, a_B + 2, b_desc0 + (uint64_t)(2 * kWShift), id3, 1u);
            uint64_t a_sl = a_row, b_sl = b_desc0;
#pragma unroll 1
            for (int sl = 1; sl < nslabs; ++sl) {
              a_sl += kSlabUnits;
              b_sl += kWSlab;
#pragma unroll
              for (int sft = 0; sft < 3; ++sft) {
                ptx::umma_bf16(dA, a_sl + sft, b_sl + (uint64_t)(sft * kWShift), id3, 1u);
                ptx::umma_bf16(dB, a_sl + kLaneUnits + sft, b_sl + (uint64_t)(sft * kWShift), id3, 1u);
              }
            }
          } else {
            const int i_lo = max(0, t - 2), i_hi = min(Lu - 1, t);
            const int blk_lo = 2 - t + i_lo, blk_hi = 2 - t + i_hi;
            for (int sl = 0; sl < nslabs; ++sl)
              for (int sft = 0; sft < 3; ++sft)
                for (int ln = 0; ln < 2; ++ln) {
                  const uint64_t a_desc = a_row + (uint64_t)(sl * kSlabUnits + ln * kLaneUnits + sft);
                  const uint64_t b_desc = b_desc0 + (uint64_t)(sl * kWSlab + sft * kWShift);
                  const uint32_t d_base = tmem_base + (uint32_t)ln * kLane;
                  if (sl == 0 && sft == 0 && opens) {
                    issue(d_base, a_desc, b_desc, j, 2, 2, 0u);                                    // overwrite the new row
                    if (blk_lo <= 1) issue(d_base, a_desc, b_desc, jb + (uint32_t)i_lo, blk_lo, 1, 1u);
                  } else {
                    issue(d_base, a_desc, b_desc, jb + (uint32_t)i_lo, blk_lo, blk_hi, 1u);
                  }
                }
          }
          ptx::umma_commit(&empty[stage]);                                   // stage back to the producer
          if (t >= 2) ptx::umma_commit(&row_done[(j - 2u) & RM]);            // output row t-2 is complete
          if (WSI_DBG(p) == 6) { ptx::umma_commit(dummy_bar); ptx::umma_commit(dummy_bar); }   // what does a commit cost?
          a_off += stage_units;
          if (++stage == S) { stage = 0; phase ^= 1u; a_off = 0; }
        }
        jb += (uint32_t)Lu;
      }
    }
  } else {
    // ================================ epilogue (2 lanes x 2 row parities x 4 warps) ======
    const uint32_t tmem_base = *tmem_holder;
    const int q = warp & 3;
    const int egrp = (warp - 2) >> 2;
    const int lane_sel = egrp >> 1;
    const uint32_t rowpar = (uint32_t)(egrp & 1);
    const int mrow = q * 32 + lane;
    constexpr bool kRegSB = (BN == 16) && !HEAD;        // folded-BN constants in registers only where the budget allows
    float r_scale[kRegSB ? BN : 1], r_bias[kRegSB ? BN : 1];
    if (kRegSB) {
#pragma unroll
      for (int j = 0; j < BN; ++j) { r_scale[kRegSB ? j : 0] = s_scale[j]; r_bias[kRegSB ? j : 0] = s_bias[j]; }
    }
    const float lo = p.relu ? 0.f : -INFINITY;
    const bool planar_out = (p.out_layout == LAYOUT_PLANAR);
    const bool planar_res = (p.res_layout == LAYOUT_PLANAR);
    const size_t chunk_step = (size_t)p.od.Wrow * 16;
    const size_t plane_row_step = (size_t)p.od.KC * p.od.P * chunk_step;
    const size_t o_step = planar_out ? chunk_step : 16, o_inc = planar_out ? plane_row_step : (size_t)p.OW * (BN * 2);
    const size_t r_step = planar_res ? chunk_step : 16, r_inc = planar_res ? plane_row_step : (size_t)p.OW * (BN * 2);
    uint32_t job = 0;
    UnitIter uc;
    uc.init(p);
    while (uc.next()) {
      const int Lu = uc.Lu;
      const int x = uc.xb * 256 + lane_sel * 128 + mrow;
      const bool valid = x < p.OW;
      size_t pix = ((size_t)uc.n * p.OH + uc.y0) * p.OW + x;
      const size_t prow = p.od.row_off(uc.n, uc.y0, 0, 0) + (size_t)(x + kRowPad) * 16;
      size_t o_off = planar_out ? prow : pix * (size_t)(BN * 2);
      size_t r_off = planar_res ? prow : pix * (size_t)(BN * 2);
      const bool has_res = (p.res != nullptr) && valid;
      for (int i = 0; i < Lu; ++i, ++job, pix += (size_t)p.OW, o_off += o_inc, r_off += r_inc) {
        if ((job & 1u) != rowpar) continue;
        const uint32_t slot = job & RM;
        uint4 rall[BN / 8];
        if (has_res) {
#pragma unroll
          for (int k = 0; k < BN / 8; ++k) rall[k] = __ldg(reinterpret_cast<const uint4*>(p.res + r_off + (size_t)k * r_step));
        }
        ptx::mbar_wait(&row_done[slot], (job / RS) & 1u, p.error_flag, 54);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)lane_sel * kLane + slot * BN;
        float4 hacc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (WSI_DBG(p) != 3)
#pragma unroll
        for (int c = 0; c < BN; c += 16) {
          uint32_t v[16];
          ptx::tmem_ld16(t_row + (uint32_t)c, v);
          ptx::tmem_ld_wait();
          ptx::tmem_st16_zero(t_row + (uint32_t)c);              // the slot goes back to the MMA issuer zeroed
          float yv[16];
          if (HEAD) {
            // d5b + final_conv: BN, ReLU and the 16 -> 4 head as packed FMAs; constants are broadcast smem reads.
            // Lane .x of every pair carries the even channels, .y the odd ones: the same two summation chains per
            // logit as the scalar formulation, so the result is bit-identical.
            // (constants: compile-time offsets into the parameter bank — FFMA with a c[][] operand, no shared-memory
            //  loads; the even / odd summation chains of the earlier packed-FMA formulation are kept, so results are
            //  bit-identical with it)
            float2 y2[8];
#pragma unroll
            for (int jp = 0; jp < 8; ++jp) {
              y2[jp].x = fmaxf(fmaf(__uint_as_float(v[2 * jp]), p.k.scale[2 * jp], p.k.bias[2 * jp]), lo);
              y2[jp].y = fmaxf(fmaf(__uint_as_float(v[2 * jp + 1]), p.k.scale[2 * jp + 1], p.k.bias[2 * jp + 1]), lo);
            }
            float* hp = reinterpret_cast<float*>(&hacc);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float ae = 0.f, ao = 0.f;
#pragma unroll
              for (int jp = 0; jp < 8; ++jp) {
                ae = fmaf(y2[jp].x, p.k.head[k * 16 + 2 * jp], ae);
                ao = fmaf(y2[jp].y, p.k.head[k * 16 + 2 * jp + 1], ao);
              }
              hp[k] = (ae + ao) + p.k.head[64 + k];
            }
            if (valid && p.out != nullptr) {
#pragma unroll
              for (int jp = 0; jp < 8; ++jp) { yv[2 * jp] = y2[jp].x; yv[2 * jp + 1] = y2[jp].y; }
            }
          } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float sc = kRegSB ? r_scale[kRegSB ? c + j : 0] : p.k.scale[c + j];
            const float bi = kRegSB ? r_bias[kRegSB ? c + j : 0] : p.k.bias[c + j];
            yv[j] = fmaf(__uint_as_float(v[j]), sc, bi);
          }
          }
          if (!HEAD && has_res) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint4 rv = rall[c / 8 + k];
              const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int tt = 0; tt < 4; ++tt) {
                yv[8 * k + 2 * tt + 0] += __uint_as_float(w[tt] << 16);
                yv[8 * k + 2 * tt + 1] += __uint_as_float(w[tt] & 0xffff0000u);
              }
            }
          }
          if (!HEAD) {
#pragma unroll
            for (int j = 0; j < 16; ++j) yv[j] = fmaxf(yv[j], lo);
          }
          if (valid && p.out != nullptr) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              uint32_t w[4];
#pragma unroll
              for (int tt = 0; tt < 4; ++tt) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(yv[8 * k + 2 * tt], yv[8 * k + 2 * tt + 1]);
                w[tt] = *reinterpret_cast<uint32_t*>(&h2);
              }
              *reinterpret_cast<uint4*>(p.out + o_off + (size_t)(c / 8 + k) * o_step) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
        if (HEAD && valid) reinterpret_cast<float4*>(p.head_out)[pix] = hacc;
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&slot_free[slot]);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(*tmem_holder, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool RowStreamOp::eligible(const std::vector<ConvInputPart>& parts, const ConvSpec& spec) {
  if (spec.ksize != 3 || spec.stride != 1 || spec.pad != 1) return false;
  if (spec.cout != 16 && spec.cout != 32 && spec.cout != 64) return false;
  if (spec.head && spec.cout != 16) return false;
  if (parts.size() != 1 || parts[0].up2) return false;
  const int C = parts[0].t.C;
  return C % 16 == 0 && C > 0 && C <= 64;
}

void RowStreamOp::build(const ConvInputPart& part, const ConvSpec& spec, const float* w_oihw, const float* scale, const float* bias,
                        const void* residual, int res_layout, void* out, int out_layout, const float* head_w, const float* head_b,
                        float* head_out, int* error_flag, int num_sms) {
  StreamParams& p = p_;
  p = StreamParams{};
  const TensorView& t = part.t;
  const int N = t.N, OH = t.H, OW = t.W, C = t.C, BN = spec.cout;
  p.d = PlanarDims::make(t.H, t.W, C, LAYOUT_PLANAR);
  relayout_ = false;
  if (t.layout == LAYOUT_NHWC) {
    stage_in_.alloc(p.d.bytes(N));
    CUDA_CHECK(cudaMemset(stage_in_.p, 0, stage_in_.bytes));
    relayout_ = true;
    relayout_src_ = t.ptr;
    p.in = stage_in_.as<uint8_t>();
  } else {
    WSI_REQUIRE(t.layout == LAYOUT_PLANAR, WSI_ERR_INVALID, "row-stream conv: operand layout %d", t.layout);
    p.in = static_cast<const uint8_t*>(t.ptr);
  }
  WSI_REQUIRE(out_layout == LAYOUT_NHWC || out_layout == LAYOUT_PLANAR, WSI_ERR_INVALID, "row-stream conv: bad output layout");
  WSI_REQUIRE(res_layout == LAYOUT_NHWC || res_layout == LAYOUT_PLANAR, WSI_ERR_INVALID, "row-stream conv: bad residual layout");
  p.nslabs = C / 16;
  p.N = N; p.OH = OH; p.OW = OW; p.Cout = BN;
  // Cout 16 / 32: two 128-column lanes per CTA (conv_rowstream2_kernel)
  // Cout 16 / 32, when the stage ring fits in half the shared memory (d5b, d4b): the one-lane kernel at TWO CTAs per SM
  // (256 TMEM columns each) — two independent pipelines per SM overlap one CTA's epilogue with the other's MMAs a little
  // better than two lanes inside one CTA (same-box A/B: conv stage -0.8 %).  Otherwise two 128-column lanes per CTA
  // (conv_rowstream2_kernel).
  const int ring2 = 256 / BN;
  const int stages2 = std::min(24, (110 * 1024 - (128 + (((C / 16) * 3 * 2 * 3 * BN * 16 + (2 * BN + 68) * 4 + 127) & ~127) + 1024)) / ((C / 16) * kStreamStageBytes));
  const bool ctas2 = BN <= 32 && stages2 + 2 >= ring2 && stages2 >= 8 && getenv("WSI_STREAM_CTAS1") == nullptr;
  p.lanes = (BN <= 32 && getenv("WSI_STREAM_LANES1") == nullptr && !ctas2) ? 2 : 1;
  p.tiles_x = (int)ceil_div(OW, 128 * p.lanes);
  const long long total = (long long)N * p.tiles_x * OH;      // (strip, output row) pairs, split evenly over the CTAs
  WSI_REQUIRE(total < (1LL << 31), WSI_ERR_UNSUPPORTED, "row-stream conv: too many rows");
  p.total_rows = (int)total;
  p.relu = spec.relu ? 1 : 0;
  p.out = static_cast<uint8_t*>(out);
  p.out_layout = out_layout;
  p.od = PlanarDims::make(OH, OW, BN, LAYOUT_PLANAR);
  p.res = static_cast<const uint8_t*>(residual);
  p.res_layout = res_layout;
  p.error_flag = error_flag;
#ifdef WSI_DEBUG_SWITCHES
  if (const char* e = getenv("WSI_STREAM_DBG")) p.dbg = atoi(e);
#endif
  // stacked weights: [slab][s][2 chunks][3*BN][8], N order = vertical tap r = 2 | 1 | 0
  std::vector<uint16_t> wp((size_t)p.nslabs * 3 * 2 * 3 * BN * 8);
  for (int sl = 0; sl < p.nslabs; ++sl)
    for (int s = 0; s < 3; ++s)
      for (int ch = 0; ch < 2; ++ch)
        for (int blk = 0; blk < 3; ++blk)
          for (int n = 0; n < BN; ++n)
            for (int e = 0; e < 8; ++e) {
              const int r = 2 - blk, ci = sl * 16 + ch * 8 + e;
              const float v = w_oihw[(((size_t)n * C + ci) * 3 + r) * 3 + s];
              wp[(((((size_t)sl * 3 + s) * 2 + ch) * 3 + blk) * BN + n) * 8 + e] = f32_to_bf16_bits(v);
            }
  upload(w_, wp);
  std::vector<float> sc(BN, 1.f), bi(BN, 0.f);
  if (scale) sc.assign(scale, scale + BN);
  if (bias) bi.assign(bias, bias + BN);
  upload(scale_, sc);
  upload(bias_, bi);
  p.w = w_.as<bf16>(); p.scale = scale_.as<float>(); p.bias = bias_.as<float>();
  for (int j = 0; j < BN; ++j) { p.k.scale[j] = sc[j]; p.k.bias[j] = bi[j]; }
  flops_ = 2.0 * N * OH * OW * (double)BN * C * 9;
  if (spec.head) {
    WSI_REQUIRE(head_w && head_b && head_out, WSI_ERR_INVALID, "fused head needs weights and an output");
    std::vector<float> hw(head_w, head_w + 64), hb(head_b, head_b + 4);
    upload(headw_, hw);
    upload(headb_, hb);
    p.head_w = headw_.as<float>(); p.head_b = headb_.as<float>(); p.head_out = head_out;
    for (int j = 0; j < 64; ++j) p.k.head[j] = hw[j];
    for (int j = 0; j < 4; ++j) p.k.head[64 + j] = hb[j];
    flops_ += 2.0 * N * OH * OW * 16 * 4;
  }
  const int w_bytes = p.nslabs * 3 * 2 * 3 * BN * 16;
  const int fixed = 128 + ((w_bytes + (2 * BN + 68) * 4 + 127) & ~127) + 1024;
  const int stage_bytes = (p.lanes == 2) ? p.nslabs * 2 * kRun2Bytes : p.nslabs * kStreamStageBytes;
  p.stages = std::min(24, ((ctas2 ? 110 : 226) * 1024 - fixed) / stage_bytes);
  if (p.lanes == 2) {
    WSI_REQUIRE(p.stages >= 3, WSI_ERR_UNSUPPORTED, "row-stream conv: not enough shared memory");
    p.ring = 256 / BN;
    smem_ = fixed + p.stages * stage_bytes;
    grid_ = (int)std::min<long long>(total, num_sms);
    CUDA_CHECK(cudaStreamSynchronize(0));
    return;
  }
  // accumulator ring: 16 output rows in flight when TMEM (512 columns) and the stage ring allow it — the chain
  // commit -> epilogue drain -> slot_free -> MMA of a later row has more slack to hide in
  p.ring = (BN <= 32 && p.stages >= 18) ? 16 : 8;
  if (const char* e = getenv("WSI_STREAM_RING")) { const int r = atoi(e); if ((r == 8 || r == 16) && r * BN <= 512) p.ring = r; }
  if (ctas2) p.ring = std::min(p.ring, 256 / BN);                       // two CTAs share the 512 TMEM columns
  WSI_REQUIRE(p.stages + 2 >= p.ring, WSI_ERR_UNSUPPORTED, "row-stream conv: not enough shared memory");   // see the epilogue wait
  smem_ = fixed + p.stages * stage_bytes;
  grid_ = (int)std::min<long long>(total, ctas2 ? 2 * num_sms : num_sms);
  CUDA_CHECK(cudaStreamSynchronize(0));
}

template <int BN, bool HEAD, int RS>
static void launch_stream_rs(const StreamParams& p, int grid, int smem, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(conv_rowstream_kernel<BN, HEAD, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured = true;
  }
  conv_rowstream_kernel<BN, HEAD, RS><<<grid, kStreamThreads, smem, s>>>(p);
  CUDA_CHECK(cudaGetLastError());
}

template <int BN, bool HEAD>
static void launch_stream(const StreamParams& p, int grid, int smem, cudaStream_t s) {
  if constexpr (BN <= 32) {
    if (p.ring == 16) { launch_stream_rs<BN, HEAD, 16>(p, grid, smem, s); return; }
  }
  launch_stream_rs<BN, HEAD, 8>(p, grid, smem, s);
}

template <int BN, bool HEAD>
static void launch_stream2(const StreamParams& p, int grid, int smem, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(conv_rowstream2_kernel<BN, HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured = true;
  }
  conv_rowstream2_kernel<BN, HEAD><<<grid, kStream2Threads, smem, s>>>(p);
  CUDA_CHECK(cudaGetLastError());
}

void RowStreamOp::launch(cudaStream_t stream, LaunchCounter* lc) const {
  if (relayout_) launch_relayout_planar(relayout_src_, stage_in_.p, p_.N, p_.OH, p_.OW, p_.nslabs * 16, LAYOUT_PLANAR, stream, lc);
  const bool head = p_.head_out != nullptr;
  if (p_.lanes == 2) {
    if (p_.Cout == 16) { if (head) launch_stream2<16, true>(p_, grid_, smem_, stream); else launch_stream2<16, false>(p_, grid_, smem_, stream); }
    else launch_stream2<32, false>(p_, grid_, smem_, stream);
    if (lc) lc->n++;
    return;
  }
  if (p_.Cout == 16) { if (head) launch_stream<16, true>(p_, grid_, smem_, stream); else launch_stream<16, false>(p_, grid_, smem_, stream); }
  else if (p_.Cout == 32) launch_stream<32, false>(p_, grid_, smem_, stream);
  else launch_stream<64, false>(p_, grid_, smem_, stream);
  if (lc) lc->n++;
}

}  // namespace wsi
