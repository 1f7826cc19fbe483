// conv_upstream.cu — see conv_upstream.cuh.
#include "conv_upstream.cuh"

#include <algorithm>
#include <cstdlib>

namespace wsi {

namespace uptx {
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// no-swizzle K-major operand: 8-row core matrices of 128 contiguous bytes, SBO between 8-row groups, LBO between
// the two 8-channel chunks of a K = 16 step
__device__ __forceinline__ uint64_t nosw_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor: D fp32, A/B bf16 K-major, M = 128, N given at run time
__device__ __forceinline__ uint32_t idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}
}  // namespace uptx

// Work split: the (strip, source row) space — strips of 128 source columns, strip-major — is cut into gridDim.x
// equal contiguous ranges; a unit is the part of one strip inside the CTA's range: source rows [a0, a0 + La),
// i.e. output rows [2*a0, 2*(a0 + La)), computed by the steps a = a0-1 .. a0+La.
struct UpUnit {
  int g, g_end, h, tiles_x;
  int n, xb, a0, La;
  __device__ __forceinline__ void init(const UpParams& p) {
    g = (int)((long long)p.total_rows * blockIdx.x / gridDim.x);
    g_end = (int)((long long)p.total_rows * (blockIdx.x + 1) / gridDim.x);
    h = p.h;
    tiles_x = p.tiles_x;
  }
  __device__ __forceinline__ bool next() {
    if (g >= g_end) return false;
    const int strip = g / h;
    a0 = g - strip * h;
    La = min(h - a0, g_end - g);
    n = strip / tiles_x;
    xb = strip - n * tiles_x;
    g += La;
    return true;
  }
};

template <int BN, int RS>
__global__ void __launch_bounds__(kUpThreads, 1) conv_upstream_kernel(const __grid_constant__ UpParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  constexpr uint32_t RM = RS - 1;
  constexpr uint32_t kTmemCols = 2 * RS * BN;        // two column parities
  static_assert(kTmemCols == 512, "TMEM ring: 2 parities x RS rows x BN columns");
  const int w_al = (p.w_bytes + 127) & ~127;
  uint8_t* s_w = smem;
  float* s_scale = reinterpret_cast<float*>(smem + w_al);
  float* s_bias = s_scale + BN;
  uint8_t* s_stage = smem + w_al + 2 * BN * 4;       // BN * 8 is a multiple of 128
  const int S = p.stages;
  const int stage_bytes = p.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + (size_t)S * stage_bytes);
  uint64_t* full = bars;                   // producer -> MMA: input row item landed
  uint64_t* empty = bars + S;              // MMA -> producer: the MMAs reading this item retired
  // per ring slot, so that the epilogue group that owns a slot (job & 3; RS is a multiple of 4) sees EVERY phase
  // of its barriers — a parity wait on a barrier whose earlier phases the waiter skipped passes prematurely
  uint64_t* row_done = bars + 2 * S;       // MMA -> epilogue: the output row in this slot is complete
  uint64_t* slot_free = bars + 2 * S + RS; // epilogue -> MMA: accumulator slot drained
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * S + 2 * RS);
  const bool has_skip = (p.skip != nullptr);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.wts);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = threadIdx.x; i < p.w_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < BN; i += blockDim.x) { s_scale[i] = p.scale[i]; s_bias[i] = p.bias[i]; }
    uint4* st = reinterpret_cast<uint4*>(s_stage);
    for (int i = threadIdx.x; i < S * stage_bytes / 16; i += blockDim.x) st[i] = make_uint4(0, 0, 0, 0);
    uptx::fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < RS; ++i) {
      ptx::mbar_init(&row_done[i], 1);
      ptx::mbar_init(&slot_free[i], 128);
    }
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
      const size_t run_u = (size_t)p.du.Wrow * 16, run_s = (size_t)p.ds.Wrow * 16;
      const int nruns_u = 2 * p.nslabs_u, nruns_s = 4 * p.nslabs_s;
      auto load_item = [&](const uint8_t* src, size_t run_step, int nruns, uint32_t bytes) {
        ptx::mbar_wait(&empty[stage], phase ^ 1u, p.error_flag, 41);
        ptx::mbar_expect_tx(&full[stage], (uint32_t)nruns * bytes);
        uint32_t d = dst;
#pragma unroll 2
        for (int r = 0; r < nruns; ++r) {
          uptx::bulk_g2s(d, src, bytes, &full[stage]);
          d += kRowRunBytes;
          src += run_step;
        }
        dst += (uint32_t)stage_bytes;
        if (++stage == S) { stage = 0; phase ^= 1u; dst = stage0; }
      };
      UpUnit uc;
      uc.init(p);
      while (uc.next()) {
        const int b0 = uc.xb * 128;
        const uint32_t bytes_u = (uint32_t)min(kRowHaloCols, p.du.Wrow - b0) * 16u;
        const uint32_t bytes_s = has_skip ? (uint32_t)min(kRowHaloCols, p.ds.Wrow - b0) * 16u : 0u;
        const int a_first = uc.a0 - 1, a_last = uc.a0 + uc.La;
        for (int a = a_first; a <= a_last; ++a) {
          load_item(p.u + p.du.row_off(uc.n, a, 0, 0) + (size_t)b0 * 16, run_u, nruns_u, bytes_u);
          if (has_skip) {
            // the first step only needs skip row 2a+1 (row 2a feeds rows below the unit), the last one only 2a
            if (a != a_first) load_item(p.skip + p.ds.row_off(uc.n, 2 * a, 0, 0) + (size_t)b0 * 16, run_s, nruns_s, bytes_s);
            if (a != a_last) load_item(p.skip + p.ds.row_off(uc.n, 2 * a + 1, 0, 0) + (size_t)b0 * 16, run_s, nruns_s, bytes_s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ========================================
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_holder);
    if (ptx::elect_one()) {
      // A descriptors at entry kRowPad of run 0 of stage 0 (u: chunks are consecutive runs; skip: chunk = 2 plane runs)
      const uint64_t au0 = uptx::nosw_desc(ptx::smem_u32(s_stage), kRowRunBytes, 128) + (uint64_t)kRowPad;
      const uint64_t as0 = uptx::nosw_desc(ptx::smem_u32(s_stage), 2 * kRowRunBytes, 128) + (uint64_t)kRowPad;
      const uint64_t bu0 = uptx::nosw_desc(ptx::smem_u32(s_w), 4 * BN * 16, 128);
      const uint64_t bs0 = uptx::nosw_desc(ptx::smem_u32(s_w) + (uint32_t)p.wskip_off, 3 * BN * 16, 128);
      constexpr uint32_t kRunUnits = kRowRunBytes >> 4;
      constexpr uint32_t kBuBlock = 2 * 4 * BN, kBsBlock = 2 * 3 * BN;      // 16-byte units per weight block
      constexpr uint32_t kPar = RS * BN;                                    // TMEM columns of one parity region
      const uint32_t stage_units = (uint32_t)stage_bytes >> 4;
      const uint32_t id2 = uptx::idesc(2 * BN), id3 = uptx::idesc(3 * BN), id4 = uptx::idesc(4 * BN);
      const int nslabs_u = p.nslabs_u, nslabs_s = p.nslabs_s;
      const bool no_mma = (WSI_DBG(p) == 1);
      int stage = 0;
      uint32_t phase = 0, a_off = 0;
      uint32_t jb = 0;        // output-row jobs of the units before this one
      // weight blocks [blk_lo, blk_hi] (BN columns each) into the ring slots of jobs j_lo, j_lo+1, ...
      auto issue = [&](uint32_t pbase, uint64_t a_desc, uint64_t b_desc, uint32_t j_lo, int blk_lo, int blk_hi, uint32_t accumulate) {
        int cnt = blk_hi - blk_lo + 1;
        int slot = (int)(j_lo & RM);
        int blk = blk_lo;
        while (cnt > 0) {
          const int n_here = min(cnt, RS - slot);      // split where the ring wraps
          ptx::umma_bf16(pbase + (uint32_t)(slot * BN), a_desc, b_desc + (uint64_t)(blk * BN), uptx::idesc(n_here * BN), accumulate);
          cnt -= n_here;
          blk += n_here;
          slot = 0;
        }
      };
      auto advance = [&]() {
        ptx::umma_commit(&empty[stage]);
        a_off += stage_units;
        if (++stage == S) { stage = 0; phase ^= 1u; a_off = 0; }
      };
      UpUnit uc;
      uc.init(p);
      while (uc.next()) {
        const int lo = 2 * uc.a0, hi = 2 * (uc.a0 + uc.La);
        const int a_first = uc.a0 - 1, a_last = uc.a0 + uc.La;
        for (int a = a_first; a <= a_last; ++a) {
          // ---------------- u row a -> output rows 2a-1 .. 2a+2 (blocks 0..3), rows 2a+1, 2a+2 open here ----------
          {
            const int r0 = 2 * a - 1;
            const int b_lo = max(0, lo - r0), b_hi = min(3, hi - 1 - r0);
            const int nb_lo = max(b_lo, 2);
            const uint32_t j0 = jb + (uint32_t)(r0 + b_lo - lo);          // job of block b_lo
            for (int b = nb_lo; b <= b_hi; ++b) {
              const uint32_t j = j0 + (uint32_t)(b - b_lo);
              ptx::mbar_wait(&slot_free[j & RM], ((j / RS) & 1u) ^ 1u, p.error_flag, 42);
            }
            ptx::mbar_wait(&full[stage], phase, p.error_flag, 43);
            ptx::tc_fence_after();
            const uint64_t a_base = au0 + (uint64_t)a_off;
            const uint32_t slot0 = j0 & RM;
            if (no_mma) {
            } else if (b_lo == 0 && b_hi == 3 && slot0 + 3 <= RM) {
              // interior step, no ring wrap: per (slab, tap, parity) one MMA of N = 4*BN.  The slots of the two newly
              // opened rows were zeroed by the epilogue when it drained them (or at kernel start): they accumulate like
              // the other two, no separate overwriting MMAs
              const uint32_t d0 = tmem_base + slot0 * BN;
#pragma unroll 1
              for (int sl = 0; sl < nslabs_u; ++sl) {
                const uint64_t a_sl = a_base + (uint64_t)(sl * 2 * kRunUnits);
                const uint64_t b_sl = bu0 + (uint64_t)(sl * 4 * kBuBlock);
                ptx::umma_bf16(d0, a_sl - 1, b_sl, id4, 1u);                                                               // p0, tap 0
                ptx::umma_bf16(d0 + kPar, a_sl, b_sl + (uint64_t)(2 * kBuBlock), id4, 1u);                                 // p1, tap 0
                ptx::umma_bf16(d0, a_sl, b_sl + (uint64_t)kBuBlock, id4, 1u);                                              // p0, tap 1
                ptx::umma_bf16(d0 + kPar, a_sl + 1, b_sl + (uint64_t)(3 * kBuBlock), id4, 1u);                             // p1, tap 1
              }
            } else {
              for (int sl = 0; sl < nslabs_u; ++sl)
                for (int tau = 0; tau < 2; ++tau)
                  for (int par = 0; par < 2; ++par) {
                    const uint64_t a_desc = a_base + (uint64_t)(sl * 2 * kRunUnits) + (uint64_t)(int64_t)(par + tau - 1);
                    const uint64_t b_desc = bu0 + (uint64_t)(((sl * 2 + par) * 2 + tau) * kBuBlock);
                    const uint32_t pbase = tmem_base + (uint32_t)par * kPar;
                    if (sl == 0 && tau == 0) {
                      if (nb_lo <= b_hi) issue(pbase, a_desc, b_desc, j0 + (uint32_t)(nb_lo - b_lo), nb_lo, b_hi, 0u);
                      if (b_lo <= 1) issue(pbase, a_desc, b_desc, j0, b_lo, min(b_hi, 1), 1u);
                    } else {
                      issue(pbase, a_desc, b_desc, j0, b_lo, b_hi, 1u);
                    }
                  }
            }
            advance();
          }
          // ---------------- skip rows y = 2a, 2a+1 -> output rows y-1 .. y+1 (blocks 0..2) ------------------------
          if (has_skip) {
            for (int yy = 0; yy < 2; ++yy) {
              if ((yy == 0 && a == a_first) || (yy == 1 && a == a_last)) continue;
              const int r0 = 2 * a + yy - 1;
              const int b_lo = max(0, lo - r0), b_hi = min(2, hi - 1 - r0);
              const uint32_t j0 = jb + (uint32_t)(r0 + b_lo - lo);
              ptx::mbar_wait(&full[stage], phase, p.error_flag, 44);
              ptx::tc_fence_after();
              const uint64_t a_base = as0 + (uint64_t)a_off;
              const uint32_t slot0 = j0 & RM;
              if (no_mma) {
              } else if (b_lo == 0 && b_hi == 2 && slot0 + 2 <= RM) {
                const uint32_t d0 = tmem_base + slot0 * BN;
#pragma unroll 1
                for (int sl = 0; sl < nslabs_s; ++sl) {
                  // chunk = 2 plane runs; plane 1 = +kRunUnits.  (parity, s) -> (plane, half-index offset):
                  // p0: s0 (1,-1) s1 (0,0) s2 (1,0);  p1: s0 (0,0) s1 (1,0) s2 (0,+1)
                  const uint64_t a_sl = a_base + (uint64_t)(sl * 4 * kRunUnits);
                  const uint64_t b_sl = bs0 + (uint64_t)(sl * 3 * kBsBlock);
                  ptx::umma_bf16(d0, a_sl + kRunUnits - 1, b_sl, id3, 1u);
                  ptx::umma_bf16(d0 + kPar, a_sl, b_sl, id3, 1u);
                  ptx::umma_bf16(d0, a_sl, b_sl + (uint64_t)kBsBlock, id3, 1u);
                  ptx::umma_bf16(d0 + kPar, a_sl + kRunUnits, b_sl + (uint64_t)kBsBlock, id3, 1u);
                  ptx::umma_bf16(d0, a_sl + kRunUnits, b_sl + (uint64_t)(2 * kBsBlock), id3, 1u);
                  ptx::umma_bf16(d0 + kPar, a_sl + 1, b_sl + (uint64_t)(2 * kBsBlock), id3, 1u);
                }
              } else if (b_lo <= b_hi) {
                for (int sl = 0; sl < nslabs_s; ++sl)
                  for (int s = 0; s < 3; ++s)
                    for (int par = 0; par < 2; ++par) {
                      const int q = par + s - 1;                     // source column 2b + q
                      const int plane = q & 1, delta = (q >= 0) ? q / 2 : -1;
                      const uint64_t a_desc = a_base + (uint64_t)((sl * 4 + plane) * kRunUnits) + (uint64_t)(int64_t)delta;
                      const uint64_t b_desc = bs0 + (uint64_t)((sl * 3 + s) * kBsBlock);
                      issue(tmem_base + (uint32_t)par * kPar, a_desc, b_desc, j0, b_lo, b_hi, 1u);
                    }
              }
              advance();
            }
          }
          // rows 2a-1 and 2a are complete when everything issued so far has retired
#pragma unroll 1
          for (int k = 0; k < 2; ++k) {
            const int oy = 2 * a - 1 + k;
            if (oy >= lo && oy < hi) ptx::umma_commit(&row_done[(jb + (uint32_t)(oy - lo)) & RM]);
          }
        }
        jb += (uint32_t)(2 * uc.La);
      }
    }
  } else {
    // ================================ epilogue (4 x 4 warps) ============================
    const uint32_t tmem_base = *tmem_holder;
    const int q = warp & 3;
    const int egrp = (warp - 2) >> 2;
    const int mrow = q * 32 + lane;
    constexpr uint32_t kPar = RS * BN;
    const float lo_clamp = p.relu ? 0.f : -INFINITY;
    const size_t chunk_step = (size_t)p.od.Wrow * 16;
    const size_t row_step = (size_t)p.od.KC * chunk_step;
    uint32_t job = 0;
    UpUnit uc;
    uc.init(p);
    while (uc.next()) {
      const int lo = 2 * uc.a0, hi = 2 * (uc.a0 + uc.La);
      const int b = uc.xb * 128 + mrow;                  // source column: output pixels 2b, 2b+1
      const bool valid = b < p.w;
      uint8_t* out_base = p.out + p.od.row_off(uc.n, 0, 0, 0) + (size_t)(2 * b + kRowPad) * 16;
      for (int a = uc.a0 - 1; a <= uc.a0 + uc.La; ++a) {
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
          const int oy = 2 * a - 1 + k;
          if (oy < lo || oy >= hi) continue;
          const uint32_t j = job++;
          if ((j & 3u) != (uint32_t)egrp) continue;
          const uint32_t slot = j & RM;
          ptx::mbar_wait(&row_done[slot], (j / RS) & 1u, p.error_flag, 45);
          ptx::tc_fence_after();
          const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + slot * BN;
          uint8_t* orow = out_base + (size_t)oy * row_step;
#pragma unroll
          for (int c = 0; c < BN; c += 16) {
            uint32_t v0[16], v1[16];
            ptx::tmem_ld16(t_row + (uint32_t)c, v0);
            ptx::tmem_ld16(t_row + kPar + (uint32_t)c, v1);
            ptx::tmem_ld_wait();
            ptx::tmem_st16_zero(t_row + (uint32_t)c);            // the slot goes back to the MMA issuer zeroed
            ptx::tmem_st16_zero(t_row + kPar + (uint32_t)c);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              uint32_t w0[4], w1[4];
#pragma unroll
              for (int tt = 0; tt < 4; ++tt) {
                const int jx = 8 * kk + 2 * tt;
                const float sa = p.k.scale[c + jx], sb = p.k.scale[c + jx + 1], ba = p.k.bias[c + jx], bb = p.k.bias[c + jx + 1];
                __nv_bfloat162 e = __floats2bfloat162_rn(fmaxf(fmaf(__uint_as_float(v0[jx]), sa, ba), lo_clamp),
                                                         fmaxf(fmaf(__uint_as_float(v0[jx + 1]), sb, bb), lo_clamp));
                __nv_bfloat162 o = __floats2bfloat162_rn(fmaxf(fmaf(__uint_as_float(v1[jx]), sa, ba), lo_clamp),
                                                         fmaxf(fmaf(__uint_as_float(v1[jx + 1]), sb, bb), lo_clamp));
                w0[tt] = *reinterpret_cast<uint32_t*>(&e);
                w1[tt] = *reinterpret_cast<uint32_t*>(&o);
              }
              if (valid) {
                // pixels 2b and 2b + 1 of this channel chunk: 32 contiguous bytes, one 256-bit store
                const uint32_t w[8] = {w0[0], w0[1], w0[2], w0[3], w1[0], w1[1], w1[2], w1[3]};
                ptx::st_global_256(orow + (size_t)(c / 8 + kk) * chunk_step, w);
              }
            }
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          ptx::mbar_arrive(&slot_free[slot]);
        }
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
static int up_smem_fixed(int w_bytes, int BN) { return 128 + ((w_bytes + 127) & ~127) + 2 * BN * 4 + 1024; }

bool UpStreamOp::eligible(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const void* residual, int out_layout) {
  if (spec.ksize != 3 || spec.stride != 1 || spec.pad != 1) return false;
  if (spec.cout != 16 && spec.cout != 32) return false;
  if (spec.head || residual != nullptr || out_layout != LAYOUT_PLANAR) return false;
  if (parts.empty() || parts.size() > 2 || !parts[0].up2) return false;
  const int Cu = parts[0].t.C;
  if (Cu % 16 != 0 || Cu <= 0 || Cu > 64) return false;
  int Cs = 0;
  if (parts.size() == 2) {
    if (parts[1].up2) return false;
    Cs = parts[1].t.C;
    if (Cs % 16 != 0 || Cs <= 0 || Cs > 64 || parts[1].t.W % 2 != 0) return false;
  }
  const int BN = spec.cout;
  const int w_bytes = (Cu / 16) * 4 * 128 * BN + (Cs / 16) * 3 * 96 * BN;
  const int stage_bytes = std::max(2 * (Cu / 16), 4 * (Cs / 16)) * kRowRunBytes;
  return up_smem_fixed(w_bytes, BN) + 3 * stage_bytes <= 226 * 1024;
}

void UpStreamOp::build(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const float* w_oihw, const float* scale,
                       const float* bias, void* out, int* error_flag, int num_sms) {
  WSI_REQUIRE(eligible(parts, spec, nullptr, LAYOUT_PLANAR), WSI_ERR_UNSUPPORTED, "conv is not eligible for the x2 row-stream kernel");
  UpParams& p = p_;
  p = UpParams{};
  relayouts_.clear();
  const int N = parts[0].t.N, h = parts[0].t.H, w = parts[0].t.W, BN = spec.cout;
  const int Cu = parts[0].t.C, Cs = parts.size() == 2 ? parts[1].t.C : 0, cin = Cu + Cs;
  for (size_t i = 0; i < parts.size(); ++i) {
    const auto& q = parts[i];
    const int want = (i == 0) ? LAYOUT_PLANAR : LAYOUT_PLANAR_PARITY;
    const PlanarDims d = PlanarDims::make(q.t.H, q.t.W, q.t.C, want);
    const uint8_t* base;
    if (q.t.layout == LAYOUT_NHWC) {
      // operand produced by an NHWC kernel: converted before every launch into a private planar buffer
      stage_in_[i].alloc(d.bytes(N));
      CUDA_CHECK(cudaMemset(stage_in_[i].p, 0, stage_in_[i].bytes));
      relayouts_.push_back(Relayout{q.t.ptr, stage_in_[i].p, N, q.t.H, q.t.W, q.t.C, want});
      base = stage_in_[i].as<uint8_t>();
    } else {
      WSI_REQUIRE(q.t.layout == want, WSI_ERR_INVALID, "x2 row-stream conv: operand %zu has layout %d, needs %d", i, q.t.layout, want);
      base = static_cast<const uint8_t*>(q.t.ptr);
    }
    if (i == 0) { p.u = base; p.du = d; p.nslabs_u = Cu / 16; }
    else {
      WSI_REQUIRE(q.t.H == 2 * h && q.t.W == 2 * w && q.t.N == N, WSI_ERR_INVALID, "x2 row-stream conv: skip shape mismatch");
      p.skip = base; p.ds = d; p.nslabs_s = Cs / 16;
    }
  }
  p.N = N; p.h = h; p.w = w; p.OH = 2 * h; p.OW = 2 * w; p.Cout = BN;
  p.tiles_x = (int)ceil_div(w, 128);
  const long long total = (long long)N * p.tiles_x * h;
  WSI_REQUIRE(total < (1LL << 30), WSI_ERR_UNSUPPORTED, "x2 row-stream conv: too many rows");
  p.total_rows = (int)total;
  p.relu = spec.relu ? 1 : 0;
  p.out = static_cast<uint8_t*>(out);
  p.od = PlanarDims::make(p.OH, p.OW, BN, LAYOUT_PLANAR);
  p.error_flag = error_flag;
#ifdef WSI_DEBUG_SWITCHES
  if (const char* e = getenv("WSI_UP_DBG")) p.dbg = atoi(e);
#endif

  // weights (bf16, fp32 sums rounded once).  Input channel order = torch.cat([up(u), skip], 1).
  //   u blocks   [slab][parity p][tap tau][2 chunks][4*BN rows][8]: row blk*BN + n, blk <-> output row 2a-1+blk with
  //              vertical taps {2}, {1,2}, {0,1}, {0}; horizontal taps: p=0: tau0 {0}, tau1 {1,2}; p=1: tau0 {0,1}, tau1 {2}
  //   skip blocks [slab][s][2 chunks][3*BN rows][8]: blk <-> vertical tap r = 2 - blk
  const size_t u_elems = (size_t)p.nslabs_u * 4 * 2 * 4 * BN * 8, s_elems = (size_t)p.nslabs_s * 3 * 2 * 3 * BN * 8;
  std::vector<uint16_t> wp(u_elems + s_elems);
  static const int vset[4][2] = {{2, 2}, {1, 2}, {0, 1}, {0, 0}};              // [blk] -> r range (inclusive)
  static const int hset[2][2][2] = {{{0, 0}, {1, 2}}, {{0, 1}, {2, 2}}};       // [p][tau] -> s range (inclusive)
  for (int sl = 0; sl < p.nslabs_u; ++sl)
    for (int par = 0; par < 2; ++par)
      for (int tau = 0; tau < 2; ++tau)
        for (int ch = 0; ch < 2; ++ch)
          for (int blk = 0; blk < 4; ++blk)
            for (int n = 0; n < BN; ++n)
              for (int e = 0; e < 8; ++e) {
                const int ci = sl * 16 + ch * 8 + e;
                float v = 0.f;
                for (int r = vset[blk][0]; r <= vset[blk][1]; ++r)
                  for (int s = hset[par][tau][0]; s <= hset[par][tau][1]; ++s) v += w_oihw[(((size_t)n * cin + ci) * 3 + r) * 3 + s];
                wp[((((((size_t)sl * 2 + par) * 2 + tau) * 2 + ch) * 4 + blk) * BN + n) * 8 + e] = f32_to_bf16_bits(v);
              }
  for (int sl = 0; sl < p.nslabs_s; ++sl)
    for (int s = 0; s < 3; ++s)
      for (int ch = 0; ch < 2; ++ch)
        for (int blk = 0; blk < 3; ++blk)
          for (int n = 0; n < BN; ++n)
            for (int e = 0; e < 8; ++e) {
              const int r = 2 - blk, ci = Cu + sl * 16 + ch * 8 + e;
              wp[u_elems + (((((size_t)sl * 3 + s) * 2 + ch) * 3 + blk) * BN + n) * 8 + e] =
                  f32_to_bf16_bits(w_oihw[(((size_t)n * cin + ci) * 3 + r) * 3 + s]);
            }
  upload(w_, wp);
  p.wts = w_.as<bf16>();
  p.w_bytes = (int)(wp.size() * 2);
  p.wskip_off = (int)(u_elems * 2);
  std::vector<float> sc(BN, 1.f), bi(BN, 0.f);
  if (scale) sc.assign(scale, scale + BN);
  if (bias) bi.assign(bias, bias + BN);
  upload(scale_, sc);
  upload(bias_, bi);
  p.scale = scale_.as<float>(); p.bias = bias_.as<float>();
  for (int j = 0; j < BN; ++j) { p.k.scale[j] = sc[j]; p.k.bias[j] = bi[j]; }
  flops_ = 2.0 * N * p.OH * p.OW * (double)BN * cin * 9;

  p.stage_bytes = std::max(2 * p.nslabs_u, 4 * p.nslabs_s) * kRowRunBytes;
  const int fixed = up_smem_fixed(p.w_bytes, BN);
  p.stages = std::min(20, (226 * 1024 - fixed) / p.stage_bytes);
  WSI_REQUIRE(p.stages >= 3, WSI_ERR_UNSUPPORTED, "x2 row-stream conv: not enough shared memory");
  smem_ = fixed + p.stages * p.stage_bytes;
  grid_ = (int)std::min<long long>(total, num_sms);
  CUDA_CHECK(cudaStreamSynchronize(0));
}

template <int BN, int RS>
static void launch_up(const UpParams& p, int grid, int smem, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(conv_upstream_kernel<BN, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured = true;
  }
  conv_upstream_kernel<BN, RS><<<grid, kUpThreads, smem, s>>>(p);
  CUDA_CHECK(cudaGetLastError());
}

void UpStreamOp::launch(cudaStream_t stream, LaunchCounter* lc) const {
  for (const auto& r : relayouts_) launch_relayout_planar(r.src, r.dst, r.N, r.H, r.W, r.C, r.layout, stream, lc);
  if (p_.Cout == 16) launch_up<16, 16>(p_, grid_, smem_, stream);
  else launch_up<32, 8>(p_, grid_, smem_, stream);
  if (lc) lc->n++;
}

}  // namespace wsi
