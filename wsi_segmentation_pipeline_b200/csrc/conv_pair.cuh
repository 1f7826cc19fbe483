// conv_pair.cuh — CTA-pair (tcgen05 cta_group::2) variant of the implicit-GEMM convolution, sm_100a.
//
// Why: measured on B200 (profiles/r01_notes.md) a single-CTA tcgen05.mma of 128 x N x 16 takes 64 + N/2 cycles —
// the tensor core fetches its shared-memory operands at ~64 B/clk (A: 4 KB, B: 32*N bytes) while the math needs
// only N/2 cycles — and the TMA fill of the stage ring tops out at the same ~64 B/clk per SM.  Both limits are in
// bytes per SM, so the lever is FLOPs per operand byte.  A CTA pair computes a 256 x BN tile with ONE instruction
// stream: each CTA stages its own 128 pixel rows of A and only HALF of the weight rows (BN/2), the tensor cores
// of both SMs read both halves.  Per CTA and K block that is 16 KB + BN/2 * 128 B instead of 16 KB + BN * 128 B
// for the same 128 x BN x 64 MACs: -33 % bytes at BN = 256, -25 % at BN = 128, on both limits.
//
// Protocol (ranks 0 = leader, 1 = peer of a 2-CTA cluster; barriers named as in conv_igemm.cuh):
//   * both producers wait on their OWN empty[stage] and issue their TMA loads with .cta_group::2, completing
//     transaction bytes on the LEADER's full[stage]; the leader's producer arms it with the bytes of both CTAs;
//   * only the leader's MMA warp issues tcgen05.mma.cta_group::2 (M = 256: rows 0-127 accumulate in the leader's
//     TMEM, rows 128-255 in the peer's, same columns) and tcgen05.commit...multicast::cluster arrives on
//     empty[stage] / tmem_full[acc] of BOTH CTAs;
//   * each CTA's 8 epilogue warps drain their own TMEM lanes exactly as in the single-CTA kernel and then
//     arrive (one lane per warp, release.cluster) on the leader's tmem_empty[acc] (16 arrivals).
// Tile schedule: persistent pairs; pair tile = (M-tile pair 2j, 2j+1) x N tile x parity class, N tile fastest.
// An odd M-tile count leaves the last peer with an out-of-range tile: its TMA boxes are fully out of bounds
// (zero fill) and its epilogue stores nothing.
#pragma once
#include "conv_igemm.cuh"

namespace wsi {

namespace pptx {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(ptx::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int* error_flag, int tag) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (error_flag) atomicExch(error_flag, 100 + tag);
      printf("wsi conv_pair: barrier wait timed out (tag %d, block %d, thread %d)\n", tag, blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// TMA loads whose mbarrier lives in the pair's leader CTA (cluster address)
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(ptx::smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(ptx::smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(holder)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once all prior MMAs of the pair retired
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(ptx::smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
}  // namespace pptx

template <int BN>
struct PairSmem {
  static constexpr int kBlockK = 64;
  static constexpr int kABytes = kBlockM * kBlockK * 2;             // this CTA's 128 pixel rows
  static constexpr int kBBytes = (BN / 2) * kBlockK * 2;            // this CTA's half of the weight rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (192 * 1024) / kStageBytes;        // 8 at BN = 128, 6 at BN = 256
  static constexpr int kTableBytes = 4 * 128 * (int)sizeof(KBlock);
  static constexpr int kBarBytes = 256;
  static constexpr int kScaleBytes = 2 * 512 * (int)sizeof(float);
  static constexpr int kRing = kStages * kStageBytes;
  static constexpr int kTotal = 1024 + kRing + kTableBytes + kBarBytes + kScaleBytes;
  static constexpr int kTmemCols = 2 * BN;                          // two accumulators
  static_assert(kTmemCols <= 512, "TMEM");
  static_assert(kBBytes % 1024 == 0, "swizzle atom alignment");
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
conv_igemm_pair_kernel(const __grid_constant__ AMaps amaps, const __grid_constant__ CUtensorMap bmap, const ConvParams p) {
  using S = PairSmem<BN>;
  constexpr int BLOCK_K = S::kBlockK;
  extern __shared__ uint8_t smem_raw[];
  // the dynamic window starts at the same shared address in both CTAs, so the aligned offsets agree too
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  KBlock* tbl = reinterpret_cast<KBlock*>(smem + S::kRing);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kRing + S::kTableBytes);
  uint64_t* full = bars;                               // [kStages]  used in the leader only
  uint64_t* empty = bars + S::kStages;                 // [kStages]  per CTA
  uint64_t* tmem_full = bars + 2 * S::kStages;         // [2]        per CTA
  uint64_t* tmem_empty = bars + 2 * S::kStages + 2;    // [2]        used in the leader only (16 warp arrivals)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * S::kStages + 4);
  float* s_scale = reinterpret_cast<float*>(smem + S::kRing + S::kTableBytes + S::kBarBytes);
  float* s_bias = s_scale + 512;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.num_kb;
  const uint32_t rank = pptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  for (int i = threadIdx.x; i < p.num_parity * num_kb; i += blockDim.x) tbl[i] = p.kblocks[i];
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) { s_scale[i] = p.scale[i]; s_bias[i] = p.bias[i]; }

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < S::kStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
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
  pptx::cluster_sync();          // barrier inits and the TMEM allocation of BOTH CTAs are visible from here on
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int tiles_m = p.tiles_n * p.tiles_h * p.tiles_w;
  const int pairs_m = (tiles_m + 1) >> 1;
  const int total_tiles = pairs_m * p.num_parity * p.tiles_co;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full0 = pptx::mapa(ptx::smem_u32(&full[0]), 0);     // the leader's full[] ring
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        int r = tile;
        const int ct = r % p.tiles_co; r /= p.tiles_co;
        const int par = r % p.num_parity; r /= p.num_parity;
        int mt = 2 * r + (int)rank;
        const int tw = mt % p.tiles_w; mt /= p.tiles_w;
        const int th = mt % p.tiles_h; mt /= p.tiles_h;
        const int tn = mt;                                   // == tiles_n for the out-of-range tail tile
        const int n0 = tn * p.bn, a0 = th * p.bh, b0 = tw * p.bw;
        const int co0 = ct * BN + (int)rank * (BN / 2);
        const KBlock* kb_tbl = tbl + par * num_kb;
        for (int kb = 0; kb < num_kb; ++kb) {
          const KBlock e = kb_tbl[kb];      // read BEFORE the wait (asm volatile + memory clobber would pin it after)
          ptx::mbar_wait(&empty[stage], phase ^ 1u, p.error_flag, 11);
          if (WSI_DBG(p) >= 3) {      // timing experiments (see ConvParams::dbg): 3 A only, 4 B only (2 would let the leader lap the peer)
            const bool la = (WSI_DBG(p) == 3), lb = (WSI_DBG(p) == 4);
            if (rank == 0) ptx::mbar_expect_tx(&full[stage], 2u * (uint32_t)((la ? S::kABytes : 0) + (lb ? S::kBBytes : 0)));
            uint8_t* sA = stage_base + stage * S::kStageBytes;
            const uint32_t fb = full0 + (uint32_t)(stage * 8);
            if (la) pptx::tma_load_4d_pair(sA, &amaps.m[0], fb, 0, b0, a0, n0);
            if (lb) pptx::tma_load_2d_pair(sA + S::kABytes, &bmap, fb, par * p.b_parity_stride + kb * BLOCK_K, co0);
            if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
            continue;
          }
          if (rank == 0) ptx::mbar_expect_tx(&full[stage], 2u * (uint32_t)S::kStageBytes);
          uint8_t* sA = stage_base + stage * S::kStageBytes;
          uint8_t* sB = sA + S::kABytes;
          const uint32_t fbar = full0 + (uint32_t)(stage * 8);
          // the weight tile does not depend on the table: issue it first; the table entry was read before the wait
          pptx::tma_load_2d_pair(sB, &bmap, fbar, par * p.b_parity_stride + kb * BLOCK_K, co0);
          const CUtensorMap* am = &amaps.m[0];
          switch (e.map) {
            case 1: am = &amaps.m[1]; break;
            case 2: am = &amaps.m[2]; break;
            case 3: am = &amaps.m[3]; break;
            case 4: am = &amaps.m[4]; break;
            default: break;
          }
          pptx::tma_load_4d_pair(sA, am, fbar, e.c0, b0 + e.db, a0 + e.da, n0);
          if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader only) ================================
    if (rank == 0 && ptx::elect_one()) {
      // instruction descriptor: D fp32, A/B bf16 K-major, N = BN, M = 256 over the pair
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        pptx::mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1u, p.error_flag, 12);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          pptx::mbar_wait_cluster(&full[stage], phase, p.error_flag, 13);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(stage_base + stage * S::kStageBytes);
          const uint64_t adesc = make_kmajor_desc<BLOCK_K>(a_addr);
          const uint64_t bdesc = make_kmajor_desc<BLOCK_K>(a_addr + S::kABytes);
          if (WSI_DBG(p) != 1)
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            pptx::umma_bf16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          pptx::umma_commit_pair(&empty[stage]);      // frees this stage in both CTAs
          if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
        }
        pptx::umma_commit_pair(&tmem_full[acc]);      // accumulator complete -> both epilogues
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
      const int par = r % p.num_parity; r /= p.num_parity;
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
      const int oh = p.sigma * a + (par >> 1), ow = p.sigma * b + (par & 1);
      const size_t pix = ((size_t)n * p.OH + oh) * p.OW + ow;
      const size_t off0 = pix * p.Cout + co0 + c_lo;
      const bool has_res = (p.res != nullptr) && valid;

      uint4 rcur[STEP / 8], rnext[STEP / 8];
      if (has_res) {
#pragma unroll
        for (int j = 0; j < STEP / 8; j += 2) ptx::ld_global_nc_256(p.res + off0 + 8 * j, rcur[j], rcur[j + 1]);
      }

      ptx::mbar_wait(&tmem_full[acc], acc_phase, p.error_flag, 14);
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
          uint8_t* ob = reinterpret_cast<uint8_t*>(p.out + off0 + c);
#pragma unroll
          for (int j = 0; j < STEP / 16; ++j) {        // 256-bit stores, one full sector each
            uint32_t w[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(y[16 * j + 2 * t], y[16 * j + 2 * t + 1]);
              w[t] = *reinterpret_cast<uint32_t*>(&h2);
            }
            ptx::st_global_256(ob + (size_t)j * 32, w);
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
