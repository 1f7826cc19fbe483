// conv_igemm.cuh — implicit-GEMM convolution on 5th-gen tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// Every dense contraction on the reference's hot path is a convolution over NHWC tiles
// (SURVEY.md §2a K1-K3, K5): 7x7/s2 stem (resnets_shift.py:122), 3x3 s1/s2 BasicBlock convs
// (:41-44), 1x1/s2 projections (:173-177) and the smp decoder's upsample+concat+3x3 convs.
// All of them run through ONE table-driven kernel:
//
//   D[128 output pixels, BLOCK_N out channels] += A[128 pixels, BLOCK_K] * B[BLOCK_N, BLOCK_K]^T
//
// * A (activations, bf16 NHWC) is never im2col'ed in memory: for each K block (one filter tap x
//   one channel chunk) the producer issues a 4-D TMA box load {BLOCK_K ch, bw, bh, bn} whose
//   origin is shifted by the tap offset; TMA's out-of-bounds zero fill IS the conv zero padding.
// * stride-2 convs read through "parity views" of the input (one tensor map per (row,col) parity
//   with doubled strides), so the same box load works.
// * the decoder's nearest x2 upsample + channel concat is folded into the operand addressing:
//   output pixels are processed per parity class (oh%2, ow%2), for which every tap of the
//   upsampled operand is a plain box of the half-resolution tensor, and the skip operand is a
//   parity view.  No upsampled or concatenated tensor ever exists in HBM.  Within a parity class
//   the 9 taps of the upsampled operand read only 2 x 2 distinct half-resolution pixels, so their
//   weights are summed (fp32, per parity) on the host: 4 K blocks per channel chunk instead of 9.
// * the stem reads the gather kernel's zero-padded 4-channel tiles through an overlapping-window
//   map (dim1 stride 16 B): one K block = one filter row = 8 px x 4 ch = 32 elements.
// * B (weights) is pre-packed [Cout][K] bf16, K ordered exactly as the K-block table.
// * accumulators live in TMEM (double buffered: the epilogue of tile i overlaps the MMAs of
//   tile i+1); the epilogue applies folded BN scale/bias (+ residual) (+ ReLU) in fp32 and
//   stores bf16 NHWC, or for the last decoder conv applies the fused 1x1 `final_conv` head and
//   stores fp32 logits.
// * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2-5 = epilogue.
//   Persistent CTAs (grid = min(tiles, #SM)), static round-robin tile schedule, N-tile fastest so
//   CTAs sharing an A tile run together and hit L2.
#pragma once
#include <memory>

#include "common.cuh"

namespace wsi {

constexpr int kMaxAMaps = 6;
constexpr int kBlockM = 128;
constexpr int kNumThreads = 320;     // TMA producer warp, MMA issuer warp, 8 epilogue warps

struct KBlock {      // one K block of the implicit GEMM
  int8_t map;        // which A tensor map
  int8_t da, db;     // box origin offset in (row, col) of that map's lattice
  int8_t _pad;
  int32_t c0;        // channel origin in that map
};

struct ConvParams {
  int N, OH, OW, Cout;        // output tensor [N, OH, OW, Cout]
  int sigma;                  // output lattice stride: pixel = (sigma*a + py, sigma*b + px)
  int num_parity;             // 1, or 4 when sigma == 2 (py, px in {0,1}^2)
  int A_h, A_w;               // lattice extents
  int bw, bh, bn;             // M tile = bn x bh x bw lattice points (product 128)
  int tiles_w, tiles_h, tiles_n, tiles_co;
  int num_kb;                 // K blocks per parity
  int relu;
  const float* scale;         // [Cout] folded BN scale (or 1)
  const float* bias;          // [Cout] folded BN bias (or conv bias)
  const bf16* res;            // residual, same shape as out, or nullptr
  bf16* out;                  // may be nullptr when head_out is set
  const float* head_w;        // fused 1x1 head: [4][16]
  const float* head_b;        // [4]
  float* head_out;            // [N, OH, OW, 4] fp32
  const KBlock* kblocks;      // [num_parity][num_kb]
  int b_parity_stride;        // K offset (elements) between the packed weights of consecutive parity classes (0: shared)
  int* error_flag;            // set to 1 by a timed-out barrier wait
  int out_planar;             // 1: `out` is the zero-padded channel-chunk-planar layout of conv_rowtile.cuh (consumer = a row kernel)
  long long pl_img, pl_row, pl_chunk;   // its byte strides: image, image row, 8-channel chunk row (entry = x + 8, 16 B each)
  const struct HaloGroup* hgroups;   // halo-resident kernel (conv_halo.cuh): [num_parity][num_kb] groups
  int halo_plain;             // ... plain 3x3/s1 conv: every group is one 64-channel chunk with the 9 taps in (r, s) order
  // fp32-emulated precision (NSPLIT = 3): every bf16 tensor is three planes a + b + c = the fp32 value (8 + 8 + 8
  // mantissa bits, exact), activations as channel blocks [a | b | c] of one NHWC tensor (stem: three tensors), weights
  // as three packed matrices; each K block issues the 6 plane products whose weight is >= 2^-24:
  // (a,a) (a,b) (b,a) (b,b) (a,c) (c,a).  A plane j of map m is reached by channel offset j * a_plane[m] and map
  // offset j * split_map_step.
  int a_plane[kMaxAMaps];
  int split_map_step;
  int dbg;                    // -DWSI_DEBUG_SWITCHES builds only (pair kernel timing experiments)
};

// Plane pairs (activation plane, weight plane) of the fp32 emulation, SMALLEST products first.  The tensor core adds
// into its fp32 accumulator with truncation (measured here: ~2e-8 x |acc| per 128xNx16 MMA, biased — a 4608-long K
// loop of six products loses 2.6e-5), so the order and length of the accumulation chains matter:
//   * the five correction products (weights 2^-16, 2^-16, 2^-16, 2^-8, 2^-8) are accumulated first, all K blocks, into
//     one TMEM accumulator: their chain stays ~2^-8 of the result, its truncation ~1e-10;
//   * the main product (a, a) is accumulated in short chains of kSplitChunkSteps MMAs, each drained by the epilogue
//     warps into fp32 REGISTERS (round-to-nearest adds), double-buffered against the next chain.
__device__ __constant__ const int8_t kSplitA[6] = {0, 2, 1, 0, 1, 0};
__device__ __constant__ const int8_t kSplitB[6] = {2, 0, 1, 1, 0, 0};
constexpr int kSplitChunkSteps = 8;     // MMAs (K = 16 each) per main-product chain

struct AMaps {
  CUtensorMap m[kMaxAMaps];
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of the (converged) warp; the compiler treats code under it as single-threaded, so operands of
// tcgen05 / TMA instructions move to uniform registers without a per-instruction waterfall loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU box (a hang costs a strike); after ~2 s of
// SM clocks the kernel records the failure and traps.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* error_flag, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (error_flag) atomicExch(error_flag, 100 + tag);
      printf("wsi conv_igemm: barrier wait timed out (tag %d, block %d, thread %d)\n", tag, blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 256-bit global store / read-only load (sm_100: STG.256 / LDG.256): one full 32-byte sector per thread and instruction
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
               "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_nc_256(const void* ptr, uint4& lo, uint4& hi) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(ptr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 16 accumulator columns of this warp's 32 lanes <- 0.  The row kernels hand an accumulator slot back ZEROED, so the MMA
// that opens the next output row in it can accumulate like every other one (no separate overwriting MMA per row).
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "{\n\t.reg .b32 z;\n\tmov.b32 z, 0;\n\t"
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {z,z,z,z,z,z,z,z,z,z,z,z,z,z,z,z};\n\t}" ::"r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// all accumulator columns [0, cols) <- 0 by the epilogue warps (warps 2 .., groups of 4 covering the 4 lane quarters)
__device__ __forceinline__ void tmem_zero_all(uint32_t tmem_base, int warp, int n_groups, int cols) {
  const int grp = (warp - 2) >> 2, per = cols / n_groups;
  const uint32_t tb = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  for (int c = grp * per; c < (grp + 1) * per; c += 16) tmem_st16_zero(tb + (uint32_t)c);
  tmem_st_wait();
}
}  // namespace ptx

// SM100 shared-memory matrix descriptor, K-major operand, rows of (BLOCK_K*2) bytes packed densely,
// swizzle width == row width (128/64/32 B); 8-row core-matrix groups are 8*row bytes apart (SBO).
template <int BLOCK_K>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  constexpr uint32_t row_bytes = BLOCK_K * 2;
  constexpr uint64_t layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);  // UMMA::LayoutType
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address  [0,14)
  d |= (uint64_t)1 << 16;                              // LBO (ignored for swizzled K-major) [16,30)
  d |= (uint64_t)((8u * row_bytes) >> 4) << 32;        // SBO [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell) [46,48)
  d |= layout << 61;                                   // swizzle mode [61,64)
  return d;
}

template <int BLOCK_N>
__host__ __device__ constexpr uint32_t make_idesc_bf16() {
  return (1u << 4)                       // D format  = F32
         | (1u << 7)                     // A format  = BF16
         | (1u << 10)                    // B format  = BF16
         | (0u << 15) | (0u << 16)       // A, B K-major
         | ((uint32_t)(BLOCK_N >> 3) << 17)
         | ((uint32_t)(kBlockM >> 4) << 24);
}

constexpr int kResidentBBytes = 96 * 1024;   // RESB: the whole weight tensor of the conv stays in smem

template <int BLOCK_N, int BLOCK_K, bool RESB = false>
struct ConvSmem {
  static constexpr int kABytes = kBlockM * BLOCK_K * 2;
  static constexpr int kBBytesRaw = BLOCK_N * BLOCK_K * 2;
  static constexpr int kBBytes = (kBBytesRaw + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = RESB ? kABytes : kABytes + kBBytes;
  static constexpr int kResBytes = RESB ? kResidentBBytes : 0;
  static constexpr int kStagesWanted = ((RESB ? 112 : (BLOCK_N > 128 ? 200 : 160)) * 1024) / kStageBytes;
  static constexpr int kStages = kStagesWanted > 8 ? 8 : (kStagesWanted < 2 ? 2 : kStagesWanted);
  static constexpr int kTableBytes = 4 * 128 * (int)sizeof(KBlock);   // up to 4 parities x 128 K blocks
  static constexpr int kBarBytes = 256;
  static constexpr int kScaleBytes = 2 * 512 * (int)sizeof(float);    // folded-BN scale + bias of up to 512 channels
  static constexpr int kRing = kStages * kStageBytes + kResBytes;     // stage ring (+ resident weights)
  static constexpr int kTotal = 1024 /*align slack*/ + kRing + kTableBytes + kBarBytes + kScaleBytes;
  static constexpr int kTmemCols = (2 * BLOCK_N) < 32 ? 32 : (2 * BLOCK_N);     // two accumulators (512 columns at BLOCK_N = 256)
};

// RESB: convs whose whole weight tensor fits in kResidentBBytes (one output-channel tile) load B ONCE per CTA
// and the stage ring carries only A — the mainloop is bound by the TMA box-row rate, and B is a third (N=64)
// to a half (N=128) of the rows of a K block.
template <int BLOCK_N, int BLOCK_K, bool RESB = false, int NSPLIT = 1>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_igemm_kernel(const __grid_constant__ AMaps amaps, const __grid_constant__ CUtensorMap bmap, const ConvParams p) {
  static_assert(NSPLIT == 1 || (NSPLIT == 3 && !RESB), "NSPLIT: 1 (bf16) or 3 (fp32 emulation, streamed weights)");
  constexpr int NPROD = (NSPLIT == 3) ? 6 : 1;      // plane products per K block
  using S = ConvSmem<BLOCK_N, BLOCK_K, RESB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* b_res = smem + S::kStages * S::kStageBytes;                 // RESB only
  KBlock* tbl = reinterpret_cast<KBlock*>(smem + S::kRing);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kRing + S::kTableBytes);
  uint64_t* full = bars;                       // [kStages]
  uint64_t* empty = bars + S::kStages;         // [kStages]
  uint64_t* tmem_full = bars + 2 * S::kStages;     // [2]
  uint64_t* tmem_empty = bars + 2 * S::kStages + 2;  // [2]
  uint64_t* bres_bar = bars + 2 * S::kStages + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * S::kStages + 5);
  float* s_scale = reinterpret_cast<float*>(smem + S::kRing + S::kTableBytes + S::kBarBytes);
  float* s_bias = s_scale + 512;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.num_kb;

  // K-block table and epilogue constants -> smem
  for (int i = threadIdx.x; i < p.num_parity * num_kb; i += blockDim.x) tbl[i] = p.kblocks[i];
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) { s_scale[i] = p.scale[i]; s_bias[i] = p.bias[i]; }

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < S::kStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 256);
    }
    ptx::mbar_init(bres_bar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&bmap);
    ptx::prefetch_tmap(&amaps.m[0]);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_holder, S::kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int tiles_per_par = p.tiles_n * p.tiles_h * p.tiles_w * p.tiles_co;
  const int total_tiles = tiles_per_par * p.num_parity;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      if (RESB && blockIdx.x < total_tiles) {
        ptx::mbar_expect_tx(bres_bar, (uint32_t)num_kb * S::kBBytesRaw);
        for (int kb = 0; kb < num_kb; ++kb) ptx::tma_load_2d(b_res + kb * S::kBBytes, &bmap, bres_bar, kb * BLOCK_K, 0);
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int r = tile;
        const int ct = r % p.tiles_co; r /= p.tiles_co;
        const int par = r % p.num_parity; r /= p.num_parity;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h; r /= p.tiles_h;
        const int tn = r;
        const int n0 = tn * p.bn, a0 = th * p.bh, b0 = tw * p.bw, co0 = ct * BLOCK_N;
        const KBlock* kb_tbl = tbl + par * num_kb;
        // fp32 emulation: product-major (all K blocks of product 0, then of product 1, ... the main product last)
#pragma unroll 1
        for (int pr = 0; pr < NPROD; ++pr) {
#pragma unroll 1
          for (int kb = 0; kb < num_kb; ++kb) {
            const KBlock e = kb_tbl[kb];      // read BEFORE the wait (asm volatile + memory clobber would pin it after)
            int mi = e.map, c0 = e.c0;
            if (NSPLIT == 3) { mi += kSplitA[pr] * p.split_map_step; c0 += kSplitA[pr] * p.a_plane[e.map]; }
            ptx::mbar_wait(&empty[stage], phase ^ 1u, p.error_flag, 1);
            ptx::mbar_expect_tx(&full[stage], RESB ? S::kABytes : S::kABytes + S::kBBytesRaw);
            uint8_t* sA = stage_base + stage * S::kStageBytes;
            uint8_t* sB = sA + S::kABytes;
            const CUtensorMap* am = &amaps.m[0];
            switch (mi) {
              case 1: am = &amaps.m[1]; break;
              case 2: am = &amaps.m[2]; break;
              case 3: am = &amaps.m[3]; break;
              case 4: am = &amaps.m[4]; break;
              case 5: am = &amaps.m[5]; break;
              default: break;
            }
            if (!RESB) ptx::tma_load_2d(sB, &bmap, &full[stage], par * p.b_parity_stride + (kb * NPROD + pr) * BLOCK_K, co0);
            ptx::tma_load_4d(sA, am, &full[stage], c0, b0 + e.db, a0 + e.da, n0);
            if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16<BLOCK_N>();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (RESB && blockIdx.x < total_tiles) ptx::mbar_wait(bres_bar, 0, p.error_flag, 5);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u, p.error_flag, 2);
        ptx::tc_fence_after();
        uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        constexpr int kChunkKb = (kSplitChunkSteps * 16 >= BLOCK_K) ? kSplitChunkSteps * 16 / BLOCK_K : 1;   // K blocks per main chain
        int chain0 = 0;                       // first K-block index (product-major order) of the current accumulation chain
        for (int kb = 0; kb < num_kb * NPROD; ++kb) {
          if (NSPLIT == 3 && kb >= num_kb * (NPROD - 1) && kb > chain0 && (kb - num_kb * (NPROD - 1)) % kChunkKb == 0) {
            // close the chain (corrections, or kChunkKb main K blocks): hand it to the epilogue, continue in the other buffer
            ptx::umma_commit(&tmem_full[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
            ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u, p.error_flag, 2);
            ptx::tc_fence_after();
            d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
            chain0 = kb;
          }
          ptx::mbar_wait(&full[stage], phase, p.error_flag, 3);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(stage_base + stage * S::kStageBytes);
          const uint32_t b_addr = RESB ? ptx::smem_u32(b_res + kb * S::kBBytes) : a_addr + S::kABytes;
          const uint64_t adesc = make_kmajor_desc<BLOCK_K>(a_addr);
          const uint64_t bdesc = make_kmajor_desc<BLOCK_K>(b_addr);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            // advancing 16 bf16 = 32 B along K inside the swizzle atom: +2 in the (addr>>4) field
            ptx::umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)(((kb - chain0) | k) != 0));
          }
          ptx::umma_commit(&empty[stage]);      // frees the smem stage when these MMAs retire
          if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit(&tmem_full[acc]);      // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ================================ epilogue (8 warps) ==========================
    // Two warps per TMEM lane quadrant, each draining half of the tile's columns (BLOCK_N >= 32), so
    // the per-tile drain latency halves.  Folded-BN constants come from smem (loaded once per CTA),
    // columns are processed 16/32 at a time with the residual for the NEXT chunk already in flight.
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int hsel = (warp - 2) >> 2;          // which half of the columns
    const int row = q * 32 + lane;             // M index inside the tile == TMEM lane
    constexpr int CH = (BLOCK_N >= 32) ? BLOCK_N / 2 : BLOCK_N;      // columns per warp
    constexpr int STEP = (CH >= 32) ? 32 : 16;                       // columns per iteration
    const bool idle = (BLOCK_N < 32) && (hsel == 1);
    const int c_lo = (BLOCK_N >= 32) ? hsel * CH : 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int r = tile;
      const int ct = r % p.tiles_co; r /= p.tiles_co;
      const int par = r % p.num_parity; r /= p.num_parity;
      const int tw = r % p.tiles_w; r /= p.tiles_w;
      const int th = r % p.tiles_h; r /= p.tiles_h;
      const int tn = r;
      const int co0 = ct * BLOCK_N;
      // lattice point of this row
      const int wl = row % p.bw;
      const int hl = (row / p.bw) % p.bh;
      const int nl = row / (p.bw * p.bh);
      const int n = tn * p.bn + nl, a = th * p.bh + hl, b = tw * p.bw + wl;
      const bool valid = (n < p.N) && (a < p.A_h) && (b < p.A_w) && !idle;
      const int oh = p.sigma * a + (par >> 1), ow = p.sigma * b + (par & 1);
      const size_t pix = ((size_t)n * p.OH + oh) * p.OW + ow;
      const size_t off0 = pix * (size_t)(NSPLIT * p.Cout) + co0 + c_lo;     // NSPLIT planes as channel blocks [a | b | c]
      const size_t pl_off = (size_t)n * (size_t)p.pl_img + (size_t)(oh + 1) * (size_t)p.pl_row + (size_t)(ow + 8) * 16;
      const bool has_res = (p.res != nullptr) && valid;

      // residual of the first chunk: issued before waiting for the accumulator
      uint4 rcur[STEP / 8], rnext[STEP / 8];
      if (has_res && NSPLIT == 1) {
#pragma unroll
        for (int j = 0; j < STEP / 8; ++j) rcur[j] = __ldg(reinterpret_cast<const uint4*>(p.res + off0) + j);
      }

      // fp32 emulation: the accumulation chains of this tile (corrections, then the main product in short chains) arrive
      // one by one in alternating TMEM buffers and are summed here in fp32 registers (round-to-nearest)
      float racc[NSPLIT == 3 ? CH : 1];
      if (NSPLIT == 3) {
        constexpr int kChunkKb = (kSplitChunkSteps * 16 >= BLOCK_K) ? kSplitChunkSteps * 16 / BLOCK_K : 1;
        const int n_chains = 1 + (num_kb + kChunkKb - 1) / kChunkKb;
#pragma unroll
        for (int j = 0; j < CH; ++j) racc[j] = 0.f;
#pragma unroll 1
        for (int chn = 0; chn < n_chains; ++chn) {
          ptx::mbar_wait(&tmem_full[acc], acc_phase, p.error_flag, 4);
          ptx::tc_fence_after();
          if (!idle) {
            const uint32_t t_chain = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c_lo);
#pragma unroll
            for (int c = 0; c < CH; c += STEP) {
              uint32_t v[STEP];
#pragma unroll
              for (int j = 0; j < STEP; j += 16) ptx::tmem_ld16(t_chain + (uint32_t)(c + j), *reinterpret_cast<uint32_t(*)[16]>(&v[j]));
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < STEP; ++j) racc[c + j] += __uint_as_float(v[j]);
            }
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(&tmem_empty[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      } else {
        ptx::mbar_wait(&tmem_full[acc], acc_phase, p.error_flag, 4);
        ptx::tc_fence_after();
      }
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c_lo);

      float head_acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (!idle) {
#pragma unroll (NSPLIT == 3 ? 8 : 1)
        for (int c = 0; c < CH; c += STEP) {
          uint32_t v[STEP];
          if (NSPLIT == 3) {
#pragma unroll
            for (int j = 0; j < STEP; ++j) v[j] = __float_as_uint(racc[c + j]);
          } else {
#pragma unroll
            for (int j = 0; j < STEP; j += 16) ptx::tmem_ld16(t_row + (uint32_t)(c + j), *reinterpret_cast<uint32_t(*)[16]>(&v[j]));
          }
          if (has_res && NSPLIT == 1 && c + STEP < CH) {
#pragma unroll
            for (int j = 0; j < STEP / 8; ++j) rnext[j] = __ldg(reinterpret_cast<const uint4*>(p.res + off0 + c + STEP) + j);
          }
          uint4 rsp[NSPLIT == 3 ? 3 * (STEP / 8) : 1];
          if (NSPLIT == 3 && has_res) {
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
#pragma unroll
              for (int j = 0; j < STEP / 8; ++j) rsp[pl * (STEP / 8) + j] = __ldg(reinterpret_cast<const uint4*>(p.res + off0 + (size_t)pl * p.Cout + c) + j);
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
          if (NSPLIT == 3 && has_res) {
            // residual = a + b + c, summed big to small: exact (the planes are a non-overlapping expansion of an fp32 value)
#pragma unroll
            for (int j = 0; j < STEP / 8; ++j) {
              float rv[8];
#pragma unroll
              for (int pl = 0; pl < 3; ++pl) {
                const uint4 q = rsp[pl * (STEP / 8) + j];
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float lo = __uint_as_float(w[t] << 16), hi = __uint_as_float(w[t] & 0xffff0000u);
                  rv[2 * t] = (pl == 0) ? lo : rv[2 * t] + lo;
                  rv[2 * t + 1] = (pl == 0) ? hi : rv[2 * t + 1] + hi;
                }
              }
#pragma unroll
              for (int t = 0; t < 8; ++t) y[8 * j + t] += rv[t];
            }
          }
          if (NSPLIT == 1 && has_res) {
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
          if (p.head_out != nullptr) {
            // fused final 1x1 conv (Cout == BLOCK_N == 16): logits = W[4x16] y + b
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float sacc = 0.f;
#pragma unroll
              for (int j = 0; j < 16; ++j) sacc = fmaf(y[j], __ldg(p.head_w + k * 16 + j), sacc);
              head_acc[k] += sacc;
            }
          }
          if (NSPLIT == 3 && valid && p.out != nullptr) {
            // three bf16 planes of the fp32 result: a = rn(y), b = rn(y - a), c = rn(y - a - b) (the last is exact)
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) {
              uint8_t* ob = reinterpret_cast<uint8_t*>(p.out + off0 + (size_t)pl * p.Cout + c);
#pragma unroll
              for (int j = 0; j < STEP / 8; ++j) {
                uint32_t w[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const __nv_bfloat162 h2 = __floats2bfloat162_rn(y[8 * j + 2 * t], y[8 * j + 2 * t + 1]);
                  w[t] = *reinterpret_cast<const uint32_t*>(&h2);
                  y[8 * j + 2 * t] -= __low2float(h2);
                  y[8 * j + 2 * t + 1] -= __high2float(h2);
                }
                *reinterpret_cast<uint4*>(ob + (size_t)j * 16) = make_uint4(w[0], w[1], w[2], w[3]);
              }
            }
          }
          if (NSPLIT == 1 && valid && p.out != nullptr) {
            // NHWC: 2 * STEP contiguous bytes; planar: one 16-byte entry per 8-channel chunk row
            uint8_t* ob = p.out_planar ? reinterpret_cast<uint8_t*>(p.out) + pl_off + (size_t)((co0 + c_lo + c) >> 3) * (size_t)p.pl_chunk
                                       : reinterpret_cast<uint8_t*>(p.out + off0 + c);
            const size_t ostep = p.out_planar ? (size_t)p.pl_chunk : 16;
            if (STEP == 32 && !p.out_planar) {
#pragma unroll
              for (int j = 0; j < STEP / 16; ++j) {
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
        if (valid && p.head_out != nullptr) {
          float4 o;
          o.x = head_acc[0] + __ldg(p.head_b + 0);
          o.y = head_acc[1] + __ldg(p.head_b + 1);
          o.z = head_acc[2] + __ldg(p.head_b + 2);
          o.w = head_acc[3] + __ldg(p.head_b + 3);
          reinterpret_cast<float4*>(p.head_out)[pix] = o;
        }
      }
      if (NSPLIT == 1) {
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tmem_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, S::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// Activation layouts in HBM (bf16).  NHWC everywhere except around the row-tile kernel, whose
// operands are zero-padded channel-chunk-planar (conv_rowtile.cuh).
enum ActLayout { LAYOUT_NHWC = 0, LAYOUT_PLANAR = 1, LAYOUT_PLANAR_PARITY = 2 };

struct TensorView {          // bf16 activation tensor in HBM
  const void* ptr = nullptr;
  int N = 0, H = 0, W = 0, C = 0;
  int layout = LAYOUT_NHWC;
};

// One operand part of a conv input (the input is the channel concat of its parts).
struct ConvInputPart {
  TensorView t;
  bool up2 = false;          // nearest x2 upsampled view of t
};

struct ConvSpec {
  int ksize = 3, stride = 1, pad = 1;
  int cout = 0;
  bool relu = false;
  bool head = false;         // fused final 1x1 conv (cout must be 16)
};

class RowConvOp;   // conv_rowtile.cuh: halo-resident kernel for the small-channel 3x3 layers
class RowStemOp;   // conv_rowtile.cuh: the stem in the same style
class RowStreamOp; // conv_rowstream.cuh: row-streaming kernel for plain 3x3 convs on <= 64 channels
class UpStreamOp;  // conv_upstream.cuh: row-streaming kernel for the x2-upsampling decoder convs (Cout 16 / 32)

// A fully prepared conv launch: tensor maps, K-block table, packed weights, epilogue params.
// build() routes small-channel 3x3/s1 convs to the row-tile kernel (conv_rowtile.cuh) and
// everything else to the TMA implicit-GEMM kernel below.
class ConvOp {
 public:
  ConvOp();
  ~ConvOp();
  ConvOp(ConvOp&&) noexcept;
  ConvOp& operator=(ConvOp&&) noexcept;

  // weights: fp32 OIHW [cout][cin_total][k][k]; scale/bias may be nullptr (1 / 0).
  // out_layout: LAYOUT_NHWC, or LAYOUT_PLANAR when the consumer is a row-tile conv (row-tile producers only).
  void build(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const float* w_oihw,
             const float* scale, const float* bias, const void* residual, void* out,
             const float* head_w, const float* head_b, float* head_out, int* error_flag, int num_sms,
             int out_layout = LAYOUT_NHWC, int res_layout = LAYOUT_NHWC, int split = 0);
  bool is_rowtile() const { return (bool)row_; }
  // would build() route this conv to the row-tile kernel?  (lets the caller chain planar layouts)
  static bool routes_to_rowtile(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const void* residual);
  static bool routes_to_upstream(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const void* residual, int out_layout);
  // stem: x = gather output, zero-padded tiles [n][ph+6][pw+8][4] bf16; 7x7/s2/p3, cout 64.
  // out_layout: LAYOUT_NHWC or LAYOUT_PLANAR_PARITY (row-tile stem only; see stem_routes_to_rowtile)
  void build_stem(const void* padded_tiles, int n, int ph, int pw, const float* w_oihw /*[64,3,7,7]*/,
                  const float* scale, const float* bias, void* out, int* error_flag, int num_sms,
                  int out_layout = LAYOUT_NHWC, int split = 0);
  static bool stem_routes_to_rowtile();
  void launch(cudaStream_t stream, LaunchCounter* lc) const;
  // re-point the fused-head output (fp32 logits [N,OH,OW,4]) of a conv built with one: the engine writes each batch
  // straight into its slot of the logit ring
  void set_head_out(float* p);
  double flops() const { return flops_; }
  int block_n() const { return block_n_; }
  int block_k() const { return block_k_; }
  bool is_pair() const { return pair_ || halo_; }
  bool is_halo() const { return halo_; }
  std::string kernel_name() const;   // the __global__ function this op launches (evidence tables)

 private:
  void finish(const std::vector<KBlock>& table, int num_parity, const std::vector<uint16_t>& wpacked, int K,
              const float* scale, const float* bias, int num_sms);
  AMaps amaps_{};
  CUtensorMap bmap_{};
  ConvParams p_{};
  DevBuf w_, scale_, bias_, tbl_, headw_, headb_;
  std::unique_ptr<RowConvOp> row_;
  std::unique_ptr<RowStemOp> stem_;
  std::unique_ptr<RowStreamOp> stream_;
  std::unique_ptr<UpStreamOp> upstream_;
  int block_n_ = 0, block_k_ = 0, grid_ = 0;
  bool resb_ = false;         // weights resident in smem (see conv_igemm_kernel RESB)
  bool pair_ = false;         // CTA-pair kernel (conv_pair.cuh)
  bool halo_ = false;         // halo-resident CTA-pair kernel (conv_halo.cuh)
  static bool halo_eligible(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, int out_layout, bool head, int num_sms);
  void build_halo(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const float* w_oihw, const float* scale, const float* bias,
                  const void* residual, void* out, int out_layout, int* error_flag, int num_sms);
  DevBuf hgroups_;
  bool out_planar_ = false;   // TMA kernel writing the planar layout for a row-kernel consumer
  int split_ = 0;             // fp32 emulation: three bf16 planes per tensor, 6 plane products per K block (NSPLIT = 3)
  double flops_ = 0;
};

void init_tensor_map_api();
// hardware probe (dev tool): shifted SWIZZLE_128B A operand inside a TMA-written halo tile, see conv_igemm.cu
void debug_umma_shift(const void* A_dev, const void* B_dev, int r, int s, int pitch, int use_base_offset, float* D_dev, cudaStream_t st);   // resolves cuTensorMapEncodeTiled through the runtime (no -lcuda link)

}  // namespace wsi
