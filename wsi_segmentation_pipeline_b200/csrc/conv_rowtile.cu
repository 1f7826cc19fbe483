// conv_rowtile.cu — see conv_rowtile.cuh.
#include "conv_rowtile.cuh"

#include <algorithm>
#include <cstdlib>

namespace wsi {

namespace ptx {
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
};

template <int BN, bool HEAD, int G>
__global__ void __launch_bounds__(kRowThreads, 1) conv_rowtile_kernel(const RowParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t s_adelta[3][2][G][9];     // [mode][row parity][column parity group][tap] in 16-byte units
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int w_bytes = p.nslabs * 9 * 2 * BN * 16;
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
  // accumulator ring: kRowAccStages tiles in flight between the MMA issuer and the epilogue (the tiles are
  // short, so commit -> wait -> drain -> release latency would otherwise idle the tensor pipe)
  constexpr uint32_t kTmemCols = kRowAccStages * G * BN;
  static_assert(kTmemCols >= 32 && kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns must be a power of two");

  // resident weights + epilogue constants (generic-proxy writes, read by the async proxy -> fence)
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = threadIdx.x; i < w_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < BN; i += blockDim.x) { s_scale[i] = p.scale[i]; s_bias[i] = p.bias[i]; }
    if (HEAD)
      for (int i = threadIdx.x; i < 68; i += blockDim.x) s_hw[i] = (i < 64) ? p.head_w[i] : p.head_b[i - 64];
    ptx::fence_proxy_async();
  }
  // A-operand start offsets (16-byte units) of every tap, relative to the stage base
  for (int i = threadIdx.x; i < 3 * 2 * G * 9; i += blockDim.x) {
    const int tap = i % 9, g = (i / 9) % G, py = (i / (9 * G)) % 2, mode = i / (18 * G);
    const int r = tap / 3, sft = tap % 3;
    int par = 0, poff;
    if (mode == 0) {
      poff = r * kRowHaloCols + sft;
    } else if (mode == 1) {
      poff = (((py + r - 1) >> 1) + 1) * kRowHaloCols + (((g + sft - 1) >> 1) + 1);
    } else {
      const int qq = g + sft - 1;
      par = qq & 1;
      poff = r * kRowHaloCols + ((qq >> 1) + 1);
    }
    s_adelta[mode][py][g][tap] = (uint32_t)((par * 2 * kRowPlaneBytes + poff * 16) >> 4);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      ptx::mbar_init(&full[i], kRowProducers);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kRowAccStages; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == kRowProducerWarps) ptx::tmem_alloc(tmem_holder, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // contiguous tile range per CTA: coordinates advance incrementally (no divisions in the loops) and a
  // CTA walks down consecutive rows, so two of its three halo rows were just read by itself (L2 hits)
  const int t_begin = (int)((long long)p.total_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)p.total_tiles * (blockIdx.x + 1) / gridDim.x);

  if (warp < kRowProducerWarps) {
    // ================================ cp.async producers ================================
    // lanes = (8-channel chunk kc, 16 consecutive halo columns): 8 consecutive lanes write 128
    // contiguous smem bytes (no bank conflicts) and read whole 32-byte sectors from L2.
    // Everything that depends only on (tile, operand part) is computed once per part; a 16-channel
    // slab then costs one wait, <= 9 cp.async with one IMAD + one compare each, and one arrive.
    const int kc = lane >> 4, cxl = lane & 15;
    const int ry = warp % 3, hi = warp / 3;          // this warp's halo row; parity (mode 2) or column half
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t stage0 = ptx::smem_u32(s_stage);
    const uint32_t lane_dst = (uint32_t)(kc * kRowPlaneBytes + cxl * 16);
    TileCoord tc;
    tc.init(t_begin, p.tiles_x, p.OH);
    for (int tile = t_begin; tile < t_end; ++tile, tc.next(p.tiles_x, p.OH)) {
      int sl = 0;
      for (int pi = 0; pi < p.nparts; ++pi) {
        const bf16* ptr = p.part[pi].ptr;
        const int H = p.part[pi].H, W = p.part[pi].W, C = p.part[pi].C, mode = p.part[pi].mode;
        const int nsl = C >> 4;
        const int par = (mode == 2) ? hi : 0;
        const int col0 = (mode == 2) ? 0 : hi * 65;
        const int sy = ((mode == 1) ? (tc.y >> 1) : tc.y) - 1 + ry;
        const bool row_ok = (sy >= 0) && (sy < H);
        const int c0 = tc.xb * 128 - 1 + col0 + cxl;
        const int sx0 = (mode == 2) ? (2 * c0 + par) : c0;
        const int xstep = (mode == 2) ? 32 : 16;     // source columns per unrolled step
        // this lane's 9 (5) chunks: validity mask and byte offsets are the same for every slab of the part
        uint32_t okmask = 0;
#pragma unroll
        for (int it = 0; it < 9; ++it) {
          const int sx = sx0 + it * xstep;
          const bool in_task = (mode == 2) ? (cxl + 16 * it < kRowHaloCols) : (it < 5 && cxl + 16 * it < 65);
          if (in_task) okmask |= 1u << (16 + it);
          if (in_task && row_ok && (unsigned)sx < (unsigned)W) okmask |= 1u << it;
        }
        const bf16* rowp = ptr + ((size_t)tc.n * H + (row_ok ? sy : 0)) * W * C + kc * 8;
        const int off0 = sx0 * C, offstep = xstep * C;       // element offsets (fit 32 bits: one image row)
        const uint32_t dst_part = lane_dst + (uint32_t)(par * 2 * kRowPlaneBytes + (ry * kRowHaloCols + col0) * 16);
        for (int j = 0; j < nsl; ++j, ++sl) {
          const bf16* srcb = rowp + j * 16;
          const uint32_t dst = stage0 + (uint32_t)stage * p.stage_bytes + dst_part;
          ptx::mbar_wait(&empty[stage], phase ^ 1u, p.error_flag, 11);
#pragma unroll
          for (int it = 0; it < 9; ++it) {
            if (okmask & (1u << (16 + it))) {
              const bool ok = (okmask >> it) & 1u;
              ptx::cp_async16_zfill(dst + 256u * it, srcb + (ok ? off0 + it * offstep : 0), ok ? 16u : 0u);
            }
          }
          ptx::cp_async_arrive_noinc(&full[stage]);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == kRowProducerWarps) {
    // ================================ MMA issuer ========================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16<BN>();
      const uint64_t a_desc0 = make_nosw_desc(ptx::smem_u32(s_stage), kRowPlaneBytes, 128);
      const uint64_t b_desc0 = make_nosw_desc(ptx::smem_u32(s_w), BN * 16, 128);
      const uint32_t stage_units = (uint32_t)p.stage_bytes >> 4;
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
        int sl = 0;
        for (int pi = 0; pi < p.nparts; ++pi) {
          const int nsl = p.part[pi].C >> 4;
          const uint32_t* dl = &s_adelta[p.part[pi].mode][py][0][0];
          uint32_t delta[G * 9];
#pragma unroll
          for (int i = 0; i < G * 9; ++i) delta[i] = dl[i];
          for (int j = 0; j < nsl; ++j, ++sl) {
            const uint64_t a_st = a_desc0 + (uint64_t)((uint32_t)stage * stage_units);
            const uint64_t b_sl = b_desc0 + (uint64_t)(sl * 9 * 2 * BN);
            ptx::mbar_wait(&full[stage], phase, p.error_flag, 13);
            ptx::fence_proxy_async();
            ptx::tc_fence_after();
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int g = 0; g < G; ++g) {
                ptx::umma_bf16(tmem_base + (uint32_t)((acc * G + g) * BN), a_st + delta[g * 9 + tap], b_sl + (uint64_t)(tap * 2 * BN), idesc,
                               (uint32_t)((sl | tap) != 0));
              }
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
    // ================================ epilogue (4 warps) ================================
    // Tiles here are only 9..144 small MMAs long, so the per-tile epilogue is on the critical path:
    // folded-BN constants and the fused 1x1 head live in registers for the life of the CTA.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float r_scale[BN], r_bias[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) { r_scale[j] = s_scale[j]; r_bias[j] = s_bias[j]; }
    float r_hw[HEAD ? 64 : 1], r_hb[HEAD ? 4 : 1];
    if (HEAD) {
#pragma unroll
      for (int j = 0; j < 64; ++j) r_hw[HEAD ? j : 0] = s_hw[j];
#pragma unroll
      for (int j = 0; j < 4; ++j) r_hb[HEAD ? j : 0] = s_hb[j];
    }
    const float lo = p.relu ? 0.f : -INFINITY;
    int acc = 0;
    uint32_t acc_phase = 0;
    TileCoord tc;
    tc.init(t_begin, p.tiles_x, p.OH);
    for (int tile = t_begin; tile < t_end; ++tile, tc.next(p.tiles_x, p.OH)) {
      const size_t rowpix = ((size_t)tc.n * p.OH + tc.y) * p.OW;
      ptx::mbar_wait(&tmem_full[acc], acc_phase, p.error_flag, 14);
      ptx::tc_fence_after();
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int x = (G == 2) ? (2 * (tc.xb * 128 + row) + g) : (tc.xb * 128 + row);
        const bool valid = x < p.OW;
        const size_t pix = rowpix + x;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * G + g) * BN);
        uint32_t v[BN];
#pragma unroll
        for (int c = 0; c < BN; c += 16) ptx::tmem_ld16(t_row + (uint32_t)c, *reinterpret_cast<uint32_t(*)[16]>(&v[c]));
        ptx::tmem_ld_wait();
        float yv[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) yv[j] = fmaxf(fmaf(__uint_as_float(v[j]), r_scale[j], r_bias[j]), lo);
        if (HEAD) {
          float4 o;
          float* op = reinterpret_cast<float*>(&o);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // two independent chains per logit (ILP), summed in a fixed order
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              s0 = fmaf(yv[j], r_hw[HEAD ? k * 16 + j : 0], s0);
              s1 = fmaf(yv[j + 1], r_hw[HEAD ? k * 16 + j + 1 : 0], s1);
            }
            op[k] = (s0 + s1) + r_hb[HEAD ? k : 0];
          }
          if (valid) reinterpret_cast<float4*>(p.head_out)[pix] = o;
        }
        if (valid && p.out != nullptr) {
          uint4* op = reinterpret_cast<uint4*>(p.out + pix * BN);
#pragma unroll
          for (int j = 0; j < BN / 8; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(yv[8 * j + 2 * t], yv[8 * j + 2 * t + 1]);
              w[t] = *reinterpret_cast<uint32_t*>(&h2);
            }
            op[j] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[acc]);
      if (++acc == kRowAccStages) { acc = 0; acc_phase ^= 1u; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kRowProducerWarps) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool RowConvOp::eligible(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const void* residual) {
  if (spec.ksize != 3 || spec.stride != 1 || spec.pad != 1 || residual != nullptr) return false;
  if (spec.cout != 16 && spec.cout != 32) return false;
  if (spec.head && spec.cout != 16) return false;
  if (parts.empty() || parts.size() > 2) return false;
  int slabs = 0;
  for (auto& q : parts) {
    if (q.t.C % 16 != 0 || q.t.C > 64 || q.t.C <= 0) return false;
    slabs += q.t.C / 16;
  }
  if (slabs > kRowMaxSlabs) return false;
  if (parts[0].up2) return parts.size() == 1 || !parts[1].up2;
  return parts.size() == 1;
}

void RowConvOp::build(const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const float* w_oihw, const float* scale,
                      const float* bias, void* out, const float* head_w, const float* head_b, float* head_out, int* error_flag,
                      int num_sms) {
  WSI_REQUIRE(eligible(parts, spec, nullptr), WSI_ERR_UNSUPPORTED, "conv is not eligible for the row-tile kernel");
  RowParams& p = p_;
  p = RowParams{};
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
    rp.ptr = static_cast<const bf16*>(q.t.ptr);
    rp.H = q.t.H; rp.W = q.t.W; rp.C = q.t.C;
    rp.mode = up2 ? (q.up2 ? 1 : 2) : 0;
    if (rp.mode == 2) {
      WSI_REQUIRE(q.t.H == OH && q.t.W == OW && q.t.N == N, WSI_ERR_INVALID, "row conv: skip shape mismatch");
      has_skip = true;
    }
    for (int kc0 = 0; kc0 < q.t.C / 8; kc0 += 2) {
      p.slab_part[p.nslabs] = (int8_t)i;
      p.slab_kc0[p.nslabs] = (int16_t)kc0;
      ++p.nslabs;
    }
    cin += q.t.C;
  }
  p.N = N; p.OH = OH; p.OW = OW; p.Cout = BN; p.up2 = up2 ? 1 : 0;
  p.tiles_x = (int)ceil_div(OW, up2 ? 256 : 128);
  const long long total = (long long)N * OH * p.tiles_x;
  WSI_REQUIRE(total < (1LL << 31), WSI_ERR_UNSUPPORTED, "row conv: too many tiles");
  p.total_tiles = (int)total;
  p.relu = spec.relu ? 1 : 0;
  p.out = static_cast<bf16*>(out);
  p.error_flag = error_flag;

  // weights: [slab][tap][2 chunks][BN][8]
  std::vector<uint16_t> wp((size_t)p.nslabs * 9 * 2 * BN * 8);
  int part_off[2] = {0, parts[0].t.C};
  for (int sl = 0; sl < p.nslabs; ++sl)
    for (int tap = 0; tap < 9; ++tap)
      for (int j = 0; j < 2; ++j)
        for (int n = 0; n < BN; ++n)
          for (int e = 0; e < 8; ++e) {
            const int ci = part_off[p.slab_part[sl]] + (p.slab_kc0[sl] + j) * 8 + e;
            const float v = w_oihw[(((size_t)n * cin + ci) * 3 + tap / 3) * 3 + tap % 3];
            wp[((((size_t)sl * 9 + tap) * 2 + j) * BN + n) * 8 + e] = f32_to_bf16_bits(v);
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
  const int w_bytes = p.nslabs * 9 * 2 * BN * 16;
  const int fixed = 128 + ((w_bytes + (2 * BN + 68) * 4 + 127) & ~127) + 512;
  p.stages = std::min(8, (224 * 1024 - fixed) / p.stage_bytes);   // 3 KB left for the static tables
  WSI_REQUIRE(p.stages >= 2, WSI_ERR_UNSUPPORTED, "row conv: not enough shared memory for 2 stages");
  smem_ = fixed + p.stages * p.stage_bytes;
  grid_ = (int)std::min<long long>(total, num_sms);
  CUDA_CHECK(cudaStreamSynchronize(0));
}

template <int BN, bool HEAD, int G>
static void launch_row(const RowParams& p, int grid, int smem, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(conv_rowtile_kernel<BN, HEAD, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    configured = true;
  }
  conv_rowtile_kernel<BN, HEAD, G><<<grid, kRowThreads, smem, s>>>(p);
  CUDA_CHECK(cudaGetLastError());
}

void RowConvOp::launch(cudaStream_t stream, LaunchCounter* lc) const {
  const bool head = p_.head_out != nullptr;
  if (p_.Cout == 16) {
    if (p_.up2) { if (head) launch_row<16, true, 2>(p_, grid_, smem_, stream); else launch_row<16, false, 2>(p_, grid_, smem_, stream); }
    else        { if (head) launch_row<16, true, 1>(p_, grid_, smem_, stream); else launch_row<16, false, 1>(p_, grid_, smem_, stream); }
  } else {
    if (p_.up2) launch_row<32, false, 2>(p_, grid_, smem_, stream);
    else        launch_row<32, false, 1>(p_, grid_, smem_, stream);
  }
  if (lc) lc->n++;
}

}  // namespace wsi
