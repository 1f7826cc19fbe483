// engine.cu — libwsi_b200's context, network plans and the C-ABI entry points (include/wsi_b200.h).
//
// The sliding-window loop of the reference (utils/eval.py:190-228) becomes, per slide / row band:
//   sort tiles by canvas origin -> for each batch of tiles { gather+normalise (K0) -> stem ->
//   max-pool -> tcgen05 implicit-GEMM convs (K2/K3/K5) -> head | fused 1x1 logits } ->
//   atomic-free overlap-accumulate into an fp32 canvas (K6) -> softmax/argmax/heatmap (K7).
// Everything is stream-ordered on the caller's stream; host buffers are copied in/out at the ends.
#include <algorithm>
#include <cmath>
#include <map>
#include <memory>
#include <numeric>

#include "conv_igemm.cuh"
#include "conv_rowtile.cuh"
#include "kernels.cuh"

namespace wsi {

static thread_local std::string g_last_error;
void set_global_error(const std::string& m) { g_last_error = m; }

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

struct Act {  // bf16 activation in HBM: NHWC, or zero-padded channel-chunk-planar around the row-tile kernel
  DevBuf buf;
  int N = 0, H = 0, W = 0, C = 0;
  int layout = LAYOUT_NHWC;
  int planes = 1;          // 3 in the fp32-emulated precision: NHWC pixel = [a | b | c] bf16 planes of C channels each
  TensorView view() const { return TensorView{buf.p, N, H, W, C, layout}; }
  size_t bytes() const {
    if (layout == LAYOUT_NHWC) return (size_t)N * H * W * C * planes * 2;
    return PlanarDims::make(H, W, C, layout).bytes(N);
  }
};

enum StageId { ST_GATHER = 0, ST_STEM, ST_MAXPOOL, ST_CONV, ST_HEAD, ST_STITCH, ST_FINALISE, ST_H2D, ST_D2H, ST_COUNT };
static const char* kStageNames[ST_COUNT] = {"gather", "stem", "maxpool", "conv", "head", "stitch", "finalise", "h2d", "d2h"};

struct StageAcc {
  double ms = 0, work = 0;
  int64_t launches = 0;
};

struct EventSpan {
  cudaEvent_t a, b;
  int stage;
};

struct PinnedBuf {   // page-locked host staging (RAII)
  void* p = nullptr;
  size_t bytes = 0;
  PinnedBuf() = default;
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  void alloc(size_t n) {
    if (n <= bytes && p) return;
    if (p) cudaFreeHost(p);
    p = nullptr; bytes = 0;
    if (n == 0) n = 64;
    cudaError_t e = cudaMallocHost(&p, n);
    if (e != cudaSuccess) { p = nullptr; WSI_THROW(WSI_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", n, cudaGetErrorString(e)); }
    bytes = n;
  }
};

}  // namespace wsi

using namespace wsi;

struct NetPlan;

struct wsi_ctx {
  int device = 0;
  int num_sms = 148;
  std::string err;
  LaunchCounter lc;
  int arch = -1;
  int num_classes = 4;
  std::map<std::string, HostTensor> sd;
  float class_probs[4] = {0.f, 0.f, 0.f, 0.f};
  int64_t batch_tiles = 0;
  int stage_timing = 0;
  int op_trace = 0;                // per-conv CUDA events (wsi_op_stats); set before the plan is built
  int precision = WSI_PRECISION_BF16;
  DevBuf lut;        // f32 [3][256] normalise table
  DevBuf err_flag;   // int, set by a timed-out barrier wait inside the conv kernel
  std::unique_ptr<NetPlan> plan;
  // stage statistics
  StageAcc acc[ST_COUNT];
  std::vector<EventSpan> spans;
  std::vector<cudaEvent_t> event_pool;
  // scratch reused across slides
  // scan_resize != 1: per-axis resize tables (PIL bicubic, wsi_resample_coeffs) and the resized u8 tiles of one batch
  DevBuf rs_hb, rs_hk, rs_vb, rs_vk, rs_tmp, rs_tiles, rs_xy;
  int rs_key[4] = {0, 0, 0, 0}, rs_hks = 0, rs_vks = 0;
  DevBuf raster, maskbuf, logit_ring, classes, heatmap, tiles_dev, rect_tx, rect_rowy, rect_rowstart, rect_rows, rect_bx, rect_by, rect_cellx, rect_celly, cls_cells, tile_logits, scratch_f32, counts;
  // copy streams: chunked raster upload / strip-wise output download overlap the compute on the caller's stream
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  // pinned staging for the per-slide index arrays (tile list, rect index): uploaded without a stream sync
  PinnedBuf idx_host;
  cudaEvent_t idx_event = nullptr;      // completion of the previous call's index upload (guards idx_host reuse)
  bool idx_pending = false;
  std::vector<cudaEvent_t> order_events;   // ordering events between the caller's stream and the copy streams
  size_t order_used = 0;
  // multi-patch ensemble head `fc` (uploaded on first use after a model load)
  DevBuf ens_w1, ens_b1, ens_w2, ens_b2;
  bool ens_ready = false;
};

namespace wsi {

// ---------------------------------------------------------------------------------------------
// stage timing: CUDA events on the launching stream, resolved lazily after a sync
// ---------------------------------------------------------------------------------------------
struct StageScope {
  wsi_ctx* c;
  cudaStream_t s;
  int stage;
  int64_t l0;
  cudaEvent_t a = nullptr, b = nullptr;
  StageScope(wsi_ctx* ctx, cudaStream_t st, int stage_id, double work) : c(ctx), s(st), stage(stage_id), l0(ctx->lc.n) {
    c->acc[stage].work += work;
    if (!c->stage_timing) return;
    a = get_event();
    b = get_event();
    cudaEventRecord(a, s);
  }
  ~StageScope() {
    c->acc[stage].launches += c->lc.n - l0;
    if (!a) return;
    cudaEventRecord(b, s);
    c->spans.push_back(EventSpan{a, b, stage});
  }
  cudaEvent_t get_event() {
    if (!c->event_pool.empty()) {
      cudaEvent_t e = c->event_pool.back();
      c->event_pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};

static void resolve_spans(wsi_ctx* c) {
  for (auto& sp : c->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) c->acc[sp.stage].ms += ms;
    c->event_pool.push_back(sp.a);
    c->event_pool.push_back(sp.b);
  }
  c->spans.clear();
}

// ---------------------------------------------------------------------------------------------
// model helpers
// ---------------------------------------------------------------------------------------------
static const HostTensor& weight(const wsi_ctx* c, const std::string& name) {
  auto it = c->sd.find(name);
  if (it == c->sd.end()) WSI_THROW(WSI_ERR_NOMODEL, "state_dict entry '%s' is missing", name.c_str());
  return it->second;
}
static bool has_weight(const wsi_ctx* c, const std::string& name) { return c->sd.find(name) != c->sd.end(); }

struct Folded {
  std::vector<float> scale, bias;
};
// eval-mode BatchNorm2d (eps 1e-5, resnets_shift.py:37) folded to y = x*scale + bias
static Folded fold_bn(const wsi_ctx* c, const std::string& p, int channels) {
  const HostTensor &w = weight(c, p + ".weight"), &b = weight(c, p + ".bias"), &mu = weight(c, p + ".running_mean"),
                   &var = weight(c, p + ".running_var");
  WSI_REQUIRE(w.numel() == channels && b.numel() == channels && mu.numel() == channels && var.numel() == channels,
              WSI_ERR_NOMODEL, "BatchNorm '%s' has the wrong size (want %d)", p.c_str(), channels);
  Folded f;
  f.scale.resize(channels);
  f.bias.resize(channels);
  for (int i = 0; i < channels; ++i) {
    const float s = w.data[i] / std::sqrt(var.data[i] + 1e-5f);
    f.scale[i] = s;
    f.bias[i] = b.data[i] - mu.data[i] * s;
  }
  return f;
}

static const HostTensor& conv_weight(const wsi_ctx* c, const std::string& name, int cout, int cin, int k) {
  const HostTensor& w = weight(c, name);
  WSI_REQUIRE(w.shape.size() == 4 && w.shape[0] == cout && w.shape[1] == cin && w.shape[2] == k && w.shape[3] == k,
              WSI_ERR_NOMODEL, "conv weight '%s' has the wrong shape (want [%d,%d,%d,%d])", name.c_str(), cout, cin, k, k);
  return w;
}

}  // namespace wsi

// ---------------------------------------------------------------------------------------------
// NetPlan: one batch of `cap` tiles of ph x pw through the network, all buffers and tensor maps
// prepared once.  Trunk: resnets_shift.py:196-204 (== smp ResNetEncoder); decoder: smp Unet.
// ---------------------------------------------------------------------------------------------
struct NetPlan {
  int arch = -1, head = -1, cap = 0, ph = 0, pw = 0;
  int precise = 0;               // WSI_PRECISION_FP32: three bf16 planes per tensor, every conv on the table-driven TMA kernel
  int in_planes() const { return precise ? 3 : 1; }
  int64_t in_plane_stride() const { return (int64_t)cap * (ph + 6) * (pw + 8) * 4; }   // elements between the stem operand's planes
  struct Step {
    int kind;  // 0 conv, 1 maxpool, 2 pool+head
    int stage;
    ConvOp* op = nullptr;
    Act *in = nullptr, *out = nullptr;
  };
  std::vector<std::unique_ptr<Act>> acts;
  std::vector<std::unique_ptr<ConvOp>> ops;
  std::vector<Step> steps;
  DevBuf in_pad;                 // [cap][ph+6][pw+8][4] bf16, zero border
  DevBuf logits;                 // SEG: f32 [cap][ph][pw][4]; CLS/REG/FEATURES: f32 [cap][out_dim]
  DevBuf hw1, hb1, hw2, hb2;     // head weights
  DevBuf pooled;                 // CLS head: the pooled 512-vectors next to the logits (multi-patch ensemble head reads them)
  int n1 = 0, n2 = 0, out_dim = 0;
  Act* x4 = nullptr;
  // SEG: the last decoder conv(s) (fused 1x1 head) and the tile offset of each; their fp32 logits output is re-pointed per batch
  std::vector<std::pair<ConvOp*, int>> head_ops;
  void set_logits_base(float* base) {
    for (auto& h : head_ops) h.first->set_head_out(base + (size_t)h.second * ph * pw * 4);
  }
  double conv_flops = 0, stem_flops = 0;   // per batch of `cap` tiles (algorithmic: 2*MAC, no padding)
  std::vector<Act*> feats;                 // [x4, x3, x2, x1, x0]
  // WSI_CONV_TRACE=1: per-op device time (dev tool; events around every conv launch)
  struct OpStat { std::string desc; double ms = 0, flops = 0; int64_t count = 0; double bytes = 0; std::string kernel; };
  std::vector<OpStat> op_stats;
  struct OpSpan { cudaEvent_t a, b; int idx; };
  std::vector<OpSpan> op_spans;
  bool trace = false;
  void resolve_trace();
  void print_trace();
  ~NetPlan() { if (trace) { resolve_trace(); if (getenv("WSI_CONV_TRACE")) print_trace(); } }

  Act* new_act(int N, int H, int W, int C, int layout = LAYOUT_NHWC) {
    acts.emplace_back(new Act());
    Act* a = acts.back().get();
    a->N = N; a->H = H; a->W = W; a->C = C; a->layout = layout;
    a->planes = precise ? 3 : 1;
    a->buf.alloc(a->bytes());
    if (layout != LAYOUT_NHWC) CUDA_CHECK(cudaMemset(a->buf.p, 0, a->buf.bytes));   // the zero border IS the conv padding
    return a;
  }
  ConvOp* new_op() {
    ops.emplace_back(new ConvOp());
    return ops.back().get();
  }

  void build(wsi_ctx* c, int arch_, int head_, int cap_, int ph_, int pw_, int precise_);
  void run(wsi_ctx* c, cudaStream_t s);
};

void NetPlan::build(wsi_ctx* c, int arch_, int head_, int cap_, int ph_, int pw_, int precise_) {
  arch = arch_; head = head_; cap = cap_; ph = ph_; pw = pw_; precise = precise_ ? 1 : 0;
  WSI_REQUIRE(cap > 0 && ph > 0 && pw > 0, WSI_ERR_INVALID, "plan: empty batch");
  WSI_REQUIRE(ph % 2 == 0 && pw % 2 == 0 && ph >= 8 && pw >= 8, WSI_ERR_UNSUPPORTED, "tile size %dx%d: must be even and >= 8", ph, pw);
  if (head == WSI_HEAD_SEG) {
    WSI_REQUIRE(arch == WSI_ARCH_UNET_R18, WSI_ERR_UNSUPPORTED, "SEG head needs the U-Net model");
    WSI_REQUIRE(ph % 32 == 0 && pw % 32 == 0, WSI_ERR_UNSUPPORTED, "SEG: tile size %dx%d must be a multiple of 32 (5 encoder stages)", ph, pw);
    WSI_REQUIRE(c->num_classes == 4, WSI_ERR_UNSUPPORTED, "SEG: num_classes must be 4 (got %d)", c->num_classes);
  }
  const std::string tp = (arch == WSI_ARCH_UNET_R18) ? "encoder." : "";
  int* ef = c->err_flag.as<int>();
  const int sms = c->num_sms;

  in_pad.alloc((size_t)in_planes() * cap * (ph + 6) * (pw + 8) * 4 * 2);
  CUDA_CHECK(cudaMemset(in_pad.p, 0, in_pad.bytes));

  // ---- stem: conv1 7x7/s2/p3 + bn1 + relu (resnets_shift.py:196-198) ----
  // the row-tile stem writes x0 column-parity-planar: the max-pool reads it directly and the decoder's level-4
  // conv takes it as its skip operand without a relayout
  const int x0_layout = (!precise && ConvOp::stem_routes_to_rowtile()) ? LAYOUT_PLANAR_PARITY : LAYOUT_NHWC;
  Act* x0 = new_act(cap, ph / 2, pw / 2, 64, x0_layout);
  {
    const HostTensor& w = conv_weight(c, tp + "conv1.weight", 64, 3, 7);
    const Folded f = fold_bn(c, tp + "bn1", 64);
    ConvOp* op = new_op();
    op->build_stem(in_pad.p, cap, ph, pw, w.data.data(), f.scale.data(), f.bias.data(), x0->buf.p, ef, sms, x0_layout, precise);
    steps.push_back(Step{0, ST_STEM, op, nullptr, x0});
    stem_flops = op->flops();
    // algorithmic bytes: the padded bf16 operand read once + x0 written once
    op_stats.push_back(OpStat{"stem 7x7/s2 3->64", 0, op->flops(), 0, (double)cap * ph * pw * 8.0 + (double)cap * (ph / 2) * (pw / 2) * 64 * 2.0 * (precise ? 3 : 1),
                              op->kernel_name()});
    trace = getenv("WSI_CONV_TRACE") != nullptr || c->op_trace;
  }
  // ---- maxpool 3x3/s2/p1 (:199) ----
  // layer1's four 64->64 convs run on the row-tile kernel: the pooled tensor, the block-internal tensors and
  // the first block's output stay in the planar layout (also as residuals); layer1's output goes back to NHWC
  // for the TMA kernel (layer2, decoder level-3 skip)
  bool l1_row = false;
  {
    ConvSpec sp; sp.ksize = 3; sp.stride = 1; sp.pad = 1; sp.cout = 64; sp.relu = true;
    const ConvInputPart probe{TensorView{nullptr, cap, (x0->H - 1) / 2 + 1, (x0->W - 1) / 2 + 1, 64, LAYOUT_NHWC}, false};
    l1_row = !precise && (x0_layout == LAYOUT_PLANAR_PARITY) && ConvOp::routes_to_rowtile({probe}, sp, nullptr);
  }
  Act* p0 = new_act(cap, (x0->H - 1) / 2 + 1, (x0->W - 1) / 2 + 1, 64, l1_row ? LAYOUT_PLANAR : LAYOUT_NHWC);
  steps.push_back(Step{1, ST_MAXPOOL, nullptr, x0, p0});

  auto add_conv = [&](const std::vector<ConvInputPart>& parts, const ConvSpec& spec, const float* w, const Folded* f,
                      const Act* res, Act* out, const float* hw = nullptr, const float* hb = nullptr, float* hout = nullptr) {
    ConvOp* op = new_op();
    op->build(parts, spec, w, f ? f->scale.data() : nullptr, f ? f->bias.data() : nullptr, res ? res->buf.p : nullptr,
              out ? out->buf.p : nullptr, hw, hb, hout, ef, sms, out ? out->layout : LAYOUT_NHWC, res ? res->layout : LAYOUT_NHWC, precise);
    steps.push_back(Step{0, ST_CONV, op, nullptr, out});
    conv_flops += op->flops();
    char d[160];
    int cin_t = 0;
    for (auto& q : parts) cin_t += q.t.C;
    snprintf(d, sizeof(d), "conv%dx%d/s%d %4d->%-4d @%dx%d%s%s BN%d BK%d%s", spec.ksize, spec.ksize, spec.stride, cin_t, spec.cout,
             parts[0].up2 ? 2 * parts[0].t.H : parts[0].t.H, parts[0].up2 ? 2 * parts[0].t.W : parts[0].t.W, parts[0].up2 ? " up2" : "",
             res ? " +res" : "", op->block_n(), op->block_k(), op->is_halo() ? " x2 halo" : (op->is_pair() ? " x2" : ""));
    // algorithmic bytes: every operand read once, the output (bf16 tensor or fp32 logits) written once
    double by = 0;
    for (auto& q : parts) by += (double)q.t.N * q.t.H * q.t.W * q.t.C * 2.0;
    if (res) by += (double)res->N * res->H * res->W * res->C * 2.0;
    if (out) by += (double)out->N * out->H * out->W * out->C * 2.0;
    if (precise) by *= 3.0;
    if (hout) by += (double)parts[0].t.N * parts[0].t.H * parts[0].t.W * 16.0;
    op_stats.push_back(OpStat{d, 0, op->flops(), 0, by, op->kernel_name()});
  };

  // ---- layer1..layer4, BasicBlock (resnets_shift.py:49-65) ----
  Act* cur = p0;
  int cin = 64;
  std::vector<Act*> stage_out;
  for (int li = 1; li <= 4; ++li) {
    const int cout = 64 << (li - 1);
    for (int b = 0; b < 2; ++b) {
      const int stride = (li > 1 && b == 0) ? 2 : 1;
      const std::string q = tp + "layer" + std::to_string(li) + "." + std::to_string(b);
      const int bc_in = (b == 0) ? cin : cout;
      const int OH = (cur->H + 2 - 3) / stride + 1, OW = (cur->W + 2 - 3) / stride + 1;
      const bool planar_block = l1_row && li == 1;
      Act* t = new_act(cap, OH, OW, cout, planar_block ? LAYOUT_PLANAR : LAYOUT_NHWC);
      {
        const HostTensor& w = conv_weight(c, q + ".conv1.weight", cout, bc_in, 3);
        const Folded f = fold_bn(c, q + ".bn1", cout);
        ConvSpec sp; sp.ksize = 3; sp.stride = stride; sp.pad = 1; sp.cout = cout; sp.relu = true;
        add_conv({ConvInputPart{cur->view(), false}}, sp, w.data.data(), &f, nullptr, t);
      }
      Act* res = cur;
      if (has_weight(c, q + ".downsample.0.weight")) {
        const HostTensor& w = conv_weight(c, q + ".downsample.0.weight", cout, bc_in, 1);
        const Folded f = fold_bn(c, q + ".downsample.1", cout);
        Act* d = new_act(cap, OH, OW, cout);
        ConvSpec sp; sp.ksize = 1; sp.stride = stride; sp.pad = 0; sp.cout = cout; sp.relu = false;
        add_conv({ConvInputPart{cur->view(), false}}, sp, w.data.data(), &f, nullptr, d);
        res = d;
      } else {
        WSI_REQUIRE(stride == 1 && bc_in == cout, WSI_ERR_NOMODEL, "block %s needs a downsample projection", q.c_str());
      }
      Act* out = new_act(cap, OH, OW, cout, (planar_block && b == 0) ? LAYOUT_PLANAR : LAYOUT_NHWC);
      {
        const HostTensor& w = conv_weight(c, q + ".conv2.weight", cout, cout, 3);
        const Folded f = fold_bn(c, q + ".bn2", cout);
        ConvSpec sp; sp.ksize = 3; sp.stride = 1; sp.pad = 1; sp.cout = cout; sp.relu = true;
        add_conv({ConvInputPart{t->view(), false}}, sp, w.data.data(), &f, res, out);
      }
      cur = out;
    }
    cin = cout;
    stage_out.push_back(cur);
  }
  x4 = stage_out[3];
  feats = {stage_out[3], stage_out[2], stage_out[1], stage_out[0], x0};

  auto upload_f = [&](DevBuf& b, const HostTensor& t) { upload(b, t.data); };

  if (head == WSI_HEAD_SEG) {
    // ---- smp Unet decoder: 5 x {nearest x2, concat skip, 2 x (conv3x3 + BN + ReLU)}, final 1x1 (+bias) ----
    const int outs[5] = {256, 128, 64, 32, 16};
    // The ten decoder convs form a chain; a conv writes the planar layout iff it and its consumer both run
    // on the row-tile kernel (the small-channel, high-resolution end of the decoder).
    struct DConv { std::vector<ConvInputPart> parts; ConvSpec sp; bool row; };
    auto probe_part = [](int N, int H, int W, int C, bool up2) { return ConvInputPart{TensorView{nullptr, N, H, W, C, LAYOUT_NHWC}, up2}; };
    std::vector<DConv> plan;
    {
      int xh = x4->H, xw = x4->W, xc = x4->C;
      for (int i = 1; i <= 5; ++i) {
        Act* skip = (i <= 4) ? feats[i] : nullptr;
        const int co = outs[i - 1];
        if (skip) WSI_REQUIRE(skip->H == 2 * xh && skip->W == 2 * xw, WSI_ERR_UNSUPPORTED, "decoder level %d: skip shape mismatch", i);
        DConv a, b;
        a.parts.push_back(probe_part(cap, xh, xw, xc, true));
        if (skip) a.parts.push_back(probe_part(cap, skip->H, skip->W, skip->C, false));
        a.sp.ksize = 3; a.sp.stride = 1; a.sp.pad = 1; a.sp.cout = co; a.sp.relu = true;
        b.parts.push_back(probe_part(cap, 2 * xh, 2 * xw, co, false));
        b.sp = a.sp; b.sp.head = (i == 5);
        a.row = !precise && ConvOp::routes_to_rowtile(a.parts, a.sp, nullptr);
        b.row = !precise && ConvOp::routes_to_rowtile(b.parts, b.sp, nullptr);
        plan.push_back(a); plan.push_back(b);
        xh *= 2; xw *= 2; xc = co;
      }
    }
    Act* x = x4;
    logits.alloc((size_t)cap * ph * pw * 4 * sizeof(float));
    for (int i = 1; i <= 5; ++i) {
      Act* skip = (i <= 4) ? feats[i] : nullptr;
      const int co = outs[i - 1];
      const int ci = x->C + (skip ? skip->C : 0);
      const std::string q = "decoder.layer" + std::to_string(i) + ".block.";
      const size_t ia = 2 * (size_t)(i - 1), ib = ia + 1;
      const bool a_planar = plan[ib].row;      // consumer is a row kernel; every producer (row kernels, TMA kernel) can write planar
      const bool b_planar = (ib + 1 < plan.size()) && plan[ib].row && plan[ib + 1].row;
      Act* a = new_act(cap, 2 * x->H, 2 * x->W, co, a_planar ? LAYOUT_PLANAR : LAYOUT_NHWC);
      {
        const HostTensor& w = conv_weight(c, q + "0.block.0.weight", co, ci, 3);
        const Folded f = fold_bn(c, q + "0.block.1", co);
        std::vector<ConvInputPart> parts{ConvInputPart{x->view(), true}};
        if (skip) parts.push_back(ConvInputPart{skip->view(), false});
        add_conv(parts, plan[ia].sp, w.data.data(), &f, nullptr, a);
      }
      const HostTensor& w = conv_weight(c, q + "1.block.0.weight", co, co, 3);
      const Folded f = fold_bn(c, q + "1.block.1", co);
      if (i < 5) {
        Act* b = new_act(cap, a->H, a->W, co, b_planar ? LAYOUT_PLANAR : LAYOUT_NHWC);
        add_conv({ConvInputPart{a->view(), false}}, plan[ib].sp, w.data.data(), &f, nullptr, b);
        x = b;
      } else {
        const HostTensor& fw = conv_weight(c, "decoder.final_conv.weight", 4, 16, 1);
        const HostTensor& fb = weight(c, "decoder.final_conv.bias");
        WSI_REQUIRE(fb.numel() == 4, WSI_ERR_NOMODEL, "decoder.final_conv.bias must have 4 entries");
        add_conv({ConvInputPart{a->view(), false}}, plan[ib].sp, w.data.data(), &f, nullptr, nullptr, fw.data.data(), fb.data.data(),
                 logits.as<float>());
        head_ops.push_back({ops.back().get(), 0});
      }
    }
    out_dim = 4;
  } else {
    std::string w1n, b1n, w2n, b2n;
    if (head == WSI_HEAD_CLS) {
      if (arch == WSI_ARCH_RESNET18) { w1n = "fc0.weight"; b1n = "fc0.bias"; }                 // resnets_shift.py:140,208
      else { w1n = "classifier.fc.0.weight"; b1n = "classifier.fc.0.bias"; }                   // models/models.py:27,36
    } else if (head == WSI_HEAD_REG) {
      WSI_REQUIRE(arch == WSI_ARCH_UNET_R18, WSI_ERR_UNSUPPORTED, "REG head needs the U-Net model (Regressor, models/models.py:41-58)");
      w1n = "regressor.fc.0.weight"; b1n = "regressor.fc.0.bias"; w2n = "regressor.fc.2.weight"; b2n = "regressor.fc.2.bias";
    } else {
      WSI_REQUIRE(head == WSI_HEAD_FEATURES, WSI_ERR_INVALID, "unknown head %d", head);
    }
    if (!w1n.empty()) {
      const HostTensor &w1 = weight(c, w1n), &b1 = weight(c, b1n);
      WSI_REQUIRE(w1.shape.size() == 2 && w1.shape[1] == 512 && b1.numel() == w1.shape[0], WSI_ERR_NOMODEL, "head '%s' must be [n,512]", w1n.c_str());
      n1 = (int)w1.shape[0];
      upload_f(hw1, w1); upload_f(hb1, b1);
      out_dim = n1;
      if (!w2n.empty()) {
        const HostTensor &w2 = weight(c, w2n), &b2 = weight(c, b2n);
        WSI_REQUIRE(w2.shape.size() == 2 && w2.shape[1] == n1 && b2.numel() == w2.shape[0], WSI_ERR_NOMODEL, "head '%s' must be [n,%d]", w2n.c_str(), n1);
        n2 = (int)w2.shape[0];
        upload_f(hw2, w2); upload_f(hb2, b2);
        out_dim = n2;
      }
    } else {
      out_dim = 512;
    }
    logits.alloc((size_t)cap * out_dim * sizeof(float));
    if (head == WSI_HEAD_CLS) pooled.alloc((size_t)cap * 512 * sizeof(float));
    steps.push_back(Step{2, ST_HEAD, nullptr, x4, nullptr});
  }
  CUDA_CHECK(cudaDeviceSynchronize());
}

void NetPlan::resolve_trace() {
  for (auto& sp : op_spans) {
    float ms = 0.f;
    if (cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
      op_stats[sp.idx].ms += ms;
      op_stats[sp.idx].count++;
    }
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  op_spans.clear();
}

void NetPlan::print_trace() {
  double tot = 0;
  for (auto& o : op_stats) tot += o.ms;
  fprintf(stderr, "[wsi conv trace] cap=%d tile=%dx%d, %.2f ms total\n", cap, ph, pw, tot);
  for (auto& o : op_stats)
    if (o.count)
      fprintf(stderr, "  %-58s %8.3f ms/launch %7.1f TFLOP/s %5.1f%%\n", o.desc.c_str(), o.ms / o.count,
              o.flops / (o.ms / o.count * 1e-3) / 1e12, 100.0 * o.ms / tot);
}

void NetPlan::run(wsi_ctx* c, cudaStream_t s) {
  if (trace && op_spans.size() > 20000) resolve_trace();
  int conv_idx = 0;
  size_t i = 0;
  while (i < steps.size()) {
    const int stage = steps[i].stage;
    double work = 0;
    size_t j = i;
    while (j < steps.size() && steps[j].stage == stage) {
      if (steps[j].kind == 0) work += steps[j].op->flops();
      else if (steps[j].kind == 1) work += (double)steps[j].in->bytes() + (double)steps[j].out->bytes();
      else work += (double)steps[j].in->bytes();
      ++j;
    }
    StageScope scope(c, s, stage, work);
    for (; i < j; ++i) {
      const Step& st = steps[i];
      if (st.kind == 0) {
        if (trace) {
          OpSpan sp;
          cudaEventCreate(&sp.a);
          cudaEventCreate(&sp.b);
          sp.idx = conv_idx;
          cudaEventRecord(sp.a, s);
          st.op->launch(s, &c->lc);
          cudaEventRecord(sp.b, s);
          op_spans.push_back(sp);
        } else {
          st.op->launch(s, &c->lc);
        }
        ++conv_idx;
      } else if (st.kind == 1) {
        if (precise)
          launch_maxpool_split(st.in->buf.as<bf16>(), st.in->N, st.in->H, st.in->W, st.in->C, st.out->buf.as<bf16>(), s, &c->lc);
        else if (st.in->layout == LAYOUT_PLANAR_PARITY)
          launch_maxpool_planar(st.in->buf.p, st.in->N, st.in->H, st.in->W, st.in->C, st.out->buf.p, st.out->layout, s, &c->lc);
        else
          launch_maxpool(st.in->buf.as<bf16>(), st.in->N, st.in->H, st.in->W, st.in->C, st.out->buf.as<bf16>(), s, &c->lc);
      } else {
        const bool feat = (head == WSI_HEAD_FEATURES);
        float* feat_out = feat ? logits.as<float>() : (pooled.p ? pooled.as<float>() : nullptr);
        launch_pool_head(x4->buf.as<bf16>(), cap, x4->H * x4->W, x4->C, hw1.as<float>(), hb1.as<float>(), n1, hw2.as<float>(),
                         hb2.as<float>(), n2, feat_out, logits.as<float>(), s, &c->lc, precise ? 3 : 1);
      }
    }
  }
}

namespace wsi {

static NetPlan* get_plan(wsi_ctx* c, int head, int n, int ph, int pw) {
  WSI_REQUIRE(c->arch >= 0, WSI_ERR_NOMODEL, "wsi_model_load has not been called");
  NetPlan* p = c->plan.get();
  const int precise = (c->precision == WSI_PRECISION_FP32) ? 1 : 0;
  if (p && p->arch == c->arch && p->head == head && p->ph == ph && p->pw == pw && p->precise == precise && p->cap >= n && p->cap <= std::max(2 * n, 8)) return p;
  c->plan.reset();
  c->plan.reset(new NetPlan());
  c->plan->build(c, c->arch, head, n, ph, pw, precise);
  return c->plan.get();
}

// The device error flag (set by a conv pipeline barrier that timed out) is STICKY: no call clears it on entry, so a
// failure inside an asynchronous (device-output) call is reported by the next synchronising call or by wsi_check.
static void check_device_flag(wsi_ctx* c, cudaStream_t s) {
  int flag = 0;
  CUDA_CHECK(cudaMemcpyAsync(&flag, c->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  if (flag != 0) {
    cudaMemsetAsync(c->err_flag.p, 0, sizeof(int), s);
    WSI_THROW(WSI_ERR_CUDA, "conv kernel pipeline barrier timed out (code %d)", flag);
  }
}

static int64_t auto_batch(const wsi_ctx* c, int ph, int pw, int64_t T) {
  int64_t b = c->batch_tiles;
  if (b <= 0) {
    // Three 512^2 tiles' worth of pixels per SM and batch (444 tiles of 512^2 on 148 SMs; one per SM in the fp32-emulated
    // precision, whose activations are three times as large).  The persistent row kernels cut the batch into one contiguous
    // range per CTA: whole tiles per CTA mean no split units, and an ODD number of them keeps the CTAs' start addresses from
    // being large powers of two apart.  Same-box sweep, 20k x 20k slide (tools/gpu_batch_ab.sh, slide-Mpx/s): 74 -> 352,
    // 148 -> 361, 222 -> 362-367, 296 -> 360, 370 -> 361, 407 -> 360, 444 -> 370.5, 481 -> 361, 518 -> 363, 592 -> 358.
    const int64_t per_sm_px = (c->precision == WSI_PRECISION_FP32 ? 1LL : 3LL) * 512 * 512;
    b = std::max<int64_t>(1, per_sm_px * c->num_sms / ((int64_t)ph * pw));
    // ... rounded to a multiple of num_sms / 4: every layer's output-tile count is batch x (H x W / 128) x N tiles
    // with power-of-two factors, so this makes all of them whole waves over the SMs / SM pairs
    const int64_t q = std::max(1, c->num_sms / 4);
    if (2 * b >= q) b = std::max<int64_t>(1, (b + q / 2) / q) * q;
    b = std::min<int64_t>(b, 1036);
  }
  return std::max<int64_t>(1, std::min<int64_t>(b, T));
}

struct SlideGeom {
  int64_t dx, dy;
};

static void validate_slide(const wsi_slide_desc* sl) {
  WSI_REQUIRE(sl && sl->rgb, WSI_ERR_INVALID, "slide raster is NULL");
  WSI_REQUIRE(sl->ih > 0 && sl->iw > 0 && sl->ph > 0 && sl->pw > 0, WSI_ERR_INVALID, "bad slide geometry");
  WSI_REQUIRE(sl->row_stride >= 3 * sl->iw, WSI_ERR_INVALID, "row_stride %lld < 3*iw", (long long)sl->row_stride);
  WSI_REQUIRE(sl->row0 >= 0 && sl->rows >= 0 && sl->row0 + sl->rows <= sl->ih, WSI_ERR_INVALID, "raster rows [%lld,+%lld) outside the slide",
              (long long)sl->row0, (long long)sl->rows);
  WSI_REQUIRE(sl->resize >= 0 && sl->resize <= 16, WSI_ERR_INVALID, "resize %d outside [0, 16]", sl->resize);
  if (sl->resize > 1)
    WSI_REQUIRE(sl->ph % sl->resize == 0 && sl->pw % sl->resize == 0 && sl->pw <= 16000, WSI_ERR_INVALID,   // one window row + weights' tail in 48 KB of shared memory
                "resize %d: the tile %d x %d must be a multiple of it (ph = tile_h * scan_resize, eval_tumorbed.py:39-40) and at most 16000 wide",
                sl->resize, sl->ph, sl->pw);
}

// scan_resize (myargs.py:115): the network sees (ph / r) x (pw / r) tiles
static inline int slide_resize(const wsi_slide_desc* sl) { return sl->resize > 1 ? sl->resize : 1; }

// ordering events (no timing), reused call after call
static cudaEvent_t order_event(wsi_ctx* c) {
  if (c->order_used == c->order_events.size()) {
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->order_events.push_back(e);
  }
  return c->order_events[c->order_used++];
}

// Raster (band) resident on the device.  A host raster is uploaded in row chunks on the context's copy stream, one
// event per chunk: the caller's stream waits only for the chunk a batch of tiles needs, so the first batches run
// while the rest of the band is still crossing PCIe.  (SlideUpload::need is called with raster-local row bounds.)
struct SlideUpload {
  const uint8_t* rgb = nullptr;
  int64_t stride = 0, chunk_rows = 0;
  std::vector<cudaEvent_t> chunk_ev;
  int waited = -1;
  void need(cudaStream_t s, int64_t row_end) {
    if (chunk_ev.empty() || row_end <= 0) return;
    const int ch = (int)std::min<int64_t>((row_end - 1) / chunk_rows, (int64_t)chunk_ev.size() - 1);
    if (ch > waited) {
      CUDA_CHECK(cudaStreamWaitEvent(s, chunk_ev[ch], 0));
      waited = ch;
    }
  }
};

static SlideUpload stage_raster(wsi_ctx* c, const wsi_slide_desc* sl, int64_t rows, cudaStream_t s) {
  SlideUpload u;
  if (sl->rgb_mem == WSI_MEM_DEVICE) {
    u.stride = sl->row_stride;
    u.rgb = sl->rgb;
    return u;
  }
  const int64_t tight = 3 * sl->iw;
  c->raster.alloc((size_t)rows * tight);
  u.stride = tight;
  u.rgb = c->raster.as<uint8_t>();
  if (rows <= 0) return u;
  u.chunk_rows = std::max<int64_t>(sl->ph, (int64_t)(48 << 20) / tight);
  cudaEvent_t e0 = order_event(c);
  CUDA_CHECK(cudaEventRecord(e0, s));                         // the copy stream starts after everything already queued on s
  CUDA_CHECK(cudaStreamWaitEvent(c->h2d_stream, e0, 0));      // (an earlier call may still be reading c->raster)
  StageScope scope(c, c->h2d_stream, ST_H2D, (double)rows * tight);
  for (int64_t r0 = 0; r0 < rows; r0 += u.chunk_rows) {
    const int64_t nr = std::min(u.chunk_rows, rows - r0);
    CUDA_CHECK(cudaMemcpy2DAsync(c->raster.as<uint8_t>() + r0 * tight, tight, sl->rgb + r0 * sl->row_stride, sl->row_stride, tight, nr,
                                 cudaMemcpyHostToDevice, c->h2d_stream));
    cudaEvent_t e = order_event(c);
    CUDA_CHECK(cudaEventRecord(e, c->h2d_stream));
    u.chunk_ev.push_back(e);
  }
  return u;
}

// whole raster resident before anything else on s (single-batch entry points)
static const uint8_t* stage_raster_all(wsi_ctx* c, const wsi_slide_desc* sl, int64_t rows, int64_t* stride_out, cudaStream_t s) {
  c->order_used = 0;
  SlideUpload u = stage_raster(c, sl, rows, s);
  u.need(s, rows);
  *stride_out = u.stride;
  return u.rgb;
}

static void check_tiles(const wsi_slide_desc* sl, const int32_t* xy, int64_t n, int64_t rows) {
  for (int64_t i = 0; i < n; ++i) {
    const int64_t x = xy[2 * i], y = xy[2 * i + 1];
    WSI_REQUIRE(x >= 0 && x + sl->pw <= sl->iw, WSI_ERR_INVALID, "tile %lld: x=%lld leaves the raster", (long long)i, (long long)x);
    WSI_REQUIRE(y >= sl->row0 && y + sl->ph <= sl->row0 + rows, WSI_ERR_INVALID, "tile %lld: y=%lld outside raster rows [%lld,%lld)",
                (long long)i, (long long)y, (long long)sl->row0, (long long)(sl->row0 + rows));
  }
}

static void ensure_lut(wsi_ctx* c) {
  if (c->lut.p) return;
  // ToTensor: u8 -> f32, div 255; Normalize: sub mean, div std (utils/preprocessing.py:209-212,
  // myargs.py:127-130), evaluated in fp32 in that order — 256 possible values per channel.
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  std::vector<float> t(768);
  for (int ch = 0; ch < 3; ++ch)
    for (int v = 0; v < 256; ++v) {
      volatile float a = (float)v / 255.0f;
      volatile float b = a - mean[ch];
      volatile float d = b / stdv[ch];
      t[ch * 256 + v] = d;
    }
  upload(c->lut, t);
  CUDA_CHECK(cudaDeviceSynchronize());
}

static void ensure_resample(wsi_ctx* c, int ph, int pw, int th, int tw) {
  if (c->rs_key[0] == ph && c->rs_key[1] == pw && c->rs_key[2] == th && c->rs_key[3] == tw) return;
  // weight rows padded with zeros to a multiple of 4 taps: the kernels fetch them with 16-byte loads
  auto table = [](int in_size, int out_size, std::vector<int32_t>& b, std::vector<int32_t>& k) {
    const int ks = wsi_resample_ksize(in_size, out_size), ksp = (ks + 3) & ~3;
    std::vector<int32_t> raw((size_t)out_size * ks);
    b.assign((size_t)out_size * 2, 0);
    WSI_REQUIRE(wsi_resample_coeffs(in_size, out_size, b.data(), raw.data()) == WSI_OK, WSI_ERR_INVALID, "resize tables");
    k.assign((size_t)out_size * ksp, 0);
    for (int i = 0; i < out_size; ++i) memcpy(&k[(size_t)i * ksp], &raw[(size_t)i * ks], (size_t)ks * sizeof(int32_t));
    return ksp;
  };
  std::vector<int32_t> hb, hk, vb, vk;
  const int hks = table(pw, tw, hb, hk), vks = table(ph, th, vb, vk);
  CUDA_CHECK(cudaDeviceSynchronize());                       // an earlier call may still read the old tables
  upload(c->rs_hb, hb); upload(c->rs_hk, hk); upload(c->rs_vb, vb); upload(c->rs_vk, vk);
  CUDA_CHECK(cudaDeviceSynchronize());
  c->rs_hks = hks; c->rs_vks = vks;
  c->rs_key[0] = ph; c->rs_key[1] = pw; c->rs_key[2] = th; c->rs_key[3] = tw;
}

// K0 (+ K0r): n tiles of the raster -> the stem operand (and / or the fp32 normalised tiles).  With scan_resize r > 1 the
// pw x ph windows are first resized to (pw / r) x (ph / r) u8 tiles exactly as PIL does it (utils/dataset.py:180-181).
static void gather_batch(wsi_ctx* c, const wsi_slide_desc* sl, const uint8_t* rgb, int64_t rstride, const int32_t* xy_dev, int n, bf16* padded,
                         float* norm_out, int planes, int64_t plane_stride, cudaStream_t s) {
  const int r = slide_resize(sl);
  if (r == 1) {
    launch_gather(rgb, rstride, sl->row0, xy_dev, n, sl->ph, sl->pw, c->lut.as<float>(), padded, norm_out, s, &c->lc, planes, plane_stride);
    return;
  }
  const int th = sl->ph / r, tw = sl->pw / r;
  ensure_resample(c, sl->ph, sl->pw, th, tw);
  c->rs_tmp.alloc((size_t)n * sl->ph * tw * 3);
  c->rs_tiles.alloc((size_t)n * th * tw * 3);
  c->rs_xy.alloc((size_t)n * 2 * sizeof(int32_t));
  launch_resample_tiles(rgb, rstride, sl->row0, xy_dev, n, sl->ph, sl->pw, th, tw, c->rs_hb.as<int32_t>(), c->rs_hk.as<int32_t>(), c->rs_hks,
                        c->rs_vb.as<int32_t>(), c->rs_vk.as<int32_t>(), c->rs_vks, c->rs_tmp.as<uint8_t>(), c->rs_tiles.as<uint8_t>(),
                        c->rs_xy.as<int32_t>(), s, &c->lc);
  launch_gather(c->rs_tiles.as<uint8_t>(), (int64_t)tw * 3, 0, c->rs_xy.as<int32_t>(), n, th, tw, c->lut.as<float>(), padded, norm_out, s, &c->lc,
                planes, plane_stride);
}

// the hot path ------------------------------------------------------------------------------
// Per slide / row band: sort the tiles by canvas origin; batches of tiles go gather -> network; SEG logits land in a
// ring of the last few tile rows; whenever a run of canvas rows can receive no further tile (every later tile starts
// below it) ONE kernel sums the covering tiles of each of its pixels and finalises them (K6 + K7 fused, no canvas).
static void run_slide(wsi_ctx* c, const wsi_slide_desc* sl, const int32_t* tiles_xy, int64_t T, int head, const wsi_out_desc* out,
                      cudaStream_t s) {
  validate_slide(sl);
  WSI_REQUIRE(out && out->classes && out->heatmap, WSI_ERR_INVALID, "classes and heatmap outputs are required");
  WSI_REQUIRE(head == WSI_HEAD_SEG || head == WSI_HEAD_CLS, WSI_ERR_INVALID, "wsi_run_slide: head must be SEG or CLS");
  WSI_REQUIRE(T >= 0 && T < (1LL << 30) && (T == 0 || tiles_xy), WSI_ERR_INVALID, "bad tile list");
  WSI_REQUIRE(sl->m > 0, WSI_ERR_INVALID, "m must be positive");
  const int64_t rows_r = (sl->rows == 0 && sl->row0 == 0) ? sl->ih : sl->rows;
  const int64_t H2 = sl->H2, W2 = sl->W2;
  WSI_REQUIRE(H2 > 0 && W2 > 0 && W2 < (1LL << 31) - 512 && H2 < (1LL << 31), WSI_ERR_INVALID, "bad canvas size");
  const int64_t own0 = sl->own0, own1 = (sl->own0 == 0 && sl->own1 == 0) ? H2 : sl->own1;
  WSI_REQUIRE(own0 >= 0 && own1 >= own0 && own1 <= H2, WSI_ERR_INVALID, "bad owned rows");
  const int64_t orows = own1 - own0, plane = orows * W2;
  const int ph = sl->ph, pw = sl->pw;
  const int rs = slide_resize(sl), th = ph / rs, tw = pw / rs;                            // th x tw: what the network sees
  const int64_t dx = (int64_t)(sl->m * (double)pw), dy = (int64_t)(sl->m * (double)ph);   // utils/eval.py:186
  if (head == WSI_HEAD_SEG)
    WSI_REQUIRE(dx == pw && dy == ph, WSI_ERR_UNSUPPORTED, "SEG needs m == 1 (the reference's slice-add requires equal shapes, utils/eval.py:213-215)");
  WSI_REQUIRE(dx > 0 && dy > 0, WSI_ERR_DEGENERATE, "tile rectangle is empty on the canvas");
  check_tiles(sl, tiles_xy, T, rows_r);
  ensure_lut(c);
  c->order_used = 0;

  // ---- sort tiles by canvas origin (ty, tx): batches become spatially compact and every pixel sums its tiles in
  //      one fixed order, whatever order the caller (or a shuffling DataLoader, utils/dataset.py:192) presents ----
  std::vector<int32_t> order((size_t)T), tys((size_t)T), txs((size_t)T);
  for (int64_t i = 0; i < T; ++i) {
    order[i] = (int32_t)i;
    txs[i] = (int32_t)(sl->m * (double)tiles_xy[2 * i]);        // int(m * x), utils/eval.py:214
    tys[i] = (int32_t)(sl->m * (double)tiles_xy[2 * i + 1]);
  }
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
    if (tys[a] != tys[b]) return tys[a] < tys[b];
    return txs[a] < txs[b];
  });
  std::vector<int32_t> sty((size_t)T), rowy, rowstart;
  for (int64_t i = 0; i < T; ++i) {
    sty[i] = tys[order[i]];
    if (i == 0 || sty[i] != rowy.back()) {
      rowy.push_back(sty[i]);
      rowstart.push_back((int32_t)i);
    }
  }
  rowstart.push_back((int32_t)T);
  int cls_nbx = 0, cls_nby = 0;
  // index arrays -> pinned staging -> device, stream-ordered (no host sync)
  {
    if (c->idx_pending) { CUDA_CHECK(cudaEventSynchronize(c->idx_event)); c->idx_pending = false; }
    const size_t n_xy = (size_t)2 * T, n_tx = (size_t)T, n_ry = rowy.size(), n_rs = rowstart.size(), n_ri = rowy.size() * 8;
    // CLS: the canvas cut at every tile edge (cells of constant summed logits) and the cell index of every column / owned row
    std::vector<int32_t> bx, by, cellx, celly;
    if (head == WSI_HEAD_CLS) {
      bx.push_back(0);
      for (int64_t i = 0; i < T; ++i) {
        bx.push_back((int32_t)std::min<int64_t>(std::max<int64_t>(txs[i], 0), W2));
        bx.push_back((int32_t)std::min<int64_t>(std::max<int64_t>((int64_t)txs[i] + dx, 0), W2));
      }
      std::sort(bx.begin(), bx.end());
      bx.erase(std::unique(bx.begin(), bx.end()), bx.end());
      while (!bx.empty() && bx.back() >= W2) bx.pop_back();
      by.push_back((int32_t)own0);
      for (int32_t ry : rowy) {
        by.push_back((int32_t)std::min<int64_t>(std::max<int64_t>(ry, own0), own1));
        by.push_back((int32_t)std::min<int64_t>(std::max<int64_t>((int64_t)ry + dy, own0), own1));
      }
      std::sort(by.begin(), by.end());
      by.erase(std::unique(by.begin(), by.end()), by.end());
      while (!by.empty() && by.back() >= own1) by.pop_back();
      cellx.resize((size_t)W2);
      for (size_t k = 0; k < bx.size(); ++k) {
        const int64_t x1 = (k + 1 < bx.size()) ? bx[k + 1] : W2;
        for (int64_t x = bx[k]; x < x1; ++x) cellx[(size_t)x] = (int32_t)k;
      }
      celly.resize((size_t)orows);
      for (size_t k = 0; k < by.size(); ++k) {
        const int64_t y1 = (k + 1 < by.size()) ? by[k + 1] : own1;
        for (int64_t y = by[k]; y < y1; ++y) celly[(size_t)(y - own0)] = (int32_t)k;
      }
    }
    const size_t n_cl = bx.size() + by.size() + cellx.size() + celly.size();
    c->idx_host.alloc((n_xy + n_tx + n_ry + n_rs + n_ri + n_cl + 8) * sizeof(int32_t));
    int32_t* h = static_cast<int32_t*>(c->idx_host.p);
    int32_t *h_xy = h, *h_tx = h + n_xy, *h_ry = h_tx + n_tx, *h_rs = h_ry + n_ry, *h_ri = h_rs + n_rs;
    h_ri += (4 - ((h_ri - h) & 3)) & 3;                       // RowInfo entries are read with 16-byte loads
    for (int64_t i = 0; i < T; ++i) {
      const int32_t o = order[i];
      h_xy[2 * i] = tiles_xy[2 * o];
      h_xy[2 * i + 1] = tiles_xy[2 * o + 1];
      h_tx[i] = txs[o];
    }
    if (n_ry) memcpy(h_ry, rowy.data(), n_ry * sizeof(int32_t));
    memcpy(h_rs, rowstart.data(), n_rs * sizeof(int32_t));
    static_assert(sizeof(RowInfo) == 32, "RowInfo is 8 ints");
    for (size_t r = 0; r < n_ry; ++r) {                       // per tile row: its rects and the regular-grid prefix of their x origins
      RowInfo q{};
      q.ry = rowy[r]; q.a = rowstart[r]; q.b = rowstart[r + 1];
      q.tx0 = h_tx[q.a]; q.step = 0; q.nreg = 1;
      if (q.b - q.a >= 2 && h_tx[q.a + 1] > h_tx[q.a]) {
        q.step = h_tx[q.a + 1] - h_tx[q.a];
        q.nreg = 2;
        while (q.a + q.nreg < q.b && h_tx[q.a + q.nreg] - h_tx[q.a + q.nreg - 1] == q.step) ++q.nreg;
      }
      memcpy(h_ri + 8 * r, &q, sizeof(q));
    }
    auto up = [&](DevBuf& b, const int32_t* src, size_t n) {
      b.alloc(std::max<size_t>(n, 1) * sizeof(int32_t));
      if (n) CUDA_CHECK(cudaMemcpyAsync(b.p, src, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    };
    up(c->tiles_dev, h_xy, n_xy);
    up(c->rect_tx, h_tx, n_tx);
    up(c->rect_rowy, h_ry, n_ry);
    up(c->rect_rowstart, h_rs, n_rs);
    up(c->rect_rows, h_ri, n_ri);
    if (head == WSI_HEAD_CLS) {
      int32_t* h_cl = h_ri + n_ri;
      auto put = [&](DevBuf& b, const std::vector<int32_t>& v) {
        if (!v.empty()) memcpy(h_cl, v.data(), v.size() * sizeof(int32_t));
        up(b, h_cl, v.size());
        h_cl += v.size();
      };
      put(c->rect_bx, bx); put(c->rect_by, by); put(c->rect_cellx, cellx); put(c->rect_celly, celly);
      cls_nbx = (int)bx.size();
      cls_nby = (int)by.size();
    }
    CUDA_CHECK(cudaEventRecord(c->idx_event, s));
    c->idx_pending = true;
  }
  RectIndex ri;
  ri.tx = c->rect_tx.as<int32_t>();
  ri.row_y = c->rect_rowy.as<int32_t>();
  ri.row_start = c->rect_rowstart.as<int32_t>();
  ri.rows = c->rect_rows.as<RowInfo>();
  ri.R = (int32_t)rowy.size();
  ri.dx = (int32_t)dx;
  ri.dy = (int32_t)dy;
  ri.up = (head == WSI_HEAD_SEG) ? rs : 1;

  SlideUpload up_r;
  if (T > 0) up_r = stage_raster(c, sl, rows_r, s);

  const uint8_t* mask_dev = nullptr;
  if (sl->mask) {
    if (sl->mask_mem == WSI_MEM_DEVICE) {
      mask_dev = sl->mask;
    } else {
      StageScope scope(c, s, ST_H2D, (double)plane);
      c->maskbuf.alloc((size_t)plane);
      CUDA_CHECK(cudaMemcpyAsync(c->maskbuf.p, sl->mask, (size_t)plane, cudaMemcpyHostToDevice, s));
      mask_dev = c->maskbuf.as<uint8_t>();
    }
  }

  const bool host_out = (out->mem == WSI_MEM_HOST);
  uint8_t *cls_dev = out->classes, *heat_dev = out->heatmap;
  float *canvas_out = out->canvas, *probs_out = out->probs;
  int32_t* counts_out = out->counts;
  if (host_out) {
    c->classes.alloc((size_t)plane);
    c->heatmap.alloc((size_t)plane);
    cls_dev = c->classes.as<uint8_t>();
    heat_dev = c->heatmap.as<uint8_t>();
    const size_t want = ((out->canvas ? 1 : 0) + (out->probs ? 1 : 0)) * (size_t)plane * 4;
    if (want) c->scratch_f32.alloc(want * sizeof(float));
    float* p = c->scratch_f32.as<float>();
    if (out->canvas) { canvas_out = p; p += plane * 4; }
    if (out->probs) probs_out = p;
    if (out->counts) { c->counts.alloc((size_t)plane * sizeof(int32_t)); counts_out = c->counts.as<int32_t>(); }
  }

  FinaliseArgs fa;
  fa.W2 = W2; fa.own0 = own0; fa.own1 = own1; fa.mask = mask_dev;
  for (int i = 0; i < 4; ++i) fa.class_probs[i] = (double)c->class_probs[i];
  fa.heat_mode = (head == WSI_HEAD_CLS) ? 1 : 0;
  fa.classes = cls_dev; fa.heatmap = heat_dev; fa.canvas_out = canvas_out; fa.probs_out = probs_out;

  const int64_t B = auto_batch(c, th, tw, std::max<int64_t>(T, 1));
  NetPlan* plan = (T > 0) ? get_plan(c, head, (int)B, th, tw) : nullptr;
  const int cap = plan ? plan->cap : 1;
  const double tile_px = (double)th * tw;
  const int64_t tile_elems = (int64_t)th * tw * 4;

  // tile rows [r_lo, r_hi) that can touch canvas rows [ya, yb)
  auto row_range = [&](int64_t ya, int64_t yb, int* r_lo, int* r_hi) {
    *r_lo = (int)(std::upper_bound(rowy.begin(), rowy.end(), (int32_t)std::max<int64_t>(ya - dy, INT32_MIN)) - rowy.begin());
    *r_hi = (int)(std::upper_bound(rowy.begin(), rowy.end(), (int32_t)std::min<int64_t>(yb - 1, INT32_MAX)) - rowy.begin());
  };
  // first sorted tile that can still touch canvas row y: ty + dy > y
  auto first_needed = [&](int64_t y) { return (int64_t)(std::upper_bound(sty.begin(), sty.end(), (int32_t)std::max<int64_t>(y - dy, INT32_MIN)) - sty.begin()); };
  // rows below which no unprocessed tile starts once the first e sorted tiles are done
  auto ready_after = [&](int64_t e) { return (e >= T) ? own1 : std::min<int64_t>(std::max<int64_t>(sty[e], own0), own1); };

  int64_t ring_cap = 0;
  if (head == WSI_HEAD_SEG && T > 0) {
    // size the ring: while batch [t0, e) is written, every earlier tile that a not-yet-final row needs must survive
    int64_t y_done = own0, max_span = 1;
    for (int64_t t0 = 0; t0 < T; t0 += cap) {
      const int64_t e = std::min<int64_t>(T, t0 + cap);
      max_span = std::max(max_span, e - std::min(first_needed(y_done), t0));
      y_done = std::max(y_done, ready_after(e));
    }
    ring_cap = std::min<int64_t>(round_up(max_span, cap), round_up(T, cap));
    c->logit_ring.alloc((size_t)ring_cap * tile_elems * sizeof(float));
  } else if (head == WSI_HEAD_CLS) {
    c->tile_logits.alloc((size_t)std::max<int64_t>(T, 1) * 4 * sizeof(float));
    WSI_REQUIRE(!plan || plan->out_dim == 4, WSI_ERR_UNSUPPORTED, "CLS stitch needs 4 classes (got %d)", plan ? plan->out_dim : 0);
  }

  // finished strips leave for the host on the download stream while later batches compute
  cudaEvent_t last_d2h = nullptr;
  auto download_rows = [&](int64_t y0, int64_t y1) {
    if (!host_out || y1 <= y0) return;
    cudaEvent_t e = order_event(c);
    CUDA_CHECK(cudaEventRecord(e, s));
    CUDA_CHECK(cudaStreamWaitEvent(c->d2h_stream, e, 0));
    const size_t off = (size_t)(y0 - own0) * W2, n = (size_t)(y1 - y0) * W2;
    StageScope scope(c, c->d2h_stream, ST_D2H, (double)n * 2.0);
    CUDA_CHECK(cudaMemcpyAsync(out->classes + off, cls_dev + off, n, cudaMemcpyDeviceToHost, c->d2h_stream));
    CUDA_CHECK(cudaMemcpyAsync(out->heatmap + off, heat_dev + off, n, cudaMemcpyDeviceToHost, c->d2h_stream));
    last_d2h = order_event(c);
    CUDA_CHECK(cudaEventRecord(last_d2h, c->d2h_stream));
  };

  int64_t y_done = own0;
  double pending_logit_bytes = 0;
  for (int64_t t0 = 0; t0 < T; t0 += cap) {
    const int n = (int)std::min<int64_t>(cap, T - t0);
    {
      int64_t row_end = 0;
      for (int64_t i = t0; i < t0 + n; ++i) row_end = std::max<int64_t>(row_end, (int64_t)tiles_xy[2 * order[i] + 1] + ph - sl->row0);
      up_r.need(s, row_end);
      // algorithmic bytes: 3 read + 8 (or 24) written per network pixel; resize: + the window read, the half-resized tile
      // written and read, the resized tile written and read
      StageScope scope(c, s, ST_GATHER, 9.0 * n * tile_px + (rs > 1 ? 3.0 * n * ((double)ph * pw + 2.0 * ph * tw + tile_px) : 0.0));
      gather_batch(c, sl, up_r.rgb, up_r.stride, c->tiles_dev.as<int32_t>() + 2 * t0, n, plan->in_pad.as<bf16>(), nullptr, plan->in_planes(),
                   plan->in_plane_stride(), s);
    }
    if (head == WSI_HEAD_SEG) {
      plan->set_logits_base(c->logit_ring.as<float>() + (t0 % ring_cap) * tile_elems);
      plan->run(c, s);
      pending_logit_bytes += (double)n * tile_px * 16.0;
      const int64_t e = t0 + n, y_ready = ready_after(e);
      if (y_ready > y_done) {
        // algorithmic bytes (fused stitch + finalise): T*P*C*4 logits read once + S*2 written (+ S mask read)
        StageScope scope(c, s, ST_STITCH, pending_logit_bytes + (double)(y_ready - y_done) * W2 * (mask_dev ? 3.0 : 2.0));
        pending_logit_bytes = 0;
        int r_lo, r_hi;
        row_range(y_done, y_ready, &r_lo, &r_hi);
        launch_stitch_finalise_seg(ri, c->logit_ring.as<float4>(), (int)ring_cap, (int)first_needed(y_done), (int)e, r_lo, r_hi, y_done, y_ready, fa, s, &c->lc);
      }
      if (y_ready > y_done) { download_rows(y_done, y_ready); y_done = y_ready; }
    } else {
      plan->run(c, s);
      CUDA_CHECK(cudaMemcpyAsync(c->tile_logits.as<float>() + 4 * t0, plan->logits.p, (size_t)n * 4 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
  }

  if (head == WSI_HEAD_SEG) {
    if (y_done < own1) {                       // T == 0 (or nothing owned was reached): uncovered rows
      StageScope scope(c, s, ST_STITCH, (double)(own1 - y_done) * W2 * (mask_dev ? 3.0 : 2.0));
      launch_stitch_finalise_seg(ri, c->logit_ring.as<float4>(), (int)std::max<int64_t>(ring_cap, 1), 0, 0, 0, 0, y_done, own1, fa, s, &c->lc);
      download_rows(y_done, own1);
    }
  } else {
    {
      StageScope scope(c, s, ST_STITCH, (double)T * 16.0 + (double)plane * (mask_dev ? 3.0 : 2.0));
      c->cls_cells.alloc((size_t)std::max(cls_nbx, 1) * std::max(cls_nby, 1) * cls_cell_bytes());
      launch_stitch_finalise_cls(ri, c->tile_logits.as<float4>(), (int)T, c->rect_bx.as<int32_t>(), cls_nbx, c->rect_by.as<int32_t>(), cls_nby,
                                 c->rect_cellx.as<int32_t>(), c->rect_celly.as<int32_t>(), c->cls_cells.p, fa, s, &c->lc);
    }
    download_rows(own0, own1);
  }
  if (counts_out) launch_counts(ri, (int)T, W2, own0, own1, counts_out, s, &c->lc);
  if (last_d2h) CUDA_CHECK(cudaStreamWaitEvent(s, last_d2h, 0));     // the call's outputs are ordered on the caller's stream
  up_r.need(s, rows_r);                                               // ... and so is the last read of the caller's raster

  if (host_out) {
    StageScope scope(c, s, ST_D2H, (double)plane * 4.0 * ((out->canvas ? 4 : 0) + (out->probs ? 4 : 0) + (out->counts ? 1 : 0)));
    if (out->canvas) CUDA_CHECK(cudaMemcpyAsync(out->canvas, canvas_out, (size_t)plane * 4 * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (out->probs) CUDA_CHECK(cudaMemcpyAsync(out->probs, probs_out, (size_t)plane * 4 * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (out->counts) CUDA_CHECK(cudaMemcpyAsync(out->counts, counts_out, (size_t)plane * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  }
  if (out->tile_logits) {
    WSI_REQUIRE(head == WSI_HEAD_CLS, WSI_ERR_UNSUPPORTED, "tile_logits is a CLS-only output");
    std::vector<float> sorted((size_t)T * 4), orig((size_t)T * 4);
    CUDA_CHECK(cudaMemcpyAsync(sorted.data(), c->tile_logits.p, (size_t)T * 4 * sizeof(float), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    for (int64_t i = 0; i < T; ++i) memcpy(&orig[(size_t)order[i] * 4], &sorted[(size_t)i * 4], 4 * sizeof(float));
    CUDA_CHECK(cudaMemcpy(out->tile_logits, orig.data(), (size_t)T * 4 * sizeof(float), host_out ? cudaMemcpyHostToHost : cudaMemcpyHostToDevice));
  }
  // Host outputs: the call returns when they are complete (and reports a conv pipeline failure).  Device outputs:
  // fully asynchronous; the sticky error flag is reported by the next synchronising call or by wsi_check.
  if (host_out || c->stage_timing) {
    check_device_flag(c, s);
    resolve_spans(c);
  }
}

// one batch through the network: x f32 NCHW (normalised) or tiles cut from a raster
static void forward_common(wsi_ctx* c, NetPlan* plan, int n, int head, float* out, int mem, cudaStream_t s) {
  plan->set_logits_base(plan->logits.as<float>());                            // wsi_run_slide points them into its logit ring
  plan->run(c, s);
  const int ph = plan->ph, pw = plan->pw;
  const size_t out_elems = (head == WSI_HEAD_SEG) ? (size_t)n * 4 * ph * pw : (size_t)n * plan->out_dim;
  float* dst = out;
  if (mem == WSI_MEM_HOST) {
    c->scratch_f32.alloc(out_elems * sizeof(float));
    dst = c->scratch_f32.as<float>();
  }
  if (head == WSI_HEAD_SEG) {
    StageScope scope(c, s, ST_HEAD, (double)out_elems * 8.0);
    launch_nhwc4_to_nchw(plan->logits.as<float>(), n, ph, pw, dst, s, &c->lc);
  } else {
    CUDA_CHECK(cudaMemcpyAsync(dst, plan->logits.p, out_elems * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  if (mem == WSI_MEM_HOST) CUDA_CHECK(cudaMemcpyAsync(out, dst, out_elems * sizeof(float), cudaMemcpyDeviceToHost, s));
  check_device_flag(c, s);
  resolve_spans(c);
}

}  // namespace wsi

// =============================================================================================
// C-ABI
// =============================================================================================
#define WSI_API_BEGIN try {
#define WSI_API_END(ctxp)                                  \
  }                                                        \
  catch (const ::wsi::Error& e) {                          \
    if (ctxp) (ctxp)->err = e.msg;                         \
    ::wsi::set_global_error(e.msg);                        \
    return e.status;                                       \
  }                                                        \
  catch (const std::bad_alloc&) {                          \
    ::wsi::set_global_error("out of host memory");         \
    return WSI_ERR_NOMEM;                                  \
  }                                                        \
  catch (const std::exception& e) {                        \
    if (ctxp) (ctxp)->err = e.what();                      \
    ::wsi::set_global_error(e.what());                     \
    return WSI_ERR_INVALID;                                \
  }                                                        \
  return WSI_OK;

extern "C" {

const char* wsi_version(void) { return "wsi_b200 0.1 (sm_100a)"; }

int wsi_ctx_create(int device, wsi_ctx** out) {
  wsi_ctx* c = nullptr;
  WSI_API_BEGIN
  WSI_REQUIRE(out, WSI_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  WSI_REQUIRE(e == cudaSuccess && ndev > 0, WSI_ERR_CUDA, "no CUDA device (%s); libwsi_b200 has no CPU fallback", cudaGetErrorString(e));
  WSI_REQUIRE(device >= 0 && device < ndev, WSI_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
  CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  WSI_REQUIRE(prop.major == 10, WSI_ERR_CUDA, "device %d is sm_%d%d; libwsi_b200 is built for sm_100a only", device, prop.major, prop.minor);
  std::unique_ptr<wsi_ctx> ctx(new wsi_ctx());
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->err_flag.alloc(sizeof(int));
  CUDA_CHECK(cudaMemset(ctx->err_flag.p, 0, sizeof(int)));
  CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&ctx->idx_event, cudaEventDisableTiming));
  init_tensor_map_api();
  *out = ctx.release();
  WSI_API_END(c)
}

int wsi_ctx_destroy(wsi_ctx* ctx) {
  if (!ctx) return WSI_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (auto& sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (auto e : ctx->event_pool) cudaEventDestroy(e);
  for (auto e : ctx->order_events) cudaEventDestroy(e);
  if (ctx->idx_event) cudaEventDestroy(ctx->idx_event);
  if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  delete ctx;
  return WSI_OK;
}

const char* wsi_last_error(wsi_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int wsi_set_option(wsi_ctx* ctx, const char* key, int64_t value) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && key, WSI_ERR_INVALID, "NULL argument");
  const std::string k(key);
  if (k == "batch_tiles") { WSI_REQUIRE(value >= 0 && value <= 4096, WSI_ERR_INVALID, "batch_tiles out of range"); ctx->batch_tiles = value; }
  else if (k == "stage_timing") ctx->stage_timing = value ? 1 : 0;
  else if (k == "op_trace") { ctx->op_trace = value ? 1 : 0; ctx->plan.reset(); }
  else if (k == "precision") {
    WSI_REQUIRE(value == WSI_PRECISION_BF16 || value == WSI_PRECISION_FP32, WSI_ERR_INVALID, "precision must be WSI_PRECISION_BF16 (0) or WSI_PRECISION_FP32 (1)");
    ctx->precision = (int)value;
  }
  else WSI_THROW(WSI_ERR_INVALID, "unknown option '%s'", key);
  WSI_API_END(ctx)
}

int wsi_set_class_probs(wsi_ctx* ctx, const float* p, int n) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && p && n == 4, WSI_ERR_INVALID, "class_probs must have 4 entries");
  for (int i = 0; i < 4; ++i) ctx->class_probs[i] = p[i];
  WSI_API_END(ctx)
}

int64_t wsi_kernel_launches(wsi_ctx* ctx) { return ctx ? ctx->lc.n : 0; }

int wsi_model_load(wsi_ctx* ctx, int arch, const wsi_tensor_desc* tensors, int n, int num_classes) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && tensors && n > 0, WSI_ERR_INVALID, "NULL argument");
  WSI_REQUIRE(arch == WSI_ARCH_RESNET18 || arch == WSI_ARCH_UNET_R18, WSI_ERR_INVALID, "unknown arch %d", arch);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ctx->plan.reset();
  ctx->ens_ready = false;
  ctx->sd.clear();
  ctx->arch = -1;
  for (int i = 0; i < n; ++i) {
    const wsi_tensor_desc& t = tensors[i];
    WSI_REQUIRE(t.name && t.data && t.ndim >= 0 && t.ndim <= 4, WSI_ERR_INVALID, "tensor %d is malformed", i);
    HostTensor h;
    int64_t numel = 1;
    for (int d = 0; d < t.ndim; ++d) { WSI_REQUIRE(t.shape[d] >= 0, WSI_ERR_INVALID, "tensor '%s': negative dim", t.name); h.shape.push_back(t.shape[d]); numel *= t.shape[d]; }
    h.data.assign(t.data, t.data + numel);
    ctx->sd[t.name] = std::move(h);
  }
  ctx->num_classes = num_classes;
  ctx->arch = arch;
  // fail early on an incomplete trunk
  (void)weight(ctx, (arch == WSI_ARCH_UNET_R18 ? std::string("encoder.") : std::string("")) + "conv1.weight");
  WSI_API_END(ctx)
}

int wsi_run_slide(wsi_ctx* ctx, const wsi_slide_desc* slide, const int32_t* tiles_xy, int64_t n_tiles, int head,
                  const wsi_out_desc* out, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx, WSI_ERR_INVALID, "ctx is NULL");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  run_slide(ctx, slide, tiles_xy, n_tiles, head, out, (cudaStream_t)stream);
  WSI_API_END(ctx)
}

int wsi_forward_batch(wsi_ctx* ctx, const float* x, int64_t n, int32_t h, int32_t w, int head, float* out, int mem, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && x && out && n > 0 && n < (1 << 20), WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  NetPlan* plan = get_plan(ctx, head, (int)n, h, w);
  const float* xd = x;
  const size_t in_bytes = (size_t)n * 3 * h * w * sizeof(float);
  DevBuf tmp;
  if (mem == WSI_MEM_HOST) {
    tmp.alloc(in_bytes);
    CUDA_CHECK(cudaMemcpyAsync(tmp.p, x, in_bytes, cudaMemcpyHostToDevice, s));
    xd = tmp.as<float>();
  }
  {
    StageScope scope(ctx, s, ST_GATHER, (double)n * h * w * 18.0);
    launch_pack_nchw(xd, (int)n, h, w, plan->in_pad.as<bf16>(), s, &ctx->lc, 0, plan->in_planes(), plan->in_plane_stride());
  }
  forward_common(ctx, plan, (int)n, head, out, mem, s);
  WSI_API_END(ctx)
}

int wsi_forward_batch_tta(wsi_ctx* ctx, const float* x, int64_t n, int32_t h, int32_t w, int head, float* out, int mem, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && x && out && n > 0 && n < (1 << 20), WSI_ERR_INVALID, "bad argument");
  WSI_REQUIRE(head == WSI_HEAD_REG || head == WSI_HEAD_CLS, WSI_ERR_UNSUPPORTED, "TTA averages a per-tile vector head (CLS or REG)");
  WSI_REQUIRE(h == w, WSI_ERR_UNSUPPORTED, "TTA views transpose the tile: %dx%d is not square", h, w);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  NetPlan* plan = get_plan(ctx, head, (int)n, h, w);
  const float* xd = x;
  const size_t in_bytes = (size_t)n * 3 * h * w * sizeof(float);
  DevBuf tmp, acc;
  if (mem == WSI_MEM_HOST) {
    tmp.alloc(in_bytes);
    CUDA_CHECK(cudaMemcpyAsync(tmp.p, x, in_bytes, cudaMemcpyHostToDevice, s));
    xd = tmp.as<float>();
  }
  const int64_t out_elems = n * plan->out_dim;
  float* dst = out;
  if (mem == WSI_MEM_HOST) {
    acc.alloc((size_t)out_elems * sizeof(float));
    dst = acc.as<float>();
  }
  for (int view = 0; view < 4; ++view) {                                 // utils/eval.py:305-318 / :386-399
    {
      StageScope scope(ctx, s, ST_GATHER, (double)n * h * w * 18.0);
      launch_pack_nchw(xd, (int)n, h, w, plan->in_pad.as<bf16>(), s, &ctx->lc, view, plan->in_planes(), plan->in_plane_stride());
    }
    plan->run(ctx, s);
    launch_tta_accumulate(dst, plan->logits.as<float>(), out_elems, view == 0, view == 3 ? 4.f : 0.f, s, &ctx->lc);
  }
  if (mem == WSI_MEM_HOST) CUDA_CHECK(cudaMemcpyAsync(out, dst, (size_t)out_elems * sizeof(float), cudaMemcpyDeviceToHost, s));
  check_device_flag(ctx, s);
  resolve_spans(ctx);
  WSI_API_END(ctx)
}

int wsi_forward_patches(wsi_ctx* ctx, const float* xs, int64_t B, int32_t P, int32_t h, int32_t w, float* y, float* ens, int mem,
                        void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && xs && y && ens && B > 0 && P > 0 && B * P < (1 << 20), WSI_ERR_INVALID, "bad argument");
  WSI_REQUIRE(ctx->arch == WSI_ARCH_RESNET18, WSI_ERR_UNSUPPORTED, "the multi-patch forward is resnets_shift.ResNet's (WSI_ARCH_RESNET18)");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = B * P;
  // ensemble head weights: fc = Linear(P*512, P*256) + ReLU + Linear(P*256, 4) (resnets_shift.py:133-139)
  const HostTensor &w1 = weight(ctx, "fc.0.weight"), &b1 = weight(ctx, "fc.0.bias"), &w2 = weight(ctx, "fc.2.weight"), &b2 = weight(ctx, "fc.2.bias");
  WSI_REQUIRE(w1.shape.size() == 2 && w1.shape[1] == (int64_t)P * 512 && b1.numel() == w1.shape[0] && w2.shape.size() == 2 &&
                  w2.shape[1] == w1.shape[0] && b2.numel() == w2.shape[0],
              WSI_ERR_NOMODEL, "fc.0 / fc.2 do not match %d patches x 512 features", P);
  const int n_hid = (int)w1.shape[0], n_out = (int)w2.shape[0];
  NetPlan* plan = get_plan(ctx, WSI_HEAD_CLS, (int)n, h, w);
  DevBuf tmp, hid, ens_d;
  const float* xd = xs;
  const size_t in_bytes = (size_t)n * 3 * h * w * sizeof(float);
  if (mem == WSI_MEM_HOST) {
    tmp.alloc(in_bytes);
    CUDA_CHECK(cudaMemcpyAsync(tmp.p, xs, in_bytes, cudaMemcpyHostToDevice, s));
    xd = tmp.as<float>();
  }
  {
    StageScope scope(ctx, s, ST_GATHER, (double)n * h * w * 18.0);
    launch_pack_nchw(xd, (int)n, h, w, plan->in_pad.as<bf16>(), s, &ctx->lc, 0, plan->in_planes(), plan->in_plane_stride());
  }
  forward_common(ctx, plan, (int)n, WSI_HEAD_CLS, y, mem, s);            // trunk + avgpool + fc0 for all P*B patches
  if (!ctx->ens_ready) {
    upload(ctx->ens_w1, w1.data, s); upload(ctx->ens_b1, b1.data, s); upload(ctx->ens_w2, w2.data, s); upload(ctx->ens_b2, b2.data, s);
    ctx->ens_ready = true;
  }
  hid.alloc((size_t)B * n_hid * sizeof(float));
  float* ens_dst = ens;
  if (mem == WSI_MEM_HOST) {
    ens_d.alloc((size_t)B * n_out * sizeof(float));
    ens_dst = ens_d.as<float>();
  }
  {
    StageScope scope(ctx, s, ST_HEAD, (double)n_hid * P * 512 * 4.0);
    launch_ensemble_head(plan->pooled.as<float>(), (int)B, P, ctx->ens_w1.as<float>(), ctx->ens_b1.as<float>(), n_hid, ctx->ens_w2.as<float>(),
                         ctx->ens_b2.as<float>(), n_out, hid.as<float>(), ens_dst, s, &ctx->lc);
  }
  if (mem == WSI_MEM_HOST) CUDA_CHECK(cudaMemcpyAsync(ens, ens_dst, (size_t)B * n_out * sizeof(float), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  WSI_API_END(ctx)
}

int wsi_forward_tiles(wsi_ctx* ctx, const wsi_slide_desc* slide, const int32_t* tiles_xy, int64_t n_tiles, int head, float* out,
                      int mem, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && out && tiles_xy && n_tiles > 0 && n_tiles < (1 << 20), WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  validate_slide(slide);
  const int64_t rows_r = (slide->rows == 0 && slide->row0 == 0) ? slide->ih : slide->rows;
  check_tiles(slide, tiles_xy, n_tiles, rows_r);
  ensure_lut(ctx);
  const int rs = slide_resize(slide);
  NetPlan* plan = get_plan(ctx, head, (int)n_tiles, slide->ph / rs, slide->pw / rs);
  int64_t rstride = 0;
  const uint8_t* rgb = stage_raster_all(ctx, slide, rows_r, &rstride, s);
  std::vector<int32_t> xy(tiles_xy, tiles_xy + 2 * n_tiles);
  upload(ctx->tiles_dev, xy, s);
  CUDA_CHECK(cudaStreamSynchronize(s));
  {
    StageScope scope(ctx, s, ST_GATHER, 9.0 * n_tiles * (slide->ph / rs) * (slide->pw / rs));
    gather_batch(ctx, slide, rgb, rstride, ctx->tiles_dev.as<int32_t>(), (int)n_tiles, plan->in_pad.as<bf16>(), nullptr, plan->in_planes(),
                 plan->in_plane_stride(), s);
  }
  forward_common(ctx, plan, (int)n_tiles, head, out, mem, s);
  WSI_API_END(ctx)
}

int wsi_resize_argmax(wsi_ctx* ctx, const float* canvas, int64_t H, int64_t W, int64_t H2, int64_t W2, uint8_t* classes,
                      float* pred_or_null, int mem, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && canvas && classes && H > 0 && W > 0 && H2 > 0 && W2 > 0, WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t in_bytes = (size_t)4 * H * W * sizeof(float), out_px = (size_t)H2 * W2;
  if (mem == WSI_MEM_HOST) {
    DevBuf in, cls, pr;
    in.alloc(in_bytes);
    cls.alloc(out_px);
    if (pred_or_null) pr.alloc(out_px * 4 * sizeof(float));
    CUDA_CHECK(cudaMemcpyAsync(in.p, canvas, in_bytes, cudaMemcpyHostToDevice, s));
    launch_resize_argmax(in.as<float>(), H, W, H2, W2, cls.as<uint8_t>(), pred_or_null ? pr.as<float>() : nullptr, s, &ctx->lc);
    CUDA_CHECK(cudaMemcpyAsync(classes, cls.p, out_px, cudaMemcpyDeviceToHost, s));
    if (pred_or_null) CUDA_CHECK(cudaMemcpyAsync(pred_or_null, pr.p, out_px * 4 * sizeof(float), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  } else {
    launch_resize_argmax(canvas, H, W, H2, W2, classes, pred_or_null, s, &ctx->lc);
  }
  WSI_API_END(ctx)
}

int wsi_find_nuclei(wsi_ctx* ctx, const uint8_t* rgb, int64_t row_stride, int rgb_mem, int64_t H, int64_t W, double mu_percent,
                    uint8_t* mask, int mask_mem, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && rgb && mask && H > 0 && W > 0 && row_stride >= 3 * W, WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  // skimage.color.rgb2hsv: the u8 image becomes float64 by MULTIPLICATION with 1/255 (skimage.util.dtype._convert:
  // np.multiply(image, 1. / imax_in)), then S = ptp / max with S = 0 where ptp == 0; mask = S > mu_percent
  std::vector<uint32_t> bits(2048, 0);
  const double inv255 = 1.0 / 255.0;
  for (int mx = 0; mx < 256; ++mx)
    for (int mn = 0; mn <= mx; ++mn) {
      const double a = mx * inv255, b = mn * inv255, delta = a - b;
      const double sat = (delta == 0.0) ? 0.0 : delta / a;
      if (sat > mu_percent) bits[(mx * 256 + mn) >> 5] |= 1u << ((mx * 256 + mn) & 31);
    }
  DevBuf lut, in, out;
  upload(lut, bits, s);
  const uint8_t* rgb_d = rgb;
  int64_t stride_d = row_stride;
  if (rgb_mem == WSI_MEM_HOST) {
    in.alloc((size_t)H * W * 3);
    CUDA_CHECK(cudaMemcpy2DAsync(in.p, (size_t)W * 3, rgb, (size_t)row_stride, (size_t)W * 3, (size_t)H, cudaMemcpyHostToDevice, s));
    rgb_d = in.as<uint8_t>();
    stride_d = 3 * W;
  }
  uint8_t* mask_d = mask;
  if (mask_mem == WSI_MEM_HOST) {
    out.alloc((size_t)H * W);
    mask_d = out.as<uint8_t>();
  }
  launch_find_nuclei(rgb_d, stride_d, H, W, lut.as<uint32_t>(), mask_d, s, &ctx->lc);
  if (mask_mem == WSI_MEM_HOST) CUDA_CHECK(cudaMemcpyAsync(mask, mask_d, (size_t)H * W, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));      // lut / staging buffers die here
  WSI_API_END(ctx)
}

int wsi_plan_tiles_gpu(wsi_ctx* ctx, int64_t ih, int64_t iw, int32_t ph, int32_t pw, int32_t sh, int32_t sw, const uint8_t* mask,
                       int mask_mem, int64_t mh, int64_t mw, double m, int32_t** xy_out, int64_t* n_out, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && xy_out && n_out && mask && ih > 0 && iw > 0 && ph > 0 && pw > 0 && sh > 0 && sw > 0 && m > 0 && mh > 0 && mw > 0,
              WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  *xy_out = nullptr;
  *n_out = 0;
  // candidates in the reference's order (utils/dataset.py:147-166): main grid, right column, bottom row, no corner
  std::vector<int64_t> ys, xs;
  for (int64_t y = 1; y < ih - 1 - ph; y += sh) ys.push_back(y);
  for (int64_t x = 1; x < iw - 1 - pw; x += sw) xs.push_back(x);
  const int64_t x_last = iw - 1 - pw, y_last = ih - 1 - ph;
  std::vector<int32_t> cand;
  cand.reserve((ys.size() * (xs.size() + 1) + xs.size()) * 2);
  for (int64_t y : ys) for (int64_t x : xs) { cand.push_back((int32_t)x); cand.push_back((int32_t)y); }
  for (int64_t y : ys) { cand.push_back((int32_t)x_last); cand.push_back((int32_t)y); }
  for (int64_t x : xs) { cand.push_back((int32_t)x); cand.push_back((int32_t)y_last); }
  const int64_t n = (int64_t)cand.size() / 2;
  const int64_t dx = (int64_t)((double)pw * m), dy = (int64_t)((double)ph * m);
  std::vector<int64_t> win((size_t)2 * n);
  for (int64_t i = 0; i < n; ++i) {
    // the reference would index the mask with a negative origin
    WSI_REQUIRE(cand[2 * i] >= 0 && cand[2 * i + 1] >= 0, WSI_ERR_DEGENERATE, "tile origin (%d,%d) is negative", cand[2 * i], cand[2 * i + 1]);
    win[2 * i] = (int64_t)((double)cand[2 * i] * m);
    win[2 * i + 1] = (int64_t)((double)cand[2 * i + 1] * m);
  }
  std::vector<int32_t> keep;
  if (n > 0) {
    DevBuf mask_buf, win_d, cnt_d, size_d;
    const uint8_t* mask_d = mask;
    if (mask_mem == WSI_MEM_HOST) {
      mask_buf.alloc((size_t)mh * mw);
      CUDA_CHECK(cudaMemcpyAsync(mask_buf.p, mask, (size_t)mh * mw, cudaMemcpyHostToDevice, s));
      mask_d = mask_buf.as<uint8_t>();
    }
    upload(win_d, win, s);
    cnt_d.alloc((size_t)n * sizeof(uint32_t));
    size_d.alloc((size_t)n * sizeof(int64_t));
    launch_window_count(mask_d, mh, mw, win_d.as<int64_t>(), n, dx, dy, cnt_d.as<uint32_t>(), size_d.as<int64_t>(), s, &ctx->lc);
    std::vector<uint32_t> cnt((size_t)n);
    std::vector<int64_t> size((size_t)n);
    CUDA_CHECK(cudaMemcpyAsync(cnt.data(), cnt_d.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(size.data(), size_d.p, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    for (int64_t i = 0; i < n; ++i) {
      // isforeground: python int / python int >= 0.05; an empty window raises ZeroDivisionError in the reference
      WSI_REQUIRE(size[(size_t)i] > 0, WSI_ERR_DEGENERATE, "tile (%d,%d): empty mask window", cand[2 * i], cand[2 * i + 1]);
      if ((double)cnt[(size_t)i] / (double)size[(size_t)i] >= 0.05) { keep.push_back(cand[2 * i]); keep.push_back(cand[2 * i + 1]); }
    }
  }
  int32_t* buf = (int32_t*)malloc(std::max<size_t>(keep.size(), 2) * sizeof(int32_t));
  WSI_REQUIRE(buf != nullptr, WSI_ERR_NOMEM, "out of host memory");
  if (!keep.empty()) memcpy(buf, keep.data(), keep.size() * sizeof(int32_t));
  *xy_out = buf;
  *n_out = (int64_t)keep.size() / 2;
  WSI_API_END(ctx)
}

int wsi_debug_umma_shift(wsi_ctx* ctx, const void* A_dev, const void* B_dev, int r, int s, int pitch, int use_base_offset, float* D_dev,
                         void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && A_dev && B_dev && D_dev, WSI_ERR_INVALID, "NULL argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  debug_umma_shift(A_dev, B_dev, r, s, pitch, use_base_offset, D_dev, (cudaStream_t)stream);
  CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  WSI_API_END(ctx)
}

int wsi_synth_slide(wsi_ctx* ctx, int64_t ih, int64_t iw, uint32_t seed, int64_t y0, int64_t y1, const uint8_t* lut, uint8_t* rgb_dev,
                    int64_t row_stride, uint8_t* mask_dev_or_null, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && lut && rgb_dev && y0 >= 0 && y1 >= y0 && y1 <= ih && iw > 0 && row_stride >= 3 * iw, WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  DevBuf l;
  l.alloc(16 * 8 * 3);
  CUDA_CHECK(cudaMemcpyAsync(l.p, lut, 16 * 8 * 3, cudaMemcpyHostToDevice, s));
  launch_synth(ih, iw, seed, y0, y1, l.as<uint8_t>(), rgb_dev, row_stride, mask_dev_or_null, s, &ctx->lc);
  CUDA_CHECK(cudaStreamSynchronize(s));
  WSI_API_END(ctx)
}

int wsi_debug_conv(wsi_ctx* ctx, const void* x, int n, int h, int w, int cin, const float* wt, int cout, int ksize, int stride, int pad,
                   const float* scale, const float* bias, const void* res, int relu, int up2, const void* skip, int cskip, void* y,
                   void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && x && wt && y, WSI_ERR_INVALID, "NULL argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<ConvInputPart> parts;
  parts.push_back(ConvInputPart{TensorView{x, n, h, w, cin}, up2 != 0});
  if (skip) parts.push_back(ConvInputPart{TensorView{skip, n, up2 ? 2 * h : h, up2 ? 2 * w : w, cskip}, false});
  ConvSpec sp;
  sp.ksize = ksize; sp.stride = stride; sp.pad = pad; sp.cout = cout; sp.relu = relu != 0;
  ConvOp op;
  if (ConvOp::routes_to_upstream(parts, sp, res, LAYOUT_PLANAR)) {
    // the x2 row-stream kernel only writes the planar layout (what the engine chains): run it that way and hand NHWC back
    const int OH = 2 * h, OW = 2 * w;
    const PlanarDims od = PlanarDims::make(OH, OW, cout, LAYOUT_PLANAR);
    DevBuf tmp;
    tmp.alloc(od.bytes(n));
    CUDA_CHECK(cudaMemsetAsync(tmp.p, 0, tmp.bytes, s));
    op.build(parts, sp, wt, scale, bias, res, tmp.p, nullptr, nullptr, nullptr, ctx->err_flag.as<int>(), ctx->num_sms, LAYOUT_PLANAR);
    op.launch(s, &ctx->lc);
    launch_relayout_nhwc(tmp.p, y, n, OH, OW, cout, s, &ctx->lc);
    check_device_flag(ctx, s);
  } else {
    op.build(parts, sp, wt, scale, bias, res, y, nullptr, nullptr, nullptr, ctx->err_flag.as<int>(), ctx->num_sms);
    op.launch(s, &ctx->lc);
    check_device_flag(ctx, s);
  }
  WSI_API_END(ctx)
}

int wsi_debug_conv_f32(wsi_ctx* ctx, const float* x, int n, int h, int w, int cin, const float* wt, int cout, int ksize, int stride, int pad,
                       const float* scale, const float* bias, const float* res, int relu, int up2, const float* skip, int cskip, float* y,
                       void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && x && wt && y, WSI_ERR_INVALID, "NULL argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int hin = up2 ? 2 * h : h, win = up2 ? 2 * w : w;
  const int oh = (hin + 2 * pad - ksize) / stride + 1, ow = (win + 2 * pad - ksize) / stride + 1;
  DevBuf xs, ss, rs, ys;
  xs.alloc((size_t)n * h * w * cin * 3 * 2);
  launch_split_planes(x, (int64_t)n * h * w, cin, xs.as<bf16>(), s, &ctx->lc);
  std::vector<ConvInputPart> parts;
  parts.push_back(ConvInputPart{TensorView{xs.p, n, h, w, cin}, up2 != 0});
  if (skip) {
    ss.alloc((size_t)n * hin * win * cskip * 3 * 2);
    launch_split_planes(skip, (int64_t)n * hin * win, cskip, ss.as<bf16>(), s, &ctx->lc);
    parts.push_back(ConvInputPart{TensorView{ss.p, n, hin, win, cskip}, false});
  }
  if (res) {
    rs.alloc((size_t)n * oh * ow * cout * 3 * 2);
    launch_split_planes(res, (int64_t)n * oh * ow, cout, rs.as<bf16>(), s, &ctx->lc);
  }
  ys.alloc((size_t)n * oh * ow * cout * 3 * 2);
  ConvSpec sp;
  sp.ksize = ksize; sp.stride = stride; sp.pad = pad; sp.cout = cout; sp.relu = relu != 0;
  ConvOp op;
  op.build(parts, sp, wt, scale, bias, res ? rs.p : nullptr, ys.p, nullptr, nullptr, nullptr, ctx->err_flag.as<int>(), ctx->num_sms, LAYOUT_NHWC,
           LAYOUT_NHWC, 1);
  op.launch(s, &ctx->lc);
  launch_merge_planes(ys.as<bf16>(), (int64_t)n * oh * ow, cout, y, s, &ctx->lc);
  check_device_flag(ctx, s);
  WSI_API_END(ctx)
}

// ---- peer-mapped result buffer (multi-GPU, one process per GPU on one node) ---------------------------------------
int wsi_ipc_alloc(wsi_ctx* ctx, int64_t bytes, void** dev_ptr, uint8_t* handle /*[64]*/) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && dev_ptr && handle && bytes > 0, WSI_ERR_INVALID, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  WSI_REQUIRE(e == cudaSuccess, WSI_ERR_NOMEM, "cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    WSI_THROW(WSI_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  }
  memcpy(handle, &h, 64);
  *dev_ptr = p;
  WSI_API_END(ctx)
}

int wsi_ipc_open(wsi_ctx* ctx, const uint8_t* handle, void** dev_ptr) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && handle && dev_ptr, WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    WSI_THROW(WSI_ERR_CUDA, "cudaIpcOpenMemHandle failed: %s (peer access over NVLink / PCIe unavailable between these processes)", cudaGetErrorString(e));
  }
  *dev_ptr = p;
  WSI_API_END(ctx)
}

int wsi_ipc_close(wsi_ctx* ctx, void* dev_ptr) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && dev_ptr, WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  CUDA_CHECK(cudaDeviceSynchronize());
  CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
  WSI_API_END(ctx)
}

int wsi_ipc_free(wsi_ctx* ctx, void* dev_ptr) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && dev_ptr, WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  CUDA_CHECK(cudaDeviceSynchronize());
  CUDA_CHECK(cudaFree(dev_ptr));
  WSI_API_END(ctx)
}

// Page-lock caller-owned host memory (e.g. this rank's rows of a POSIX-shared result buffer) so that copies to / from it are
// asynchronous DMA.  One registration for the whole range (a copy may not span two registrations).  On failure the CUDA
// error state is cleared and WSI_ERR_NOMEM is returned: the caller falls back to pageable or private pinned memory.
int wsi_host_register(void* ptr, int64_t bytes) {
  wsi_ctx* none = nullptr;
  WSI_API_BEGIN
  WSI_REQUIRE(ptr && bytes > 0, WSI_ERR_INVALID, "bad argument");
  cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    WSI_THROW(WSI_ERR_NOMEM, "cudaHostRegister(%lld bytes) failed: %s", (long long)bytes, cudaGetErrorString(e));
  }
  WSI_API_END(none)
}

int wsi_host_unregister(void* ptr) {
  wsi_ctx* none = nullptr;
  WSI_API_BEGIN
  WSI_REQUIRE(ptr, WSI_ERR_INVALID, "bad argument");
  cudaHostUnregister(ptr);
  cudaGetLastError();
  WSI_API_END(none)
}

int wsi_check(wsi_ctx* ctx, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx, WSI_ERR_INVALID, "ctx is NULL");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  check_device_flag(ctx, (cudaStream_t)stream);
  resolve_spans(ctx);
  WSI_API_END(ctx)
}

int wsi_debug_gather(wsi_ctx* ctx, const wsi_slide_desc* slide, const int32_t* tiles_xy, int n, float* norm_out, void* padded_out,
                     void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && tiles_xy && n > 0, WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  validate_slide(slide);
  const int64_t rows_r = (slide->rows == 0 && slide->row0 == 0) ? slide->ih : slide->rows;
  check_tiles(slide, tiles_xy, n, rows_r);
  ensure_lut(ctx);
  int64_t rstride = 0;
  const uint8_t* rgb = stage_raster_all(ctx, slide, rows_r, &rstride, s);
  std::vector<int32_t> xy(tiles_xy, tiles_xy + 2 * (size_t)n);
  upload(ctx->tiles_dev, xy, s);
  CUDA_CHECK(cudaStreamSynchronize(s));
  const int rs = slide_resize(slide);
  if (padded_out) CUDA_CHECK(cudaMemsetAsync(padded_out, 0, (size_t)n * (slide->ph / rs + 6) * (slide->pw / rs + 8) * 8, s));
  gather_batch(ctx, slide, rgb, rstride, ctx->tiles_dev.as<int32_t>(), n, static_cast<bf16*>(padded_out), norm_out, 1, 0, s);
  CUDA_CHECK(cudaStreamSynchronize(s));
  WSI_API_END(ctx)
}

int wsi_debug_stem(wsi_ctx* ctx, const wsi_slide_desc* slide, const int32_t* tiles_xy, int n, const float* wt, const float* scale,
                   const float* bias, void* y, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && tiles_xy && n > 0 && wt && y, WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  validate_slide(slide);
  WSI_REQUIRE(slide_resize(slide) == 1, WSI_ERR_UNSUPPORTED, "wsi_debug_stem: resize is not applied here");
  const int64_t rows_r = (slide->rows == 0 && slide->row0 == 0) ? slide->ih : slide->rows;
  check_tiles(slide, tiles_xy, n, rows_r);
  ensure_lut(ctx);
  int64_t rstride = 0;
  const uint8_t* rgb = stage_raster_all(ctx, slide, rows_r, &rstride, s);
  std::vector<int32_t> xy(tiles_xy, tiles_xy + 2 * (size_t)n);
  upload(ctx->tiles_dev, xy, s);
  CUDA_CHECK(cudaStreamSynchronize(s));
  DevBuf pad;
  pad.alloc((size_t)n * (slide->ph + 6) * (slide->pw + 8) * 8);
  CUDA_CHECK(cudaMemsetAsync(pad.p, 0, pad.bytes, s));
  launch_gather(rgb, rstride, slide->row0, ctx->tiles_dev.as<int32_t>(), n, slide->ph, slide->pw, ctx->lut.as<float>(), pad.as<bf16>(),
                nullptr, s, &ctx->lc);
  ConvOp op;
  op.build_stem(pad.p, n, slide->ph, slide->pw, wt, scale, bias, y, ctx->err_flag.as<int>(), ctx->num_sms);
  op.launch(s, &ctx->lc);
  check_device_flag(ctx, s);
  WSI_API_END(ctx)
}

int wsi_debug_maxpool(wsi_ctx* ctx, const void* x, int n, int h, int w, int c, void* y, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && x && y && c % 8 == 0, WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  launch_maxpool(static_cast<const bf16*>(x), n, h, w, c, static_cast<bf16*>(y), (cudaStream_t)stream, &ctx->lc);
  CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  WSI_API_END(ctx)
}

int wsi_stage_stats(wsi_ctx* ctx, const char* stage, double* ms_out, int64_t* launches_out, double* work_out) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && stage, WSI_ERR_INVALID, "NULL argument");
  for (int i = 0; i < ST_COUNT; ++i)
    if (strcmp(stage, kStageNames[i]) == 0) {
      if (ms_out) *ms_out = ctx->acc[i].ms;
      if (launches_out) *launches_out = ctx->acc[i].launches;
      if (work_out) *work_out = ctx->acc[i].work;
      return WSI_OK;
    }
  WSI_THROW(WSI_ERR_INVALID, "unknown stage '%s'", stage);
  WSI_API_END(ctx)
}

int wsi_op_stats(wsi_ctx* ctx, int idx, char* desc, int desc_cap, char* kernel, int kernel_cap, double* ms_per_launch, double* flops,
                 double* bytes, int64_t* count) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx, WSI_ERR_INVALID, "ctx is NULL");
  NetPlan* p = ctx->plan.get();
  WSI_REQUIRE(p && p->trace, WSI_ERR_INVALID, "no traced plan: set option op_trace = 1 before running");
  if (idx < 0 || idx >= (int)p->op_stats.size()) return WSI_ERR_INVALID;      // end of the list: not an error worth a message
  CUDA_CHECK(cudaSetDevice(ctx->device));
  CUDA_CHECK(cudaDeviceSynchronize());
  p->resolve_trace();
  const NetPlan::OpStat& o = p->op_stats[idx];
  if (desc && desc_cap > 0) snprintf(desc, (size_t)desc_cap, "%s", o.desc.c_str());
  if (kernel && kernel_cap > 0) snprintf(kernel, (size_t)kernel_cap, "%s", o.kernel.c_str());
  if (ms_per_launch) *ms_per_launch = o.count ? o.ms / (double)o.count : 0.0;
  if (flops) *flops = o.flops;
  if (bytes) *bytes = o.bytes;
  if (count) *count = o.count;
  WSI_API_END(ctx)
}

int wsi_stage_reset(wsi_ctx* ctx) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx, WSI_ERR_INVALID, "ctx is NULL");
  resolve_spans(ctx);
  for (int i = 0; i < ST_COUNT; ++i) ctx->acc[i] = StageAcc();
  WSI_API_END(ctx)
}

}  // extern "C"

namespace wsi {
int ctx_device(const wsi_ctx* c) { return c->device; }
LaunchCounter* ctx_launch_counter(wsi_ctx* c) { return &c->lc; }
}  // namespace wsi

// ---- tumour-bed post-processing ---------------------------------------------------------------
namespace wsi {

struct P2 { int64_t x, y; };
static inline __int128 cross3(const P2& o, const P2& a, const P2& b) { return (__int128)(a.x - o.x) * (b.y - o.y) - (__int128)(a.y - o.y) * (b.x - o.x); }
static inline int64_t floor_div(int64_t a, int64_t b) { int64_t q = a / b, r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q; }
static inline int64_t ceil_div_s(int64_t a, int64_t b) { return -floor_div(-a, b); }

// skimage.morphology.convex_hull_image (offset_coordinates=True, include_borders=True) restated with exact integer
// arithmetic in doubled coordinates: candidate points = the first / last set pixel of every row (the hull of the set
// pixels is the hull of those), each replaced by the four midpoints of its pixel edges (x +- 1/2, y), (x, y +- 1/2);
// hull by Andrew's monotone chain; a pixel centre belongs to the image iff it lies inside or ON the hull polygon.
// Returns per row the inclusive pixel range [xl, xr] (xl > xr: empty).
static void hull_row_ranges(const std::vector<int32_t>& xmin, const std::vector<int32_t>& xmax, int64_t H, std::vector<int32_t>& xl, std::vector<int32_t>& xr) {
  std::vector<P2> pts;
  for (int64_t y = 0; y < H; ++y) {
    if (xmax[(size_t)y] < 0) continue;
    const int32_t ends[2] = {xmin[(size_t)y], xmax[(size_t)y]};
    for (int e = 0; e < (ends[0] == ends[1] ? 1 : 2); ++e) {
      const int64_t X = 2 * (int64_t)ends[e], Y = 2 * y;
      pts.push_back({X - 1, Y}); pts.push_back({X + 1, Y}); pts.push_back({X, Y - 1}); pts.push_back({X, Y + 1});
    }
  }
  xl.assign((size_t)H, 1);
  xr.assign((size_t)H, 0);
  if (pts.empty()) return;
  std::sort(pts.begin(), pts.end(), [](const P2& a, const P2& b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
  pts.erase(std::unique(pts.begin(), pts.end(), [](const P2& a, const P2& b) { return a.x == b.x && a.y == b.y; }), pts.end());
  std::vector<P2> hull(2 * pts.size());
  size_t k = 0;
  for (size_t i = 0; i < pts.size(); ++i) {
    while (k >= 2 && cross3(hull[k - 2], hull[k - 1], pts[i]) <= 0) --k;
    hull[k++] = pts[i];
  }
  for (size_t i = pts.size() - 1, t = k + 1; i-- > 0;) {
    while (k >= t && cross3(hull[k - 2], hull[k - 1], pts[i]) <= 0) --k;
    hull[k++] = pts[i];
  }
  hull.resize(k > 1 ? k - 1 : k);
  const size_t V = hull.size();
  for (int64_t y = 0; y < H; ++y) {
    const int64_t Y = 2 * y;
    // exact min / max of X over the intersection of the line Y with the polygon: rationals num / den (den > 0)
    bool any = false;
    int64_t lo_n = 0, lo_d = 1, hi_n = 0, hi_d = 1;
    auto upd = [&](int64_t n, int64_t d) {
      if (!any) { lo_n = hi_n = n; lo_d = hi_d = d; any = true; return; }
      if ((__int128)n * lo_d < (__int128)lo_n * d) { lo_n = n; lo_d = d; }
      if ((__int128)n * hi_d > (__int128)hi_n * d) { hi_n = n; hi_d = d; }
    };
    for (size_t i = 0; i < V; ++i) {
      const P2 &p = hull[i], &q = hull[(i + 1) % V];
      if ((p.y < Y && q.y < Y) || (p.y > Y && q.y > Y)) continue;
      if (p.y == q.y) { upd(p.x, 1); upd(q.x, 1); continue; }
      int64_t d = q.y - p.y, n = p.x * d + (q.x - p.x) * (Y - p.y);
      if (d < 0) { d = -d; n = -n; }
      upd(n, d);
    }
    if (!any) continue;
    // pixel x is inside iff lo <= 2x <= hi
    xl[(size_t)y] = (int32_t)ceil_div_s(lo_n, 2 * lo_d);
    xr[(size_t)y] = (int32_t)floor_div(hi_n, 2 * hi_d);
  }
}

}  // namespace wsi

extern "C" {

int wsi_hull_rows(const int32_t* xmin, const int32_t* xmax, int64_t H, int32_t* xl, int32_t* xr) {
  wsi_ctx* none = nullptr;
  WSI_API_BEGIN
  WSI_REQUIRE(xmin && xmax && xl && xr && H > 0, WSI_ERR_INVALID, "bad argument");
  std::vector<int32_t> a(xmin, xmin + H), b(xmax, xmax + H), l, r;
  hull_row_ranges(a, b, H, l, r);
  memcpy(xl, l.data(), (size_t)H * sizeof(int32_t));
  memcpy(xr, r.data(), (size_t)H * sizeof(int32_t));
  WSI_API_END(none)
}

int wsi_morph(wsi_ctx* ctx, const uint8_t* src, int64_t H, int64_t W, int op, int k, uint8_t* dst, int mem, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && src && dst && H > 0 && W > 0 && op >= 0 && op <= 3, WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)H * W;
  DevBuf in, tmp, out;
  tmp.alloc(n);
  const uint8_t* sd = src;
  uint8_t* dd = dst;
  if (mem == WSI_MEM_HOST) {
    in.alloc(n); out.alloc(n);
    CUDA_CHECK(cudaMemcpyAsync(in.p, src, n, cudaMemcpyHostToDevice, s));
    sd = in.as<uint8_t>(); dd = out.as<uint8_t>();
  }
  // cv2.morphologyEx: OPEN = dilate(erode(src)), CLOSE = erode(dilate(src)), same kernel and anchor for both passes
  const bool first_max = (op == WSI_MORPH_DILATE || op == WSI_MORPH_CLOSE);
  launch_morph(sd, H, W, k, first_max, tmp.as<uint8_t>(), dd, nullptr, s, &ctx->lc);
  if (op == WSI_MORPH_OPEN || op == WSI_MORPH_CLOSE) launch_morph(dd, H, W, k, !first_max, tmp.as<uint8_t>(), dd, nullptr, s, &ctx->lc);
  if (mem == WSI_MEM_HOST) CUDA_CHECK(cudaMemcpyAsync(dst, dd, n, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  WSI_API_END(ctx)
}

int wsi_tumor_bed(wsi_ctx* ctx, const uint8_t* src, int64_t H, int64_t W, const uint8_t* rule_lut, int open_k, int dilate_k, uint8_t* opened_out,
                  uint8_t* hull_out, uint8_t* outline_out, int64_t* n_open_out, int mem, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && src && rule_lut && H > 0 && W > 0 && W < (1LL << 30) && open_k >= 1, WSI_ERR_INVALID, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)H * W;
  const bool host = (mem == WSI_MEM_HOST);
  DevBuf in, lut, bin, tmp, hull, perim, outline, cnt, ext;
  const uint8_t* sd = src;
  if (host) { in.alloc(n); CUDA_CHECK(cudaMemcpyAsync(in.p, src, n, cudaMemcpyHostToDevice, s)); sd = in.as<uint8_t>(); }
  lut.alloc(256);
  CUDA_CHECK(cudaMemcpyAsync(lut.p, rule_lut, 256, cudaMemcpyHostToDevice, s));
  cnt.alloc(sizeof(unsigned long long));
  CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), s));
  tmp.alloc(n);
  uint8_t* opened = (!host && opened_out) ? opened_out : (bin.alloc(n), bin.as<uint8_t>());
  launch_lut(sd, (int64_t)n, lut.as<uint8_t>(), opened, nullptr, s, &ctx->lc);                                   // tb = rule(src)
  launch_morph(opened, H, W, open_k, false, tmp.as<uint8_t>(), opened, nullptr, s, &ctx->lc);                    // cv2.MORPH_OPEN: erode ...
  launch_morph(opened, H, W, open_k, true, tmp.as<uint8_t>(), opened, cnt.as<unsigned long long>(), s, &ctx->lc);   // ... then dilate
  unsigned long long n_open = 0;
  CUDA_CHECK(cudaMemcpyAsync(&n_open, cnt.p, sizeof(n_open), cudaMemcpyDeviceToHost, s));
  std::vector<int32_t> xmin, xmax, xl, xr;
  if (hull_out || outline_out) {
    ext.alloc((size_t)H * 4 * sizeof(int32_t));
    int32_t* e = ext.as<int32_t>();
    launch_row_extent(opened, H, W, e, e + H, s, &ctx->lc);
    xmin.resize((size_t)H); xmax.resize((size_t)H);
    CUDA_CHECK(cudaMemcpyAsync(xmin.data(), e, (size_t)H * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(xmax.data(), e + H, (size_t)H * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  }
  CUDA_CHECK(cudaStreamSynchronize(s));
  if (n_open_out) *n_open_out = (int64_t)n_open;
  if (host && opened_out) CUDA_CHECK(cudaMemcpyAsync(opened_out, opened, n, cudaMemcpyDeviceToHost, s));
  if (hull_out || outline_out) {
    hull_row_ranges(xmin, xmax, H, xl, xr);                   // O(H x hull vertices) on the host: 2H points in, a few hundred vertices out
    int32_t* e = ext.as<int32_t>();
    CUDA_CHECK(cudaMemcpyAsync(e + 2 * H, xl.data(), (size_t)H * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(e + 3 * H, xr.data(), (size_t)H * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    uint8_t* hd = (!host && hull_out) ? hull_out : (hull.alloc(n), hull.as<uint8_t>());
    launch_fill_rows(e + 2 * H, e + 3 * H, H, W, hd, s, &ctx->lc);                                               // chull(tb)
    if (host && hull_out) CUDA_CHECK(cudaMemcpyAsync(hull_out, hd, n, cudaMemcpyDeviceToHost, s));
    if (outline_out) {
      perim.alloc(n);
      launch_bwperim(hd, H, W, perim.as<uint8_t>(), s, &ctx->lc);                                                // bwperim(tb_pred)
      uint8_t* od = (!host) ? outline_out : (outline.alloc(n), outline.as<uint8_t>());
      if (dilate_k > 1) launch_morph(perim.as<uint8_t>(), H, W, dilate_k, true, tmp.as<uint8_t>(), od, nullptr, s, &ctx->lc);   // cv2.dilate
      else CUDA_CHECK(cudaMemcpyAsync(od, perim.p, n, cudaMemcpyDeviceToDevice, s));
      if (host) CUDA_CHECK(cudaMemcpyAsync(outline_out, od, n, cudaMemcpyDeviceToHost, s));
    }
  }
  CUDA_CHECK(cudaStreamSynchronize(s));
  WSI_API_END(ctx)
}

int wsi_overlay(wsi_ctx* ctx, const uint8_t* rgb, const uint8_t* heat, int64_t H, int64_t W, int mode, const uint8_t* on_lut, const uint8_t* im,
                const uint8_t* perim, uint8_t* out, int mem, void* stream) {
  WSI_API_BEGIN
  WSI_REQUIRE(ctx && rgb && heat && out && H > 0 && W > 0 && (mode == WSI_OVERLAY_HEAT || mode == WSI_OVERLAY_BED), WSI_ERR_INVALID, "bad argument");
  WSI_REQUIRE(mode != WSI_OVERLAY_HEAT || on_lut, WSI_ERR_INVALID, "WSI_OVERLAY_HEAT needs the 256-level rule table");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)H * W;
  const bool host = (mem == WSI_MEM_HOST);
  DevBuf b_rgb, b_heat, b_im, b_perim, b_out, lut;
  auto dev_in = [&](DevBuf& b, const uint8_t* p, size_t bytes) -> const uint8_t* {
    if (!p || !host) return p;
    b.alloc(bytes);
    CUDA_CHECK(cudaMemcpyAsync(b.p, p, bytes, cudaMemcpyHostToDevice, s));
    return b.as<uint8_t>();
  };
  const uint8_t *d_rgb = dev_in(b_rgb, rgb, 3 * n), *d_heat = dev_in(b_heat, heat, n), *d_im = dev_in(b_im, im, n), *d_perim = dev_in(b_perim, perim, n);
  uint8_t* d_out = host ? (b_out.alloc(3 * n), b_out.as<uint8_t>()) : out;
  if (mode == WSI_OVERLAY_HEAT) {
    lut.alloc(256);
    CUDA_CHECK(cudaMemcpyAsync(lut.p, on_lut, 256, cudaMemcpyHostToDevice, s));
    launch_overlay_heat(d_rgb, d_heat, (int64_t)n, lut.as<uint8_t>(), d_out, s, &ctx->lc);
  } else {
    launch_overlay_bed(d_rgb, d_heat, d_im, d_perim, (int64_t)n, d_out, s, &ctx->lc);
  }
  if (host) CUDA_CHECK(cudaMemcpyAsync(out, d_out, 3 * n, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  WSI_API_END(ctx)
}

}  // extern "C"
