// libspef_b200.so -- C ABI (include/spef_b200.h): context, weight packer (BN fold), layer plan, launches.
// The network topology follows src/modeling/backbone/mobilenet_v2.py:232-271 and
// src/modeling/common/pytorch_layers.py:65-98 of the reference; state_dict keys follow SURVEY Appendix B.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <chrono>
#include <map>
#include <string>
#include <vector>
#include <utility>

#include "../../include/spef_b200.h"
#include "common.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm_tcgen05_v2.cuh"
#include "dwconv_tma.cuh"
#include "fused_block.cuh"
#include "fused_block_t.cuh"
#include "dw_project.cuh"
#include "conv_pool.cuh"
#include "kernels_conv.cuh"
#include "kernels_post.cuh"
#include "kernels_ingest.cuh"
#include "decode_stream.cuh"

using namespace spef;

namespace {

enum LayerKind { K_STEM = 0, K_PW = 1, K_DW = 2, K_POOL = 3, K_HEAD = 4 };
enum BufId { BUF_IMG = -1, BUF_P = 0, BUF_Q = 1, BUF_H1 = 2, BUF_H2 = 3, BUF_POOL = 4, BUF_HEAD = 5 };

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

struct Layer {
  int kind = 0;
  std::string prefix;  // state_dict prefix of the ConvBnAct ("features.features.3.conv.1"), empty for pool/head
  int cin = 0, cout = 0, hin = 0, win = 0, hout = 0, wout = 0, stride = 1, relu = 0;
  int residual = 0;    // project conv of a residual block: D += block input
  int src = 0, dst = 0, res_buf = -1;
  float* w_f32 = nullptr;   // stem [27][32] | dw [9][C] | pw/head Wt [K][Npad] (SIMT operand)
  bf16* w_bf16 = nullptr;   // pw/head [Npad][K] (tcgen05 operand, PyTorch layout)
  float* bias = nullptr;    // zero padded
  int n_pad = 0;            // head: n_ori + n_pos rounded up to 8; else cout
  // tcgen05 plan
  int block_n = 0, stages = 0;
  size_t smem = 0;
  CUtensorMap tmA, tmW;
  bool tmW_ready = false;
  int plan_batch = -1;      // batch size the cached activation tensor maps (tmA/tmD/tmX) were encoded for
  // TMA depthwise plan
  int dw_cv = 0;
  int dw_tx = 4;            // outputs per thread along x of the TMA depthwise kernel
  dw::DwParams dwp;
  CUtensorMap tmX;
  // host copies of the folded bias (pw, dw) and the packed [9][C] depthwise weights: inputs of the fused-block plan
  std::vector<float> h_bias, h_wdw;
  std::vector<bf16> h_wb;   // pointwise weights [N][K] as uploaded (BF16 engine)
};

// One InvertedResidual block as a single fused kernel (fused_block.cuh): layers [first, first + n_layers)
struct Block {
  int first = 0, n_layers = 0;   // expand (optional), depthwise, project
  int i_exp = -1, i_dw = -1, i_proj = -1;
  bool fusable = false;
  int ng = 2;                    // worker groups
  fb::FbParams prm;
  size_t smem = 0;
  float* aux = nullptr;          // device [n_chunks][AUX_FLOATS]
  CUtensorMap tmX, tmWe, tmWp;
  bool tmW_ready = false;
  const void* tmX_ptr = nullptr;
  int tmX_batch = -1;
  // channel-lane variant (fused_block_t.cuh): permuted / zero-padded weights, one 128-slot chunk per TMEM pass
  bool t_ok = false;
  int t_ng = 2;                  // worker groups of the channel-lane kernel
  fbt::FbtParams tprm;
  size_t t_smem = 0;
  bf16* t_we = nullptr;          // [n_chunks*128][Cin]
  bf16* t_wp = nullptr;          // [Cout][n_chunks*128]
  float* t_aux = nullptr;        // [n_chunks][11][128]
  CUtensorMap t_tmX, t_tmWe, t_tmWp;
  bool t_tmW_ready = false;
  const void* t_tmX_ptr = nullptr;
  int t_tmX_batch = -1;
  // stem fused into the t = 1 block (fused_block_t.cuh, STEM instantiation): block 0 only
  bool s_ok = false;
  fbt::FbtParams sprm;
  size_t s_smem = 0;
  bf16* s_we = nullptr;          // window matrix [224][64]: rows 96..127 = the stem's folded weights [32 ch][27 taps -> 32]
  float* s_aux = nullptr;
  CUtensorMap s_tmImg, s_tmWe, s_tmWp;
  bool s_tmW_ready = false;
  const void* s_img_ptr = nullptr;
  int s_img_batch = -1, s_img_u8 = -1;
  // expand GEMM + fused depthwise -> project kernel (dw_project.cuh): the wide blocks that have no single-kernel plan
  bool dp_ok = false;
  dwp::DwpParams dprm;
  size_t dp_smem = 0;
  float* dp_wdw = nullptr;       // [k_chunks][10][64] depthwise weights + bias per K chunk
  CUtensorMap dp_tmX, dp_tmW;
  bool dp_tmW_ready = false;
  const void* dp_tmX_ptr = nullptr;
  int dp_tmX_batch = -1;
};

const double kBnEps = 1e-5;  // torch.nn.BatchNorm2d default (pytorch_layers.py:55-56)

}  // namespace

struct spef_ctx {
  spef_config cfg;
  std::string err;
  std::map<std::string, HostTensor> host_tensors;
  std::vector<Layer> layers;
  std::vector<Block> blocks;
  int fuse = 1;        // fused InvertedResidual kernels on the BF16 tcgen05 path (SPEF_FUSE=0 disables)
  int fb_gw = 4;       // warps per worker group of the fused kernel (SPEF_FB_GW = 4 | 8; 4 measured faster: more registers per thread)
  int dwp_enable = 1;  // SPEF_DWP=0: per-layer kernels for the blocks without a single-kernel plan
  int dwp_w_stages = 0; // SPEF_DWP_WST (developer A/B): cap on the project-weight ring depth of the depthwise -> project kernel
  int dwp_opt_skip = 0;
  int dwp_min_tiles = 37;  // fewer row tiles than this (a few-image step): per-layer kernels instead of the depthwise -> project kernel (SPEF_DWP_MIN_TILES)
  int dwp_s2_box_kb = 60;  // stride-2 depthwise -> project: largest input box (SPEF_DWP_S2_BOX_KB)
  int dwp_force = 0;   // SPEF_DWP_FORCE=1 (tests): expand GEMM + depthwise->project kernel for every block it can run, ahead of the single-kernel plans
  int fb_max_cin = 64; // fuse blocks with Cin <= this (SPEF_FB_MAX_CIN); wider blocks measured faster as three kernels
  int stem_patch = 1;  // stem input patches staged by TMA (SPEF_STEM_PATCH=0: gather the 27 taps from global memory)
  int stem_prod = 2;   // im2col producer groups (128 threads each) of the tcgen05 stem (SPEF_STEM_PROD = 1 | 2)
  int fbt_a2_bufs = 2; // A2 buffers per worker group of the channel-lane kernel where shared memory allows (SPEF_FBT_A2 = 1 | 2)
  int stem_fuse = 1;   // stem conv fused into the first InvertedResidual block's kernel (SPEF_STEM_FUSE=0: separate stem launch)
  int fbt_max_pstages = 4; // project accumulator stages of the channel-lane kernel, as many as TMEM has columns for (SPEF_FBT_PSTAGES)
  int fb_variant = 1;  // 1: channel-lane fused kernel where it applies, else the staged one; 0: staged kernel only (SPEF_FB_VARIANT)
  int fb_trace_block = -1;  // SPEF_FB_TRACE=<block index>: dump CTA-0 clock64 timestamps of that fused block to stderr
  bool finalized = false;
  int num_sms = 148;
  size_t smem_optin = 0;
  tc::EncodeTiledFn encode = nullptr;
  long long* trace_dev = nullptr;  // SPEF_GEMM_TRACE=<layer index>: dump CTA-0 timestamps of that layer to stderr
  int trace_layer = -1;
  int image_u8 = 0;    // SPEF_IMG_U8: images are uint8, the stem divides by 255
  int stem_simt = 0;   // SPEF_STEM_SIMT=1: CUDA-core stem instead of the tcgen05 implicit GEMM (cross-check; float images only)
  int fb_debug_skip = 0;   // SPEF_FB_DEBUG_SKIP: timing experiments of the staged fused kernel (wrong results)
  int fbt_no_stack = 0;    // SPEF_FBT_NO_STACK=1: no strip stacking in the channel-lane plan
  int head_wide = 0;       // SPEF_HEAD_WIDE=1: 256-column tiles for the head GEMM
  // Programmatic dependent launch along the forward chain (common.cuh; SPEF_PDL=0: plain stream order).  The kernels trigger their
  // successor at once only for batches that leave SMs idle (SPEF_PDL_MAX_BATCH).  Measured: one stream of single frames 0.292 -> 0.278 ms,
  // 64 streams 88.5 k -> 94.6 k frames/s; at batch 256 the early trigger costs 1.8 % (the waiting successor takes the SMs the other
  // lane's kernel would have filled the tail with) while the attribute alone gains 1.4 % (166.5 k -> 168.8 k images/s).
  int pdl = 1, pdl_max_batch = 96;
  bool pdl_on = false, pdl_early = false;   // this call's launches: attribute / explicit early trigger (set per entry point from the batch)
  int pool_fuse = 1;       // last 1x1 conv + global average pool as one kernel (conv_pool.cuh); SPEF_POOL_FUSE=0: two launches
  CUtensorMap cp_tmW, cp_tmX;
  bool cp_w_ready = false;
  const void* cp_x_ptr = nullptr;
  int cp_x_batch = -1;
  int gemm_nsw = 4;    // store warps of the v2 epilogue (SPEF_GEMM_NSW = 4 | 8; 8 only with one drain group)
  int gemm_ndg = 2;    // drain groups of the v2 epilogue (SPEF_GEMM_NDG = 1 | 2)
  size_t esz = 2;
  // activations
  void* act[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t act_elems[4] = {0, 0, 0, 0};
  void* pooled = nullptr;      // [max_batch, 1280] activation type
  float* head_out = nullptr;   // [max_batch, n_pad] f32
  int head_pad = 0;
  int plan_batch = -1;
  // tables
  double* ori_tab64 = nullptr;   // [n][4] float64 bins for the encode kernels (the reference encodes in float64)
  double* pos_tab64 = nullptr;
  float4* ori_tab = nullptr;
  float* ori_tab_soa = nullptr;  // [4][ori_tab_ld]: one plane per quaternion component, zero padded (decode_ori_stream_kernel)
  int ori_tab_ld = 0;
  int rz_gray_kernel = 1;        // greyscale fast path of spef_resize_frames (SPEF_RESIZE_GRAY=0: generic kernel)
  int dw_small_plan = 1;         // task-filling tile plans for the 15x24 / 8x12 depthwise layers (SPEF_DW_SMALL=0: first plan)
  int decode_cfg = -1;           // -1: by batch size (launch_decode_stream)
  int decode_stream = 1;         // spef_decode_ori takes the streaming kernel for 16-byte aligned rows (SPEF_DECODE_STREAM=0: decode_ori_kernel)
  int ori_n = 0;
  float4* pos_tab = nullptr;
  int pos_n = 0;
  // evaluation accumulators
  double* eval_sums = nullptr;  // [8]
  // workspaces for the *_host entry points and fused predict
  float* ws_images = nullptr;
  // pipelined host evaluation: two staging slots (images + targets + per-image results), a copy stream and events
  float* pipe_images[2] = {nullptr, nullptr};
  float* pipe_qt[2] = {nullptr, nullptr};
  float* pipe_tt[2] = {nullptr, nullptr};
  float* pipe_per[2] = {nullptr, nullptr};
  cudaStream_t pipe_copy_stream = nullptr;
  cudaEvent_t pipe_copied[2] = {nullptr, nullptr};   // H2D of slot s complete
  cudaEvent_t pipe_consumed[2] = {nullptr, nullptr}; // compute that read slot s complete
  int pipe_slot = 0;
  long long pipe_submitted = 0;
  // packed upload of float images (host_pack.cpp): the host rounds the pixels to BF16 -- the stem's own first step -- chunk by chunk
  // into a pinned staging slot, the chunks cross the bus at half the bytes and a widening kernel on the copy stream restores the
  // float tensor the stem reads.  0: plain copy.  Default 1 on the BF16 tcgen05 engine (SPEF_HOST_PACK=0 / spef_set_host_pack).
  int host_pack = 0;
  uint16_t* pack_host[2] = {nullptr, nullptr};   // pinned, [max_batch * 3 * H * W] bf16
  uint16_t* pack_dev[2] = {nullptr, nullptr};
  cudaEvent_t pack_uploaded[2] = {nullptr, nullptr};   // the DMA engine is done reading pack_host[s]
  cudaEvent_t pack_t0[2] = {nullptr, nullptr}, pack_t1[2] = {nullptr, nullptr};   // around the plain slice's copy (upload_packed)
  size_t pack_plain_bytes[2] = {0, 0};
  double pack_rc = 60e9, pack_rd = 50e9;   // measured while running: float bytes / s the host threads convert, bytes / s of a plain copy
  double pack_frac = -1.0;                 // SPEF_PACK_FRAC (developer): fixed packed fraction instead of the balance
  double pack_last_frac = 0.0;
  int pack_on = 1;                         // the measured rates say packing pays (upload_packed)
  unsigned long long pack_calls = 0;
  float* ws_quat = nullptr;     // [max_batch,4]
  float* ws_pos = nullptr;      // [max_batch,3]
  float* ws_qt = nullptr;       // [max_batch,4]
  float* ws_tt = nullptr;       // [max_batch,3]
  float* ws_soft = nullptr;     // [max_batch, max(n_ori,n_pos)] generic pdf/logit staging (lazily grown)
  size_t ws_soft_elems = 0;
  float* ws_soft2 = nullptr;
  size_t ws_soft2_elems = 0;
  float* ws_hinv = nullptr;
  float* ws_per_image = nullptr;  // [max_batch,2]
  int32_t* ws_argmax = nullptr;
  uint32_t* ws_flags = nullptr;
  double* ws_sums = nullptr;      // [8]
  // temporal state
  int t_streams = 0;
  float* t_ori_state = nullptr;
  float* t_pos_state = nullptr;
  int* t_has = nullptr;          // [4][S]: ori filter, pos filter, prev still, prev video
  float* t_prev_still = nullptr; // [S,4]
  float* t_prev_video = nullptr;
  float* t_ws[8] = {nullptr};    // scratch outputs when the caller passes NULL
  // spef_temporal_step as a CUDA graph (batch-1 / few-stream latency: ~40 launches per frame are launch-bound): one instantiated
  // graph per call signature (image pointer, n_streams, apply_filter, output pointers), captured the second time a signature is seen
  struct TGraph { const void* img; int S, filt; spef_temporal_out out; cudaGraphExec_t exec; int64_t launches; };
  std::vector<TGraph> tgraphs;
  TGraph tg_last{};                // signature of the previous direct (non-graph) call
  cudaStream_t cap_stream = nullptr;
  int temporal_graph = 1;          // SPEF_TEMPORAL_GRAPH=0 disables
  // spef_eval_batch as a CUDA graph: the ~35 launches of a step replayed as one graph save the launch gaps (B = 256: 1.92 -> 1.87 ms
  // per step, B = 64: 0.74 -> 0.69 ms).  One graph per call signature, captured when a signature comes back (the host-buffer
  // routes call with the ctx's own one or two device buffers); SPEF_EVAL_GRAPH=0 disables
  struct EGraph { const void *img, *qt, *tt, *per; int B; cudaGraphExec_t exec; int64_t launches; };
  std::vector<EGraph> egraphs;
  std::vector<EGraph> eg_seen;     // signatures of recent direct calls
  int eval_graph = 1;
  int temporal_graph_max_streams = 256;   // SPEF_TEMPORAL_GRAPH_MAX: the frame step of up to this many streams is replayed as a graph (64 streams: 7 % faster)
  // ingest plan (spef_resize_frames): taps of both axes for the last (src_h, src_w) seen, one device allocation
  int rz_fixed = 0, rz_gray = 0, rz_gray_off = 0;
  int rz_sh = 0, rz_sw = 0, rz_hks = 0, rz_vks = 0, rz_band = 0, rz_max_rows = 0, rz_pitch = 0;
  int* rz_tab = nullptr;
  // bookkeeping
  int64_t launches = 0;
  std::vector<cudaEvent_t> events;
};

static std::string g_create_err;
static cudaError_t decode_stream_init();

static int fail(spef_ctx* c, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_err = buf;
  return code;
}

#define CK(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) return fail(ctx, SPEF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CK_LAUNCH(name)                                                                                 \
  do {                                                                                                  \
    cudaError_t e_ = cudaGetLastError();                                                                \
    if (e_ != cudaSuccess) return fail(ctx, SPEF_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    ctx->launches++;                                                                                    \
  } while (0)


// Launch with programmatic stream serialization (PDL): the kernel may begin while its predecessor in the stream is still draining; it
// calls griddepcontrol.wait before touching anything the predecessor wrote (common.cuh).  Only kernels that do so are launched this way.
template <typename... KArgs, typename... Args>
static cudaError_t launch_chain(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ------------------------------------------------------------------------------------------------------
// topology
// ------------------------------------------------------------------------------------------------------
static void build_layers(spef_ctx* ctx) {
  const int settings[7][4] = {{1, 16, 1, 1}, {6, 24, 2, 2}, {6, 32, 3, 2}, {6, 64, 4, 2}, {6, 96, 3, 1}, {6, 160, 3, 2}, {6, 320, 1, 1}};
  std::vector<Layer>& L = ctx->layers;
  L.clear();
  ctx->blocks.clear();
  int H = ctx->cfg.img_h, W = ctx->cfg.img_w;
  auto down = [](int x) { return (x + 2 - 3) / 2 + 1; };
  {
    Layer s;
    s.kind = K_STEM; s.prefix = "features.features.0"; s.cin = 3; s.cout = 32; s.hin = H; s.win = W;
    s.hout = down(H); s.wout = down(W); s.stride = 2; s.relu = 1; s.src = BUF_IMG; s.dst = BUF_P;
    L.push_back(s);
    H = s.hout; W = s.wout;
  }
  int cur = BUF_P, cin = 32, idx = 1;
  for (int g = 0; g < 7; ++g) {
    const int t = settings[g][0], c = settings[g][1], n = settings[g][2], s = settings[g][3];
    for (int i = 0; i < n; ++i, ++idx) {
      const int stride = (i == 0) ? s : 1;
      const int hidden = cin * t;
      const bool res = (stride == 1 && cin == c);
      char pre[64];
      int j = 0, src = cur;
      Block blk;
      blk.first = (int)L.size();
      if (t != 1) {
        blk.i_exp = (int)L.size();
        Layer e;
        snprintf(pre, sizeof(pre), "features.features.%d.conv.%d", idx, j++);
        e.kind = K_PW; e.prefix = pre; e.cin = cin; e.cout = hidden; e.hin = e.hout = H; e.win = e.wout = W;
        e.relu = 1; e.src = src; e.dst = BUF_H1;
        L.push_back(e);
        src = BUF_H1;
      }
      Layer d;
      snprintf(pre, sizeof(pre), "features.features.%d.conv.%d", idx, j++);
      d.kind = K_DW; d.prefix = pre; d.cin = d.cout = hidden; d.hin = H; d.win = W; d.stride = stride;
      d.hout = (stride == 2) ? down(H) : H; d.wout = (stride == 2) ? down(W) : W; d.relu = 1; d.src = src; d.dst = BUF_H2;
      blk.i_dw = (int)L.size();
      L.push_back(d);
      H = d.hout; W = d.wout;
      Layer p;
      snprintf(pre, sizeof(pre), "features.features.%d.conv.%d", idx, j++);
      p.kind = K_PW; p.prefix = pre; p.cin = hidden; p.cout = c; p.hin = p.hout = H; p.win = p.wout = W; p.relu = 0;
      p.residual = res ? 1 : 0; p.src = BUF_H2; p.dst = (cur == BUF_P) ? BUF_Q : BUF_P; p.res_buf = res ? cur : -1;
      blk.i_proj = (int)L.size();
      L.push_back(p);
      blk.n_layers = (int)L.size() - blk.first;
      ctx->blocks.push_back(blk);
      cur = p.dst;
      cin = c;
    }
  }
  {
    Layer l;
    l.kind = K_PW; l.prefix = "features.features.18"; l.cin = cin; l.cout = 1280; l.hin = l.hout = H; l.win = l.wout = W;
    l.relu = 1; l.src = cur; l.dst = BUF_H1;
    L.push_back(l);
  }
  {
    Layer pl;
    pl.kind = K_POOL; pl.cin = pl.cout = 1280; pl.hin = H; pl.win = W; pl.hout = pl.wout = 1; pl.src = BUF_H1; pl.dst = BUF_POOL;
    L.push_back(pl);
  }
  {
    Layer h;
    h.kind = K_HEAD; h.cin = 1280; h.cout = ctx->cfg.n_ori + ctx->cfg.n_pos; h.hin = h.win = h.hout = h.wout = 1;
    h.src = BUF_POOL; h.dst = BUF_HEAD;
    L.push_back(h);
  }
  for (Layer& l : L) l.n_pad = (l.kind == K_HEAD) ? ((l.cout + 7) / 8) * 8 : l.cout;
}

static void* buf_ptr(spef_ctx* ctx, int id) {
  if (id >= 0 && id < 4) return ctx->act[id];
  if (id == BUF_POOL) return ctx->pooled;
  if (id == BUF_HEAD) return ctx->head_out;
  return nullptr;
}

// ------------------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------------------
// instantiated graphs hold weight / table / state pointers and the kernel choice of the moment they were captured
static void drop_graphs(spef_ctx* ctx) {
  for (auto& g : ctx->tgraphs) cudaGraphExecDestroy(g.exec);
  ctx->tgraphs.clear();
  ctx->tg_last = spef_ctx::TGraph{};
  for (auto& g : ctx->egraphs) cudaGraphExecDestroy(g.exec);
  ctx->egraphs.clear();
  ctx->eg_seen.clear();
}

extern "C" int spef_abi_version(void) { return SPEF_ABI_VERSION; }

extern "C" const char* spef_last_error(const spef_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" int spef_create(spef_ctx** out, const spef_config* cfg) {
  spef_ctx* ctx = nullptr;  // for CK(): errors go to g_create_err until the ctx exists
  if (!out || !cfg) return fail(nullptr, SPEF_ERR_INVALID, "spef_create: null argument");
  if (cfg->struct_size != (int32_t)sizeof(spef_config)) return fail(nullptr, SPEF_ERR_INVALID, "spef_create: struct_size %d != %d", cfg->struct_size, (int)sizeof(spef_config));
  if (cfg->img_h < 32 || cfg->img_w < 32 || cfg->n_ori < 1 || cfg->n_pos < 1 || cfg->max_batch < 1)
    return fail(nullptr, SPEF_ERR_INVALID, "spef_create: bad geometry (img %dx%d, n_ori %d, n_pos %d, max_batch %d)", cfg->img_h, cfg->img_w, cfg->n_ori, cfg->n_pos, cfg->max_batch);
  if (cfg->precision != SPEF_FP32 && cfg->precision != SPEF_BF16) return fail(nullptr, SPEF_ERR_INVALID, "spef_create: precision must be SPEF_FP32 or SPEF_BF16");
  if (!cfg->pos_classification && cfg->n_pos != 3) return fail(nullptr, SPEF_ERR_INVALID, "spef_create: regression position head must have n_pos = 3");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, SPEF_ERR_CUDA, "spef_create: no CUDA device (%s); libspef_b200 has no CPU fallback", cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, SPEF_ERR_INVALID, "spef_create: device %d out of range (%d devices)", cfg->device, ndev);
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail(nullptr, SPEF_ERR_UNSUPPORTED, "spef_create: device is sm_%d%d; this library is built for sm_100a (B200) only", prop.major, prop.minor);

  ctx = new spef_ctx();
  ctx->cfg = *cfg;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  ctx->esz = (cfg->precision == SPEF_BF16) ? 2 : 4;
  if (const char* e4 = getenv("SPEF_GEMM_TRACE")) { ctx->trace_layer = atoi(e4); cudaMalloc((void**)&ctx->trace_dev, 256 * 16 * sizeof(long long)); }
  if (const char* e17 = getenv("SPEF_RESIZE_GRAY")) ctx->rz_gray_kernel = atoi(e17);
  if (const char* e15 = getenv("SPEF_DW_SMALL")) ctx->dw_small_plan = atoi(e15);
  if (const char* e13 = getenv("SPEF_DECODE_CFG")) ctx->decode_cfg = atoi(e13);
  if (const char* e12 = getenv("SPEF_DECODE_STREAM")) ctx->decode_stream = atoi(e12);
  if (const char* e5 = getenv("SPEF_STEM_SIMT")) ctx->stem_simt = atoi(e5) ? 1 : 0;
  if (const char* e5 = getenv("SPEF_FB_DEBUG_SKIP")) ctx->fb_debug_skip = atoi(e5);
  if (getenv("SPEF_FBT_NO_STACK")) ctx->fbt_no_stack = 1;
  ctx->host_pack = (cfg->precision == SPEF_BF16 && cfg->pw_impl == 0) ? 1 : 0;
  if (const char* e = getenv("SPEF_HOST_PACK")) ctx->host_pack = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SPEF_PACK_FRAC")) ctx->pack_frac = atof(e);
  // first guess of the conversion rate: the host's threads are shared by the ranks of the node (torchrun sets LOCAL_WORLD_SIZE), so eight
  // ranks start with the packed upload off and let the probe slices decide, one or two ranks start with it on
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) { const int r = atoi(e); if (r > 1) ctx->pack_rc /= (double)r; }
  if (getenv("SPEF_HEAD_WIDE")) ctx->head_wide = 1;
  if (const char* e = getenv("SPEF_POOL_FUSE")) ctx->pool_fuse = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SPEF_PDL")) ctx->pdl = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SPEF_DWP_OPT_SKIP")) ctx->dwp_opt_skip = atoi(e);
  if (const char* e = getenv("SPEF_DWP_MIN_TILES")) ctx->dwp_min_tiles = atoi(e);
  if (const char* e = getenv("SPEF_DWP_S2_BOX_KB")) ctx->dwp_s2_box_kb = atoi(e);
  if (const char* e = getenv("SPEF_PDL_MAX_BATCH")) ctx->pdl_max_batch = atoi(e);
  if (const char* e5 = getenv("SPEF_TEMPORAL_GRAPH")) ctx->temporal_graph = atoi(e5) ? 1 : 0;
  if (const char* e5 = getenv("SPEF_TEMPORAL_GRAPH_MAX")) ctx->temporal_graph_max_streams = atoi(e5);
  if (const char* e7 = getenv("SPEF_GEMM_NDG")) ctx->gemm_ndg = (atoi(e7) == 1) ? 1 : 2;
  if (const char* e6 = getenv("SPEF_GEMM_NSW")) ctx->gemm_nsw = (atoi(e6) == 8 && ctx->gemm_ndg == 1) ? 8 : 4;
  if (const char* e8 = getenv("SPEF_FUSE")) ctx->fuse = atoi(e8) ? 1 : 0;
  if (const char* e8 = getenv("SPEF_DWP")) ctx->dwp_enable = atoi(e8) ? 1 : 0;
  if (const char* e8 = getenv("SPEF_DWP_FORCE")) ctx->dwp_force = atoi(e8) ? 1 : 0;
  if (const char* e8 = getenv("SPEF_DWP_WST")) ctx->dwp_w_stages = atoi(e8);
  if (const char* e8 = getenv("SPEF_EVAL_GRAPH")) ctx->eval_graph = atoi(e8) ? 1 : 0;
  if (const char* e9 = getenv("SPEF_FB_GW")) ctx->fb_gw = (atoi(e9) == 8) ? 8 : 4;
  if (const char* e11 = getenv("SPEF_FB_MAX_CIN")) ctx->fb_max_cin = atoi(e11);
  if (const char* e12 = getenv("SPEF_FB_VARIANT")) ctx->fb_variant = atoi(e12) ? 1 : 0;
  if (const char* e15 = getenv("SPEF_STEM_PATCH")) ctx->stem_patch = atoi(e15) ? 1 : 0;
  if (const char* e14 = getenv("SPEF_STEM_PROD")) { int v = atoi(e14); ctx->stem_prod = (v == 1 || v == 4) ? v : 2; }
  if (const char* e16 = getenv("SPEF_FBT_A2")) ctx->fbt_a2_bufs = (atoi(e16) == 1) ? 1 : 2;
  if (const char* e20 = getenv("SPEF_STEM_FUSE")) ctx->stem_fuse = atoi(e20) ? 1 : 0;
  if (const char* e19 = getenv("SPEF_FBT_PSTAGES")) { int v = atoi(e19); ctx->fbt_max_pstages = (v >= 1 && v <= fbt::MAX_PROJ) ? v : fbt::MAX_PROJ; }
  if (const char* e10 = getenv("SPEF_FB_TRACE")) { ctx->fb_trace_block = atoi(e10); if (!ctx->trace_dev) cudaMalloc((void**)&ctx->trace_dev, 256 * 16 * sizeof(long long)); }
  build_layers(ctx);

  // cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    delete ctx;
    return fail(nullptr, SPEF_ERR_CUDA, "spef_create: cuTensorMapEncodeTiled not available from the driver");
  }
  ctx->encode = (tc::EncodeTiledFn)fn;

  // activation buffers (elements per image): P/Q = block in/out, H1 = expand out / last conv out, H2 = depthwise out
  size_t need[4] = {0, 0, 0, 0};
  for (const Layer& l : ctx->layers) {
    if (l.dst >= 0 && l.dst < 4) {
      size_t n = (size_t)l.hout * l.wout * l.cout;
      if (n > need[l.dst]) need[l.dst] = n;
    }
  }
  const size_t B = (size_t)cfg->max_batch;
  auto dalloc = [&](void** p, size_t bytes) -> bool { return cudaMalloc(p, bytes ? bytes : 16) == cudaSuccess; };
  bool ok = true;
  for (int i = 0; i < 4; ++i) {
    ctx->act_elems[i] = need[i] * B;
    ok = ok && dalloc(&ctx->act[i], ctx->act_elems[i] * ctx->esz);
  }
  ctx->head_pad = ctx->layers.back().n_pad;
  ok = ok && dalloc(&ctx->pooled, B * 1280 * ctx->esz);
  ok = ok && dalloc((void**)&ctx->head_out, B * ctx->head_pad * sizeof(float));
  ok = ok && dalloc((void**)&ctx->eval_sums, 8 * sizeof(double));
  ok = ok && dalloc((void**)&ctx->ws_quat, B * 4 * sizeof(float));
  ok = ok && dalloc((void**)&ctx->ws_pos, B * 3 * sizeof(float));
  ok = ok && dalloc((void**)&ctx->ws_qt, B * 4 * sizeof(float));
  ok = ok && dalloc((void**)&ctx->ws_tt, B * 3 * sizeof(float));
  ok = ok && dalloc((void**)&ctx->ws_hinv, B * 16 * sizeof(float));
  ok = ok && dalloc((void**)&ctx->ws_per_image, B * 2 * sizeof(float));
  ok = ok && dalloc((void**)&ctx->ws_argmax, B * sizeof(int32_t));
  ok = ok && dalloc((void**)&ctx->ws_flags, B * sizeof(uint32_t));
  ok = ok && dalloc((void**)&ctx->ws_sums, 8 * sizeof(double));
  if (!ok) {
    std::string m = std::string("spef_create: cudaMalloc failed: ") + cudaGetErrorString(cudaGetLastError());
    spef_destroy(ctx);
    return fail(nullptr, SPEF_ERR_CUDA, "%s", m.c_str());
  }
  cudaMemset(ctx->eval_sums, 0, 8 * sizeof(double));
  if (decode_stream_init() != cudaSuccess) {
    spef_destroy(ctx);
    return fail(nullptr, SPEF_ERR_CUDA, "spef_create: cannot reserve shared memory for the decode kernel");
  }
  *out = ctx;
  return SPEF_OK;
}

extern "C" void spef_destroy(spef_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  for (Layer& l : ctx->layers) {
    cudaFree(l.w_f32);
    cudaFree(l.w_bf16);
    cudaFree(l.bias);
  }
  for (Block& b : ctx->blocks) { cudaFree(b.aux); cudaFree(b.t_we); cudaFree(b.t_wp); cudaFree(b.t_aux); cudaFree(b.dp_wdw); cudaFree(b.s_we); cudaFree(b.s_aux); }
  for (int i = 0; i < 4; ++i) cudaFree(ctx->act[i]);
  void* ptrs[] = {ctx->pooled, ctx->head_out, ctx->ori_tab64, ctx->pos_tab64, ctx->ori_tab, ctx->pos_tab, ctx->eval_sums, ctx->ws_images, ctx->ws_quat,
                  ctx->ws_pos, ctx->ws_qt, ctx->ws_tt, ctx->ws_soft, ctx->ws_soft2, ctx->ws_hinv, ctx->ws_per_image,
                  ctx->ws_argmax, ctx->ws_flags, ctx->ws_sums, ctx->t_ori_state, ctx->t_pos_state, ctx->t_has,
                  ctx->t_prev_still, ctx->t_prev_video, ctx->rz_tab, ctx->ori_tab_soa};
  for (void* p : ptrs) cudaFree(p);
  for (int i = 0; i < 8; ++i) cudaFree(ctx->t_ws[i]);
  for (cudaEvent_t ev : ctx->events) cudaEventDestroy(ev);
  for (int s = 0; s < 2; ++s) {
    if (ctx->pack_host[s]) cudaFreeHost(ctx->pack_host[s]);
    cudaFree(ctx->pack_dev[s]);
    if (ctx->pack_uploaded[s]) cudaEventDestroy(ctx->pack_uploaded[s]);
    if (ctx->pack_t0[s]) cudaEventDestroy(ctx->pack_t0[s]);
    if (ctx->pack_t1[s]) cudaEventDestroy(ctx->pack_t1[s]);
    cudaFree(ctx->pipe_images[s]); cudaFree(ctx->pipe_qt[s]); cudaFree(ctx->pipe_tt[s]); cudaFree(ctx->pipe_per[s]);
    if (ctx->pipe_copied[s]) cudaEventDestroy(ctx->pipe_copied[s]);
    if (ctx->pipe_consumed[s]) cudaEventDestroy(ctx->pipe_consumed[s]);
  }
  if (ctx->pipe_copy_stream) cudaStreamDestroy(ctx->pipe_copy_stream);
  drop_graphs(ctx);
  if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
  delete ctx;
}

// ------------------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------------------
extern "C" int spef_load_tensor(spef_ctx* ctx, const char* key, const float* data, const int64_t* shape, int32_t ndim) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (!key || !data || ndim < 0 || ndim > 4 || (ndim > 0 && !shape)) return fail(ctx, SPEF_ERR_INVALID, "spef_load_tensor: bad argument");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] < 0) return fail(ctx, SPEF_ERR_INVALID, "spef_load_tensor(%s): negative dimension", key);
    t.shape.push_back(shape[i]);
    n *= (size_t)shape[i];
  }
  t.data.assign(data, data + n);
  ctx->host_tensors[key] = std::move(t);
  ctx->finalized = false;
  return SPEF_OK;
}

static const HostTensor* find_tensor(spef_ctx* ctx, const std::string& key, std::initializer_list<int64_t> shape) {
  auto it = ctx->host_tensors.find(key);
  if (it == ctx->host_tensors.end()) {
    fail(ctx, SPEF_ERR_STATE, "spef_finalize_weights: missing state_dict tensor '%s'", key.c_str());
    return nullptr;
  }
  if (it->second.shape != std::vector<int64_t>(shape)) {
    std::string got;
    for (int64_t d : it->second.shape) got += std::to_string(d) + ",";
    std::string want;
    for (int64_t d : shape) want += std::to_string(d) + ",";
    fail(ctx, SPEF_ERR_INVALID, "spef_finalize_weights: tensor '%s' has shape [%s], expected [%s]", key.c_str(), got.c_str(), want.c_str());
    return nullptr;
  }
  return &it->second;
}

static inline float bf16_round_host(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <typename T>
static bool upload(T** dst, const std::vector<T>& src) {
  if (*dst) { cudaFree(*dst); *dst = nullptr; }
  if (cudaMalloc((void**)dst, src.size() * sizeof(T)) != cudaSuccess) return false;
  return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}

// Fused InvertedResidual plan (fused_block.cuh): tile shape, shared-memory stages, per-chunk aux array.
static int plan_blocks(spef_ctx* ctx) {
  std::vector<Layer>& L = ctx->layers;
  const size_t limit = ctx->smem_optin;
  for (Block& b : ctx->blocks) {
    b.fusable = false;
    b.tmW_ready = false;
    b.tmX_ptr = nullptr;
    const bool has_exp = b.i_exp >= 0;   // t = 1 block (no expand conv): the x tile is the hidden tile
    const Layer& d = L[b.i_dw];
    const Layer& e = has_exp ? L[b.i_exp] : d;
    const Layer& pj = L[b.i_proj];
    fb::FbParams& q = b.prm;
    memset(&q, 0, sizeof(q));
    q.H = e.hin; q.W = e.win; q.Cin = e.cin; q.Ch = has_exp ? e.cout : d.cin; q.Cout = pj.cout; q.Ho = d.hout; q.Wo = d.wout;
    if (!has_exp && q.Ch > fb::HC) continue;
    const int S = d.stride;
    fb::pick_tile(q.Ho, q.Wo, S, &q.TH, &q.TW);
    q.THI = (q.TH - 1) * S + 3; q.TWI = (q.TW - 1) * S + 3;
    q.tiles_y = cdiv(q.Ho, q.TH); q.tiles_x = cdiv(q.Wo, q.TW);
    q.kc_in = cdiv(q.Cin, 64); q.n_chunks = cdiv(q.Ch, fb::HC); q.cpad = ((q.Cout + 15) / 16) * 16;
    if (q.Cin > ctx->fb_max_cin) continue;
    if (q.cpad > 256) continue;  // project accumulator must fit one MMA N and the TMEM columns left of the expand stages
    q.proj_stages = (q.cpad <= 128) ? 2 : 1;
    // TMEM columns: expand stages of 128 columns each, then the project accumulator stages
    q.n_acc = (q.cpad <= 64) ? 3 : 2;
    q.proj_col0 = q.n_acc * 128;
    q.proj_stride = (q.proj_stages == 2) ? ((q.cpad <= 64) ? 64 : 128) : 0;
    q.residual = pj.residual; q.has_expand = has_exp ? 1 : 0;
    // shared-memory plan, best first: two worker groups before one, resident weights before a ring, two x stages before one.
    // A weight ring needs >= NG + 1 stages: the MMA thread issues expand(n + NG) before project(n) frees the stage of chunk n.
    bool found = false;
    for (int ng = 2; ng >= 1 && !found; --ng) {
      struct Opt { int w, res, x; };
      std::vector<Opt> opts;
      if (q.n_chunks <= fb::MAX_W_STAGES) { opts.push_back({q.n_chunks, 1, 2}); if (has_exp) opts.push_back({q.n_chunks, 1, 1}); }
      if (q.n_chunks > ng + 1) {
        opts.push_back({ng + 2, 0, 2}); opts.push_back({ng + 1, 0, 2}); opts.push_back({ng + 2, 0, 1}); opts.push_back({ng + 1, 0, 1});
      }
      for (const Opt& o : opts) {
        q.w_stages = o.w; q.resident = o.res; q.x_stages = o.x;
        if (fb::smem_bytes(q, ng) <= limit) { found = true; b.ng = ng; b.smem = fb::smem_bytes(q, ng); break; }
      }
    }
    if (!found) continue;
    std::vector<float> aux((size_t)q.n_chunks * fb::AUX_FLOATS, 0.f);
    for (int ch = 0; ch < q.Ch; ++ch) {
      float* a = aux.data() + (size_t)(ch / fb::HC) * fb::AUX_FLOATS;
      const int j = ch % fb::HC;
      a[j] = has_exp ? e.h_bias[ch] : 0.f;
      a[fb::HC + j] = d.h_bias[ch];
      for (int k = 0; k < 9; ++k) a[(2 + k) * fb::HC + j] = d.h_wdw[(size_t)k * q.Ch + ch];
    }
    if (!upload(&b.aux, aux)) return fail(ctx, SPEF_ERR_CUDA, "spef_finalize_weights: upload failed");
    b.fusable = true;
  }
  return SPEF_OK;
}

// Channel-lane fused plan (fused_block_t.cuh): tile rows, TMEM / shared-memory stages and the permuted weight arrays.
static int plan_blocks_t(spef_ctx* ctx) {
  std::vector<Layer>& L = ctx->layers;
  const size_t limit = ctx->smem_optin;
  for (Block& b : ctx->blocks) {
    b.t_ok = false;
    b.t_tmW_ready = false;
    b.t_tmX_ptr = nullptr;
    const bool has_exp = b.i_exp >= 0;
    const Layer& d = L[b.i_dw];
    const Layer& e = has_exp ? L[b.i_exp] : d;
    const Layer& pj = L[b.i_proj];
    const int S = d.stride, TW = (S == 1) ? 12 : 6, TWI = (TW - 1) * S + 3;
    const int Ch = has_exp ? e.cout : d.cin;
    // Strip stacking (fused_block_t.cuh): t = 1 block -> identity "expand", four strips in the four TMEM lane quarters (needs Ch == 32,
    // stride 1, no skip, Cout <= 16); blocks whose hidden width wastes lanes at 128 channels per pass (144, 192) -> two strips of 64 slots
    int stack = 1;
    if (!has_exp) {
      if (!(Ch == 32 && S == 1 && !pj.residual && pj.cout <= 16 && !ctx->fbt_no_stack)) continue;
      stack = 4;
    } else if (S == 1 && e.cin <= 64 && !ctx->fbt_no_stack && cdiv(Ch, 64) < 2 * cdiv(Ch, fbt::CL) && d.wout % (TW * 2) == 0) {
      // (stride-2 blocks measured no faster stacked: their items are dominated by the conversion of the 4x larger hidden tile)
      stack = 2;
    }
    const int LS = fbt::CL / stack;                  // channel slots per strip
    if (d.wout % (TW * stack) != 0 || Ch % 4 != 0) continue;   // exact tiling in x: only the two halo columns can leave the image
    fbt::FbtParams& q = b.tprm;
    memset(&q, 0, sizeof(q));
    q.H = e.hin; q.W = e.win; q.Cin = (stack > 1) ? stack * 64 : e.cin; q.Cout = pj.cout; q.Ho = d.hout; q.Wo = d.wout;
    q.cx = e.cin; q.stack = stack; q.kst_stack = has_exp ? cdiv(e.cin, 16) : 2;
    q.TW = TW; q.TWI = TWI;
    long long best = -1;
    for (int TH = 4; TH <= 6; ++TH) {              // instantiated kernels: S = 1: TH 5..6, S = 2: TH 4
      const int thi = (TH - 1) * S + 3;
      if ((S == 1 && TH < 5) || (S == 2 && TH > 4)) continue;
      if (thi * TWI > 128 || TH * TW > 128) continue;
      const long long cost = (long long)cdiv(q.Ho, TH) * thi;   // hidden rows computed per image column of tiles
      if (best < 0 || cost < best) { best = cost; q.TH = TH; }
    }
    if (best < 0) continue;
    q.THI = (q.TH - 1) * S + 3;
    q.tiles_y = cdiv(q.Ho, q.TH); q.tiles_x = q.Wo / (TW * stack);
    q.n_px = ((q.THI * TWI + 15) / 16) * 16;
    q.kc_in = cdiv(q.Cin, 64); q.n_chunks = cdiv(Ch, LS); q.cpad = ((q.Cout + 15) / 16) * 16;
    const int we_rows = LS * (2 * stack - 1);       // window matrix [zeros | chunk | zeros]: 224 (stack 4), 192 (stack 2)
    q.we_bytes = (stack > 1) ? we_rows * 128 : q.kc_in * fbt::CL * 128;
    // stack = 2: the chunks share their zero blocks in one resident region [Z | C0 | Z | C1 | ... | Z] (fused_block_t.cuh, wz_bytes)
    const bool we_shared = (stack == 2 && q.n_chunks >= 2 && q.n_chunks <= fbt::MAX_W_STAGES);
    if (we_shared) { q.wz_bytes = (2 * q.n_chunks + 1) * 64 * 128; q.we_bytes = 0; }
    if (q.cpad > 128 || q.n_chunks > fbt::MAX_W_STAGES) continue;
    q.residual = pj.residual;
    // worker groups, TMEM expand stages (n_px columns each, one more than groups when they fit) and the project
    // accumulator(s) behind them; shared memory: resident weights before a ring, as many x stages as fit
    bool found = false;
    for (int ng = 2; ng >= 2 && !found; --ng) {   // (three worker groups never fit next to four TMEM expand stages with the shipped tile shapes)
      // stacked: the per-strip accumulators sit cpad columns apart (the epilogue's x32 load over-reads into the next one)
      const int pcols = (stack > 1) ? ((q.cpad * stack + 31) / 32) * 32 : ((q.cpad + 31) / 32) * 32;
      q.proj_sub = (stack == 2) ? ((q.cpad + 31) / 32) * 32 : q.cpad;   // (the epilogue reads 32 columns at a time when Cout > 16)
      int n_acc = 0, pstages = 0;
      for (int na = ng + 1; na >= ng && !n_acc; --na)
        for (int ps = ctx->fbt_max_pstages; ps >= 1 && !n_acc; --ps)
          if (na * q.n_px + ps * pcols <= 512) { n_acc = na; pstages = ps; }
      if (!n_acc) continue;
      q.n_acc = n_acc; q.acc_stride = q.n_px; q.proj_col0 = n_acc * q.n_px; q.proj_stages = pstages; q.proj_stride = (pstages >= 2) ? pcols : 0;
      struct Opt { int w, res, x; };
      std::vector<Opt> opts;
      // a stacked tile is four TMA boxes of 64-byte pixel rows (~2700 cycles): it must be prefetched behind the previous item
      for (int xs = 4; xs >= (stack == 4 ? 2 : 1); --xs) opts.push_back({q.n_chunks, 1, xs});
      if (q.n_chunks > 3 && !we_shared) for (int xs = 2; xs >= 1; --xs) { opts.push_back({4, 0, xs}); opts.push_back({3, 0, xs}); }
      // two A2 buffers per group where they fit next to at least two x stages (the workers then never wait for the project MMA
      // of their previous item), else one
      for (int a2b = ctx->fbt_a2_bufs; a2b >= 1 && !found; --a2b) {
        q.a2_bufs = a2b;
        for (const Opt& o : opts) {
          if (a2b == 2 && o.x < 2) continue;
          q.w_stages = o.w; q.resident = o.res; q.x_stages = o.x;
          if (fbt::smem_bytes(q, ng) <= limit) { found = true; b.t_ng = ng; b.t_smem = fbt::smem_bytes(q, ng); break; }
        }
      }
    }
    if (!found) continue;
    // channel -> (chunk, quarter, lane): every chunk spreads its channels evenly over the TMEM lane quarters of a strip; the
    // strips of a stacked tile hold the same channels (slot = strip * LS + quarter-in-strip * 32 + lane)
    const int QS = 4 / stack;                        // lane quarters per strip
    const int per_q = Ch / QS, base = per_q / q.n_chunks, rem = per_q % q.n_chunks;
    const size_t we_cols = (stack > 1) ? 64 : (size_t)q.Cin;
    std::vector<bf16> we((we_shared ? (size_t)(2 * q.n_chunks + 1) * 64 : (stack > 1 ? (size_t)q.n_chunks * we_rows : (size_t)q.n_chunks * fbt::CL)) * we_cols, __float2bfloat16_rn(0.f));
    std::vector<bf16> wp((size_t)q.Cout * q.n_chunks * fbt::CL, __float2bfloat16_rn(0.f));
    std::vector<float> aux((size_t)q.n_chunks * fbt::AUX_ROWS * fbt::CL, 0.f);
    if (Ch % QS != 0) continue;
    int ch = 0;
    for (int c = 0; c < q.n_chunks; ++c) {
      const int nvq = base + (c < rem ? 1 : 0);
      if (nvq > 32) return fail(ctx, SPEF_ERR_INVALID, "plan_blocks_t: internal error (nvq = %d)", nvq);
      float* a = aux.data() + (size_t)c * fbt::AUX_ROWS * fbt::CL;
      for (int qq = 0; qq < QS; ++qq)
        for (int l = 0; l < nvq; ++l, ++ch) {
          const int in_strip = qq * 32 + l;          // slot inside the strip's block of LS lanes
          // expand weights: one copy (the window matrix places it for every strip), identity for the t = 1 block
          if (stack > 1) {
            bf16* row = we.data() + (we_shared ? ((size_t)(2 * c + 1) * 64 + in_strip) : ((size_t)c * we_rows + (size_t)(stack - 1) * LS + in_strip)) * 64;
            if (has_exp) for (int k = 0; k < e.cin; ++k) row[k] = e.h_wb[(size_t)ch * e.cin + k];
            else row[ch] = __float2bfloat16_rn(1.f);
          } else {
            for (int k = 0; k < q.Cin; ++k) we[((size_t)c * fbt::CL + in_strip) * q.Cin + k] = e.h_wb[(size_t)ch * q.Cin + k];
          }
          for (int st = 0; st < stack; ++st) {       // per-slot project weights and depthwise constants, replicated per strip
            const int slot = st * LS + in_strip;
            for (int co = 0; co < q.Cout; ++co) wp[(size_t)co * q.n_chunks * fbt::CL + (size_t)c * fbt::CL + slot] = pj.h_wb[(size_t)co * Ch + ch];
            // the workers carry h' = h - be = max(acc, -be) (fused_block_t.cuh): row 0 = -be, row 1 = bd + be * sum(w)
            double wsum = 0.0;
            for (int k = 0; k < 9; ++k) wsum += (double)d.h_wdw[(size_t)k * Ch + ch];
            a[slot] = has_exp ? -e.h_bias[ch] : 0.f;
            a[fbt::CL + slot] = has_exp ? (float)((double)d.h_bias[ch] + (double)e.h_bias[ch] * wsum) : d.h_bias[ch];
            for (int k = 0; k < 9; ++k) a[(2 + k) * fbt::CL + slot] = d.h_wdw[(size_t)k * Ch + ch];
          }
        }
    }
    if (ch != Ch) return fail(ctx, SPEF_ERR_INVALID, "plan_blocks_t: internal error (%d of %d channels placed)", ch, Ch);
    if (!upload(&b.t_we, we) || !upload(&b.t_wp, wp) || !upload(&b.t_aux, aux)) return fail(ctx, SPEF_ERR_CUDA, "spef_finalize_weights: upload failed");
    b.t_ok = true;
  }
  return SPEF_OK;
}

// Stem fused into block 0 (the t = 1 block): the channel-lane plan of that block with the stem conv as its expand conv
// (reference: mobilenet_v2.py:252-262).  Same tiles, strips, project weights and depthwise constants; the window matrix carries the
// stem's folded BF16 weights instead of the identity, the aux rows the stem bias (fused_block_t.cuh: h' = max(acc, -be)).
static int plan_stem_block(spef_ctx* ctx) {
  if (ctx->blocks.empty() || ctx->layers.empty()) return SPEF_OK;
  Block& b = ctx->blocks[0];
  b.s_ok = false; b.s_tmW_ready = false; b.s_img_ptr = nullptr;
  const Layer& s = ctx->layers[0];
  if (!(b.t_ok && b.i_exp < 0 && b.first == 1 && b.tprm.stack == 4 && b.t_ng == 2 && b.tprm.TH == 6 && b.tprm.n_chunks == 1)) return SPEF_OK;
  if (!(s.kind == K_STEM && s.cout == 32 && s.hin == 2 * s.hout && s.win == 2 * s.wout && s.win % 16 == 0 && s.h_wb.size() == 32 * 32)) return SPEF_OK;
  const Layer& d = ctx->layers[b.i_dw];
  fbt::FbtParams q = b.tprm;
  q.stem = 1; q.x_stages = 2; q.a2_bufs = 1; q.kst_stack = 2;
  q.patch_stride = ((104 * (2 * q.THI + 1) * 3 * 4 + 1023) / 1024) * 1024;   // sized for float images (uint8 patches are smaller)
  q.patch_stages = 0;
  for (int ps = 4; ps >= 2 && !q.patch_stages; --ps) {
    q.patch_stages = ps;
    if (fbt::smem_bytes(q, 2) > ctx->smem_optin) q.patch_stages = 0;
  }
  if (!q.patch_stages) return SPEF_OK;
  const int we_rows = q.we_bytes >> 7;               // 224
  std::vector<bf16> we((size_t)we_rows * 64, __float2bfloat16_rn(0.f));
  std::vector<float> aux((size_t)fbt::AUX_ROWS * fbt::CL, 0.f);
  const int Ch = 32, LS = 32;
  for (int ch = 0; ch < Ch; ++ch) {
    bf16* row = we.data() + ((size_t)(4 - 1) * LS + ch) * 64;
    for (int k = 0; k < 27; ++k) row[k] = s.h_wb[(size_t)ch * 32 + k];
    double wsum = 0.0;
    for (int k = 0; k < 9; ++k) wsum += (double)d.h_wdw[(size_t)k * Ch + ch];
    for (int st = 0; st < 4; ++st) {
      const int slot = st * LS + ch;
      aux[slot] = -s.h_bias[ch];
      aux[fbt::CL + slot] = (float)((double)d.h_bias[ch] + (double)s.h_bias[ch] * wsum);
      for (int k = 0; k < 9; ++k) aux[(2 + k) * fbt::CL + slot] = d.h_wdw[(size_t)k * Ch + ch];
    }
  }
  if (!upload(&b.s_we, we) || !upload(&b.s_aux, aux)) return fail(ctx, SPEF_ERR_CUDA, "spef_finalize_weights: upload failed");
  b.sprm = q;
  b.s_smem = fbt::smem_bytes(q, 2);
  b.s_ok = true;
  return SPEF_OK;
}

// Depthwise -> project plan (dw_project.cuh): blocks with an expand conv whose hidden width is a multiple of 64 (stride 1, and stride 2
// without a skip connection).
static int plan_blocks_dp(spef_ctx* ctx) {
  std::vector<Layer>& L = ctx->layers;
  for (Block& b : ctx->blocks) {
    b.dp_ok = false;
    b.dp_tmW_ready = false;
    b.dp_tmX_ptr = nullptr;
    if (b.i_exp < 0) continue;
    const Layer& d = L[b.i_dw];
    const Layer& pj = L[b.i_proj];
    if ((d.stride != 1 && d.stride != 2) || d.cin % 64 != 0 || pj.cout % 32 != 0 || d.wout % 4 != 0 || d.wout > 128) continue;
    if (d.stride == 2 && (pj.residual || d.hout != (d.hin + 1) / 2 || d.wout != (d.win + 1) / 2)) continue;
    dwp::DwpParams& q = b.dprm;
    memset(&q, 0, sizeof(q));
    q.S = d.stride;
    q.H = d.hout; q.W = d.wout; q.C = d.cin; q.N = pj.cout;
    q.WB = (q.S == 1) ? q.W + 2 : 2 * q.W + 1;
    // tile = TH full-width output rows, at most 128 pixels; stride 2: the input box (2 TH + 1 rows of 2 W + 1 pixels) also has to stay
    // small enough for three stages next to the operand rings
    for (int th = 128 / q.W; th >= 1 && !q.TH; --th)
      if (q.H % th == 0 && (q.S == 1 || (2 * th + 1) * q.WB * 128 <= ctx->dwp_s2_box_kb * 1024 || th == 1)) q.TH = th;
    q.tiles_y = q.H / q.TH; q.n_px = q.TH * q.W; q.k_chunks = q.C / 64;
    // producer task = (4 channels, 2 columns, R rows): R the largest divisor of TH (<= 5) for which the tile's tasks fill a whole
    // number G of equal producer groups (G * tasks * 16 threads = all producer threads); group g computes the K chunks g, g + G, ...
    q.R = 0;
    for (int r = 5; r >= 1 && !q.R; --r) {
      if (q.TH % r != 0 || q.W % 2 != 0 || (q.S == 2 && r != 4 && r != 2 && r != 1)) continue;
      const int threads = (q.W / 2) * (q.TH / r) * 16;
      if (threads <= 32 * dwp::PROD_WARPS && threads % 32 == 0 && (32 * dwp::PROD_WARPS) % threads == 0 && 32 * dwp::PROD_WARPS / threads <= 2) { q.R = r; q.G = 32 * dwp::PROD_WARPS / threads; }
    }
    if (!q.R || q.G > q.k_chunks || q.G > 2) continue;
    q.n_half = (q.N <= 256) ? 1 : 2; q.nh = q.N / q.n_half;
    if (q.nh > 256 || q.nh % 16 != 0 || q.N > 512) continue;
    q.acc_stride = q.N; q.acc_stages = (2 * q.N <= 512) ? 2 : 1;
    q.in_bytes = ((q.S == 1) ? q.TH + 2 : 2 * q.TH + 1) * q.WB * 128; q.in_stride = ((q.in_bytes + 1023) / 1024) * 1024;
    bool found = false;
    // {A stages, project-weight stages, input stages}: a producer group holds its input stage for G chunk times, so four input boxes
    // (two of them prefetched); the weight ring is decoupled from the A stages (a deeper one, up to 6, measured no faster)
    const int opts[12][3] = {{4, 4, 4}, {3, 4, 4}, {3, 4, 3}, {3, 3, 4}, {3, 3, 3}, {2, 4, 4}, {2, 4, 3}, {2, 3, 3}, {2, 2, 4}, {2, 2, 3}, {2, 3, 2}, {2, 2, 2}};
    const int opt_skip = ctx->dwp_opt_skip;   // SPEF_DWP_OPT_SKIP (developer): skip the first n candidates
    int oi = 0;
    for (const auto& o : opts) {
      if (oi++ < opt_skip) continue;
      q.ab_stages = o[0]; q.w_stages = o[1]; q.in_stages = o[2];
      if (ctx->dwp_w_stages > 0 && q.w_stages > ctx->dwp_w_stages) continue;
      if (dwp::smem_bytes(q) <= ctx->smem_optin) { found = true; break; }
    }
    if (!found) continue;
    {
      int a = q.in_stages, g2 = q.G;
      while (g2) { const int t = a % g2; a = g2; g2 = t; }   // gcd(in_stages, G)
      q.in_period = q.in_stages / a;
    }
    b.dp_smem = dwp::smem_bytes(q);
    std::vector<float> w((size_t)q.k_chunks * dwp::WDW_CHUNK_FLOATS);
    for (int ch = 0; ch < q.C; ++ch) {
      float* base = w.data() + (size_t)(ch / 64) * dwp::WDW_CHUNK_FLOATS + ch % 64;
      for (int k = 0; k < 9; ++k) base[k * 64] = d.h_wdw[(size_t)k * q.C + ch];
      base[9 * 64] = d.h_bias[ch];
    }
    if (!upload(&b.dp_wdw, w)) return fail(ctx, SPEF_ERR_CUDA, "spef_finalize_weights: upload failed");
    b.dp_ok = true;
  }
  return SPEF_OK;
}

extern "C" int spef_finalize_weights(spef_ctx* ctx) {
  if (!ctx) return SPEF_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  const bool use_bf16 = ctx->cfg.precision == SPEF_BF16;
  for (Layer& l : ctx->layers) {
    if (l.kind == K_POOL) continue;
    std::vector<float> wf, bias;
    int N = l.cout, K = l.cin;
    if (l.kind == K_HEAD) {
      const int n_ori = ctx->cfg.n_ori, n_pos = ctx->cfg.n_pos;
      const HostTensor* wo = find_tensor(ctx, "head.ori.1.weight", {n_ori, 1280});
      const HostTensor* bo = wo ? find_tensor(ctx, "head.ori.1.bias", {n_ori}) : nullptr;
      const HostTensor* wp = bo ? find_tensor(ctx, "head.pos.0.weight", {n_pos, 1280}) : nullptr;
      const HostTensor* bp = wp ? find_tensor(ctx, "head.pos.0.bias", {n_pos}) : nullptr;
      if (!bp) return ctx->err.find("missing") != std::string::npos ? SPEF_ERR_STATE : SPEF_ERR_INVALID;
      wf.assign((size_t)l.n_pad * K, 0.f);
      bias.assign(l.n_pad, 0.f);
      memcpy(wf.data(), wo->data.data(), wo->data.size() * 4);
      memcpy(wf.data() + (size_t)n_ori * K, wp->data.data(), wp->data.size() * 4);
      memcpy(bias.data(), bo->data.data(), n_ori * 4);
      memcpy(bias.data() + n_ori, bp->data.data(), n_pos * 4);
    } else {
      const int64_t kk = (l.kind == K_PW) ? 1 : 3;
      const int64_t cin_w = (l.kind == K_DW) ? 1 : l.cin;
      const HostTensor* w = find_tensor(ctx, l.prefix + ".0.weight", {l.cout, cin_w, kk, kk});
      const HostTensor* g = w ? find_tensor(ctx, l.prefix + ".1.weight", {l.cout}) : nullptr;
      const HostTensor* b = g ? find_tensor(ctx, l.prefix + ".1.bias", {l.cout}) : nullptr;
      const HostTensor* m = b ? find_tensor(ctx, l.prefix + ".1.running_mean", {l.cout}) : nullptr;
      const HostTensor* v = m ? find_tensor(ctx, l.prefix + ".1.running_var", {l.cout}) : nullptr;
      if (!v) return ctx->err.find("missing") != std::string::npos ? SPEF_ERR_STATE : SPEF_ERR_INVALID;
      const size_t per = (size_t)cin_w * kk * kk;
      wf.resize((size_t)l.cout * per);
      bias.resize(l.cout);
      for (int co = 0; co < l.cout; ++co) {  // BN fold in double
        const double s = (double)g->data[co] / std::sqrt((double)v->data[co] + kBnEps);
        for (size_t i = 0; i < per; ++i) wf[co * per + i] = (float)((double)w->data[co * per + i] * s);
        bias[co] = (float)((double)b->data[co] - (double)m->data[co] * s);
      }
    }
    std::vector<float> packed;
    if (l.kind == K_STEM) {  // [32,3,3,3] -> [27][32]
      if (use_bf16) {  // implicit-GEMM operand [N=32][K=32] (27 taps + 5 zero columns), and the same rounded values for SIMT
        std::vector<bf16> wb(32 * 32, __float2bfloat16_rn(0.f));
        for (int co = 0; co < 32; ++co)
          for (int k = 0; k < 27; ++k) {
            wb[co * 32 + k] = __float2bfloat16_rn(wf[co * 27 + k]);
            wf[co * 27 + k] = __bfloat162float(wb[co * 32 + k]);
          }
        if (!upload(&l.w_bf16, wb)) return fail(ctx, SPEF_ERR_CUDA, "spef_finalize_weights: upload failed");
        l.tmW_ready = false;
        l.h_wb = wb;
      }
      packed.resize(27 * 32);
      for (int co = 0; co < 32; ++co)
        for (int k = 0; k < 27; ++k) packed[k * 32 + co] = wf[co * 27 + k];
    } else if (l.kind == K_DW) {  // [C,1,3,3] -> [9][C]
      packed.resize((size_t)9 * l.cout);
      for (int c = 0; c < l.cout; ++c)
        for (int k = 0; k < 9; ++k) packed[(size_t)k * l.cout + c] = wf[(size_t)c * 9 + k];
    } else {  // pw / head: [Npad][K]; SIMT operand is the transpose [K][Npad]
      const int Np = l.n_pad;
      if (use_bf16) {
        std::vector<bf16> wb((size_t)Np * K);
        for (size_t i = 0; i < wb.size(); ++i) {
          wb[i] = __float2bfloat16_rn(wf[i]);
          wf[i] = __bfloat162float(wb[i]);  // SIMT cross-check sees exactly the tensor-core operand
        }
        if (!upload(&l.w_bf16, wb)) return fail(ctx, SPEF_ERR_CUDA, "spef_finalize_weights: upload failed");
        l.tmW_ready = false;
        ctx->cp_w_ready = false;
        if (l.kind == K_PW) l.h_wb = wb;
      }
      packed.resize((size_t)K * Np);
      for (int n = 0; n < Np; ++n)
        for (int k = 0; k < K; ++k) packed[(size_t)k * Np + n] = wf[(size_t)n * K + k];
      (void)N;
    }
    bias.resize(tc::bias_floats(l.n_pad), 0.f);
    if (l.kind == K_PW || l.kind == K_DW || l.kind == K_STEM) l.h_bias = bias;
    if (l.kind == K_DW) l.h_wdw = packed;
    if (!upload(&l.w_f32, packed) || !upload(&l.bias, bias)) return fail(ctx, SPEF_ERR_CUDA, "spef_finalize_weights: upload failed");
    if (l.kind == K_DW) {  // TMA tile plan (BF16 path)
      l.dw_cv = (l.cin % 64 == 0) ? 8 : ((l.cin % 48 == 0) ? 6 : 4);
      dw::DwParams& d = l.dwp;
      d.B = 0; d.H = l.hin; d.W = l.win; d.C = l.cin; d.Ho = l.hout; d.Wo = l.wout; d.relu = l.relu;
      const int wo4 = ((l.wout + 3) / 4) * 4;
      d.TH = (l.stride == 1) ? 8 : 4;
      d.TW = (l.stride == 1) ? (wo4 < 32 ? wo4 : 32) : (wo4 < 16 ? wo4 : 16);
      if (l.stride == 1 && l.dw_cv == 6 && d.TW > 24) d.TW = 24;
      // small maps of the 64-channel-chunk kernel (256 threads = 32 (strip, row) tasks per pass): fill the pass
      if (l.dw_cv == 8 && ctx->dw_small_plan) {
        if (l.wout == 12) { l.dw_tx = 3; d.TW = 12; if (l.stride == 2) d.TH = 8; }   // 4 strips of 3 columns x 8 rows = 32 tasks (was 3 x 8 at stride 1, 3 x 4 at stride 2: 35 -> 30.6 us; the stride-2 box needs 109 KB of shared memory)
        else if (l.stride == 1 && l.wout == 24 && l.hout == 15) { d.TH = 5; }     // 6 strips x 5 rows = 30 tasks, three tiles of 5 rows (was 48 tasks = two passes, 8 + 7 rows)
      }
      d.THI = (d.TH - 1) * l.stride + 3;
      d.TWI = (d.TW - 1) * l.stride + 3;
      // bank-conflict-free row pitch for the rows-fastest lane order (dwconv_tma.cuh): 64-byte pixels need an odd box
      // width, 96-byte pixels a width = 1 (mod 4); the extra columns are in-bounds neighbours (L2 hits) or OOB zeros
      if (l.stride == 1 && l.dw_cv == 4) d.TWI |= 1;
      if (l.stride == 1 && l.dw_cv == 6) d.TWI += (5 - d.TWI % 4) % 4;
      d.tiles_y = cdiv(l.hout, d.TH);
      d.tiles_x = cdiv(l.wout, d.TW);
      d.nchunks = cdiv(l.cin, l.dw_cv * 8);
    }
    if (l.kind == K_STEM) {  // stem as an implicit GEMM: M = B*Ho*Wo, N = 32, K = 27 -> 32
      l.block_n = 32;
      l.stages = tc::pick_stages_v2(32, 32, 32, ctx->smem_optin - (tc::PATCH_STAGES * tc::PATCH_STAGE_BYTES + 1280));
      l.smem = tc::smem_bytes_v2(32, l.stages, 32, 32) + tc::PATCH_STAGES * tc::PATCH_STAGE_BYTES + 1280;   // + stem patch ring (1024-byte aligned)
    }
    if (l.kind == K_PW || l.kind == K_HEAD) {
      l.block_n = tc::pick_block_n(l.n_pad);
      // head GEMM: M = batch is at most a few 128-row tiles, so 256-column tiles would keep 14 CTAs busy at B = 256 (22.5 us to
      // stream 4.4 MB of weights); 64-column tiles spread the same weights over 4x as many CTAs
      if (l.kind == K_HEAD && ctx->cfg.max_batch <= 2048 && !ctx->head_wide) l.block_n = 64;
      l.stages = tc::pick_stages_v2(l.block_n, l.n_pad, l.cin, ctx->smem_optin);
      l.smem = tc::smem_bytes_v2(l.block_n, l.stages, l.n_pad, l.cin);
    }
  }
  if (use_bf16) {
    const int so = (int)ctx->smem_optin;
    int rcb = plan_blocks(ctx);
    if (rcb) return rcb;
    rcb = plan_blocks_t(ctx);
    if (rcb) return rcb;
    rcb = plan_blocks_dp(ctx);
    if (rcb) return rcb;
    rcb = plan_stem_block(ctx);
    if (rcb) return rcb;
    CK(cudaFuncSetAttribute(fbt::fused_block_t_kernel<1, 6, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
#define SPEF_FBT_ATTR(S_, TH_) \
    CK(cudaFuncSetAttribute(fbt::fused_block_t_kernel<S_, TH_, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    SPEF_FBT_ATTR(1, 6) SPEF_FBT_ATTR(1, 5) SPEF_FBT_ATTR(2, 4)
#undef SPEF_FBT_ATTR
    CK(cudaFuncSetAttribute(fbt::fused_block_t_kernel<1, 6, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(dwp::dw_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(fb::fused_block_kernel<1, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(fb::fused_block_kernel<2, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(fb::fused_block_kernel<1, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(fb::fused_block_kernel<2, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(fb::fused_block_kernel<1, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(fb::fused_block_kernel<2, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(fb::fused_block_kernel<1, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(fb::fused_block_kernel<2, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(tc::pw_gemm_tcgen05_v2_kernel<false, 1, 4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(tc::pw_gemm_tcgen05_v2_kernel<true, 1, 4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(tc::pw_gemm_tcgen05_v2_kernel<false, 1, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(tc::pw_gemm_tcgen05_v2_kernel<true, 1, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(tc::pw_gemm_tcgen05_v2_kernel<false, 2, 4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(tc::pw_gemm_tcgen05_v2_kernel<true, 2, 4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(tc::pw_gemm_tcgen05_v2_kernel<false, 1, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(tc::pw_gemm_tcgen05_v2_kernel<false, 1, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    CK(cudaFuncSetAttribute(tc::pw_gemm_tcgen05_v2_kernel<false, 1, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, so));
    const int dw_smem = 112 * 1024;
    CK(cudaFuncSetAttribute(dw::dwconv3x3_tma_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem));
    CK(cudaFuncSetAttribute(dw::dwconv3x3_tma_kernel<1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem));
    CK(cudaFuncSetAttribute(dw::dwconv3x3_tma_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem));
    CK(cudaFuncSetAttribute(dw::dwconv3x3_tma_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem));
    CK(cudaFuncSetAttribute(dw::dwconv3x3_tma_kernel<1, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem));
    CK(cudaFuncSetAttribute(dw::dwconv3x3_tma_kernel<2, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem));
    CK(cudaFuncSetAttribute(dw::dwconv3x3_tma_kernel<2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem));
    CK(cudaFuncSetAttribute(dw::dwconv3x3_tma_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem));
  }
  drop_graphs(ctx);   // captured launches hold the old weight pointers
  ctx->host_tensors.clear();
  ctx->plan_batch = -1;
  for (Layer& l : ctx->layers) l.plan_batch = -1;
  ctx->finalized = true;
  return SPEF_OK;
}

// ------------------------------------------------------------------------------------------------------
// histograms
// ------------------------------------------------------------------------------------------------------
extern "C" int spef_set_ori_histogram(spef_ctx* ctx, const double* q, int32_t n) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (!q || n < 1) return fail(ctx, SPEF_ERR_INVALID, "spef_set_ori_histogram: bad argument");
  CK(cudaSetDevice(ctx->cfg.device));
  drop_graphs(ctx);   // captured decode launches hold the old table pointers and sizes
  std::vector<float4> t(n);
  for (int i = 0; i < n; ++i) t[i] = make_float4((float)q[i * 4], (float)q[i * 4 + 1], (float)q[i * 4 + 2], (float)q[i * 4 + 3]);
  if (!upload(&ctx->ori_tab, t)) return fail(ctx, SPEF_ERR_CUDA, "spef_set_ori_histogram: upload failed");
  if (!upload(&ctx->ori_tab64, std::vector<double>(q, q + (size_t)n * 4))) return fail(ctx, SPEF_ERR_CUDA, "spef_set_ori_histogram: upload failed");
  ctx->ori_tab_ld = (n + dstream::SUB - 1) / dstream::SUB * dstream::SUB;
  std::vector<float> soa((size_t)4 * ctx->ori_tab_ld, 0.f);
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < 4; ++c) soa[(size_t)c * ctx->ori_tab_ld + i] = (float)q[i * 4 + c];
  if (!upload(&ctx->ori_tab_soa, soa)) return fail(ctx, SPEF_ERR_CUDA, "spef_set_ori_histogram: upload failed");
  ctx->ori_n = n;
  return SPEF_OK;
}

extern "C" int spef_set_pos_histogram(spef_ctx* ctx, const double* x, int32_t n) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (!x || n < 1) return fail(ctx, SPEF_ERR_INVALID, "spef_set_pos_histogram: bad argument");
  CK(cudaSetDevice(ctx->cfg.device));
  drop_graphs(ctx);
  std::vector<float4> t(n);
  for (int i = 0; i < n; ++i) t[i] = make_float4((float)x[i * 3], (float)x[i * 3 + 1], (float)x[i * 3 + 2], 0.f);
  if (!upload(&ctx->pos_tab, t)) return fail(ctx, SPEF_ERR_CUDA, "spef_set_pos_histogram: upload failed");
  {
    std::vector<double> t64((size_t)n * 4, 0.0);
    for (int i = 0; i < n; ++i) { t64[(size_t)i * 4] = x[i * 3]; t64[(size_t)i * 4 + 1] = x[i * 3 + 1]; t64[(size_t)i * 4 + 2] = x[i * 3 + 2]; }
    if (!upload(&ctx->pos_tab64, t64)) return fail(ctx, SPEF_ERR_CUDA, "spef_set_pos_histogram: upload failed");
  }
  ctx->pos_n = n;
  return SPEF_OK;
}

// ------------------------------------------------------------------------------------------------------
// layer launches
// ------------------------------------------------------------------------------------------------------
template <typename T>
static int launch_cuda_core_layer(spef_ctx* ctx, const Layer& l, const void* in, const void* res, void* out, int B, cudaStream_t st) {
  if (l.kind == K_STEM) {
    const long long total = (long long)B * l.hout * ((l.wout + 1) / 2) * 4;
    stem_conv3x3s2_kernel<T><<<(unsigned)cdivll(total, 256), 256, 0, st>>>((const float*)in, l.w_f32, l.bias, (T*)out, B, l.hin, l.win, l.hout, l.wout);
    CK_LAUNCH("stem_conv3x3s2_kernel");
  } else if (l.kind == K_DW) {
    const int CG = l.cin / 8;
    if (l.stride == 1) {
      const long long total = (long long)B * l.hout * cdiv(l.wout, 4) * CG;
      dwconv3x3_kernel<T, 1, 4><<<(unsigned)cdivll(total, 256), 256, 0, st>>>((const T*)in, l.w_f32, l.bias, (T*)out, B, l.hin, l.win, l.cin, l.hout, l.wout, l.relu);
    } else {
      const long long total = (long long)B * l.hout * cdiv(l.wout, 2) * CG;
      dwconv3x3_kernel<T, 2, 2><<<(unsigned)cdivll(total, 256), 256, 0, st>>>((const T*)in, l.w_f32, l.bias, (T*)out, B, l.hin, l.win, l.cin, l.hout, l.wout, l.relu);
    }
    CK_LAUNCH("dwconv3x3_kernel");
  } else if (l.kind == K_POOL) {
    global_mean_kernel<T><<<B * cdiv(l.cin / 8, 32), 256, 0, st>>>((const T*)in, (T*)out, B, l.hin * l.win, l.cin);
    CK_LAUNCH("global_mean_kernel");
  } else {  // K_PW / K_HEAD on CUDA cores
    const int M = B * l.hout * l.wout, N = l.n_pad, K = l.cin;
    dim3 grid(cdiv(M, 64), cdiv(N, 64));
    if (l.kind == K_HEAD) {
      pw_gemm_simt_kernel<T, float><<<grid, 256, 0, st>>>((const T*)in, l.w_f32, l.bias, nullptr, (float*)out, M, N, K, N, 0);
    } else {
      pw_gemm_simt_kernel<T, T><<<grid, 256, 0, st>>>((const T*)in, l.w_f32, l.bias, (const T*)res, (T*)out, M, N, K, N, l.relu);
    }
    CK_LAUNCH("pw_gemm_simt_kernel");
  }
  return SPEF_OK;
}

template <int S, int CV, int TX = 4>
static void launch_dw_inst(bool pdl, bool pdl_early, const CUtensorMap& tm, const Layer& l, bf16* out, int grid, size_t smem, cudaStream_t st) {
  dw::DwParams q = l.dwp;
  q.pdl_early = pdl_early ? 1 : 0;
  launch_chain(pdl, dw::dwconv3x3_tma_kernel<S, CV, TX>, dim3(grid), dim3(32 * CV), smem, st, tm, (const float*)l.w_f32, (const float*)l.bias, out, q);
}

static int launch_dw_tma_layer(spef_ctx* ctx, Layer& l, const void* in, void* out, int B, cudaStream_t st, bool cached_maps) {
  CUtensorMap local;
  CUtensorMap* tm = cached_maps ? &l.tmX : &local;
  if (!cached_maps || l.plan_batch != B) {
    if (!dw::make_tmap_nhwc(ctx->encode, tm, in, B, l.hin, l.win, l.cin, l.dw_cv, l.dwp.TWI, l.dwp.THI))
      return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(X) failed for %s", l.prefix.c_str());
    if (cached_maps) l.plan_batch = B;
  }
  l.dwp.B = B;
  if ((l.dwp.TW / l.dw_tx) * l.dwp.TH > 2 * 32) return fail(ctx, SPEF_ERR_INVALID, "depthwise tile plan of %s has more than two tasks per thread", l.prefix.c_str());
  const long long sp_tiles = (long long)B * l.dwp.tiles_y * l.dwp.tiles_x;
  const int per_sm = (l.dw_cv == 4) ? 4 : 2;
  // grid = nchunks * (CTAs per chunk): every chunk gets the same number of CTAs, at most the resident capacity
  long long per_chunk = ((long long)per_sm * ctx->num_sms) / l.dwp.nchunks;
  if (per_chunk < 1) per_chunk = 1;
  if (per_chunk > sp_tiles) per_chunk = sp_tiles;
  const int grid = (int)(per_chunk * l.dwp.nchunks);
  const size_t smem = dw::smem_bytes(l.dwp, l.dw_cv);
  bf16* o = (bf16*)out;
  if (l.stride == 1) {
    if (l.dw_cv == 8 && l.dw_tx == 3) launch_dw_inst<1, 8, 3>(ctx->pdl_on, ctx->pdl_early, *tm, l, o, grid, smem, st);
    else if (l.dw_cv == 8) launch_dw_inst<1, 8>(ctx->pdl_on, ctx->pdl_early, *tm, l, o, grid, smem, st);
    else if (l.dw_cv == 6) launch_dw_inst<1, 6>(ctx->pdl_on, ctx->pdl_early, *tm, l, o, grid, smem, st);
    else launch_dw_inst<1, 4>(ctx->pdl_on, ctx->pdl_early, *tm, l, o, grid, smem, st);
  } else {
    if (l.dw_cv == 8 && l.dw_tx == 3) launch_dw_inst<2, 8, 3>(ctx->pdl_on, ctx->pdl_early, *tm, l, o, grid, smem, st);
    else if (l.dw_cv == 8) launch_dw_inst<2, 8>(ctx->pdl_on, ctx->pdl_early, *tm, l, o, grid, smem, st);
    else if (l.dw_cv == 6) launch_dw_inst<2, 6>(ctx->pdl_on, ctx->pdl_early, *tm, l, o, grid, smem, st);
    else launch_dw_inst<2, 4>(ctx->pdl_on, ctx->pdl_early, *tm, l, o, grid, smem, st);
  }
  CK_LAUNCH("dwconv3x3_tma_kernel");
  return SPEF_OK;
}

static int launch_stem_tcgen05(spef_ctx* ctx, Layer& l, const void* images, void* out, int B, cudaStream_t st) {
  if (!l.tmW_ready) {
    if (!tc::make_tmap_2d(ctx->encode, &l.tmW, l.w_bf16, false, 32, 32, 32, 32)) return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(W) failed for the stem");
    l.tmW_ready = true;
  }
  tc::GemmParams p;
  memset(&p, 0, sizeof(p));
  p.bias = l.bias; p.residual = nullptr; p.M = B * l.hout * l.wout; p.N = 32; p.K = 32; p.block_n = 32; p.num_stages = l.stages; p.relu = 1;
  p.store_mode = 0; p.out = out; p.ldd = 32; p.trace = nullptr; p.pdl_early = 0;
  p.img_u8 = ctx->image_u8;
  p.img = (const float*)images; p.img_h = l.hin; p.img_w = l.win; p.out_h = l.hout; p.out_w = l.wout;
  // patch mode: 2 x 64 output-pixel tiles whose input patch is staged by TMA (needs exact tiling and 16-byte image rows)
  CUtensorMap tmImg = l.tmW;
  p.patch_mode = (ctx->stem_patch && l.hout % 2 == 0 && l.wout % 64 == 0 && l.hin == 2 * l.hout && l.win == 2 * l.wout &&
                  (l.win * (ctx->image_u8 ? 1 : 4)) % 16 == 0 && ((uintptr_t)images % 16) == 0) ? 1 : 0;
  if (p.patch_mode) {
    p.patch_tiles_x = l.wout / 64;
    p.patch_w = ctx->image_u8 ? 128 : 32;        // 128-byte column chunks
    p.patch_chunks = ctx->image_u8 ? 2 : 5;      // >= patch_x0 + 2 * 64 pixels
    p.patch_x0 = ctx->image_u8 ? 16 : 4;         // 16 bytes: the innermost TMA coordinate must be 16-byte aligned
    const size_t esz = ctx->image_u8 ? 1 : 4;
    const cuuint64_t gdim[3] = {(cuuint64_t)l.win, (cuuint64_t)l.hin, (cuuint64_t)3 * B};   // planes of all images: NCHW is [B*3][H][W]
    const cuuint64_t gstride[2] = {(cuuint64_t)l.win * esz, (cuuint64_t)l.win * l.hin * esz};
    const cuuint32_t box[3] = {(cuuint32_t)p.patch_w, 5, 3};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ctx->encode(&tmImg, ctx->image_u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(images),
                             gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(image) failed for the stem");
  }
  const int tiles = cdiv(p.M, tc::BLOCK_M);
  const int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
  if (ctx->stem_prod == 4) tc::pw_gemm_tcgen05_v2_kernel<false, 1, 4, 4><<<grid, 128 + 512 + 128 + 128, l.smem, st>>>(tmImg, l.tmW, p);
  else if (ctx->stem_prod == 2) tc::pw_gemm_tcgen05_v2_kernel<false, 1, 4, 2><<<grid, 128 + 256 + 128 + 128, l.smem, st>>>(tmImg, l.tmW, p);
  else tc::pw_gemm_tcgen05_v2_kernel<false, 1, 4, 1><<<grid, 128 + 128 + 128 + 128, l.smem, st>>>(tmImg, l.tmW, p);
  CK_LAUNCH("pw_gemm_tcgen05_v2_kernel<im2col stem>");
  return SPEF_OK;
}

static int launch_tcgen05_layer(spef_ctx* ctx, Layer& l, const void* in, const void* res, void* out, int B, cudaStream_t st, bool cached_maps) {
  const int M = B * l.hout * l.wout, N = l.n_pad, K = l.cin;
  const bool f32out = (l.kind == K_HEAD);
  if (!l.tmW_ready) {
    if (!tc::make_tmap_2d(ctx->encode, &l.tmW, l.w_bf16, false, N, K, K, l.block_n)) return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(W) failed for %s", l.prefix.c_str());
    l.tmW_ready = true;
  }
  CUtensorMap tA_local;
  CUtensorMap* tA = cached_maps ? &l.tmA : &tA_local;
  if (!cached_maps || l.plan_batch != B) {
    if (!tc::make_tmap_2d(ctx->encode, tA, in, false, M, K, K, tc::BLOCK_M)) return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed for %s (M=%d K=%d)", l.prefix.c_str(), M, K);
    if (cached_maps) l.plan_batch = B;
  }
  tc::GemmParams p;
  memset(&p, 0, sizeof(p));
  p.bias = l.bias; p.residual = (const bf16*)res; p.M = M; p.N = N; p.K = K; p.block_n = l.block_n; p.num_stages = l.stages; p.relu = l.relu;
  p.store_mode = 0; p.out = out; p.ldd = N; p.pdl_early = ctx->pdl_early ? 1 : 0;
  p.img = nullptr; p.img_h = p.img_w = p.out_h = p.out_w = 0; p.img_u8 = 0;
  const bool trace = ctx->trace_dev && (&l == &ctx->layers[ctx->trace_layer < (int)ctx->layers.size() && ctx->trace_layer >= 0 ? ctx->trace_layer : 0]) && ctx->trace_layer >= 0;
  p.trace = nullptr;
  if (trace) cudaMemsetAsync(ctx->trace_dev, 0, 256 * 16 * sizeof(long long), st);
  const int tiles = cdiv(M, tc::BLOCK_M) * cdiv(N, l.block_n);
  const int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
  {
    const int nsw = ctx->gemm_nsw, ndg = ctx->gemm_ndg, nt2 = 128 + 128 * ndg + 32 * nsw;
#define SPEF_V2_LAUNCH(F32, NDG_, NSW_) launch_chain(ctx->pdl_on, tc::pw_gemm_tcgen05_v2_kernel<F32, NDG_, NSW_, 0>, dim3(grid), dim3(nt2), l.smem, st, *tA, l.tmW, p)
    if (f32out) {
      if (ndg == 2) SPEF_V2_LAUNCH(true, 2, 4); else if (nsw == 8) SPEF_V2_LAUNCH(true, 1, 8); else SPEF_V2_LAUNCH(true, 1, 4);
    } else {
      if (ndg == 2) SPEF_V2_LAUNCH(false, 2, 4); else if (nsw == 8) SPEF_V2_LAUNCH(false, 1, 8); else SPEF_V2_LAUNCH(false, 1, 4);
    }
#undef SPEF_V2_LAUNCH
  }
  CK_LAUNCH("pw_gemm_tcgen05_v2_kernel");
  if (trace) {
    static int dumped = 0;
    if (dumped++ == 3) {  // 4th call: warmed up
      std::vector<long long> h(256 * 8);
      cudaStreamSynchronize(st);
      cudaMemcpy(h.data(), ctx->trace_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      const long long t0 = h[0];
      fprintf(stderr, "TRACE layer %s M=%d N=%d K=%d block_n=%d stages=%d (cycles rel. to first MMA start)\n tile: mma_start mma_commit | drain: tmem_full ld_done staged | store: sfull copied | mma: first smem stage ready (v2 layout; last box of the tile)\n", l.prefix.c_str(), M, N, K, l.block_n, l.stages);
      for (int t = 0; t < 40; ++t) {
        fprintf(stderr, "%3d:", t);
        for (int j = 0; j < 8; ++j) fprintf(stderr, " %8lld", h[t * 8 + j] ? h[t * 8 + j] - t0 : -1);
        fprintf(stderr, "\n");
      }
    }
  }
  return SPEF_OK;
}

// true when block bi runs as one fused kernel in the current configuration
// 0: per-layer kernels, 1: staged fused kernel (fused_block.cuh), 2: channel-lane fused kernel (fused_block_t.cuh),
// 3: expand GEMM + fused depthwise -> project kernel (dw_project.cuh)
static int block_variant(const spef_ctx* ctx, int bi) {
  if (!(ctx->fuse && ctx->cfg.precision == SPEF_BF16 && ctx->cfg.pw_impl == 0 && bi >= 0 && bi < (int)ctx->blocks.size())) return 0;
  const Block& b = ctx->blocks[bi];
  if (ctx->dwp_force && b.dp_ok) return 3;
  if (ctx->fb_variant == 1 && b.t_ok) return 2;
  if (b.fusable) return 1;
  return (ctx->dwp_enable && b.dp_ok) ? 3 : 0;
}
static bool block_is_fused(const spef_ctx* ctx, int bi) { return block_variant(ctx, bi) != 0; }

static int launch_fused_block_t(spef_ctx* ctx, Block& b, const void* in, void* out, int B, cudaStream_t st) {
  std::vector<Layer>& L = ctx->layers;
  fbt::FbtParams& q = b.tprm;
  if (!b.t_tmW_ready) {
    if (!(q.wz_bytes ? tc::make_tmap_2d(ctx->encode, &b.t_tmWe, b.t_we, false, (long long)(q.wz_bytes >> 7), 64, 64, 64)
          : q.stack > 1 ? tc::make_tmap_2d(ctx->encode, &b.t_tmWe, b.t_we, false, (long long)q.n_chunks * (q.we_bytes >> 7), 64, 64, q.we_bytes >> 7)
                       : tc::make_tmap_2d(ctx->encode, &b.t_tmWe, b.t_we, false, (long long)q.n_chunks * fbt::CL, q.Cin, q.Cin, fbt::CL)) ||
        !tc::make_tmap_2d(ctx->encode, &b.t_tmWp, b.t_wp, false, q.Cout, (long long)q.n_chunks * fbt::CL, (long long)q.n_chunks * fbt::CL, q.cpad))
      return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(W') failed for fused block at layer %d", b.first);
    b.t_tmW_ready = true;
  }
  if (b.t_tmX_ptr != in || b.t_tmX_batch != B) {
    if (!fb::make_tmap_x(ctx->encode, &b.t_tmX, in, B, q.H, q.W, q.cx, q.TWI, q.THI))
      return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(X) failed for fused block at layer %d", b.first);
    b.t_tmX_ptr = in; b.t_tmX_batch = B;
  }
  q.x = (const bf16*)in; q.y = (bf16*)out; q.B = B; q.aux = b.t_aux; q.bp = L[b.i_proj].bias;
  const bool trace = ctx->trace_dev && ctx->fb_trace_block >= 0 && &b == &ctx->blocks[ctx->fb_trace_block < (int)ctx->blocks.size() ? ctx->fb_trace_block : 0];
  q.trace = trace ? ctx->trace_dev : nullptr;
  q.pdl_early = ctx->pdl_early ? 1 : 0;
  if (trace) cudaMemsetAsync(ctx->trace_dev, 0, 256 * 16 * sizeof(long long), st);
  const long long tiles = (long long)B * q.tiles_y * q.tiles_x;
  if (tiles >= (1 << 22)) return fail(ctx, SPEF_ERR_UNSUPPORTED, "fused block: %lld tiles exceed the 2^22 limit of the tile index arithmetic", tiles);
  if ((long long)B * q.Ho * q.Wo >= (1LL << 31)) return fail(ctx, SPEF_ERR_UNSUPPORTED, "fused block: %lld output pixels exceed the 32-bit pixel offsets of the epilogue", (long long)B * q.Ho * q.Wo);
  const int grid = (int)(tiles < ctx->num_sms ? tiles : ctx->num_sms);
  const int nthr = 32 * (fbt::CTRL_WARPS + b.t_ng * fbt::GWT);
#define SPEF_FBT_LAUNCH(S_, TH_) launch_chain(ctx->pdl_on, fbt::fused_block_t_kernel<S_, TH_, 2, true>, dim3(grid), dim3(nthr), b.t_smem, st, b.t_tmX, b.t_tmWe, b.t_tmWp, q)
  const int S = L[b.i_dw].stride;
  if (b.i_exp < 0) {   // t = 1 block: no expand conv (strip-stacked four ways, two worker groups)
    if (!(S == 1 && q.TH == 6 && b.t_ng == 2)) return fail(ctx, SPEF_ERR_UNSUPPORTED, "fused block: no kernel instance for the t = 1 block with tile height %d", q.TH);
    launch_chain(ctx->pdl_on, fbt::fused_block_t_kernel<1, 6, 2, false>, dim3(grid), dim3(nthr), b.t_smem, st, b.t_tmX, b.t_tmWe, b.t_tmWp, q);
  }
  else if (S == 1 && q.TH == 6) SPEF_FBT_LAUNCH(1, 6);
  else if (S == 1 && q.TH == 5) SPEF_FBT_LAUNCH(1, 5);
  else if (S == 2 && q.TH == 4) SPEF_FBT_LAUNCH(2, 4);
  else return fail(ctx, SPEF_ERR_UNSUPPORTED, "fused block: no kernel instance for stride %d tile height %d", S, q.TH);
#undef SPEF_FBT_LAUNCH
  CK_LAUNCH("fused_block_t_kernel");
  if (trace) {
    static int dumped = 0;
    if (dumped++ == 3) {
      std::vector<long long> h(64 * 16);
      cudaStreamSynchronize(st);
      cudaMemcpy(h.data(), ctx->trace_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      long long t0 = 0;
      for (int j = 0; j < 16; ++j) if (h[j] && (!t0 || h[j] < t0)) t0 = h[j];
      fprintf(stderr, "FBT TRACE block at layer %d: tile %dx%d (in %dx%d, n_px %d) chunks %d w_stages %d resident %d x_stages %d grid %d B %d\n"
              " n: E enter/waited/exit - P waited/exit | worker: top acc_full rows_done - - - synced | epilogue: proj_full done\n",
              b.first, q.TH, q.TW, q.THI, q.TWI, q.n_px, q.n_chunks, q.w_stages, q.resident, q.x_stages, grid, B);
      for (int t = 0; t < 48; ++t) {
        fprintf(stderr, "%3d:", t);
        for (int j = 0; j < 16; ++j) fprintf(stderr, " %7lld", h[t * 16 + j] ? h[t * 16 + j] - t0 : -1);
        fprintf(stderr, "\n");
      }
    }
  }
  return SPEF_OK;
}

static int run_layer(spef_ctx* ctx, Layer& l, const void* in, const void* res, void* out, int B, cudaStream_t st, bool cached_maps);

// the stem conv and the first InvertedResidual block as ONE launch: images [B,3,H,W] (f32 or uint8) -> block output [B,H/2,W/2,16]
static bool stem_block_fused(const spef_ctx* ctx) {
  return ctx->stem_fuse && !ctx->stem_simt && !ctx->blocks.empty() && ctx->blocks[0].s_ok && block_variant(ctx, 0) == 2;
}
static int launch_stem_block(spef_ctx* ctx, const void* images, void* out, int B, cudaStream_t st) {
  Block& b = ctx->blocks[0];
  std::vector<Layer>& L = ctx->layers;
  fbt::FbtParams& q = b.sprm;
  const Layer& s = L[0];
  if (((uintptr_t)images % 16) != 0) return fail(ctx, SPEF_ERR_INVALID, "fused stem: the image batch must be 16-byte aligned");
  if (!b.s_tmW_ready) {
    if (!tc::make_tmap_2d(ctx->encode, &b.s_tmWe, b.s_we, false, (long long)(q.we_bytes >> 7), 64, 64, q.we_bytes >> 7) ||
        !tc::make_tmap_2d(ctx->encode, &b.s_tmWp, b.t_wp, false, q.Cout, (long long)q.n_chunks * fbt::CL, (long long)q.n_chunks * fbt::CL, q.cpad))
      return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(W') failed for the fused stem");
    b.s_tmW_ready = true;
  }
  q.img_u8 = ctx->image_u8;
  q.patch_w = ctx->image_u8 ? 128 : 104;       // >= patch_x0 + 2 * 48 + 1 columns; rows of a multiple of 16 bytes
  q.patch_x0 = ctx->image_u8 ? 16 : 4;         // 16 bytes: the innermost TMA coordinate must be 16-byte aligned
  if (b.s_img_ptr != images || b.s_img_batch != B || b.s_img_u8 != ctx->image_u8) {
    const size_t esz = ctx->image_u8 ? 1 : 4;
    const cuuint64_t gdim[3] = {(cuuint64_t)s.win, (cuuint64_t)s.hin, (cuuint64_t)3 * B};   // planes of all images: NCHW is [B*3][H][W]
    const cuuint64_t gstride[2] = {(cuuint64_t)s.win * esz, (cuuint64_t)s.win * s.hin * esz};
    const cuuint32_t box[3] = {(cuuint32_t)q.patch_w, (cuuint32_t)(2 * q.THI + 1), 3};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ctx->encode(&b.s_tmImg, ctx->image_u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(images),
                             gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(image) failed for the fused stem");
    b.s_img_ptr = images; b.s_img_batch = B; b.s_img_u8 = ctx->image_u8;
  }
  q.x = nullptr; q.y = (bf16*)out; q.B = B; q.aux = b.s_aux; q.bp = L[b.i_proj].bias; q.trace = nullptr;
  q.pdl_early = ctx->pdl_early ? 1 : 0;
  const long long tiles = (long long)B * q.tiles_y * q.tiles_x;
  if (tiles >= (1 << 22)) return fail(ctx, SPEF_ERR_UNSUPPORTED, "fused stem: %lld tiles exceed the 2^22 limit of the tile index arithmetic", tiles);
  const int grid = (int)(tiles < ctx->num_sms ? tiles : ctx->num_sms);
  const int nthr = 32 * (fbt::CTRL_WARPS + 2 * fbt::GWT);
  launch_chain(ctx->pdl_on, fbt::fused_block_t_kernel<1, 6, 2, true, true>, dim3(grid), dim3(nthr), b.s_smem, st, b.s_tmImg, b.s_tmWe, b.s_tmWp, q);
  CK_LAUNCH("fused_block_t_kernel<stem>");
  return SPEF_OK;
}

// hidden: the expand conv's output [B,H,W,C]; res: the block input (skip connection) or nullptr
static int launch_dw_project(spef_ctx* ctx, Block& b, const void* hidden, const void* res, void* out, int B, cudaStream_t st) {
  std::vector<Layer>& L = ctx->layers;
  dwp::DwpParams& q = b.dprm;
  const Layer& pj = L[b.i_proj];
  if (!b.dp_tmW_ready) {
    if (!tc::make_tmap_2d(ctx->encode, &b.dp_tmW, pj.w_bf16, false, q.N, q.C, q.C, q.nh))
      return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(Wp) failed for the depthwise-project kernel at layer %d", b.i_dw);
    b.dp_tmW_ready = true;
  }
  if (b.dp_tmX_ptr != hidden || b.dp_tmX_batch != B) {
    const Layer& dl = L[b.i_dw];
    if (!dw::make_tmap_nhwc(ctx->encode, &b.dp_tmX, hidden, B, dl.hin, dl.win, q.C, 8, q.WB, (q.S == 1) ? q.TH + 2 : 2 * q.TH + 1))
      return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(X) failed for the depthwise-project kernel at layer %d", b.i_dw);
    b.dp_tmX_ptr = hidden; b.dp_tmX_batch = B;
  }
  q.pdl_early = ctx->pdl_early ? 1 : 0;
  q.B = B; q.wdw = b.dp_wdw; q.bias = pj.bias; q.residual = pj.residual ? (const bf16*)res : nullptr; q.out = (bf16*)out;
  const long long tiles = (long long)B * q.tiles_y;
  const int grid = (int)(tiles < ctx->num_sms ? tiles : ctx->num_sms);
  launch_chain(ctx->pdl_on, dwp::dw_project_kernel, dim3(grid), dim3(dwp::NT), b.dp_smem, st, b.dp_tmX, b.dp_tmW, q);
  CK_LAUNCH("dw_project_kernel");
  return SPEF_OK;
}

static int launch_fused_block(spef_ctx* ctx, Block& b, const void* in, void* out, int B, cudaStream_t st) {
  const int variant = block_variant(ctx, (int)(&b - ctx->blocks.data()));
  if (variant == 3) {
    // teacher-forced block (spef_block_forward): expand into the hidden buffer of the forward pass, then the fused kernel
    Layer& e = ctx->layers[b.i_exp];
    int rc = run_layer(ctx, e, in, nullptr, buf_ptr(ctx, e.dst), B, st, false);
    if (rc) return rc;
    return launch_dw_project(ctx, b, buf_ptr(ctx, e.dst), in, out, B, st);
  }
  if (variant == 2) return launch_fused_block_t(ctx, b, in, out, B, st);
  std::vector<Layer>& L = ctx->layers;
  fb::FbParams& q = b.prm;
  if (!b.tmW_ready) {
    const Layer& pj = L[b.i_proj];
    const Layer& e = b.i_exp >= 0 ? L[b.i_exp] : pj;   // t = 1 block: tmWe is never used by the kernel
    if (!tc::make_tmap_2d(ctx->encode, &b.tmWe, e.w_bf16, false, b.i_exp >= 0 ? q.Ch : q.Cout, b.i_exp >= 0 ? q.Cin : q.Ch, b.i_exp >= 0 ? q.Cin : q.Ch, fb::HC) ||
        !tc::make_tmap_2d(ctx->encode, &b.tmWp, pj.w_bf16, false, q.Cout, q.Ch, q.Ch, q.cpad))
      return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(W) failed for fused block at layer %d", b.first);
    b.tmW_ready = true;
  }
  if (b.tmX_ptr != in || b.tmX_batch != B) {
    if (!fb::make_tmap_x(ctx->encode, &b.tmX, in, B, q.H, q.W, q.Cin, q.TWI, q.THI))
      return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(X) failed for fused block at layer %d", b.first);
    b.tmX_ptr = in; b.tmX_batch = B;
  }
  q.x = (const bf16*)in; q.y = (bf16*)out; q.B = B; q.aux = b.aux; q.bp = L[b.i_proj].bias;
  const bool trace = ctx->trace_dev && ctx->fb_trace_block >= 0 && &b == &ctx->blocks[ctx->fb_trace_block < (int)ctx->blocks.size() ? ctx->fb_trace_block : 0];
  q.trace = trace ? ctx->trace_dev : nullptr;
  q.debug_skip = ctx->fb_debug_skip;
  if (trace) cudaMemsetAsync(ctx->trace_dev, 0, 256 * 16 * sizeof(long long), st);
  const long long tiles = (long long)B * q.tiles_y * q.tiles_x;
  const int grid = (int)(tiles < ctx->num_sms ? tiles : ctx->num_sms);
  const int S = L[b.i_dw].stride, gw = ctx->fb_gw;
  const int nthr = 32 * (fb::CTRL_WARPS + b.ng * gw);
#define SPEF_FB_LAUNCH(S_, NG_, GW_) fb::fused_block_kernel<S_, NG_, GW_><<<grid, nthr, b.smem, st>>>(b.tmX, b.tmWe, b.tmWp, q)
  if (gw == 8) {
    if (S == 1) { if (b.ng == 2) SPEF_FB_LAUNCH(1, 2, 8); else SPEF_FB_LAUNCH(1, 1, 8); }
    else        { if (b.ng == 2) SPEF_FB_LAUNCH(2, 2, 8); else SPEF_FB_LAUNCH(2, 1, 8); }
  } else {
    if (S == 1) { if (b.ng == 2) SPEF_FB_LAUNCH(1, 2, 4); else SPEF_FB_LAUNCH(1, 1, 4); }
    else        { if (b.ng == 2) SPEF_FB_LAUNCH(2, 2, 4); else SPEF_FB_LAUNCH(2, 1, 4); }
  }
#undef SPEF_FB_LAUNCH
  CK_LAUNCH("fused_block_kernel");
  if (trace) {
    static int dumped = 0;
    if (dumped++ == 3) {  // 4th call: warmed up
      std::vector<long long> h(64 * 16);
      cudaStreamSynchronize(st);
      cudaMemcpy(h.data(), ctx->trace_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      long long t0 = 0;
      for (int j = 0; j < 16; ++j) if (h[j] && (!t0 || h[j] < t0)) t0 = h[j];
      fprintf(stderr, "FB TRACE block at layer %d: tile %dx%d chunks %d ng %d w_stages %d resident %d x_stages %d grid %d B %d\n"
              " n: E enter/waited/exit  P enter/waited/exit | worker: top acc_full drain_loop synced a2_empty dw_loop synced | epilogue: proj_full done\n",
              b.first, q.TH, q.TW, q.n_chunks, b.ng, q.w_stages, q.resident, q.x_stages, grid, B);
      for (int t = 0; t < 48; ++t) {
        fprintf(stderr, "%3d:", t);
        for (int j = 0; j < 16; ++j) fprintf(stderr, " %7lld", h[t * 16 + j] ? h[t * 16 + j] - t0 : -1);
        fprintf(stderr, "\n");
      }
    }
  }
  return SPEF_OK;
}

static int run_layer(spef_ctx* ctx, Layer& l, const void* in, const void* res, void* out, int B, cudaStream_t st, bool cached_maps) {
  const bool use_bf16 = ctx->cfg.precision == SPEF_BF16;
  if (l.kind == K_STEM && ctx->image_u8 && !(use_bf16 && ctx->cfg.pw_impl == 0 && !ctx->stem_simt))
    return fail(ctx, SPEF_ERR_UNSUPPORTED, "uint8 images are implemented on the BF16 tcgen05 stem only");
  if (use_bf16 && (l.kind == K_PW || l.kind == K_HEAD) && ctx->cfg.pw_impl == 0) return launch_tcgen05_layer(ctx, l, in, res, out, B, st, cached_maps);
  if (use_bf16 && l.kind == K_DW && ctx->cfg.pw_impl == 0) return launch_dw_tma_layer(ctx, l, in, out, B, st, cached_maps);
  if (use_bf16 && l.kind == K_STEM && ctx->cfg.pw_impl == 0 && !ctx->stem_simt) return launch_stem_tcgen05(ctx, l, in, out, B, st);
  if (use_bf16) return launch_cuda_core_layer<bf16>(ctx, l, in, res, out, B, st);
  return launch_cuda_core_layer<float>(ctx, l, in, res, out, B, st);
}

static int check_ready(spef_ctx* ctx, int B, const char* who) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (!ctx->finalized) return fail(ctx, SPEF_ERR_STATE, "%s: weights not finalised (spef_load_tensor + spef_finalize_weights first)", who);
  if (B < 1 || B > ctx->cfg.max_batch) return fail(ctx, SPEF_ERR_INVALID, "%s: batch %d outside [1, max_batch=%d]", who, B, ctx->cfg.max_batch);
  ctx->pdl_on = false; ctx->pdl_early = false;   // single-launch entry points: plain stream order (forward_internal decides for the chain)
  return SPEF_OK;
}

// last 1x1 conv (ConvBnAct 320 -> 1280) followed by the global average pool: one kernel on the BF16 tcgen05 path (conv_pool.cuh)
static int conv_pool_ipt(const Layer& l) {
  const int hw = l.hout * l.wout;
  int ipt = 256 / (hw > 0 ? hw : 1);
  while (ipt > 1 && (ipt * hw) % 16 != 0) --ipt;
  return ipt;
}
static bool conv_pool_ok(const spef_ctx* ctx, int i) {
  if (!(ctx->fuse && ctx->pool_fuse && ctx->cfg.precision == SPEF_BF16 && ctx->cfg.pw_impl == 0)) return false;
  if (i + 1 >= (int)ctx->layers.size()) return false;
  const Layer& l = ctx->layers[i];
  const Layer& pl = ctx->layers[i + 1];
  const int hw = l.hout * l.wout;
  return l.kind == K_PW && pl.kind == K_POOL && !l.residual && l.cout % 128 == 0 && l.cin % 8 == 0 && hw % 32 == 0 && hw <= 256 &&
         (conv_pool_ipt(l) * hw) % 16 == 0 && l.w_bf16 != nullptr;
}
static int launch_conv_pool(spef_ctx* ctx, Layer& l, const void* in, void* pooled, int B, cudaStream_t st) {
  cpool::ConvPoolParams p;
  memset(&p, 0, sizeof(p));
  p.pdl_early = ctx->pdl_early ? 1 : 0;
  p.bias = l.bias; p.out = (bf16*)pooled; p.B = B; p.HW = l.hout * l.wout; p.K = l.cin; p.C = l.cout; p.relu = l.relu;
  p.ipt = conv_pool_ipt(l);
  p.n_ct = l.cout / 128;
  const int n_items = cdiv(B, p.ipt);
  p.cpc = ctx->num_sms / p.n_ct;
  if (p.cpc < 1) p.cpc = 1;
  if (p.cpc > n_items) p.cpc = n_items;
  p.x_stages = cpool::MAX_X_STAGES;
  while (p.x_stages > 2 && cpool::smem_bytes(p) > ctx->smem_optin) --p.x_stages;
  const size_t smem = cpool::smem_bytes(p);
  if (smem > ctx->smem_optin) return fail(ctx, SPEF_ERR_UNSUPPORTED, "conv + pool kernel: %zu bytes of shared memory needed", smem);
  if (!ctx->cp_w_ready) {
    if (!tc::make_tmap_2d(ctx->encode, &ctx->cp_tmW, l.w_bf16, false, l.cout, l.cin, l.cin, 128)) return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(W) failed for the conv + pool kernel");
    CK(cudaFuncSetAttribute(cpool::conv_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    ctx->cp_w_ready = true;
  }
  if (ctx->cp_x_ptr != in || ctx->cp_x_batch != B) {
    if (!tc::make_tmap_2d(ctx->encode, &ctx->cp_tmX, in, false, (long long)B * p.HW, l.cin, l.cin, p.ipt * p.HW)) return fail(ctx, SPEF_ERR_CUDA, "cuTensorMapEncodeTiled(X) failed for the conv + pool kernel");
    ctx->cp_x_ptr = in; ctx->cp_x_batch = B;
  }
  launch_chain(ctx->pdl_on, cpool::conv_pool_kernel, dim3(p.n_ct * p.cpc), dim3(cpool::NT), smem, st, ctx->cp_tmW, ctx->cp_tmX, p);
  CK_LAUNCH("conv_pool_kernel");
  return SPEF_OK;
}

static int forward_internal(spef_ctx* ctx, const float* images, int B, cudaStream_t st, float* layer_ms) {
  ctx->pdl_on = ctx->pdl != 0;
  ctx->pdl_early = ctx->pdl && B <= ctx->pdl_max_batch;
  cudaEvent_t* ev = nullptr;
  const int nl = (int)ctx->layers.size();
  if (layer_ms) {
    while ((int)ctx->events.size() < nl + 1) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      ctx->events.push_back(e);
    }
    ev = ctx->events.data();
    CK(cudaEventRecord(ev[0], st));
  }
  size_t next_block = 0;
  for (int i = 0; i < nl; ++i) {
    Layer& l = ctx->layers[i];
    while (next_block < ctx->blocks.size() && ctx->blocks[next_block].first < i) ++next_block;
    if (i == 0 && l.kind == K_STEM && stem_block_fused(ctx)) {
      // stem + first block as one launch (its time is reported in the stem's slot)
      Block& b = ctx->blocks[0];
      int rc = launch_stem_block(ctx, images, buf_ptr(ctx, ctx->layers[b.i_proj].dst), B, st);
      if (rc) return rc;
      for (int j = 0; j <= b.n_layers; ++j)
        if (ev) CK(cudaEventRecord(ev[j + 1], st));
      i += b.n_layers;
      continue;
    }
    // the depthwise -> project kernel runs one CTA per (image, row tile): a few-image step (temporal streams) would leave most SMs
    // idle through its K-chunk loop, so small batches keep the per-layer kernels (batch 1: 0.37 ms per frame with it, 0.30 without)
    int variant = 0;
    if (next_block < ctx->blocks.size() && ctx->blocks[next_block].first == i) {
      variant = block_variant(ctx, (int)next_block);
      if (variant == 3 && !ctx->dwp_force && (long long)B * ctx->blocks[next_block].dprm.tiles_y < ctx->dwp_min_tiles) variant = 0;
    }
    if (variant == 3) {
      // expand conv as a GEMM, then depthwise + project in one kernel (its time is reported in the slot of the depthwise layer)
      Block& b = ctx->blocks[next_block];
      Layer& pj = ctx->layers[b.i_proj];
      int rc = run_layer(ctx, l, buf_ptr(ctx, l.src), nullptr, buf_ptr(ctx, l.dst), B, st, true);
      if (rc) return rc;
      if (ev) CK(cudaEventRecord(ev[i + 1], st));
      rc = launch_dw_project(ctx, b, buf_ptr(ctx, l.dst), pj.res_buf >= 0 ? buf_ptr(ctx, pj.res_buf) : nullptr, buf_ptr(ctx, pj.dst), B, st);
      if (rc) return rc;
      if (ev) { CK(cudaEventRecord(ev[i + 2], st)); CK(cudaEventRecord(ev[i + 3], st)); }
      i += 2;
      continue;
    }
    if (variant != 0) {
      // one kernel for expand + depthwise + project; its time is reported in the slot of the block's first layer
      Block& b = ctx->blocks[next_block];
      int rc = launch_fused_block(ctx, b, buf_ptr(ctx, l.src), buf_ptr(ctx, ctx->layers[b.i_proj].dst), B, st);
      if (rc) return rc;
      for (int j = 0; j < b.n_layers; ++j)
        if (ev) CK(cudaEventRecord(ev[i + j + 1], st));
      i += b.n_layers - 1;
      continue;
    }
    if (conv_pool_ok(ctx, i)) {
      // last 1x1 conv + global average pool as one kernel (its time is reported in the conv's slot)
      int rc = launch_conv_pool(ctx, l, buf_ptr(ctx, l.src), buf_ptr(ctx, ctx->layers[i + 1].dst), B, st);
      if (rc) return rc;
      if (ev) { CK(cudaEventRecord(ev[i + 1], st)); CK(cudaEventRecord(ev[i + 2], st)); }
      i += 1;
      continue;
    }
    const void* in = (l.src == BUF_IMG) ? (const void*)images : buf_ptr(ctx, l.src);
    const void* res = (l.res_buf >= 0) ? buf_ptr(ctx, l.res_buf) : nullptr;
    int rc = run_layer(ctx, l, in, res, buf_ptr(ctx, l.dst), B, st, true);
    if (rc) return rc;
    if (ev) CK(cudaEventRecord(ev[i + 1], st));
  }
  ctx->plan_batch = B;
  if (layer_ms) {
    CK(cudaEventSynchronize(ev[nl]));
    for (int i = 0; i < nl; ++i) CK(cudaEventElapsedTime(&layer_ms[i], ev[i], ev[i + 1]));
  }
  return SPEF_OK;
}

static int copy_head_out(spef_ctx* ctx, int B, float* ori_out, float* pos_out, cudaStream_t st) {
  const size_t pitch = (size_t)ctx->head_pad * sizeof(float);
  if (ori_out) CK(cudaMemcpy2DAsync(ori_out, (size_t)ctx->cfg.n_ori * 4, ctx->head_out, pitch, (size_t)ctx->cfg.n_ori * 4, B, cudaMemcpyDeviceToDevice, st));
  if (pos_out) CK(cudaMemcpy2DAsync(pos_out, (size_t)ctx->cfg.n_pos * 4, ctx->head_out + ctx->cfg.n_ori, pitch, (size_t)ctx->cfg.n_pos * 4, B, cudaMemcpyDeviceToDevice, st));
  return SPEF_OK;
}

extern "C" int spef_forward(spef_ctx* ctx, const float* images_dev, int32_t B, float* ori_out, float* pos_out, void* stream) {
  int rc = check_ready(ctx, B, "spef_forward");
  if (rc) return rc;
  if (!images_dev) return fail(ctx, SPEF_ERR_INVALID, "spef_forward: images_dev is NULL");
  CK(cudaSetDevice(ctx->cfg.device));
  rc = forward_internal(ctx, images_dev, B, (cudaStream_t)stream, nullptr);
  if (rc) return rc;
  return copy_head_out(ctx, B, ori_out, pos_out, (cudaStream_t)stream);
}

extern "C" int spef_forward_timed(spef_ctx* ctx, const float* images_dev, int32_t B, float* ori_out, float* pos_out, float* layer_ms, void* stream) {
  int rc = check_ready(ctx, B, "spef_forward_timed");
  if (rc) return rc;
  if (!images_dev || !layer_ms) return fail(ctx, SPEF_ERR_INVALID, "spef_forward_timed: NULL argument");
  CK(cudaSetDevice(ctx->cfg.device));
  rc = forward_internal(ctx, images_dev, B, (cudaStream_t)stream, layer_ms);
  if (rc) return rc;
  return copy_head_out(ctx, B, ori_out, pos_out, (cudaStream_t)stream);
}

extern "C" int spef_set_image_dtype(spef_ctx* ctx, int32_t dt) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (dt != SPEF_IMG_F32 && dt != SPEF_IMG_U8) return fail(ctx, SPEF_ERR_INVALID, "spef_set_image_dtype: unknown dtype %d", dt);
  if (dt == SPEF_IMG_U8 && ctx->cfg.precision != SPEF_BF16) return fail(ctx, SPEF_ERR_UNSUPPORTED, "uint8 images need the BF16 engine");
  if (ctx->image_u8 != (dt == SPEF_IMG_U8)) drop_graphs(ctx);
  ctx->image_u8 = (dt == SPEF_IMG_U8);
  return SPEF_OK;
}

extern "C" int spef_num_layers(const spef_ctx* ctx) { return ctx ? (int)ctx->layers.size() : 0; }

extern "C" int spef_layer_info(const spef_ctx* ctx, int32_t i, int32_t* kind, int32_t* cin, int32_t* cout, int32_t* hin, int32_t* win,
                               int32_t* hout, int32_t* wout, int32_t* stride, int32_t* relu, int32_t* has_res) {
  if (!ctx || i < 0 || i >= (int)ctx->layers.size()) return SPEF_ERR_INVALID;
  const Layer& l = ctx->layers[i];
  if (kind) *kind = l.kind;
  if (cin) *cin = l.cin;
  if (cout) *cout = (l.kind == K_HEAD) ? l.n_pad : l.cout;
  if (hin) *hin = l.hin;
  if (win) *win = l.win;
  if (hout) *hout = l.hout;
  if (wout) *wout = l.wout;
  if (stride) *stride = l.stride;
  if (relu) *relu = l.relu;
  if (has_res) *has_res = l.residual;
  return SPEF_OK;
}

extern "C" int spef_layer_forward(spef_ctx* ctx, int32_t i, const void* in, const void* res, void* out, int32_t B, void* stream) {
  int rc = check_ready(ctx, B, "spef_layer_forward");
  if (rc) return rc;
  if (i < 0 || i >= (int)ctx->layers.size() || !in || !out) return fail(ctx, SPEF_ERR_INVALID, "spef_layer_forward: bad argument");
  CK(cudaSetDevice(ctx->cfg.device));
  Layer& l = ctx->layers[i];
  return run_layer(ctx, l, in, l.residual ? res : nullptr, out, B, (cudaStream_t)stream, false);
}

extern "C" int spef_num_blocks(const spef_ctx* ctx) { return ctx ? (int)ctx->blocks.size() : 0; }

extern "C" int spef_block_info(const spef_ctx* ctx, int32_t i, int32_t* first_layer, int32_t* n_layers, int32_t* fused,
                               int32_t* tile_h, int32_t* tile_w, int32_t* groups, int32_t* w_stages, int32_t* resident) {
  if (!ctx || i < 0 || i >= (int)ctx->blocks.size()) return SPEF_ERR_INVALID;
  const Block& b = ctx->blocks[i];
  const int v = block_variant(ctx, i);
  if (first_layer) *first_layer = b.first;
  if (n_layers) *n_layers = b.n_layers;
  if (fused) *fused = v;
  if (tile_h) *tile_h = v == 3 ? b.dprm.TH : v == 2 ? b.tprm.TH : (v == 1 ? b.prm.TH : 0);
  if (tile_w) *tile_w = v == 3 ? b.dprm.W : v == 2 ? b.tprm.TW : (v == 1 ? b.prm.TW : 0);
  if (groups) *groups = v == 2 ? b.t_ng : (v == 1 ? b.ng : 0);
  if (w_stages) *w_stages = v == 2 ? b.tprm.w_stages : (v == 1 ? b.prm.w_stages : 0);
  if (resident) *resident = v == 2 ? b.tprm.resident : (v == 1 ? b.prm.resident : 0);
  return SPEF_OK;
}

extern "C" int spef_set_fusion(spef_ctx* ctx, int32_t on) {
  if (!ctx) return SPEF_ERR_INVALID;
  drop_graphs(ctx);
  ctx->fuse = on ? 1 : 0;
  return SPEF_OK;
}

extern "C" int spef_set_stem_fusion(spef_ctx* ctx, int32_t on) {
  if (!ctx) return SPEF_ERR_INVALID;
  drop_graphs(ctx);
  ctx->stem_fuse = on ? 1 : 0;
  return SPEF_OK;
}

extern "C" int spef_pool_fusion_active(const spef_ctx* ctx) {
  if (!ctx || !ctx->finalized) return 0;
  for (int i = 0; i + 1 < (int)ctx->layers.size(); ++i)
    if (conv_pool_ok(ctx, i)) return 1;
  return 0;
}
extern "C" int spef_stem_fusion_active(const spef_ctx* ctx) { return (ctx && ctx->finalized && stem_block_fused(ctx)) ? 1 : 0; }

extern "C" int spef_stem_block_forward(spef_ctx* ctx, const void* images_dev, void* out_dev, int32_t B, void* stream) {
  int rc = check_ready(ctx, B, "spef_stem_block_forward");
  if (rc) return rc;
  if (!images_dev || !out_dev) return fail(ctx, SPEF_ERR_INVALID, "spef_stem_block_forward: NULL argument");
  if (!stem_block_fused(ctx)) return fail(ctx, SPEF_ERR_UNSUPPORTED, "spef_stem_block_forward: the stem is not fused into the first block in this configuration");
  CK(cudaSetDevice(ctx->cfg.device));
  return launch_stem_block(ctx, images_dev, out_dev, B, (cudaStream_t)stream);
}

extern "C" int spef_block_forward(spef_ctx* ctx, int32_t i, const void* in, void* out, int32_t B, void* stream) {
  int rc = check_ready(ctx, B, "spef_block_forward");
  if (rc) return rc;
  if (i < 0 || i >= (int)ctx->blocks.size() || !in || !out) return fail(ctx, SPEF_ERR_INVALID, "spef_block_forward: bad argument");
  if (!block_is_fused(ctx, i)) return fail(ctx, SPEF_ERR_UNSUPPORTED, "spef_block_forward: block %d has no fused kernel in this configuration", i);
  CK(cudaSetDevice(ctx->cfg.device));
  return launch_fused_block(ctx, ctx->blocks[i], in, out, B, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------
// ingest: camera frames -> the tensor the stem reads
// ------------------------------------------------------------------------------------------------------
static int resize_plan(spef_ctx* ctx, int sh, int sw, int C) {
  const int oh = ctx->cfg.img_h, ow = ctx->cfg.img_w;
  if (ctx->rz_tab && ctx->rz_sh == sh && ctx->rz_sw == sw) return SPEF_OK;
  const ingest::AxisTaps h = ingest::make_axis_taps(sw, ow), v = ingest::make_axis_taps(sh, oh);
  // layout of the table block (int32): hfirst[ow] hcount[ow] hcoef[hks][ow] vfirst[oh] vcount[oh] vcoef[oh][vks]
  std::vector<int> tab;
  // horizontal taps: when the filter has at most KWIN taps (and the row at least that many pixels) every output gets a
  // window of exactly KWIN taps that stays inside the row -- shifted left at the right edge, zero coefficients elsewhere --
  // so that the kernel reads at compile-time offsets without predicates
  ctx->rz_fixed = (h.ksize <= ingest::KWIN && sw >= ingest::KWIN) ? 1 : 0;
  const int hk = ctx->rz_fixed ? ingest::KWIN : h.ksize;
  std::vector<int> hfirst(h.first), hcoef((size_t)hk * ow, 0);
  for (int o = 0; o < ow; ++o) {
    int shift = 0;
    if (ctx->rz_fixed && h.first[o] + ingest::KWIN > sw) { shift = h.first[o] + ingest::KWIN - sw; hfirst[o] = sw - ingest::KWIN; }
    for (int j = 0; j < h.count[o]; ++j) hcoef[(size_t)(j + shift) * ow + o] = h.coef[(size_t)o * h.ksize + j];
  }
  tab.insert(tab.end(), hfirst.begin(), hfirst.end());
  tab.insert(tab.end(), h.count.begin(), h.count.end());
  tab.insert(tab.end(), hcoef.begin(), hcoef.end());
  // greyscale fast path (resize_aa_gray_kernel): first aligned word of every window and the coefficients as three byte planes
  // on that word grid
  ctx->rz_gray = (ctx->rz_fixed && sw % 16 == 0) ? 1 : 0;
  std::vector<int> gword(ow, 0), gplane((size_t)12 * ow, 0);
  if (ctx->rz_gray) {
    for (int o = 0; o < ow; ++o) {
      const int x0 = hfirst[o], sh4 = x0 & 3;
      gword[o] = x0 >> 2;
      for (int j = 0; j < ingest::KWIN; ++j) {
        const unsigned k = (unsigned)hcoef[(size_t)j * ow + o];
        const int pos = j + sh4, w = pos >> 2, bb = pos & 3;
        for (int q = 0; q < 3; ++q) gplane[(size_t)(q * 4 + w) * ow + o] |= (int)(((k >> (8 * q)) & 0xffu) << (8 * bb));
      }
    }
  }
  tab.insert(tab.end(), v.first.begin(), v.first.end());
  tab.insert(tab.end(), v.count.begin(), v.count.end());
  tab.insert(tab.end(), v.coef.begin(), v.coef.end());
  ctx->rz_gray_off = (int)tab.size();
  tab.insert(tab.end(), gword.begin(), gword.end());
  tab.insert(tab.end(), gplane.begin(), gplane.end());
  cudaFree(ctx->rz_tab);
  ctx->rz_tab = nullptr;
  CK(cudaMalloc(&ctx->rz_tab, tab.size() * sizeof(int)));
  CK(cudaMemcpy(ctx->rz_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
  ctx->rz_sh = sh; ctx->rz_sw = sw; ctx->rz_hks = hk; ctx->rz_vks = v.ksize;
  ctx->rz_pitch = (ow + 15) & ~15;
  // rows per CTA: as many as keep the filtered rows of a 3-channel band within 96 KB of shared memory (two CTAs per SM)
  for (int band = 8; band >= 1; band >>= 1) {
    int rows = 0;
    for (int y0 = 0; y0 < oh; y0 += band) {
      const int y1 = std::min(y0 + band, oh) - 1;
      rows = std::max(rows, v.first[y1] + v.count[y1] - v.first[y0]);
    }
    ctx->rz_band = band;
    ctx->rz_max_rows = rows;
    if ((size_t)3 * rows * ctx->rz_pitch <= (size_t)96 * 1024) break;
  }
  if ((size_t)C * ctx->rz_max_rows * ctx->rz_pitch > ctx->smem_optin)
    return fail(ctx, SPEF_ERR_UNSUPPORTED, "spef_resize_frames: a %d-row filter window of %d-pixel rows does not fit in shared memory", ctx->rz_max_rows, ow);
  return SPEF_OK;
}

template <int KREG, int C>
static cudaError_t launch_resize(const ingest::ResizeParams& p, int B, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(ingest::resize_aa_kernel<KREG, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int threads = std::min(384, (p.ow + 31) & ~31);
  ingest::resize_aa_kernel<KREG, C><<<dim3((p.oh + p.band - 1) / p.band, B), threads, smem, st>>>(p);
  return cudaGetLastError();
}

extern "C" int spef_resize_frames(spef_ctx* ctx, const uint8_t* frames_dev, int32_t B, int32_t src_h, int32_t src_w, int32_t channels,
                                  void* images_out_dev, int32_t out_dtype, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (!frames_dev || !images_out_dev || B < 1 || B > 65535 || src_h < 1 || src_w < 1)
    return fail(ctx, SPEF_ERR_INVALID, "spef_resize_frames: bad argument");
  if (channels != 1 && channels != 3) return fail(ctx, SPEF_ERR_INVALID, "spef_resize_frames: frames must have 1 (grey) or 3 (RGB) channels, got %d", channels);
  if (out_dtype != SPEF_IMG_F32 && out_dtype != SPEF_IMG_U8) return fail(ctx, SPEF_ERR_INVALID, "spef_resize_frames: unknown output dtype %d", out_dtype);
  CK(cudaSetDevice(ctx->cfg.device));
  int rc = resize_plan(ctx, src_h, src_w, channels);
  if (rc) return rc;
  const int oh = ctx->cfg.img_h, ow = ctx->cfg.img_w;
  ingest::ResizeParams p;
  p.src = frames_dev; p.dst = images_out_dev;
  p.hfirst = ctx->rz_tab; p.hcount = p.hfirst + ow; p.hcoef = p.hcount + ow;
  p.vfirst = p.hcoef + (size_t)ctx->rz_hks * ow; p.vcount = p.vfirst + oh; p.vcoef = p.vcount + oh;
  p.sh = src_h; p.sw = src_w; p.C = channels; p.oh = oh; p.ow = ow; p.hks = ctx->rz_hks; p.vks = ctx->rz_vks;
  p.band = ctx->rz_band; p.max_rows = ctx->rz_max_rows; p.pitch = ctx->rz_pitch; p.out_f32 = (out_dtype == SPEF_IMG_F32);
  const size_t smem = (size_t)channels * p.max_rows * p.pitch;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  constexpr int RB = 8;
  const size_t gsmem = 2 * ((size_t)RB * src_w + 16) + (size_t)p.max_rows * p.pitch;
  if (channels == 1 && ctx->rz_gray && ctx->rz_gray_kernel && (reinterpret_cast<uintptr_t>(frames_dev) & 15) == 0 && ((size_t)src_h * src_w) % 16 == 0 &&
      gsmem <= ctx->smem_optin) {
    ingest::GrayPlan g;
    g.hword = ctx->rz_tab + ctx->rz_gray_off;
    g.hplane = reinterpret_cast<const uint32_t*>(g.hword + ow);
    e = cudaFuncSetAttribute(ingest::resize_aa_gray_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem);
    if (e == cudaSuccess) {
      const int threads = std::min(384, (ow + 31) & ~31);
      ingest::resize_aa_gray_kernel<RB><<<dim3((oh + p.band - 1) / p.band, B), threads, gsmem, st>>>(p, g);
      e = cudaGetLastError();
    }
  } else if (ctx->rz_fixed) e = (channels == 1) ? launch_resize<ingest::KWIN, 1>(p, B, smem, st) : launch_resize<ingest::KWIN, 3>(p, B, smem, st);
  else e = (channels == 1) ? launch_resize<0, 1>(p, B, smem, st) : launch_resize<0, 3>(p, B, smem, st);
  if (e != cudaSuccess) return fail(ctx, SPEF_ERR_CUDA, "launch of resize_aa_kernel failed: %s", cudaGetErrorString(e));
  ctx->launches++;
  return SPEF_OK;
}

// ------------------------------------------------------------------------------------------------------
// post-processing
// ------------------------------------------------------------------------------------------------------
template <int NW, int RING, int PF, bool LOGITS>
static auto decode_stream_pick(bool amax, bool precise) {
  return precise ? (amax ? dstream::decode_ori_stream_kernel<NW, RING, PF, true, true, LOGITS> : dstream::decode_ori_stream_kernel<NW, RING, PF, false, true, LOGITS>)
                 : (amax ? dstream::decode_ori_stream_kernel<NW, RING, PF, true, false, LOGITS> : dstream::decode_ori_stream_kernel<NW, RING, PF, false, false, LOGITS>);
}
template <int NW, int RING, int PF>
static cudaError_t decode_stream_attr() {
  const int smem = (int)dstream::smem_bytes(NW, RING);
  cudaError_t e = cudaSuccess;
  for (int v = 0; v < 8 && e == cudaSuccess; ++v)
    e = (v & 4) ? cudaFuncSetAttribute(decode_stream_pick<NW, RING, PF, true>(v & 1, v & 2), cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                : cudaFuncSetAttribute(decode_stream_pick<NW, RING, PF, false>(v & 1, v & 2), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  return e;
}
// once per context (not at launch time: launches may be under CUDA-graph capture)
static cudaError_t decode_stream_init() {
  cudaError_t e = decode_stream_attr<8, 4, 0>();
  if (e == cudaSuccess) e = decode_stream_attr<16, 2, 0>();
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dstream::decode_ori_half_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dstream::half_smem_bytes());
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dstream::decode_ori_half_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dstream::half_smem_bytes());
  return e;
}

template <int NW, int RING, int PF>
static int launch_decode_stream_cfg(spef_ctx* ctx, const float* in, int ld, int B, int n, int is_logits, float* soft, float* quat, float* hinv,
                                    int32_t* amax, uint32_t* flags, cudaStream_t st) {
  const size_t smem = dstream::smem_bytes(NW, RING);
  const int grid = std::min(cdiv(B, NW), ctx->num_sms);
  auto kern = is_logits ? decode_stream_pick<NW, RING, PF, true>(amax != nullptr, hinv != nullptr)
                        : decode_stream_pick<NW, RING, PF, false>(amax != nullptr, hinv != nullptr);
  launch_chain(ctx->pdl_on, kern, dim3(grid), dim3(NW * 32), smem, st, in, ld, B, n, (const float*)ctx->ori_tab_soa, ctx->ori_tab_ld, soft, quat, hinv, (int*)amax, flags);
  CK_LAUNCH("decode_ori_stream_kernel");
  return SPEF_OK;
}

// The streaming decode kernel (one persistent CTA per SM, SoA table in shared memory, logits through per-warp rings of bulk
// async copies, packed FP32 pairs).  Warps x ring slots per CTA: batches that fill the GPU take 16 x 2 (more warps beat a deeper
// ring, and an L2 prefetch 6 steps ahead only cost issue slots: profiles/r01_decode_ab.txt), smaller ones 8 x 4 (more CTAs).
// SPEF_DECODE_CFG overrides: 0 = 8 x 4, 2 = 16 x 2.
static int launch_decode_stream(spef_ctx* ctx, const float* in, int ld, int B, int n, int is_logits, float* soft, float* quat, float* hinv,
                                int32_t* amax, uint32_t* flags, cudaStream_t st) {
  const int cfg = ctx->decode_cfg >= 0 ? ctx->decode_cfg : ((long long)B >= 16LL * ctx->num_sms ? 2 : 0);
  // small histograms at batches that fill the GPU, quaternions only: half a warp per image (decode_ori_half_kernel); SPEF_DECODE_CFG=3 forces it
  if ((cfg == 2 || ctx->decode_cfg == 3) && ctx->decode_cfg != 2 && n <= dstream::HN && ctx->ori_tab_ld <= 2 * dstream::HN && !soft && !hinv && !amax) {
    const int pairs = (B + 1) / 2;
    const int grid = std::min(cdiv(pairs, dstream::HNW), ctx->num_sms);
    if (is_logits) launch_chain(ctx->pdl_on, dstream::decode_ori_half_kernel<true>, dim3(grid), dim3(dstream::HNW * 32), dstream::half_smem_bytes(), st, in, ld, B, n, (const float*)ctx->ori_tab_soa, ctx->ori_tab_ld, quat, flags);
    else launch_chain(ctx->pdl_on, dstream::decode_ori_half_kernel<false>, dim3(grid), dim3(dstream::HNW * 32), dstream::half_smem_bytes(), st, in, ld, B, n, (const float*)ctx->ori_tab_soa, ctx->ori_tab_ld, quat, flags);
    CK_LAUNCH("decode_ori_half_kernel");
    return SPEF_OK;
  }
  if (cfg == 0) return launch_decode_stream_cfg<8, 4, 0>(ctx, in, ld, B, n, is_logits, soft, quat, hinv, amax, flags, st);
  return launch_decode_stream_cfg<16, 2, 0>(ctx, in, ld, B, n, is_logits, soft, quat, hinv, amax, flags, st);
}

static int decode_ori_ld(spef_ctx* ctx, const float* in, int ld, int B, int n, int is_logits, float* soft, float* quat, float* hinv,
                         int32_t* amax, uint32_t* flags, cudaStream_t st) {
  if (!ctx->ori_tab) return fail(ctx, SPEF_ERR_STATE, "decode_ori: orientation histogram not set (spef_set_ori_histogram)");
  if (n != ctx->ori_n) return fail(ctx, SPEF_ERR_INVALID, "decode_ori: n = %d but the histogram has %d bins", n, ctx->ori_n);
  if (!in || !quat || B < 1) return fail(ctx, SPEF_ERR_INVALID, "decode_ori: bad argument");
  // images per warp: 32 (eigen-solves fully lane-parallel) once there are enough images to fill the GPU that way
  const long long fill = (long long)ctx->num_sms * 16;  // warps needed for ~16 warps per SM
  // large batches: the streaming kernel (persistent CTAs of 8 warps, SoA table in shared memory, packed FP32 pairs)
  const bool vec_ok = (n % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) &&
                      (!soft || (reinterpret_cast<uintptr_t>(soft) & 15) == 0);
  if (vec_ok && ctx->decode_stream != 0) {
    return launch_decode_stream(ctx, in, ld, B, n, is_logits, soft, quat, hinv, amax, flags, st);
  }
  if ((long long)B >= 32 * fill) launch_chain(ctx->pdl_on, decode_ori_kernel<32>, dim3(cdiv(cdiv(B, 32), 4)), dim3(128), 0, st, in, ld, B, n, is_logits, (const float4*)ctx->ori_tab, soft, quat, hinv, (int*)amax, flags);
  else if ((long long)B >= 8 * fill) launch_chain(ctx->pdl_on, decode_ori_kernel<8>, dim3(cdiv(cdiv(B, 8), 4)), dim3(128), 0, st, in, ld, B, n, is_logits, (const float4*)ctx->ori_tab, soft, quat, hinv, (int*)amax, flags);
  else launch_chain(ctx->pdl_on, decode_ori_kernel<1>, dim3(cdiv(B, 4)), dim3(128), 0, st, in, ld, B, n, is_logits, (const float4*)ctx->ori_tab, soft, quat, hinv, (int*)amax, flags);
  CK_LAUNCH("decode_ori_kernel");
  return SPEF_OK;
}

static int decode_pos_ld(spef_ctx* ctx, const float* in, int ld, int B, int n, int is_logits, float* soft, float* pos, uint32_t* flags, cudaStream_t st) {
  if (!ctx->pos_tab) return fail(ctx, SPEF_ERR_STATE, "decode_pos: position histogram not set (spef_set_pos_histogram)");
  if (n != ctx->pos_n) return fail(ctx, SPEF_ERR_INVALID, "decode_pos: n = %d but the histogram has %d bins", n, ctx->pos_n);
  if (!in || !pos || B < 1) return fail(ctx, SPEF_ERR_INVALID, "decode_pos: bad argument");
  launch_chain(ctx->pdl_on, decode_pos_kernel, dim3(cdiv(B, 4)), dim3(128), 0, st, in, ld, B, n, is_logits, (const float4*)ctx->pos_tab, soft, pos, flags);
  CK_LAUNCH("decode_pos_kernel");
  return SPEF_OK;
}

extern "C" int spef_decode_ori(spef_ctx* ctx, const float* in, int32_t B, int32_t n, int32_t is_logits, float* soft, float* quat,
                               float* hinv, int32_t* amax, uint32_t* flags, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  return decode_ori_ld(ctx, in, n, B, n, is_logits, soft, quat, hinv, amax, flags, (cudaStream_t)stream);
}

extern "C" int spef_decode_pos(spef_ctx* ctx, const float* in, int32_t B, int32_t n, int32_t is_logits, float* soft, float* pos,
                               uint32_t* flags, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  return decode_pos_ld(ctx, in, n, B, n, is_logits, soft, pos, flags, (cudaStream_t)stream);
}

// ---- encode (label side) and error statistics ---------------------------------------------------------------
extern "C" int spef_encode_ori(spef_ctx* ctx, const double* quat_dev, int32_t B, int32_t n, double variance, const uint8_t* masked_dev,
                               float* out_dev, uint32_t* flags_dev, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (!quat_dev || !out_dev || B < 1 || !(variance > 0.0)) return fail(ctx, SPEF_ERR_INVALID, "spef_encode_ori: bad argument");
  if (!ctx->ori_tab64) return fail(ctx, SPEF_ERR_STATE, "spef_encode_ori: orientation histogram not set (spef_set_ori_histogram)");
  if (n != ctx->ori_n) return fail(ctx, SPEF_ERR_INVALID, "spef_encode_ori: n = %d but the histogram has %d bins", n, ctx->ori_n);
  CK(cudaSetDevice(ctx->cfg.device));
  encode_kernel<true><<<B, 256, 0, (cudaStream_t)stream>>>(quat_dev, B, n, 1.0 / (2.0 * variance), ctx->ori_tab64, masked_dev, out_dev, flags_dev);
  CK_LAUNCH("encode_kernel<ori>");
  return SPEF_OK;
}

extern "C" int spef_encode_pos(spef_ctx* ctx, const double* pos_dev, int32_t B, int32_t n, double variance, float* out_dev, uint32_t* flags_dev,
                               void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (!pos_dev || !out_dev || B < 1 || !(variance > 0.0)) return fail(ctx, SPEF_ERR_INVALID, "spef_encode_pos: bad argument");
  if (!ctx->pos_tab64) return fail(ctx, SPEF_ERR_STATE, "spef_encode_pos: position histogram not set (spef_set_pos_histogram)");
  if (n != ctx->pos_n) return fail(ctx, SPEF_ERR_INVALID, "spef_encode_pos: n = %d but the histogram has %d bins", n, ctx->pos_n);
  CK(cudaSetDevice(ctx->cfg.device));
  encode_kernel<false><<<B, 256, 0, (cudaStream_t)stream>>>(pos_dev, B, n, 1.0 / (2.0 * variance), ctx->pos_tab64, nullptr, out_dev, flags_dev);
  CK_LAUNCH("encode_kernel<pos>");
  return SPEF_OK;
}

extern "C" int spef_error_stats(spef_ctx* ctx, const float* x_dev, int32_t stride, int32_t n, double* out_host, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (!x_dev || !out_host || stride < 1 || n < 1) return fail(ctx, SPEF_ERR_INVALID, "spef_error_stats: bad argument");
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = nullptr;
  double* out_d = nullptr;
  CK(cudaMallocAsync((void**)&ws, (size_t)n * 8, st));
  CK(cudaMallocAsync((void**)&out_d, 4 * sizeof(double), st));
  error_stats_kernel<<<1, 1024, 0, st>>>(x_dev, stride, n, ws, ws + n, out_d);
  cudaError_t le = cudaGetLastError();
  if (le == cudaSuccess) { ctx->launches++; cudaMemcpyAsync(out_host, out_d, 4 * sizeof(double), cudaMemcpyDeviceToHost, st); }
  cudaFreeAsync(ws, st);
  cudaFreeAsync(out_d, st);
  if (le != cudaSuccess) return fail(ctx, SPEF_ERR_CUDA, "launch of error_stats_kernel failed: %s", cudaGetErrorString(le));
  CK(cudaStreamSynchronize(st));
  return SPEF_OK;
}

static int score_internal(spef_ctx* ctx, const float* qp, const float* tp, const float* qt, const float* tt, int32_t B, double* sums,
                          float* per_image, const uint32_t* flags, void* stream) {
  if (!qp || !tp || !qt || !tt || !sums || B < 1) return fail(ctx, SPEF_ERR_INVALID, "spef_score: bad argument");
  CK(cudaSetDevice(ctx->cfg.device));
  int grid = cdiv(B, 256);
  if (grid > 4 * ctx->num_sms) grid = 4 * ctx->num_sms;
  launch_chain(ctx->pdl_on, score_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, qp, tp, qt, tt, B, sums, per_image, (const uint32_t*)flags);
  CK_LAUNCH("score_kernel");
  return SPEF_OK;
}

extern "C" int spef_score(spef_ctx* ctx, const float* qp, const float* tp, const float* qt, const float* tt, int32_t B, double* sums,
                          float* per_image, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  return score_internal(ctx, qp, tp, qt, tt, B, sums, per_image, nullptr, stream);
}

// ------------------------------------------------------------------------------------------------------
// fused predict / evaluation
// ------------------------------------------------------------------------------------------------------
static int predict_internal(spef_ctx* ctx, const float* images, int B, float* ori_soft, float* quat, float* pos_soft, float* pos,
                            int32_t* amax, uint32_t* flags, cudaStream_t st) {
  int rc = forward_internal(ctx, images, B, st, nullptr);
  if (rc) return rc;
  if (flags) CK(cudaMemsetAsync(flags, 0, (size_t)B * sizeof(uint32_t), st));
  rc = decode_ori_ld(ctx, ctx->head_out, ctx->head_pad, B, ctx->cfg.n_ori, 1, ori_soft, quat, nullptr, amax, flags, st);
  if (rc) return rc;
  if (ctx->cfg.pos_classification) {
    rc = decode_pos_ld(ctx, ctx->head_out + ctx->cfg.n_ori, ctx->head_pad, B, ctx->cfg.n_pos, 1, pos_soft, pos, flags, st);
    if (rc) return rc;
  } else {
    if (pos_soft) return fail(ctx, SPEF_ERR_INVALID, "predict: pos_soft requested but the position head is a regression head");
    rc = copy_head_out(ctx, B, nullptr, pos, st);
    if (rc) return rc;
  }
  return SPEF_OK;
}

extern "C" int spef_predict(spef_ctx* ctx, const float* images_dev, int32_t B, float* ori_soft, float* quat, float* pos_soft, float* pos,
                            int32_t* amax, uint32_t* flags, void* stream) {
  int rc = check_ready(ctx, B, "spef_predict");
  if (rc) return rc;
  if (!images_dev || !quat || !pos) return fail(ctx, SPEF_ERR_INVALID, "spef_predict: NULL argument");
  CK(cudaSetDevice(ctx->cfg.device));
  return predict_internal(ctx, images_dev, B, ori_soft, quat, pos_soft, pos, amax, flags, (cudaStream_t)stream);
}

static int grow(spef_ctx* ctx, float** p, size_t* have, size_t need) {
  if (*have >= need) return SPEF_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *have = 0;
  CK(cudaMalloc((void**)p, need * sizeof(float)));
  *have = need;
  return SPEF_OK;
}

static int ensure_ws_images(spef_ctx* ctx) {
  if (ctx->ws_images) return SPEF_OK;
  CK(cudaMalloc((void**)&ctx->ws_images, (size_t)ctx->cfg.max_batch * 3 * ctx->cfg.img_h * ctx->cfg.img_w * sizeof(float)));
  return SPEF_OK;
}

extern "C" int spef_predict_host(spef_ctx* ctx, const float* images_host, int32_t B, float* ori_soft_h, float* quat_h, float* pos_soft_h,
                                 float* pos_h, int32_t* amax_h, uint32_t* flags_h, void* stream) {
  int rc = check_ready(ctx, B, "spef_predict_host");
  if (rc) return rc;
  if (!images_host || !quat_h || !pos_h) return fail(ctx, SPEF_ERR_INVALID, "spef_predict_host: NULL argument");
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = ensure_ws_images(ctx))) return rc;
  const size_t img_bytes = (size_t)B * 3 * ctx->cfg.img_h * ctx->cfg.img_w * (ctx->image_u8 ? 1 : sizeof(float));
  CK(cudaMemcpyAsync(ctx->ws_images, images_host, img_bytes, cudaMemcpyHostToDevice, st));
  float* soft_d = nullptr;
  float* psoft_d = nullptr;
  if (ori_soft_h) {
    if ((rc = grow(ctx, &ctx->ws_soft, &ctx->ws_soft_elems, (size_t)ctx->cfg.max_batch * ctx->cfg.n_ori))) return rc;
    soft_d = ctx->ws_soft;
  }
  if (pos_soft_h) {
    if ((rc = grow(ctx, &ctx->ws_soft2, &ctx->ws_soft2_elems, (size_t)ctx->cfg.max_batch * ctx->cfg.n_pos))) return rc;
    psoft_d = ctx->ws_soft2;
  }
  rc = predict_internal(ctx, ctx->ws_images, B, soft_d, ctx->ws_quat, psoft_d, ctx->ws_pos, amax_h ? ctx->ws_argmax : nullptr, ctx->ws_flags, st);
  if (rc) return rc;
  if (ori_soft_h) CK(cudaMemcpyAsync(ori_soft_h, soft_d, (size_t)B * ctx->cfg.n_ori * 4, cudaMemcpyDeviceToHost, st));
  if (pos_soft_h) CK(cudaMemcpyAsync(pos_soft_h, psoft_d, (size_t)B * ctx->cfg.n_pos * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(quat_h, ctx->ws_quat, (size_t)B * 16, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(pos_h, ctx->ws_pos, (size_t)B * 12, cudaMemcpyDeviceToHost, st));
  if (amax_h) CK(cudaMemcpyAsync(amax_h, ctx->ws_argmax, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  if (flags_h) CK(cudaMemcpyAsync(flags_h, ctx->ws_flags, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SPEF_OK;
}

extern "C" int spef_eval_reset(spef_ctx* ctx, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemsetAsync(ctx->eval_sums, 0, 8 * sizeof(double), (cudaStream_t)stream));
  return SPEF_OK;
}

extern "C" double* spef_eval_sums_dev(spef_ctx* ctx) { return ctx ? ctx->eval_sums : nullptr; }

extern "C" int spef_eval_batch(spef_ctx* ctx, const float* images_dev, const float* qt, const float* tt, int32_t B, float* per_image, void* stream) {
  int rc = check_ready(ctx, B, "spef_eval_batch");
  if (rc) return rc;
  if (!images_dev || !qt || !tt) return fail(ctx, SPEF_ERR_INVALID, "spef_eval_batch: NULL argument");
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  auto run = [&](cudaStream_t s_) -> int {
    int r = predict_internal(ctx, images_dev, B, nullptr, ctx->ws_quat, nullptr, ctx->ws_pos, nullptr, ctx->ws_flags, s_);
    if (r) return r;
    // the decode guard flags of this batch are counted into sums[6] / sums[7] (the reference raises ValueError from decode())
    return score_internal(ctx, ctx->ws_quat, ctx->ws_pos, qt, tt, B, ctx->eval_sums, per_image, ctx->ws_flags, s_);
  };
  if (ctx->eval_graph) {
    // same mechanism as spef_temporal_step: a signature that comes back is captured once (on an internal stream: the caller's may be
    // the legacy default stream, which cannot capture) and replayed from then on; everything else the step touches is owned by the ctx
    spef_ctx::EGraph sig{images_dev, qt, tt, per_image, B, nullptr, 0};
    auto same = [&](const spef_ctx::EGraph& g) { return g.img == sig.img && g.qt == sig.qt && g.tt == sig.tt && g.per == sig.per && g.B == sig.B; };
    for (auto& g : ctx->egraphs)
      if (same(g)) {
        CK(cudaGraphLaunch(g.exec, st));
        ctx->launches += g.launches;
        return SPEF_OK;
      }
    bool seen = false;
    for (auto& g : ctx->eg_seen) seen = seen || same(g);
    if (seen) {
      if (!ctx->cap_stream) CK(cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
      const int64_t l0 = ctx->launches;
      CK(cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeThreadLocal));
      rc = run(ctx->cap_stream);
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(ctx->cap_stream, &graph);
      const int64_t n_launch = ctx->launches - l0;
      ctx->launches = l0;   // nothing has run yet
      if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        ctx->eval_graph = 0;   // capture is not possible in this process: direct launches from now on
        if (rc) return rc;
      } else {
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie == cudaSuccess) {
          if (ctx->egraphs.size() >= 4) { cudaGraphExecDestroy(ctx->egraphs.front().exec); ctx->egraphs.erase(ctx->egraphs.begin()); }
          sig.exec = exec; sig.launches = n_launch;
          ctx->egraphs.push_back(sig);
          CK(cudaGraphLaunch(exec, st));
          ctx->launches += n_launch;
          return SPEF_OK;
        }
        cudaGetLastError();
        ctx->eval_graph = 0;
      }
    } else {
      if (ctx->eg_seen.size() >= 4) ctx->eg_seen.erase(ctx->eg_seen.begin());
      ctx->eg_seen.push_back(sig);
    }
  }
  return run(st);
}

extern "C" int spef_eval_batch_host(spef_ctx* ctx, const float* images_host, const float* qt_h, const float* tt_h, int32_t B,
                                    float* per_image_h, void* stream) {
  int rc = check_ready(ctx, B, "spef_eval_batch_host");
  if (rc) return rc;
  if (!images_host || !qt_h || !tt_h) return fail(ctx, SPEF_ERR_INVALID, "spef_eval_batch_host: NULL argument");
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = ensure_ws_images(ctx))) return rc;
  CK(cudaMemcpyAsync(ctx->ws_images, images_host, (size_t)B * 3 * ctx->cfg.img_h * ctx->cfg.img_w * (ctx->image_u8 ? 1 : sizeof(float)), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->ws_qt, qt_h, (size_t)B * 16, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->ws_tt, tt_h, (size_t)B * 12, cudaMemcpyHostToDevice, st));
  rc = spef_eval_batch(ctx, ctx->ws_images, ctx->ws_qt, ctx->ws_tt, B, per_image_h ? ctx->ws_per_image : nullptr, stream);
  if (rc) return rc;
  if (per_image_h) {
    CK(cudaMemcpyAsync(per_image_h, ctx->ws_per_image, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  return SPEF_OK;
}

namespace spef_host {
void pack_bf16(const float* src, uint16_t* dst, size_t n);
int pack_threads();
}

// BF16 pixels -> the float tensor of the image contract (exact widening); 8 pixels per thread
__global__ void __launch_bounds__(256) widen_bf16_kernel(const uint4* __restrict__ in, float4* __restrict__ out, size_t n8) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = in[i];
    out[2 * i] = make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
    out[2 * i + 1] = make_float4(__uint_as_float(v.z << 16), __uint_as_float(v.z & 0xffff0000u), __uint_as_float(v.w << 16), __uint_as_float(v.w & 0xffff0000u));
  }
}

static bool pack_applies(const spef_ctx* ctx) {
  // only where the consumer's first step is the same rounding: the tcgen05 stem (GEMM or fused into block 1) of the BF16 engine
  return ctx->host_pack && !ctx->image_u8 && ctx->cfg.precision == SPEF_BF16 && ctx->cfg.pw_impl == 0 && !ctx->stem_simt &&
         ((size_t)3 * ctx->cfg.img_h * ctx->cfg.img_w) % 8 == 0;
}

static int ensure_pack(spef_ctx* ctx) {
  if (ctx->pack_host[0]) return SPEF_OK;
  const size_t n = (size_t)ctx->cfg.max_batch * 3 * ctx->cfg.img_h * ctx->cfg.img_w;
  for (int s = 0; s < 2; ++s) {
    CK(cudaHostAlloc((void**)&ctx->pack_host[s], n * 2, cudaHostAllocDefault));
    CK(cudaMalloc((void**)&ctx->pack_dev[s], n * 2));
    CK(cudaEventCreateWithFlags(&ctx->pack_uploaded[s], cudaEventDisableTiming));
    CK(cudaEventCreate(&ctx->pack_t0[s]));
    CK(cudaEventCreate(&ctx->pack_t1[s]));
  }
  return SPEF_OK;
}

// images_host (float) -> dst_dev (float) on `cs` through staging slot s.  The batch is split: the tail goes as it is (the DMA engine
// needs no help with it and starts at once), the head is rounded to BF16 chunk by chunk on the host threads, each chunk handed to the DMA
// engine while the next one converts, and widened on the device.  The split balances the two resources from what this context measures
// while it runs -- r_c, the float bytes per second the host threads convert (host clock around the conversions, i.e. with the DMA traffic
// competing for the same memory), and r_d, the bytes per second of the plain slice's copy (events around it):
//   head fraction f:  f / r_c = (f / 2 + 1 - f) / r_d   ->   f = r_c / (r_d + r_c / 2),   capped at 15/16 (there is always a slice to time).
static int upload_packed(spef_ctx* ctx, int s, const float* images_host, float* dst_dev, int B, cudaStream_t cs, cudaEvent_t dst_free) {
  int rc = ensure_pack(ctx);
  if (rc) return rc;
  const size_t n = (size_t)B * 3 * ctx->cfg.img_h * ctx->cfg.img_w;
  // the staging slot is free once the copies of its previous use have left the host (normally long ago: two submits back)
  CK(cudaEventSynchronize(ctx->pack_uploaded[s]));
  if (ctx->pack_plain_bytes[s] > 0) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->pack_t0[s], ctx->pack_t1[s]) == cudaSuccess && ms > 0.f) {
      const double obs = (double)ctx->pack_plain_bytes[s] / (ms * 1e-3);
      ctx->pack_rd = obs > ctx->pack_rd * 0.95 ? obs : ctx->pack_rd * 0.95;   // the best recent rate: a copy that shared the engine with the other lane reads low
    } else {
      (void)cudaGetLastError();
    }
    ctx->pack_plain_bytes[s] = 0;
  }
  double f = ctx->pack_rc / (ctx->pack_rd + 0.5 * ctx->pack_rc);
  if (ctx->pack_frac >= 0.0) f = ctx->pack_frac;
  const size_t BLK = 65536;   // split and chunk boundaries: whole 64 K-pixel blocks (staging rows stay 64-byte aligned, n % 8 holds for the widening)
  const size_t nblk = n / BLK;
  size_t head_blk = (size_t)(f * (double)nblk + 0.5);
  const size_t max_head = ctx->pack_frac >= 0.0 ? nblk : nblk - (nblk + 15) / 16;
  if (head_blk > max_head) head_blk = max_head;
  // The balance assumes two independent resources; they are not when the host's memory system is what limits the copies (8 ranks on
  // one 32-core host: the plain copies already run at the concurrent H2D roof, the conversion threads crawl at 3 GB/s and packing a
  // tenth of the batch cost 12 % end to end).  The model's own gain, 1 + r_c / (2 r_d), is the test: below 1.25 the batch goes
  // as it is, and every sixteenth submit packs a small probe slice so that r_c keeps being measured.
  if (ctx->pack_frac < 0.0) {
    const bool worth = ctx->pack_rc >= (ctx->pack_on ? 0.5 : 0.6) * ctx->pack_rd;
    ctx->pack_on = worth ? 1 : 0;
    if (!worth) head_blk = ((ctx->pack_calls & 15) == 0 && nblk >= 64) ? 8 : 0;
  }
  ctx->pack_calls++;
  size_t head = head_blk * BLK;
  if (nblk == 0 || ctx->pack_frac >= 1.0) head = n;             // a small batch (or a forced full pack): everything packed
  if (dst_free) CK(cudaStreamWaitEvent(cs, dst_free, 0));
  if (head < n) {
    CK(cudaEventRecord(ctx->pack_t0[s], cs));
    CK(cudaMemcpyAsync(dst_dev + head, images_host + head, (n - head) * 4, cudaMemcpyHostToDevice, cs));
    CK(cudaEventRecord(ctx->pack_t1[s], cs));
    ctx->pack_plain_bytes[s] = (n - head) * 4;
  }
  if (head > 0) {
    size_t chunk = ((head / 8 + BLK - 1) / BLK) * BLK;   // ~8 chunks
    if (chunk == 0 || head <= 16 * BLK) chunk = head;   // a probe slice is one chunk: its rate is what gets measured
    const auto t0 = std::chrono::steady_clock::now();
    for (size_t o = 0; o < head; o += chunk) {
      const size_t len = (o + chunk <= head) ? chunk : head - o;
      spef_host::pack_bf16(images_host + o, ctx->pack_host[s] + o, len);
      CK(cudaMemcpyAsync(ctx->pack_dev[s] + o, ctx->pack_host[s] + o, len * 2, cudaMemcpyHostToDevice, cs));
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (sec > 0 && head >= BLK) ctx->pack_rc = 0.5 * ctx->pack_rc + 0.5 * ((double)head * 4 / sec);
    CK(cudaEventRecord(ctx->pack_uploaded[s], cs));
    widen_bf16_kernel<<<ctx->num_sms * 4, 256, 0, cs>>>(reinterpret_cast<const uint4*>(ctx->pack_dev[s]), reinterpret_cast<float4*>(dst_dev), head / 8);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  const double fr = n ? (double)head / (double)n : 0.0;
  ctx->pack_last_frac = ctx->pack_calls <= 1 ? fr : 0.75 * ctx->pack_last_frac + 0.25 * fr;   // recent average
  return SPEF_OK;
}

static int ensure_pipe(spef_ctx* ctx) {
  if (ctx->pipe_copy_stream) return SPEF_OK;
  const size_t B = (size_t)ctx->cfg.max_batch;
  CK(cudaStreamCreateWithFlags(&ctx->pipe_copy_stream, cudaStreamNonBlocking));
  for (int s = 0; s < 2; ++s) {
    CK(cudaMalloc((void**)&ctx->pipe_images[s], B * 3 * ctx->cfg.img_h * ctx->cfg.img_w * sizeof(float)));
    CK(cudaMalloc((void**)&ctx->pipe_qt[s], B * 16));
    CK(cudaMalloc((void**)&ctx->pipe_tt[s], B * 12));
    CK(cudaMalloc((void**)&ctx->pipe_per[s], B * 8));
    CK(cudaEventCreateWithFlags(&ctx->pipe_copied[s], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->pipe_consumed[s], cudaEventDisableTiming));
  }
  return SPEF_OK;
}

extern "C" int spef_pack_bf16_host(const float* src_host, uint16_t* dst_host, int64_t n) {
  if (!src_host || !dst_host || n < 0) return SPEF_ERR_INVALID;
  spef_host::pack_bf16(src_host, dst_host, (size_t)n);
  return SPEF_OK;
}

extern "C" int spef_set_host_pack(spef_ctx* ctx, int32_t on) {
  if (!ctx) return SPEF_ERR_INVALID;
  ctx->host_pack = on ? 1 : 0;
  return SPEF_OK;
}

extern "C" int spef_host_pack_info(const spef_ctx* ctx, int32_t* active, int32_t* threads, double* stats) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (active) *active = pack_applies(ctx) ? 1 : 0;
  if (threads) *threads = pack_applies(ctx) ? spef_host::pack_threads() : 0;
  if (stats) { stats[0] = ctx->pack_last_frac; stats[1] = ctx->pack_rc; stats[2] = ctx->pack_rd; }
  return SPEF_OK;
}

extern "C" int spef_eval_submit_host(spef_ctx* ctx, const float* images_host, const float* qt_h, const float* tt_h, int32_t B,
                                     float* per_image_h, void* stream) {
  int rc = check_ready(ctx, B, "spef_eval_submit_host");
  if (rc) return rc;
  if (!images_host || !qt_h || !tt_h) return fail(ctx, SPEF_ERR_INVALID, "spef_eval_submit_host: NULL argument");
  CK(cudaSetDevice(ctx->cfg.device));
  if ((rc = ensure_pipe(ctx))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int s = ctx->pipe_slot;
  // the copy into slot s may start once the compute that last read slot s (two submits ago) has finished
  if (pack_applies(ctx)) {
    if ((rc = upload_packed(ctx, s, images_host, ctx->pipe_images[s], B, ctx->pipe_copy_stream, ctx->pipe_submitted >= 2 ? ctx->pipe_consumed[s] : nullptr))) return rc;
  } else {
    if (ctx->pipe_submitted >= 2) CK(cudaStreamWaitEvent(ctx->pipe_copy_stream, ctx->pipe_consumed[s], 0));
    CK(cudaMemcpyAsync(ctx->pipe_images[s], images_host, (size_t)B * 3 * ctx->cfg.img_h * ctx->cfg.img_w * (ctx->image_u8 ? 1 : sizeof(float)), cudaMemcpyHostToDevice, ctx->pipe_copy_stream));
  }
  CK(cudaMemcpyAsync(ctx->pipe_qt[s], qt_h, (size_t)B * 16, cudaMemcpyHostToDevice, ctx->pipe_copy_stream));
  CK(cudaMemcpyAsync(ctx->pipe_tt[s], tt_h, (size_t)B * 12, cudaMemcpyHostToDevice, ctx->pipe_copy_stream));
  CK(cudaEventRecord(ctx->pipe_copied[s], ctx->pipe_copy_stream));
  CK(cudaStreamWaitEvent(st, ctx->pipe_copied[s], 0));
  rc = spef_eval_batch(ctx, ctx->pipe_images[s], ctx->pipe_qt[s], ctx->pipe_tt[s], B, per_image_h ? ctx->pipe_per[s] : nullptr, stream);
  if (rc) return rc;
  if (per_image_h) CK(cudaMemcpyAsync(per_image_h, ctx->pipe_per[s], (size_t)B * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(ctx->pipe_consumed[s], st));
  ctx->pipe_slot ^= 1;
  ctx->pipe_submitted++;
  return SPEF_OK;
}

extern "C" int spef_eval_wait(spef_ctx* ctx, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  if (ctx->pipe_copy_stream) CK(cudaStreamSynchronize(ctx->pipe_copy_stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return SPEF_OK;
}

extern "C" int spef_eval_read(spef_ctx* ctx, double* sums_host, void* stream) {
  if (!ctx || !sums_host) return SPEF_ERR_INVALID;
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemcpyAsync(sums_host, ctx->eval_sums, 8 * sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return SPEF_OK;
}

// ---- host-buffer post-processing -----------------------------------------------------------------------
extern "C" int spef_decode_ori_host(spef_ctx* ctx, const float* in_h, int32_t B, int32_t n, int32_t is_logits, float* soft_h, float* quat_h,
                                    float* hinv_h, int32_t* amax_h, uint32_t* flags_h, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (!in_h || !quat_h || B < 1 || n < 1) return fail(ctx, SPEF_ERR_INVALID, "spef_decode_ori_host: bad argument");
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  const size_t elems = (size_t)B * n;
  if ((rc = grow(ctx, &ctx->ws_soft, &ctx->ws_soft_elems, elems))) return rc;
  if (soft_h && (rc = grow(ctx, &ctx->ws_soft2, &ctx->ws_soft2_elems, elems))) return rc;
  float* quat_d; float* hinv_d = nullptr; int32_t* amax_d = nullptr; uint32_t* flags_d;
  CK(cudaMallocAsync((void**)&quat_d, (size_t)B * 16, st));
  CK(cudaMallocAsync((void**)&flags_d, (size_t)B * 4, st));
  if (hinv_h) CK(cudaMallocAsync((void**)&hinv_d, (size_t)B * 64, st));
  if (amax_h) CK(cudaMallocAsync((void**)&amax_d, (size_t)B * 4, st));
  CK(cudaMemsetAsync(flags_d, 0, (size_t)B * 4, st));
  CK(cudaMemcpyAsync(ctx->ws_soft, in_h, elems * 4, cudaMemcpyHostToDevice, st));
  rc = decode_ori_ld(ctx, ctx->ws_soft, n, B, n, is_logits, soft_h ? ctx->ws_soft2 : nullptr, quat_d, hinv_d, amax_d, flags_d, st);
  if (rc == SPEF_OK) {
    if (soft_h) cudaMemcpyAsync(soft_h, ctx->ws_soft2, elems * 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(quat_h, quat_d, (size_t)B * 16, cudaMemcpyDeviceToHost, st);
    if (hinv_h) cudaMemcpyAsync(hinv_h, hinv_d, (size_t)B * 64, cudaMemcpyDeviceToHost, st);
    if (amax_h) cudaMemcpyAsync(amax_h, amax_d, (size_t)B * 4, cudaMemcpyDeviceToHost, st);
    if (flags_h) cudaMemcpyAsync(flags_h, flags_d, (size_t)B * 4, cudaMemcpyDeviceToHost, st);
  }
  cudaFreeAsync(quat_d, st);
  cudaFreeAsync(flags_d, st);
  if (hinv_d) cudaFreeAsync(hinv_d, st);
  if (amax_d) cudaFreeAsync(amax_d, st);
  if (rc) return rc;
  CK(cudaStreamSynchronize(st));
  return SPEF_OK;
}

extern "C" int spef_decode_pos_host(spef_ctx* ctx, const float* in_h, int32_t B, int32_t n, int32_t is_logits, float* soft_h, float* pos_h,
                                    uint32_t* flags_h, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (!in_h || !pos_h || B < 1 || n < 1) return fail(ctx, SPEF_ERR_INVALID, "spef_decode_pos_host: bad argument");
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  const size_t elems = (size_t)B * n;
  if ((rc = grow(ctx, &ctx->ws_soft, &ctx->ws_soft_elems, elems))) return rc;
  if (soft_h && (rc = grow(ctx, &ctx->ws_soft2, &ctx->ws_soft2_elems, elems))) return rc;
  float* pos_d; uint32_t* flags_d;
  CK(cudaMallocAsync((void**)&pos_d, (size_t)B * 12, st));
  CK(cudaMallocAsync((void**)&flags_d, (size_t)B * 4, st));
  CK(cudaMemsetAsync(flags_d, 0, (size_t)B * 4, st));
  CK(cudaMemcpyAsync(ctx->ws_soft, in_h, elems * 4, cudaMemcpyHostToDevice, st));
  rc = decode_pos_ld(ctx, ctx->ws_soft, n, B, n, is_logits, soft_h ? ctx->ws_soft2 : nullptr, pos_d, flags_d, st);
  if (rc == SPEF_OK) {
    if (soft_h) cudaMemcpyAsync(soft_h, ctx->ws_soft2, elems * 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(pos_h, pos_d, (size_t)B * 12, cudaMemcpyDeviceToHost, st);
    if (flags_h) cudaMemcpyAsync(flags_h, flags_d, (size_t)B * 4, cudaMemcpyDeviceToHost, st);
  }
  cudaFreeAsync(pos_d, st);
  cudaFreeAsync(flags_d, st);
  if (rc) return rc;
  CK(cudaStreamSynchronize(st));
  return SPEF_OK;
}

extern "C" int spef_score_host(spef_ctx* ctx, const float* qp_h, const float* tp_h, const float* qt_h, const float* tt_h, int32_t B,
                               double* sums_h, float* per_image_h, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (!qp_h || !tp_h || !qt_h || !tt_h || !sums_h || B < 1) return fail(ctx, SPEF_ERR_INVALID, "spef_score_host: bad argument");
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  float* buf;  // qp | qt | tp | tt | per_image
  const size_t nB = (size_t)B;
  CK(cudaMallocAsync((void**)&buf, nB * (4 + 4 + 3 + 3 + 2) * sizeof(float), st));
  float* qp = buf; float* qt = qp + nB * 4; float* tp = qt + nB * 4; float* tt = tp + nB * 3; float* pi = tt + nB * 3;
  CK(cudaMemcpyAsync(qp, qp_h, nB * 16, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(qt, qt_h, nB * 16, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(tp, tp_h, nB * 12, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(tt, tt_h, nB * 12, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(ctx->ws_sums, 0, 8 * sizeof(double), st));
  int rc = spef_score(ctx, qp, tp, qt, tt, B, ctx->ws_sums, per_image_h ? pi : nullptr, stream);
  if (rc == SPEF_OK) {
    cudaMemcpyAsync(sums_h, ctx->ws_sums, 8 * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (per_image_h) cudaMemcpyAsync(per_image_h, pi, nB * 8, cudaMemcpyDeviceToHost, st);
  }
  cudaFreeAsync(buf, st);
  if (rc) return rc;
  CK(cudaStreamSynchronize(st));
  return SPEF_OK;
}

// ------------------------------------------------------------------------------------------------------
// temporal
// ------------------------------------------------------------------------------------------------------
extern "C" int spef_temporal_reset(spef_ctx* ctx, int32_t S, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (S < 1 || S > ctx->cfg.max_batch) return fail(ctx, SPEF_ERR_INVALID, "spef_temporal_reset: n_streams %d outside [1, max_batch=%d]", S, ctx->cfg.max_batch);
  if (!ctx->cfg.pos_classification) return fail(ctx, SPEF_ERR_UNSUPPORTED, "temporal filtering requires a classification position head (src/temporal/inference.py:160-161)");
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  if (S != ctx->t_streams) {
    drop_graphs(ctx);   // captured launches hold the old state pointers
    ctx->tg_last = spef_ctx::TGraph{};
    void** ptrs[] = {(void**)&ctx->t_ori_state, (void**)&ctx->t_pos_state, (void**)&ctx->t_has, (void**)&ctx->t_prev_still, (void**)&ctx->t_prev_video};
    for (void** p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }
    for (int i = 0; i < 8; ++i) { if (ctx->t_ws[i]) cudaFree(ctx->t_ws[i]); ctx->t_ws[i] = nullptr; }
    const size_t no = ctx->cfg.n_ori, np = ctx->cfg.n_pos, s = S;
    CK(cudaMalloc((void**)&ctx->t_ori_state, s * no * 4));
    CK(cudaMalloc((void**)&ctx->t_pos_state, s * np * 4));
    CK(cudaMalloc((void**)&ctx->t_has, s * 4 * sizeof(int)));
    CK(cudaMalloc((void**)&ctx->t_prev_still, s * 16));
    CK(cudaMalloc((void**)&ctx->t_prev_video, s * 16));
    const size_t sz[8] = {s * no * 4, s * np * 4, s * 16, s * 12, s * no * 4, s * np * 4, s * 16, s * 12};
    for (int i = 0; i < 8; ++i) CK(cudaMalloc((void**)&ctx->t_ws[i], sz[i]));
    ctx->t_streams = S;
  }
  CK(cudaMemsetAsync(ctx->t_has, 0, (size_t)S * 4 * sizeof(int), st));
  return SPEF_OK;
}

static int temporal_from_logits(spef_ctx* ctx, const float* ori_logits, int ld_o, const float* pos_logits, int ld_p, int S,
                                int apply_filter, const spef_temporal_out* o, cudaStream_t st) {
  if (S != ctx->t_streams) return fail(ctx, SPEF_ERR_STATE, "temporal step: n_streams %d != %d set by spef_temporal_reset", S, ctx->t_streams);
  spef_temporal_out z;
  memset(&z, 0, sizeof(z));
  if (o) z = *o;
  float* still_os = z.still_ori_soft ? z.still_ori_soft : ctx->t_ws[0];
  float* still_ps = z.still_pos_soft ? z.still_pos_soft : ctx->t_ws[1];
  float* still_q = z.still_quat ? z.still_quat : ctx->t_ws[2];
  float* still_p = z.still_pos ? z.still_pos : ctx->t_ws[3];
  float* vid_os = z.video_ori_soft ? z.video_ori_soft : ctx->t_ws[4];
  float* vid_ps = z.video_pos_soft ? z.video_pos_soft : ctx->t_ws[5];
  float* vid_q = z.video_quat ? z.video_quat : ctx->t_ws[6];
  float* vid_p = z.video_pos ? z.video_pos : ctx->t_ws[7];
  float* d_o = z.ori_distance ? z.ori_distance : (float*)ctx->ws_per_image;          // [S] scratch
  float* d_p = z.pos_distance ? z.pos_distance : (float*)ctx->ws_per_image + S;
  const int no = ctx->cfg.n_ori, np = ctx->cfg.n_pos;
  int* has = ctx->t_has;
  if (z.flags) CK(cudaMemsetAsync(z.flags, 0, (size_t)S * 4, st));
  int rc;
  // still pose: softmax + decode (inference.py:131-133 -> spe_torch.py:75-76)
  if ((rc = decode_ori_ld(ctx, ori_logits, ld_o, S, no, 1, still_os, still_q, nullptr, nullptr, z.flags, st))) return rc;
  if ((rc = decode_pos_ld(ctx, pos_logits, ld_p, S, np, 1, still_ps, still_p, z.flags, st))) return rc;
  launch_chain(ctx->pdl_on, quat_continuity_kernel, dim3(cdiv(S, 128)), dim3(128), 0, st, still_q, ctx->t_prev_still, has + 2 * S, S);  // inference.py:136-144
  CK_LAUNCH("quat_continuity_kernel");
  if (!apply_filter) return SPEF_OK;
  // adaptive pdf filters (inference.py:38-39, 164-165)
  launch_chain(ctx->pdl_on, temporal_filter_kernel, dim3(S), dim3(256), 0, st, (const float*)still_os, no, ctx->t_ori_state, has, 0.8f, 16.49f, vid_os, d_o);
  CK_LAUNCH("temporal_filter_kernel");
  launch_chain(ctx->pdl_on, temporal_filter_kernel, dim3(S), dim3(256), 0, st, (const float*)still_ps, np, ctx->t_pos_state, has + S, 0.5f, 48.64f, vid_ps, d_p);
  CK_LAUNCH("temporal_filter_kernel");
  // decode of the filtered pdfs (inference.py:167-168)
  if ((rc = decode_ori_ld(ctx, vid_os, no, S, no, 0, nullptr, vid_q, nullptr, nullptr, z.flags, st))) return rc;
  if ((rc = decode_pos_ld(ctx, vid_ps, np, S, np, 0, nullptr, vid_p, z.flags, st))) return rc;
  launch_chain(ctx->pdl_on, quat_continuity_kernel, dim3(cdiv(S, 128)), dim3(128), 0, st, vid_q, ctx->t_prev_video, has + 3 * S, S);  // inference.py:173-180
  CK_LAUNCH("quat_continuity_kernel");
  return SPEF_OK;
}

extern "C" int spef_temporal_step_logits(spef_ctx* ctx, const float* ori_logits, const float* pos_logits, int32_t S,
                                         const spef_temporal_out* out, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (!ori_logits || !pos_logits) return fail(ctx, SPEF_ERR_INVALID, "spef_temporal_step_logits: NULL argument");
  CK(cudaSetDevice(ctx->cfg.device));
  return temporal_from_logits(ctx, ori_logits, ctx->cfg.n_ori, pos_logits, ctx->cfg.n_pos, S, 1, out, (cudaStream_t)stream);
}

extern "C" int spef_temporal_step(spef_ctx* ctx, const float* images_dev, int32_t S, int32_t apply_filter, const spef_temporal_out* out, void* stream) {
  int rc = check_ready(ctx, S, "spef_temporal_step");
  if (rc) return rc;
  if (!images_dev) return fail(ctx, SPEF_ERR_INVALID, "spef_temporal_step: NULL argument");
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  if (S != ctx->t_streams) return fail(ctx, SPEF_ERR_STATE, "temporal step: n_streams %d != %d set by spef_temporal_reset", S, ctx->t_streams);
  // Few streams: the step is ~40 short launches, i.e. launch-bound (0.42 ms per frame at batch 1 against 0.30 ms as one graph).
  // A call signature seen twice in a row is captured once (on an internal stream: the caller's may be the legacy default
  // stream, which cannot capture) and replayed from then on; callers that hand over new pointers every frame stay on direct
  // launches.  Everything the step touches besides its arguments is owned by the ctx (weights, activations, filter state).
  if (ctx->temporal_graph && S <= ctx->temporal_graph_max_streams) {
    spef_ctx::TGraph sig{};
    sig.img = images_dev; sig.S = S; sig.filt = apply_filter ? 1 : 0;
    if (out) sig.out = *out;
    auto same = [&](const spef_ctx::TGraph& g) {
      return g.img == sig.img && g.S == sig.S && g.filt == sig.filt && memcmp(&g.out, &sig.out, sizeof(sig.out)) == 0;
    };
    for (auto& g : ctx->tgraphs)
      if (same(g)) {
        CK(cudaGraphLaunch(g.exec, st));
        ctx->launches += g.launches;
        return SPEF_OK;
      }
    if (ctx->tg_last.S == S && same(ctx->tg_last)) {
      if (!ctx->cap_stream) CK(cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
      const int64_t l0 = ctx->launches;
      CK(cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeThreadLocal));
      rc = forward_internal(ctx, images_dev, S, ctx->cap_stream, nullptr);
      if (!rc) rc = temporal_from_logits(ctx, ctx->head_out, ctx->head_pad, ctx->head_out + ctx->cfg.n_ori, ctx->head_pad, S, apply_filter, out, ctx->cap_stream);
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(ctx->cap_stream, &graph);
      const int64_t n_launch = ctx->launches - l0;
      ctx->launches = l0;   // nothing has run yet
      if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        ctx->temporal_graph = 0;   // capture is not possible in this process: direct launches from now on
        if (rc) return rc;
      } else {
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie == cudaSuccess) {
          if (ctx->tgraphs.size() >= 4) { cudaGraphExecDestroy(ctx->tgraphs.front().exec); ctx->tgraphs.erase(ctx->tgraphs.begin()); }
          sig.exec = exec; sig.launches = n_launch;
          ctx->tgraphs.push_back(sig);
          CK(cudaGraphLaunch(exec, st));
          ctx->launches += n_launch;
          return SPEF_OK;
        }
        cudaGetLastError();
        ctx->temporal_graph = 0;
      }
    }
    ctx->tg_last = sig;
  }
  if ((rc = forward_internal(ctx, images_dev, S, st, nullptr))) return rc;
  return temporal_from_logits(ctx, ctx->head_out, ctx->head_pad, ctx->head_out + ctx->cfg.n_ori, ctx->head_pad, S, apply_filter, out, st);
}

extern "C" int spef_pdf_filter(spef_ctx* ctx, const float* cur, int32_t S, int32_t n, float* state, int32_t* has_state, float n_coef,
                               float alpha, float* out, float* distance, void* stream) {
  if (!ctx) return SPEF_ERR_INVALID;
  if (!cur || !state || !has_state || !out || !distance || S < 1 || n < 1) return fail(ctx, SPEF_ERR_INVALID, "spef_pdf_filter: bad argument");
  if (cur == out) return fail(ctx, SPEF_ERR_INVALID, "spef_pdf_filter: cur and out must not alias");
  CK(cudaSetDevice(ctx->cfg.device));
  temporal_filter_kernel<<<S, 256, 0, (cudaStream_t)stream>>>(cur, n, state, has_state, n_coef, alpha, out, distance);
  CK_LAUNCH("temporal_filter_kernel");
  return SPEF_OK;
}

// ------------------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------------------
extern "C" int64_t spef_launch_count(const spef_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int spef_forward_cost(const spef_ctx* ctx, int32_t B, double* bytes_out, double* flops_out) {
  if (!ctx) return SPEF_ERR_INVALID;
  double bytes = 0.0, flops = 0.0;
  const double e = (double)ctx->esz;
  for (const Layer& l : ctx->layers) {
    const double in_el = (double)l.hin * l.win * l.cin, out_el = (double)l.hout * l.wout * ((l.kind == K_HEAD) ? l.cout : l.cout);
    switch (l.kind) {
      case K_STEM: bytes += in_el * 4 + out_el * e; flops += 2.0 * 27 * out_el; break;
      case K_DW: bytes += (in_el + out_el) * e; flops += 2.0 * 9 * out_el; break;
      case K_PW: bytes += (in_el + out_el + (l.residual ? out_el : 0.0)) * e; flops += 2.0 * l.cin * out_el; break;
      case K_POOL: bytes += (in_el + out_el) * e; flops += in_el; break;
      case K_HEAD: bytes += in_el * e + out_el * 4; flops += 2.0 * l.cin * out_el; break;
    }
  }
  if (bytes_out) *bytes_out = bytes * B;
  if (flops_out) *flops_out = flops * B;
  return SPEF_OK;
}

// debug: the host-side tap tables of spef_resize_frames (float64 coefficient computation, 22-bit rounding), so that the CPU
// test-suite can pin them against the oracle / Pillow without a GPU.  coef_out holds out_size * ksize values (ksize returned).
extern "C" int spef_debug_resize_taps_host(int32_t in_size, int32_t out_size, int32_t* first_out, int32_t* count_out, int32_t* coef_out,
                                           int32_t coef_capacity, int32_t* ksize_out) {
  if (in_size < 1 || out_size < 1 || !first_out || !count_out || !coef_out || !ksize_out) return SPEF_ERR_INVALID;
  const ingest::AxisTaps t = ingest::make_axis_taps(in_size, out_size);
  *ksize_out = t.ksize;
  if ((long long)out_size * t.ksize > coef_capacity) return SPEF_ERR_INVALID;
  for (int o = 0; o < out_size; ++o) { first_out[o] = t.first[o]; count_out[o] = t.count[o]; }
  for (size_t i = 0; i < t.coef.size(); ++i) coef_out[i] = t.coef[i];
  return SPEF_OK;
}

// debug: the eigen-solve of the streaming decode kernel (f32 Jacobi + f64 polish, cofactor inverse) compiled for the host
extern "C" int spef_debug_decode_solve_host(const double* sums /*[11]*/, int32_t is_logits, float* quat /*[4]*/, float* hinv /*[16] or NULL*/) {
  if (!sums || !quat) return SPEF_ERR_INVALID;
  double tot[11];
  for (int k = 0; k < 11; ++k) tot[k] = sums[k];
  float q[4];
  if (!dstream::solve_core(tot, is_logits != 0, q, hinv)) return SPEF_ERR_INVALID;
  for (int k = 0; k < 4; ++k) quat[k] = q[k];
  return SPEF_OK;
}

// debug: the device Jacobi solver compiled for the host, so that the CPU test-suite can pin it against LAPACK
extern "C" int spef_debug_jacobi4_host(const double* a_in /*[16]*/, double* evals /*[4]*/, double* evecs /*[16], columns*/) {
  if (!a_in || !evals || !evecs) return SPEF_ERR_INVALID;
  double a[4][4], v[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) a[i][j] = a_in[i * 4 + j];
  jacobi4(a, v);
  for (int i = 0; i < 4; ++i) {
    evals[i] = a[i][i];
    for (int j = 0; j < 4; ++j) evecs[i * 4 + j] = v[i][j];
  }
  return SPEF_OK;
}
