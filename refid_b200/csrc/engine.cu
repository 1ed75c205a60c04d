// The REFID network engine: a static launch plan for one (B,T,H,W) problem.
//
// `Engine::build_network` restates FinalBidirectionAttenfusion.forward (reference
// basicsr/models/archs/XXNet_final_attenfusion_arch.py:130-218) as a sequence of tap-GEMM launches with fused
// epilogues plus a handful of memory-bound kernels, recording a tape; replaying the tape in reverse emits the
// backward plan (data-gradient tap-GEMMs with fused activation-derivative masks, weight-gradient GEMMs, reductions).
// All tensor maps are encoded once at plan time; forward/backward are then pure launch loops on the caller's stream.
//
// Gradient convention: for an activation tensor X = act(Z) the gradient buffer holds dL/dZ ("GZ"): every consumer
// multiplies its contribution by act'(.) (a linear mask) while accumulating, so the producer of X can feed the buffer
// straight into its weight-gradient and data-gradient GEMMs.
#include <algorithm>
#include <functional>
#include <stdlib.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/refid_b200.h"
#include "convop.cuh"
#include "elementwise.cuh"

namespace refid {

namespace {

enum SiteKind { SK_CONV3 = 0, SK_CONV1 = 1, SK_DOWN4 = 2, SK_UP2 = 3, SK_ROWS5 = 4, SK_RAW = 5 };

struct Site {
  std::string key;
  int kind, taps, R, Cc, nbias;
  long w_off, b_off;        // float offsets in the flat vector
  long fwd_off, dgrad_off;  // bf16 element offsets in the pack region (-1: none)
};

struct Pend {
  int g;     // gradient pool buffer (reference counted)
  long off;  // byte offset of the addend in the workspace
};

struct Ten {
  int N = 0, H = 0, W = 0, C = 0, pitch = 0;
  long off = -1;  // byte offset in the workspace
  int act = ACT_NONE;
  float slope = 0.f;
  long mask_off = -1;  // byte offset of the tensor act'(.) is evaluated on (own data, or the saved pre-activation)
  long bits_off = -1;  // byte offset of the activation sign bits (one word per pixel and 32 channels; see EpiDesc), -1: none
  int bits_pitch = 0;  // words per pixel
  bool need_grad = true;
  bool f32acc = false;  // gradient accumulated in fp32 (many contributions over time), converted when finalized
  long gfoff = -1;
  bool gfwritten = false;
  int gidx = -1;   // gradient pool buffer
  long goff = -1;  // byte offset of the bf16 GZ buffer
  bool gwritten = false;
  int parent = -1;
  long rel = 0;              // byte offset relative to the parent (views)
  int series = -1, slot = -1;  // time series this tensor is one slot of (contiguous over the T steps)
  std::vector<Pend> pending;  // gradient buffers still to be added (residual / skip-sum passthrough)
  bool contiguous() const { return pitch == C; }
  long elems() const { return (long)N * H * W * C; }
};

// A per-step tensor role (e.g. "the trunk's v at level 1, forward sweep") stored for all T steps in ONE contiguous array
// of T+1 slots, slot = time step (+1 in the forward sweep; the spare slot is the zero initial state of recurrent
// series).  Contiguity is what lets the weight gradient of a site be ONE GEMM over all T*B images at the end of the
// backward pass instead of T small ones (the same weights are applied at every step).
struct Series {
  long base = -1;       // byte offset of slot 0 in the workspace
  size_t slot_bytes = 0;
  int nslots = 0;       // T+1 on a training plan (everything is kept for the backward pass); 1 or 2 on a forward-only plan
  long gbase = -1;      // persistent gradient (dL/dZ) array, allocated when the producing site's wgrad is batched
  bool need_g = false;  // ... or when a deferred gradient needs all T output gradients at once
};

struct GBuf {
  size_t off, bytes;
  int refs;
};

struct ConvOp {
  int kind = CK_3X3;
  int site = -1;
  int in[2] = {-1, -1};
  int nin = 1;
  int out = -1;
  int act = ACT_NONE;
  float slope = 0.f;
  int res = -1, res2 = -1;     // residual addends (added before the activation)
  int post = -1, out2 = -1;    // out2 = out + post (skip sums)
  bool nchw_out = false;       // pred: fp32 NCHW store into the caller's output tensor
  long nchw_toff = 0, nchw_nstride = 0;
  int nchw_B = 0;              // > 0: the launch covers a chunk of steps, image n = (step n / B, sample n % B)
  long nchw_tstride = 0;
  bool no_tape = false;
  bool defer_in1 = false;  // in[1] is the same tensor at every step: its gradient is computed once from the summed dL/dZ
  bool rep_in1 = false;    // in[1] has fewer images than in[0] and repeats along the image axis (chunk ops; implies defer_in1)
  bool post_no_grad = false;  // the skip-sum addend's gradient is handed over elsewhere (once for all T steps)
  bool out2_grad_preadded = false;  // out2's gradient was registered as a pending addend of `out` before this level's backward
};

typedef std::function<int(cudaStream_t)> Launch;

// launch classes for the roofline accounting (bench.py)
// 4 / 5: the dominant kernel = stride-1 3x3 convs with >= 64 channels on both sides on the halo-conv engine (the EvR
// trunks, bottleneck, decoder trunks); 0 / 1: every other conv-shaped launch (1x1, 32-channel, stride-2, heads, pred)
enum LaunchClass { LC_CONV_FWD = 0, LC_CONV_DGRAD = 1, LC_WGRAD = 2, LC_OTHER = 3, LC_MAIN3_FWD = 4, LC_MAIN3_DGRAD = 5, LC_COUNT = 6 };
struct LaunchMeta {
  int cls;
  double flops;  // algorithmic FLOPs (2*MAC on un-padded channel counts); 0 for memory-bound kernels
  std::string label;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct Engine {
  refid_cfg cfg;
  int Kp_img = 0, Kp_ev = 0;
  std::vector<Site> sites;
  std::map<std::string, int> site_idx;
  long flat_floats = 0, pack_elems = 0, pack_max = 0;
  std::vector<PackDesc> packs;
  PackDesc* packs_dev = nullptr;

  // plan state
  bool planned = false, dry = false;
  int B = 0, T = 0, H = 0, W = 0, train = 0;
  int opt_infer_fp16 = 1;  // forward-only plans store activations / packed weights as fp16 (11 significant bits) instead of bf16
  int opt_graphs = 1;      // replay forward / backward as CUDA graphs
  int f16 = 0;             // storage flavour of the current plan
  bool packed_f16 = false;
  char* ws = nullptr;
  float* wmaster = nullptr;
  __nv_bfloat16* wpackb = nullptr;
  float* gflat = nullptr;
  std::vector<Ten> tens;
  std::map<std::string, int> named;
  size_t act_top = 0, gbase = 0, gtop = 0;
  std::vector<GBuf> gbufs;
  std::multimap<size_t, int> gfree;
  std::vector<Launch> fwd, bwd;
  std::vector<LaunchMeta> fwd_meta, bwd_meta;
  std::vector<Launch>* cur = nullptr;
  std::vector<std::function<int()>> tape;
  std::vector<std::pair<size_t, size_t>> f32_zero;
  std::vector<int> release_after;  // pool buffers of pending addends consumed by the launch being built
  std::vector<Series> series;
  std::map<std::string, int> series_idx;
  int cur_slot = -1;       // slot of the per-step tensors being created (-1: outside the time sweeps)
  int step_seq = 0;        // running index of new_tensor calls inside the current step (= the tensor's role)
  std::string cur_sdir;    // "b" / "f": sweep the current step belongs to
  std::vector<std::vector<ConvOp>> site_calls;  // forward conv calls per site (plan time)
  std::vector<char> site_batched;               // weight/bias gradient of this site is one batched launch
  std::vector<int> site_rep;                    // bit k: input k of the batched site is ONE tensor repeated every step

  // CUDA graphs of the two launch lists.  The plan is static and its buffers fixed; only the caller's tensors (x, event,
  // out / grad_out) vary between calls, so an instantiated graph is keyed on those pointers.  A key is replayed from its
  // graph from its second sighting on (the first runs eagerly: one-off pointers are not worth a capture, and every kernel
  // is loaded and configured before any capture starts); a few graphs per list are kept, least recently used evicted.
  struct GraphSlot {
    const void* k[3] = {nullptr, nullptr, nullptr};
    cudaGraphExec_t exec = nullptr;
    unsigned long stamp = 0;
    int seen = 0;
  };
  static constexpr int kGraphSlots = 8;
  GraphSlot fwd_graphs[kGraphSlots], bwd_graphs[kGraphSlots];
  unsigned long graph_clock = 0;
  cudaStream_t cap_stream = nullptr;
  std::string graph_error;
  long graph_captures = 0, graph_replays = 0, graph_eager = 0, graph_failures = 0;

  void drop_graphs() {
    for (GraphSlot* gs : {fwd_graphs, bwd_graphs})
      for (int i = 0; i < kGraphSlots; ++i) {
        if (gs[i].exec) cudaGraphExecDestroy(gs[i].exec);
        gs[i] = GraphSlot();
      }
  }

  int run_eager(std::vector<Launch>& ls, bool zero_grads, cudaStream_t st) {
    if (zero_grads) REFID_CUDA_CHECK(cudaMemsetAsync(gflat, 0, (size_t)flat_floats * 4, st));
    for (auto& l : ls)
      if (l(st)) return 1;
    return 0;
  }

  int run_list(std::vector<Launch>& ls, GraphSlot* slots, const void* k0, const void* k1, const void* k2, bool zero_grads,
               cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (!opt_graphs || cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      ++graph_eager;  // graphs off, or the caller is capturing this stream itself: plain launches join its capture
      return run_eager(ls, zero_grads, st);
    }
    GraphSlot* hit = nullptr;
    GraphSlot* lru = &slots[0];
    for (int i = 0; i < kGraphSlots; ++i) {
      if (slots[i].seen && slots[i].k[0] == k0 && slots[i].k[1] == k1 && slots[i].k[2] == k2) hit = &slots[i];
      if (slots[i].stamp < lru->stamp) lru = &slots[i];
    }
    if (hit && hit->exec) {
      hit->stamp = ++graph_clock;
      ++graph_replays;
      REFID_CUDA_CHECK(cudaGraphLaunch(hit->exec, st));
      return 0;
    }
    if (!hit) {  // first sighting: remember the key, run eagerly
      if (lru->exec) cudaGraphExecDestroy(lru->exec);
      *lru = GraphSlot();
      lru->k[0] = k0;
      lru->k[1] = k1;
      lru->k[2] = k2;
      lru->seen = 1;
      lru->stamp = ++graph_clock;
      ++graph_eager;
      return run_eager(ls, zero_grads, st);
    }
    // second sighting: capture, instantiate, launch.  The capture runs on the engine's own stream: the caller's stream is
    // usually PyTorch's current stream = the legacy default stream, which cannot be captured.  Capturing executes nothing,
    // so no ordering between the two streams is involved; the instantiated graph is then launched on the caller's stream.
    hit->stamp = ++graph_clock;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaSuccess;
    if (!cap_stream) ce = cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal);
    bool ok = ce == cudaSuccess;
    if (ok) {
      const int rc = run_eager(ls, zero_grads, cap_stream);
      ce = cudaStreamEndCapture(cap_stream, &graph);
      if (rc != 0) graph_error = std::string("launch failed during capture: ") + get_error();
      ok = rc == 0 && ce == cudaSuccess && graph != nullptr;
    }
    if (ok) {
      ce = cudaGraphInstantiate(&hit->exec, graph, 0);
      ok = ce == cudaSuccess;
    }
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {  // not capturable on this driver: keep working with plain launches (visible in refid_graph_stats / _error)
      if (ce != cudaSuccess) graph_error = cudaGetErrorString(ce);
      cudaGetLastError();
      hit->exec = nullptr;
      opt_graphs = 0;
      ++graph_failures;
      ++graph_eager;
      return run_eager(ls, zero_grads, st);
    }
    ++graph_captures;
    ++graph_replays;
    REFID_CUDA_CHECK(cudaGraphLaunch(hit->exec, st));
    return 0;
  }

  // per-call io
  const float* io_x = nullptr;
  const float* io_ev = nullptr;
  const float* io_gout = nullptr;
  float* io_out = nullptr;

  // ------------------------------------------------------------------------------------------
  // parameter sites
  // ------------------------------------------------------------------------------------------
  int add_site(const std::string& key, int kind, int taps, int R, int Cc, int nbias, bool fwd_pack, bool dgrad_pack) {
    Site s;
    s.key = key;
    s.kind = kind;
    s.taps = taps;
    s.R = R;
    s.Cc = Cc;
    s.nbias = nbias;
    s.w_off = flat_floats;
    flat_floats += (long)taps * R * Cc;
    s.b_off = -1;
    if (nbias > 0) {
      s.b_off = flat_floats;
      flat_floats += nbias;
    }
    flat_floats = (flat_floats + 3) / 4 * 4;  // keep every entry 16-byte aligned
    s.fwd_off = s.dgrad_off = -1;
    const long n = (long)taps * R * Cc;
    auto add_pack = [&](int transpose, const int* tapmap) {
      PackDesc d;
      memset(&d, 0, sizeof(d));
      d.src_off = s.w_off;
      d.dst_off = pack_elems;
      d.ntaps = taps;
      d.R = R;
      d.Cc = Cc;
      d.transpose = transpose;
      for (int i = 0; i < 16; ++i) d.tapmap[i] = (signed char)(i < taps ? tapmap[i] : 0);
      packs.push_back(d);
      const long off = pack_elems;
      pack_elems = (pack_elems + n + 63) / 64 * 64;  // 128-byte aligned matrices
      if (n > pack_max) pack_max = n;
      return off;
    };
    int ident[16], flip[16], down[16];
    for (int i = 0; i < 16; ++i) {
      ident[i] = i;
      flip[i] = taps - 1 - i;
    }
    static const int kidx[2][2] = {{1, 3}, {0, 2}};  // [output parity][slot] -> kernel index (convop.cu: down_dgrad_off)
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b) down[(py * 2 + px) * 4 + a * 2 + b] = kidx[py][a] * 4 + kidx[px][b];
    switch (kind) {
      case SK_CONV3:
      case SK_CONV1:
      case SK_ROWS5:
        if (fwd_pack) s.fwd_off = add_pack(1, ident);
        if (dgrad_pack) s.dgrad_off = add_pack(0, flip);
        break;
      case SK_DOWN4:
        if (fwd_pack) {
          // forward on the halo-conv engine: [tap9][Cout][parity q * Cin + ci] (K = 4*Cin), zero where view q does not
          // use the tap; view q = (py,px), tap (dy,dx): ky = 2*dy + py + 1, kx = 2*dx + px + 1
          s.fwd_off = pack_elems;
          for (int q = 0; q < 4; ++q) {
            PackDesc d;
            memset(&d, 0, sizeof(d));
            d.src_off = s.w_off;
            d.dst_off = pack_elems + (long)q * R;
            d.ntaps = 9;
            d.R = R;
            d.Cc = Cc;
            d.transpose = 1;
            d.dst_tap_stride = 4L * R * Cc;
            d.dst_pitch = 4 * R;
            for (int t9 = 0; t9 < 9; ++t9) {
              const int ky = 2 * (t9 / 3 - 1) + (q >> 1) + 1, kx = 2 * (t9 % 3 - 1) + (q & 1) + 1;
              d.tapmap[t9] = (signed char)((ky >= 0 && ky < 4 && kx >= 0 && kx < 4) ? ky * 4 + kx : -1);
            }
            packs.push_back(d);
          }
          pack_elems = (pack_elems + 36L * R * Cc + 63) / 64 * 64;
          if (9L * R * Cc > pack_max) pack_max = 9L * R * Cc;
          (void)ident;
        }
        if (dgrad_pack) {
          // data-gradient on the halo-conv engine: [tap9 of the dY neighbourhood][output parity q][Cin][Cout], zero blocks
          // where the parity does not use the tap (ky = qy + 1 - 2*dy, kx = qx + 1 - 2*dx outside the 4x4 kernel)
          s.dgrad_off = pack_elems;
          for (int q = 0; q < 4; ++q) {
            PackDesc d;
            memset(&d, 0, sizeof(d));
            d.src_off = s.w_off;
            d.dst_off = pack_elems + (long)q * R * Cc;
            d.ntaps = 9;
            d.R = R;
            d.Cc = Cc;
            d.transpose = 0;
            d.dst_tap_stride = 4L * R * Cc;
            for (int t9 = 0; t9 < 9; ++t9) {
              const int ky = (q >> 1) + 1 - 2 * (t9 / 3 - 1), kx = (q & 1) + 1 - 2 * (t9 % 3 - 1);
              d.tapmap[t9] = (signed char)((ky >= 0 && ky < 4 && kx >= 0 && kx < 4) ? ky * 4 + kx : -1);
            }
            packs.push_back(d);
          }
          const long n36 = 36L * R * Cc;
          pack_elems = (pack_elems + n36 + 63) / 64 * 64;
          if (9L * R * Cc > pack_max) pack_max = 9L * R * Cc;
          (void)down;
        }
        break;
      case SK_UP2:
        if (fwd_pack) s.fwd_off = add_pack(0, ident);
        if (dgrad_pack) s.dgrad_off = add_pack(1, ident);
        break;
      default:
        break;
    }
    site_idx[key] = (int)sites.size();
    sites.push_back(s);
    return (int)sites.size() - 1;
  }

  int site(const std::string& key) {
    auto it = site_idx.find(key);
    if (it == site_idx.end()) {
      set_error("unknown parameter site '%s'", key.c_str());
      return -1;
    }
    return it->second;
  }

  void build_sites() {
    const int b = cfg.base_num_channels;
    Kp_img = (5 * cfg.img_chn + 31) / 32 * 32;
    Kp_ev = (5 * cfg.ev_chn + 31) / 32 * 32;
    auto trunk = [&](const std::string& p, int c) {
      add_site(p + ".main.0", SK_CONV3, 9, 2 * c, c, c, true, true);
      add_site(p + ".main.2.0.conv1", SK_CONV3, 9, c, c, c, true, true);
      add_site(p + ".main.2.0.conv2", SK_CONV3, 9, c, c, c, true, true);
    };
    add_site("head_img", SK_ROWS5, 5, Kp_img, b, b, true, false);
    add_site("head", SK_ROWS5, 5, Kp_ev, b, b, true, false);
    for (int l = 0; l < 3; ++l) {
      const int cin = b << l, c = b << (l + 1);
      const std::string p = "img_encoders." + std::to_string(l);
      add_site(p + ".conv_1", SK_CONV3, 9, cin, c, c, true, true);
      add_site(p + ".conv_2", SK_CONV3, 9, c, c, c, true, true);
      add_site(p + ".identity", SK_CONV1, 1, cin, c, c, true, true);
      add_site(p + ".down", SK_DOWN4, 16, c, c, 0, true, true);
    }
    add_site("enc0_in", SK_CONV3, 9, b, 4 * b, 4 * b, true, true);  // both directions' level-0 in-convs, stacked on Cout
    const char* dirs[2] = {"encoders_backward", "encoders_forward"};
    for (int d = 0; d < 2; ++d)
      for (int l = 0; l < 3; ++l) {
        const int cin = b << l, c = b << (l + 1);
        const std::string p = std::string(dirs[d]) + "." + std::to_string(l);
        if (l == 2) add_site(p + ".conv", SK_CONV3, 9, cin, c, c, true, true);
        if (l == 1) {
          const std::string a = p + ".atten_fuse";
          add_site(a + ".conv1", SK_CONV1, 1, cin, cin, cin, true, true);    // norm1 affine folded in
          add_site(a + ".conv1_e", SK_CONV1, 1, cin, cin, cin, true, true);  // norm1_e affine folded in
          add_site(a + ".conv2", SK_RAW, 1, cin, 9, cin, false, false);
          add_site(a + ".conv2_e", SK_RAW, 1, cin, 9, cin, false, false);
          add_site(a + ".se_1.1", SK_RAW, 1, cin / 2, cin, cin / 2, false, false);
          add_site(a + ".se_1.3", SK_RAW, 1, cin, cin / 2, cin, false, false);
          add_site(a + ".conv3", SK_CONV1, 1, 2 * cin, cin, cin, true, true);   // beta folded in
          add_site(a + ".conv4", SK_CONV1, 1, cin, 2 * cin, 2 * cin, true, true);  // norm2 affine folded in
          add_site(a + ".conv5s", SK_CONV1, 1, 3 * cin, c, c, true, true);  // [conv_y_side | gamma*conv5] on K = [y ; gelu]
        }
        trunk(p + ".recurrent_block.forward_trunk", c);
        if (d == 1) add_site(p + ".fuse_two_dir", SK_CONV1, 1, 2 * c, c, c, true, true);
        if (!(d == 0 && l == 2)) add_site(p + ".down", SK_DOWN4, 16, c, c, 0, true, true);
      }
    for (int i = 0; i < 2; ++i) {
      const std::string p = "resblocks." + std::to_string(i);
      add_site(p + ".conv1", SK_CONV3, 9, 8 * b, 8 * b, 8 * b, true, true);
      add_site(p + ".conv2", SK_CONV3, 9, 8 * b, 8 * b, 8 * b, true, true);
    }
    for (int i = 0; i < 3; ++i) {
      const int cin = (8 * b) >> i, c = cin / 2;
      const std::string p = "decoders." + std::to_string(i);
      add_site(p + ".transposed_conv2d", SK_UP2, 4, c, cin, c, true, true);
      trunk(p + ".forward_trunk", c);
    }
    add_site("pred", SK_CONV3, 9, b, 32, 32, true, true);  // Cout padded out_chn -> 32
  }

  // ------------------------------------------------------------------------------------------
  // memory
  // ------------------------------------------------------------------------------------------
  __nv_bfloat16* P(long off) const { return reinterpret_cast<__nv_bfloat16*>(ws + off); }
  float* PF(long off) const { return reinterpret_cast<float*>(ws + off); }

  long act_alloc(size_t bytes) {
    const size_t off = act_top;
    act_top = align_up(act_top + bytes, 1024);
    return (long)off;
  }

  // Forward-only plans recycle per-step buffers, so the workspace is O(1) in T: a per-step role tensor lives in ONE slot
  // (it is dead at the end of its step), a recurrent state in two (step t reads slot t-1 and writes slot t, modulo 2).
  int new_series(size_t slot_bytes, int infer_slots = 1) {
    Series sr;
    sr.slot_bytes = slot_bytes;
    sr.nslots = train ? T + 1 : infer_slots;
    sr.base = act_alloc(slot_bytes * (size_t)sr.nslots);
    series.push_back(sr);
    return (int)series.size() - 1;
  }

  int series_tensor(int sidx, int slot, int N, int Hh, int Ww, int C, const std::string& name = "") {
    Ten t;
    t.N = N;
    t.H = Hh;
    t.W = Ww;
    t.C = C;
    t.pitch = C;
    t.series = sidx;
    t.slot = slot;
    t.off = series[sidx].base + (long)(slot % series[sidx].nslots) * (long)series[sidx].slot_bytes;
    tens.push_back(t);
    if (!name.empty()) named[name] = (int)tens.size() - 1;
    return (int)tens.size() - 1;
  }

  int new_tensor(int N, int Hh, int Ww, int C, const std::string& name = "") {
    if (cur_slot >= 0) {  // inside a time step: the tensor is one slot of its role's series
      const std::string key = cur_sdir + "#" + std::to_string(step_seq++);
      const size_t bytes = (size_t)N * Hh * Ww * C * 2;
      auto it = series_idx.find(key);
      int sidx;
      if (it == series_idx.end()) {
        sidx = new_series(bytes);
        series_idx[key] = sidx;
      } else {
        sidx = it->second;
        if (series[sidx].slot_bytes != bytes) {
          set_error("series %s: slot size changed between steps", key.c_str());
          return -1;
        }
      }
      return series_tensor(sidx, cur_slot, N, Hh, Ww, C, name);
    }
    Ten t;
    t.N = N;
    t.H = Hh;
    t.W = Ww;
    t.C = C;
    t.pitch = C;
    t.off = act_alloc((size_t)t.elems() * 2);
    tens.push_back(t);
    if (!name.empty()) named[name] = (int)tens.size() - 1;
    return (int)tens.size() - 1;
  }

  int view(int parent, int n0, int N, int c0, int C) {
    Ten t = tens[parent];
    t.pending.clear();
    t.rel = ((long)n0 * t.H * t.W * t.pitch + c0) * 2;
    t.off += t.rel;
    if (t.mask_off >= 0) t.mask_off += t.rel;
    if (t.bits_off >= 0) t.bits_off += ((long)n0 * t.H * t.W * t.bits_pitch + c0 / 32) * 4;
    t.N = N;
    t.C = C;
    t.parent = parent;
    t.gidx = -1;
    t.goff = -1;
    t.gwritten = false;
    tens.push_back(t);
    return (int)tens.size() - 1;
  }

  void mark_f32acc(int id) {
    if (!train) return;
    Ten& t = tens[id];
    t.f32acc = true;
    t.gfoff = act_alloc((size_t)t.elems() * 4);
    f32_zero.push_back({(size_t)t.gfoff, (size_t)t.elems() * 4});
  }

  int galloc(size_t bytes) {
    bytes = align_up(bytes, 1024);
    auto it = gfree.find(bytes);
    if (it != gfree.end()) {
      const int idx = it->second;
      gfree.erase(it);
      gbufs[idx].refs = 1;
      return idx;
    }
    GBuf g;
    g.off = gbase + gtop;
    g.bytes = bytes;
    g.refs = 1;
    gtop += bytes;
    gbufs.push_back(g);
    return (int)gbufs.size() - 1;
  }
  void gunref(int idx) {
    if (idx < 0) return;
    if (--gbufs[idx].refs == 0) gfree.insert({gbufs[idx].bytes, idx});
  }

  void ensure_gbuf(int id) {
    if (tens[id].goff >= 0) return;
    if (tens[id].parent >= 0) {
      const int p = tens[id].parent;
      ensure_gbuf(p);
      tens[p].gwritten = true;  // children fill the parent's buffer piecewise (each element exactly once)
      tens[id].gidx = tens[p].gidx;
      tens[id].goff = tens[p].goff + tens[id].rel;
      return;
    }
    if (tens[id].series >= 0 && series[tens[id].series].gbase >= 0) {  // persistent per-step gradient array
      const Series& sr = series[tens[id].series];
      tens[id].gidx = -1;
      tens[id].goff = sr.gbase + (long)tens[id].slot * (long)sr.slot_bytes;
      return;
    }
    tens[id].gidx = galloc((size_t)tens[id].elems() * 2);
    tens[id].goff = (long)gbufs[tens[id].gidx].off;
  }

  std::string cur_label;  // site / op the launches being emitted belong to (profiling only)
  void emit(Launch l, int cls = LC_OTHER, double flops = 0.0, const char* tag = "") {
    if (dry) return;
    cur->push_back(std::move(l));
    (cur == &fwd ? fwd_meta : bwd_meta).push_back(LaunchMeta{cls, flops, cur_label + tag});
  }

  bool main3x3(const ConvOp& op) const {
    const Site& s = sites[op.site];
    return op.kind == CK_3X3 && s.R >= 64 && s.Cc >= 64 && s.key != "pred";
  }

  // algorithmic FLOPs of one convolution (forward = data-gradient = weight-gradient)
  double conv_flops(const ConvOp& op) const {
    const Site& s = sites[op.site];
    const Ten& in0 = tens[op.in[0]];
    const double pix = (double)in0.N * in0.H * in0.W;
    if (op.kind == CK_ROWS5) return 2.0 * pix * 25.0 * (s.key == "head" ? cfg.ev_chn : cfg.img_chn) * s.Cc;
    if (op.kind == CK_UP2) return 2.0 * pix * 4.0 * s.Cc * s.R;
    if (op.kind == CK_DOWN4) return 2.0 * (pix / 4.0) * 16.0 * s.R * s.Cc;
    const double cout = (s.key == "pred") ? cfg.out_chn : s.Cc;
    return 2.0 * pix * s.taps * s.R * cout;
  }

  // ------------------------------------------------------------------------------------------
  // gradient bookkeeping (plan time)
  // ------------------------------------------------------------------------------------------
  // Describes how a GEMM epilogue (or an elementwise kernel) adds its contribution into tensor `id`'s gradient.
  struct Target {
    __nv_bfloat16* dst = nullptr;
    const __nv_bfloat16* pre = nullptr;   // existing contents (accumulate)
    const __nv_bfloat16* pre2 = nullptr;  // one pending passthrough addend
    const __nv_bfloat16* sv = nullptr;
    const uint32_t* sv_bits = nullptr;  // sign-bit form of an LReLU / ReLU mask (preferred by the halo-conv epilogue)
    int sv_bits_pitch = 0;
    int act = ACT_NONE;
    float slope = 0.f;
    float* dstf = nullptr;
    int pitch = 0;
  };

  int flush_pending(int id, bool keep_one) {
    // Fold pending passthrough addends into the gradient buffer with the masked-accumulate kernel.
    while ((int)tens[id].pending.size() > (keep_one ? 1 : 0)) {
      if (!tens[id].contiguous()) {
        set_error("pending gradient on a strided view");
        return 1;
      }
      ensure_gbuf(id);
      AddMaskArgs a = {};
      const int g0 = tens[id].pending.back().g;
      a.a = P(tens[id].pending.back().off);
      tens[id].pending.pop_back();
      int g1 = -1;
      if ((int)tens[id].pending.size() > (keep_one ? 1 : 0)) {
        g1 = tens[id].pending.back().g;
        a.b = P(tens[id].pending.back().off);
        tens[id].pending.pop_back();
      }
      if (tens[id].act != ACT_NONE) {
        a.sv = P(tens[id].mask_off);
        a.act = tens[id].act == ACT_GELU ? ACT_MULT : tens[id].act;
        a.slope = tens[id].slope;
      }
      a.dst = P(tens[id].goff);
      a.dst_acc = tens[id].gwritten ? 1 : 0;
      a.n = tens[id].elems();
      tens[id].gwritten = true;
      emit([a](cudaStream_t s) { return launch_addmask(a, s); }, LC_OTHER, 0.0, (std::string("addmask:flush") + ":" + std::to_string(tens[id].H) + "x" + std::to_string(tens[id].C)).c_str());
      gunref(g0);
      gunref(g1);
    }
    return 0;
  }

  int target(int id, Target* t) {
    *t = Target();
    t->pitch = tens[id].pitch;
    if (tens[id].f32acc) {
      t->dstf = PF(tens[id].gfoff);
      tens[id].gfwritten = true;
      return 0;
    }
    if (flush_pending(id, true)) return 1;
    ensure_gbuf(id);
    t->dst = P(tens[id].goff);
    if (tens[id].gwritten) t->pre = t->dst;
    if (!tens[id].pending.empty()) {
      const int g = tens[id].pending.back().g;
      t->pre2 = P(tens[id].pending.back().off);
      tens[id].pending.pop_back();
      release_after.push_back(g);
    }
    if (tens[id].act != ACT_NONE) {
      t->sv = P(tens[id].mask_off);
      t->act = tens[id].act == ACT_GELU ? ACT_MULT : tens[id].act;
      t->slope = tens[id].slope;
      if (tens[id].act == ACT_LRELU && tens[id].bits_off >= 0) {
        t->sv_bits = reinterpret_cast<const uint32_t*>(ws + tens[id].bits_off);
        t->sv_bits_pitch = tens[id].bits_pitch;
      }
    }
    tens[id].gwritten = true;
    return 0;
  }
  void release_consumed() {
    for (int g : release_after) gunref(g);
    release_after.clear();
  }

  int add_pending(int id, int src_gidx, long src_off) {
    if (id < 0 || !tens[id].need_grad) return 0;
    if (tens[id].f32acc) {
      AddMaskArgs a = {};
      a.a = P(src_off);
      a.dstf = PF(tens[id].gfoff);
      a.n = tens[id].elems();
      tens[id].gfwritten = true;
      emit([a](cudaStream_t s) { return launch_addmask(a, s); }, LC_OTHER, 0.0, (std::string("addmask:f32acc") + ":" + std::to_string(tens[id].H) + "x" + std::to_string(tens[id].C)).c_str());
      return 0;
    }
    if (src_gidx >= 0) gbufs[src_gidx].refs++;
    tens[id].pending.push_back(Pend{src_gidx, src_off});
    return 0;
  }

  // Complete tensor id's gradient; *gz = nullptr when no gradient reaches it.
  int finalize(int id, const __nv_bfloat16** gz) {
    *gz = nullptr;
    if (!tens[id].need_grad) return 0;
    if (tens[id].parent >= 0 && !tens[id].gwritten && tens[id].pending.empty()) {
      const int p = tens[id].parent;
      if (tens[p].gwritten || tens[p].goff >= 0) {
        tens[id].gidx = tens[p].gidx;
        tens[id].goff = tens[p].goff + tens[id].rel;
        *gz = P(tens[id].goff);
      }
      return 0;
    }
    if (tens[id].f32acc) {
      if (!tens[id].gfwritten) return 0;
      ensure_gbuf(id);
      AddMaskArgs a = {};
      a.f = PF(tens[id].gfoff);
      if (tens[id].act != ACT_NONE) {
        a.sv = P(tens[id].mask_off);
        a.act = tens[id].act == ACT_GELU ? ACT_MULT : tens[id].act;
        a.slope = tens[id].slope;
      }
      a.dst = P(tens[id].goff);
      a.n = tens[id].elems();
      tens[id].gwritten = true;
      emit([a](cudaStream_t s) { return launch_addmask(a, s); }, LC_OTHER, 0.0, (std::string("addmask:f32fin") + ":" + std::to_string(tens[id].H) + "x" + std::to_string(tens[id].C)).c_str());
      *gz = P(tens[id].goff);
      return 0;
    }
    if (flush_pending(id, false)) return 1;
    if (!tens[id].gwritten) return 0;
    *gz = P(tens[id].goff);
    return 0;
  }
  void release_grad(int id) {
    if (tens[id].parent >= 0) return;
    if (tens[id].gidx >= 0) {
      gunref(tens[id].gidx);
      tens[id].gidx = -1;
    }
  }

  // ------------------------------------------------------------------------------------------
  // convolution ops
  // ------------------------------------------------------------------------------------------
  int conv(ConvOp op, const std::string& name = "") {
    if (op.site < 0) return -1;
    const Site& s = sites[op.site];
    const Ten in0 = tens[op.in[0]];
    const int cout = (op.kind == CK_UP2) ? s.R : s.Cc;
    int oh = in0.H, ow = in0.W;
    if (op.kind == CK_DOWN4) {
      oh /= 2;
      ow /= 2;
    } else if (op.kind == CK_UP2) {
      oh *= 2;
      ow *= 2;
    }
    if (op.out < 0 && !op.nchw_out) {
      op.out = new_tensor(in0.N, oh, ow, cout, name);
      tens[op.out].act = op.act;
      tens[op.out].slope = op.slope;
      if (op.act == ACT_LRELU) tens[op.out].mask_off = tens[op.out].off;
      if (op.act == ACT_GELU && train) tens[op.out].mask_off = act_alloc((size_t)tens[op.out].elems() * 2);
    }
    else if (op.out >= 0 && !name.empty()) named[name] = op.out;
    // training plans: LeakyReLU / ReLU outputs also get their derivative mask as sign bits (1/16 of the tensor's bytes);
    // the data-gradient epilogues that target the tensor read those instead of the tensor (allocated in dry plans too)
    if (train && op.act == ACT_LRELU && op.out >= 0 && !op.nchw_out && tens[op.out].parent < 0 && tens[op.out].bits_off < 0 &&
        cout % 32 == 0) {
      const int bt = new_tensor(in0.N, oh, ow, cout / 16);  // cout / 32 words = cout / 16 16-bit elements per pixel
      tens[bt].need_grad = false;
      tens[op.out].bits_off = tens[bt].off;
      tens[op.out].bits_pitch = cout / 32;
    }
    if (op.post >= 0 && op.out2 < 0) op.out2 = new_tensor(in0.N, oh, ow, cout, name.empty() ? "" : name + "+");
    if (!dry) {
      cur_label = s.key;
      ConvDesc d;
      memset(&d, 0, sizeof(d));
      d.kind = op.kind;
      d.nsrc = op.nin;
      int cin_total = 0;
      for (int k = 0; k < op.nin; ++k) {
        const Ten& t = tens[op.in[k]];
        d.src[k] = {P(t.off), t.C, t.pitch, (k == 1 && op.rep_in1) ? t.N : 0};
        cin_total += t.C;
      }
      d.N = in0.N;
      d.H = in0.H;
      d.W = in0.W;
      d.w = wpackb + s.fwd_off;
      d.w_rows = (long)s.taps * cout;
      d.w_cols = cin_total;
      d.wrows_per_tap = cout;
      d.w_row0 = 0;
      d.f16 = f16;
      if (op.kind == CK_DOWN4) {  // forward of the stride-2 conv runs as a masked 3x3 over the four parity views
        d.kind = CK_DOWN4_HALO;
        d.w_rows = 9L * cout;
        d.w_cols = 4 * cin_total;
      }
      const int expect_cin = (op.kind == CK_UP2) ? s.Cc : s.R;
      if (cin_total != expect_cin) {
        set_error("conv %s: input channels %d != %d", s.key.c_str(), cin_total, expect_cin);
        return -1;
      }
      OutGroup g;
      memset(&g, 0, sizeof(g));
      g.channels = cout;
      EpiDesc& e = g.epi;
      if (op.nchw_out) {
        e.C = cout;
        e.nchw_C = cfg.out_chn;
        e.nchw_nstride = op.nchw_nstride;
        e.nchw_B = op.nchw_B;
        e.nchw_tstride = op.nchw_tstride;
      } else {
        const Ten& o = tens[op.out];
        e.out = P(o.off);
        e.C = o.pitch;
        if (op.act == ACT_GELU && o.mask_off >= 0) e.out_pre = P(o.mask_off);
        if (op.act == ACT_LRELU && o.bits_off >= 0) {
          e.out_bits = reinterpret_cast<uint32_t*>(ws + o.bits_off);
          e.out_bits_pitch = o.bits_pitch;
        }
      }
      e.bias = s.b_off >= 0 ? wmaster + s.b_off : nullptr;
      e.act = op.act;
      e.slope = op.slope;
      if (op.res >= 0) e.pre = P(tens[op.res].off);
      if (op.res2 >= 0) e.pre2 = P(tens[op.res2].off);
      if (op.out2 >= 0) {
        e.out2 = P(tens[op.out2].off);
        e.post = P(tens[op.post].off);
      }
      TapGemmLaunch l;
      if (build_conv(d, &g, 1, &l)) return -1;
      if (!l.use_halo && !op.nchw_out) tens[op.out].bits_off = -1;  // only the halo-conv epilogue writes the sign bits
      if (op.nchw_out) {
        Engine* self = this;
        const long toff = op.nchw_toff;
        emit([l, self, toff](cudaStream_t st) mutable {
          for (int i = 0; i < l.num_epi(); ++i) l.epi(i)->out_nchw = self->io_out + toff;
          return run_conv(l, st);
        }, main3x3(op) ? LC_MAIN3_FWD : LC_CONV_FWD, conv_flops(op), ":fwd");
      } else {
        emit([l](cudaStream_t st) mutable { return run_conv(l, st); }, main3x3(op) ? LC_MAIN3_FWD : LC_CONV_FWD, conv_flops(op), ":fwd");
      }
    }
    if (train && !op.no_tape) {
      Engine* self = this;
      tape.push_back([self, op]() { return self->conv_bwd(op); });
      if (site_calls.size() < sites.size()) site_calls.resize(sites.size());
      site_calls[op.site].push_back(op);
      if (op.defer_in1 && op.out >= 0 && tens[op.out].series >= 0) series[tens[op.out].series].need_g = true;
    }
    return op.out2 >= 0 && op.out < 0 ? op.out2 : op.out;
  }

  int emit_colsum(const __nv_bfloat16* gz, long rows, int C, float* dst) {
    emit([gz, rows, C, dst](cudaStream_t s) { return launch_colsum(gz, rows, C, dst, s); }, LC_OTHER, 0.0, ":colsum");
    return 0;
  }

  int conv_bwd(const ConvOp& op) {
    const Site& s = sites[op.site];
    cur_label = "";
    if (op.out2 >= 0) {
      const __nv_bfloat16* gs = nullptr;
      if (finalize(op.out2, &gs)) return 1;
      if (gs) {
        const int g = tens[op.out2].gidx;
        const long go = tens[op.out2].goff;
        if (!op.out2_grad_preadded) add_pending(op.out, g, go);
        if (!op.post_no_grad) add_pending(op.post, g, go);
      }
      release_grad(op.out2);
    }
    const __nv_bfloat16* gz = nullptr;
    if (finalize(op.out, &gz)) return 1;
    if (!gz) {
      REFID_REQUIRE(!(op.site < (int)site_batched.size() && site_batched[op.site]),
                    "no gradient reaches one step of batched site %s", s.key.c_str());
      return 0;
    }
    cur_label = s.key;
    const Ten o = tens[op.out];
    const Ten in0 = tens[op.in[0]];
    const int cout = o.C;
    int cin_total = 0;
    for (int k = 0; k < op.nin; ++k) cin_total += tens[op.in[k]].C;
    const bool batched = op.site < (int)site_batched.size() && site_batched[op.site];
    // bias gradient
    if (s.b_off >= 0 && !batched) {
      REFID_REQUIRE(o.contiguous(), "bias gradient on a strided tensor (%s)", s.key.c_str());
      emit_colsum(gz, (long)o.N * o.H * o.W, cout, gflat + s.b_off);
    }
    // weight gradient (per call; batched sites get one launch over all steps at the end of the backward plan)
    if (!dry && !batched) {
      ConvDesc d;
      memset(&d, 0, sizeof(d));
      ActSrc q;
      if (op.kind == CK_UP2) {
        d.kind = CK_UP2_DGRAD;
        d.src[0] = {gz, o.C, o.pitch};
        d.nsrc = 1;
        d.N = o.N;
        d.H = o.H;
        d.W = o.W;
        q = {P(in0.off), in0.C, in0.pitch};
      } else {
        d.kind = op.kind;
        d.nsrc = op.nin;
        for (int k = 0; k < op.nin; ++k) {
          const Ten& t = tens[op.in[k]];
          d.src[k] = {P(t.off), t.C, t.pitch, (k == 1 && op.rep_in1) ? t.N : 0};
        }
        d.N = in0.N;
        d.H = in0.H;
        d.W = in0.W;
        q = {gz, o.C, o.pitch};
      }
      WgradLaunch wl;
      if (build_wgrad(d, q, gflat + s.w_off, &wl)) return 1;
      emit([wl](cudaStream_t st) mutable { return run_wgrad(wl, st); }, LC_WGRAD, conv_flops(op), ":wgrad");
    }
    // data gradient
    if (s.dgrad_off >= 0) {
      int k = 0, chan0 = 0;
      while (k < op.nin) {
        if (!tens[op.in[k]].need_grad || (k == 1 && op.defer_in1)) {
          chan0 += tens[op.in[k]].C;
          ++k;
          continue;
        }
        // maximal run of inputs that need a gradient -> one launch
        OutGroup groups[2];
        memset(groups, 0, sizeof(groups));
        int ng = 0, w_row0 = chan0, grad_ch = 0;
        while (k < op.nin && tens[op.in[k]].need_grad && !(k == 1 && op.defer_in1)) {
          Target t;
          if (target(op.in[k], &t)) return 1;
          OutGroup& g = groups[ng++];
          g.channels = tens[op.in[k]].C;
          grad_ch += g.channels;
          g.epi.out = t.dst;
          g.epi.pre = t.pre;
          g.epi.pre2 = t.pre2;
          g.epi.sv = t.sv;
          g.epi.sv_bits = t.sv_bits;
          g.epi.sv_bits_pitch = t.sv_bits_pitch;
          g.epi.act = t.act;
          g.epi.slope = t.slope;
          g.epi.out_f32 = t.dstf;
          g.epi.C = t.pitch;
          chan0 += tens[op.in[k]].C;
          ++k;
        }
        if (!dry) {
          const bool down_halo = op.kind == CK_DOWN4 && o.C % 64 == 0 && ng == 1 && grad_ch == cin_total;
          REFID_REQUIRE(op.kind != CK_DOWN4 || down_halo, "down conv %s: unsupported channel count for the data-gradient", s.key.c_str());
          const int launches = 1;
          for (int par = 0; par < launches; ++par) {
            ConvDesc d;
            memset(&d, 0, sizeof(d));
                  d.src[0] = {gz, o.C, o.pitch};
            d.nsrc = 1;
            d.N = o.N;
            d.H = o.H;
            d.W = o.W;
            d.w = wpackb + s.dgrad_off;
            d.w_cols = cout;
            if (op.kind == CK_DOWN4) {
              d.kind = CK_DOWN4_DGRAD_HALO;
              d.w_rows = 36L * cin_total;
              d.wrows_per_tap = 4 * cin_total;
              d.w_row0 = w_row0;
            } else if (op.kind == CK_UP2) {
              d.kind = CK_UP2_DGRAD;
              d.w_rows = 4L * cin_total;
              d.wrows_per_tap = cin_total;
              d.w_row0 = w_row0;
            } else {
              d.kind = op.kind;
              d.w_rows = (long)s.taps * cin_total;
              d.wrows_per_tap = cin_total;
              d.w_row0 = w_row0;
            }
            TapGemmLaunch l;
            if (build_conv(d, groups, ng, &l)) return 1;
            emit([l](cudaStream_t st) mutable { return run_conv(l, st); }, main3x3(op) ? LC_MAIN3_DGRAD : LC_CONV_DGRAD,
                 conv_flops(op) * grad_ch / cin_total / launches, ":dgrad");
          }
        }
        release_consumed();
      }
    }
    cur_label = "";
    if (op.res >= 0) add_pending(op.res, tens[op.out].gidx, tens[op.out].goff);
    if (op.res2 >= 0) add_pending(op.res2, tens[op.out].gidx, tens[op.out].goff);
    release_grad(op.out);
    return 0;
  }

  // ------------------------------------------------------------------------------------------
  // EGACA pieces (fusion_modules.py:290-333)
  // ------------------------------------------------------------------------------------------
  int ln(int x, const std::string& name = "", int preset = -1) {
    const Ten tx = tens[x];
    const int y = preset >= 0 ? preset : new_tensor(tx.N, tx.H, tx.W, tx.C, name);
    const long npix = (long)tx.N * tx.H * tx.W;
    {
      const __nv_bfloat16 *px = P(tx.off);
      __nv_bfloat16* py = P(tens[y].off);
      cur_label = "";
      const int h16 = f16;
      emit([px, py, npix, h16](cudaStream_t s) { return launch_ln_fwd(px, py, npix, s, h16); }, LC_OTHER, 0.0, "ln_fwd");
    }
    if (train) {
      Engine* self = this;
      tape.push_back([self, x, y, npix]() {
        const __nv_bfloat16* gy = nullptr;
        if (self->finalize(y, &gy)) return 1;
        if (!gy || !self->tens[x].need_grad) return 0;
        if (self->tens[x].act != ACT_NONE) {
          set_error("LayerNorm input with an activation mask is unsupported");
          return 1;
        }
        Target t;
        if (self->target(x, &t)) return 1;
        const __nv_bfloat16* px = self->P(self->tens[x].off);
        self->emit([px, gy, t, npix](cudaStream_t s) {
          return launch_ln_bwd(px, gy, t.pre2, t.dst, t.pre ? 1 : 0, t.dstf, npix, s);
        }, LC_OTHER, 0.0, "ln_bwd");
        self->release_consumed();
        self->release_grad(y);
        return 0;
      });
    }
    return y;
  }

  // depthwise 3x3 + GELU (+ pooled sums); returns the GELU output tensor, *pre_out the saved pre-activation
  int dw(int a, int site_id, long pool_off, const std::string& name = "") {
    const Ten ta = tens[a];
    const Site s = sites[site_id];
    const int g = new_tensor(ta.N, ta.H, ta.W, ta.C, name);
    tens[g].act = ACT_GELU;
    if (train) tens[g].mask_off = act_alloc((size_t)ta.elems() * 2);
    {
      const __nv_bfloat16* pa = P(ta.off);
      __nv_bfloat16 *pd = train ? P(tens[g].mask_off) : nullptr, *pg = P(tens[g].off);
      const int h16 = f16;
      const float *w = wmaster + s.w_off, *b = wmaster + s.b_off;
      float* pool = pool_off >= 0 ? PF(pool_off) : nullptr;
      const int N = ta.N, Hh = ta.H, Ww = ta.W;
      emit([pa, w, b, pd, pg, pool, N, Hh, Ww, h16](cudaStream_t st) { return launch_dw_fwd(pa, w, b, pd, pg, pool, N, Hh, Ww, st, h16); }, LC_OTHER, 0.0, "dw_fwd");
    }
    if (train) {
      Engine* self = this;
      tape.push_back([self, a, g, s]() {
        const __nv_bfloat16* gz = nullptr;
        if (self->finalize(g, &gz)) return 1;
        if (!gz) return 0;
        const Ten ta = self->tens[a];
        self->ensure_gbuf(a);
        self->tens[a].gwritten = true;
        const __nv_bfloat16* pa = self->P(ta.off);
        __nv_bfloat16* ga = self->P(self->tens[a].goff);
        const float* w = self->wmaster + s.w_off;
        float *gw = self->gflat + s.w_off, *gb = self->gflat + s.b_off;
        const int N = ta.N, Hh = ta.H, Ww = ta.W;
        self->emit([gz, pa, w, ga, gw, gb, N, Hh, Ww](cudaStream_t st) { return launch_dw_bwd(gz, pa, w, ga, gw, gb, N, Hh, Ww, st); }, LC_OTHER, 0.0, "dw_bwd");
        self->release_grad(g);
        return 0;
      });
    }
    return g;
  }

  // Per-direction arrays of the folded gate (EGACA, fusion_modules.py:251-259,312-317): slot = step index of the sweep.
  struct GateCtx {
    long s_off = -1;    // fp32 [slots][N][64]      the gate s
    long wf_off = -1;   // 16-bit [slots][N][64][128]  conv3's forward weights scaled by s on their K side
    long wd_off = -1;   // 16-bit [slots][N][128][64]  the same, data-gradient layout (training plans)
    long m_off = -1;    // fp32 [slots][N][128][64]  per-sample weight gradients of the scaled conv (training plans)
    int slots = 0;
    long cap = 0, used = 0;  // samples (= step x image pairs) the arrays hold / handed out so far
  };
  GateCtx gate_ctx[2];

  void gate_ctx_alloc(int dir, int N) {
    GateCtx& g = gate_ctx[dir];
    g = GateCtx();
    g.slots = train ? T : 1;
    g.cap = (long)g.slots * N;
    g.s_off = act_alloc((size_t)g.slots * N * 64 * 4);
    g.wf_off = act_alloc((size_t)g.slots * N * 8192 * 2);
    if (train) {
      g.wd_off = act_alloc((size_t)g.slots * N * 8192 * 2);
      g.m_off = act_alloc((size_t)g.slots * N * 8192 * 4);
    }
  }

  // Image n of a conv-shaped launch as its own ConvDesc / OutGroup set (tiny grids, where one 128-pixel tile spans
  // several images and per-image weights / outputs cannot be addressed inside one launch).
  static void slice_image(ConvDesc* d, OutGroup* g, int ng, int n, int OH, int OW) {
    for (int k = 0; k < d->nsrc; ++k) d->src[k].ptr += (size_t)n * d->H * d->W * d->src[k].pitch;
    d->N = 1;
    for (int i = 0; i < ng; ++i) {
      EpiDesc& e = g[i].epi;
      const size_t off = (size_t)n * OH * OW * e.C;
      if (e.out) e.out += off;
      if (e.out2) e.out2 += off;
      if (e.out_pre) e.out_pre += off;
      if (e.out_f32) e.out_f32 += off;
      if (e.post) e.post += off;
      if (e.pre) e.pre += off;
      if (e.pre2) e.pre2 += off;
      if (e.sv) e.sv += off;
      if (e.bias) e.bias += (size_t)n * e.bias_nstride;
    }
  }

  // 1x1 conv whose weights differ per image (`w` = [N][rows_per_image][K] 16-bit): one tap-GEMM launch when every
  // 128-pixel tile lies inside one image, else one launch per image.
  int emit_per_image_conv(ConvDesc d, const OutGroup* groups, int ng, const __nv_bfloat16* w, int rows_per_image, int cls,
                          double flops, const char* tag) {
    if (dry) return 0;
    d.w = w;
    d.w_row0 = 0;
    d.wrows_per_tap = rows_per_image;
    if (one_image_per_tile(d.N, d.H, d.W)) {
      d.w_rows = (long)d.N * rows_per_image;
      d.w_img_rows = rows_per_image;
      TapGemmLaunch l;
      if (build_conv(d, groups, ng, &l)) return 1;
      emit([l](cudaStream_t st) mutable { return run_conv(l, st); }, cls, flops, tag);
      return 0;
    }
    const int N = d.N;
    for (int n = 0; n < N; ++n) {
      ConvDesc dn = d;
      OutGroup gn[2];
      for (int i = 0; i < ng; ++i) gn[i] = groups[i];
      slice_image(&dn, gn, ng, n, d.H, d.W);
      dn.w = w + (size_t)n * rows_per_image * d.w_cols;
      dn.w_rows = rows_per_image;
      dn.w_img_rows = 0;
      TapGemmLaunch l;
      if (build_conv(dn, gn, ng, &l)) return 1;
      emit([l](cudaStream_t st) mutable { return run_conv(l, st); }, cls, flops / N, tag);
    }
    return 0;
  }

  // One EGACA evaluation for direction `dir`: event feature xe (per step), image feature branch g_i (hoisted).
  // The squeeze-excite gate is folded into conv3 (the conv that consumes the gated features): the SE kernel scales this
  // sample's conv3 weights by s on their K side, conv3 then reads g_i / g_e directly, and in the backward pass the gate's
  // gradient comes out of a per-sample weight-gradient GEMM -- the gated tensor, its gradient and the two reduction passes
  // over them do not exist (r1: gate_fwd + gate_bwd_reduce + gate_bwd_apply, ~0.3 GB of traffic per call).
  // Chunk views of all-T tensors for the intermediates that feed 1x1-conv sites (so that those sites' weight gradients are one
  // launch over all chunks, see decide_batching); -1: per-call allocations.
  struct EgacaPresets {
    int n_e, a_e, y, n_y, g4, u;
    EgacaPresets() : n_e(-1), a_e(-1), y(-1), n_y(-1), g4(-1), u(-1) {}
  };
  int egaca_step(int dir, int xe, int xi, int g_i, const std::string& nm) { return egaca_step(dir, xe, xi, g_i, nm, EgacaPresets()); }
  int egaca_step(int dir, int xe, int xi, int g_i, const std::string& nm, EgacaPresets ps) {
    const std::string a = std::string(dir ? "encoders_forward" : "encoders_backward") + ".1.atten_fuse";
    const Ten te = tens[xe];
    const int N = te.N;
    const long hw = (long)te.H * te.W;
    const int n_e = ln(xe, "", ps.n_e);
    ConvOp c1;
    c1.kind = CK_1X1;
    c1.site = site(a + ".conv1_e");
    c1.in[0] = n_e;
    c1.out = ps.a_e;
    const int a_e = conv(c1);
    if (a_e < 0) return -1;
    GateCtx& gc = gate_ctx[dir];
    // training plans keep every (step, image) pair's gate arrays: this call's N images (one step, or a chunk of steps) take
    // the next N sample slots; forward-only plans reuse slot 0
    REFID_REQUIRE(gc.slots > 0 && (!train || gc.used + N <= gc.cap), "EGACA gate arrays not allocated / exhausted");
    const long samp = train ? gc.used : 0;
    if (train) gc.used += N;
    // small fp32 state: saved mean / hidden, pooled gradient; per-block partial sums of the global pool
    const int parts = dw_pool_parts(te.H, te.W);
    const long small = act_alloc((size_t)N * (64 * 2 + 32) * 4);
    const long mean_off = small, gpool_off = small + N * 64 * 4, z_off = small + N * 128 * 4;
    const long s_off = gc.s_off + samp * 64 * 4;
    const long wf_off = gc.wf_off + samp * 8192 * 2;
    const long wd_off = train ? gc.wd_off + samp * 8192 * 2 : -1;
    const long m_off = train ? gc.m_off + samp * 8192 * 4 : -1;
    const long pool_off = act_alloc((size_t)N * parts * 64 * 4);
    const int g_e = dw(a_e, site(a + ".conv2_e"), pool_off, nm + ".g_e");
    const int s1 = site(a + ".se_1.1"), s2 = site(a + ".se_1.3"), s3 = site(a + ".conv3");
    if (s1 < 0 || s2 < 0 || s3 < 0) return -1;
    SeParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.w1 = wmaster + sites[s1].w_off;
    sp.b1 = wmaster + sites[s1].b_off;
    sp.w2 = wmaster + sites[s2].w_off;
    sp.b2 = wmaster + sites[s2].b_off;
    sp.gw1 = gflat ? gflat + sites[s1].w_off : nullptr;
    sp.gb1 = gflat ? gflat + sites[s1].b_off : nullptr;
    sp.gw2 = gflat ? gflat + sites[s2].w_off : nullptr;
    sp.gb2 = gflat ? gflat + sites[s2].b_off : nullptr;
    sp.w3 = wmaster + sites[s3].w_off;  // [k = 128][co = 64] fp32 (beta folded in by the host)
    sp.wf = P(wf_off);
    sp.wd = train ? P(wd_off) : nullptr;
    sp.f16 = f16;
    const float inv_hw = 1.f / (float)hw;
    {
      float *pool = PF(pool_off), *sg = PF(s_off), *mean = PF(mean_off), *z = PF(z_off);
      cur_label = a;
      emit([pool, parts, inv_hw, sp, sg, mean, z, N](cudaStream_t st) { return launch_se_fwd(pool, parts, inv_hw, sp, sg, mean, z, N, st); }, LC_OTHER, 0.0, ":se_fwd");
    }
    // y = conv3(cat(g_i, g_e)) with this step's per-sample weights, + xe + xi (fusion_modules.py:317-321)
    const int y = ps.y >= 0 ? ps.y : new_tensor(N, te.H, te.W, 64, nm + ".y");
    {
      ConvDesc d;
      memset(&d, 0, sizeof(d));
      d.kind = CK_1X1;
      d.nsrc = 2;
      d.src[0] = {P(tens[g_i].off), 64, tens[g_i].pitch};
      d.src[1] = {P(tens[g_e].off), 64, tens[g_e].pitch};
      d.N = N;
      d.H = te.H;
      d.W = te.W;
      d.w_cols = 128;
      d.f16 = f16;
      OutGroup g;
      memset(&g, 0, sizeof(g));
      g.channels = 64;
      g.epi.out = P(tens[y].off);
      g.epi.C = 64;
      g.epi.bias = wmaster + sites[s3].b_off;
      g.epi.pre = P(tens[xe].off);
      g.epi.pre2 = P(tens[xi].off);
      cur_label = sites[s3].key;
      if (emit_per_image_conv(d, &g, 1, P(wf_off), 64, LC_CONV_FWD, 2.0 * N * hw * 128 * 64, ":fwd")) return -1;
    }
    if (train) {
      Engine* self = this;
      tape.push_back([self, y, xe, xi, g_i, g_e, sp, inv_hw, N, hw, s3, s_off, mean_off, z_off, gpool_off, wd_off, m_off]() mutable {
        const __nv_bfloat16* gz = nullptr;
        if (self->finalize(y, &gz)) return 1;
        if (!gz) return 0;
        const Site& cs3 = self->sites[s3];
        const Ten ty = self->tens[y];
        self->cur_label = cs3.key;
        self->emit_colsum(gz, (long)N * hw, 64, self->gflat + cs3.b_off);
        float *sg = self->PF(s_off), *mean = self->PF(mean_off), *z = self->PF(z_off), *gpool = self->PF(gpool_off),
              *M = self->PF(m_off);
        // (1) per-sample weight gradient of the scaled conv: M[n][k][co] = sum_pix cat(g_i, g_e)[k] * gz[co]
        self->emit([M, N](cudaStream_t st) {
          REFID_CUDA_CHECK(cudaMemsetAsync(M, 0, (size_t)N * 8192 * 4, st));
          return 0;
        });
        if (!self->dry) {
          ConvDesc d;
          memset(&d, 0, sizeof(d));
          d.kind = CK_1X1;
          d.nsrc = 2;
          d.src[0] = {self->P(self->tens[g_i].off), 64, self->tens[g_i].pitch};
          d.src[1] = {self->P(self->tens[g_e].off), 64, self->tens[g_e].pitch};
          d.N = N;
          d.H = ty.H;
          d.W = ty.W;
          ActSrc q = {gz, 64, ty.pitch};
          const double fl = 2.0 * N * hw * 128 * 64;
          if (one_image_per_tile(N, ty.H, ty.W)) {
            d.out_img_stride = 8192;
            WgradLaunch wl;
            if (build_wgrad(d, q, M, &wl)) return 1;
            self->emit([wl](cudaStream_t st) mutable { return run_wgrad(wl, st); }, LC_WGRAD, fl, ":wgrad");
          } else {
            for (int n = 0; n < N; ++n) {
              ConvDesc dn = d;
              slice_image(&dn, nullptr, 0, n, ty.H, ty.W);
              ActSrc qn = q;
              qn.ptr += (size_t)n * hw * ty.pitch;
              WgradLaunch wl;
              if (build_wgrad(dn, qn, M + (size_t)n * 8192, &wl)) return 1;
              self->emit([wl](cudaStream_t st) mutable { return run_wgrad(wl, st); }, LC_WGRAD, fl / N, ":wgrad");
            }
          }
        }
        // (2) gate gradient from M, squeeze-excite backward -> pooled gradient per sample and channel
        SeParams sb = sp;
        sb.mwg = M;
        self->emit([sg, mean, z, inv_hw, sb, gpool, N](cudaStream_t st) {
          return launch_se_bwd(nullptr, sg, mean, z, inv_hw, sb, gpool, N, st);
        }, LC_OTHER, 0.0, ":se_bwd");
        // (3) data gradient with the scaled weights: d g_i accumulates in fp32 over the steps (one tensor for all T), d z_e =
        //     (W^T gz . s + pooled gradient) * gelu'(z_e) goes straight into the depthwise conv's gradient buffer
        self->ensure_gbuf(g_e);
        self->tens[g_e].gwritten = true;
        Target tgi;  // d g_i: fp32 accumulation when g_i is one tensor for all T steps, a plain 16-bit target for a per-chunk copy
        if (self->tens[g_i].f32acc) self->tens[g_i].gfwritten = true;
        else if (self->target(g_i, &tgi)) return 1;
        if (!self->dry) {
          ConvDesc d;
          memset(&d, 0, sizeof(d));
          d.kind = CK_1X1;
          d.nsrc = 1;
          d.src[0] = {gz, 64, ty.pitch};
          d.N = N;
          d.H = ty.H;
          d.W = ty.W;
          d.w_cols = 64;
          OutGroup g[2];
          memset(g, 0, sizeof(g));
          g[0].channels = 64;
          if (self->tens[g_i].f32acc) {
            g[0].epi.out_f32 = self->PF(self->tens[g_i].gfoff);
          } else {
            g[0].epi.out = tgi.dst;
            g[0].epi.pre = tgi.pre;
            g[0].epi.pre2 = tgi.pre2;
            g[0].epi.out_f32 = tgi.dstf;
          }
          g[0].epi.C = 64;
          g[1].channels = 64;
          g[1].epi.out = self->P(self->tens[g_e].goff);
          g[1].epi.C = 64;
          g[1].epi.bias = gpool;
          g[1].epi.bias_nstride = 64;
          g[1].epi.sv = self->P(self->tens[g_e].mask_off);
          g[1].epi.act = ACT_MULT;
          if (self->emit_per_image_conv(d, g, 2, self->P(wd_off), 128, LC_CONV_DGRAD, 2.0 * N * hw * 128 * 64, ":dgrad")) return 1;
        }
        self->release_consumed();
        self->cur_label = "";
        self->add_pending(xe, self->tens[y].gidx, self->tens[y].goff);
        self->add_pending(xi, self->tens[y].gidx, self->tens[y].goff);
        self->release_grad(y);
        return 0;
      });
    }
    const int n_y = ln(y, "", ps.n_y);
    ConvOp c4;
    c4.kind = CK_1X1;
    c4.site = site(a + ".conv4");
    c4.in[0] = n_y;
    c4.act = ACT_GELU;
    c4.out = ps.g4;
    const int g4 = conv(c4);
    if (g4 < 0) return -1;
    ConvOp c5;
    c5.kind = CK_1X1;
    c5.site = site(a + ".conv5s");
    c5.in[0] = y;
    c5.in[1] = g4;
    c5.nin = 2;
    c5.out = ps.u;
    return conv(c5, nm + ".u");
  }

  // conv3's weight gradient of one direction: sum over the sweep's (step, sample) pairs of M * s (after the sweep's BPTT)
  int emit_gate_wgrad(int dir) {
    const GateCtx& gc = gate_ctx[dir];
    if (!train || gc.used == 0) return 0;
    const int s3 = site(std::string(dir ? "encoders_forward" : "encoders_backward") + ".1.atten_fuse.conv3");
    if (s3 < 0) return 1;
    const float *M = PF(gc.m_off), *sg = PF(gc.s_off);
    float* gw3 = gflat + sites[s3].w_off;
    const int count = (int)gc.used;
    cur_label = sites[s3].key;
    emit([M, sg, count, gw3](cudaStream_t st) { return launch_gate_wgrad(M, sg, count, gw3, st); }, LC_OTHER, 0.0, ":wgrad_gate");
    cur_label = "";
    return 0;
  }

  // ------------------------------------------------------------------------------------------
  // the network
  // ------------------------------------------------------------------------------------------
  int zero_tensor(int N, int Hh, int Ww, int C, std::vector<std::pair<size_t, size_t>>* zero_list) {
    const int id = new_tensor(N, Hh, Ww, C);
    tens[id].need_grad = false;
    zero_list->push_back({(size_t)tens[id].off, (size_t)tens[id].elems() * 2});
    return id;
  }

  int trunk(const std::string& p, int u, int hprev, const std::string& nm, int post, int out2_preset, int* out2,
            int h_preset = -1, bool post_no_grad = false, bool out2_grad_preadded = false) {
    ConvOp a;
    a.site = site(p + ".main.0");
    a.in[0] = u;
    a.in[1] = hprev;
    a.nin = 2;
    a.act = ACT_LRELU;
    a.slope = 0.1f;
    const int v = conv(a, nm + ".v");
    if (v < 0) return -1;
    ConvOp b;
    b.site = site(p + ".main.2.0.conv1");
    b.in[0] = v;
    b.act = ACT_LRELU;
    b.slope = 0.f;
    const int r = conv(b);
    if (r < 0) return -1;
    ConvOp c;
    c.site = site(p + ".main.2.0.conv2");
    c.in[0] = r;
    c.res = v;
    c.post = post;
    c.post_no_grad = post_no_grad;
    c.out2_grad_preadded = out2_grad_preadded;
    c.out2 = out2_preset;
    c.out = h_preset;
    const int h = conv(c, nm + ".h");
    if (h < 0) return -1;
    if (out2) *out2 = post >= 0 ? (out2_preset >= 0 ? out2_preset : (int)tens.size() - 1) : -1;
    return h;
  }

  // ------------------------------------------------------------------------------------------
  // time-chunked, level-major schedule (training plans)
  // ------------------------------------------------------------------------------------------
  // Only the three convs of a recurrent trunk depend on the previous time step.  Everything else of an encoder level --
  // EGACA, the in-conv, fuse_two_dir, the stride-2 `down` conv -- depends on the level below at the SAME step, so on a
  // training plan (which keeps every step's activations anyway) the sweeps run level by level and those ops run once per
  // CHUNK of k steps on k*B images instead of once per step: ~8x fewer launches of kernels that are too short (10-45 us at
  // 64-128 channels and 128^2 .. 64^2 pixels) to amortise their prologue, pipeline fill and tail.  Per-step tensors a chunk
  // op produces are views of ONE all-T tensor per role, so the per-step consumers (trunks, decoders) and the T-batched
  // weight gradients of their sites see the same contiguous layout as before.  Image-branch features, which enter every
  // step, are replicated k times per chunk (`rep`); the backward of that copy sums the k gradient slices.
  int fuse_all[3] = {-1, -1, -1};  // all-T fuse_two_dir outputs per level (forward sweep)
  // steps per chunk (refid_set_option "tchunk" / REFID_TCHUNK; 0: step-major schedule as on forward-only plans).  The default
  // asks for as many as the 192-image limit of EGACA's per-sample tables allows: all T = 23 at B = 8 (measured at B = 1:
  // 749 -> 783 frames/s against chunks of 8; HighREV B = 2: 348 -> 357; B = 8: 126.7 -> 125.0 ms)
  int tchunk = 64;

  // All-T tensor of one per-step role; act / slope describe the producing conv's activation (views inherit the masks).
  int alloc_all(int N, int Hh, int Ww, int C, int act, float slope, const std::string& name = "") {
    const int id = new_tensor(N, Hh, Ww, C, name);
    if (id < 0) return -1;
    tens[id].act = act;
    tens[id].slope = slope;
    if (act == ACT_GELU && train) tens[id].mask_off = act_alloc((size_t)tens[id].elems() * 2);  // saved gelu'(z)
    if (act == ACT_LRELU) {
      tens[id].mask_off = tens[id].off;
      if (train && C % 32 == 0) {  // derivative mask as sign bits (see conv())
        const int bt = new_tensor(N, Hh, Ww, C / 16);
        tens[bt].need_grad = false;
        tens[id].bits_off = tens[bt].off;
        tens[id].bits_pitch = C / 32;
      }
    }
    return id;
  }

  // k copies of x along the image axis.  Backward: sum of the k gradient slices -> one pending addend of x.
  int rep(int x, int k) {
    const Ten tx = tens[x];
    if (!tx.contiguous()) {
      set_error("rep: strided source");
      return -1;
    }
    const int y = new_tensor(k * tx.N, tx.H, tx.W, tx.C);
    if (y < 0) return -1;
    const size_t bytes = (size_t)tx.elems() * 2;
    {
      const __nv_bfloat16* src = P(tx.off);
      __nv_bfloat16* dst = P(tens[y].off);
      const long elems = tx.elems();
      cur_label = "";
      emit([src, dst, elems, k](cudaStream_t st) { return launch_repeat(src, elems, k, dst, st); }, LC_OTHER, 0.0, "rep");
    }
    if (train && tx.need_grad) {
      Engine* self = this;
      tape.push_back([self, x, y, k, bytes]() {
        const __nv_bfloat16* gy = nullptr;
        if (self->finalize(y, &gy)) return 1;
        if (!gy) return 0;
        const int tmp = self->galloc(bytes);
        __nv_bfloat16* sum = self->P((long)self->gbufs[tmp].off);
        const long elems = (long)(bytes / 2);
        self->cur_label = "";
        self->emit([gy, elems, k, sum](cudaStream_t st) { return launch_sum_series(gy, elems, k, sum, st); }, LC_OTHER, 0.0, "rep:sum");
        self->release_grad(y);
        if (self->add_pending(x, tmp, (long)self->gbufs[tmp].off)) return 1;
        self->gunref(tmp);
        return 0;
      });
    }
    return y;
  }

  // Tape hook placed right AFTER a chunk tensor's producer (so it runs right BEFORE the producer's backward): the per-step
  // views of the chunk only ever receive passthrough addends (skip sums); fold them into the chunk's gradient buffer.
  void push_flush_hook(int chunk_view, std::vector<int> step_views) {
    if (!train) return;
    Engine* self = this;
    tape.push_back([self, chunk_view, step_views]() {
      // common case: every step view holds exactly one addend and the addends are consecutive in memory (the decoder trunk's
      // skip-sum gradients, slices of one all-T gradient buffer): ONE masked-accumulate launch over the chunk
      {
        bool one = !step_views.empty();
        const long vb = (long)self->tens[step_views[0]].elems() * 2;
        for (size_t i = 0; i < step_views.size() && one; ++i) {
          const Ten& tv = self->tens[step_views[i]];
          if (tv.pending.size() != 1 || tv.gwritten || tv.pending[0].off != self->tens[step_views[0]].pending[0].off + (long)i * vb) one = false;
        }
        if (one && self->tens[chunk_view].contiguous() && self->tens[chunk_view].pending.empty()) {
          self->ensure_gbuf(chunk_view);
          AddMaskArgs a = {};
          a.a = self->P(self->tens[step_views[0]].pending[0].off);
          if (self->tens[chunk_view].act != ACT_NONE) {
            a.sv = self->P(self->tens[chunk_view].mask_off);
            a.act = self->tens[chunk_view].act == ACT_GELU ? ACT_MULT : self->tens[chunk_view].act;
            a.slope = self->tens[chunk_view].slope;
          }
          a.dst = self->P(self->tens[chunk_view].goff);
          a.dst_acc = self->tens[chunk_view].gwritten ? 1 : 0;
          a.n = (long)step_views.size() * (vb / 2);
          self->tens[chunk_view].gwritten = true;
          const std::string tag = "addmask:flush_chunk:" + std::to_string(self->tens[chunk_view].H) + "x" + std::to_string(self->tens[chunk_view].C);
          self->emit([a](cudaStream_t s) { return launch_addmask(a, s); }, LC_OTHER, 0.0, tag.c_str());
          for (int v : step_views) {
            self->gunref(self->tens[v].pending[0].g);
            self->tens[v].pending.clear();
            self->ensure_gbuf(v);
            self->tens[v].gwritten = true;
          }
          return 0;
        }
      }
      for (int v : step_views) {
        if (self->tens[v].pending.empty()) continue;
        if (self->tens[chunk_view].gwritten && !self->tens[v].gwritten) {  // a whole-chunk consumer wrote this region already
          self->ensure_gbuf(v);
          self->tens[v].gwritten = true;
        }
        if (self->flush_pending(v, false)) return 1;
      }
      if (!self->tens[chunk_view].gwritten) {
        // the step views carried the chunk's only gradients so far: a later whole-chunk write (the producer's own pending
        // addends) must accumulate -- valid only when every step region has been written
        int nw = 0;
        for (int v : step_views) nw += self->tens[v].gwritten ? 1 : 0;
        REFID_REQUIRE(nw == 0 || nw == (int)step_views.size(), "chunk gradient written for %d of %d steps", nw, (int)step_views.size());
        if (nw) {
          self->ensure_gbuf(chunk_view);
          self->tens[chunk_view].gwritten = true;
        }
      }
      return 0;
    });
  }
  // Tape hook placed right BEFORE the chunk consumer of a recurrent state series (so it runs right AFTER that consumer's
  // backward, which wrote dL/dh for the chunk's slots of the persistent series gradient array): the per-step state tensors
  // share that memory, later contributions (the next step's trunk) must accumulate.
  void push_mark_hook(int chunk_tensor, std::vector<int> step_ids) {
    if (!train) return;
    Engine* self = this;
    tape.push_back([self, chunk_tensor, step_ids]() {
      if (!self->tens[chunk_tensor].gwritten) return 0;
      for (int id : step_ids) {
        self->ensure_gbuf(id);
        self->tens[id].gwritten = true;
      }
      return 0;
    });
  }

  // One view id per (all-T tensor, chunk): producer, whole-chunk consumers and the hooks must agree on the id, because the
  // gradient bookkeeping (written / pending) is kept per tensor id.
  std::map<std::pair<int, int>, int> cview_cache;
  int cview(int all_id, int n0, int nk) {
    auto key = std::make_pair(all_id, n0);
    auto it = cview_cache.find(key);
    if (it != cview_cache.end()) return it->second;
    const int v = view(all_id, n0, nk, 0, tens[all_id].C);
    cview_cache[key] = v;
    return v;
  }

  struct SweepOut {
    int h_last[3];                 // final recurrent state per level
    int f_all[3];                  // forward sweep: all-T fuse_two_dir outputs per level
    bool fuse_deferred[3];         // ... whose gradient w.r.t. the final backward state is computed once, from their sum
    int d_all[3], x3_all;          // forward sweep: all-T `down` outputs per level, level-2 output + image feature
    std::vector<std::pair<int, int>> chunks;  // (t0, k) in sweep order
    std::vector<int> dn[3], cx3;   // forward sweep: per-step views of the `down` outputs and of the level-2 output + image feature
  };

  // One direction's encoder sweep, level by level, k steps per chunk.  dir 0: t = T-1 .. 0 (XXNet_final_attenfusion_arch.py
  // :172-181); dir 1: t = 0 .. T-1 with fuse_two_dir against the final backward states hb_final (:185-200).
  int sweep_chunked(int dir, int u0_all, const int xb[3], int g_i, const int h_series[3], const int h_pad[3], const int* hb_final,
                    SweepOut* out) {
    const int b = cfg.base_num_channels;
    const std::string dname = dir ? "encoders_forward" : "encoders_backward";
    const std::string dtag = dir ? "f" : "b";
    int kmax = tchunk;
    // per-sample tables of the folded gate live in shared memory (halo-conv epilogue): 512 bytes per image of the chunk
    const int max_imgs = (int)(kHaloMaxBiasTable / 512);
    if (kmax * B > max_imgs) kmax = max_imgs / B;
    if (kmax < 1) kmax = 1;
    struct Chunk { int t0, k; };
    std::vector<Chunk> chunks;  // in sweep order; images of a chunk are ordered by ascending t
    if (dir) for (int t0 = 0; t0 < T; t0 += kmax) chunks.push_back({t0, T - t0 < kmax ? T - t0 : kmax});
    else for (int t1 = T; t1 > 0; t1 -= kmax) chunks.push_back({t1 - kmax > 0 ? t1 - kmax : 0, t1 < kmax ? t1 : kmax});
    out->chunks.clear();
    for (const Chunk& ck : chunks) out->chunks.push_back({ck.t0, ck.k});
    int x_all = -1;  // input of the current level for all T (output of the level below), level >= 1
    for (int l = 0; l < 3; ++l) {
      const int Hl = H >> l, Wl = W >> l, Cl = (2 * b) << l;  // trunk resolution / channels
      const std::string p = dname + "." + std::to_string(l);
      // all-T tensors of this level's chunk-produced roles
      int u_all = -1, f_all = -1, d_all = -1, x_next = -1;
      if (l == 0) u_all = u0_all;
      else if (l == 1) u_all = alloc_all(T * B, Hl, Wl, Cl, ACT_NONE, 0.f, dtag + ".u1_all");
      else u_all = alloc_all(T * B, Hl, Wl, Cl, ACT_LRELU, 0.04f, dtag + ".u2_all");
      if (u_all < 0) return 1;
      const bool has_down = dir == 1 || l < 2;  // encoders_backward.2.down is dead
      const bool has_post = has_down && ((dir == 1 && l >= 1) || (dir == 0 && l == 1));
      if (dir == 1) f_all = alloc_all(T * B, Hl, Wl, Cl, ACT_LRELU, 0.2f, dtag + ".fuse" + std::to_string(l));
      if (has_down) {
        d_all = alloc_all(T * B, Hl / 2, Wl / 2, Cl, ACT_NONE, 0.f, dtag + ".dn" + std::to_string(l));
        x_next = has_post ? alloc_all(T * B, Hl / 2, Wl / 2, Cl, ACT_NONE, 0.f, dtag + ".x" + std::to_string(l + 1)) : d_all;
        if (d_all < 0 || x_next < 0) return 1;
      }
      if (dir == 1) {
        out->dn[l].assign(T, -1);
        if (l == 2) out->cx3.assign(T, -1);
      }
      // fuse_two_dir reads the final backward state modulo B (no copies) unless a weight-gradient tile of this tiny grid would
      // span images that do not repeat with period B; then the state is replicated per chunk like the image features
      bool hb_modulo = dir == 1;
      for (const Chunk& ck : chunks) {
        int tw, th, tn;
        pick_tile(ck.k * B, Hl, Wl, &tw, &th, &tn);
        if (B % tn) hb_modulo = false;
      }
      if (dir == 1) out->fuse_deferred[l] = hb_modulo;
      int eg_all[5] = {-1, -1, -1, -1, -1};  // level 1: EGACA intermediates that feed 1x1-conv sites, for all T
      if (l == 1) {
        const int Hx = tens[x_all].H, Wx = tens[x_all].W, Cx = tens[x_all].C;
        eg_all[0] = alloc_all(T * B, Hx, Wx, Cx, ACT_NONE, 0.f);      // LayerNorm(x_e)
        eg_all[1] = alloc_all(T * B, Hx, Wx, Cx, ACT_NONE, 0.f);      // conv1_e
        eg_all[2] = alloc_all(T * B, Hx, Wx, Cx, ACT_NONE, 0.f);      // y = conv3(...) + x_e + x_i
        eg_all[3] = alloc_all(T * B, Hx, Wx, Cx, ACT_NONE, 0.f);      // LayerNorm(y)
        eg_all[4] = alloc_all(T * B, Hx, Wx, 2 * Cx, ACT_GELU, 0.f);  // GELU(conv4)
        for (int i = 0; i < 5; ++i)
          if (eg_all[i] < 0) return 1;
      }
      int hprev = h_pad[l];
      // chunk ops on a chunk's states: fuse_two_dir (forward sweep), the stride-2 `down` conv (+ image-feature skip sum).
      // They are emitted ONE CHUNK LATE (after the next chunk's trunks): the first trunk of the next chunk reads this chunk's
      // last state, and on the reversed tape its gradient contribution must come AFTER the whole-chunk write of dL/dh by the
      // chunk consumer's data-gradient (which stores, the per-step contributions accumulate).
      auto emit_post = [&](int t0, int k, const std::vector<int>& h_ids) -> int {
        const int n0 = t0 * B, nk = k * B;
        const int slot0 = dir ? t0 + 1 : t0;
        const int hc = series_tensor(h_series[l], slot0, nk, Hl, Wl, Cl);
        series[h_series[l]].need_g = true;
        push_mark_hook(hc, h_ids);
        int dsrc = hc;
        if (dir == 1) {
          // the final backward state enters every step: the conv reads its B images modulo B, its gradient is ONE 1x1
          // data-gradient GEMM on the sum of all T output gradients (deferred_state_grad_all, between the sweeps on the tape)
          ConvOp cf;
          cf.kind = CK_1X1;
          cf.site = site(p + ".fuse_two_dir");
          cf.in[0] = hc;
          cf.nin = 2;
          if (hb_modulo) {
            cf.in[1] = hb_final[l];
            cf.rep_in1 = cf.defer_in1 = true;
          } else {
            cf.in[1] = rep(hb_final[l], k);
            if (cf.in[1] < 0) return 1;
          }
          cf.out = cview(f_all, n0, nk);
          cf.act = ACT_LRELU;
          cf.slope = 0.2f;
          if (conv(cf) < 0) return 1;
          if (tens[cf.out].bits_off < 0) tens[f_all].bits_off = -1;  // (the conv fell back to an engine without sign bits)
          dsrc = cf.out;
        }
        ConvOp cd;
        cd.kind = CK_DOWN4;
        cd.site = site(p + ".down");
        cd.in[0] = dsrc;
        cd.out = cview(d_all, n0, nk);
        if (has_post) {
          cd.post = rep(xb[l], k);
          if (cd.post < 0) return 1;
          cd.out2 = cview(x_next, n0, nk);
        }
        if (conv(cd) < 0) return 1;
        if (dir == 1) {
          // per-step views for the decoders' skip sums (passthrough addends only: folded in by the hook) and, at level 2,
          // the bottleneck's input (written through gradient targets)
          std::vector<int> dviews;
          for (int i = 0; i < k; ++i) {
            const int t = t0 + i;
            out->dn[l][t] = view(d_all, t * B, B, 0, Cl);
            dviews.push_back(out->dn[l][t]);
            if (l == 2) out->cx3[t] = view(x_next, t * B, B, 0, Cl);
          }
          push_flush_hook(cd.out, dviews);
        }
        return 0;
      };
      int pend_t0 = -1, pend_k = 0;
      std::vector<int> pend_ids;
      for (const Chunk& ck : chunks) {
        const int t0 = ck.t0, k = ck.k, n0 = t0 * B, nk = k * B;
        // ---- chunk op: this level's trunk input u for the chunk's k steps
        if (l == 1) {
          const int xi_r = rep(xb[0], k), gi_r = rep(g_i, k);
          if (xi_r < 0 || gi_r < 0) return 1;
          EgacaPresets ps;
          ps.n_e = cview(eg_all[0], n0, nk);
          ps.a_e = cview(eg_all[1], n0, nk);
          ps.y = cview(eg_all[2], n0, nk);
          ps.n_y = cview(eg_all[3], n0, nk);
          ps.g4 = cview(eg_all[4], n0, nk);
          ps.u = cview(u_all, n0, nk);
          if (egaca_step(dir, cview(x_all, n0, nk), xi_r, gi_r, dtag + ".c" + std::to_string(t0), ps) < 0) return 1;
        } else if (l == 2) {
          ConvOp ci;
          ci.site = site(p + ".conv");
          ci.in[0] = cview(x_all, n0, nk);
          ci.out = cview(u_all, n0, nk);
          ci.act = ACT_LRELU;
          ci.slope = 0.04f;
          if (conv(ci) < 0) return 1;
          if (tens[ci.out].bits_off < 0) tens[u_all].bits_off = -1;
        }
        // ---- the recurrent trunks, step by step inside the chunk
        std::vector<int> h_ids;
        for (int i = 0; i < k; ++i) {
          const int t = dir ? t0 + i : t0 + k - 1 - i;
          cur_slot = dir ? t + 1 : t;
          cur_sdir = dtag + std::to_string(l);
          step_seq = 0;
          const int u = l == 0 ? view(u_all, t * B, B, dir ? 2 * b : 0, 2 * b) : view(u_all, t * B, B, 0, Cl);
          const std::string nm = dtag + ".t" + std::to_string(t) + ".l" + std::to_string(l);
          const int hslot = series_tensor(h_series[l], cur_slot, B, Hl, Wl, Cl);
          const int h = trunk(p + ".recurrent_block.forward_trunk", u, hprev, nm, -1, -1, nullptr, hslot);
          cur_slot = -1;
          if (h < 0) return 1;
          hprev = h;
          h_ids.push_back(h);
        }
        if (!has_down) continue;
        if (pend_t0 >= 0 && emit_post(pend_t0, pend_k, pend_ids)) return 1;
        pend_t0 = t0;
        pend_k = k;
        pend_ids = h_ids;
      }
      if (pend_t0 >= 0 && emit_post(pend_t0, pend_k, pend_ids)) return 1;
      out->h_last[l] = hprev;
      out->f_all[l] = f_all;
      out->d_all[l] = d_all;
      if (l == 2) out->x3_all = x_next;
      x_all = x_next;
    }
    return 0;
  }

  // Chunked tail of the forward sweep (training plans): the bottleneck ResidualBlocks and the decoders' transposed convs have
  // no recurrence either -- at 32^2 .. 64^2 pixels one step is a quarter of a wave of tiles -- so they run per chunk; only
  // the decoder trunks (and `pred`, which writes the caller's tensor step by step) run per step, level by level.
  int tail_chunked(const SweepOut& so, int head, int sp_all, int sd[3], const int sd_series[3]) {
    const int b = cfg.base_num_channels;
    const int y_all = alloc_all(T * B, H >> 3, W >> 3, 8 * b, ACT_NONE, 0.f, "f.y_all");
    if (y_all < 0) return 1;
    for (auto& ck : so.chunks) {
      const int n0 = ck.first * B, nk = ck.second * B;
      int xin = cview(so.x3_all, n0, nk);
      for (int i = 0; i < 2; ++i) {
        const std::string p = "resblocks." + std::to_string(i);
        ConvOp c1;
        c1.site = site(p + ".conv1");
        c1.in[0] = xin;
        c1.act = ACT_LRELU;
        c1.slope = 0.f;
        const int r = conv(c1);
        if (r < 0) return 1;
        ConvOp c2;
        c2.site = site(p + ".conv2");
        c2.in[0] = r;
        c2.res = xin;
        c2.act = ACT_LRELU;
        c2.slope = 0.f;
        if (i == 1) {
          c2.post = cview(so.d_all[2], n0, nk);
          c2.out2 = cview(y_all, n0, nk);
        }
        const int o = conv(c2);
        if (o < 0) return 1;
        xin = (i == 1) ? c2.out2 : o;
      }
    }
    int in_all = y_all;
    for (int i = 0; i < 3; ++i) {
      const std::string p = "decoders." + std::to_string(i);
      const int Hd = H >> (2 - i), Wd = W >> (2 - i), Cd = (4 * b) >> i;
      const int up_all = alloc_all(T * B, Hd, Wd, Cd, ACT_NONE, 0.f, "f.up" + std::to_string(i));
      const int o2_all = i < 2 ? alloc_all(T * B, Hd, Wd, Cd, ACT_NONE, 0.f, "f.dec" + std::to_string(i)) : sp_all;
      if (up_all < 0 || o2_all < 0) return 1;
      if (i == 2 && train && tens[head].need_grad) {
        // `head` (the image head's output) is the skip-sum addend of EVERY step's last decoder: its gradient is the sum over T
        // of the gradients of sp_all (one all-T buffer, complete once `pred` and this level have been back-propagated): one
        // reduction instead of T fp32 read-modify-write passes.  Runs after this level's backward (pushed before its ops).
        Engine* self = this;
        const long slot_elems = (long)B * Hd * Wd * Cd;
        tape.push_back([self, head, sp_all, slot_elems]() {
          if (self->tens[sp_all].goff < 0) return 0;
          const int tmp = self->galloc((size_t)slot_elems * 2);
          __nv_bfloat16* sum = self->P((long)self->gbufs[tmp].off);
          const __nv_bfloat16* gy = self->P(self->tens[sp_all].goff);
          const int steps = self->T;
          self->cur_label = "";
          self->emit([gy, slot_elems, steps, sum](cudaStream_t st) { return launch_sum_series(gy, slot_elems, steps, sum, st); },
                     LC_OTHER, 0.0, "head:sum_skip_grads");
          if (self->add_pending(head, tmp, (long)self->gbufs[tmp].off)) return 1;
          self->gunref(tmp);
          return 0;
        });
      }
      std::vector<std::pair<int, int>> pre_pairs;  // (decoder state of step t, step view of the level's all-T skip-sum output)
      static const bool no_preadd = getenv("REFID_NO_PREADD") != nullptr;  // diagnostic
      const bool preadd = train && !no_preadd;
      for (auto& ck : so.chunks) {
        const int t0 = ck.first, k = ck.second, n0 = t0 * B, nk = k * B;
        ConvOp up;
        up.kind = CK_UP2;
        up.site = site(p + ".transposed_conv2d");
        up.in[0] = cview(in_all, n0, nk);
        up.out = cview(up_all, n0, nk);
        if (conv(up) < 0) return 1;
        for (int t = t0; t < t0 + k; ++t) {
          cur_slot = t + 1;
          cur_sdir = "fd" + std::to_string(i);
          step_seq = 0;
          const std::string nm = "f.t" + std::to_string(t) + ".dec" + std::to_string(i);
          const int u = view(up_all, t * B, B, 0, Cd);
          const int post = i < 2 ? so.dn[1 - i][t] : head;
          const int preset = view(o2_all, t * B, B, 0, Cd);
          int o2 = -1;
          const int hs = series_tensor(sd_series[i], t + 1, B, Hd, Wd, Cd);
          const int st = trunk(p + ".forward_trunk", u, sd[i], nm, post, preset, &o2, hs, i == 2, preadd);
          if (preadd && st >= 0) pre_pairs.push_back({st, preset});
          if (st < 0) {
            cur_slot = -1;
            return 1;
          }
          sd[i] = st;
          cur_slot = -1;
        }
        if (i == 2) {  // prediction for the chunk's steps straight into the caller's (B,T,out_chn,H,W) tensor
          ConvOp cp;
          cp.site = site("pred");
          cp.in[0] = cview(sp_all, n0, nk);
          cp.nchw_out = true;
          cp.no_tape = true;
          cp.nchw_toff = (long)t0 * cfg.out_chn * H * W;
          cp.nchw_nstride = (long)T * cfg.out_chn * H * W;
          cp.nchw_B = B;
          cp.nchw_tstride = (long)cfg.out_chn * H * W;
          conv(cp);
        }
      }
      if (preadd) {
        // The gradient of the skip-sum output of every step (a slice of ONE all-T buffer, written by the next level's chunk
        // ops / `pred` before this level is back-propagated) also is a gradient of that step's decoder state.  Registered
        // here -- the first thing this level's backward does --, it rides as the epilogue addend of the NEXT step's main.0
        // data-gradient (which writes dL/ds_t) instead of costing one masked-accumulate pass per step and level afterwards.
        Engine* self = this;
        tape.push_back([self, pre_pairs]() {
          for (auto& pr : pre_pairs) {
            const int v = pr.second, par = self->tens[v].parent;
            if (par < 0 || !(self->tens[par].gwritten || self->tens[par].goff >= 0)) {
              set_error("decoder skip-sum gradient is not available before its level's backward");
              return 1;
            }
            self->ensure_gbuf(v);
            if (self->add_pending(pr.first, self->tens[v].gidx, self->tens[v].goff)) return 1;
          }
          return 0;
        });
      }
      in_all = o2_all;
    }
    return 0;
  }

  // Per-step tail of the forward sweep: bottleneck ResidualBlocks, the three recurrent decoders and `pred` for step t.
  // curx = this step's level-2 encoder output (+ image feature), dn[l] = the encoder skips, sd[] = running decoder states.
  int step_tail(int t, int curx, const int dn[3], int head, int sp_all, int sd[3], const int sd_series[3]) {
    const int b = cfg.base_num_channels;
    // bottleneck: two ResidualBlocks (recurrent_sub_modules.py:468-503)
    int xin = curx;
    for (int i = 0; i < 2; ++i) {
      const std::string p = "resblocks." + std::to_string(i);
      ConvOp c1;
      c1.site = site(p + ".conv1");
      c1.in[0] = xin;
      c1.act = ACT_LRELU;
      c1.slope = 0.f;
      const int r = conv(c1);
      if (r < 0) return 1;
      ConvOp c2;
      c2.site = site(p + ".conv2");
      c2.in[0] = r;
      c2.res = xin;
      c2.act = ACT_LRELU;
      c2.slope = 0.f;
      if (i == 1) c2.post = dn[2];
      const int o = conv(c2, "f.t" + std::to_string(t) + ".res" + std::to_string(i));
      if (o < 0) return 1;
      xin = (i == 1) ? (int)tens.size() - 1 : o;
    }
    // decoders (TransposeRecurrentConvLayer, :370-408); input = previous + encoder skip (skip_sum, :211)
    for (int i = 0; i < 3; ++i) {
      const std::string p = "decoders." + std::to_string(i);
      const std::string nm = "f.t" + std::to_string(t) + ".dec" + std::to_string(i);
      ConvOp up;
      up.kind = CK_UP2;
      up.site = site(p + ".transposed_conv2d");
      up.in[0] = xin;
      const int u = conv(up, nm + ".up");
      if (u < 0) return 1;
      int post, preset = -1, o2 = -1;
      if (i < 2) {
        post = dn[1 - i];
      } else {
        post = head;
        if (train) preset = view(sp_all, t * B, B, 0, b);
      }
      const int s = trunk(p + ".forward_trunk", u, sd[i], nm, post, preset, &o2,
                          series_tensor(sd_series[i], t + 1, B, H >> (2 - i), W >> (2 - i), (4 * b) >> i));
      if (s < 0) return 1;
      sd[i] = s;
      xin = o2;
    }
    // prediction for this step straight into the caller's (B,T,out_chn,H,W) tensor
    ConvOp cp;
    cp.site = site("pred");
    cp.in[0] = xin;
    cp.nchw_out = true;
    cp.no_tape = true;
    cp.nchw_toff = (long)t * cfg.out_chn * H * W;
    cp.nchw_nstride = (long)T * cfg.out_chn * H * W;
    conv(cp);
    return 0;
  }

  int build_network() {
    const int b = cfg.base_num_channels;
    std::vector<std::pair<size_t, size_t>> zero_fwd;
    Engine* self = this;
    // ---- image branch: head_img, three ImageEncoderConvBlocks (recurrent_sub_modules.py:22-49)
    const int ximg = new_tensor(B, H, W, Kp_img);
    tens[ximg].need_grad = false;
    {
      __nv_bfloat16* o = P(tens[ximg].off);
      const int Bc = B, Cin = cfg.img_chn, Hh = H, Ww = W, Kp = Kp_img;
      const int h16 = f16;
      emit([self, o, Bc, Cin, Hh, Ww, Kp, h16](cudaStream_t st) { return launch_unroll5(self->io_x, o, Bc, 1, Cin, Hh, Ww, Kp, st, h16); }, LC_OTHER, 0.0, "unroll5");
    }
    ConvOp hi;
    hi.kind = CK_ROWS5;
    hi.site = site("head_img");
    hi.in[0] = ximg;
    hi.act = ACT_LRELU;
    hi.slope = 0.2f;
    const int head = conv(hi, "head_img");
    if (head < 0) return 1;
    mark_f32acc(head);
    int xb[3];
    int f = head;
    for (int l = 0; l < 3; ++l) {
      const std::string p = "img_encoders." + std::to_string(l);
      ConvOp c1;
      c1.site = site(p + ".conv_1");
      c1.in[0] = f;
      c1.act = ACT_LRELU;
      c1.slope = 0.2f;
      const int t1 = conv(c1);
      ConvOp ci;
      ci.kind = CK_1X1;
      ci.site = site(p + ".identity");
      ci.in[0] = f;
      const int idn = conv(ci);
      if (t1 < 0 || idn < 0) return 1;
      ConvOp c2;
      c2.site = site(p + ".conv_2");
      c2.in[0] = t1;
      c2.act = ACT_LRELU;
      c2.slope = 0.2f;
      c2.post = idn;
      if (conv(c2) < 0) return 1;
      const int sum = (int)tens.size() - 1;
      ConvOp cd;
      cd.kind = CK_DOWN4;
      cd.site = site(p + ".down");
      cd.in[0] = sum;
      xb[l] = conv(cd, "xb" + std::to_string(l));
      if (xb[l] < 0) return 1;
      mark_f32acc(xb[l]);
      f = xb[l];
    }
    // ---- event head + both directions' level-0 in-convs (recurrence independent): ONE launch each over all T slices on a
    //      training plan; step by step on a forward-only plan, where only u0 (both directions' conv outputs) is kept for
    //      all T and the unrolled input / head output live in one-step buffers (workspace O(1) in T apart from u0)
    int u0_all = -1;
    if (train) {
      const int xev = new_tensor(T * B, H, W, Kp_ev);
      tens[xev].need_grad = false;
      {
        __nv_bfloat16* o = P(tens[xev].off);
        const int Bc = B, Tc = T, Cin = cfg.ev_chn, Hh = H, Ww = W, Kp = Kp_ev;
        const int h16 = f16;
        emit([self, o, Bc, Tc, Cin, Hh, Ww, Kp, h16](cudaStream_t st) { return launch_unroll5(self->io_ev, o, Bc, Tc, Cin, Hh, Ww, Kp, st, h16); }, LC_OTHER, 0.0, "unroll5");
      }
      ConvOp he;
      he.kind = CK_ROWS5;
      he.site = site("head");
      he.in[0] = xev;
      he.act = ACT_LRELU;
      he.slope = 0.2f;
      const int e_all = conv(he, "e_all");
      if (e_all < 0) return 1;
      ConvOp cu;
      cu.site = site("enc0_in");
      cu.in[0] = e_all;
      cu.act = ACT_LRELU;
      cu.slope = 0.04f;  // ConvLayer's LeakyReLU(0.2) applied twice (recurrent_sub_modules.py:283-285)
      u0_all = conv(cu, "u0_all");
      if (u0_all < 0) return 1;
    } else {
      u0_all = new_tensor(T * B, H, W, 4 * b, "u0_all");
      const int xev = new_tensor(B, H, W, Kp_ev);
      const int e_t = new_tensor(B, H, W, b);
      tens[e_t].act = ACT_LRELU;
      tens[e_t].slope = 0.2f;
      for (int t = 0; t < T; ++t) {
        __nv_bfloat16* o = P(tens[xev].off);
        const int Bc = B, Tc = T, Cin = cfg.ev_chn, Hh = H, Ww = W, Kp = Kp_ev, h16 = f16;
        emit([self, o, Bc, Tc, Cin, Hh, Ww, Kp, h16, t](cudaStream_t st) { return launch_unroll5(self->io_ev, o, Bc, Tc, Cin, Hh, Ww, Kp, st, h16, t, 1); }, LC_OTHER, 0.0, "unroll5");
        ConvOp he;
        he.kind = CK_ROWS5;
        he.site = site("head");
        he.in[0] = xev;
        he.out = e_t;
        he.act = ACT_LRELU;
        he.slope = 0.2f;
        if (conv(he) < 0) return 1;
        ConvOp cu;
        cu.site = site("enc0_in");
        cu.in[0] = e_t;
        cu.out = view(u0_all, t * B, B, 0, 4 * b);
        cu.act = ACT_LRELU;
        cu.slope = 0.04f;
        if (conv(cu) < 0) return 1;
      }
    }
    // ---- EGACA image branch, once per direction (image feature is constant over time)
    int g_i[2];
    for (int d = 0; d < 2; ++d) {
      const std::string a = std::string(d ? "encoders_forward" : "encoders_backward") + ".1.atten_fuse";
      const int n_i = ln(xb[0]);
      ConvOp c1;
      c1.kind = CK_1X1;
      c1.site = site(a + ".conv1");
      c1.in[0] = n_i;
      const int a_i = conv(c1);
      if (a_i < 0) return 1;
      g_i[d] = dw(a_i, site(a + ".conv2"), -1, d ? "f.g_i" : "b.g_i");
      mark_f32acc(g_i[d]);
    }
    const char* dirs[2] = {"encoders_backward", "encoders_forward"};
    gate_ctx_alloc(0, B);
    gate_ctx_alloc(1, B);
    // ---- backward sweep t = T-1 .. 0 (XXNet_final_attenfusion_arch.py:172-181)
    // Recurrent states live in (T+1)-slot series: slot t = h_t, the spare slot is the zero initial state.
    auto state_series = [&](int Hh, int Ww, int C, int pad_slot, int* pad_tensor) {
      const int sidx = new_series((size_t)B * Hh * Ww * C * 2, 2);
      *pad_tensor = series_tensor(sidx, pad_slot, B, Hh, Ww, C);
      tens[*pad_tensor].need_grad = false;
      zero_fwd.push_back({(size_t)tens[*pad_tensor].off, (size_t)tens[*pad_tensor].elems() * 2});
      return sidx;
    };
    int hb[3], hb_series[3];
    for (int l = 0; l < 3; ++l) hb_series[l] = state_series(H >> l, W >> l, (2 * b) << l, T, &hb[l]);
    const bool chunked = train && tchunk > 0;  // level-major, time-chunked schedule (see sweep_chunked)
    SweepOut so_b, so_f;
    if (chunked) {
      const int pad[3] = {hb[0], hb[1], hb[2]};
      if (sweep_chunked(0, u0_all, xb, g_i[0], hb_series, pad, nullptr, &so_b)) return 1;
      for (int l = 0; l < 3; ++l) hb[l] = so_b.h_last[l];
    }
    for (int t = T - 1; t >= 0 && !chunked; --t) {
      cur_slot = t;
      cur_sdir = "b";
      step_seq = 0;
      int curx = -1;
      for (int l = 0; l < 3; ++l) {
        const std::string p = std::string(dirs[0]) + "." + std::to_string(l);
        const std::string nm = "b.t" + std::to_string(t) + ".l" + std::to_string(l);
        int u;
        if (l == 0) {
          u = view(u0_all, t * B, B, 0, 2 * b);
        } else if (l == 1) {
          u = egaca_step(0, curx, xb[0], g_i[0], nm);
        } else {
          ConvOp ci;
          ci.site = site(p + ".conv");
          ci.in[0] = curx;
          ci.act = ACT_LRELU;
          ci.slope = 0.04f;
          u = conv(ci, nm + ".u");
        }
        if (u < 0) return 1;
        const int h = trunk(p + ".recurrent_block.forward_trunk", u, hb[l], nm, -1, -1, nullptr,
                            series_tensor(hb_series[l], t, B, H >> l, W >> l, (2 * b) << l));
        if (h < 0) return 1;
        hb[l] = h;
        if (l < 2) {
          ConvOp cd;
          cd.kind = CK_DOWN4;
          cd.site = site(p + ".down");
          cd.in[0] = h;
          if (l == 1) cd.post = xb[1];
          const int dn = conv(cd, nm + ".d");
          if (dn < 0) return 1;
          curx = (l == 1) ? (int)tens.size() - 1 : dn;
        }
      }
    }
    cur_slot = -1;
    // Only the FINAL backward state reaches the forward sweep (list aliasing at :181): W_h . h_b is
    // linear in a tensor that is the same at all T steps, so its gradient is W_h^T . (sum_t dL/dZ_t) -- ONE 1x1
    // data-gradient GEMM per level on the summed output gradients of fuse_two_dir, placed on the tape between the two
    // sweeps (i.e. after the whole forward sweep has been back-propagated, before the backward sweep's BPTT starts).
    if (train) {
      const int hb0 = hb[0], hb1 = hb[1], hb2 = hb[2];
      tape.push_back([self, hb0, hb1, hb2, chunked]() {
        const int hbs[3] = {hb0, hb1, hb2};
        for (int l = 0; l < 3; ++l)
          if (chunked ? (self->fuse_all[l] >= 0 && self->deferred_state_grad_all(l, hbs[l], self->fuse_all[l]))
                      : self->deferred_state_grad(l, hbs[l]))
            return 1;
        return 0;
      });
    }
    // ---- forward sweep t = 0 .. T-1 (:185-216)
    int hf[3], sd[3], hf_series[3], sd_series[3];
    for (int l = 0; l < 3; ++l) hf_series[l] = state_series(H >> l, W >> l, (2 * b) << l, 0, &hf[l]);
    for (int i = 0; i < 3; ++i) sd_series[i] = state_series(H >> (2 - i), W >> (2 - i), (4 * b) >> i, 0, &sd[i]);
    const int sp_all = train ? new_tensor(T * B, H, W, b, "sp_all") : -1;  // pred's input for all T steps (batched pred backward)
    if (chunked) {
      const int pad[3] = {hf[0], hf[1], hf[2]};
      if (sweep_chunked(1, u0_all, xb, g_i[1], hf_series, pad, hb, &so_f)) return 1;
      for (int l = 0; l < 3; ++l) fuse_all[l] = so_f.fuse_deferred[l] ? so_f.f_all[l] : -1;
      static const bool step_tail_env = getenv("REFID_STEP_TAIL") != nullptr;  // diagnostic: per-step bottleneck / decoders
      if (!step_tail_env) {
        if (tail_chunked(so_f, head, sp_all, sd, sd_series)) return 1;
      } else {
        for (int t = 0; t < T; ++t) {
          cur_slot = t + 1;
          cur_sdir = "fd";
          step_seq = 0;
          const int dn[3] = {so_f.dn[0][t], so_f.dn[1][t], so_f.dn[2][t]};
          if (step_tail(t, so_f.cx3[t], dn, head, sp_all, sd, sd_series)) return 1;
        }
      }
    }
    for (int t = 0; t < T && !chunked; ++t) {
      cur_slot = t + 1;
      cur_sdir = "f";
      step_seq = 0;
      int curx = -1;
      int dn[3];
      for (int l = 0; l < 3; ++l) {
        const std::string p = std::string(dirs[1]) + "." + std::to_string(l);
        const std::string nm = "f.t" + std::to_string(t) + ".l" + std::to_string(l);
        int u;
        if (l == 0) {
          u = view(u0_all, t * B, B, 2 * b, 2 * b);
        } else if (l == 1) {
          u = egaca_step(1, curx, xb[0], g_i[1], nm);
        } else {
          ConvOp ci;
          ci.site = site(p + ".conv");
          ci.in[0] = curx;
          ci.act = ACT_LRELU;
          ci.slope = 0.04f;
          u = conv(ci, nm + ".u");
        }
        if (u < 0) return 1;
        const int h = trunk(p + ".recurrent_block.forward_trunk", u, hf[l], nm, -1, -1, nullptr,
                            series_tensor(hf_series[l], t + 1, B, H >> l, W >> l, (2 * b) << l));
        if (h < 0) return 1;
        hf[l] = h;
        ConvOp cf;
        cf.kind = CK_1X1;
        cf.site = site(p + ".fuse_two_dir");
        cf.in[0] = h;
        cf.in[1] = hb[l];
        cf.nin = 2;
        cf.defer_in1 = true;
        cf.act = ACT_LRELU;
        cf.slope = 0.2f;
        const int hfu = conv(cf);
        if (hfu < 0) return 1;
        ConvOp cd;
        cd.kind = CK_DOWN4;
        cd.site = site(p + ".down");
        cd.in[0] = hfu;
        if (l >= 1) cd.post = xb[l];
        dn[l] = conv(cd, nm + ".d");
        if (dn[l] < 0) return 1;
        curx = (l >= 1) ? (int)tens.size() - 1 : dn[l];
      }
      if (step_tail(t, curx, dn, head, sp_all, sd, sd_series)) return 1;
    }
    cur_slot = -1;
    // pred backward is batched over all T steps: gout -> bf16 NHWC (padded to 32 channels), wgrad + dgrad once
    if (train) {
      const int ps = site("pred");
      tape.push_back([self, sp_all, ps]() {
        Ten t;
        t.N = self->T * self->B;
        t.H = self->H;
        t.W = self->W;
        t.C = t.pitch = 32;
        t.off = 0;
        self->tens.push_back(t);
        const int pg = (int)self->tens.size() - 1;
        self->ensure_gbuf(pg);
        self->tens[pg].gwritten = true;
        __nv_bfloat16* dst = self->P(self->tens[pg].goff);
        const int Bc = self->B, Tc = self->T, Cv = self->cfg.out_chn, Hh = self->H, Ww = self->W;
        self->emit([self, dst, Bc, Tc, Cv, Hh, Ww](cudaStream_t st) {
          return launch_gout_pack(self->io_gout, dst, Bc, Tc, Cv, Hh, Ww, st);
        }, LC_OTHER, 0.0, "gout_pack");
        ConvOp op;
        op.site = ps;
        op.in[0] = sp_all;
        op.out = pg;
        return self->conv_bwd(op);
      });
    }
    // workspace zeroing for the t = 0 recurrent states goes first in the forward list
    if (!dry) {
      std::vector<Launch> pre;
      for (auto& z : zero_fwd) {
        char* p = ws + z.first;
        const size_t n = z.second;
        pre.push_back([p, n](cudaStream_t st) {
          REFID_CUDA_CHECK(cudaMemsetAsync(p, 0, n, st));
          return 0;
        });
      }
      fwd.insert(fwd.begin(), pre.begin(), pre.end());
      fwd_meta.insert(fwd_meta.begin(), pre.size(), LaunchMeta{LC_OTHER, 0.0, "memset:state0"});
    }
    decide_batching();
    return 0;
  }

  int deferred_state_grad(int l, int hb_id) {
    const int si = site("encoders_forward." + std::to_string(l) + ".fuse_two_dir");
    if (si < 0) return 1;
    const Site& s = sites[si];
    const auto& calls = site_calls[si];
    REFID_REQUIRE(!calls.empty(), "fuse_two_dir has no calls");
    int lo = 1 << 30;
    for (auto& c : calls) lo = tens[c.out].slot < lo ? tens[c.out].slot : lo;
    const Ten o = tens[calls[0].out];
    const Series& sr = series[o.series];
    REFID_REQUIRE(sr.gbase >= 0 && (int)calls.size() == T, "fuse_two_dir gradients are not kept as a series");
    const long slot_elems = (long)(sr.slot_bytes / 2);
    return deferred_state_grad_core(s, hb_id, P(sr.gbase + (long)lo * (long)sr.slot_bytes), slot_elems, o);
  }

  // chunked schedule: the fuse_two_dir outputs of all T steps are ONE tensor (f_all), so are their gradients
  int deferred_state_grad_all(int l, int hb_id, int f_all) {
    const int si = site("encoders_forward." + std::to_string(l) + ".fuse_two_dir");
    if (si < 0) return 1;
    Ten o = tens[f_all];
    REFID_REQUIRE(o.goff >= 0 && o.N == T * B, "fuse_two_dir (level %d): no output gradients were written", l);
    o.N = B;
    return deferred_state_grad_core(sites[si], hb_id, P(o.goff), o.elems(), o, true);
  }

  // dL/dh_b = W_h^T . (sum over the T steps of dL/dZ_t): gbase = T consecutive gradient slices of slot_elems elements
  // as_pending (chunked schedule): the result becomes a pending addend of h_b instead of a write into its gradient buffer --
  // there the buffer is a slot of the state series' gradient array, which the chunk consumer of the backward sweep (`down`'s
  // data-gradient over the chunk) STORES into later on the reversed tape.
  int deferred_state_grad_core(const Site& s, int hb_id, const __nv_bfloat16* gbase, long slot_elems, const Ten& o,
                               bool as_pending = false) {
    cur_label = s.key;
    const int tmp = galloc((size_t)slot_elems * 2);
    __nv_bfloat16* sum = P((long)gbufs[tmp].off);
    const int steps = T;
    emit([gbase, slot_elems, steps, sum](cudaStream_t st) { return launch_sum_series(gbase, slot_elems, steps, sum, st); },
         LC_OTHER, 0.0, ":sum_gz");
    Target t;
    int pend = -1;
    if (as_pending) {
      REFID_REQUIRE(tens[hb_id].act == ACT_NONE && tens[hb_id].contiguous(), "deferred state gradient: unexpected state tensor");
      pend = galloc((size_t)tens[hb_id].elems() * 2);
      t.dst = P((long)gbufs[pend].off);
      t.pitch = tens[hb_id].pitch;
    } else if (target(hb_id, &t)) {
      return 1;
    }
    if (!dry) {
      const int c = o.C;  // fuse: (h | h_b) 2c -> c; the data-gradient pack is [2c rows][c cols], h_b rows start at c
      OutGroup g;
      memset(&g, 0, sizeof(g));
      g.channels = tens[hb_id].C;
      g.epi.out = t.dst;
      g.epi.pre = t.pre;
      g.epi.pre2 = t.pre2;
      g.epi.sv = t.sv;
      g.epi.act = t.act;
      g.epi.slope = t.slope;
      g.epi.out_f32 = t.dstf;
      g.epi.C = t.pitch;
      ConvDesc d;
      memset(&d, 0, sizeof(d));
      d.kind = CK_1X1;
      d.src[0] = {sum, c, c, 0};
      d.nsrc = 1;
      d.N = o.N;
      d.H = o.H;
      d.W = o.W;
      d.w = wpackb + s.dgrad_off;
      d.w_cols = c;
      d.w_rows = 2L * c;
      d.wrows_per_tap = 2 * c;
      d.w_row0 = c;
      TapGemmLaunch lch;
      if (build_conv(d, &g, 1, &lch)) return 1;
      emit([lch](cudaStream_t st) mutable { return run_conv(lch, st); }, LC_CONV_DGRAD,
           2.0 * (double)o.N * o.H * o.W * c * c * T, ":dgrad_hb");
    }
    release_consumed();
    gunref(tmp);
    cur_label = "";
    if (pend >= 0) {
      if (add_pending(hb_id, pend, (long)gbufs[pend].off)) return 1;
      gunref(pend);
    }
    return 0;
  }

  // A site called once per time step with contiguous per-step operands gets ONE weight/bias-gradient launch over all
  // T*B images; its per-step output gradients are then kept in a persistent series array instead of the pool.
  void decide_batching() {
    site_batched.assign(sites.size(), 0);
    site_rep.assign(sites.size(), 0);
    site_calls.resize(sites.size());
    if (!train) return;
    for (auto& sr : series)
      if (sr.need_g && sr.gbase < 0) sr.gbase = act_alloc(sr.slot_bytes * (size_t)(T + 1));
    static const bool disabled = getenv("REFID_NO_BATCHED_WGRAD") != nullptr;
    if (disabled) return;
    for (size_t si = 0; si < sites.size(); ++si) {
      auto& calls = site_calls[si];
      if (calls.size() < 2) continue;
      bool ok = true;
      for (auto& c : calls)
        if (c.out < 0 || c.nchw_out || tens[c.out].series < 0 || tens[c.out].parent >= 0) ok = false;
      if (!ok) continue;
      std::sort(calls.begin(), calls.end(), [&](const ConvOp& a, const ConvOp& b2) { return tens[a.out].slot < tens[b2.out].slot; });
      const Ten o0 = tens[calls[0].out];
      int rep = 0;
      for (int k = 0; k < calls[0].nin; ++k)
        if (calls.size() > 1 && tens[calls[1].in[k]].off == tens[calls[0].in[k]].off) {
          // the same tensor at every step (the final backward state in fuse_two_dir): repeat it along the image axis
          const Ten& a0 = tens[calls[0].in[k]];
          int tw, th, tn;
          pick_tile(a0.N, a0.H, a0.W, &tw, &th, &tn);
          if (calls[0].kind == CK_1X1 && a0.N % tn == 0) rep |= 1 << k;
        }
      for (size_t i = 0; i < calls.size() && ok; ++i) {
        const Ten& o = tens[calls[i].out];
        if (o.series != o0.series || o.slot != o0.slot + (int)i || calls[i].nin != calls[0].nin ||
            calls[i].kind != calls[0].kind || !o.contiguous())
          ok = false;
        for (int k = 0; k < calls[0].nin && ok; ++k) {
          const Ten& a0 = tens[calls[0].in[k]];
          const Ten& a = tens[calls[i].in[k]];
          const long stride = (long)a0.N * a0.H * a0.W * a0.pitch * 2;
          if (a.C != a0.C || a.pitch != a0.pitch || a.N != a0.N || a.H != a0.H || a.W != a0.W ||
              a.off != a0.off + ((rep >> k) & 1 ? 0 : (long)i * stride))
            ok = false;
        }
      }
      if (!ok) continue;
      site_batched[si] = 1;
      site_rep[si] = rep;
      Series& sr = series[o0.series];
      if (sr.gbase < 0) sr.gbase = act_alloc(sr.slot_bytes * (size_t)(T + 1));
    }
    // Chunk ops (time-chunked schedule): the calls of a site write consecutive chunk views of ONE all-T tensor and read
    // consecutive chunks of one memory range, so the site's weight gradient is again one launch over all T*B images; the
    // output gradients of all chunks are contiguous in the all-T tensor's gradient buffer (which is never recycled).
    for (size_t si = 0; si < sites.size(); ++si) {
      auto& calls = site_calls[si];
      if (site_batched[si] || calls.size() < 2) continue;
      bool ok = true;
      for (auto& c : calls)
        if (c.out < 0 || c.nchw_out || tens[c.out].parent < 0 || tens[c.out].parent != tens[calls[0].out].parent ||
            !tens[c.out].contiguous() || c.kind != calls[0].kind || c.nin != calls[0].nin || c.rep_in1 != calls[0].rep_in1)
          ok = false;
      if (!ok) continue;
      std::sort(calls.begin(), calls.end(), [&](const ConvOp& a, const ConvOp& b2) { return tens[a.out].off < tens[b2.out].off; });
      auto bytes = [&](int id) { return (long)tens[id].N * tens[id].H * tens[id].W * tens[id].pitch * 2; };
      int rep = 0;
      long oacc = 0, iacc[2] = {0, 0};
      for (size_t i = 0; i < calls.size() && ok; ++i) {
        const Ten& o = tens[calls[i].out];
        const Ten& o0 = tens[calls[0].out];
        if (o.off != o0.off + oacc || o.C != o0.C || o.H != o0.H || o.W != o0.W) ok = false;
        oacc += bytes(calls[i].out);
        for (int k = 0; k < calls[0].nin && ok; ++k) {
          const Ten& a0 = tens[calls[0].in[k]];
          const Ten& a = tens[calls[i].in[k]];
          if (a.C != a0.C || a.pitch != a0.pitch || a.H != a0.H || a.W != a0.W || !a.contiguous()) ok = false;
          if (k == 1 && calls[0].rep_in1) {
            if (a.off != a0.off || a.N != a0.N) ok = false;
            rep |= 2;
          } else {
            if (a.off != a0.off + iacc[k]) ok = false;
            iacc[k] += bytes(calls[i].in[k]);
          }
        }
      }
      if (!ok) continue;
      site_batched[si] = 2;
      site_rep[si] = rep;
    }
  }

  // Emitted after the whole tape: one bias column-sum and one weight-gradient GEMM per batched site.
  int emit_batched_grads() {
    for (size_t si = 0; si < sites.size(); ++si) {
      if (!site_batched[si]) continue;
      const Site& s = sites[si];
      const auto& calls = site_calls[si];
      const ConvOp& op = calls[0];
      const Ten o = tens[op.out];
      const Ten in0 = tens[op.in[0]];
      int n_out = 0, n_in = 0;  // images over all calls (per-step calls: n * B; chunk ops: the chunks' image counts)
      double flops = 0.0;
      for (auto& c : calls) {
        n_out += tens[c.out].N;
        n_in += tens[c.in[0]].N;
        flops += conv_flops(c);
      }
      const __nv_bfloat16* gz;
      if (site_batched[si] == 2) {
        ensure_gbuf(op.out);  // first chunk view -> its region of the all-T tensor's gradient buffer
        gz = P(tens[op.out].goff);
      } else {
        const Series& sr = series[o.series];
        gz = P(sr.gbase + (long)o.slot * (long)sr.slot_bytes);
      }
      cur_label = s.key;
      if (dry) continue;
      ConvDesc d;
      memset(&d, 0, sizeof(d));
      d.bias_grad = s.b_off >= 0 ? gflat + s.b_off : nullptr;
      ActSrc q;
      if (op.kind == CK_UP2) {
        d.kind = CK_UP2_DGRAD;
        d.src[0] = {gz, o.C, o.pitch};
        d.nsrc = 1;
        d.N = n_out;
        d.H = o.H;
        d.W = o.W;
        q = {P(in0.off), in0.C, in0.pitch};
      } else {
        d.kind = op.kind;
        d.nsrc = op.nin;
        for (int k = 0; k < op.nin; ++k) {
          const Ten& t = tens[op.in[k]];
          d.src[k] = {P(t.off), t.C, t.pitch, (site_rep[si] >> k) & 1 ? t.N : 0};
        }
        d.N = n_in;
        d.H = in0.H;
        d.W = in0.W;
        q = {gz, o.C, o.pitch};
      }
      WgradLaunch wl;
      if (build_wgrad(d, q, gflat + s.w_off, &wl)) return 1;
      if (s.b_off >= 0 && !wl.bias_done) emit_colsum(gz, (long)n_out * o.H * o.W, o.C, gflat + s.b_off);
      emit([wl](cudaStream_t st) mutable { return run_wgrad(wl, st); }, LC_WGRAD, flops, ":wgrad");
    }
    cur_label = "";
    return 0;
  }

  int plan(int B_, int T_, int H_, int W_, int train_, void* workspace, void* wpack, float* grad_flat, bool dry_) {
    REFID_REQUIRE(B_ >= 1 && T_ >= 1 && H_ >= 8 && W_ >= 8 && H_ % 8 == 0 && W_ % 8 == 0,
                  "plan: need B,T >= 1 and H,W multiples of 8 (got B%d T%d H%d W%d)", B_, T_, H_, W_);
    B = B_;
    T = T_;
    H = H_;
    W = W_;
    train = train_;
    dry = dry_;
    f16 = (!train && opt_infer_fp16) ? 1 : 0;
    if (!dry) drop_graphs();
    ws = static_cast<char*>(workspace);
    wmaster = static_cast<float*>(wpack);
    wpackb = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(wpack) + align_up((size_t)flat_floats * 4, 1024));
    gflat = grad_flat;
    REFID_REQUIRE(dry || !train || gflat, "plan: grad_flat is required when train != 0");
    tens.clear();
    named.clear();
    gbufs.clear();
    gfree.clear();
    fwd.clear();
    bwd.clear();
    fwd_meta.clear();
    bwd_meta.clear();
    tape.clear();
    f32_zero.clear();
    release_after.clear();
    series.clear();
    series_idx.clear();
    site_calls.clear();
    site_batched.clear();
    cview_cache.clear();
    cur_slot = -1;
    act_top = gtop = 0;
    planned = false;
    cur = &fwd;
    if (build_network()) return 1;
    gbase = act_top;
    cur = &bwd;
    if (train) {
      for (auto& z : f32_zero) {
        char* p = ws + z.first;
        const size_t n = z.second;
        emit([p, n](cudaStream_t st) {
          REFID_CUDA_CHECK(cudaMemsetAsync(p, 0, n, st));
          return 0;
        });
      }
      for (int i = (int)tape.size() - 1; i >= 0; --i)
        if (tape[i]()) return 1;
      if (emit_batched_grads()) return 1;
      if (emit_gate_wgrad(0) || emit_gate_wgrad(1)) return 1;
    }
    tape.clear();
    planned = !dry;
    return 0;
  }
};

}  // namespace refid

using refid::Engine;

extern "C" {

int refid_create(const refid_cfg* cfg, refid_handle* out) {
  using namespace refid;
  REFID_REQUIRE(cfg && out, "refid_create: null argument");
  REFID_REQUIRE(cfg->base_num_channels == 32, "refid_create: base_num_channels must be 32 (got %d)", cfg->base_num_channels);
  REFID_REQUIRE(cfg->img_chn >= 1 && cfg->img_chn <= 64 && cfg->ev_chn >= 1 && cfg->ev_chn <= 64,
                "refid_create: img_chn/ev_chn out of range");
  REFID_REQUIRE(cfg->out_chn >= 1 && cfg->out_chn <= 8, "refid_create: out_chn must be 1..8");
  Engine* e = new Engine();
  e->cfg = *cfg;
  if (const char* g = getenv("REFID_GRAPHS")) e->opt_graphs = (g[0] == '0') ? 0 : 1;  // diagnostics (ncu launch lists)
  if (const char* g = getenv("REFID_TCHUNK")) e->tchunk = atoi(g) < 0 ? 0 : (atoi(g) > 64 ? 64 : atoi(g));
  e->build_sites();
  *out = reinterpret_cast<refid_handle>(e);
  return 0;
}

int refid_destroy(refid_handle h) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (!e) return 0;
  if (e->packs_dev) cudaFree(e->packs_dev);
  e->drop_graphs();
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  delete e;
  return 0;
}

int refid_num_param_entries(refid_handle h) { return (int)reinterpret_cast<Engine*>(h)->sites.size(); }

int refid_param_entry_at(refid_handle h, int idx, refid_param_entry* out) {
  using namespace refid;
  Engine* e = reinterpret_cast<Engine*>(h);
  REFID_REQUIRE(idx >= 0 && idx < (int)e->sites.size(), "param entry %d out of range", idx);
  const auto& s = e->sites[idx];
  memset(out, 0, sizeof(*out));
  snprintf(out->key, sizeof(out->key), "%s", s.key.c_str());
  out->kind = s.kind;
  out->taps = s.taps;
  out->R = s.R;
  out->Cc = s.Cc;
  out->nbias = s.nbias;
  out->w_off = s.w_off;
  out->b_off = s.b_off;
  return 0;
}

long refid_flat_floats(refid_handle h) { return reinterpret_cast<Engine*>(h)->flat_floats; }

size_t refid_wpack_bytes(refid_handle h) {
  Engine* e = reinterpret_cast<Engine*>(h);
  return refid::align_up((size_t)e->flat_floats * 4, 1024) + (size_t)e->pack_elems * 2 + 1024;
}

int refid_workspace_bytes(refid_handle h, int B, int T, int H, int W, int train, size_t* out) {
  using namespace refid;
  Engine* e = reinterpret_cast<Engine*>(h);
  REFID_REQUIRE(e && out, "refid_workspace_bytes: null argument");
  // a dry plan on a scratch engine with the same configuration and options: the handle's own plan (if any) stays intact
  Engine tmp;
  tmp.cfg = e->cfg;
  tmp.opt_infer_fp16 = e->opt_infer_fp16;
  tmp.tchunk = e->tchunk;
  tmp.build_sites();
  if (tmp.plan(B, T, H, W, train, nullptr, nullptr, nullptr, true)) return 1;
  *out = tmp.gbase + tmp.gtop + 4096;
  return 0;
}

int refid_plan(refid_handle h, int B, int T, int H, int W, int train, void* workspace, void* wpack, float* grad_flat) {
  using namespace refid;
  Engine* e = reinterpret_cast<Engine*>(h);
  REFID_REQUIRE(workspace && wpack, "refid_plan: null buffer");
  REFID_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(wpack) & 255) == 0,
                "refid_plan: workspace and wpack must be 256-byte aligned");
  return e->plan(B, T, H, W, train, workspace, wpack, grad_flat, false);
}

int refid_pack_weights(refid_handle h, const float* flat, void* stream) {
  using namespace refid;
  Engine* e = reinterpret_cast<Engine*>(h);
  REFID_REQUIRE(e->planned, "refid_pack_weights: call refid_plan first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!e->packs_dev) {
    REFID_CUDA_CHECK(cudaMalloc(&e->packs_dev, e->packs.size() * sizeof(PackDesc)));  // 10s of KB of descriptors, not data
    REFID_CUDA_CHECK(cudaMemcpy(e->packs_dev, e->packs.data(), e->packs.size() * sizeof(PackDesc), cudaMemcpyHostToDevice));
  }
  REFID_CUDA_CHECK(cudaMemcpyAsync(e->wmaster, flat, (size_t)e->flat_floats * 4, cudaMemcpyDeviceToDevice, st));
  return launch_pack(flat, e->wpackb, e->packs_dev, (int)e->packs.size(), e->pack_max, st, e->f16);
}

int refid_forward(refid_handle h, const float* x, const float* event, float* out, void* stream) {
  using namespace refid;
  Engine* e = reinterpret_cast<Engine*>(h);
  REFID_REQUIRE(e->planned, "refid_forward: no plan");
  REFID_REQUIRE(x && event && out, "refid_forward: null tensor");
  REFID_REQUIRE_HEALTHY("refid_forward");
  e->io_x = x;
  e->io_ev = event;
  e->io_out = out;
  return e->run_list(e->fwd, e->fwd_graphs, x, event, out, false, static_cast<cudaStream_t>(stream));
}

int refid_backward(refid_handle h, const float* grad_out, void* stream) {
  using namespace refid;
  Engine* e = reinterpret_cast<Engine*>(h);
  REFID_REQUIRE(e->planned && e->train, "refid_backward: no training plan");
  REFID_REQUIRE(grad_out, "refid_backward: null tensor");
  REFID_REQUIRE_HEALTHY("refid_backward");
  e->io_gout = grad_out;
  return e->run_list(e->bwd, e->bwd_graphs, grad_out, nullptr, nullptr, true, static_cast<cudaStream_t>(stream));
}

int refid_set_option(refid_handle h, const char* name, long value) {
  using namespace refid;
  Engine* e = reinterpret_cast<Engine*>(h);
  REFID_REQUIRE(e && name, "refid_set_option: null argument");
  const std::string n(name);
  if (n == "tchunk") {
    REFID_REQUIRE(value >= 0 && value <= 64, "refid_set_option: tchunk must be in [0, 64]");
    e->tchunk = (int)value;
    return 0;
  }
  if (n == "infer_fp16") {
    REFID_REQUIRE(!e->planned, "refid_set_option: infer_fp16 must be set before refid_plan");
    e->opt_infer_fp16 = value ? 1 : 0;
  } else if (n == "graphs") {
    e->opt_graphs = value ? 1 : 0;
    if (!value) e->drop_graphs();
  } else {
    set_error("refid_set_option: unknown option '%s'", name);
    return 1;
  }
  return 0;
}

int refid_graph_stats(refid_handle h, long out[4]) {
  Engine* e = reinterpret_cast<Engine*>(h);
  out[0] = e->graph_captures;
  out[1] = e->graph_replays;
  out[2] = e->graph_eager;
  out[3] = e->graph_failures;
  return 0;
}

int refid_plan_storage(refid_handle h) { return reinterpret_cast<Engine*>(h)->f16; }

const char* refid_graph_error(refid_handle h) { return reinterpret_cast<Engine*>(h)->graph_error.c_str(); }

// Re-runs forward (+ backward) of the current plan on the last call's tensors with a CUDA event pair around every
// launch and sums device time, launch count and algorithmic FLOPs per launch class
// (0 conv forward tap-GEMM, 1 data-gradient tap-GEMM, 2 weight-gradient GEMM, 3 memory-bound kernels / memsets).
int refid_profile(refid_handle h, int with_backward, double* ms, double* flops, long* launches, void* stream) {
  using namespace refid;
  Engine* e = reinterpret_cast<Engine*>(h);
  REFID_REQUIRE(e->planned && e->io_x && e->io_out, "refid_profile: run refid_forward first");
  REFID_REQUIRE(!with_backward || (e->train && e->io_gout), "refid_profile: run refid_backward first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int c = 0; c < LC_COUNT; ++c) {
    ms[c] = 0.0;
    flops[c] = 0.0;
    launches[c] = 0;
  }
  const size_t n = e->fwd.size() + (with_backward ? e->bwd.size() : 0);
  std::vector<cudaEvent_t> ev(2 * n);
  for (auto& x : ev) REFID_CUDA_CHECK(cudaEventCreate(&x));
  size_t k = 0;
  if (with_backward) REFID_CUDA_CHECK(cudaMemsetAsync(e->gflat, 0, (size_t)e->flat_floats * 4, st));
  for (int pass = 0; pass < (with_backward ? 2 : 1); ++pass) {
    auto& ls = pass ? e->bwd : e->fwd;
    for (size_t i = 0; i < ls.size(); ++i, ++k) {
      REFID_CUDA_CHECK(cudaEventRecord(ev[2 * k], st));
      if (ls[i](st)) return 1;
      REFID_CUDA_CHECK(cudaEventRecord(ev[2 * k + 1], st));
    }
  }
  REFID_CUDA_CHECK(cudaStreamSynchronize(st));
  k = 0;
  for (int pass = 0; pass < (with_backward ? 2 : 1); ++pass) {
    auto& ms_meta = pass ? e->bwd_meta : e->fwd_meta;
    for (size_t i = 0; i < ms_meta.size(); ++i, ++k) {
      float t = 0.f;
      REFID_CUDA_CHECK(cudaEventElapsedTime(&t, ev[2 * k], ev[2 * k + 1]));
      ms[ms_meta[i].cls] += t;
      flops[ms_meta[i].cls] += ms_meta[i].flops;
      launches[ms_meta[i].cls] += 1;
    }
  }
  for (auto& x : ev) cudaEventDestroy(x);
  return 0;
}

// Same replay as refid_profile, one CSV row per launch: pass,index,class,label,ms,gflop.
int refid_profile_csv(refid_handle h, int with_backward, const char* path, void* stream) {
  using namespace refid;
  Engine* e = reinterpret_cast<Engine*>(h);
  REFID_REQUIRE(e->planned && e->io_x && e->io_out, "refid_profile_csv: run refid_forward first");
  REFID_REQUIRE(!with_backward || (e->train && e->io_gout), "refid_profile_csv: run refid_backward first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = e->fwd.size() + (with_backward ? e->bwd.size() : 0);
  std::vector<cudaEvent_t> ev(2 * n);
  for (auto& x : ev) REFID_CUDA_CHECK(cudaEventCreate(&x));
  size_t k = 0;
  if (with_backward) REFID_CUDA_CHECK(cudaMemsetAsync(e->gflat, 0, (size_t)e->flat_floats * 4, st));
  for (int pass = 0; pass < (with_backward ? 2 : 1); ++pass) {
    auto& ls = pass ? e->bwd : e->fwd;
    for (size_t i = 0; i < ls.size(); ++i, ++k) {
      REFID_CUDA_CHECK(cudaEventRecord(ev[2 * k], st));
      if (ls[i](st)) return 1;
      REFID_CUDA_CHECK(cudaEventRecord(ev[2 * k + 1], st));
    }
  }
  REFID_CUDA_CHECK(cudaStreamSynchronize(st));
  FILE* f = fopen(path, "w");
  REFID_REQUIRE(f != nullptr, "refid_profile_csv: cannot open %s", path);
  fprintf(f, "pass,index,class,label,ms,gflop\n");
  k = 0;
  for (int pass = 0; pass < (with_backward ? 2 : 1); ++pass) {
    auto& mm = pass ? e->bwd_meta : e->fwd_meta;
    for (size_t i = 0; i < mm.size(); ++i, ++k) {
      float t = 0.f;
      REFID_CUDA_CHECK(cudaEventElapsedTime(&t, ev[2 * k], ev[2 * k + 1]));
      fprintf(f, "%s,%zu,%d,%s,%.6f,%.6f\n", pass ? "bwd" : "fwd", i, mm[i].cls, mm[i].label.c_str(), t, mm[i].flops * 1e-9);
    }
  }
  fclose(f);
  for (auto& x : ev) cudaEventDestroy(x);
  return 0;
}

int refid_set_pdl(int enable) { return refid::set_pdl(enable); }

int refid_num_launches(refid_handle h, int* fwd, int* bwd) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (fwd) *fwd = (int)e->fwd.size();
  if (bwd) *bwd = (int)e->bwd.size();
  return 0;
}

int refid_debug_tensor(refid_handle h, const char* name, void** ptr, int* N, int* H, int* W, int* C, int* pitch) {
  using namespace refid;
  Engine* e = reinterpret_cast<Engine*>(h);
  auto it = e->named.find(name);
  REFID_REQUIRE(it != e->named.end(), "no tensor named '%s'", name);
  const auto& t = e->tens[it->second];
  *ptr = e->ws + t.off;
  *N = t.N;
  *H = t.H;
  *W = t.W;
  *C = t.C;
  *pitch = t.pitch;
  return 0;
}

}  // extern "C"
