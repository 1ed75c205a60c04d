// C-ABI entry points that expose single kernels for unit tests and ncu captures (declared in include/refid_b200.h).
#include "convop.cuh"

using namespace refid;

extern "C" {

const char* refid_last_error(void) { return get_error(); }

int refid_abort_flag(unsigned int* out) { return read_and_clear_abort_flag(0, out); }
unsigned int refid_abort_pending(void) { return abort_pending(); }

int refid_test_conv(int kind, int parity, const void* in0, int C0, const void* in1, int C1, int N, int H, int W,
                    const void* w, long w_rows, int w_cols, int wrows_per_tap, int w_row0, int Cout, const float* bias,
                    const void* pre, const void* sv, int act, float slope, void* out, void* out_b, void* out2,
                    const void* post, float* out_f32, void* stream) {
  ConvDesc d;
  memset(&d, 0, sizeof(d));
  d.kind = kind;
  d.parity = parity;
  d.src[0] = {static_cast<const __nv_bfloat16*>(in0), C0, C0};
  d.nsrc = 1;
  if (in1) {
    d.src[1] = {static_cast<const __nv_bfloat16*>(in1), C1, C1};
    d.nsrc = 2;
  }
  d.N = N;
  d.H = H;
  d.W = W;
  d.w = static_cast<const __nv_bfloat16*>(w);
  d.w_rows = w_rows;
  d.w_cols = w_cols;
  d.wrows_per_tap = wrows_per_tap;
  d.w_row0 = w_row0;
  OutGroup g[2];
  memset(g, 0, sizeof(g));
  int ng = 1;
  const int cg = out_b ? Cout / 2 : Cout;
  g[0].channels = cg;
  g[0].epi.out = static_cast<__nv_bfloat16*>(out);
  g[0].epi.out2 = static_cast<__nv_bfloat16*>(out2);
  g[0].epi.post = static_cast<const __nv_bfloat16*>(post);
  g[0].epi.out_f32 = out_f32;
  g[0].epi.pre = static_cast<const __nv_bfloat16*>(pre);
  g[0].epi.sv = static_cast<const __nv_bfloat16*>(sv);
  g[0].epi.bias = bias;
  g[0].epi.C = cg;
  g[0].epi.act = act;
  g[0].epi.slope = slope;
  if (out_b) {
    g[1] = g[0];
    g[1].epi.out = static_cast<__nv_bfloat16*>(out_b);
    g[1].epi.bias = bias ? bias + cg : nullptr;
    ng = 2;
  }
  TapGemmLaunch l;
  if (build_conv(d, g, ng, &l)) return 1;
  return run_conv(l, static_cast<cudaStream_t>(stream));
}

int refid_test_wgrad(int kind, const void* p0, int C0, const void* p1, int C1, int N, int H, int W, const void* q, int CQ,
                     float* out, void* stream) {
  ConvDesc d;
  memset(&d, 0, sizeof(d));
  d.kind = kind;
  d.src[0] = {static_cast<const __nv_bfloat16*>(p0), C0, C0};
  d.nsrc = 1;
  if (p1) {
    d.src[1] = {static_cast<const __nv_bfloat16*>(p1), C1, C1};
    d.nsrc = 2;
  }
  d.N = N;
  d.H = H;
  d.W = W;
  WgradLaunch l;
  if (build_wgrad(d, ActSrc{static_cast<const __nv_bfloat16*>(q), CQ, CQ}, out, &l)) return 1;
  return run_wgrad(l, static_cast<cudaStream_t>(stream));
}


}  // extern "C"
