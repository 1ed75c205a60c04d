/* refid_b200 -- C ABI of the B200-native (sm_100a) backend for REFID's FinalBidirectionAttenfusion forward/backward.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference has no FFI of its own (it is pure
 * PyTorch); the interface replaced is the Python call `net_g(x=lq, event=voxel)` made by the model wrappers
 *   basicsr/models/twoImage_event_recurrent_model.py:276,324   (train / eval)
 *   basicsr/models/Test_twoImage_event_recurrent_model.py:222  (test)
 * on a module built by `define_network` (basicsr/models/archs/__init__.py:43-46) from
 *   basicsr/models/archs/XXNet_final_attenfusion_arch.py:81-218 (FinalBidirectionAttenfusion),
 * plus the backward pass autograd derives from it (`l_total.backward()`, twoImage_event_recurrent_model.py:303).
 * The ctypes binding a maintainer adds is shown in INTEGRATION.md and shipped as refid_b200/engine.py.
 *
 * Conventions
 *   - plain pointers and sizes only; every device buffer is allocated and owned by the caller (PyTorch);
 *     the library borrows raw device pointers and never allocates device memory for data;
 *   - every function returns 0 on success; on failure the message is available from refid_last_error()
 *     (thread-local); no exceptions cross the ABI and nothing is silently "handled";
 *   - all work is enqueued asynchronously on the caller's CUDA stream (cudaStream_t passed as void*);
 *   - one handle may be used from any thread, one call at a time.
 */
#ifndef REFID_B200_H
#define REFID_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct refid_engine* refid_handle;

/* Constructor arguments of FinalBidirectionAttenfusion that change tensor shapes
 * (XXNet_final_attenfusion_arch.py:90-92).  base_num_channels must be 32, num_encoders 3, num_block 1,
 * num_residual_blocks 2, skip_type 'sum', norm None (the only configuration any shipped option file uses). */
typedef struct {
  int img_chn;
  int ev_chn;
  int out_chn; /* <= 8; 3 in every shipped config */
  int base_num_channels;
} refid_cfg;

/* One entry of the flat fp32 parameter vector ("gradient layout") the engine consumes and whose gradient it
 * produces.  kind: 0 conv3x3, 1 conv1x1, 2 conv4x4s2, 3 convT2x2s2, 4 head 5x5 (x-unrolled rows), 5 raw vector.
 * Weight layout: [taps][R][Cc] fp32 at float offset w_off (conv: R = Cin (or padded 5*Cin), Cc = Cout, tap = ky*kw+kx;
 * convT: R = Cout, Cc = Cin, tap = a*2+b).  Bias: nbias floats at b_off (-1: none). */
typedef struct {
  char key[96];
  int kind, taps, R, Cc, nbias;
  long w_off, b_off;
} refid_param_entry;

const char* refid_last_error(void);

int refid_create(const refid_cfg* cfg, refid_handle* out);
int refid_destroy(refid_handle h);

int refid_num_param_entries(refid_handle h);
int refid_param_entry_at(refid_handle h, int idx, refid_param_entry* out);
long refid_flat_floats(refid_handle h);    /* length of the flat fp32 parameter / gradient vectors */
size_t refid_wpack_bytes(refid_handle h);  /* persistent device buffer: fp32 master copy + bf16 UMMA-ready packs */

/* Bytes of workspace for one (B,T,H,W) problem; H, W % 8 == 0.  train != 0: every activation the backward needs for all
 * T steps plus the gradient pool; train == 0: per-step buffers are recycled, O(1) in T apart from the hoisted event head. */
int refid_workspace_bytes(refid_handle h, int B, int T, int H, int W, int train, size_t* out);

/* Build the launch plan (TMA tensor maps are encoded here) for fixed buffers.  grad_flat may be NULL when train == 0. */
int refid_plan(refid_handle h, int B, int T, int H, int W, int train, void* workspace, void* wpack, float* grad_flat);

/* flat (device, fp32, refid_flat_floats) -> master copy + bf16 packed weights inside wpack. */
int refid_pack_weights(refid_handle h, const float* flat, void* stream);

/* x: (B,img_chn,H,W) fp32 NCHW; event: (B,T,ev_chn,H,W) fp32; out: (B,T,out_chn,H,W) fp32.  Device pointers. */
int refid_forward(refid_handle h, const float* x, const float* event, float* out, void* stream);
/* grad_out: (B,T,out_chn,H,W) fp32.  Zeroes grad_flat, then accumulates d(loss)/d(flat) into it.
 * Must follow a refid_forward on the same plan (train != 0). */
int refid_backward(refid_handle h, const float* grad_out, void* stream);

/* Parameter tensors <-> flat vector in one launch each way (replaces ~300 framework ops per training step).  Entry: `ptr` =
 * the source tensor (gather) or the destination gradient tensor (scatter), device fp32, contiguous; n = taps*R*Cc floats at
 * float offset flat_off.  mode 0: plain copy; mode 1: conv weight (Cc,R,taps) = PyTorch (Cout,Cin,kh,kw) <-> [tap][R][Cc].
 * refid_flat_gather zero-fills `flat` first (alignment gaps, padded rows).  The table is host memory, read during the call. */
typedef struct {
  const void* ptr;
  long flat_off;
  int mode, taps, R, Cc;
} refid_flat_entry;
int refid_flat_gather(const refid_flat_entry* entries, int n, float* flat, long flat_floats, void* stream);
int refid_flat_scatter(const refid_flat_entry* entries, int n, const float* gflat, void* stream);

/* Engine options (name, value):
 *   "infer_fp16" (default 1; set before refid_plan): forward-only plans (train == 0) keep activations and packed weights
 *       as fp16 instead of bf16 -- same bytes and tensor-core rate, 11 instead of 8 significant bits, which brings the
 *       output within ~1e-3 max-abs / 2.5e-4 RMS of the fp32 reference (the PSNR-parity bar); training plans are bf16.
 *   "graphs" (default 1): refid_forward / refid_backward replay their launch list as a CUDA graph once the same
 *       (x, event, out) / grad_out pointers are seen a second time; 0 = plain launches.
 *   "tchunk" (default 64, i.e. as many as the limit allows; set before refid_plan / refid_workspace_bytes; env REFID_TCHUNK): training plans run the two
 *       sweeps level by level and every op without a recurrence (EGACA, in-convs, fuse_two_dir, `down`, bottleneck,
 *       transposed convs) once per chunk of min(tchunk, 192 / B, T) time steps on that many x B images instead of once per step; only the
 *       recurrent trunks run step by step.  0 = the step-major schedule (the order of the reference's Python loop, which
 *       forward-only plans always use).  Results are identical up to fp32 summation order of the weight gradients.
 * refid_graph_stats: {graphs captured, graph replays, eager runs, capture failures}.  refid_plan_storage: 1 = the current
 * plan stores fp16, 0 = bf16. */
int refid_set_option(refid_handle h, const char* name, long value);
int refid_graph_stats(refid_handle h, long out[4]);
const char* refid_graph_error(refid_handle h); /* why the last capture failed ("" if none did) */
int refid_plan_storage(refid_handle h);

/* Roofline accounting: re-runs forward (+ backward) of the current plan on the tensors of the last call with a CUDA
 * event pair around every launch; sums device milliseconds, algorithmic FLOPs and launch counts per class:
 * [0] other conv-shaped forward launches (1x1, 32-channel, stride-2, heads, pred), [1] their data-gradients,
 * [2] weight-gradient GEMMs, [3] memory-bound kernels, [4] stride-1 3x3 convs with >= 64 channels (halo-conv engine,
 * the dominant kernel) forward, [5] their data-gradients. */
int refid_profile(refid_handle h, int with_backward, double ms[6], double flops[6], long launches[6], void* stream);

/* Same replay, one CSV row per launch (pass,index,class,label,ms,gflop) written to `path`. */
int refid_profile_csv(refid_handle h, int with_backward, const char* path, void* stream);

/* Introspection used by tests and profiling. */
int refid_num_launches(refid_handle h, int* fwd, int* bwd);
/* Device pointer + shape of a named intermediate activation (NHWC bf16, `pitch` channels per pixel). */
int refid_debug_tensor(refid_handle h, const char* name, void** ptr, int* N, int* H, int* W, int* C, int* pitch);
int refid_abort_flag(unsigned int* out); /* non-zero: a bounded mbarrier wait timed out inside a kernel; synchronises, clears */
/* Same condition without synchronising or clearing (a pinned host word the timing-out kernel writes): once non-zero,
 * refid_forward / refid_backward refuse to run until refid_abort_flag() has been called. */
unsigned int refid_abort_pending(void);
/* Programmatic dependent launch (each kernel's set-up overlaps the tail of the kernel before it) on / off for all
 * launches issued afterwards; default on unless the environment has REFID_PDL=0.  Returns the previous setting. */
int refid_set_pdl(int enable);

/* Charbonnier loss, value and gradient in one pass (replaces CharbonnierLoss.forward + its autograd
 * backward, basicsr/models/losses/losses.py:28-30,143-173): loss = loss_weight * (mean | sum) sqrt((pred-target)^2 + eps),
 * grad (nullable) = d loss / d pred.  fp32, n elements, 16-byte aligned; `scratch` = refid_charbonnier_scratch_bytes()
 * bytes of device memory; `loss` = one device float.  Asynchronous on `stream`, bit-reproducible. */
int refid_charbonnier_scratch_bytes(void);
int refid_charbonnier(const float* pred, const float* target, float* grad, void* scratch, float* loss, long n, float eps,
                      float loss_weight, int reduction_mean, void* stream);
/* x[0..n) *= *scalar (one device float; e.g. the upstream gradient of the loss); no memory traffic when it is exactly 1. */
int refid_scale_by_device_scalar(float* x, const float* scalar, long n, void* stream);

/* Global-norm gradient clip + Adam / AdamW step over all parameter tensors as two multi-tensor passes (SURVEY.md 8f rank
 * 2; replaces `torch.nn.utils.clip_grad_norm_(net_g.parameters(), 0.01)` + `optimizer_g.step()`,
 * basicsr/models/twoImage_event_recurrent_model.py:304-307, optimizer set-up :67-95).  `numel[i]` elements per tensor, fixed
 * at create; per step the caller passes host arrays of fp32 device pointers (a NULL gradient skips that parameter, as torch
 * does).  max_norm <= 0: no clipping.  `steps[i]` = 1-based step count of tensor i (bias correction).  decoupled = 1: AdamW, 0: Adam with
 * L2 weight decay.  norm_out (optional, device, 2 floats) receives {total gradient norm, clip coefficient}.
 * Gradients are NOT rescaled in memory (the reference's clipped .grad tensors are never read again).  Asynchronous. */
int refid_optim_create(int ntensors, const long* numel, void** handle);
int refid_optim_destroy(void* handle);
int refid_optim_step(void* handle, float* const* params, const float* const* grads, float* const* exp_avg,
                     float* const* exp_avg_sq, float max_norm, float lr, float beta1, float beta2, float eps,
                     float weight_decay, const long* steps, int decoupled, float* norm_out, void* stream);

/* tensor2img quantisation + PSNR sums on the GPU (replaces `tensor2img` on result and gt followed by
 * `calculate_psnr`, basicsr/utils/img_util.py:59-121, basicsr/metrics/psnr_ssim.py:9-61, per validation frame).  pred / gt:
 * `frames` x (C,H,W) fp32 on the device.  Every value is clamped to [0,1], x255, rounded half to even.  With gt != NULL:
 * ssd[f] = exact integer sum of squared uint8 differences inside the crop border, max_pred[f] = largest quantised pred
 * value there (calculate_psnr's `max_value` rule); PSNR = 20 log10(255 / sqrt(ssd / count)) is formed by the caller in
 * double.  img_pred / img_gt (nullable): the uint8 images, (H,W,C) per frame, channels reversed (RGB -> BGR) if asked. */
int refid_quant_psnr(const float* pred, const float* gt, int frames, int C, int H, int W, int crop_border,
                     int reverse_channels, unsigned long long* ssd, unsigned int* max_pred, unsigned char* img_pred,
                     unsigned char* img_gt, void* stream);

/* Full-frame validation tiling (replaces `grids` / `grids_voxel` / `grids_inverse`,
 * basicsr/models/twoImage_event_recurrent_model.py:128-270, transposes :115-126).  idx: ncrops x {i, j, trans_idx} int32 on the
 * device (placement computed by the caller as the reference does, :201-243); trans_idx 0..7 = rot90 x (k % 4) of the crop,
 * W-flipped first when k >= 4.  crop: (planes,H,W) -> (ncrops,planes,cs,cs).  merge: (ncrops,planes,cs,cs) -> (planes,H,W),
 * every pixel the mean of the crops covering it (accumulated in crop order, as the reference's loop does). */
int refid_grids_crop(const float* src, int planes, int H, int W, const int* idx, int ncrops, int crop_size, float* dst,
                     void* stream);
int refid_grids_merge(const float* parts, int planes, int H, int W, const int* idx, int ncrops, int crop_size, float* dst,
                      void* stream);

/* Training-sample assembly (second half; replaces the crop / augment / channel-packing steps of
 * `__getitem__`, basicsr/data/image_npy_dataset.py:189-232 with basicsr/data/transforms.py:88-129,163-238):
 * dst[p][y][x] = plane plane_table[p] (>= 0: plane of src_a, < 0: plane -1-v of src_b; both (planes,H,W) fp32) at the crop
 * window (top,left,ph,pw), horizontally / vertically flipped, then transposed if rot90 (dst is (nplanes, pw, ph) then). */
int refid_crop_flip_gather(const float* src_a, const float* src_b, int H, int W, const int* plane_table, int nplanes,
                           int top, int left, int ph, int pw, int hflip, int vflip, int rot90, float* dst, void* stream);

/* Event -> voxel-grid rasterisation (replaces `events_to_voxel_grid`, basicsr/data/event_util.py:6-66).
 * events: n rows of float32 [timestamp, x, y, polarity] on the device (the reference's array layout), 16-byte aligned;
 * voxel: (num_bins,height,width) fp32, or (height,width,num_bins) with hwc != 0; scratch: num_bins*height*width*8 bytes.
 * Time span = first and last row's stamps, polarity 0 counts as -1, bilinear weights in double, accumulated exactly in
 * fixed point (bit-reproducible); events outside the grid are dropped.  Asynchronous on `stream`. */
int refid_events_to_voxel(const float* events, long n, int num_bins, int width, int height, int hwc, void* scratch,
                          float* voxel, void* stream);

/* Single-kernel entry points (unit tests, ncu captures). */
int refid_test_conv(int kind, int parity, const void* in0, int C0, const void* in1, int C1, int N, int H, int W,
                    const void* w, long w_rows, int w_cols, int wrows_per_tap, int w_row0, int Cout, const float* bias,
                    const void* pre, const void* sv, int act, float slope, void* out, void* out_b, void* out2,
                    const void* post, float* out_f32, void* stream);
int refid_test_wgrad(int kind, const void* p0, int C0, const void* p1, int C1, int N, int H, int W, const void* q, int CQ,
                     float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
