/*
 * plc.h -- C ABI of the B200-native ConvLSTM recurrence (libplc.so).
 *
 * The reference (Tomzhuiowewie/Pl-ConvLSTM-GAN) has no FFI: its hot path is the Python
 * nn.Module `ConvLSTMCell` (src/models/convlstm.py:4-28) called from the T-loop of
 * `Generator.forward` (src/models/generator.py:156-171).  This header is the boundary a
 * replacement binds instead of `nn.Conv2d` + the ~10 ATen pointwise ops of that module;
 * every entry point cites the reference lines it replaces.  Host-side mirror of the
 * nn.Module API: pl-convlstm-gan_b200/nn.py; binding stub: INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; all device buffers are caller-owned.
 *   - activations are NHWC ("channels_last"): element (b, y, x, ch) at ((b*H + y)*W + x)*C + ch.
 *     The logical shape stays the reference's [B, C, H, W] (torch.channels_last strides).
 *   - every function returns PLC_OK (0) or a negative PlcStatus; the message is available
 *     from plc_last_error() (thread-local).  There is NO fallback path: an unsupported
 *     shape, a misaligned pointer or a missing sm_100 device is an error.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - re-entrant; callable from PyTorch's autograd worker threads.
 */
#ifndef PLC_H_
#define PLC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLC_ABI_VERSION 1

typedef enum PlcStatus {
  PLC_OK = 0,
  PLC_ERR_BAD_DESC = -1,     /* null / inconsistent descriptor                         */
  PLC_ERR_UNSUPPORTED = -2,  /* shape or mode this build has no kernel for              */
  PLC_ERR_ALIGNMENT = -3,    /* pointer not 16-byte aligned / channel count constraint  */
  PLC_ERR_NULL_ARG = -4,
  PLC_ERR_CUDA = -5,         /* CUDA runtime / driver error (message has the detail)    */
  PLC_ERR_WORKSPACE = -6     /* workspace too small                                     */
} PlcStatus;

typedef enum PlcMode {
  /* bf16 operands, fp32 accumulation in TMEM (tcgen05), fp32 cell state, tanh.approx gates.
     x, h: bf16; c: fp32.  Requires Cin % 8 == 0 (0 allowed), Ch % 16 == 0.                */
  PLC_MODE_BF16_TC = 0,
  /* fp32 validation mode: fp32 SIMT FMA, exact expf/tanhf.  x, h, c: fp32.  Any Cin, Ch.  */
  PLC_MODE_FP32 = 1
} PlcMode;

/* which packed-weight image plc_pack_weight produces */
typedef enum PlcPackKind {
  PLC_PACK_FWD = 0,   /* gate conv of convlstm.py:18 (also used for gate recompute in backward) */
  PLC_PACK_DGRAD = 1  /* transposed/flipped image for d(cat(x,h)) = conv_transpose(dZ, W)      */
} PlcPackKind;

/* One ConvLSTM cell step problem.  Mirrors ConvLSTMCell(input_dim=Cin, hidden_dim=Ch,
 * kernel_size=k, bias=has_bias) of convlstm.py:5-14 applied to a [B, *, H, W] batch.       */
typedef struct PlcCellDesc {
  int32_t B, H, W;
  int32_t Cin;      /* channels of x (0 = no input tensor: forecaster first layer)          */
  int32_t Ch;       /* hidden channels; conv has 4*Ch output channels in i,f,o,g order      */
  int32_t k;        /* odd kernel size; zero "same" padding k/2, stride 1 (convlstm.py:12)  */
  int32_t mode;     /* PlcMode                                                              */
  int32_t has_bias; /* convlstm.py:13                                                       */
} PlcCellDesc;

/* ---- introspection ------------------------------------------------------------------- */
int plc_abi_version(void);
const char* plc_last_error(void);
/* 1 if device `dev` can run this library (compute capability 10.x), else 0 */
int plc_device_supported(int dev);

/* ---- weights --------------------------------------------------------------------------
 * Replaces: the implicit weight layout of nn.Conv2d at convlstm.py:8-14.
 * `w_oihw` is the reference parameter `conv.weight` [4Ch, Cin+Ch, k, k] fp32 (device),
 * x channels first then h channels; output channels in gate order i, f, o, g.            */
size_t plc_packed_weight_bytes(const PlcCellDesc* d, int pack_kind);
int plc_pack_weight(const PlcCellDesc* d, int pack_kind, const float* w_oihw, void* w_packed, void* stream);

/* ---- forward cell step ------------------------------------------------------------------
 * Replaces convlstm.py:17-28 (cat, conv, split, 3x sigmoid + tanh, c/h update) in ONE kernel.
 *   x      [B,H,W,Cin]  (NULL iff Cin == 0)      h_prev [B,H,W,Ch]      c_prev [B,H,W,Ch]
 *   bias   [4Ch] fp32 in reference order (NULL iff !has_bias)
 *   h_out  [B,H,W,Ch]   c_out [B,H,W,Ch]   (must not alias h_prev / c_prev may alias c_out)
 *   gates_out: optional [B,H,W,4Ch] (i,f,o,g activations, mode dtype for x) or NULL.
 * dtypes: PLC_MODE_BF16_TC: x,h bf16, c fp32.  PLC_MODE_FP32: all fp32.                     */
/* Zero-initial-state form of plc_cell_fwd: every sequence of the reference starts from h = c = 0
 * (generator.py:156-160).  When this returns 1 for the descriptor, plc_cell_fwd accepts h_prev = c_prev = NULL and
 * skips the h taps (half of the K loop) and the c_prev read instead of multiplying by zero tensors; results are
 * bit-identical to passing zeros.  Returns 0 when the shape/mode needs the zero tensors (fp32 mode, Cin = 0, or
 * narrow channel boxes whose x taps do not end on a K-stage boundary).                                        */
int plc_cell_fwd_zero_state_ok(const PlcCellDesc* d);
int plc_cell_fwd(const PlcCellDesc* d, const void* x, const void* h_prev, const void* c_prev,
                 const void* w_packed_fwd, const float* bias, void* h_out, void* c_out, void* gates_out,
                 void* stream);

/* ---- backward (BPTT) cell step -----------------------------------------------------------
 * Replaces autograd of convlstm.py:17-28: recomputes the gates from the
 * saved (x, h_prev, c_prev), forms dZ, then dgrad and wgrad.
 *   dh      [B,H,W,Ch]  gradient w.r.t. h_out, dtype of h
 *   dh2     [B,H,W,Ch]  optional second contribution, added to dh inside the kernel (NULL = none):
 *                       in BPTT h_t feeds both step t+1 of the same layer and the layer above.
 *   dc_next [B,H,W,Ch]  gradient w.r.t. c_out, fp32 (NULL = zeros)
 *   dx      [B,H,W,Cin] out (NULL = not needed), dtype of x
 *   dh_prev [B,H,W,Ch]  out, dtype of h;   dc_prev [B,H,W,Ch] out fp32 (may alias dc_next)
 *   dW_acc  the library's fp32 weight-gradient ACCUMULATOR IMAGE (plc_wgrad_acc_bytes(d) bytes, zeroed by the caller
 *           before the first step, accumulated into across all T steps).  bf16 mode: packed [4Ch][column blocks*64]
 *           so the kernel can use 16-byte vector reductions; fp32 mode: the OIHW layout itself.  Convert once per
 *           backward pass with plc_wgrad_unpack, which ADDS it into `conv.weight.grad` [4Ch, Cin+Ch, k, k].
 *   db_acc  [4Ch] fp32, accumulated (NULL iff !has_bias)
 *   workspace: plc_bwd_workspace_bytes(d) bytes of scratch (holds dZ).                          */
size_t plc_bwd_workspace_bytes(const PlcCellDesc* d);
size_t plc_wgrad_acc_bytes(const PlcCellDesc* d);
int plc_wgrad_unpack(const PlcCellDesc* d, const float* dW_acc, float* dW_oihw, void* stream);
int plc_cell_bwd(const PlcCellDesc* d, const void* x, const void* h_prev, const void* c_prev,
                 const void* w_packed_fwd, const void* w_packed_dgrad, const float* bias, const void* dh,
                 const void* dh2, const float* dc_next, void* dx, void* dh_prev, float* dc_prev, float* dW_acc, float* db_acc,
                 void* workspace, size_t workspace_bytes, void* stream);

/* Weight / bias gradient alone: dW_acc, db_acc += sum over the pixels of d->B images of dZ x cat(x, h_prev) -- the third
 * stage of plc_cell_bwd, which skips it when called with dW_acc = NULL.  `dz` [B,H,W,4Ch] is what plc_cell_bwd left in
 * its workspace.  The images need not come from one time step: a rollout that keeps dZ of all T steps (a [T*B,H,W,4Ch]
 * ring next to the state rings, whose slices ARE x and h_prev of every step) runs ONE call with B = T*B per layer
 * instead of T -- the accumulator flush (partial tiles reduced into HBM by red.add), prologue and tail of a launch are
 * fixed costs that dominate at small per-GPU batches.                                           */
int plc_cell_wgrad(const PlcCellDesc* d, const void* x, const void* h_prev, const void* dz, float* dW_acc, float* db_acc,
                   void* stream);

/* ---- saved-gates BPTT (optional; trades HBM for tensor time and power) -----------------------------------------
 * The reference's autograd keeps the activated gates of every step (convlstm.py:21-24 outputs are saved tensors);
 * plc_cell_bwd recomputes them instead (one extra gate contraction per step, no extra memory).  Where memory allows,
 * the forward pass can keep them -- bf16, 8 bytes per hidden element per step -- and the backward gate-gradient kernel
 * then runs WITHOUT its mainloop: it streams the saved gates (bulk copies) next to c_prev / dc_next / dh / dh2 and is
 * HBM-bound.  On a power-capped B200 that removes a quarter of the cell path's tensor-core energy.
 *   plc_saved_gates_bytes: size of the per-step buffer, or 0 when the descriptor has no saved-gates form (fp32 mode,
 *     Ch % 64 != 0) -- then use plc_cell_fwd / plc_cell_bwd.  The layout is private to the library (tile-major,
 *     see ConvTcParams::gates_saved) and tied to the descriptor: pass the same descriptor to both calls.
 *   plc_cell_fwd_save: plc_cell_fwd that also fills `gates_saved` (h_prev = c_prev = NULL zero-state form allowed).
 *   plc_cell_bwd_saved: plc_cell_bwd without w_packed_fwd / bias, reading `gates_saved` of the same step.  Gradients
 *     differ from plc_cell_bwd only by the bf16 rounding of the stored gates (inside the 1e-2 bf16 budget).        */
size_t plc_saved_gates_bytes(const PlcCellDesc* d);
int plc_cell_fwd_save(const PlcCellDesc* d, const void* x, const void* h_prev, const void* c_prev,
                      const void* w_packed_fwd, const float* bias, void* h_out, void* c_out, void* gates_saved,
                      void* stream);
int plc_cell_bwd_saved(const PlcCellDesc* d, const void* x, const void* h_prev, const void* c_prev,
                       const void* gates_saved, const void* w_packed_dgrad, const void* dh, const void* dh2,
                       const float* dc_next, void* dx, void* dh_prev, float* dc_prev, float* dW_acc, float* db_acc,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- generic "same" convolution on the same tensor-core core (bf16; SURVEY.md section 8f "next-1") ----------------
 * Replaces the plain nn.Conv2d layers of the reference Generator body: UpsampleBlock conv + PixelShuffle(2) + ReLU
 * (generator.py:10-28), post_process convs (generator.py:67-71), attention convs (attention.py:6-10).
 *   x   [B,H,W,Cin] bf16 NHWC;  out [B,H,W,Cout] bf16, or with pixel_shuffle [B,2H,2W,Cout/4] (torch PixelShuffle(2)
 *   semantics: natural output channel n = c*4 + i*2 + j -> out[b, 2y+i, 2x+j, c]);  relu applied last.
 *   Cin % 8 == 0, Cout % 8 == 0 (pixel_shuffle: Cout % 32 == 0): callers zero-pad channels.
 * plc_conv_pack_weight: w_oihw [Cout,Cin,k,k] fp32 -> packed image (PLC_PACK_FWD also writes bias_packed [Cout] in
 *   packed column order; PLC_PACK_DGRAD the flipped/transposed image).
 * Backward: plc_conv_grad_mask forms dZ [B,H,W,Cout] (natural channel order) = dY * (Y > 0) undoing the shuffle;
 *   plc_conv_bwd: dx [B,H,W,Cin] (nullable), dW_acc = accumulator image (plc_conv_wgrad_acc_bytes; convert with
 *   plc_conv_wgrad_unpack, which adds into [Cout,Cin,k,k]), db_acc [Cout] fp32 += (nullable).                   */
typedef struct PlcConvDesc {
  int32_t B, H, W, Cin, Cout, k;
  int32_t relu;           /* apply ReLU to the output                              */
  int32_t pixel_shuffle;  /* fuse PixelShuffle(2) into the store                   */
  int32_t has_bias;
} PlcConvDesc;
size_t plc_conv_packed_weight_bytes(const PlcConvDesc* d, int pack_kind);
int plc_conv_pack_weight(const PlcConvDesc* d, int pack_kind, const float* w_oihw, const float* bias, void* w_packed,
                         float* bias_packed, void* stream);
int plc_conv_fwd(const PlcConvDesc* d, const void* x, const void* w_packed_fwd, const float* bias_packed, void* out,
                 void* stream);
/* plc_conv_fwd with an fp32 output tensor [B,H,W,Cout] (no PixelShuffle): for the LAST layer of a model
 * (generator.py:67-71 post_process[2], 32 -> 1), so the predicted rain keeps the fp32 accumulator instead of being
 * rounded to 8 mantissa bits before the loss.                                                                    */
int plc_conv_fwd_f32(const PlcConvDesc* d, const void* x, const void* w_packed_fwd, const float* bias_packed, float* out,
                     void* stream);
int plc_conv_grad_mask(const PlcConvDesc* d, const void* y, const void* dy, void* dz, void* stream);
size_t plc_conv_wgrad_acc_bytes(const PlcConvDesc* d);
int plc_conv_wgrad_unpack(const PlcConvDesc* d, const float* dW_acc, float* dW_oihw, void* stream);
int plc_conv_bwd(const PlcConvDesc* d, const void* x, const void* dz, const void* w_packed_dgrad, void* dx,
                 float* dW_acc, float* db_acc, void* stream);
/* Weight gradient of a NARROW-input layer (the front-end init_conv of generator.py:50-55,166-168: 1 rain + 2 coordinate
 * channels): with cin_true * k * k <= 32 the whole receptive field of a pixel fits ONE 32-column im2col row, and the
 * reduction over pixels becomes a plain [Cout x 32] GEMM instead of k*k column blocks that are 5/8 zero padding.
 * plc_conv_im2col_narrow: x [B,H,W,8] bf16 (channels >= cin_true are zero) -> col [B,H,W,32] bf16,
 * col[pixel][tap * cin_true + c] = x[pixel + tap offset][c], zero outside the image.  The caller then runs
 * plc_conv_bwd (dx = NULL) with a k = 1, Cin = 32 descriptor on `col` and maps dW1[Cout][tap * cin_true + c] back
 * to [Cout][c][ky][kx].  `d` describes the ORIGINAL layer (Cin = 8, k).                                          */
int plc_conv_im2col_narrow(const PlcConvDesc* d, int cin_true, const void* x, void* col, void* stream);

/* ---- strided 2-D / 3-D convolution on the same core (bf16; the discriminator of the GAN training step) ----------
 * No reference counterpart (the reference has no discriminator: SURVEY.md section 0); the eager spec is
 * oracle/gan_oracle.py (torch F.conv2d / F.conv3d).  North-star: "the discriminator's strided 2D/3D convolutions reuse
 * the same implicit-GEMM core" -- these entry points run conv_igemm_tc_kernel with strided (elementStrides) and, for a
 * time kernel, 5-D [C, W, H, T, B] tensor maps; TMA zero fill is the zero padding in all three dimensions.
 *   x   [B, T, H, W, Cin]  bf16 (T = 1, kt = 1, stride_t = 1: a 2-D conv over B images)
 *   w   [Cout, Cin, kt, k, k] fp32 (torch Conv3d layout; Conv2d layout when kt = 1), bias [Cout]
 *   out [B, To, Ho, Wo, Cout] bf16, To = (T-1)/stride_t + 1, Ho = (H-1)/stride + 1 (padding kt/2, k/2, k/2)
 *   act: 0 none, 1 ReLU, 2 LeakyReLU(slope), applied after the bias.
 * Backward: plc_convnd_grad_mask forms dZ = dY * act'(Y) (layers with an activation; otherwise dZ == dY).
 * plc_convnd_bwd: dx [B,T,H,W,Cin] (nullable).  For a strided layer the transposed conv is decomposed by output phase
 * (parity class of the dX position in every strided dimension): each phase is a stride-1 conv of dZ over its own tap
 * subset that writes its sub-lattice of dX -- no zero-inserted copy of dZ, exactly the layer's MMAs in total; the
 * PLC_PACK_DGRAD image holds one packed weight image per phase.  dW_acc = accumulator image
 * (plc_convnd_wgrad_acc_bytes; plc_convnd_wgrad_unpack ADDS it into [Cout,Cin,kt,k,k]); db_acc [Cout] += (nullable). */
typedef struct PlcConvNdDesc {
  int32_t B, T, H, W;          /* input grid                                   */
  int32_t Cin, Cout;           /* multiples of 8 (callers zero-pad)            */
  int32_t kt, k;               /* odd kernel sizes (time, space), <= 7         */
  int32_t stride_t, stride;    /* 1 or 2                                       */
  int32_t act;                 /* 0 none, 1 ReLU, 2 LeakyReLU                  */
  float slope;                 /* LeakyReLU negative slope                     */
  int32_t has_bias;
} PlcConvNdDesc;
int plc_convnd_out_shape(const PlcConvNdDesc* d, int* T_out, int* H_out, int* W_out);
size_t plc_convnd_packed_weight_bytes(const PlcConvNdDesc* d, int pack_kind);
int plc_convnd_pack_weight(const PlcConvNdDesc* d, int pack_kind, const float* w, const float* bias, void* w_packed,
                           float* bias_packed, void* stream);
int plc_convnd_fwd(const PlcConvNdDesc* d, const void* x, const void* w_packed_fwd, const float* bias_packed, void* out,
                   void* stream);
int plc_convnd_grad_mask(const PlcConvNdDesc* d, const void* y, const void* dy, void* dz, void* stream);
size_t plc_convnd_wgrad_acc_bytes(const PlcConvNdDesc* d);
int plc_convnd_wgrad_unpack(const PlcConvNdDesc* d, const float* dW_acc, float* dW, void* stream);
int plc_convnd_bwd(const PlcConvNdDesc* d, const void* x, const void* dz, const void* w_packed_dgrad, void* dx,
                   float* dW_acc, float* db_acc, void* stream);

/* ---- frame-level first layer (discriminator conv1): 3x3 conv straight from fp32 frames --------------------------
 * out[n, yo, xo, co] = act(bias[co] + sum_{ci, ky, kx} w[co, ci, ky, kx] * frames[n, ci, yo*s + ky - 1, xo*s + kx - 1])
 *   frames [N, Cf, H, W] fp32 (Cf = 1..4; a clip [B, T, Cf, H, W] is N = B*T frames), zero padding 1, stride s = 1 or 2
 *   w [Cout, Cf, 3, 3] fp32 (torch layout, unpacked), bias [Cout]; Cout in {8, 16, 32, 64, 128, 256}
 *   out / y / dy [N, Ho, Wo, Cout] bf16, Ho = (H-1)/s + 1;  act: 0 none, 1 ReLU, 2 LeakyReLU(slope)
 * With one rain channel the layer is 9 multiply-adds per output value and HBM-bound: SIMT kernels that read the frames
 * where they lie replace the implicit-GEMM path for it (csrc/frame_conv.cuh has the numbers).  Eager spec:
 * oracle/gan_oracle.py conv1.  plc_frameconv_bwd takes the forward output y and dY and applies act'(y) on the fly:
 * dframes [N, Cf, H, W] fp32 is overwritten (nullable), dW [Cout, Cf, 3, 3] and db [Cout] are ACCUMULATED (nullable). */
typedef struct PlcFrameConvDesc {
  int32_t N, Cf, H, W, Cout, stride, act;
  float slope;
  int32_t has_bias;
} PlcFrameConvDesc;
int plc_frameconv_out_shape(const PlcFrameConvDesc* d, int* H_out, int* W_out);
int plc_frameconv_fwd(const PlcFrameConvDesc* d, const float* frames, const float* w_oihw, const float* bias, void* out,
                      void* stream);
int plc_frameconv_bwd(const PlcFrameConvDesc* d, const float* frames, const float* w_oihw, const void* y, const void* dy,
                      float* dframes, float* dW, float* db, void* stream);

/* ---- per-launch timing (bench.py's roofline / step breakdown) -------------------------------------
 * plc_timing_enable(1) clears the record and makes every kernel launch of the library record a CUDA event pair on
 * its launching stream, tagged with a PlcKernelKind; plc_timing_enable(0) stops recording (and clears).
 * plc_timing_collect synchronises the recorded events, writes up to `capacity` (kind, milliseconds, algorithmic FLOPs)
 * triples in launch order, clears the record and returns the number of launches recorded (which may exceed
 * `capacity`).  FLOPs = 2*M*N*K over the TRUE (unpadded) channel counts for the contractions, 0 for other kernels
 * (`flops` may be NULL).  A kind groups
 * the launches one ABI call makes for that purpose (normally exactly one kernel).  Never enable it inside a region
 * whose time is reported: the events cost host time and serialise nothing but are not free.                    */
typedef enum PlcKernelKind {
  PLC_K_CELL_FWD = 0,       /* fused cell step, full K loop (x and h taps)                                    */
  PLC_K_CELL_FWD_ZERO = 1,  /* fused cell step, zero-initial-state form (x taps only)                          */
  PLC_K_BWD_GATES = 2,      /* gate recompute + dZ / dc_prev                                                   */
  PLC_K_BWD_DGRAD = 3,      /* dx, dh_prev                                                                     */
  PLC_K_BWD_WGRAD = 4,      /* dW, db                                                                          */
  PLC_K_CONV_FWD = 5,       /* plain / strided / 3-D conv forward (generator body, discriminator)             */
  PLC_K_CONV_DGRAD = 6,
  PLC_K_CONV_WGRAD = 7,
  PLC_K_FRONTEND = 8,
  PLC_K_HEAD = 9,
  PLC_K_LOSS = 10,
  PLC_K_PACK = 11,          /* weight packing                                                                  */
  PLC_K_ELEMENTWISE = 12    /* layout conversions, gradient masks, accumulator unpack                          */
} PlcKernelKind;
int plc_timing_enable(int on);
int plc_timing_collect(int* kinds, float* ms, double* flops, int capacity);

/* ---- debug ---------------------------------------------------------------------------------
 * Developer aid (tools/kprof.py): when set to a zeroed device buffer of at least 148*16 uint64, the tensor-core conv
 * kernels record per-CTA cycle counters (MMA warp total / waiting for TMEM / waiting for TMA, epilogue busy / idle).
 * NULL (default) disables it.  Not part of the reference-facing surface.                                   */
int plc_debug_set_prof(void* device_buf_u64);
/* Force the tensor-core kernels onto cta_group 1 or 2 (0 = automatic choice by problem size); used by the parity
 * tests to cover the CTA-pair path on small shapes.                                                          */
int plc_debug_set_cta_group(int cta_group);
/* Haloed-patch pipeline of the tensor-core conv kernels (one activation patch per tile serves all k*k taps as shifted
 * UMMA views): -1 = automatic (default), 0 = never, 1 = whenever the kernel supports it.  Parity tests run both.   */
int plc_debug_set_patch(int mode);

/* ---- layout helpers (HBM-bound elementwise kernels) ---------------------------------------
 * The reference keeps NCHW fp32 tensors (generator.py:156-160).  These convert between that and
 * the NHWC working layout at sequence entry / exit.  `C_dst >= C_src` zero-pads channels.       */
int plc_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int B, int C_src, int C_dst, int H, int W,
                              void* stream);
int plc_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int B, int C, int H, int W, void* stream);

/* ---- frame front-end / head (either side of the recurrence; SURVEY.md section 8f "next-1") -------
 * plc_frontend_fwd replaces generator.py:166-168 + coordconv.py:3-10 per step:
 *     x_t = relu(init_conv(add_coord_channels(frame_t)))
 *   frames [N, Cf, H, W] fp32 NCHW (N = B or B*T); w_oihw = init_conv.weight [C, Cf+2, 3, 3]; bias [C] or NULL
 *   out    [N, H, W, C_stride] NHWC, channels [0,C) written (bf16 in PLC_MODE_BF16_TC, fp32 in PLC_MODE_FP32)
 * plc_head_fwd: 1x1 conv C -> 1 on the top layer's h (encoder-forecaster output head, repo-defined):
 *   h [npix, C] NHWC (mode dtype), w [C] fp32, bias [1] or NULL, out [npix] fp32.                   */
int plc_frontend_fwd(const float* frames, int N, int Cf, int H, int W, const float* w_oihw, const float* bias, int C,
                     int C_stride, int mode, void* out, void* stream);
int plc_head_fwd(const void* h, long npix, int C, const float* w, const float* bias, int mode, float* out,
                 void* stream);

/* The same front-end (generator.py:166-168 + coordconv.py:3-10, all T frames of a batch in one launch) on the tensor
 * cores with the im2col done by the kernel's producer warps: frames [B,T,Cf,H,W] fp32 -> out [T*B,H,W,64] bf16 =
 * relu(conv3x3(cat(frame, row/(H-1), col/(W-1)))) with w_oihw = init_conv.weight [64,Cf+2,3,3] fp32 (read directly, no
 * packing step), bias [64] or NULL.  Available for 1..3 frame channels and C = C_stride = 64
 * (plc_frontend_tc_supported); other shapes use plc_frames_to_nhwc + plc_conv_fwd.                              */
int plc_frontend_tc_supported(int Cf, int C, int C_stride);
int plc_frontend_tc_fwd(const float* frames, int B, int T, int Cf, int H, int W, const float* w_oihw, const float* bias,
                        int C, void* out, void* stream);

/* frames [B,T,Cf,H,W] fp32 (layout of the reference's rain_lr, generator.py:96) -> [T*B,H,W,Cp] bf16, T-major, with
 * the coordinate planes of add_coord_channels (coordconv.py:3-10) appended and zero padding to Cp (Cp % 8 == 0,
 * Cp >= Cf+2): the NHWC input of init_conv when it runs on the tensor-core conv (plc_conv_fwd).                */
int plc_frames_to_nhwc(const float* frames, int B, int T, int Cf, int H, int W, int Cp, void* out, void* stream);

/* backward of plc_head_fwd (bf16 mode): dh [npix, C] bf16 = dy * w;  dw_acc [C] += sum dy * h;  db_acc [1] += sum dy
 * (db_acc may be NULL).  dy [npix] fp32.                                                                       */
int plc_head_bwd(const void* h, long npix, int C, const float* w, const float* dy, void* dh, float* dw_acc,
                 float* db_acc, void* stream);

/* ---- CombinedLoss (SURVEY.md section 8f "next-2") ---------------------------------------------
 * plc_combined_loss replaces CombinedLoss.forward (src/losses/combined_loss.py:173-191) and its autograd:
 * station point supervision (:79-141), conservation by area pooling (:64-74), spatial gradient (:146-155), temporal
 * consistency (:160-168), in fp32.
 *   pred      [B,T,1,H*scale,W*scale] fp32      lr_input [B,T,1,H,W] fp32
 *   s_coords  [N,2] int64 (row, col) on the LR grid; mapped to HR pixels with ((c + 0.5) * coord_scale - 0.5) truncated
 *   s_values  [T,N] (svals_has_batch = 0) or [B,T,N] fp32, NaN = no observation
 *   terms_out [5] fp32 DEVICE: total, point, conserve, smooth, temporal          (no host synchronisation)
 *   dpred     [B,T,1,H*scale,W*scale] fp32 = grad_scale * d total / d pred, or NULL for loss values only
 *   grad_scale DEVICE scalar (the upstream gradient of `total`), or NULL for 1
 *   workspace plc_loss_workspace_bytes(d) bytes, 16-byte aligned
 * Only integer upsampling ratios are supported (the Generator only produces those).                          */
typedef struct PlcLossDesc {
  int32_t B, T, H, W;       /* LR grid                                                   */
  int32_t scale;            /* HR = LR * scale                                           */
  int32_t n_stations;
  int32_t svals_has_batch;
  int32_t weight_mode;      /* 0 off, 1 log, 2 sqrt, 3 stratified (combined_loss.py:22-59) */
  float coord_scale;        /* scale_factor argument of CombinedLoss.forward             */
  float lambda_point, lambda_conserve, lambda_smooth, lambda_temporal;
} PlcLossDesc;
size_t plc_loss_workspace_bytes(const PlcLossDesc* d);
int plc_combined_loss(const PlcLossDesc* d, const float* pred, const float* lr_input, const long long* s_coords,
                      const float* s_values, void* workspace, float* terms_out, float* dpred, const float* grad_scale,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PLC_H_ */
