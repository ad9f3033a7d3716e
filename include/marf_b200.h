/*
 * marf_b200.h — C ABI of the B200-native planar bundle-adjusting training step.
 *
 * This is the drop-in boundary for ONE path of the reference
 * (thomasjaron/masking-bundle-adjusting-neural-radiance-fields): what
 * `Model.train_iteration` executes between `optim.zero_grad()` and `optim.step()`
 * (model/planar.py:192-196), i.e. `Graph.forward` (model/planar.py:329-353),
 * `Graph.compute_loss` (:355-380), `Model.summarize_loss` (:172-185) and the autograd
 * backward of all of it, plus the forward-only render used by
 * `Model.predict_entire_image` (:211-217).
 *
 * The reference has no FFI of its own (it is eager PyTorch); the binding a maintainer adds
 * is the ctypes stub shown in INTEGRATION.md (and shipped as marf_b200/_lib.py).
 *
 * Conventions
 *   - plain C types only; every pointer in the *_io structs is a DEVICE pointer owned by the
 *     caller (they are torch.Tensor.data_ptr()s in the Python host) unless marked "host";
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*); no entry
 *     point synchronises the device or the stream except marf_create / marf_destroy and the three
 *     diagnostics that say so (marf_tc_selftest, marf_tf32_gemm, marf_debug_read_bf16, marf_profile_read), so every
 *     step / render / optimizer call can be captured into a CUDA graph;
 *   - every entry point returns 0 on success, else a marf_status / cudaError_t value;
 *     marf_last_error() returns a human-readable message for the last failure on that handle;
 *   - nothing throws across the ABI; asynchronous CUDA faults surface at the next call;
 *   - one handle per device/rank, not thread-safe;
 *   - there is NO CPU fallback: without a CUDA device marf_create fails.
 */
#ifndef MARF_B200_H_
#define MARF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MARF_ABI_VERSION 2
#define MARF_MAX_LAYERS 12

typedef struct marf_handle marf_handle;

enum marf_status {
  MARF_OK = 0,
  MARF_ERR_INVALID = -1,      /* bad argument / unsupported configuration           */
  MARF_ERR_CUDA = -2,         /* a CUDA runtime call failed (see marf_last_error)     */
  MARF_ERR_NO_DEVICE = -3,    /* no sm_100 device: the product has no CPU path       */
  MARF_ERR_UNSUPPORTED = -4   /* valid reference option this build does not implement */
};

enum marf_mask_mode {
  MARF_MASK_NONE = 0,         /* use_masks=False: mean((p-l)^2)                     model/planar.py:384-386 */
  MARF_MASK_DISK = 1,         /* use_masks=True:  sum(((p-l)m)^2)/(3 sum m)         model/planar.py:388-390 */
  MARF_MASK_IMPLICIT = 2      /* use_implicit_mask=True: m = mask head prediction   model/planar.py:339-352 */
};

enum marf_precision {
  MARF_FP32 = 0,              /* fp32 CUDA-core arithmetic end to end (parity mode, <=1e-3 vs reference)   */
  MARF_BF16 = 1               /* bf16 operands on tcgen05 tensor cores, fp32 accumulate in TMEM (perf mode) */
};

/* Static description of the job; mirrors the options/planar.yaml keys the path consumes. */
typedef struct marf_config {
  int32_t abi_version;        /* MARF_ABI_VERSION */
  int32_t device;             /* CUDA ordinal */
  int32_t precision;          /* enum marf_precision */
  /* geometry: warp.py:9-21 */
  int32_t H, W;               /* canvas (opt.H, opt.W) */
  int32_t patch_H, patch_W;   /* opt.patch_H, opt.patch_W */
  int32_t use_cropped;        /* opt.use_cropped_images: 1 -> grid is the centre crop, 0 -> full canvas */
  /* sharding: this handle owns `batch` patches starting at global patch `patch_offset`, and of each
   * of them the rows [row_offset, row_offset+rows) of the h-row patch grid.  Single GPU: batch =
   * batch_global, patch_offset = 0, row_offset = 0, rows = h. */
  int32_t batch_global;       /* opt.batch_size */
  int32_t batch, patch_offset;
  int32_t rows, row_offset;
  /* neural image: model/planar.py:410-471 */
  int32_t L;                  /* arch.posenc.L_2D, 0 = `--arch.posenc!` */
  int32_t n_layers;           /* number of Linear layers */
  int32_t layer_out[MARF_MAX_LAYERS];  /* k_out per layer (arch.layers[1:]) */
  uint32_t skip_mask;         /* bit li set <=> li in arch.skip (input concat) */
  int32_t c2f_enabled;        /* barf_c2f is not None */
  float c2f_start, c2f_end;
  /* masks / mask head: model/planar.py:319-327,475-518 */
  int32_t mask_mode;          /* enum marf_mask_mode */
  int32_t mask_n_layers;      /* 5 */
  int32_t mask_layer_out[MARF_MAX_LAYERS];  /* 256,256,256,256,1 */
  int32_t mask_uv_freqs;      /* 10  (PosEmbedding(9,10)) */
  int32_t mask_embed_dim;     /* 128 (embedding_view) */
  /* edge branch: inputs.py:50-69, model/planar.py:336,366-369 */
  int32_t use_edges;
  int32_t edge_label_channels;/* channels of images.edges (1: computed from the grey image) */
  /* workspace policy */
  int64_t max_chunk_pixels;   /* 0 = library default; pixel-samples processed per pass */
  /* (ABI 2) rows of embedding_view (opt.N_vocab, model/planar.py:327): colour indices trunc(rgb) outside
   * [0, mask_n_vocab) raise IndexError in the reference; here they are clamped and COUNTED in
   * loss_sums[MARF_BAD_INDEX] (bf16 mode: the class table serves indices {0,1} only, anything else counts). 0 = 1500. */
  int32_t mask_n_vocab;
  int32_t reserved0;
} marf_config;

/* Per-step inputs/outputs.  Shapes use the LOCAL shard: n = batch*rows*w pixel-samples. */
typedef struct marf_step_io {
  /* ---- parameters (fp32, row-major, torch layouts) */
  const float* const* mlp_w;  /* host array[n_layers] of device ptrs, W_l [k_out,k_in]   model/planar.py:421 */
  const float* const* mlp_b;  /* host array[n_layers] of device ptrs, b_l [k_out] */
  const float* warp;          /* [batch_global,8] sl(3) parameters                        model/planar.py:310 */
  const float* const* mask_w; /* host array[mask_n_layers] or NULL                        model/planar.py:480-484 */
  const float* const* mask_b;
  const float* embed;         /* embedding_view.weight [n_vocab,mask_embed_dim] or NULL   model/planar.py:327 */
  /* ---- data (resident; re-read every step) */
  const float* rgb;           /* [batch,3,rows,w] targets in [0,1]                        inputs.py:26-33 */
  const float* masks;         /* [batch,1,rows,w] validity (1=valid) or NULL              inputs.py:119 */
  const float* masks_eroded;  /* [batch,1,rows,w] or NULL (edge loss, disk-mask mode)     inputs.py:120 */
  const double* edges;        /* [batch,edge_label_channels,rows,w] float64 or NULL       inputs.py:125 */
  int64_t data_version;       /* bump when rgb/masks/embed CONTENTS change (cached derived inputs) */
  /* ---- scalars */
  float progress;             /* neural_image.progress (c2f schedule)                     model/planar.py:208,465 */
  float c_rgb, c_mask, c_edge;/* d(loss.all)/d(rgb|mask|edge loss): summarize_loss x render expansion */
  double norm_rgb;            /* host: GLOBAL normaliser of the rgb loss (3*sum m, or 3*N without masks);
                                 0 -> derive from the local shard (single-rank use) */
  double norm_edge;           /* host: same for the edge loss; 0 -> local */
  /* ---- outputs */
  float* const* g_mlp_w;      /* host array of device ptrs; OVERWRITTEN with d(all)/dW_l (local shard's share) */
  float* const* g_mlp_b;
  float* g_warp;              /* [batch_global,8]; rows of patches this shard does not own are zeroed */
  float* const* g_mask_w;     /* or NULL */
  float* const* g_mask_b;
  float* rgb_pred;            /* optional [batch,rows*w,3]  (var.rgb_prediction)          model/planar.py:334 */
  float* mask_pred;           /* optional [batch,rows*w,1]  (var.mask_prediction)         model/planar.py:351 */
  double* edge_pred;          /* optional [batch,3,rows,w] float64 (var.edge_prediction)  model/planar.py:336 */
  double* loss_sums;          /* device double[MARF_N_SUMS], see below */
} marf_step_io;

/* loss_sums slots (all LOCAL partial sums after marf_step_forward; the caller may all-reduce
 * the buffer before marf_step_backward; losses are then
 *   rgb  = S_RGB / N_RGB        (N_RGB = 3*sum m, or 3*N without masks)
 *   mask = S_MASK / N_MASK      (N_MASK = N pixel-samples)
 *   edge = S_EDGE / N_EDGE ) */
enum marf_sum_slot {
  MARF_S_RGB = 0, MARF_N_RGB = 1, MARF_S_MASK = 2, MARF_N_MASK = 3,
  MARF_S_EDGE = 4, MARF_N_EDGE = 5, MARF_NONFINITE = 6, MARF_BAD_INDEX = 7, MARF_N_SUMS = 8
};

/* Forward-only render of the neural image (Model.predict_entire_image, model/planar.py:211-217,
 * and the <=1e-3 per-pixel output parity check). */
typedef struct marf_render_io {
  const float* const* mlp_w;
  const float* const* mlp_b;
  const float* warp;          /* [n_patches,8] or NULL (identity: un-warped grid) */
  int32_t n_patches;          /* 1 for predict_entire_image */
  int32_t crop;               /* get_normalized_pixel_grid(crop=...)  warp.py:33 */
  float progress;
  float* rgb;                 /* out [n_patches, P, 3], P = patch_H*patch_W (crop) or H*W */
} marf_render_io;

int marf_abi_version(void);
/* Allocates the handle and its workspace on cfg->device.  Fails (MARF_ERR_NO_DEVICE) without a GPU. */
int marf_create(const marf_config* cfg, marf_handle** out);
int marf_destroy(marf_handle* h);
const char* marf_last_error(const marf_handle* h);     /* h may be NULL: last create failure */

/* forward + loss + backward in one call (single rank, or multi-rank with static normalisers). */
int marf_step(marf_handle* h, const marf_step_io* io, void* stream);
/* two-phase variant for a global normaliser that depends on the forward pass (implicit masks, N>1):
 * forward writes local loss_sums; the caller all-reduces them on `stream`; backward consumes them. */
int marf_step_forward(marf_handle* h, const marf_step_io* io, void* stream);
int marf_step_backward(marf_handle* h, const marf_step_io* io, void* stream);

int marf_render(marf_handle* h, const marf_render_io* io, void* stream);

/* NeuralImageFunction.forward(coord_2d) for caller-supplied coordinates (model/planar.py:429-449): xy [n,2] are the
 * ALREADY WARPED normalised coordinates (what Graph.forward passes at model/planar.py:334); positional encoding with the
 * c2f weights of `progress`, the MLP and the sigmoid run on device; rgb [n,3].  Any n >= 1 (processed in passes). */
int marf_forward_points(marf_handle* h, const float* const* mlp_w, const float* const* mlp_b, const float* xy, int64_t n,
                        float progress, float* rgb, void* stream);

/* geometry helpers on device (warp.py:83-108): H = expm(A(h)) and the warped crop corners. */
int marf_sl3_to_SL3(marf_handle* h, const float* warp, int32_t n, float* out9, void* stream);
int marf_warp_corners(marf_handle* h, const float* warp, int32_t n, float* out_n_4_2, void* stream);
/* Warp.warp_grid (warp.py:70-81) for arbitrary points: xy [n,p,2], warp [n,8] -> out [n,p,2]. */
int marf_warp_points(marf_handle* h, const float* xy, const float* warp, int32_t n, int32_t p, float* out, void* stream);

/* edge map of a [n,c,rows,w] fp32 image batch -> float64 (inputs.compute_edges, inputs.py:50-69). */
int marf_compute_edges(marf_handle* h, const float* images, int32_t n, int32_t c, int32_t rows, int32_t w,
                       double* out, void* stream);

/* Device-side Adam over a list of tensors in one launch (torch.optim.Adam defaults: no weight decay, no amsgrad;
 * model/planar.py:98-99,197), plus the `warp.fix_first` reset of the loop tail (model/planar.py:157-158):
 * after the update the first `zero_count` elements of tensor `zero_tensor` are set to 0 (moments keep evolving,
 * as in the reference).  `step` is the 1-based step count used for the bias corrections. */
typedef struct marf_adam_io {
  int32_t n_tensors;
  float* const* params;       /* host array of device ptrs */
  const float* const* grads;
  float* const* exp_avg;
  float* const* exp_avg_sq;
  const int64_t* numel;       /* host array */
  const float* lr;            /* host array, per tensor */
  float beta1, beta2, eps;
  int64_t step;
  int32_t zero_tensor;        /* -1: none */
  int32_t zero_count;
} marf_adam_io;
int marf_adam_step(marf_handle* h, const marf_adam_io* io, void* stream);

/* Loss scalars of the last step from its device-side sums, one launch (replaces the dozen 0-dim tensor ops of
 * Graph.compute_loss model/planar.py:362-378 + Model.summarize_loss :172-185): out5 = {rgb, mask, edge, render, all} with
 * render = (1-alpha) rgb + 0.5 mask + alpha edge and all = w[0] render + w[1] rgb + w[2] mask + w[3] edge
 * (w = 10**loss_weight, 0 for a disabled term; host array).  sums/out5 are device pointers. */
int marf_loss_scalars(marf_handle* h, const double* sums, double alpha, const double* weights4, double* out5, void* stream);

/* One-shot all-reduce (sum) over NVLink peer memory for the small exchanges of a data-parallel step (loss sums, gradient
 * buffer): every rank passes the device pointers of all ranks' inputs and flag arrays (symmetric allocations, e.g.
 * torch.distributed._symmetric_memory; flag array = 2*world+1 zero-initialised uint32), its rank, and a round number `seq`
 * that increases by one per call of the group.  dtype 0 = fp32, 1 = fp64.  out == peer_in[rank] (in place) is allowed for
 * n <= 256.  When the launch has completed no peer reads this rank's input any more.  Replaces the NCCL all-reduce the
 * host side would otherwise issue (`dist.all_reduce` in marf_b200/planar.py); no handle needed. */
int marf_peer_allreduce(int device, int dtype, const void* const* peer_in, uint32_t* const* peer_flags, int rank, int world,
                        void* out, long long n, uint32_t seq, void* stream);

/* diagnostic: run ONE tensor-core kernel (tcgen05) on fp32 device arrays that are rounded to bf16 inside.
 * mode 0: relu(A[rows,K] W[N,K]^T + aux[N]); 1: (A W^T)*(aux[rows,N]>0); 2: plain fp32 out, N=64;
 * 3: out[N,K] = A[rows,N]^T aux[rows,K] (the dW kernel).  Synchronises `stream`.  Used by tests/ only. */
int marf_tc_selftest(int device, int mode, int rows, int K, int N, const float* A, const float* W, const float* aux,
                     float* out, void* stream);

/* diagnostic: run ONE 3xTF32 tensor-core GEMM of the precision=fp32 path (csrc/tc_tf32.cuh) on fp32 device arrays.
 * mode 0: C[M,N] = epi(A[M,K] W[N,K]^T)   (forward layer; W row-major [N, ldw])
 * mode 1: C[M,N] = epi(A[M,K] W[K,N])     (dX layer;      W row-major [K, ldw])
 * mode 2: C[N,K] += A[M,N]^T W[M,K]       (dW layer;      W = the second activation matrix [M, ldw]; aux, if given, [N] += column sums of A)
 * epi (modes 0/1) 0: + aux[N]; 1: relu(+ aux[N]); 2: plain; 3: zero where aux[M, ldaux] <= 0.
 * M must be a multiple of 128, every ld a multiple of 4.  Synchronises `stream`.  Used by tests/ only. */
int marf_tf32_gemm(marf_handle* h, int mode, int epi, int M, int N, int K, const float* A, int lda, const float* W, int ldw,
                   float* C, int ldc, const float* aux, int ldaux, void* stream);

/* diagnostic: fp32 copy of a resident bf16 buffer of the bf16 path's last chunk (which 0: input of `layer`, [rows, ld];
 * 1: gradient w.r.t. the output of `layer`, [rows, 256]).  Synchronises `stream`.  Used by tests/ and profiles/tools only. */
int marf_debug_read_bf16(marf_handle* h, int chain, int which, int layer, float* out, long long rows, void* stream);

/* Per-kernel timing for bench.py's roofline line.  While enabled, the bf16 path records a CUDA event pair on the launching
 * stream around every launch of its five tensor-core kernel classes (this suspends programmatic dependent launch across
 * those launches, so enable it for a separate measurement pass, not for the headline timing).  Classes: the fused forward
 * chain, the fused backward launch (dX chains + every dW GEMM; or, with MARF_NO_BWD_FUSE, the dX chain and the merged dW launch
 * separately), the dX0 + encoding-backward GEMM.  marf_profile_read
 * synchronises the recorded events, writes per class the summed milliseconds and the launch count, and clears them. */
#define MARF_PROF_CLASSES 5
enum { MARF_PROF_CHAIN_FWD = 0, MARF_PROF_CHAIN_DX = 1, MARF_PROF_DW = 2, MARF_PROF_BWD = 3, MARF_PROF_DX0 = 4 };
int marf_profile(marf_handle* h, int enable);
int marf_profile_read(marf_handle* h, double* ms, int64_t* launches, int n_classes);

/* bookkeeping for bench.py / tests: kernels launched by this handle since creation, bytes of workspace. */
int64_t marf_launch_count(const marf_handle* h);
int64_t marf_workspace_bytes(const marf_handle* h);

#ifdef __cplusplus
}
#endif
#endif  /* MARF_B200_H_ */
