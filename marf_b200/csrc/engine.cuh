// Host-side engine behind the C ABI: workspace, chunking, kernel sequencing.
#pragma once
#include "common.cuh"

namespace marf {

struct Chain {                 // one MLP in padded fp32 workspace form
  int n = 0;
  int k_in[MARF_MAX_LAYERS], k_out[MARF_MAX_LAYERS];
  int ld_in[MARF_MAX_LAYERS], ld_out[MARF_MAX_LAYERS];
  uint32_t skip_mask = 0;
  int d_in = 0;                // width of the encoded input (for skip concat)
  float* Wp[MARF_MAX_LAYERS];  // [ld_out, ld_in] zero padded
  float* bp[MARF_MAX_LAYERS];  // [ld_out]
  float* gWp[MARF_MAX_LAYERS];
  float* gbp[MARF_MAX_LAYERS];
  float* act[MARF_MAX_LAYERS + 1];  // act[l]: input of layer l [chunk, ld_in[l]]; act[n]: logits [chunk, ld_out[n-1]]
  bool need_dx0 = false;       // image MLP: yes (warp gradient); mask head: no
  // 3xTF32 tensor-core path (tc_tf32.cuh): the weights split into big / small tf32 planes, [2 * pad16(N), pad32(K)], rebuilt by
  // pack_chain every step: Wsp_f = W (forward: N = out, K = in), Wt = W^T (dX: N = in, K = out); nullptr for layers that stay
  // on k_sgemm (N or K < 32)
  float* Wsp_f[MARF_MAX_LAYERS] = {};
  float* Wt[MARF_MAX_LAYERS] = {};
  // sign bits of the input of layer l (= ReLU output of layer l - 1), [rows, bits_ld[l]] words, written by the tensor-core
  // forward layer and read by the dX layer instead of the fp32 input; bits_ok[l]: written by the last forward of this chunk
  uint32_t* bits[MARF_MAX_LAYERS] = {};
  int bits_ld[MARF_MAX_LAYERS] = {};
  bool bits_ok[MARF_MAX_LAYERS] = {};
  int max_ld = 0;
};

struct Bf16State;              // tcgen05 path (bf16_path.cu)

}  // namespace marf

struct marf_handle {
  marf_config cfg;
  marf::Geo geo;
  int h, w;                    // traversed patch grid
  long long n_local;           // batch*rows*w
  int chunk;                   // padded pixel-samples per pass (multiple of 128)
  int n_chunks;
  int render_rows;             // rows of the fp32 activation buffers (forward-only render path)
  marf::Chain img, msk;
  float* Hm = nullptr;         // [batch_global,9]
  double* G = nullptr;         // [batch,9]
  float* dYa = nullptr;        // ping-pong gradient buffers [chunk, max_ld]
  float* dYb = nullptr;
  float* dX0acc = nullptr;     // skip connections: accumulated gradient wrt the encoded input
  float* dXscratch = nullptr;  // skip connections: raw dX of a concat layer
  float* pred_rgb = nullptr;   // [n_local,3] when edges are on (or caller did not pass rgb_pred)
  float* pred_mask = nullptr;  // [n_local]
  double* edge_mag = nullptr;  // [batch,3,rows,w]
  double* edge_pred = nullptr;
  double* sums_static = nullptr;  // [2]: 3*sum(masks), 3*sum(masks_eroded) of the local shard
  marf::LossCoef* coef = nullptr;
  double* bad_index = nullptr;    // sticky per data version: colour indices outside the embedding table (-> loss_sums[MARF_BAD_INDEX])
  int64_t data_version_seen = INT64_MIN;
  bool feats_valid = false;    // mask-head input features cached in msk.act[0] (single chunk only)
  bool acts_valid = false;     // forward activations of the (single) chunk are resident for backward
  marf::Bf16State* bf16 = nullptr;
  bool fp32_tc = true;         // precision=fp32: wide layers as 3xTF32 on the tensor cores (MARF_FP32_TC=0: CUDA-core SGEMMs only)
  int n_sms = 148;
  int64_t launches = 0;
  int64_t ws_bytes = 0;
  // per-kernel-class timing (marf_profile): CUDA event pairs recorded on the launching stream around the tensor-core launches
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev[MARF_PROF_CLASSES];
  std::vector<cudaEvent_t> prof_pool;
  std::vector<void*> allocs;
  std::string err;
};

namespace marf {
// implemented in api.cu
int fail(marf_handle* h, int code, const std::string& msg);
void* ws_alloc(marf_handle* h, size_t bytes, bool zero = true);
void set_schedule(marf_handle* h, float progress);
// implemented in bf16_path.cu
int bf16_create(marf_handle* h);
void bf16_destroy(marf_handle* h);
int bf16_step(marf_handle* h, const marf_step_io* io, cudaStream_t st);
int bf16_forward(marf_handle* h, const marf_step_io* io, cudaStream_t st);
int bf16_backward(marf_handle* h, const marf_step_io* io, cudaStream_t st);
bool bf16_supported(const marf_handle* h, const marf_step_io* io, std::string* why);
}  // namespace marf
