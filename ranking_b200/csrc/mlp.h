// Host-side plan of the scorer tower: parameter offsets and workspace carving.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "tfr_b200.h"

namespace tfr {

struct MlpPlan {
  int n_dense;
  int dims[TFR_MLP_MAX_LAYERS + 1];
  int activation;
  // flat parameter layout (floats)
  size_t w_off[TFR_MLP_MAX_LAYERS], b_off[TFR_MLP_MAX_LAYERS], n_params;
  // workspace layout (floats)
  size_t act_off[TFR_MLP_MAX_LAYERS];  // post-activation output of hidden layer d
  size_t dz_off[2];                    // ping-pong dZ buffers
  size_t partial_off, partial_stride;  // split partials for dW / db
  int splits, rows_per_split;
  size_t whi_off, wlo_off;             // TF32 hi / lo copies of the flat parameters
  // tensor-core path: per-tile column sums and the fine-grain output-layer slots
  size_t tile_off, tile_stride;        // [4 * ceil(M / 128)][tile_stride]
  int tile_slots;
  size_t oslot_off, oslot_stride;      // [ceil(M / out_rows)][oslot_stride]
  int out_rows, out_slots;
  size_t ws_floats;
};

// Returns 0 on success and fills `p`; sets the error string otherwise.
int make_mlp_plan(const tfr_mlp_cfg* cfg, int M, MlpPlan* p);

int mlp_simt_fwd(const float* X, int M, const MlpPlan& p, const float* params,
                 const uint8_t* mask, float* ws, float* scores, cudaStream_t st);
int mlp_simt_bwd(const float* X, int M, const MlpPlan& p, const float* params,
                 const float* dscores, const uint8_t* mask, float* ws, float* grads,
                 cudaStream_t st);

// tensor-core (tcgen05) path; passes = 1 (TF32) or 3 (3xTF32, fp32-faithful)
int mlp_tc_fwd(const float* X, int M, const MlpPlan& p, const float* params,
               const uint8_t* mask, float* ws, float* scores, int passes, cudaStream_t st);
int mlp_tc_bwd(const float* X, int M, const MlpPlan& p, const float* params,
               const float* dscores, const uint8_t* mask, float* ws, float* grads,
               int passes, cudaStream_t st);

// pieces of the CUDA-core path reused by the tensor-core path
int mlp_out_layer_fwd(const float* H, int M, int K, int O, const float* W, const float* bias,
                      const uint8_t* mask, float* scores, cudaStream_t st);
int mlp_out_layer_bwd(const float* H, int M, int K, int O, const float* W, const float* dS,
                      const uint8_t* mask, int act, int rows_per, int splits, float* dH,
                      float* partial, size_t pstride, cudaStream_t st);
int mlp_out_layer_bwd2(const float* H, int M, int K, int O, const float* W, const float* dS,
                       const uint8_t* mask, int act, int rows_per, float* dH, float* slots,
                       size_t slot_stride, cudaStream_t st);
int mlp_regroup_sum(const float* src, int slots_in, size_t src_stride, size_t src_off, int n,
                    int group, float* dst, int slots_out, size_t dst_stride, size_t dst_off,
                    cudaStream_t st);
int mlp_reduce2(const float* srcA, int slotsA, size_t strideA, size_t nA, const float* srcB,
                int slotsB, size_t strideB, size_t nB, float* out, cudaStream_t st);
int mlp_colsum(const float* dZ, int M, int N, int rows_per, int splits, float* partial,
               size_t pstride, size_t col_offset, cudaStream_t st);
int mlp_reduce_partials(const float* partial, int splits, size_t stride, size_t n, float* out,
                        cudaStream_t st);

}  // namespace tfr
