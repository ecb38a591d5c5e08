// Pieces shared by the per-list loss kernels (loss_kernels.cu, pairwise_tri.cu): list
// staging in shared memory, counting ranks, lambda-weight setup, MUFU helpers.
#pragma once

#include <math_constants.h>

#include "common.cuh"

namespace tfr {

constexpr int kLossThreads = 256;
constexpr int kMaxListSize = 8192;

struct LamDev {
  int kind, topn, gain_fn, disc_fn, normalized;
  float alpha;
  const float* gain_table;
  const float* disc_table;
};

static inline LamDev make_lam(const tfr_lambda_cfg* h) {
  LamDev d;
  if (h == nullptr) {
    d = LamDev{TFR_LAMBDA_NONE, 0, 0, 0, 0, 0.f, nullptr, nullptr};
  } else {
    d = LamDev{h->kind, h->topn, h->gain_fn, h->disc_fn, h->normalized,
               h->smooth_fraction, h->gain_table, h->disc_table};
  }
  return d;
}

static inline int check_lam(const tfr_lambda_cfg* h) {
  if (h == nullptr) return TFR_OK;
  TFR_REQUIRE(h->kind >= TFR_LAMBDA_NONE && h->kind <= TFR_LAMBDA_PRECISION,
              "lambda kind %d is not a tfr_lambda_kind", h->kind);
  TFR_REQUIRE(h->smooth_fraction >= 0.f && h->smooth_fraction <= 1.f,
              "smooth_fraction %g should be in range [0, 1].", h->smooth_fraction);
  TFR_REQUIRE(h->gain_fn >= 0 && h->gain_fn <= TFR_GAIN_TABLE, "bad gain_fn %d", h->gain_fn);
  TFR_REQUIRE(h->disc_fn >= 0 && h->disc_fn <= TFR_DISC_TABLE, "bad disc_fn %d", h->disc_fn);
  TFR_REQUIRE(h->gain_fn != TFR_GAIN_TABLE || h->gain_table != nullptr,
              "gain_fn TABLE needs gain_table");
  TFR_REQUIRE(h->disc_fn != TFR_DISC_TABLE || h->disc_table != nullptr,
              "disc_fn TABLE needs disc_table");
  return TFR_OK;
}

__host__ __device__ inline bool lam_needs_rank(int kind) {
  return kind >= TFR_LAMBDA_DCG;
}

// Shared-memory view of one list.
struct ListView {
  float* z;     // [N] logits / temperature
  float* l;     // [N] raw labels
  float* w;     // [N] row-item weight (0 where the label is invalid)
  float* g;     // [N] lambda gains
  float* disc;  // [N + 2] rank discount table, disc[r] = d(r)
  int* rank;    // [N] 1-based ranks
  float* red;   // [32] reduction scratch
  unsigned char* mv;  // [N] valid per mask (or label >= 0)
  unsigned char* lv;  // [N] label >= 0
};

__host__ __device__ inline size_t list_smem_bytes(int N) {
  return (size_t)(4 * N + (N + 2) + N + 32) * 4 + 2 * (size_t)N + 16;
}

__device__ inline ListView carve(unsigned char* base, int N) {
  ListView v;
  float* f = reinterpret_cast<float*>(base);
  v.z = f; f += N;
  v.l = f; f += N;
  v.w = f; f += N;
  v.g = f; f += N;
  v.disc = f; f += N + 2;
  v.rank = reinterpret_cast<int*>(f); f += N;
  v.red = f; f += 32;
  v.mv = reinterpret_cast<unsigned char*>(f);
  v.lv = v.mv + N;
  return v;
}

// rank_i = 1 + #{j that precede i}, where j precedes i if it is valid and i is not,
// or (same validity) s_j > s_i, or s_j == s_i and j < i.  The reference gives
// invalid entries the score zmin - 1e-6 (losses_impl.py:497-499) so that they
// sort last; in fp32 that sentinel can be absorbed (|zmin| >= 16) and then ties
// with the worst valid item are broken randomly.  Here invalid entries are
// strictly last and ties are broken by index (shuffle_ties=False semantics).
__device__ inline void compute_ranks(const ListView& v, int N, float /*zmin*/) {
  // One thread per item; the (validity, score) reads of the partners are smem broadcasts.
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const bool vi = v.mv[i];
    const float si = v.z[i];
    int cnt = 0;
#pragma unroll 4
    for (int j = 0; j < N; ++j) {
      const bool vj = v.mv[j];
      const float sj = v.z[j];
      const bool same = vi == vj;
      // different validity: the valid one precedes; both invalid: by index
      const bool by_score = (sj > si) || (sj == si && j < i);
      const bool before = same ? (vi ? by_score : j < i) : vj;
      cnt += before;
    }
    v.rank[i] = cnt + 1;
  }
}

// sum_{k <= topn} gain(l_(k)) * disc(k) over labels sorted descending
// (inverse_max_dcg, losses_impl.py:109-134); `gain` holds per-item gains of the
// cleaned labels `cl`.  Returns the sum to every thread.
template <typename DiscFn>
__device__ inline float ideal_dcg(const float* cl, const float* gain, int N,
                                  int topn, float* red, DiscFn disc) {
  // One thread per item: its rank among the labels by counting (label reads are smem
  // broadcasts), then every lane evaluates its own discount.  (A warp-per-row count
  // with a lane-0 discount serialises ~50 transcendental instructions per row.)
  float part = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float li = cl[i];
    int cnt = 0;
#pragma unroll 4
    for (int j = 0; j < N; ++j) {
      const float lj = cl[j];
      cnt += (lj > li) || (lj == li && j < i);
    }
    if (cnt + 1 <= topn) part += gain[i] * disc(cnt + 1);
  }
  return block_sum(part, red);
}

template <int LAM>
__device__ __forceinline__ float pair_lambda(const LamDev& lam, const ListView& v,
                                             int N, int topn, int i, int j) {
  if (LAM == TFR_LAMBDA_NONE) return 1.f;
  if (LAM == TFR_LAMBDA_LABEL_DIFF) return fabsf(v.l[i] - v.l[j]);
  if (!(v.lv[i] && v.lv[j])) return 0.f;
  const int ri = v.rank[i], rj = v.rank[j];
  const float dg = fabsf(v.g[i] - v.g[j]);
  if (LAM == TFR_LAMBDA_PRECISION) {
    return ((ri <= topn) != (rj <= topn)) ? dg : 0.f;
  }
  const int d = ri > rj ? ri - rj : rj - ri;
  float pd;
  if (LAM == TFR_LAMBDA_DCG) {
    const bool in_top = (ri <= topn) || (rj <= topn);
    const float u = (d > 0 && in_top) ? fabsf(v.disc[d] - v.disc[d + 1]) : 0.f;
    const float di = ri > topn ? 0.f : v.disc[ri];
    const float dj = rj > topn ? 0.f : v.disc[rj];
    pd = (1.f - lam.alpha) * u + lam.alpha * fabsf(di - dj);
    if (!in_top) pd = 0.f;
  } else {  // V2 / YETI
    const int mx = ri > rj ? ri : rj;
    const float mult = mx > topn ? 1.f / (1.f - v.disc[mx]) : 1.f;
    pd = d > 0 ? fabsf(v.disc[d] - v.disc[d + 1]) * mult : 0.f;
    if (LAM == TFR_LAMBDA_YETI && d != 1) pd = 0.f;
  }
  return dg * pd * (float)N;
}

// Fill disc table, gains (and inverse max DCG normalisation) for a lambda.
__device__ inline void setup_lambda(const LamDev& lam, const ListView& v, int b,
                                    int N, int tid) {
  if (!lam_needs_rank(lam.kind)) return;
  for (int r = tid; r < N + 2; r += blockDim.x)
    v.disc[r] = lam.disc_fn == TFR_DISC_TABLE ? lam.disc_table[r]
                                              : disc_of(lam.disc_fn, (float)r);
  for (int i = tid; i < N; i += blockDim.x) {
    const float cl = v.lv[i] ? v.l[i] : 0.f;
    float g;
    if (lam.gain_fn == TFR_GAIN_TABLE) g = lam.gain_table[(size_t)b * N + i];
    else if (lam.kind == TFR_LAMBDA_PRECISION) g = cl >= 1.f ? 1.f : 0.f;
    else g = gain_of(lam.gain_fn, cl);
    v.g[i] = g;
  }
  __syncthreads();
  if (lam.normalized && lam.kind != TFR_LAMBDA_PRECISION) {
    // cleaned labels are needed for the ideal ordering; reuse rank[] as scratch
    // is not possible (ranks may already be there), so recompute cl on the fly
    // through a small lambda over v.l / v.lv.
    const int topn = lam.topn > 0 ? min(lam.topn, N) : N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    float part = 0.f;
    for (int i = warp; i < N; i += nwarps) {
      const float li = v.lv[i] ? v.l[i] : 0.f;
      int cnt = 0;
      for (int j = lane; j < N; j += 32) {
        const float lj = v.lv[j] ? v.l[j] : 0.f;
        cnt += (lj > li) || (lj == li && j < i);
      }
      cnt = warp_sum_int(cnt);
      if (lane == 0 && cnt + 1 <= topn) part += v.g[i] * v.disc[cnt + 1];
    }
    const float s = block_sum(part, v.red);
    const float inv = s > 0.f ? 1.f / s : 0.f;
    for (int i = tid; i < N; i += blockDim.x) v.g[i] *= inv;
    __syncthreads();
  }
}

// Load one list into shared memory.  Returns the row minimum of z (all entries).
__device__ inline float load_list(const ListView& v, const float* scores,
                                  const float* labels, const float* item_w,
                                  int w_per_item, const uint8_t* mask, int b,
                                  int N, float temperature) {
  float zmin = CUDART_INF_F;
  const size_t off = (size_t)b * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float z = scores[off + i] / temperature;
    const float lab = labels[off + i];
    const bool lvalid = lab >= 0.f;
    const bool mvalid = mask ? (mask[off + i] != 0) : lvalid;
    float wv = 1.f;
    if (item_w) wv = w_per_item ? item_w[off + i] : item_w[b];
    v.z[i] = z;
    v.l[i] = lab;
    v.w[i] = lvalid ? wv : 0.f;
    v.mv[i] = mvalid;
    v.lv[i] = lvalid;
    zmin = fminf(zmin, z);
  }
  zmin = block_min(zmin, v.red);  // includes the barrier that publishes the loads
  return zmin;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float exp2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {   // one MUFU.RCP (~1 ulp), no fix-up
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float log2f_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------
template <int PHI>
__device__ __forceinline__ void phi_eval(float x, float& f, float& df) {
  if (PHI == TFR_PHI_LOGISTIC) {
    // relu(-x) + log1p(exp(-|x|)); d/dx = -sigmoid(-x)   (losses_impl.py:936-940)
    // MUFU path: e = 2^(-|x| log2 e), log1p(e) = lg2(1 + e) ln 2, 1/(1+e) by rcp
    const float e = exp2f_approx(-fabsf(x) * kLog2e);
    const float rc = rcp_approx(1.f + e);
    f = fmaxf(-x, 0.f) + log2f_approx(1.f + e) * kLn2;
    df = -(x >= 0.f ? e * rc : rc);
  } else if (PHI == TFR_PHI_HINGE) {
    const float m = 1.f - x;  // relu(1 - x)            (losses_impl.py:946-948)
    f = fmaxf(m, 0.f);
    df = m > 0.f ? -1.f : 0.f;
  } else {
    // sigmoid(-x); d/dx = -sigmoid(x) sigmoid(-x)       (losses_impl.py:954-958)
    const float e = exp2f_approx(-fabsf(x) * kLog2e);
    const float inv = rcp_approx(1.f + e);
    f = x > 0.f ? e * inv : inv;
    df = -e * inv * inv;
  }
}

}  // namespace tfr
