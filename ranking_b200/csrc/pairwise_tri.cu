// K1 (triangular form): PairwiseLogistic / Hinge / SoftZeroOne losses with every
// LambdaWeight, forward + backward, ONE phi evaluation per unordered pair.
//
// Replaces the both-ends walk of pairwise_loss_kernel for list sizes <= 1024
// (losses_impl.py:61-74, 483-537, 871-884, 933-958 and the LambdaWeight family
// :170-454).  What changes with respect to that kernel:
//
//   * a pair {a, b} is visited once: with s = sign(l_a - l_b) the preferred item is
//     known, phi is evaluated on s (z_a - z_b) and the gradient term goes to row a with
//     +s and to column b with -s (per-lane column registers, as in the ApproxNDCG
//     kernel), so the 2-3 MUFU operations of phi are paid N^2/2 times, not N^2;
//   * no three-way branch: the pair weight is a select, inactive pairs (equal labels,
//     padding) carry a NaN label so both label comparisons fail;
//   * when the lambda needs ranks the list is SORTED by score in shared memory (64-bit
//     key bitonic sort: invalid last, score descending, ties by index — the same order
//     as the counting ranks of compute_ranks).  In walk order rank = position + 1, so
//     |r_a - r_b| = b - a and the discount-difference table u[d] = |d(d) - d(d+1)| is
//     read at consecutive addresses by consecutive lanes (no bank conflicts, no
//     rank gathers); the top-n tests become index comparisons that are uniform per row;
//     padded items sit at the tail and their tiles are skipped;
//   * small lists share a CTA: 32 / 64 / 128 / 256 threads per list for N <= 32 / 64 /
//     128 / > 128 (named barriers per list), so a 256-thread CTA works on 8 / 4 / 2 / 1
//     lists;
//   * columns are processed in chunks of <= 8 tiles of 32, so the register footprint
//     and the code size do not grow with N.
//
// PairwiseMSELoss (all ordered pairs, different weighting) and N > 1024 stay on
// pairwise_loss_kernel.
#include "loss_common.cuh"

namespace tfr {

constexpr int kTriThreads = 256;

struct Grp {
  int tid, nthr, lane, warp, nwarps, bar;
  float* red;   // [8] reduction scratch of this list's group
};

__device__ __forceinline__ void gsync(const Grp& g) {
  if (g.nthr == 32) {
    __syncwarp();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(g.bar), "r"(g.nthr) : "memory");
  }
}

__device__ __forceinline__ float gsum(float v, const Grp& g) {
  v = warp_sum(v);
  if (g.nwarps == 1) return v;
  gsync(g);   // red may still be read from a previous reduction
  if (g.lane == 0) g.red[g.warp] = v;
  gsync(g);
  float r = 0.f;
  for (int w = 0; w < g.nwarps; ++w) r += g.red[w];
  return r;
}

__device__ __forceinline__ uint32_t desc_key(float x) {
  if (x == 0.f) x = 0.f;   // -0 == +0
  uint32_t u = __float_as_uint(x);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending-order preserving
  return ~u;                                         // descending
}

__device__ inline void group_bitonic_sort(unsigned long long* keys, int P, const Grp& g) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int h = g.tid; h < (P >> 1); h += g.nthr) {
        const int a = ((h & ~(j - 1)) << 1) | (h & (j - 1));
        const int b = a | j;
        const bool asc = (a & k) == 0;
        const unsigned long long x = keys[a], y = keys[b];
        if ((x > y) == asc) {
          keys[a] = y;
          keys[b] = x;
        }
      }
      gsync(g);
    }
  }
}

// Shared memory of one list (floats unless noted), N4 = N rounded up to 4:
//   z, l, w, g, dt [N4] each   walk-order logits, pair labels, weights, gains, d~ / mult
//   u [32 T + 64]              discount differences, u[d] valid for d in [-32, 32 T + 32)
//                              (zero outside [1, N]: inactive lanes read without a bounds test)
//   perm [N4] (int)            original index of a walk position
//   rowg [N4], rowl [N4]       row sums (gradient, row loss)
//   scratch                    max(staging, column partials): staging = z0, l0, w0, g0 [N4]
//                              + keys [P] (u64) + flags [N4 bytes]; column partials =
//                              nwarps x N4 (x 2 with row losses)
//   red [8]
__host__ __device__ inline size_t tri_list_floats(int N, int nwarps, int rows) {
  const size_t N4 = (size_t)(N + 3) & ~(size_t)3;
  size_t P = 32;
  while (P < (size_t)N) P <<= 1;
  const size_t staging = 4 * N4 + 2 * P + (N4 + 3) / 4 + 4;
  const size_t colpart = (size_t)nwarps * N4 * (rows ? 2 : 1);
  const size_t scratch = staging > colpart ? staging : colpart;
  const size_t usize = (((size_t)N + 31) & ~(size_t)31) + 64;
  return 5 * N4 + usize + N4 + 2 * N4 + ((scratch + 3) & ~(size_t)3) + 8;
}

template <int PHI, int LAM, int TC, bool ROWS>
__global__ void __launch_bounds__(kTriThreads, 2)
pairwise_tri_kernel(const float* __restrict__ scores, const float* __restrict__ labels,
                    const float* __restrict__ item_w, int w_per_item,
                    const uint8_t* __restrict__ mask, int B, int N, float temperature,
                    LamDev lam, float grad_scale, int want_sort, float* __restrict__ grad,
                    float* __restrict__ row_loss, float* __restrict__ loss_sum,
                    float* __restrict__ w_sum, float* __restrict__ nnz,
                    int32_t* __restrict__ ranks_out) {
  extern __shared__ __align__(16) float smem_f[];
  // ---- group (= list) geometry -------------------------------------------------
  constexpr int GT = TC >= 8 ? 256 : 32 * TC;   // threads per list
  constexpr int LPC = kTriThreads / GT;         // lists per CTA
  Grp g;
  const int gid = threadIdx.x / GT;
  g.tid = threadIdx.x % GT;
  g.nthr = GT;
  g.lane = threadIdx.x & 31;
  g.warp = g.tid >> 5;
  g.nwarps = GT / 32;
  g.bar = 1 + gid;
  const int b = blockIdx.x * LPC + gid;
  if (b >= B) return;   // whole groups leave together; named barriers are per group

  const int N4 = (N + 3) & ~3;
  int P = 32;
  while (P < N) P <<= 1;
  float* base = smem_f + (size_t)gid * tri_list_floats(N, g.nwarps, ROWS ? 1 : 0);
  float* sz = base;
  float* sl = sz + N4;
  float* sw = sl + N4;
  float* sg = sw + N4;
  float* sdt = sg + N4;
  const int usize = ((N + 31) & ~31) + 64;
  float* su = sdt + N4 + 32;   // su[-32 .. usize - 32)
  int* perm = reinterpret_cast<int*>(sdt + N4 + usize);
  float* rowg = reinterpret_cast<float*>(perm + N4);
  float* rowl = rowg + N4;
  float* scratch = rowl + N4;
  // staging view of the scratch region
  float* z0 = scratch;
  float* l0 = z0 + N4;
  float* w0 = l0 + N4;
  float* g0 = w0 + N4;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(g0 + N4);
  unsigned char* flags = reinterpret_cast<unsigned char*>(keys + P);   // bit0 mv, bit1 lv
  {
    size_t P_ = P;
    const size_t staging = 4 * (size_t)N4 + 2 * P_ + ((size_t)N4 + 3) / 4 + 4;
    const size_t colpart = (size_t)g.nwarps * N4 * (ROWS ? 2 : 1);
    const size_t sc = staging > colpart ? staging : colpart;
    g.red = scratch + ((sc + 3) & ~(size_t)3);
  }

  constexpr bool kRanked = LAM >= TFR_LAMBDA_DCG;
  const bool do_sort = kRanked || want_sort;
  const size_t off = (size_t)b * N;

  // ---- stage the list ------------------------------------------------------------
  for (int i = g.tid; i < N; i += GT) {
    const float z = scores[off + i] / temperature;
    const float lab = labels[off + i];
    const bool lvalid = lab >= 0.f;
    const bool mvalid = mask ? (mask[off + i] != 0) : lvalid;
    float wv = 1.f;
    if (item_w) wv = w_per_item ? item_w[off + i] : item_w[b];
    z0[i] = z;
    l0[i] = lab;
    w0[i] = lvalid ? wv : 0.f;
    flags[i] = (mvalid ? 1 : 0) | (lvalid ? 2 : 0);
    if (kRanked) {
      const float cl = lvalid ? lab : 0.f;
      float gn;
      if (lam.gain_fn == TFR_GAIN_TABLE) gn = lam.gain_table[off + i];
      else if (LAM == TFR_LAMBDA_PRECISION) gn = cl >= 1.f ? 1.f : 0.f;
      else gn = gain_of(lam.gain_fn, cl);
      g0[i] = gn;
    }
  }
  gsync(g);

  // ---- lambda tables ---------------------------------------------------------------
  const int topn = lam.topn > 0 ? lam.topn : N;
  float gnorm = 1.f;
  if (kRanked) {
    // inverse max DCG (losses_impl.py:109-134): labels sorted descending, ties by index
    if (lam.normalized && LAM != TFR_LAMBDA_PRECISION) {
      for (int i = g.tid; i < P; i += GT) {
        unsigned long long key = ~0ull;
        if (i < N) {
          const float cl = (flags[i] & 2) ? l0[i] : 0.f;
          key = ((unsigned long long)desc_key(cl) << 20) | (unsigned long long)i;
        }
        keys[i] = key;
      }
      gsync(g);
      group_bitonic_sort(keys, P, g);
      const int tn = lam.topn > 0 ? min(lam.topn, N) : N;
      float part = 0.f;
      for (int k = g.tid; k < tn; k += GT) {
        const int i = (int)(keys[k] & 0xFFFFFu);
        const float d = lam.disc_fn == TFR_DISC_TABLE ? lam.disc_table[k + 1]
                                                      : disc_of(lam.disc_fn, (float)(k + 1));
        part += g0[i] * d;
      }
      const float s = gsum(part, g);
      gnorm = s > 0.f ? 1.f / s : 0.f;
      gsync(g);
    }
    // u[d] = |d(d) - d(d+1)| (x N, x (1 - alpha) for DCG), walk-order d~ / multiplier
    const float nf = (float)N;
    const float us = LAM == TFR_LAMBDA_DCG ? (1.f - lam.alpha) * nf : nf;
    for (int d = g.tid - 32; d < usize - 32; d += GT) {
      float v = 0.f;
      if (d >= 1 && d <= N) {
        const float a = lam.disc_fn == TFR_DISC_TABLE ? lam.disc_table[d]
                                                      : disc_of(lam.disc_fn, (float)d);
        const float c = lam.disc_fn == TFR_DISC_TABLE ? lam.disc_table[d + 1]
                                                      : disc_of(lam.disc_fn, (float)(d + 1));
        v = fabsf(a - c) * us;
      }
      su[d] = v;
    }
    for (int p = g.tid; p < N; p += GT) {
      const int r = p + 1;
      const float d = lam.disc_fn == TFR_DISC_TABLE ? lam.disc_table[r]
                                                    : disc_of(lam.disc_fn, (float)r);
      float v;
      if (LAM == TFR_LAMBDA_DCG) v = r <= topn ? d * lam.alpha * nf : 0.f;
      else v = r > topn ? 1.f / (1.f - d) : 1.f;   // V2 / Yeti multiplier of max(r_a, r_b)
      sdt[p] = v;
    }
  }

  // ---- walk order: sorted by (invalid, score desc, index) or the storage order -------
  if (do_sort) {
    for (int i = g.tid; i < P; i += GT) {
      unsigned long long key = ~0ull;
      if (i < N) {
        const bool mv = (flags[i] & 1) != 0;   // invalid items: last, among themselves by index
        key = ((unsigned long long)(mv ? 0 : 1) << 52) |
              ((unsigned long long)desc_key(mv ? z0[i] : 0.f) << 20) | (unsigned long long)i;
      }
      keys[i] = key;
    }
    gsync(g);
    group_bitonic_sort(keys, P, g);
  }
  int nv_part = 0;
  for (int p = g.tid; p < N; p += GT) {
    const int i = do_sort ? (int)(keys[p] & 0xFFFFFu) : p;
    const unsigned char f = flags[i];
    // pair-eligible: mask-valid (and, for the rank-based lambdas, label-valid: their
    // weight is zero otherwise)
    const bool ok = kRanked ? (f == 3) : ((f & 1) != 0);
    perm[p] = i;
    sz[p] = z0[i];
    sl[p] = ok ? l0[i] : CUDART_NAN_F;
    sw[p] = w0[i];
    if (kRanked) sg[p] = g0[i] * gnorm;
    rowg[p] = 0.f;
    if (ROWS) rowl[p] = 0.f;
    if (ok) nv_part = max(nv_part, p + 1);
    if (ranks_out) ranks_out[off + i] = p + 1;
  }
  // number of walk positions that can take part in a pair (exclusive upper bound)
  int nv = nv_part;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nv = max(nv, __shfl_xor_sync(0xffffffffu, nv, o));
  if (g.nwarps > 1) {
    gsync(g);
    if (g.lane == 0) g.red[g.warp] = __int_as_float(nv);
    gsync(g);
    for (int w = 0; w < g.nwarps; ++w) nv = max(nv, __float_as_int(g.red[w]));
  }
  gsync(g);   // staging is dead from here on: the scratch region becomes column partials
  float* colg = scratch;
  float* coll = scratch + (size_t)g.nwarps * N4;

  // ---- pair loop -------------------------------------------------------------------
  float acc_l = 0.f, acc_w = 0.f, acc_n = 0.f;
  const int T = (N + 31) >> 5;                 // column tiles of the list
  const int tmax = (nv + 31) >> 5;             // tiles that hold a pair-eligible item
  const int nchunks = (tmax + TC - 1) / TC;
  // DCG: pairs need r_a <= topn or r_b <= topn; with a < b that is a < topn (per row)
  const int row_end = LAM == TFR_LAMBDA_DCG ? min(nv, topn) : nv;

  for (int c = 0; c < nchunks; ++c) {
    const int ct0 = c * TC;                    // first column tile of the chunk
    float zc[TC], lc[TC], wc[TC], gc[TC], dc[TC], cg[TC], cl[TC];
#pragma unroll
    for (int t = 0; t < TC; ++t) {
      const int j = g.lane + 32 * (ct0 + t);
      const bool in = j < N;
      zc[t] = in ? sz[j] : 0.f;
      lc[t] = in ? sl[j] : CUDART_NAN_F;
      wc[t] = in ? sw[j] : 0.f;
      gc[t] = (kRanked && in) ? sg[j] : 0.f;
      dc[t] = (kRanked && in) ? sdt[j] : 0.f;
      cg[t] = 0.f;
      cl[t] = 0.f;
    }
    // one row against the column tiles [tb, TC) of this chunk; DIAG: tile tb holds the row
    auto do_row = [&](int i, int tb, bool diag) {
      const float zi = sz[i], li = sl[i], wi = sw[i];
      const float gi = kRanked ? sg[i] : 0.f;
      const float di = kRanked ? sdt[i] : 0.f;
      float rg = 0.f, rl = 0.f;
#pragma unroll
      for (int t = 0; t < TC; ++t) {
        if (t < tb || ct0 + t >= tmax) continue;
        const int j = g.lane + 32 * (ct0 + t);
        const float lj = lc[t];
        const bool a_hi = li > lj, b_hi = lj > li;      // NaN labels: both false
        bool active = a_hi || b_hi;
        if (diag && t == tb) active = active && (j > i);
        const float d = zi - zc[t];
        float x = a_hi ? d : -d;                        // logit of preferred - other
        x = active ? x : 0.f;                           // keeps phi finite on padding
        float f, df;
        phi_eval<PHI>(x, f, df);
        float lw = 1.f;
        if (LAM == TFR_LAMBDA_LABEL_DIFF) {
          lw = fabsf(li - lj);
        } else if (LAM == TFR_LAMBDA_DCG) {
          lw = fabsf(gi - gc[t]) * (su[j - i] + fabsf(di - dc[t]));
        } else if (LAM == TFR_LAMBDA_DCG_V2) {
          lw = fabsf(gi - gc[t]) * su[j - i] * dc[t];
        } else if (LAM == TFR_LAMBDA_YETI) {
          lw = (j - i == 1) ? fabsf(gi - gc[t]) * su[1] * dc[t] : 0.f;
        } else if (LAM == TFR_LAMBDA_PRECISION) {
          lw = ((i < topn) != (j < topn)) ? fabsf(gi - gc[t]) : 0.f;
        }
        const float W = active ? lw * (a_hi ? wi : wc[t]) : 0.f;
        const float Wf = W * f;
        acc_l += Wf;
        acc_w += W;
        acc_n += W != 0.f ? 1.f : 0.f;
        const float tt = a_hi ? W * df : -(W * df);     // d loss / d z_a ; z_b gets -tt
        rg += tt;
        cg[t] -= tt;
        if (ROWS) {
          rl += a_hi ? Wf : 0.f;
          cl[t] += a_hi ? 0.f : Wf;
        }
      }
      rg = warp_sum(rg);
      if (ROWS) rl = warp_sum(rl);
      if (g.lane == 0) {        // a row always belongs to the same warp: no race
        rowg[i] += rg;
        if (ROWS) rowl[i] += rl;
      }
    };
    // rows strictly above the chunk: every column tile of the chunk is to their right
    {
      const int iend = min(row_end, 32 * ct0);
      for (int i = g.warp; i < iend; i += g.nwarps) do_row(i, 0, false);
    }
    // rows inside the chunk's own tiles: triangular
#pragma unroll
    for (int tb = 0; tb < TC; ++tb) {
      const int i0 = 32 * (ct0 + tb);
      const int iend = min(row_end, i0 + 32);
      for (int i = i0 + g.warp; i < iend; i += g.nwarps) do_row(i, tb, true);
    }
    // column partials of this chunk (each warp owns a row of the partial matrix)
#pragma unroll
    for (int t = 0; t < TC; ++t) {
      const int j = g.lane + 32 * (ct0 + t);
      if (j < N) {
        colg[(size_t)g.warp * N4 + j] = cg[t];
        if (ROWS) coll[(size_t)g.warp * N4 + j] = cl[t];
      }
    }
  }
  gsync(g);

  // ---- outputs ----------------------------------------------------------------------
  const float inv_t = grad_scale / temperature;
  const int covered = min(N, nchunks * TC * 32);   // columns that own a partial
  for (int p = g.tid; p < N; p += GT) {
    float a = rowg[p], r = ROWS ? rowl[p] : 0.f;
    if (p < covered)
      for (int w = 0; w < g.nwarps; ++w) {
        a += colg[(size_t)w * N4 + p];
        if (ROWS) r += coll[(size_t)w * N4 + p];
      }
    const int i = perm[p];
    if (grad) grad[off + i] = a * inv_t;
    if (ROWS) row_loss[off + i] = r;
  }
  acc_l = gsum(acc_l, g);
  acc_w = gsum(acc_w, g);
  acc_n = gsum(acc_n, g);
  if (g.tid == 0) {
    loss_sum[b] = acc_l;
    if (w_sum) w_sum[b] = acc_w;
    if (nnz) nnz[b] = acc_n;
  }
}

template <int PHI, int LAM, int TC, bool ROWS>
static int launch_tri_one(cudaStream_t st, const float* scores, const float* labels,
                          const float* item_w, int w_per_item, const uint8_t* mask, int B,
                          int N, float temperature, LamDev lam, float grad_scale, float* grad,
                          float* row_loss, float* loss_sum, float* w_sum, float* nnz,
                          int32_t* ranks_out) {
  constexpr int GT = TC >= 8 ? 256 : 32 * TC;
  constexpr int LPC = kTriThreads / GT;
  const size_t smem = (size_t)LPC * tri_list_floats(N, GT / 32, ROWS ? 1 : 0) * sizeof(float);
  auto kern = pairwise_tri_kernel<PHI, LAM, TC, ROWS>;
  if (smem > 48 * 1024)
    TFR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  const int grid = (B + LPC - 1) / LPC;
  kern<<<grid, kTriThreads, smem, st>>>(scores, labels, item_w, w_per_item, mask, B, N,
                                        temperature, lam, grad_scale, ranks_out != nullptr,
                                        grad, row_loss, loss_sum, w_sum, nnz, ranks_out);
  TFR_LAUNCH_OK();
  return TFR_OK;
}

template <int PHI, int LAM>
static int launch_tri_lam(cudaStream_t st, const float* scores, const float* labels,
                          const float* item_w, int w_per_item, const uint8_t* mask, int B,
                          int N, float temperature, LamDev lam, float grad_scale, float* grad,
                          float* row_loss, float* loss_sum, float* w_sum, float* nnz,
                          int32_t* ranks_out) {
#define TFR_TRI(TC_, ROWS_)                                                                 \
  return launch_tri_one<PHI, LAM, TC_, ROWS_>(st, scores, labels, item_w, w_per_item, mask, \
                                              B, N, temperature, lam, grad_scale, grad,     \
                                              row_loss, loss_sum, w_sum, nnz, ranks_out)
  const int T = (N + 31) / 32;
  if (row_loss) {
    if (T <= 1) TFR_TRI(1, true);
    if (T <= 2) TFR_TRI(2, true);
    if (T <= 4) TFR_TRI(4, true);
    TFR_TRI(8, true);
  } else {
    if (T <= 1) TFR_TRI(1, false);
    if (T <= 2) TFR_TRI(2, false);
    if (T <= 4) TFR_TRI(4, false);
    TFR_TRI(8, false);
  }
#undef TFR_TRI
}

template <int PHI>
static int launch_tri_phi(int lamkind, cudaStream_t st, const float* scores,
                          const float* labels, const float* item_w, int w_per_item,
                          const uint8_t* mask, int B, int N, float temperature, LamDev lam,
                          float grad_scale, float* grad, float* row_loss, float* loss_sum,
                          float* w_sum, float* nnz, int32_t* ranks_out) {
#define TFR_CASE(L)                                                                        \
  case L:                                                                                  \
    return launch_tri_lam<PHI, L>(st, scores, labels, item_w, w_per_item, mask, B, N,      \
                                  temperature, lam, grad_scale, grad, row_loss, loss_sum,  \
                                  w_sum, nnz, ranks_out);
  switch (lamkind) {
    TFR_CASE(TFR_LAMBDA_NONE)
    TFR_CASE(TFR_LAMBDA_LABEL_DIFF)
    TFR_CASE(TFR_LAMBDA_DCG)
    TFR_CASE(TFR_LAMBDA_DCG_V2)
    TFR_CASE(TFR_LAMBDA_YETI)
    TFR_CASE(TFR_LAMBDA_PRECISION)
  }
#undef TFR_CASE
  set_error("bad lambda kind %d", lamkind);
  return TFR_INVALID_ARGUMENT;
}

// Entry used by tfr_pairwise_loss_fwd_bwd (loss_kernels.cu) for phi != MSE, N <= 1024.
int launch_pairwise_tri(int phi, cudaStream_t st, const float* scores, const float* labels,
                        const float* item_w, int w_per_item, const uint8_t* mask, int B, int N,
                        float temperature, const LamDev& lam, float grad_scale, float* grad,
                        float* row_loss, float* loss_sum, float* w_sum, float* nnz,
                        int32_t* ranks_out) {
  switch (phi) {
    case TFR_PHI_LOGISTIC:
      return launch_tri_phi<TFR_PHI_LOGISTIC>(lam.kind, st, scores, labels, item_w, w_per_item,
                                              mask, B, N, temperature, lam, grad_scale, grad,
                                              row_loss, loss_sum, w_sum, nnz, ranks_out);
    case TFR_PHI_HINGE:
      return launch_tri_phi<TFR_PHI_HINGE>(lam.kind, st, scores, labels, item_w, w_per_item,
                                           mask, B, N, temperature, lam, grad_scale, grad,
                                           row_loss, loss_sum, w_sum, nnz, ranks_out);
    case TFR_PHI_SOFT_ZERO_ONE:
      return launch_tri_phi<TFR_PHI_SOFT_ZERO_ONE>(lam.kind, st, scores, labels, item_w,
                                                   w_per_item, mask, B, N, temperature, lam,
                                                   grad_scale, grad, row_loss, loss_sum, w_sum,
                                                   nnz, ranks_out);
  }
  set_error("phi %d has no triangular kernel", phi);
  return TFR_INVALID_ARGUMENT;
}

}  // namespace tfr
