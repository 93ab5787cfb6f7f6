// The TOLD training step on the device: TDMPC.update (algorithm/tdmpc.py:191-245) -- forward over
// the H-step latent unroll, TD targets (_td_target, :180-189), backward through time, clip_grad_norm_, Adam, the policy update
// (update_pi, :165-178) and the target EMA (helper.ema, helper.py:48-52) -- for one replay batch, fp32 tensors throughout.
//
// Layer-by-layer over the whole batch (rows = batch, or horizon x batch where the steps are independent): every nn.Linear is one
// GEMM call (forward Y = X W^T + b, data gradient dX = dY W, weight gradient dW += dY^T X) -- on the tensor cores at fp32 accuracy
// (gemm_tc.cuh: operands split into hi + lo bf16, three tcgen05 MMAs per product) with the tiled fp32 CUDA-core kernel below for the
// few-row / few-column products; the activations the backward pass needs stay in a caller-provided workspace.  The step's launch
// sequence is captured once and replayed as a CUDA graph (launch_train_step at the end of this file).  Only the dynamics chain is sequential in t; the reward and Q heads
// of all steps, the TD targets of all steps and the policy update over all H + 1 latents run as single batched passes.
//
// Parity: tests/test_gpu_train.py against oracle/tdmpc_oracle.py::update_step (torch autograd over the oracle's forward) and the
// fixtures the unmodified TDMPC.update produced (tests/golden/train_*.npz).
#include <cmath>
#include <cstdint>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ GEMM
// C[M,N] = op(A) op(B) (+ beta C) (+ bias[n]) (ELU).  A: [M,K] row-major (TA = false) or [K,M] (TA = true); B: [K,N] (TB = false) or
// [N,K] (TB = true: nn.Linear weights).  BM x 64 x 16 tiles (BM = 64 or 128), 256 threads, (BM / 16) x 4 outputs per thread.
constexpr int GB = 64, GK = 16;

template <bool TA, bool TB, int BM>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                    float* __restrict__ C, int ldc, float beta, const float* __restrict__ bias, int elu, int kslice) {
  // split-k (kslice > 0): block z multiplies the k range [z * kslice, (z + 1) * kslice) into its own partial C + z * M * ldc
  const int kbeg = kslice > 0 ? blockIdx.z * kslice : 0;
  if (kslice > 0) { K = min(K, kbeg + kslice); C += (size_t)blockIdx.z * M * ldc; }
  constexpr int RM = BM / 16;                                // rows per thread: 4 (64-row tile) or 8 (128-row tile)
  __shared__ __align__(16) float As[GK][BM + 4];
  __shared__ __align__(16) float Bs[GK][GB + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * GB;
  float acc[RM][4];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  // global -> registers one k-tile ahead of the shared-memory tile being multiplied (the loads fly during the FMAs)
  float ra[RM], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      const int idx = tid + r * 256;
      int m, k;
      if (TA) { k = idx / BM; m = idx % BM; } else { m = idx >> 4; k = idx & 15; }
      const int gm = m0 + m, gk = k0 + k;
      ra[r] = (gm < M && gk < K) ? (TA ? __ldg(A + (size_t)gk * lda + gm) : __ldg(A + (size_t)gm * lda + gk)) : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;
      int n, kb;
      if (TB) { n = idx >> 4; kb = idx & 15; } else { kb = idx >> 6; n = idx & 63; }
      const int gn = n0 + n, gkb = k0 + kb;
      rb[r] = (gn < N && gkb < K) ? (TB ? __ldg(B + (size_t)gn * ldb + gkb) : __ldg(B + (size_t)gkb * ldb + gn)) : 0.0f;
    }
  };
  fetch(kbeg);
  for (int k0 = kbeg; k0 < K; k0 += GK) {
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      const int idx = tid + r * 256;
      if (TA) As[idx / BM][idx % BM] = ra[r]; else As[idx & 15][idx >> 4] = ra[r];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;
      if (TB) Bs[idx & 15][idx >> 4] = rb[r]; else Bs[idx >> 6][idx & 63] = rb[r];
    }
    __syncthreads();
    if (k0 + GK < K) fetch(k0 + GK);
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float av[RM];
#pragma unroll
      for (int i4 = 0; i4 < RM / 4; ++i4) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * RM + i4 * 4]);
        av[i4 * 4] = a.x; av[i4 * 4 + 1] = a.y; av[i4 * 4 + 2] = a.z; av[i4 * 4 + 3] = a.w;
      }
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int gm = m0 + ty * RM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + gn);
      if (beta != 0.0f) v += beta * C[(size_t)gm * ldc + gn];
      if (elu) v = elu1(v);
      C[(size_t)gm * ldc + gn] = v;
    }
  }
}

constexpr size_t SPLITK_FLOATS = (size_t)4 << 20;      // 16 MB: up to 16 partials of a 512 x 512 weight gradient

struct Ctx {
  cudaStream_t st;
  float* splitk = nullptr;        // partials of the split-k weight-gradient GEMMs
  size_t splitk_floats = 0;
  int launches = 0;
  bool ok = true;
  int gemm_mode = 1;              // 1: split-bf16 tcgen05 GEMMs (gemm_tc.cuh), 0: the SIMT kernel only (TDMPC_TRAIN_GEMM=simt)
  int tc_bn = 128;                // tile width of the tensor-core GEMM for N > 64 (TDMPC_TRAIN_BN=64: twice the tiles, twice the A re-reads)
  void check() { if (cudaGetLastError() != cudaSuccess) ok = false; ++launches; }
};

template <int BM>
void gemm_bm(Ctx& c, bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, float beta,
             const float* bias, bool elu, int kslice = 0, int nz = 1) {
  dim3 grid((N + GB - 1) / GB, (M + BM - 1) / BM, nz);
  if (ta && tb) sgemm_kernel<true, true, BM><<<grid, 256, 0, c.st>>>(M, N, K, A, lda, B, ldb, C, ldc, beta, bias, elu ? 1 : 0, kslice);
  else if (ta) sgemm_kernel<true, false, BM><<<grid, 256, 0, c.st>>>(M, N, K, A, lda, B, ldb, C, ldc, beta, bias, elu ? 1 : 0, kslice);
  else if (tb) sgemm_kernel<false, true, BM><<<grid, 256, 0, c.st>>>(M, N, K, A, lda, B, ldb, C, ldc, beta, bias, elu ? 1 : 0, kslice);
  else sgemm_kernel<false, false, BM><<<grid, 256, 0, c.st>>>(M, N, K, A, lda, B, ldb, C, ldc, beta, bias, elu ? 1 : 0, kslice);
  c.check();
}
// C[i] = beta * C[i] + sum_z P[z][i]  (fixed order: the split-k result does not depend on scheduling)
__global__ void splitk_reduce_kernel(const float* __restrict__ P, int nz, size_t n, float* __restrict__ C, float beta,
                                     const float* __restrict__ bias = nullptr, int N = 1, int elu = 0, size_t zstride = 0,
                                     float* __restrict__ db = nullptr, int db_n = 0, int db_acc = 0) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (zstride == 0) zstride = n;
  if (i >= n + (size_t)db_n) return;
  float s = 0.0f;
  for (int z = 0; z < nz; ++z) s += P[(size_t)z * zstride + i];
  if (i >= n) {                                                      // the row sums that rode along with a weight-gradient GEMM: db
    const size_t j = i - n;
    db[j] = db_acc ? db[j] + s : s;
    return;
  }
  if (bias) s += __ldg(bias + (int)(i % (size_t)N));               // same order as the GEMM epilogues: bias, beta C, ELU
  if (beta != 0.0f) s += beta * C[i];
  if (elu) s = elu1(s);
  C[i] = s;
}
// db / db_acc / db_done: optional bias gradient db[m] (+)= sum_k op(A)[m][k] of a weight-gradient product (ta && !tb); *db_done says
// whether this call produced it (only the split-k tensor-core path does; the caller runs colsum otherwise)
void gemm(Ctx& c, bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, float beta = 0.0f,
          const float* bias = nullptr, bool elu = false, float* db = nullptr, bool db_acc = false, bool* db_done = nullptr) {
  // Tensor-core path (gemm_tc.cuh): every GEMM with a tile's worth of rows and columns; the few-row / few-column products (the last
  // Linear of the reward and Q heads, the policy's A-wide output) stay on the SIMT kernel.
  if (c.gemm_mode == 1 && !(ta && tb) && M >= 32 && N >= 16 && K >= 16) {
    const int BN = N <= 64 ? 64 : c.tc_bn;
    const long long tiles = (long long)((M + gemm_tc::BM - 1) / gemm_tc::BM) * ((N + BN - 1) / BN);
    int nz = 1, kslice = 0;
    float* out = C; int ldo = ldc; float b = beta;
    // One CTA per SM and a serial chain of k-blocks inside it: a product of a few tiles (the weight gradients: small M x N, the
    // batch as a long k; the per-step Linears of the dynamics chain: 512 rows) is cut along k into slices that fill one wave, each
    // slice writes a partial, and a reduction adds them in a fixed order (and applies bias / beta / ELU).
    if (c.splitk && K >= 2 * gemm_tc::BK && tiles * 2 <= 148 && ldc == N && (size_t)M * N <= c.splitk_floats / 2) {
      nz = (int)(148 / tiles);
      const int maxz = (int)(c.splitk_floats / ((size_t)M * N));
      if (nz > maxz) nz = maxz;
      if (nz > 32) nz = 32;
      kslice = ((K + nz - 1) / nz + gemm_tc::BK - 1) / gemm_tc::BK * gemm_tc::BK;
      nz = (K + kslice - 1) / kslice;
      if (nz >= 2) { out = c.splitk; ldo = N; b = 0.0f; } else { nz = 1; kslice = 0; }
    }
    const bool part = nz >= 2;
    const size_t mn = (size_t)M * N;
    const bool fuse = part && db && ta && !tb && (size_t)nz * (mn + M) <= c.splitk_floats;
    const long long zs = part ? (long long)(mn + (fuse ? M : 0)) : 0;
    float* bsum = fuse ? c.splitk + mn : nullptr;
    const bool done = BN == 128 ? gemm_tc::launch_bn<128>(ta, tb, M, N, K, A, lda, B, ldb, out, ldo, b, part ? nullptr : bias, part ? false : elu, kslice, nz, c.st, zs, bsum)
                                : gemm_tc::launch_bn<64>(ta, tb, M, N, K, A, lda, B, ldb, out, ldo, b, part ? nullptr : bias, part ? false : elu, kslice, nz, c.st, zs, bsum);
    if (done) {
      c.check();
      if (part) {
        const size_t tot = mn + (fuse ? M : 0);
        splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c.st>>>(c.splitk, nz, mn, C, beta, bias, N, elu ? 1 : 0, (size_t)zs, fuse ? db : nullptr,
                                                                               fuse ? M : 0, db_acc ? 1 : 0);
        c.check();
        if (fuse && db_done) *db_done = true;
      }
      return;
    }
  }
  // Weight gradients (dW = dY^T X: small M x N, the batch as a long k): a few tiles would leave most SMs idle for 200+ us, so the
  // k range is cut into slices that fill the GPU, each slice writes a partial, and a reduction adds them in a fixed order.
  const long long tiles64 = (long long)((M + 63) / 64) * ((N + GB - 1) / GB);
  if (ta && !tb && !bias && !elu && c.splitk && K >= 512 && tiles64 < 120 && ldc == N && (size_t)M * N <= c.splitk_floats / 2) {
    int nz = (int)((148 + tiles64 - 1) / tiles64);
    const int maxz = (int)(c.splitk_floats / ((size_t)M * N));
    if (nz > maxz) nz = maxz;
    if (nz > 32) nz = 32;
    int kslice = ((K + nz - 1) / nz + GK - 1) / GK * GK;
    nz = (K + kslice - 1) / kslice;
    if (nz >= 2) {
      gemm_bm<64>(c, ta, tb, M, N, K, A, lda, B, ldb, c.splitk, N, 0.0f, nullptr, false, kslice, nz);
      const size_t n = (size_t)M * N;
      splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.st>>>(c.splitk, nz, n, C, beta);
      c.check();
      return;
    }
  }
  // 128-row tiles (8 x 4 outputs per thread) once they still fill the 148 SMs, 64-row tiles otherwise
  const long long tiles128 = (long long)((M + 127) / 128) * ((N + GB - 1) / GB);
  if (tiles128 >= 296) gemm_bm<128>(c, ta, tb, M, N, K, A, lda, B, ldb, C, ldc, beta, bias, elu);
  else gemm_bm<64>(c, ta, tb, M, N, K, A, lda, B, ldb, C, ldc, beta, bias, elu);
}
// nn.Linear forward over `rows`: Y = X W^T + b (ELU)
void linear_fwd(Ctx& c, int rows, int N, int K, const float* X, int ldx, const float* W, const float* b, float* Y, int ldy, bool elu) {
  gemm(c, false, true, rows, N, K, X, ldx, W, K, Y, ldy, 0.0f, b, elu);
}

// out[n] (+)= sum_r X[r][n]   (bias gradients, LayerNorm beta gradients)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int rows, int N, int ld, float* __restrict__ out, int accumulate) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s = 0.0f;
  if (n < N)
    for (int r = ty; r < rows; r += 8) s += X[(size_t)r * ld + n];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    out[n] = accumulate ? out[n] + t : t;
  }
}
// row slice blockIdx.y of the column sums -> P[blockIdx.y][n]
__global__ void __launch_bounds__(256) colsum_split_kernel(const float* __restrict__ X, int rows, int rs, int N, int ld, float* __restrict__ P) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rs, r1 = min(rows, r0 + rs);
  float s = 0.0f;
  if (n < N)
    for (int r = r0 + ty; r < r1; r += 8) s += X[(size_t)r * ld + n];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    P[(size_t)blockIdx.y * N + n] = t;
  }
}
void colsum(Ctx& c, const float* X, int rows, int N, int ld, float* out, bool accumulate) {
  // many rows, few columns: row slices into partials, then the fixed-order reduction (as the split-k GEMM)
  const int nz = rows >= 1024 && c.splitk ? 16 : 1;
  if (nz > 1 && (size_t)nz * N <= c.splitk_floats) {
    const int rs = (rows + nz - 1) / nz;
    colsum_split_kernel<<<dim3((N + 31) / 32, nz), 256, 0, c.st>>>(X, rows, rs, N, ld, c.splitk);
    c.check();
    splitk_reduce_kernel<<<(N + 255) / 256, 256, 0, c.st>>>(c.splitk, nz, (size_t)N, out, accumulate ? 1.0f : 0.0f);
    c.check();
    return;
  }
  colsum_kernel<<<(N + 31) / 32, 256, 0, c.st>>>(X, rows, N, ld, out, accumulate ? 1 : 0);
  c.check();
}
// out[n] (+)= sum_r X[r][n] * Y[r][n]   (LayerNorm gamma gradients)
__global__ void __launch_bounds__(256) colsum_prod_kernel(const float* __restrict__ X, const float* __restrict__ Y, int rows, int N, float* __restrict__ out, int accumulate) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s = 0.0f;
  if (n < N)
    for (int r = ty; r < rows; r += 8) s = fmaf(X[(size_t)r * N + n], Y[(size_t)r * N + n], s);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    out[n] = accumulate ? out[n] + t : t;
  }
}

// row slice blockIdx.y of the column sums of X * Y -> P[blockIdx.y][n]
__global__ void __launch_bounds__(256) colsum_prod_split_kernel(const float* __restrict__ X, const float* __restrict__ Y, int rows, int rs, int N, float* __restrict__ P) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rs, r1 = min(rows, r0 + rs);
  float s = 0.0f;
  if (n < N)
    for (int r = r0 + ty; r < r1; r += 8) s = fmaf(X[(size_t)r * N + n], Y[(size_t)r * N + n], s);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    P[(size_t)blockIdx.y * N + n] = t;
  }
}
// LayerNorm gamma gradients: 16 blocks of one slice each left the GPU idle for 83 us per call; row slices + the fixed-order reduction
void colsum_prod(Ctx& c, const float* X, const float* Y, int rows, int N, float* out) {
  const int nz = rows >= 1024 && c.splitk ? 16 : 1;
  if (nz > 1 && (size_t)nz * N <= c.splitk_floats) {
    const int rs = (rows + nz - 1) / nz;
    colsum_prod_split_kernel<<<dim3((N + 31) / 32, nz), 256, 0, c.st>>>(X, Y, rows, rs, N, c.splitk);
    c.check();
    splitk_reduce_kernel<<<(N + 255) / 256, 256, 0, c.st>>>(c.splitk, nz, (size_t)N, out, 0.0f);
    c.check();
    return;
  }
  colsum_prod_kernel<<<(N + 31) / 32, 256, 0, c.st>>>(X, Y, rows, N, out, 0);
  c.check();
}

// nn.Linear backward: dW (+)= dY^T X, db (+)= colsum(dY), dX = dY W (optional, (+)= with beta_x)
void linear_bwd(Ctx& c, int rows, int N, int K, const float* dY, int ldy, const float* X, int ldx, const float* W, float* dW, float* db, bool accumulate,
                float* dX, int lddx, float beta_x = 0.0f) {
  bool db_done = false;
  if (dW) gemm(c, true, false, N, K, rows, dY, ldy, X, ldx, dW, K, accumulate ? 1.0f : 0.0f, nullptr, false, db, accumulate, &db_done);
  if (db && !db_done) colsum(c, dY, rows, N, ldy, db, accumulate);
  if (dX) gemm(c, false, false, rows, K, N, dY, ldy, W, K, dX, lddx, beta_x);
}

// ------------------------------------------------------------------------------------------------ elementwise
// dX *= ELU'(y) with y the ELU OUTPUT (y > 0 ? 1 : y + 1), or tanh'(y) = 1 - y^2
__global__ void act_bwd_kernel(float* __restrict__ d, const float* __restrict__ y, size_t n, int tanh_act) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = y[i];
  d[i] *= tanh_act ? (1.0f - v * v) : (v > 0.0f ? 1.0f : v + 1.0f);
}
void act_bwd(Ctx& c, float* d, const float* y, size_t n, bool tanh_act) {
  act_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.st>>>(d, y, n, tanh_act ? 1 : 0);
  c.check();
}

// x[r] = [z[r] | a[r]]
__global__ void concat_kernel(const float* __restrict__ z, const float* __restrict__ a, int rows, int L, int A, float* __restrict__ x) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int W = L + A;
  if (i >= (size_t)rows * W) return;
  const size_t r = i / W; const int k = (int)(i - r * W);
  x[i] = k < L ? z[r * L + k] : a[r * A + (k - L)];
}
void concat(Ctx& c, const float* z, const float* a, int rows, int L, int A, float* x) {
  const size_t n = (size_t)rows * (L + A);
  concat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.st>>>(z, a, rows, L, A, x);
  c.check();
}

// nn.LayerNorm(N) over a row (biased variance, eps 1e-5) + tanh / ELU.  In place: pre -> xhat; y = act(xhat * g + b); rstd[r] kept.
// Warp per row.
__global__ void __launch_bounds__(256) ln_act_fwd_kernel(float* __restrict__ x, int rows, int N, const float* __restrict__ g, const float* __restrict__ b,
                                                         float* __restrict__ y, float* __restrict__ rstd_out, int tanh_act) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float* xr = x + (size_t)warp * N;
  float s = 0.0f;
  for (int k = lane; k < N; k += 32) s += xr[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / (float)N;
  float v = 0.0f;
  for (int k = lane; k < N; k += 32) { const float d = xr[k] - mu; v = fmaf(d, d, v); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = 1.0f / sqrtf(v / (float)N + 1e-5f);
  for (int k = lane; k < N; k += 32) {
    const float xh = (xr[k] - mu) * rstd;
    xr[k] = xh;
    const float o = xh * __ldg(g + k) + __ldg(b + k);
    y[(size_t)warp * N + k] = tanh_act ? tanhf(o) : elu1(o);
  }
  if (lane == 0 && rstd_out) rstd_out[warp] = rstd;
}
void ln_act_fwd(Ctx& c, float* x, int rows, int N, const float* g, const float* b, float* y, float* rstd, bool tanh_act) {
  ln_act_fwd_kernel<<<(rows * 32 + 255) / 256, 256, 0, c.st>>>(x, rows, N, g, b, y, rstd, tanh_act ? 1 : 0);
  c.check();
}
// Backward of y = act(LN(x)): in: d = dL/dy (overwritten with dL/dx), y, xhat, rstd, gamma.  dout (= dL/d(LN output), needed for the
// gamma / beta gradients) is written to `dln` when given.
__global__ void __launch_bounds__(256) ln_act_bwd_kernel(float* __restrict__ d, const float* __restrict__ y, const float* __restrict__ xhat,
                                                         const float* __restrict__ rstd, const float* __restrict__ g, int rows, int N,
                                                         float* __restrict__ dln, int tanh_act) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const size_t base = (size_t)warp * N;
  float s1 = 0.0f, s2 = 0.0f;
  for (int k = lane; k < N; k += 32) {
    const float yv = y[base + k];
    const float dl = d[base + k] * (tanh_act ? (1.0f - yv * yv) : (yv > 0.0f ? 1.0f : yv + 1.0f));     // dL/d(LN output)
    if (dln) dln[base + k] = dl;
    const float dxh = dl * __ldg(g + k);
    d[base + k] = dxh;
    s1 += dxh; s2 = fmaf(dxh, xhat[base + k], s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  const float m1 = s1 / (float)N, m2 = s2 / (float)N, rs = rstd[warp];
  for (int k = lane; k < N; k += 32) d[base + k] = rs * (d[base + k] - m1 - xhat[base + k] * m2);
}

// pi(): a = clamp(tanh(pre) + clip(eps * std, +-0.3), -1 + 1e-6, 1 - 1e-6); mu = tanh(pre) kept for the backward pass (helper.py:86-96)
__global__ void pi_sample_kernel(const float* __restrict__ pre, size_t n, float min_std, NoiseSrc nz, float* __restrict__ mu_out, float* __restrict__ a_out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float mu = tanhf(pre[i]);
  if (mu_out) mu_out[i] = mu;
  a_out[i] = policy_sample(mu, min_std, min_std > 0.0f ? noise_at(nz, i) : 0.0f);
}

// N = 1 heads: out[r] = x[r] . w + b   (warp per row)
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ x, int rows, int K, const float* __restrict__ w, const float* __restrict__ b,
                                                     float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float s = 0.0f;
  for (int k = lane; k < K; k += 32) s = fmaf(x[(size_t)warp * K + k], __ldg(w + k), s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[warp] = s + __ldg(b);
}
void rowdot(Ctx& c, const float* x, int rows, int K, const float* w, const float* b, float* out) {
  rowdot_kernel<<<(rows * 32 + 255) / 256, 256, 0, c.st>>>(x, rows, K, w, b, out);
  c.check();
}
// dX[r][k] = dout[r] * w[k]
__global__ void outer_kernel(const float* __restrict__ dout, const float* __restrict__ w, int rows, int K, float* __restrict__ dX) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * K) return;
  const size_t r = i / K;
  dX[i] = dout[r] * __ldg(w + (i - r * K));
}

// ------------------------------------------------------------------------------------------------ losses
struct LossParams {
  int B, H, L;
  const float *zs, *next_z, *rp, *q1, *q2, *tq1, *tq2, *reward, *weights;
  float discount, c_cons, c_rew, c_val;
  float rho_pow[TDMPC_MAX_H];
  float* td;          // [H*B] out
  float *d_z, *d_rp, *d_q1, *d_q2;     // gradient seeds: [H*B*L] (w.r.t. z_{t+1}), [H*B] x 3
  float* per_sample;  // [5*B]: consistency, reward, value, priority (clamped), total*weight
  float* priorities;
};
// one WARP per sample (lanes over the latent dimension): the loss sums of tdmpc.py:205-219 and the gradient of
// (total * weights).mean() / H w.r.t. every prediction.  (One thread per sample walked H x L strided floats twice: 70 us at batch 512.)
__global__ void __launch_bounds__(256) loss_kernel(LossParams p) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= p.B) return;
  float cons = 0.0f, rew = 0.0f, val = 0.0f, pri = 0.0f;
  for (int t = 0; t < p.H; ++t) {
    const size_t r = (size_t)t * p.B + b;
    const float rt = p.reward[r];
    const float td = __fadd_rn(rt, __fmul_rn(p.discount, fminf(p.tq1[r], p.tq2[r])));
    if (lane == 0) p.td[r] = td;
    const float* z = p.zs + ((size_t)(t + 1) * p.B + b) * p.L;
    const float* nz = p.next_z + r * p.L;
    float ms = 0.0f;
    for (int l = lane; l < p.L; l += 32) { const float d = z[l] - nz[l]; ms = fmaf(d, d, ms); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ms += __shfl_xor_sync(0xffffffffu, ms, o);      // xor butterfly: every lane holds the same sum
    const float d1 = p.q1[r] - td, d2 = p.q2[r] - td, dr = p.rp[r] - rt, rho = p.rho_pow[t];
    cons = __fadd_rn(cons, __fmul_rn(rho, ms / (float)p.L));
    rew = __fadd_rn(rew, __fmul_rn(rho, dr * dr));
    val = __fadd_rn(val, __fmul_rn(rho, __fadd_rn(d1 * d1, d2 * d2)));
    pri = __fadd_rn(pri, __fmul_rn(rho, __fadd_rn(fabsf(d1), fabsf(d2))));
  }
  const float total = __fadd_rn(__fadd_rn(__fmul_rn(p.c_cons, fminf(cons, 1e4f)), __fmul_rn(p.c_rew, fminf(rew, 1e4f))), __fmul_rn(p.c_val, fminf(val, 1e4f)));
  const float w = p.weights[b];
  if (lane == 0) {
    p.per_sample[b] = cons; p.per_sample[p.B + b] = rew; p.per_sample[2 * p.B + b] = val; p.per_sample[3 * p.B + b] = total; p.per_sample[4 * p.B + b] = total * w;
    p.priorities[b] = fminf(pri, 1e4f);
  }
  // d weighted_loss / d total[b] = w / B, times the 1/H hook (tdmpc.py:224); clamp(max) passes the gradient where x <= max
  const float s = w / (float)p.B / (float)p.H;
  const float gc = cons <= 1e4f ? s * p.c_cons : 0.0f, gr = rew <= 1e4f ? s * p.c_rew : 0.0f, gv = val <= 1e4f ? s * p.c_val : 0.0f;
  for (int t = 0; t < p.H; ++t) {
    const size_t r = (size_t)t * p.B + b;
    const float rho = p.rho_pow[t];
    const float td = __fadd_rn(p.reward[r], __fmul_rn(p.discount, fminf(p.tq1[r], p.tq2[r])));
    if (lane == 0) {
      p.d_rp[r] = gr * rho * 2.0f * (p.rp[r] - p.reward[r]);
      p.d_q1[r] = gv * rho * 2.0f * (p.q1[r] - td);
      p.d_q2[r] = gv * rho * 2.0f * (p.q2[r] - td);
    }
    const float* z = p.zs + ((size_t)(t + 1) * p.B + b) * p.L;
    const float* nz = p.next_z + r * p.L;
    float* dz = p.d_z + r * p.L;
    const float k = gc * rho * 2.0f / (float)p.L;
    for (int l = lane; l < p.L; l += 32) dz[l] = k * (z[l] - nz[l]);
  }
}

// out[i] = mean over n of x[i*n .. ]   (single block; metrics)
__global__ void __launch_bounds__(256) means_kernel(const float* __restrict__ x, int n, int count, float* __restrict__ out) {
  __shared__ double red[256];
  for (int i = 0; i < count; ++i) {
    double s = 0.0;
    for (int k = threadIdx.x; k < n; k += 256) s += (double)x[(size_t)i * n + k];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) out[i] = (float)(red[0] / (double)n);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ optimiser
struct Seg { float* p; float* pt; size_t off, n; };      // parameter tensor, its target twin, offset / size in the flat arenas
constexpr int MAX_SEGS = 48;
struct SegTable { Seg s[MAX_SEGS]; int count; };

// sum of squares of the flat gradient range [lo, hi) -> acc (double, atomics)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, size_t lo, size_t hi, double* __restrict__ acc) {
  __shared__ double red[256];
  double s = 0.0;
  for (size_t i = lo + (size_t)blockIdx.x * 256 + threadIdx.x; i < hi; i += (size_t)gridDim.x * 256) s += (double)g[i] * (double)g[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) atomicAdd(acc, red[0]);
}
// clip_grad_norm_: total_norm = sqrt(acc0 (+ acc1)); coef = min(1, max_norm / (total_norm + 1e-6)); norm -> metrics slot
__global__ void clip_coef_kernel(const double* acc, int nacc, float max_norm, float* coef, float* norm_out) {
  double s = 0.0;
  for (int i = 0; i < nacc; ++i) s += acc[i];
  const float norm = (float)sqrt(s);
  *norm_out = norm;
  const float c = max_norm / (norm + 1e-6f);
  *coef = c < 1.0f ? c : 1.0f;
}
// torch.optim.Adam (single-tensor formulas, no amsgrad / weight decay) over the segments [seg_lo, seg_hi) of the table; gradients are
// scaled by *coef (and stored scaled, as clip_grad_norm_ does in place).  State arenas are indexed by (off - state_base).
struct AdamParams {
  float beta1, beta2, eps;
  const float* dyn;                                    // device: {step_size = lr / (1 - beta1^step), sqrt(1 - beta2^step)}, formed in double on the host
  float* grads; float* m; float* v; size_t state_base; // like the reference; in device memory so that a captured graph can be replayed
  const float* coef;
};
__global__ void __launch_bounds__(256) adam_kernel(SegTable tab, int seg_lo, int seg_hi, AdamParams a) {
  const float coef = *a.coef;
  const float step_size = __ldg(a.dyn), bc2_sqrt = __ldg(a.dyn + 1);
  for (int si = seg_lo; si < seg_hi; ++si) {
    const Seg sg = tab.s[si];
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < sg.n; i += (size_t)gridDim.x * 256) {
      const float g = a.grads[sg.off + i] * coef;
      a.grads[sg.off + i] = g;
      const size_t k = sg.off - a.state_base + i;
      const float m = a.m[k] + (1.0f - a.beta1) * (g - a.m[k]);            // exp_avg.lerp_(grad, 1 - beta1)
      const float v = a.beta2 * a.v[k] + (1.0f - a.beta2) * g * g;          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
      a.m[k] = m; a.v[k] = v;
      const float denom = sqrtf(v) / bc2_sqrt + a.eps;
      sg.p[i] = sg.p[i] - step_size * (m / denom);                          // param.addcdiv_(exp_avg, denom, value = -step_size)
    }
  }
}
// helper.ema: p_target.lerp_(p, tau)
__global__ void __launch_bounds__(256) ema_kernel(SegTable tab, float tau) {
  for (int si = 0; si < tab.count; ++si) {
    const Seg sg = tab.s[si];
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < sg.n; i += (size_t)gridDim.x * 256) {
      const float t = sg.pt[i];
      sg.pt[i] = t + tau * (sg.p[i] - t);
    }
  }
}

// min(Q1, Q2) of the policy update and its routing: pi_loss terms and dL/dQ1, dL/dQ2 (torch.minimum: ties split evenly)
__global__ void pi_loss_kernel(const float* __restrict__ q1, const float* __restrict__ q2, int B, int T, const float* __restrict__ rho_pow,
                               float* __restrict__ dq1, float* __restrict__ dq2, float* __restrict__ qmin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int t = i / B;
  const float a = q1[i], b = q2[i];
  const float g = -rho_pow[t] / (float)B;                       // pi_loss += -Q.mean() * rho^t
  qmin[i] = -rho_pow[t] * fminf(a, b);
  dq1[i] = a < b ? g : (a == b ? 0.5f * g : 0.0f);
  dq2[i] = b < a ? g : (a == b ? 0.5f * g : 0.0f);
}
// pi_loss = sum_t mean_b(-rho^t Qmin)  -> single value
__global__ void __launch_bounds__(256) pi_loss_sum_kernel(const float* __restrict__ qmin, int B, int T, float* __restrict__ out) {
  __shared__ double red[256];
  double tot = 0.0;
  for (int t = 0; t < T; ++t) {
    double s = 0.0;
    for (int k = threadIdx.x; k < B; k += 256) s += (double)qmin[(size_t)t * B + k];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    tot += red[0] / (double)B;
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)tot;
}
// da[r][a] = dx[r][L + a] * (1 - mu^2)   (straight-through clamp of TruncatedNormal.sample, then tanh)
__global__ void pi_dpre_kernel(const float* __restrict__ dx, const float* __restrict__ mu, int rows, int L, int A, float* __restrict__ dpre) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * A) return;
  const size_t r = i / A; const int a = (int)(i - r * A);
  const float m = mu[i];
  dpre[i] = dx[r * (L + A) + L + a] * (1.0f - m * m);
}
// dz[b][l] (+)= dx[b][l]  (the latent part of a [z, a] gradient)
__global__ void take_z_kernel(const float* __restrict__ dx, int rows, int L, int A, float* __restrict__ dz, int accumulate) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * L) return;
  const size_t r = i / L; const int l = (int)(i - r * L);
  const float v = dx[r * (L + A) + l];
  dz[i] = accumulate ? dz[i] + v : v;
}
__global__ void add_kernel(float* __restrict__ d, const float* __restrict__ s, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] += s[i];
}

inline unsigned nblk(size_t n) { return (unsigned)((n + 255) / 256); }

// parameter table in state_dict order
struct Mlp3 { float* w[3]; float* b[3]; };
struct QNet { float* w[3]; float* b[3]; float* g[2]; float* e[2]; };
struct Net { float* ew[2]; float* eb[2]; Mlp3 dyn, rew, pi; QNet q[2]; };

void net_from(const tdmpc_weights& w, Net& n) {
  auto mut = [](const float* p) { return const_cast<float*>(p); };
  for (int i = 0; i < 2; ++i) { n.ew[i] = mut(w.enc_w[i]); n.eb[i] = mut(w.enc_b[i]); }
  for (int i = 0; i < 3; ++i) {
    n.dyn.w[i] = mut(w.dyn_w[i]); n.dyn.b[i] = mut(w.dyn_b[i]);
    n.rew.w[i] = mut(w.rew_w[i]); n.rew.b[i] = mut(w.rew_b[i]);
    n.pi.w[i] = mut(w.pi_w[i]); n.pi.b[i] = mut(w.pi_b[i]);
    for (int h = 0; h < 2; ++h) { n.q[h].w[i] = mut(w.q_w[h][i]); n.q[h].b[i] = mut(w.q_b[h][i]); }
  }
  for (int h = 0; h < 2; ++h)
    for (int i = 0; i < 2; ++i) { n.q[h].g[i] = mut(w.q_ln_g[h][i]); n.q[h].e[i] = mut(w.q_ln_b[h][i]); }
}

}  // namespace

// number of parameters, and the [begin, end) range of the _pi block, in the flat state_dict order
long long train_param_layout(const tdmpc_model_cfg& c, long long* pi_begin, long long* pi_end) {
  const long long D = c.obs_dim, A = c.action_dim, L = c.latent_dim, M = c.mlp_dim, E = c.enc_dim;
  long long n = E * D + E + L * E + L;
  n += 2 * (M * (L + A) + M + M * M + M) + (L * M + L) + (M + 1);       // dynamics, reward
  if (pi_begin) *pi_begin = n;
  n += M * L + M + M * M + M + A * M + A;
  if (pi_end) *pi_end = n;
  n += 2 * (M * (L + A) + M + 2 * M + M * M + M + 2 * M + M + 1);
  return n;
}

static long long train_core_bytes(const tdmpc_model_cfg& c, int B, int H);
// Front of the workspace: the DynBlock, then fixed staging slots for the step's inputs and the priorities (graph mode: the captured
// kernels read / write these, the caller's tensors -- new addresses every step -- are copied in and out around the replay).
struct FrontLayout { size_t dyn, pri, obs, nxt, act, rew, wts, n1, n2, total; };     // offsets in floats
static FrontLayout front_layout(const tdmpc_model_cfg& c, int B, int H) {
  const size_t D = c.obs_dim, A = c.action_dim, R = (size_t)B * H, R1 = (size_t)B * (H + 1);
  FrontLayout f{};
  size_t o = 0;
  auto take = [&](size_t n) { const size_t at = o; o += (n + 63) / 64 * 64; return at; };     // 256-byte aligned slots
  f.dyn = take(64); f.pri = take(B); f.obs = take((size_t)B * D); f.nxt = take(R * D); f.act = take(R * A); f.rew = take(R); f.wts = take(B);
  f.n1 = take(R * A); f.n2 = take(R1 * A);
  f.total = o;
  return f;
}

long long train_workspace_bytes(const tdmpc_model_cfg& c, int B, int H) {
  return train_core_bytes(c, B, H) + (long long)(front_layout(c, B, H).total * sizeof(float));
}

static long long train_core_bytes(const tdmpc_model_cfg& c, int B, int H) {
  const size_t M = c.mlp_dim, L = c.latent_dim, A = c.action_dim, E = c.enc_dim, W = L + A;
  const size_t R = (size_t)B * H, R1 = (size_t)B * (H + 1);
  size_t f = 0;
  f += R1 * E + R1 * L;                // e1 (encoder hidden; reused per pass), zs
  f += R1 * W + 2 * R * M;             // x, d1, d2
  f += 2 * R * M + R;                  // r1, r2, rp
  f += 2 * (4 * R1 * M + 3 * R1);      // per Q head: xhat_a, qa, xhat_b, qb, rstd_a, rstd_b, qv
  f += R * L + R * L + R1 * M * 2 + 2 * R1 * A + R1 * W + 2 * R;      // next_z, nzo, p1, p2, pa, a_pi, xt, tq1, tq2
  f += R + R * L + 3 * R;              // td, d_z, d_rp, d_q1, d_q2
  f += R1 * W + 2 * R1 * M + R1 * M;   // dx, dh1, dh2, dln
  f += 2 * (size_t)B * L + (size_t)B * W + (size_t)B * E;   // dz carry, dzn, dxd, de1
  f += 5 * (size_t)B + 3 * R1 + R1 * A + 64;                 // per_sample, pi-loss scratch (dq1, dq2, qmin), dpre, control words
  f += SPLITK_FLOATS;                                         // partials of the split-k weight-gradient GEMMs / column sums
  return (long long)(f * sizeof(float) + 256);
}

// The values that change from one step to the next, in device memory (written by one small H2D copy before the step), so that the
// launch sequence itself is the same every step and can be replayed as a CUDA graph.
struct DynBlock {
  uint64_t call;                      // Philox call counter (NoiseSrc::call_dev)
  float adam[4];                      // model optimiser {step_size, sqrt(1 - beta2^step)}, policy optimiser {same}
  float rho[TDMPC_MAX_H + 1];         // rho^t (update_pi)
};
static_assert(sizeof(DynBlock) <= 64 * sizeof(float), "DynBlock outgrew its workspace slot");

// returns launches (> 0) or -1
static int train_step_body(const tdmpc_model_cfg& cfg, const tdmpc_train_args& a, cudaStream_t st, const DynBlock* dyn) {
  const int B = a.batch, H = a.horizon, D = cfg.obs_dim, A = cfg.action_dim, L = cfg.latent_dim, M = cfg.mlp_dim, E = cfg.enc_dim, W = L + A;
  const int R = B * H, R1 = B * (H + 1);
  Ctx c; c.st = st;
  {
    static const int mode = [] { const char* e = getenv("TDMPC_TRAIN_GEMM"); return (e && e[0] == 's') ? 0 : 1; }();
    static const int bn = [] { const char* e = getenv("TDMPC_TRAIN_BN"); return (e && atoi(e) == 64) ? 64 : 128; }();
    c.gemm_mode = mode; c.tc_bn = bn;
  }
  Net on, tg;
  net_from(a.online, on);
  net_from(a.target, tg);

  // ---- workspace carve-up
  float* ws = reinterpret_cast<float*>(a.workspace);
  auto take = [&](size_t n) { float* p = ws; ws += n; return p; };
  float* e1 = take((size_t)R1 * E);
  float* zs = take((size_t)R1 * L);
  float* x = take((size_t)R1 * W);
  float* d1 = take((size_t)R * M); float* d2 = take((size_t)R * M);
  float* r1 = take((size_t)R * M); float* r2 = take((size_t)R * M); float* rp = take(R);
  float *xa[2], *qa[2], *xb[2], *qb[2], *rsa[2], *rsb[2], *qv[2];
  for (int h = 0; h < 2; ++h) {
    xa[h] = take((size_t)R1 * M); qa[h] = take((size_t)R1 * M); xb[h] = take((size_t)R1 * M); qb[h] = take((size_t)R1 * M);
    rsa[h] = take(R1); rsb[h] = take(R1); qv[h] = take(R1);
  }
  float* next_z = take((size_t)R * L); float* nzo = take((size_t)R * L);
  float* p1 = take((size_t)R1 * M); float* p2 = take((size_t)R1 * M); float* pa = take((size_t)R1 * A); float* api = take((size_t)R1 * A);
  float* xt = take((size_t)R1 * W); float* tq1 = take(R); float* tq2 = take(R);
  float* td = take(R); float* d_z = take((size_t)R * L); float* d_rp = take(R); float* d_q1 = take(R); float* d_q2 = take(R);
  float* dx = take((size_t)R1 * W); float* dh1 = take((size_t)R1 * M); float* dh2 = take((size_t)R1 * M); float* dln = take((size_t)R1 * M);
  float* dzc = take((size_t)B * L); float* dzn = take((size_t)B * L); float* dxd = take((size_t)B * W); float* de1 = take((size_t)B * E);
  float* per_sample = take(5 * (size_t)B);
  float* pq1 = take(R1); float* pq2 = take(R1); float* qmin = take(R1); float* dpre = take((size_t)R1 * A);
  ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 15) & ~(uintptr_t)15);
  float* skp = take(SPLITK_FLOATS);
  c.splitk = skp; c.splitk_floats = SPLITK_FLOATS;
  float* ctl = take(64);                               // [0..3] doubles: sum-of-squares accumulators; [8] clip coef; [16..] rho powers
  if ((char*)ws - (char*)a.workspace > a.workspace_bytes) return -1;
  double* acc = reinterpret_cast<double*>(ctl);
  float* coef = ctl + 8;
  const float* rho_dev = dyn->rho;

  // ---- gradient arena offsets (state_dict order)
  long long pi_lo, pi_hi;
  const long long total = train_param_layout(cfg, &pi_lo, &pi_hi);
  float* G = a.grads;
  SegTable tab{}; tab.count = 0;
  size_t off = 0;
  auto seg = [&](float* p, float* pt, size_t n) { tab.s[tab.count++] = Seg{p, pt, off, n}; float* g = G + off; off += n; return g; };
  struct LinG { float* w; float* b; };
  auto lin = [&](float* w, float* b, float* wt, float* bt, int N, int K) { LinG g; g.w = seg(w, wt, (size_t)N * K); g.b = seg(b, bt, N); return g; };
  LinG g_e0 = lin(on.ew[0], on.eb[0], tg.ew[0], tg.eb[0], E, D), g_e1 = lin(on.ew[1], on.eb[1], tg.ew[1], tg.eb[1], L, E);
  const int seg_model0 = 0;
  LinG g_d[3] = {lin(on.dyn.w[0], on.dyn.b[0], tg.dyn.w[0], tg.dyn.b[0], M, W), lin(on.dyn.w[1], on.dyn.b[1], tg.dyn.w[1], tg.dyn.b[1], M, M),
                 lin(on.dyn.w[2], on.dyn.b[2], tg.dyn.w[2], tg.dyn.b[2], L, M)};
  LinG g_r[3] = {lin(on.rew.w[0], on.rew.b[0], tg.rew.w[0], tg.rew.b[0], M, W), lin(on.rew.w[1], on.rew.b[1], tg.rew.w[1], tg.rew.b[1], M, M),
                 lin(on.rew.w[2], on.rew.b[2], tg.rew.w[2], tg.rew.b[2], 1, M)};
  const int seg_pi0 = tab.count;
  LinG g_p[3] = {lin(on.pi.w[0], on.pi.b[0], tg.pi.w[0], tg.pi.b[0], M, L), lin(on.pi.w[1], on.pi.b[1], tg.pi.w[1], tg.pi.b[1], M, M),
                 lin(on.pi.w[2], on.pi.b[2], tg.pi.w[2], tg.pi.b[2], A, M)};
  const int seg_pi1 = tab.count;
  struct QG { LinG l[3]; float* g[2]; float* e[2]; } g_q[2];
  for (int h = 0; h < 2; ++h) {
    g_q[h].l[0] = lin(on.q[h].w[0], on.q[h].b[0], tg.q[h].w[0], tg.q[h].b[0], M, W);
    g_q[h].g[0] = seg(on.q[h].g[0], tg.q[h].g[0], M); g_q[h].e[0] = seg(on.q[h].e[0], tg.q[h].e[0], M);
    g_q[h].l[1] = lin(on.q[h].w[1], on.q[h].b[1], tg.q[h].w[1], tg.q[h].b[1], M, M);
    g_q[h].g[1] = seg(on.q[h].g[1], tg.q[h].g[1], M); g_q[h].e[1] = seg(on.q[h].e[1], tg.q[h].e[1], M);
    g_q[h].l[2] = lin(on.q[h].w[2], on.q[h].b[2], tg.q[h].w[2], tg.q[h].b[2], 1, M);
  }
  if ((long long)off != total || tab.count > MAX_SEGS) return -1;
  if (cudaMemsetAsync(G, 0, sizeof(float) * (size_t)total, st) != cudaSuccess) return -1;
  if (cudaMemsetAsync(ctl, 0, 64 * sizeof(float), st) != cudaSuccess) return -1;

  LossParams lp{};
  {
    double r = 1.0;
    for (int t = 0; t <= H && t < TDMPC_MAX_H; ++t) { lp.rho_pow[t] = (float)r; r *= a.rho; }
  }

  NoiseSrc nz_td{a.noise_pi, a.seed, 0, NS_PI, &dyn->call};
  NoiseSrc nz_up{a.noise_pi_update, a.seed, 0, NS_TERM, &dyn->call};

  // Q head forward over `rows` of xin [rows, W]: saves xhat / rstd / activations for the backward pass
  auto q_fwd = [&](const QNet& q, const float* xin, int rows, float* xha, float* ya, float* rsa_, float* xhb, float* yb, float* rsb_, float* out) {
    linear_fwd(c, rows, M, W, xin, W, q.w[0], q.b[0], xha, M, false);
    ln_act_fwd(c, xha, rows, M, q.g[0], q.e[0], ya, rsa_, true);
    linear_fwd(c, rows, M, M, ya, M, q.w[1], q.b[1], xhb, M, false);
    ln_act_fwd(c, xhb, rows, M, q.g[1], q.e[1], yb, rsb_, false);
    rowdot(c, yb, rows, M, q.w[2], q.b[2], out);
  };
  // Q head backward from dout [rows]: parameter gradients into `gq` (nullptr: frozen, update_pi) and dX (+)= into dxo [rows, W]
  auto q_bwd = [&](const QNet& q, QG* gq, const float* xin, int rows, const float* dout, const float* xha, const float* ya, const float* rsa_,
                   const float* xhb, const float* yb, const float* rsb_, float* dxo, float beta_x) {
    outer_kernel<<<nblk((size_t)rows * M), 256, 0, st>>>(dout, q.w[2], rows, M, dh2); c.check();       // d yb
    if (gq) { gemm(c, true, false, 1, M, rows, dout, 1, yb, M, gq->l[2].w, M, 0.0f); colsum(c, dout, rows, 1, 1, gq->l[2].b, false); }
    ln_act_bwd_kernel<<<(rows * 32 + 255) / 256, 256, 0, st>>>(dh2, yb, xhb, rsb_, q.g[1], rows, M, gq ? dln : nullptr, 0); c.check();   // dh2 = d pre2
    if (gq) {
      colsum_prod(c, dln, xhb, rows, M, gq->g[1]);
      colsum(c, dln, rows, M, M, gq->e[1], false);
    }
    linear_bwd(c, rows, M, M, dh2, M, ya, M, q.w[1], gq ? gq->l[1].w : nullptr, gq ? gq->l[1].b : nullptr, false, dh1, M);        // dh1 = d ya
    ln_act_bwd_kernel<<<(rows * 32 + 255) / 256, 256, 0, st>>>(dh1, ya, xha, rsa_, q.g[0], rows, M, gq ? dln : nullptr, 1); c.check();   // dh1 = d pre
    if (gq) {
      colsum_prod(c, dln, xha, rows, M, gq->g[0]);
      colsum(c, dln, rows, M, M, gq->e[0], false);
    }
    linear_bwd(c, rows, M, W, dh1, M, xin, W, q.w[0], gq ? gq->l[0].w : nullptr, gq ? gq->l[0].b : nullptr, false, dxo, W, beta_x);
  };

  // ================================================================================ forward (tdmpc.py:198-219)
  linear_fwd(c, B, E, D, a.obs, D, on.ew[0], on.eb[0], e1, E, true);
  linear_fwd(c, B, L, E, e1, E, on.ew[1], on.eb[1], zs, L, false);                        // z = h(obs); e1[0:B] stays for the backward pass
  for (int t = 0; t < H; ++t) {                                                          // z_{t+1} = dynamics([z_t, a_t])
    float* xt_ = x + (size_t)t * B * W;
    concat(c, zs + (size_t)t * B * L, a.action + (size_t)t * B * A, B, L, A, xt_);
    linear_fwd(c, B, M, W, xt_, W, on.dyn.w[0], on.dyn.b[0], d1 + (size_t)t * B * M, M, true);
    linear_fwd(c, B, M, M, d1 + (size_t)t * B * M, M, on.dyn.w[1], on.dyn.b[1], d2 + (size_t)t * B * M, M, true);
    linear_fwd(c, B, L, M, d2 + (size_t)t * B * M, M, on.dyn.w[2], on.dyn.b[2], zs + (size_t)(t + 1) * B * L, L, false);
  }
  linear_fwd(c, R, M, W, x, W, on.rew.w[0], on.rew.b[0], r1, M, true);                   // reward head, all steps
  linear_fwd(c, R, M, M, r1, M, on.rew.w[1], on.rew.b[1], r2, M, true);
  rowdot(c, r2, R, M, on.rew.w[2], on.rew.b[2], rp);
  for (int h = 0; h < 2; ++h) q_fwd(on.q[h], x, R, xa[h], qa[h], rsa[h], xb[h], qb[h], rsb[h], qv[h]);      // Q1, Q2 = Q(z_t, a_t)
  // targets (no_grad): next_z = h_target(next_obs); td = reward + gamma * min Q_target(h(next_obs), pi(h(next_obs)))
  float* et = e1 + (size_t)B * E;                                                        // scratch rows behind the encoder activations
  linear_fwd(c, R, E, D, a.next_obses, D, tg.ew[0], tg.eb[0], et, E, true);
  linear_fwd(c, R, L, E, et, E, tg.ew[1], tg.eb[1], next_z, L, false);
  linear_fwd(c, R, E, D, a.next_obses, D, on.ew[0], on.eb[0], et, E, true);
  linear_fwd(c, R, L, E, et, E, on.ew[1], on.eb[1], nzo, L, false);
  linear_fwd(c, R, M, L, nzo, L, on.pi.w[0], on.pi.b[0], p1, M, true);
  linear_fwd(c, R, M, M, p1, M, on.pi.w[1], on.pi.b[1], p2, M, true);
  linear_fwd(c, R, A, M, p2, M, on.pi.w[2], on.pi.b[2], pa, A, false);
  pi_sample_kernel<<<nblk((size_t)R * A), 256, 0, st>>>(pa, (size_t)R * A, a.min_std, nz_td, nullptr, api); c.check();
  concat(c, nzo, api, R, L, A, xt);
  // the target heads need no saved activations: the policy-update buffers (free until then) take them
  q_fwd(tg.q[0], xt, R, dh1, dh2, pq1, dln, p1, pq2, tq1);
  q_fwd(tg.q[1], xt, R, dh1, dh2, pq1, dln, p1, pq2, tq2);
  // losses + gradient seeds
  lp.B = B; lp.H = H; lp.L = L;
  lp.zs = zs; lp.next_z = next_z; lp.rp = rp; lp.q1 = qv[0]; lp.q2 = qv[1]; lp.tq1 = tq1; lp.tq2 = tq2; lp.reward = a.reward; lp.weights = a.weights;
  lp.discount = (float)a.discount; lp.c_cons = a.consistency_coef; lp.c_rew = a.reward_coef; lp.c_val = a.value_coef;
  lp.td = td; lp.d_z = d_z; lp.d_rp = d_rp; lp.d_q1 = d_q1; lp.d_q2 = d_q2; lp.per_sample = per_sample; lp.priorities = a.priorities;
  loss_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(lp); c.check();
  means_kernel<<<1, 256, 0, st>>>(per_sample, B, 5, a.metrics); c.check();               // consistency, reward, value, total, weighted

  // ================================================================================ backward (weighted_loss.backward())
  // Q heads and reward head of all steps at once: dx = dL/d[z_t, a_t]
  q_bwd(on.q[0], &g_q[0], x, R, d_q1, xa[0], qa[0], rsa[0], xb[0], qb[0], rsb[0], dx, 0.0f);
  q_bwd(on.q[1], &g_q[1], x, R, d_q2, xa[1], qa[1], rsa[1], xb[1], qb[1], rsb[1], dx, 1.0f);
  outer_kernel<<<nblk((size_t)R * M), 256, 0, st>>>(d_rp, on.rew.w[2], R, M, dh2); c.check();
  gemm(c, true, false, 1, M, R, d_rp, 1, r2, M, g_r[2].w, M, 0.0f); colsum(c, d_rp, R, 1, 1, g_r[2].b, false);
  act_bwd(c, dh2, r2, (size_t)R * M, false);
  linear_bwd(c, R, M, M, dh2, M, r1, M, on.rew.w[1], g_r[1].w, g_r[1].b, false, dh1, M);
  act_bwd(c, dh1, r1, (size_t)R * M, false);
  linear_bwd(c, R, M, W, dh1, M, x, W, on.rew.w[0], g_r[0].w, g_r[0].b, false, dx, W, 1.0f);
  // back through the dynamics chain
  for (int t = H - 1; t >= 0; --t) {
    const size_t rt = (size_t)t * B;
    // dL/dz_{t+1}: consistency seed + what the later steps sent back
    if (cudaMemcpyAsync(dzn, d_z + rt * L, sizeof(float) * (size_t)B * L, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return -1;
    if (t < H - 1) { add_kernel<<<nblk((size_t)B * L), 256, 0, st>>>(dzn, dzc, (size_t)B * L); c.check(); }
    linear_bwd(c, B, L, M, dzn, L, d2 + rt * M, M, on.dyn.w[2], g_d[2].w, g_d[2].b, true, dh2, M);
    act_bwd(c, dh2, d2 + rt * M, (size_t)B * M, false);
    linear_bwd(c, B, M, M, dh2, M, d1 + rt * M, M, on.dyn.w[1], g_d[1].w, g_d[1].b, true, dh1, M);
    act_bwd(c, dh1, d1 + rt * M, (size_t)B * M, false);
    linear_bwd(c, B, M, W, dh1, M, x + rt * W, W, on.dyn.w[0], g_d[0].w, g_d[0].b, true, dxd, W);
    // dL/dz_t = latent part of (dynamics + Q + reward) gradients
    take_z_kernel<<<nblk((size_t)B * L), 256, 0, st>>>(dxd, B, L, A, dzc, 0); c.check();
    take_z_kernel<<<nblk((size_t)B * L), 256, 0, st>>>(dx + rt * W, B, L, A, dzc, 1); c.check();
  }
  // encoder
  linear_bwd(c, B, L, E, dzc, L, e1, E, on.ew[1], g_e1.w, g_e1.b, false, de1, E);
  act_bwd(c, de1, e1, (size_t)B * E, false);
  linear_bwd(c, B, E, D, de1, E, a.obs, D, on.ew[0], g_e0.w, g_e0.b, false, nullptr, 0);

  // ================================================================================ clip_grad_norm_ + Adam (tdmpc.py:226-228)
  sumsq_kernel<<<128, 256, 0, st>>>(G, 0, (size_t)pi_lo, acc); c.check();
  sumsq_kernel<<<128, 256, 0, st>>>(G, (size_t)pi_hi, (size_t)total, acc + 1); c.check();
  clip_coef_kernel<<<1, 1, 0, st>>>(acc, 2, (float)a.grad_clip_norm, coef, a.metrics + 5); c.check();
  {
    AdamParams ap{(float)a.beta1, (float)a.beta2, (float)a.adam_eps, dyn->adam, G, a.adam_m, a.adam_v, 0, coef};
    adam_kernel<<<296, 256, 0, st>>>(tab, seg_model0, seg_pi0, ap); c.check();
    adam_kernel<<<296, 256, 0, st>>>(tab, seg_pi1, tab.count, ap); c.check();
  }

  // ================================================================================ update_pi (tdmpc.py:165-178): the Q heads just stepped
  linear_fwd(c, R1, M, L, zs, L, on.pi.w[0], on.pi.b[0], p1, M, true);
  linear_fwd(c, R1, M, M, p1, M, on.pi.w[1], on.pi.b[1], p2, M, true);
  linear_fwd(c, R1, A, M, p2, M, on.pi.w[2], on.pi.b[2], pa, A, false);
  pi_sample_kernel<<<nblk((size_t)R1 * A), 256, 0, st>>>(pa, (size_t)R1 * A, a.min_std, nz_up, pa, api); c.check();   // pa := mu = tanh(pre)
  concat(c, zs, api, R1, L, A, xt);
  for (int h = 0; h < 2; ++h) q_fwd(on.q[h], xt, R1, xa[h], qa[h], rsa[h], xb[h], qb[h], rsb[h], qv[h]);
  pi_loss_kernel<<<nblk(R1), 256, 0, st>>>(qv[0], qv[1], B, H + 1, rho_dev, pq1, pq2, qmin); c.check();
  pi_loss_sum_kernel<<<1, 256, 0, st>>>(qmin, B, H + 1, a.metrics + 6); c.check();
  q_bwd(on.q[0], nullptr, xt, R1, pq1, xa[0], qa[0], rsa[0], xb[0], qb[0], rsb[0], dx, 0.0f);
  q_bwd(on.q[1], nullptr, xt, R1, pq2, xa[1], qa[1], rsa[1], xb[1], qb[1], rsb[1], dx, 1.0f);
  pi_dpre_kernel<<<nblk((size_t)R1 * A), 256, 0, st>>>(dx, pa, R1, L, A, dpre); c.check();
  linear_bwd(c, R1, A, M, dpre, A, p2, M, on.pi.w[2], g_p[2].w, g_p[2].b, false, dh2, M);
  act_bwd(c, dh2, p2, (size_t)R1 * M, false);
  linear_bwd(c, R1, M, M, dh2, M, p1, M, on.pi.w[1], g_p[1].w, g_p[1].b, false, dh1, M);
  act_bwd(c, dh1, p1, (size_t)R1 * M, false);
  linear_bwd(c, R1, M, L, dh1, M, zs, L, on.pi.w[0], g_p[0].w, g_p[0].b, false, nullptr, 0);
  sumsq_kernel<<<128, 256, 0, st>>>(G, (size_t)pi_lo, (size_t)pi_hi, acc + 2); c.check();
  clip_coef_kernel<<<1, 1, 0, st>>>(acc + 2, 1, (float)a.grad_clip_norm, coef + 1, a.metrics + 7); c.check();
  {
    AdamParams ap{(float)a.beta1, (float)a.beta2, (float)a.adam_eps, dyn->adam + 2, G, a.pi_adam_m, a.pi_adam_v, (size_t)pi_lo, coef + 1};
    adam_kernel<<<296, 256, 0, st>>>(tab, seg_pi0, seg_pi1, ap); c.check();
  }
  // ================================================================================ target network (tdmpc.py:236-237)
  if (a.update_target) { ema_kernel<<<296, 256, 0, st>>>(tab, (float)a.tau); c.check(); }
  return c.ok ? c.launches : -1;
}

// ------------------------------------------------------------------------------------------------ the step as a CUDA graph
// A training step is ~290 short launches, and at the reference's batch sizes the host cannot issue them as fast as the GPU retires
// them (measured: 10.8 us per launch, 3.1 ms per step at batch 512, whatever the kernels cost).  The launch sequence depends only on
// shapes, hyper-parameters and the addresses of persistent buffers (parameters, optimiser state, workspace), so it is captured once
// per such key and replayed; what changes per step lives in the DynBlock, and the caller's input / output tensors are copied to /
// from fixed slots at the front of the workspace.  TDMPC_TRAIN_GRAPH=0 launches the kernels directly.
namespace {
struct GraphKey { tdmpc_model_cfg cfg; tdmpc_train_args args; };
struct GraphEntry { GraphKey key; cudaGraphExec_t exec; int launches; unsigned long long used; };
constexpr int MAX_GRAPHS = 8;
GraphEntry g_graphs[MAX_GRAPHS];
int g_graph_count = 0;
unsigned long long g_graph_clock = 0;
}  // namespace

int launch_train_step(const tdmpc_model_cfg& cfg, const tdmpc_train_args& a, cudaStream_t st) {
  static const int graph_mode = [] { const char* e = getenv("TDMPC_TRAIN_GRAPH"); return (e && e[0] == '0') ? 0 : 1; }();
  const int B = a.batch, H = a.horizon, D = cfg.obs_dim, A = cfg.action_dim;
  const size_t R = (size_t)B * H, R1 = (size_t)B * (H + 1);
  const FrontLayout fl = front_layout(cfg, B, H);
  float* front = reinterpret_cast<float*>(a.workspace);
  DynBlock* dyn = reinterpret_cast<DynBlock*>(front + fl.dyn);

  DynBlock hb{};
  hb.call = a.call;
  {
    const double s1 = (double)(a.adam_step + 1), s2 = (double)(a.pi_adam_step + 1);
    hb.adam[0] = (float)(a.lr / (1.0 - pow(a.beta1, s1))); hb.adam[1] = (float)sqrt(1.0 - pow(a.beta2, s1));
    hb.adam[2] = (float)(a.lr / (1.0 - pow(a.beta1, s2))); hb.adam[3] = (float)sqrt(1.0 - pow(a.beta2, s2));
    double r = 1.0;
    for (int t = 0; t <= H && t <= TDMPC_MAX_H; ++t) { hb.rho[t] = (float)r; r *= a.rho; }
  }
  // pageable source: the runtime stages these few bytes before the call returns, so the stack copy may die
  if (cudaMemcpyAsync(dyn, &hb, sizeof(hb), cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;

  tdmpc_train_args b = a;
  b.workspace = front + fl.total;
  b.workspace_bytes = a.workspace_bytes - (long long)(fl.total * sizeof(float));
  if (!graph_mode) { const int n = train_step_body(cfg, b, st, dyn); return n < 0 ? n : n + 1; }

  // ---- inputs into their slots
  int copies = 0;
  auto stage = [&](const void* src, size_t off, size_t n) -> const float* {
    if (cudaMemcpyAsync(front + off, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) copies = -1000000;
    ++copies;
    return front + off;
  };
  b.obs = stage(a.obs, fl.obs, (size_t)B * D);
  b.next_obses = stage(a.next_obses, fl.nxt, R * D);
  b.action = stage(a.action, fl.act, R * A);
  b.reward = stage(a.reward, fl.rew, R);
  b.weights = stage(a.weights, fl.wts, B);
  if (a.noise_pi) b.noise_pi = stage(a.noise_pi, fl.n1, R * A);
  if (a.noise_pi_update) b.noise_pi_update = stage(a.noise_pi_update, fl.n2, R1 * A);
  b.priorities = front + fl.pri;
  if (copies < 0) return -1;

  GraphKey key;
  memset(&key, 0, sizeof(key));
  memcpy(&key.cfg, &cfg, sizeof(cfg));
  memcpy(&key.args, &b, sizeof(b));
  key.args.call = 0; key.args.adam_step = 0; key.args.pi_adam_step = 0;
  GraphEntry* e = nullptr;
  for (int i = 0; i < g_graph_count; ++i)
    if (memcmp(&g_graphs[i].key, &key, sizeof(key)) == 0) { e = &g_graphs[i]; break; }
  int n;
  if (e && e->exec) {
    if (cudaGraphLaunch(e->exec, st) != cudaSuccess) return -1;
    n = e->launches;
    e->used = ++g_graph_clock;
  } else {
    // first step with this key: launch directly (this also sets the kernels' shared-memory attributes), then record the same
    // sequence on a capture stream for the next step -- capturing executes nothing
    n = train_step_body(cfg, b, st, dyn);
    if (n < 0) return n;
    if (!e) {
      if (g_graph_count < MAX_GRAPHS) e = &g_graphs[g_graph_count++];
      else {
        e = &g_graphs[0];
        for (int i = 1; i < MAX_GRAPHS; ++i) if (g_graphs[i].used < e->used) e = &g_graphs[i];
        if (e->exec) cudaGraphExecDestroy(e->exec);
      }
      e->key = key; e->exec = nullptr; e->launches = n; e->used = ++g_graph_clock;
      cudaStream_t cap = nullptr;
      cudaGraph_t graph = nullptr;
      if (cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) == cudaSuccess) {
        if (cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
          const int nc = train_step_body(cfg, b, cap, dyn);
          const cudaError_t end = cudaStreamEndCapture(cap, &graph);
          if (nc == n && end == cudaSuccess && graph) {
            if (cudaGraphInstantiate(&e->exec, graph, 0) != cudaSuccess) e->exec = nullptr;
          }
          if (graph) cudaGraphDestroy(graph);
        }
        cudaStreamDestroy(cap);
      }
      cudaGetLastError();                       // a failed capture leaves this key on direct launches
    }
  }
  if (cudaMemcpyAsync(a.priorities, front + fl.pri, sizeof(float) * B, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return -1;
  return n + copies + 2;
}
