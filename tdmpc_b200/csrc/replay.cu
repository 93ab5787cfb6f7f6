// Prioritised replay sampling on the device: the arithmetic of
// ReplayBuffer.sample (algorithm/helper.py:506-534) without the D2H copy of the whole priority
// vector and the host `np.random.choice` the reference does on every update.
//
//   probs   = priorities[:total] ** per_alpha;  probs /= probs.sum()                  (fp32, helper.py:507-508)
//   idxs    = np.random.choice(total, batch, p=probs, replace=not full)               (helper.py:510-511)
//             numpy: cdf = cumsum(p) in float64, cdf /= cdf[-1], searchsorted(u, side='right');
//             replace=False: keep the first occurrence of every index, zero p of the found ones, redraw the rest
//   weights = (total * probs[idxs]) ** (-per_beta);  weights /= weights.max()        (helper.py:512-513)
//   obs / next_obs / action / reward gathers over horizon+1 steps                     (helper.py:515-528)
//
// The cdf is never materialised: the priorities are cut into blocks of RB_BLK elements, one pass
// gives the block sums (float64), a single CTA scans them, and every draw (one warp each) bisects
// the block prefix and then re-walks ONE block with a warp scan.  Algorithmic bytes per sample call:
// 2 x 4 x total (two passes over the priorities) + batch x RB_BLK x 4 for the walks (L2 hits).
// HBM-bound, latency-scale: at total = 1 M that is 8 MB, about 1.3 us at the measured copy peak, so
// the cost is the launches (7 for replace=True).
#include "common.cuh"

#include <cstdio>
#include <string>

void tdmpc_set_global_error(const char* msg);   // api.cu: text behind tdmpc_last_error(NULL)

namespace {

constexpr int RB_BLK = 1024;        // priorities per scan block
constexpr int RB_THREADS = 256;
constexpr int RB_MAX_PARTS = 1024;  // partial sums of pass 1
constexpr int RB_MAX_BATCH = 4096;  // candidates of one round are de-duplicated in shared memory

// control words at the head of the workspace
struct ReplayCtl {
  long long n_found;     // distinct indices committed so far (numpy: n_uniq)
  long long cursor;      // uniforms consumed so far
  long long rounds;      // draw rounds executed
  double cdf_total;      // cumsum(p)[-1] of the current round
  float prob_sum;        // probs.sum() in fp32 (the divisor of helper.py:508)
  float _pad;
};

struct ReplayWs {
  ReplayCtl* ctl;
  double* part;              // [RB_MAX_PARTS]
  double* bsum;              // [nb] block sums of the normalised probabilities
  double* bpre;              // [nb] exclusive prefix of bsum
  long long* cand;           // [batch] candidates of the current round
  unsigned char* taken;      // [total] 1 = already drawn (replace = False)
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t ws_layout(int64_t total, int32_t batch, void* base, ReplayWs* ws) {
  const size_t nb = (size_t)((total + RB_BLK - 1) / RB_BLK);
  size_t off = 0;
  char* b = static_cast<char*>(base);
  auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return b ? b + o : nullptr; };
  void* ctl = take(sizeof(ReplayCtl));
  void* part = take(sizeof(double) * RB_MAX_PARTS);
  void* bsum = take(sizeof(double) * nb);
  void* bpre = take(sizeof(double) * nb);
  void* cand = take(sizeof(long long) * (size_t)batch);
  void* taken = take((size_t)total);
  if (ws) *ws = ReplayWs{(ReplayCtl*)ctl, (double*)part, (double*)bsum, (double*)bpre, (long long*)cand, (unsigned char*)taken};
  return off;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over the block, result valid in every thread
__device__ double block_sum(double v, double* sh /* [32] */) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = (l < nw) ? sh[l] : 0.0;
  return warp_sum(t);
}

// priorities ** alpha in fp32 (torch: elementwise pow on a float32 tensor)
__device__ __forceinline__ float pow_alpha(float p, float alpha) { return powf(p, alpha); }

// pass 1: partial float64 sums of priorities ** alpha (grid-stride), then the last step folds them to fp32
__global__ void replay_pow_partials(const float* __restrict__ pri, long long total, float alpha, double* __restrict__ part) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    acc += (double)pow_alpha(__ldg(pri + i), alpha);
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

__global__ void replay_begin(const double* __restrict__ part, int nparts, ReplayCtl* ctl) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += part[i];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    ctl->n_found = 0; ctl->cursor = 0; ctl->rounds = 0; ctl->cdf_total = 0.0;
    ctl->prob_sum = (float)acc;      // the sum is carried in float64 and rounded once (torch's own fp32 reduction order is unspecified)
  }
}

// the float64 value numpy sees for element i: float32(p_i^alpha / sum), 0 once it has been drawn without replacement
__device__ __forceinline__ double prob_at(const float* __restrict__ pri, const unsigned char* __restrict__ taken, long long i,
                                          float alpha, float psum) {
  if (taken && taken[i]) return 0.0;
  return (double)(pow_alpha(__ldg(pri + i), alpha) / psum);
}

// pass 2: block sums of the (masked) normalised probabilities
__global__ void replay_block_sums(const float* __restrict__ pri, const unsigned char* __restrict__ taken, long long total,
                                  float alpha, int batch, const ReplayCtl* __restrict__ ctl, double* __restrict__ bsum) {
  __shared__ double sh[32];
  if (ctl->n_found >= batch) return;
  const float psum = ctl->prob_sum;
  const long long base = (long long)blockIdx.x * RB_BLK;
  double acc = 0.0;
  for (int k = threadIdx.x; k < RB_BLK; k += blockDim.x) {
    const long long i = base + k;
    if (i < total) acc += prob_at(pri, taken, i, alpha, psum);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) bsum[blockIdx.x] = acc;
}

// single CTA: exclusive scan of the block sums
__global__ void replay_scan_blocks(const double* __restrict__ bsum, double* __restrict__ bpre, long long nb, int batch, ReplayCtl* ctl) {
  __shared__ double wsum[32];
  __shared__ double carry_s;
  if (ctl->n_found >= batch) return;
  if (threadIdx.x == 0) carry_s = 0.0;
  __syncthreads();
  const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (long long t0 = 0; t0 < nb; t0 += blockDim.x) {
    const long long i = t0 + threadIdx.x;
    const double v = i < nb ? bsum[i] : 0.0;
    double s = v;                                              // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, s, o); if (l >= o) s += u; }
    if (l == 31) wsum[w] = s;
    __syncthreads();
    double woff = 0.0;
    for (int k = 0; k < w; ++k) woff += wsum[k];
    const double carry = carry_s;
    if (i < nb) bpre[i] = carry + woff + (s - v);
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = carry + woff + s;
    __syncthreads();
  }
  if (threadIdx.x == 0) ctl->cdf_total = carry_s;
}

// one warp per draw: first index with cdf[i] / cdf[-1] > u  (searchsorted side='right')
__global__ void replay_search(const float* __restrict__ pri, const unsigned char* __restrict__ taken, long long total, float alpha,
                              int batch, const double* __restrict__ uniforms, NoiseSrc rng, const ReplayCtl* __restrict__ ctl,
                              const double* __restrict__ bsum, const double* __restrict__ bpre, long long nb,
                              long long* __restrict__ cand) {
  const long long found = ctl->n_found;
  const int m = (int)(batch - found);
  const int j = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= m) return;
  const long long e = ctl->cursor + j;
  const double u = uniforms ? uniforms[e] : philox_uniform64(rng, (uint64_t)e);
  const double C = ctl->cdf_total;
  const float psum = ctl->prob_sum;
  // smallest block whose inclusive prefix / C exceeds u
  long long lo = 0, hi = nb - 1;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((bpre[mid] + bsum[mid]) / C > u) hi = mid; else lo = mid + 1;
  }
  long long pick = -1, last_pos = -1;
  for (long long b = lo; b < nb && pick < 0; ++b) {       // walks on past b = lo only if the re-summed block rounds differently
    double run = bpre[b];
    for (int c = 0; c < RB_BLK / 32 && pick < 0; ++c) {
      const long long i = b * RB_BLK + c * 32 + lane;
      const double q = i < total ? prob_at(pri, taken, i, alpha, psum) : 0.0;
      double s = q;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
      const unsigned hit = __ballot_sync(0xffffffffu, q > 0.0 && (run + s) / C > u);
      const unsigned pos = __ballot_sync(0xffffffffu, q > 0.0);
      if (hit) pick = b * RB_BLK + c * 32 + (__ffs(hit) - 1);
      if (pos) last_pos = b * RB_BLK + c * 32 + (31 - __clz(pos));
      run += __shfl_sync(0xffffffffu, s, 31);
    }
  }
  if (pick < 0) pick = last_pos >= 0 ? last_pos : total - 1;     // u beyond the re-summed total: the last drawable element
  if (lane == 0) cand[j] = pick;
}

// single CTA: numpy's "keep the first occurrence, in draw order" + bookkeeping of the round
__global__ void replay_commit(const long long* __restrict__ cand, int batch, int replace, ReplayCtl* ctl, unsigned char* taken,
                              long long* __restrict__ idxs) {
  __shared__ long long sc[RB_MAX_BATCH];
  __shared__ int wsum[32];
  __shared__ int carry_s;
  const long long found = ctl->n_found;
  const int m = (int)(batch - found);
  if (m <= 0) return;
  for (int j = threadIdx.x; j < m; j += blockDim.x) sc[j] = cand[j];
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int t0 = 0; t0 < m; t0 += blockDim.x) {
    const int j = t0 + threadIdx.x;
    int keep = 0;
    if (j < m) {
      keep = 1;
      if (!replace) {
        const long long v = sc[j];
        for (int k = 0; k < j; ++k) if (sc[k] == v) { keep = 0; break; }
      }
    }
    int s = keep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (l >= o) s += t; }
    if (l == 31) wsum[w] = s;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < w; ++k) woff += wsum[k];
    const int carry = carry_s;
    if (keep) {
      idxs[found + carry + woff + s - 1] = sc[j];
      if (!replace) taken[sc[j]] = 1;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = carry + woff + s;
    __syncthreads();
  }
  if (threadIdx.x == 0) { ctl->n_found = found + carry_s; ctl->cursor += m; ctl->rounds += 1; }
}

// single CTA: importance weights (helper.py:512-513) and the status words
__global__ void replay_weights(const float* __restrict__ pri, long long total, float alpha, float beta, int batch,
                               const ReplayCtl* __restrict__ ctl, long long* __restrict__ idxs, float* __restrict__ weights,
                               int* __restrict__ status) {
  __shared__ float sh[32];
  const int found = (int)ctl->n_found;
  const float psum = ctl->prob_sum;
  float mx = 0.0f;
  for (int j = threadIdx.x; j < batch; j += blockDim.x) {
    float wgt = 0.0f;
    if (j < found) {
      const float p = pow_alpha(__ldg(pri + idxs[j]), alpha) / psum;
      wgt = powf((float)total * p, -beta);
    } else {
      idxs[j] = -1;
    }
    weights[j] = wgt;
    mx = fmaxf(mx, wgt);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = (threadIdx.x & 31) < (blockDim.x >> 5) ? sh[threadIdx.x & 31] : 0.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  for (int j = threadIdx.x; j < batch; j += blockDim.x) weights[j] = weights[j] / mx;
  if (threadIdx.x == 0 && status) { status[0] = found; status[1] = (int)ctl->rounds; }
}

// grid (batch, horizon + 2): slice 0 is `obs`, slice s >= 1 is step t = s - 1 of next_obs / action / reward
__global__ void replay_gather(tdmpc_replay_gather_args a) {
  const int j = blockIdx.x, s = blockIdx.y;
  const long long idx = a.idxs[j];
  if (idx < 0) return;
  const int D = a.obs_dim, A = a.action_dim, B = a.batch;
  if (s == 0) {
    const float* src = a.obs_buf + idx * D;
    float* dst = a.obs + (long long)j * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) dst[d] = __ldg(src + d);
    return;
  }
  const int t = s - 1;
  const long long row = idx + t;
  // helper.py:526-527: on the last step, a transition that ends its episode reads the episode's stored final observation
  const bool last = (t == a.horizon) && ((row + 1) % a.episode_length == 0);
  const float* src = last ? a.last_obs_buf + (row / a.episode_length) * D : a.obs_buf + (row + 1) * D;
  float* dst = a.next_obs + ((long long)t * B + j) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) dst[d] = __ldg(src + d);
  for (int d = threadIdx.x; d < A; d += blockDim.x) a.action[((long long)t * B + j) * A + d] = __ldg(a.action_buf + row * A + d);
  if (threadIdx.x == 0) a.reward[(long long)t * B + j] = __ldg(a.reward_buf + row);
}

int fail(int code, const char* msg) { tdmpc_set_global_error(msg); return code; }

}  // namespace

extern "C" int64_t tdmpc_replay_workspace_bytes(int64_t total, int32_t batch) {
  if (total <= 0 || batch <= 0) return 0;
  return (int64_t)ws_layout(total, batch, nullptr, nullptr);
}

extern "C" int tdmpc_replay_sample(const tdmpc_replay_sample_args* a, void* stream_) {
  if (!a) return fail(TDMPC_ERR_INVALID, "replay_sample: null args");
  if (!a->priorities || !a->idxs || !a->weights || !a->workspace) return fail(TDMPC_ERR_INVALID, "replay_sample: null device pointer");
  if (a->total <= 0 || a->batch <= 0 || a->batch > RB_MAX_BATCH) return fail(TDMPC_ERR_INVALID, "replay_sample: total <= 0 or batch outside [1, 4096]");
  if (!a->replace && a->batch > a->total) return fail(TDMPC_ERR_INVALID, "replay_sample: cannot take more distinct samples than the buffer holds");
  const int rounds = a->replace ? 1 : (a->max_rounds > 0 ? a->max_rounds : 8);
  if (a->uniforms && a->n_uniforms < (int64_t)a->batch) return fail(TDMPC_ERR_INVALID, "replay_sample: fewer injected uniforms than draws");
  ReplayWs ws;
  const size_t need = ws_layout(a->total, a->batch, a->workspace, &ws);
  if ((int64_t)need > a->workspace_bytes) return fail(TDMPC_ERR_WORKSPACE, "replay_sample: workspace smaller than tdmpc_replay_workspace_bytes()");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const long long total = a->total;
  const long long nb = (total + RB_BLK - 1) / RB_BLK;
  int nparts = (int)((total + RB_THREADS * 8 - 1) / (RB_THREADS * 8));
  nparts = nparts < 1 ? 1 : (nparts > RB_MAX_PARTS ? RB_MAX_PARTS : nparts);
  unsigned char* taken = a->replace ? nullptr : ws.taken;
  if (taken && cudaMemsetAsync(taken, 0, (size_t)total, st) != cudaSuccess) return fail(TDMPC_ERR_CUDA, "replay_sample: memset failed");
  NoiseSrc rng{nullptr, a->seed, a->call, NS_REPLAY, nullptr};
  replay_pow_partials<<<nparts, RB_THREADS, 0, st>>>(a->priorities, total, a->per_alpha, ws.part);
  replay_begin<<<1, RB_THREADS, 0, st>>>(ws.part, nparts, ws.ctl);
  // injected uniforms: a round can only be run if the buffer still covers its draws (numpy consumes batch - n_found per round);
  // the caller sizes it (n_uniforms) and the status words tell how far the draw got
  for (int r = 0; r < rounds; ++r) {
    if (a->uniforms && r > 0 && a->n_uniforms < (int64_t)a->batch * (r + 1)) break;
    replay_block_sums<<<(unsigned)nb, RB_THREADS, 0, st>>>(a->priorities, taken, total, a->per_alpha, a->batch, ws.ctl, ws.bsum);
    replay_scan_blocks<<<1, 1024, 0, st>>>(ws.bsum, ws.bpre, nb, a->batch, ws.ctl);
    replay_search<<<(a->batch * 32 + RB_THREADS - 1) / RB_THREADS, RB_THREADS, 0, st>>>(
        a->priorities, taken, total, a->per_alpha, a->batch, a->uniforms, rng, ws.ctl, ws.bsum, ws.bpre, nb, ws.cand);
    replay_commit<<<1, 1024, 0, st>>>(ws.cand, a->batch, a->replace, ws.ctl, taken, reinterpret_cast<long long*>(a->idxs));
  }
  replay_weights<<<1, 1024, 0, st>>>(a->priorities, total, a->per_alpha, a->per_beta, a->batch, ws.ctl, reinterpret_cast<long long*>(a->idxs), a->weights, a->status);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { char b[256]; snprintf(b, sizeof(b), "replay_sample: launch failed: %s", cudaGetErrorString(e)); return fail(TDMPC_ERR_CUDA, b); }
  return TDMPC_OK;
}

extern "C" int tdmpc_replay_gather(const tdmpc_replay_gather_args* a, void* stream_) {
  if (!a) return fail(TDMPC_ERR_INVALID, "replay_gather: null args");
  if (!a->idxs || !a->obs_buf || !a->last_obs_buf || !a->action_buf || !a->reward_buf || !a->obs || !a->next_obs || !a->action || !a->reward)
    return fail(TDMPC_ERR_INVALID, "replay_gather: null device pointer");
  if (a->batch <= 0 || a->horizon < 0 || a->horizon + 2 > 65535 || a->obs_dim <= 0 || a->action_dim <= 0 || a->episode_length <= 0)
    return fail(TDMPC_ERR_INVALID, "replay_gather: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  replay_gather<<<dim3((unsigned)a->batch, (unsigned)(a->horizon + 2)), 128, 0, st>>>(*a);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { char b[256]; snprintf(b, sizeof(b), "replay_gather: launch failed: %s", cudaGetErrorString(e)); return fail(TDMPC_ERR_CUDA, b); }
  return TDMPC_OK;
}
