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
// the block prefix and then re-walks ONE block with a warp scan.  priorities ** alpha is evaluated once (pass 1 keeps it
// in the workspace).  Algorithmic bytes per sample call: 3 x 4 x total (read priorities, write and re-read the powers)
// + batch x RB_BLK x 4 for the walks (L2 hits).
// HBM-bound, latency-scale: at total = 1 M that is 12 MB, about 1.9 us at the measured copy peak, so
// the cost is the launches: 3 kernels + 1 memset for replace=True (every stage's serial tail runs in the last block
// to finish, instead of in a single-CTA kernel of its own), 2 more kernels per redraw round.
#include "common.cuh"

#include <cstdio>
#include <string>

void tdmpc_set_global_error(const char* msg);   // api.cu: text behind tdmpc_last_error(NULL)

namespace {

constexpr int RB_BLK = 1024;        // priorities per scan block
constexpr int RB_THREADS = 256;
constexpr int RB_MAX_PARTS = 1024;  // partial sums of pass 1
constexpr int RB_MAX_BATCH = 4096;  // candidates of one round are de-duplicated in shared memory

// control words at the head of the workspace (zeroed by a memset at the start of every fresh draw)
struct ReplayCtl {
  long long n_found;     // distinct indices committed so far (numpy: n_uniq)
  long long cursor;      // uniforms consumed so far
  long long rounds;      // draw rounds executed
  double cdf_total;      // cumsum(p)[-1] of the current round
  float prob_sum;        // probs.sum() in fp32 (the divisor of helper.py:508)
  unsigned int done[3];  // "last block finishes the job" tickets of the three kernels
};

struct ReplayWs {
  ReplayCtl* ctl;
  unsigned char* taken;      // [total] 1 = already drawn (replace = False); directly behind ctl: one memset clears both
  double* part;              // [RB_MAX_PARTS]
  double* bsum;              // [nb] block sums of the normalised probabilities
  double* bpre;              // [nb] exclusive prefix of bsum
  long long* cand;           // [batch] candidates of the current round
  float* pw;                 // [total] priorities ** alpha, computed once by pass 1 (accurate powf is the expensive part of every sweep)
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t ws_layout(int64_t total, int32_t batch, void* base, ReplayWs* ws) {
  const size_t nb = (size_t)((total + RB_BLK - 1) / RB_BLK);
  size_t off = 0;
  char* b = static_cast<char*>(base);
  auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return b ? b + o : nullptr; };
  void* ctl = take(sizeof(ReplayCtl));
  void* taken = take((size_t)total);
  void* part = take(sizeof(double) * RB_MAX_PARTS);
  void* bsum = take(sizeof(double) * nb);
  void* bpre = take(sizeof(double) * nb);
  void* cand = take(sizeof(long long) * (size_t)batch);
  void* pw = take(sizeof(float) * (size_t)total);
  if (ws) *ws = ReplayWs{(ReplayCtl*)ctl, (unsigned char*)taken, (double*)part, (double*)bsum, (double*)bpre, (long long*)cand, (float*)pw};
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

// Every kernel below ends with "the last block to finish does the serial tail" (a ticket in ReplayCtl::done), which
// saves the launch of a single-CTA kernel per stage: the whole draw is 3 launches instead of 7.
__device__ bool last_block_done(unsigned int* ticket) {
  __shared__ bool last;
  __threadfence();                       // this block's global writes before the ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
    if (last) *ticket = 0;               // ready for the next round
  }
  __syncthreads();
  if (last) __threadfence();
  return last;
}

// priorities ** alpha in fp32 (torch: elementwise pow on a float32 tensor)
__device__ __forceinline__ float pow_alpha(float p, float alpha) { return powf(p, alpha); }

// launch 1: partial float64 sums of priorities ** alpha (grid-stride); the last block folds them into probs.sum()
__global__ void replay_pow_sum(const float* __restrict__ pri, long long total, float alpha, float* __restrict__ pw, double* part, ReplayCtl* ctl) {
  __shared__ double sh[32];
  double acc = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    float v[4];                                   // four independent loads in flight before the first powf
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = i0 + k * stride < total ? __ldg(pri + i0 + k * stride) : 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i0 + k * stride < total) { const float w = pow_alpha(v[k], alpha); pw[i0 + k * stride] = w; acc += (double)w; }
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
  if (!last_block_done(&ctl->done[0])) return;
  acc = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) acc += __ldcg(part + i);
  acc = block_sum(acc, sh);
  // the sum is carried in float64 and rounded once (torch's own fp32 reduction order is unspecified)
  if (threadIdx.x == 0) ctl->prob_sum = (float)acc;
}

// the float64 value numpy sees for element i: float32(p_i^alpha / sum), 0 once it has been drawn without replacement
__device__ __forceinline__ double prob_at(const float* __restrict__ pw, const unsigned char* taken, long long i, float psum) {
  if (taken && taken[i]) return 0.0;
  return (double)(__ldg(pw + i) / psum);
}

// exclusive scan of the block sums by one CTA
__device__ void scan_block_sums(const double* bsum, double* bpre, long long nb, ReplayCtl* ctl) {
  __shared__ double wsum[32];
  __shared__ double carry_s;
  if (threadIdx.x == 0) carry_s = 0.0;
  __syncthreads();
  const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (long long t0 = 0; t0 < nb; t0 += blockDim.x) {
    const long long i = t0 + threadIdx.x;
    const double v = i < nb ? __ldcg(bsum + i) : 0.0;
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

// launch 2 (per round): block sums of the (masked) normalised probabilities; the last block scans them
__global__ void replay_block_sums(const float* __restrict__ pw, const unsigned char* taken, long long total,
                                  int batch, ReplayCtl* ctl, double* bsum, double* bpre) {
  __shared__ double sh[32];
  if (ctl->n_found >= batch) return;
  const float psum = ctl->prob_sum;
  const long long base = (long long)blockIdx.x * RB_BLK;
  double acc = 0.0;
  static_assert(RB_BLK == 4 * RB_THREADS, "one pass of four loads per thread");
  float v[4];
  bool live[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {                    // loads first, arithmetic after
    const long long i = base + threadIdx.x + k * RB_THREADS;
    live[k] = i < total && !(taken && taken[i]);
    v[k] = live[k] ? __ldg(pw + i) : 0.0f;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) if (live[k]) acc += (double)(v[k] / psum);
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) bsum[blockIdx.x] = acc;
  if (!last_block_done(&ctl->done[1])) return;
  scan_block_sums(bsum, bpre, (long long)gridDim.x, ctl);
}

// numpy's "keep the first occurrence, in draw order" + bookkeeping of the round (one CTA)
__device__ void commit_round(const long long* cand, int batch, int replace, ReplayCtl* ctl, unsigned char* taken, long long* idxs,
                             long long* sc /* shared [RB_MAX_BATCH] */) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  const long long found = ctl->n_found;
  const int m = (int)(batch - found);
  for (int j = threadIdx.x; j < m; j += blockDim.x) sc[j] = __ldcg(cand + j);
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int t0 = 0; t0 < m; t0 += blockDim.x) {
    const int j = t0 + threadIdx.x;
    int keep = 0;
    if (j < m && sc[j] >= 0) {
      keep = 1;
      if (!replace) {
        const long long v = sc[j];                           // branch-free sweep, two candidates per 16-byte shared load
        const longlong2* sc2 = reinterpret_cast<const longlong2*>(sc);
        int dup = 0;
#pragma unroll 4
        for (int k = 0; k < (j >> 1); ++k) { const longlong2 c = sc2[k]; dup |= (c.x == v) | (c.y == v); }
        if ((j & 1) && sc[j - 1] == v) dup = 1;
        keep = !dup;
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
  __syncthreads();
}

// importance weights (helper.py:512-513) and the status words (one CTA)
__device__ void finish_weights(const float* __restrict__ pw, long long total, float beta, int batch,
                               const ReplayCtl* ctl, long long* idxs, float* weights, int* status) {
  __shared__ float sh[32];
  const int found = (int)ctl->n_found;
  const float psum = ctl->prob_sum;
  float mx = 0.0f;
  for (int j = threadIdx.x; j < batch; j += blockDim.x) {
    float wgt = 0.0f;
    if (j < found) {
      const float p = __ldg(pw + idxs[j]) / psum;
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
  for (int j = threadIdx.x; j < batch; j += blockDim.x) weights[j] = weights[j] / mx;   // each thread re-reads its own writes
  if (threadIdx.x == 0 && status) { status[0] = found; status[1] = (int)ctl->rounds; }
}

// launch 3 (per round), one warp per draw: first index with cdf[i] / cdf[-1] > u  (searchsorted side='right');
// the last block commits the round and, when the draw is complete or this was the last round enqueued, writes the weights
__global__ void replay_search(const float* __restrict__ pw, unsigned char* taken, long long total, float beta,
                              int batch, int replace, int final_round, const double* __restrict__ uniforms, NoiseSrc rng,
                              ReplayCtl* ctl, const double* __restrict__ bsum, const double* __restrict__ bpre, long long nb,
                              long long* cand, long long* idxs, float* weights, int* status) {
  __shared__ __align__(16) long long sc[RB_MAX_BATCH];
  const long long found = ctl->n_found;
  if (found >= batch) return;
  const int m = (int)(batch - found);
  const int j = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (j < m) {
    const long long e = ctl->cursor + j;
    const double u = uniforms ? uniforms[e] : philox_uniform64(rng, (uint64_t)e);
    const double C = ctl->cdf_total;
    const float psum = ctl->prob_sum;
    long long pick = -1;
    if (C > 0.0) {           // else nothing drawable is left (numpy: "Fewer non-zero entries in p than size")
      // smallest block whose inclusive prefix / C exceeds u
      long long lo = 0, hi = nb - 1;
      while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((bpre[mid] + bsum[mid]) / C > u) hi = mid; else lo = mid + 1;
      }
      // two levels inside a block: every lane sums RB_BLK/32 = 32 consecutive elements, one warp scan finds the lane whose
      // run crosses u, a second warp scan over that lane's 32 elements finds the element
      long long last_chunk = -1;
      for (long long b = lo; b < nb && pick < 0; ++b) {     // walks on past b = lo only if the re-summed block rounds differently
        const long long base = b * RB_BLK;
        double tot = 0.0;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
          const long long i = base + lane * 32 + k;
          if (i < total) tot += prob_at(pw, taken, i, psum);
        }
        double s = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
        const double run = bpre[b];
        const unsigned hit = __ballot_sync(0xffffffffu, tot > 0.0 && (run + s) / C > u);
        const unsigned pos = __ballot_sync(0xffffffffu, tot > 0.0);
        if (pos) last_chunk = base + (31 - __clz(pos)) * 32;
        if (!hit) continue;
        const int L = __ffs(hit) - 1;
        const double excl = run + __shfl_sync(0xffffffffu, s - tot, L);
        const long long i = base + L * 32 + lane;
        const double q = i < total ? prob_at(pw, taken, i, psum) : 0.0;
        double s2 = q;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, s2, o); if (lane >= o) s2 += t; }
        const unsigned hit2 = __ballot_sync(0xffffffffu, q > 0.0 && (excl + s2) / C > u);
        const unsigned pos2 = __ballot_sync(0xffffffffu, q > 0.0);
        pick = base + L * 32 + (hit2 ? __ffs(hit2) - 1 : 31 - __clz(pos2));   // no hit: the chunk re-summed one ulp lower -> its last drawable element
      }
      if (pick < 0 && last_chunk >= 0) {                     // u beyond the re-summed total: the last drawable element of the buffer
        const long long i = last_chunk + lane;
        const unsigned pos2 = __ballot_sync(0xffffffffu, i < total && prob_at(pw, taken, i, psum) > 0.0);
        if (pos2) pick = last_chunk + (31 - __clz(pos2));
      }
    }
    if (lane == 0) cand[j] = pick;
  }
  if (!last_block_done(&ctl->done[2])) return;
  commit_round(cand, batch, replace, ctl, taken, idxs, sc);
  if (final_round || ctl->n_found >= batch) finish_weights(pw, total, beta, batch, ctl, idxs, weights, status);
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
  int nparts = (int)((total + RB_THREADS * 4 - 1) / (RB_THREADS * 4));
  nparts = nparts < 1 ? 1 : (nparts > RB_MAX_PARTS ? RB_MAX_PARTS : nparts);
  unsigned char* taken = a->replace ? nullptr : ws.taken;
  NoiseSrc rng{nullptr, a->seed, a->call, NS_REPLAY, nullptr};
  if (!a->resume) {
    // one memset clears the control words and (without replacement) the drawn-already mask that sits right behind them
    const size_t clear = taken ? (size_t)((char*)ws.taken - (char*)ws.ctl) + (size_t)total : sizeof(ReplayCtl);
    if (cudaMemsetAsync(ws.ctl, 0, clear, st) != cudaSuccess) return fail(TDMPC_ERR_CUDA, "replay_sample: memset failed");
    replay_pow_sum<<<nparts, RB_THREADS, 0, st>>>(a->priorities, total, a->per_alpha, ws.pw, ws.part, ws.ctl);
  }
  // injected uniforms: a round can only be run if the buffer still covers its draws (numpy consumes batch - n_found per round);
  // the caller sizes it (n_uniforms) and the status words tell how far the draw got
  int last = 0;
  for (int r = 0; r < rounds; ++r)
    if (!(a->uniforms && (r > 0 || a->resume) && a->n_uniforms < (int64_t)a->batch * (r + 1 + (a->resume > 0 ? a->resume : 0)))) last = r + 1;
    else break;
  if (last == 0) return fail(TDMPC_ERR_INVALID, "replay_sample: injected uniforms exhausted");
  for (int r = 0; r < last; ++r) {
    replay_block_sums<<<(unsigned)nb, RB_THREADS, 0, st>>>(ws.pw, taken, total, a->batch, ws.ctl, ws.bsum, ws.bpre);
    replay_search<<<(a->batch * 32 + RB_THREADS - 1) / RB_THREADS, RB_THREADS, 0, st>>>(
        ws.pw, taken, total, a->per_beta, a->batch, a->replace, r == last - 1, a->uniforms, rng, ws.ctl,
        ws.bsum, ws.bpre, nb, ws.cand, reinterpret_cast<long long*>(a->idxs), a->weights, a->status);
  }
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
