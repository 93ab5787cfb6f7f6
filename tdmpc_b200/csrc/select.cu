// Elite selection + sampling-distribution update + next-iteration sampling: one CTA per planning
// problem.  Restates algorithm/tdmpc.py:130-134 (sampling), :138-149 (top-k, softmax weights,
// weighted mean/std, clamp, momentum) and :152-163 (categorical pick, exploration noise, metrics).
//
// top-k: in-CTA bitonic sort of (value, index) pairs -- strides < 32 run on warp shuffles, larger
// strides through shared memory -- giving the K largest values sorted descending with ties broken
// towards the lower index (torch.topk leaves tie order unspecified, tdmpc.py:138).
// The weighted reductions are warp-shuffle trees; nothing but [B,H,A] mean/std, the K elite
// indices and the next iteration's action rows leave the SM.
#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int NT_MAX = 1024;     // select_kernel on a small batch: one thread per sorted element
constexpr int MAX_SORT = 4096;
constexpr int MAX_K = 256;
constexpr int MAX_HA = TDMPC_MAX_H * 64;

__device__ __forceinline__ bool before(float ka, int ia, float kb, int ib) {   // a sorts before b (descending)
  return ka > kb || (ka == kb && ia < ib);
}

__device__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  // (the tree over 32 slots adds zeros first when the block has 8 warps: the same sums, bit for bit, as the 8-slot tree)
  float t = (threadIdx.x < blockDim.x / 32) ? red[threadIdx.x] : 0.0f;
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// Sorts n (power of two, >= 64) pairs in shared memory, descending.
__device__ void bitonic_sort_desc(float* key, int* idx, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const int l = i ^ j;
          if (l > i) {
            const float ka = key[i], kb = key[l];
            const int ia = idx[i], ib = idx[l];
            const bool desc = (i & k) == 0;
            const bool sw = desc ? before(kb, ib, ka, ia) : before(ka, ia, kb, ib);
            if (sw) { key[i] = kb; key[l] = ka; idx[i] = ib; idx[l] = ia; }
          }
        }
        __syncthreads();
      } else {
        // all remaining strides of this k are < 32: finish them in registers with shuffles
        for (int base = (threadIdx.x >> 5) * 32; base < n; base += blockDim.x) {
          const int i = base + (threadIdx.x & 31);
          float ka = key[i]; int ia = idx[i];
          for (int jj = j; jj > 0; jj >>= 1) {
            const float kb = __shfl_xor_sync(0xffffffffu, ka, jj);
            const int ib = __shfl_xor_sync(0xffffffffu, ia, jj);
            const bool lower = (i & jj) == 0;
            const bool desc = (i & k) == 0;
            // the lower lane keeps the element that sorts first when desc, else the other one
            const bool mine_first = before(ka, ia, kb, ib);
            const bool keep = (lower == desc) ? mine_first : !mine_first;
            if (!keep) { ka = kb; ia = ib; }
          }
          key[i] = ka; idx[i] = ia;
        }
        __syncthreads();
        break;
      }
    }
  }
}

// Unit-variance noise of trajectory n at step t, action a.  Built-in generator: white rows read one
// Philox normal; colored rows (iCEM) are S[t][:] . white[:, n, a] with S the power-law synthesis
// matrix along the horizon axis (restates colorednoise.powerlaw_psd_gaussian as a linear map).
__device__ __forceinline__ float sample_noise(const SampleSpec& sp, int b, int t, int n, int a, int A) {
  const uint64_t row0 = sp.cem_base + (uint64_t)b * sp.cem_bstride;
  if (sp.cem.inj || sp.color_len == 0 || n < sp.n_white)
    return noise_at(sp.cem, row0 + ((uint64_t)t * sp.cem_rows + n) * A + a);
  const float* S = sp.color_mats + (n < sp.n_white + sp.n_pink ? 0 : sp.color_len * sp.color_len) + t * sp.color_len;
  float acc = 0.0f;
  for (int j = 0; j < sp.color_len; ++j)
    acc = fmaf(__ldg(S + j), philox_normal(sp.cem, row0 + ((uint64_t)j * sp.cem_rows + n) * A + a), acc);
  return acc;
}

// Rows of one iteration, in the reference's concatenation order [sampled | reused elites | pi]
// (tdmpc.py:130-134; tdmpc_icem_similarity_mlp.py:230-233).  `reuse` rows must already be in place
// (init_kernel writes the shifted elites itself) unless reuse != nullptr.
// `part` of `nparts` CTAs share one problem's rows (element-interleaved); the single-CTA callers pass 0 of 1.
__device__ void sample_rows(int b, const SampleSpec& sp, int Tcap, int H, int A, const float* mean_s, const float* std_s,
                            float* actions, const float* pi_actions, int P0, const float* reuse, int reuse_rows,
                            int part = 0, int nparts = 1) {
  const int HA = H * A;
  const int first = part * blockDim.x + threadIdx.x, stride = nparts * blockDim.x;
  float* out = actions + (size_t)b * Tcap * HA;
  const uint64_t row0 = sp.cem_base + (uint64_t)b * sp.cem_bstride;
  if (!sp.cem.inj && sp.color_len == 0 && (A & 3) == 0 && (row0 & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    // white built-in noise, A a multiple of 4: the 4 actions of one (row, step, a..a+3) group are the 4 normals of ONE
    // Philox block -- one block and two Box-Muller pairs per 4 outputs instead of per output, one 16-byte store
    for (int q = first; q < sp.N * HA / 4; q += stride) {
      const int e = q * 4;
      const int n = e / HA, ta = e - n * HA;
      const int t = ta / A, a = ta - t * A;
      float z[4];
      philox_normal4(sp.cem, (row0 + ((uint64_t)t * sp.cem_rows + n) * A + a) >> 2, z);
      float4 v;
      v.x = fminf(fmaxf(mean_s[ta] + std_s[ta] * z[0], -1.0f), 1.0f);
      v.y = fminf(fmaxf(mean_s[ta + 1] + std_s[ta + 1] * z[1], -1.0f), 1.0f);
      v.z = fminf(fmaxf(mean_s[ta + 2] + std_s[ta + 2] * z[2], -1.0f), 1.0f);
      v.w = fminf(fmaxf(mean_s[ta + 3] + std_s[ta + 3] * z[3], -1.0f), 1.0f);
      if (sp.best_a && n == 0) v = make_float4(mean_s[ta], mean_s[ta + 1], mean_s[ta + 2], mean_s[ta + 3]);
      *reinterpret_cast<float4*>(out + e) = v;
    }
  } else
  for (int e = first; e < sp.N * HA; e += stride) {
    const int n = e / HA, ta = e - n * HA;
    const int t = ta / A, a = ta - t * A;
    float v = mean_s[ta] + std_s[ta] * sample_noise(sp, b, t, n, a, A);   // mean + std * noise, clamp(-1, 1)  (tdmpc.py:130-132)
    v = fminf(fmaxf(v, -1.0f), 1.0f);
    if (sp.best_a && n == 0) v = mean_s[ta];        // iCEM best-a (tdmpc_icem_similarity_mlp.py:213-214)
    out[e] = v;
  }
  if (reuse)
    for (int e = first; e < sp.E * HA; e += stride)
      out[(size_t)sp.N * HA + e] = reuse[(size_t)b * reuse_rows * HA + e];
  for (int e = first; e < sp.P * HA; e += stride)
    out[(size_t)(sp.N + sp.E) * HA + e] = pi_actions[(size_t)b * P0 * HA + e];
}

// Sampling alone, grid (B, nparts): used instead of the fused tail of select_kernel / init_kernel when there are too few
// problems to fill the GPU with one CTA each (B = 1: a single CTA would draw all N*H*A normals of an iteration).
// Same Philox addressing per element, so the rows are bit-identical to the fused path's.
__global__ void __launch_bounds__(NT) sample_kernel(SampleParams p) {
  __shared__ float mean_s[MAX_HA], std_s[MAX_HA];
  const int b = blockIdx.x, HA = p.H * p.A;
  for (int e = threadIdx.x; e < HA; e += NT) { mean_s[e] = p.mean[(size_t)b * HA + e]; std_s[e] = p.std[(size_t)b * HA + e]; }
  __syncthreads();
  sample_rows(b, p.spec, p.Tcap, p.H, p.A, mean_s, std_s, p.actions, p.pi_actions, p.P0, p.reuse, p.reuse_rows, blockIdx.y, gridDim.y);
}

__global__ void __launch_bounds__(NT_MAX) select_kernel(SelectParams p) {
  extern __shared__ __align__(16) unsigned char sraw[];
  const int b = blockIdx.x;
  const int T = p.T, K = p.K, H = p.H, A = p.A, HA = H * A;
  int n2 = 64;
  while (n2 < T) n2 <<= 1;
  float* key = reinterpret_cast<float*>(sraw);
  int* idx = reinterpret_cast<int*>(key + n2);
  float* w = reinterpret_cast<float*>(idx + n2);       // [K]
  float* mean_s = w + MAX_K;                            // [HA]
  float* std_s = mean_s + HA;                           // [HA]
  float* red = std_s + HA;                              // [32]
  // p.stage: the K elite rows [K][HA] gathered into shared memory by all threads at once.  The weighted reductions below are
  // one thread per (t, a) element walking the K elites in order (the order is part of the parity contract), so with few
  // elements (cartpole: H*A = 5) a handful of threads would otherwise wait out 2 K dependent global-memory round trips --
  // most of this kernel's time when a single environment is planned.  Same values, same order: bit-identical results.
  float* erow = red + 32;
  __shared__ int pick_j;

  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    key[i] = i < T ? p.values[(size_t)b * p.Tcap + i] : -INFINITY;
    idx[i] = i < T ? i : 0x7fffffff;
  }
  __syncthreads();
  bitonic_sort_desc(key, idx, n2);

  // score = exp(temperature * (v - max)); score /= score.sum()      (tdmpc.py:142-144)
  const float vmax = key[0];
  float sc = 0.0f;
  if (threadIdx.x < K) sc = expf(p.temperature * (key[threadIdx.x] - vmax));
  const float ssum = block_sum(sc, red);
  if (threadIdx.x < K) { sc = sc / ssum; w[threadIdx.x] = sc; }
  const float s2 = block_sum(threadIdx.x < K ? sc : 0.0f, red) + 1e-9f;   // score.sum(0) + 1e-9  (:145)
  if (threadIdx.x < K) {
    p.elite_idx[(size_t)b * K + threadIdx.x] = idx[threadIdx.x];
    if (p.score) p.score[(size_t)b * K + threadIdx.x] = sc;
    if (p.tap_elite_idx) p.tap_elite_idx[((size_t)p.iter * p.B + b) * K + threadIdx.x] = idx[threadIdx.x];
  }
  __syncthreads();

  // weighted mean / std over the elites, clamp, momentum                (tdmpc.py:145-149)
  const float* act_b = p.actions + (size_t)b * p.Tcap * HA;
  const float one_m = p.one_minus_momentum;
  const bool staged = p.stage != 0;
  if (staged) {
    for (int e = threadIdx.x; e < K * HA; e += blockDim.x) {
      const int k = e / HA;
      erow[e] = act_b[(size_t)idx[k] * HA + (e - k * HA)];
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < HA; e += blockDim.x) {
    float m = 0.0f;
    if (staged) for (int k = 0; k < K; ++k) m += w[k] * erow[k * HA + e];
    else for (int k = 0; k < K; ++k) m += w[k] * act_b[(size_t)idx[k] * HA + e];
    m = m / s2;
    float var = 0.0f;
    if (staged) for (int k = 0; k < K; ++k) { const float d = erow[k * HA + e] - m; var += w[k] * (d * d); }
    else for (int k = 0; k < K; ++k) { const float d = act_b[(size_t)idx[k] * HA + e] - m; var += w[k] * (d * d); }
    float sd = sqrtf(var / s2);
    sd = fminf(fmaxf(sd, p.std_floor), p.std_max);
    const float old = p.mean[(size_t)b * HA + e];
    const float mn = __fadd_rn(__fmul_rn(p.momentum, old), __fmul_rn(one_m, m));
    mean_s[e] = mn; std_s[e] = sd;
    p.mean[(size_t)b * HA + e] = mn;
    p.std[(size_t)b * HA + e] = sd;
    if (p.tap_mean) p.tap_mean[((size_t)p.iter * p.B + b) * HA + e] = mn;
    if (p.tap_std) p.tap_std[((size_t)p.iter * p.B + b) * HA + e] = sd;
  }
  if (p.elite_actions)
    for (int e = threadIdx.x; e < K * HA; e += blockDim.x) {
      const int k = e / HA;
      p.elite_actions[(size_t)b * K * HA + e] = staged ? erow[e] : act_b[(size_t)idx[k] * HA + (e - k * HA)];
    }
  __syncthreads();

  if (p.do_sample) {
    sample_rows(b, p.next, p.Tcap, H, A, mean_s, std_s, p.actions, p.pi_actions, p.P0, p.elite_actions, K);
    return;
  }
  if (!p.do_output) return;

  // ---- outputs (tdmpc.py:152-163)
  // j ~ Categorical(score): numpy's choice = inverse CDF on the float64 cumsum, side='right'
  if (threadIdx.x == 0) {
    const double u = p.u_pick ? p.u_pick[b] : philox_uniform64(p.pick, (uint64_t)b);
    double tot = 0.0;
    for (int k = 0; k < K; ++k) tot += (double)w[k];
    double c = 0.0; int j = 0;
    for (int k = 0; k < K; ++k) { c += (double)w[k]; if (c / tot <= u) j = k + 1; }
    pick_j = j < K ? j : K - 1;
  }
  float rsum = 0.0f;
  if (p.last_reward && threadIdx.x < NT)        // the first NT threads, NT apart: the same partial sums whatever the block size
    for (int i = threadIdx.x; i < T; i += NT) rsum += p.last_reward[(size_t)b * p.Tcap + i];
  rsum = block_sum(rsum, red);
  float ssd = 0.0f;
  for (int a = threadIdx.x; a < A; a += blockDim.x) ssd += std_s[a];
  ssd = block_sum(ssd, red);
  const int j = pick_j;
  for (int a = threadIdx.x; a < A; a += blockDim.x) {
    float v = staged ? erow[j * HA + a] : act_b[(size_t)idx[j] * HA + a];   // elite_actions[0, j]
    if (!p.eval_mode) v += std_s[a] * noise_at(p.out_noise, (uint64_t)b * A + a);   // a += std * randn(A), no clamp
    p.action[(size_t)b * A + a] = v;
  }
  for (int e = threadIdx.x; e < HA; e += blockDim.x) p.prev_mean[(size_t)b * HA + e] = mean_s[e];   // self._prev_mean = mean
  if (threadIdx.x == 0) {
    p.metrics[(size_t)b * 2 + 0] = rsum / (float)T;       // reward.mean() of the last step, last iteration
    p.metrics[(size_t)b * 2 + 1] = ssd / (float)A;        // std[0].mean()
  }
}

// mean/std initialisation with warm start (tdmpc.py:122-125) + sampling of iteration 0.
__global__ void __launch_bounds__(NT) init_kernel(InitParams p) {
  __shared__ float mean_s[MAX_HA], std_s[MAX_HA];
  const int b = blockIdx.x, H = p.H, A = p.A, HA = H * A;
  const bool t0 = p.t0_all >= 0 ? (p.t0_all != 0) : (p.t0[b] != 0);
  const bool warm = !t0 && p.has_prev;
  for (int e = threadIdx.x; e < HA; e += NT) {
    const int t = e / A;
    float m = 0.0f;
    if (warm) {
      if (t < H - 1) m = p.prev_mean[(size_t)b * HA + e + A];          // mean[:-1] = prev_mean[1:]
      else if (p.warm_last) m = p.prev_mean[(size_t)b * HA + e];        // iCEM: mean[-1] = prev_mean[-1]
    }
    mean_s[e] = m; std_s[e] = p.init_std;
    p.mean[(size_t)b * HA + e] = m; p.std[(size_t)b * HA + e] = p.init_std;
  }
  __syncthreads();
  // iteration-0 reused rows: previous step's elites advanced by one step; the tail that runs past
  // the stored horizon is sampled afresh from (mean, std) with its own noise stream
  const SampleSpec& sp = p.first;
  if (sp.E > 0) {
    float* out = p.actions + ((size_t)b * p.Tcap + sp.N) * HA;
    for (int e = threadIdx.x; e < sp.E * HA; e += NT) {
      const int k = e / HA, ta = e - k * HA;
      const int t = ta / A, a = ta - t * A;
      float v;
      if (!p.shift_elites) v = p.elite_actions[((size_t)b * p.K + k) * HA + ta];
      else if (t < p.elite_h - 1) v = p.elite_actions[((size_t)b * p.K + k) * (size_t)p.elite_h * A + (size_t)(t + 1) * A + a];
      else {
        const uint64_t row0 = (uint64_t)b * H * sp.E * A;
        float xi;
        if (p.shift.inj || !p.shift_mat) xi = noise_at(p.shift, row0 + ((uint64_t)t * sp.E + k) * A + a);
        else {
          xi = 0.0f;
          for (int j = 0; j < sp.color_len; ++j)
            xi = fmaf(__ldg(p.shift_mat + t * sp.color_len + j), philox_normal(p.shift, row0 + ((uint64_t)j * sp.E + k) * A + a), xi);
        }
        v = fminf(fmaxf(mean_s[ta] + std_s[ta] * xi, -1.0f), 1.0f);
      }
      out[e] = v;
    }
  }
  if (!p.skip_sample) sample_rows(b, sp, p.Tcap, H, A, mean_s, std_s, p.actions, p.pi_actions, p.P0, nullptr, 0);
}

__global__ void fill_noise_kernel(int stream_id, uint64_t seed, uint64_t call, uint64_t count, void* out) {
  NoiseSrc s{nullptr, seed, call, stream_id};
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
    if (stream_id == NS_PICK) reinterpret_cast<double*>(out)[i] = philox_uniform64(s, i);
    else reinterpret_cast<float*>(out)[i] = philox_normal(s, i);
  }
}

// W [N][K] (nn.Linear) -> Wt [Kp][N] at row offset koff, zero padded rows untouched (pre-zeroed).
__global__ void transpose_pack_kernel(const float* __restrict__ w, int N, int K, float* __restrict__ wt) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)N * K) {
    const int n = (int)(i / K), k = (int)(i - (size_t)n * K);
    wt[(size_t)k * N + n] = w[i];
  }
}

// eval-mode BatchNorm1d as y = x * scale + shift  (helper.py:179-184, 657-659; eps 1e-5)
__global__ void bn_affine_kernel(const float* g, const float* b, const float* rm, const float* rv, float eps, int n, float* scale, float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float s = g[i] / sqrtf(rv[i] + eps);
    scale[i] = s;
    shift[i] = b[i] - rm[i] * s;
  }
}

}  // namespace

void launch_bn_affine(const float* g, const float* b, const float* rm, const float* rv, float eps, int n, float* scale, float* shift, cudaStream_t st) {
  bn_affine_kernel<<<(n + 255) / 256, 256, 0, st>>>(g, b, rm, rv, eps, n, scale, shift);
}

// Elite rows are staged in shared memory (SelectParams::stage) when they are small (<= 16 KB: no occupancy cost) or when the
// batch is too small to hide the gather latency behind other CTAs; the opt-in limit is raised once per device.
constexpr size_t SELECT_SMEM_OPTIN = 160 * 1024;
int select_configure() {
  return cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SELECT_SMEM_OPTIN) == cudaSuccess ? 0 : -1;
}
void launch_select(const SelectParams& p0, int small_batch, cudaStream_t st) {
  SelectParams p = p0;
  int n2 = 64;
  while (n2 < p.T) n2 <<= 1;
  size_t smem = (size_t)n2 * 8 + sizeof(float) * (MAX_K + 2 * p.H * p.A + 32);
  const size_t rows = sizeof(float) * (size_t)p.K * p.H * p.A;
  p.stage = (rows <= 16384 || (small_batch && smem + rows <= SELECT_SMEM_OPTIN)) ? 1 : 0;
  if (p.stage) smem += rows;
  // One CTA per problem.  On a GPU-filling batch 256 threads (several CTAs per SM); on a small batch the kernel is a latency chain
  // on a handful of SMs -- 1024 threads make every pass of the bitonic sort one round instead of four (B = 1: 28 -> 14 us).
  // Every reduction is arranged so that both block sizes give bit-identical results.
  select_kernel<<<p.B, small_batch ? NT_MAX : NT, smem, st>>>(p);
}

__global__ void bump_counter_kernel(unsigned long long* c) { *c += 1ull; }
void launch_bump_counter(uint64_t* c, cudaStream_t st) { bump_counter_kernel<<<1, 1, 0, st>>>(reinterpret_cast<unsigned long long*>(c)); }
void launch_init(const InitParams& p, cudaStream_t st) { init_kernel<<<p.B, NT, 0, st>>>(p); }
void launch_sample(const SampleParams& p, int nparts, cudaStream_t st) { sample_kernel<<<dim3(p.B, nparts), NT, 0, st>>>(p); }

void launch_fill_noise(int stream_id, uint64_t seed, uint64_t call, uint64_t count, void* out, cudaStream_t st) {
  const int blocks = (int)((count + 255) / 256 < 4096 ? (count + 255) / 256 : 4096);
  fill_noise_kernel<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(stream_id, seed, call, count, out);
}

void launch_transpose_pack(const float* w, int N, int K, int Kp, int Kdst_off, float* wt, cudaStream_t st) {
  (void)Kp; (void)Kdst_off;
  const size_t n = (size_t)N * K;
  transpose_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, N, K, wt);
}

int select_max_rows() { return MAX_SORT; }
int select_max_elites() { return MAX_K; }
int select_max_ha() { return MAX_HA; }
