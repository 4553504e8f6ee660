// The small dense parts of the iteration as hand-written kernels (they were ~450 cuBLAS / ATen
// launches per iteration): EqualisedLinear forward / backward for all layers of a pass in ONE
// launch (the `to_style` linears of a decode, the StyleExtractor head), the MappingNetwork with
// style mixing and the theta interpolation fused, and the style-cycle loss with its backward.
// K is 6 or 512 here -- far below a tensor-core tile: SIMT, fp32, one thread
// or one warp per output; the point is the launch count, not the FLOPs.
#include "common.cuh"

namespace otm {

// ---------------------------------------------------------------------------
// multi-job EqualisedLinear (reference layers.py:27-43): y = x @ (c W)^T + b, c = 1/sqrt(K)
// ---------------------------------------------------------------------------
struct LinJobs {
  otm_linear_job j[OTM_MAX_LINEAR_JOBS];
  int n_jobs;
  int out_begin[OTM_MAX_LINEAR_JOBS + 1];  // prefix sums of the per-job work-item counts
};

__device__ __forceinline__ int find_job(const LinJobs& J, int item) {
  int lo = 0;
#pragma unroll 1
  for (int i = 1; i < J.n_jobs; ++i)
    if (item >= J.out_begin[i]) lo = i;
  return lo;
}

// one thread per output element (small K)
__global__ void __launch_bounds__(256) linear_fwd_thread_kernel(const __grid_constant__ LinJobs J) {
  const int total = J.out_begin[J.n_jobs];
  for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
    const int ji = find_job(J, item);
    const otm_linear_job& jb = J.j[ji];
    const int r = item - J.out_begin[ji];
    const int n = r / jb.o, o = r - n * jb.o;
    const float* xr = jb.x + (long long)n * jb.x_row_stride;
    const float* wr = jb.w + (long long)o * jb.k;
    float acc = 0.f;
    for (int k = 0; k < jb.k; ++k) acc = fmaf(xr[k], wr[k], acc);
    jb.y[(long long)n * jb.o + o] = acc * rsqrtf((float)jb.k) + (jb.bias ? jb.bias[o] : 0.f);
  }
}

// one warp per output element (large K: the 512 -> w_dim StyleExtractor head)
__global__ void __launch_bounds__(256) linear_fwd_warp_kernel(const __grid_constant__ LinJobs J) {
  const int total = J.out_begin[J.n_jobs];
  const int lane = threadIdx.x % 32;
  const int warps = gridDim.x * (blockDim.x / 32);
  for (int item = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32; item < total; item += warps) {
    const int ji = find_job(J, item);
    const otm_linear_job& jb = J.j[ji];
    const int r = item - J.out_begin[ji];
    const int n = r / jb.o, o = r - n * jb.o;
    const float* xr = jb.x + (long long)n * jb.x_row_stride;
    const float* wr = jb.w + (long long)o * jb.k;
    float acc = 0.f;
    for (int k = lane; k < jb.k; k += 32) acc = fmaf(xr[k], wr[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) jb.y[(long long)n * jb.o + o] = acc * rsqrtf((float)jb.k) + (jb.bias ? jb.bias[o] : 0.f);
  }
}

// backward, parameters: dW[o,k] += c * sum_n dy[n,o] x[n,k] ;  db[o] += sum_n dy[n,o]
// one thread per (o, k) plus one per o for the bias; the n loop is short (a batch)
__global__ void __launch_bounds__(256) linear_bwd_param_kernel(const __grid_constant__ LinJobs J) {
  const int total = J.out_begin[J.n_jobs];
  for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
    const int ji = find_job(J, item);
    const otm_linear_job& jb = J.j[ji];
    if (!jb.dy) continue;
    const int r = item - J.out_begin[ji];
    const int o = r / (jb.k + 1), k = r - o * (jb.k + 1);
    float acc = 0.f;
    if (k == jb.k) {
      if (!jb.dbias) continue;
      for (int n = 0; n < jb.n; ++n) acc += jb.dy[(long long)n * jb.o + o];
      jb.dbias[o] += acc;
    } else {
      if (!jb.dw) continue;
      for (int n = 0; n < jb.n; ++n)
        acc = fmaf(jb.dy[(long long)n * jb.o + o], jb.x[(long long)n * jb.x_row_stride + k], acc);
      jb.dw[(long long)o * jb.k + k] += acc * rsqrtf((float)jb.k);
    }
  }
}

// backward, input: dx[n,k] += c * sum_o dy[n,o] W[o,k]  (atomic: several jobs may share one dx,
// e.g. the two to_style linears of a ModulatedResnetBlock read the same w[i]; a stride-0 x row
// = a broadcast input accumulates over n as well).  One warp per (n, k): lanes stride over o.
__global__ void __launch_bounds__(256) linear_bwd_input_kernel(const __grid_constant__ LinJobs J) {
  const int total = J.out_begin[J.n_jobs];
  const int lane = threadIdx.x % 32;
  const int warps = gridDim.x * (blockDim.x / 32);
  for (int item = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32; item < total; item += warps) {
    const int ji = find_job(J, item);
    const otm_linear_job& jb = J.j[ji];
    if (!jb.dy || !jb.dx) continue;
    const int r = item - J.out_begin[ji];
    const int n = r / jb.k, k = r - n * jb.k;
    float acc = 0.f;
    for (int o = lane; o < jb.o; o += 32) acc = fmaf(jb.dy[(long long)n * jb.o + o], jb.w[(long long)o * jb.k + k], acc);
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(jb.dx + (long long)n * jb.dx_row_stride + k, acc * rsqrtf((float)jb.k));
  }
}

// ---------------------------------------------------------------------------
// MappingNetwork (reference builder.py:16-132) with the style mixing and the domain-variable
// interpolation of get_single_w / get_two_w fused:
//   s = net(z / max(|z|, 1e-12)),  net = [Linear, LeakyReLU(0.2)] x (L-1), Linear, ReLU
//   w_j[blk, b, :] = d_j[b] * (blk < cross ? s(z1[b]) : s(z2[b]))            (lerp(0, s, d))
// One thread per sample; F <= 32 features, L <= 8 layers.
// ---------------------------------------------------------------------------
constexpr int MAP_F = OTM_MAX_STYLE_DIM;
constexpr int MAP_L = OTM_MAX_MAPPING_LAYERS;

struct MapP {
  const float* w[MAP_L];
  const float* b[MAP_L];
  float* dw[MAP_L];
  float* db[MAP_L];
  int F, L;
};

// forward of the net for one sample; act[l] = input of layer l (act[L] = output)
__device__ __forceinline__ void map_forward(const MapP& P, const float* z, float (&act)[MAP_L + 1][MAP_F]) {
  const int F = P.F;
  float ss = 0.f;
  for (int i = 0; i < F; ++i) ss += z[i] * z[i];
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize(z, dim=1)
  for (int i = 0; i < F; ++i) act[0][i] = z[i] * inv;
  const float c = rsqrtf((float)F);
  for (int l = 0; l < P.L; ++l) {
    for (int o = 0; o < F; ++o) {
      float a = 0.f;
      for (int k = 0; k < F; ++k) a = fmaf(act[l][k], P.w[l][o * F + k], a);
      a = a * c + P.b[l][o];
      act[l + 1][o] = (l == P.L - 1) ? fmaxf(a, 0.f) : (a > 0.f ? a : 0.2f * a);
    }
  }
}

// backward for one sample: g = dL/d(output).  Writes the layer inputs and the back-propagated
// deltas of this sample to shared memory ([L][spb][F] each); the CTA then reduces them over its
// samples without any contended atomic (96 threads adding into the same 84 parameters took 30 us
// with global atomics and 200 us with shared-memory CAS loops).
__device__ __forceinline__ void map_backward(const MapP& P, const float (&act)[MAP_L + 1][MAP_F],
                                             float (&g)[MAP_F], float* s_act, float* s_delta, int t,
                                             int spb) {
  const int F = P.F;
  const float c = rsqrtf((float)F);
  for (int l = P.L - 1; l >= 0; --l) {
    float delta[MAP_F];
    for (int o = 0; o < F; ++o) {
      const float y = act[l + 1][o];  // sign(y) == sign(pre-activation) for ReLU and LeakyReLU
      delta[o] = (l == P.L - 1) ? (y > 0.f ? g[o] : 0.f) : (y > 0.f ? g[o] : 0.2f * g[o]);
      s_delta[(l * spb + t) * F + o] = delta[o];
      s_act[(l * spb + t) * F + o] = act[l][o];
    }
    if (l > 0) {
      for (int k = 0; k < F; ++k) {
        float a = 0.f;
        for (int o = 0; o < F; ++o) a = fmaf(delta[o], P.w[l][o * F + k], a);
        g[k] = a * c;
      }
    }
  }
}

struct MapSampleP {
  MapP net;
  const float* z1;
  const float* z2;           // == z1: no style mixing
  const long long* cross;    // device scalar: blocks [0, cross) take s(z1); NULL = all z1
  const float* d[2];         // per-sample domain variable (NULL: d_const)
  float d_const[2];
  float* out[2];             // [n_blocks, B, F]; out[1] may be NULL
  const float* dout[2];      // backward: gradients w.r.t. out[j] (NULL = none)
  int B, n_blocks;
};

__global__ void __launch_bounds__(128) mapping_sample_fwd_kernel(const __grid_constant__ MapSampleP P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const int F = P.net.F;
  float act[MAP_L + 1][MAP_F];
  float s1[MAP_F], s2[MAP_F];
  map_forward(P.net, P.z1 + (long long)b * F, act);
  for (int i = 0; i < F; ++i) s1[i] = act[P.net.L][i];
  if (P.z2 != P.z1) {
    map_forward(P.net, P.z2 + (long long)b * F, act);
    for (int i = 0; i < F; ++i) s2[i] = act[P.net.L][i];
  } else {
    for (int i = 0; i < F; ++i) s2[i] = s1[i];
  }
  const int cross = P.cross ? (int)*P.cross : P.n_blocks;
  for (int j = 0; j < 2; ++j) {
    if (!P.out[j]) continue;
    const float d = P.d[j] ? P.d[j][b] : P.d_const[j];
    for (int blk = 0; blk < P.n_blocks; ++blk)
      for (int i = 0; i < F; ++i)
        P.out[j][((long long)blk * P.B + b) * F + i] = d * (blk < cross ? s1[i] : s2[i]);
  }
}

__global__ void __launch_bounds__(128) mapping_sample_bwd_kernel(const __grid_constant__ MapSampleP P,
                                                                 int spb) {
  extern __shared__ float map_sm[];  // [L][spb][F] layer inputs, [L][spb][F] deltas
  const int F = P.net.F, L = P.net.L;
  float* s_act = map_sm;
  float* s_delta = map_sm + L * spb * F;
  const int t = threadIdx.x;
  const int b = blockIdx.x * spb + t;
  const bool live = t < spb && b < P.B;
  const int n_live = min(spb, P.B - blockIdx.x * spb);
  const float c = rsqrtf((float)F);
  float g1[MAP_F], g2[MAP_F];
  if (live) {
    const int cross = P.cross ? (int)*P.cross : P.n_blocks;
    for (int i = 0; i < F; ++i) { g1[i] = 0.f; g2[i] = 0.f; }
    for (int j = 0; j < 2; ++j) {
      if (!P.dout[j]) continue;
      const float d = P.d[j] ? P.d[j][b] : P.d_const[j];
      for (int blk = 0; blk < P.n_blocks; ++blk)
        for (int i = 0; i < F; ++i) {
          const float g = d * P.dout[j][((long long)blk * P.B + b) * F + i];
          if (blk < cross) g1[i] += g; else g2[i] += g;
        }
    }
    if (P.z2 == P.z1)
      for (int i = 0; i < F; ++i) g1[i] += g2[i];
  }
  const int passes = P.z2 == P.z1 ? 1 : 2;
  for (int pass = 0; pass < passes; ++pass) {
    if (live) {
      float act[MAP_L + 1][MAP_F];
      map_forward(P.net, (pass == 0 ? P.z1 : P.z2) + (long long)b * F, act);
      if (pass == 0) map_backward(P.net, act, g1, s_act, s_delta, t, spb);
      else map_backward(P.net, act, g2, s_act, s_delta, t, spb);
    }
    __syncthreads();
    // dW[l][o][k] += c * sum_b delta[l][b][o] * in[l][b][k] ;  db[l][o] += sum_b delta[l][b][o]
    for (int i = t; i < L * F * (F + 1); i += blockDim.x) {
      const int l = i / (F * (F + 1)), r = i - l * F * (F + 1);
      const int o = r / (F + 1), k = r - o * (F + 1);
      float acc = 0.f;
      if (k == F) {
        for (int bb = 0; bb < n_live; ++bb) acc += s_delta[(l * spb + bb) * F + o];
        if (acc != 0.f) atomicAdd(P.net.db[l] + o, acc);
      } else {
        for (int bb = 0; bb < n_live; ++bb)
          acc = fmaf(s_delta[(l * spb + bb) * F + o], s_act[(l * spb + bb) * F + k], acc);
        if (acc != 0.f) atomicAdd(P.net.dw[l] + o * F + k, acc * c);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// style_cycle_loss_func (reference loss.py:60-75) on [B, F] tensors:
//   ah = a / max(|a|, 1e-12), bh likewise (F.normalize);
//   cos = (ah / max(|ah|, 1e-8)) . (bh / max(|bh|, 1e-8))        (F.cosine_similarity)
//   loss = 1 - mean_b cos + ratio * mean_{b,i} (ah - bh)^2
// One CTA; thread b handles sample b, writes scale * dloss/da and /db, block-reduces the loss.
// ---------------------------------------------------------------------------
struct StyleCycleP {
  const float* a;
  const float* b;
  long long a_stride, b_stride;
  float* da;  // [B, F] dense or NULL
  float* db;
  float* out;
  float ratio, scale;
  int B, F;
};

// v = x / max(|x|, eps); returns the clamped norm
__device__ __forceinline__ float normalize_to(const float* x, int F, float eps, float (&v)[MAP_F]) {
  float ss = 0.f;
  for (int i = 0; i < F; ++i) ss += x[i] * x[i];
  const float nrm = fmaxf(sqrtf(ss), eps);
  for (int i = 0; i < F; ++i) v[i] = x[i] / nrm;
  return nrm;
}
// backward of v = x / max(|x|, eps) given dL/dv in g (in place -> dL/dx); `clamped` = |x| <= eps
__device__ __forceinline__ void normalize_bwd(const float (&v)[MAP_F], int F, float nrm, bool clamped,
                                              float (&g)[MAP_F]) {
  float dot = 0.f;
  if (!clamped)
    for (int i = 0; i < F; ++i) dot += g[i] * v[i];
  for (int i = 0; i < F; ++i) g[i] = (g[i] - (clamped ? 0.f : v[i] * dot)) / nrm;
}

__global__ void __launch_bounds__(256) style_cycle_kernel(const __grid_constant__ StyleCycleP P) {
  const int F = P.F;
  float part = 0.f;
  for (int b = threadIdx.x; b < P.B; b += blockDim.x) {
    const float* a = P.a + (long long)b * P.a_stride;
    const float* bb = P.b + (long long)b * P.b_stride;
    float ah[MAP_F], bh[MAP_F], au[MAP_F], bu[MAP_F];
    float ssa = 0.f, ssb = 0.f;
    for (int i = 0; i < F; ++i) { ssa += a[i] * a[i]; ssb += bb[i] * bb[i]; }
    const bool ca = sqrtf(ssa) <= 1e-12f, cb = sqrtf(ssb) <= 1e-12f;
    const float na = normalize_to(a, F, 1e-12f, ah);
    const float nb = normalize_to(bb, F, 1e-12f, bh);
    float s2a = 0.f, s2b = 0.f;
    for (int i = 0; i < F; ++i) { s2a += ah[i] * ah[i]; s2b += bh[i] * bh[i]; }
    const bool cua = sqrtf(s2a) <= 1e-8f, cub = sqrtf(s2b) <= 1e-8f;
    const float nua = normalize_to(ah, F, 1e-8f, au);
    const float nub = normalize_to(bh, F, 1e-8f, bu);
    float cosv = 0.f, mse = 0.f;
    for (int i = 0; i < F; ++i) { cosv += au[i] * bu[i]; const float d = ah[i] - bh[i]; mse += d * d; }
    part += -cosv / (float)P.B + P.ratio * mse / (float)(P.B * F);
    if (P.da || P.db) {
      // d loss / d au = -bu / B ; through the cosine's own normalisation, plus the mse term
      float ga[MAP_F], gb[MAP_F];
      for (int i = 0; i < F; ++i) { ga[i] = -bu[i] / (float)P.B; gb[i] = -au[i] / (float)P.B; }
      normalize_bwd(au, F, nua, cua, ga);
      normalize_bwd(bu, F, nub, cub, gb);
      const float km = 2.f * P.ratio / (float)(P.B * F);
      for (int i = 0; i < F; ++i) { const float d = ah[i] - bh[i]; ga[i] += km * d; gb[i] -= km * d; }
      normalize_bwd(ah, F, na, ca, ga);
      normalize_bwd(bh, F, nb, cb, gb);
      for (int i = 0; i < F; ++i) {
        if (P.da) P.da[(long long)b * F + i] = P.scale * ga[i];
        if (P.db) P.db[(long long)b * F + i] = P.scale * gb[i];
      }
    }
  }
  __shared__ float red[8];
  part = warp_sum(part);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) P.out[0] = 1.f + v;
  }
}

static int make_jobs(const otm_linear_job* jobs, int n_jobs, int mode, LinJobs* J, const char* what) {
  OTM_REQUIRE(jobs && n_jobs >= 1 && n_jobs <= OTM_MAX_LINEAR_JOBS, "%s: 1..%d jobs", what,
              OTM_MAX_LINEAR_JOBS);
  J->n_jobs = n_jobs;
  J->out_begin[0] = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const otm_linear_job& j = jobs[i];
    OTM_REQUIRE(j.x && j.w && j.n >= 1 && j.k >= 1 && j.o >= 1, "%s: job %d has a null/empty operand", what, i);
    J->j[i] = j;
    long long items = mode == 0 ? (long long)j.n * j.o : mode == 1 ? (long long)j.o * (j.k + 1)
                                                                  : (long long)j.n * j.k;
    OTM_REQUIRE(J->out_begin[i] + items < (1ll << 30), "%s: too large", what);
    J->out_begin[i + 1] = J->out_begin[i] + (int)items;
  }
  return OTM_OK;
}

}  // namespace otm

using namespace otm;

extern "C" {

int otm_linear_fwd(const otm_linear_job* jobs, int32_t n_jobs, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  LinJobs J;
  int rc = make_jobs(jobs, n_jobs, 0, &J, "linear_fwd");
  if (rc) return rc;
  int kmax = 0;
  for (int i = 0; i < n_jobs; ++i) {
    OTM_REQUIRE(jobs[i].y, "linear_fwd: job %d has no output", i);
    kmax = jobs[i].k > kmax ? jobs[i].k : kmax;
  }
  const int total = J.out_begin[n_jobs];
  if (kmax >= 64) {
    int grid = (total + 7) / 8;
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    linear_fwd_warp_kernel<<<grid, 256, 0, st>>>(J);
  } else {
    int grid = (total + 255) / 256;
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    linear_fwd_thread_kernel<<<grid, 256, 0, st>>>(J);
  }
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

int otm_linear_bwd(const otm_linear_job* jobs, int32_t n_jobs, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  LinJobs J;
  int rc = make_jobs(jobs, n_jobs, 1, &J, "linear_bwd");
  if (rc) return rc;
  bool any_param = false, any_input = false;
  for (int i = 0; i < n_jobs; ++i) {
    any_param |= jobs[i].dy && (jobs[i].dw || jobs[i].dbias);
    any_input |= jobs[i].dy && jobs[i].dx;
  }
  if (any_param) {
    const int total = J.out_begin[n_jobs];
    int grid = (total + 255) / 256;
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    linear_bwd_param_kernel<<<grid, 256, 0, st>>>(J);
    OTM_LAUNCH_CHECK();
  }
  if (any_input) {
    rc = make_jobs(jobs, n_jobs, 2, &J, "linear_bwd");
    if (rc) return rc;
    const int total = J.out_begin[n_jobs];
    int grid = (total + 7) / 8;
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    linear_bwd_input_kernel<<<grid, 256, 0, st>>>(J);
    OTM_LAUNCH_CHECK();
  }
  return OTM_OK;
}

static int fill_map(const otm_mapping_args* a, MapSampleP* P, bool bwd) {
  OTM_REQUIRE(a && a->z1 && a->features >= 1 && a->features <= OTM_MAX_STYLE_DIM && a->n_layers >= 1 &&
                  a->n_layers <= OTM_MAX_MAPPING_LAYERS && a->batch >= 1 && a->n_blocks >= 1,
              "mapping: features <= %d, layers <= %d", OTM_MAX_STYLE_DIM, OTM_MAX_MAPPING_LAYERS);
  P->net.F = a->features;
  P->net.L = a->n_layers;
  for (int l = 0; l < a->n_layers; ++l) {
    OTM_REQUIRE(a->w[l] && a->b[l], "mapping: layer %d has no parameters", l);
    P->net.w[l] = a->w[l]; P->net.b[l] = a->b[l];
    P->net.dw[l] = a->dw[l]; P->net.db[l] = a->db[l];
    if (bwd) OTM_REQUIRE(a->dw[l] && a->db[l], "mapping_bwd: layer %d has no gradient buffers", l);
  }
  P->z1 = a->z1;
  P->z2 = a->z2 ? a->z2 : a->z1;
  P->cross = (const long long*)a->cross;
  for (int j = 0; j < 2; ++j) {
    P->d[j] = a->d[j]; P->d_const[j] = a->d_const[j];
    P->out[j] = a->out[j]; P->dout[j] = a->dout[j];
  }
  P->B = a->batch;
  P->n_blocks = a->n_blocks;
  return OTM_OK;
}

int otm_mapping_fwd(const otm_mapping_args* a, otm_stream stream) {
  MapSampleP P;
  int rc = fill_map(a, &P, false);
  if (rc) return rc;
  OTM_REQUIRE(a->out[0], "mapping_fwd: no output");
  mapping_sample_fwd_kernel<<<(a->batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

int otm_mapping_bwd(const otm_mapping_args* a, otm_stream stream) {
  MapSampleP P;
  int rc = fill_map(a, &P, true);
  if (rc) return rc;
  OTM_REQUIRE(a->dout[0] || a->dout[1], "mapping_bwd: no upstream gradient");
  // samples per CTA: as many as fit 40 KB of [L][spb][F] x 2 staging (128 at the default 6 x 2)
  int spb = (40 * 1024) / (int)(sizeof(float) * 2 * a->n_layers * a->features);
  if (spb > 128) spb = 128;
  const size_t smem = sizeof(float) * 2 * a->n_layers * spb * a->features;
  mapping_sample_bwd_kernel<<<(a->batch + spb - 1) / spb, 128, smem, (cudaStream_t)stream>>>(P, spb);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

int otm_loss_style_cycle(const float* a, int64_t a_stride, const float* b, int64_t b_stride,
                         int32_t batch, int32_t features, float ratio, float scale, float* out,
                         float* da, float* db, otm_stream stream) {
  OTM_REQUIRE(a && b && out && batch >= 1 && features >= 1 && features <= OTM_MAX_STYLE_DIM,
              "loss_style_cycle: features <= %d", OTM_MAX_STYLE_DIM);
  StyleCycleP P{a, b, a_stride, b_stride, da, db, out, ratio, scale, batch, features};
  style_cycle_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(P);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

}  // extern "C"
