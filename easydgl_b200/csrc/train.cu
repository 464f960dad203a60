// train.cu - the training-mode forward of Sequential.train (EasyDGL.py:140-189, CTSMA.py:82-124, Base.py:119-131),
// i.e. everything between the encoder and the scalar `loss` (both dropout rates 0; the backward pass and the Adam
// step are not part of this library):
//   gather   tf.batch_gather(seqs_outs, masked_positions)                         EasyDGL.py:141-142
//   ce       -log(softmax(logits) + 1e-5)[label], weight = (label != 0)           EasyDGL.py:155,178-185
//   tpp      MAU.biased_likelihood(lam @ masked positions, mark(label), spans)    temporal.py:317-333, EasyDGL.py:159-175
//   l2       tf.losses.get_regularization_loss() = l2_reg * sum(table^2) / 2       coding.py:13-44
// Every reduction is a fixed-order tree in double precision, so the loss is reproducible bit for bit.
#include <math.h>

#include "common.cuh"

namespace edgl {

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ Y, const int64_t* __restrict__ pos,
                                                          int L, int M, int d, long long rows,
                                                          float* __restrict__ out, int* __restrict__ err) {
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const long long b = r / M;
  long long p = pos[r];
  if (p < 0 || p >= L) {  // tf.batch_gather raises on the CPU; report instead of reading out of bounds
    if (lane == 0) atomicExch(err, 1);
    p = 0;
  }
  const float4* src = reinterpret_cast<const float4*>(Y + (b * L + p) * d);
  float4* dst = reinterpret_cast<float4*>(out + r * d);
  for (int i = lane; i < d / 4; i += 32) dst[i] = src[i];
}

__device__ __forceinline__ float block_reduce_256(float v, float* red, bool is_max) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : v + u;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) t = is_max ? fmaxf(t, red[w]) : t + red[w];
  return t;
}

// one CTA per row of logits [rows, N] (pitch ld): pe[row0 + row] = weight * -log(softmax(x)[label] + 1e-5)
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logits, int ld, int N,
                                                      const int64_t* __restrict__ labels, long long row0,
                                                      float* __restrict__ pe, float* __restrict__ wt,
                                                      int* __restrict__ err) {
  __shared__ float red[8];
  const float* x = logits + (long long)blockIdx.x * ld;
  const long long row = row0 + blockIdx.x;
  long long lab = labels[row];
  if (lab < 0 || lab >= N) {  // tf.one_hot gives an all-zero row (loss 0); flag it - the labels are item ids
    if (threadIdx.x == 0) atomicExch(err, 2);
    lab = 0;
  }
  float m = -INFINITY;
  for (int i = threadIdx.x; i < N; i += 256) m = fmaxf(m, x[i]);
  m = block_reduce_256(m, red, true);
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += 256) s += expf(x[i] - m);
  s = block_reduce_256(s, red, false);
  if (threadIdx.x == 0) {
    const float p = expf(x[lab] - m) / s;                 // tf.nn.softmax
    const float w = lab != 0 ? 1.f : 0.f;                  // label_weights (EasyDGL.py:180)
    pe[row] = w * -logf(p + 1e-5f);                        // :155,182-183
    wt[row] = w;
  }
}

// One thread per (head * B + b, m): the three per-element terms of MAU.biased_likelihood.
//   positions != null (EasyDGL): lam row = positions[b,m], span = clip(t[p] - t[p-1], 0, 100) on UNSCALED timestamps,
//                                span[0] = span[1] (EasyDGL.py:161-163);
//   positions == null (CTSMA)  : lam row = m, span = t[m+1] - t[m], unclipped (CTSMA.py:100).
__global__ void __launch_bounds__(256) tpp_terms_kernel(const float* __restrict__ lam, const int64_t* __restrict__ positions,
                                                        const int64_t* __restrict__ labels,
                                                        const uint8_t* __restrict__ mark8, int mark_rows,
                                                        const float* __restrict__ ts, int ts_len, int B, int L, int M,
                                                        int heads, int E, float* __restrict__ ell,
                                                        float* __restrict__ nu, float* __restrict__ cnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)heads * B * M;
  if (i >= n) return;
  const int m = (int)(i % M);
  const long long hb = i / M;
  const int b = (int)(hb % B);
  long long lab = labels[(long long)b * M + m];
  if (lab < 0 || lab >= mark_rows) lab = 0;
  const uint8_t* nm = mark8 + lab * E;
  int p;
  float span;
  const float* t = ts + (long long)b * ts_len;
  if (positions) {
    p = (int)positions[(long long)b * M + m];
    p = p < 0 ? 0 : (p >= L ? L - 1 : p);
    const int q = p == 0 ? 1 : p;
    span = fminf(fmaxf(t[q] - t[q - 1], 0.f), 100.f);
  } else {
    p = m;
    span = t[m + 1] - t[m];
  }
  const float* l = lam + (hb * L + p) * E;
  float nsum = 0.f, ei = 0.f, tot = 0.f;
  for (int e = 0; e < E; ++e) {
    const float v = (float)nm[e];
    nsum += v;
    ei = fmaf(l[e], v, ei);
    tot += l[e];
  }
  const float sg = nsum > 0.f ? 1.f : 0.f;                // tf.sign of a non-negative sum (temporal.py:321)
  ei *= sg;
  tot *= sg;
  ell[i] = ei == 0.f ? 0.f : logf(ei);                    // log(where(ei == 0, 1, ei)) (:324)
  nu[i] = tot * span * 0.5f;                              // :327-328
  cnt[i] = nsum;                                          // :331
}

// out[slot] (double) = scale * sum(x[0..n)) (or of squares), one CTA, fixed order
__global__ void __launch_bounds__(1024) reduce_kernel(const float* __restrict__ x, long long n, int squares, double scale,
                                                      double* __restrict__ out, int accumulate) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) {
    const double v = (double)x[i];
    s += squares ? v * v : v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += red[w];
    *out = (accumulate ? *out : 0.0) + scale * t;
  }
}

// acc: [0] sum w*pe, [1] sum w, [2] l2, [3..3+3*nb) per block {sum ell, sum nu, sum cnt};  loss_out = {loss, ce, l2, ct}
__global__ void loss_combine_kernel(const double* __restrict__ acc, int num_blocks, double ct_scale,
                                    float* __restrict__ loss_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double ce = acc[0] / (acc[1] + 1e-5);              // EasyDGL.py:183-185
  double ct = 0.0;
  for (int i = 0; i < num_blocks; ++i) {
    const double* a = acc + 3 + 3 * i;
    ct += ct_scale * (-(a[0] - a[1]) / a[2]);              // temporal.py:332; EasyDGL.py:175 / CTSMA.py:110
  }
  loss_out[0] = (float)(ce + acc[2] + ct);
  loss_out[1] = (float)ce;
  loss_out[2] = (float)acc[2];
  loss_out[3] = (float)ct;
}

int launch_gather_rows(const float* Y, const int64_t* pos, int L, int M, int d, long long rows, float* out, int* err,
                       cudaStream_t st) {
  if (rows == 0) return 0;
  gather_rows_kernel<<<cdiv(rows * 32, 256), 256, 0, st>>>(Y, pos, L, M, d, rows, out, err);
  EDGL_LAUNCH_CHECK();
  return 0;
}
int launch_ce_rows(const float* logits, int ld, int N, const int64_t* labels, long long row0, int rows, float* pe,
                   float* wt, int* err, cudaStream_t st) {
  if (rows == 0) return 0;
  ce_rows_kernel<<<rows, 256, 0, st>>>(logits, ld, N, labels, row0, pe, wt, err);
  EDGL_LAUNCH_CHECK();
  return 0;
}
int launch_tpp_terms(const float* lam, const int64_t* positions, const int64_t* labels, const uint8_t* mark8,
                     int mark_rows, const float* ts, int ts_len, int B, int L, int M, int heads, int E, float* ell,
                     float* nu, float* cnt, cudaStream_t st) {
  const long long n = (long long)heads * B * M;
  if (n == 0) return 0;
  tpp_terms_kernel<<<cdiv(n, 256), 256, 0, st>>>(lam, positions, labels, mark8, mark_rows, ts, ts_len, B, L, M, heads, E,
                                                  ell, nu, cnt);
  EDGL_LAUNCH_CHECK();
  return 0;
}
int launch_reduce(const float* x, long long n, int squares, double scale, double* out, int accumulate, cudaStream_t st) {
  reduce_kernel<<<1, 1024, 0, st>>>(x, n, squares, scale, out, accumulate);
  EDGL_LAUNCH_CHECK();
  return 0;
}
int launch_loss_combine(const double* acc, int num_blocks, double ct_scale, float* loss_out, cudaStream_t st) {
  loss_combine_kernel<<<1, 32, 0, st>>>(acc, num_blocks, ct_scale, loss_out);
  EDGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace edgl
