// norm.cu - Base.layernorm (Base.py:12-67).
// begin_norm_axis=1: mean / population variance over (L,C) JOINTLY per sample (Base.py:51-52, Q2),
// eps 1e-12, gamma/beta over the last axis, tf.nn.batch_normalization form
//   out = x * inv + (beta - mean * inv),  inv = rsqrt(var + eps) * gamma   (Base.py:57-63).
// One CTA per sequence; the sample is cached in shared memory when it fits so HBM sees one read and
// one write per element; two-pass (mean, then squared deviations) like tf.nn.moments.
#include "common.cuh"

namespace edgl {

__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect red[] reuse
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < 8) ? red[lane] : 0.f;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return __shfl_sync(0xffffffffu, t, 0);
}

template <bool CACHE>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int L, int C,
                                                        float* __restrict__ out, int last_only,
                                                        unsigned int* __restrict__ out_amax, long long out_stride) {
  extern __shared__ __align__(16) float xs[];
  __shared__ float red[8];
  const long long n = (long long)L * C;
  const float* xp = x + (long long)blockIdx.x * n;
  const int n4 = (int)(n >> 2);  // n % 4 == 0 is guaranteed by the launcher
  float s = 0.f;
  for (int i = threadIdx.x; i < n4; i += 256) {
    const float4 v = reinterpret_cast<const float4*>(xp)[i];
    if (CACHE) reinterpret_cast<float4*>(xs)[i] = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = block_sum_256(s, red) / (float)n;
  float q = 0.f;
  for (int i = threadIdx.x; i < n4; i += 256) {
    const float4 v = CACHE ? reinterpret_cast<const float4*>(xs)[i] : reinterpret_cast<const float4*>(xp)[i];
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float var = block_sum_256(q, red) / (float)n;
  const float rstd = __frcp_rn(sqrtf(var + 1e-12f));
  float omax = 0.f;  // max |out| over what this thread writes
  if (last_only) {
    const float* src = (CACHE ? xs : xp) + (long long)(L - 1) * C;
    float* op = out + (long long)blockIdx.x * out_stride;
    for (int c = threadIdx.x; c < C; c += 256) {
      const float inv = rstd * gamma[c];
      const float o = src[c] * inv + (beta[c] - mean * inv);
      op[c] = o;
      omax = fmaxf(omax, fabsf(o));
    }
  } else {
    float* op = out + (long long)blockIdx.x * n;
    const int c4n = C >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) {
      const float4 v = CACHE ? reinterpret_cast<const float4*>(xs)[i] : reinterpret_cast<const float4*>(xp)[i];
      const int c = (i % c4n) * 4;
      const float4 g = *reinterpret_cast<const float4*>(gamma + c);
      const float4 bt = *reinterpret_cast<const float4*>(beta + c);
      float4 o;
      float inv;
      inv = rstd * g.x; o.x = v.x * inv + (bt.x - mean * inv);
      inv = rstd * g.y; o.y = v.y * inv + (bt.y - mean * inv);
      inv = rstd * g.z; o.z = v.z * inv + (bt.z - mean * inv);
      inv = rstd * g.w; o.w = v.w * inv + (bt.w - mean * inv);
      reinterpret_cast<float4*>(op)[i] = o;
      omax = fmaxf(omax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
    }
  }
  if (out_amax) amax_publish(out_amax, omax, threadIdx.x & 31);
}

// ---- LayerNorm folded into the tensor-core dense layers (LnEpi, common.cuh)
// One warp per sequence: the L * nparts row partials are summed in a fixed order in double precision;
// var = E[x^2] - mean^2 (clamped at 0), rstd = 1 / sqrt(var + 1e-12)  (Base.py:51-56).  Optionally the warp also
// writes y = LN(x_last) for the sequence's stored last row (EasyDGL.py:139,146).
__global__ void __launch_bounds__(256) ln_finalize_kernel(const float2* __restrict__ parts, int nparts, int B, int L,
                                                          int C, float2* __restrict__ rs,
                                                          const float* __restrict__ x_last,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ y,
                                                          unsigned int* __restrict__ y_amax, long long y_stride) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float2* p = parts + (size_t)warp * L * nparts;
  const int n = L * nparts;
  double s = 0.0, q = 0.0;
  for (int i = lane; i < n; i += 32) {
    const float2 v = p[i];
    s += (double)v.x;
    q += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const double cnt = (double)L * (double)C;
  const double mean = s / cnt;
  double var = q / cnt - mean * mean;
  if (!(var > 0.0)) var = 0.0;
  const float meanf = (float)mean;
  const float rstd = (float)(1.0 / sqrt(var + 1e-12));
  if (lane == 0) rs[warp] = make_float2(meanf, rstd);
  if (y) {
    float omax = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float inv = rstd * gamma[c];
      const float o = x_last[(size_t)warp * C + c] * inv + (beta[c] - meanf * inv);
      y[(size_t)warp * y_stride + c] = o;
      omax = fmaxf(omax, fabsf(o));
    }
    if (y_amax) amax_publish(y_amax, omax, lane);
  }
}

int launch_ln_finalize(const float2* parts, int nparts, int B, int L, int C, float2* rs, const float* x_last,
                       const float* gamma, const float* beta, float* y, unsigned int* y_amax, cudaStream_t st,
                       long long y_stride) {
  if (B == 0) return 0;
  if (y_stride == 0) y_stride = C;
  ln_finalize_kernel<<<cdiv((long long)B * 32, 256), 256, 0, st>>>(parts, nparts, B, L, C, rs, x_last, gamma, beta, y,
                                                                     y_amax, y_stride);
  EDGL_LAUNCH_CHECK();
  return 0;
}

int launch_layernorm(const float* x, const float* gamma, const float* beta, int B, int L, int C, float* out,
                     bool last_only, cudaStream_t st, unsigned int* out_amax, long long out_stride) {
  if (out_stride == 0) out_stride = C;
  EDGL_REQUIRE(C % 4 == 0, "layernorm: channel count must be a multiple of 4 (got %d)", C);
  EDGL_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(beta) & 15) == 0,
               "layernorm: pointers must be 16-byte aligned");
  if (B == 0) return 0;
  const size_t bytes = (size_t)L * C * sizeof(float);
  if (bytes <= 200 * 1024) {
    auto kern = layernorm_kernel<true>;
    EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    kern<<<B, 256, bytes, st>>>(x, gamma, beta, L, C, out, last_only ? 1 : 0, out_amax, out_stride);
  } else {
    layernorm_kernel<false><<<B, 256, 0, st>>>(x, gamma, beta, L, C, out, last_only ? 1 : 0, out_amax, out_stride);
  }
  EDGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace edgl
