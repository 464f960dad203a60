// time_attn.cu - the attention cores of the reference's time-aware baseline layers:
//   TiMultiHeadAttention  temporal.py:15-109   (TiSASRec: learned position + clipped-interval embeddings on K and V)
//   TfMultiHeadAttention  temporal.py:112-185  (TGAT: position code + Bochner / Mercer time kernel cos(w dt + phi) on K)
//   TgMultiHeadAttention  temporal.py:188-264  (TGSRec: keys and values are projections of [key | time code(q, k)])
// One kernel, exact fp32 on the CUDA cores (these layers are not on the north-star path: parity first).  The pairwise
// [B, Tq, Tk, C] code tensors of the reference are never materialised: a code element is looked up / evaluated where
// it is consumed.
//
//   s[q,k] = (Q_q . K_k + Q_q . PK_k + time_k(q,k)) / sqrt(dh)       masks: key padding, causality  -> softmax P
//   o[q]   = sum_k P[q,k] (V_k + PV_k + time_v(q,k)) + residual[q]    (Ti: P *= query mask first)
//   mode 1 (Ti): time_k = Q_q . TK[iv[q,k]], time_v = TV[iv[q,k]]     (tables [vocab, C], iv int64)
//   mode 2 (Tf): time_k = Q_q . cos(iv[q,k] f + phi)  (head slice of the C code dims), no time_v
//   mode 3 (Tg): time_k = U_{q,head} . cos(iv[q,k] f + phi) over ALL C code dims, U = W_k2[:, head] Q_q[head] (the time
//                half of the K projection applied to the query side); the value side returns
//                TC[q,head,:] = sum_k P[q,k] cos(iv[q,k] f + phi), which the caller multiplies by W_v2[:, head]
// One CTA per (sequence, head): the head slices of K, V, PK, PV live in shared memory, each warp takes query rows in
// turn, lanes split the keys.
#include <math.h>

#include "common.cuh"

namespace edgl {

namespace {

constexpr float kFillT = -4294967296.0f;  // float(-2**32+1), temporal.py:69,78

struct TimeAttnArgs {
  const float *Q, *K, *V;          // [B,Tq,C], [B,Tk,C], [B,Tk,C]
  const uint8_t *kmask, *qmask;    // [B,Tk] (null = no key masking), [B,Tq] or null
  const float *pos_k, *pos_v;      // [Tk,C] or null
  int mode;
  const int64_t* iv_i;             // mode 1: [B,Tq,Tk]
  const float* iv_f;               // modes 2, 3
  const float *tk, *tv;            // mode 1 tables [vocab,C]
  int vocab;
  const float *freq, *phase;       // [C]
  const float* U;                  // mode 3: [B,Tq,h,C]
  float* TC;                       // mode 3: [B,Tq,h,C]
  const float* R;                  // residual [B,Tq,C] or null
  float* O;                        // [B,Tq,C]
  int B, Tq, Tk, C, h, causal;
};

template <int DH>
__global__ void __launch_bounds__(128) time_attention_kernel(TimeAttnArgs a) {
  extern __shared__ float sm[];
  const int Tk = a.Tk, Tq = a.Tq, C = a.C;
  float* Ks = sm;                      // [Tk][DH] = K + PK
  float* Vs = Ks + (size_t)Tk * DH;    // [Tk][DH] = V + PV
  float* sc = Vs + (size_t)Tk * DH;    // [4 warps][Tk] scores / probabilities
  float* fr = sc + 4 * (size_t)Tk;     // modes 2, 3: freq, phase of the code dims this CTA needs
  const int b = blockIdx.x / a.h, hh = blockIdx.x % a.h;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ncode = a.mode == 3 ? C : DH;          // code dims used per (q, k)
  const int code0 = a.mode == 3 ? 0 : hh * DH;     // first code dim
  float* ph = fr + ncode;
  for (int i = tid; i < Tk * DH; i += 128) {
    const int k = i / DH, d = i % DH;
    const size_t g = ((size_t)b * Tk + k) * C + hh * DH + d;
    Ks[i] = a.K[g] + (a.pos_k ? a.pos_k[(size_t)k * C + hh * DH + d] : 0.f);
    Vs[i] = a.V[g] + (a.pos_v ? a.pos_v[(size_t)k * C + hh * DH + d] : 0.f);
  }
  if (a.mode >= 2)
    for (int i = tid; i < ncode; i += 128) {
      fr[i] = a.freq[code0 + i];
      ph[i] = a.phase[code0 + i];
    }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)DH);  // temporal.py:62
  float* my = sc + (size_t)warp * Tk;
  for (int q = warp; q < Tq; q += 4) {
    const size_t qrow = (size_t)b * Tq + q;
    float Qr[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) Qr[d] = a.Q[qrow * C + hh * DH + d];
    const float* Uq = a.mode == 3 ? a.U + (qrow * a.h + hh) * C : nullptr;
    // ---- scores
    float m = -INFINITY;
    for (int k = lane; k < Tk; k += 32) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) s = fmaf(Qr[d], Ks[k * DH + d], s);
      const size_t pair = qrow * Tk + k;
      if (a.mode == 1) {
        long long iv = a.iv_i[pair];
        iv = iv < 0 ? 0 : (iv >= a.vocab ? a.vocab - 1 : iv);
        const float* tr = a.tk + (size_t)iv * C + hh * DH;
#pragma unroll
        for (int d = 0; d < DH; ++d) s = fmaf(Qr[d], tr[d], s);
      } else if (a.mode == 2) {
        const float x = a.iv_f[pair];
#pragma unroll
        for (int d = 0; d < DH; ++d) s = fmaf(Qr[d], cosf(fmaf(x, fr[d], ph[d])), s);
      } else if (a.mode == 3) {
        const float x = a.iv_f[pair];
        for (int c = 0; c < C; ++c) s = fmaf(Uq[c], cosf(fmaf(x, fr[c], ph[c])), s);
      }
      s *= scale;
      if (a.kmask && !a.kmask[(size_t)b * Tk + k]) s = kFillT;
      if (a.causal && k > q) s = kFillT;
      my[k] = s;
      m = fmaxf(m, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
    for (int k = lane; k < Tk; k += 32) {
      const float p = expf(my[k] - m);
      my[k] = p;
      l += p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    float inv = 1.0f / l;
    if (a.qmask && !a.qmask[qrow]) inv = 0.f;  // temporal.py:87-90: query masking (Ti)
    __syncwarp();
    // ---- weighted sum
    float acc[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) acc[d] = 0.f;
    for (int k = lane; k < Tk; k += 32) {
      const float p = my[k] * inv;
#pragma unroll
      for (int d = 0; d < DH; ++d) acc[d] = fmaf(p, Vs[k * DH + d], acc[d]);
      if (a.mode == 1 && a.tv) {
        long long iv = a.iv_i[qrow * Tk + k];
        iv = iv < 0 ? 0 : (iv >= a.vocab ? a.vocab - 1 : iv);
        const float* tr = a.tv + (size_t)iv * C + hh * DH;
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] = fmaf(p, tr[d], acc[d]);
      }
    }
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      float v = acc[d];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[d] = v;
    }
    if (lane == 0) {
#pragma unroll
      for (int d = 0; d < DH; ++d) {
        const size_t g = qrow * C + hh * DH + d;
        a.O[g] = acc[d] + (a.R ? a.R[g] : 0.f);
      }
    }
    if (a.mode == 3) {  // TC[q, head, c] = sum_k P[q,k] cos(iv f_c + phi_c): lanes split the code dims
      float* tc = a.TC + (qrow * a.h + hh) * C;
      for (int c = lane; c < C; c += 32) {
        float t = 0.f;
        for (int k = 0; k < Tk; ++k) t = fmaf(my[k] * inv, cosf(fmaf(a.iv_f[qrow * Tk + k], fr[c], ph[c])), t);
        tc[c] = t;
      }
    }
    __syncwarp();
  }
}

__global__ void row_nonzero_kernel(const float* __restrict__ x, long long rows, int C, uint8_t* __restrict__ out) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += fabsf(x[r * C + c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = s != 0.f ? 1 : 0;  // tf.sign(tf.reduce_sum(tf.abs(x), -1)), temporal.py:65
}

// module.normalize.layernorm (normalize.py:9-19): moments over the LAST axis, (x - mean) / sqrt(var + eps) * gamma + beta
__global__ void rownorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                               const float* __restrict__ beta, long long rows, int C, float eps, float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* p = x + r * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += p[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = p[c] - mean;
    v = fmaf(d, d, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float den = sqrtf(v / (float)C + eps);
  for (int c = lane; c < C; c += 32) out[r * C + c] = gamma[c] * ((p[c] - mean) / den) + beta[c];
}

template <int DH>
int launch_t(const TimeAttnArgs& a, cudaStream_t st) {
  const int ncode = a.mode == 3 ? a.C : DH;
  const size_t smem = ((size_t)2 * a.Tk * DH + 4 * (size_t)a.Tk + 2 * (size_t)ncode) * sizeof(float);
  EDGL_REQUIRE(smem <= 227 * 1024, "time_attention: Tk = %d does not fit shared memory", a.Tk);
  auto kern = time_attention_kernel<DH>;
  EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)(a.B * a.h), 128, smem, st>>>(a);
  EDGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int launch_row_nonzero(const float* x, long long rows, int C, uint8_t* out, cudaStream_t st) {
  if (rows == 0) return 0;
  row_nonzero_kernel<<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(x, rows, C, out);
  EDGL_LAUNCH_CHECK();
  return 0;
}

int launch_rownorm(const float* x, const float* gamma, const float* beta, long long rows, int C, float eps, float* out,
                   cudaStream_t st) {
  if (rows == 0) return 0;
  rownorm_kernel<<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(x, gamma, beta, rows, C, eps, out);
  EDGL_LAUNCH_CHECK();
  return 0;
}

int launch_time_attention(const float* Q, const float* K, const float* V, const uint8_t* kmask, const uint8_t* qmask,
                          const float* pos_k, const float* pos_v, int mode, const void* intervals, const float* tk,
                          const float* tv, int vocab, const float* freq, const float* phase, const float* U, float* TC,
                          const float* R, int B, int Tq, int Tk, int C, int h, int causal, float* out, cudaStream_t st) {
  EDGL_REQUIRE(h >= 1 && C % h == 0, "time_attention: num_units %d not divisible by num_heads %d", C, h);
  EDGL_REQUIRE(mode >= 0 && mode <= 3, "time_attention: unknown time mode %d", mode);
  EDGL_REQUIRE(mode == 0 || intervals, "time_attention: intervals missing");
  EDGL_REQUIRE(mode != 1 || (tk && vocab >= 1), "time_attention: interval table missing");
  EDGL_REQUIRE(mode < 2 || (freq && phase), "time_attention: basis_freq / phase missing");
  EDGL_REQUIRE(mode != 3 || (U && TC), "time_attention: mode 3 needs U and TC");
  if (B == 0 || Tq == 0) return 0;
  TimeAttnArgs a;
  a.Q = Q; a.K = K; a.V = V; a.kmask = kmask; a.qmask = qmask; a.pos_k = pos_k; a.pos_v = pos_v; a.mode = mode;
  a.iv_i = mode == 1 ? static_cast<const int64_t*>(intervals) : nullptr;
  a.iv_f = mode >= 2 ? static_cast<const float*>(intervals) : nullptr;
  a.tk = tk; a.tv = tv; a.vocab = vocab; a.freq = freq; a.phase = phase; a.U = U; a.TC = TC; a.R = R; a.O = out;
  a.B = B; a.Tq = Tq; a.Tk = Tk; a.C = C; a.h = h; a.causal = causal;
  switch (C / h) {
    case 8: return launch_t<8>(a, st);
    case 16: return launch_t<16>(a, st);
    case 32: return launch_t<32>(a, st);
    case 64: return launch_t<64>(a, st);
    default: return set_error(-1, "time_attention: head dim %d unsupported (supported: 8, 16, 32, 64)", C / h);
  }
}

}  // namespace edgl
