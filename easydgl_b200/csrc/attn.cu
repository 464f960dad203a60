// attn.cu - fused self-modulating attention core.
// Replaces the body of T.BiMAU.__call__ (temporal.py:412-447) / T.MAU.__call__ (temporal.py:345-385)
// after the Q/K/V/T projections, including T.MAU.intensity (temporal.py:281-315):
//   S = Q K^T / sqrt(dh); key mask (-2^32+1); [causal mask]; P = softmax(S); H = P T;
//   Z = sigmoid([H, span] W1 + b1); lam_e = s_e * log(1 + exp((Z_e . w_e) / s_e));
//   G[q,k] = sum_e lam[q,e] * marks[k,e]; [G[q,q] = 1]; O = (G o P) V + residual.
// The reference runs ~60 TF ops with [hB,L,L] and two [hB,L,L,E] HBM round trips; here one CTA owns
// one (sequence, head): K/V/T/marks and the intensity MLP weights live in shared memory, each thread
// owns one query row, S/P/G never leave registers.
#include <stdlib.h>

#include "attn_mma.cuh"
#include "common.cuh"

namespace edgl {

constexpr float kMaskFill = -4294967296.0f;  // float(-2**32 + 1), temporal.py:358,425

template <int DH>
__device__ __forceinline__ float dot_dh(const float* q, const float* __restrict__ kr) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < DH; j += 4) {
    const float4 kv = *reinterpret_cast<const float4*>(kr + j);
    s = fmaf(q[j + 0], kv.x, s);
    s = fmaf(q[j + 1], kv.y, s);
    s = fmaf(q[j + 2], kv.z, s);
    s = fmaf(q[j + 3], kv.w, s);
  }
  return s;
}

// Intensity MLP for one row: H[DH], span -> lam[EMAX] (entries >= E untouched).
// W1 [DH+1, DH*E] row-major (row DH multiplies the interval), b1 [DH*E], w [E, DH], sc[E] = exp(scaling).
template <int DH, int EMAX>
__device__ __forceinline__ void intensity_row(const float* H, float span, const float* W1, const float* b1,
                                              const float* w, const float* sc, int E, float* lam) {
  const int NC = DH * E;
#pragma unroll
  for (int e = 0; e < EMAX; ++e) {
    if (e < E) {
      float acc = 0.f;
#pragma unroll 1
      for (int j0 = 0; j0 < DH; j0 += 4) {
        const int c = e * DH + j0;
        const float4 bb = *reinterpret_cast<const float4*>(b1 + c);
        const float4 ws = *reinterpret_cast<const float4*>(W1 + DH * NC + c);
        float z0 = fmaf(span, ws.x, bb.x), z1 = fmaf(span, ws.y, bb.y), z2 = fmaf(span, ws.z, bb.z),
              z3 = fmaf(span, ws.w, bb.w);
#pragma unroll
        for (int i = 0; i < DH; ++i) {
          const float4 wv = *reinterpret_cast<const float4*>(W1 + i * NC + c);
          z0 = fmaf(H[i], wv.x, z0);
          z1 = fmaf(H[i], wv.y, z1);
          z2 = fmaf(H[i], wv.z, z2);
          z3 = fmaf(H[i], wv.w, z3);
        }
        const float4 we = *reinterpret_cast<const float4*>(w + e * DH + j0);
        acc = fmaf(__frcp_rn(1.f + expf(-z0)), we.x, acc);  // tf.nn.sigmoid, temporal.py:290
        acc = fmaf(__frcp_rn(1.f + expf(-z1)), we.y, acc);
        acc = fmaf(__frcp_rn(1.f + expf(-z2)), we.z, acc);
        acc = fmaf(__frcp_rn(1.f + expf(-z3)), we.w, acc);
      }
      const float s = sc[e];
      const float x = __fdiv_rn(acc, s);       // temporal.py:305
      lam[e] = s * logf(1.f + expf(x));        // temporal.py:306 (naive softplus, overflows like TF: Q6)
    }
  }
}

struct AttnSmem {
  float *Ks, *Vs, *Ts, *Ms, *km, *W1, *b1, *w, *sc;
};

template <int DH>
__host__ __device__ inline size_t attn_smem_floats(int L, int E) {
  return (size_t)3 * L * DH + (size_t)L * E + L + (size_t)(DH + 1) * DH * E + (size_t)DH * E + (size_t)E * DH + E;
}

template <int DH, int EMAX>
__global__ void __launch_bounds__(128) attention_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int L = a.L, E = a.E, B = a.B;
  const int b = blockIdx.x / a.h, hh = blockIdx.x % a.h;
  const int tid = threadIdx.x;
  float* Ks = smem;
  float* Vs = Ks + (size_t)L * DH;
  float* Ts = Vs + (size_t)L * DH;
  float* Ms = Ts + (size_t)L * DH;               // [L][E] marks as float
  float* W1 = Ms + (size_t)L * E;                // [(DH+1)][DH*E]
  float* b1 = W1 + (size_t)(DH + 1) * DH * E;
  float* w = b1 + DH * E;
  float* sc = w + E * DH;
  float* km = sc + E;                            // [L] 1 = real key

  const long long row0 = (long long)b * L;
  // ---- stage K/V/T head slices (float4, 64 B contiguous per row at DH=16)
  constexpr int V4 = DH / 4;
  for (int i = tid; i < L * V4; i += 128) {
    const int k = i / V4, j = (i % V4) * 4;
    const long long r = row0 + k;
    *reinterpret_cast<float4*>(Ks + k * DH + j) = *reinterpret_cast<const float4*>(a.K + r * a.ldk + hh * DH + j);
    *reinterpret_cast<float4*>(Vs + k * DH + j) = *reinterpret_cast<const float4*>(a.V + r * a.ldv + hh * DH + j);
    *reinterpret_cast<float4*>(Ts + k * DH + j) = *reinterpret_cast<const float4*>(a.T + r * a.ldt + hh * DH + j);
  }
  for (int i = tid; i < L * E; i += 128) Ms[i] = (float)a.marks[row0 * E + i];  // tf.to_float, temporal.py:311
  for (int i = tid; i < L; i += 128) km[i] = a.kmask[row0 + i] ? 1.f : 0.f;
  for (int i = tid; i < (DH + 1) * DH * E; i += 128) W1[i] = a.int_w[i];
  for (int i = tid; i < DH * E; i += 128) {
    b1[i] = a.int_b[i];
    w[i] = a.int_weight[i];
  }
  for (int i = tid; i < E; i += 128) sc[i] = expf(a.int_scaling[i]);  // temporal.py:302
  __syncthreads();

  const float sqrt_dh = sqrtf((float)DH);  // K_.get_shape()[-1] ** 0.5, temporal.py:355,422
  for (int q0 = 0; q0 < L; q0 += 128) {
    const int q = q0 + tid;
    if (q >= L) continue;
    const long long row = row0 + q;
    float qv[DH];
#pragma unroll
    for (int j = 0; j < DH; j += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(a.Q + row * a.ldq + hh * DH + j);
      qv[j] = t4.x; qv[j + 1] = t4.y; qv[j + 2] = t4.z; qv[j + 3] = t4.w;
    }
    // pass A: row max of the masked scores
    float m = -INFINITY;
    for (int k = 0; k < L; ++k) {
      float s = __fdiv_rn(dot_dh<DH>(qv, Ks + k * DH), sqrt_dh);
      if (km[k] == 0.f || (a.causal && k > q)) s = kMaskFill;
      m = fmaxf(m, s);
    }
    // pass B: softmax denominator and H = P T
    float l = 0.f;
    float Hq[DH];
#pragma unroll
    for (int j = 0; j < DH; ++j) Hq[j] = 0.f;
    for (int k = 0; k < L; ++k) {
      float s = __fdiv_rn(dot_dh<DH>(qv, Ks + k * DH), sqrt_dh);
      if (km[k] == 0.f || (a.causal && k > q)) s = kMaskFill;
      const float p = expf(s - m);
      l += p;
#pragma unroll
      for (int j = 0; j < DH; j += 4) {
        const float4 tv = *reinterpret_cast<const float4*>(Ts + k * DH + j);
        Hq[j] = fmaf(p, tv.x, Hq[j]);
        Hq[j + 1] = fmaf(p, tv.y, Hq[j + 1]);
        Hq[j + 2] = fmaf(p, tv.z, Hq[j + 2]);
        Hq[j + 3] = fmaf(p, tv.w, Hq[j + 3]);
      }
    }
    const float inv_l = __frcp_rn(l);
#pragma unroll
    for (int j = 0; j < DH; ++j) Hq[j] *= inv_l;
    // intensity MLP
    float lam[EMAX];
#pragma unroll
    for (int e = 0; e < EMAX; ++e) lam[e] = 0.f;
    intensity_row<DH, EMAX>(Hq, a.spans[row], W1, b1, w, sc, E, lam);
    if (a.lam) {
      float* lp = a.lam + (((long long)hh * B + b) * L + q) * E;  // head-major, temporal.py:413-416
#pragma unroll
      for (int e = 0; e < EMAX; ++e)
        if (e < E) lp[e] = lam[e];
    }
    // pass C: O = (G o P) V
    float Oq[DH];
#pragma unroll
    for (int j = 0; j < DH; ++j) Oq[j] = 0.f;
    for (int k = 0; k < L; ++k) {
      float s = __fdiv_rn(dot_dh<DH>(qv, Ks + k * DH), sqrt_dh);
      if (km[k] == 0.f || (a.causal && k > q)) s = kMaskFill;
      const float p = expf(s - m) * inv_l;
      float g = 0.f;
      const float* mk = Ms + k * E;
#pragma unroll
      for (int e = 0; e < EMAX; ++e)
        if (e < E) g = fmaf(lam[e], mk[e], g);
      if (a.diag_one && k == q) g = 1.f;  // tf.linalg.set_diag, temporal.py:438-439
      const float gp = g * p;             // temporal.py:441
#pragma unroll
      for (int j = 0; j < DH; j += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(Vs + k * DH + j);
        Oq[j] = fmaf(gp, vv.x, Oq[j]);
        Oq[j + 1] = fmaf(gp, vv.y, Oq[j + 1]);
        Oq[j + 2] = fmaf(gp, vv.z, Oq[j + 2]);
        Oq[j + 3] = fmaf(gp, vv.w, Oq[j + 3]);
      }
    }
    float* op = a.O + row * a.ldo + hh * DH;
    const float* rp = a.R ? a.R + row * a.ldr + hh * DH : nullptr;
#pragma unroll
    for (int j = 0; j < DH; j += 4) {
      float4 o = make_float4(Oq[j], Oq[j + 1], Oq[j + 2], Oq[j + 3]);
      if (rp) {  // residual, temporal.py:385,447
        const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
        o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
      }
      *reinterpret_cast<float4*>(op + j) = o;
    }
  }
}

template <int DH, int EMAX>
static int launch_attention_stream_t(const AttnArgs& a, cudaStream_t st);

template <int DH, int EMAX>
static int launch_attention_t(const AttnArgs& a, cudaStream_t st) {
  const size_t smem = attn_smem_floats<DH>(a.L, a.E) * sizeof(float);
  if (smem > 227 * 1024) return launch_attention_stream_t<DH, EMAX>(a, st);  // long sequences: stream the keys
  auto kern = attention_kernel<DH, EMAX>;
  EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)(a.B * a.h), 128, smem, st>>>(a);
  EDGL_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Key-streaming CUDA-core variant for long sequences (any L; e.g. BASELINE config C5: L=512, dh=32): the
// same three passes as attention_kernel, but K / T / V / marks are streamed through shared memory in chunks
// of KT keys, so shared memory no longer grows with L.  Correctness fallback, not a tuned kernel.
constexpr int KT = 64;

template <int DH, int EMAX>
__global__ void __launch_bounds__(128) attention_stream_kernel(AttnArgs a, int w1_smem) {
  extern __shared__ __align__(16) float smem[];
  const int L = a.L, E = a.E, B = a.B;
  const int b = blockIdx.x / a.h, hh = blockIdx.x % a.h;
  const int tid = threadIdx.x;
  float* Kc = smem;                    // [KT][DH]
  float* Xc = Kc + KT * DH;            // [KT][DH]  T (pass B) or V (pass C)
  float* Mc = Xc + KT * DH;            // [KT][E]
  float* kmc = Mc + KT * E;            // [KT]
  float* b1 = kmc + KT;                // [DH*E]
  float* w = b1 + DH * E;              // [E*DH]
  float* sc = w + E * DH;              // [E]
  float* W1s = sc + ((E + 3) & ~3);    // [(DH+1)][DH*E] when it fits
  const float* W1 = w1_smem ? W1s : a.int_w;
  const long long row0 = (long long)b * L;
  if (w1_smem)
    for (int i = tid; i < (DH + 1) * DH * E; i += 128) W1s[i] = a.int_w[i];
  for (int i = tid; i < DH * E; i += 128) {
    b1[i] = a.int_b[i];
    w[i] = a.int_weight[i];
  }
  for (int i = tid; i < E; i += 128) sc[i] = expf(a.int_scaling[i]);
  __syncthreads();
  constexpr int V4 = DH / 4;
  auto load_chunk = [&](int k0, const float* X, int ldx, bool with_marks) {
    __syncthreads();  // previous chunk fully consumed
    for (int i = tid; i < KT * V4; i += 128) {
      const int k = i / V4, j = (i % V4) * 4;
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), xx = kk;
      if (k0 + k < L) {
        const long long r = row0 + k0 + k;
        kk = *reinterpret_cast<const float4*>(a.K + r * a.ldk + hh * DH + j);
        if (X) xx = *reinterpret_cast<const float4*>(X + r * ldx + hh * DH + j);
      }
      *reinterpret_cast<float4*>(Kc + k * DH + j) = kk;
      *reinterpret_cast<float4*>(Xc + k * DH + j) = xx;
    }
    if (with_marks)
      for (int i = tid; i < KT * E; i += 128) {
        const int k = i / E;
        Mc[i] = (k0 + k < L) ? (float)a.marks[(row0 + k0 + k) * E + (i % E)] : 0.f;
      }
    for (int i = tid; i < KT; i += 128) kmc[i] = (k0 + i < L) ? (a.kmask[row0 + k0 + i] ? 1.f : 0.f) : -1.f;
    __syncthreads();
  };
  const float sqrt_dh = sqrtf((float)DH);
  for (int q0 = 0; q0 < L; q0 += 128) {
    const int q = q0 + tid;
    const bool act = q < L;
    const long long row = row0 + (act ? q : L - 1);
    float qv[DH];
#pragma unroll
    for (int j = 0; j < DH; j += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(a.Q + row * a.ldq + hh * DH + j);
      qv[j] = t4.x; qv[j + 1] = t4.y; qv[j + 2] = t4.z; qv[j + 3] = t4.w;
    }
    auto score = [&](int k0, int k) {
      float s = __fdiv_rn(dot_dh<DH>(qv, Kc + k * DH), sqrt_dh);
      if (kmc[k] == 0.f || (a.causal && k0 + k > q)) s = kMaskFill;
      return s;
    };
    float m = -INFINITY;
    for (int k0 = 0; k0 < L; k0 += KT) {  // pass A: row max
      load_chunk(k0, nullptr, 0, false);
      const int kn = min(KT, L - k0);
      for (int k = 0; k < kn; ++k) m = fmaxf(m, score(k0, k));
    }
    float l = 0.f, Hq[DH];
#pragma unroll
    for (int j = 0; j < DH; ++j) Hq[j] = 0.f;
    for (int k0 = 0; k0 < L; k0 += KT) {  // pass B: denominator and H = P T
      load_chunk(k0, a.T, a.ldt, false);
      const int kn = min(KT, L - k0);
      for (int k = 0; k < kn; ++k) {
        const float p = expf(score(k0, k) - m);
        l += p;
#pragma unroll
        for (int j = 0; j < DH; ++j) Hq[j] = fmaf(p, Xc[k * DH + j], Hq[j]);
      }
    }
    const float inv_l = __frcp_rn(l);
#pragma unroll
    for (int j = 0; j < DH; ++j) Hq[j] *= inv_l;
    float lam[EMAX];
#pragma unroll
    for (int e = 0; e < EMAX; ++e) lam[e] = 0.f;
    intensity_row<DH, EMAX>(Hq, a.spans[row], W1, b1, w, sc, E, lam);
    if (a.lam && act) {
      float* lp = a.lam + (((long long)hh * B + b) * L + q) * E;
#pragma unroll
      for (int e = 0; e < EMAX; ++e)
        if (e < E) lp[e] = lam[e];
    }
    float Oq[DH];
#pragma unroll
    for (int j = 0; j < DH; ++j) Oq[j] = 0.f;
    for (int k0 = 0; k0 < L; k0 += KT) {  // pass C: O = (G o P) V
      load_chunk(k0, a.V, a.ldv, true);
      const int kn = min(KT, L - k0);
      for (int k = 0; k < kn; ++k) {
        const float p = expf(score(k0, k) - m) * inv_l;
        float g = 0.f;
#pragma unroll
        for (int e = 0; e < EMAX; ++e)
          if (e < E) g = fmaf(lam[e], Mc[k * E + e], g);
        if (a.diag_one && k0 + k == q) g = 1.f;
        const float gp = g * p;
#pragma unroll
        for (int j = 0; j < DH; ++j) Oq[j] = fmaf(gp, Xc[k * DH + j], Oq[j]);
      }
    }
    if (act) {
      float* op = a.O + row * a.ldo + hh * DH;
      const float* rp = a.R ? a.R + row * a.ldr + hh * DH : nullptr;
#pragma unroll
      for (int j = 0; j < DH; j += 4) {
        float4 o = make_float4(Oq[j], Oq[j + 1], Oq[j + 2], Oq[j + 3]);
        if (rp) {
          const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
          o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
        }
        *reinterpret_cast<float4*>(op + j) = o;
      }
    }
  }
}

template <int DH, int EMAX>
static int launch_attention_stream_t(const AttnArgs& a, cudaStream_t st) {
  const size_t fixed = (size_t)2 * KT * DH + (size_t)KT * a.E + KT + (size_t)2 * DH * a.E + ((a.E + 3) & ~3);
  const size_t w1 = (size_t)(DH + 1) * DH * a.E;
  const int w1_smem = (fixed + w1) * sizeof(float) <= 200 * 1024;
  const size_t smem = (fixed + (w1_smem ? w1 : 0)) * sizeof(float);
  auto kern = attention_stream_kernel<DH, EMAX>;
  EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)(a.B * a.h), 128, smem, st>>>(a, w1_smem);
  EDGL_LAUNCH_CHECK();
  return 0;
}

int launch_attention(const AttnArgs& a, cudaStream_t st) {
  static const char mode = [] {
    const char* e = getenv("EDGL_ATTN");
    if (!e) return 'd';
    if (e[0] == 't' && e[1] == 'c' && e[2] == '2') return '2';
    return e[0];
  }();
  return launch_attention_mode(a, st, mode);
}

int launch_attention_mode(const AttnArgs& a, cudaStream_t st, char mode) {
  EDGL_REQUIRE(a.d % a.h == 0, "num_units %d not divisible by num_heads %d", a.d, a.h);
  const int dh = a.d / a.h;
  EDGL_REQUIRE(a.E >= 1 && a.E <= 32, "num_events must be in [1,32] (got %d)", a.E);
  EDGL_REQUIRE(a.E % 4 == 0, "num_events must be a multiple of 4 (got %d)", a.E);
  EDGL_REQUIRE((a.ldq % 4 == 0) && (a.ldk % 4 == 0) && (a.ldv % 4 == 0) && (a.ldt % 4 == 0) &&
                   (a.ldo % 4 == 0) && (!a.R || a.ldr % 4 == 0),
               "attention: leading dimensions must be multiples of 4");
  if (a.amax_published) *a.amax_published = false;
  if (a.B == 0) return 0;
  // tensor-core path (attn_mma.cuh) for the shapes it is instantiated for; EDGL_ATTN=simt forces the
  // CUDA-core kernel below (same results to fp32 rounding; used by the parity tests to cover both)
  const bool force_simt = mode == 's';
  // EDGL_ATTN=tc2: the tcgen05 / TMA kernel with two items in flight (attn_tc2.cu; dh = 16, E = 16, L <= 112).  Parity
  // green at fp32 level, but its serial per-item chain (five MMA round trips, 8 warps per item) runs at 1.61 ms at C2
  // against 1.35 ms for the mma.sync kernel below, so it is opt-in (DESIGN.md section 4 has the per-phase cycle table)
  if (mode == '2') {
    const int r = launch_attention_tc2(a, st);
    if (r == 0 && a.amax_published) *a.amax_published = a.out_amax != nullptr;
    if (r <= 0) return r;
  }
  // EDGL_ATTN=tc: the tcgen05 / TMEM kernel (attn_tc.cu; dh = 16, E = 16, L <= 128).  It passes the same parity
  // tests but, with one thread per query row, is latency-bound (2.4 ms at C2 against 1.35 ms for the default
  // attn_f16.cu); docs/attn_tc2_design.md describes the column-parallel successor.
  if (mode == 't') {
    const int r = launch_attention_tc(a, st);
    if (r <= 0) return r;
  }
  // next (and EDGL_ATTN=f16): the scaled 3xFP16 mma.sync kernel (attn_f16.cu; dh = 16, E = 16, L <= 208).
  // EDGL_ATTN=mma selects the 3xTF32 mma.sync kernel for every shape, as before.
  if (mode == 'f' || mode == 'd' || mode == '2') {
    const int r = launch_attention_f16(a, st);
    if (r == 0 && a.amax_published) *a.amax_published = a.out_amax != nullptr;
    if (r <= 0) return r;
  }
  if (!force_simt) {  // EDGL_ATTN=mma: the mma.sync kernel; EDGL_ATTN=simt: the CUDA-core kernel
    int r = 1;
    if (dh == 8) r = launch_attention_mma_dh8(a, st);
    else if (dh == 16) r = launch_attention_mma_dh16(a, st);
    else if (dh == 32) r = launch_attention_mma_dh32(a, st);
    if (r <= 0) return r;
  }
#define EDGL_ATT(DHV)                                                     \
  if (dh == DHV) {                                                        \
    if (a.E <= 16) return launch_attention_t<DHV, 16>(a, st);             \
    return launch_attention_t<DHV, 32>(a, st);                            \
  }
  EDGL_ATT(8)
  EDGL_ATT(16)
  EDGL_ATT(32)
  EDGL_ATT(64)
#undef EDGL_ATT
  return set_error(-1, "attention: head dim %d unsupported (supported: 8, 16, 32, 64)", dh);
}

// Stand-alone T.MAU.intensity (temporal.py:281-315): H [hB,L,dh] -> G [hB,L,L], lam [hB,L,E].
template <int DH, int EMAX>
__global__ void __launch_bounds__(128) intensity_kernel(const float* __restrict__ H, const float* __restrict__ spans,
                                                        const uint8_t* __restrict__ marks,
                                                        const float* __restrict__ int_w,
                                                        const float* __restrict__ int_b,
                                                        const float* __restrict__ int_weight,
                                                        const float* __restrict__ int_scaling, int B, int L,
                                                        int h, int E, float* __restrict__ G,
                                                        float* __restrict__ lam_out) {
  extern __shared__ __align__(16) float smem[];
  float* Ms = smem;                  // [L][E]
  float* sc = Ms + (size_t)L * E;    // [E]
  const int hb = blockIdx.x, b = hb % B;  // head-major: hb = head*B + b
  for (int i = threadIdx.x; i < L * E; i += 128) Ms[i] = (float)marks[(long long)b * L * E + i];
  for (int i = threadIdx.x; i < E; i += 128) sc[i] = expf(int_scaling[i]);
  __syncthreads();
  for (int q = threadIdx.x; q < L; q += 128) {
    float Hq[DH];
    const float* hp = H + ((long long)hb * L + q) * DH;
#pragma unroll
    for (int j = 0; j < DH; ++j) Hq[j] = hp[j];
    float lam[EMAX];
#pragma unroll
    for (int e = 0; e < EMAX; ++e) lam[e] = 0.f;
    intensity_row<DH, EMAX>(Hq, spans[(long long)b * L + q], int_w, int_b, int_weight, sc, E, lam);
    if (lam_out) {
#pragma unroll
      for (int e = 0; e < EMAX; ++e)
        if (e < E) lam_out[((long long)hb * L + q) * E + e] = lam[e];
    }
    if (G) {
      for (int k = 0; k < L; ++k) {
        float g = 0.f;
#pragma unroll
        for (int e = 0; e < EMAX; ++e)
          if (e < E) g = fmaf(lam[e], Ms[k * E + e], g);
        G[((long long)hb * L + q) * L + k] = g;
      }
    }
  }
}

int launch_intensity(const float* H, const float* spans, const uint8_t* marks, const float* int_w,
                     const float* int_b, const float* int_weight, const float* int_scaling, int B, int L, int h,
                     int dh, int E, float* G, float* lam, cudaStream_t st) {
  EDGL_REQUIRE(E >= 1 && E <= 32 && E % 4 == 0, "num_events must be a multiple of 4 in [4,32] (got %d)", E);
  if (B == 0) return 0;
  const size_t smem = ((size_t)L * E + E) * sizeof(float);
  EDGL_REQUIRE(smem <= 227 * 1024, "intensity: L*E too large for shared memory");
#define EDGL_INT(DHV, EM)                                                                                  \
  {                                                                                                        \
    auto kern = intensity_kernel<DHV, EM>;                                                                 \
    EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    kern<<<(unsigned)(B * h), 128, smem, st>>>(H, spans, marks, int_w, int_b, int_weight, int_scaling, B, L, \
                                               h, E, G, lam);                                              \
    EDGL_LAUNCH_CHECK();                                                                                   \
    return 0;                                                                                              \
  }
  if (dh == 8) { if (E <= 16) EDGL_INT(8, 16) else EDGL_INT(8, 32) }
  if (dh == 16) { if (E <= 16) EDGL_INT(16, 16) else EDGL_INT(16, 32) }
  if (dh == 32) { if (E <= 16) EDGL_INT(32, 16) else EDGL_INT(32, 32) }
  if (dh == 64) { if (E <= 16) EDGL_INT(64, 16) else EDGL_INT(64, 32) }
#undef EDGL_INT
  return set_error(-1, "intensity: head dim %d unsupported (supported: 8, 16, 32, 64)", dh);
}

}  // namespace edgl
