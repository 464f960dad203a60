// attn_mma.cuh - tensor-core version of the fused self-modulating attention core.
// Same math as attn.cu (temporal.py:345-385 / 412-447 / 281-315); every contraction
//   S = Q K^T, H = P T, Z = [H,span] W1, G = lam M^T, O = (G o P) V
// runs on mma.sync.m16n8k8 TF32 with the 3xTF32 split (x = hi + lo, hi = the 19 bits the tensor core
// reads, lo = x - hi exact in fp32; D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi) so results stay at fp32
// accuracy (parity bar: top-K sets identical to an fp32 reference).  Marks are small integers, exact
// in TF32, so G needs two MMAs instead of three.
//
// One CTA per (sequence, head); each warp owns 16 query rows; S/P/G live in registers in the MMA
// accumulator layout.  The key index of every P-as-A-operand contraction is permuted
// (slot t <-> key 2t, slot t+4 <-> key 2t+1) so the accumulator fragment of one MMA is directly the
// A fragment of the next - no shuffles, no shared-memory round trip.
#pragma once
#include "common.cuh"

namespace edgl {

constexpr float kFillMma = -4294967296.0f;  // float(-2**32+1), temporal.py:358,425
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// lo part of the 3xTF32 split: x - (x with the 13 low mantissa bits cleared); exact in fp32
__device__ __forceinline__ uint32_t tf32_lo(float x) {
  return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u));
}

// c += A * B with A given as fp32 values (split here) and B as two fp32 values (split here)
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], float b0,
                                     float b1) {
  const uint32_t bh0 = __float_as_uint(b0), bh1 = __float_as_uint(b1);
  mma_tf32(c, al, bh0, bh1);
  mma_tf32(c, ah, tf32_lo(b0), tf32_lo(b1));
  mma_tf32(c, ah, bh0, bh1);
}

__device__ __forceinline__ void split4(const float (&x)[4], uint32_t (&h)[4], uint32_t (&l)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __float_as_uint(x[i]);
    l[i] = tf32_lo(x[i]);
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DH, int E>
struct AttnMmaLayout {
  static constexpr int SK = DH + 4;       // K/V/T row stride (floats): conflict-free fragment loads
  static constexpr int SE = E + 4;        // marks row stride
  static constexpr int NC = DH * E;       // intensity MLP width
  static constexpr int SW = NC + 4;       // W1 row stride (== 4 mod 32)
  __host__ __device__ static size_t floats(int LP) {
    return (size_t)3 * LP * SK + (size_t)LP * SE + (size_t)DH * SW + 3 * NC + E + LP;
  }
};

// NT = number of 8-key tiles held in registers (L <= 8*NT)
template <int DH, int E, int NT>
__global__ void __launch_bounds__(256, (NT <= 16) ? 2 : 1) attention_mma_kernel(AttnArgs a) {
  using LY = AttnMmaLayout<DH, E>;
  constexpr int SK = LY::SK, SE = LY::SE, NC = LY::NC, SW = LY::SW;
  constexpr int KS = DH / 8;   // k-steps over the head dim / n-tiles of a [.,DH] output
  constexpr int ES = E / 8;    // k-steps over events
  constexpr int LP = NT * 8;   // padded key count
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;
  float* Vs = Ks + LP * SK;
  float* Ts = Vs + LP * SK;
  float* Ms = Ts + LP * SK;          // [LP][SE] marks as float (tf.to_float, temporal.py:311)
  float* W1 = Ms + LP * SE;          // [DH][SW]  rows 0..DH-1 of int_w
  float* wsp = W1 + DH * SW;         // [NC] row DH of int_w (multiplies the interval)
  float* b1 = wsp + NC;              // [NC]
  float* wv = b1 + NC;               // [NC] int_weight flattened [E][DH]
  float* sc = wv + NC;               // [E] exp(scaling)
  float* km = sc + E;                // [LP] 1 = real key, 0 = padding id, -1 = beyond L

  const int L = a.L, B = a.B;
  const int b = blockIdx.x / a.h, hh = blockIdx.x % a.h;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const long long row0 = (long long)b * L;

  // ---------------------------------------------------------------- stage operands in shared memory
  constexpr int V4 = DH / 4;
  for (int i = tid; i < LP * V4; i += nthr) {
    const int k = i / V4, j = (i % V4) * 4;
    float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk, tt = kk;
    if (k < L) {
      const long long r = row0 + k;
      kk = *reinterpret_cast<const float4*>(a.K + r * a.ldk + hh * DH + j);
      vv = *reinterpret_cast<const float4*>(a.V + r * a.ldv + hh * DH + j);
      tt = *reinterpret_cast<const float4*>(a.T + r * a.ldt + hh * DH + j);
    }
    *reinterpret_cast<float4*>(Ks + k * SK + j) = kk;
    *reinterpret_cast<float4*>(Vs + k * SK + j) = vv;
    *reinterpret_cast<float4*>(Ts + k * SK + j) = tt;
  }
  for (int i = tid; i < LP * E; i += nthr) {
    const int k = i / E, e = i % E;
    Ms[k * SE + e] = (k < L) ? (float)a.marks[(row0 + k) * E + e] : 0.f;
  }
  for (int i = tid; i < LP; i += nthr) km[i] = (i < L) ? (a.kmask[row0 + i] ? 1.f : 0.f) : -1.f;
  for (int i = tid; i < DH * NC; i += nthr) W1[(i / NC) * SW + (i % NC)] = a.int_w[i];
  for (int i = tid; i < NC; i += nthr) {
    wsp[i] = a.int_w[DH * NC + i];
    b1[i] = a.int_b[i];
    wv[i] = a.int_weight[i];
  }
  for (int i = tid; i < E; i += nthr) sc[i] = expf(a.int_scaling[i]);  // temporal.py:302
  __syncthreads();

  const float inv_sqrt_dh = 1.0f / sqrtf((float)DH);  // temporal.py:355,422
  const int num_mt = (L + 15) >> 4;
  for (int mt = warp; mt < num_mt; mt += (nthr >> 5)) {
    const int q0 = mt * 16;
    const int qa = q0 + g, qb = q0 + g + 8;                      // this thread's two query rows
    const long long ra = row0 + (qa < L ? qa : L - 1), rb = row0 + (qb < L ? qb : L - 1);

    // ---- Q fragments (A operand of S = Q K^T)
    uint32_t qh[KS][4], ql[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      float x[4];
      x[0] = a.Q[ra * a.ldq + hh * DH + ks * 8 + t];
      x[1] = a.Q[rb * a.ldq + hh * DH + ks * 8 + t];
      x[2] = a.Q[ra * a.ldq + hh * DH + ks * 8 + t + 4];
      x[3] = a.Q[rb * a.ldq + hh * DH + ks * 8 + t + 4];
      split4(x, qh[ks], ql[ks]);
    }
    // ---- S = Q K^T  (accumulators P[nt][c]: rows g / g+8, keys nt*8 + 2t + (c&1))
    float P[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      P[nt][0] = P[nt][1] = P[nt][2] = P[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const float* kp = Ks + (nt * 8 + g) * SK + ks * 8 + t;
        mma3(P[nt], qh[ks], ql[ks], kp[0], kp[4]);
      }
    }
    // ---- scale, key mask, causal mask, softmax (rows live in a quad: 2 shuffles per reduction)
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = nt * 8 + 2 * t + (c & 1);
        const int qr = (c < 2) ? qa : qb;
        const float kmv = km[col];
        float s = P[nt][c] * inv_sqrt_dh;
        if (kmv == 0.f || (a.causal && col > qr)) s = kFillMma;
        if (kmv < 0.f) s = -INFINITY;  // beyond L: not a key at all
        P[nt][c] = s;
        if (c < 2) ma = fmaxf(ma, s); else mb = fmaxf(mb, s);
      }
    }
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    float la = 0.f, lb = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float m = (c < 2) ? ma : mb;
        const float p = ex2_approx((P[nt][c] - m) * kLog2e);
        P[nt][c] = p;
        if (c < 2) la += p; else lb += p;
      }
    }
    la += __shfl_xor_sync(0xffffffffu, la, 1);
    la += __shfl_xor_sync(0xffffffffu, la, 2);
    lb += __shfl_xor_sync(0xffffffffu, lb, 1);
    lb += __shfl_xor_sync(0xffffffffu, lb, 2);
    const float ia = __frcp_rn(la), ib = __frcp_rn(lb);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      P[nt][0] *= ia; P[nt][1] *= ia; P[nt][2] *= ib; P[nt][3] *= ib;
    }
    // ---- H = P T   (k = keys, permuted: slot t <-> key 2t, slot t+4 <-> key 2t+1)
    float H[KS][4];
#pragma unroll
    for (int n = 0; n < KS; ++n) H[n][0] = H[n][1] = H[n][2] = H[n][3] = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float pa[4] = {P[nt][0], P[nt][2], P[nt][1], P[nt][3]};
      uint32_t ph[4], pl[4];
      split4(pa, ph, pl);
#pragma unroll
      for (int n = 0; n < KS; ++n) {
        const float* tp = Ts + (nt * 8 + 2 * t) * SK + n * 8 + g;
        mma3(H[n], ph, pl, tp[0], tp[SK]);
      }
    }
    // ---- intensity MLP: Z = sigmoid([H, span] W1 + b1); dot with w per event (temporal.py:287-305)
    const float spa = a.spans[ra], spb = a.spans[rb];
    uint32_t hh_[KS][4], hl_[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const float ha[4] = {H[ks][0], H[ks][2], H[ks][1], H[ks][3]};
      split4(ha, hh_[ks], hl_[ks]);
    }
    float lsa[E], lsb[E];  // per-event dot products for rows qa / qb (full sums after the quad reduce)
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float pa = 0.f, pb = 0.f;
#pragma unroll
      for (int jj = 0; jj < KS; ++jj) {
        const int n0 = (e * KS + jj) * 8;
        float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const float* wp = W1 + (ks * 8 + 2 * t) * SW + n0 + g;
          mma3(z, hh_[ks], hl_[ks], wp[0], wp[SW]);
        }
        const int c0 = n0 + 2 * t;
        const float2 bb = *reinterpret_cast<const float2*>(b1 + c0);
        const float2 ws = *reinterpret_cast<const float2*>(wsp + c0);
        const float2 we = *reinterpret_cast<const float2*>(wv + c0);
        const float z0 = z[0] + fmaf(spa, ws.x, bb.x), z1 = z[1] + fmaf(spa, ws.y, bb.y);
        const float z2 = z[2] + fmaf(spb, ws.x, bb.x), z3 = z[3] + fmaf(spb, ws.y, bb.y);
        // sigmoid = 1 / (1 + 2^(-z log2 e))
        pa = fmaf(__frcp_rn(1.f + ex2_approx(-z0 * kLog2e)), we.x, pa);
        pa = fmaf(__frcp_rn(1.f + ex2_approx(-z1 * kLog2e)), we.y, pa);
        pb = fmaf(__frcp_rn(1.f + ex2_approx(-z2 * kLog2e)), we.x, pb);
        pb = fmaf(__frcp_rn(1.f + ex2_approx(-z3 * kLog2e)), we.y, pb);
      }
      pa += __shfl_xor_sync(0xffffffffu, pa, 1);
      pa += __shfl_xor_sync(0xffffffffu, pa, 2);
      pb += __shfl_xor_sync(0xffffffffu, pb, 1);
      pb += __shfl_xor_sync(0xffffffffu, pb, 2);
      lsa[e] = pa;
      lsb[e] = pb;
    }
    // ---- lam_e = s_e log(1 + exp(x / s_e))  (temporal.py:305-306); lane t owns events t, t+4, t+8, ...
    uint32_t lh[ES][4], ll[ES][4];
#pragma unroll
    for (int ks = 0; ks < ES; ++ks) {
      float lam4[4];
#pragma unroll
      for (int hsel = 0; hsel < 2; ++hsel) {
        const int ebase = ks * 8 + hsel * 4;  // events ebase + t
        float xa = lsa[ebase], xb = lsb[ebase];
        if (t == 1) { xa = lsa[ebase + 1]; xb = lsb[ebase + 1]; }
        if (t == 2) { xa = lsa[ebase + 2]; xb = lsb[ebase + 2]; }
        if (t == 3) { xa = lsa[ebase + 3]; xb = lsb[ebase + 3]; }
        const float s = sc[ebase + t];
        const float va = s * logf(1.f + expf(__fdiv_rn(xa, s)));
        const float vb = s * logf(1.f + expf(__fdiv_rn(xb, s)));
        lam4[hsel * 2 + 0] = va;  // a0 / a2 : row qa
        lam4[hsel * 2 + 1] = vb;  // a1 / a3 : row qb
        if (a.lam) {
          if (qa < L) a.lam[(((long long)hh * B + b) * L + qa) * E + ebase + t] = va;  // head-major, temporal.py:413
          if (qb < L) a.lam[(((long long)hh * B + b) * L + qb) * E + ebase + t] = vb;
        }
      }
      split4(lam4, lh[ks], ll[ks]);
    }
    // ---- G = lam M^T (marks exact in TF32: 2 MMAs), set_diag, gate: P <- G o P   (temporal.py:309-313,438-441)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float G[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < ES; ++ks) {
        const float* mp = Ms + (nt * 8 + g) * SE + ks * 8 + t;
        const uint32_t m0 = __float_as_uint(mp[0]), m1 = __float_as_uint(mp[4]);
        mma_tf32(G, ll[ks], m0, m1);
        mma_tf32(G, lh[ks], m0, m1);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = nt * 8 + 2 * t + (c & 1);
        const int qr = (c < 2) ? qa : qb;
        const float gg = (a.diag_one && col == qr) ? 1.f : G[c];
        P[nt][c] *= gg;
      }
    }
    // ---- O = (G o P) V
    float O[KS][4];
#pragma unroll
    for (int n = 0; n < KS; ++n) O[n][0] = O[n][1] = O[n][2] = O[n][3] = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float pa[4] = {P[nt][0], P[nt][2], P[nt][1], P[nt][3]};
      uint32_t ph[4], pl[4];
      split4(pa, ph, pl);
#pragma unroll
      for (int n = 0; n < KS; ++n) {
        const float* vp = Vs + (nt * 8 + 2 * t) * SK + n * 8 + g;
        mma3(O[n], ph, pl, vp[0], vp[SK]);
      }
    }
    // ---- residual + store (temporal.py:385,447)
#pragma unroll
    for (int n = 0; n < KS; ++n) {
      const int col = hh * DH + n * 8 + 2 * t;
      if (qa < L) {
        float2 o = make_float2(O[n][0], O[n][1]);
        if (a.R) {
          const float2 r = *reinterpret_cast<const float2*>(a.R + (row0 + qa) * a.ldr + col);
          o.x += r.x; o.y += r.y;
        }
        *reinterpret_cast<float2*>(a.O + (row0 + qa) * a.ldo + col) = o;
      }
      if (qb < L) {
        float2 o = make_float2(O[n][2], O[n][3]);
        if (a.R) {
          const float2 r = *reinterpret_cast<const float2*>(a.R + (row0 + qb) * a.ldr + col);
          o.x += r.x; o.y += r.y;
        }
        *reinterpret_cast<float2*>(a.O + (row0 + qb) * a.ldo + col) = o;
      }
    }
  }
}

template <int DH, int E, int NT>
int launch_attention_mma_t(const AttnArgs& a, cudaStream_t st) {
  using LY = AttnMmaLayout<DH, E>;
  const size_t smem = LY::floats(NT * 8) * sizeof(float);
  if (smem > 227 * 1024) return 1;  // caller falls back
  auto kern = attention_mma_kernel<DH, E, NT>;
  EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int warps = (a.L + 15) / 16;
  if (warps > 8) warps = 8;
  kern<<<(unsigned)(a.B * a.h), warps * 32, smem, st>>>(a);
  EDGL_LAUNCH_CHECK();
  return 0;
}

// returns 0 = launched, 1 = shape not covered by this instantiation set, <0 = error
template <int DH>
int launch_attention_mma_dh(const AttnArgs& a, cudaStream_t st) {
#define EDGL_NT(EV)                                                          \
  if (a.L <= 32) return launch_attention_mma_t<DH, EV, 4>(a, st);            \
  if (a.L <= 104) return launch_attention_mma_t<DH, EV, 13>(a, st);          \
  if (a.L <= 128) return launch_attention_mma_t<DH, EV, 16>(a, st);          \
  if (a.L <= 208) return launch_attention_mma_t<DH, EV, 26>(a, st);          \
  return 1;
  if (a.E == 16) { EDGL_NT(16) }
  if (a.E == 8) { EDGL_NT(8) }
#undef EDGL_NT
  return 1;
}

int launch_attention_mma_dh8(const AttnArgs& a, cudaStream_t st);
int launch_attention_mma_dh16(const AttnArgs& a, cudaStream_t st);
int launch_attention_mma_dh32(const AttnArgs& a, cudaStream_t st);

}  // namespace edgl
