// attn_f16_common.cuh - helpers shared by the scaled 3xFP16 mma.sync attention kernels (attn_f16.cu: keys held in
// registers, L <= 208; attn_f16_long.cu: keys streamed in chunks, any L): the m16n8k16 MMA, the (hi, lo) fp16 split,
// packed fp32x2 arithmetic, power-of-two scales, the constant pack of the intensity MLP, ldmatrix / cp.async wrappers
// and the P-as-A-operand product.
#pragma once
#include <cuda_fp16.h>

#include "attn_mma.cuh"

namespace edgl {
namespace f16c {


__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_h2(float e0, float e1) {  // e0 -> low half (lower k index)
  const __half2 h = __floats2half2_rn(e0, e1);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2 - one issue slot for two IEEE operations on an aligned
// register pair; bit-identical to the scalar instructions).  The kernel is bound by issue slots, so every
// elementwise fp32 step on two neighbouring accumulator columns is written this way.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// two (already scaled) fp32 values -> packed hi pair, packed lo pair
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const float h0 = __uint_as_float(__float_as_uint(x0) & 0xffffe000u);
  const float h1 = __uint_as_float(__float_as_uint(x1) & 0xffffe000u);
  hi = pack_h2(h0, h1);
  float l0, l1;
  upk2(sub2(pk2(x0, x1), pk2(h0, h1)), l0, l1);
  lo = pack_h2(l0, l1);
}

// m >= 0: s = 2^k with m*s in [2^14, 2^15), is = 1/s (both exact; exponent clamped so neither is denormal)
__device__ __forceinline__ void pow2_scale(float m, float& s, float& is) {
  int e = (int)((__float_as_uint(m) >> 23) & 0xffu);
  e = min(max(e, 15), 239);
  s = __uint_as_float((uint32_t)(268 - e) << 23);
  is = __uint_as_float((uint32_t)(e - 14) << 23);
}

template <int DH>
struct F16Layout {
  static constexpr int E = 16;
  static constexpr int KD = DH / 16;  // k16 steps over the head dim
  static constexpr int ND = DH / 8;   // 8-wide n tiles of a [., DH] output
  static constexpr int NC = DH * E;   // intensity MLP width
  static constexpr int MT = NC / 8;   // 8-column tiles of the MLP output
  // words per K / W1^T row: KD blocks of 16 words ({hi,hi,lo,lo} x 4 lanes); = 16 (mod 32) keeps LDS.128 conflict free
  static constexpr int SKW = KD * 16 + ((KD % 2 == 0) ? 16 : 0);
  // constant pack (global image == shared image), byte offsets
  static constexpr int BW_OFF = NC * SKW * 4;               // float4 per column pair {b1[c], b1[c+1], wsp[c], wsp[c+1]}
  static constexpr int WV_OFF = BW_OFF + (NC / 2) * 16;     // int_weight [NC]
  static constexpr int SC_OFF = WV_OFF + NC * 4;            // exp(scaling) [E]
  static constexpr int MISC_OFF = SC_OFF + E * 4;           // {1 / scale(W1), 0, 0, 0}
  static constexpr int PACK_BYTES = MISC_OFF + 16;
  // K / V / T rows in shared memory: [key]{hi[DH] | lo[DH] | 16 B pad} fp16; an odd number of 16-byte chunks per
  // row keeps the eight row addresses of an ldmatrix phase on distinct bank groups
  static constexpr int RB = 4 * DH + 16;
  __host__ __device__ static constexpr size_t smem_bytes(int NT) {
    const int NB = (NT + 1) / 2, LP = NT * 8, RP = NB * 16;
    return (size_t)PACK_BYTES + 3 * (size_t)RP * RB + (size_t)LP * 32 + (size_t)LP * 4 + 32;
  }
};

// Built once per commit: W1 (rows 0..DH-1 of int_w, times -log2 e) transposed to [column][dim] fp16 (hi, lo)
// fragments, the span row and bias (times -log2 e), int_weight, exp(scaling) and the W1 scale.
template <int DH>
__global__ void __launch_bounds__(256) mlp_pack_kernel(const float* __restrict__ int_w, const float* __restrict__ int_b,
                                                       const float* __restrict__ int_weight,
                                                       const float* __restrict__ int_scaling,
                                                       unsigned char* __restrict__ pack) {
  using LY = F16Layout<DH>;
  constexpr int NC = LY::NC, KD = LY::KD, SKW = LY::SKW, E = LY::E;
  __shared__ unsigned int mx;
  if (threadIdx.x == 0) mx = 0u;
  __syncthreads();
  float m = 0.f;
  for (int i = threadIdx.x; i < DH * NC; i += blockDim.x) m = fmaxf(m, fabsf(kLog2e * int_w[i]));
  atomicMax(&mx, __float_as_uint(m));
  __syncthreads();
  float sw, isw;
  pow2_scale(__uint_as_float(mx), sw, isw);
  uint32_t* w1 = reinterpret_cast<uint32_t*>(pack);
  for (int i = threadIdx.x; i < NC * SKW; i += blockDim.x) w1[i] = 0u;
  __syncthreads();
  for (int i = threadIdx.x; i < NC * KD * 4; i += blockDim.x) {
    const int c = i / (KD * 4), ks = (i / 4) % KD, t = i % 4;
    auto W = [&](int s) { return -kLog2e * int_w[(size_t)(ks * 16 + s) * NC + c] * sw; };
    uint4 v;
    split2(W(2 * t), W(2 * t + 1), v.x, v.z);          // b0: k slots 2t, 2t+1
    split2(W(2 * t + 8), W(2 * t + 9), v.y, v.w);      // b1: k slots 2t+8, 2t+9
    *reinterpret_cast<uint4*>(w1 + (size_t)c * SKW + ks * 16 + t * 4) = v;
  }
  float4* bw = reinterpret_cast<float4*>(pack + LY::BW_OFF);
  for (int i = threadIdx.x; i < NC / 2; i += blockDim.x)
    bw[i] = make_float4(-kLog2e * int_b[2 * i], -kLog2e * int_b[2 * i + 1], -kLog2e * int_w[(size_t)DH * NC + 2 * i],
                        -kLog2e * int_w[(size_t)DH * NC + 2 * i + 1]);
  float* wv = reinterpret_cast<float*>(pack + LY::WV_OFF);
  for (int i = threadIdx.x; i < NC; i += blockDim.x) wv[i] = int_weight[i];
  float* sc = reinterpret_cast<float*>(pack + LY::SC_OFF);
  for (int i = threadIdx.x; i < E; i += blockDim.x) sc[i] = expf(int_scaling[i]);  // temporal.py:302
  float* misc = reinterpret_cast<float*>(pack + LY::MISC_OFF);
  if (threadIdx.x < 4) misc[threadIdx.x] = threadIdx.x == 0 ? isw : 0.f;
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// out[ND][4] = A[NT][.] (accumulator layout: rows g / g+8, keys nt*8 + 2t + (c&1)) times X, X staged row-major as
// [key]{hi[DH] | lo[DH] | pad} fp16 and read with ldmatrix.trans (xs = this lane's row address for key block 0,
// dims 0..15, hi).  Two key blocks are in flight on separate accumulators, so an accumulator is touched once per
// 2*ND MMAs.
// ABL: precision ablation (common.cuh): bit 2 drops the P_lo X_hi product, bit 3 the P_hi X_lo product.
template <int DH, int NT, int ABL = 0>
__device__ __forceinline__ void pv_product16(const float (&P)[NT][4], uint32_t xs, float (&out)[DH / 8][4]) {
  constexpr int ND = DH / 8, NB = (NT + 1) / 2, RB = F16Layout<DH>::RB;
  float acc[2][ND][4];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int n = 0; n < ND; ++n) acc[p][n][0] = acc[p][n][1] = acc[p][n][2] = acc[p][n][3] = 0.f;
#pragma unroll
  for (int j0 = 0; j0 < NB; j0 += 2) {
    uint32_t ah[2][4], al[2][4];
    uint32_t xh[2][ND / 2][4], xl[2][ND / 2][4];
#pragma unroll
    for (int p = 0; p < 2; ++p)
      if (j0 + p < NB) {
        const int j = j0 + p;
        split2(P[2 * j][0], P[2 * j][1], ah[p][0], al[p][0]);
        split2(P[2 * j][2], P[2 * j][3], ah[p][1], al[p][1]);
        if (2 * j + 1 < NT) {
          split2(P[2 * j + 1][0], P[2 * j + 1][1], ah[p][2], al[p][2]);
          split2(P[2 * j + 1][2], P[2 * j + 1][3], ah[p][3], al[p][3]);
        } else {
          ah[p][2] = ah[p][3] = al[p][2] = al[p][3] = 0u;
        }
#pragma unroll
        for (int np = 0; np < ND / 2; ++np) {
          ldsm_x4_trans(xh[p][np], xs + j * 16 * RB + np * 32);
          ldsm_x4_trans(xl[p][np], xs + j * 16 * RB + np * 32 + DH * 2);
        }
      }
    if constexpr ((ABL & 4) == 0) {
#pragma unroll
      for (int p = 0; p < 2; ++p)
        if (j0 + p < NB)
#pragma unroll
          for (int n = 0; n < ND; ++n) mma_f16(acc[p][n], al[p], xh[p][n >> 1][(n & 1) * 2], xh[p][n >> 1][(n & 1) * 2 + 1]);
    }
    if constexpr ((ABL & 8) == 0) {
#pragma unroll
      for (int p = 0; p < 2; ++p)
        if (j0 + p < NB)
#pragma unroll
          for (int n = 0; n < ND; ++n) mma_f16(acc[p][n], ah[p], xl[p][n >> 1][(n & 1) * 2], xl[p][n >> 1][(n & 1) * 2 + 1]);
    }
#pragma unroll
    for (int p = 0; p < 2; ++p)
      if (j0 + p < NB)
#pragma unroll
        for (int n = 0; n < ND; ++n) mma_f16(acc[p][n], ah[p], xh[p][n >> 1][(n & 1) * 2], xh[p][n >> 1][(n & 1) * 2 + 1]);
  }
#pragma unroll
  for (int n = 0; n < ND; ++n)
#pragma unroll
    for (int c = 0; c < 4; ++c) out[n][c] = acc[0][n][c] + acc[1][n][c];
}

__device__ __forceinline__ float absmax4(float m, const float4& v) {
  return fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
}
// 4 consecutive dims of one key -> 8 bytes of hi and 8 bytes of lo in the key's row
__device__ __forceinline__ void put4(unsigned char* row, int dim0, int DH2, const float4& v, float s) {
  uint2 hi, lo;
  split2(v.x * s, v.y * s, hi.x, lo.x);
  split2(v.z * s, v.w * s, hi.y, lo.y);
  *reinterpret_cast<uint2*>(row + dim0 * 2) = hi;
  *reinterpret_cast<uint2*>(row + DH2 + dim0 * 2) = lo;
}


}  // namespace f16c
}  // namespace edgl
