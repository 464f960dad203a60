// tc_common.cuh - pieces shared by the two tcgen05 GEMMs (gemm_tc.cu: 3xTF32, gemm_f16.cu: scaled 3xFP16): the epilogue
// (TMEM -> bias / activation / residual -> global, staged or direct), the GELU,
// mbarrier / TMA / tcgen05 fence, commit and TMEM-load wrappers, and the host-side tensor-map encoder.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace edgl {
namespace tcc {

constexpr int BK = 32;  // elements per k-block of either kernel (the TMA box is [box_rows][BK])

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// see gemm_tc.cu: thread (g = lane/4, t = lane%4): r[4j], r[4j+1] = row g, columns 8j + 2t, 2t+1; r[4j+2], r[4j+3] = row g+8
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------- epilogue
// Warp roles of both kernels: warps 0-3 control (TMA producer, MMA issuer, TMEM allocator, spare), then EPI_WARPS
// epilogue warps (EPI_PARTS per TMEM lane quarter, taking every EPI_PARTS-th 16-column chunk), then 4 splitter warps.
// The layers are bound by their epilogues (bias / GELU / residual / LayerNorm pieces on 128 x BN fp32 values per tile
// with TMEM -> shared -> global hops in between), so the epilogue gets 12 warps; setmaxnreg moves the registers the
// control and splitter warps do not need over to them.
#ifndef EDGL_EPI_PARTS
#define EDGL_EPI_PARTS 3
#endif
constexpr int EPI_PARTS = EDGL_EPI_PARTS;
constexpr int EPI_WARPS = 4 * EPI_PARTS;
constexpr int SPLIT_WARP0 = 4 + EPI_WARPS;
constexpr int TC_THREADS = (SPLIT_WARP0 + 4) * 32;
constexpr int BM = 128;      // UMMA M of both kernels
constexpr int CW = 16;       // epilogue sub-chunk width (columns per tcgen05.ld)
constexpr int CP = CW + 4;   // padded pitch of the epilogue staging tile

// GELU(x) = x * 0.5 * (1 + erf(x / sqrt 2)) (EasyDGL.py:31-32, Q18) with a branch-free erf:
//   erf(t) = 1 - 2^(-t * g(t)),  g = degree-7 minimax fit of -log2(erfc(t)) / t on [0, 4] (erf(t >= 4) = 1 in fp32).
// Max |erf error| 1.0e-7, i.e. the rounding of an fp32 erff; the resulting GELU differs from the float64 one by at
// most 1.1e-7 * |x| - the same bound as the erff-based fp32 form (fit and emulation: DESIGN.md 4).  14 instructions
// instead of ~45: the GELU epilogues (FF1, transform) are bound by exactly these issue slots.
__device__ __forceinline__ float gelu_fit(float x) {
  const float t = fminf(fabsf(x) * 0.70710678118654752440f, 4.0f);
  float p = 4.5358559873420745e-05f;
  p = fmaf(p, t, -0.00044550723396241665f);
  p = fmaf(p, t, 0.0014894399791955948f);
  p = fmaf(p, t, 0.0007746326737105846f);
  p = fmaf(p, t, -0.02825368382036686f);
  p = fmaf(p, t, 0.14848162233829498f);
  p = fmaf(p, t, 0.9184163808822632f);
  p = fmaf(p, t, 1.6279085874557495f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-t * p));
  return fmaf(fabsf(x), fmaf(e, -0.5f, 0.5f), 0.5f * x);
}

// the same fit on two values at once with packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2: one issue slot for two IEEE
// operations, bit-identical to the scalar form): 9 instead of 14 issue slots per element
__device__ __forceinline__ void gelu_fit2(float& x0, float& x1) {
  typedef unsigned long long f2;
  auto pk = [](float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; };
  auto fma2 = [](f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; };
  auto mul2 = [](f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; };
  const float a0 = fabsf(x0), a1 = fabsf(x1);
  const float t0 = fminf(a0 * 0.70710678118654752440f, 4.0f), t1 = fminf(a1 * 0.70710678118654752440f, 4.0f);
  const f2 t = pk(t0, t1);
  f2 p = pk(4.5358559873420745e-05f, 4.5358559873420745e-05f);
  p = fma2(p, t, pk(-0.00044550723396241665f, -0.00044550723396241665f));
  p = fma2(p, t, pk(0.0014894399791955948f, 0.0014894399791955948f));
  p = fma2(p, t, pk(0.0007746326737105846f, 0.0007746326737105846f));
  p = fma2(p, t, pk(-0.02825368382036686f, -0.02825368382036686f));
  p = fma2(p, t, pk(0.14848162233829498f, 0.14848162233829498f));
  p = fma2(p, t, pk(0.9184163808822632f, 0.9184163808822632f));
  p = fma2(p, t, pk(1.6279085874557495f, 1.6279085874557495f));
  float q0, q1;
  const f2 tp = mul2(pk(-t0, -t1), p);
  asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(tp));
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
  const f2 w = fma2(pk(e0, e1), pk(-0.5f, -0.5f), pk(0.5f, 0.5f));
  const f2 r = fma2(pk(a0, a1), w, mul2(pk(0.5f, 0.5f), pk(x0, x1)));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(r));
}

__device__ __forceinline__ float gelu_erf_exact(float x) {
  return x * (0.5f * (1.0f + erff(x * 0.70710678118654752440f)));  // EasyDGL.py:31-32
}

// One 128 x BN accumulator tile -> global memory, for one epilogue warp (TMEM lane quarter q, column part `part`).
// Two paths per 16-column chunk:
//  * staged: tcgen05.ld hands every thread one accumulator ROW; writing rows straight to global would cost 32 cache
//    lines per instruction, so the 32x16 chunk is transposed through a padded shared-memory tile and handled 8 rows x
//    64 B per warp instruction: bias / periodic bias / residual loads and the store are 16-byte-per-lane accesses, and
//    every load of a chunk is issued before the first store (C and R may alias as far as the compiler knows);
//  * direct (p.epi_direct): 16x256b fragments, rows stored sector by sector from registers, no shared memory.
// SCALED (gemm_f16.cu): the accumulator is multiplied by `inv` (1 / (sa * sw)) first and max|C| is tracked in cmax.
// LNF (bit mask, gemm_tc.cu only): fused LayerNorm pieces of LnEpi (common.cuh) - LN_RES: the residual rows are
// LayerNorm(R); LN_STATS: per-row partial sums of the stored values; LN_LAST: store only the last row of every
// sequence.  (LN_A, the normalisation of the A operand, lives in the splitter warps.)
// LN_FILT (same bit mask, not a LayerNorm piece): the epilogue stores nothing and appends the values that pass the
// row's threshold to the top-K candidate lists of p.flt (TopkFilter, common.cuh).
constexpr int LN_A = 1, LN_RES = 2, LN_STATS = 4, LN_LAST = 8, LN_FILT = 16;

template <int BN, int ACT, bool SCALED, int LNF = 0, class P>
__device__ __forceinline__ void epilogue_tile(const P& p, uint32_t tmem_acc, int m0, int n0, int q, int part, int lane,
                                              float* stg, bool all_al, const int (&pbo)[4], const int (&pbd)[4], float inv,
                                              float& cmax) {
  constexpr int LPR = CW / 4;        // lanes per row (4)
  constexpr int RPI = 32 / LPR;      // rows per warp instruction (8)
  constexpr int NIT = 32 / RPI;      // iterations per sub-chunk (4)
  const int cl = (lane % LPR) * 4;   // this lane's 4 columns inside a sub-chunk
  const int rsub = lane / LPR;       // this lane's row inside a group of RPI
  const int rbase = m0 + q * 32 + rsub;  // rows rbase + RPI*itr
  auto acc_val = [&](uint32_t bits) {
    float x = __uint_as_float(bits);
    if constexpr (SCALED) x *= inv;
    return x;
  };
  if constexpr ((LNF & LN_FILT) != 0) {
    // ---- top-K candidate filter (logits layer: bias only).  tcgen05.ld hands this thread 16 consecutive columns of
    // ONE row; the value is formed exactly as the storing paths form it (accumulator, column 0 = bias, + bias), so a
    // candidate carries the bits the materialised logits would.  One atomic per (row, 16 columns) that has hits.
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    const float thr = row_ok ? p.flt.thr[(size_t)row * p.flt.thr_stride] : INFINITY;
    const bool bias_al = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
#pragma unroll 1
    for (int c0 = part * CW; c0 < BN; c0 += EPI_PARTS * CW) {
      uint32_t r[CW];
      tmem_ld16(tmem_acc + c0 + ((uint32_t)(q * 32) << 16), r);
      const int colb = n0 + c0;
      if (colb >= p.N) continue;  // warp-uniform
      float v[CW];
      if (bias_al && colb + CW <= p.N) {
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(p.bias + colb + 4 * j);
          v[4 * j] = b4.x; v[4 * j + 1] = b4.y; v[4 * j + 2] = b4.z; v[4 * j + 3] = b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < CW; ++j) v[j] = (p.bias && colb + j < p.N) ? p.bias[colb + j] : 0.f;
      }
      unsigned int hits = 0;
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        float x = acc_val(r[j]);
        if (p.col0_bias_only && colb + j == 0) x = 0.f;
        v[j] += x;
        if (colb + j < p.N && v[j] >= thr) hits |= 1u << j;
      }
      if (hits != 0u) {
        unsigned int pos = atomicAdd(p.flt.cnt + row, (unsigned int)__popc(hits));
        unsigned long long* dst = p.flt.cand + (size_t)row * p.flt.cap;
#pragma unroll
        for (int j = 0; j < CW; ++j)
          if (hits & (1u << j)) {
            if (pos < (unsigned int)p.flt.cap) dst[pos] = compose(f2key(v[j]), (uint32_t)(colb + j));
            ++pos;
          }
      }
    }
    return;
  }
  // fused LayerNorm pieces: (mean, rstd) of the sequence each of this lane's rows belongs to, the rows' running
  // partial sums, and the output row of a last-row-only store (-1: not a last row)
  float2 lnr[NIT];
  float st_s[NIT], st_q[NIT];
  int lnlast[NIT];
  if constexpr (LNF != 0) {
#pragma unroll
    for (int itr = 0; itr < NIT; ++itr) {
      lnr[itr] = make_float2(0.f, 1.f);
      st_s[itr] = st_q[itr] = 0.f;
      lnlast[itr] = -1;
      const int row = rbase + RPI * itr;
      if (row < p.M) {
        const int sq = row / p.ln.L;
        if constexpr ((LNF & LN_RES) != 0) lnr[itr] = p.ln.r_rs[sq];
        if (row - sq * p.ln.L == p.ln.L - 1) lnlast[itr] = sq;
      }
    }
  }
#pragma unroll 1
  for (int c0 = part * CW; c0 < BN; c0 += EPI_PARTS * CW) {
    if (LNF == 0 && p.epi_direct && all_al && n0 + c0 + CW <= p.N) {
      // ---- direct path: two 16-row halves; every store instruction writes 8 rows x one 32-byte sector
      const int g = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        uint32_t r8[8];
        tmem_ld_16x256b_x2(tmem_acc + c0 + ((uint32_t)(q * 32 + hr * 16) << 16), r8);
        float2 add[2][2], res[2][2];
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int row = m0 + q * 32 + hr * 16 + h8 * 8 + g, col = n0 + c0 + 8 * j + t2;
            float2 a2 = p.bias ? *reinterpret_cast<const float2*>(p.bias + col) : make_float2(0.f, 0.f);
            if (p.pbias) {
              const float2 pp = *reinterpret_cast<const float2*>(p.pbias + pbd[hr * 2 + h8] + col);
              a2.x += pp.x; a2.y += pp.y;
            }
            add[h8][j] = a2;
            res[h8][j] = (p.R && row < p.M) ? *reinterpret_cast<const float2*>(p.R + (size_t)row * p.ldr + col)
                                            : make_float2(0.f, 0.f);
          }
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int row = m0 + q * 32 + hr * 16 + h8 * 8 + g, col = n0 + c0 + 8 * j + t2;
            float2 v = make_float2(acc_val(r8[4 * j + 2 * h8]), acc_val(r8[4 * j + 2 * h8 + 1]));
            if (p.col0_bias_only && col == 0) v.x = 0.f;
            v.x += add[h8][j].x; v.y += add[h8][j].y;
            if (ACT == ACT_GELU) {
              if (p.gelu_fit) gelu_fit2(v.x, v.y);
              else { v.x = gelu_erf_exact(v.x); v.y = gelu_erf_exact(v.y); }
            }
            if (ACT == ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
            v.x += res[h8][j].x; v.y += res[h8][j].y;
            if (row < p.M) {
              *reinterpret_cast<float2*>(p.C + (size_t)row * p.ldc + col) = v;
              if constexpr (SCALED) cmax = fmaxf(cmax, fmaxf(fabsf(v.x), fabsf(v.y)));
            }
          }
      }
      continue;
    }
    uint32_t r[CW];
    tmem_ld16(tmem_acc + c0 + ((uint32_t)(q * 32) << 16), r);
    if (n0 + c0 >= p.N) continue;  // warp-uniform
#pragma unroll
    for (int j = 0; j < CW / 4; ++j)
      *reinterpret_cast<float4*>(stg + lane * CP + 4 * j) =
          make_float4(acc_val(r[4 * j]), acc_val(r[4 * j + 1]), acc_val(r[4 * j + 2]), acc_val(r[4 * j + 3]));
    __syncwarp();
    const int col = n0 + c0 + cl;
    if (all_al && n0 + c0 + CW <= p.N) {
      // ---- fast path: whole sub-chunk inside N, everything 16-byte aligned
      float4 rr[NIT], pp[NIT];
      if (p.R) {
#pragma unroll
        for (int itr = 0; itr < NIT; ++itr) {
          const int row = rbase + RPI * itr;
          rr[itr] = row < p.M ? *reinterpret_cast<const float4*>(p.R + (size_t)row * p.ldr + col)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if constexpr ((LNF & LN_RES) != 0) {  // the residual is LayerNorm(R): (R - mean) rstd gamma + beta
          const float4 g4 = *reinterpret_cast<const float4*>(p.ln.r_g + col);
          const float4 e4 = *reinterpret_cast<const float4*>(p.ln.r_b + col);
#pragma unroll
          for (int itr = 0; itr < NIT; ++itr) {
            const float rs = lnr[itr].y, nm = -lnr[itr].x * lnr[itr].y;
            rr[itr].x = fmaf(fmaf(rr[itr].x, rs, nm), g4.x, e4.x);
            rr[itr].y = fmaf(fmaf(rr[itr].y, rs, nm), g4.y, e4.y);
            rr[itr].z = fmaf(fmaf(rr[itr].z, rs, nm), g4.z, e4.z);
            rr[itr].w = fmaf(fmaf(rr[itr].w, rs, nm), g4.w, e4.w);
          }
        }
      }
      if (p.pbias) {
#pragma unroll
        for (int itr = 0; itr < NIT; ++itr) pp[itr] = *reinterpret_cast<const float4*>(p.pbias + pbo[itr] + col);
      }
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias) b4 = *reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
      for (int itr = 0; itr < NIT; ++itr) {
        const int row = rbase + RPI * itr;
        float4 v = *reinterpret_cast<const float4*>(stg + (rsub + RPI * itr) * CP + cl);
        if (p.col0_bias_only && col == 0) v.x = 0.f;
        if (p.pbias) { v.x += pp[itr].x; v.y += pp[itr].y; v.z += pp[itr].z; v.w += pp[itr].w; }
        v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
        if (ACT == ACT_GELU) {
          if (p.gelu_fit) { gelu_fit2(v.x, v.y); gelu_fit2(v.z, v.w); }
          else { v.x = gelu_erf_exact(v.x); v.y = gelu_erf_exact(v.y); v.z = gelu_erf_exact(v.z); v.w = gelu_erf_exact(v.w); }
        }
        if (ACT == ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (p.R) { v.x += rr[itr].x; v.y += rr[itr].y; v.z += rr[itr].z; v.w += rr[itr].w; }
        if (row < p.M) {
          if constexpr ((LNF & LN_STATS) != 0) {
            st_s[itr] += (v.x + v.y) + (v.z + v.w);
            st_q[itr] += fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w);
          }
          if constexpr ((LNF & LN_LAST) != 0) {
            if (lnlast[itr] >= 0) *reinterpret_cast<float4*>(p.C + (size_t)lnlast[itr] * p.ldc + col) = v;
          } else {
            *reinterpret_cast<float4*>(p.C + (size_t)row * p.ldc + col) = v;
          }
          if constexpr (SCALED) cmax = fmaxf(cmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
      }
    } else {
      // ---- generic path (N tail / unaligned operands): scalar, guarded
#pragma unroll 1
      for (int itr = 0; itr < NIT; ++itr) {
        const int row = rbase + RPI * itr;
        if (row >= p.M) continue;
#pragma unroll 1
        for (int e = 0; e < 4; ++e) {
          const int c = col + e;
          if (c >= p.N) continue;
          float x = stg[(rsub + RPI * itr) * CP + cl + e];
          if (p.col0_bias_only && c == 0) x = 0.f;
          if (p.pbias) x += p.pbias[(size_t)(row % p.pperiod) * p.N + c];
          if (p.bias) x += p.bias[c];
          if (ACT == ACT_GELU) x = (p.gelu_fit ? gelu_fit(x) : gelu_erf_exact(x));
          if (ACT == ACT_RELU) x = fmaxf(x, 0.f);
          if (p.R) x += p.R[(size_t)row * p.ldr + c];
          p.C[(size_t)row * p.ldc + c] = x;
          if constexpr (SCALED) cmax = fmaxf(cmax, fabsf(x));
        }
      }
    }
    __syncwarp();
  }
  if constexpr ((LNF & LN_STATS) != 0) {
    // the four lanes of a row hold its partial sums over this warp's column chunks: one (sum, sum of squares) pair
    // per (row, n-tile, column half), added in a fixed order (deterministic)
    const int nparts = EPI_PARTS * ((p.N + BN - 1) / BN), pidx = EPI_PARTS * (n0 / BN) + part;
#pragma unroll
    for (int itr = 0; itr < NIT; ++itr) {
      float s = st_s[itr], q2 = st_q[itr];
      s += __shfl_xor_sync(0xffffffffu, s, 1); q2 += __shfl_xor_sync(0xffffffffu, q2, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2); q2 += __shfl_xor_sync(0xffffffffu, q2, 2);
      const int row = rbase + RPI * itr;
      if ((lane % LPR) == 0 && row < p.M) p.ln.stats[(size_t)row * nparts + pidx] = make_float2(s, q2);
    }
  }
}

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// 2-D tensor [rows][cols] of fp32 (128B swizzle) or fp16 (64B swizzle); box = [box_rows][32 elements]
static int make_map(CUtensorMap* m, const void* ptr, bool f16, long long rows, long long cols, long long ld_elems,
                    int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(-3, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * (f16 ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-3, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

}  // namespace tcc
}  // namespace edgl
