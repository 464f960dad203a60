// attn_tc.cu - Blackwell-native (tcgen05 / TMEM) fused self-modulating attention core for dh = 16, E = 16,
// L <= 128 (BASELINE configs C2 and C3).  Same math as attn.cu / attn_mma.cuh (temporal.py:281-315, 345-385,
// 412-447), 3xTF32 everywhere.
//
// One persistent CTA per SM; TWO (sequence, head) items are in flight per CTA ("groups"), each with its own
// 128 row threads (TMEM lane r = query row r = key row r), its own MMA-issuing thread, its own 256 TMEM
// columns, operand tiles and mbarriers - so the SFU-bound phases of one item (softmax, 256 sigmoids per row)
// overlap the tensor-core phases of the other.  Per item:
//
//   S  = Q K^T            SS  A=Q (smem)          B=K (smem)         -> TMEM[0,112)     (pre-scaled by log2e/sqrt(dh))
//   softmax               ld S -> exp2 -> P_un kept in REGISTERS; (hi, lo) copies st to TMEM[0,112) / [112,224)
//   Hu = P_un T           TS  A=P (TMEM)          B=T^T (smem)       -> TMEM[224,240)
//   Z  = [H,span,1] W1'   TS  A=[H,span,1] (TMEM[0,64)) B=W1'^T (smem) -> TMEM[64,128) / [128,192): four 64-column
//                                                                       quarters double buffered against the sigmoid
//   G  = lam M^T          TS  A=lam (TMEM[192,224)) B=marks (smem)   -> TMEM[0,112)
//   gate                  ld G, times the register copy of P (set_diag) -> st hi in place, lo to TMEM[112,224)
//   Ou = (G o P) V        TS  A=GoP (TMEM)        B=V^T (smem)       -> TMEM[240,256);  O = Ou / l + residual
//
// Shared-memory operands are written by the row threads directly in the K-major SWIZZLE_128B layout the UMMA
// descriptors of gemm_tc.cu use (rows of 128 B, 16-byte chunk c of row r stored at chunk c ^ (r & 7)); Q and K
// tiles hold hi in columns 0-15 and lo in columns 16-31.  W1' = -log2(e) * [W1 ; w_span ; b1] (shared by both
// groups) so the MMA yields -z*log2(e) and sigmoid = rcp(1 + ex2(.)).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace edgl {
namespace atc {

constexpr int DH = 16, E = 16, NC = DH * E;      // 256 MLP columns
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kFill = -4294967296.0f;          // float(-2**32+1), temporal.py:358,425

// ---- TMEM column map of one group (256 columns; group g starts at column 256*g)
constexpr uint32_t C_S = 0;     // 112: S -> P_un hi ... later G -> (G o P) hi
constexpr uint32_t C_PL = 112;  // 112: P_un lo ... later (G o P) lo
constexpr uint32_t C_AH = 0;    // 32 : [H, span, 1, 0...] hi (24 used)   (MLP phase: P regions are dead, P is in registers)
constexpr uint32_t C_AL = 32;   // 32 : lo
constexpr uint32_t C_Z = 64;    // 128: Z double buffer (2 x 64)
constexpr uint32_t C_LH = 192;  // 16 : lam hi
constexpr uint32_t C_LL = 208;  // 16 : lam lo
constexpr uint32_t C_G = 0;     // 112: G accumulator (after the MLP)
constexpr uint32_t C_HU = 224;  // 16 : H_un accumulator
constexpr uint32_t C_O = 240;   // 16 : O_un accumulator
constexpr uint32_t C_GROUP = 256;

// ---- shared memory map (bytes); every operand tile is 1024-byte aligned
constexpr int T128 = 128 * 128;                  // [128 rows][32 floats] tile
constexpr int OFF_Q = 0;                         // Q  : hi cols 0-15 | lo cols 16-31
constexpr int OFF_K = OFF_Q + T128;              // K  : hi | lo
constexpr int OFF_MK = OFF_K + T128;             // marks [128 keys][32] (16 events used)
constexpr int TT = 4 * 16 * 128;                 // T^T / V^T: 4 key-atoms of [16 rows][32 keys]
constexpr int OFF_TH = OFF_MK + T128, OFF_TL = OFF_TH + TT, OFF_VH = OFF_TL + TT, OFF_VL = OFF_VH + TT;
constexpr int GROUP_BYTES = OFF_VL + TT;         // 80 KB per group
constexpr int W1B = 256 * 128;                   // W1'^T [256 n][32 k] (24 used)
constexpr int OFF_WH = 2 * GROUP_BYTES, OFF_WL = OFF_WH + W1B;
constexpr int OFF_WV = OFF_WL + W1B;             // int_weight [256] floats
constexpr int OFF_SC = OFF_WV + 1024;            // exp(scaling) [16]
constexpr int OFF_KM = OFF_SC + 64;              // key min-mask [2][128]
constexpr int OFF_BAR = OFF_KM + 1024;           // mbarriers [2][B_COUNT] + tmem slot
constexpr int SMEM_BYTES = OFF_BAR + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_tf32(int n) {  // D=f32, A=B=tf32, K-major, M=128
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
      "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

#define EDGL_R16(r) r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14], r[15]
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&f)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lo_of(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// byte offset of element (r, k) in a [rows][32 floats] K-major SWIZZLE_128B tile
__device__ __forceinline__ int sw128(int r, int k) {
  return (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + ((k & 3) << 2);
}
// store 4 consecutive k (k % 4 == 0) of row r as (hi, lo) into two tiles
__device__ __forceinline__ void put4(uint8_t* th, uint8_t* tl, int r, int k, float4 v) {
  const int o = sw128(r, k);
  *reinterpret_cast<float4*>(th + o) = v;
  *reinterpret_cast<float4*>(tl + o) = make_float4(lo_of(v.x), lo_of(v.y), lo_of(v.z), lo_of(v.w));
}

enum { B_STAGED = 0, B_S, B_P, B_HU, B_A, B_Z0, B_Z1, B_ZF0, B_ZF1, B_L, B_G, B_GP, B_O, B_COUNT };

struct GroupCtx {
  uint8_t* sm;       // this group's operand tiles
  uint8_t* smw;      // shared W1' tiles (OFF_WH / OFF_WL relative to the CTA base)
  uint64_t* bar;     // this group's mbarriers
  float* km;         // this group's key min-mask
  const float* wv;
  const float* sc;
  uint32_t tm;       // this group's TMEM base column
  int first_item, item_stride;
};

// =================================================================== MMA issuer (one thread per group)
// A single thread issues ~124 MMAs per item, so the per-MMA instruction overhead IS the tensor-pipe feed rate:
// every shared-memory descriptor is built once (the operand tiles never move) and advanced by adding the
// byte offset >> 4 to its start-address field; the loops are unrolled.
__device__ void run_mma(const AttnArgs& a, const GroupCtx& g, int num_items) {
  const int L = a.L;
  const int NS = (L + 15) & ~15, KSP = (L + 7) >> 3;
  const uint32_t idS = idesc_tf32(NS), id16 = idesc_tf32(16), id64 = idesc_tf32(64);
  const uint64_t dq = umma_desc(smem_u32(g.sm + OFF_Q)), dk = umma_desc(smem_u32(g.sm + OFF_K)),
                 dmk = umma_desc(smem_u32(g.sm + OFF_MK)), dth = umma_desc(smem_u32(g.sm + OFF_TH)),
                 dtl = umma_desc(smem_u32(g.sm + OFF_TL)), dvh = umma_desc(smem_u32(g.sm + OFF_VH)),
                 dvl = umma_desc(smem_u32(g.sm + OFF_VL)), dwh = umma_desc(smem_u32(g.smw)),
                 dwl = umma_desc(smem_u32(g.smw + W1B));
  const uint32_t tm = g.tm;
  uint64_t* bar = g.bar;
  // P V product: A = (hi, lo) copies of a [128 x keys] matrix in TMEM, B = X^T (hi, lo) tiles, 3 products per key step
  auto pv = [&](uint32_t dacc, uint64_t dxh, uint64_t dxl) {
#pragma unroll 1
    for (int ks0 = 0; ks0 < KSP; ks0 += 4) {  // one 32-key atom (2048 B) per outer step
      const uint64_t ah = dxh + (uint64_t)(ks0 >> 2) * 128, al = dxl + (uint64_t)(ks0 >> 2) * 128;
      const uint32_t ta = tm + C_S + ks0 * 8, tl_ = tm + C_PL + ks0 * 8;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (ks0 + k < KSP) {
          mma_ts(dacc, tl_ + k * 8, ah + k * 2, id16, (ks0 + k) != 0);
          mma_ts(dacc, ta + k * 8, al + k * 2, id16, 1);
          mma_ts(dacc, ta + k * 8, ah + k * 2, id16, 1);
        }
      }
    }
  };
  uint32_t n = 0;
  for (int item = g.first_item; item < num_items; item += g.item_stride, ++n) {
    const uint32_t ph = n & 1;
    // ---- S = Q K^T   (hi k-steps at byte offsets 0, 32; lo at 64, 96 of the 128-byte rows)
    mbar_wait(&bar[B_STAGED], ph);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < DH / 8; ++ks) {
      const uint64_t oh = ks * 2, ol = 4 + ks * 2;  // byte offsets >> 4
      mma_ss(tm + C_S, dq + ol, dk + oh, idS, ks != 0);
      mma_ss(tm + C_S, dq + oh, dk + ol, idS, 1);
      mma_ss(tm + C_S, dq + oh, dk + oh, idS, 1);
    }
    umma_commit(&bar[B_S]);
    // ---- Hu = P_un T
    mbar_wait(&bar[B_P], ph);
    tc_fence_after();
    pv(tm + C_HU, dth, dtl);
    umma_commit(&bar[B_HU]);
    // ---- Z = [H, span, 1] W1'  in four 64-column quarters, two TMEM buffers
    mbar_wait(&bar[B_A], ph);
    tc_fence_after();
#pragma unroll
    for (int qz = 0; qz < 4; ++qz) {
      const int buf = qz & 1;
      const uint32_t use = 2 * n + (qz >> 1);  // how many times this buffer has been handed out before
      mbar_wait(&bar[B_ZF0 + buf], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t dz = tm + C_Z + buf * 64;
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        const uint64_t o = (uint64_t)(qz * 64 * 128 + ks * 32) >> 4;
        mma_ts(dz, tm + C_AL + ks * 8, dwh + o, id64, ks != 0);
        mma_ts(dz, tm + C_AH + ks * 8, dwl + o, id64, 1);
        mma_ts(dz, tm + C_AH + ks * 8, dwh + o, id64, 1);
      }
      umma_commit(&bar[B_Z0 + buf]);
    }
    // ---- G = lam M^T  (marks are exact in TF32: two products)
    mbar_wait(&bar[B_L], ph);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < E / 8; ++ks) {
      mma_ts(tm + C_G, tm + C_LL + ks * 8, dmk + ks * 2, idS, ks != 0);
      mma_ts(tm + C_G, tm + C_LH + ks * 8, dmk + ks * 2, idS, 1);
    }
    umma_commit(&bar[B_G]);
    // ---- Ou = (G o P) V
    mbar_wait(&bar[B_GP], ph);
    tc_fence_after();
    pv(tm + C_O, dvh, dvl);
    umma_commit(&bar[B_O]);
  }
}

// =================================================================== row threads (r = query row = key row)
__device__ void run_rows(const AttnArgs& a, const GroupCtx& g, int num_items, int r, int bar_id) {
  const int L = a.L;
  const int NS = (L + 15) & ~15;
  uint8_t* sm = g.sm;
  uint64_t* bar = g.bar;
  float* km = g.km;
  const uint32_t tml = g.tm + ((uint32_t)((r >> 5) * 32) << 16);  // TMEM address of this warp's lane quarter
  const bool row_ok = r < L;
  const float sc2 = kLog2e / sqrtf((float)DH);
  uint32_t n = 0;
  for (int item = g.first_item; item < num_items; item += g.item_stride, ++n) {
    const uint32_t ph = n & 1;
    const int b = item / a.h, hh = item % a.h;
    const long long grow = (long long)b * L + (row_ok ? r : L - 1);
    // ---------------- stage this row's operands (rows >= L are zero / masked)
    {
      float4 q4[4], k4[4], t4[4], v4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        q4[j] = row_ok ? *reinterpret_cast<const float4*>(a.Q + grow * a.ldq + hh * DH + 4 * j) : z;
        k4[j] = row_ok ? *reinterpret_cast<const float4*>(a.K + grow * a.ldk + hh * DH + 4 * j) : z;
        t4[j] = row_ok ? *reinterpret_cast<const float4*>(a.T + grow * a.ldt + hh * DH + 4 * j) : z;
        v4[j] = row_ok ? *reinterpret_cast<const float4*>(a.V + grow * a.ldv + hh * DH + 4 * j) : z;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        q4[j].x *= sc2; q4[j].y *= sc2; q4[j].z *= sc2; q4[j].w *= sc2;  // scores arrive in the log2 domain
        *reinterpret_cast<float4*>(sm + OFF_Q + sw128(r, 4 * j)) = q4[j];
        *reinterpret_cast<float4*>(sm + OFF_Q + sw128(r, 16 + 4 * j)) =
            make_float4(lo_of(q4[j].x), lo_of(q4[j].y), lo_of(q4[j].z), lo_of(q4[j].w));
        *reinterpret_cast<float4*>(sm + OFF_K + sw128(r, 4 * j)) = k4[j];
        *reinterpret_cast<float4*>(sm + OFF_K + sw128(r, 16 + 4 * j)) =
            make_float4(lo_of(k4[j].x), lo_of(k4[j].y), lo_of(k4[j].z), lo_of(k4[j].w));
      }
      // T^T, V^T: element (dim jj, key r) -> atom r/32, row jj, k = r%32
      const int atom = (r >> 5) * 2048, kk = r & 31;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float tv[4] = {t4[j].x, t4[j].y, t4[j].z, t4[j].w}, vv[4] = {v4[j].x, v4[j].y, v4[j].z, v4[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int o = atom + sw128(4 * j + e, kk);
          *reinterpret_cast<float*>(sm + OFF_TH + o) = tv[e];
          *reinterpret_cast<float*>(sm + OFF_TL + o) = lo_of(tv[e]);
          *reinterpret_cast<float*>(sm + OFF_VH + o) = vv[e];
          *reinterpret_cast<float*>(sm + OFF_VL + o) = lo_of(vv[e]);
        }
      }
      // marks row (uint8 -> float, exact), key mask
      uint4 mraw = make_uint4(0, 0, 0, 0);
      if (row_ok) mraw = *reinterpret_cast<const uint4*>(a.marks + grow * E);
      const uint32_t mw[4] = {mraw.x, mraw.y, mraw.z, mraw.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(sm + OFF_MK + sw128(r, 4 * j)) =
            make_float4((float)(mw[j] & 255u), (float)((mw[j] >> 8) & 255u), (float)((mw[j] >> 16) & 255u),
                        (float)(mw[j] >> 24));
      km[r] = row_ok ? (a.kmask[grow] ? INFINITY : kFill) : -INFINITY;
    }
    const float span = a.spans[grow];
    float4 res[4];  // residual row, fetched early so its latency hides behind the whole item
#pragma unroll
    for (int j = 0; j < 4; ++j)
      res[j] = (a.R && row_ok) ? *reinterpret_cast<const float4*>(a.R + grow * a.ldr + hh * DH + 4 * j)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    mbar_arrive(&bar[B_STAGED]);
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // km[] is read by the other row threads below

    // ---------------- softmax over the S row: pass 1 = max, pass 2 = exp / sum; P_un stays in registers
    mbar_wait(&bar[B_S], ph);
    tc_fence_after();
    float P[8][16];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c * 16 < NS) {
        tmem_ld16(tml + C_S + c * 16, P[c]);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float v = fminf(P[c][i], km[c * 16 + i]);
          if (a.causal && c * 16 + i > r) v = fminf(v, kFill);
          P[c][i] = v;
          m = fmaxf(m, v);
        }
      }
    }
    float l = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c * 16 < NS) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p = ex2a(P[c][i] - m);
          P[c][i] = p;
          l += p;
          hi[i] = __float_as_uint(p);
          lo[i] = __float_as_uint(lo_of(p));
        }
        tmem_st16(tml + C_S + c * 16, hi);
        tmem_st16(tml + C_PL + c * 16, lo);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(&bar[B_P]);
    const float inv_l = __frcp_rn(l);

    // ---------------- H = Hu / l ; A operand of the MLP = [H (16) | span | 1 | 0 x 6]
    mbar_wait(&bar[B_HU], ph);
    tc_fence_after();
    {
      float hu[16];
      uint32_t hi[16], lo[16];
      tmem_ld16(tml + C_HU, hu);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float hval = hu[i] * inv_l;
        hi[i] = __float_as_uint(hval);
        lo[i] = __float_as_uint(lo_of(hval));
      }
      tmem_st16(tml + C_AH, hi);
      tmem_st16(tml + C_AL, lo);
#pragma unroll
      for (int i = 0; i < 16; ++i) { hi[i] = 0u; lo[i] = 0u; }
      hi[0] = __float_as_uint(span); lo[0] = __float_as_uint(lo_of(span));
      hi[1] = __float_as_uint(1.0f);
      tmem_st16(tml + C_AH + 16, hi);
      tmem_st16(tml + C_AL + 16, lo);
      tmem_st_wait();
    }
    tc_fence_before();
    mbar_arrive(&bar[B_A]);

    // ---------------- sigmoid-dot epilogue of the MLP quarters -> per-event pre-activations
    float ls[E];
#pragma unroll
    for (int e = 0; e < E; ++e) ls[e] = 0.f;
#pragma unroll
    for (int qz = 0; qz < 4; ++qz) {
      const int buf = qz & 1;
      const uint32_t use = 2 * n + (qz >> 1);
      mbar_wait(&bar[B_Z0 + buf], use & 1);
      tc_fence_after();
#pragma unroll
      for (int ev = 0; ev < 4; ++ev) {  // one event = 16 columns
        float z[16];
        tmem_ld16(tml + C_Z + buf * 64 + ev * 16, z);
        const float* wp = g.wv + (qz * 4 + ev) * DH;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp + j);
          acc = fmaf(rcpa(1.f + ex2a(z[j + 0])), w4.x, acc);  // z holds -z*log2(e): tf.nn.sigmoid (temporal.py:290)
          acc = fmaf(rcpa(1.f + ex2a(z[j + 1])), w4.y, acc);
          acc = fmaf(rcpa(1.f + ex2a(z[j + 2])), w4.z, acc);
          acc = fmaf(rcpa(1.f + ex2a(z[j + 3])), w4.w, acc);
        }
        ls[qz * 4 + ev] = acc;
      }
      tc_fence_before();
      mbar_arrive(&bar[B_ZF0 + buf]);
    }
    // ---------------- lam_e = s_e ln(1 + exp(x / s_e))   (temporal.py:305-306, naive softplus, Q6)
    {
      uint32_t hi[16], lo[16];
      float lam[16];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float s = g.sc[e];
        const float rs = rcpa(s) * kLog2e, sl = s * 0.69314718055994531f;
        lam[e] = sl * lg2a(1.f + ex2a(ls[e] * rs));
        hi[e] = __float_as_uint(lam[e]);
        lo[e] = __float_as_uint(lo_of(lam[e]));
      }
      tmem_st16(tml + C_LH, hi);
      tmem_st16(tml + C_LL, lo);
      tmem_st_wait();
      if (a.lam && row_ok) {
        float* lp = a.lam + (((long long)hh * a.B + b) * L + r) * E;  // head-major, temporal.py:413-416
#pragma unroll
        for (int e = 0; e < E; e += 4) *reinterpret_cast<float4*>(lp + e) = make_float4(lam[e], lam[e + 1], lam[e + 2], lam[e + 3]);
      }
    }
    tc_fence_before();
    mbar_arrive(&bar[B_L]);

    // ---------------- gate: G o P  (set_diag for BiMAU), temporal.py:438-441; hi in place of G, lo beside it
    mbar_wait(&bar[B_G], ph);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c * 16 < NS) {
        float gg[16];
        uint32_t hi[16], lo[16];
        tmem_ld16(tml + C_G + c * 16, gg);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float gv = (a.diag_one && c * 16 + i == r) ? 1.f : gg[i];
          const float gp = gv * P[c][i];
          hi[i] = __float_as_uint(gp);
          lo[i] = __float_as_uint(lo_of(gp));
        }
        tmem_st16(tml + C_S + c * 16, hi);
        tmem_st16(tml + C_PL + c * 16, lo);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(&bar[B_GP]);

    // ---------------- O = Ou / l + residual   (temporal.py:385,447)
    mbar_wait(&bar[B_O], ph);
    tc_fence_after();
    {
      float o[16];
      tmem_ld16(tml + C_O, o);
      if (row_ok) {
        float* op = a.O + grow * a.ldo + hh * DH;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(op + 4 * j) =
              make_float4(o[4 * j] * inv_l + res[j].x, o[4 * j + 1] * inv_l + res[j].y, o[4 * j + 2] * inv_l + res[j].z,
                          o[4 * j + 3] * inv_l + res[j].w);
      }
    }
    tc_fence_before();  // TMEM reads of this item are done before the next item's operands / MMAs
  }
}

__global__ void __launch_bounds__(320, 1) attention_tc_kernel(AttnArgs a, int num_items) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * B_COUNT);
  float* wv = reinterpret_cast<float*>(sm + OFF_WV);
  float* sc = reinterpret_cast<float*>(sm + OFF_SC);
  const int tid = threadIdx.x, warp = tid >> 5;
  if ((smem_u32(sm) & 1023u) != 0) __trap();  // the swizzled tiles need a 1024-byte aligned base

  if (tid == 0) {
    for (int g = 0; g < 2; ++g)
      for (int i = 0; i < B_COUNT; ++i) {
        const bool simt = (i == B_STAGED || i == B_P || i == B_A || i == B_ZF0 || i == B_ZF1 || i == B_L || i == B_GP);
        mbar_init(&bars[g * B_COUNT + i], simt ? 128 : 1);
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ---- static operands: W1'^T (hi, lo), int_weight, exp(scaling); zero the operand tiles once (padding stays 0)
  for (int i = tid; i < OFF_WH / 16; i += blockDim.x) reinterpret_cast<float4*>(sm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < NC * 32; i += blockDim.x) {
    const int n = i >> 5, k = i & 31;
    float v = 0.f;
    if (k <= DH) v = -kLog2e * a.int_w[k * NC + n];     // rows 0..15: W1, row 16: interval weight
    else if (k == DH + 1) v = -kLog2e * a.int_b[n];      // row 17: bias (A column 17 is the constant 1)
    const int o = sw128(n, k);
    *reinterpret_cast<float*>(sm + OFF_WH + o) = v;
    *reinterpret_cast<float*>(sm + OFF_WL + o) = lo_of(v);
  }
  for (int i = tid; i < NC; i += blockDim.x) wv[i] = a.int_weight[i];
  for (int i = tid; i < E; i += blockDim.x) sc[i] = expf(a.int_scaling[i]);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;

  const int grp = warp < 8 ? (warp >> 2) : (warp - 8);  // warps 0-3 / 4-7: row threads; warp 8 / 9: MMA issuers
  GroupCtx g;
  g.sm = sm + grp * GROUP_BYTES;
  g.smw = sm + OFF_WH;
  g.bar = bars + grp * B_COUNT;
  g.km = reinterpret_cast<float*>(sm + OFF_KM) + grp * 128;
  g.wv = wv;
  g.sc = sc;
  g.tm = tm + grp * C_GROUP;
  g.first_item = blockIdx.x * 2 + grp;
  g.item_stride = gridDim.x * 2;
  if (warp >= 8) {
    if ((tid & 31) == 0) run_mma(a, g, num_items);
  } else {
    run_rows(a, g, num_items, tid & 127, 1 + grp);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512));
  }
}

}  // namespace atc

// 0 = launched, 1 = shape not covered (caller falls back), <0 = error
int launch_attention_tc(const AttnArgs& a, cudaStream_t st) {
  using namespace atc;
  if (a.d / a.h != DH || a.E != E || a.L > 112 || a.L < 8) return 1;  // S (<= 112 cols) must not reach P_lo
  if ((a.ldq | a.ldk | a.ldv | a.ldt | a.ldo) % 4 || (a.R && a.ldr % 4)) return 1;
  if (reinterpret_cast<uintptr_t>(a.marks) & 15) return 1;
  static int num_sms = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  const long long items = (long long)a.B * a.h;
  if (items == 0) return 0;
  auto kern = attention_tc_kernel;
  EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const long long pairs = (items + 1) / 2;
  const int grid = pairs < num_sms ? (int)pairs : num_sms;
  kern<<<grid, 320, SMEM_BYTES, st>>>(a, (int)items);
  EDGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace edgl
