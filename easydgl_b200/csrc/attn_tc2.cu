// attn_tc2.cu - the fused self-modulating attention core (temporal.py:281-315, 345-385, 412-447) on tcgen05 / TMEM / TMA
// for dh = 16, E = 16, L <= 112 (BASELINE configs C1-C3): the default attention kernel of the pipeline.
//
//   S = Q K^T, P = softmax_k(mask(S / sqrt(dh))), H = P T, Z = sigmoid([H, span] W1 + b1), x_e = Z_e . w_e,
//   lam_e = s_e log(1 + exp(x_e / s_e)), G = lam M^T (set_diag(G, 1) for BiMAU), O = (G o P) V + residual
//
// Arithmetic: the SCALED 3xFP16 split of attn_f16.cu / gemm_f16.cu (x * 2^k = hi + lo, hi = top 11 significant bits,
// D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi in fp32), i.e. fp32-level accuracy on `tcgen05.mma kind::f16`.
//
// One persistent CTA per SM (512 threads) works on TWO (sequence, head) items at a time ("slots" of 256 threads with
// 256 TMEM columns each), so the MUFU-bound phases of one item (256 sigmoids and 100 exps per row) overlap the
// tensor-core / TMA latencies of the other.  Inside a slot the TMEM lane quarter of a warp (warp % 4) fixes its 32
// query rows and the two warps of a quarter split every tile by COLUMNS (key halves in the softmax / gate, event
// halves in the MLP, dim halves in the H / O epilogues), so a thread keeps 56 probabilities in registers from the
// softmax to the gate.  Per item and slot (TMEM columns relative to the slot base):
//
//   TMA    Q, K, V, T tiles [L x 16] fp32 straight from the projection buffer -> shared (64B-swizzled boxes),
//          issued one item ahead; converted to fp16 (hi | lo) operand tiles while the previous item's last MMA runs
//   S      SS  A = Q (row scale)          B = K (tile scale)            -> [0, NS)
//   P      softmax in the log2 domain; P' = 2^14 p stays in registers, fp16x2 (hi, lo) copies -> TMEM [0, NS)
//   H      TS  A = P'_hi (TMEM)           B = [T_hi | T_lo] (MN-major, N = 32) -> [112, 144);  A = P'_lo -> [144, 176)
//   Z      SS  A = H (hi | lo)            B = W1'^T (N = 256, static)   -> [0, 256)
//   lam    sigmoid-dot per event (8 events per thread; the per-column constants are kernel parameters, i.e. constant-
//          bank operands of the FFMAs: no loads in the loop), softplus; one static scale from a bound on lam
//   G      SS  A = lam (hi | lo)          B = marks (exact)             -> [128, 128 + NS)
//   gate   G o P' (set_diag) -> fp16x2 (hi, lo)                          -> TMEM [0, NS)
//   O      TS  A = G o P'                 B = [V_hi | V_lo] (MN-major)  -> [128, 160) / [160, 192);  + residual -> global
//
// All operand tiles are SWIZZLE_64B with rows of 64 B = 16 hi | 16 lo halves: K-major for Q, K, H, lam, marks and W1
// (the layout gemm_f16.cu uses), MN-major for V and T (row = key = the K index of the product, the 32 halves of a row
// are the N index), so all four projections are converted by the same row-wise routine.  One mbarrier per slot tracks MMA completion (tcgen05.commit), one the TMA
// transaction; the 256 threads of a slot meet on a named barrier before their elected thread issues the next MMAs.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <math.h>

#include <cmath>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "tc_common.cuh"

namespace edgl {
namespace at2 {

using namespace tcc;

constexpr int DH = 16, E = 16, NC = DH * E;
constexpr int NTHR = 512, SLOT_THR = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kFill = -4294967296.0f;  // float(-2**32+1), temporal.py:358,425

// ---- constant pack: the B operand of the MLP product, built once per edgl_commit
constexpr int PK_W1 = 0;                         // W1'^T [256 n][hi 16 | lo 16] fp16, SWIZZLE_64B      16 KB
constexpr int PK_BYTES = PK_W1 + NC * 64;

// Per-column constants of the intensity MLP as a KERNEL PARAMETER: with the loops unrolled every access is a
// constant-bank operand of an FFMA, so the 256 sigmoids per row need no loads at all.
struct MlpConst {
  float b[NC];     // -log2(e) * int_b
  float wsp[NC];   // -log2(e) * (interval row of int_w)
  float w[NC];     // int_weight, [event][dim] = the column order of the MLP
  float rs[E];     // log2(e) / s_e,  s_e = exp(scaling_e)   (temporal.py:302)
  float sl2[E];    // s_e * ln 2
  float isw;       // 1 / scale(W1)
  float sl, isl;   // scale of lam (a power of two from a static bound on lam) and its inverse
  float pad;
};

// ---- TMEM columns of a slot
constexpr uint32_t C_S = 0, C_H1 = 112, C_H2 = 144, C_Z = 0, C_G = 128, C_O1 = 128, C_O2 = 160, C_SLOT = 256;
constexpr uint32_t kBMajorMN = 1u << 16;  // instruction descriptor: B operand is MN-major

template <int NS>
struct Lay {
  static constexpr int NSH = NS / 2;              // keys per thread
  static constexpr int A16 = 128 * 64;            // [128 rows][hi 16 | lo 16]
  static constexpr int KT = ((NS * 64 + 1023) / 1024) * 1024;  // K / marks / T / V tiles: [NS rows][64 B]
  static constexpr int XT = KT;
  static constexpr int RAWT = ((NS * 64 + 1023) / 1024) * 1024;  // one raw fp32 tile [NS][16], 1 KB aligned
  // per-slot offsets
  static constexpr int O_A16A = 0;                // Q, later lam
  static constexpr int O_A16B = O_A16A + A16;     // H
  static constexpr int O_BK = O_A16B + A16;
  static constexpr int O_BM = O_BK + KT;
  static constexpr int O_BT = O_BM + KT;
  static constexpr int O_BV = O_BT + XT;          // two buffers
  static constexpr int O_RAW = ((O_BV + 2 * XT + 1023) / 1024) * 1024;  // Q, K, V, T raw
  static constexpr int O_KM = O_RAW + 4 * RAWT;   // float [128] key min-mask
  static constexpr int O_ISQ = O_KM + 512;        // float [128] 1 / scale(Q row)
  static constexpr int O_X = O_ISQ + 512;         // float [3][2][128] exchanges: softmax max, softmax sum, (spare)
  static constexpr int O_WRED = O_X + 3072;       // float [8 warps][4] tile maxima (K, T, V, marks sum)
  static constexpr int SLOT = ((O_WRED + 128 + 1023) / 1024) * 1024;
  // CTA-wide
  static constexpr int O_PACK = 2 * SLOT;
  static constexpr int O_BAR = O_PACK + ((PK_BYTES + 127) / 128) * 128;  // mbarriers [2 slots][2] + tmem slot
  static constexpr int BYTES = O_BAR + 64 + 1024 /* base alignment slack */;
};

// ------------------------------------------------------------------------------------------ small device helpers
__device__ __forceinline__ uint64_t desc64(uint32_t saddr) {  // K-major SWIZZLE_64B, 8-row atoms 512 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_f16(int n) {  // D = f32, A = B = f16, K-major, M = 128, N = n
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
      "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float* f) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void slot_sync(int slot) { asm volatile("bar.sync %0, 256;" ::"r"(slot + 1) : "memory"); }

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ uint32_t pack_h2(float e0, float e1) {  // e0 -> low half (lower k index)
  const __half2 h = __floats2half2_rn(e0, e1);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// two (already scaled) fp32 values -> packed hi pair, packed lo pair
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const float h0 = __uint_as_float(__float_as_uint(x0) & 0xffffe000u);
  const float h1 = __uint_as_float(__float_as_uint(x1) & 0xffffe000u);
  hi = pack_h2(h0, h1);
  lo = pack_h2(x0 - h0, x1 - h1);
}
// m >= 0: s = 2^k with m*s in [2^14, 2^15), is = 1/s (both exact; exponent clamped so neither is denormal)
__device__ __forceinline__ void pow2_scale(float m, float& s, float& is) {
  int e = (int)((__float_as_uint(m) >> 23) & 0xffu);
  e = min(max(e, 15), 239);
  s = __uint_as_float((uint32_t)(268 - e) << 23);
  is = __uint_as_float((uint32_t)(e - 14) << 23);
}
// byte offset of 16-byte chunk c4 (0..3) of row r in a K-major SWIZZLE_64B tile
__device__ __forceinline__ uint32_t sw64(int r, int c4) {
  return (uint32_t)((r >> 3) * 512 + (r & 7) * 64 + (((c4 ^ (r >> 1)) & 3) << 4));
}

// ------------------------------------------------------------------------------------------ constant pack
__global__ void __launch_bounds__(256) tc2_pack_kernel(const float* __restrict__ int_w, float sw,
                                                       unsigned char* __restrict__ pack) {
  // W1'^T: row n = MLP column, k = input dim; hi in halves 0..15, lo in halves 16..31 of the 64-byte row
  for (int i = threadIdx.x; i < NC * 2; i += blockDim.x) {
    const int n = i >> 1, half8 = i & 1;  // 8 dims per chunk
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k0 = half8 * 8 + 2 * j;
      split2(-kLog2e * int_w[(size_t)k0 * NC + n] * sw, -kLog2e * int_w[(size_t)(k0 + 1) * NC + n] * sw, hi[j], lo[j]);
    }
    *reinterpret_cast<uint4*>(pack + PK_W1 + sw64(n, half8)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(pack + PK_W1 + sw64(n, 2 + half8)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ------------------------------------------------------------------------------------------ the kernel
// Shared memory is addressed through explicit .shared instructions on 32-bit addresses: the operand tiles are carved
// out of a dynamically aligned base, which hides the address space from the compiler (generic LD / ST with a
// descriptor shuffle per access otherwise).
__device__ __forceinline__ float4 lds4(uint32_t a) {  // data other threads write between barriers
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float lds1(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 ldc4(uint32_t a) {  // constants of the pack (never written after the prologue)
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 ldc2(uint32_t a) {
  float2 v;
  asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts1(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts4(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts4u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts_h2(uint32_t a_lo, uint32_t a_hi, uint32_t packed) {  // two halves to two addresses
  asm volatile(
      "{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tst.shared.b16 [%0], l;\n\tst.shared.b16 [%1], h;\n\t}\n" ::"r"(a_lo),
      "r"(a_hi), "r"(packed)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma_commit_s(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_tile(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

struct SlotCtx {
  uint32_t sm;        // this slot's shared memory (shared-space address)
  uint32_t pack;      // the constant pack (shared-space address)
  uint32_t mma_bar;   // tcgen05.commit target
  uint32_t raw_bar;   // TMA transaction barrier
  uint32_t tm;        // TMEM base of the slot
  int slot, ts, q, ch, r, lane;
};

template <int NS, bool DBG>
__device__ __forceinline__ void dbg_put(const AttnArgs& a, int item, int phase, int r, int col, float v) {
  if constexpr (DBG) {
    if (a.dbg) a.dbg[(((size_t)item * 8 + phase) * 128 + r) * 256 + col] = v;
  }
}

__device__ __forceinline__ float absmax4(float m, const float4& v) {
  return fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
}
// 16 fp32 values (4 float4) * s -> a 64-byte (hi 16 | lo 16) row of a SWIZZLE_64B tile
__device__ __forceinline__ void put_row16(uint32_t tile, int row, const float4 (&v)[4], float s) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    split2(v[j].x * s, v[j].y * s, hi[2 * j], lo[2 * j]);
    split2(v[j].z * s, v[j].w * s, hi[2 * j + 1], lo[2 * j + 1]);
  }
  sts4u(tile + sw64(row, 0), hi[0], hi[1], hi[2], hi[3]);
  sts4u(tile + sw64(row, 1), hi[4], hi[5], hi[6], hi[7]);
  sts4u(tile + sw64(row, 2), lo[0], lo[1], lo[2], lo[3]);
  sts4u(tile + sw64(row, 3), lo[4], lo[5], lo[6], lo[7]);
}
// Convert the raw fp32 tiles of `item` (landed by TMA) into the fp16 operand tiles.  vbuf = V buffer to fill.
// marks row / key mask byte of row t of `item` (threads of group 0): requested early by the caller, consumed by
// convert_item
struct KeyMeta {
  uint4 mraw;
  float kv;
};
__device__ __forceinline__ KeyMeta load_key_meta(const AttnArgs& a, const SlotCtx& c, int item) {
  KeyMeta k;
  k.mraw = make_uint4(0u, 0u, 0u, 0u);
  k.kv = -INFINITY;
  const int t = c.ts & 127;
  if ((c.ts >> 7) == 0 && t < a.L) {
    const long long row = (long long)(item / a.h) * a.L + t;
    k.mraw = __ldg(reinterpret_cast<const uint4*>(a.marks + row * E));
    k.kv = a.kmask[row] ? INFINITY : -INFINITY;
  }
  return k;
}

template <int NS>
__device__ __forceinline__ void convert_item(const AttnArgs& a, const SlotCtx& c, int item, int vbuf, const KeyMeta& meta) {
  using LY = Lay<NS>;
  const int L = a.L;
  const int t = c.ts & 127, grp = c.ts >> 7;  // grp 0: Q, K, marks of row t;  grp 1: T, V of key t
  const uint32_t sm = c.sm;
  const uint32_t rsw = (uint32_t)((t >> 1) & 3);
  const uint32_t rawrow = sm + LY::O_RAW + t * 64;
  auto raw4 = [&](int which, int j) {  // 16-byte chunk j of row t of raw tile `which` (TMA SWIZZLE_64B)
    return lds4(rawrow + which * LY::RAWT + ((j ^ rsw) << 4));
  };
  const bool in_l = t < L;
  float4 x0[4], x1[4];  // grp 0: K row (x0); grp 1: T row (x0), V row (x1)
#pragma unroll
  for (int j = 0; j < 4; ++j) x0[j] = x1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float m0 = 0.f, m1 = 0.f, msum = 0.f;
  if (grp == 0) {
    const uint4 mraw = meta.mraw;
    const float kv = meta.kv;
    // ---- Q row: per-row scale
    float4 qv[4];
    float qm = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      qv[j] = in_l ? raw4(0, j) : make_float4(0.f, 0.f, 0.f, 0.f);
      qm = absmax4(qm, qv[j]);
    }
    float sq, isq;
    pow2_scale(qm, sq, isq);
    sts1(sm + LY::O_ISQ + t * 4, isq);
    put_row16(sm + LY::O_A16A, t, qv, sq);
    // ---- K row (raw values kept until the tile maximum is known)
    if (t < NS) {
      if (in_l) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          x0[j] = raw4(1, j);
          m0 = absmax4(m0, x0[j]);
        }
      }
      sts1(sm + LY::O_KM + t * 4, kv);  // min-mask: +inf real key, -inf masked id or beyond L
      // bytes -> fp16 exactly through 1024 + b (0x6400 | b) - 1024; two events per word, natural order
      const uint32_t mw[4] = {mraw.x, mraw.y, mraw.z, mraw.w};
      uint32_t mh[8];
      unsigned int s8 = 0;
      const __half2 k1024 = __floats2half2_rn(1024.f, 1024.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t lo2 = __byte_perm(mw[j], 0x64646464u, 0x4140u), hi2 = __byte_perm(mw[j], 0x64646464u, 0x4342u);
        const __half2 r0 = __hsub2(*reinterpret_cast<const __half2*>(&lo2), k1024);
        const __half2 r1 = __hsub2(*reinterpret_cast<const __half2*>(&hi2), k1024);
        mh[2 * j] = *reinterpret_cast<const uint32_t*>(&r0);
        mh[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&r1);
        s8 += __vsadu4(mw[j], 0u);  // sum of the four bytes
      }
      msum = (float)s8;
      sts4u(sm + LY::O_BM + sw64(t, 0), mh[0], mh[1], mh[2], mh[3]);
      sts4u(sm + LY::O_BM + sw64(t, 1), mh[4], mh[5], mh[6], mh[7]);
    }
  } else if (in_l) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x0[j] = raw4(3, j);  // T
      x1[j] = raw4(2, j);  // V
      m0 = absmax4(m0, x0[j]);
      m1 = absmax4(m1, x1[j]);
    }
  }
  // ---- tile maxima: warp reduce, one row of wred per warp, everybody combines after the barrier
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    msum = fmaxf(msum, __shfl_xor_sync(0xffffffffu, msum, o));
  }
  if (c.lane == 0) {
    const uint32_t wr = sm + LY::O_WRED + (c.ts >> 5) * 16;  // columns: 0 = K, 1 = T, 2 = V, 3 = marks sum
    if (grp == 0) sts4(wr, m0, 0.f, 0.f, msum); else sts4(wr, 0.f, m0, m1, 0.f);
  }
  slot_sync(c.slot);
  float4 mx = lds4(sm + LY::O_WRED);
#pragma unroll
  for (int w = 1; w < 8; ++w) {
    const float4 v = lds4(sm + LY::O_WRED + w * 16);
    mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
  }
  if (in_l) {  // rows / keys >= L were zeroed once and are never written
    if (grp == 0) {
      float sk, isk;
      pow2_scale(mx.x, sk, isk);
      put_row16(sm + LY::O_BK, t, x0, sk);
    } else {
      float st, ist, sv, isv;
      pow2_scale(mx.y, st, ist);
      pow2_scale(mx.z, sv, isv);
      // MN-major B tiles: row = key, 64 bytes = (hi 16 | lo 16) dims - the same row format as K
      put_row16(sm + LY::O_BT, t, x0, st);
      put_row16(sm + LY::O_BV + vbuf * LY::XT, t, x1, sv);
    }
  }
  fence_proxy_async();  // generic-proxy writes -> visible to the tensor core
}

// The sigmoid-dot epilogue of the MLP for the 8 events (128 columns) of column half CH: lam[i] = intensity of event
// 8 CH + i.  Fully unrolled so that every per-column constant is a constant-bank operand.
template <int CH, int NS, bool DBG>
__device__ __forceinline__ void mlp_events(const AttnArgs& a, const MlpConst& mc, uint32_t tml, float span, float zscale,
                                           int item, int r, float (&lam)[8]) {
  // 16 chunks of 8 accumulator columns (two per event); chunk c+1 is requested before chunk c is processed, so the
  // TMEM round trip hides behind 8 sigmoids
  float zb[2][8];
  tmem_ld8_nowait(tml + C_Z + CH * 128, zb[0]);
  float acc = 0.f;
#pragma unroll
  for (int cidx = 0; cidx < 16; ++cidx) {
    tmem_ld_wait();
    if (cidx + 1 < 16) tmem_ld8_nowait(tml + C_Z + CH * 128 + (cidx + 1) * 8, zb[(cidx + 1) & 1]);
    const int ev = CH * 8 + (cidx >> 1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = CH * 128 + cidx * 8 + j;
      // zz = -log2(e) * ([H, span] W1 + b1): sigmoid = 1 / (1 + 2^zz)   (tf.nn.sigmoid, temporal.py:290)
      const float zz = fmaf(span, mc.wsp[col], fmaf(zb[cidx & 1][j], zscale, mc.b[col]));
      acc = fmaf(rcpa(1.f + ex2a(zz)), mc.w[col], acc);
    }
    if (cidx & 1) {
      // lam_e = s_e log(1 + exp(x / s_e))  (temporal.py:305-306: the naive softplus, overflows like the reference, Q6)
      lam[cidx >> 1] = mc.sl2[ev] * lg2a(1.f + ex2a(acc * mc.rs[ev]));
      if constexpr (DBG) dbg_put<NS, DBG>(a, item, 3, r, ev, acc);
      acc = 0.f;
    }
  }
}

template <int NS, bool DBG, bool PROF>
__device__ __forceinline__ void run_slot(const AttnArgs& a, const MlpConst& mc, const SlotCtx& c, const CUtensorMap* mq,
                                         const CUtensorMap* mk, const CUtensorMap* mv, const CUtensorMap* mt,
                                         int num_items, int first_item, int stride, bool alternate) {
  using LY = Lay<NS>;
  constexpr int NSH = LY::NSH;
  const int L = a.L;
  const uint32_t sm = c.sm;
  const int r = c.r, ch = c.ch;
  const uint32_t tml = c.tm + ((uint32_t)(c.q * 32) << 16);  // this warp's TMEM lane quarter
  const int k0 = ch * NSH;                                    // first key of this thread
  const bool row_ok = r < L;
  const float sc2 = kLog2e / sqrtf((float)DH);                // temporal.py:355,422; scores in the log2 domain
  const uint32_t xmax = sm + LY::O_X, xsum = xmax + 1024;     // [2][128] floats each
  const uint32_t kmrow = sm + LY::O_KM + k0 * 4;
  const float sl = mc.sl, isl = mc.isl;                       // static scale of lam
  // set_diag (BiMAU): this thread's diagonal key slot, and whether this WARP has a diagonal inside an 8-key chunk
  const int jd = a.diag_one ? r - k0 : -1000;
  const int jd_lo = a.diag_one ? c.q * 32 - k0 : -1000, jd_hi = jd_lo + 31;
  uint32_t mph = 0, rph = 0;
  // per-phase cycle counters of the slot's first thread (PROF builds only: tools/attn_selftest ... prof)
  long long tacc[14];
  long long tprev = 0;
  if constexpr (PROF) {
#pragma unroll
    for (int i = 0; i < 14; ++i) tacc[i] = 0;
    tprev = clock64();
  }
#define AT2_TICK(i)                                  \
  if constexpr (PROF) {                              \
    const long long tnow_ = clock64();               \
    tacc[i] += tnow_ - tprev;                        \
    tprev = tnow_;                                   \
  }

  auto issue_tma = [&](int item) {
    const int b = item / a.h, hh = item % a.h;
    const uint32_t raw = sm + LY::O_RAW;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(c.raw_bar), "r"((uint32_t)(4 * L * 64))
                 : "memory");
    tma_tile(mq, c.raw_bar, raw + 0 * LY::RAWT, hh * DH, b * L);
    tma_tile(mk, c.raw_bar, raw + 1 * LY::RAWT, hh * DH, b * L);
    tma_tile(mv, c.raw_bar, raw + 2 * LY::RAWT, hh * DH, b * L);
    tma_tile(mt, c.raw_bar, raw + 3 * LY::RAWT, hh * DH, b * L);
  };
  auto issue_s = [&]() {
    const uint32_t aq = sm + LY::O_A16A, bk = sm + LY::O_BK;
    constexpr uint32_t id = idesc_f16(NS);
    mma_ss(c.tm + C_S, desc64(aq + 32), desc64(bk), id, 0);      // Q_lo K_hi
    mma_ss(c.tm + C_S, desc64(aq), desc64(bk + 32), id, 1);      // Q_hi K_lo
    mma_ss(c.tm + C_S, desc64(aq), desc64(bk), id, 1);           // Q_hi K_hi
    umma_commit_s(c.mma_bar);
  };
  // A = (hi, lo) fp16x2 copies of a [128 x NS] matrix in TMEM [0, NS); B = [X_hi | X_lo] rows of 16 keys per k-step,
  // MN-major: d1 = A_hi [X_hi | X_lo], d2 = A_lo [X_hi | X_lo] (its second half, lo * lo, is not read)
  auto issue_pv = [&](uint32_t d1, uint32_t d2, uint32_t xs) {
    constexpr uint32_t id32 = idesc_f16(32) | kBMajorMN;
#pragma unroll
    for (int ks = 0; ks < NS / 16; ++ks) {
      const uint64_t bd = desc64(xs + ks * 1024);
      mma_ts(d1, c.tm + C_S + ks * 8, bd, id32, ks != 0);
      mma_ts(d2, c.tm + C_S + NS / 2 + ks * 8, bd, id32, ks != 0);
    }
    umma_commit_s(c.mma_bar);
  };

  int item = first_item;
  if (item >= num_items) return;
  // ---- prologue: first item's tiles
  if (c.ts == 0) issue_tma(item);
  {
    const KeyMeta meta0 = load_key_meta(a, c, item);
    mbar_wait_s(c.raw_bar, rph); rph ^= 1;
    convert_item<NS>(a, c, item, 0, meta0);
  }
  slot_sync(c.slot);
  if (c.ts == 0) {
    if (item + stride < num_items) issue_tma(item + stride);
    tc_fence_after();
    issue_s();
  }

  for (int n = 0; item < num_items; ++n, item += stride) {
    const int b = item / a.h, hh = item % a.h;
    const long long grow = (long long)b * L + (row_ok ? r : L - 1);
    const int next = item + stride;
    const bool has_next = next < num_items;
    const float span = __ldg(a.spans + grow);
    // tile scales of this item (written by convert_item; stable until the next conversion)
    float4 mx = lds4(sm + LY::O_WRED);
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 v = lds4(sm + LY::O_WRED + w * 16);
      mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
    }
    float sk, isk, st, ist, sv, isv;
    pow2_scale(mx.x, sk, isk);
    pow2_scale(mx.y, st, ist);
    pow2_scale(mx.z, sv, isv);
    const float mxm = mx.w;
    const float isq = lds1(sm + LY::O_ISQ + r * 4);

    AT2_TICK(13)  // loop head: scales
    // ================================================================ softmax
    // The masks act on the RAW accumulator (min with +inf / -inf), the positive row factor ar = log2e / sqrt(dh) /
    // (scale(Q row) scale(K)) is applied inside the exponential.  A masked key's score is the finite -2^32+1 in the
    // reference (temporal.py:358,425), which only matters when EVERY key of a row is masked: then all L scores are
    // equal and the attention is uniform over the L keys (Q8) - handled below.
    mbar_wait_s(c.mma_bar, mph); mph ^= 1;
    tc_fence_after();
    AT2_TICK(0)  // wait S
    float P[NSH];
#pragma unroll
    for (int j = 0; j < NSH; j += 8) tmem_ld8_nowait(tml + C_S + k0 + j, P + j);
    tmem_ld_wait();
    {
      const float ar = sc2 * isq * isk;
      float m = -INFINITY;
      if (a.causal) {
        const int jr = r - k0;  // keys j > jr are in the future of row r  (temporal.py:362-367)
#pragma unroll
        for (int j = 0; j < NSH; j += 4) {
          const float4 kmv = lds4(kmrow + j * 4);
          const float kk[4] = {kmv.x, kmv.y, kmv.z, kmv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float t = fminf(P[j + i], kk[i]);
            t = (j + i > jr) ? -INFINITY : t;
            P[j + i] = t;
            m = fmaxf(m, t);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < NSH; j += 4) {
          const float4 kmv = lds4(kmrow + j * 4);
          const float kk[4] = {kmv.x, kmv.y, kmv.z, kmv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float t = fminf(P[j + i], kk[i]);
            P[j + i] = t;
            m = fmaxf(m, t);
          }
        }
      }
      sts1(xmax + (ch * 128 + r) * 4, m);
      slot_sync(c.slot);  // also: every S column has been read before P overwrites the region
      AT2_TICK(1)  // softmax pass 1 + barrier
      m = fmaxf(lds1(xmax + r * 4), lds1(xmax + (128 + r) * 4));
      if constexpr (DBG) {
#pragma unroll
        for (int j = 0; j < NSH; ++j) dbg_put<NS, DBG>(a, item, 0, r, k0 + j, P[j] * ar);
      }
      // P' = 2^14 exp2((t - m) ar): the unnormalised probabilities, scaled into the fp16 range
      float lsum = 0.f;
      if (m == -INFINITY) {
        // every key of the row is masked: uniform over the L keys of the sequence
#pragma unroll
        for (int j = 0; j < NSH; ++j) P[j] = (k0 + j < L) ? 16384.f : 0.f;
#pragma unroll
        for (int j = 0; j < NSH; ++j) lsum += P[j];
      } else {
        const float moff = fmaf(-m, ar, 14.f);
#pragma unroll
        for (int j = 0; j < NSH; ++j) {
          const float p = ex2a(fmaf(P[j], ar, moff));
          P[j] = p;
          lsum += p;
        }
      }
#pragma unroll
      for (int j = 0; j < NSH; j += 8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split2(P[j + 2 * i], P[j + 2 * i + 1], hi[i], lo[i]);
        tmem_st4(tml + C_S + (k0 + j) / 2, hi);
        tmem_st4(tml + C_S + NS / 2 + (k0 + j) / 2, lo);
      }
      sts1(xsum + (ch * 128 + r) * 4, lsum);
      tmem_st_wait();
    }
    tc_fence_before();
    slot_sync(c.slot);
    if (c.ts == 0) {
      tc_fence_after();
      issue_pv(c.tm + C_H1, c.tm + C_H2, sm + LY::O_BT);
    }
    AT2_TICK(2)  // softmax pass 2, P stores, barrier, issue
    // ================================================================ H = P T / l  -> A operand of the MLP
    const float linv = __frcp_rn(lds1(xsum + r * 4) + lds1(xsum + (128 + r) * 4));  // 1 / (2^14 l)
    if constexpr (DBG) {
#pragma unroll
      for (int j = 0; j < NSH; ++j) dbg_put<NS, DBG>(a, item, 1, r, k0 + j, P[j] * linv);
    }
    mbar_wait_s(c.mma_bar, mph); mph ^= 1;
    tc_fence_after();
    AT2_TICK(3)  // wait P T
    {
      float h1[8], h2[8], h3[8];
      tmem_ld8_nowait(tml + C_H1 + ch * 8, h1);       // P_hi T_hi
      tmem_ld8_nowait(tml + C_H1 + 16 + ch * 8, h2);  // P_hi T_lo
      tmem_ld8_nowait(tml + C_H2 + ch * 8, h3);       // P_lo T_hi
      tmem_ld_wait();
      uint32_t hi[4], lo[4];
      float hv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) hv[i] = (h3[i] + h2[i] + h1[i]) * linv;  // = H * scale(T), |.| < 2^15
#pragma unroll
      for (int i = 0; i < 4; ++i) split2(hv[2 * i], hv[2 * i + 1], hi[i], lo[i]);
      sts4u(sm + LY::O_A16B + sw64(r, ch), hi[0], hi[1], hi[2], hi[3]);
      sts4u(sm + LY::O_A16B + sw64(r, 2 + ch), lo[0], lo[1], lo[2], lo[3]);
      if constexpr (DBG) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dbg_put<NS, DBG>(a, item, 2, r, ch * 8 + i, hv[i] * ist);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    slot_sync(c.slot);
    if (c.ts == 0) {
      tc_fence_after();
      const uint32_t ah = sm + LY::O_A16B, bw = c.pack + PK_W1;
      constexpr uint32_t id = idesc_f16(NC);
      mma_ss(c.tm + C_Z, desc64(ah + 32), desc64(bw), id, 0);
      mma_ss(c.tm + C_Z, desc64(ah), desc64(bw + 32), id, 1);
      mma_ss(c.tm + C_Z, desc64(ah), desc64(bw), id, 1);
      umma_commit_s(c.mma_bar);
    }
    AT2_TICK(4)  // H conversion, barrier, issue
    // ================================================================ intensity MLP epilogue: 8 events per thread
    // the next item's marks row / key mask byte: requested here, consumed by convert_item after the gate
    KeyMeta meta_next;
    meta_next.mraw = make_uint4(0u, 0u, 0u, 0u);
    meta_next.kv = -INFINITY;
    if (has_next) meta_next = load_key_meta(a, c, next);
    const float zscale = ist * mc.isw;  // accumulator -> -z log2(e)
    mbar_wait_s(c.mma_bar, mph); mph ^= 1;
    tc_fence_after();
    AT2_TICK(5)  // wait MLP
    // The two slots take turns in this MUFU-bound phase (slot 0 first, then strictly alternating): left alone, two
    // identical chains fall into lockstep and share the MUFU during the sigmoids and the issue slots during
    // everything else; with the hand-over one slot's sigmoids run under the other slot's ALU phases.  Named barrier 3
    // = "slot 0 has finished its sigmoids", 4 = "slot 1 has"; the item counts of the slots differ by at most one
    // (slot 0 has the extra), so every sync has its arrive and at most one arrive stays unmatched at the end.
    if (alternate) {
      if (c.slot == 0) {
        if (n > 0) asm volatile("bar.sync 4, 512;" ::: "memory");
      } else {
        asm volatile("bar.sync 3, 512;" ::: "memory");
      }
    }
    float lam[8];
    if (ch == 0) mlp_events<0, NS, DBG>(a, mc, tml, span, zscale, item, r, lam);
    else mlp_events<1, NS, DBG>(a, mc, tml, span, zscale, item, r, lam);
    if (alternate) {
      if (c.slot == 0) asm volatile("bar.arrive 3, 512;" ::: "memory");
      else asm volatile("bar.arrive 4, 512;" ::: "memory");
    }
    AT2_TICK(6)  // sigmoids
    if (a.lam && row_ok) {
      float* lp = a.lam + (((long long)hh * a.B + b) * L + r) * E + ch * 8;  // head-major, temporal.py:413-416
      *reinterpret_cast<float4*>(lp) = make_float4(lam[0], lam[1], lam[2], lam[3]);
      *reinterpret_cast<float4*>(lp + 4) = make_float4(lam[4], lam[5], lam[6], lam[7]);
    }
    {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split2(lam[2 * i] * sl, lam[2 * i + 1] * sl, hi[i], lo[i]);
      sts4u(sm + LY::O_A16A + sw64(r, ch), hi[0], hi[1], hi[2], hi[3]);
      sts4u(sm + LY::O_A16A + sw64(r, 2 + ch), lo[0], lo[1], lo[2], lo[3]);
      if constexpr (DBG) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dbg_put<NS, DBG>(a, item, 4, r, ch * 8 + i, lam[i]);
      }
    }
    fence_proxy_async();
    tc_fence_before();  // the Z columns are free again once every thread has passed the barrier below
    slot_sync(c.slot);
    if (c.ts == 0) {
      tc_fence_after();
      const uint32_t al = sm + LY::O_A16A, bm = sm + LY::O_BM;
      constexpr uint32_t id = idesc_f16(NS);
      mma_ss(c.tm + C_G, desc64(al + 32), desc64(bm), id, 0);  // marks are exact in fp16: two products
      mma_ss(c.tm + C_G, desc64(al), desc64(bm), id, 1);
      umma_commit_s(c.mma_bar);
    }
    AT2_TICK(7)  // lam stores, barrier, issue
    // ================================================================ gate: (G o P') -> fp16x2 (hi, lo) in TMEM [0, NS)
    // G' = G * sl <= sl * max(lam) * (largest marks row sum) < 2^15 * that sum; P' <= 2^14; BiMAU forces G[q,q] = 1,
    // i.e. G' = sl.  The product is brought under 2^15 with a power of two derived from those bounds (no second pass).
    float cgp, icgp;
    {
      const float bound = fmaxf(32768.f * fmaxf(mxm, 1.f), a.diag_one ? sl : 0.f);  // >= max G'
      int e = (int)((__float_as_uint(bound) >> 23) & 0xffu) + 1;                     // 2^(e-127) > bound
      e = min(max(e, 16), 240);
      cgp = __uint_as_float((uint32_t)(255 - e) << 23);       // 2 / 2^(e-127): P' G' cgp < 2^15
      icgp = __uint_as_float((uint32_t)(e - 1) << 23);        // 1 / cgp
    }
    mbar_wait_s(c.mma_bar, mph); mph ^= 1;
    tc_fence_after();
    AT2_TICK(8)  // wait G
    float gb[2][8];
    tmem_ld8_nowait(tml + C_G + k0, gb[0]);
#pragma unroll
    for (int j = 0; j < NSH; j += 8) {
      tmem_ld_wait();
      if (j + 8 < NSH) tmem_ld8_nowait(tml + C_G + k0 + j + 8, gb[((j >> 3) + 1) & 1]);
      const float(&g)[8] = gb[(j >> 3) & 1];
      uint32_t hi[4], lo[4];
      float gp[8];
      if (jd_hi >= j && jd_lo <= j + 7) {  // warp-uniform: some row of this warp has its diagonal in this chunk
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float gv = (j + i == jd) ? sl : g[i];  // tf.linalg.set_diag(G, 1), temporal.py:438-439
          gp[i] = (P[j + i] * cgp) * gv;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) gp[i] = (P[j + i] * cgp) * g[i];
      }
      if constexpr (DBG) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dbg_put<NS, DBG>(a, item, 5, r, k0 + j + i, ((j + i == jd) ? sl : g[i]) * isl);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) split2(gp[2 * i], gp[2 * i + 1], hi[i], lo[i]);
      tmem_st4(tml + C_S + (k0 + j) / 2, hi);
      tmem_st4(tml + C_S + NS / 2 + (k0 + j) / 2, lo);
    }
    tmem_st_wait();
    tc_fence_before();
    slot_sync(c.slot);
    if (c.ts == 0) {
      tc_fence_after();
      issue_pv(c.tm + C_O1, c.tm + C_O2, sm + LY::O_BV + (n & 1) * LY::XT);
    }
    AT2_TICK(9)  // gate, barrier, issue
    // residual row (requested now, used after the last MMA)
    float4 res0 = make_float4(0.f, 0.f, 0.f, 0.f), res1 = res0;
    if (a.R && row_ok) {
      const float* rp = a.R + grow * a.ldr + hh * DH + ch * 8;
      res0 = __ldg(reinterpret_cast<const float4*>(rp));
      res1 = __ldg(reinterpret_cast<const float4*>(rp + 4));
    }
    // ================================================================ next item's operands (under the (G o P) V MMAs)
    if (has_next) {
      mbar_wait_s(c.raw_bar, rph); rph ^= 1;
      convert_item<NS>(a, c, next, (n + 1) & 1, meta_next);
    }
    AT2_TICK(10)  // next item's conversion (incl. its TMA wait)
    mbar_wait_s(c.mma_bar, mph); mph ^= 1;  // (G o P) V done: its TMEM operand region may be overwritten by the next S
    tc_fence_after();
    slot_sync(c.slot);
    if (c.ts == 0 && has_next) {
      if (next + stride < num_items) issue_tma(next + stride);  // the raw tiles were consumed by convert_item
      tc_fence_after();
      issue_s();
    }
    AT2_TICK(11)  // wait (G o P) V, barrier, issue S
    // ================================================================ O = (G o P) V / l + residual  (temporal.py:385,447)
    {
      float o1[8], o2[8], o3[8];
      tmem_ld8_nowait(tml + C_O1 + ch * 8, o1);
      tmem_ld8_nowait(tml + C_O1 + 16 + ch * 8, o2);
      tmem_ld8_nowait(tml + C_O2 + ch * 8, o3);
      tmem_ld_wait();
      const float f1 = linv * isv, f2 = isl * icgp;
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = ((o3[i] + o2[i] + o1[i]) * f1) * f2;
      o[0] += res0.x; o[1] += res0.y; o[2] += res0.z; o[3] += res0.w;
      o[4] += res1.x; o[5] += res1.y; o[6] += res1.z; o[7] += res1.w;
      float omax = 0.f;
      if (row_ok) {
        float* op = a.O + grow * a.ldo + hh * DH + ch * 8;
        *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(op + 4) = make_float4(o[4], o[5], o[6], o[7]);
#pragma unroll
        for (int i = 0; i < 8; ++i) omax = fmaxf(omax, fabsf(o[i]));
      }
      if constexpr (DBG) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dbg_put<NS, DBG>(a, item, 6, r, ch * 8 + i, o[i]);
      }
      if (a.out_amax) amax_publish(a.out_amax, omax, c.lane);  // consumed by the scaled 3xFP16 attention-out GEMM
    }
    tc_fence_before();  // this thread's TMEM reads are complete before the barriers that precede the next MMAs
    AT2_TICK(12)  // epilogue
  }
  if constexpr (PROF) {
    if (a.prof && c.ts == 0) {
#pragma unroll
      for (int i = 0; i < 14; ++i) a.prof[((size_t)blockIdx.x * 2 + c.slot) * 16 + i] = tacc[i];
    }
  }
#undef AT2_TICK
}

template <int NS, bool DBG, bool PROF>
__global__ void __launch_bounds__(NTHR, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                     const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapT,
                     const __grid_constant__ MlpConst mc, const AttnArgs a, int num_items) {
  using LY = Lay<NS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* pack = base + LY::O_PACK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + LY::O_BAR);  // [slot][0 = mma, 1 = raw]
  const uint32_t base_s = smem_u32(base);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapK)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapV)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapT)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // operand tiles start as zeros: rows / keys >= L are never written again and must stay finite
  for (int i = tid; i < (2 * LY::SLOT) / 16; i += NTHR) reinterpret_cast<uint4*>(base)[i] = make_uint4(0u, 0u, 0u, 0u);
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.mlp_pack2);
    for (int i = tid; i < PK_BYTES / 16; i += NTHR) reinterpret_cast<uint4*>(pack)[i] = __ldg(src + i);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;

  SlotCtx c;
  c.slot = tid >> 8;
  c.ts = tid & 255;
  c.lane = tid & 31;
  const int ws = c.ts >> 5;
  c.q = ws & 3;   // = (global warp) % 4: the TMEM lane quarter this warp may access
  c.ch = ws >> 2;
  c.r = c.q * 32 + c.lane;
  c.sm = base_s + c.slot * LY::SLOT;
  c.pack = base_s + LY::O_PACK;
  c.mma_bar = base_s + LY::O_BAR + (c.slot * 2 + 0) * 8;
  c.raw_bar = base_s + LY::O_BAR + (c.slot * 2 + 1) * 8;
  c.tm = tm + c.slot * C_SLOT;
  if (num_items < 0) {  // experiment (EDGL_TC2_ONESLOT): slot 0 alone works on every item of the CTA
    if (c.slot == 0) run_slot<NS, DBG, PROF>(a, mc, c, &mapQ, &mapK, &mapV, &mapT, -num_items, blockIdx.x, gridDim.x, false);
  } else {
    run_slot<NS, DBG, PROF>(a, mc, c, &mapQ, &mapK, &mapV, &mapT, num_items, blockIdx.x * 2 + c.slot, gridDim.x * 2, true);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512));
  }
}

// ------------------------------------------------------------------------------------------ host
// [rows][cols] fp32, box = [box_rows][16 columns] landing as 64-byte rows with the 64B swizzle
int make_map_tile(CUtensorMap* m, const float* ptr, long long rows, int cols, int ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(-3, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)DH, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-3, "cuTensorMapEncodeTiled (attention tile) failed (%d)", (int)r);
  return 0;
}

// tensor maps are cached per (pointer, rows, ld, L, d): the projection buffer of a handle never moves
struct MapKey {
  const void* p; long long rows; int ld, L, d;
  bool operator<(const MapKey& o) const { return std::tie(p, rows, ld, L, d) < std::tie(o.p, o.rows, o.ld, o.L, o.d); }
};
int cached_map(CUtensorMap* out, const float* ptr, long long rows, int d, int ld, int L) {
  static std::mutex mu;
  static std::map<MapKey, CUtensorMap> cache;
  std::lock_guard<std::mutex> g(mu);
  const MapKey k{ptr, rows, ld, L, d};
  auto it = cache.find(k);
  if (it == cache.end()) {
    CUtensorMap m;
    EDGL_TRY(make_map_tile(&m, ptr, rows, d, ld, L));
    if (cache.size() > 4096) cache.clear();
    it = cache.emplace(k, m).first;
  }
  *out = it->second;
  return 0;
}

// host copies of the MLP constants, keyed by the device pack they were built with (launch_attention_tc2_pack)
std::mutex g_mc_mu;
std::map<const void*, MlpConst> g_mc;

template <int NS>
int launch_t(const AttnArgs& a, cudaStream_t st, int num_sms) {
  using LY = Lay<NS>;
  static_assert(LY::BYTES <= 227 * 1024, "attn_tc2: shared memory");
  MlpConst mc;
  {
    std::lock_guard<std::mutex> g(g_mc_mu);
    auto it = g_mc.find(a.mlp_pack2);
    if (it == g_mc.end()) return set_error(-2, "attn_tc2: the MLP pack was not built by launch_attention_tc2_pack");
    mc = it->second;
  }
  const long long rows = (long long)a.B * a.L;
  CUtensorMap mq, mk, mv, mt;
  EDGL_TRY(cached_map(&mq, a.Q, rows, a.d, a.ldq, a.L));
  EDGL_TRY(cached_map(&mk, a.K, rows, a.d, a.ldk, a.L));
  EDGL_TRY(cached_map(&mv, a.V, rows, a.d, a.ldv, a.L));
  EDGL_TRY(cached_map(&mt, a.T, rows, a.d, a.ldt, a.L));
  long long items = (long long)a.B * a.h;
  const long long pairs = (items + 1) / 2;
  const int grid = pairs < num_sms ? (int)pairs : num_sms;
  static const bool one_slot = getenv("EDGL_TC2_ONESLOT") != nullptr;
  if (one_slot) items = -items;
  if (a.prof) {
    auto kern = attention_tc2_kernel<NS, false, true>;
    EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LY::BYTES));
    kern<<<grid, NTHR, LY::BYTES, st>>>(mq, mk, mv, mt, mc, a, (int)items);
  } else if (a.dbg) {
    auto kern = attention_tc2_kernel<NS, true, false>;
    EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LY::BYTES));
    kern<<<grid, NTHR, LY::BYTES, st>>>(mq, mk, mv, mt, mc, a, (int)items);
  } else {
    auto kern = attention_tc2_kernel<NS, false, false>;
    EDGL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LY::BYTES));
    kern<<<grid, NTHR, LY::BYTES, st>>>(mq, mk, mv, mt, mc, a, (int)items);
  }
  EDGL_LAUNCH_CHECK();
  return 0;
}

}  // namespace at2

size_t attention_tc2_pack_bytes(int dh, int E) { return (dh == at2::DH && E == at2::E) ? (size_t)at2::PK_BYTES : 0; }

int launch_attention_tc2_pack(const float* int_w, const float* int_b, const float* int_weight, const float* int_scaling,
                              int dh, int E, void* pack, cudaStream_t st) {
  using namespace at2;
  EDGL_REQUIRE(attention_tc2_pack_bytes(dh, E) != 0, "attention pack: dh=%d E=%d not supported", dh, E);
  // the per-column constants travel as a kernel parameter: fetch the weights once (edgl_commit synchronises anyway)
  std::vector<float> w((size_t)(DH + 1) * NC), b(NC), wt(NC), sc(at2::E);
  EDGL_CUDA(cudaMemcpyAsync(w.data(), int_w, w.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
  EDGL_CUDA(cudaMemcpyAsync(b.data(), int_b, b.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
  EDGL_CUDA(cudaMemcpyAsync(wt.data(), int_weight, wt.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
  EDGL_CUDA(cudaMemcpyAsync(sc.data(), int_scaling, sc.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
  EDGL_CUDA(cudaStreamSynchronize(st));
  auto pow2_for = [](float m, float& s, float& is) {  // s = 2^k with m * s in [2^14, 2^15); same clamps as the device
    int e = 0;
    if (!(m > 0.f) || !std::isfinite(m)) m = std::isfinite(m) ? 0.f : 3.0e38f;
    if (m > 0.f) { frexpf(m, &e); e += 126; } else e = 0;   // biased exponent of m
    e = e < 15 ? 15 : (e > 239 ? 239 : e);
    s = ldexpf(1.f, 268 - e - 127);
    is = ldexpf(1.f, e - 14 - 127);
  };
  MlpConst mc;
  float wmax = 0.f;
  for (int i = 0; i < DH * NC; ++i) wmax = fmaxf(wmax, fabsf(kLog2e * w[i]));
  float sw, isw;
  pow2_for(wmax, sw, isw);
  for (int i = 0; i < NC; ++i) {
    mc.b[i] = -kLog2e * b[i];
    mc.wsp[i] = -kLog2e * w[(size_t)DH * NC + i];
    mc.w[i] = wt[i];
  }
  // lam_e = s_e log(1 + exp(x_e / s_e)) <= max(x_e, 0) + s_e ln 2 with x_e = sum_j sigmoid(.) w_ej <= sum_j max(w_ej, 0)
  float lam_bound = 0.f;
  for (int e = 0; e < at2::E; ++e) {
    const float s = expf(sc[e]);  // temporal.py:302
    mc.rs[e] = kLog2e / s;
    mc.sl2[e] = s * 0.69314718055994531f;
    float xpos = 0.f;
    for (int j = 0; j < DH; ++j) xpos += fmaxf(wt[e * DH + j], 0.f);
    lam_bound = fmaxf(lam_bound, xpos + mc.sl2[e]);
  }
  mc.isw = isw;
  pow2_for(lam_bound, mc.sl, mc.isl);
  mc.pad = 0.f;
  {
    std::lock_guard<std::mutex> g(g_mc_mu);
    g_mc[pack] = mc;
  }
  tc2_pack_kernel<<<1, 256, 0, st>>>(int_w, sw, reinterpret_cast<unsigned char*>(pack));
  EDGL_LAUNCH_CHECK();
  return 0;
}

// 0 = launched, 1 = shape not covered (caller falls back), <0 = error
int launch_attention_tc2(const AttnArgs& a, cudaStream_t st) {
  using namespace at2;
  if (a.d / a.h != DH || a.E != E || a.L > 112 || a.L < 1 || !a.mlp_pack2) return 1;
  if ((a.ldq | a.ldk | a.ldv | a.ldt | a.ldo) % 4 || (a.R && a.ldr % 4) || a.d % 4) return 1;
  if ((reinterpret_cast<uintptr_t>(a.marks) & 15) || (reinterpret_cast<uintptr_t>(a.mlp_pack2) & 15)) return 1;
  if ((reinterpret_cast<uintptr_t>(a.Q) | reinterpret_cast<uintptr_t>(a.K) | reinterpret_cast<uintptr_t>(a.V) |
       reinterpret_cast<uintptr_t>(a.T) | reinterpret_cast<uintptr_t>(a.O)) & 15)
    return 1;
  if (a.R && (reinterpret_cast<uintptr_t>(a.R) & 15)) return 1;
  if (a.lam && (reinterpret_cast<uintptr_t>(a.lam) & 15)) return 1;
  static int num_sms = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  if ((long long)a.B * a.h == 0) return 0;
  if (a.L <= 32) return launch_t<32>(a, st, num_sms);
  if (a.L <= 64) return launch_t<64>(a, st, num_sms);
  return launch_t<112>(a, st, num_sms);
}

}  // namespace edgl
